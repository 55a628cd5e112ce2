// Stage 3: 3-D bilateral solver, /root/reference/bilateral_solver3d.py:211-245.
//
// The reference hashes 6-D lattice coordinates, sorts them (np.unique, 77% of its run time at 256^3)
// and builds CSR splat/blur matrices.  For the grey reference volume the hot path always passes
// (predict_ntf.py:92 `cvol.expand(3, ...)`) the chroma bins are constants, so the lattice is a dense
// (Gx, Gy, Gz, L) box with an occupancy mask: no sort, no sparse matrices.
//   splat  = scatter-add of voxels into their cell (fp64 atomics)       [:66-67,87-88]
//   blur   = 12*y + sum over the 4 varying lattice dims of y[+-1], occupied cells only   [:71-81,93-99]
//   bistochastize: 10 fixed-point sweeps                                 [:107-118]
//   solve  = Jacobi-preconditioned CG, matrix-free, scipy's stopping rule [:128-154]
//   slice  = gather + float32 cast + nan_to_num                          [:90-91,245]
// All grid vectors are fp64 like the reference.  Several targets sharing one reference volume are
// solved together (blockIdx.y = right-hand side); the grid (tens of MB) stays L2-resident.
#include "common.cuh"

namespace {

struct Grid {
    int W, H, D;
    int gx, gy, gz, L;
    double sigma;
    int64_t ncell;
};

struct RhsScalars {      // one per right-hand side, device memory
    double bb;           // ||b||^2
    double rr;           // ||r||^2 of the current residual
    double rho;          // r . Minv r of the current residual
    double rho_prev;
    double pq;
    double rr_next, rho_next;   // accumulated by the update kernel
    int done;
    int iters;
};

__device__ __forceinline__ int sbin(int i, double sigma) { return static_cast<int>(static_cast<double>(i) / sigma); }

__device__ __forceinline__ int64_t cell_of(const Grid& g, int x, int y, int z, int luma_bin) {
    return ((static_cast<int64_t>(sbin(x, g.sigma)) * g.gy + sbin(y, g.sigma)) * g.gz + sbin(z, g.sigma)) * g.L + luma_bin;
}

// ---- Sobel confidence (:176-181): fp32 central differences, zero padding, on r/255 -------------------
// The volume may be processed in z-slabs [z0, z0+zs) (multi-GPU): r is always the FULL reference volume, the
// per-voxel arrays (c, t, out) are slab-local (W, H, zs).
__global__ void __launch_bounds__(256) sobel_kernel(const uint8_t* __restrict__ r, int W, int H, int D, int z0, int zs,
                                                    float* __restrict__ c_raw, float* __restrict__ c_max) {
    const int64_t n = static_cast<int64_t>(W) * H * zs;
    float local = 0.0f;
    for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int z = z0 + static_cast<int>(i % zs), y = static_cast<int>((i / zs) % H), x = static_cast<int>(i / (static_cast<int64_t>(zs) * H));
        auto at = [&](int xx, int yy, int zz) -> float {
            if (xx < 0 || xx >= W || yy < 0 || yy >= H || zz < 0 || zz >= D) return 0.0f;
            return static_cast<float>(__ldg(r + (static_cast<int64_t>(xx) * H + yy) * D + zz)) / 255.0f;
        };
        const float gz = 0.5f * at(x, y, z + 1) - 0.5f * at(x, y, z - 1);
        const float gy = 0.5f * at(x, y + 1, z) - 0.5f * at(x, y - 1, z);
        const float gx = 0.5f * at(x + 1, y, z) - 0.5f * at(x - 1, y, z);
        const float c = sqrtf(gz * gz + gy * gy + gx * gx);
        c_raw[i] = c;
        local = fmaxf(local, c);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) local = fmaxf(local, __shfl_xor_sync(0xffffffffu, local, o));
    if ((threadIdx.x & 31) == 0) atomicMax(reinterpret_cast<int*>(c_max), __float_as_int(local));  // c >= 0
}
// Same stencil, 4 z-adjacent voxels per thread from aligned 32-bit words of the five lines (D, z0, zs multiples of 4): the
// scalar kernel issues 6 byte loads per voxel and ran at 0.43 TB/s.
__global__ void __launch_bounds__(256) sobel_vec4_kernel(const uint8_t* __restrict__ r, int W, int H, int D, int z0, int zs,
                                                         float* __restrict__ c_raw, float* __restrict__ c_max) {
    const int zq = zs >> 2;
    const int64_t n4 = static_cast<int64_t>(W) * H * zq;
    float local = 0.0f;
    for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n4;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int z = z0 + 4 * static_cast<int>(i % zq), y = static_cast<int>((i / zq) % H), x = static_cast<int>(i / (static_cast<int64_t>(zq) * H));
        const uint8_t* line = r + (static_cast<int64_t>(x) * H + y) * D + z;
        auto word = [&](bool ok, const uint8_t* q) -> uint32_t { return ok ? __ldg(reinterpret_cast<const uint32_t*>(q)) : 0u; };
        const uint32_t c0 = word(true, line);
        const uint32_t xm = word(x > 0, line - static_cast<int64_t>(H) * D), xp = word(x < W - 1, line + static_cast<int64_t>(H) * D);
        const uint32_t ym = word(y > 0, line - D), yp = word(y < H - 1, line + D);
        const float zlo = z > 0 ? static_cast<float>(__ldg(line - 1)) / 255.0f : 0.0f;
        const float zhi = z + 4 < D ? static_cast<float>(__ldg(line + 4)) / 255.0f : 0.0f;
        float ctr[6];
        ctr[0] = zlo;
        ctr[5] = zhi;
        float out4[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) ctr[1 + k] = static_cast<float>((c0 >> (8 * k)) & 255u) / 255.0f;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float gz = 0.5f * ctr[k + 2] - 0.5f * ctr[k];
            const float gy = 0.5f * (static_cast<float>((yp >> (8 * k)) & 255u) / 255.0f) - 0.5f * (static_cast<float>((ym >> (8 * k)) & 255u) / 255.0f);
            const float gx = 0.5f * (static_cast<float>((xp >> (8 * k)) & 255u) / 255.0f) - 0.5f * (static_cast<float>((xm >> (8 * k)) & 255u) / 255.0f);
            out4[k] = sqrtf(gz * gz + gy * gy + gx * gx);
            local = fmaxf(local, out4[k]);
        }
        *reinterpret_cast<float4*>(c_raw + 4 * i) = make_float4(out4[0], out4[1], out4[2], out4[3]);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) local = fmaxf(local, __shfl_xor_sync(0xffffffffu, local, o));
    if ((threadIdx.x & 31) == 0) atomicMax(reinterpret_cast<int*>(c_max), __float_as_int(local));  // c >= 0
}

__global__ void __launch_bounds__(256) confidence_finish_kernel(float* __restrict__ c, int64_t n,
                                                                const float* __restrict__ c_max) {
    const float m = *c_max;
    for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x)
        c[i] = m - c[i];
}

// ---- splat: m = S 1, wbar = S c, b_k = S (t_k * c) ----------------------------------------------------
// conf is either the confidence itself (c_max == nullptr) or the raw Sobel magnitude, finished here as
// c = max(c) - c (:237) with the GLOBAL maximum in *c_max.  The accumulators are interleaved per cell,
// acc[cell][m, wbar, b_0 .. b_{nrhs-1}]: the 2 + nrhs atomics of a voxel land in one or two 128-byte lines instead of
// 2 + nrhs different planes (the planar layout moved 11 GB through DRAM for 5 GB of input at 512^3).
__global__ void __launch_bounds__(256) splat_kernel(Grid g, int z0, int zs, const uint8_t* __restrict__ r,
                                                    const int* __restrict__ lut, const float* __restrict__ conf,
                                                    const float* __restrict__ c_max, const float* __restrict__ t, int nrhs,
                                                    double* __restrict__ acc /* ncell x (2 + nrhs) */) {
    __shared__ int s_lut[256];
    s_lut[threadIdx.x] = lut[threadIdx.x];
    __syncthreads();
    const int64_t n = static_cast<int64_t>(g.W) * g.H * zs;
    const float cm = c_max ? *c_max : 0.0f;
    const int lane = threadIdx.x & 31;
    const int stride = 2 + nrhs;
    // Lanes are z-adjacent voxels, so several of them fall into the same cell (one spatial bin is 7 voxels long): the L2
    // atomic units are what bounds this kernel (195 G fp64 atomics / s measured), so lanes with equal cells first add up
    // inside the warp (match_any + a rank tree over the peer set) and only the first of them issues the 2 + nrhs atomics.
    for (int64_t i0 = (blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x) - lane; i0 < n;
         i0 += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int64_t i = i0 + lane;
        const bool valid = i < n;
        long long cell = -1 - lane;                                   // invalid lanes: unique keys, never written
        double c = 0.0;
        if (valid) {
            const int z = z0 + static_cast<int>(i % zs), y = static_cast<int>((i / zs) % g.H), x = static_cast<int>(i / (static_cast<int64_t>(zs) * g.H));
            cell = cell_of(g, x, y, z, s_lut[__ldg(r + (static_cast<int64_t>(x) * g.H + y) * g.D + z)]);
            c = static_cast<double>(c_max ? cm - conf[i] : conf[i]);
        }
        const unsigned peers = __match_any_sync(0xffffffffu, cell);
        const int rank = __popc(peers & ((1u << lane) - 1u));
        unsigned partner[5];
#pragma unroll
        for (int s5 = 0; s5 < 5; ++s5) partner[s5] = __fns(peers, lane, (1 << s5) + 1);       // peer `1 << s5` ranks above, or ~0
        auto seg_sum = [&](double v) {
#pragma unroll
            for (int s5 = 0; s5 < 5; ++s5) {
                const bool take = partner[s5] != 0xffffffffu && (rank & ((2 << s5) - 1)) == 0;
                const double o = __shfl_sync(0xffffffffu, v, take ? static_cast<int>(partner[s5]) : lane);
                if (take) v += o;
            }
            return v;
        };
        const double cnt = seg_sum(valid ? 1.0 : 0.0);
        const double csum = seg_sum(c);
        double* a = acc + (cell > 0 ? cell : 0) * stride;
        const bool lead = valid && rank == 0;
        if (lead) {
            atomicAdd(a, cnt);
            atomicAdd(a + 1, csum);
        }
        for (int k = 0; k < nrhs; ++k) {
            const double bk = seg_sum(valid ? static_cast<double>(t[k * n + i]) * c : 0.0);
            if (lead) atomicAdd(a + 2 + k, bk);
        }
    }
}

// =========================================================================================================
// Grid stage on the OCCUPIED cells only.  The dense (gx, gy, gz, L) box is ~80 % empty (a 7^3-voxel spatial bin touches
// a handful of its 52 luma bins), and with 8 right-hand sides the 25 PCG iterations stream ~110 grid vectors each:
// on the dense box that is HBM traffic nobody needs (146 ms at 512^3).  So, like the reference's np.unique vertex set
// (:60-61) but without a sort: the occupancy mask is compacted by a prefix sum (dense order is kept), every vertex gets its 8
// lattice neighbours as compact indices (:71-81, -1 = missing), and all solver vectors live in compact, right-hand-side-
// interleaved form [vertex][KP] (KP = nrhs rounded up to a power of two): per-vertex scalars and the neighbour table are
// read once for all right-hand sides, a neighbour gather is one contiguous 8 * KP-byte piece, and one thread = one
// (vertex, rhs) pair.  The arithmetic per cell is the dense formulation's, term for term.
// =========================================================================================================
constexpr int SCAN_BLOCK = 1024;         // cells per scan block (256 threads x 4)

__global__ void __launch_bounds__(256) occ_count_kernel(int64_t ncell, const double* __restrict__ acc, int stride, int* __restrict__ block_sums) {
    const int64_t base = static_cast<int64_t>(blockIdx.x) * SCAN_BLOCK;
    int c = 0;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        const int64_t j = base + u * 256 + threadIdx.x;
        c += (j < ncell && acc[j * stride] > 0.0) ? 1 : 0;
    }
    c = __syncthreads_count(c & 1) + 2 * __syncthreads_count(c & 2) + 4 * __syncthreads_count(c & 4);
    if (threadIdx.x == 0) block_sums[blockIdx.x] = c;
}
// exclusive scan of the block sums in place (one CTA), total -> *nv
__global__ void __launch_bounds__(1024) scan_block_sums_kernel(int* __restrict__ block_sums, int nblocks, int* __restrict__ nv) {
    __shared__ int warp_tot[32];
    __shared__ int carry_s;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int base = 0; base < nblocks; base += 1024) {
        const int i = base + threadIdx.x;
        const int v = i < nblocks ? block_sums[i] : 0;
        int incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) warp_tot[wid] = incl;
        __syncthreads();
        if (wid == 0) {
            int w = warp_tot[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, w, o);
                if (lane >= o) w += t;
            }
            warp_tot[lane] = w;                                   // inclusive over warps
        }
        __syncthreads();
        const int carry = carry_s;
        const int excl = carry + (wid > 0 ? warp_tot[wid - 1] : 0) + incl - v;
        if (i < nblocks) block_sums[i] = excl;
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = carry + warp_tot[31];
        __syncthreads();
    }
    if (threadIdx.x == 0) *nv = carry_s;
}
// idx_map[dense] = compact index or -1; cells[compact] = dense index (dense order preserved)
__global__ void __launch_bounds__(256) compact_kernel(int64_t ncell, const double* __restrict__ acc, int stride, const int* __restrict__ block_sums,
                                                      int* __restrict__ idx_map, int* __restrict__ cells) {
    __shared__ int warp_cnt[8];
    const int64_t base = static_cast<int64_t>(blockIdx.x) * SCAN_BLOCK;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int run = block_sums[blockIdx.x];
    for (int u = 0; u < 4; ++u) {
        const int64_t j = base + u * 256 + threadIdx.x;
        const bool occ = j < ncell && acc[j * stride] > 0.0;
        const unsigned bal = __ballot_sync(0xffffffffu, occ);
        if (lane == 0) warp_cnt[wid] = __popc(bal);
        __syncthreads();
        int before = 0, total = 0;
#pragma unroll
        for (int w = 0; w < 8; ++w) {
            before += w < wid ? warp_cnt[w] : 0;
            total += warp_cnt[w];
        }
        const int v = run + before + __popc(bal & ((1u << lane) - 1u));
        if (j < ncell) idx_map[j] = occ ? v : -1;
        if (occ) cells[v] = static_cast<int>(j);
        run += total;
        __syncthreads();
    }
}
// neighbour table in the blur's summation order: l-1, l+1, z-1, z+1, y-1, y+1, x-1, x+1; plus the per-vertex counts
__global__ void __launch_bounds__(256) neighbours_kernel(Grid g, const int* __restrict__ nv_p, const int* __restrict__ cells,
                                                         const int* __restrict__ idx_map, const double* __restrict__ acc, int stride,
                                                         int* __restrict__ nbr, double* __restrict__ m_cnt_c, double* __restrict__ wbar_c) {
    const int nv = *nv_p;
    const int64_t sz = g.L, sy = static_cast<int64_t>(g.gz) * g.L, sx = static_cast<int64_t>(g.gy) * g.gz * g.L;
    for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < nv; v += gridDim.x * blockDim.x) {
        const int64_t j = cells[v];
        const int l = static_cast<int>(j % g.L);
        const int64_t s = j / g.L;
        const int cz = static_cast<int>(s % g.gz), cy = static_cast<int>((s / g.gz) % g.gy), cx = static_cast<int>(s / (static_cast<int64_t>(g.gz) * g.gy));
        int* nb = nbr + static_cast<int64_t>(v) * 8;
        nb[0] = l > 0 ? idx_map[j - 1] : -1;
        nb[1] = l < g.L - 1 ? idx_map[j + 1] : -1;
        nb[2] = cz > 0 ? idx_map[j - sz] : -1;
        nb[3] = cz < g.gz - 1 ? idx_map[j + sz] : -1;
        nb[4] = cy > 0 ? idx_map[j - sy] : -1;
        nb[5] = cy < g.gy - 1 ? idx_map[j + sy] : -1;
        nb[6] = cx > 0 ? idx_map[j - sx] : -1;
        nb[7] = cx < g.gx - 1 ? idx_map[j + sx] : -1;
        m_cnt_c[v] = acc[j * stride];
        wbar_c[v] = acc[j * stride + 1];
    }
}
// blur of a compact scalar vector: 12 y + sum over existing neighbours (:93-99, 2 * dim with dim = 6)
__device__ __forceinline__ double blur1(const double* __restrict__ y, const int* __restrict__ nb, int v) {
    double acc = 12.0 * y[v];
    const int4 a = *reinterpret_cast<const int4*>(nb), b = *reinterpret_cast<const int4*>(nb + 4);
    if (a.x >= 0) acc += y[a.x];
    if (a.y >= 0) acc += y[a.y];
    if (a.z >= 0) acc += y[a.z];
    if (a.w >= 0) acc += y[a.w];
    if (b.x >= 0) acc += y[b.x];
    if (b.y >= 0) acc += y[b.y];
    if (b.z >= 0) acc += y[b.z];
    if (b.w >= 0) acc += y[b.w];
    return acc;
}
__global__ void __launch_bounds__(256) fill_ones_kernel(const int* __restrict__ nv_p, double* __restrict__ n) {
    const int nv = *nv_p;
    for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < nv; v += gridDim.x * blockDim.x) n[v] = 1.0;
}
// n_out = sqrt(n * m / blur(n)) (:111-112)
__global__ void __launch_bounds__(256) bistoch_step_kernel(const int* __restrict__ nv_p, const int* __restrict__ nbr, const double* __restrict__ m_cnt,
                                                           const double* __restrict__ n_in, double* __restrict__ n_out) {
    const int nv = *nv_p;
    for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < nv; v += gridDim.x * blockDim.x)
        n_out[v] = sqrt(n_in[v] * m_cnt[v] / blur1(n_in, nbr + static_cast<int64_t>(v) * 8, v));
}
// m' = n * blur(n) (:115); Minv = 1 / max(lam (m' - 12 n^2) + wbar, diag_min) (:141-142)
__global__ void __launch_bounds__(256) bistoch_finish_kernel(const int* __restrict__ nv_p, const int* __restrict__ nbr, const double* __restrict__ n,
                                                             const double* __restrict__ wbar, double lam, double diag_min,
                                                             double* __restrict__ m_out, double* __restrict__ minv) {
    const int nv = *nv_p;
    for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < nv; v += gridDim.x * blockDim.x) {
        const double mm = n[v] * blur1(n, nbr + static_cast<int64_t>(v) * 8, v);
        m_out[v] = mm;
        minv[v] = 1.0 / fmax(lam * (mm - 12.0 * n[v] * n[v]) + wbar[v], diag_min);
    }
}

// ---- per-right-hand-side block reduction: thread t holds the partial sums of rhs k = t & (KP - 1) --------------------
__device__ __forceinline__ void block_add_rhs(double v, double u, int KP, int nrhs, RhsScalars* sc, int which) {
    __shared__ double sv[256], su[256];
    sv[threadIdx.x] = v;
    su[threadIdx.x] = u;
    __syncthreads();
    if (static_cast<int>(threadIdx.x) < KP && static_cast<int>(threadIdx.x) < nrhs) {
        double a = 0.0, b = 0.0;
        for (int i = threadIdx.x; i < 256; i += KP) { a += sv[i]; b += su[i]; }
        RhsScalars& s = sc[threadIdx.x];
        if (which == 0) atomicAdd(&s.bb, a);
        else if (which == 1) { atomicAdd(&s.rr_next, a); atomicAdd(&s.rho_next, b); }
        else atomicAdd(&s.pq, a);
    }
    __syncthreads();
}
// blur of an interleaved vector for right-hand side k
__device__ __forceinline__ double blurk(const double* __restrict__ y, const int* __restrict__ nb, int64_t self, int kshift, int k) {
    double acc = 12.0 * y[self];
    const int4 a = *reinterpret_cast<const int4*>(nb), b = *reinterpret_cast<const int4*>(nb + 4);
    if (a.x >= 0) acc += y[(static_cast<int64_t>(a.x) << kshift) + k];
    if (a.y >= 0) acc += y[(static_cast<int64_t>(a.y) << kshift) + k];
    if (a.z >= 0) acc += y[(static_cast<int64_t>(a.z) << kshift) + k];
    if (a.w >= 0) acc += y[(static_cast<int64_t>(a.w) << kshift) + k];
    if (b.x >= 0) acc += y[(static_cast<int64_t>(b.x) << kshift) + k];
    if (b.y >= 0) acc += y[(static_cast<int64_t>(b.y) << kshift) + k];
    if (b.z >= 0) acc += y[(static_cast<int64_t>(b.z) << kshift) + k];
    if (b.w >= 0) acc += y[(static_cast<int64_t>(b.w) << kshift) + k];
    return acc;
}
// Items i = (vertex << kshift) + k walk the interleaved vectors contiguously; the grid stride is a multiple of KP, so a
// thread keeps its k.
#define FOR_ITEMS(i, v, k)                                                                                              \
    const int nv = *nv_p;                                                                                               \
    const int KP = 1 << kshift;                                                                                         \
    const int k = threadIdx.x & (KP - 1);                                                                               \
    const int64_t n_items = static_cast<int64_t>(nv) << kshift;                                                         \
    for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x, v = i >> kshift; i < n_items;         \
         i += static_cast<int64_t>(gridDim.x) * blockDim.x, v = i >> kshift)

// b_c = compacted b; y0 = b / wbar (:144); np = n * y0; bb = ||b||^2
__global__ void __launch_bounds__(256) pcg_init_y_kernel(const int* __restrict__ nv_p, int kshift, int nrhs, const int* __restrict__ cells,
                                                         const double* __restrict__ acc, const double* __restrict__ n, const double* __restrict__ wbar,
                                                         double* __restrict__ bc, double* __restrict__ y, double* __restrict__ np, RhsScalars* sc) {
    double bb = 0.0;
    FOR_ITEMS(i, v, k) {
        if (k < nrhs) {
            const double bv = acc[static_cast<int64_t>(cells[v]) * (2 + nrhs) + 2 + k];
            const double y0 = bv / wbar[v];
            bc[i] = bv;
            y[i] = y0;
            np[i] = n[v] * y0;
            bb += bv * bv;
        }
    }
    block_add_rhs(bb, 0.0, KP, nrhs, sc, 0);
}
// r = b - A y0; accumulates rr and rho = r . Minv r
__global__ void __launch_bounds__(256) pcg_init_r_kernel(const int* __restrict__ nv_p, int kshift, int nrhs, const int* __restrict__ nbr,
                                                         const double* __restrict__ m, const double* __restrict__ n, const double* __restrict__ wbar,
                                                         const double* __restrict__ minv, const double* __restrict__ bc, const double* __restrict__ y,
                                                         const double* __restrict__ np, double lam, double* __restrict__ r, RhsScalars* sc) {
    double rr = 0.0, rho = 0.0;
    FOR_ITEMS(i, v, k) {
        if (k < nrhs) {
            const double blurred = blurk(np, nbr + v * 8, i, kshift, k);
            const double ay = lam * (m[v] * y[i] - n[v] * blurred) + wbar[v] * y[i];
            const double res = bc[i] - ay;
            r[i] = res;
            rr += res * res;
            rho += res * res * minv[v];
        }
    }
    block_add_rhs(rr, rho, KP, nrhs, sc, 1);
}
// top of a PCG iteration: convergence test (scipy: ||r|| < rtol*||b||), scalar rotation
__global__ void pcg_advance_kernel(RhsScalars* sc, int nrhs, double tol_rel, int first) {
    const int k = threadIdx.x;
    if (k >= nrhs) return;
    RhsScalars& s = sc[k];
    if (s.done) return;
    s.rho_prev = first ? 1.0 : s.rho;
    s.rr = s.rr_next;
    s.rho = s.rho_next;
    s.rr_next = 0.0;
    s.rho_next = 0.0;
    s.pq = 0.0;
    if (sqrt(s.rr) < tol_rel * sqrt(s.bb)) s.done = 1;   // NaN compares false, like scipy
}
// p = z + (rho/rho_prev) p with z = Minv r (p = z on the first iteration); np = n * p
__global__ void __launch_bounds__(256) pcg_direction_kernel(const int* __restrict__ nv_p, int kshift, int nrhs, const double* __restrict__ n,
                                                            const double* __restrict__ minv, const double* __restrict__ r, double* __restrict__ p,
                                                            double* __restrict__ np, const RhsScalars* sc, int first) {
    const int kk = threadIdx.x & ((1 << kshift) - 1);
    const bool live = kk < nrhs && !sc[kk].done;
    const double beta = (live && !first) ? sc[kk].rho / sc[kk].rho_prev : 0.0;
    FOR_ITEMS(i, v, k) {
        if (live) {
            const double z = minv[v] * r[i];
            const double pn = first ? z : z + beta * p[i];
            p[i] = pn;
            np[i] = n[v] * pn;
        }
    }
}
// q = A p; pq += p . q
__global__ void __launch_bounds__(256) pcg_apply_kernel(const int* __restrict__ nv_p, int kshift, int nrhs, const int* __restrict__ nbr,
                                                        const double* __restrict__ m, const double* __restrict__ n, const double* __restrict__ wbar,
                                                        const double* __restrict__ p, const double* __restrict__ np, double lam, double* __restrict__ q,
                                                        RhsScalars* sc) {
    const int kk = threadIdx.x & ((1 << kshift) - 1);
    const bool live = kk < nrhs && !sc[kk].done;
    double pq = 0.0;
    FOR_ITEMS(i, v, k) {
        if (live) {
            const double blurred = blurk(np, nbr + v * 8, i, kshift, k);
            const double pj = p[i];
            const double qj = lam * (m[v] * pj - n[v] * blurred) + wbar[v] * pj;
            q[i] = qj;
            pq += pj * qj;
        }
    }
    block_add_rhs(pq, 0.0, KP, nrhs, sc, 2);
}
// y += alpha p; r -= alpha q; accumulate the next rr / rho
__global__ void __launch_bounds__(256) pcg_update_kernel(const int* __restrict__ nv_p, int kshift, int nrhs, const double* __restrict__ minv,
                                                         const double* __restrict__ p, const double* __restrict__ q, double* __restrict__ y,
                                                         double* __restrict__ r, RhsScalars* sc) {
    const int kk = threadIdx.x & ((1 << kshift) - 1);
    const bool live = kk < nrhs && !sc[kk].done;
    const double alpha = live ? sc[kk].rho / sc[kk].pq : 0.0;
    double rr = 0.0, rho = 0.0;
    FOR_ITEMS(i, v, k) {
        if (live) {
            y[i] += alpha * p[i];
            const double res = r[i] - alpha * q[i];
            r[i] = res;
            rr += res * res;
            rho += res * res * minv[v];
        }
    }
    block_add_rhs(rr, rho, KP, nrhs, sc, 1);
    if (blockIdx.x == 0 && static_cast<int>(threadIdx.x) < nrhs && !sc[threadIdx.x].done) atomicAdd(&sc[threadIdx.x].iters, 1);
}
// compact interleaved solution -> the dense, cell-major (ncell, nrhs) layout the slice stage reads (zero on unoccupied cells):
// the nrhs values of a voxel's cell are one contiguous piece
__global__ void __launch_bounds__(256) scatter_y_kernel(const int* __restrict__ nv_p, int kshift, int nrhs, const int* __restrict__ cells,
                                                        const double* __restrict__ yc, double* __restrict__ y_dense) {
    FOR_ITEMS(i, v, k) {
        if (k < nrhs) y_dense[static_cast<int64_t>(cells[v]) * nrhs + k] = yc[i];
    }
}
#undef FOR_ITEMS

// ---- slice + float32 cast + nan_to_num (:153,245) -------------------------------------------------------
__global__ void __launch_bounds__(256) slice_kernel(Grid g, int z0, int zs, const uint8_t* __restrict__ r,
                                                    const int* __restrict__ lut, const double* __restrict__ y, int nrhs,
                                                    float* __restrict__ out) {
    __shared__ int s_lut[256];
    s_lut[threadIdx.x] = lut[threadIdx.x];
    __syncthreads();
    const int64_t n = static_cast<int64_t>(g.W) * g.H * zs;
    for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int z = z0 + static_cast<int>(i % zs), yy = static_cast<int>((i / zs) % g.H), x = static_cast<int>(i / (static_cast<int64_t>(zs) * g.H));
        const int64_t cell = cell_of(g, x, yy, z, s_lut[__ldg(r + (static_cast<int64_t>(x) * g.H + yy) * g.D + z)]);
        for (int k = 0; k < nrhs; ++k) {
            float v = static_cast<float>(y[cell * nrhs + k]);
            if (isnan(v)) v = 0.0f;
            else if (isinf(v)) v = v > 0.0f ? 3.402823466e+38f : -3.402823466e+38f;
            out[k * n + i] = v;
        }
    }
}

__global__ void copy_iters_kernel(const RhsScalars* sc, int nrhs, int* out) {
    if (threadIdx.x < nrhs) out[threadIdx.x] = sc[threadIdx.x].iters;
}

int make_grid(const vittf_bls_params* p, Grid* g) {
    VITTF_REQUIRE(p->W > 0 && p->H > 0 && p->D > 0, "bls: empty volume");
    VITTF_REQUIRE(p->sigma_spatial > 0 && p->luma_bins > 0 && p->luma_bins <= 256, "bls: bad grid parameters");
    g->W = p->W; g->H = p->H; g->D = p->D;
    g->sigma = p->sigma_spatial;
    g->gx = static_cast<int>(static_cast<double>(p->W - 1) / p->sigma_spatial) + 1;
    g->gy = static_cast<int>(static_cast<double>(p->H - 1) / p->sigma_spatial) + 1;
    g->gz = static_cast<int>(static_cast<double>(p->D - 1) / p->sigma_spatial) + 1;
    g->L = p->luma_bins;
    g->ncell = static_cast<int64_t>(g->gx) * g->gy * g->gz * g->L;
    return VITTF_OK;
}

inline int64_t align_up(int64_t v, int64_t a) { return (v + a - 1) / a * a; }

unsigned blocks_for(int64_t n) {
    const int64_t bl = ceil_div_ll(n, 256), cap = static_cast<int64_t>(vittf_num_sms()) * 16;
    return static_cast<unsigned>(bl < cap ? bl : cap);
}
int check_slab(const vittf_bls_params* p, int z0, int z1) {
    VITTF_REQUIRE(z0 >= 0 && z1 > z0 && z1 <= p->D, "bls: bad z-slab [%d,%d) of depth %d", z0, z1, p->D);
    return VITTF_OK;
}

}  // namespace

extern "C" int64_t vittf_bls_grid_cells(const vittf_bls_params* p) {
    Grid g;
    if (!p || make_grid(p, &g) != VITTF_OK) return -1;
    return g.ncell;
}

// scratch of the grid stage, sized for the worst case (every cell occupied; the solve touches the first nv entries only):
// int32 block sums, vertex count, dense->compact map, compact->dense list, 8 neighbours per vertex; fp64 per-vertex
// m_cnt, wbar, m', n (x2), Minv; fp64 interleaved [vertex][KP] b, y, r, p, n*p, q; per-rhs scalars
static int kp_shift(int nrhs) {
    int s = 0;
    while ((1 << s) < nrhs) ++s;
    return s;
}
extern "C" int64_t vittf_bls_grid_workspace_bytes(const vittf_bls_params* p, int nrhs) {
    Grid g;
    if (!p || nrhs <= 0 || nrhs > 64 || make_grid(p, &g) != VITTF_OK) return -1;
    if (g.ncell >= (1ll << 31)) return -1;
    const int64_t nblocks = ceil_div_ll(g.ncell, SCAN_BLOCK);
    const int64_t kp = 1ll << kp_shift(nrhs);
    return align_up((nblocks + 64) * 4, 256) + 2 * align_up(g.ncell * 4, 256) + align_up(g.ncell * 32, 256) + 6 * align_up(g.ncell * 8, 256) +
           6 * align_up(g.ncell * 8 * kp, 256) + align_up(nrhs * sizeof(RhsScalars), 256);
}

extern "C" int64_t vittf_bls_workspace_bytes(const vittf_bls_params* p, int nrhs) {
    Grid g;
    if (!p || nrhs <= 0 || make_grid(p, &g) != VITTF_OK) return -1;
    const int64_t npix = static_cast<int64_t>(p->W) * p->H * p->D;
    // acc (m_cnt | wbar | b) + y + grid scratch + Sobel magnitude + its maximum
    return align_up((2 + nrhs) * g.ncell * 8, 256) + align_up(g.ncell * 8 * nrhs, 256) + vittf_bls_grid_workspace_bytes(p, nrhs) +
           align_up(npix * 4, 256) + 256;
}

extern "C" int vittf_bls_sobel_slab(const uint8_t* r_u8, int W, int H, int D, int z0, int z1, float* c_raw_slab, float* c_max,
                                    void* stream) {
    VITTF_REQUIRE(r_u8 && c_raw_slab && c_max && W > 0 && H > 0 && D > 0 && z0 >= 0 && z1 > z0 && z1 <= D,
                  "vittf_bls_sobel_slab: bad arguments");
    const int64_t n = static_cast<int64_t>(W) * H * (z1 - z0);
    if (D % 4 == 0 && z0 % 4 == 0 && (z1 - z0) % 4 == 0 && (reinterpret_cast<uintptr_t>(r_u8) & 3) == 0 &&
        (reinterpret_cast<uintptr_t>(c_raw_slab) & 15) == 0)
        sobel_vec4_kernel<<<blocks_for(n / 4), 256, 0, static_cast<cudaStream_t>(stream)>>>(r_u8, W, H, D, z0, z1 - z0, c_raw_slab, c_max);
    else
        sobel_kernel<<<blocks_for(n), 256, 0, static_cast<cudaStream_t>(stream)>>>(r_u8, W, H, D, z0, z1 - z0, c_raw_slab, c_max);
    VITTF_CHECK_CUDA(cudaGetLastError());
    vittf_count_launches(1);
    return VITTF_OK;
}

extern "C" int vittf_sobel_confidence(const uint8_t* r_u8, int W, int H, int D, float* out, float* scratch_max, void* stream) {
    VITTF_REQUIRE(r_u8 && out && scratch_max && W > 0 && H > 0 && D > 0, "vittf_sobel_confidence: bad arguments");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const int64_t n = static_cast<int64_t>(W) * H * D;
    VITTF_CHECK_CUDA(cudaMemsetAsync(scratch_max, 0, sizeof(float), s));
    VITTF_CHECK(vittf_bls_sobel_slab(r_u8, W, H, D, 0, D, out, scratch_max, stream));
    confidence_finish_kernel<<<blocks_for(n), 256, 0, s>>>(out, n, scratch_max);
    VITTF_CHECK_CUDA(cudaGetLastError());
    vittf_count_launches(1);
    return VITTF_OK;
}

extern "C" int vittf_bls_splat_slab(const vittf_bls_params* p, const float* t_slab, const uint8_t* r_u8, const float* conf_slab,
                                    const float* c_max, const int* luma_lut, int nrhs, int z0, int z1, double* acc,
                                    void* stream) {
    VITTF_REQUIRE(p && t_slab && r_u8 && conf_slab && luma_lut && acc, "vittf_bls_splat_slab: null pointer");
    VITTF_REQUIRE(nrhs > 0 && nrhs <= 64, "vittf_bls_splat_slab: nrhs must be in [1,64]");
    Grid g;
    VITTF_CHECK(make_grid(p, &g));
    VITTF_CHECK(check_slab(p, z0, z1));
    const int64_t n = static_cast<int64_t>(g.W) * g.H * (z1 - z0);
    splat_kernel<<<blocks_for(n), 256, 0, static_cast<cudaStream_t>(stream)>>>(g, z0, z1 - z0, r_u8, luma_lut, conf_slab, c_max,
                                                                               t_slab, nrhs, acc);
    VITTF_CHECK_CUDA(cudaGetLastError());
    vittf_count_launches(1);
    return VITTF_OK;
}

extern "C" int vittf_bls_grid_solve(const vittf_bls_params* p, int nrhs, const double* acc, double* y, int* iters_out,
                                    void* workspace, int64_t workspace_bytes, void* stream) {
    VITTF_REQUIRE(p && acc && y && workspace, "vittf_bls_grid_solve: null pointer");
    VITTF_REQUIRE(nrhs > 0 && nrhs <= 64, "vittf_bls_grid_solve: nrhs must be in [1,64]");
    Grid g;
    VITTF_CHECK(make_grid(p, &g));
    const int64_t need = vittf_bls_grid_workspace_bytes(p, nrhs);
    if (workspace_bytes < need) {
        vittf_set_error("vittf_bls_grid_solve: workspace of %lld B is smaller than the %lld B required", (long long)workspace_bytes, (long long)need);
        return VITTF_ERR_NOMEM;
    }
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const int kshift = kp_shift(nrhs);
    const int64_t kp = 1ll << kshift;
    const int64_t nblocks = ceil_div_ll(g.ncell, SCAN_BLOCK);
    const int64_t vec = align_up(g.ncell * 8, 256), per = align_up(g.ncell * 8 * kp, 256);
    uint8_t* base = static_cast<uint8_t*>(workspace);
    auto take = [&](int64_t bytes) { uint8_t* r = base; base += bytes; return r; };
    int* block_sums = reinterpret_cast<int*>(take(align_up((nblocks + 64) * 4, 256)));
    int* nv = block_sums + nblocks;                      // vertex count, device resident: no host round trip in the solve
    int* idx_map = reinterpret_cast<int*>(take(align_up(g.ncell * 4, 256)));
    int* cells = reinterpret_cast<int*>(take(align_up(g.ncell * 4, 256)));
    int* nbr = reinterpret_cast<int*>(take(align_up(g.ncell * 32, 256)));
    double* m_cnt_c = reinterpret_cast<double*>(take(vec));
    double* wbar_c = reinterpret_cast<double*>(take(vec));
    double* m = reinterpret_cast<double*>(take(vec));
    double* n_a = reinterpret_cast<double*>(take(vec));
    double* n_b = reinterpret_cast<double*>(take(vec));
    double* minv = reinterpret_cast<double*>(take(vec));
    double* bc = reinterpret_cast<double*>(take(per));
    double* yc = reinterpret_cast<double*>(take(per));
    double* r = reinterpret_cast<double*>(take(per));
    double* pd = reinterpret_cast<double*>(take(per));
    double* np = reinterpret_cast<double*>(take(per));
    double* q = reinterpret_cast<double*>(take(per));
    RhsScalars* sc = reinterpret_cast<RhsScalars*>(take(align_up(nrhs * sizeof(RhsScalars), 256)));
    VITTF_CHECK_CUDA(cudaMemsetAsync(sc, 0, nrhs * sizeof(RhsScalars), s));
    VITTF_CHECK_CUDA(cudaMemsetAsync(y, 0, static_cast<size_t>(nrhs) * g.ncell * 8, s));
    // ---- vertex set: compaction of the occupancy mask (the reference's np.unique, :60-61, without the sort) ----
    occ_count_kernel<<<static_cast<unsigned>(nblocks), 256, 0, s>>>(g.ncell, acc, 2 + nrhs, block_sums);
    scan_block_sums_kernel<<<1, 1024, 0, s>>>(block_sums, static_cast<int>(nblocks), nv);
    compact_kernel<<<static_cast<unsigned>(nblocks), 256, 0, s>>>(g.ncell, acc, 2 + nrhs, block_sums, idx_map, cells);
    // grids sized for a typical occupancy; every kernel below strides over the device-resident vertex count
    const unsigned gv = blocks_for(g.ncell / 4 + 1), gk = blocks_for((g.ncell / 4 + 1) * kp);
    neighbours_kernel<<<gv, 256, 0, s>>>(g, nv, cells, idx_map, acc, 2 + nrhs, nbr, m_cnt_c, wbar_c);
    // ---- bistochastisation (:107-118) ----
    fill_ones_kernel<<<gv, 256, 0, s>>>(nv, n_a);
    double* n_cur = n_a;
    double* n_nxt = n_b;
    for (int it = 0; it < 10; ++it) {
        bistoch_step_kernel<<<gv, 256, 0, s>>>(nv, nbr, m_cnt_c, n_cur, n_nxt);
        double* tmp = n_cur; n_cur = n_nxt; n_nxt = tmp;
    }
    bistoch_finish_kernel<<<gv, 256, 0, s>>>(nv, nbr, n_cur, wbar_c, p->lam, p->A_diag_min, m, minv);
    // ---- Jacobi-PCG with scipy's stopping rule (:128-154) ----
    pcg_init_y_kernel<<<gk, 256, 0, s>>>(nv, kshift, nrhs, cells, acc, n_cur, wbar_c, bc, yc, np, sc);
    pcg_init_r_kernel<<<gk, 256, 0, s>>>(nv, kshift, nrhs, nbr, m, n_cur, wbar_c, minv, bc, yc, np, p->lam, r, sc);
    for (int it = 0; it < p->cg_maxiter; ++it) {
        pcg_advance_kernel<<<1, 64, 0, s>>>(sc, nrhs, p->cg_tol, it == 0);
        pcg_direction_kernel<<<gk, 256, 0, s>>>(nv, kshift, nrhs, n_cur, minv, r, pd, np, sc, it == 0);
        pcg_apply_kernel<<<gk, 256, 0, s>>>(nv, kshift, nrhs, nbr, m, n_cur, wbar_c, pd, np, p->lam, q, sc);
        pcg_update_kernel<<<gk, 256, 0, s>>>(nv, kshift, nrhs, minv, pd, q, yc, r, sc);
    }
    scatter_y_kernel<<<gk, 256, 0, s>>>(nv, kshift, nrhs, cells, yc, y);
    if (iters_out) copy_iters_kernel<<<1, 64, 0, s>>>(sc, nrhs, iters_out);
    VITTF_CHECK_CUDA(cudaGetLastError());
    vittf_count_launches(3 + 1 + 1 + 10 + 1 + 2 + 4 * p->cg_maxiter + 1 + (iters_out ? 1 : 0));
    return VITTF_OK;
}

extern "C" int vittf_bls_slice_slab(const vittf_bls_params* p, const uint8_t* r_u8, const int* luma_lut, const double* y, int nrhs,
                                    int z0, int z1, float* out_slab, void* stream) {
    VITTF_REQUIRE(p && r_u8 && luma_lut && y && out_slab && nrhs > 0, "vittf_bls_slice_slab: bad arguments");
    Grid g;
    VITTF_CHECK(make_grid(p, &g));
    VITTF_CHECK(check_slab(p, z0, z1));
    const int64_t n = static_cast<int64_t>(g.W) * g.H * (z1 - z0);
    slice_kernel<<<blocks_for(n), 256, 0, static_cast<cudaStream_t>(stream)>>>(g, z0, z1 - z0, r_u8, luma_lut, y, nrhs, out_slab);
    VITTF_CHECK_CUDA(cudaGetLastError());
    vittf_count_launches(1);
    return VITTF_OK;
}

// Single-GPU composition of the stages above (the whole volume is one slab).
extern "C" int vittf_bls_solve(const vittf_bls_params* p, const float* t, const uint8_t* r_u8, const float* conf,
                               const int* luma_lut, int nrhs, float* out, int* iters_out, void* workspace,
                               int64_t workspace_bytes, void* stream) {
    VITTF_REQUIRE(p && t && r_u8 && luma_lut && out && workspace, "vittf_bls_solve: null pointer");
    VITTF_REQUIRE(nrhs > 0 && nrhs <= 64, "vittf_bls_solve: nrhs must be in [1,64]");
    Grid g;
    VITTF_CHECK(make_grid(p, &g));
    const int64_t need = vittf_bls_workspace_bytes(p, nrhs);
    if (workspace_bytes < need) {
        vittf_set_error("vittf_bls_solve: workspace of %lld B is smaller than the %lld B required", (long long)workspace_bytes, (long long)need);
        return VITTF_ERR_NOMEM;
    }
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const int64_t npix = static_cast<int64_t>(g.W) * g.H * g.D;
    uint8_t* base = static_cast<uint8_t*>(workspace);
    auto take = [&](int64_t bytes) { uint8_t* r = base; base += bytes; return r; };
    const int64_t acc_bytes = align_up((2 + nrhs) * g.ncell * 8, 256);
    double* acc = reinterpret_cast<double*>(take(acc_bytes));
    double* y = reinterpret_cast<double*>(take(align_up(g.ncell * 8 * nrhs, 256)));
    const int64_t grid_bytes = vittf_bls_grid_workspace_bytes(p, nrhs);
    void* grid_ws = take(grid_bytes);
    float* c_buf = reinterpret_cast<float*>(take(align_up(npix * 4, 256)));
    float* c_max = reinterpret_cast<float*>(take(256));
    VITTF_CHECK_CUDA(cudaMemsetAsync(acc, 0, acc_bytes, s));
    if (!conf) {
        VITTF_CHECK_CUDA(cudaMemsetAsync(c_max, 0, sizeof(float), s));
        VITTF_CHECK(vittf_bls_sobel_slab(r_u8, g.W, g.H, g.D, 0, g.D, c_buf, c_max, stream));
        VITTF_CHECK(vittf_bls_splat_slab(p, t, r_u8, c_buf, c_max, luma_lut, nrhs, 0, g.D, acc, stream));
    } else {
        VITTF_CHECK(vittf_bls_splat_slab(p, t, r_u8, conf, nullptr, luma_lut, nrhs, 0, g.D, acc, stream));
    }
    VITTF_CHECK(vittf_bls_grid_solve(p, nrhs, acc, y, iters_out, grid_ws, grid_bytes, stream));
    VITTF_CHECK(vittf_bls_slice_slab(p, r_u8, luma_lut, y, nrhs, 0, g.D, out, stream));
    return VITTF_OK;
}
