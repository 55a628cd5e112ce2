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

__global__ void __launch_bounds__(256) confidence_finish_kernel(float* __restrict__ c, int64_t n,
                                                                const float* __restrict__ c_max) {
    const float m = *c_max;
    for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x)
        c[i] = m - c[i];
}

// ---- splat: m = S 1, wbar = S c, b_k = S (t_k * c) ----------------------------------------------------
// conf is either the confidence itself (c_max == nullptr) or the raw Sobel magnitude, finished here as
// c = max(c) - c (:237) with the GLOBAL maximum in *c_max.
__global__ void __launch_bounds__(256) splat_kernel(Grid g, int z0, int zs, const uint8_t* __restrict__ r,
                                                    const int* __restrict__ lut, const float* __restrict__ conf,
                                                    const float* __restrict__ c_max, const float* __restrict__ t, int nrhs,
                                                    double* __restrict__ m, double* __restrict__ wbar,
                                                    double* __restrict__ b /* nrhs x ncell */) {
    __shared__ int s_lut[256];
    s_lut[threadIdx.x] = lut[threadIdx.x];
    __syncthreads();
    const int64_t n = static_cast<int64_t>(g.W) * g.H * zs;
    const float cm = c_max ? *c_max : 0.0f;
    for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int z = z0 + static_cast<int>(i % zs), y = static_cast<int>((i / zs) % g.H), x = static_cast<int>(i / (static_cast<int64_t>(zs) * g.H));
        const int64_t cell = cell_of(g, x, y, z, s_lut[__ldg(r + (static_cast<int64_t>(x) * g.H + y) * g.D + z)]);
        const double c = static_cast<double>(c_max ? cm - conf[i] : conf[i]);
        atomicAdd(m + cell, 1.0);
        atomicAdd(wbar + cell, c);
        for (int k = 0; k < nrhs; ++k) atomicAdd(b + k * g.ncell + cell, static_cast<double>(t[k * n + i]) * c);
    }
}

// ---- blur of an occupied-masked vector -----------------------------------------------------------------
__device__ __forceinline__ double blur_at(const Grid& g, const double* __restrict__ y, int64_t j) {
    const int l = static_cast<int>(j % g.L);
    const int64_t s = j / g.L;
    const int cz = static_cast<int>(s % g.gz), cy = static_cast<int>((s / g.gz) % g.gy), cx = static_cast<int>(s / (static_cast<int64_t>(g.gz) * g.gy));
    const int64_t sz = g.L, sy = static_cast<int64_t>(g.gz) * g.L, sx = static_cast<int64_t>(g.gy) * g.gz * g.L;
    double acc = 12.0 * y[j];   // 2 * dim with dim = 6 lattice dimensions (:96)
    if (l > 0) acc += y[j - 1];
    if (l < g.L - 1) acc += y[j + 1];
    if (cz > 0) acc += y[j - sz];
    if (cz < g.gz - 1) acc += y[j + sz];
    if (cy > 0) acc += y[j - sy];
    if (cy < g.gy - 1) acc += y[j + sy];
    if (cx > 0) acc += y[j - sx];
    if (cx < g.gx - 1) acc += y[j + sx];
    return acc;
}

__global__ void __launch_bounds__(256) occ_init_kernel(int64_t ncell, const double* __restrict__ m, double* __restrict__ n) {
    for (int64_t j = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; j < ncell;
         j += static_cast<int64_t>(gridDim.x) * blockDim.x)
        n[j] = m[j] > 0.0 ? 1.0 : 0.0;
}
// n_out = sqrt(n * m / blur(n)) on occupied cells (:111-112)
__global__ void __launch_bounds__(256) bistoch_step_kernel(Grid g, const double* __restrict__ m, const double* __restrict__ n_in,
                                                           double* __restrict__ n_out) {
    for (int64_t j = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; j < g.ncell;
         j += static_cast<int64_t>(gridDim.x) * blockDim.x)
        n_out[j] = m[j] > 0.0 ? sqrt(n_in[j] * m[j] / blur_at(g, n_in, j)) : 0.0;
}
// m' = n * blur(n) (:115); Minv = 1 / max(lam (m' - 12 n^2) + wbar, diag_min) (:141-142)
__global__ void __launch_bounds__(256) bistoch_finish_kernel(Grid g, const double* __restrict__ m_cnt, const double* __restrict__ n,
                                                             const double* __restrict__ wbar, double lam, double diag_min,
                                                             double* __restrict__ m_out, double* __restrict__ minv) {
    for (int64_t j = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; j < g.ncell;
         j += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const bool occ = m_cnt[j] > 0.0;
        const double mm = occ ? n[j] * blur_at(g, n, j) : 0.0;
        m_out[j] = mm;
        minv[j] = occ ? 1.0 / fmax(lam * (mm - 12.0 * n[j] * n[j]) + wbar[j], diag_min) : 0.0;
    }
}

// ---- block reduction helper: adds `v` (and optionally `u`) into device doubles -------------------------
__device__ __forceinline__ void block_add2(double v, double u, double* dst_v, double* dst_u) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        v += __shfl_xor_sync(0xffffffffu, v, o);
        u += __shfl_xor_sync(0xffffffffu, u, o);
    }
    __shared__ double sv[8], su[8];
    const int wdx = threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 0) { sv[wdx] = v; su[wdx] = u; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int i = 1; i < 8; ++i) { v += sv[i]; u += su[i]; }
        atomicAdd(dst_v, v);
        if (dst_u) atomicAdd(dst_u, u);
    }
}

// y0 = b / wbar on occupied cells (:144); np = n*y0 for the first operator application; bb = ||b||^2
__global__ void __launch_bounds__(256) pcg_init_y_kernel(Grid g, const double* __restrict__ m_cnt, const double* __restrict__ n,
                                                         const double* __restrict__ wbar, const double* __restrict__ b,
                                                         double* __restrict__ y, double* __restrict__ np, RhsScalars* sc) {
    const int k = blockIdx.y;
    const double* bk = b + k * g.ncell;
    double bb = 0.0;
    for (int64_t j = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; j < g.ncell;
         j += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const double y0 = m_cnt[j] > 0.0 ? bk[j] / wbar[j] : 0.0;
        y[k * g.ncell + j] = y0;
        np[k * g.ncell + j] = n[j] * y0;
        bb += bk[j] * bk[j];
    }
    block_add2(bb, 0.0, &sc[k].bb, nullptr);
}
// r = b - A y0; accumulates rr and rho = r . Minv r
__global__ void __launch_bounds__(256) pcg_init_r_kernel(Grid g, const double* __restrict__ m, const double* __restrict__ n,
                                                         const double* __restrict__ wbar, const double* __restrict__ minv,
                                                         const double* __restrict__ b, const double* __restrict__ y,
                                                         const double* __restrict__ np, double lam, double* __restrict__ r,
                                                         RhsScalars* sc) {
    const int k = blockIdx.y;
    const int64_t off = k * g.ncell;
    double rr = 0.0, rho = 0.0;
    for (int64_t j = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; j < g.ncell;
         j += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const double blurred = m[j] > 0.0 ? blur_at(g, np + off, j) : 0.0;   // m' > 0 exactly on occupied cells
        const double ay = lam * (m[j] * y[off + j] - n[j] * blurred) + wbar[j] * y[off + j];
        const double res = b[off + j] - ay;
        r[off + j] = res;
        rr += res * res;
        rho += res * res * minv[j];
    }
    block_add2(rr, rho, &sc[k].rr_next, &sc[k].rho_next);
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
__global__ void __launch_bounds__(256) pcg_direction_kernel(Grid g, const double* __restrict__ n, const double* __restrict__ minv,
                                                            const double* __restrict__ r, double* __restrict__ p,
                                                            double* __restrict__ np, const RhsScalars* sc, int first) {
    const int k = blockIdx.y;
    if (sc[k].done) return;
    const double beta = first ? 0.0 : sc[k].rho / sc[k].rho_prev;
    const int64_t off = k * g.ncell;
    for (int64_t j = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; j < g.ncell;
         j += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const double z = minv[j] * r[off + j];
        const double pn = first ? z : z + beta * p[off + j];
        p[off + j] = pn;
        np[off + j] = n[j] * pn;
    }
}
// q = A p; pq += p . q
__global__ void __launch_bounds__(256) pcg_apply_kernel(Grid g, const double* __restrict__ m, const double* __restrict__ n,
                                                        const double* __restrict__ wbar, const double* __restrict__ p,
                                                        const double* __restrict__ np, double lam, double* __restrict__ q,
                                                        RhsScalars* sc) {
    const int k = blockIdx.y;
    if (sc[k].done) return;
    const int64_t off = k * g.ncell;
    double pq = 0.0;
    for (int64_t j = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; j < g.ncell;
         j += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const double blurred = m[j] > 0.0 ? blur_at(g, np + off, j) : 0.0;
        const double pj = p[off + j];
        const double qj = lam * (m[j] * pj - n[j] * blurred) + wbar[j] * pj;
        q[off + j] = qj;
        pq += pj * qj;
    }
    block_add2(pq, 0.0, &sc[k].pq, nullptr);
}
// y += alpha p; r -= alpha q; accumulate the next rr / rho
__global__ void __launch_bounds__(256) pcg_update_kernel(Grid g, const double* __restrict__ minv, const double* __restrict__ p,
                                                         const double* __restrict__ q, double* __restrict__ y,
                                                         double* __restrict__ r, RhsScalars* sc) {
    const int k = blockIdx.y;
    if (sc[k].done) return;
    const double alpha = sc[k].rho / sc[k].pq;
    const int64_t off = k * g.ncell;
    double rr = 0.0, rho = 0.0;
    for (int64_t j = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; j < g.ncell;
         j += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        y[off + j] += alpha * p[off + j];
        const double res = r[off + j] - alpha * q[off + j];
        r[off + j] = res;
        rr += res * res;
        rho += res * res * minv[j];
    }
    block_add2(rr, rho, &sc[k].rr_next, &sc[k].rho_next);
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(&sc[k].iters, 1);
}

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
            float v = static_cast<float>(y[k * g.ncell + cell]);
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

// scratch of the grid stage: m, n_a, n_b, minv (ncell each) + r, p, n*p, q (nrhs * ncell each) + scalars
extern "C" int64_t vittf_bls_grid_workspace_bytes(const vittf_bls_params* p, int nrhs) {
    Grid g;
    if (!p || nrhs <= 0 || make_grid(p, &g) != VITTF_OK) return -1;
    return 4 * align_up(g.ncell * 8, 256) + 4 * align_up(g.ncell * 8 * nrhs, 256) + align_up(nrhs * sizeof(RhsScalars), 256);
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
                                                                               t_slab, nrhs, acc, acc + g.ncell, acc + 2 * g.ncell);
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
    const int64_t vec = align_up(g.ncell * 8, 256), per = align_up(g.ncell * 8 * nrhs, 256);
    uint8_t* base = static_cast<uint8_t*>(workspace);
    auto take = [&](int64_t bytes) { uint8_t* r = base; base += bytes; return r; };
    const double* m_cnt = acc;
    const double* wbar = acc + g.ncell;
    const double* b = acc + 2 * g.ncell;
    double* m = reinterpret_cast<double*>(take(vec));
    double* n_a = reinterpret_cast<double*>(take(vec));
    double* n_b = reinterpret_cast<double*>(take(vec));
    double* minv = reinterpret_cast<double*>(take(vec));
    double* r = reinterpret_cast<double*>(take(per));
    double* pd = reinterpret_cast<double*>(take(per));
    double* np = reinterpret_cast<double*>(take(per));
    double* q = reinterpret_cast<double*>(take(per));
    RhsScalars* sc = reinterpret_cast<RhsScalars*>(take(align_up(nrhs * sizeof(RhsScalars), 256)));
    const unsigned gc = blocks_for(g.ncell);
    VITTF_CHECK_CUDA(cudaMemsetAsync(sc, 0, nrhs * sizeof(RhsScalars), s));
    occ_init_kernel<<<gc, 256, 0, s>>>(g.ncell, m_cnt, n_a);
    double* n_cur = n_a;
    double* n_nxt = n_b;
    for (int it = 0; it < 10; ++it) {
        bistoch_step_kernel<<<gc, 256, 0, s>>>(g, m_cnt, n_cur, n_nxt);
        double* tmp = n_cur; n_cur = n_nxt; n_nxt = tmp;
    }
    bistoch_finish_kernel<<<gc, 256, 0, s>>>(g, m_cnt, n_cur, wbar, p->lam, p->A_diag_min, m, minv);
    dim3 gk(gc, nrhs);
    pcg_init_y_kernel<<<gk, 256, 0, s>>>(g, m_cnt, n_cur, wbar, b, y, np, sc);
    pcg_init_r_kernel<<<gk, 256, 0, s>>>(g, m, n_cur, wbar, minv, b, y, np, p->lam, r, sc);
    for (int it = 0; it < p->cg_maxiter; ++it) {
        pcg_advance_kernel<<<1, 64, 0, s>>>(sc, nrhs, p->cg_tol, it == 0);
        pcg_direction_kernel<<<gk, 256, 0, s>>>(g, n_cur, minv, r, pd, np, sc, it == 0);
        pcg_apply_kernel<<<gk, 256, 0, s>>>(g, m, n_cur, wbar, pd, np, p->lam, q, sc);
        pcg_update_kernel<<<gk, 256, 0, s>>>(g, minv, pd, q, y, r, sc);
    }
    if (iters_out) copy_iters_kernel<<<1, 64, 0, s>>>(sc, nrhs, iters_out);
    VITTF_CHECK_CUDA(cudaGetLastError());
    vittf_count_launches(10 + 1 + 1 + 2 + 4 * p->cg_maxiter + (iters_out ? 1 : 0));
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
