// Stage 2: prototype similarity.  Replaces the arithmetic of
//   /root/reference/infer.py:48-72            sample_features3d (grid_sample)
//   /root/reference/predict_ntf.py:59-72      einsum + where/pow + mean           (REF-NTF)
//   /root/reference/old/cluster_dino.py:306-322  normalize + einsum + clamp/pow + max (LEGACY)
// and the north-star composition (up-sample FEATURES trilinearly -> normalise -> dot -> class max)
// WITHOUT ever materialising the up-sampled feature volume:
//   pass 1 (sim_lowres)   reads the (F, n_lr) feature volume once -> dots (A, n_lr) and the 14 Gram
//                         scalars per voxel that determine |interp(f)|^2 inside every cell;
//   pass 2 (sim_upsample) per output voxel: 8-corner trilinear combination of the dots, norm from the
//                         Gram, non-linearity, per-class reduction; writes C maps.
#include <stdlib.h>

#include <type_traits>

#include "common.cuh"
#include "sim_internal.h"

namespace {

template <typename T>
__device__ __forceinline__ float ldf(const T* p, int64_t i);
template <>
__device__ __forceinline__ float ldf<__half>(const __half* p, int64_t i) { return __half2float(__ldg(p + i)); }
template <>
__device__ __forceinline__ float ldf<float>(const float* p, int64_t i) { return __ldg(p + i); }

// ---------------------------------------------------------------------------------------------
// grid_sample(align_corners=False, padding_mode='zeros'), one CTA per annotation.
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(128) sample_prototypes_kernel(const T* __restrict__ feats, int F, int w, int h, int d,
                                                                const float* __restrict__ rel, int mode,
                                                                float* __restrict__ out) {
    const int a = blockIdx.x;
    // rel is (X,Y,Z); feature dims (w,h,d) are the same order (the reference flips to grid_sample's x=last-dim order)
    const float px = ((rel[a * 3 + 0] + 1.0f) * w - 1.0f) / 2.0f;
    const float py = ((rel[a * 3 + 1] + 1.0f) * h - 1.0f) / 2.0f;
    const float pz = ((rel[a * 3 + 2] + 1.0f) * d - 1.0f) / 2.0f;
    const int64_t n = static_cast<int64_t>(w) * h * d;
    if (mode == 0) {
        const int ix = static_cast<int>(nearbyintf(px)), iy = static_cast<int>(nearbyintf(py)),
                  iz = static_cast<int>(nearbyintf(pz));
        const bool ok = ix >= 0 && ix < w && iy >= 0 && iy < h && iz >= 0 && iz < d;
        const int64_t v = (static_cast<int64_t>(ix) * h + iy) * d + iz;
        for (int f = threadIdx.x; f < F; f += blockDim.x) out[static_cast<int64_t>(a) * F + f] = ok ? ldf<T>(feats, f * n + v) : 0.0f;
        return;
    }
    const float fx = floorf(px), fy = floorf(py), fz = floorf(pz);
    const int x0 = static_cast<int>(fx), y0 = static_cast<int>(fy), z0 = static_cast<int>(fz);
    const float tx = px - fx, ty = py - fy, tz = pz - fz;
    float wgt[8];
    int64_t idx[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        const int cx = x0 + (c >> 2), cy = y0 + ((c >> 1) & 1), cz = z0 + (c & 1);
        const bool ok = cx >= 0 && cx < w && cy >= 0 && cy < h && cz >= 0 && cz < d;
        const float wx = (c >> 2) ? tx : 1.0f - tx, wy = ((c >> 1) & 1) ? ty : 1.0f - ty, wz = (c & 1) ? tz : 1.0f - tz;
        wgt[c] = ok ? wx * wy * wz : 0.0f;
        idx[c] = ok ? (static_cast<int64_t>(cx) * h + cy) * d + cz : 0;
    }
    for (int f = threadIdx.x; f < F; f += blockDim.x) {
        float acc = 0.0f;
#pragma unroll
        for (int c = 0; c < 8; ++c) acc += wgt[c] * ldf<T>(feats, f * n + idx[c]);
        out[static_cast<int64_t>(a) * F + f] = acc;
    }
}

// ---------------------------------------------------------------------------------------------
// pass 1: dots + Gram at feature resolution.  One thread per voxel, AT prototypes in registers,
// prototypes staged (transposed) in shared memory per F-chunk and read as broadcast float4.
// 13 "forward" neighbour offsets (first non-zero component positive), enumeration shared with pass 2.
// ---------------------------------------------------------------------------------------------
constexpr int FCHUNK = 32;

__host__ __device__ __forceinline__ void fwd_offset(int o, int& dx, int& dy, int& dz) {
    // o = 0: self; 1..9: dx=1, (dy,dz) in {-1,0,1}^2; 10..12: dx=0, dy=1, dz in {-1,0,1}; 13: (0,0,1)
    if (o == 0) { dx = dy = dz = 0; }
    else if (o <= 9) { dx = 1; dy = (o - 1) / 3 - 1; dz = (o - 1) % 3 - 1; }
    else if (o <= 12) { dx = 0; dy = 1; dz = o - 11; }
    else { dx = 0; dy = 0; dz = 1; }
}
// index of the Gram slot for a relative offset (any sign); `swap` tells the caller to anchor at the other voxel
__device__ __forceinline__ int gram_slot(int dx, int dy, int dz, bool& swap) {
    swap = dx < 0 || (dx == 0 && (dy < 0 || (dy == 0 && dz < 0)));
    if (swap) { dx = -dx; dy = -dy; dz = -dz; }
    if (dx == 1) return 1 + (dy + 1) * 3 + (dz + 1);
    if (dy == 1) return 11 + dz;
    return dz == 1 ? 13 : 0;
}

template <typename T, int AT, bool GRAM>
__global__ void __launch_bounds__(128) sim_lowres_kernel(const T* __restrict__ feats, int F, int w, int h, int d,
                                                         const float* __restrict__ protos, int A, int a_base,
                                                         float* __restrict__ dots, float* __restrict__ gram) {
    __shared__ __align__(16) float s_p[FCHUNK][AT];
    const int64_t n = static_cast<int64_t>(w) * h * d;
    const int64_t v = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
    const bool live = v < n;
    const int64_t vv = live ? v : n - 1;
    int64_t nb[14];
    bool nb_ok[14];
    if (GRAM) {
        const int iz = static_cast<int>(vv % d), iy = static_cast<int>((vv / d) % h), ix = static_cast<int>(vv / (static_cast<int64_t>(d) * h));
#pragma unroll
        for (int o = 0; o < 14; ++o) {
            int dx, dy, dz;
            fwd_offset(o, dx, dy, dz);
            const int jx = ix + dx, jy = iy + dy, jz = iz + dz;
            nb_ok[o] = jx >= 0 && jx < w && jy >= 0 && jy < h && jz >= 0 && jz < d;
            nb[o] = nb_ok[o] ? (static_cast<int64_t>(jx) * h + jy) * d + jz : vv;
        }
    }
    float acc[AT];
    float g[14];
#pragma unroll
    for (int i = 0; i < AT; ++i) acc[i] = 0.0f;
#pragma unroll
    for (int o = 0; o < 14; ++o) g[o] = 0.0f;

    for (int f0 = 0; f0 < F; f0 += FCHUNK) {
        __syncthreads();
        for (int i = threadIdx.x; i < FCHUNK * AT; i += blockDim.x) {
            const int ff = i / AT, aa = i - ff * AT;
            const int a = a_base + aa;
            s_p[ff][aa] = (a < A && f0 + ff < F) ? protos[static_cast<int64_t>(a) * F + f0 + ff] : 0.0f;
        }
        __syncthreads();
        const int fmax = min(FCHUNK, F - f0);
        for (int ff = 0; ff < fmax; ++ff) {
            const T* plane = feats + static_cast<int64_t>(f0 + ff) * n;
            const float x = ldf<T>(plane, vv);
#pragma unroll
            for (int i = 0; i < AT; i += 4) {
                const float4 pv = *reinterpret_cast<const float4*>(&s_p[ff][i]);
                acc[i + 0] = fmaf(x, pv.x, acc[i + 0]);
                acc[i + 1] = fmaf(x, pv.y, acc[i + 1]);
                acc[i + 2] = fmaf(x, pv.z, acc[i + 2]);
                acc[i + 3] = fmaf(x, pv.w, acc[i + 3]);
            }
            if (GRAM) {
                g[0] = fmaf(x, x, g[0]);
#pragma unroll
                for (int o = 1; o < 14; ++o) g[o] = fmaf(x, ldf<T>(plane, nb[o]), g[o]);
            }
        }
    }
    if (!live) return;
#pragma unroll
    for (int i = 0; i < AT; ++i)
        if (a_base + i < A) dots[static_cast<int64_t>(a_base + i) * n + v] = acc[i];
    if (GRAM) {
#pragma unroll
        for (int o = 0; o < 14; ++o) gram[static_cast<int64_t>(o) * n + v] = nb_ok[o] ? g[o] : 0.0f;
    }
}


// ---------------------------------------------------------------------------------------------
// pass 1, prototype dots on the tensor cores (fp16 feature volumes): dots[a][v] = sum_f feats[f][v] * protos[a][f] is a
// (n_lr x F) x (F x A) GEMM whose A operand is stored "voxel-major" -- exactly what ldmatrix.trans delivers.  Warp-level
// mma.sync (m16n8k16, fp32 accumulate) is ample here: the kernel only has to keep up with the feature stream from HBM
// (F * n_lr * 2 bytes), which the FMA pipe could not (it bounded the dots at ~130 us for 32 prototypes at 64^3).
// One CTA = 128 consecutive voxels (4 warps x 32); feature planes arrive as 128B-swizzled TMA boxes of 64 voxels x 32 f;
// the prototypes (rounded to fp16, unit vectors: 3e-5 on a dot) sit in shared memory, rows padded against bank conflicts.
// ---------------------------------------------------------------------------------------------
constexpr int DM_BM = 128, DM_BK = 32, DM_STAGES = 4;
constexpr int DM_STAGE_BYTES = 2 * DM_BK * 128;           // two boxes of (32 f) x (64 voxels x 2 B)

__device__ __forceinline__ void ldmatrix_x4(uint32_t addr, uint32_t (&r)[4]) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t addr, uint32_t (&r)[4]) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void mma_f16_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

template <int NT>   // N-tiles of 8 prototypes handled by the CTA (A <= 8 * NT)
__global__ void __launch_bounds__(128) sim_dots_mma_kernel(const __grid_constant__ CUtensorMap tm_f, int F, int64_t n,
                                                           const float* __restrict__ protos, int A, int a_base,
                                                           float* __restrict__ dots, int64_t sa, int64_t sv, int64_t v_first) {
    extern __shared__ __align__(1024) uint8_t dm_smem_raw[];
    uint8_t* dm_smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(dm_smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* s_t = dm_smem;                                                   // [stage][box][32 f][64 v] halves, swizzled
    __half* s_p = reinterpret_cast<__half*>(dm_smem + DM_STAGES * DM_STAGE_BYTES);   // [8 NT][F + 8]
    const int pstride = F + 8;
    uint64_t* full = reinterpret_cast<uint64_t*>(s_p + static_cast<size_t>(8 * NT) * pstride);
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int64_t v0 = v_first + static_cast<int64_t>(blockIdx.x) * DM_BM;      // n = end of the voxel range
    const int nchunk = (F + DM_BK - 1) / DM_BK;
    if (tid == 0) {
        for (int i = 0; i < DM_STAGES; ++i) ptx::mbar_init(&full[i], 1);
        ptx::fence_barrier_init();
    }
    __syncthreads();
    auto issue = [&](int c) {
        uint64_t* bar = &full[c % DM_STAGES];
        uint8_t* dst = s_t + (c % DM_STAGES) * DM_STAGE_BYTES;
        ptx::mbar_arrive_expect_tx(bar, DM_STAGE_BYTES);
        ptx::tma_load_2d(dst, &tm_f, bar, static_cast<int>(v0), c * DM_BK);
        ptx::tma_load_2d(dst + DM_BK * 128, &tm_f, bar, static_cast<int>(v0) + 64, c * DM_BK);
    };
    if (tid == 0) {
        ptx::prefetch_tmap(&tm_f);
        for (int c = 0; c < DM_STAGES && c < nchunk; ++c) issue(c);
    }
    for (int i = tid; i < 8 * NT * pstride; i += 128) {
        const int aa = i / pstride, ff = i - aa * pstride;
        const int a = a_base + aa;
        s_p[i] = __float2half_rn(a < A && ff < F ? protos[static_cast<size_t>(a) * F + ff] : 0.0f);
    }
    __syncthreads();
    float acc[2][NT][4];
#pragma unroll
    for (int m = 0; m < 2; ++m)
#pragma unroll
        for (int j = 0; j < NT; ++j)
#pragma unroll
            for (int e = 0; e < 4; ++e) acc[m][j][e] = 0.0f;
    // ldmatrix row address of this lane inside a stage: matrix i = lane / 8, row r = lane % 8
    const int mi = lane >> 3, mr = lane & 7;
    const uint32_t p_base = ptx::smem_u32(s_p);
    for (int c = 0; c < nchunk; ++c) {
        ptx::mbar_wait(&full[c % DM_STAGES], (c / DM_STAGES) & 1);
        const uint32_t st = ptx::smem_u32(s_t + (c % DM_STAGES) * DM_STAGE_BYTES);
#pragma unroll
        for (int ks = 0; ks < DM_BK / 16; ++ks) {
            uint32_t af[2][4];
#pragma unroll
            for (int m = 0; m < 2; ++m) {
                // A fragment of voxels [wid*32 + m*16, +16): matrices (m0:8,k0:8) (m8:16,k0:8) (m0:8,k8:16) (m8:16,k8:16)
                const int vloc = wid * 32 + m * 16 + (mi & 1) * 8;      // voxel offset inside the CTA tile
                const int f = ks * 16 + (mi >> 1) * 8 + mr;              // feature row inside the stage
                const int box = vloc >> 6, chunk = (vloc & 63) >> 3;
                ldmatrix_x4_trans(st + box * (DM_BK * 128) + f * 128 + ((chunk ^ (f & 7)) << 4), af[m]);
            }
#pragma unroll
            for (int j = 0; j < NT; j += 2) {
                // B fragments of prototypes [8j, 8j+16): matrices (n0:8,k0:8) (n0:8,k8:16) (n8:16,k0:8) (n8:16,k8:16)
                uint32_t bf[4];
                const int prow = 8 * j + (mi >> 1) * 8 + mr;
                const int pk = c * DM_BK + ks * 16 + (mi & 1) * 8;
                if (j + 1 < NT || (NT & 1) == 0) {
                    ldmatrix_x4(p_base + (prow * pstride + pk) * 2, bf);
                } else {                                                 // odd tail: second tile does not exist
                    ldmatrix_x4(p_base + ((8 * j + mr) * pstride + pk) * 2, bf);
                }
#pragma unroll
                for (int m = 0; m < 2; ++m) {
                    mma_f16_16816(acc[m][j], af[m], bf[0], bf[1]);
                    if (j + 1 < NT) mma_f16_16816(acc[m][j + 1], af[m], bf[2], bf[3]);
                }
            }
        }
        __syncthreads();                                   // every warp is done with this stage
        if (tid == 0 && c + DM_STAGES < nchunk) issue(c + DM_STAGES);
    }
    // D fragment: d0,d1 -> (row lane/4, cols 2*(lane%4) + {0,1}); d2,d3 -> row + 8
#pragma unroll
    for (int m = 0; m < 2; ++m)
#pragma unroll
        for (int j = 0; j < NT; ++j)
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int a = a_base + 8 * j + 2 * (lane & 3) + (e & 1);
                const int64_t v = v0 + wid * 32 + m * 16 + (lane >> 2) + (e >> 1) * 8;
                if (a < A && v < n) dots[a * sa + v * sv] = acc[m][j][e];
            }
}

template <int NT>
int launch_dots_mma_one(const CUtensorMap& tm, int F, int64_t n, const float* protos, int A, int a_base, float* dots,
                        int64_t sa, int64_t sv, int64_t v_first, cudaStream_t s) {
    const size_t smem = static_cast<size_t>(DM_STAGES) * DM_STAGE_BYTES + static_cast<size_t>(8 * NT) * (F + 8) * 2 + 64 + 1024;
    auto kern = sim_dots_mma_kernel<NT>;
    static PerDeviceMemo configured;
    if (smem > configured.cur()) {
        VITTF_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
        configured.cur() = smem;
    }
    kern<<<static_cast<unsigned>(ceil_div_ll(n - v_first, DM_BM)), 128, smem, s>>>(tm, F, n, protos, A, a_base, dots, sa, sv, v_first);
    vittf_count_launches(1);
    return VITTF_OK;
}

// all prototypes, 64 per launch
int launch_dots_mma(const __half* feats, int F, int64_t n, const float* protos, int A, float* dots, cudaStream_t s, int a_first = 0,
                    int64_t sa = -1, int64_t sv = 1, int64_t v_first = 0, int64_t v_end = -1) {
    if (sa < 0) sa = n;
    if (v_end < 0) v_end = n;
    CUtensorMap tm;
    const uint64_t dims[2] = {static_cast<uint64_t>(n), static_cast<uint64_t>(F)};
    const uint64_t strides[1] = {static_cast<uint64_t>(n) * 2};
    const uint32_t box[2] = {64, DM_BK};
    VITTF_CHECK(vittf_make_tmap(&tm, feats, 2, 2, dims, strides, box, true));
    for (int a_base = a_first; a_base < A; a_base += 64) {
        const int rem = A - a_base;
        int rc;
        if (rem > 32) rc = launch_dots_mma_one<8>(tm, F, v_end, protos, A, a_base, dots, sa, sv, v_first, s);
        else if (rem > 16) rc = launch_dots_mma_one<4>(tm, F, v_end, protos, A, a_base, dots, sa, sv, v_first, s);
        else if (rem > 8) rc = launch_dots_mma_one<2>(tm, F, v_end, protos, A, a_base, dots, sa, sv, v_first, s);
        else rc = launch_dots_mma_one<1>(tm, F, v_end, protos, A, a_base, dots, sa, sv, v_first, s);
        VITTF_CHECK(rc);
    }
    VITTF_CHECK_CUDA(cudaGetLastError());
    return VITTF_OK;
}

// ---------------------------------------------------------------------------------------------
// pass 1 fused on the tensor cores: prototype dots AND the 14 Gram planes from ONE read of the fp16 feature volume.
// The Gram entry of voxel v and forward neighbour v + (dx, dy, dz) is a dot product over F, so for a tile of 16 z-adjacent
// voxels (an m16 tile) the products with the 8-voxel tiles of the five neighbour z-lines (dx, dy) in {(0,0), (0,1), (1,-1),
// (1,0), (1,1)} are m16n8k16 MMAs whose +-1 diagonals are the wanted entries: 3 + 4 x 4 n-tiles per k16 step, plus A / 8
// n-tiles for the prototypes -- 23 HMMA instead of 14 x 16 x 16 + 32 x 16 x 16 FMAs.  ldmatrix.trans turns the staged
// [f][line][z] rows into A fragments, and the SAME registers are the B fragments of the two 8-voxel halves (a 16 x 16 block
// loaded once serves both roles).  Persistent CTAs: 8 consumer warps (2 lines x 64 z = 8 m-tiles per tile) + 1 TMA producer
// warp that keeps a ring of 16-feature stages full across tile boundaries (3 boxes per stage: lines y0..y0+2 of plane x,
// y0-1..y0+1 and y0+2 of plane x+1; out-of-volume coordinates are zero-filled = the "no neighbour" convention).  The box is
// 88 voxels long in z ([tz0 - 8, tz0 + 80)) so that the per-feature strides (528 / 176 B) are odd multiples of 16 B:
// every ldmatrix phase is bank-conflict free.
// ---------------------------------------------------------------------------------------------
constexpr int GM_ZB = 88, GM_TZ = 64, GM_KF = 16;
constexpr int GM_L1 = GM_ZB * 2;                          // bytes of one staged z-line of one feature
constexpr int GM_L3 = 3 * GM_L1;                          // per-feature stride of a 3-line box
constexpr int GM_OFF_P = GM_KF * GM_L3, GM_OFF_Q = 2 * GM_KF * GM_L3;
constexpr int GM_STAGE = GM_KF * (2 * GM_L3 + GM_L1);     // 19712
constexpr int GM_CONSUMERS = 8;

__device__ __forceinline__ void ldmatrix_x2_trans(uint32_t addr, uint32_t& r0, uint32_t& r1) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(addr));
}

template <int NT>   // n-tiles of 8 prototypes (A <= 8 * NT)
__global__ void __maxnreg__(168)
    sim_lowres_mma_kernel(const __grid_constant__ CUtensorMap tm3, const __grid_constant__ CUtensorMap tm1, int F, int w, int h, int d,
                          const float* __restrict__ protos, int A, float* __restrict__ dots, float* __restrict__ gram,
                          int tiles_y, int tiles_z, int ntiles, int nstages, int64_t sa, int64_t sv, int xa) {
    extern __shared__ __align__(1024) uint8_t gm_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(gm_raw) + 127) & ~uintptr_t(127));
    uint8_t* s_t = smem;                                                                   // [stage][X | P | Q]
    __half* s_p = reinterpret_cast<__half*>(smem + static_cast<size_t>(nstages) * GM_STAGE);   // [8 NT][F + 8]
    const int pstride = F + 8;
    uint64_t* full = reinterpret_cast<uint64_t*>(s_p + static_cast<size_t>(8 * NT) * pstride);
    uint64_t* empty = full + nstages;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int nchunk = F / GM_KF;
    if (tid == 0) {
        for (int i = 0; i < nstages; ++i) {
            ptx::mbar_init(&full[i], 1);
            ptx::mbar_init(&empty[i], GM_CONSUMERS);
        }
        ptx::fence_barrier_init();
    }
    // prototype panel (fp16, rows padded by 8 halves): one prototype per warp pass, 4 independent loads per lane in flight
    for (int a = wid; a < 8 * NT; a += GM_CONSUMERS + 1) {
        __half* row = s_p + static_cast<size_t>(a) * pstride;
        const float* src = protos + static_cast<size_t>(a) * F;
        for (int f0 = lane; f0 < pstride; f0 += 128) {
            float v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) v[u] = (a < A && f0 + 32 * u < F) ? __ldg(src + f0 + 32 * u) : 0.0f;
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (f0 + 32 * u < pstride) row[f0 + 32 * u] = __float2half_rn(v[u]);
        }
    }
    __syncthreads();

    if (wid == GM_CONSUMERS) {
        // ---- TMA producer ------------------------------------------------------------------------------------
        if (lane == 0) {
            ptx::prefetch_tmap(&tm3);
            ptx::prefetch_tmap(&tm1);
            int st = 0;
            uint32_t ph = 0;                                  // parity of the phase that releases the stage's previous use
            bool first_round = true;                          // (nothing to wait for while the ring fills for the first time)
            for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
                const int zb = tile % tiles_z, yb = (tile / tiles_z) % tiles_y, x = xa + tile / (tiles_z * tiles_y);
                const int tz = zb * GM_TZ - 8, y0 = yb * 2;
                for (int c = 0; c < nchunk; ++c) {
                    if (!first_round) ptx::mbar_wait_quiet(&empty[st], ph);
                    uint8_t* dst = s_t + static_cast<size_t>(st) * GM_STAGE;
                    ptx::mbar_arrive_expect_tx(&full[st], GM_STAGE);
                    ptx::tma_load_4d(dst, &tm3, &full[st], tz, y0, x, c * GM_KF);
                    ptx::tma_load_4d(dst + GM_OFF_P, &tm3, &full[st], tz, y0 - 1, x + 1, c * GM_KF);
                    ptx::tma_load_4d(dst + GM_OFF_Q, &tm1, &full[st], tz, y0 + 2, x + 1, c * GM_KF);
                    if (++st == nstages) {
                        st = 0;
                        if (!first_round) ph ^= 1u;
                        first_round = false;
                    }
                }
            }
        }
        return;
    }

    // ---- consumers: warp = (line li, z tile zt) ----------------------------------------------------------------
    const int li = wid >> 2, zt = wid & 3;
    const int g = lane >> 2, t = lane & 3;
    const int mi = lane >> 3, mr = lane & 7;
    // per-lane ldmatrix row offsets for the two per-feature strides
    // A order (own line): matrices (v 0-7, f 0-7) (v 8-15, f 0-7) (v 0-7, f 8-15) (v 8-15, f 8-15);
    // B order (neighbour windows): (v 0-7, f 0-7) (v 0-7, f 8-15) (v 8-15, f 0-7) (v 8-15, f 8-15) -> register pairs = B fragments
    const uint32_t lo3a = (8 * (mi >> 1) + mr) * GM_L3 + 16 * (mi & 1);
    const uint32_t lo3 = (8 * (mi & 1) + mr) * GM_L3 + 16 * (mi >> 1), lo1 = (8 * (mi & 1) + mr) * GM_L1 + 16 * (mi >> 1);
    const uint32_t lo3x2 = (8 * (mi & 1) + mr) * GM_L3;                      // x2: matrices (f 0-7), (f 8-15) of 8 voxels
    const uint32_t zc0 = (8 + 16 * zt) * 2;                                  // byte offset of the m-tile's first voxel in a staged line
    // neighbour lines (dx, dy): (0,1), (1,-1), (1,0), (1,1)
    uint32_t nb_off[4];
    nb_off[0] = (li + 1) * GM_L1 + lo3;
    nb_off[1] = GM_OFF_P + li * GM_L1 + lo3;
    nb_off[2] = GM_OFF_P + (li + 1) * GM_L1 + lo3;
    nb_off[3] = li == 0 ? GM_OFF_P + 2 * GM_L1 + lo3 : GM_OFF_Q + lo1;
    const uint32_t own_off = li * GM_L1 + lo3a + zc0, own_x2 = li * GM_L1 + lo3x2 + zc0 + 32;
    const uint32_t st_base = ptx::smem_u32(s_t);
    const uint32_t p_base = ptx::smem_u32(s_p);
    const int64_t n = static_cast<int64_t>(w) * h * d;

    float acc_o[3][4], acc_n[4][4][4], acc_d[NT][4];
    auto zero_acc = [&]() {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
#pragma unroll
            for (int i = 0; i < 3; ++i) acc_o[i][e] = 0.0f;
#pragma unroll
            for (int k = 0; k < 4; ++k)
#pragma unroll
                for (int i = 0; i < 4; ++i) acc_n[k][i][e] = 0.0f;
#pragma unroll
            for (int j = 0; j < NT; ++j) acc_d[j][e] = 0.0f;
        }
    };
    zero_acc();
    int st = 0;
    uint32_t ph = 0;
    uint32_t sb = st_base;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        uint32_t pk = p_base + (((NT > 1 ? (mi >> 1) * 8 : 0) + mr) * pstride + (mi & 1) * 8) * 2;   // panel ldmatrix row of this lane
#pragma unroll 1
        for (int c = 0; c < nchunk; ++c) {
            ptx::mbar_wait_quiet(&full[st], ph);
            // all fragments of the k-step first (11 + NT/2 ldmatrix in flight), then the MMAs
            uint32_t a[4], o2[2], w0[4][4], w1[4][4], bf[(NT + 1) / 2][4];
            ldmatrix_x4_trans(sb + own_off, a);
            ldmatrix_x2_trans(sb + own_x2, o2[0], o2[1]);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                ldmatrix_x4_trans(sb + nb_off[k] + zc0 - 16, w0[k]);         // voxels [z0 - 8, z0 + 8)
                ldmatrix_x4_trans(sb + nb_off[k] + zc0 + 16, w1[k]);         // voxels [z0 + 8, z0 + 24)
            }
            // B fragments of prototypes [8j, 8j+16): matrices (n0:8,k0:8) (n0:8,k8:16) (n8:16,k0:8) (n8:16,k8:16)
#pragma unroll
            for (int j = 0; j < NT; j += 2) ldmatrix_x4(pk + 16 * j * pstride, bf[j / 2]);
            pk += GM_KF * 2;
            mma_f16_16816(acc_o[0], a, a[0], a[2]);
            mma_f16_16816(acc_o[1], a, a[1], a[3]);
            mma_f16_16816(acc_o[2], a, o2[0], o2[1]);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                mma_f16_16816(acc_n[k][0], a, w0[k][0], w0[k][1]);
                mma_f16_16816(acc_n[k][1], a, w0[k][2], w0[k][3]);
                mma_f16_16816(acc_n[k][2], a, w1[k][0], w1[k][1]);
                mma_f16_16816(acc_n[k][3], a, w1[k][2], w1[k][3]);
            }
#pragma unroll
            for (int j = 0; j < NT; j += 2) {
                mma_f16_16816(acc_d[j], a, bf[j / 2][0], bf[j / 2][1]);
                if (j + 1 < NT) mma_f16_16816(acc_d[j + 1], a, bf[j / 2][2], bf[j / 2][3]);
            }
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(&empty[st]);
            sb += GM_STAGE;
            if (++st == nstages) {
                st = 0;
                ph ^= 1;
                sb = st_base;
            }
        }
        // ---- tile epilogue: the wanted diagonals of the accumulators -> gram planes, prototype columns -> dots --------
        const int zb = tile % tiles_z, yb = (tile / tiles_z) % tiles_y, x = xa + tile / (tiles_z * tiles_y);
        const int y = yb * 2 + li, z0 = zb * GM_TZ + 16 * zt;
        const bool yok = y < h;
        const int64_t vb = (static_cast<int64_t>(x) * h + y) * d + z0;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int row = g + 8 * (e >> 1);
            const bool ok = yok && z0 + row < d;
            const int cb = 2 * t + (e & 1) - row;                           // column - row, before the tile offset
#pragma unroll
            for (int j = 0; j < NT; ++j) {
                const int aa = 8 * j + 2 * t + (e & 1);
                if (ok && aa < A) dots[aa * sa + (vb + row) * sv] = acc_d[j][e];
            }
            if (gram != nullptr) {
#pragma unroll
                for (int i = 0; i < 3; ++i) {
                    const int dz = 8 * i + cb;
                    if (ok && dz == 0) gram[vb + row] = acc_o[i][e];
                    if (ok && dz == 1) gram[13 * n + vb + row] = acc_o[i][e];
                }
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int slot0 = k == 0 ? 11 : 3 * k - 1;              // slot of dz = 0: (0,1) -> 11, (1,-1) -> 2, (1,0) -> 5, (1,1) -> 8
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int dz = 8 * i - 8 + cb;
                        if (ok && dz >= -1 && dz <= 1) gram[static_cast<int64_t>(slot0 + dz) * n + vb + row] = acc_n[k][i][e];
                    }
                }
            }
        }
        zero_acc();
    }
}

template <int NT>
int launch_lowres_mma_one(const CUtensorMap& tm3, const CUtensorMap& tm1, int F, int w, int h, int d, const float* protos, int A,
                          float* dots, float* gram, int64_t sa, int64_t sv, int xa, int xb, cudaStream_t s) {
    const int tiles_z = ceil_div(d, GM_TZ), tiles_y = ceil_div(h, 2);
    const int ntiles = tiles_z * tiles_y * (xb - xa);
    const size_t panel = static_cast<size_t>(8 * NT) * (F + 8) * 2;
    int nstages = static_cast<int>((220 * 1024 - panel - 256) / GM_STAGE);
    nstages = nstages > 8 ? 8 : nstages;
    if (nstages < 3) return -1;
    const size_t smem = static_cast<size_t>(nstages) * GM_STAGE + panel + 2 * nstages * 8 + 128;
    auto kern = sim_lowres_mma_kernel<NT>;
    static PerDeviceMemo configured;
    if (smem > configured.cur()) {
        VITTF_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
        configured.cur() = smem;
    }
    const int grid = ntiles < vittf_num_sms() ? ntiles : vittf_num_sms();
    kern<<<grid, 32 * (GM_CONSUMERS + 1), smem, s>>>(tm3, tm1, F, w, h, d, protos, A, dots, gram, tiles_y, tiles_z, ntiles, nstages, sa, sv, xa);
    vittf_count_launches(1);
    return VITTF_OK;
}

// first <= 64 prototypes + Gram planes fused (one read of the feature volume); prototypes beyond 64 on the dots-only
// tensor-core kernel
#ifndef GM_MAX_PROTOS
#define GM_MAX_PROTOS 64
#endif
int launch_lowres_mma(const __half* feats, int F, int w, int h, int d, const float* protos, int A, float* dots, float* gram,
                      cudaStream_t s, int layout = 0, int xa = 0, int xb = -1) {
    if (xb < 0) xb = w;
    // layout 0: dots (A, n_lr); 1: (n_lr, A4) with A4 = A rounded up to 4 (the tcgen05 up-sampling kernel gathers the corner
    // dots of 4 prototypes with one 16-byte load)
    const int64_t n_lr = static_cast<int64_t>(w) * h * d;
    const int64_t sa = layout ? 1 : n_lr, sv = layout ? ((A + 3) & ~3) : 1;
    CUtensorMap tm3, tm1;
    const uint64_t dims[4] = {static_cast<uint64_t>(d), static_cast<uint64_t>(h), static_cast<uint64_t>(w), static_cast<uint64_t>(F)};
    const uint64_t strides[3] = {static_cast<uint64_t>(d) * 2, static_cast<uint64_t>(h) * d * 2, static_cast<uint64_t>(w) * h * d * 2};
    const uint32_t box3[4] = {GM_ZB, 3, 1, GM_KF}, box1[4] = {GM_ZB, 1, 1, GM_KF};
    VITTF_CHECK(vittf_make_tmap(&tm3, feats, 2, 4, dims, strides, box3, false));
    VITTF_CHECK(vittf_make_tmap(&tm1, feats, 2, 4, dims, strides, box1, false));
    const int a0 = A < GM_MAX_PROTOS ? A : GM_MAX_PROTOS;
    int rc;
    if (a0 > 32) rc = launch_lowres_mma_one<8>(tm3, tm1, F, w, h, d, protos, a0, dots, gram, sa, sv, xa, xb, s);
    else if (a0 > 16) rc = launch_lowres_mma_one<4>(tm3, tm1, F, w, h, d, protos, a0, dots, gram, sa, sv, xa, xb, s);
    else if (a0 > 8) rc = launch_lowres_mma_one<2>(tm3, tm1, F, w, h, d, protos, a0, dots, gram, sa, sv, xa, xb, s);
    else rc = launch_lowres_mma_one<1>(tm3, tm1, F, w, h, d, protos, a0, dots, gram, sa, sv, xa, xb, s);
    if (rc != VITTF_OK) return rc;
    if (A > GM_MAX_PROTOS) VITTF_CHECK(launch_dots_mma(feats, F, n_lr, protos, A, dots, s, GM_MAX_PROTOS, sa, sv, static_cast<int64_t>(xa) * h * d, static_cast<int64_t>(xb) * h * d));
    VITTF_CHECK_CUDA(cudaGetLastError());
    return VITTF_OK;
}

// ---------------------------------------------------------------------------------------------
// pass 2: per output voxel.  Index rule of F.interpolate(mode='trilinear', align_corners=False):
//   src = max((dst + 0.5) * in/out - 0.5, 0); i0 = floor(src); i1 = min(i0 + 1, in - 1); t = src - i0.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void src_index(int dst, int in, int out, int& i0, int& i1, float& t) {
    if (in == out) { i0 = i1 = dst; t = 0.0f; return; }
    const float scale = static_cast<float>(in) / static_cast<float>(out);
    float src = (dst + 0.5f) * scale - 0.5f;
    src = src < 0.0f ? 0.0f : src;
    i0 = static_cast<int>(src);
    i1 = i0 + (i0 < in - 1 ? 1 : 0);
    t = src - i0;
}

__device__ __forceinline__ float pow_unit(float x, float e) {
    // x in [0, 1] (or 0 after thresholding); exact fast paths for the exponents the reference uses
    if (e == 2.0f) return x * x;
    if (e == 2.5f) return x * x * sqrtf(x);
    if (e == 1.0f) return x;
    return x > 0.0f ? __powf(x, e) : 0.0f;
}

__global__ void __launch_bounds__(256) sim_upsample_kernel(UpParams q) {
    extern __shared__ int s_off[];
    for (int i = threadIdx.x; i <= q.C; i += blockDim.x) s_off[i] = q.class_offsets[i];
    __syncthreads();
    const int zs = q.z1 - q.z0;
    const int64_t n_out = static_cast<int64_t>(q.x1 - q.x0) * q.H * zs;
    const int64_t n_lr = static_cast<int64_t>(q.w) * q.h * q.d;
    for (int64_t o = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; o < n_out;
         o += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int oz = static_cast<int>(o % zs), oy = static_cast<int>((o / zs) % q.H),
                  ox = q.x0 + static_cast<int>(o / (static_cast<int64_t>(zs) * q.H));
        int x0, x1, y0, y1, c0, c1;
        float tx, ty, tz;
        src_index(ox, q.w, q.W, x0, x1, tx);
        src_index(oy, q.h, q.H, y0, y1, ty);
        src_index(oz + q.z0, q.d, q.D, c0, c1, tz);
        int64_t idx[8];
        float wgt[8];
        int cx[8], cy[8], cz[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            cx[c] = (c >> 2) ? x1 : x0;
            cy[c] = ((c >> 1) & 1) ? y1 : y0;
            cz[c] = (c & 1) ? c1 : c0;
            idx[c] = (static_cast<int64_t>(cx[c]) * q.h + cy[c]) * q.d + cz[c];
            wgt[c] = ((c >> 2) ? tx : 1.0f - tx) * (((c >> 1) & 1) ? ty : 1.0f - ty) * ((c & 1) ? tz : 1.0f - tz);
        }
        float inv_norm = 1.0f;
        if (q.mode == VITTF_SIM_NS) {
            float n2 = 0.0f;
#pragma unroll
            for (int a = 0; a < 8; ++a) {
                if (wgt[a] == 0.0f) continue;
                n2 = fmaf(wgt[a] * wgt[a], __ldg(q.gram + idx[a]), n2);
#pragma unroll
                for (int b = a + 1; b < 8; ++b) {
                    if (wgt[b] == 0.0f) continue;
                    bool swap;
                    const int slot = gram_slot(cx[b] - cx[a], cy[b] - cy[a], cz[b] - cz[a], swap);
                    const float gv = __ldg(q.gram + static_cast<int64_t>(slot) * n_lr + (swap ? idx[b] : idx[a]));
                    n2 = fmaf(2.0f * wgt[a] * wgt[b], gv, n2);
                }
            }
            inv_norm = 1.0f / fmaxf(sqrtf(fmaxf(n2, 0.0f)), 1e-12f);     // F.normalize eps
        } else if (q.mode == VITTF_SIM_LEGACY) {
            inv_norm = 1.0f / fmaxf(sqrtf(__ldg(q.gram + idx[0])), 1e-12f);
        }
        for (int c = 0; c < q.C; ++c) {
            const bool mean_mode = q.mode == VITTF_SIM_REFNTF || q.mode == VITTF_SIM_CLAMP_MEAN;
            float red = mean_mode ? 0.0f : -INFINITY;
            for (int a = s_off[c]; a < s_off[c + 1]; ++a) {
                const float* da = q.dots + static_cast<int64_t>(a) * n_lr;
                float s = 0.0f;
#pragma unroll
                for (int k = 0; k < 8; ++k)
                    if (wgt[k] != 0.0f) s = fmaf(wgt[k], __ldg(da + idx[k]), s);
                if (q.mode == VITTF_SIM_REFNTF) {
                    red += pow_unit(s >= q.threshold ? s : 0.0f, q.exponent);
                } else if (q.mode == VITTF_SIM_CLAMP_MEAN) {
                    red += pow_unit(fminf(fmaxf(s, 0.0f), 1.0f), q.exponent);       // infer.py:104
                } else {
                    red = fmaxf(red, s);  // clamp(0,1)**e is monotone: reduce first, transform once
                }
            }
            float r;
            if (mean_mode) r = red / static_cast<float>(s_off[c + 1] - s_off[c]);
            else r = pow_unit(fminf(fmaxf(red * inv_norm, 0.0f), 1.0f), q.exponent);
            q.out[static_cast<int64_t>(c) * n_out + o] = r;
        }
    }
}



// ---------------------------------------------------------------------------------------------
// pass 2 on the tensor cores (NS mode, U in {4, 8}).  Inside a low-res cell the U^3 outputs of one prototype are
//   out[o] = sum_{corner c} W[o][c] * dots[corner c],      W[o][c] = wx * wy * wz  -- the SAME U^3 x 8 matrix for every cell
// (border cells clamp their corner INDICES, not their weights), so a tile of 16 cells is the GEMM
//   D (16 cells x 64 outputs) = Corners (16 x 8) * W^T (8 x 64)
// per prototype.  fp32 accuracy on bf16 tensor cores: corners split hi + mid (16 mantissa bits), weights split hi + lo
// (they have up to 12 significant bits, bf16 keeps 8) -> [hi | hi] x [W_hi | W_lo] + [mid | mid] x [W_hi | W_lo]: two m16n8k16 per 16 x 8 tile,
// relative error <= 2^-16.  The class max is a running FMNMX over the accumulator fragments (cells = rows, so no shuffles),
// 1 / |interp(f)| comes from the Gram planes on the FMA pipe once per tile (amortised over all prototypes), outputs leave
// as float2 along z.  A warp owns 16 consecutive cells of the flattened (cy, cz) cell grid at one cx -- the half cells
// at the borders are ordinary rows, so no lane idles.  Instruction count per output voxel: ~90 (cell kernel: ~350).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void mma_bf16_16816_z(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%10,%10,%10,%10};"
                 : "=f"(d[0]), "=f"(d[1]), "=f"(d[2]), "=f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1), "f"(0.0f));
}
__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ float bf16_lo_f(uint32_t p) { return __uint_as_float(p << 16); }
__device__ __forceinline__ float bf16_hi_f(uint32_t p) { return __uint_as_float(p & 0xffff0000u); }

__device__ __forceinline__ const float* up_ptr_at(const float* base, uint32_t off) {      // one IMAD.WIDE.U32
    uint64_t r;
    asm("mad.wide.u32 %0, %1, 4, %2;" : "=l"(r) : "r"(off), "l"(reinterpret_cast<uint64_t>(base)));
    return reinterpret_cast<const float*>(r);
}

// B fragment words of W^T for output (kx, ky, kz) and the corner pair k = (2t, 2t+1) = (bx = t>>1, by = t&1, bz = 0/1)
template <int U>
__device__ __forceinline__ uint2 up_w_frag(int kx, int ky, int kz, int t) {
    const float tx = (kx + 0.5f) / U, ty = (ky + 0.5f) / U, tz = (kz + 0.5f) / U;
    const float wxy = ((t >> 1) ? tx : 1.0f - tx) * ((t & 1) ? ty : 1.0f - ty);
    const float w0 = wxy * (1.0f - tz), w1 = wxy * tz;                    // exact: <= 12 significant bits
    const uint32_t hi = ptx::pack_bf16x2(w0, w1);
    const uint32_t lo = ptx::pack_bf16x2(w0 - bf16_lo_f(hi), w1 - bf16_hi_f(hi));
    return make_uint2(hi, lo);
}

__device__ __forceinline__ float rsqrt_fast(float x) { float y; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float* up_out_at(float* base, int off) {                         // one IMAD.WIDE
    uint64_t r;
    asm("mad.wide.s32 %0, %1, 4, %2;" : "=l"(r) : "r"(off), "l"(reinterpret_cast<uint64_t>(base)));
    return reinterpret_cast<float*>(r);
}
// bits [lo, hi) of an 8-bit mask
__device__ __forceinline__ uint32_t bit_range(int lo, int hi) {
    lo = lo < 0 ? 0 : lo;
    hi = hi > 8 ? 8 : hi;
    return hi > lo ? ((1u << hi) - 1u) & ~((1u << lo) - 1u) : 0u;
}

constexpr int UP_PD = 4;            // prefetch distance of the corner dots, in prototypes
// CTAs per SM of the factor-4 instantiation: 3 (168 registers, 64-164 B of spills in the prototype loop) measured 324 us
// against 227 us for 2 (203 registers, no spills) at 384 x 64^3 -> 256^3, 32 prototypes -- the kernel is short of issue
// slots, not of warps
#ifndef UP_BLOCKS_U4
#define UP_BLOCKS_U4 2
#endif
#define UP_MIN_BLOCKS(U) ((U) == 4 ? UP_BLOCKS_U4 : 2)

template <int U, int EXPK>
__global__ void __launch_bounds__(128, UP_MIN_BLOCKS(U)) sim_upsample_mma_kernel(UpParams q, int cz_lo, int ncz, int tiles_per_x, int total_tasks, int cx_lo) {
    constexpr int NSUB = U == 8 ? 8 : 1;             // sub-blocks of 64 outputs per cell
    extern __shared__ __align__(16) uint8_t um_smem[];
    int* s_off = reinterpret_cast<int*>(um_smem);
    uint2* s_w = reinterpret_cast<uint2*>(um_smem + (((q.C + 1) * 4 + 15) & ~15));   // [sub][ntile][lane]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;

    for (int i = threadIdx.x; i <= q.C; i += 128) s_off[i] = q.class_offsets[i];
#pragma unroll 1
    for (int i = threadIdx.x; i < NSUB * 8 * 32; i += 128) {
        const int l = i & 31, sj = i >> 5;
        if (U == 8) s_w[i] = up_w_frag<U>(sj >> 3, sj & 7, l >> 2, l & 3);
        else s_w[i] = up_w_frag<U>(sj >> 1, 2 * (sj & 1) + (l >> 4), (l >> 2) & 3, l & 3);
    }
    __syncthreads();
    const int task = static_cast<int>(blockIdx.x) * 4 + warp;
    if (task >= total_tasks) return;
    const int cxi = task / tiles_per_x, tile = task - cxi * tiles_per_x;
    const int cx = cx_lo + cxi;
    const int w = q.w, h = q.h, d = q.d;
    const int ncells = (h + 1) * ncz;
    const uint32_t n_lr = static_cast<uint32_t>(w) * h * d;
    const int zs = q.z1 - q.z0;
    const int x0c = cx < 0 ? 0 : cx, x1c = cx + 1 > w - 1 ? w - 1 : cx + 1;
    const int oxb = U * cx + U / 2;

    // ---- the two cells (rows g, g + 8) of this thread -------------------------------------------------------
    uint32_t off[2][2];              // A-fragment corner offsets: (row, bz)
    int obase[2];                    // output offset of (kx, ky, kz) = (0, 0, 0) + the thread's fixed (ky, kz) part
    float inv[2][8][2];
    float Q[2][2][10];               // z-contracted Gram per row and per kz of the thread: (a, b) pairs of (bx, by) corners
    const int kz0 = U == 4 ? 2 * (t & 1) : 2 * t;
    const int ky_t = U == 4 ? (t >> 1) : 0;          // thread's fixed part of ky (U == 4: ky = 2 (j & 1) + (t >> 1))
    bool zok[2][2];
    int oyb[2], ycl[2][2], zcl[2][2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        int id = tile * 16 + g + 8 * r;
        const bool rlive = id < ncells;
        id = rlive ? id : ncells - 1;
        const int cyi = id / ncz;
        const int cy = cyi - 1, cz = cz_lo + (id - cyi * ncz);
        const int y0c = cy < 0 ? 0 : cy, y1c = cy + 1 > h - 1 ? h - 1 : cy + 1;
        const int z0c = cz < 0 ? 0 : cz, z1c = cz + 1 > d - 1 ? d - 1 : cz + 1;
        const uint32_t line = (static_cast<uint32_t>((t >> 1) ? x1c : x0c) * h + ((t & 1) ? y1c : y0c)) * d;
        off[r][0] = line + z0c;
        off[r][1] = line + z1c;
        const int oz = U * cz + U / 2 + kz0;
        zok[r][0] = rlive && oz >= q.z0 && oz < q.z1 && oz >= 0 && oz < q.D;
        zok[r][1] = rlive && oz + 1 >= q.z0 && oz + 1 < q.z1 && oz + 1 >= 0 && oz + 1 < q.D;
        oyb[r] = U * cy + U / 2 + ky_t;
        obase[r] = ((oxb - q.x0) * q.H + oyb[r]) * zs + (oz - q.z0);
        ycl[r][0] = y0c;
        ycl[r][1] = y1c;
        zcl[r][0] = z0c;
        zcl[r][1] = z1c;
    }
    // the corner dots of the first UP_PD prototypes go out BEFORE the Gram phase: their L2 round trip overlaps the Gram
    // loads and the norm arithmetic below
    // (prototypes a + 1 .. a + UP_PD are in flight while a runs on the tensor cores: the loads are L2 hits, ~1 us away)
    const float* da = q.dots;
    const float* const da_last = q.dots + static_cast<size_t>(q.A - 1) * n_lr;
    float nv[UP_PD][4];
    // (volatile: the load must stay behind the consumption of the ring slot it refills, or the compiler loads into a
    // temporary and copies it into the slot -- a copy that waits for the load)
    auto ldv = [](const float* p) { float v; asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(v) : "l"(p)); return v; };
    auto fetch = [&](float (&v)[4]) {
        v[0] = ldv(up_ptr_at(da, off[0][0]));
        v[1] = ldv(up_ptr_at(da, off[0][1]));
        v[2] = ldv(up_ptr_at(da, off[1][0]));
        v[3] = ldv(up_ptr_at(da, off[1][1]));
        da = da < da_last ? da + n_lr : da;                              // (past the end: re-read the last prototype)
    };
#pragma unroll
    for (int i = 0; i < UP_PD; ++i) fetch(nv[i]);
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        // Gram of the 8 corners -> contract z for the thread's two kz.  Distinct corners differ by exactly the cell's
        // clamped extents (ex, ey, ez in {0, 1}); the slot of a pair is looked up from those.
        const int y0c = ycl[r][0], y1c = ycl[r][1], z0c = zcl[r][0], z1c = zcl[r][1];
        const int ex = x1c - x0c, ey = y1c - y0c, ez = z1c - z0c;
        const uint32_t base = (static_cast<uint32_t>(x0c) * h + y0c) * d + z0c;
        const uint32_t sx = ex ? static_cast<uint32_t>(h) * d : 0u, sy = ey ? static_cast<uint32_t>(d) : 0u, sz = ez ? 1u : 0u;
        float G[8][8];
        const float* gb = q.gram + base;
        if (ex & ey & ez) {
            // interior cell: slots and anchor corners are compile-time, plane and corner offsets warp-uniform
#pragma unroll
            for (int a = 0; a < 8; ++a)
#pragma unroll
                for (int b = a; b < 8; ++b) {
                    bool swap;
                    const int slot = gram_slot((b >> 2) - (a >> 2), ((b >> 1) & 1) - ((a >> 1) & 1), (b & 1) - (a & 1), swap);
                    const int an = swap ? b : a;
                    const size_t uo = static_cast<size_t>(slot) * n_lr + static_cast<size_t>(an >> 2) * (static_cast<uint32_t>(h) * d) +
                                      static_cast<size_t>((an >> 1) & 1) * static_cast<uint32_t>(d) + (an & 1);
                    const float gv = __ldg(gb + uo);
                    G[a][b] = gv;
                    G[b][a] = gv;
                }
        } else {
#pragma unroll
            for (int a = 0; a < 8; ++a)
#pragma unroll
                for (int b = a; b < 8; ++b) {
                    bool swap;
                    const int slot = gram_slot(((b >> 2) - (a >> 2)) * ex, (((b >> 1) & 1) - ((a >> 1) & 1)) * ey, ((b & 1) - (a & 1)) * ez, swap);
                    const int an = swap ? b : a;        // anchor corner
                    const uint32_t idx = ((an >> 2) ? sx : 0u) + (((an >> 1) & 1) ? sy : 0u) + ((an & 1) ? sz : 0u);
                    const float gv = __ldg(gb + static_cast<size_t>(slot) * n_lr + idx);
                    G[a][b] = gv;
                    G[b][a] = gv;
                }
        }
#pragma unroll
        for (int zi = 0; zi < 2; ++zi) {
            const float tz = (kz0 + zi + 0.5f) / U;
            const float w00 = (1.0f - tz) * (1.0f - tz), w01 = (1.0f - tz) * tz, w11 = tz * tz;
            int e = 0;
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int b = a; b < 4; ++b, ++e)
                    Q[r][zi][e] = w00 * G[2 * a][2 * b] + w01 * (G[2 * a][2 * b + 1] + G[2 * a + 1][2 * b]) + w11 * G[2 * a + 1][2 * b + 1];
        }
    }
    const bool vec_ok = (q.z0 & 1) == 0 && (zs & 1) == 0 && (reinterpret_cast<uintptr_t>(q.out) & 7) == 0;
    const int Hzs = q.H * zs;
    const size_t n_out = static_cast<size_t>(q.x1 - q.x0) * Hzs;

    uint32_t whi[8], wlo[8];
    if (U == 4) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const uint2 f = s_w[j * 32 + lane];
            whi[j] = f.x;
            wlo[j] = f.y;
        }
    }
    // store masks, bit (r * 8 + j), for the two z-adjacent elements of a pair.  Valid kx (ky) form a range.
    // U == 4: j = 2 kx + kyi with ky = 2 kyi + ky_t;  U == 8: j = ky (kx = sub-block, handled per sub-block)
    uint32_t m0 = 0, m1 = 0;
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        uint32_t mr;
        if (U == 4) {
            const uint32_t mx = bit_range(q.x0 - oxb, q.x1 - oxb);                // bit kx
            const uint32_t mx2 = ((mx & 1) * 3u) | (((mx >> 1) & 1) * 12u) | (((mx >> 2) & 1) * 48u) | (((mx >> 3) & 1) * 192u);
            const uint32_t my = ((oyb[r] >= 0 && oyb[r] < q.H) ? 0x55u : 0u) | ((oyb[r] + 2 >= 0 && oyb[r] + 2 < q.H) ? 0xaau : 0u);
            mr = mx2 & my;
        } else {
            mr = bit_range(-oyb[r], q.H - oyb[r]);
        }
        m0 |= zok[r][0] ? mr << (8 * r) : 0u;
        m1 |= zok[r][1] ? mr << (8 * r) : 0u;
    }
    // warp-uniform: every pair is stored whole (or not at all) as one 8-byte word
    const bool fast_st = __all_sync(0xffffffffu, vec_ok && m0 == m1);

#pragma unroll 1
    for (int sub = 0; sub < NSUB; ++sub) {
        uint32_t ms0 = m0, ms1 = m1;
        if (U == 8) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const uint2 f = s_w[(sub * 8 + j) * 32 + lane];
                whi[j] = f.x;
                wlo[j] = f.y;
            }
            const bool xok = oxb + sub >= q.x0 && oxb + sub < q.x1;
            ms0 = xok ? m0 : 0u;
            ms1 = xok ? m1 : 0u;
        }
        // ---- 1 / |interp(f)| of the thread's 32 outputs: contract x, then y --------------------------------
        // Q index e of (a, b), a <= b in (bx, by) = 00, 01, 10, 11: (0,0)=0 (0,1)=1 (0,2)=2 (0,3)=3 (1,1)=4 (1,2)=5 (1,3)=6 (2,2)=7 (2,3)=8 (3,3)=9
#pragma unroll
        for (int kxi = 0; kxi < (U == 4 ? 4 : 1); ++kxi) {
            const int kx = U == 4 ? kxi : sub;
            const float tx = (kx + 0.5f) / U;
            const float x00 = (1.0f - tx) * (1.0f - tx), x01 = (1.0f - tx) * tx, x11 = tx * tx;
#pragma unroll
            for (int r = 0; r < 2; ++r)
#pragma unroll
                for (int zi = 0; zi < 2; ++zi) {
                    const float* qq = Q[r][zi];
                    // R[by][by'] = sum_{bx, bx'} wx wx' Q[(bx,by)][(bx',by')]
                    const float r00 = x00 * qq[0] + 2.0f * x01 * qq[2] + x11 * qq[7];
                    const float r01 = x00 * qq[1] + x01 * (qq[3] + qq[5]) + x11 * qq[8];
                    const float r11 = x00 * qq[4] + 2.0f * x01 * qq[6] + x11 * qq[9];
#pragma unroll
                    for (int kyi = 0; kyi < (U == 4 ? 2 : 8); ++kyi) {
                        const int j = U == 4 ? 2 * kxi + kyi : kyi;
                        const int ky = U == 4 ? 2 * kyi + ky_t : kyi;
                        const float ty = (ky + 0.5f) / U;
                        const float n2 = (1.0f - ty) * (1.0f - ty) * r00 + 2.0f * ty * (1.0f - ty) * r01 + ty * ty * r11;
                        inv[r][j][zi] = rsqrt_fast(fmaxf(n2, 1e-24f));       // 1 / max(|v|, 1e-12)
                    }
                }
        }

        // ---- per class: running max over its prototypes of the tensor-core interpolated dots ----------------
        if (U == 8 && sub > 0) {
            da = q.dots;
#pragma unroll
            for (int i = 0; i < UP_PD; ++i) fetch(nv[i]);
        }
        float best[8][4];            // accumulator-fragment layout: [n-tile][(row g, col 2t) (g, 2t+1) (g+8, 2t) (g+8, 2t+1)]
        auto tr = [&](float sim) {
            const float x = fminf(fmaxf(sim, 0.0f), 1.0f);
            if (EXPK == 0) return x * x;
            if (EXPK == 1) return x * x * sqrtf(x);
            if (EXPK == 2) return x;
            return x > 0.0f ? __powf(x, q.exponent) : 0.0f;
        };
        int c = 0, a_begin = 0, a_end = s_off[1];
        float* oc = q.out;
        // classes without prototypes (max over nothing = -inf -> clamp -> 0, as the generic kernel): zero planes
        auto skip_empty = [&]() {
            while (c < q.C && a_end == a_begin) {
#pragma unroll
                for (int rj = 0; rj < 16; ++rj) {                       // (static indices: obase must stay in registers)
                    const int r = rj >> 3, j = rj & 7;
                    const int joff = U == 4 ? (j >> 1) * Hzs + 2 * (j & 1) * zs : sub * Hzs + j * zs;
                    float* dst = up_out_at(oc, obase[r] + joff);
                    if ((ms0 >> rj) & 1) dst[0] = 0.0f;
                    if ((ms1 >> rj) & 1) dst[1] = 0.0f;
                }
                ++c;
                oc += n_out;
                a_end = c < q.C ? s_off[c + 1] : -1;
            }
        };
        skip_empty();
        // flat loop over all prototypes, unrolled by the prefetch distance so that the ring of in-flight corner dots is
        // indexed statically (a register copy of a value still in flight would wait for it)
#pragma unroll 1
        for (int a0 = 0; a0 < q.A; a0 += UP_PD) {
#pragma unroll
            for (int i = 0; i < UP_PD; ++i) {
                if (a0 + i >= q.A) break;
                const float v00 = nv[i][0], v01 = nv[i][1], v10 = nv[i][2], v11 = nv[i][3];
                // K = [corner hi | corner hi] x [W_hi | W_lo], then [corner mid | corner mid] x the same B pair
                uint32_t ah[4], am[4];
                ah[0] = ptx::pack_bf16x2(v00, v01);
                ah[1] = ptx::pack_bf16x2(v10, v11);
                am[0] = ptx::pack_bf16x2(v00 - bf16_lo_f(ah[0]), v01 - bf16_hi_f(ah[0]));
                am[1] = ptx::pack_bf16x2(v10 - bf16_lo_f(ah[1]), v11 - bf16_hi_f(ah[1]));
                // opaque copies: the fragment is a register quad, the compiler must see four distinct values to keep it alive
                asm volatile("mov.b32 %0, %1;" : "=r"(ah[2]) : "r"(ah[0]));
                asm volatile("mov.b32 %0, %1;" : "=r"(ah[3]) : "r"(ah[1]));
                asm volatile("mov.b32 %0, %1;" : "=r"(am[2]) : "r"(am[0]));
                asm volatile("mov.b32 %0, %1;" : "=r"(am[3]) : "r"(am[1]));
                fetch(nv[i]);
                if (a0 + i == a_begin) {
                    // first prototype of the class: the MMAs write the running maximum directly (no reset, no max)
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        mma_bf16_16816_z(best[j], ah, whi[j], wlo[j]);
                        mma_bf16_16816(best[j], am, whi[j], wlo[j]);
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        float dd[4];
                        mma_bf16_16816_z(dd, ah, whi[j], wlo[j]);
                        mma_bf16_16816(dd, am, whi[j], wlo[j]);
#pragma unroll
                        for (int e = 0; e < 4; ++e) best[j][e] = fmaxf(best[j][e], dd[e]);
                    }
                }
                if (a0 + i + 1 != a_end) continue;
                // ---- last prototype of class c: normalise, clamp(0,1)^e, store ------------------------------
                if (fast_st) {
#pragma unroll
                    for (int r = 0; r < 2; ++r)
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const int joff = U == 4 ? (j >> 1) * Hzs + 2 * (j & 1) * zs : sub * Hzs + j * zs;
                            const float r0 = tr(best[j][2 * r] * inv[r][j][0]), r1 = tr(best[j][2 * r + 1] * inv[r][j][1]);
                            if ((ms0 >> (r * 8 + j)) & 1) *reinterpret_cast<float2*>(up_out_at(oc, obase[r] + joff)) = make_float2(r0, r1);
                        }
                } else {
#pragma unroll
                    for (int r = 0; r < 2; ++r)
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const int joff = U == 4 ? (j >> 1) * Hzs + 2 * (j & 1) * zs : sub * Hzs + j * zs;
                            const float r0 = tr(best[j][2 * r] * inv[r][j][0]), r1 = tr(best[j][2 * r + 1] * inv[r][j][1]);
                            float* dst = up_out_at(oc, obase[r] + joff);
                            if ((ms0 >> (r * 8 + j)) & 1) dst[0] = r0;
                            if ((ms1 >> (r * 8 + j)) & 1) dst[1] = r1;
                        }
                }
                ++c;
                oc += n_out;
                a_begin = a_end;
                a_end = c < q.C ? s_off[c + 1] : -1;
                skip_empty();
            }
        }
    }
}

template <int U>
int launch_upsample_mma(const UpParams& q, cudaStream_t s) {
    auto fdiv = [](int a, int b) { return a >= 0 ? a / b : -((-a + b - 1) / b); };
    const int cz_lo = fdiv(q.z0 - U / 2, U), cz_hi = fdiv(q.z1 - 1 - U / 2, U);
    const int ncz = cz_hi - cz_lo + 1;
    const int tiles_per_x = ceil_div((q.h + 1) * ncz, 16);
    const int cx_lo = fdiv(q.x0 - U / 2, U), cx_hi = fdiv(q.x1 - 1 - U / 2, U);      // x cells overlapping the slab
    const long long total = static_cast<long long>(cx_hi - cx_lo + 1) * tiles_per_x;
    if (total > 0x7fffffff / 4) return -1;
    const size_t smem = (((q.C + 1) * 4 + 15) & ~15) + (U == 8 ? 64 : 8) * 32 * sizeof(uint2);
    const unsigned grid = static_cast<unsigned>((total + 3) / 4);
    const int tt = static_cast<int>(total);
    if (q.exponent == 2.0f) sim_upsample_mma_kernel<U, 0><<<grid, 128, smem, s>>>(q, cz_lo, ncz, tiles_per_x, tt, cx_lo);
    else if (q.exponent == 2.5f) sim_upsample_mma_kernel<U, 1><<<grid, 128, smem, s>>>(q, cz_lo, ncz, tiles_per_x, tt, cx_lo);
    else if (q.exponent == 1.0f) sim_upsample_mma_kernel<U, 2><<<grid, 128, smem, s>>>(q, cz_lo, ncz, tiles_per_x, tt, cx_lo);
    else sim_upsample_mma_kernel<U, 3><<<grid, 128, smem, s>>>(q, cz_lo, ncz, tiles_per_x, tt, cx_lo);
    return 0;
}

// ---------------------------------------------------------------------------------------------
__global__ void class_max_init_kernel(float* out, int C) {
    if (threadIdx.x < C) out[threadIdx.x] = -INFINITY;
}
__global__ void __launch_bounds__(256) class_max_kernel(const float* __restrict__ sims, int64_t n, float* out) {
    const float* s = sims + static_cast<int64_t>(blockIdx.y) * n;
    float m = -INFINITY;
    for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x)
        m = fmaxf(m, __ldg(s + i));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) {
        if (m >= 0.0f) atomicMax(reinterpret_cast<int*>(out + blockIdx.y), __float_as_int(m));
        else atomicMin(reinterpret_cast<unsigned int*>(out + blockIdx.y), __float_as_uint(m));
    }
}


// predict_ntf.py:95-100: (255 / (0.99 * max_c) * sim).to(uint8) -- the C-style float -> int -> uint8 conversion of the
// reference's CPU cast, values in (255, 257.6] wrap -- followed by F.interpolate(mode='nearest'):
// src = min(floor(dst * in/out), in - 1).  The two commute, so only the kept voxels are read.
__global__ void __launch_bounds__(256) quantize_maps_kernel(const float* __restrict__ sims, int W, int H, int zs, int z0, int D,
                                                            const float* __restrict__ class_max, int Wo, int Ho, int Do,
                                                            int zo0, int zos, uint8_t* __restrict__ out) {
    const int c = blockIdx.y;
    const float scale = 255.0f / (0.99f * class_max[c]);
    const float sx = static_cast<float>(W) / Wo, sy = static_cast<float>(H) / Ho, sz = static_cast<float>(D) / Do;
    const int64_t n_out = static_cast<int64_t>(Wo) * Ho * zos;
    const float* src = sims + static_cast<int64_t>(c) * W * H * zs;
    uint8_t* dst = out + static_cast<int64_t>(c) * n_out;
    for (int64_t o = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; o < n_out;
         o += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int oz = static_cast<int>(o % zos) + zo0, oy = static_cast<int>((o / zos) % Ho),
                  ox = static_cast<int>(o / (static_cast<int64_t>(zos) * Ho));
        const int ix = min(static_cast<int>(floorf(ox * sx)), W - 1), iy = min(static_cast<int>(floorf(oy * sy)), H - 1),
                  iz = min(static_cast<int>(floorf(oz * sz)), D - 1) - z0;
        const float v = scale * __ldg(src + (static_cast<int64_t>(ix) * H + iy) * zs + iz);
        dst[o] = static_cast<uint8_t>(static_cast<int>(v));
    }
}

template <typename T>
__global__ void __launch_bounds__(256) labels_kernel(const T* __restrict__ sims, int C, int64_t n,
                                                     const int* __restrict__ thr, int mode, uint8_t* __restrict__ out) {
    for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        if (mode == 0) {  // predict_ntf.py:203-215
            float best = 0.0f;
            int lab = 0;
            for (int c = 0; c < C; ++c) {
                const float s = static_cast<float>(sims[static_cast<int64_t>(c) * n + i]);
                if (s > static_cast<float>(thr[c]) && s > best) { best = s; lab = c + 1; }
            }
            out[i] = static_cast<uint8_t>(lab);
        } else {          // argmax(0), first maximum wins
            float best = static_cast<float>(sims[i]);
            int lab = 0;
            for (int c = 1; c < C; ++c) {
                const float s = static_cast<float>(sims[static_cast<int64_t>(c) * n + i]);
                if (s > best) { best = s; lab = c; }
            }
            out[i] = static_cast<uint8_t>(lab);
        }
    }
}

template <typename T, bool GRAM>
int launch_lowres(const T* feats, int F, int w, int h, int d, const float* protos, int A, float* dots, float* gram,
                  cudaStream_t s) {
    const int64_t n = static_cast<int64_t>(w) * h * d;
    const unsigned grid = static_cast<unsigned>(ceil_div_ll(n, 128));
    int a_base = 0;
    bool first = true;
    while (a_base < A || first) {
        const int rem = A - a_base;
        // Gram is accumulated by the first prototype group only
        if (rem > 16) {
            if (first && GRAM) sim_lowres_kernel<T, 32, true><<<grid, 128, 0, s>>>(feats, F, w, h, d, protos, A, a_base, dots, gram);
            else sim_lowres_kernel<T, 32, false><<<grid, 128, 0, s>>>(feats, F, w, h, d, protos, A, a_base, dots, gram);
            a_base += 32;
            vittf_count_launches(1);
        } else if (rem > 8) {
            if (first && GRAM) sim_lowres_kernel<T, 16, true><<<grid, 128, 0, s>>>(feats, F, w, h, d, protos, A, a_base, dots, gram);
            else sim_lowres_kernel<T, 16, false><<<grid, 128, 0, s>>>(feats, F, w, h, d, protos, A, a_base, dots, gram);
            a_base += 16;
            vittf_count_launches(1);
        } else {
            if (first && GRAM) sim_lowres_kernel<T, 8, true><<<grid, 128, 0, s>>>(feats, F, w, h, d, protos, A, a_base, dots, gram);
            else sim_lowres_kernel<T, 8, false><<<grid, 128, 0, s>>>(feats, F, w, h, d, protos, A, a_base, dots, gram);
            a_base += 8;
            vittf_count_launches(1);
        }
        first = false;
    }
    VITTF_CHECK_CUDA(cudaGetLastError());
    return VITTF_OK;
}

}  // namespace

extern "C" int vittf_sample_prototypes(const void* feats, int feat_dtype, int F, int w, int h, int d, const float* rel,
                                       int A, int mode, float* out, void* stream) {
    VITTF_REQUIRE(feats && rel && out, "vittf_sample_prototypes: null pointer");
    VITTF_REQUIRE(F > 0 && w > 0 && h > 0 && d > 0 && A > 0, "vittf_sample_prototypes: empty problem");
    VITTF_REQUIRE(mode == 0 || mode == 1, "vittf_sample_prototypes: mode must be 0 (nearest) or 1 (trilinear)");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (feat_dtype == VITTF_F16)
        sample_prototypes_kernel<__half><<<A, 128, 0, s>>>(static_cast<const __half*>(feats), F, w, h, d, rel, mode, out);
    else if (feat_dtype == VITTF_F32)
        sample_prototypes_kernel<float><<<A, 128, 0, s>>>(static_cast<const float*>(feats), F, w, h, d, rel, mode, out);
    else
        VITTF_REQUIRE(false, "vittf_sample_prototypes: features must be fp16 or fp32");
    VITTF_CHECK_CUDA(cudaGetLastError());
    vittf_count_launches(1);
    return VITTF_OK;
}

// One A/B switch for the whole stage: VITTF_SIM_GENERIC forces the generic kernels of both passes (any dtype / shape / mode),
// which a test exercises against the same oracle as the tensor-core paths.
static bool sim_generic_only() {
    static const bool on = getenv("VITTF_SIM_GENERIC") != nullptr;
    return on;
}

extern "C" int vittf_sim_lowres_layout(int feat_dtype, int F, int w, int h, int d, const void* feats) {
    const size_t panel = static_cast<size_t>(32) * (F + 8) * 2;
    return !sim_generic_only() && feat_dtype == VITTF_F16 && F > 0 && F % DM_BK == 0 && d % 8 == 0 &&
           (reinterpret_cast<uintptr_t>(feats) & 15) == 0 && (220 * 1024 - panel - 256) / GM_STAGE >= 3;
}

extern "C" int vittf_sim_lowres(const void* feats, int feat_dtype, int F, int w, int h, int d, const float* protos, int A,
                                float* dots, float* gram, int dots_layout, int xa, int xb, void* stream) {
    VITTF_REQUIRE(feats && protos && dots, "vittf_sim_lowres: null pointer");
    VITTF_REQUIRE(F > 0 && w > 0 && h > 0 && d > 0 && A > 0, "vittf_sim_lowres: empty problem");
    VITTF_REQUIRE(xa >= 0 && xb > xa && xb <= w, "vittf_sim_lowres: x-plane range [%d,%d) outside [0,%d)", xa, xb, w);
    const bool fused = vittf_sim_lowres_layout(feat_dtype, F, w, h, d, feats) != 0;
    VITTF_REQUIRE(dots_layout == 0 || (dots_layout == 1 && fused),
                  "vittf_sim_lowres: dots_layout %d is not available for this input (ask vittf_sim_lowres_layout)", dots_layout);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (fused) {
        // fp16 features, F % 32 == 0, d % 8 == 0: dots + Gram from one read of the volume on the tensor cores
        const int rc = launch_lowres_mma(static_cast<const __half*>(feats), F, w, h, d, protos, A, dots, gram, s, dots_layout, xa, xb);
        if (rc != -1) return rc;
        VITTF_REQUIRE(dots_layout == 0, "vittf_sim_lowres: fused pass unavailable for F=%d", F);
    }
    if (feat_dtype == VITTF_F16) {
        const __half* f = static_cast<const __half*>(feats);
        return gram ? launch_lowres<__half, true>(f, F, w, h, d, protos, A, dots, gram, s)
                    : launch_lowres<__half, false>(f, F, w, h, d, protos, A, dots, gram, s);
    } else if (feat_dtype == VITTF_F32) {
        const float* f = static_cast<const float*>(feats);
        return gram ? launch_lowres<float, true>(f, F, w, h, d, protos, A, dots, gram, s)
                    : launch_lowres<float, false>(f, F, w, h, d, protos, A, dots, gram, s);
    }
    VITTF_REQUIRE(false, "vittf_sim_lowres: features must be fp16 or fp32");
    return VITTF_OK;
}

extern "C" int vittf_sim_upsample(const float* dots, const float* gram, int w, int h, int d, int A,
                                  const int* class_offsets, int C, int W, int H, int D, int x0, int x1, int z0, int z1, int mode,
                                  float threshold, float exponent, int dots_layout, float* out, void* stream) {
    VITTF_REQUIRE(dots && class_offsets && out, "vittf_sim_upsample: null pointer");
    VITTF_REQUIRE(dots_layout == 0 || dots_layout == 1, "vittf_sim_upsample: dots_layout must be 0 (A, n_lr) or 1 (n_lr, A4)");
    VITTF_REQUIRE(mode == VITTF_SIM_NS || mode == VITTF_SIM_REFNTF || mode == VITTF_SIM_LEGACY || mode == VITTF_SIM_CLAMP_MEAN,
                  "vittf_sim_upsample: unknown mode %d", mode);
    VITTF_REQUIRE(mode == VITTF_SIM_REFNTF || mode == VITTF_SIM_CLAMP_MEAN || gram, "vittf_sim_upsample: NS/LEGACY modes need the Gram planes");
    VITTF_REQUIRE(C > 0 && A > 0 && W > 0 && H > 0 && D > 0 && z0 >= 0 && z1 > z0 && z1 <= D && x0 >= 0 && x1 > x0 && x1 <= W,
                  "vittf_sim_upsample: bad sizes (C=%d A=%d out=%dx%dx%d x=[%d,%d) z=[%d,%d))", C, A, W, H, D, x0, x1, z0, z1);
    UpParams q{dots, gram, class_offsets, out, w, h, d, A, C, W, H, D, z0, z1, x0, x1, mode, threshold, exponent};
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (dots_layout == 1) {
        // voxel-major dots: NS mode with one integer factor 2 / 4 / 8 on the tcgen05 cell-tile kernel (sim_up_tc.cu)
        VITTF_REQUIRE(mode == VITTF_SIM_NS && vittf_launch_upsample_tc(q, dots_layout, s) == 0,
                      "vittf_sim_upsample: the voxel-major dots layout is only read by the tcgen05 kernel (NS mode, factor 2 / 4 / 8)");
        VITTF_CHECK_CUDA(cudaGetLastError());
        vittf_count_launches(1);
        return VITTF_OK;
    }
    // NS mode, one factor 4 or 8 on all three axes: warp-level mma.sync cell-tile kernel (cheaper per prototype than the
    // tcgen05 kernel: no TMEM round trip per prototype -- profiles/r2_sim_kernels.md)
    if (!sim_generic_only() && mode == VITTF_SIM_NS && (W == 4 * w || W == 8 * w) && H * static_cast<int64_t>(w) == static_cast<int64_t>(h) * W &&
        D * static_cast<int64_t>(w) == static_cast<int64_t>(d) * W && static_cast<int64_t>(x1 - x0) * H * (z1 - z0) < (1ll << 30) &&
        static_cast<int64_t>(w) * h * d < (1ll << 31)) {
        const int rc = W == 4 * w ? launch_upsample_mma<4>(q, s) : launch_upsample_mma<8>(q, s);
        if (rc == 0) {
            VITTF_CHECK_CUDA(cudaGetLastError());
            vittf_count_launches(1);
            return VITTF_OK;
        }
    }
    // every other scale factor, mode and shape
    const int64_t n_out = static_cast<int64_t>(x1 - x0) * H * (z1 - z0);
    int64_t blocks = ceil_div_ll(n_out, 256);
    const int64_t cap = static_cast<int64_t>(vittf_num_sms()) * 32;
    if (blocks > cap) blocks = cap;
    sim_upsample_kernel<<<static_cast<unsigned>(blocks), 256, (C + 1) * sizeof(int), s>>>(q);
    VITTF_CHECK_CUDA(cudaGetLastError());
    vittf_count_launches(1);
    return VITTF_OK;
}

extern "C" int vittf_class_max(const float* sims, int C, int64_t n, float* out, void* stream) {
    VITTF_REQUIRE(sims && out && C > 0 && C <= 1024 && n > 0, "vittf_class_max: bad arguments");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    class_max_init_kernel<<<1, 1024, 0, s>>>(out, C);
    dim3 grid(vittf_num_sms() * 2, C);
    class_max_kernel<<<grid, 256, 0, s>>>(sims, n, out);
    VITTF_CHECK_CUDA(cudaGetLastError());
    vittf_count_launches(2);
    return VITTF_OK;
}

extern "C" int vittf_quantize_maps_u8(const float* sims, int C, int W, int H, int D, int z0, int z1, const float* class_max,
                                      int Wo, int Ho, int Do, int zo0, int zo1, uint8_t* out, void* stream) {
    VITTF_REQUIRE(sims && class_max && out, "vittf_quantize_maps_u8: null pointer");
    VITTF_REQUIRE(C > 0 && C <= 65535 && W > 0 && H > 0 && D > 0 && z0 >= 0 && z1 > z0 && z1 <= D && Wo > 0 && Ho > 0 && Do > 0 &&
                      zo0 >= 0 && zo1 >= zo0 && zo1 <= Do,
                  "vittf_quantize_maps_u8: bad sizes (C=%d in=%dx%dx%d z=[%d,%d) out=%dx%dx%d zo=[%d,%d))", C, W, H, D, z0, z1, Wo,
                  Ho, Do, zo0, zo1);
    if (zo1 == zo0) return VITTF_OK;
    // every requested output plane must read a source plane of the slab
    const float sz = static_cast<float>(D) / Do;
    auto src_plane = [&](int oz) { const int i = static_cast<int>(floorf(oz * sz)); return i < D - 1 ? i : D - 1; };
    VITTF_REQUIRE(src_plane(zo0) >= z0 && src_plane(zo1 - 1) < z1, "vittf_quantize_maps_u8: output planes [%d,%d) read outside the slab [%d,%d)",
                  zo0, zo1, z0, z1);
    const int64_t n_out = static_cast<int64_t>(Wo) * Ho * (zo1 - zo0);
    int64_t blocks = ceil_div_ll(n_out, 256);
    const int64_t cap = static_cast<int64_t>(vittf_num_sms()) * 8;
    if (blocks > cap) blocks = cap;
    dim3 grid(static_cast<unsigned>(blocks), C);
    quantize_maps_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(sims, W, H, z1 - z0, z0, D, class_max, Wo, Ho, Do, zo0,
                                                                              zo1 - zo0, out);
    VITTF_CHECK_CUDA(cudaGetLastError());
    vittf_count_launches(1);
    return VITTF_OK;
}

extern "C" int vittf_labels(const void* sims, int sims_dtype, int C, int64_t n, const int* thresholds_u8, int mode,
                            uint8_t* out, void* stream) {
    VITTF_REQUIRE(sims && out && C > 0 && C < 255 && n > 0, "vittf_labels: bad arguments");
    VITTF_REQUIRE(mode == 1 || thresholds_u8, "vittf_labels: thresholds required for mode 0");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    int64_t blocks = ceil_div_ll(n, 256);
    const int64_t cap = static_cast<int64_t>(vittf_num_sms()) * 16;
    if (blocks > cap) blocks = cap;
    if (sims_dtype == VITTF_U8)
        labels_kernel<uint8_t><<<static_cast<unsigned>(blocks), 256, 0, s>>>(static_cast<const uint8_t*>(sims), C, n, thresholds_u8, mode, out);
    else if (sims_dtype == VITTF_F32)
        labels_kernel<float><<<static_cast<unsigned>(blocks), 256, 0, s>>>(static_cast<const float*>(sims), C, n, thresholds_u8, mode, out);
    else
        VITTF_REQUIRE(false, "vittf_labels: similarity maps must be uint8 or fp32");
    VITTF_CHECK_CUDA(cudaGetLastError());
    vittf_count_launches(1);
    return VITTF_OK;
}
