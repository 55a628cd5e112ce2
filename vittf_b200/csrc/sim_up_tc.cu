// Pass 2 of the prototype-similarity stage on the 5th-generation tensor cores (north-star order: trilinear
// up-sampling of the FEATURES -> L2 normalise -> dot -> clamp(0,1)^e -> per-class MAX, never materialised;
// reference ops being composed: /root/reference/predict_ntf.py:65,87, old/cluster_dino.py:307,318,322).
//
// With align_corners=False and an integer factor U, a block of outputs that shares its low-res corner voxels is
//     out[o] = sum_{corner c} W[o][c] * dots[corner c]
// with the SAME weight matrix W for every block (borders clamp corner INDICES, not weights).  So for one prototype
//     D (128 blocks x NOUT outputs) = Corners (128 x KC) * W^T (KC x NOUT)
// is one tcgen05.mma tile: M = 128 rows = 128 output blocks, N = NOUT, K = KC corner dots (kind::tf32, K = 8 per
// instruction).  fp32 accuracy on tf32 operands: a row holds [hi(KC) | lo(KC)] with hi = the dot truncated to tf32 and
// lo = dot - hi (exact), W^T rows hold [W | W]; every weight is an odd multiple of 2^-k with <= 11 significant bits for
// U in {2, 4} (exact in tf32), for U = 8 (12 bits) a third group of K-slices adds hi * W_lo.  Relative error <= 2^-21.
//
// Block shapes (outputs per row x corners per row), chosen so that a row's z runs are 16-byte aligned:
//   U = 4 : 4 x 4 x 4 outputs of one (cx, cy) cell and one ALIGNED z chunk [4j, 4j+4)  -- corners 2 x 2 x 3 (z: j-1, j, j+1)
//   U = 8 : 1 x 8 x 8 outputs: one output x plane of a (cy, cz) cell (z run 8cz+4 .. 8cz+11) -- corners 2 x 2 x 2
//   U = 2 : 2 x 2 x 4 outputs of one (cx, cy) cell and one aligned z chunk [4j, 4j+4)   -- corners 2 x 2 x 4 (z: 2j-1 .. 2j+2)
//
// One persistent CTA per SM, 512 threads (setmaxnreg: 128 / 128 / 216 / 40 registers per warpgroup), three roles connected by
// mbarriers (no TMA: the operand is gathered):
//   warps 8-11 producers : thread r owns row r of the tile.  Per tile it contracts the cell's Gram scalars (pass 1's 14
//                          planes) into 1 / |interp(f)| for its NOUT outputs and parks them in TMEM (tcgen05.st); per
//                          prototype it gathers the KC corner dots, splits hi / lo and parks the row in TMEM (tcgen05.st):
//                          the A operand of the MMA comes from TMEM, W^T (built once per CTA, 128-byte swizzled) from smem.
//   warp 12    issuer    : tcgen05.mma (2-4 K-slices per prototype) into one of 4 accumulator buffers of 64 TMEM columns.
//   warps 0-7  epilogue  : two warpgroups, each owning half of the accumulator columns: a thread reads its 32 outputs of row r
//                          (tcgen05.ld, two prototypes per round trip), keeps the running class maximum in registers
//                          (rows = blocks: no shuffles), and at the end of a class multiplies by the parked 1 / norm,
//                          saturates, applies the exponent and stores 16-byte runs along z.
// Per output voxel and prototype the CUDA cores execute one FMNMX (+ ~0.8 gather / split instructions on the producer
// side); the kernel is bound by the HBM writes of the maps (C * 4 B per voxel).
#include <stdlib.h>

#include "common.cuh"
#include "sim_internal.h"

namespace {

constexpr int TC_THREADS = 512;             // warpgroups: epilogue A | epilogue B | producers | issuer (+3 idle warps that give up their registers)
constexpr int TC_NBUF = 4;                   // accumulator buffers
constexpr int TC_BUF_COLS = 64;
constexpr int TC_COL_NORM = TC_NBUF * TC_BUF_COLS;    // two 64-column buffers of 1 / norm
constexpr int TC_STAGES = 2;                 // A-operand stages IN TMEM: 128 lanes x 64 columns = TWO prototypes (one tf32 per column)
constexpr int TC_STAGE_COLS = 64;
constexpr int TC_COL_A = TC_COL_NORM + 2 * TC_BUF_COLS;
constexpr int TC_TMEM_COLS = 512;            // 4 x 64 accumulators + 2 x 64 norms + 2 x 64 operand stages

template <int U> struct Geo;
template <> struct Geo<4> { static constexpr int BX = 4, BY = 4, BZ = 4, CZ = 3, NSUB = 1; };
template <> struct Geo<8> { static constexpr int BX = 1, BY = 8, BZ = 8, CZ = 2, NSUB = 8; };
template <> struct Geo<2> { static constexpr int BX = 2, BY = 2, BZ = 4, CZ = 4, NSUB = 1; };

// z weights of output kz inside a row: the two contributing corners are zlo(kz), zlo(kz) + 1 with weights (wza, wzb)
template <int U> __host__ __device__ constexpr int zlo(int kz) { return U == 8 ? 0 : U == 4 ? (kz >> 1) : ((kz + 1) >> 1); }
template <int U> __host__ __device__ constexpr float wzb(int kz) {
    // F.interpolate(align_corners=False): src = (o + 0.5) / U - 0.5, weight of the upper corner = frac(src)
    return U == 8 ? (kz + 0.5f) / 8.0f
         : U == 4 ? (kz == 0 ? 0.625f : kz == 1 ? 0.875f : kz == 2 ? 0.125f : 0.375f)
                  : ((kz & 1) ? 0.25f : 0.75f);
}
template <int U> __host__ __device__ constexpr float wza(int kz) { return 1.0f - wzb<U>(kz); }
// x / y weights of output k in a cell: upper corner (k + 0.5) / U
template <int U> __host__ __device__ constexpr float wcell(int k, int upper) { return upper ? (k + 0.5f) / U : 1.0f - (k + 0.5f) / U; }

struct TcParams {
    const float* dots;
    const float* gram;
    const int* class_offsets;
    float* out;
    int w, h, d, A, C;
    int W, H, D, z0, z1;
    int x0, x1;        // output x-slab
    int xrow0;         // U = 2, 4: first x cell (cx) of the slab
    float exponent;
    int nz;            // rows along z per (x, y): z chunks (U = 2, 4) or z cells (U = 8) overlapping the slab
    int zrow0;         // first z chunk / z cell
    int rows_per_x;    // U = 8: rows of one output x plane, (h + 1) * nz
    int tiles_per_x;   // U = 8: tiles of 128 rows per output x plane
    int64_t n_rows;    // U = 2, 4: all rows
    int n_tiles;
    int vec_ok;        // z runs may be stored as 16-byte vectors
    int A4;            // row pitch of the voxel-major dots: A rounded up to a multiple of 4
};

// D[tmem] (+)= A[tmem] * B[smem]^T, tf32: the A operand is one row per TMEM lane, one K element per 32-bit column
__device__ __forceinline__ void umma_ts_tf32(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n"
        "}\n" ::"r"(d_tmem),
        "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// tf32 x tf32 -> fp32, both operands K-major, dense
__host__ __device__ constexpr uint32_t idesc_tf32_f32(int m, int n) {
    return (1u << 4) | (2u << 7) | (2u << 10) | (static_cast<uint32_t>(n >> 3) << 17) | (static_cast<uint32_t>(m >> 4) << 24);
}
// 16-byte piece `chunk` of row `row` in a 128-byte-swizzled K-major tile (1024-byte aligned base)
__device__ __forceinline__ uint32_t swz(uint32_t base, int row, int chunk) {
    return base + static_cast<uint32_t>(row) * 128u + static_cast<uint32_t>((chunk ^ (row & 7)) << 4);
}
__device__ __forceinline__ float tf32_hi(float v) { return __uint_as_float(__float_as_uint(v) & 0xFFFFE000u); }
__device__ __forceinline__ float rsqrt_approx(float x) { float y; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

// One lane polls, the warp joins: 12 warps spinning on try_wait with all their lanes take the issue slots the working warps
// of the same scheduler need (measured in the attention kernel too).
__device__ __forceinline__ void warp_wait(uint64_t* bar, uint32_t parity) {
    if ((threadIdx.x & 31) == 0) ptx::mbar_wait_quiet(bar, parity);
    __syncwarp();
}

template <int EXPK>
__device__ __forceinline__ float pow_sat(float x, float e) {      // x already in [0, 1]
    if (EXPK == 0) return x * x;
    if (EXPK == 1) return x;
    if (EXPK == 2) return x * x * sqrtf(x);
    return x > 0.0f ? __powf(x, e) : 0.0f;
}

// Row -> block coordinates.  ix: cell cx + 1 (U = 2, 4) or output x (U = 8); iy: cell cy + 1; iz: z chunk / z cell.
template <int U>
__device__ __forceinline__ bool decode_row(const TcParams& p, int tile, int r, int& ix, int& iy, int& iz) {
    if (U == 8) {
        const int xi = tile / p.tiles_per_x;
        ix = p.x0 + xi;                                   // output x plane
        const int rr = (tile - xi * p.tiles_per_x) * 128 + r;
        const bool live = rr < p.rows_per_x;
        const int rc = live ? rr : p.rows_per_x - 1;
        iy = rc / p.nz;
        iz = p.zrow0 + (rc - iy * p.nz);
        return live;
    } else {
        int64_t rr = static_cast<int64_t>(tile) * 128 + r;
        const bool live = rr < p.n_rows;
        rr = live ? rr : p.n_rows - 1;
        const int per_x = (p.h + 1) * p.nz;
        const int xi = static_cast<int>(rr / per_x);
        ix = p.xrow0 + xi + 1;                            // cell cx + 1
        const int rem = static_cast<int>(rr - static_cast<int64_t>(xi) * per_x);
        iy = rem / p.nz;
        iz = p.zrow0 + (rem - iy * p.nz);
        return live;
    }
}
__device__ __forceinline__ int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

#ifdef SIM_TRACE
__device__ long long g_tc_trace[3][256];
#define TC_TRACE(role, slot) do { if (blockIdx.x == 0 && (slot) < 256) g_tc_trace[role][slot] = clock64(); } while (0)
#else
#define TC_TRACE(role, slot) do {} while (0)
#endif

template <int U, int EXPK>
__global__ void __launch_bounds__(TC_THREADS, 1) sim_upsample_tc_kernel(TcParams p) {
    using G = Geo<U>;
    constexpr int BX = G::BX, BY = G::BY, BZ = G::BZ, CZ = G::CZ, KC = 4 * G::CZ, NOUT = BX * BY * BZ;
    constexpr int NSL = 2 * KC / 8;                       // K-slices of [hi | lo]
    constexpr int B_BYTES = NOUT * 128;                   // one W^T matrix
    extern __shared__ uint8_t tc_smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(tc_smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* s_b = smem;                                          // NSUB x B_BYTES (B_BYTES is a multiple of 1024 or the only one)
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_b + ((G::NSUB * B_BYTES + 1023) & ~1023));
    uint64_t* a_full = bars;                    // [TC_STAGES] 128 producer arrivals
    uint64_t* a_empty = a_full + TC_STAGES;     // [TC_STAGES] tcgen05.commit
    uint64_t* d_full = a_empty + TC_STAGES;     // [TC_NBUF]   tcgen05.commit
    uint64_t* d_empty = d_full + TC_NBUF;       // [TC_NBUF]   128 epilogue arrivals
    uint64_t* n_full = d_empty + TC_NBUF;       // [2] 128 producer arrivals: 1 / norm of the tile is in TMEM
    uint64_t* n_empty = n_full + 2;             // [2] 128 epilogue arrivals
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(n_empty + 2);
    int* s_off = reinterpret_cast<int*>(tmem_slot + 2);           // class offsets

    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
    if (threadIdx.x == 0) {
        for (int i = 0; i < TC_STAGES; ++i) { ptx::mbar_init(&a_full[i], 128); ptx::mbar_init(&a_empty[i], 1); }
        constexpr int EPI_THREADS = NOUT > 32 ? 256 : 128;       // both epilogue warpgroups or only the first
        for (int i = 0; i < TC_NBUF; ++i) { ptx::mbar_init(&d_full[i], 1); ptx::mbar_init(&d_empty[i], EPI_THREADS); }
        for (int i = 0; i < 2; ++i) { ptx::mbar_init(&n_full[i], 128); ptx::mbar_init(&n_empty[i], EPI_THREADS); }
        ptx::fence_barrier_init();
    }
    if (warp == 12) ptx::tmem_alloc<TC_TMEM_COLS>(tmem_slot);
    for (int i = threadIdx.x; i <= p.C; i += TC_THREADS) s_off[i] = p.class_offsets[i];
    // W^T: row = output j = (kx * BY + ky) * BZ + kz (U = 8: one matrix per kx, j = ky * 8 + kz), floats [W(KC) | W(KC) | W_lo(KC)],
    // corner k = (xi * 2 + yi) * CZ + zi
    for (int i = threadIdx.x; i < G::NSUB * NOUT * 32; i += TC_THREADS) {
        const int f = i & 31, j = (i >> 5) % NOUT, sub = i / (32 * NOUT);
        const int kz = j % BZ, ky = (j / BZ) % BY, kx = U == 8 ? sub : j / (BZ * BY);
        float v = 0.0f;
        const int part = f / KC, k = f - part * KC;               // part 0, 1: W (tf32 part), 2: W_lo (U = 8)
        if (part < (U == 8 ? 3 : 2)) {
            const int zi = k % CZ, yi = (k / CZ) & 1, xi = k / (2 * CZ);
            const float wz = zi == zlo<U>(kz) ? wza<U>(kz) : zi == zlo<U>(kz) + 1 ? wzb<U>(kz) : 0.0f;
            const float wfull = wcell<U>(kx, xi) * wcell<U>(ky, yi) * wz;      // exact products of dyadic fractions
            v = part < 2 ? tf32_hi(wfull) : wfull - tf32_hi(wfull);
        }
        const uint32_t addr = swz(ptx::smem_u32(s_b) + sub * B_BYTES, j, f >> 2) + (f & 3) * 4;
        asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
    }
    ptx::fence_proxy_async();                   // W^T is read by the tensor core (async proxy)
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t n_lr = static_cast<uint32_t>(p.w) * p.h * p.d;
    const int zs = p.z1 - p.z0;

    if (warp >= 12) {
        // ---------------------------------------------------------------- MMA issuer (one elected lane of warp 12)
        ptx::setmaxnreg_dec<40>();
        if (warp == 12 && ptx::elect_one()) {
            constexpr uint32_t idesc = idesc_tf32_f32(128, NOUT);
            const uint64_t b_desc0 = ptx::smem_desc_k_sw128(ptx::smem_u32(s_b));
            uint32_t g = 0, gs = 0;                                  // prototypes / operand stages (pairs) issued so far
            for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
                const int sub = U == 8 ? ((p.x0 + tile / p.tiles_per_x + 4) & 7) : 0;  // kx of the tile's output x plane
                const uint64_t b_desc = b_desc0 + static_cast<uint64_t>(sub * (B_BYTES >> 4));
                for (int a = 0; a < p.A; a += 2, ++gs) {
                    const uint32_t st = gs % TC_STAGES;
                    ptx::mbar_wait_quiet(&a_full[st], (gs / TC_STAGES) & 1);
                    for (int q = 0; q < 2 && a + q < p.A; ++q, ++g) {
                        const uint32_t buf = g % TC_NBUF;
                        ptx::mbar_wait_quiet(&d_empty[buf], ((g / TC_NBUF) & 1) ^ 1);
                        ptx::tc_fence_after();
                        const uint32_t a_t = tmem_base + TC_COL_A + st * TC_STAGE_COLS + 32 * q;
                        const uint32_t d_t = tmem_base + buf * TC_BUF_COLS;
#ifndef TC_EXP_NO_MMA
#pragma unroll
                        for (int s = 0; s < NSL; ++s) umma_ts_tf32(d_t, a_t + 8 * s, b_desc + 2 * s, idesc, s != 0);
                        if (U == 8) {
#pragma unroll
                            for (int s = 0; s < KC / 8; ++s) umma_ts_tf32(d_t, a_t + 8 * s, b_desc + 2 * (NSL + s), idesc, 1);
                        }
#else
                        (void)a_t; (void)d_t; (void)b_desc; (void)idesc;
#endif
                        ptx::tc_commit(&d_full[buf]);
                        TC_TRACE(1, g);
                    }
                    ptx::tc_commit(&a_empty[st]);
                }
            }
        }
    } else if (warp >= 8) {
        // ---------------------------------------------------------------- producers: 1 / norm + corner-dot rows
        ptx::setmaxnreg_inc<216>();
        const int r = threadIdx.x - 256;
        const uint32_t lane_base = static_cast<uint32_t>((warp & 3) * 32) << 16;
        uint32_t g = 0, tn = 0;
        for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++tn) {
            int ix, iy, iz;
            if (r == 0) TC_TRACE(0, 200 + 2 * tn);
            decode_row<U>(p, tile, r, ix, iy, iz);
            // clamped corner coordinates
            int xc[2], yc[2], zc[CZ];
            float wx0 = 0.0f, wx1 = 0.0f;             // U = 8: runtime x weights of the row's output plane
            {
                int cx;
                if (U == 8) {
                    cx = (ix + 4) / 8 - 1;            // floor((ox - 4) / 8)
                    const int kx = ix - 4 - 8 * cx;
                    wx1 = (kx + 0.5f) / 8.0f;
                    wx0 = 1.0f - wx1;
                } else {
                    cx = ix - 1;
                }
                xc[0] = clampi(cx, 0, p.w - 1);
                xc[1] = clampi(cx + 1, 0, p.w - 1);
                const int cy = iy - 1;
                yc[0] = clampi(cy, 0, p.h - 1);
                yc[1] = clampi(cy + 1, 0, p.h - 1);
                const int zb = U == 8 ? iz : U == 4 ? iz - 1 : 2 * iz - 1;
#pragma unroll
                for (int i = 0; i < CZ; ++i) zc[i] = clampi(zb + i, 0, p.d - 1);
            }
            uint32_t idx[KC];
#pragma unroll
            for (int k = 0; k < KC; ++k) idx[k] = (static_cast<uint32_t>(xc[k / (2 * CZ)]) * p.h + yc[(k / CZ) & 1]) * p.d + zc[k % CZ];
            // ---- Gram scalars of the row's corner pairs, ALL requested before anything is consumed (volatile loads keep
            // their program order: ~70 L2 round trips overlap instead of queueing behind each other) ----
            auto ldv = [](const float* q) { float t; asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(t) : "l"(q)); return t; };
            auto gram_ld = [&](int xa, int ya, int za, int xb, int yb, int zb) {
                int dx = xb - xa, dy = yb - ya, dz = zb - za;
                const bool swap = dx < 0 || (dx == 0 && (dy < 0 || (dy == 0 && dz < 0)));
                const int ax = swap ? xb : xa, ay = swap ? yb : ya, az = swap ? zb : za;
                if (swap) { dx = -dx; dy = -dy; dz = -dz; }
                const int slot = dx == 1 ? 1 + (dy + 1) * 3 + (dz + 1) : dy == 1 ? 11 + dz : dz == 1 ? 13 : 0;
                return ldv(p.gram + static_cast<size_t>(slot) * n_lr + (static_cast<uint32_t>(ax) * p.h + ay) * p.d + az);
            };
            constexpr int NT = U == 8 ? 10 : 10 * BZ;     // contracted Gram the norm arithmetic works from
            float T[NT];
            if constexpr (U == 8) {
                // x first (runtime weights of the plane): Gx over the unordered pairs of (y, z) corners p = yi * 2 + zi
                float g00[10], g01[10], g10[10], g11[10];
                int pi = 0;
#pragma unroll
                for (int pa = 0; pa < 4; ++pa)
#pragma unroll
                    for (int qb = pa; qb < 4; ++qb, ++pi) {
                        const int ya = yc[pa >> 1], za = zc[pa & 1], yb = yc[qb >> 1], zb2 = zc[qb & 1];
                        g00[pi] = gram_ld(xc[0], ya, za, xc[0], yb, zb2);
                        g01[pi] = gram_ld(xc[0], ya, za, xc[1], yb, zb2);
                        g10[pi] = pa == qb ? 0.0f : gram_ld(xc[1], ya, za, xc[0], yb, zb2);
                        g11[pi] = gram_ld(xc[1], ya, za, xc[1], yb, zb2);
                    }
                pi = 0;
#pragma unroll
                for (int pa = 0; pa < 4; ++pa)
#pragma unroll
                    for (int qb = pa; qb < 4; ++qb, ++pi)
                        T[pi] = wx0 * wx0 * g00[pi] + wx0 * wx1 * (g01[pi] + (pa == qb ? g01[pi] : g10[pi])) + wx1 * wx1 * g11[pi];
            } else {
                // z first: T[pair of (x, y) corners][kz]
                float gs[10][CZ], gu[10][CZ - 1], gl[10][CZ - 1];
                int pi = 0;
#pragma unroll
                for (int pa = 0; pa < 4; ++pa)
#pragma unroll
                    for (int qb = pa; qb < 4; ++qb, ++pi) {
                        const int xa = xc[pa >> 1], ya = yc[pa & 1], xb = xc[qb >> 1], yb = yc[qb & 1];
#pragma unroll
                        for (int zi = 0; zi < CZ; ++zi) gs[pi][zi] = gram_ld(xa, ya, zc[zi], xb, yb, zc[zi]);
#pragma unroll
                        for (int zi = 0; zi < CZ - 1; ++zi) {
                            gu[pi][zi] = gram_ld(xa, ya, zc[zi], xb, yb, zc[zi + 1]);
                            gl[pi][zi] = pa == qb ? 0.0f : gram_ld(xa, ya, zc[zi + 1], xb, yb, zc[zi]);
                        }
                    }
                pi = 0;
#pragma unroll
                for (int pa = 0; pa < 4; ++pa)
#pragma unroll
                    for (int qb = pa; qb < 4; ++qb, ++pi)
#pragma unroll
                        for (int kz = 0; kz < BZ; ++kz) {
                            const int zl = zlo<U>(kz);
                            const float a0 = wza<U>(kz), a1 = wzb<U>(kz);
                            T[pi * BZ + kz] = a0 * a0 * gs[pi][zl] + a0 * a1 * (gu[pi][zl] + (pa == qb ? gu[pi][zl] : gl[pi][zl])) +
                                              a1 * a1 * gs[pi][zl + 1];
                        }
            }
            // ---- corner dots: pass 1 wrote them voxel-major, (n_lr, A4), so ONE 16-byte load brings a corner's dots with 4
            // prototypes.  PG groups of 4 prototypes are in flight per thread (requested a group ahead of their use: the
            // gather is a chain of L2 round trips); the first PG groups go out now and overlap the norm arithmetic below.
            // Refills are UNCONDITIONAL (group index clamped): a conditional one makes the compiler load into a temporary and
            // copy it into the ring slot -- a copy that waits for the load it was meant to hide.
            constexpr int PG = KC == 8 ? 3 : (KC == 12 ? 2 : 1);
            const int ngroups = (p.A + 3) >> 2;
            float4 v4[PG][KC];
            auto ldv4 = [](const float* q) {
                float4 t;
                asm volatile("ld.global.nc.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(t.x), "=f"(t.y), "=f"(t.z), "=f"(t.w) : "l"(q));
                return t;
            };
            uint32_t voff[KC];
#pragma unroll
            for (int k = 0; k < KC; ++k) voff[k] = idx[k] * static_cast<uint32_t>(p.A4);
#pragma unroll
            for (int gi = 0; gi < PG; ++gi) {
                const int grp = gi < ngroups ? gi : ngroups - 1;
#pragma unroll
                for (int k = 0; k < KC; ++k) v4[gi][k] = ldv4(p.dots + voff[k] + 4 * grp);
            }

            // ---- 1 / |interp(f)| of the row's NOUT outputs -> TMEM ----
            const uint32_t t_norm = tmem_base + lane_base + TC_COL_NORM + (tn & 1) * TC_BUF_COLS;
            warp_wait(&n_empty[tn & 1], ((tn >> 1) & 1) ^ 1);
            ptx::tc_fence_after();
            {
                constexpr int CH = NOUT < 32 ? NOUT : 32;
#pragma unroll
                for (int hh = 0; hh < NOUT / CH; ++hh) {
                    uint32_t nv[CH];
#pragma unroll
                    for (int i = 0; i < CH; ++i) {
                        const int j = hh * CH + i, kz = j % BZ, ky = (j / BZ) % BY, kx = j / (BZ * BY);
                        float n2 = 0.0f;
                        if constexpr (U == 8) {
                            const float a0 = wza<U>(kz), a1 = wzb<U>(kz), b0 = wcell<U>(ky, 0), b1 = wcell<U>(ky, 1);
                            // pairs in T: (0,0) (0,1) (0,2) (0,3) (1,1) (1,2) (1,3) (2,2) (2,3) (3,3), corner = yi * 2 + zi
                            const float t00 = a0 * a0 * T[0] + 2.0f * a0 * a1 * T[1] + a1 * a1 * T[4];
                            const float t11 = a0 * a0 * T[7] + 2.0f * a0 * a1 * T[8] + a1 * a1 * T[9];
                            const float t01 = a0 * a0 * T[2] + a0 * a1 * (T[3] + T[5]) + a1 * a1 * T[6];
                            n2 = b0 * b0 * t00 + 2.0f * b0 * b1 * t01 + b1 * b1 * t11;
                        } else {
                            int pj = 0;
#pragma unroll
                            for (int pa = 0; pa < 4; ++pa)
#pragma unroll
                                for (int qb = pa; qb < 4; ++qb, ++pj) {
                                    const float wp = wcell<U>(kx, pa >> 1) * wcell<U>(ky, pa & 1), wq = wcell<U>(kx, qb >> 1) * wcell<U>(ky, qb & 1);
                                    n2 = fmaf((pa == qb ? 1.0f : 2.0f) * wp * wq, T[pj * BZ + kz], n2);
                                }
                        }
                        nv[i] = __float_as_uint(rsqrt_approx(fmaxf(n2, 1e-24f)));      // 1 / max(|v|, 1e-12), F.normalize
                    }
                    if constexpr (CH == 32) ptx::tmem_st32(t_norm + hh * 32, nv);
                    else ptx::tmem_st16(t_norm, nv);
                }
            }
            ptx::tc_wait_st();
            ptx::tc_fence_before();
            ptx::mbar_arrive(&n_full[tn & 1]);
            if (r == 0) TC_TRACE(0, 201 + 2 * tn);

            // ---- one operand row per prototype, [hi(KC) | lo(KC)], parked in TMEM (tcgen05.st: lane = row): no shared-memory
            // staging and no generic->async proxy fence -- that fence is a MEMBAR which also waits for the prefetched
            // loads in flight and would put one L2 round trip back into every prototype.  A stage holds TWO prototypes
            // (columns [0, 2 KC) and [32, 32 + 2 KC)): one barrier round trip and one tcgen05.wait::st per pair. ----
            const uint32_t t_a = tmem_base + lane_base + TC_COL_A;
            auto put_row = [&](const float (&x)[KC], uint32_t dst) {
                uint32_t row[2 * KC];
#pragma unroll
                for (int k = 0; k < KC; ++k) {
                    const float hi = tf32_hi(x[k]);
                    row[k] = __float_as_uint(hi);
                    row[KC + k] = __float_as_uint(x[k] - hi);
                }
                if constexpr (KC == 16) {
                    ptx::tmem_st32(dst, row);
                } else if constexpr (KC == 8) {
                    ptx::tmem_st16(dst, row);
                } else {
                    uint32_t r0[16], r1[8];
#pragma unroll
                    for (int k = 0; k < 16; ++k) r0[k] = row[k];
#pragma unroll
                    for (int k = 0; k < 8; ++k) r1[k] = row[16 + k];
                    ptx::tmem_st16(dst, r0);
                    ptx::tmem_st8(dst + 16, r1);
                }
            };
            for (int g0 = 0; g0 < ngroups; g0 += PG) {
#pragma unroll
                for (int gi = 0; gi < PG; ++gi) {
                    const int grp = g0 + gi;
                    float x[4][KC];
#pragma unroll
                    for (int k = 0; k < KC; ++k) {
                        x[0][k] = v4[gi][k].x;
                        x[1][k] = v4[gi][k].y;
                        x[2][k] = v4[gi][k].z;
                        x[3][k] = v4[gi][k].w;
                    }
                    {
                        const int gn = grp + PG < ngroups ? grp + PG : ngroups - 1;
#pragma unroll
                        for (int k = 0; k < KC; ++k) v4[gi][k] = ldv4(p.dots + voff[k] + 4 * gn);
                    }
#pragma unroll
                    for (int pr = 0; pr < 2; ++pr) {
                        const int a = 4 * grp + 2 * pr;
                        if (a < p.A) {
                            const uint32_t st = g % TC_STAGES;
                            warp_wait(&a_empty[st], ((g / TC_STAGES) & 1) ^ 1);
                            ptx::tc_fence_after();
                            const uint32_t dst = t_a + st * TC_STAGE_COLS;
#ifndef TC_EXP_NO_PROD
                            put_row(x[2 * pr], dst);
                            if (a + 1 < p.A) put_row(x[2 * pr + 1], dst + 32);
                            ptx::tc_wait_st();
#else
                            (void)dst;
#endif
                            ptx::tc_fence_before();
                            ptx::mbar_arrive(&a_full[st]);
                            if (r == 0) TC_TRACE(0, 2 * g);
                            ++g;
                        }
                    }
                }
            }
        }
    } else {
        // ---------------------------------------------------------------- epilogue: class max, normalise, store
        // Two warpgroups split the accumulator COLUMNS (warps 0-3: outputs 0..31 of a row, warps 4-7: 32..63): a single
        // warp per scheduler cannot hide its own dependent-issue and TMEM-read latencies (measured: 0.19 IPC).
        constexpr int CW = NOUT < 32 ? NOUT : 32;              // columns per thread
        constexpr int RUNS = CW / BZ;                          // z runs per thread
        const int half = warp >> 2;
        if (half * CW >= NOUT) {
            // (U = 2: 16 outputs per row, the second warpgroup has none)
        } else {
        const int r = threadIdx.x & 127;
        const uint32_t lane_base = static_cast<uint32_t>((warp & 3) * 32) << 16;
        const uint32_t col0 = half * CW;
        const int64_t n_out = static_cast<int64_t>(p.x1 - p.x0) * p.H * zs;
        uint32_t g = 0, tn = 0;
        for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++tn) {
            int ix, iy, iz;
            const bool live = decode_row<U>(p, tile, r, ix, iy, iz);
            // output coordinates of (kx, ky, kz) = (0, 0, 0) and validity masks
            const int ox0 = U == 8 ? ix : U * (ix - 1) + U / 2;
            const int oy0 = U * (iy - 1) + U / 2;
            const int oz0 = U == 8 ? 8 * iz + 4 : 4 * iz;
            uint32_t xmask = 0, ymask = 0, zmask = 0;
#pragma unroll
            for (int k = 0; k < BX; ++k) xmask |= (ox0 + k >= p.x0 && ox0 + k < p.x1) ? (1u << k) : 0u;
#pragma unroll
            for (int k = 0; k < BY; ++k) ymask |= (oy0 + k >= 0 && oy0 + k < p.H) ? (1u << k) : 0u;
#pragma unroll
            for (int k = 0; k < BZ; ++k) zmask |= (oz0 + k >= p.z0 && oz0 + k < p.z1) ? (1u << k) : 0u;
            if (!live) xmask = 0;
            const bool zfull = zmask == (1u << BZ) - 1 && p.vec_ok;
            // run q of this thread = (kx, ky) = divmod(half * RUNS + q, BY): its store offset and its validity bit
            uint32_t runmask = 0;
#pragma unroll
            for (int q = 0; q < RUNS; ++q) {
                const int run = half * RUNS + q;               // `half` is warp-uniform, q compile-time
                runmask |= (((xmask >> (run / BY)) & 1u) & ((ymask >> (run % BY)) & 1u)) << q;
            }
            const int64_t obase = (static_cast<int64_t>(ox0 - p.x0 + (half * RUNS) / BY) * p.H + oy0 + (half * RUNS) % BY) * zs + (oz0 - p.z0);

            float cls[CW];
#pragma unroll
            for (int j = 0; j < CW; ++j) cls[j] = -INFINITY;
            const uint32_t t_norm = tmem_base + lane_base + TC_COL_NORM + (tn & 1) * TC_BUF_COLS + col0;
            bool norm_ready = false;
            int c = 0;
            auto finalize = [&](int cc) {
                if (!norm_ready) {
                    warp_wait(&n_full[tn & 1], (tn >> 1) & 1);
                    ptx::tc_fence_after();
                    norm_ready = true;
                }
                float* oc = p.out + static_cast<int64_t>(cc) * n_out + obase;
                uint32_t nv[CW];
                if constexpr (CW == 32) ptx::tmem_ld32(t_norm, nv);
                else ptx::tmem_ld16(t_norm, nv);
                ptx::tc_wait_ld();
#pragma unroll
                for (int q = 0; q < RUNS; ++q) {
                    // runs of this thread are consecutive (kx, ky): (RUNS divides BY or is a multiple of it)
                    const int dq_x = RUNS > BY ? q / BY : 0, dq_y = RUNS > BY ? q % BY : q;
                    float o[BZ];
#pragma unroll
                    for (int k = 0; k < BZ; ++k) {
                        const int j = q * BZ + k;
                        o[k] = pow_sat<EXPK>(__saturatef(cls[j] * __uint_as_float(nv[j])), p.exponent);
                        cls[j] = -INFINITY;
                    }
                    if ((runmask >> q) & 1u) {
                        float* dst = oc + (static_cast<int64_t>(dq_x) * p.H + dq_y) * zs;
                        if (zfull) {
#pragma unroll
                            for (int k = 0; k < BZ; k += 4) *reinterpret_cast<float4*>(dst + k) = make_float4(o[k], o[k + 1], o[k + 2], o[k + 3]);
                        } else {
#pragma unroll
                            for (int k = 0; k < BZ; ++k)
                                if ((zmask >> k) & 1u) dst[k] = o[k];
                        }
                    }
                }
            };
            // Two prototypes of the same class per TMEM round trip (tcgen05.wait::ld waits for every outstanding load, so
            // the read latency is paid once per pair) and ONE 3-input maximum per output for the pair.
            int a = 0;
            while (a < p.A) {
                while (c < p.C && s_off[c + 1] <= a) finalize(c++);       // classes that ended before prototype a (incl. empty ones)
                const uint32_t buf = g % TC_NBUF;
                const uint32_t t_d = tmem_base + lane_base + buf * TC_BUF_COLS + col0;
                const bool two = a + 1 < s_off[c + 1];                    // a + 1 is a prototype of the same class
                uint32_t v0[CW];
                warp_wait(&d_full[buf], (g / TC_NBUF) & 1);
                ptx::tc_fence_after();
#ifndef TC_EXP_NO_LD
                if constexpr (CW == 32) ptx::tmem_ld32(t_d, v0);
                else ptx::tmem_ld16(t_d, v0);
#else
#pragma unroll
                for (int i = 0; i < CW; ++i) v0[i] = t_d + i;
#endif
                if (two) {
                    const uint32_t buf1 = (g + 1) % TC_NBUF;
                    const uint32_t t_d1 = tmem_base + lane_base + buf1 * TC_BUF_COLS + col0;
                    uint32_t u0[CW];
                    warp_wait(&d_full[buf1], ((g + 1) / TC_NBUF) & 1);
                    ptx::tc_fence_after();
#ifndef TC_EXP_NO_LD
                    if constexpr (CW == 32) ptx::tmem_ld32(t_d1, u0);
                    else ptx::tmem_ld16(t_d1, u0);
#else
#pragma unroll
                    for (int i = 0; i < CW; ++i) u0[i] = t_d1 + i;
#endif
                    ptx::tc_wait_ld();
                    ptx::tc_fence_before();
                    ptx::mbar_arrive(&d_empty[buf]);
                    ptx::mbar_arrive(&d_empty[buf1]);
#pragma unroll
                    for (int i = 0; i < CW; ++i)
                        asm("max.f32 %0, %0, %1, %2;" : "+f"(cls[i]) : "f"(__uint_as_float(v0[i])), "f"(__uint_as_float(u0[i])));
                    if (r == 0 && half == 0) TC_TRACE(2, g + 1);
                    a += 2;
                    g += 2;
                } else {
                    ptx::tc_wait_ld();
                    ptx::tc_fence_before();
                    ptx::mbar_arrive(&d_empty[buf]);
#pragma unroll
                    for (int i = 0; i < CW; ++i) cls[i] = fmaxf(cls[i], __uint_as_float(v0[i]));
                    if (r == 0 && half == 0) TC_TRACE(2, g);
                    a += 1;
                    g += 1;
                }
            }
            while (c < p.C) finalize(c++);
            if (r == 0 && half == 0) TC_TRACE(2, 200 + tn);
            if (!norm_ready) warp_wait(&n_full[tn & 1], (tn >> 1) & 1);
            ptx::tc_fence_before();
            ptx::mbar_arrive(&n_empty[tn & 1]);
        }
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 12) ptx::tmem_dealloc<TC_TMEM_COLS>(tmem_base);
}

template <int U>
size_t tc_smem_bytes(int C) {
    using G = Geo<U>;
    constexpr int NOUT = G::BX * G::BY * G::BZ;
    const size_t b = (static_cast<size_t>(G::NSUB) * NOUT * 128 + 1023) & ~static_cast<size_t>(1023);
    return 1024 + b + (2 * TC_STAGES + 2 * TC_NBUF + 4) * 8 + 16 + (C + 1) * 4;
}

template <int U, int EXPK>
int launch_tc(const TcParams& p, cudaStream_t s) {
    const size_t smem = tc_smem_bytes<U>(p.C);
    if (smem > 200 * 1024) return -1;
    auto kern = sim_upsample_tc_kernel<U, EXPK>;
    static PerDeviceMemo configured;
    if (smem > configured.cur()) {
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)) != cudaSuccess) return -1;
        configured.cur() = smem;
    }
    const int grid = p.n_tiles < vittf_num_sms() ? p.n_tiles : vittf_num_sms();
    kern<<<grid, TC_THREADS, smem, s>>>(p);
#ifdef SIM_TRACE
    if (getenv("VITTF_SIM_TRACE_DUMP")) {
        static long long h[3][256];
        cudaDeviceSynchronize();
        cudaMemcpyFromSymbol(h, g_tc_trace, sizeof(h));
        const long long t0 = h[0][200];
        printf("producer tile starts / norm done:");
        for (int i = 200; i < 208; ++i) printf(" %lld", h[0][i] - t0);
        printf("\nepilogue tile ends:");
        for (int i = 200; i < 204; ++i) printf(" %lld", h[2][i] - t0);
        printf("\n proto: producer-arrive  mma-commit  epilogue-done\n");
        for (int i = 0; i < 70; ++i) printf("%3d: %8lld %8lld %8lld\n", i, h[0][i] - t0, h[1][i] - t0, h[2][i] - t0);
    }
#endif
    return 0;
}

template <int U>
int launch_tc_u(const UpParams& q, cudaStream_t s) {
    TcParams p{};
    p.dots = q.dots; p.gram = q.gram; p.class_offsets = q.class_offsets; p.out = q.out;
    p.w = q.w; p.h = q.h; p.d = q.d; p.A = q.A; p.C = q.C;
    p.W = q.W; p.H = q.H; p.D = q.D; p.z0 = q.z0; p.z1 = q.z1;
    p.x0 = q.x0; p.x1 = q.x1;
    p.exponent = q.exponent;
    p.A4 = (q.A + 3) & ~3;
    auto fdiv = [](int a, int b) { return (a >= 0 ? a : a - b + 1) / b; };
    const int zs = q.z1 - q.z0;
    int64_t tiles;
    if (U == 8) {
        const int c_lo = fdiv(q.z0 - 4, 8), c_hi = fdiv(q.z1 - 1 - 4, 8);      // z cells -1 .. d-1 overlapping the slab
        p.zrow0 = c_lo;
        p.nz = c_hi - c_lo + 1;
        p.rows_per_x = (q.h + 1) * p.nz;
        p.tiles_per_x = (p.rows_per_x + 127) / 128;
        tiles = static_cast<int64_t>(q.x1 - q.x0) * p.tiles_per_x;
    } else {
        p.zrow0 = q.z0 / 4;
        p.nz = (q.z1 + 3) / 4 - p.zrow0;
        const int cx_lo = fdiv(q.x0 - U / 2, U), cx_hi = fdiv(q.x1 - 1 - U / 2, U);    // x cells -1 .. w-1 overlapping the slab
        p.xrow0 = cx_lo;
        p.n_rows = static_cast<int64_t>(cx_hi - cx_lo + 1) * (q.h + 1) * p.nz;
        tiles = (p.n_rows + 127) / 128;
    }
    if (tiles <= 0 || tiles > 0x7fffffff) return -1;
    p.n_tiles = static_cast<int>(tiles);
    p.vec_ok = (zs % 4 == 0) && (q.z0 % 4 == 0) && ((reinterpret_cast<uintptr_t>(q.out) & 15) == 0);
    if (q.exponent == 2.0f) return launch_tc<U, 0>(p, s);
    if (q.exponent == 1.0f) return launch_tc<U, 1>(p, s);
    if (q.exponent == 2.5f) return launch_tc<U, 2>(p, s);
    return launch_tc<U, 3>(p, s);
}

}  // namespace

int vittf_launch_upsample_tc(const UpParams& q, int dots_layout, cudaStream_t stream) {
    if (dots_layout != 1) return -1;              // the gather reads voxel-major dots (n_lr, A4)
    if (!q.gram || q.C < 1 || q.C > 4096 || q.A < 1) return -1;
    if (static_cast<int64_t>(q.w) * q.h * q.d * ((q.A + 3) & ~3) >= (1ll << 32)) return -1;
    const bool uniform = q.W % q.w == 0 && static_cast<int64_t>(q.H) * q.w == static_cast<int64_t>(q.h) * q.W &&
                         static_cast<int64_t>(q.D) * q.w == static_cast<int64_t>(q.d) * q.W;
    if (!uniform) return -1;
    if (static_cast<int64_t>(q.w) * q.h * q.d >= (1ll << 31)) return -1;
    const int U = q.W / q.w;
    if (U == 8) return launch_tc_u<8>(q, stream);
    if (U == 4) return launch_tc_u<4>(q, stream);
    if (U == 2 && q.D % 4 == 0) return launch_tc_u<2>(q, stream);
    return -1;
}
