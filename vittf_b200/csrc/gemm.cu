// Linear layers of the ViT (attn.qkv / attn.proj / mlp.fc1 / mlp.fc2 of the hub DINO model
// that /root/reference/infer.py:177 runs): C = A W^T + bias with fused epilogues.
//
// B200 design: persistent, warp-specialised kernel, one CTA per SM -- or one CTA PAIR per TPC (CG = 2, below).
//   warp 0    : TMA producer  (cp.async.bulk.tensor, 128B-swizzled 128x64 / BNx64 bf16 tiles, STAGES-deep ring)
//   warp 1    : MMA issuer    (one thread, tcgen05.mma kind::f16, fp32 accum in TMEM: cta_group::1 128 x BN x 16, or, in
//                              the leader CTA of a pair, cta_group::2 256 x BN x 16)
//   warp 2    : TMEM allocator
//   warps 4-11: epilogue      (tcgen05.ld 32x32b, one accumulator row per thread; two warps per TMEM lane
//                              quarter, each owning half of the tile's columns).  Results are staged in a
//                              128B-swizzled shared-memory tile per warp and leave the SM as TMA tile stores;
//                              the residual update x += A W^T + b is a TMA reduce-add (the read-modify-write
//                              happens in L2, the SM never reads the fp32 residual stream).
// The accumulator is double-buffered in TMEM (2 x BN columns) so the epilogue of tile i overlaps
// the main loop of tile i+1 -- essential here because K is short (384..3072).
//
// LayerNorm without a LayerNorm pass (the pre-LN blocks compute LN(x) W^T + b twice per block).  With gamma folded into
// the weight (W' = W * gamma, host side) and mean / rstd of the row known,
//     LN(x) W^T + b  =  rstd * (x W'^T  -  mean * colsum(W'))  +  (b + W beta),
// so the CONSUMER GEMMs (qkv, fc1, K features) take the raw bf16 copy of the residual stream as their A operand and
// apply the two row scalars in the epilogue (2 FMAs per element instead of 1 add).  The PRODUCER GEMMs (proj, fc2:
// VITTF_EPI_BIAS_RESID_LN) own the residual stream: their epilogue reads x, adds the accumulator, writes x back, writes
// the bf16 copy (TMA store) and the per-row partial sums (sum, sum of squares) of its 128-column slice -- the consumer adds
// the 2 N / BN partials of a row.  An epilogue thread owns one accumulator ROW, so the fp32 stream is kept in a row-tiled
// layout xt[M/32][D/4][32 rows][4 columns]: a warp's 16-byte accesses are 512 contiguous bytes, nothing is staged through
// shared memory, and the row sums need no shuffles.  Against the reduce-add epilogue + layernorm_kernel this removes one
// read of x (fp32) per LayerNorm and the launch itself.  Accuracy: the A operand is bf16(x) instead of bf16(LN(x)), so the
// rounding error of an output grows with |row mean| / row std (the mean term is subtracted AFTER the product): equal to the
// unfused path for centred rows, 2^-9 * mean / std relative otherwise; the variance E[x^2] - mean^2 is formed from fp32 sums
// (relative error ~1e-6 * mean^2 / var).  ViT residual streams have |mean| < std (massive activations sit in single channels
// and raise the std, not the mean); tests/test_gpu_gemm.py::test_gemm_ln_consumer_gelu bounds the error against the unfused
// computation on rows with |mean| / std up to 3.
#include <stdlib.h>

#include "common.cuh"

namespace {

constexpr int BM = 128;
constexpr int BK = 64;  // 64 bf16 = 128 B = one swizzle row
// EW epilogue warps: 8 (two per TMEM lane quarter, half of the tile's columns each) or 16 (four per quarter, a quarter of
// the columns each).  With K = 768 (qkv, fc1 of ViT-B) a 128 x 256 tile's main loop is ~6000 tensor cycles and 8 warps need
// about as long for its epilogue (dependent chains: TMEM load -> bias / LayerNorm scalars -> GELU -> pack -> staging -> TMA):
// with the epilogue skipped the same launches ran 23 % faster (profiles/r2_gemm_epilogue_bound.log).  16 warps (64 columns
// = ONE 64-column staging tile per thread and tile) are selectable for the bf16 epilogues of the 256-wide pair tiles
// (VITTF_GEMM_EPI16) but lose in the power-capped step; what did pay was shortening the 8-warp epilogue's live ranges
// (one 32-column chunk at a time: no spills, 2063 -> 2028 ms at equal clocks).
constexpr int gemm_threads(int ew) { return 128 + ew * 32; }
constexpr int STAGING_BYTES = 32 * 128;  // one 32-row x 128-byte tile per epilogue warp

// CG = 2: CTA pair (cluster of 2, tcgen05 cta_group::2).  The pair computes a 256 x BN tile: each CTA stages ITS 128 rows of
// A and ITS half of the W rows (BN / 2), the leader's MMA reads both halves, each CTA's tensor memory receives its 128
// accumulator rows.  Per CTA and K block that is 32 KB from L2 instead of 48 KB for the same 128 x 256 x 64 MACs (the
// single-CTA kernel at 1.28 PFLOP/s pulls ~52 B/clk/SM from the L2 slices).  Un-throttled (under ncu) the pairs run as
// fast as the single-CTA tiles; in the power-capped step they are 7 % faster (GEMM time 836 -> 775 ms per volume).
template <int BN, int CG = 1, int EW = 8>
struct GemmCfg {
    static constexpr int B_ROWS = BN / CG;                     // W rows staged per CTA
    static constexpr int STAGES = B_ROWS <= 128 ? (EW == 16 ? 5 : 6) : 4;    // (16 staging tiles take one stage's room)
    static constexpr int ACC_STRIDE = BN <= 128 ? 128 : 256;   // TMEM column offset of the second accumulator
    static constexpr int A_BYTES = BM * BK * 2;
    static constexpr int B_BYTES = B_ROWS * BK * 2;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int TMEM_COLS = BN <= 128 ? 256 : 512;
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + EW * STAGING_BYTES + 1024 + 512;
};

struct GemmParams {
    const float* bias;
    void* out;
    void* out2;
    int M, N, K;
    int tokens, tok_pad, heads;
    // LayerNorm folded into the linear layers around it (see the header comment):
    const float* colsum;      // consumer: (N) sum_k W'[n][k]; nullptr = plain bias epilogue
    const float2* stats;      // consumer: (m_pad, VITTF_LN_SLOTS) partial (sum, sum of squares) of every row of the fp32 source of A
    float inv_k, eps;
    float* xt;                // producer (VITTF_EPI_BIAS_RESID_LN): row-tiled fp32 residual stream, read-modify-write
    float2* stats_out;        // producer: (m_pad, VITTF_LN_SLOTS), slots [0, 2 N / BN) written
    int64_t m_pad;            // M rounded up to a multiple of 128
};

// GELU(x) = 0.5 x (1 + erf(x / sqrt 2)), erf from Abramowitz-Stegun 7.1.28
//   erf(z) = 1 - (1 + a1 z + ... + a6 z^6)^-16,  |err| <= 3e-7  (far below the bf16 output rounding),
// evaluated on PAIRS of values with packed fp32x2 instructions (FFMA2 / FMUL2): ~10 issue slots and one
// MUFU.RCP per element instead of the ~30 instructions of erff(), so the fc1 epilogue hides under the MMAs.
__device__ __forceinline__ void gelu_erf2(float& x0, float& x1) {
    const ptx::F2 x = ptx::f2_make(x0, x1);
    const ptx::F2 ax = ptx::f2_make(fabsf(x0), fabsf(x1));
    const ptx::F2 z = ptx::f2_mul(ax, ptx::f2_make(0.70710678118654752f, 0.70710678118654752f));
    ptx::F2 p = ptx::f2_fma(ptx::f2_make(0.0000430638f, 0.0000430638f), z, ptx::f2_make(0.0002765672f, 0.0002765672f));
    p = ptx::f2_fma(p, z, ptx::f2_make(0.0001520143f, 0.0001520143f));
    p = ptx::f2_fma(p, z, ptx::f2_make(0.0092705272f, 0.0092705272f));
    p = ptx::f2_fma(p, z, ptx::f2_make(0.0422820123f, 0.0422820123f));
    p = ptx::f2_fma(p, z, ptx::f2_make(0.0705230784f, 0.0705230784f));
    p = ptx::f2_fma(p, z, ptx::f2_make(1.0f, 1.0f));
    float p0, p1;
    ptx::f2_get(p, p0, p1);
    ptx::F2 r = ptx::f2_make(ptx::rcp_approx(p0), ptx::rcp_approx(p1));
    r = ptx::f2_mul(r, r);
    r = ptx::f2_mul(r, r);
    r = ptx::f2_mul(r, r);
    r = ptx::f2_mul(r, r);                                                      // p^-16 = 1 - erf
    const ptx::F2 erfv = ptx::f2_fma(r, ptx::f2_make(-1.0f, -1.0f), ptx::f2_make(1.0f, 1.0f));
    const ptx::F2 half = ptx::f2_make(0.5f, 0.5f);
    const ptx::F2 g = ptx::f2_fma(ptx::f2_mul(ax, half), erfv, ptx::f2_mul(x, half));          // 0.5 (x + |x| erf(|x| / sqrt 2))
    ptx::f2_get(g, x0, x1);
}

// Same function through tanh: erf(x / sqrt 2) = tanh(x (c0 + c1 x^2 + c2 x^4)) with coefficients fitted to the exact (erf)
// GELU, max |error| 2.5e-5 over the real line (the textbook two-term tanh form is off by 4.7e-4) -- 300x below the bf16
// rounding of the hidden activations it feeds.  6 packed FMA-pipe instructions + 2 FMNMX + 2 MUFU.TANH per PAIR instead
// of 16 + 2 MUFU.RCP: the fc1 epilogue (32768 elements per 128 x 256 tile) needed ~4000 FMA-pipe cycles per tile against
// ~3000 tensor-pipe cycles of its mainloop.  x^2 is clamped at 64 (tanh is +-1 in fp32 there; the quartic term would
// turn the argument's sign beyond |x| = 11).  -DVITTF_GELU_ERF restores the Abramowitz-Stegun version above.
__device__ __forceinline__ float tanh_approx(float x) {
    float y;
    asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ void gelu_tanh2(float& x0, float& x1) {
    const ptx::F2 x = ptx::f2_make(x0, x1);
    float s0, s1;
    ptx::f2_get(ptx::f2_mul(x, x), s0, s1);
    const ptx::F2 s = ptx::f2_make(fminf(s0, 64.0f), fminf(s1, 64.0f));
    ptx::F2 w = ptx::f2_fma(s, ptx::f2_make(-0.000351516787706616f, -0.000351516787706616f), ptx::f2_make(0.037005646018342414f, 0.037005646018342414f));
    w = ptx::f2_fma(w, s, ptx::f2_make(0.7975078842834491f, 0.7975078842834491f));
    float u0, u1;
    ptx::f2_get(ptx::f2_mul(x, w), u0, u1);
    const ptx::F2 t = ptx::f2_make(tanh_approx(u0), tanh_approx(u1));
    const ptx::F2 hx = ptx::f2_mul(x, ptx::f2_make(0.5f, 0.5f));
    ptx::f2_get(ptx::f2_fma(hx, t, hx), x0, x1);
}

// byte offset of 16-byte chunk `chunk` of row `row` inside a 128B-swizzled 32 x 128 B staging tile
__device__ __forceinline__ uint32_t swz(int row, int chunk) { return row * 128 + ((chunk ^ (row & 7)) << 4); }

template <int BN, int EPI, int CG, int EW>
__global__ void __launch_bounds__(gemm_threads(EW), 1)
    gemm_bf16_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b,
                     const __grid_constant__ CUtensorMap tm_out, GemmParams p) {
    using Cfg = GemmCfg<BN, CG, EW>;
    static_assert(CG == 1 || CG == 2, "single CTA or CTA pair");
    static_assert(EW == 8 || EW == 16, "two or four epilogue warps per TMEM lane quarter");
    static_assert(EW == 8 || (EPI != VITTF_EPI_BIAS_RESID_F32 && EPI != VITTF_EPI_BIAS_RESID_LN && EPI != VITTF_EPI_KFEAT_F16),
                  "the residual-stream epilogues (128-column partial sums) and the K-feature epilogue run on 8 warps");
    constexpr int COLS = BN / (EW / 4);                        // columns of the tile one epilogue warp owns
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* staging = smem + Cfg::STAGES * Cfg::STAGE_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(staging + EW * STAGING_BYTES);
    uint64_t* full = bars;
    uint64_t* empty = bars + Cfg::STAGES;
    uint64_t* tmem_full = bars + 2 * Cfg::STAGES;
    uint64_t* tmem_empty = tmem_full + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);   // provably warp-uniform role index
    const int lane = threadIdx.x & 31;
    // a tile = BM * CG rows x BN columns, owned by one CTA or one CTA pair; `unit` walks the tiles of this CTA (pair)
    const int rank = CG == 2 ? static_cast<int>(ptx::cluster_ctarank()) : 0;
    const int unit0 = blockIdx.x / CG, unit_stride = gridDim.x / CG;
    const int m_tiles = (p.M + BM * CG - 1) / (BM * CG);
    const int n_tiles = p.N / BN;
    const int num_tiles = m_tiles * n_tiles;
    const int num_kb = p.K / BK;

    if (threadIdx.x == 0) {
        for (int i = 0; i < Cfg::STAGES; ++i) {
            ptx::mbar_init(&full[i], 1);
            ptx::mbar_init(&empty[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            ptx::mbar_init(&tmem_full[i], 1);
            ptx::mbar_init(&tmem_empty[i], CG * EW * 32);     // the leader's barrier collects both CTAs' epilogue threads
        }
        ptx::fence_barrier_init();
    }
    if (warp == 2) {
        if constexpr (CG == 2) ptx::tmem_alloc_pair<Cfg::TMEM_COLS>(tmem_slot);
        else ptx::tmem_alloc<Cfg::TMEM_COLS>(tmem_slot);
    }
    ptx::tc_fence_before();
    if constexpr (CG == 2) ptx::cluster_sync_all();      // the peer's barriers are initialised before anything is signalled on them
    else __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // Roles are dispatched per WARP (uniform) and the issuing lane is chosen with elect.sync: a branch on
    // threadIdx.x would make ptxas wrap every TMA / MMA instruction in a divergence ("waterfall") loop.
    // Only that one lane polls the mbarriers.
    // EW = 16: the CTA's register pool is what it was launched with, 640 threads x 96; the producer / issuer / allocator
    // warpgroup gives 40 per thread back (128 x 40 = 5120) and the four epilogue warpgroups take 8 more each (512 x 8 = 4096
    // <= 5120: a claim larger than what was released would block forever).  The setmaxnreg instructions sit INSIDE the role
    // branches: ptxas only allocates against the raised limit in code they dominate.
    if (warp < 4) {
      if constexpr (EW == 16) ptx::setmaxnreg_dec<56>();
      if (warp == 0 && ptx::elect_one()) {
        // ---------------- TMA producer ----------------
        ptx::prefetch_tmap(&tm_a);
        ptx::prefetch_tmap(&tm_b);
        int stage = 0;
        uint32_t phase = 0;
        for (int tile = unit0; tile < num_tiles; tile += unit_stride) {
            const int m_blk = tile / n_tiles, n_blk = tile - m_blk * n_tiles;
            for (int kb = 0; kb < num_kb; ++kb) {
                ptx::mbar_wait(&empty[stage], phase ^ 1);
                uint8_t* sa = smem + stage * Cfg::STAGE_BYTES;
                if constexpr (CG == 2) {
                    // both CTAs' bytes are counted on the leader's barrier (the only one its MMA issuer waits on)
                    if (rank == 0) ptx::mbar_arrive_expect_tx(&full[stage], 2 * Cfg::STAGE_BYTES);
                    ptx::tma_load_2d_pair(sa, &tm_a, &full[stage], kb * BK, (m_blk * 2 + rank) * BM);
                    ptx::tma_load_2d_pair(sa + Cfg::A_BYTES, &tm_b, &full[stage], kb * BK, n_blk * BN + rank * Cfg::B_ROWS);
                } else {
                    ptx::mbar_arrive_expect_tx(&full[stage], Cfg::STAGE_BYTES);
                    ptx::tma_load_2d(sa, &tm_a, &full[stage], kb * BK, m_blk * BM);
                    ptx::tma_load_2d(sa + Cfg::A_BYTES, &tm_b, &full[stage], kb * BK, n_blk * BN);
                }
                if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
            }
        }
      } else if (warp == 1 && rank == 0 && ptx::elect_one()) {
        // ---------------- MMA issuer ----------------
        constexpr uint32_t idesc = ptx::idesc_bf16_f32(BM * CG, BN);
        int stage = 0;
        uint32_t phase = 0;
        int it = 0;
        for (int tile = unit0; tile < num_tiles; tile += unit_stride, ++it) {
            const int as = it & 1;
            if constexpr (CG == 2) ptx::mbar_wait_cluster(&tmem_empty[as], ((it >> 1) & 1) ^ 1);
            else ptx::mbar_wait(&tmem_empty[as], ((it >> 1) & 1) ^ 1);
            ptx::tc_fence_after();
            const uint32_t d_tmem = tmem_base + as * Cfg::ACC_STRIDE;
            for (int kb = 0; kb < num_kb; ++kb) {
                ptx::mbar_wait(&full[stage], phase);
                ptx::tc_fence_after();
                const uint32_t sa = ptx::smem_u32(smem + stage * Cfg::STAGE_BYTES);
                const uint64_t adesc = ptx::smem_desc_k_sw128(sa);
                const uint64_t bdesc = ptx::smem_desc_k_sw128(sa + Cfg::A_BYTES);
                if constexpr (CG == 2) {
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k) ptx::umma_ss_pair(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
                    ptx::tc_commit_pair(&empty[stage]);          // frees the stage in both CTAs
                    if (kb == num_kb - 1) ptx::tc_commit_pair(&tmem_full[as]);
                } else {
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k)  // +32 B per K=16 step inside the swizzle row
                        ptx::umma_ss(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
                    ptx::tc_commit(&empty[stage]);
                    if (kb == num_kb - 1) ptx::tc_commit(&tmem_full[as]);
                }
                if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
            }
        }
      }
    } else {
        if constexpr (EW == 16) ptx::setmaxnreg_inc<104>();
        // ---------------- epilogue ----------------
        const int q = warp & 3;                 // TMEM lane quarter this warp may access
        const int half = (warp - 4) >> 2;       // which half (EW = 8) / quarter (EW = 16) of the tile's columns
        uint8_t* stg = staging + (warp - 4) * STAGING_BYTES;
        constexpr int CHUNKS = COLS / 32;
        if (ptx::elect_one()) ptx::prefetch_tmap(&tm_out);
        int it = 0;
        // accumulator drained: tell the (leader's) MMA issuer
        auto release_acc = [&](int as) {
            ptx::tc_fence_before();
            if constexpr (CG == 2) ptx::mbar_arrive_leader(&tmem_empty[as]);
            else ptx::mbar_arrive(&tmem_empty[as]);
        };
        for (int tile = unit0; tile < num_tiles; tile += unit_stride, ++it) {
            const int n_blk = tile % n_tiles;
            const int m_blk = (tile / n_tiles) * CG + rank;      // this CTA's 128-row block
            const int as = it & 1;
            const int row0 = m_blk * BM + q * 32;
            const int row = row0 + lane;
            const int n0 = n_blk * BN + half * COLS;
            // LayerNorm fold, consumer side: the two row scalars (plain epilogue: rstd = 1, -mean * rstd = 0)
            float rstd = 1.0f, nmr = 0.0f;
            const float* cs = p.bias;
            if (p.colsum != nullptr) {
                // the row's partial sums: 64 contiguous bytes, four independent 16-byte loads (unused slots hold zeros)
                static_assert(VITTF_LN_SLOTS == 8, "four float4 loads per row");
                const float4* st = reinterpret_cast<const float4*>(p.stats) + static_cast<size_t>(row) * 4;
                const float4 t0 = __ldg(st), t1 = __ldg(st + 1), t2 = __ldg(st + 2), t3 = __ldg(st + 3);
                const float sum = ((t0.x + t0.z) + (t1.x + t1.z)) + ((t2.x + t2.z) + (t3.x + t3.z));
                const float sq = ((t0.y + t0.w) + (t1.y + t1.w)) + ((t2.y + t2.w) + (t3.y + t3.w));
                const float mean = sum * p.inv_k;
                rstd = rsqrtf(fmaxf(sq * p.inv_k - mean * mean, 0.0f) + p.eps);
                nmr = -mean * rstd;
                cs = p.colsum;
            }
            // producer side: this thread's row of the tile in the row-tiled stream; the first 32 columns are requested
            // before the accumulator is waited for
            float4* xrow = nullptr;
            float4 xo[8];
            if constexpr (EPI == VITTF_EPI_BIAS_RESID_LN) {
                // a warp's slice of the tile (32 rows x BN/2 columns) is ONE contiguous block of the row-tiled stream
                auto slice = [&](int mb, int nb) {
                    return reinterpret_cast<float4*>(p.xt) + (static_cast<size_t>(mb * 4 + q) * (p.N / 4) + (nb * BN + half * COLS) / 4) * 32;
                };
                xrow = slice(m_blk, n_blk) + lane;
#pragma unroll
                for (int i = 0; i < 8; ++i) xo[i] = __ldcs(xrow + i * 32);
            }
            ptx::mbar_wait(&tmem_full[as], (it >> 1) & 1);
            ptx::tc_fence_after();
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * Cfg::ACC_STRIDE + half * COLS;
            // stage_out(): the 32 x 128 B tile in `stg` is complete -> hand it to the TMA engine
            auto stage_out = [&](int col, bool reduce) {
                ptx::fence_proxy_async();
                __syncwarp();
                if (ptx::elect_one()) {          // deterministic leader: the same lane commits and waits
                    if (reduce) ptx::tma_reduce_add_2d(&tm_out, stg, col, row0);
                    else ptx::tma_store_2d(&tm_out, stg, col, row0);
                    ptx::bulk_commit();
                }
            };
            auto staging_free = [&]() {          // previous TMA of this warp has finished reading `stg`
                if (ptx::elect_one()) ptx::bulk_wait_read<0>();
                __syncwarp();
            };
            auto load_biased = [&](int c, float (&v)[32]) {
                uint32_t acc[32];
                ptx::tmem_ld32(taddr + c * 32, acc);
                ptx::tc_wait_ld();
                const int n = n0 + c * 32;
#pragma unroll
                for (int i = 0; i < 32; i += 4) {
                    const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + n + i));
                    const float4 s4 = __ldg(reinterpret_cast<const float4*>(cs + n + i));
                    v[i + 0] = fmaf(__uint_as_float(acc[i + 0]), rstd, fmaf(nmr, s4.x, b.x));
                    v[i + 1] = fmaf(__uint_as_float(acc[i + 1]), rstd, fmaf(nmr, s4.y, b.y));
                    v[i + 2] = fmaf(__uint_as_float(acc[i + 2]), rstd, fmaf(nmr, s4.z, b.z));
                    v[i + 3] = fmaf(__uint_as_float(acc[i + 3]), rstd, fmaf(nmr, s4.w, b.w));
                }
            };
            auto put_bf16 = [&](int chunk0, const float (&v)[32]) {   // 32 values -> 4 chunks of 8 bf16
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    *reinterpret_cast<uint4*>(stg + swz(lane, chunk0 + i)) =
                        make_uint4(ptx::pack_bf16x2(v[8 * i + 0], v[8 * i + 1]), ptx::pack_bf16x2(v[8 * i + 2], v[8 * i + 3]),
                                   ptx::pack_bf16x2(v[8 * i + 4], v[8 * i + 5]), ptx::pack_bf16x2(v[8 * i + 6], v[8 * i + 7]));
            };

            if constexpr (EPI == VITTF_EPI_BIAS_RESID_F32) {
#pragma unroll
                for (int c = 0; c < CHUNKS; ++c) {
                    float v[32];
                    load_biased(c, v);
                    if (c == CHUNKS - 1) {       // accumulator fully read: give it back to the MMA warp
                        release_acc(as);
                    }
                    staging_free();
#pragma unroll
                    for (int i = 0; i < 8; ++i)
                        *reinterpret_cast<float4*>(stg + swz(lane, i)) = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
                    stage_out(n0 + c * 32, true);
                }
            } else if constexpr (EPI == VITTF_EPI_BIAS_RESID_LN) {
                static_assert(CHUNKS % 2 == 0, "the bf16 copy is staged in 64-column tiles");
                float sum = 0.0f, sq = 0.0f;
#pragma unroll
                for (int c = 0; c < CHUNKS; ++c) {
                    float v[32];
                    load_biased(c, v);
                    if (c == CHUNKS - 1) {
                        release_acc(as);
                    }
                    float4 xn[8];
                    if (c + 1 < CHUNKS) {        // next 32 columns of x: in flight while this chunk is finished
#pragma unroll
                        for (int i = 0; i < 8; ++i) xn[i] = __ldcs(xrow + ((c + 1) * 8 + i) * 32);
                    }
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        v[4 * i + 0] += xo[i].x;
                        v[4 * i + 1] += xo[i].y;
                        v[4 * i + 2] += xo[i].z;
                        v[4 * i + 3] += xo[i].w;
                        xrow[(c * 8 + i) * 32] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
                        sum += (v[4 * i] + v[4 * i + 1]) + (v[4 * i + 2] + v[4 * i + 3]);
                        sq = fmaf(v[4 * i], v[4 * i], sq);
                        sq = fmaf(v[4 * i + 1], v[4 * i + 1], sq);
                        sq = fmaf(v[4 * i + 2], v[4 * i + 2], sq);
                        sq = fmaf(v[4 * i + 3], v[4 * i + 3], sq);
                    }
                    if ((c & 1) == 0) staging_free();
                    put_bf16((c & 1) * 4, v);
                    if (c & 1) stage_out(n0 + (c - 1) * 32, false);
                    if (c + 1 < CHUNKS) {
#pragma unroll
                        for (int i = 0; i < 8; ++i) xo[i] = xn[i];
                    }
                }
                p.stats_out[static_cast<size_t>(row) * VITTF_LN_SLOTS + n_blk * 2 + half] = make_float2(sum, sq);
            } else if constexpr (EPI == VITTF_EPI_BIAS_BF16 || EPI == VITTF_EPI_BIAS_GELU_BF16 || EPI == VITTF_EPI_QKV_SPLIT) {
                static_assert(CHUNKS % 2 == 0, "bf16 epilogues stage 64-column tiles");
                const int two_d = (p.N / 3) * 2;
#pragma unroll
                for (int c = 0; c < CHUNKS; c += 2) {
                    const int n = n0 + c * 32;
                    // V third of the QKV projection: stored transposed per (image, head): vt[(img*heads+h)*64 + d][token];
                    // lanes hold consecutive tokens, so each store instruction writes 64 contiguous bytes
                    const bool vt_third = EPI == VITTF_EPI_QKV_SPLIT && n >= two_d;          // (warp-uniform)
                    // one 32-column chunk at a time (32 values + the TMEM load live): the two chunks share a staging tile
#pragma unroll
                    for (int hh = 0; hh < 2; ++hh) {
                        float v[32];
                        load_biased(c + hh, v);
                        if (c == CHUNKS - 2 && hh == 1) release_acc(as);
                        if (vt_third) {
                            if (row < p.M) {
                                const int img = row / p.tokens;
                                const int tok = row - img * p.tokens;
                                __nv_bfloat16* dst = static_cast<__nv_bfloat16*>(p.out2) +
                                                     (static_cast<size_t>(img) * p.heads * 64 + (n - two_d) + 32 * hh) * p.tok_pad + tok;
#pragma unroll
                                for (int i = 0; i < 32; ++i) dst[static_cast<size_t>(i) * p.tok_pad] = __float2bfloat16_rn(v[i]);
                            }
                            continue;
                        }
                        if constexpr (EPI == VITTF_EPI_BIAS_GELU_BF16) {
#pragma unroll
                            for (int i = 0; i < 32; i += 2) {
#ifdef VITTF_GELU_ERF
                                gelu_erf2(v[i], v[i + 1]);
#else
                                gelu_tanh2(v[i], v[i + 1]);
#endif
                            }
                        }
                        if (hh == 0) staging_free();
                        put_bf16(4 * hh, v);
                    }
                    if (!vt_third) stage_out(n, false);
                }
            } else {  // VITTF_EPI_KFEAT_F16: CLS rows dropped (infer.py:202 `[:, 1:]`), fp16, direct stores
                const int img = row / p.tokens;
                const int tok = row - img * p.tokens;
                const bool live = row < p.M && tok != 0;
                const size_t orow = static_cast<size_t>(img) * (p.tokens - 1) + (tok - 1);
#pragma unroll
                for (int c = 0; c < CHUNKS; ++c) {
                    float v[32];
                    load_biased(c, v);
                    if (c == CHUNKS - 1) {
                        release_acc(as);
                    }
                    if (live) {
                        uint4* dst = reinterpret_cast<uint4*>(static_cast<__half*>(p.out) + orow * p.N + n0 + c * 32);
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            __half2 h0 = __floats2half2_rn(v[8 * i + 0], v[8 * i + 1]);
                            __half2 h1 = __floats2half2_rn(v[8 * i + 2], v[8 * i + 3]);
                            __half2 h2 = __floats2half2_rn(v[8 * i + 4], v[8 * i + 5]);
                            __half2 h3 = __floats2half2_rn(v[8 * i + 6], v[8 * i + 7]);
                            dst[i] = make_uint4(*reinterpret_cast<uint32_t*>(&h0), *reinterpret_cast<uint32_t*>(&h1),
                                                *reinterpret_cast<uint32_t*>(&h2), *reinterpret_cast<uint32_t*>(&h3));
                        }
                    }
                }
            }
        }
        if (ptx::elect_one()) ptx::bulk_wait_all();   // staging tiles must outlive their TMA reads; writes complete
    }
    ptx::tc_fence_before();
    if constexpr (CG == 2) {
        ptx::cluster_sync_all();      // neither CTA leaves (or frees tensor memory) while the pair's MMAs / remote arrives are in flight
        if (warp == 2) ptx::tmem_dealloc_pair<Cfg::TMEM_COLS>(tmem_base);
    } else {
        __syncthreads();
        if (warp == 2) ptx::tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
    }
}

template <int BN, int EPI, int CG = 1, int EW = 8>
int launch_gemm(const void* A, const void* W, const GemmParams& p, cudaStream_t stream) {
    using Cfg = GemmCfg<BN, CG, EW>;
    CUtensorMap tm_a, tm_b, tm_out;
    {
        uint64_t dims[2] = {static_cast<uint64_t>(p.K), static_cast<uint64_t>(p.M)};
        uint64_t strides[1] = {static_cast<uint64_t>(p.K) * 2};
        uint32_t box[2] = {BK, BM};
        VITTF_CHECK(vittf_make_tmap(&tm_a, A, 2, 2, dims, strides, box, true));
    }
    {
        uint64_t dims[2] = {static_cast<uint64_t>(p.K), static_cast<uint64_t>(p.N)};
        uint64_t strides[1] = {static_cast<uint64_t>(p.K) * 2};
        uint32_t box[2] = {BK, static_cast<uint32_t>(Cfg::B_ROWS)};
        VITTF_CHECK(vittf_make_tmap(&tm_b, W, 2, 2, dims, strides, box, true));
    }
    if (EPI == VITTF_EPI_BIAS_RESID_F32) {            // fp32 (M, N), 32 x 32 tiles, reduce-add
        uint64_t dims[2] = {static_cast<uint64_t>(p.N), static_cast<uint64_t>(p.M)};
        uint64_t strides[1] = {static_cast<uint64_t>(p.N) * 4};
        uint32_t box[2] = {32, 32};
        VITTF_CHECK(vittf_make_tmap(&tm_out, p.out, 4, 2, dims, strides, box, true));
    } else if (EPI == VITTF_EPI_KFEAT_F16) {
        tm_out = tm_a;                                // unused by this epilogue
    } else {                                          // bf16 (M, ldc), 64 x 32 tiles
        const uint64_t ldc = EPI == VITTF_EPI_QKV_SPLIT ? static_cast<uint64_t>(p.N / 3) * 2 : static_cast<uint64_t>(p.N);
        uint64_t dims[2] = {ldc, static_cast<uint64_t>(p.M)};
        uint64_t strides[1] = {ldc * 2};
        uint32_t box[2] = {64, 32};
        VITTF_CHECK(vittf_make_tmap(&tm_out, p.out, 2, 2, dims, strides, box, true));
    }
    auto kern = gemm_bf16_kernel<BN, EPI, CG, EW>;
    static PerDeviceMemo configured;
    if (!configured.cur()) {
        VITTF_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
        configured.cur() = 1;
    }
    const int tiles = ceil_div(p.M, BM * CG) * (p.N / BN);          // tiles of a CTA (CG = 1) or of a CTA pair
    const int units = vittf_num_sms() / CG;
    const int grid = (tiles < units ? tiles : units) * CG;
    if (CG == 1) {
        kern<<<grid, gemm_threads(EW), Cfg::SMEM_BYTES, stream>>>(tm_a, tm_b, tm_out, p);
    } else {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(grid);
        cfg.blockDim = dim3(gemm_threads(EW));
        cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
        cfg.stream = stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;          // CTA pairs: the two CTAs of a cluster share a TPC
        attr[0].val.clusterDim.x = CG;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        VITTF_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, tm_a, tm_b, tm_out, p));
    }
    VITTF_CHECK_CUDA(cudaGetLastError());
    vittf_count_launches(1);
    return VITTF_OK;
}

}  // namespace

namespace {
int gemm_dispatch(const void* A, const void* W, GemmParams p, int epi, cudaStream_t s) {
    const int N = p.N;
    static const bool pairs = getenv("VITTF_GEMM_NO_PAIRS") == nullptr;       // A/B switch: single-CTA tiles only
    // A/B switch: 16 epilogue warps for the bf16 epilogues of the pair tiles.  Measured in the step (profiles/
    // r2_ab_step_epilogue_warps.log): 2045 / 2051 ms against 2027 / 2029 ms for 8 warps -- stand-alone the K = 768 GEMMs are
    // epilogue-bound (23 % faster with the epilogue skipped), under the power cap the extra warps cost more than they hide
    static const bool wide_epi = getenv("VITTF_GEMM_EPI16") != nullptr;
    // (256 x 128 pair tiles for N % 256 != 0 were measured and dropped: ViT-S/8 step 462.8 ms with them, 457.8 ms without)
    // widest tile that divides N (a wider tile halves the shared-memory operand traffic per MMA); the bf16
    // epilogues stage 64-column tiles per warp, so they use 256 or 128; the fp32 reduce-add also takes 192
#define VITTF_GEMM_BF16OUT(E)                                                       \
    do {                                                                            \
        if (N % 256 == 0 && pairs && wide_epi && E != VITTF_EPI_BIAS_RESID_LN)      \
            return launch_gemm<256, E == VITTF_EPI_BIAS_RESID_LN ? VITTF_EPI_BIAS_BF16 : E, 2, 16>(A, W, p, s); \
        if (N % 256 == 0 && pairs) return launch_gemm<256, E, 2>(A, W, p, s);      \
        if (N % 256 == 0) return launch_gemm<256, E>(A, W, p, s);                   \
        return launch_gemm<128, E>(A, W, p, s);                                     \
    } while (0)
    switch (epi) {
        case VITTF_EPI_BIAS_BF16: VITTF_GEMM_BF16OUT(VITTF_EPI_BIAS_BF16);
        case VITTF_EPI_BIAS_GELU_BF16: VITTF_GEMM_BF16OUT(VITTF_EPI_BIAS_GELU_BF16);
        case VITTF_EPI_BIAS_RESID_F32:
            if (N % 256 == 0 && pairs) return launch_gemm<256, VITTF_EPI_BIAS_RESID_F32, 2>(A, W, p, s);
            if (N % 256 == 0) return launch_gemm<256, VITTF_EPI_BIAS_RESID_F32>(A, W, p, s);
            if (N % 192 == 0) return launch_gemm<192, VITTF_EPI_BIAS_RESID_F32>(A, W, p, s);
            return launch_gemm<128, VITTF_EPI_BIAS_RESID_F32>(A, W, p, s);
        case VITTF_EPI_BIAS_RESID_LN:
            VITTF_REQUIRE(p.xt && p.stats_out, "vittf_gemm_bf16_ln: the residual-stream epilogue needs xt and stats_out");
            VITTF_REQUIRE(vittf_gemm_ln_slots(N) <= VITTF_LN_SLOTS, "vittf_gemm_bf16_ln: N=%d needs %d partial-sum slots (max %d)", N,
                          vittf_gemm_ln_slots(N), VITTF_LN_SLOTS);
            VITTF_GEMM_BF16OUT(VITTF_EPI_BIAS_RESID_LN);
        case VITTF_EPI_QKV_SPLIT:
            VITTF_REQUIRE(p.out2 && N % 3 == 0 && (N / 3) % 64 == 0 && p.tokens > 0 && p.tok_pad >= p.tokens && p.M % p.tokens == 0,
                          "vittf_gemm_bf16: bad QKV split arguments (N=%d tokens=%d tok_pad=%d M=%d)", N, p.tokens,
                          p.tok_pad, p.M);
            p.heads = N / 3 / 64;
            VITTF_GEMM_BF16OUT(VITTF_EPI_QKV_SPLIT);
        case VITTF_EPI_KFEAT_F16:
            VITTF_REQUIRE(p.tokens > 1 && p.M % p.tokens == 0, "vittf_gemm_bf16: bad K-feature arguments");
            return launch_gemm<128, VITTF_EPI_KFEAT_F16>(A, W, p, s);
        default: VITTF_REQUIRE(false, "vittf_gemm_bf16: unknown epilogue %d", epi);
    }
#undef VITTF_GEMM_BF16OUT
    return VITTF_OK;
}
}  // namespace

extern "C" int vittf_gemm_bf16(const void* A, const void* W, const float* bias, void* out, void* out2, int M, int N,
                               int K, int epi, int tokens, int tok_pad, void* stream) {
    VITTF_REQUIRE(A && W && bias && out, "vittf_gemm_bf16: null pointer");
    VITTF_REQUIRE(M > 0 && N > 0 && K > 0, "vittf_gemm_bf16: empty problem M=%d N=%d K=%d", M, N, K);
    VITTF_REQUIRE(K % BK == 0, "vittf_gemm_bf16: K=%d must be a multiple of %d", K, BK);
    VITTF_REQUIRE(N % 128 == 0, "vittf_gemm_bf16: N=%d must be a multiple of 128", N);
    VITTF_REQUIRE(epi != VITTF_EPI_BIAS_RESID_LN, "vittf_gemm_bf16: epilogue %d needs vittf_gemm_bf16_ln", epi);
    GemmParams p{};
    p.bias = bias; p.out = out; p.out2 = out2; p.M = M; p.N = N; p.K = K; p.tokens = tokens; p.tok_pad = tok_pad;
    return gemm_dispatch(A, W, p, epi, static_cast<cudaStream_t>(stream));
}

extern "C" int vittf_gemm_ln_slots(int N) { return N <= 0 ? -1 : 2 * N / (N % 256 == 0 ? 256 : 128); }

extern "C" int vittf_gemm_bf16_ln(const void* A, const void* W, const float* bias, void* out, void* out2, int M, int N,
                                  int K, int epi, int tokens, int tok_pad, const vittf_ln_fold* ln, void* stream) {
    VITTF_REQUIRE(A && W && bias && out && ln, "vittf_gemm_bf16_ln: null pointer");
    VITTF_REQUIRE(M > 0 && N > 0 && K > 0, "vittf_gemm_bf16_ln: empty problem M=%d N=%d K=%d", M, N, K);
    VITTF_REQUIRE(K % BK == 0, "vittf_gemm_bf16_ln: K=%d must be a multiple of %d", K, BK);
    VITTF_REQUIRE(N % 128 == 0, "vittf_gemm_bf16_ln: N=%d must be a multiple of 128", N);
    VITTF_REQUIRE(ln->m_pad >= M && ln->m_pad % 256 == 0, "vittf_gemm_bf16_ln: m_pad=%lld must be M rounded up to a multiple of 256",
                  (long long)ln->m_pad);
    GemmParams p{};
    p.bias = bias; p.out = out; p.out2 = out2; p.M = M; p.N = N; p.K = K; p.tokens = tokens; p.tok_pad = tok_pad;
    p.m_pad = ln->m_pad;
    if (ln->colsum) {                 // consumer: A is the raw bf16 copy of the stream, W the gamma-folded weight
        VITTF_REQUIRE(ln->stats, "vittf_gemm_bf16_ln: colsum given without row statistics");
        VITTF_REQUIRE(epi != VITTF_EPI_BIAS_RESID_F32, "vittf_gemm_bf16_ln: the reduce-add epilogue takes no LayerNorm fold");
        p.colsum = ln->colsum;
        p.stats = reinterpret_cast<const float2*>(ln->stats);
        p.inv_k = 1.0f / static_cast<float>(K);
        p.eps = ln->eps;
    }
    p.xt = ln->xt;
    p.stats_out = reinterpret_cast<float2*>(ln->stats_out);
    return gemm_dispatch(A, W, p, epi, static_cast<cudaStream_t>(stream));
}
