// Fused multi-head attention of the DINO ViT blocks (softmax(q k^T * hd^-0.5) v, head dim 64) that
// /root/reference/infer.py:177 runs through the hub model.  The reference materialises the N x N
// score matrix per head (N = 4097 at 512^2 input); here it never leaves the SM.
//
// B200 design (one CTA per 256 query rows of one (image, head), 384 threads):
//   warp 0    : TMA producer -- Q tiles once, then a 3-stage ring of K (128x64) and V^T (64x128) tiles
//   warp 1, 3 : MMA issuers, one per query tile -- S = Q K^T (tcgen05.mma SS, 128x128x16, fp32 in TMEM), O += P V
//               (tcgen05.mma TS: P is read straight from TMEM, V^T from smem).  One issuing thread for both tiles
//               was itself the critical path (clock64 traces: ~100 cycles per MMA of uniform-datapath bookkeeping).
//   warp 2    : TMEM allocator (S_A, S_B: 2 x 128 cols; P_A, P_B: 2 x 64 cols; O_A, O_B: 2 x 64 cols = all 512).
//               P has its own columns, so a softmax warp hands its S buffer back as soon as the scores are in
//               registers and the MMA warp computes S(j+1) while the exponentials of block j are still running.
//   warps 4-7 : softmax of query tile A, one score row per thread (tcgen05.ld 32x32b), online max with lazy
//   warps 8-11: softmax of query tile B  rescaling of O (only when the row max grows by > 2^8), P written back
//                                         to TMEM as packed bf16.
// What bounds the kernel at head dim 64 is the exponential pipe (16 MUFU/clk/SM vs 4096 MAC/clk/SM), so the
// softmax loop is built around keeping it busy (all measured with clock64 traces on B200):
//   * the two softmax warps that share an SM sub-partition (tile A / tile B) take strict turns on the
//     exponential section through named barriers (ping-pong); left alone they drift into lock-step;
//   * a quarter of the exponentials run as a Cody-Waite + cubic polynomial on the FMA pipe;
//   * everything else is packed / 3-input instructions (FFMA2, FADD2, FMNMX3) with short dependency chains;
//   * the last key block (tokens = 4097 = 32 * 128 + 1) is a 128x16 MMA and the only masked code path; the 32 full
//     blocks carry no masking instructions (the predicated selects used to be 35 % of the loop);
//   * roles are dispatched per warp with elect.sync so that ptxas keeps TMA/MMA descriptors in uniform registers
//     (a branch on threadIdx.x cost ~75 cycles per MMA in divergence loops).
#include <stdlib.h>

#include "common.cuh"

namespace {

constexpr int HD = 64;
constexpr int BQ = 128;
constexpr int BKV = 128;
#ifndef KV_STAGES_V
#define KV_STAGES_V 3
#endif
constexpr int KV_STAGES = KV_STAGES_V;
constexpr int ATT_THREADS = 384;
constexpr int Q_TILE_BYTES = BQ * HD * 2;        // 16 KB
constexpr int K_TILE_BYTES = BKV * HD * 2;       // 16 KB
constexpr int V_HALF_BYTES = HD * 64 * 2;        // 8 KB: 64 d-rows x 64 keys
constexpr int KV_STAGE_BYTES = K_TILE_BYTES + 2 * V_HALF_BYTES;   // 32 KB
constexpr int ATT_SMEM_BYTES = 2 * Q_TILE_BYTES + KV_STAGES * KV_STAGE_BYTES + 1024 + 256;
constexpr int TMEM_COLS = 512;
constexpr int COL_S = 0;     // + t*128
constexpr int COL_P = 256;   // + t*64   packed bf16 pairs
constexpr int COL_O = 384;   // + t*64
constexpr float RESCALE_THRESHOLD = 8.0f;  // log2 units

struct AttnParams {
    __nv_bfloat16* out;
    int tokens, heads, D;
    float scale_log2e;
    int* flags;          // per item: FAST writes 1 when a row left the safe exponent range, SAFE (re)computes flagged items only
    int n_items, units;  // items = (image, head, unit of 2 query tiles), unit fastest
};

#ifdef ATTN_TRACE
__device__ long long g_trace[3 * 40 * 4];
#define TRACE(role, j, k) do { if (blockIdx.x == 1 && blockIdx.y == 0 && blockIdx.z == 0 && (j) < 40) g_trace[((role) * 40 + (j)) * 4 + (k)] = clock64(); } while (0)
#else
#define TRACE(role, j, k) do {} while (0)
#endif
// named barriers (ids 1..8; 0 is __syncthreads): the two softmax warps that share a scheduler take turns on the
// exponential pipe
__device__ __forceinline__ void named_sync(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }
__device__ __forceinline__ void named_arrive(int id, int nthreads) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }
__device__ __forceinline__ float max3(float a, float b, float c) {
    float d;
    asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
    return d;
}
// (x0, x1) = (s0, s1) * c + nm   -- one packed FFMA2
__device__ __forceinline__ void ffma2(float& x0, float& x1, float s0, float s1, float c, float nm) {
    asm("{\n"
        ".reg .b64 ra, rb, rc, rd;\n"
        "mov.b64 ra, {%2, %3};\n"
        "mov.b64 rb, {%4, %4};\n"
        "mov.b64 rc, {%5, %5};\n"
        "fma.rn.f32x2 rd, ra, rb, rc;\n"
        "mov.b64 {%0, %1}, rd;\n"
        "}\n"
        : "=f"(x0), "=f"(x1)
        : "f"(s0), "f"(s1), "f"(c), "f"(nm));
}

// 2^x for a pair of values WITHOUT the MUFU pipe (Cody-Waite range reduction + degree-3 polynomial on the
// FMA pipe, relative error 8e-5 -- invisible after the bf16 rounding of P).  Used for a fixed fraction of the
// exponentials so that the 16/clk/SM MUFU unit and the FMA pipes work side by side.
__device__ __forceinline__ void exp2_poly2(float x0, float x1, float& p0, float& p1) {
    const float magic = 12582912.0f;                       // 1.5 * 2^23: adding it rounds to the nearest integer
    x0 = fmaxf(x0, -126.0f);                               // -inf (masked keys) / deep underflow -> 2^-126 ~ 0
    x1 = fmaxf(x1, -126.0f);
    const ptx::F2 x = ptx::f2_make(x0, x1);
    const ptx::F2 t = ptx::f2_add(x, ptx::f2_make(magic, magic));
    const ptx::F2 n = ptx::f2_add(t, ptx::f2_make(-magic, -magic));
    const ptx::F2 fr = ptx::f2_fma(n, ptx::f2_make(-1.0f, -1.0f), x);            // x - round(x) in [-0.5, 0.5]
    ptx::F2 q = ptx::f2_fma(fr, ptx::f2_make(0.05508868f, 0.05508868f), ptx::f2_make(0.24260405f, 0.24260405f));
    q = ptx::f2_fma(q, fr, ptx::f2_make(0.69327623f, 0.69327623f));
    q = ptx::f2_fma(q, fr, ptx::f2_make(0.99992895f, 0.99992895f));
    float t0, t1, q0, q1;
    ptx::f2_get(t, t0, t1);
    ptx::f2_get(q, q0, q1);
    // add round(x) to the exponent field: the integer sits in the low mantissa bits of t
    p0 = __int_as_float(__float_as_int(q0) + (__float_as_int(t0) << 23));
    p1 = __int_as_float(__float_as_int(q1) + (__float_as_int(t1) << 23));
}
#ifndef POLY_GUARD_V
#define POLY_GUARD_V 1     // max-free pass: 1 = one 3-input |x| maximum per polynomial pair feeds the range flag, 0 = clamp every input
#endif
#ifndef POLY_DEG_V
#define POLY_DEG_V 2      // measured in the step (512^3, ViT-B/8): degree 2 = 4.27 ms per launch, degree 3 = 4.33 ms
#endif
// 2^x for the max-free pass: as above without the four per-pair clamps.  The exponent field only holds round(x) in
// [-126, 127]; instead of clamping, the largest |x| that went through the polynomial is tracked (ONE 3-input maximum per
// pair) and the work item is handed to the safe pass when it exceeds 126 -- exactly like the range check on the row sum.
__device__ __forceinline__ void exp2_poly2_fast(float x0, float x1, float& p0, float& p1, float& track) {
    const float magic = 12582912.0f;
    if (POLY_GUARD_V == 1) {
        asm("max.f32 %0, %0, %1, %2;" : "+f"(track) : "f"(fabsf(x0)), "f"(fabsf(x1)));
    } else {
        x0 = fmaxf(fminf(x0, 126.0f), -126.0f);
        x1 = fmaxf(fminf(x1, 126.0f), -126.0f);
    }
    const ptx::F2 x = ptx::f2_make(x0, x1);
    const ptx::F2 t = ptx::f2_add(x, ptx::f2_make(magic, magic));
    const ptx::F2 n = ptx::f2_add(t, ptx::f2_make(-magic, -magic));
    const ptx::F2 fr = ptx::f2_fma(n, ptx::f2_make(-1.0f, -1.0f), x);
    ptx::F2 q;
    if (POLY_DEG_V == 3) {
        q = ptx::f2_fma(fr, ptx::f2_make(0.05508868f, 0.05508868f), ptx::f2_make(0.24260405f, 0.24260405f));
        q = ptx::f2_fma(q, fr, ptx::f2_make(0.69327623f, 0.69327623f));
        q = ptx::f2_fma(q, fr, ptx::f2_make(0.99992895f, 0.99992895f));
    } else {                                               // relative error 1.7e-3: below the rounding of P to bf16 (2^-9)
        q = ptx::f2_fma(fr, ptx::f2_make(0.2384257f, 0.2384257f), ptx::f2_make(0.7034428f, 0.7034428f));
        q = ptx::f2_fma(q, fr, ptx::f2_make(1.000443f, 1.000443f));
    }
    float t0, t1, q0, q1;
    ptx::f2_get(t, t0, t1);
    ptx::f2_get(q, q0, q1);
    p0 = __int_as_float(__float_as_int(q0) + (__float_as_int(t0) << 23));
    p1 = __int_as_float(__float_as_int(q1) + (__float_as_int(t1) << 23));
}
#ifndef POLY_EVERY_V
#define POLY_EVERY_V 4     // measured on B200: 4 (25 %) > 0 > 2
#endif
#ifndef HANDOVER_CH
#define HANDOVER_CH 3      // hand the pipe over after the last column chunk (measured: 3 >= 0 ~ 1 ~ 2 with persistent CTAs)
#endif
#ifndef POLY_NUM
#define POLY_NUM 1         // POLY_NUM of every POLY_EVERY_V pairs of exponentials run on the FMA pipe (0 = none)
#endif
#ifndef PINGPONG_V
#define PINGPONG_V 1
#endif
constexpr bool PINGPONG = PINGPONG_V != 0;
// pair index -> polynomial (FMA pipe) or MUFU; the polynomial pairs are spread evenly through the row
__device__ __forceinline__ constexpr bool poly_slot(int pair) {
    const int r = pair % POLY_EVERY_V;
    return POLY_NUM > 0 && (r + 1) * POLY_NUM / POLY_EVERY_V != r * POLY_NUM / POLY_EVERY_V;
}

// FAST = true: the max-free first pass (see the header comment); FAST = false: online-softmax with running row max.
// PERSISTENT: gridDim.x CTAs (one per SM) walk the work items item = blockIdx.x, + gridDim.x, ...; an item is the pair of
// 128-row query tiles `unit` of one (image, head).  Barriers, the K/V ring and TMEM live across items, so the loads and the
// first score MMA of the next item overlap the epilogue of the current one, and no CTA is launched per item.
template <bool FAST>
__global__ void __launch_bounds__(ATT_THREADS, 1)
    attention_kernel(const __grid_constant__ CUtensorMap tm_qk, const __grid_constant__ CUtensorMap tm_vt, AttnParams p) {
    extern __shared__ uint8_t smem_raw[];
    // Every role re-derives its pointers and block counts from the kernel parameters behind an opaque zero (ROLE_LOCALS):
    // values computed once up here and used by all roles would stay live across the setmaxnreg boundaries, and ptxas
    // parks them in local memory -- reloads in the MMA issuers' loop are exactly what the persistent form must avoid.
#define ROLE_LOCALS                                                                                                      \
    uint32_t z_ = 0;                                                                                                     \
    asm volatile("" : "+r"(z_));                                                                                         \
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + z_ + 1023) & ~uintptr_t(1023));  \
    uint8_t* s_q = smem;                                                                                                 \
    uint8_t* s_kv = smem + 2 * Q_TILE_BYTES;                                                                             \
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_kv + KV_STAGES * KV_STAGE_BYTES);                                     \
    uint64_t* q_full = bars;                  /* Q tiles of the item have landed */                                      \
    uint64_t* q_empty = bars + 1;             /* both tiles have issued their last score MMA of the item */              \
    uint64_t* kv_full = bars + 2;                                                                                        \
    uint64_t* kv_empty = kv_full + KV_STAGES;                                                                            \
    uint64_t* s_full = kv_empty + KV_STAGES;  /* [2] scores of the next block are in TMEM */                             \
    uint64_t* s_free = s_full + 2;            /* [2] all 128 rows of S are in registers */                               \
    uint64_t* p_ready = s_free + 2;           /* [2] P written */                                                        \
    uint64_t* pv_done = p_ready + 2;          /* [2] P V finished (P buffer reusable, O stable) */                       \
    uint64_t* o_free = pv_done + 2;           /* [2] the epilogue has read O: the next item may overwrite it */          \
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_free + 2);                                                       \
    /* keys: nfull blocks of 128 + one tail block that is only as wide as it has to be (tokens = 4097 = 32 * 128 + 1 */  \
    /* at 512^2 input: the tail costs a 128x16 MMA instead of a 33rd full block) */                                      \
    const int tokens_ = p.tokens + static_cast<int>(z_);                                                                 \
    const int nfull = tokens_ / BKV;                                                                                     \
    const int tail = tokens_ - nfull * BKV;         /* valid keys in the tail block (0 = none) */                        \
    const int tail_n = (tail + 15) & ~15;            /* its MMA N (scores) / K (P V) extent */                            \
    const int nkv = nfull + (tail > 0 ? 1 : 0);                                                                          \
    /* every role walks the same item sequence; the safe pass only visits items the first pass flagged */                \
    auto visit = [&](int item) { return FAST || p.flags == nullptr || p.flags[item] != 0; };                             \
    auto two_tiles = [&](int item) { return (item % p.units) * 2 * BQ + BQ < p.tokens; };                                \
    (void)s_q; (void)q_full; (void)q_empty; (void)kv_full; (void)kv_empty; (void)s_full; (void)s_free; (void)p_ready;     \
    (void)pv_done; (void)o_free; (void)tmem_slot; (void)tail_n; (void)nkv; (void)visit; (void)two_tiles

    if (!FAST && p.flags != nullptr) {
        // second pass: leave at once unless one of this CTA's items was flagged (all threads scan in parallel; walking the
        // items role by role costs a dependent global load per item -- 24 us for nothing in the common case)
        int any = 0;
        for (int item = blockIdx.x + threadIdx.x * gridDim.x; item < p.n_items; item += blockDim.x * gridDim.x) any |= p.flags[item];
        if (!__syncthreads_or(any)) return;
    }
    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);   // provably warp-uniform role index
    const int lane = threadIdx.x & 31;
    {
        ROLE_LOCALS;
    if (threadIdx.x == 0) {   // one-time setup, not on the hot path
        ptx::mbar_init(q_full, 1);
        ptx::mbar_init(q_empty, 2);
        for (int i = 0; i < KV_STAGES; ++i) {
            ptx::mbar_init(&kv_full[i], 1);
            ptx::mbar_init(&kv_empty[i], 2);         // one tcgen05.commit per query tile (tile B's issuer shadows one-tile items)
        }
        for (int i = 0; i < 2; ++i) {
            ptx::mbar_init(&s_full[i], 1);
            ptx::mbar_init(&s_free[i], 128);
            ptx::mbar_init(&p_ready[i], 128);
            ptx::mbar_init(&pv_done[i], 1);
            ptx::mbar_init(&o_free[i], 128);
        }
        ptx::fence_barrier_init();
    }
    if (warp == 2) ptx::tmem_alloc<TMEM_COLS>(tmem_slot);
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();

    // Roles are dispatched per WARP (uniform) and the single issuing lane is chosen with elect.sync: a branch on
    // threadIdx.x would make ptxas wrap every TMA / MMA instruction in a divergence ("waterfall") loop, because
    // their descriptor operands must live in uniform registers.  Only that one lane polls the mbarriers -- a full
    // warp spinning on try_wait measurably starves the softmax warps that share its scheduler.
    // Register re-balancing: the producer / MMA / allocator warpgroup needs few registers, the two softmax warpgroups
    // hold a whole 128-wide score row per thread.
    if (warp < 4) {
      ptx::setmaxnreg_dec<64>();     // the CTA owns 384 x 168 registers: 128 x (168 - 64) released >= 256 x (216 - 168) claimed below
      if (warp == 0) {
        // ---------------- TMA producer (one elected lane) ----------------
        if (ptx::elect_one()) {
            ROLE_LOCALS;
            ptx::prefetch_tmap(&tm_qk);
            ptx::prefetch_tmap(&tm_vt);
            int g = 0, np = 0;               // key blocks / items loaded so far
            for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
                if (!visit(item)) continue;
                const int unit = item % p.units, head = (item / p.units) % p.heads, img = item / (p.units * p.heads);
                const int q0 = unit * 2 * BQ;
                const int n_tiles = q0 + BQ < p.tokens ? 2 : 1;
                ptx::mbar_wait_quiet(q_empty, (np & 1) ^ 1);           // the previous item's score MMAs are done with Q
                ptx::mbar_arrive_expect_tx(q_full, n_tiles * Q_TILE_BYTES);
                for (int t = 0; t < n_tiles; ++t)
                    ptx::tma_load_3d(s_q + t * Q_TILE_BYTES, &tm_qk, q_full, head * HD, q0 + t * BQ, img);
                const int vt_row = (img * p.heads + head) * HD;
                for (int j = 0; j < nkv; ++j, ++g) {
                    const int st = g % KV_STAGES;
                    ptx::mbar_wait_quiet(&kv_empty[st], ((g / KV_STAGES) & 1) ^ 1);
                    uint8_t* dst = s_kv + st * KV_STAGE_BYTES;
                    ptx::mbar_arrive_expect_tx(&kv_full[st], KV_STAGE_BYTES);
                    ptx::tma_load_3d(dst, &tm_qk, &kv_full[st], p.D + head * HD, j * BKV, img);
                    ptx::tma_load_2d(dst + K_TILE_BYTES, &tm_vt, &kv_full[st], j * BKV, vt_row);
                    ptx::tma_load_2d(dst + K_TILE_BYTES + V_HALF_BYTES, &tm_vt, &kv_full[st], j * BKV + 64, vt_row);
                }
                ++np;
            }
        }
      } else if ((warp == 1 || warp == 3) && ptx::elect_one()) {
        // ---------------- MMA issuers: warp 1 drives query tile A, warp 3 tile B ----------------
        // One thread per tile keeps the two tiles decoupled (an in-order issuer blocked on tile B's barrier cannot
        // delay tile A) and its instruction stream short: the clock64 traces showed the single issuing thread, not
        // the tensor pipe, on the critical path (~250 dependent uniform-datapath instructions per key block).
        ROLE_LOCALS;
        const uint32_t tmem_base = *tmem_slot;
        const int t = warp >> 1;
        constexpr uint32_t idesc_s_full = ptx::idesc_bf16_f32(BQ, BKV);
        const uint32_t idesc_s_tail = ptx::idesc_bf16_f32(BQ, tail_n);
        constexpr uint32_t idesc_o = ptx::idesc_bf16_f32(BQ, HD);
        constexpr uint32_t STAGE_DESC = KV_STAGE_BYTES >> 4;    // descriptor address units are 16 B
        constexpr uint32_t V_DESC = K_TILE_BYTES >> 4, VH_DESC = V_HALF_BYTES >> 4;
        const uint32_t d_s = tmem_base + COL_S + t * 128, d_o = tmem_base + COL_O + t * 64, a_p = tmem_base + COL_P + t * 64;
        const uint64_t q_desc = ptx::smem_desc_k_sw128(ptx::smem_u32(s_q) + t * Q_TILE_BYTES);
        const uint64_t kv_desc0 = ptx::smem_desc_k_sw128(ptx::smem_u32(s_kv));
        auto issue_s = [&](uint64_t k_desc, uint32_t idesc) {
#pragma unroll
            for (int k = 0; k < HD / 16; ++k) ptx::umma_ss(d_s, q_desc + 2 * k, k_desc + 2 * k, idesc, k != 0);
        };
        int g = 0, np = 0;       // key blocks / items seen so far (shared barriers: every phase is waited on, in order)
        int gt = 0, nt = 0;      // key blocks / items this tile has processed (its own barriers)
        for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
            if (!visit(item)) continue;
            const bool mine = t == 0 || two_tiles(item);
            ptx::mbar_wait_quiet(q_full, np & 1);
            if (!mine) {
                // one-tile item: tile B's issuer only shadows the shared barriers so that their counts and phases stay uniform
                for (int j = 0; j < nkv; ++j, ++g) {
                    ptx::mbar_wait_quiet(&kv_full[g % KV_STAGES], (g / KV_STAGES) & 1);
                    ptx::tc_commit(&kv_empty[g % KV_STAGES]);
                }
                ptx::tc_commit(q_empty);
                ++np;
                continue;
            }
            ptx::mbar_wait_quiet(&kv_full[g % KV_STAGES], (g / KV_STAGES) & 1);
            if (nt > 0) ptx::mbar_wait_quiet(&s_free[t], (gt - 1) & 1);      // the previous item's last scores were read
            ptx::tc_fence_after();
            issue_s(kv_desc0 + (g % KV_STAGES) * STAGE_DESC, nfull > 0 ? idesc_s_full : idesc_s_tail);
            ptx::tc_commit(&s_full[t]);
            if (nkv == 1) ptx::tc_commit(q_empty);
            for (int j = 0; j < nkv; ++j, ++g) {
                const int st = g % KV_STAGES, st1 = (g + 1) % KV_STAGES;
                if (j + 1 < nkv) {           // S(j+1) as soon as block j's scores sit in registers
                    ptx::mbar_wait_quiet(&kv_full[st1], ((g + 1) / KV_STAGES) & 1);
                    if (t == 0) TRACE(2, j, 0);
                    ptx::mbar_wait_quiet(&s_free[t], (gt + j) & 1);
                    ptx::tc_fence_after();
                    issue_s(kv_desc0 + st1 * STAGE_DESC, j + 1 < nfull ? idesc_s_full : idesc_s_tail);
                    ptx::tc_commit(&s_full[t]);
                    if (j + 2 == nkv) ptx::tc_commit(q_empty);         // last score MMA of the item: Q may be replaced
                    if (t == 0) TRACE(2, j, 1);
                }
                ptx::mbar_wait_quiet(&p_ready[t], (gt + j) & 1);
                if (j == 0 && nt > 0) ptx::mbar_wait_quiet(&o_free[t], (nt - 1) & 1);   // the previous item's O has been read out
                if (t == 0) TRACE(2, j, 2);
                ptx::tc_fence_after();
                const uint64_t v_desc = kv_desc0 + st * STAGE_DESC + V_DESC;
                if (j < nfull) {
                    // P: packed bf16 pairs, 8 TMEM columns per 16 keys; V^T: two 64-key halves
#pragma unroll
                    for (int ks = 0; ks < BKV / 16; ++ks)
                        ptx::umma_ts(d_o, a_p + ks * 8, v_desc + (ks >> 2) * VH_DESC + 2 * (ks & 3), idesc_o, (j | ks) != 0);
                } else {
                    for (int ks = 0; ks < tail_n / 16; ++ks)
                        ptx::umma_ts(d_o, a_p + ks * 8, v_desc + (ks >> 2) * VH_DESC + 2 * (ks & 3), idesc_o, (j | ks) != 0);
                }
                ptx::tc_commit(&pv_done[t]);
                ptx::tc_commit(&kv_empty[st]);     // this tile is done with the stage (the barrier counts both tiles)
                if (t == 0) TRACE(2, j, 3);
            }
            gt += nkv;
            ++nt;
            ++np;
        }
      }
    } else {
        ptx::setmaxnreg_inc<216>();
        // ---------------- softmax + output ----------------
        ROLE_LOCALS;
        const uint32_t tmem_base = *tmem_slot;
        int last2 = -1;                                  // last two-tile item of this CTA (ends the ping-pong token chain)
        for (int item = blockIdx.x; item < p.n_items; item += gridDim.x)
            if (visit(item) && two_tiles(item)) last2 = item;
        const int t = (warp - 4) >> 2;
        const int q = warp & 3;
        const uint32_t lane_base = static_cast<uint32_t>(q * 32) << 16;
        const uint32_t t_s = tmem_base + lane_base + COL_S + t * 128;
        const uint32_t t_p = tmem_base + lane_base + COL_P + t * 64;
        const uint32_t t_o = tmem_base + lane_base + COL_O + t * 64;
        // Ping-pong on the exponential (MUFU) pipe: warps 4+q (tile A) and 8+q (tile B) share an SM sub-partition.
        // Left alone they run their exponentials at the same time (pipe oversubscribed) and their TMEM loads /
        // row maxima at the same time (pipe idle); strict alternation keeps the pipe busy.  The turn token is handed
        // back and forth across items; tile B's very last block keeps it (nobody follows).
        const int bar_mine = 1 + q * 2 + t, bar_other = 1 + q * 2 + (t ^ 1);
        if (PINGPONG && last2 >= 0 && t == 1) named_arrive(bar_other, 64);   // tile A goes first
        int gt = 0;              // key blocks this tile has processed
        for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
            if (!visit(item)) continue;
            // only what the block loop needs stays live through it (the loop body sits at the register budget); the item's
            // coordinates are recomputed for the epilogue
            const bool has_b = two_tiles(item);
            const bool is_last2 = item == last2;
            if (t == 1 && !has_b) continue;
            float l = 0.0f;
            if constexpr (FAST) {
                // ---- max-free pass: q arrives pre-scaled, so a score IS the base-2 exponent and p = 2^s needs no running
                // maximum as long as the exponents stay inside what bf16 P and fp32 O / l can hold; rows that leave that
                // range are detected on l afterwards and their item is redone by the safe kernel.  Per score: MUFU (or the
                // FMA-pipe polynomial), half an FADD2 and half an F2FP -- no scale FMA, no max, no rescale logic.
                float l_f = 0.0f, track = 0.0f;
                for (int j = 0; j < nfull; ++j) {
                    if (q == 0 && lane == 0) TRACE(t, j, 0);
                    ptx::mbar_wait_quiet(&s_full[t], (gt + j) & 1);
                    if (q == 0 && lane == 0) TRACE(t, j, 1);
                    ptx::tc_fence_after();
                    uint32_t s[4][32];
#pragma unroll
                    for (int ch = 0; ch < 4; ++ch) ptx::tmem_ld32(t_s + ch * 32, s[ch]);
                    ptx::tc_wait_ld();
                    ptx::tc_fence_before();
                    ptx::mbar_arrive(&s_free[t]);          // the MMA warp may overwrite S with the next block's scores
                    if (j > 0) ptx::mbar_wait_quiet(&pv_done[t], (gt + j - 1) & 1);   // P buffer free again (normally long since)
                    if (PINGPONG && has_b) named_sync(bar_mine, 64);
                    if (q == 0 && lane == 0) TRACE(t, j, 2);
                    ptx::F2 sums[4] = {{0ull}, {0ull}, {0ull}, {0ull}};
#pragma unroll
                    for (int ch = 0; ch < 4; ++ch) {
                        uint32_t pk[16];
#pragma unroll
                        for (int i = 0; i < 16; ++i) {
                            const float x0 = __uint_as_float(s[ch][2 * i]), x1 = __uint_as_float(s[ch][2 * i + 1]);
                            float p0, p1;
                            if (poly_slot(ch * 16 + i)) {
                                exp2_poly2_fast(x0, x1, p0, p1, track);
                            } else {
                                p0 = ptx::ex2_approx(x0);
                                p1 = ptx::ex2_approx(x1);
                            }
                            sums[i & 3] = ptx::f2_add(sums[i & 3], ptx::f2_make(p0, p1));
                            pk[i] = ptx::pack_bf16x2(p0, p1);
                        }
                        ptx::tmem_st16(t_p + ch * 16, pk);
                        if (PINGPONG && ch == HANDOVER_CH && has_b && !(t == 1 && j == nkv - 1 && is_last2)) named_arrive(bar_other, 64);
                    }
                    float a0, a1, b0, b1;
                    ptx::f2_get(ptx::f2_add(sums[0], sums[1]), a0, a1);
                    ptx::f2_get(ptx::f2_add(sums[2], sums[3]), b0, b1);
                    l_f += (a0 + a1) + (b0 + b1);
                    ptx::tc_wait_st();
                    if (q == 0 && lane == 0) TRACE(t, j, 3);
                    ptx::tc_fence_before();
                    ptx::mbar_arrive(&p_ready[t]);
                }
                if (tail > 0) {
                    const int j = nfull;
                    ptx::mbar_wait_quiet(&s_full[t], (gt + j) & 1);
                    ptx::tc_fence_after();
                    if (j > 0) ptx::mbar_wait_quiet(&pv_done[t], (gt + j - 1) & 1);
                    if (PINGPONG && has_b) named_sync(bar_mine, 64);
                    float sum = 0.0f;
                    for (int c0 = 0; c0 < tail_n; c0 += 16) {
                        uint32_t v[16], pk[8];
                        ptx::tmem_ld16(t_s + c0, v);
                        ptx::tc_wait_ld();
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const float p0 = c0 + 2 * i < tail ? ptx::ex2_approx(__uint_as_float(v[2 * i])) : 0.0f;
                            const float p1 = c0 + 2 * i + 1 < tail ? ptx::ex2_approx(__uint_as_float(v[2 * i + 1])) : 0.0f;
                            sum += p0 + p1;
                            pk[i] = ptx::pack_bf16x2(p0, p1);
                        }
                        ptx::tmem_st8(t_p + (c0 >> 1), pk);
                    }
                    l_f += sum;
                    ptx::tc_fence_before();
                    ptx::mbar_arrive(&s_free[t]);          // keeps the barrier's phase count uniform (one per key block)
                    if (PINGPONG && has_b && (t == 0 || !is_last2)) named_arrive(bar_other, 64);
                    ptx::tc_wait_st();
                    ptx::tc_fence_before();
                    ptx::mbar_arrive(&p_ready[t]);
                }
                // 2^-100 < l < 2^100 bounds every p (and p * v) far away from fp32 overflow and from total underflow;
                // anything else (inf, NaN, 0) sends the item to the safe pass.  Padding rows beyond `tokens` are ignored.
                const bool row_live = (item % p.units) * 2 * BQ + t * BQ + q * 32 + lane < p.tokens;
                if (row_live && !(l_f > 7.888609052210118e-31f && l_f < 1.2676506002282294e30f && track <= 126.0f)) p.flags[item] = 1;
                l = l_f;
            } else {
            const float c = p.scale_log2e;
            float m_used = -INFINITY;
            // lazy running-max update: O and l are rescaled only when the row max grew by more than 2^8
            auto raise_max = [&](float m_blk, int j) {
                const bool grow = m_blk > m_used + RESCALE_THRESHOLD;
                if (__any_sync(0xffffffffu, grow)) {
                    const float m_new = grow ? m_blk : m_used;
                    const float f = ptx::ex2_approx(m_used - m_new);  // 0 on the first block, 1 if unchanged
                    if (j > 0) {
                        ptx::mbar_wait_quiet(&pv_done[t], (gt + j - 1) & 1);     // P V of block j-1 has landed in O
                        ptx::tc_fence_after();
                        uint32_t o[32];
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            ptx::tmem_ld32(t_o + h * 32, o);
                            ptx::tc_wait_ld();
#pragma unroll
                            for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * f);
                            ptx::tmem_st32(t_o + h * 32, o);
                        }
                    }
                    l *= f;
                    m_used = m_new;
                }
            };
            // ---- full 128-key blocks: no masking anywhere on this path ----
            for (int j = 0; j < nfull; ++j) {
                if (q == 0 && lane == 0) TRACE(t, j, 0);
                ptx::mbar_wait_quiet(&s_full[t], (gt + j) & 1);
                if (q == 0 && lane == 0) TRACE(t, j, 1);
                ptx::tc_fence_after();
                uint32_t s[4][32];
#pragma unroll
                for (int ch = 0; ch < 4; ++ch) ptx::tmem_ld32(t_s + ch * 32, s[ch]);
                ptx::tc_wait_ld();
                ptx::tc_fence_before();
                ptx::mbar_arrive(&s_free[t]);          // the MMA warp may overwrite S with the next block's scores
                // row max: 8 independent chains of 3-input max
                float mx[8];
#pragma unroll
                for (int a = 0; a < 8; ++a) mx[a] = __uint_as_float(s[a >> 1][(a & 1) * 16]);
#pragma unroll
                for (int a = 0; a < 8; ++a) {
                    const uint32_t* sv = &s[a >> 1][(a & 1) * 16];
#pragma unroll
                    for (int i = 1; i < 15; i += 2) mx[a] = max3(mx[a], __uint_as_float(sv[i]), __uint_as_float(sv[i + 1]));
                    mx[a] = fmaxf(mx[a], __uint_as_float(sv[15]));
                }
                raise_max(max3(max3(mx[0], mx[1], mx[2]), max3(mx[3], mx[4], mx[5]), fmaxf(mx[6], mx[7])) * c, j);
                if (j > 0) ptx::mbar_wait_quiet(&pv_done[t], (gt + j - 1) & 1);   // P buffer free again (normally long since)
                if (PINGPONG && has_b) named_sync(bar_mine, 64);       // my turn on the exponential pipe
                if (q == 0 && lane == 0) TRACE(t, j, 2);
                const float nm = -m_used;
                ptx::F2 sums[4] = {{0ull}, {0ull}, {0ull}, {0ull}};    // 4 independent packed (2 x fp32) row-sum chains
#pragma unroll
                for (int ch = 0; ch < 4; ++ch) {
                    uint32_t pk[16];
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        float x0, x1;
                        ffma2(x0, x1, __uint_as_float(s[ch][2 * i]), __uint_as_float(s[ch][2 * i + 1]), c, nm);
                        float p0, p1;
                        if (poly_slot(ch * 16 + i)) {
                            exp2_poly2(x0, x1, p0, p1);
                        } else {
                            p0 = ptx::ex2_approx(x0);
                            p1 = ptx::ex2_approx(x1);
                        }
                        sums[i & 3] = ptx::f2_add(sums[i & 3], ptx::f2_make(p0, p1));
                        pk[i] = ptx::pack_bf16x2(p0, p1);
                    }
                    ptx::tmem_st16(t_p + ch * 16, pk);
                    // hand the exponential pipe to the other tile's warp a little before my last chunk drains
                    if (PINGPONG && ch == HANDOVER_CH && has_b && !(t == 1 && j == nkv - 1 && is_last2)) named_arrive(bar_other, 64);
                }
                {
                    float a0, a1, b0, b1;
                    ptx::f2_get(ptx::f2_add(sums[0], sums[1]), a0, a1);
                    ptx::f2_get(ptx::f2_add(sums[2], sums[3]), b0, b1);
                    l += (a0 + a1) + (b0 + b1);
                }
                ptx::tc_wait_st();
                if (q == 0 && lane == 0) TRACE(t, j, 3);
                ptx::tc_fence_before();
                ptx::mbar_arrive(&p_ready[t]);
            }
            // ---- tail block: tail (< 128) valid keys in tail_n = ceil16(tail) score columns; the only masked path ----
            if (tail > 0) {
                const int j = nfull;
                ptx::mbar_wait_quiet(&s_full[t], (gt + j) & 1);
                ptx::tc_fence_after();
                float mx = -INFINITY;
                for (int c0 = 0; c0 < tail_n; c0 += 16) {
                    uint32_t v[16];
                    ptx::tmem_ld16(t_s + c0, v);
                    ptx::tc_wait_ld();
#pragma unroll
                    for (int i = 0; i < 16; ++i)
                        if (c0 + i < tail) mx = fmaxf(mx, __uint_as_float(v[i]));
                }
                raise_max(mx * c, j);
                if (j > 0) ptx::mbar_wait_quiet(&pv_done[t], (gt + j - 1) & 1);
                if (PINGPONG && has_b) named_sync(bar_mine, 64);
                const float nm = -m_used;
                float sum = 0.0f;
                for (int c0 = 0; c0 < tail_n; c0 += 16) {
                    uint32_t v[16], pk[8];
                    ptx::tmem_ld16(t_s + c0, v);
                    ptx::tc_wait_ld();
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float p0 = c0 + 2 * i < tail ? ptx::ex2_approx(fmaf(__uint_as_float(v[2 * i]), c, nm)) : 0.0f;
                        const float p1 = c0 + 2 * i + 1 < tail ? ptx::ex2_approx(fmaf(__uint_as_float(v[2 * i + 1]), c, nm)) : 0.0f;
                        sum += p0 + p1;
                        pk[i] = ptx::pack_bf16x2(p0, p1);
                    }
                    ptx::tmem_st8(t_p + (c0 >> 1), pk);
                }
                l += sum;
                ptx::tc_fence_before();
                ptx::mbar_arrive(&s_free[t]);          // keeps the barrier's phase count uniform (one per key block)
                if (PINGPONG && has_b && (t == 0 || !is_last2)) named_arrive(bar_other, 64);
                ptx::tc_wait_st();
                ptx::tc_fence_before();
                ptx::mbar_arrive(&p_ready[t]);
            }
            }
            // ---- output: O / l -> bf16, token-major (B*tokens, D) at column head*64 ----
            ptx::mbar_wait_quiet(&pv_done[t], (gt + nkv - 1) & 1);   // the last P V has landed in O
            ptx::tc_fence_after();
            uint32_t o[2][32];
            ptx::tmem_ld32(t_o, o[0]);
            ptx::tmem_ld32(t_o + 32, o[1]);
            ptx::tc_wait_ld();
            ptx::tc_fence_before();
            ptx::mbar_arrive(&o_free[t]);                      // the next item's first P V may overwrite O
            gt += nkv;
            const int unit = item % p.units, head = (item / p.units) % p.heads, img = item / (p.units * p.heads);
            const int row = unit * 2 * BQ + t * BQ + q * 32 + lane;
            const float inv = 1.0f / l;
            if (row < p.tokens) {
                __nv_bfloat16* dst = p.out + (static_cast<size_t>(img) * p.tokens + row) * p.D + head * HD;
#pragma unroll
                for (int h = 0; h < 2; ++h)
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        uint4 pkt;
                        pkt.x = ptx::pack_bf16x2(__uint_as_float(o[h][8 * i + 0]) * inv, __uint_as_float(o[h][8 * i + 1]) * inv);
                        pkt.y = ptx::pack_bf16x2(__uint_as_float(o[h][8 * i + 2]) * inv, __uint_as_float(o[h][8 * i + 3]) * inv);
                        pkt.z = ptx::pack_bf16x2(__uint_as_float(o[h][8 * i + 4]) * inv, __uint_as_float(o[h][8 * i + 5]) * inv);
                        pkt.w = ptx::pack_bf16x2(__uint_as_float(o[h][8 * i + 6]) * inv, __uint_as_float(o[h][8 * i + 7]) * inv);
                        reinterpret_cast<uint4*>(dst + h * 32)[i] = pkt;
                    }
            }
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        ROLE_LOCALS;
        ptx::tmem_dealloc<TMEM_COLS>(*tmem_slot);
    }
#undef ROLE_LOCALS
}

}  // namespace

namespace {
int attention_launch(const void* qk, const void* vt, void* out, int B, int tokens, int tok_pad, int heads, float scale,
                     int* flags, bool fast, cudaStream_t stream) {
    const int D = heads * HD;
    CUtensorMap tm_qk, tm_vt;
    {
        uint64_t dims[3] = {static_cast<uint64_t>(2 * D), static_cast<uint64_t>(tokens), static_cast<uint64_t>(B)};
        uint64_t strides[2] = {static_cast<uint64_t>(2 * D) * 2, static_cast<uint64_t>(2 * D) * 2 * tokens};
        uint32_t box[3] = {HD, BQ, 1};
        VITTF_CHECK(vittf_make_tmap(&tm_qk, qk, 2, 3, dims, strides, box, true));
    }
    {
        uint64_t dims[2] = {static_cast<uint64_t>(tok_pad), static_cast<uint64_t>(B) * heads * HD};
        uint64_t strides[1] = {static_cast<uint64_t>(tok_pad) * 2};
        uint32_t box[2] = {64, HD};
        VITTF_CHECK(vittf_make_tmap(&tm_vt, vt, 2, 2, dims, strides, box, true));
    }
    static PerDeviceMemo configured;
    if (!configured.cur()) {
        VITTF_CHECK_CUDA(cudaFuncSetAttribute(attention_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM_BYTES));
        VITTF_CHECK_CUDA(cudaFuncSetAttribute(attention_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM_BYTES));
        configured.cur() = 1;
    }
    const int units = ceil_div(tokens, 2 * BQ);
    AttnParams p{static_cast<__nv_bfloat16*>(out), tokens, heads, D, scale, flags, units * heads * B, units};
    const int grid = p.n_items < vittf_num_sms() ? p.n_items : vittf_num_sms();
    if (fast) attention_kernel<true><<<grid, ATT_THREADS, ATT_SMEM_BYTES, stream>>>(tm_qk, tm_vt, p);
    else attention_kernel<false><<<grid, ATT_THREADS, ATT_SMEM_BYTES, stream>>>(tm_qk, tm_vt, p);
    VITTF_CHECK_CUDA(cudaGetLastError());
    vittf_count_launches(1);
    return VITTF_OK;
}
}  // namespace

static void attn_trace_dump() {
#ifdef ATTN_TRACE
    if (getenv("VITTF_ATTN_TRACE_DUMP")) {
        static long long h[3 * 40 * 4];
        cudaDeviceSynchronize();
        cudaMemcpyFromSymbol(h, g_trace, sizeof(h));
        long long t0 = h[0];
        for (int r = 0; r < 3; ++r)
            for (int j = 2; j < 14; ++j) {
                printf("role %d blk %2d:", r, j);
                for (int k = 0; k < 4; ++k) printf(" %7lld", h[(r * 40 + j) * 4 + k] - t0);
                printf("\n");
            }
    }
#endif
}

extern "C" int64_t vittf_attention_workspace_bytes(int B, int tokens, int heads) {
    if (B <= 0 || tokens <= 0 || heads <= 0) return -1;
    return static_cast<int64_t>(ceil_div(tokens, 2 * BQ)) * heads * B * sizeof(int);
}

// q pre-scaled by hd^-0.5 * log2(e) (the engine folds it into the qkv weights): max-free first pass + safe pass over the
// CTAs it flagged (normally none: the second launch exits after one flag read per CTA).
extern "C" int vittf_attention_prescaled(const void* qk, const void* vt, void* out, int B, int tokens, int tok_pad, int heads,
                                         void* workspace, int64_t workspace_bytes, void* stream) {
    VITTF_REQUIRE(qk && vt && out && workspace, "vittf_attention_prescaled: null pointer");
    VITTF_REQUIRE(B > 0 && tokens > 0 && heads > 0, "vittf_attention_prescaled: empty problem");
    VITTF_REQUIRE(tok_pad >= tokens && tok_pad % 8 == 0, "vittf_attention_prescaled: tok_pad=%d must be >= tokens and a multiple of 8",
                  tok_pad);
    const int64_t need = vittf_attention_workspace_bytes(B, tokens, heads);
    VITTF_REQUIRE(workspace_bytes >= need, "vittf_attention_prescaled: workspace of %lld B is smaller than the %lld B required",
                  (long long)workspace_bytes, (long long)need);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    int* flags = static_cast<int*>(workspace);
    static const bool safe_only = getenv("VITTF_ATTN_SAFE_ONLY") != nullptr;      // A/B switch
    if (safe_only) return attention_launch(qk, vt, out, B, tokens, tok_pad, heads, 1.0f, nullptr, false, s);
    VITTF_CHECK_CUDA(cudaMemsetAsync(flags, 0, need, s));
    VITTF_CHECK(attention_launch(qk, vt, out, B, tokens, tok_pad, heads, 1.0f, flags, true, s));
    attn_trace_dump();
    return attention_launch(qk, vt, out, B, tokens, tok_pad, heads, 1.0f, flags, false, s);
}

extern "C" int vittf_attention(const void* qk, const void* vt, void* out, int B, int tokens, int tok_pad, int heads,
                               void* stream) {
    VITTF_REQUIRE(qk && vt && out, "vittf_attention: null pointer");
    VITTF_REQUIRE(B > 0 && tokens > 0 && heads > 0, "vittf_attention: empty problem");
    VITTF_REQUIRE(tok_pad >= tokens && tok_pad % 8 == 0, "vittf_attention: tok_pad=%d must be >= tokens and a multiple of 8",
                  tok_pad);
    VITTF_CHECK(attention_launch(qk, vt, out, B, tokens, tok_pad, heads, 0.125f * 1.4426950408889634f, nullptr, false,
                                 static_cast<cudaStream_t>(stream)));
    attn_trace_dump();
    return VITTF_OK;
}
