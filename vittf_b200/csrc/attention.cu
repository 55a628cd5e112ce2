// Fused multi-head attention of the DINO ViT blocks (softmax(q k^T * hd^-0.5) v, head dim 64) that
// /root/reference/infer.py:177 runs through the hub model.  The reference materialises the N x N
// score matrix per head (N = 4097 at 512^2 input); here it never leaves the SM.
//
// B200 design (one CTA per 256 query rows of one (image, head), 384 threads):
//   warp 0    : TMA producer -- Q tiles once, then a 3-stage ring of K (128x64) and V^T (64x128) tiles
//   warp 1    : MMA issuer   -- S = Q K^T (tcgen05.mma SS, 128x128x16, fp32 in TMEM), O += P [V | 1] (tcgen05.mma TS:
//                               P is read straight from TMEM, V^T from smem), for two query tiles ping-pong.
//                               A constant "ones" row appended to every V^T tile makes the tensor core produce the
//                               softmax denominator in accumulator column 64 -- no per-element adds on the CUDA cores.
//   warp 2    : TMEM allocator (S_A, S_B: 2 x 128 cols; O_A, O_B: 2 x 80 cols; P aliases S)
//   warps 4-7 : softmax of query tile A, one score row per thread (tcgen05.ld 32x32b), online max with lazy
//   warps 8-11: softmax of query tile B  rescaling of O (only when the row max grows by > 2^8), P written back
//                                         to TMEM as packed bf16.
// The tensor pipe works on tile B while the CUDA cores do the exponentials of tile A and vice versa.
// The exponentials are the bound at head dim 64 (16 MUFU/clk/SM vs 4096 MAC/clk/SM): everything else in the
// softmax loop is trimmed to packed / 3-input instructions (FFMA2, FMNMX3) with short dependency chains.
#include "common.cuh"

namespace {

constexpr int HD = 64;
constexpr int BQ = 128;
constexpr int BKV = 128;
constexpr int NV = HD + 16;                       // V^T rows + the ones row (+15 zero rows): MMA N = 80
constexpr int KV_STAGES = 3;
constexpr int ATT_THREADS = 384;
constexpr int Q_TILE_BYTES = BQ * HD * 2;        // 16 KB
constexpr int K_TILE_BYTES = BKV * HD * 2;       // 16 KB
constexpr int V_ROWS_BYTES = HD * 64 * 2;        // 8 KB: 64 d-rows x 64 keys
constexpr int ONES_BYTES = 16 * 64 * 2;          // 2 KB: 16 rows x 64 keys
constexpr int V_HALF_BYTES = V_ROWS_BYTES + ONES_BYTES;   // 10 KB, rows 64..79 follow rows 0..63 (8-row groups 1 KB apart)
constexpr int KV_STAGE_BYTES = K_TILE_BYTES + 2 * V_HALF_BYTES;   // 36 KB
constexpr int ATT_SMEM_BYTES = 2 * Q_TILE_BYTES + KV_STAGES * KV_STAGE_BYTES + 1024 + 256;
constexpr int TMEM_COLS = 512;
constexpr int COL_S = 0;     // + t*128
constexpr int COL_O = 256;   // + t*128 (80 columns used: 64 outputs, column 64 = row sum)
constexpr float RESCALE_THRESHOLD = 8.0f;  // log2 units

struct AttnParams {
    __nv_bfloat16* out;
    int tokens, heads, D;
    float scale_log2e;
};

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
constexpr int POLY_EVERY = 0;   // every POLY_EVERY-th pair of exponentials runs on the FMA pipe (0 = none)

__global__ void __launch_bounds__(ATT_THREADS, 1)
    attention_kernel(const __grid_constant__ CUtensorMap tm_qk, const __grid_constant__ CUtensorMap tm_vt, AttnParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* s_q = smem;
    uint8_t* s_kv = smem + 2 * Q_TILE_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_kv + KV_STAGES * KV_STAGE_BYTES);
    uint64_t* q_full = bars;
    uint64_t* kv_full = bars + 1;
    uint64_t* kv_empty = kv_full + KV_STAGES;
    uint64_t* s_full = kv_empty + KV_STAGES;  // [2]
    uint64_t* p_ready = s_full + 2;           // [2]
    uint64_t* o_final = p_ready + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_final + 1);

    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);   // provably warp-uniform role index
    const int lane = threadIdx.x & 31;
    const int unit = blockIdx.x, head = blockIdx.y, img = blockIdx.z;
    const int q0 = unit * 2 * BQ;
    const bool has_b = q0 + BQ < p.tokens;
    const int n_tiles = has_b ? 2 : 1;
    const int nkv = (p.tokens + BKV - 1) / BKV;

    if (threadIdx.x == 0) {   // one-time setup, not on the hot path
        ptx::mbar_init(q_full, 1);
        for (int i = 0; i < KV_STAGES; ++i) {
            ptx::mbar_init(&kv_full[i], 1);
            ptx::mbar_init(&kv_empty[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            ptx::mbar_init(&s_full[i], 1);
            ptx::mbar_init(&p_ready[i], 128);
        }
        ptx::mbar_init(o_final, 1);
        ptx::fence_barrier_init();
    }
    if (warp == 2) ptx::tmem_alloc<TMEM_COLS>(tmem_slot);
    if (warp == 3) {
        // constant rows 64..79 of every V^T half tile: row 64 = 1.0 (bf16 0x3F80), rows 65..79 = 0.  Every row
        // is constant, so the 128-byte swizzle permutation inside a row does not matter.
        for (int i = lane; i < KV_STAGES * 2 * (ONES_BYTES / 16); i += 32) {
            const int tile = i / (ONES_BYTES / 16), chunk = i % (ONES_BYTES / 16);   // 16-byte chunks, 8 per row
            const uint32_t v = (chunk < 8) ? 0x3F803F80u : 0u;
            uint8_t* base = s_kv + (tile >> 1) * KV_STAGE_BYTES + K_TILE_BYTES + (tile & 1) * V_HALF_BYTES + V_ROWS_BYTES;
            *reinterpret_cast<uint4*>(base + chunk * 16) = make_uint4(v, v, v, v);
        }
        ptx::fence_proxy_async();   // generic-proxy writes -> visible to the tensor core's async proxy
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // register re-balancing: the producer / MMA / allocator warpgroup needs few registers, the two softmax
    // warpgroups hold a whole 128-wide score row per thread
    // (each role region below is entered right after its own setmaxnreg so that ptxas sees the budget)

    // Roles are dispatched per WARP (uniform) and the single issuing lane is chosen with elect.sync: a branch on
    // threadIdx.x would make ptxas wrap every TMA / MMA instruction in a divergence ("waterfall") loop, because
    // their descriptor operands must live in uniform registers.  Only that one lane polls the mbarriers -- a full
    // warp spinning on try_wait measurably starves the softmax warps that share its scheduler.
    if (warp < 4) {
      ptx::setmaxnreg_dec<64>();
      if (warp == 0) {
        // ---------------- TMA producer (one elected lane) ----------------
        if (ptx::elect_one()) {
            ptx::prefetch_tmap(&tm_qk);
            ptx::prefetch_tmap(&tm_vt);
            ptx::mbar_arrive_expect_tx(q_full, n_tiles * Q_TILE_BYTES);
            for (int t = 0; t < n_tiles; ++t)
                ptx::tma_load_3d(s_q + t * Q_TILE_BYTES, &tm_qk, q_full, head * HD, q0 + t * BQ, img);
            const int vt_row = (img * p.heads + head) * HD;
            for (int j = 0; j < nkv; ++j) {
                const int st = j % KV_STAGES;
                ptx::mbar_wait(&kv_empty[st], ((j / KV_STAGES) & 1) ^ 1);
                uint8_t* dst = s_kv + st * KV_STAGE_BYTES;
                ptx::mbar_arrive_expect_tx(&kv_full[st], K_TILE_BYTES + 2 * V_ROWS_BYTES);
                ptx::tma_load_3d(dst, &tm_qk, &kv_full[st], p.D + head * HD, j * BKV, img);
                ptx::tma_load_2d(dst + K_TILE_BYTES, &tm_vt, &kv_full[st], j * BKV, vt_row);
                ptx::tma_load_2d(dst + K_TILE_BYTES + V_HALF_BYTES, &tm_vt, &kv_full[st], j * BKV + 64, vt_row);
            }
        }
      } else if (warp == 1 && ptx::elect_one()) {
        // ---------------- MMA issuer ----------------
        constexpr uint32_t idesc_s = ptx::idesc_bf16_f32(BQ, BKV);
        constexpr uint32_t idesc_o = ptx::idesc_bf16_f32(BQ, NV);
        const uint32_t q_addr = ptx::smem_u32(s_q);
        const uint32_t kv_addr = ptx::smem_u32(s_kv);
        auto issue_s = [&](int t, int st) {
            const uint64_t adesc = ptx::smem_desc_k_sw128(q_addr + t * Q_TILE_BYTES);
            const uint64_t bdesc = ptx::smem_desc_k_sw128(kv_addr + st * KV_STAGE_BYTES);
#pragma unroll
            for (int k = 0; k < HD / 16; ++k)
                ptx::umma_ss(tmem_base + COL_S + t * 128, adesc + 2 * k, bdesc + 2 * k, idesc_s, k != 0);
        };
        auto issue_pv = [&](int t, int st, bool acc) {
            const uint32_t v_addr = kv_addr + st * KV_STAGE_BYTES + K_TILE_BYTES;
#pragma unroll
            for (int ks = 0; ks < BKV / 16; ++ks) {
                const uint64_t bdesc = ptx::smem_desc_k_sw128(v_addr + (ks >> 2) * V_HALF_BYTES) + 2 * (ks & 3);
                // P: packed bf16 pairs, 8 TMEM columns per 16 keys
                ptx::umma_ts(tmem_base + COL_O + t * 128, tmem_base + COL_S + t * 128 + ks * 8, bdesc, idesc_o,
                             acc || ks != 0);
            }
        };
        ptx::mbar_wait(q_full, 0);
        ptx::mbar_wait(&kv_full[0], 0);
        ptx::tc_fence_after();
        for (int t = 0; t < n_tiles; ++t) {
            issue_s(t, 0);
            ptx::tc_commit(&s_full[t]);
        }
        for (int j = 0; j < nkv; ++j) {
            const int st = j % KV_STAGES;
            const bool more = j + 1 < nkv;
            if (more) {
                ptx::mbar_wait(&kv_full[(j + 1) % KV_STAGES], ((j + 1) / KV_STAGES) & 1);
                ptx::tc_fence_after();
            }
            for (int t = 0; t < n_tiles; ++t) {
                ptx::mbar_wait(&p_ready[t], j & 1);
                ptx::tc_fence_after();
                issue_pv(t, st, j > 0);
                if (more) {
                    issue_s(t, (j + 1) % KV_STAGES);
                    ptx::tc_commit(&s_full[t]);  // also certifies that P V of block j is complete
                }
                if (t == n_tiles - 1) {
                    ptx::tc_commit(&kv_empty[st]);
                    if (!more) ptx::tc_commit(o_final);
                }
            }
        }
      }
    } else {
        ptx::setmaxnreg_inc<216>();
        // ---------------- softmax + output ----------------
        const int t = (warp - 4) >> 2;
        if (t < n_tiles) {
            const int q = warp & 3;
            const uint32_t lane_base = static_cast<uint32_t>(q * 32) << 16;
            const uint32_t t_s = tmem_base + lane_base + COL_S + t * 128;
            const uint32_t t_o = tmem_base + lane_base + COL_O + t * 128;
            const float c = p.scale_log2e;
            float m_used = -INFINITY;
            for (int j = 0; j < nkv; ++j) {
                ptx::mbar_wait(&s_full[t], j & 1);
                ptx::tc_fence_after();
                uint32_t s[4][32];
#pragma unroll
                for (int ch = 0; ch < 4; ++ch) ptx::tmem_ld32(t_s + ch * 32, s[ch]);
                ptx::tc_wait_ld();
                const int valid = p.tokens - j * BKV;  // >= 1
                if (valid < BKV) {
#pragma unroll
                    for (int ch = 0; ch < 4; ++ch)
#pragma unroll
                        for (int i = 0; i < 32; ++i)
                            if (ch * 32 + i >= valid) s[ch][i] = 0xff800000u;  // -inf
                }
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
                const float m_blk = max3(max3(mx[0], mx[1], mx[2]), max3(mx[3], mx[4], mx[5]), fmaxf(mx[6], mx[7])) * c;
                const bool grow = m_blk > m_used + RESCALE_THRESHOLD;
                if (__any_sync(0xffffffffu, grow)) {
                    const float m_new = grow ? m_blk : m_used;
                    const float f = ptx::ex2_approx(m_used - m_new);  // 0 on the first block, 1 if unchanged
                    if (j > 0) {
                        uint32_t o[32];
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            ptx::tmem_ld32(t_o + h * 32, o);
                            ptx::tc_wait_ld();
#pragma unroll
                            for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * f);
                            ptx::tmem_st32(t_o + h * 32, o);
                        }
                        uint32_t l16[16];                       // column 64 = running denominator
                        ptx::tmem_ld16(t_o + 64, l16);
                        ptx::tc_wait_ld();
#pragma unroll
                        for (int i = 0; i < 16; ++i) l16[i] = __float_as_uint(__uint_as_float(l16[i]) * f);
                        ptx::tmem_st16(t_o + 64, l16);
                    }
                    m_used = m_new;
                }
                const float nm = -m_used;
#pragma unroll
                for (int ch = 0; ch < 4; ++ch) {
                    uint32_t pk[16];
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        float x0, x1;
                        ffma2(x0, x1, __uint_as_float(s[ch][2 * i]), __uint_as_float(s[ch][2 * i + 1]), c, nm);
                        if (POLY_EVERY > 0 && (i % POLY_EVERY) == POLY_EVERY - 1) {
                            float p0, p1;
                            exp2_poly2(x0, x1, p0, p1);
                            pk[i] = ptx::pack_bf16x2(p0, p1);
                        } else {
                            pk[i] = ptx::pack_bf16x2(ptx::ex2_approx(x0), ptx::ex2_approx(x1));
                        }
                    }
                    ptx::tmem_st16(t_s + ch * 16, pk);  // P aliases the first 64 columns of S (row-private)
                }
                ptx::tc_wait_st();
                ptx::tc_fence_before();
                ptx::mbar_arrive(&p_ready[t]);
            }
            // ---- output: O / l -> bf16, token-major (B*tokens, D) at column head*64 ----
            ptx::mbar_wait(o_final, 0);
            ptx::tc_fence_after();
            const int row = q0 + t * BQ + q * 32 + lane;
            uint32_t l16[16];
            ptx::tmem_ld16(t_o + 64, l16);
            ptx::tc_wait_ld();
            const float inv = 1.0f / __uint_as_float(l16[0]);
            __nv_bfloat16* dst = p.out + (static_cast<size_t>(img) * p.tokens + row) * p.D + head * HD;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                uint32_t o[32];
                ptx::tmem_ld32(t_o + h * 32, o);
                ptx::tc_wait_ld();
                if (row < p.tokens) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        uint4 pkt;
                        pkt.x = ptx::pack_bf16x2(__uint_as_float(o[8 * i + 0]) * inv, __uint_as_float(o[8 * i + 1]) * inv);
                        pkt.y = ptx::pack_bf16x2(__uint_as_float(o[8 * i + 2]) * inv, __uint_as_float(o[8 * i + 3]) * inv);
                        pkt.z = ptx::pack_bf16x2(__uint_as_float(o[8 * i + 4]) * inv, __uint_as_float(o[8 * i + 5]) * inv);
                        pkt.w = ptx::pack_bf16x2(__uint_as_float(o[8 * i + 6]) * inv, __uint_as_float(o[8 * i + 7]) * inv);
                        reinterpret_cast<uint4*>(dst + h * 32)[i] = pkt;
                    }
                }
            }
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 2) ptx::tmem_dealloc<TMEM_COLS>(tmem_base);
}

}  // namespace

extern "C" int vittf_attention(const void* qk, const void* vt, void* out, int B, int tokens, int tok_pad, int heads,
                               void* stream) {
    VITTF_REQUIRE(qk && vt && out, "vittf_attention: null pointer");
    VITTF_REQUIRE(B > 0 && tokens > 0 && heads > 0, "vittf_attention: empty problem");
    VITTF_REQUIRE(tok_pad >= tokens && tok_pad % 8 == 0, "vittf_attention: tok_pad=%d must be >= tokens and a multiple of 8",
                  tok_pad);
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
    static bool configured = false;
    if (!configured) {
        VITTF_CHECK_CUDA(
            cudaFuncSetAttribute(attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM_BYTES));
        configured = true;
    }
    AttnParams p{static_cast<__nv_bfloat16*>(out), tokens, heads, D, 0.125f * 1.4426950408889634f};
    dim3 grid(ceil_div(tokens, 2 * BQ), heads, B);
    attention_kernel<<<grid, ATT_THREADS, ATT_SMEM_BYTES, static_cast<cudaStream_t>(stream)>>>(tm_qk, tm_vt, p);
    VITTF_CHECK_CUDA(cudaGetLastError());
    vittf_count_launches(1);
    return VITTF_OK;
}
