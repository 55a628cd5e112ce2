// Linear layers of the ViT (attn.qkv / attn.proj / mlp.fc1 / mlp.fc2 of the hub DINO model
// that /root/reference/infer.py:177 runs): C = A W^T + bias with fused epilogues.
//
// B200 design: persistent, warp-specialised kernel, one CTA per SM.
//   warp 0  : TMA producer  (cp.async.bulk.tensor, 128B-swizzled 128x64 / BNx64 bf16 tiles, STAGES-deep ring)
//   warp 1  : MMA issuer    (one thread, tcgen05.mma cta_group::1 kind::f16, 128 x BN x 16, fp32 accum in TMEM)
//   warp 2  : TMEM allocator
//   warps 4-7: epilogue     (tcgen05.ld 32x32b, one accumulator row per thread, bias / GELU / residual / split)
// The accumulator is double-buffered in TMEM (2 x BN columns) so the epilogue of tile i overlaps
// the main loop of tile i+1 -- essential here because K is short (384..3072).
#include "common.cuh"

namespace {

constexpr int BM = 128;
constexpr int BK = 64;  // 64 bf16 = 128 B = one swizzle row
constexpr int GEMM_THREADS = 256;

template <int BN>
struct GemmCfg {
    static constexpr int STAGES = BN <= 128 ? 6 : 4;
    static constexpr int A_BYTES = BM * BK * 2;
    static constexpr int B_BYTES = BN * BK * 2;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int TMEM_COLS = 2 * BN <= 256 ? 256 : 512;
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 512;
};

struct GemmParams {
    const float* bias;
    void* out;
    void* out2;
    int M, N, K;
    int tokens, tok_pad, heads;
};

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }

// Epilogue for 32 consecutive accumulator columns [n, n+32) of one row.
template <int EPI>
__device__ __forceinline__ void epilogue_store(const GemmParams& p, int row, int n, const uint32_t (&acc)[32]) {
    float v[32];
#pragma unroll
    for (int i = 0; i < 32; i += 4) {
        const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + n + i));
        v[i + 0] = __uint_as_float(acc[i + 0]) + b.x;
        v[i + 1] = __uint_as_float(acc[i + 1]) + b.y;
        v[i + 2] = __uint_as_float(acc[i + 2]) + b.z;
        v[i + 3] = __uint_as_float(acc[i + 3]) + b.w;
    }
    if (row >= p.M) return;
    if constexpr (EPI == VITTF_EPI_BIAS_BF16 || EPI == VITTF_EPI_BIAS_GELU_BF16) {
        if constexpr (EPI == VITTF_EPI_BIAS_GELU_BF16) {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = gelu_erf(v[i]);
        }
        uint4* dst = reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(p.out) + static_cast<size_t>(row) * p.N + n);
#pragma unroll
        for (int i = 0; i < 4; ++i)
            dst[i] = make_uint4(ptx::pack_bf16x2(v[8 * i + 0], v[8 * i + 1]), ptx::pack_bf16x2(v[8 * i + 2], v[8 * i + 3]),
                                ptx::pack_bf16x2(v[8 * i + 4], v[8 * i + 5]), ptx::pack_bf16x2(v[8 * i + 6], v[8 * i + 7]));
    } else if constexpr (EPI == VITTF_EPI_BIAS_RESID_F32) {
        float4* dst = reinterpret_cast<float4*>(static_cast<float*>(p.out) + static_cast<size_t>(row) * p.N + n);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            float4 r = dst[i];
            r.x += v[4 * i + 0];
            r.y += v[4 * i + 1];
            r.z += v[4 * i + 2];
            r.w += v[4 * i + 3];
            dst[i] = r;
        }
    } else if constexpr (EPI == VITTF_EPI_QKV_SPLIT) {
        const int two_d = (p.N / 3) * 2;
        if (n < two_d) {  // Q and K thirds stay token-major
            uint4* dst =
                reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(p.out) + static_cast<size_t>(row) * two_d + n);
#pragma unroll
            for (int i = 0; i < 4; ++i)
                dst[i] =
                    make_uint4(ptx::pack_bf16x2(v[8 * i + 0], v[8 * i + 1]), ptx::pack_bf16x2(v[8 * i + 2], v[8 * i + 3]),
                               ptx::pack_bf16x2(v[8 * i + 4], v[8 * i + 5]), ptx::pack_bf16x2(v[8 * i + 6], v[8 * i + 7]));
        } else {  // V third is stored transposed per (image, head): vt[(b*heads+h)*64 + d][token]
            const int img = row / p.tokens;
            const int tok = row - img * p.tokens;
            const int dcol = n - two_d;  // head*64 + d
            __nv_bfloat16* dst = static_cast<__nv_bfloat16*>(p.out2) +
                                 (static_cast<size_t>(img) * p.heads * 64 + dcol) * p.tok_pad + tok;
#pragma unroll
            for (int i = 0; i < 32; ++i) dst[static_cast<size_t>(i) * p.tok_pad] = __float2bfloat16_rn(v[i]);
        }
    } else if constexpr (EPI == VITTF_EPI_KFEAT_F16) {
        const int img = row / p.tokens;
        const int tok = row - img * p.tokens;
        if (tok == 0) return;  // CLS dropped (infer.py:202 `[:, 1:]`)
        const size_t orow = static_cast<size_t>(img) * (p.tokens - 1) + (tok - 1);
        uint4* dst = reinterpret_cast<uint4*>(static_cast<__half*>(p.out) + orow * p.N + n);
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

template <int BN, int EPI>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
    gemm_bf16_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b, GemmParams p) {
    using Cfg = GemmCfg<BN>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::STAGES * Cfg::STAGE_BYTES);
    uint64_t* full = bars;
    uint64_t* empty = bars + Cfg::STAGES;
    uint64_t* tmem_full = bars + 2 * Cfg::STAGES;
    uint64_t* tmem_empty = tmem_full + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int m_tiles = (p.M + BM - 1) / BM;
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
            ptx::mbar_init(&tmem_empty[i], 128);
        }
        ptx::fence_barrier_init();
    }
    if (warp == 2) ptx::tmem_alloc<Cfg::TMEM_COLS>(tmem_slot);
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (threadIdx.x == 0) {
        // ---------------- TMA producer ----------------
        ptx::prefetch_tmap(&tm_a);
        ptx::prefetch_tmap(&tm_b);
        int stage = 0;
        uint32_t phase = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            const int m_blk = tile / n_tiles, n_blk = tile - m_blk * n_tiles;
            for (int kb = 0; kb < num_kb; ++kb) {
                ptx::mbar_wait(&empty[stage], phase ^ 1);
                uint8_t* sa = smem + stage * Cfg::STAGE_BYTES;
                ptx::mbar_arrive_expect_tx(&full[stage], Cfg::STAGE_BYTES);
                ptx::tma_load_2d(sa, &tm_a, &full[stage], kb * BK, m_blk * BM);
                ptx::tma_load_2d(sa + Cfg::A_BYTES, &tm_b, &full[stage], kb * BK, n_blk * BN);
                if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (threadIdx.x == 32) {
        // ---------------- MMA issuer ----------------
        constexpr uint32_t idesc = ptx::idesc_bf16_f32(BM, BN);
        int stage = 0;
        uint32_t phase = 0;
        int it = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
            const int as = it & 1;
            ptx::mbar_wait(&tmem_empty[as], ((it >> 1) & 1) ^ 1);
            ptx::tc_fence_after();
            const uint32_t d_tmem = tmem_base + as * BN;
            for (int kb = 0; kb < num_kb; ++kb) {
                ptx::mbar_wait(&full[stage], phase);
                ptx::tc_fence_after();
                const uint32_t sa = ptx::smem_u32(smem + stage * Cfg::STAGE_BYTES);
                const uint64_t adesc = ptx::smem_desc_k_sw128(sa);
                const uint64_t bdesc = ptx::smem_desc_k_sw128(sa + Cfg::A_BYTES);
#pragma unroll
                for (int k = 0; k < BK / 16; ++k)  // +32 B per K=16 step inside the swizzle row
                    ptx::umma_ss(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
                ptx::tc_commit(&empty[stage]);
                if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
            }
            ptx::tc_commit(&tmem_full[as]);
        }
    } else if (warp >= 4) {
        // ---------------- epilogue ----------------
        const int q = warp & 3;
        int it = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
            const int m_blk = tile / n_tiles, n_blk = tile - m_blk * n_tiles;
            const int as = it & 1;
            ptx::mbar_wait(&tmem_full[as], (it >> 1) & 1);
            ptx::tc_fence_after();
            const int row = m_blk * BM + q * 32 + lane;
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * BN;
#pragma unroll 1
            for (int c = 0; c < BN / 32; ++c) {
                uint32_t acc[32];
                ptx::tmem_ld32(taddr + c * 32, acc);
                ptx::tc_wait_ld();
                epilogue_store<EPI>(p, row, n_blk * BN + c * 32, acc);
            }
            ptx::tc_fence_before();
            ptx::mbar_arrive(&tmem_empty[as]);
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 2) ptx::tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
}

template <int BN, int EPI>
int launch_gemm(const void* A, const void* W, const GemmParams& p, cudaStream_t stream) {
    using Cfg = GemmCfg<BN>;
    CUtensorMap tm_a, tm_b;
    {
        uint64_t dims[2] = {static_cast<uint64_t>(p.K), static_cast<uint64_t>(p.M)};
        uint64_t strides[1] = {static_cast<uint64_t>(p.K) * 2};
        uint32_t box[2] = {BK, BM};
        VITTF_CHECK(vittf_make_tmap(&tm_a, A, 2, 2, dims, strides, box, true));
    }
    {
        uint64_t dims[2] = {static_cast<uint64_t>(p.K), static_cast<uint64_t>(p.N)};
        uint64_t strides[1] = {static_cast<uint64_t>(p.K) * 2};
        uint32_t box[2] = {BK, BN};
        VITTF_CHECK(vittf_make_tmap(&tm_b, W, 2, 2, dims, strides, box, true));
    }
    auto kern = gemm_bf16_kernel<BN, EPI>;
    static bool configured = false;
    if (!configured) {
        VITTF_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
        configured = true;
    }
    const int tiles = ceil_div(p.M, BM) * (p.N / BN);
    const int grid = tiles < vittf_num_sms() ? tiles : vittf_num_sms();
    kern<<<grid, GEMM_THREADS, Cfg::SMEM_BYTES, stream>>>(tm_a, tm_b, p);
    VITTF_CHECK_CUDA(cudaGetLastError());
    vittf_count_launches(1);
    return VITTF_OK;
}

}  // namespace

extern "C" int vittf_gemm_bf16(const void* A, const void* W, const float* bias, void* out, void* out2, int M, int N,
                               int K, int epi, int tokens, int tok_pad, void* stream) {
    VITTF_REQUIRE(A && W && bias && out, "vittf_gemm_bf16: null pointer");
    VITTF_REQUIRE(M > 0 && N > 0 && K > 0, "vittf_gemm_bf16: empty problem M=%d N=%d K=%d", M, N, K);
    VITTF_REQUIRE(K % BK == 0, "vittf_gemm_bf16: K=%d must be a multiple of %d", K, BK);
    VITTF_REQUIRE(N % 128 == 0, "vittf_gemm_bf16: N=%d must be a multiple of 128", N);
    GemmParams p{bias, out, out2, M, N, K, tokens, tok_pad, 0};
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    switch (epi) {
        case VITTF_EPI_BIAS_BF16: return launch_gemm<128, VITTF_EPI_BIAS_BF16>(A, W, p, s);
        case VITTF_EPI_BIAS_GELU_BF16: return launch_gemm<128, VITTF_EPI_BIAS_GELU_BF16>(A, W, p, s);
        case VITTF_EPI_BIAS_RESID_F32: return launch_gemm<128, VITTF_EPI_BIAS_RESID_F32>(A, W, p, s);
        case VITTF_EPI_QKV_SPLIT:
            VITTF_REQUIRE(out2 && N % 3 == 0 && (N / 3) % 64 == 0 && tokens > 0 && tok_pad >= tokens && M % tokens == 0,
                          "vittf_gemm_bf16: bad QKV split arguments (N=%d tokens=%d tok_pad=%d M=%d)", N, tokens,
                          tok_pad, M);
            p.heads = N / 3 / 64;
            return launch_gemm<128, VITTF_EPI_QKV_SPLIT>(A, W, p, s);
        case VITTF_EPI_KFEAT_F16:
            VITTF_REQUIRE(tokens > 1 && M % tokens == 0, "vittf_gemm_bf16: bad K-feature arguments");
            return launch_gemm<128, VITTF_EPI_KFEAT_F16>(A, W, p, s);
        default: VITTF_REQUIRE(false, "vittf_gemm_bf16: unknown epilogue %d", epi);
    }
    return VITTF_OK;
}
