// Micro-benchmark of the exponential section of attention.cu's max-free (FAST) pass in isolation: scores come from
// TMEM (LDTM), p = 2^s by MUFU or by an FMA-pipe polynomial for NUM of every DEN pairs, packed row sums, bf16 pack, STTM.
// Variants of the polynomial: how the exponent range is guarded (4 FMNMX clamps per pair / ONE 3-input |x| maximum per
// pair tracked for a range flag), its degree (3 / 2), and how round(x) enters the exponent field (IMAD / SHL + FMUL2).
// Prints cycles per 128-score row block per warp for 1 and 2 warps per SM sub-partition, next to the pipe floors.
#include <cstdio>
#include <cstdint>
#include "../../vittf_b200/csrc/common.cuh"

template <int GUARD, int DEG, int EXPINS>
__device__ __forceinline__ void exp2_poly2(float x0, float x1, float& p0, float& p1, float& track) {
    if (GUARD == 0) {
        x0 = fmaxf(fminf(x0, 126.0f), -126.0f);
        x1 = fmaxf(fminf(x1, 126.0f), -126.0f);
    } else if (GUARD == 1) {
        asm("max.f32 %0, %0, %1, %2;" : "+f"(track) : "f"(fabsf(x0)), "f"(fabsf(x1)));
    }
    const float magic = EXPINS == 1 ? 12582912.0f + 127.0f : 12582912.0f;
    const ptx::F2 x = ptx::f2_make(x0, x1);
    const ptx::F2 t = ptx::f2_add(x, ptx::f2_make(magic, magic));
    const ptx::F2 n = ptx::f2_add(t, ptx::f2_make(-magic, -magic));
    const ptx::F2 fr = ptx::f2_fma(n, ptx::f2_make(-1.0f, -1.0f), x);
    ptx::F2 q;
    if (DEG == 3) {
        q = ptx::f2_fma(fr, ptx::f2_make(0.05508868f, 0.05508868f), ptx::f2_make(0.24260405f, 0.24260405f));
        q = ptx::f2_fma(q, fr, ptx::f2_make(0.69327623f, 0.69327623f));
        q = ptx::f2_fma(q, fr, ptx::f2_make(0.99992895f, 0.99992895f));
    } else {
        q = ptx::f2_fma(fr, ptx::f2_make(0.2402265f, 0.2402265f), ptx::f2_make(0.6931472f, 0.6931472f));
        q = ptx::f2_fma(q, fr, ptx::f2_make(1.0017247f, 1.0017247f));
    }
    float t0, t1;
    ptx::f2_get(t, t0, t1);
    if (EXPINS == 0) {
        float q0, q1;
        ptx::f2_get(q, q0, q1);
        p0 = __int_as_float(__float_as_int(q0) + (__float_as_int(t0) << 23));
        p1 = __int_as_float(__float_as_int(q1) + (__float_as_int(t1) << 23));
    } else {
        const float s0 = __int_as_float(__float_as_int(t0) << 23), s1 = __int_as_float(__float_as_int(t1) << 23);
        ptx::f2_get(ptx::f2_mul(q, ptx::f2_make(s0, s1)), p0, p1);
    }
}

template <int NUM, int DEN, int GUARD, int DEG, int EXPINS>
__global__ void __launch_bounds__(256, 1) k(uint32_t* out, long long* cyc, int iters) {
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) ptx::tmem_alloc<512>(&slot);
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t lane_base = static_cast<uint32_t>((warp & 3) * 32) << 16;
    const uint32_t t_s = slot + lane_base + ((warp >> 2) & 1) * 128;
    const uint32_t t_p = slot + lane_base + 256 + ((warp >> 2) & 1) * 64;
    {   // defined scores in TMEM: small negative exponents
        uint32_t z[32];
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
#pragma unroll
            for (int i = 0; i < 32; ++i) z[i] = __float_as_uint(-0.01f * ((threadIdx.x * 131 + ch * 32 + i) % 977));
            ptx::tmem_st32(t_s + ch * 32, z);
        }
        ptx::tc_wait_st();
    }
    float l = 0.0f, track = 0.0f;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        uint32_t s[4][32];
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) ptx::tmem_ld32(t_s + ch * 32, s[ch]);
        ptx::tc_wait_ld();
        ptx::F2 sums[4] = {{0ull}, {0ull}, {0ull}, {0ull}};
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
            uint32_t pk[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const float x0 = __uint_as_float(s[ch][2 * i]), x1 = __uint_as_float(s[ch][2 * i + 1]);
                const int pair = ch * 16 + i, r = pair % DEN;
                const bool poly = NUM > 0 && (r + 1) * NUM / DEN != r * NUM / DEN;
                float p0, p1;
                if (poly) exp2_poly2<GUARD, DEG, EXPINS>(x0, x1, p0, p1, track);
                else { p0 = ptx::ex2_approx(x0); p1 = ptx::ex2_approx(x1); }
                sums[i & 3] = ptx::f2_add(sums[i & 3], ptx::f2_make(p0, p1));
                pk[i] = ptx::pack_bf16x2(p0, p1);
            }
            ptx::tmem_st16(t_p + ch * 16, pk);
        }
        float a0, a1, b0, b1;
        ptx::f2_get(ptx::f2_add(sums[0], sums[1]), a0, a1);
        ptx::f2_get(ptx::f2_add(sums[2], sums[3]), b0, b1);
        l += (a0 + a1) + (b0 + b1);
        ptx::tc_wait_st();
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = __float_as_uint(l + track);
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 0) ptx::tmem_dealloc<512>(slot);
}

template <int NUM, int DEN, int GUARD, int DEG, int EXPINS>
void run() {
    uint32_t* out; long long* cyc;
    cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
    const int iters = 400;
    double r[2];
    int idx = 0;
    for (int threads : {128, 256}) {
        k<NUM, DEN, GUARD, DEG, EXPINS><<<148, threads>>>(out, cyc, iters);
        k<NUM, DEN, GUARD, DEG, EXPINS><<<148, threads>>>(out, cyc, iters);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) printf("CUDA error: %s\n", cudaGetErrorString(e));
        long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
        double cy = 0; for (int i = 0; i < 148; ++i) cy += h[i]; cy /= 148;
        r[idx++] = cy / iters / (threads / 128);
    }
    const double p = static_cast<double>(NUM) / DEN;
    printf("poly %d/%-2d guard %d deg %d expins %d : %7.1f (1 warp/SMSP) %7.1f (2 warps/SMSP: per warp-slot)   MUFU floor %6.1f\n", NUM, DEN, GUARD,
           DEG, EXPINS, r[0], r[1], 1024.0 * (1.0 - p));
    cudaFree(out); cudaFree(cyc);
}


// Split rows: every thread owns HALF a score row (64 scores) of its tile, 4 warps per SM sub-partition (2 tiles x 2 column
// halves) instead of 2 -- what 16 softmax warps per CTA would run.  PRELOAD: both 32-score chunks are requested before the
// first is consumed (64 score registers live) / each chunk is loaded right before it is consumed (32 live; the other three
// warps of the scheduler hide the TMEM latency).  Prints cycles per iteration = per 8192 scores of an SM sub-partition,
// i.e. comparable to 2 x the "2 warps/SMSP" column above.
template <int NUM, int DEN, int DEG, bool PRELOAD>
__global__ void __launch_bounds__(512, 1) k_split(uint32_t* out, long long* cyc, int iters) {
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) ptx::tmem_alloc<512>(&slot);
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const int q = warp & 3, t = (warp >> 2) & 1, h = warp >> 3;
    const uint32_t lane_base = static_cast<uint32_t>(q * 32) << 16;
    const uint32_t t_s = slot + lane_base + t * 128 + h * 64;
    const uint32_t t_p = slot + lane_base + 256 + t * 64 + h * 32;
    {
        uint32_t z[32];
#pragma unroll
        for (int ch = 0; ch < 2; ++ch) {
#pragma unroll
            for (int i = 0; i < 32; ++i) z[i] = __float_as_uint(-0.01f * ((threadIdx.x * 131 + ch * 32 + i) % 977));
            ptx::tmem_st32(t_s + ch * 32, z);
        }
        ptx::tc_wait_st();
    }
    float l = 0.0f, track = 0.0f;
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        uint32_t s[2][32];
        if (PRELOAD) {
            ptx::tmem_ld32(t_s, s[0]);
            ptx::tmem_ld32(t_s + 32, s[1]);
            ptx::tc_wait_ld();
        }
        ptx::F2 sums[4] = {{0ull}, {0ull}, {0ull}, {0ull}};
#pragma unroll
        for (int ch = 0; ch < 2; ++ch) {
            if (!PRELOAD) {
                ptx::tmem_ld32(t_s + ch * 32, s[ch]);
                ptx::tc_wait_ld();
            }
            uint32_t pk[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const float x0 = __uint_as_float(s[ch][2 * i]), x1 = __uint_as_float(s[ch][2 * i + 1]);
                const int pair = ch * 16 + i, r = pair % DEN;
                const bool poly = NUM > 0 && (r + 1) * NUM / DEN != r * NUM / DEN;
                float p0, p1;
                if (poly) exp2_poly2<1, DEG, 0>(x0, x1, p0, p1, track);
                else { p0 = ptx::ex2_approx(x0); p1 = ptx::ex2_approx(x1); }
                sums[i & 3] = ptx::f2_add(sums[i & 3], ptx::f2_make(p0, p1));
                pk[i] = ptx::pack_bf16x2(p0, p1);
            }
            ptx::tmem_st16(t_p + ch * 16, pk);
        }
        float a0, a1, b0, b1;
        ptx::f2_get(ptx::f2_add(sums[0], sums[1]), a0, a1);
        ptx::f2_get(ptx::f2_add(sums[2], sums[3]), b0, b1);
        l += (a0 + a1) + (b0 + b1);
        ptx::tc_wait_st();
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = __float_as_uint(l + track);
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 0) ptx::tmem_dealloc<512>(slot);
}

template <int NUM, int DEN, int DEG, bool PRELOAD>
void run_split() {
    uint32_t* out; long long* cyc;
    cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
    const int iters = 400;
    k_split<NUM, DEN, DEG, PRELOAD><<<148, 512>>>(out, cyc, iters);
    k_split<NUM, DEN, DEG, PRELOAD><<<148, 512>>>(out, cyc, iters);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) printf("CUDA error: %s\n", cudaGetErrorString(e));
    long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double cy = 0; for (int i = 0; i < 148; ++i) cy += h[i]; cy /= 148;
    const double p = static_cast<double>(NUM) / DEN;
    printf("split rows, poly %d/%-2d deg %d %s : %7.1f cycles per 8192 scores of an SMSP (4 warps/SMSP)   MUFU floor %6.1f\n", NUM, DEN, DEG,
           PRELOAD ? "preload " : "per-chunk", cy / iters, 2048.0 * (1.0 - p));
    cudaFree(out); cudaFree(cyc);
}

int main() {
    printf("cycles per 128-score row block (32 rows); MUFU floor = 128 * (1 - p) * 8 cycles per warp\n");
    run<0, 4, 0, 3, 0>();
    run<1, 4, 0, 3, 0>();      // the kernel's current mix
    run<1, 4, 1, 3, 0>();
    run<1, 4, 1, 3, 1>();
    run<1, 4, 1, 2, 0>();
    run<3, 8, 0, 3, 0>();
    run<3, 8, 1, 3, 0>();
    run<3, 8, 1, 3, 1>();
    run<3, 8, 1, 2, 0>();
    run<3, 8, 1, 2, 1>();
    run<7, 16, 1, 3, 0>();
    run<7, 16, 1, 3, 1>();
    run<7, 16, 1, 2, 0>();
    run<1, 2, 0, 3, 0>();
    run<1, 2, 1, 3, 0>();
    run<1, 2, 1, 3, 1>();
    run<1, 2, 1, 2, 0>();
    run<1, 2, 1, 2, 1>();
    run<1, 2, 2, 3, 0>();      // no guard at all (lower bound of the polynomial cost)
    run<5, 8, 1, 2, 0>();
    run_split<1, 4, 2, true>();
    run_split<1, 4, 2, false>();
    run_split<3, 8, 2, true>();
    run_split<3, 8, 2, false>();
    run_split<1, 2, 2, false>();
    run_split<1, 4, 3, false>();
    run_split<0, 4, 2, false>();
    return 0;
}
