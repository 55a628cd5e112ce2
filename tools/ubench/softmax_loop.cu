// Micro-benchmark of the exponential section of attention.cu in isolation (no MMA, no barriers): cycles per
// 128-score row block for one or two warps per SM sub-partition, for several instruction mixes.
#include <cstdio>
#include <cstdint>
#include "../../vittf_b200/csrc/common.cuh"

__device__ __forceinline__ void ffma2(float& x0, float& x1, float s0, float s1, float c, float nm) {
    asm("{\n.reg .b64 ra, rb, rc, rd;\nmov.b64 ra, {%2, %3};\nmov.b64 rb, {%4, %4};\nmov.b64 rc, {%5, %5};\n"
        "fma.rn.f32x2 rd, ra, rb, rc;\nmov.b64 {%0, %1}, rd;\n}\n"
        : "=f"(x0), "=f"(x1) : "f"(s0), "f"(s1), "f"(c), "f"(nm));
}
__device__ __forceinline__ void exp2_poly2(float x0, float x1, float& p0, float& p1) {
    const float magic = 12582912.0f;
    x0 = fmaxf(x0, -126.0f);
    x1 = fmaxf(x1, -126.0f);
    const ptx::F2 x = ptx::f2_make(x0, x1);
    const ptx::F2 t = ptx::f2_add(x, ptx::f2_make(magic, magic));
    const ptx::F2 n = ptx::f2_add(t, ptx::f2_make(-magic, -magic));
    const ptx::F2 fr = ptx::f2_fma(n, ptx::f2_make(-1.0f, -1.0f), x);
    ptx::F2 q = ptx::f2_fma(fr, ptx::f2_make(0.05508868f, 0.05508868f), ptx::f2_make(0.24260405f, 0.24260405f));
    q = ptx::f2_fma(q, fr, ptx::f2_make(0.69327623f, 0.69327623f));
    q = ptx::f2_fma(q, fr, ptx::f2_make(0.99992895f, 0.99992895f));
    float t0, t1, q0, q1;
    ptx::f2_get(t, t0, t1);
    ptx::f2_get(q, q0, q1);
    p0 = __int_as_float(__float_as_int(q0) + (__float_as_int(t0) << 23));
    p1 = __int_as_float(__float_as_int(q1) + (__float_as_int(t1) << 23));
}

// OPT bits: 1 = scale FFMA2, 2 = MUFU exp (else pass-through), 4 = row sum, 8 = pack + STTM, 16 = poly on every 4th pair,
//           32 = row max phase (FMNMX3) before the section, 64 = poly 3 of 8
template <int OPT>
__global__ void __launch_bounds__(256, 1) k(uint32_t* out, long long* cyc, int iters, float c) {
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) ptx::tmem_alloc<512>(&slot);
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t t_p = slot + (static_cast<uint32_t>((warp & 3) * 32) << 16) + ((warp >> 2) & 1) * 64;
    float s[4][32];
#pragma unroll
    for (int ch = 0; ch < 4; ++ch)
#pragma unroll
        for (int i = 0; i < 32; ++i) s[ch][i] = -0.01f * ((threadIdx.x * 131 + ch * 32 + i) % 977);
    float l = 0.0f, m = 0.0f;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        if (OPT & 32) {
            float mx[8];
#pragma unroll
            for (int a = 0; a < 8; ++a) mx[a] = s[a >> 1][(a & 1) * 16];
#pragma unroll
            for (int a = 0; a < 8; ++a) {
                const float* sv = &s[a >> 1][(a & 1) * 16];
#pragma unroll
                for (int i = 1; i < 15; i += 2) asm("max.f32 %0, %0, %1, %2;" : "+f"(mx[a]) : "f"(sv[i]), "f"(sv[i + 1]));
                mx[a] = fmaxf(mx[a], sv[15]);
            }
            m = fmaxf(fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3])), fmaxf(fmaxf(mx[4], mx[5]), fmaxf(mx[6], mx[7]))) * 1e-30f;
        }
        if (OPT & 128) {
#pragma unroll
            for (int ch = 0; ch < 4; ++ch) ptx::tmem_ld32(t_p - ((warp >> 2) & 1) * 64 + 256 + ch * 32, reinterpret_cast<uint32_t(&)[32]>(s[ch]));
            ptx::tc_wait_ld();
        }
        const float nm = -(l * 1e-30f) - m;          // loop-carried: nothing can be hoisted out of the iteration
        ptx::F2 sums[4] = {{0ull}, {0ull}, {0ull}, {0ull}};
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
            uint32_t pk[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                float x0 = s[ch][2 * i], x1 = s[ch][2 * i + 1];
                if (!(OPT & 128)) { x0 += nm; x1 += nm; }
                if (OPT & 1) ffma2(x0, x1, s[ch][2 * i], s[ch][2 * i + 1], c, nm);
                float p0 = x0, p1 = x1;
                const int pair = ch * 16 + i;
                const bool poly = ((OPT & 256) && (pair % 8) == 7) || ((OPT & 512) && ((pair % 16) == 4 || (pair % 16) == 9 || (pair % 16) == 15)) ||
                                  ((OPT & 16) && (pair % 4) == 3) || ((OPT & 64) && ((pair % 8) == 2 || (pair % 8) == 5 || (pair % 8) == 7));
                if (OPT & 2) {
                    if (poly) exp2_poly2(x0, x1, p0, p1);
                    else { p0 = ptx::ex2_approx(x0); p1 = ptx::ex2_approx(x1); }
                }
                if (OPT & 4) sums[i & 3] = ptx::f2_add(sums[i & 3], ptx::f2_make(p0, p1));
                else l += p0 * p1;
                pk[i] = (OPT & 8) ? ptx::pack_bf16x2(p0, p1) : __float_as_uint(p0) ^ __float_as_uint(p1);
            }
            if (OPT & 8) ptx::tmem_st16(t_p + ch * 16, pk);
            else {
#pragma unroll
                for (int i = 0; i < 16; ++i) l += __uint_as_float(pk[i] & 0x3fffffff);
            }
        }
        if (OPT & 4) {
            float a0, a1, b0, b1;
            ptx::f2_get(ptx::f2_add(sums[0], sums[1]), a0, a1);
            ptx::f2_get(ptx::f2_add(sums[2], sums[3]), b0, b1);
            l += (a0 + a1) + (b0 + b1);
        }
        if (OPT & 8) ptx::tc_wait_st();
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = __float_as_uint(l);
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 0) ptx::tmem_dealloc<512>(slot);
}

template <int OPT>
void run(const char* name) {
    uint32_t* out; long long* cyc;
    cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
    const int iters = 500;
    for (int threads : {128, 256}) {
        k<OPT><<<148, threads>>>(out, cyc, iters, 0.18f);
        k<OPT><<<148, threads>>>(out, cyc, iters, 0.18f);
        cudaError_t e = cudaDeviceSynchronize();
        long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
        double cy = 0; for (int i = 0; i < 148; ++i) cy += h[i]; cy /= 148;
        printf("%-44s warps/SMSP %d: %7.1f cyc per row block per warp-slot (%s)\n", name, threads / 128, cy / iters / (threads / 128),
               cudaGetErrorString(e));
    }
    cudaFree(out); cudaFree(cyc);
}

int main() {
    run<2>("MUFU only");
    run<1 | 2>("FFMA2 + MUFU");
    run<1 | 2 | 4>("FFMA2 + MUFU + FADD2 sums");
    run<1 | 2 | 8>("FFMA2 + MUFU + pack/STTM");
    run<1 | 2 | 4 | 8>("full, no poly");
    run<1 | 2 | 4 | 8 | 16>("full, poly 1/4");
    run<1 | 2 | 4 | 8 | 64>("full, poly 3/8");
    run<1 | 2 | 4 | 8 | 16 | 32>("full, poly 1/4, + row max");
    run<1 | 2 | 4 | 8 | 64 | 32>("full, poly 3/8, + row max");
    run<1 | 4 | 8>("no exp at all (FFMA2 + sums + pack)");
    run<128 | 2 | 4 | 8>("LDTM-fed: MUFU + sums + pack, no poly");
    run<128 | 2 | 4 | 8 | 256>("LDTM-fed: MUFU + sums + pack, poly 1/8");
    run<128 | 2 | 4 | 8 | 512>("LDTM-fed: MUFU + sums + pack, poly 3/16");
    run<128 | 2 | 4 | 8 | 16>("LDTM-fed: MUFU + sums + pack, poly 1/4");
    run<128 | 2 | 4 | 8 | 64>("LDTM-fed: MUFU + sums + pack, poly 3/8");
    run<128 | 2 | 8 | 16>("LDTM-fed: MUFU + pack, poly 1/4, no sums");
    run<128 | 2 | 8 | 64>("LDTM-fed: MUFU + pack, poly 3/8, no sums");
    run<128 | 1 | 2 | 4 | 8 | 16 | 32>("LDTM-fed: current kernel mix (scale, max, poly 1/4)");
    return 0;
}
