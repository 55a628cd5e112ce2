// Micro-benchmark: MUFU.EX2 issue rate per SM for f32, f16x2 and bf16x2 operands (decides the softmax design
// of attention.cu).  One CTA per SM; cycles from clock64 around an unrolled dependent-free loop.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int MODE>
__global__ void k(uint32_t* out, long long* cyc, int iters) {
    uint32_t r[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) r[i] = 0x3c003800u + threadIdx.x * 8 + i;   // small positive halves / a float
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+r"(r[i]));
            if (MODE == 1) asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(r[i]));
            if (MODE == 2) asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(r[i]));
            if (MODE == 3) asm volatile("tanh.approx.f16x2 %0, %0;" : "+r"(r[i]));
        }
    }
    long long t1 = clock64();
    uint32_t acc = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) acc ^= r[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name, int threads) {
    uint32_t* out; long long* cyc;
    cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
    const int iters = 2048;
    k<MODE><<<148, threads>>>(out, cyc, iters);
    k<MODE><<<148, threads>>>(out, cyc, iters);
    cudaDeviceSynchronize();
    long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double c = 0; for (int i = 0; i < 148; ++i) c += h[i]; c /= 148;
    double instr = double(threads) * 8 * iters;
    printf("%-28s threads %4d: %.2f lane-instr/clk/SM (%.2f results/clk/SM)  err=%s\n", name, threads, instr / c,
           instr / c * (MODE == 0 ? 1 : 2), cudaGetErrorString(cudaGetLastError()));
    cudaFree(out); cudaFree(cyc);
}

int main() {
    for (int th : {128, 256, 512}) {
        run<0>("ex2.approx.ftz.f32", th);
        run<1>("ex2.approx.f16x2", th);
        run<2>("ex2.approx.ftz.bf16x2", th);
        run<3>("tanh.approx.f16x2", th);
    }
    return 0;
}
