// Micro-benchmark: issue cost (cycles per warp instruction per SM sub-partition) of the instructions in the
// softmax inner loop, alone and in pairs, to find which share a pipe.  1 and 2 warps per sub-partition.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

enum { FFMA, FFMA2, FADD2, F2FP, PRMT, FMNMX, FMNMX3, IMAD, MUFU, SHL, IADD, NKIND };
const char* names[] = {"FFMA", "FFMA2", "FADD2", "F2FP.BF16", "PRMT", "FMNMX", "FMNMX3", "IMAD", "MUFU.EX2", "SHF/SHL", "IADD3"};

template <int KIND>
__device__ __forceinline__ void op(uint32_t& a, uint32_t& b, unsigned long long& w) {
    if (KIND == FFMA) asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(*(float*)&a) : "f"(*(float*)&b));
    if (KIND == FFMA2) asm volatile("fma.rn.f32x2 %0, %0, %0, %0;" : "+l"(w));
    if (KIND == FADD2) asm volatile("add.rn.f32x2 %0, %0, %0;" : "+l"(w));
    if (KIND == F2FP) asm volatile("cvt.rn.bf16x2.f32 %0, %0, %1;" : "+r"(a) : "f"(*(float*)&b));
    if (KIND == PRMT) asm volatile("prmt.b32 %0, %0, %1, 0x7632;" : "+r"(a) : "r"(b));
    if (KIND == FMNMX) asm volatile("max.f32 %0, %0, %1;" : "+f"(*(float*)&a) : "f"(*(float*)&b));
    if (KIND == FMNMX3) asm volatile("max.f32 %0, %0, %1, %1;" : "+f"(*(float*)&a) : "f"(*(float*)&b));
    if (KIND == IMAD) asm volatile("mad.lo.s32 %0, %0, %1, %1;" : "+r"(a) : "r"(b));
    if (KIND == MUFU) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(*(float*)&a));
    if (KIND == SHL) asm volatile("shl.b32 %0, %0, 3;" : "+r"(a));
    if (KIND == IADD) asm volatile("add.s32 %0, %0, %1;" : "+r"(a) : "r"(b));
}

template <int K1, int K2>
__global__ void k(uint32_t* out, long long* cyc, int iters) {
    uint32_t a[8], b[8]; unsigned long long w[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { a[i] = threadIdx.x + i; b[i] = 0x3f800000u + i; w[i] = 0x3f8000003f800000ull + i; }
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            op<K1>(a[i], b[i], w[i]);
            if (K2 >= 0) op<(K2 >= 0 ? K2 : 0)>(b[i], a[(i + 1) & 7], w[(i + 4) & 7]);
        }
    }
    long long t1 = clock64();
    uint32_t acc = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) acc ^= a[i] ^ b[i] ^ (uint32_t)w[i] ^ (uint32_t)(w[i] >> 32);
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int K1, int K2>
void run() {
    uint32_t* out; long long* cyc;
    cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
    const int iters = 1024;
    printf("%-10s %-10s", names[K1], K2 >= 0 ? names[K2 >= 0 ? K2 : 0] : "-");
    for (int threads : {128, 256, 512}) {
        k<K1, K2><<<148, threads>>>(out, cyc, iters);
        k<K1, K2><<<148, threads>>>(out, cyc, iters);
        cudaDeviceSynchronize();
        long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
        double cy = 0; for (int i = 0; i < 148; ++i) cy += h[i]; cy /= 148;
        const double groups = double(iters) * 8 * (threads / 128);   // (K1[,K2]) groups issued per sub-partition
        printf("  %dw/SMSP: %5.2f cyc/group", threads / 128, cy / groups);
    }
    printf("\n");
    cudaFree(out); cudaFree(cyc);
}

int main() {
    run<FFMA, -1>(); run<FFMA2, -1>(); run<FADD2, -1>(); run<F2FP, -1>(); run<PRMT, -1>(); run<FMNMX, -1>(); run<FMNMX3, -1>();
    run<IMAD, -1>(); run<MUFU, -1>(); run<SHL, -1>(); run<IADD, -1>();
    run<FFMA2, F2FP>(); run<FFMA2, PRMT>(); run<FFMA2, FMNMX3>(); run<FFMA2, IMAD>(); run<FFMA2, FADD2>(); run<FFMA2, FFMA>();
    run<F2FP, PRMT>(); run<F2FP, FMNMX3>(); run<F2FP, IMAD>(); run<MUFU, FFMA2>(); run<MUFU, F2FP>(); run<FMNMX3, PRMT>();
    run<FFMA2, IADD>(); run<FFMA2, SHL>(); run<FFMA, IADD>(); run<FFMA, FMNMX>();
    return 0;
}
