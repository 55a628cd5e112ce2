// Micro-benchmark: TMEM read bandwidth (tcgen05.ld 32x32b) per SM with 4 / 8 reading warps, with and without
// concurrent MUFU work.  Decides whether the softmax warps of attention.cu are TMEM-read bound.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

template <int X>
__device__ __forceinline__ void ld(uint32_t taddr, uint32_t* r);
template <>
__device__ __forceinline__ void ld<32>(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,"
        "%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}

// MODE 0: loads only; MODE 1: loads + 128 MUFU per iteration on the loaded data
template <int MODE>
__global__ void k(uint32_t* out, long long* cyc, int iters) {
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t base = slot + (static_cast<uint32_t>((warp & 3) * 32) << 16) + ((warp >> 2) & 1) * 128;
    uint32_t acc = 0;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        uint32_t r[4][32];
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) ld<32>(base + ch * 32, r[ch]);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int ch = 0; ch < 4; ++ch)
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                if (MODE == 1) {
                    float y;
                    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(__uint_as_float(r[ch][i])));
                    acc ^= __float_as_uint(y);
                } else {
                    acc ^= r[ch][i];
                }
            }
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(slot) : "memory");
}

template <int MODE>
void run(const char* name, int threads) {
    uint32_t* out; long long* cyc;
    cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
    const int iters = 1000;
    k<MODE><<<148, threads>>>(out, cyc, iters);
    k<MODE><<<148, threads>>>(out, cyc, iters);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double c = 0; for (int i = 0; i < 148; ++i) c += h[i]; c /= 148;
    const double bytes = double(threads) * 128 * 4 * iters;
    printf("%-22s warps %2d: %.1f cyc / iteration (128 cols/thread), %.1f B/clk/SM  err=%s\n", name, threads / 32, c / iters,
           bytes / c, cudaGetErrorString(e));
    cudaFree(out); cudaFree(cyc);
}

int main() {
    for (int th : {32, 128, 256}) {
        run<0>("ldtm only", th);
        run<1>("ldtm + 128 ex2", th);
    }
    return 0;
}
