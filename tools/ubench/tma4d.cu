// Stand-alone check of a 4-D cp.async.bulk.tensor load (the staging of sim_lowres_tma_kernel): which box shapes /
// start coordinates the hardware accepts, and that out-of-volume elements arrive as zeros.
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__global__ void k(const __grid_constant__ CUtensorMap tm, int c0, int c1, int c2, int c3, int bytes, __half* out) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + ((bytes + 127) / 128) * 128);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
        asm volatile(
            "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::
                "r"(smem_u32(smem)), "l"(&tm), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
    }
    uint32_t ok = 0;
    long long t0 = clock64();
    while (!ok && clock64() - t0 < 20000000LL)
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(smem_u32(bar)) : "memory");
    if (threadIdx.x == 0 && !ok) printf("  TIMEOUT\n");
    for (int i = threadIdx.x; i < bytes / 2; i += blockDim.x) out[i] = reinterpret_cast<__half*>(smem)[i];
}

int main() {
    EncodeFn enc = nullptr;
    cudaDriverEntryPointQueryResult qr;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", reinterpret_cast<void**>(&enc), cudaEnableDefault, &qr);
    const int F = 16, W = 16, H = 16, D = 16;
    std::vector<__half> h(static_cast<size_t>(F) * W * H * D);
    for (size_t i = 0; i < h.size(); ++i) h[i] = __float2half(static_cast<float>(i % 2039));
    __half *dv, *dout;
    cudaMalloc(&dv, h.size() * 2);
    cudaMalloc(&dout, 1 << 20);
    cudaMemcpy(dv, h.data(), h.size() * 2, cudaMemcpyHostToDevice);
    struct Case { cuuint32_t box[4]; int c[4]; };
    const Case cases[] = {{{16, 4, 2, 8}, {0, 0, 0, 0}}, {{24, 4, 2, 8}, {0, 0, 0, 0}}, {{24, 10, 5, 8}, {0, 0, 0, 0}},
                          {{32, 10, 5, 8}, {-8, -1, 0, 0}}, {{32, 10, 5, 8}, {8, 7, 12, 8}}, {{24, 10, 5, 8}, {-1, -1, 0, 0}}};   // the last one traps: unaligned start
    for (const Case& cs : cases) {
        CUtensorMap tm;
        cuuint64_t dims[4] = {D, H, W, F};
        cuuint64_t strides[3] = {D * 2, H * D * 2, static_cast<cuuint64_t>(W) * H * D * 2};
        cuuint32_t estr[4] = {1, 1, 1, 1};
        CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, dv, dims, strides, cs.box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        const int bytes = cs.box[0] * cs.box[1] * cs.box[2] * cs.box[3] * 2;
        printf("box %u x %u x %u x %u at (%d,%d,%d,%d): encode %d, %d bytes\n", cs.box[0], cs.box[1], cs.box[2], cs.box[3], cs.c[0],
               cs.c[1], cs.c[2], cs.c[3], (int)r, bytes);
        if (r != CUDA_SUCCESS) continue;
        cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
        k<<<1, 128, bytes + 256>>>(tm, cs.c[0], cs.c[1], cs.c[2], cs.c[3], bytes, dout);
        cudaError_t e = cudaDeviceSynchronize();
        printf("  run: %s\n", cudaGetErrorString(e));
        if (e != cudaSuccess) return 1;
        std::vector<__half> o(bytes / 2);
        cudaMemcpy(o.data(), dout, bytes, cudaMemcpyDeviceToHost);
        long bad = 0;
        size_t idx = 0;
        for (unsigned f = 0; f < cs.box[3]; ++f)
            for (unsigned x = 0; x < cs.box[2]; ++x)
                for (unsigned y = 0; y < cs.box[1]; ++y)
                    for (unsigned z = 0; z < cs.box[0]; ++z, ++idx) {
                        const int gz = cs.c[0] + z, gy = cs.c[1] + y, gx = cs.c[2] + x, gf = cs.c[3] + f;
                        const bool in = gz >= 0 && gz < D && gy >= 0 && gy < H && gx >= 0 && gx < W && gf >= 0 && gf < F;
                        const float want = in ? __half2float(h[((static_cast<size_t>(gf) * W + gx) * H + gy) * D + gz]) : 0.0f;
                        if (__half2float(o[idx]) != want) ++bad;
                    }
        printf("  mismatches: %ld\n", bad);
    }
    return 0;
}
