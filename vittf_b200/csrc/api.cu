// Error plumbing, device queries and tensor-map construction shared by all entry points.
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

#include <atomic>

static thread_local char g_last_error[1024] = "";
static std::atomic<long long> g_launches{0};

void vittf_count_launches(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
extern "C" int64_t vittf_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }
extern "C" void vittf_launch_count_reset(void) { g_launches.store(0, std::memory_order_relaxed); }

void vittf_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_last_error, sizeof(g_last_error), fmt, ap);
    va_end(ap);
}

extern "C" const char* vittf_last_error(void) { return g_last_error; }
extern "C" int vittf_version(void) { return 100; }

extern "C" int vittf_device_arch(int* out_arch) {
    VITTF_REQUIRE(out_arch, "vittf_device_arch: null pointer");
    int dev = 0, major = 0, minor = 0;
    VITTF_CHECK_CUDA(cudaGetDevice(&dev));
    VITTF_CHECK_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
    VITTF_CHECK_CUDA(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev));
    *out_arch = major * 10 + minor;
    return VITTF_OK;
}

int vittf_num_sms() {
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (cached[dev] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached[dev] = n;
    }
    return cached[dev];
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

int vittf_make_tmap(CUtensorMap* out, const void* base, int elem_bytes, int rank, const uint64_t* dims,
                    const uint64_t* strides_bytes, const uint32_t* box, bool swizzle128) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) {
        vittf_set_error("cuTensorMapEncodeTiled entry point not available (driver too old?)");
        return VITTF_ERR_CUDA;
    }
    VITTF_REQUIRE(elem_bytes == 2 || elem_bytes == 4, "vittf_make_tmap: only 16-bit and fp32 elements supported");
    VITTF_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0, "vittf_make_tmap: base pointer must be 16-byte aligned");
    cuuint64_t gdims[5];
    cuuint64_t gstrides[4];
    cuuint32_t gbox[5];
    cuuint32_t estr[5];
    for (int i = 0; i < rank; ++i) {
        gdims[i] = dims[i];
        gbox[i] = box[i];
        estr[i] = 1;
        if (i > 0) {
            VITTF_REQUIRE(strides_bytes[i - 1] % 16 == 0, "vittf_make_tmap: stride %d (%llu B) not a multiple of 16", i,
                          (unsigned long long)strides_bytes[i - 1]);
            gstrides[i - 1] = strides_bytes[i - 1];
        }
    }
    CUresult r = fn(out, elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32,
                    static_cast<cuuint32_t>(rank), const_cast<void*>(base), gdims,
                    gstrides, gbox, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        vittf_set_error("cuTensorMapEncodeTiled failed with CUresult %d (rank %d, dims %llu x %llu, box %u x %u)", (int)r,
                        rank, (unsigned long long)dims[0], (unsigned long long)dims[1], box[0], box[1]);
        return VITTF_ERR_CUDA;
    }
    return VITTF_OK;
}
