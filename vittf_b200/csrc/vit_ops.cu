// Streaming kernels around the ViT GEMMs: intensity range, folded patch embedding, LayerNorm,
// slice-axis average pooling + 3-axis fp16 merge.  All HBM-bound; written for coalesced 128-bit
// accesses and grids that are multiples of the SM count.
#include <float.h>

#include <stdlib.h>

#include <type_traits>

#include "common.cuh"

namespace {

// ---------------------------------------------------------------------------------------------
// norm_minmax (infer.py:32-34): global min / max of the volume.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void atomic_min_float(float* addr, float v) {
    if (v >= 0.0f) atomicMin(reinterpret_cast<int*>(addr), __float_as_int(v));
    else atomicMax(reinterpret_cast<unsigned int*>(addr), __float_as_uint(v));
}
__device__ __forceinline__ void atomic_max_float(float* addr, float v) {
    if (v >= 0.0f) atomicMax(reinterpret_cast<int*>(addr), __float_as_int(v));
    else atomicMin(reinterpret_cast<unsigned int*>(addr), __float_as_uint(v));
}

template <typename T>
__device__ __forceinline__ float load_as_float(const T* p, int64_t i);
template <>
__device__ __forceinline__ float load_as_float<uint8_t>(const uint8_t* p, int64_t i) { return static_cast<float>(p[i]); }
template <>
__device__ __forceinline__ float load_as_float<__half>(const __half* p, int64_t i) { return __half2float(p[i]); }
template <>
__device__ __forceinline__ float load_as_float<float>(const float* p, int64_t i) { return p[i]; }

__global__ void minmax_init_kernel(float* out2) {
    out2[0] = INFINITY;
    out2[1] = -INFINITY;
}

template <typename T>
__global__ void __launch_bounds__(256) minmax_kernel(const T* __restrict__ vol, int64_t n, float* out2) {
    constexpr int VEC = 16 / sizeof(T);
    float lo = INFINITY, hi = -INFINITY;
    const int64_t nvec = n / VEC;
    const uint4* v4 = reinterpret_cast<const uint4*>(vol);
    for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < nvec;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const uint4 raw = __ldg(v4 + i);
        const T* e = reinterpret_cast<const T*>(&raw);
#pragma unroll
        for (int k = 0; k < VEC; ++k) {
            const float f = load_as_float<T>(e, k);
            lo = fminf(lo, f);
            hi = fmaxf(hi, f);
        }
    }
    if (blockIdx.x == 0)
        for (int64_t i = nvec * VEC + threadIdx.x; i < n; i += blockDim.x) {
            const float f = load_as_float<T>(vol, i);
            lo = fminf(lo, f);
            hi = fmaxf(hi, f);
        }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
        hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    }
    __shared__ float s_lo[8], s_hi[8];
    if ((threadIdx.x & 31) == 0) {
        s_lo[threadIdx.x >> 5] = lo;
        s_hi[threadIdx.x >> 5] = hi;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w) {
            lo = fminf(lo, s_lo[w]);
            hi = fmaxf(hi, s_hi[w]);
        }
        atomic_min_float(out2, lo);
        atomic_max_float(out2 + 1, hi);
    }
}

// ---------------------------------------------------------------------------------------------
// Patch embedding with the input pipeline folded in:
//   slice -> (v-min)/(max-min) -> 3 identical channels -> ImageNet mean/std -> NN resize -> conv p x p
// collapses to  token = sum_taps W'[tap][d] * g(src(tap)) + b'[d]  (+ pos-embed), W' = sum_c W_c / std_c.
// One CTA per (image, patch row); the p*p x f1 gathered grey values live in shared memory.
// ---------------------------------------------------------------------------------------------
struct PatchParams {
    const void* vol;
    int X, Y, Z, axis, s0;
    int a, b;          // source image rows / cols for this axis
    int im0, im1, p, D, f0, f1;
    const float* minmax;
    const float* w;    // (p*p, D)
    const float* bias; // (D)
    const float* pos;  // (1+f0*f1, D), row 0 already holds cls + pos[0]
    float* out;        // (B, 1+f0*f1, D)
};

__device__ __forceinline__ int nearest_src(int dst, int in, int out, float scale) {
    // ATen nearest (legacy) index rule used by F.interpolate(mode='nearest'), infer.py:177
    if (in == out) return dst;
    if (out == 2 * in) return dst >> 1;
    const int s = static_cast<int>(floorf(dst * scale));
    return s < in - 1 ? s : in - 1;
}

template <typename T>
__global__ void __launch_bounds__(256) patch_embed_kernel(PatchParams q) {
    extern __shared__ float s_g[];  // [f1][p*p]
    const int img = blockIdx.y;
    const int py = blockIdx.x;  // == f0 -> CLS row
    const int ntok = 1 + q.f0 * q.f1;
    float* out_img = q.out + static_cast<size_t>(img) * ntok * q.D;
    if (py == q.f0) {
        for (int d = threadIdx.x; d < q.D; d += blockDim.x) out_img[d] = q.pos[d];
        return;
    }
    const int taps = q.p * q.p;
    const float lo = q.minmax[0], hi = q.minmax[1];
    const float inv = 1.0f / (hi - lo);
    const float sc0 = static_cast<float>(q.a) / static_cast<float>(q.im0);
    const float sc1 = static_cast<float>(q.b) / static_cast<float>(q.im1);
    const int s = q.s0 + img;
    const T* vol = static_cast<const T*>(q.vol);
    for (int i = threadIdx.x; i < q.f1 * taps; i += blockDim.x) {
        const int px = i / taps, t = i - px * taps;
        const int u = t / q.p, v = t - u * q.p;
        const int r = nearest_src(py * q.p + u, q.a, q.im0, sc0);
        const int c = nearest_src(px * q.p + v, q.b, q.im1, sc1);
        int64_t idx;
        if (q.axis == 2) idx = (static_cast<int64_t>(r) * q.Y + c) * q.Z + s;        // rows X, cols Y
        else if (q.axis == 1) idx = (static_cast<int64_t>(r) * q.Y + s) * q.Z + c;   // rows X, cols Z
        else idx = (static_cast<int64_t>(s) * q.Y + r) * q.Z + c;                    // rows Y, cols Z
        s_g[i] = (load_as_float<T>(vol, idx) - lo) * inv;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < q.f1 * q.D; i += blockDim.x) {
        const int px = i / q.D, d = i - px * q.D;
        const int tok = 1 + py * q.f1 + px;
        float acc = q.bias[d] + q.pos[static_cast<size_t>(tok) * q.D + d];
        const float* g = s_g + px * taps;
#pragma unroll 8
        for (int t = 0; t < taps; ++t) acc = fmaf(g[t], __ldg(q.w + static_cast<size_t>(t) * q.D + d), acc);
        out_img[static_cast<size_t>(tok) * q.D + d] = acc;
    }
}

// Register-tiled variant for patch 8 (the benchmark backbones): the folded (64 taps x 384 d) weight panel stays in
// shared memory while a persistent CTA walks (image, patch row) items; per item the 64 taps of up to 64 patches are
// gathered once, and every thread accumulates 8 patches x 12 channels (96 FMAs per 2 broadcast LDS.128 + 12 LDS.32).
// The generic kernel above issued one LDS and one LDG per FMA and was 5 % of the ViT step.
constexpr int PE_TAPS = 64, PE_DCH = 384, PE_PX = 64;

template <typename T>
__global__ void __launch_bounds__(256) patch_embed_tiled_kernel(PatchParams q, int n_items) {
    extern __shared__ float s_pe[];
    float* s_w = s_pe;                          // [64 taps][384 d]
    float* s_g = s_pe + PE_TAPS * PE_DCH;       // [64 taps][64 px]
    const int d0 = blockIdx.y * PE_DCH;         // channel chunk of this CTA
    for (int i = threadIdx.x; i < PE_TAPS * PE_DCH; i += 256) {
        const int t = i / PE_DCH, d = i - t * PE_DCH;
        s_w[i] = q.w[static_cast<size_t>(t) * q.D + d0 + d];
    }
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int ntok = 1 + q.f0 * q.f1;
    const float lo = q.minmax[0], inv = 1.0f / (q.minmax[1] - q.minmax[0]);
    const float sc0 = static_cast<float>(q.a) / static_cast<float>(q.im0);
    const float sc1 = static_cast<float>(q.b) / static_cast<float>(q.im1);
    const T* vol = static_cast<const T*>(q.vol);
    float bias[12];
#pragma unroll
    for (int k = 0; k < 12; ++k) bias[k] = q.bias[d0 + tx + 32 * k];
    const int rows_per_img = q.f0 + 1;          // patch rows + one CLS item
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int img = item / rows_per_img, py = item - img * rows_per_img;
        float* out_img = q.out + static_cast<size_t>(img) * ntok * q.D;
        if (py == q.f0) {                       // CLS token: cls + pos[0]
            for (int d = threadIdx.x; d < PE_DCH; d += 256) out_img[d0 + d] = q.pos[d0 + d];
            continue;
        }
        const int s = q.s0 + img;
        for (int px0 = 0; px0 < q.f1; px0 += PE_PX) {
            __syncthreads();                    // previous gather fully consumed (and the weight panel is in place)
            for (int i = threadIdx.x; i < PE_TAPS * PE_PX; i += 256) {
                const int t = i >> 6, px = i & 63;          // consecutive threads: consecutive patches (conflict-free stores)
                const int u = t >> 3, v = t & 7;
                float val = 0.0f;
                if (px0 + px < q.f1) {
                    const int r = nearest_src(py * 8 + u, q.a, q.im0, sc0);
                    const int c = nearest_src((px0 + px) * 8 + v, q.b, q.im1, sc1);
                    int64_t idx;
                    if (q.axis == 2) idx = (static_cast<int64_t>(r) * q.Y + c) * q.Z + s;        // rows X, cols Y
                    else if (q.axis == 1) idx = (static_cast<int64_t>(r) * q.Y + s) * q.Z + c;   // rows X, cols Z
                    else idx = (static_cast<int64_t>(s) * q.Y + r) * q.Z + c;                    // rows Y, cols Z
                    val = (load_as_float<T>(vol, idx) - lo) * inv;
                }
                s_g[t * PE_PX + px] = val;
            }
            __syncthreads();
            float acc[8][12];
            const int tok0 = 1 + py * q.f1 + px0 + ty * 8;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const bool live = px0 + ty * 8 + j < q.f1;
#pragma unroll
                for (int k = 0; k < 12; ++k)
                    acc[j][k] = bias[k] + (live ? __ldg(q.pos + static_cast<size_t>(tok0 + j) * q.D + d0 + tx + 32 * k) : 0.0f);
            }
#pragma unroll 4
            for (int t = 0; t < PE_TAPS; ++t) {
                const float4 g0 = *reinterpret_cast<const float4*>(s_g + t * PE_PX + ty * 8);       // warp-wide broadcast
                const float4 g1 = *reinterpret_cast<const float4*>(s_g + t * PE_PX + ty * 8 + 4);
                const float g[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
                float w[12];
#pragma unroll
                for (int k = 0; k < 12; ++k) w[k] = s_w[t * PE_DCH + tx + 32 * k];                 // conflict-free
#pragma unroll
                for (int j = 0; j < 8; ++j)
#pragma unroll
                    for (int k = 0; k < 12; ++k) acc[j][k] = fmaf(g[j], w[k], acc[j][k]);
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                if (px0 + ty * 8 + j >= q.f1) continue;
                float* dst = out_img + static_cast<size_t>(tok0 + j) * q.D + d0 + tx;
#pragma unroll
                for (int k = 0; k < 12; ++k) dst[32 * k] = acc[j][k];                              // 128 B per warp store
            }
        }
    }
}

// Tensor-core variant for patch 8: per (image, patch row) the 64 x 64 (patch, tap) matrix of gathered grey values times
// the folded (64 taps x 384 d) weight panel is a GEMM.  fp16 operands, fp32 accumulation (mma.sync m16n8k16): the grey
// values g in [0, 1] lose <= 2^-12 absolute (a 16th of a uint8 step), the weights are split hi + lo (K = 128: [g | g] x
// [W_hi ; W_lo]) so they enter at fp32 accuracy.  The panel lives in shared memory for the whole (persistent) kernel as
// [d][k] rows padded to 272 B, the gathered tile as [patch][tap] rows of 144 B -- odd multiples of 16 B, so every ldmatrix
// phase is bank-conflict free.  16 warps, each 16 patches x 96 channels (12 accumulator tiles); the gather of the next
// item overlaps the MMAs of the current one (two tile buffers).  The FMA-pipe kernel above took ~500 us per 64 slices
// (37 % of the fp32 peak, 3.9 % of the ViT step); the output write (403 MB per 64 slices) is the floor here.
constexpr int PM_WK = 136, PM_GK = 72, PM_WARPS = 16, PM_THREADS = 32 * PM_WARPS;

__device__ __forceinline__ void pm_ldmatrix_x4(uint32_t addr, uint32_t (&r)[4]) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void pm_mma(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

template <typename T>
__global__ void __launch_bounds__(PM_THREADS, 1) patch_embed_mma_kernel(PatchParams q, int n_imgs) {
    extern __shared__ __align__(16) uint8_t pm_raw[];
    __half* s_w = reinterpret_cast<__half*>(pm_raw);            // [384 d][136]: k 0..63 = W_hi[tap], 64..127 = W_lo[tap]
    __half* s_g = s_w + PE_DCH * PM_WK;                         // [hi | lo][2 buffers][64 patches][72]
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int d0 = blockIdx.y * PE_DCH;
    // uint8 / fp16 volumes: the A operand is the RAW voxel value (exact in fp16); min-max normalisation moves into the
    // weights (W' / (hi - lo), scaled by 2^e into fp16's normal range, undone in the epilogue) and the bias
    // (b' - lo / (hi - lo) * sum_taps W').  fp32 volumes: A = fp16((v - lo) / (hi - lo)), <= 2^-12 absolute.
    constexpr bool RAW = !std::is_same<T, float>::value;
    const float lo = q.minmax[0], inv = 1.0f / (q.minmax[1] - q.minmax[0]);
    int sexp = 0;
    if (RAW) frexpf(q.minmax[1] - q.minmax[0], &sexp);                     // hi - lo = m 2^e, m in [0.5, 1)
    const float wscale = RAW ? ldexpf(inv, sexp) : 1.0f, oscale = RAW ? ldexpf(1.0f, -sexp) : 1.0f;
    for (int i = tid; i < PE_TAPS * PE_DCH; i += PM_THREADS) {
        const int t = i / PE_DCH, n = i - t * PE_DCH;
        const float v = q.w[static_cast<size_t>(t) * q.D + d0 + n] * wscale;
        const __half hi = __float2half_rn(v);
        s_w[n * PM_WK + t] = hi;
        s_w[n * PM_WK + PE_TAPS + t] = __float2half_rn(v - __half2float(hi));
    }
    // column sums of the folded weights (fixed summation order: identical in every CTA and launch)
    float* s_cs = reinterpret_cast<float*>(s_g + 4 * PE_PX * PM_GK);
    if (RAW) {
        for (int n = tid; n < PE_DCH; n += PM_THREADS) {
            float cs = 0.0f;
#pragma unroll 16
            for (int tp = 0; tp < PE_TAPS; ++tp) cs += q.w[static_cast<size_t>(tp) * q.D + d0 + n];
            s_cs[n] = cs;
        }
        __syncthreads();
    }
    const int ntok = 1 + q.f0 * q.f1;
    // CLS tokens: cls + pos[0] (row 0 of the pos table)
    for (int i = blockIdx.x * PM_THREADS + tid; i < n_imgs * PE_DCH; i += gridDim.x * PM_THREADS) {
        const int img = i / PE_DCH, d = i - img * PE_DCH;
        q.out[static_cast<size_t>(img) * ntok * q.D + d0 + d] = q.pos[d0 + d];
    }
    const float sc0 = static_cast<float>(q.a) / static_cast<float>(q.im0);
    const float sc1 = static_cast<float>(q.b) / static_cast<float>(q.im1);
    const T* vol = static_cast<const T*>(q.vol);
    const int chunks = (q.f1 + PE_PX - 1) / PE_PX;
    const int n_items = n_imgs * q.f0 * chunks;

    // gather: one warp instruction = the 64 taps of one patch (lane = tap pair), 4 patches per warp
    auto gather = [&](int item, __half* dst) {
        const int pxc = item % chunks, py = (item / chunks) % q.f0, img = item / (chunks * q.f0);
        const int s = q.s0 + img;
        const int u = lane >> 2, v = (lane & 3) * 2;                       // taps (u, v) and (u, v + 1)
        const int r = nearest_src(py * 8 + u, q.a, q.im0, sc0);
#pragma unroll
        for (int j = 0; j < PE_PX / PM_WARPS; ++j) {
            const int pl = wid * (PE_PX / PM_WARPS) + j, px = pxc * PE_PX + pl;
            float g0 = 0.0f, g1 = 0.0f;
            if (px < q.f1) {
                const int c0 = nearest_src(px * 8 + v, q.b, q.im1, sc1), c1 = nearest_src(px * 8 + v + 1, q.b, q.im1, sc1);
                int64_t i0, i1;
                if (q.axis == 2) { i0 = (static_cast<int64_t>(r) * q.Y + c0) * q.Z + s; i1 = (static_cast<int64_t>(r) * q.Y + c1) * q.Z + s; }
                else if (q.axis == 1) { i0 = (static_cast<int64_t>(r) * q.Y + s) * q.Z + c0; i1 = i0 + (c1 - c0); }
                else { i0 = (static_cast<int64_t>(s) * q.Y + r) * q.Z + c0; i1 = i0 + (c1 - c0); }
                g0 = load_as_float<T>(vol, i0);
                g1 = load_as_float<T>(vol, i1);
                if (!RAW) {
                    g0 = (g0 - lo) * inv;
                    g1 = (g1 - lo) * inv;
                }
            }
            const __half2 gh = __floats2half2_rn(g0, g1);
            *reinterpret_cast<__half2*>(dst + pl * PM_GK + 2 * lane) = gh;
            if (!RAW) {                                   // residual tile: fp32 volumes enter as g_hi + g_lo
                const float2 back = __half22float2(gh);
                *reinterpret_cast<__half2*>(dst + 2 * PE_PX * PM_GK + pl * PM_GK + 2 * lane) = __floats2half2_rn(g0 - back.x, g1 - back.y);
            }
        }
    };

    const int mq = wid & 3, nq = wid >> 2;                 // patches [16 mq, +16), channels [96 nq, +96) of the chunk
    const int g = lane >> 2, t = lane & 3, mi = lane >> 3, mr = lane & 7;
    const uint32_t a_lane = ((16 * mq + (mi & 1) * 8 + mr) * PM_GK + (mi >> 1) * 8) * 2;      // + 16 ks cols
    const uint32_t b_lane = ptx::smem_u32(s_w) + ((96 * nq + (mi >> 1) * 8 + mr) * PM_WK + (mi & 1) * 8) * 2;
    float bias[12][2];
#pragma unroll
    for (int j = 0; j < 12; ++j)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const int d = d0 + 96 * nq + 8 * j + 2 * t + e;
            bias[j][e] = q.bias[d] - (RAW ? lo * inv * s_cs[d - d0] : 0.0f);
        }
    int buf = 0;
    if (blockIdx.x < n_items) gather(blockIdx.x, s_g);
    __syncthreads();
    const float iscale = 1.0f / oscale;                    // (powers of two: exact)
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        // the accumulators start as (bias + pos-embed) 2^e: these loads land straight in the accumulator registers and
        // are in flight while the next item is gathered
        const int pxc = item % chunks, py = (item / chunks) % q.f0, img = item / (chunks * q.f0);
        float acc[12][4];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int px = pxc * PE_PX + 16 * mq + g + 8 * h;
            const size_t tok = 1 + static_cast<size_t>(py) * q.f1 + (px < q.f1 ? px : 0);
            const float* pos = q.pos + tok * q.D + d0 + 96 * nq + 2 * t;
#pragma unroll
            for (int j = 0; j < 12; ++j) {
                const float2 pe = __ldg(reinterpret_cast<const float2*>(pos + 8 * j));
                acc[j][2 * h] = (pe.x + bias[j][0]) * iscale;
                acc[j][2 * h + 1] = (pe.y + bias[j][1]) * iscale;
            }
        }
        if (item + gridDim.x < n_items) gather(item + gridDim.x, s_g + (buf ^ 1) * PE_PX * PM_GK);
        const uint32_t ga = ptx::smem_u32(s_g + buf * PE_PX * PM_GK) + a_lane;
        // k-steps 0..3: g x W_hi, 4..7: g x W_lo, (fp32 volumes) 8..11: g_lo x W_hi
#pragma unroll
        for (int kk = 0; kk < (RAW ? 8 : 12); ++kk) {
            uint32_t a0[4];
            pm_ldmatrix_x4(ga + (kk >= 8 ? 2 * PE_PX * PM_GK * 2 : 0) + (kk & 3) * 32, a0);
#pragma unroll
            for (int jp = 0; jp < 6; ++jp) {
                uint32_t bf[4];
                pm_ldmatrix_x4(b_lane + (16 * jp * PM_WK + 16 * (kk >= 8 ? kk - 8 : kk)) * 2, bf);
                pm_mma(acc[2 * jp], a0, bf[0], bf[1]);
                pm_mma(acc[2 * jp + 1], a0, bf[2], bf[3]);
            }
        }
        // epilogue: undo the 2^e scale, fp32 tokens (float2 = a full 32-byte sector per row and quad)
        float* out_img = q.out + static_cast<size_t>(img) * ntok * q.D;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int px = pxc * PE_PX + 16 * mq + g + 8 * h;
            if (px >= q.f1) continue;
            const size_t tok = 1 + static_cast<size_t>(py) * q.f1 + px;
            float* dst = out_img + tok * q.D + d0 + 96 * nq + 2 * t;
#pragma unroll
            for (int j = 0; j < 12; ++j)
                *reinterpret_cast<float2*>(dst + 8 * j) = make_float2(acc[j][2 * h] * oscale, acc[j][2 * h + 1] * oscale);
        }
        __syncthreads();
        buf ^= 1;
    }
}

// ---------------------------------------------------------------------------------------------
// LayerNorm (eps 1e-6) fp32 -> bf16, one warp per row, row held in registers.
// ---------------------------------------------------------------------------------------------
template <int V4>  // float4 per lane: D = 128 * V4
__global__ void __launch_bounds__(256) layernorm_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                        const float* __restrict__ b, __nv_bfloat16* __restrict__ y,
                                                        int64_t rows) {
    const int lane = threadIdx.x & 31;
    const int64_t row = blockIdx.x * static_cast<int64_t>(blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows) return;
    constexpr int D = 128 * V4;
    const float4* xr = reinterpret_cast<const float4*>(x + row * D);
    float4 v[V4];
    float sum = 0.0f;
#pragma unroll
    for (int i = 0; i < V4; ++i) {
        v[i] = xr[lane + 32 * i];
        sum += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float mean = sum * (1.0f / D);
    float var = 0.0f;
#pragma unroll
    for (int i = 0; i < V4; ++i) {
        const float a = v[i].x - mean, c = v[i].y - mean, e = v[i].z - mean, f = v[i].w - mean;
        var += (a * a + c * c) + (e * e + f * f);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) var += __shfl_xor_sync(0xffffffffu, var, o);
    const float rstd = rsqrtf(var * (1.0f / D) + 1e-6f);
    uint2* yr = reinterpret_cast<uint2*>(y + row * D);
#pragma unroll
    for (int i = 0; i < V4; ++i) {
        const float4 ww = __ldg(reinterpret_cast<const float4*>(w) + lane + 32 * i);
        const float4 bb = __ldg(reinterpret_cast<const float4*>(b) + lane + 32 * i);
        const float o0 = (v[i].x - mean) * rstd * ww.x + bb.x;
        const float o1 = (v[i].y - mean) * rstd * ww.y + bb.y;
        const float o2 = (v[i].z - mean) * rstd * ww.z + bb.z;
        const float o3 = (v[i].w - mean) * rstd * ww.w + bb.w;
        yr[lane + 32 * i] = make_uint2(ptx::pack_bf16x2(o0, o1), ptx::pack_bf16x2(o2, o3));
    }
}

// ---------------------------------------------------------------------------------------------
// Entry into the LayerNorm-folded GEMM chain (gemm.cu): the row-major fp32 tokens of the patch embedding become the
// row-tiled residual stream xt[m_pad/32][D/4][32][4], its raw bf16 copy and the per-row (sum, sum of squares).  One CTA
// per group of 32 rows, 128 columns at a time: rows are read and the bf16 copy written with a warp per row (512 / 256
// contiguous bytes per instruction), the fp32 values cross a padded shared-memory tile and leave as the 16 KB contiguous
// block that 32 rows x 128 columns occupy in the tiled layout.  Runs once per forward (the blocks' LayerNorms themselves
// never run as a pass).
// ---------------------------------------------------------------------------------------------
template <int V4>
__global__ void __launch_bounds__(256) ln_prepare_kernel(const float* __restrict__ x, float* __restrict__ xt,
                                                         __nv_bfloat16* __restrict__ xb, float2* __restrict__ stats,
                                                         int64_t rows) {
    constexpr int D = 128 * V4;
    __shared__ float4 tile[32][33];                       // [row][column group], pitch 33: the transposed reads are conflict-free
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t row0 = static_cast<int64_t>(blockIdx.x) * 32;
    float4* xt_grp = reinterpret_cast<float4*>(xt) + row0 * (D / 4);       // this group's block of the tiled stream
    float sum[4] = {0.0f, 0.0f, 0.0f, 0.0f}, sq[4] = {0.0f, 0.0f, 0.0f, 0.0f};
    for (int ch = 0; ch < V4; ++ch) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int r = warp * 4 + k;
            const int64_t row = row0 + r;
            float4 v = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
            if (row < rows) {
                v = reinterpret_cast<const float4*>(x + row * D)[ch * 32 + lane];
                reinterpret_cast<uint2*>(xb + row * D)[ch * 32 + lane] = make_uint2(ptx::pack_bf16x2(v.x, v.y), ptx::pack_bf16x2(v.z, v.w));
            }
            sum[k] += (v.x + v.y) + (v.z + v.w);
            sq[k] += (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w);
            tile[r][lane] = v;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int e = threadIdx.x + 256 * k;          // element (column group j, row r) of the chunk, r fastest
            xt_grp[(ch * 32) * 32 + e] = tile[e & 31][e >> 5];
        }
        __syncthreads();
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            sum[k] += __shfl_xor_sync(0xffffffffu, sum[k], o);
            sq[k] += __shfl_xor_sync(0xffffffffu, sq[k], o);
        }
        const int64_t row = row0 + warp * 4 + k;
        if (lane < VITTF_LN_SLOTS) stats[row * VITTF_LN_SLOTS + lane] = lane == 0 ? make_float2(sum[k], sq[k]) : make_float2(0.0f, 0.0f);
    }
}

// ---------------------------------------------------------------------------------------------
// AdaptiveAvgPool3d along the slice axis + permute to (D, fX, fY, fZ) + optional fp16 running sum.
// k: (S, T = f0*f1, D) fp16, D fastest.  One CTA transposes a 32(token) x 64(d) tile through smem.
// ---------------------------------------------------------------------------------------------
struct PoolParams {
    const __half* k;
    __half* out;
    int S, T, D, n_out;
    int slice0, n_local, o0;   // k holds global slices [slice0, slice0+n_local); blockIdx.z + o0 = output slab
    int out_o0;                // slab index of the output array's first slab (0: full-size array, o0: compact block of this rank)
    int f1;
    int64_t sd, s0, s1, so;  // output strides (elements) of d, i0, i1, o
    int accumulate;
};

__global__ void __launch_bounds__(256) pool_axis_kernel(PoolParams q) {
    __shared__ float tile[32][65];
    const int t0 = blockIdx.x * 32, d0 = blockIdx.y * 64, o = blockIdx.z + q.o0;
    // AdaptiveAvgPool window [floor(o*S/n), ceil((o+1)*S/n)) in global slice indices, then local
    const int w0 = static_cast<int>((static_cast<int64_t>(o) * q.S) / q.n_out) - q.slice0;
    const int w1 = static_cast<int>((static_cast<int64_t>(o + 1) * q.S + q.n_out - 1) / q.n_out) - q.slice0;
    // load: thread -> (token row = tid/8, 8 halves at d = (tid%8)*8)
    {
        const int tr = threadIdx.x >> 3, dc = (threadIdx.x & 7) * 8;
        float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        if (t0 + tr < q.T && d0 + dc < q.D) {
            // four slices of the window in flight per thread (the loop is latency-bound otherwise); the fp32 sum keeps the
            // slice order
            const __half* src = q.k + (static_cast<size_t>(w0) * q.T + (t0 + tr)) * q.D + d0 + dc;
            const size_t sstride = static_cast<size_t>(q.T) * q.D;
            for (int s = w0; s < w1; s += 4, src += 4 * sstride) {
                uint4 raw[4];
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    raw[u] = s + u < w1 ? __ldg(reinterpret_cast<const uint4*>(src + u * sstride)) : make_uint4(0, 0, 0, 0);
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    if (s + u >= w1) break;
                    const __half2* h = reinterpret_cast<const __half2*>(&raw[u]);
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float2 f = __half22float2(h[i]);
                        acc[2 * i] += f.x;
                        acc[2 * i + 1] += f.y;
                    }
                }
            }
        }
        const float cnt = static_cast<float>(w1 - w0);
#pragma unroll
        for (int i = 0; i < 8; ++i) tile[tr][dc + i] = acc[i] / cnt;
    }
    __syncthreads();
    // store: thread -> (d = tid/32 + 8*r, token = tid%32)
    const int tl = threadIdx.x & 31;
    const int t = t0 + tl;
    if (t >= q.T) return;
    const int i0 = t / q.f1, i1 = t - i0 * q.f1;
    const int64_t base = i0 * q.s0 + i1 * q.s1 + (o - q.out_o0) * q.so;
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        const int dl = (threadIdx.x >> 5) + 8 * r;
        if (d0 + dl >= q.D) break;
        __half* dst = q.out + (d0 + dl) * q.sd + base;
        const __half pooled = __float2half_rn(tile[tl][dl]);
        *dst = q.accumulate ? __hadd(*dst, pooled) : pooled;
    }
}


// z pass (slices along Z): the pooled slab index o is the FASTEST output dimension, so the kernel above writes one
// 2-byte element per 128-byte line (measured 0.4 TB/s against 4.1 TB/s for the y / x passes).  Here one CTA pools 8
// consecutive slabs of its 32 x 64 (token, d) tile into shared memory and writes 16 contiguous bytes per (d, token).
constexpr int POOLZ_OB = 8;
__global__ void __launch_bounds__(256) pool_axis_z_kernel(PoolParams q, int o_end) {
    __shared__ __align__(16) __half tile[POOLZ_OB][32][72];
    const int t0 = blockIdx.x * 32, d0 = blockIdx.y * 64, ob = blockIdx.z * POOLZ_OB + q.o0;
    const int tr = threadIdx.x >> 3, dc = (threadIdx.x & 7) * 8;
    const bool live = t0 + tr < q.T && d0 + dc < q.D;
    const size_t sstride = static_cast<size_t>(q.T) * q.D;
#pragma unroll 1
    for (int oo = 0; oo < POOLZ_OB; ++oo) {
        const int o = ob + oo;
        float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        float cnt = 1.0f;
        if (live && o < o_end) {
            const int w0 = static_cast<int>((static_cast<int64_t>(o) * q.S) / q.n_out) - q.slice0;
            const int w1 = static_cast<int>((static_cast<int64_t>(o + 1) * q.S + q.n_out - 1) / q.n_out) - q.slice0;
            cnt = static_cast<float>(w1 - w0);
            const __half* src = q.k + (static_cast<size_t>(w0) * q.T + (t0 + tr)) * q.D + d0 + dc;
            for (int s = w0; s < w1; s += 4, src += 4 * sstride) {
                uint4 raw[4];
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    raw[u] = s + u < w1 ? __ldg(reinterpret_cast<const uint4*>(src + u * sstride)) : make_uint4(0, 0, 0, 0);
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    if (s + u >= w1) break;
                    const __half2* h = reinterpret_cast<const __half2*>(&raw[u]);
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float2 f = __half22float2(h[i]);
                        acc[2 * i] += f.x;
                        acc[2 * i + 1] += f.y;
                    }
                }
            }
        }
        __half2 hv[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) hv[i] = __floats2half2_rn(acc[2 * i] / cnt, acc[2 * i + 1] / cnt);
        *reinterpret_cast<uint4*>(&tile[oo][tr][dc]) = *reinterpret_cast<const uint4*>(hv);
    }
    __syncthreads();
    const bool full = ob + POOLZ_OB <= o_end;
#pragma unroll 1
    for (int it = 0; it < 8; ++it) {
        const int id = it * 256 + threadIdx.x, tl = id & 31, dl = id >> 5;
        const int t = t0 + tl;
        if (t >= q.T || d0 + dl >= q.D) continue;
        const int i0 = t / q.f1, i1 = t - i0 * q.f1;
        __half* dst = q.out + (d0 + dl) * q.sd + i0 * q.s0 + i1 * q.s1 + (ob - q.out_o0);      // so == 1
        __half v[POOLZ_OB];
#pragma unroll
        for (int oo = 0; oo < POOLZ_OB; ++oo) v[oo] = tile[oo][tl][dl];
        if (full) {
            uint4 w = *reinterpret_cast<const uint4*>(v);
            if (q.accumulate) {
                const uint4 old = *reinterpret_cast<const uint4*>(dst);
                const __half2* a = reinterpret_cast<const __half2*>(&old);
                __half2* b = reinterpret_cast<__half2*>(&w);
#pragma unroll
                for (int i = 0; i < 4; ++i) b[i] = __hadd2(a[i], b[i]);
            }
            *reinterpret_cast<uint4*>(dst) = w;
        } else {
            for (int oo = 0; oo < POOLZ_OB && ob + oo < o_end; ++oo) dst[oo] = q.accumulate ? __hadd(dst[oo], v[oo]) : v[oo];
        }
    }
}

// fp16 running sum of per-axis volumes (infer.py:332): out = fp16(out + in), 128-bit vectorised
__global__ void __launch_bounds__(256) accumulate_f16_kernel(__half* __restrict__ out, const __half* __restrict__ in, int64_t n) {
    const int64_t nvec = n / 8;
    uint4* o4 = reinterpret_cast<uint4*>(out);
    const uint4* i4 = reinterpret_cast<const uint4*>(in);
    for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < nvec;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        uint4 a = o4[i];
        const uint4 b = __ldg(i4 + i);
        __half2* ah = reinterpret_cast<__half2*>(&a);
        const __half2* bh = reinterpret_cast<const __half2*>(&b);
#pragma unroll
        for (int k = 0; k < 4; ++k) ah[k] = __hadd2(ah[k], bh[k]);
        o4[i] = a;
    }
    if (blockIdx.x == 0)
        for (int64_t i = nvec * 8 + threadIdx.x; i < n; i += blockDim.x) out[i] = __hadd(out[i], in[i]);
}

// Multi-GPU merge: the all-gather of the ranks' compact per-axis blocks lands rank-major, (world, D, e0, e1, e2)
// with the slab axis extent divided by `world`; this kernel un-permutes it into (D, fX, fY, fZ) on the fly and either
// assigns (first axis: fp16(0 + z) = z) or adds in fp16 (infer.py:332), VEC halves along z per thread.
template <int VEC>
__global__ void __launch_bounds__(256) accumulate_gathered_kernel(__half* __restrict__ out, const __half* __restrict__ st, int world, int D,
                                                                  int fX, int fY, int fZ, int axis, int accumulate) {
    const int64_t n = static_cast<int64_t>(D) * fX * fY * fZ / VEC;
    const int zv = fZ / VEC;
    const int nloc = (axis == 2 ? fZ : axis == 1 ? fY : fX) / world;
    for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n; i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int z = static_cast<int>(i % zv) * VEC;
        int64_t t = i / zv;
        const int y = static_cast<int>(t % fY);
        t /= fY;
        const int x = static_cast<int>(t % fX);
        const int d = static_cast<int>(t / fX);
        int64_t src;
        if (axis == 2) { const int r = z / nloc; src = (((static_cast<int64_t>(r) * D + d) * fX + x) * fY + y) * nloc + (z - r * nloc); }
        else if (axis == 1) { const int r = y / nloc; src = (((static_cast<int64_t>(r) * D + d) * fX + x) * nloc + (y - r * nloc)) * fZ + z; }
        else { const int r = x / nloc; src = (((static_cast<int64_t>(r) * D + d) * nloc + (x - r * nloc)) * fY + y) * fZ + z; }
        if (VEC == 8) {
            uint4 b = __ldg(reinterpret_cast<const uint4*>(st + src));
            if (accumulate) {
                const uint4 a = *reinterpret_cast<const uint4*>(out + i * 8);
                const __half2* ah = reinterpret_cast<const __half2*>(&a);
                __half2* bh = reinterpret_cast<__half2*>(&b);
#pragma unroll
                for (int k = 0; k < 4; ++k) bh[k] = __hadd2(ah[k], bh[k]);
            }
            *reinterpret_cast<uint4*>(out + i * 8) = b;
        } else {
            const __half b = st[src];
            out[i] = accumulate ? __hadd(out[i], b) : b;
        }
    }
}

}  // namespace

extern "C" int vittf_accumulate_gathered_f16(void* out_f16, const void* staging_f16, int world, int D, int fX, int fY, int fZ, int axis,
                                             int accumulate, void* stream) {
    VITTF_REQUIRE(out_f16 && staging_f16, "vittf_accumulate_gathered_f16: null pointer");
    VITTF_REQUIRE(world > 0 && D > 0 && fX > 0 && fY > 0 && fZ > 0 && axis >= 0 && axis <= 2, "vittf_accumulate_gathered_f16: bad sizes");
    const int ext = axis == 2 ? fZ : axis == 1 ? fY : fX;
    VITTF_REQUIRE(ext % world == 0, "vittf_accumulate_gathered_f16: %d slabs do not divide over %d ranks", ext, world);
    const int nloc = ext / world;
    const bool vec = fZ % 8 == 0 && (axis != 2 || nloc % 8 == 0) &&
                     ((reinterpret_cast<uintptr_t>(out_f16) | reinterpret_cast<uintptr_t>(staging_f16)) & 15) == 0;
    const int64_t n = static_cast<int64_t>(D) * fX * fY * fZ / (vec ? 8 : 1);
    int64_t blocks = ceil_div_ll(n, 256);
    const int64_t cap = static_cast<int64_t>(vittf_num_sms()) * 8;
    if (blocks > cap) blocks = cap;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (vec)
        accumulate_gathered_kernel<8><<<static_cast<unsigned>(blocks), 256, 0, s>>>(static_cast<__half*>(out_f16), static_cast<const __half*>(staging_f16),
                                                                                    world, D, fX, fY, fZ, axis, accumulate);
    else
        accumulate_gathered_kernel<1><<<static_cast<unsigned>(blocks), 256, 0, s>>>(static_cast<__half*>(out_f16), static_cast<const __half*>(staging_f16),
                                                                                    world, D, fX, fY, fZ, axis, accumulate);
    VITTF_CHECK_CUDA(cudaGetLastError());
    vittf_count_launches(1);
    return VITTF_OK;
}

extern "C" int vittf_accumulate_f16(void* out_f16, const void* in_f16, int64_t n, void* stream) {
    VITTF_REQUIRE(out_f16 && in_f16 && n > 0, "vittf_accumulate_f16: bad arguments");
    VITTF_REQUIRE(((reinterpret_cast<uintptr_t>(out_f16) | reinterpret_cast<uintptr_t>(in_f16)) & 15) == 0,
                  "vittf_accumulate_f16: buffers must be 16-byte aligned");
    int64_t blocks = ceil_div_ll(n / 8 + 1, 256);
    const int64_t cap = static_cast<int64_t>(vittf_num_sms()) * 8;
    if (blocks > cap) blocks = cap;
    accumulate_f16_kernel<<<static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<__half*>(out_f16), static_cast<const __half*>(in_f16), n);
    VITTF_CHECK_CUDA(cudaGetLastError());
    vittf_count_launches(1);
    return VITTF_OK;
}

extern "C" int vittf_minmax(const void* vol, int64_t n, int dtype, float* out2, void* stream) {
    VITTF_REQUIRE(vol && out2 && n > 0, "vittf_minmax: bad arguments");
    VITTF_REQUIRE((reinterpret_cast<uintptr_t>(vol) & 15) == 0, "vittf_minmax: volume must be 16-byte aligned");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    minmax_init_kernel<<<1, 1, 0, s>>>(out2);
    const int grid = vittf_num_sms() * 8;
    switch (dtype) {
        case VITTF_U8: minmax_kernel<uint8_t><<<grid, 256, 0, s>>>(static_cast<const uint8_t*>(vol), n, out2); break;
        case VITTF_F16: minmax_kernel<__half><<<grid, 256, 0, s>>>(static_cast<const __half*>(vol), n, out2); break;
        case VITTF_F32: minmax_kernel<float><<<grid, 256, 0, s>>>(static_cast<const float*>(vol), n, out2); break;
        default: VITTF_REQUIRE(false, "vittf_minmax: unsupported dtype %d", dtype);
    }
    VITTF_CHECK_CUDA(cudaGetLastError());
    vittf_count_launches(2);
    return VITTF_OK;
}

extern "C" int vittf_patch_embed(const void* vol, int vol_dtype, int X, int Y, int Z, int axis, int s0, int s1, int im0,
                                 int im1, int patch, int D, const float* minmax2, const float* patch_w,
                                 const float* patch_b, const float* pos_embed, float* out_tokens, void* stream) {
    VITTF_REQUIRE(vol && minmax2 && patch_w && patch_b && pos_embed && out_tokens, "vittf_patch_embed: null pointer");
    VITTF_REQUIRE(axis >= 0 && axis <= 2, "vittf_patch_embed: axis must be 0 (x), 1 (y) or 2 (z)");
    const int dims[3] = {X, Y, Z};
    VITTF_REQUIRE(s0 >= 0 && s1 > s0 && s1 <= dims[axis], "vittf_patch_embed: slice range [%d,%d) outside axis of %d", s0,
                  s1, dims[axis]);
    VITTF_REQUIRE(im0 % patch == 0 && im1 % patch == 0 && im0 > 0 && im1 > 0, "vittf_patch_embed: image %dx%d not a multiple of patch %d",
                  im0, im1, patch);
    PatchParams q;
    q.vol = vol; q.X = X; q.Y = Y; q.Z = Z; q.axis = axis; q.s0 = s0;
    q.a = axis == 0 ? Y : X;
    q.b = axis == 2 ? Y : Z;
    q.im0 = im0; q.im1 = im1; q.p = patch; q.D = D; q.f0 = im0 / patch; q.f1 = im1 / patch;
    q.minmax = minmax2; q.w = patch_w; q.bias = patch_b; q.pos = pos_embed; q.out = out_tokens;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    static const bool pe_no_mma = getenv("VITTF_PE_NO_MMA") != nullptr;     // A/B switch: FMA-pipe tiled kernel
    if (patch == 8 && D % PE_DCH == 0 && !pe_no_mma) {   // tensor-core persistent kernel (patch-8 backbones)
        const size_t smem_m = (static_cast<size_t>(PE_DCH) * PM_WK + 4 * PE_PX * PM_GK) * sizeof(__half) + PE_DCH * sizeof(float);
        const int n_imgs = s1 - s0;
        const int n_items = n_imgs * q.f0 * ((q.f1 + PE_PX - 1) / PE_PX);
        dim3 grid_m(n_items < vittf_num_sms() ? n_items : vittf_num_sms(), D / PE_DCH);
#define LAUNCH_PEM(T)                                                                                                   \
    do {                                                                                                                \
        static PerDeviceMemo configured;                                                                                 \
        if (!configured.cur()) {                                                                                              \
            VITTF_CHECK_CUDA(cudaFuncSetAttribute(patch_embed_mma_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                                  static_cast<int>(smem_m)));                                          \
            configured.cur() = 1;                                                                                          \
        }                                                                                                               \
        patch_embed_mma_kernel<T><<<grid_m, PM_THREADS, smem_m, s>>>(q, n_imgs);                                               \
    } while (0)
        switch (vol_dtype) {
            case VITTF_U8: LAUNCH_PEM(uint8_t); break;
            case VITTF_F16: LAUNCH_PEM(__half); break;
            case VITTF_F32: LAUNCH_PEM(float); break;
            default: VITTF_REQUIRE(false, "vittf_patch_embed: unsupported volume dtype %d", vol_dtype);
        }
#undef LAUNCH_PEM
        VITTF_CHECK_CUDA(cudaGetLastError());
        vittf_count_launches(1);
        return VITTF_OK;
    }
    if (patch == 8 && D % PE_DCH == 0) {        // register-tiled persistent kernel (patch-8 backbones)
        const size_t smem_t = static_cast<size_t>(PE_TAPS) * (PE_DCH + PE_PX) * sizeof(float);
        const int n_items = (s1 - s0) * (q.f0 + 1);
        dim3 grid_t(n_items < vittf_num_sms() ? n_items : vittf_num_sms(), D / PE_DCH);
#define LAUNCH_PET(T)                                                                                                   \
    do {                                                                                                                \
        static PerDeviceMemo configured;                                                                                 \
        if (!configured.cur()) {                                                                                              \
            VITTF_CHECK_CUDA(cudaFuncSetAttribute(patch_embed_tiled_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                                  static_cast<int>(smem_t)));                                          \
            configured.cur() = 1;                                                                                          \
        }                                                                                                               \
        patch_embed_tiled_kernel<T><<<grid_t, 256, smem_t, s>>>(q, n_items);                                            \
    } while (0)
        switch (vol_dtype) {
            case VITTF_U8: LAUNCH_PET(uint8_t); break;
            case VITTF_F16: LAUNCH_PET(__half); break;
            case VITTF_F32: LAUNCH_PET(float); break;
            default: VITTF_REQUIRE(false, "vittf_patch_embed: unsupported volume dtype %d", vol_dtype);
        }
#undef LAUNCH_PET
        VITTF_CHECK_CUDA(cudaGetLastError());
        vittf_count_launches(1);
        return VITTF_OK;
    }
    const size_t smem = static_cast<size_t>(q.f1) * patch * patch * sizeof(float);
    VITTF_REQUIRE(smem <= 200 * 1024, "vittf_patch_embed: image row too wide (%zu B of shared memory)", smem);
    dim3 grid(q.f0 + 1, s1 - s0);
#define LAUNCH_PE(T)                                                                                          \
    do {                                                                                                      \
        if (smem > 48 * 1024)                                                                                 \
            VITTF_CHECK_CUDA(cudaFuncSetAttribute(patch_embed_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                                  static_cast<int>(smem)));                                  \
        patch_embed_kernel<T><<<grid, 256, smem, s>>>(q);                                                     \
    } while (0)
    switch (vol_dtype) {
        case VITTF_U8: LAUNCH_PE(uint8_t); break;
        case VITTF_F16: LAUNCH_PE(__half); break;
        case VITTF_F32: LAUNCH_PE(float); break;
        default: VITTF_REQUIRE(false, "vittf_patch_embed: unsupported volume dtype %d", vol_dtype);
    }
#undef LAUNCH_PE
    VITTF_CHECK_CUDA(cudaGetLastError());
    vittf_count_launches(1);
    return VITTF_OK;
}

extern "C" int vittf_layernorm(const float* x, const float* w, const float* b, void* y_bf16, int64_t rows, int D,
                               void* stream) {
    VITTF_REQUIRE(x && w && b && y_bf16 && rows > 0, "vittf_layernorm: bad arguments");
    VITTF_REQUIRE(D % 128 == 0 && D >= 128 && D <= 1024, "vittf_layernorm: D=%d must be a multiple of 128 in [128,1024]", D);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const unsigned grid = static_cast<unsigned>(ceil_div_ll(rows, 8));
    __nv_bfloat16* y = static_cast<__nv_bfloat16*>(y_bf16);
    switch (D / 128) {
        case 1: layernorm_kernel<1><<<grid, 256, 0, s>>>(x, w, b, y, rows); break;
        case 2: layernorm_kernel<2><<<grid, 256, 0, s>>>(x, w, b, y, rows); break;
        case 3: layernorm_kernel<3><<<grid, 256, 0, s>>>(x, w, b, y, rows); break;
        case 4: layernorm_kernel<4><<<grid, 256, 0, s>>>(x, w, b, y, rows); break;
        case 5: layernorm_kernel<5><<<grid, 256, 0, s>>>(x, w, b, y, rows); break;
        case 6: layernorm_kernel<6><<<grid, 256, 0, s>>>(x, w, b, y, rows); break;
        case 7: layernorm_kernel<7><<<grid, 256, 0, s>>>(x, w, b, y, rows); break;
        default: layernorm_kernel<8><<<grid, 256, 0, s>>>(x, w, b, y, rows); break;
    }
    VITTF_CHECK_CUDA(cudaGetLastError());
    vittf_count_launches(1);
    return VITTF_OK;
}

extern "C" int vittf_ln_prepare(const float* x, float* xt, void* xb_bf16, float* stats, int64_t rows, int64_t m_pad, int D,
                                void* stream) {
    VITTF_REQUIRE(x && xt && xb_bf16 && stats && rows > 0, "vittf_ln_prepare: bad arguments");
    VITTF_REQUIRE(D % 128 == 0 && D >= 128 && D <= 1024, "vittf_ln_prepare: D=%d must be a multiple of 128 in [128,1024]", D);
    VITTF_REQUIRE(m_pad >= rows && m_pad % 256 == 0, "vittf_ln_prepare: m_pad must be rows rounded up to a multiple of 256");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const unsigned grid = static_cast<unsigned>(ceil_div_ll(rows, 32));          // whole 32-row groups: padding rows are written as zeros
    __nv_bfloat16* y = static_cast<__nv_bfloat16*>(xb_bf16);
    float2* st = reinterpret_cast<float2*>(stats);
    switch (D / 128) {
        case 1: ln_prepare_kernel<1><<<grid, 256, 0, s>>>(x, xt, y, st, rows); break;
        case 2: ln_prepare_kernel<2><<<grid, 256, 0, s>>>(x, xt, y, st, rows); break;
        case 3: ln_prepare_kernel<3><<<grid, 256, 0, s>>>(x, xt, y, st, rows); break;
        case 4: ln_prepare_kernel<4><<<grid, 256, 0, s>>>(x, xt, y, st, rows); break;
        case 5: ln_prepare_kernel<5><<<grid, 256, 0, s>>>(x, xt, y, st, rows); break;
        case 6: ln_prepare_kernel<6><<<grid, 256, 0, s>>>(x, xt, y, st, rows); break;
        case 7: ln_prepare_kernel<7><<<grid, 256, 0, s>>>(x, xt, y, st, rows); break;
        default: ln_prepare_kernel<8><<<grid, 256, 0, s>>>(x, xt, y, st, rows); break;
    }
    VITTF_CHECK_CUDA(cudaGetLastError());
    vittf_count_launches(1);
    return VITTF_OK;
}

extern "C" int vittf_pool_axis(const void* k_f16, int S, int slice0, int n_local, int f0, int f1, int D, int axis,
                               int n_out, int o0, int o1, void* out_f16, int accumulate, int compact, void* stream) {
    VITTF_REQUIRE(k_f16 && out_f16, "vittf_pool_axis: null pointer");
    // n_out > S is legal (AdaptiveAvgPool3d replicates slices: every window [floor(o*S/n), ceil((o+1)*S/n)) is non-empty)
    VITTF_REQUIRE(S > 0 && f0 > 0 && f1 > 0 && D > 0 && n_out > 0, "vittf_pool_axis: bad sizes");
    VITTF_REQUIRE(o0 >= 0 && o1 > o0 && o1 <= n_out, "vittf_pool_axis: slab range [%d,%d) outside [0,%d)", o0, o1, n_out);
    {
        const int need0 = static_cast<int>((static_cast<int64_t>(o0) * S) / n_out);
        const int need1 = static_cast<int>((static_cast<int64_t>(o1) * S + n_out - 1) / n_out);
        VITTF_REQUIRE(slice0 <= need0 && slice0 + n_local >= need1,
                      "vittf_pool_axis: slabs [%d,%d) need slices [%d,%d) but k holds [%d,%d)", o0, o1, need0, need1, slice0,
                      slice0 + n_local);
    }
    VITTF_REQUIRE(D % 8 == 0, "vittf_pool_axis: D must be a multiple of 8");
    VITTF_REQUIRE(axis >= 0 && axis <= 2, "vittf_pool_axis: axis must be 0, 1 or 2");
    PoolParams q;
    q.k = static_cast<const __half*>(k_f16);
    q.out = static_cast<__half*>(out_f16);
    q.S = S; q.T = f0 * f1; q.D = D; q.n_out = n_out; q.f1 = f1; q.accumulate = accumulate;
    q.slice0 = slice0; q.n_local = n_local; q.o0 = o0;
    // output (D, A, B, C) contiguous; which of A,B,C are i0 / i1 / o depends on the slicing axis.  compact: the array holds
    // only the slabs [o0, o1) of this rank (the block the all-gather of the multi-GPU path sends)
    q.out_o0 = compact ? o0 : 0;
    const int ext = compact ? o1 - o0 : n_out;
    int64_t A, B, C;
    if (axis == 2) { A = f0; B = f1; C = ext; q.s0 = B * C; q.s1 = C; q.so = 1; }          // (D, fX, fY, o)
    else if (axis == 1) { A = f0; B = ext; C = f1; q.s0 = B * C; q.so = C; q.s1 = 1; }     // (D, fX, o, fZ)
    else { A = ext; B = f0; C = f1; q.so = B * C; q.s0 = C; q.s1 = 1; }                    // (D, o, fY, fZ)
    q.sd = A * B * C;
    dim3 grid(ceil_div(q.T, 32), ceil_div(D, 64), o1 - o0);
    if (axis == 2 && ext % POOLZ_OB == 0 && (o0 - q.out_o0) % POOLZ_OB == 0 && (reinterpret_cast<uintptr_t>(out_f16) & 15) == 0) {
        grid.z = ceil_div(o1 - o0, POOLZ_OB);
        pool_axis_z_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(q, o1);
    } else {
        pool_axis_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(q);
    }
    VITTF_CHECK_CUDA(cudaGetLastError());
    vittf_count_launches(1);
    return VITTF_OK;
}
