// Streaming kernels around the ViT GEMMs: intensity range, folded patch embedding, LayerNorm,
// slice-axis average pooling + 3-axis fp16 merge.  All HBM-bound; written for coalesced 128-bit
// accesses and grids that are multiples of the SM count.
#include <float.h>

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
// AdaptiveAvgPool3d along the slice axis + permute to (D, fX, fY, fZ) + optional fp16 running sum.
// k: (S, T = f0*f1, D) fp16, D fastest.  One CTA transposes a 32(token) x 64(d) tile through smem.
// ---------------------------------------------------------------------------------------------
struct PoolParams {
    const __half* k;
    __half* out;
    int S, T, D, n_out;
    int slice0, n_local, o0;   // k holds global slices [slice0, slice0+n_local); blockIdx.z + o0 = output slab
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
    const int64_t base = i0 * q.s0 + i1 * q.s1 + o * q.so;
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
        __half* dst = q.out + (d0 + dl) * q.sd + i0 * q.s0 + i1 * q.s1 + ob;      // so == 1
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

}  // namespace

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
    if (patch == 8 && D % PE_DCH == 0) {        // register-tiled persistent kernel (patch-8 backbones)
        const size_t smem_t = static_cast<size_t>(PE_TAPS) * (PE_DCH + PE_PX) * sizeof(float);
        const int n_items = (s1 - s0) * (q.f0 + 1);
        dim3 grid_t(n_items < vittf_num_sms() ? n_items : vittf_num_sms(), D / PE_DCH);
#define LAUNCH_PET(T)                                                                                                   \
    do {                                                                                                                \
        static bool configured = false;                                                                                 \
        if (!configured) {                                                                                              \
            VITTF_CHECK_CUDA(cudaFuncSetAttribute(patch_embed_tiled_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                                  static_cast<int>(smem_t)));                                          \
            configured = true;                                                                                          \
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

extern "C" int vittf_pool_axis(const void* k_f16, int S, int slice0, int n_local, int f0, int f1, int D, int axis,
                               int n_out, int o0, int o1, void* out_f16, int accumulate, void* stream) {
    VITTF_REQUIRE(k_f16 && out_f16, "vittf_pool_axis: null pointer");
    VITTF_REQUIRE(S > 0 && f0 > 0 && f1 > 0 && D > 0 && n_out > 0 && n_out <= S, "vittf_pool_axis: bad sizes");
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
    // output (D, A, B, C) contiguous; which of A,B,C are i0 / i1 / o depends on the slicing axis
    int64_t A, B, C;
    if (axis == 2) { A = f0; B = f1; C = n_out; q.s0 = B * C; q.s1 = C; q.so = 1; }          // (D, fX, fY, o)
    else if (axis == 1) { A = f0; B = n_out; C = f1; q.s0 = B * C; q.so = C; q.s1 = 1; }     // (D, fX, o, fZ)
    else { A = n_out; B = f0; C = f1; q.so = B * C; q.s0 = C; q.s1 = 1; }                    // (D, o, fY, fZ)
    q.sd = A * B * C;
    dim3 grid(ceil_div(q.T, 32), ceil_div(D, 64), o1 - o0);
    if (axis == 2 && n_out % POOLZ_OB == 0 && o0 % POOLZ_OB == 0 && (reinterpret_cast<uintptr_t>(out_f16) & 15) == 0) {
        grid.z = ceil_div(o1 - o0, POOLZ_OB);
        pool_axis_z_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(q, o1);
    } else {
        pool_axis_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(q);
    }
    VITTF_CHECK_CUDA(cudaGetLastError());
    vittf_count_launches(1);
    return VITTF_OK;
}
