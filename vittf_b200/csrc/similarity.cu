// Stage 2: prototype similarity.  Replaces the arithmetic of
//   /root/reference/infer.py:48-72            sample_features3d (grid_sample)
//   /root/reference/predict_ntf.py:59-72      einsum + where/pow + mean           (REF-NTF)
//   /root/reference/old/cluster_dino.py:306-322  normalize + einsum + clamp/pow + max (LEGACY)
// and the north-star composition (up-sample FEATURES trilinearly -> normalise -> dot -> class max)
// WITHOUT ever materialising the up-sampled feature volume:
//   pass 1 (sim_lowres)   reads the (F, n_lr) feature volume once -> dots (A, n_lr) and the 14 Gram
//                         scalars per voxel that determine |interp(f)|^2 inside every cell;
//   pass 2 (sim_upsample) per output voxel: 8-corner trilinear combination of the dots, norm from the
//                         Gram, non-linearity, per-class reduction; writes C maps.
#include "common.cuh"

namespace {

template <typename T>
__device__ __forceinline__ float ldf(const T* p, int64_t i);
template <>
__device__ __forceinline__ float ldf<__half>(const __half* p, int64_t i) { return __half2float(__ldg(p + i)); }
template <>
__device__ __forceinline__ float ldf<float>(const float* p, int64_t i) { return __ldg(p + i); }

// ---------------------------------------------------------------------------------------------
// grid_sample(align_corners=False, padding_mode='zeros'), one CTA per annotation.
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(128) sample_prototypes_kernel(const T* __restrict__ feats, int F, int w, int h, int d,
                                                                const float* __restrict__ rel, int mode,
                                                                float* __restrict__ out) {
    const int a = blockIdx.x;
    // rel is (X,Y,Z); feature dims (w,h,d) are the same order (the reference flips to grid_sample's x=last-dim order)
    const float px = ((rel[a * 3 + 0] + 1.0f) * w - 1.0f) / 2.0f;
    const float py = ((rel[a * 3 + 1] + 1.0f) * h - 1.0f) / 2.0f;
    const float pz = ((rel[a * 3 + 2] + 1.0f) * d - 1.0f) / 2.0f;
    const int64_t n = static_cast<int64_t>(w) * h * d;
    if (mode == 0) {
        const int ix = static_cast<int>(nearbyintf(px)), iy = static_cast<int>(nearbyintf(py)),
                  iz = static_cast<int>(nearbyintf(pz));
        const bool ok = ix >= 0 && ix < w && iy >= 0 && iy < h && iz >= 0 && iz < d;
        const int64_t v = (static_cast<int64_t>(ix) * h + iy) * d + iz;
        for (int f = threadIdx.x; f < F; f += blockDim.x) out[static_cast<int64_t>(a) * F + f] = ok ? ldf<T>(feats, f * n + v) : 0.0f;
        return;
    }
    const float fx = floorf(px), fy = floorf(py), fz = floorf(pz);
    const int x0 = static_cast<int>(fx), y0 = static_cast<int>(fy), z0 = static_cast<int>(fz);
    const float tx = px - fx, ty = py - fy, tz = pz - fz;
    float wgt[8];
    int64_t idx[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        const int cx = x0 + (c >> 2), cy = y0 + ((c >> 1) & 1), cz = z0 + (c & 1);
        const bool ok = cx >= 0 && cx < w && cy >= 0 && cy < h && cz >= 0 && cz < d;
        const float wx = (c >> 2) ? tx : 1.0f - tx, wy = ((c >> 1) & 1) ? ty : 1.0f - ty, wz = (c & 1) ? tz : 1.0f - tz;
        wgt[c] = ok ? wx * wy * wz : 0.0f;
        idx[c] = ok ? (static_cast<int64_t>(cx) * h + cy) * d + cz : 0;
    }
    for (int f = threadIdx.x; f < F; f += blockDim.x) {
        float acc = 0.0f;
#pragma unroll
        for (int c = 0; c < 8; ++c) acc += wgt[c] * ldf<T>(feats, f * n + idx[c]);
        out[static_cast<int64_t>(a) * F + f] = acc;
    }
}

// ---------------------------------------------------------------------------------------------
// pass 1: dots + Gram at feature resolution.  One thread per voxel, AT prototypes in registers,
// prototypes staged (transposed) in shared memory per F-chunk and read as broadcast float4.
// 13 "forward" neighbour offsets (first non-zero component positive), enumeration shared with pass 2.
// ---------------------------------------------------------------------------------------------
constexpr int FCHUNK = 32;

__host__ __device__ __forceinline__ void fwd_offset(int o, int& dx, int& dy, int& dz) {
    // o = 0: self; 1..9: dx=1, (dy,dz) in {-1,0,1}^2; 10..12: dx=0, dy=1, dz in {-1,0,1}; 13: (0,0,1)
    if (o == 0) { dx = dy = dz = 0; }
    else if (o <= 9) { dx = 1; dy = (o - 1) / 3 - 1; dz = (o - 1) % 3 - 1; }
    else if (o <= 12) { dx = 0; dy = 1; dz = o - 11; }
    else { dx = 0; dy = 0; dz = 1; }
}
// index of the Gram slot for a relative offset (any sign); `swap` tells the caller to anchor at the other voxel
__device__ __forceinline__ int gram_slot(int dx, int dy, int dz, bool& swap) {
    swap = dx < 0 || (dx == 0 && (dy < 0 || (dy == 0 && dz < 0)));
    if (swap) { dx = -dx; dy = -dy; dz = -dz; }
    if (dx == 1) return 1 + (dy + 1) * 3 + (dz + 1);
    if (dy == 1) return 11 + dz;
    return dz == 1 ? 13 : 0;
}

template <typename T, int AT, bool GRAM>
__global__ void __launch_bounds__(128) sim_lowres_kernel(const T* __restrict__ feats, int F, int w, int h, int d,
                                                         const float* __restrict__ protos, int A, int a_base,
                                                         float* __restrict__ dots, float* __restrict__ gram) {
    __shared__ __align__(16) float s_p[FCHUNK][AT];
    const int64_t n = static_cast<int64_t>(w) * h * d;
    const int64_t v = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
    const bool live = v < n;
    const int64_t vv = live ? v : n - 1;
    int64_t nb[14];
    bool nb_ok[14];
    if (GRAM) {
        const int iz = static_cast<int>(vv % d), iy = static_cast<int>((vv / d) % h), ix = static_cast<int>(vv / (static_cast<int64_t>(d) * h));
#pragma unroll
        for (int o = 0; o < 14; ++o) {
            int dx, dy, dz;
            fwd_offset(o, dx, dy, dz);
            const int jx = ix + dx, jy = iy + dy, jz = iz + dz;
            nb_ok[o] = jx >= 0 && jx < w && jy >= 0 && jy < h && jz >= 0 && jz < d;
            nb[o] = nb_ok[o] ? (static_cast<int64_t>(jx) * h + jy) * d + jz : vv;
        }
    }
    float acc[AT];
    float g[14];
#pragma unroll
    for (int i = 0; i < AT; ++i) acc[i] = 0.0f;
#pragma unroll
    for (int o = 0; o < 14; ++o) g[o] = 0.0f;

    for (int f0 = 0; f0 < F; f0 += FCHUNK) {
        __syncthreads();
        for (int i = threadIdx.x; i < FCHUNK * AT; i += blockDim.x) {
            const int ff = i / AT, aa = i - ff * AT;
            const int a = a_base + aa;
            s_p[ff][aa] = (a < A && f0 + ff < F) ? protos[static_cast<int64_t>(a) * F + f0 + ff] : 0.0f;
        }
        __syncthreads();
        const int fmax = min(FCHUNK, F - f0);
        for (int ff = 0; ff < fmax; ++ff) {
            const T* plane = feats + static_cast<int64_t>(f0 + ff) * n;
            const float x = ldf<T>(plane, vv);
#pragma unroll
            for (int i = 0; i < AT; i += 4) {
                const float4 pv = *reinterpret_cast<const float4*>(&s_p[ff][i]);
                acc[i + 0] = fmaf(x, pv.x, acc[i + 0]);
                acc[i + 1] = fmaf(x, pv.y, acc[i + 1]);
                acc[i + 2] = fmaf(x, pv.z, acc[i + 2]);
                acc[i + 3] = fmaf(x, pv.w, acc[i + 3]);
            }
            if (GRAM) {
                g[0] = fmaf(x, x, g[0]);
#pragma unroll
                for (int o = 1; o < 14; ++o) g[o] = fmaf(x, ldf<T>(plane, nb[o]), g[o]);
            }
        }
    }
    if (!live) return;
#pragma unroll
    for (int i = 0; i < AT; ++i)
        if (a_base + i < A) dots[static_cast<int64_t>(a_base + i) * n + v] = acc[i];
    if (GRAM) {
#pragma unroll
        for (int o = 0; o < 14; ++o) gram[static_cast<int64_t>(o) * n + v] = nb_ok[o] ? g[o] : 0.0f;
    }
}

// ---------------------------------------------------------------------------------------------
// pass 2: per output voxel.  Index rule of F.interpolate(mode='trilinear', align_corners=False):
//   src = max((dst + 0.5) * in/out - 0.5, 0); i0 = floor(src); i1 = min(i0 + 1, in - 1); t = src - i0.
// ---------------------------------------------------------------------------------------------
struct UpParams {
    const float* dots;
    const float* gram;
    const int* class_offsets;
    float* out;
    int w, h, d, A, C;
    int W, H, D, z0, z1;
    int mode;
    float threshold, exponent;
};

__device__ __forceinline__ void src_index(int dst, int in, int out, int& i0, int& i1, float& t) {
    if (in == out) { i0 = i1 = dst; t = 0.0f; return; }
    const float scale = static_cast<float>(in) / static_cast<float>(out);
    float src = (dst + 0.5f) * scale - 0.5f;
    src = src < 0.0f ? 0.0f : src;
    i0 = static_cast<int>(src);
    i1 = i0 + (i0 < in - 1 ? 1 : 0);
    t = src - i0;
}

__device__ __forceinline__ float pow_unit(float x, float e) {
    // x in [0, 1] (or 0 after thresholding); exact fast paths for the exponents the reference uses
    if (e == 2.0f) return x * x;
    if (e == 2.5f) return x * x * sqrtf(x);
    if (e == 1.0f) return x;
    return x > 0.0f ? __powf(x, e) : 0.0f;
}

__global__ void __launch_bounds__(256) sim_upsample_kernel(UpParams q) {
    extern __shared__ int s_off[];
    for (int i = threadIdx.x; i <= q.C; i += blockDim.x) s_off[i] = q.class_offsets[i];
    __syncthreads();
    const int zs = q.z1 - q.z0;
    const int64_t n_out = static_cast<int64_t>(q.W) * q.H * zs;
    const int64_t n_lr = static_cast<int64_t>(q.w) * q.h * q.d;
    for (int64_t o = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; o < n_out;
         o += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int oz = static_cast<int>(o % zs), oy = static_cast<int>((o / zs) % q.H),
                  ox = static_cast<int>(o / (static_cast<int64_t>(zs) * q.H));
        int x0, x1, y0, y1, c0, c1;
        float tx, ty, tz;
        src_index(ox, q.w, q.W, x0, x1, tx);
        src_index(oy, q.h, q.H, y0, y1, ty);
        src_index(oz + q.z0, q.d, q.D, c0, c1, tz);
        int64_t idx[8];
        float wgt[8];
        int cx[8], cy[8], cz[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            cx[c] = (c >> 2) ? x1 : x0;
            cy[c] = ((c >> 1) & 1) ? y1 : y0;
            cz[c] = (c & 1) ? c1 : c0;
            idx[c] = (static_cast<int64_t>(cx[c]) * q.h + cy[c]) * q.d + cz[c];
            wgt[c] = ((c >> 2) ? tx : 1.0f - tx) * (((c >> 1) & 1) ? ty : 1.0f - ty) * ((c & 1) ? tz : 1.0f - tz);
        }
        float inv_norm = 1.0f;
        if (q.mode == VITTF_SIM_NS) {
            float n2 = 0.0f;
#pragma unroll
            for (int a = 0; a < 8; ++a) {
                if (wgt[a] == 0.0f) continue;
                n2 = fmaf(wgt[a] * wgt[a], __ldg(q.gram + idx[a]), n2);
#pragma unroll
                for (int b = a + 1; b < 8; ++b) {
                    if (wgt[b] == 0.0f) continue;
                    bool swap;
                    const int slot = gram_slot(cx[b] - cx[a], cy[b] - cy[a], cz[b] - cz[a], swap);
                    const float gv = __ldg(q.gram + static_cast<int64_t>(slot) * n_lr + (swap ? idx[b] : idx[a]));
                    n2 = fmaf(2.0f * wgt[a] * wgt[b], gv, n2);
                }
            }
            inv_norm = 1.0f / fmaxf(sqrtf(fmaxf(n2, 0.0f)), 1e-12f);     // F.normalize eps
        } else if (q.mode == VITTF_SIM_LEGACY) {
            inv_norm = 1.0f / fmaxf(sqrtf(__ldg(q.gram + idx[0])), 1e-12f);
        }
        for (int c = 0; c < q.C; ++c) {
            float red = q.mode == VITTF_SIM_REFNTF ? 0.0f : -INFINITY;
            for (int a = s_off[c]; a < s_off[c + 1]; ++a) {
                const float* da = q.dots + static_cast<int64_t>(a) * n_lr;
                float s = 0.0f;
#pragma unroll
                for (int k = 0; k < 8; ++k)
                    if (wgt[k] != 0.0f) s = fmaf(wgt[k], __ldg(da + idx[k]), s);
                if (q.mode == VITTF_SIM_REFNTF) {
                    red += pow_unit(s >= q.threshold ? s : 0.0f, q.exponent);
                } else {
                    red = fmaxf(red, s);  // clamp(0,1)**e is monotone: reduce first, transform once
                }
            }
            float r;
            if (q.mode == VITTF_SIM_REFNTF) r = red / static_cast<float>(s_off[c + 1] - s_off[c]);
            else r = pow_unit(fminf(fmaxf(red * inv_norm, 0.0f), 1.0f), q.exponent);
            q.out[static_cast<int64_t>(c) * n_out + o] = r;
        }
    }
}

// ---------------------------------------------------------------------------------------------
__global__ void class_max_init_kernel(float* out, int C) {
    if (threadIdx.x < C) out[threadIdx.x] = -INFINITY;
}
__global__ void __launch_bounds__(256) class_max_kernel(const float* __restrict__ sims, int64_t n, float* out) {
    const float* s = sims + static_cast<int64_t>(blockIdx.y) * n;
    float m = -INFINITY;
    for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x)
        m = fmaxf(m, __ldg(s + i));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) {
        if (m >= 0.0f) atomicMax(reinterpret_cast<int*>(out + blockIdx.y), __float_as_int(m));
        else atomicMin(reinterpret_cast<unsigned int*>(out + blockIdx.y), __float_as_uint(m));
    }
}

template <typename T>
__global__ void __launch_bounds__(256) labels_kernel(const T* __restrict__ sims, int C, int64_t n,
                                                     const int* __restrict__ thr, int mode, uint8_t* __restrict__ out) {
    for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        if (mode == 0) {  // predict_ntf.py:203-215
            float best = 0.0f;
            int lab = 0;
            for (int c = 0; c < C; ++c) {
                const float s = static_cast<float>(sims[static_cast<int64_t>(c) * n + i]);
                if (s > static_cast<float>(thr[c]) && s > best) { best = s; lab = c + 1; }
            }
            out[i] = static_cast<uint8_t>(lab);
        } else {          // argmax(0), first maximum wins
            float best = static_cast<float>(sims[i]);
            int lab = 0;
            for (int c = 1; c < C; ++c) {
                const float s = static_cast<float>(sims[static_cast<int64_t>(c) * n + i]);
                if (s > best) { best = s; lab = c; }
            }
            out[i] = static_cast<uint8_t>(lab);
        }
    }
}

template <typename T, bool GRAM>
int launch_lowres(const T* feats, int F, int w, int h, int d, const float* protos, int A, float* dots, float* gram,
                  cudaStream_t s) {
    const int64_t n = static_cast<int64_t>(w) * h * d;
    const unsigned grid = static_cast<unsigned>(ceil_div_ll(n, 128));
    int a_base = 0;
    bool first = true;
    while (a_base < A || first) {
        const int rem = A - a_base;
        // Gram is accumulated by the first prototype group only
        if (rem > 16) {
            if (first && GRAM) sim_lowres_kernel<T, 32, true><<<grid, 128, 0, s>>>(feats, F, w, h, d, protos, A, a_base, dots, gram);
            else sim_lowres_kernel<T, 32, false><<<grid, 128, 0, s>>>(feats, F, w, h, d, protos, A, a_base, dots, gram);
            a_base += 32;
            vittf_count_launches(1);
        } else if (rem > 8) {
            if (first && GRAM) sim_lowres_kernel<T, 16, true><<<grid, 128, 0, s>>>(feats, F, w, h, d, protos, A, a_base, dots, gram);
            else sim_lowres_kernel<T, 16, false><<<grid, 128, 0, s>>>(feats, F, w, h, d, protos, A, a_base, dots, gram);
            a_base += 16;
            vittf_count_launches(1);
        } else {
            if (first && GRAM) sim_lowres_kernel<T, 8, true><<<grid, 128, 0, s>>>(feats, F, w, h, d, protos, A, a_base, dots, gram);
            else sim_lowres_kernel<T, 8, false><<<grid, 128, 0, s>>>(feats, F, w, h, d, protos, A, a_base, dots, gram);
            a_base += 8;
            vittf_count_launches(1);
        }
        first = false;
    }
    VITTF_CHECK_CUDA(cudaGetLastError());
    return VITTF_OK;
}

}  // namespace

extern "C" int vittf_sample_prototypes(const void* feats, int feat_dtype, int F, int w, int h, int d, const float* rel,
                                       int A, int mode, float* out, void* stream) {
    VITTF_REQUIRE(feats && rel && out, "vittf_sample_prototypes: null pointer");
    VITTF_REQUIRE(F > 0 && w > 0 && h > 0 && d > 0 && A > 0, "vittf_sample_prototypes: empty problem");
    VITTF_REQUIRE(mode == 0 || mode == 1, "vittf_sample_prototypes: mode must be 0 (nearest) or 1 (trilinear)");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (feat_dtype == VITTF_F16)
        sample_prototypes_kernel<__half><<<A, 128, 0, s>>>(static_cast<const __half*>(feats), F, w, h, d, rel, mode, out);
    else if (feat_dtype == VITTF_F32)
        sample_prototypes_kernel<float><<<A, 128, 0, s>>>(static_cast<const float*>(feats), F, w, h, d, rel, mode, out);
    else
        VITTF_REQUIRE(false, "vittf_sample_prototypes: features must be fp16 or fp32");
    VITTF_CHECK_CUDA(cudaGetLastError());
    vittf_count_launches(1);
    return VITTF_OK;
}

extern "C" int vittf_sim_lowres(const void* feats, int feat_dtype, int F, int w, int h, int d, const float* protos, int A,
                                float* dots, float* gram, void* stream) {
    VITTF_REQUIRE(feats && protos && dots, "vittf_sim_lowres: null pointer");
    VITTF_REQUIRE(F > 0 && w > 0 && h > 0 && d > 0 && A > 0, "vittf_sim_lowres: empty problem");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (feat_dtype == VITTF_F16) {
        const __half* f = static_cast<const __half*>(feats);
        return gram ? launch_lowres<__half, true>(f, F, w, h, d, protos, A, dots, gram, s)
                    : launch_lowres<__half, false>(f, F, w, h, d, protos, A, dots, gram, s);
    } else if (feat_dtype == VITTF_F32) {
        const float* f = static_cast<const float*>(feats);
        return gram ? launch_lowres<float, true>(f, F, w, h, d, protos, A, dots, gram, s)
                    : launch_lowres<float, false>(f, F, w, h, d, protos, A, dots, gram, s);
    }
    VITTF_REQUIRE(false, "vittf_sim_lowres: features must be fp16 or fp32");
    return VITTF_OK;
}

extern "C" int vittf_sim_upsample(const float* dots, const float* gram, int w, int h, int d, int A,
                                  const int* class_offsets, int C, int W, int H, int D, int z0, int z1, int mode,
                                  float threshold, float exponent, float* out, void* stream) {
    VITTF_REQUIRE(dots && class_offsets && out, "vittf_sim_upsample: null pointer");
    VITTF_REQUIRE(mode == VITTF_SIM_NS || mode == VITTF_SIM_REFNTF || mode == VITTF_SIM_LEGACY,
                  "vittf_sim_upsample: unknown mode %d", mode);
    VITTF_REQUIRE(mode == VITTF_SIM_REFNTF || gram, "vittf_sim_upsample: NS/LEGACY modes need the Gram planes");
    VITTF_REQUIRE(C > 0 && A > 0 && W > 0 && H > 0 && D > 0 && z0 >= 0 && z1 > z0 && z1 <= D,
                  "vittf_sim_upsample: bad sizes (C=%d A=%d out=%dx%dx%d z=[%d,%d))", C, A, W, H, D, z0, z1);
    UpParams q{dots, gram, class_offsets, out, w, h, d, A, C, W, H, D, z0, z1, mode, threshold, exponent};
    const int64_t n_out = static_cast<int64_t>(W) * H * (z1 - z0);
    int64_t blocks = ceil_div_ll(n_out, 256);
    const int64_t cap = static_cast<int64_t>(vittf_num_sms()) * 32;
    if (blocks > cap) blocks = cap;
    sim_upsample_kernel<<<static_cast<unsigned>(blocks), 256, (C + 1) * sizeof(int), static_cast<cudaStream_t>(stream)>>>(q);
    VITTF_CHECK_CUDA(cudaGetLastError());
    vittf_count_launches(1);
    return VITTF_OK;
}

extern "C" int vittf_class_max(const float* sims, int C, int64_t n, float* out, void* stream) {
    VITTF_REQUIRE(sims && out && C > 0 && C <= 1024 && n > 0, "vittf_class_max: bad arguments");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    class_max_init_kernel<<<1, 1024, 0, s>>>(out, C);
    dim3 grid(vittf_num_sms() * 2, C);
    class_max_kernel<<<grid, 256, 0, s>>>(sims, n, out);
    VITTF_CHECK_CUDA(cudaGetLastError());
    vittf_count_launches(2);
    return VITTF_OK;
}

extern "C" int vittf_labels(const void* sims, int sims_dtype, int C, int64_t n, const int* thresholds_u8, int mode,
                            uint8_t* out, void* stream) {
    VITTF_REQUIRE(sims && out && C > 0 && C < 255 && n > 0, "vittf_labels: bad arguments");
    VITTF_REQUIRE(mode == 1 || thresholds_u8, "vittf_labels: thresholds required for mode 0");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    int64_t blocks = ceil_div_ll(n, 256);
    const int64_t cap = static_cast<int64_t>(vittf_num_sms()) * 16;
    if (blocks > cap) blocks = cap;
    if (sims_dtype == VITTF_U8)
        labels_kernel<uint8_t><<<static_cast<unsigned>(blocks), 256, 0, s>>>(static_cast<const uint8_t*>(sims), C, n, thresholds_u8, mode, out);
    else if (sims_dtype == VITTF_F32)
        labels_kernel<float><<<static_cast<unsigned>(blocks), 256, 0, s>>>(static_cast<const float*>(sims), C, n, thresholds_u8, mode, out);
    else
        VITTF_REQUIRE(false, "vittf_labels: similarity maps must be uint8 or fp32");
    VITTF_CHECK_CUDA(cudaGetLastError());
    vittf_count_launches(1);
    return VITTF_OK;
}
