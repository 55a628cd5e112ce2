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
// pass 1, shared-memory tiled (fp16 feature volumes, the cache format of infer.py): one CTA owns a 4 x 8 x 16
// brick of low-res voxels plus the forward halo its 13 Gram neighbours need; per F-chunk the 5 x 10 x 20 region
// is staged once in shared memory as half2 words (each thread copies the same two words of every feature plane,
// addresses computed once) and every thread handles TWO z-adjacent voxels: three 4-byte LDS return the z-values
// (z-1 .. z+2) of a neighbour row: 15 LDS.32 + A/4 broadcast LDS.128 per feature for 2 x (A + 14) FMAs.
// ---------------------------------------------------------------------------------------------
constexpr int TB_X = 4, TB_Y = 8, TB_Z = 16, TCH = 16;
constexpr int TR_X = TB_X + 1, TR_Y = TB_Y + 2;
constexpr int TR_W = 10;                        // 4-byte words per staged row: z in [bz0 - 2, bz0 + 18)
constexpr int TR_WORDS = TR_X * TR_Y * TR_W;    // 500 words per feature
constexpr int TR_PER_THREAD = (TR_WORDS + 255) / 256;

template <int AT, bool GRAM>
__global__ void __launch_bounds__(256) sim_lowres_tiled_kernel(const __half* __restrict__ feats, int F, int w, int h, int d,
                                                               const float* __restrict__ protos, int A, int a_base,
                                                               float* __restrict__ dots, float* __restrict__ gram) {
    __shared__ __align__(16) uint32_t s_t[TCH][TR_WORDS];          // half2 words, [rx][ry][word]
    __shared__ __align__(16) float s_p[TCH][AT];
    const int64_t n = static_cast<int64_t>(w) * h * d;
    const int bz0 = blockIdx.x * TB_Z, by0 = blockIdx.y * TB_Y, bx0 = blockIdx.z * TB_X;
    const int tid = threadIdx.x;
    const int lx = tid >> 6, ly = (tid >> 3) & 7, lz = (tid & 7) * 2;
    // staging: every thread owns up to TR_PER_THREAD fixed words of the region (addresses computed once)
    int64_t ld_off[TR_PER_THREAD];
    int ld_mode[TR_PER_THREAD];    // 0 skip, 1 zero, 2 aligned half2 load
#pragma unroll
    for (int k = 0; k < TR_PER_THREAD; ++k) {
        const int i = tid + k * 256;
        ld_mode[k] = 0;
        ld_off[k] = 0;
        if (i < TR_WORDS) {
            const int wd = i % TR_W, ry = (i / TR_W) % TR_Y, rx = i / (TR_W * TR_Y);
            const int gx = bx0 + rx, gy = by0 + ry - 1, gz = bz0 - 2 + wd * 2;     // even: d is even, word never straddles
            const bool ok = gx < w && gy >= 0 && gy < h && gz >= 0 && gz < d;
            ld_mode[k] = ok ? 2 : 1;
            ld_off[k] = ok ? (static_cast<int64_t>(gx) * h + gy) * d + gz : 0;
        }
    }
    float acc[2][AT];
    float g[2][14];
#pragma unroll
    for (int v = 0; v < 2; ++v) {
#pragma unroll
        for (int i = 0; i < AT; ++i) acc[v][i] = 0.0f;
#pragma unroll
        for (int o = 0; o < 14; ++o) g[v][o] = 0.0f;
    }
    for (int f0 = 0; f0 < F; f0 += TCH) {
        __syncthreads();
#pragma unroll
        for (int k = 0; k < TR_PER_THREAD; ++k) {
            if (ld_mode[k] == 0) continue;
#pragma unroll 4
            for (int ff = 0; ff < TCH; ++ff) {
                uint32_t v = 0;
                if (ld_mode[k] == 2 && f0 + ff < F)
                    v = __ldg(reinterpret_cast<const uint32_t*>(feats + static_cast<int64_t>(f0 + ff) * n + ld_off[k]));
                s_t[ff][tid + k * 256] = v;
            }
        }
        for (int i = tid; i < TCH * AT; i += 256) {
            const int ff = i / AT, aa = i - ff * AT;
            s_p[ff][aa] = (a_base + aa < A && f0 + ff < F) ? protos[static_cast<int64_t>(a_base + aa) * F + f0 + ff] : 0.0f;
        }
        __syncthreads();
#pragma unroll 2
        for (int ff = 0; ff < TCH; ++ff) {
            // rows: 0 (x,y)  1 (x,y+1)  2 (x+1,y-1)  3 (x+1,y)  4 (x+1,y+1); values z-1 .. z+2 of voxel pair (z, z+1)
            float r[5][4];
            const int rxs[5] = {lx, lx, lx + 1, lx + 1, lx + 1};
            const int rys[5] = {ly + 1, ly + 2, ly, ly + 1, ly + 2};
#pragma unroll
            for (int k = 0; k < 5; ++k) {
                // words cover z-2..z-1 | z..z+1 | z+2..z+3 (word index lz/2 .. lz/2+2)
                const uint32_t* pr = &s_t[ff][(rxs[k] * TR_Y + rys[k]) * TR_W + (lz >> 1)];
                const float2 a = __half22float2(*reinterpret_cast<const __half2*>(pr));
                const float2 b = __half22float2(*reinterpret_cast<const __half2*>(pr + 1));
                const float2 c = __half22float2(*reinterpret_cast<const __half2*>(pr + 2));
                r[k][0] = a.y; r[k][1] = b.x; r[k][2] = b.y; r[k][3] = c.x;
            }
#pragma unroll
            for (int i = 0; i < AT; i += 4) {
                const float4 pv = *reinterpret_cast<const float4*>(&s_p[ff][i]);
#pragma unroll
                for (int v = 0; v < 2; ++v) {
                    const float c = r[0][1 + v];
                    acc[v][i + 0] = fmaf(c, pv.x, acc[v][i + 0]);
                    acc[v][i + 1] = fmaf(c, pv.y, acc[v][i + 1]);
                    acc[v][i + 2] = fmaf(c, pv.z, acc[v][i + 2]);
                    acc[v][i + 3] = fmaf(c, pv.w, acc[v][i + 3]);
                }
            }
            if (GRAM) {
#pragma unroll
                for (int v = 0; v < 2; ++v) {
                    const float c = r[0][1 + v];
                    g[v][0] = fmaf(c, c, g[v][0]);
#pragma unroll
                    for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
                        for (int dz = -1; dz <= 1; ++dz)
                            g[v][1 + (dy + 1) * 3 + (dz + 1)] = fmaf(c, r[3 + dy][1 + v + dz], g[v][1 + (dy + 1) * 3 + (dz + 1)]);
#pragma unroll
                    for (int dz = -1; dz <= 1; ++dz) g[v][11 + dz] = fmaf(c, r[1][1 + v + dz], g[v][11 + dz]);
                    g[v][13] = fmaf(c, r[0][2 + v], g[v][13]);
                }
            }
        }
    }
    const int gx = bx0 + lx, gy = by0 + ly;
    if (gx >= w || gy >= h) return;
#pragma unroll
    for (int v = 0; v < 2; ++v) {
        const int gz = bz0 + lz + v;
        if (gz >= d) continue;
        const int64_t vox = (static_cast<int64_t>(gx) * h + gy) * d + gz;
#pragma unroll
        for (int i = 0; i < AT; ++i)
            if (a_base + i < A) dots[static_cast<int64_t>(a_base + i) * n + vox] = acc[v][i];
        if (GRAM) {
#pragma unroll
            for (int o = 0; o < 14; ++o) gram[static_cast<int64_t>(o) * n + vox] = g[v][o];
        }
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
// pass 2, fast path for integer up-sampling factors U in {2, 4, 8} (the benchmark shapes: 64^3 -> 256^3,
// 128^3 -> 512^3, 64^3 -> 512^3).  With align_corners=False every low-res CELL c in [-1, n-1] owns the U
// outputs o = U*c + U/2 + k, k in [0, U), whose weights t_k = (k + 0.5) / U do not depend on c.  One thread
// owns (cell, kx, group of 4 ky) and produces a 4 x U block of outputs (U contiguous floats along z):
//   * the 8 corner dots of a prototype are loaded once and interpolated SEPARABLY (x, then z, then y):
//     1.75 lerps per output voxel and prototype instead of 8 FMAs;
//   * |interp(f)|^2 = w^T G w is contracted separably from the 36 Gram scalars of the cell (~7 FMA / voxel);
//   * lanes run along z, so a warp writes 32 * U contiguous floats per output row.
// ---------------------------------------------------------------------------------------------
template <int U>
__global__ void __launch_bounds__(32 * U * (U / 4 > 0 ? U / 4 : 1))
    sim_upsample_cells_kernel(UpParams q) {
    constexpr int KY = U < 4 ? U : 4;          // output rows (y) per thread
    extern __shared__ int s_off[];
    for (int i = threadIdx.x + blockDim.x * (threadIdx.y + blockDim.y * threadIdx.z); i <= q.C;
         i += blockDim.x * blockDim.y * blockDim.z)
        s_off[i] = q.class_offsets[i];
    __syncthreads();
    const int n = q.w;                          // cubic low-res grid (checked by the host)
    const int cz = static_cast<int>(blockIdx.x * 32 + threadIdx.x) - 1;
    const int cy = static_cast<int>(blockIdx.y) - 1, cx = static_cast<int>(blockIdx.z) - 1;
    const int kx = threadIdx.y, ky0 = threadIdx.z * KY;
    if (cz > n - 1) return;
    const int ox = U * cx + U / 2 + kx;
    if (ox < 0 || ox >= q.W) return;
    const int oy_base = U * cy + U / 2 + ky0, oz_base = U * cz + U / 2;
    // this thread's output rows / columns that exist and lie in the requested z-slab
    const int zs = q.z1 - q.z0;
    if (oz_base + U <= q.z0 || oz_base >= q.z1) return;
    const int x0 = cx < 0 ? 0 : cx, x1 = cx + 1 > n - 1 ? n - 1 : cx + 1;
    const int y0 = cy < 0 ? 0 : cy, y1 = cy + 1 > n - 1 ? n - 1 : cy + 1;
    const int z0 = cz < 0 ? 0 : cz, z1 = cz + 1 > n - 1 ? n - 1 : cz + 1;
    const float tx = (kx + 0.5f) / U;
    const int64_t n_lr = static_cast<int64_t>(n) * n * n;
    int64_t idx[8];
#pragma unroll
    for (int c = 0; c < 8; ++c)
        idx[c] = (static_cast<int64_t>((c >> 2) ? x1 : x0) * n + (((c >> 1) & 1) ? y1 : y0)) * n + ((c & 1) ? z1 : z0);

    // ---- 1 / |interp(f)| for the KY x U outputs ---------------------------------------------------------
    float inv[KY][U];
    if (q.mode == VITTF_SIM_NS) {
        // corner Gram matrix (symmetric 8 x 8) from the 14 forward-neighbour planes
        const int cxs[2] = {x0, x1}, cys[2] = {y0, y1}, czs[2] = {z0, z1};
        float G[8][8];
#pragma unroll
        for (int a = 0; a < 8; ++a)
#pragma unroll
            for (int b = a; b < 8; ++b) {
                bool swap;
                const int slot = gram_slot(cxs[b >> 2] - cxs[a >> 2], cys[(b >> 1) & 1] - cys[(a >> 1) & 1],
                                           czs[b & 1] - czs[a & 1], swap);
                const float g = __ldg(q.gram + static_cast<int64_t>(slot) * n_lr + (swap ? idx[b] : idx[a]));
                G[a][b] = g;
                G[b][a] = g;
            }
        // contract x (fixed tx): H[(y,z)][(y',z')] = sum_{x,x'} wx wx' G[(x,y,z)][(x',y',z')]
        const float wx[2] = {1.0f - tx, tx};
        float H[4][4];
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = a; b < 4; ++b) {
                const float v = wx[0] * wx[0] * G[a][b] + wx[0] * wx[1] * (G[a][4 + b] + G[4 + a][b]) + wx[1] * wx[1] * G[4 + a][4 + b];
                H[a][b] = v;
                H[b][a] = v;
            }
#pragma unroll
        for (int kz = 0; kz < U; ++kz) {
            const float tz = (kz + 0.5f) / U;
            const float wz[2] = {1.0f - tz, tz};
            // contract z: J[y][y'] ; index of (y,z) in H is y*2+z
            float J[2][2];
#pragma unroll
            for (int a = 0; a < 2; ++a)
#pragma unroll
                for (int b = a; b < 2; ++b) {
                    const float v = wz[0] * wz[0] * H[a * 2][b * 2] + wz[0] * wz[1] * (H[a * 2][b * 2 + 1] + H[a * 2 + 1][b * 2]) +
                                    wz[1] * wz[1] * H[a * 2 + 1][b * 2 + 1];
                    J[a][b] = v;
                    J[b][a] = v;
                }
#pragma unroll
            for (int ky = 0; ky < KY; ++ky) {
                const float ty = (ky0 + ky + 0.5f) / U;
                const float n2 = (1.0f - ty) * (1.0f - ty) * J[0][0] + 2.0f * ty * (1.0f - ty) * J[0][1] + ty * ty * J[1][1];
                inv[ky][kz] = 1.0f / fmaxf(sqrtf(fmaxf(n2, 0.0f)), 1e-12f);
            }
        }
    } else {
#pragma unroll
        for (int ky = 0; ky < KY; ++ky)
#pragma unroll
            for (int kz = 0; kz < U; ++kz) inv[ky][kz] = 1.0f;
    }

    // ---- per class: max over its prototypes of the separably interpolated dots --------------------------
    const int64_t plane = static_cast<int64_t>(q.H) * zs;
    const int64_t n_out = static_cast<int64_t>(q.W) * plane;
    for (int c = 0; c < q.C; ++c) {
        float best[KY][U];
#pragma unroll
        for (int ky = 0; ky < KY; ++ky)
#pragma unroll
            for (int kz = 0; kz < U; ++kz) best[ky][kz] = -INFINITY;
        for (int a = s_off[c]; a < s_off[c + 1]; ++a) {
            const float* da = q.dots + static_cast<int64_t>(a) * n_lr;
            float v[4];                                  // x-interpolated corners, index y*2+z
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float lo = __ldg(da + idx[k]), hi = __ldg(da + idx[4 + k]);
                v[k] = fmaf(tx, hi - lo, lo);
            }
#pragma unroll
            for (int kz = 0; kz < U; ++kz) {
                const float tz = (kz + 0.5f) / U;
                const float r0 = fmaf(tz, v[1] - v[0], v[0]);      // y0 row
                const float r1 = fmaf(tz, v[3] - v[2], v[2]);      // y1 row
                const float dr = r1 - r0;
#pragma unroll
                for (int ky = 0; ky < KY; ++ky) {
                    const float ty = (ky0 + ky + 0.5f) / U;
                    best[ky][kz] = fmaxf(best[ky][kz], fmaf(ty, dr, r0));
                }
            }
        }
        float* oc = q.out + static_cast<int64_t>(c) * n_out + static_cast<int64_t>(ox) * plane;
#pragma unroll
        for (int ky = 0; ky < KY; ++ky) {
            const int oy = oy_base + ky;
            if (oy < 0 || oy >= q.H) continue;
            float r[U];
#pragma unroll
            for (int kz = 0; kz < U; ++kz)
                r[kz] = pow_unit(fminf(fmaxf(best[ky][kz] * inv[ky][kz], 0.0f), 1.0f), q.exponent);
            float* dst = oc + static_cast<int64_t>(oy) * zs + (oz_base - q.z0);
            if (oz_base >= q.z0 && oz_base + U <= q.z1 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
#pragma unroll
                for (int kz = 0; kz < U; kz += 4) {
                    if (U >= 4) *reinterpret_cast<float4*>(dst + kz) = make_float4(r[kz], r[kz + 1], r[kz + 2], r[kz + 3]);
                }
                if (U == 2) *reinterpret_cast<float2*>(dst) = make_float2(r[0], r[1]);
            } else {
#pragma unroll
                for (int kz = 0; kz < U; ++kz) {
                    const int oz = oz_base + kz;
                    if (oz >= q.z0 && oz < q.z1) dst[kz] = r[kz];
                }
            }
        }
    }
}

template <int U>
void launch_cells(const UpParams& q, cudaStream_t s) {
    constexpr int KY = U < 4 ? U : 4;
    dim3 block(32, U, U / KY);
    dim3 grid(ceil_div(q.w + 1, 32), q.w + 1, q.w + 1);
    sim_upsample_cells_kernel<U><<<grid, block, (q.C + 1) * sizeof(int), s>>>(q);
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

template <bool GRAM>
int launch_lowres_tiled(const __half* feats, int F, int w, int h, int d, const float* protos, int A, float* dots,
                        float* gram, cudaStream_t s) {
    dim3 grid(ceil_div(d, TB_Z), ceil_div(h, TB_Y), ceil_div(w, TB_X));
    bool first = true;
    for (int a_base = 0; a_base < A || first; first = false) {
        const int rem = A - a_base;
        if (rem > 16) {
            if (first && GRAM) sim_lowres_tiled_kernel<32, true><<<grid, 256, 0, s>>>(feats, F, w, h, d, protos, A, a_base, dots, gram);
            else sim_lowres_tiled_kernel<32, false><<<grid, 256, 0, s>>>(feats, F, w, h, d, protos, A, a_base, dots, gram);
            a_base += 32;
        } else if (rem > 8) {
            if (first && GRAM) sim_lowres_tiled_kernel<16, true><<<grid, 256, 0, s>>>(feats, F, w, h, d, protos, A, a_base, dots, gram);
            else sim_lowres_tiled_kernel<16, false><<<grid, 256, 0, s>>>(feats, F, w, h, d, protos, A, a_base, dots, gram);
            a_base += 16;
        } else {
            if (first && GRAM) sim_lowres_tiled_kernel<8, true><<<grid, 256, 0, s>>>(feats, F, w, h, d, protos, A, a_base, dots, gram);
            else sim_lowres_tiled_kernel<8, false><<<grid, 256, 0, s>>>(feats, F, w, h, d, protos, A, a_base, dots, gram);
            a_base += 8;
        }
        vittf_count_launches(1);
    }
    VITTF_CHECK_CUDA(cudaGetLastError());
    return VITTF_OK;
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
        if (d % 2 == 0 && (reinterpret_cast<uintptr_t>(f) & 1) == 0)   // z pairs share a 4-byte word
            return gram ? launch_lowres_tiled<true>(f, F, w, h, d, protos, A, dots, gram, s)
                        : launch_lowres_tiled<false>(f, F, w, h, d, protos, A, dots, gram, s);
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
    // integer power-of-two up-sampling of a cubic grid, max-type modes: separable cell kernel
    const bool cubic = w == h && h == d && W == H && H == D && W % w == 0;
    const int U = cubic ? W / w : 0;
    if (cubic && mode != VITTF_SIM_REFNTF && (U == 2 || U == 4 || U == 8) && mode == VITTF_SIM_NS) {
        if (U == 2) launch_cells<2>(q, static_cast<cudaStream_t>(stream));
        else if (U == 4) launch_cells<4>(q, static_cast<cudaStream_t>(stream));
        else launch_cells<8>(q, static_cast<cudaStream_t>(stream));
        VITTF_CHECK_CUDA(cudaGetLastError());
        vittf_count_launches(1);
        return VITTF_OK;
    }
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
