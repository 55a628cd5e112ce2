// Annotation samplers of /root/reference/compare_feat_sampling.py:13-33: the volume-sized
// stencil work of `sample_surface` -- scipy.ndimage.binary_erosion with generate_binary_structure(3, connectivity)
// (border_value = 0: voxels outside the volume count as background) -- as one streaming kernel; the index draw itself
// is host glue (vittf_b200/compare_feat_sampling.py).
#include "common.cuh"

namespace {

// structure = { offsets with |dx| + |dy| + |dz| <= connectivity } inside the 3 x 3 x 3 box (connectivity >= 3: all 27)
__global__ void __launch_bounds__(256) binary_erosion_kernel(const uint8_t* __restrict__ in, int W, int H, int D, int conn,
                                                             uint8_t* __restrict__ out) {
    const int64_t n = static_cast<int64_t>(W) * H * D;
    for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int z = static_cast<int>(i % D), y = static_cast<int>((i / D) % H), x = static_cast<int>(i / (static_cast<int64_t>(D) * H));
        bool keep = true;
#pragma unroll
        for (int dx = -1; dx <= 1; ++dx)
#pragma unroll
            for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
                for (int dz = -1; dz <= 1; ++dz) {
                    const int dist = (dx != 0) + (dy != 0) + (dz != 0);
                    if (dist > conn) continue;
                    const int xx = x + dx, yy = y + dy, zz = z + dz;
                    const bool inside = xx >= 0 && xx < W && yy >= 0 && yy < H && zz >= 0 && zz < D;
                    keep = keep && inside && __ldg(in + (static_cast<int64_t>(xx) * H + yy) * D + zz) != 0;
                }
        out[i] = keep ? 1 : 0;
    }
}

// ---------------------------------------------------------------------------------------------
// resample_topk (infer.py:90-94): per similarity map, thr = K-th largest value (with multiplicity), then the FIRST K
// voxels in index order with s >= thr -- `(s >= thr).nonzero()[:K]`, which is not "the K largest" when thr is tied.
// One CTA per map: K rounds of (value, lowest index) arg-max over not-yet-taken voxels, then an ordered scan.
// ---------------------------------------------------------------------------------------------
constexpr int TOPK_MAX = 64;

__global__ void __launch_bounds__(512) topk_voxels_kernel(const float* __restrict__ maps, int64_t n, int K,
                                                          long long* __restrict__ out_idx, float* __restrict__ out_thr) {
    __shared__ float s_val[16];
    __shared__ long long s_idx[16];
    __shared__ long long s_taken[TOPK_MAX];
    __shared__ float s_thr;
    __shared__ int s_count[17];
    const float* m = maps + static_cast<int64_t>(blockIdx.x) * n;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    for (int round = 0; round < K; ++round) {
        float bv = -INFINITY;
        long long bi = -1;
        for (int64_t i = tid; i < n; i += blockDim.x) {
            const float v = m[i];
            if (!(v > bv)) continue;                       // strict: the lowest index wins among equal values
            bool taken = false;
            for (int t = 0; t < round; ++t) taken = taken || s_taken[t] == i;
            if (!taken) { bv = v; bi = i; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
            const long long oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (oi >= 0 && (bi < 0 || ov > bv || (ov == bv && oi < bi))) { bv = ov; bi = oi; }
        }
        if (lane == 0) { s_val[wid] = bv; s_idx[wid] = bi; }
        __syncthreads();
        if (tid == 0) {
            for (int w = 1; w < static_cast<int>(blockDim.x >> 5); ++w)
                if (s_idx[w] >= 0 && (s_idx[0] < 0 || s_val[w] > s_val[0] || (s_val[w] == s_val[0] && s_idx[w] < s_idx[0]))) {
                    s_val[0] = s_val[w];
                    s_idx[0] = s_idx[w];
                }
            s_taken[round] = s_idx[0];
            s_thr = s_val[0];
        }
        __syncthreads();
    }
    const float thr = s_thr;
    if (tid == 0 && out_thr) out_thr[blockIdx.x] = thr;
    // ordered selection of the first K voxels with value >= thr: chunks of blockDim voxels, warp ballots for the ranks
    int found = 0;
    for (int64_t base = 0; base < n && found < K; base += blockDim.x) {
        const int64_t i = base + tid;
        const bool hit = i < n && m[i] >= thr;
        const unsigned bal = __ballot_sync(0xffffffffu, hit);
        if (lane == 0) s_count[wid + 1] = __popc(bal);
        __syncthreads();
        if (tid == 0) {
            s_count[0] = 0;
            for (int w = 0; w < static_cast<int>(blockDim.x >> 5); ++w) s_count[w + 1] += s_count[w];
        }
        __syncthreads();
        const int rank = found + s_count[wid] + __popc(bal & ((1u << lane) - 1u));
        if (hit && rank < K) out_idx[static_cast<int64_t>(blockIdx.x) * K + rank] = i;
        found += s_count[blockDim.x >> 5];
        __syncthreads();
    }
}

// take_most_dissimilar (infer.py:118-121): dist_i = 1 - mean_j cos(f_i, f_j) (F.cosine_similarity: every norm clamped at
// 1e-8) or mean_j ||f_i - f_j||.  One CTA per i.
__global__ void __launch_bounds__(256) mean_distance_kernel(const float* __restrict__ f, int N, int F, int measure,
                                                            float* __restrict__ out) {
    extern __shared__ float s_fi[];
    __shared__ float s_red[8];
    const int i = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    for (int k = tid; k < F; k += blockDim.x) s_fi[k] = f[static_cast<int64_t>(i) * F + k];
    __syncthreads();
    float ni2 = 0.0f;
    if (measure == 0)
        for (int k = 0; k < F; ++k) ni2 = fmaf(s_fi[k], s_fi[k], ni2);
    float acc = 0.0f;
    for (int j = wid; j < N; j += 8) {                     // one warp per partner row
        const float* fj = f + static_cast<int64_t>(j) * F;
        float a = 0.0f, b = 0.0f;
        for (int k = lane; k < F; k += 32) {
            const float x = s_fi[k], y = __ldg(fj + k);
            if (measure == 0) { a = fmaf(x, y, a); b = fmaf(y, y, b); }
            else { const float dlt = x - y; a = fmaf(dlt, dlt, a); }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { a += __shfl_xor_sync(0xffffffffu, a, o); b += __shfl_xor_sync(0xffffffffu, b, o); }
        if (lane == 0) acc += measure == 0 ? a / (fmaxf(sqrtf(ni2), 1e-8f) * fmaxf(sqrtf(b), 1e-8f)) : sqrtf(a);
    }
    if (lane == 0) s_red[wid] = acc;
    __syncthreads();
    if (tid == 0) {
        float t = 0.0f;
        for (int w = 0; w < 8; ++w) t += s_red[w];
        const float mean = t / static_cast<float>(N);
        out[i] = measure == 0 ? 1.0f - mean : mean;
    }
}

}  // namespace

namespace {
// ---------------------------------------------------------------------------------------------
// evaluate_similarities.py:66-71: sklearn's confusion_matrix / precision / recall / F1 / Jaccard /
// accuracy all derive from the K x K table of (true, predicted) label pairs, the only volume-sized work.  One pass over
// the two uint8 volumes (2 B / voxel, HBM-bound): every thread keeps PRIVATE counters in shared memory
// (s_cnt[pair][thread]: bank = thread, so neither atomics nor conflicts -- label volumes are long runs of one value, the
// worst case for shared atomics), 16 voxels per 128-bit load; a warp shuffle tree folds the threads, one 64-bit atomic per
// pair and CTA.
// ---------------------------------------------------------------------------------------------
__global__ void confusion_matrix_kernel(const uint8_t* __restrict__ truth, const uint8_t* __restrict__ pred, int64_t n, int K,
                                        unsigned long long* __restrict__ out, unsigned int* __restrict__ out_of_range) {
    extern __shared__ unsigned int s_cnt[];                 // [K * K][blockDim.x]
    const int tid = threadIdx.x, nt = blockDim.x, KK = K * K;
    for (int i = 0; i < KK; ++i) s_cnt[i * nt + tid] = 0;
    unsigned int bad = 0;
    auto count = [&](uint32_t t, uint32_t p) {
        if (t < static_cast<uint32_t>(K) && p < static_cast<uint32_t>(K)) s_cnt[(t * K + p) * nt + tid] += 1;
        else ++bad;
    };
    const int64_t n16 = ((reinterpret_cast<uintptr_t>(truth) | reinterpret_cast<uintptr_t>(pred)) & 15) == 0 ? n / 16 : 0;
    const uint4* t4 = reinterpret_cast<const uint4*>(truth);
    const uint4* p4 = reinterpret_cast<const uint4*>(pred);
    for (int64_t i = blockIdx.x * static_cast<int64_t>(nt) + tid; i < n16; i += static_cast<int64_t>(gridDim.x) * nt) {
        const uint4 a = __ldg(t4 + i), b = __ldg(p4 + i);
        const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, bw[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
        for (int wd = 0; wd < 4; ++wd)
#pragma unroll
            for (int k = 0; k < 4; ++k) count((aw[wd] >> (8 * k)) & 255u, (bw[wd] >> (8 * k)) & 255u);
    }
    for (int64_t i = n16 * 16 + blockIdx.x * static_cast<int64_t>(nt) + tid; i < n; i += static_cast<int64_t>(gridDim.x) * nt)
        count(truth[i], pred[i]);
    // fold the threads: each warp reduces its 32 columns of every pair
    for (int i = 0; i < KK; ++i) {
        unsigned int v = s_cnt[i * nt + tid];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
        if ((tid & 31) == 0 && v) atomicAdd(out + i, static_cast<unsigned long long>(v));
    }
    if (bad) atomicAdd(out_of_range, bad);
}
}  // namespace

extern "C" int vittf_topk_voxels(const float* maps, int n_maps, int64_t n, int K, long long* out_idx, float* out_thr, void* stream) {
    VITTF_REQUIRE(maps && out_idx && n_maps > 0 && n > 0, "vittf_topk_voxels: bad arguments");
    VITTF_REQUIRE(K > 0 && K <= TOPK_MAX && K <= n, "vittf_topk_voxels: K=%d must be in [1, min(%d, n)]", K, TOPK_MAX);
    topk_voxels_kernel<<<n_maps, 512, 0, static_cast<cudaStream_t>(stream)>>>(maps, n, K, out_idx, out_thr);
    VITTF_CHECK_CUDA(cudaGetLastError());
    vittf_count_launches(1);
    return VITTF_OK;
}

extern "C" int vittf_mean_pairwise_distance(const float* feats, int N, int F, int measure, float* out, void* stream) {
    VITTF_REQUIRE(feats && out && N > 0 && F > 0 && F <= 8192, "vittf_mean_pairwise_distance: bad arguments");
    VITTF_REQUIRE(measure == 0 || measure == 1, "vittf_mean_pairwise_distance: measure must be 0 (cosine) or 1 (euclidean)");
    mean_distance_kernel<<<N, 256, F * sizeof(float), static_cast<cudaStream_t>(stream)>>>(feats, N, F, measure, out);
    VITTF_CHECK_CUDA(cudaGetLastError());
    vittf_count_launches(1);
    return VITTF_OK;
}

extern "C" int vittf_binary_erosion(const uint8_t* mask, int W, int H, int D, int connectivity, uint8_t* out, void* stream) {
    VITTF_REQUIRE(mask && out && mask != out, "vittf_binary_erosion: null or aliased pointers");
    VITTF_REQUIRE(W > 0 && H > 0 && D > 0 && connectivity >= 1, "vittf_binary_erosion: bad arguments");
    const int64_t n = static_cast<int64_t>(W) * H * D;
    int64_t blocks = ceil_div_ll(n, 256);
    const int64_t cap = static_cast<int64_t>(vittf_num_sms()) * 16;
    if (blocks > cap) blocks = cap;
    binary_erosion_kernel<<<static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(mask, W, H, D, connectivity, out);
    VITTF_CHECK_CUDA(cudaGetLastError());
    vittf_count_launches(1);
    return VITTF_OK;
}

extern "C" int vittf_confusion_matrix(const uint8_t* truth, const uint8_t* pred, int64_t n, int K, unsigned long long* out,
                                      unsigned int* out_of_range, void* stream) {
    VITTF_REQUIRE(truth && pred && out && out_of_range && n > 0, "vittf_confusion_matrix: bad arguments");
    VITTF_REQUIRE(K >= 1 && K <= 16, "vittf_confusion_matrix: K=%d must be in [1, 16]", K);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    VITTF_CHECK_CUDA(cudaMemsetAsync(out, 0, sizeof(unsigned long long) * K * K, s));
    VITTF_CHECK_CUDA(cudaMemsetAsync(out_of_range, 0, sizeof(unsigned int), s));
    // private counters: K * K * threads * 4 B of shared memory (<= 48 KB), at least one warp
    int threads = 256;
    while (threads > 32 && static_cast<size_t>(K) * K * threads * 4 > 48 * 1024) threads >>= 1;
    int64_t blocks = ceil_div_ll(ceil_div_ll(n, 16), threads);
    const int64_t cap = static_cast<int64_t>(vittf_num_sms()) * 8;
    blocks = blocks < 1 ? 1 : blocks > cap ? cap : blocks;
    confusion_matrix_kernel<<<static_cast<unsigned>(blocks), threads, static_cast<size_t>(K) * K * threads * 4, s>>>(
        truth, pred, n, K, out, out_of_range);
    VITTF_CHECK_CUDA(cudaGetLastError());
    vittf_count_launches(1);
    return VITTF_OK;
}
