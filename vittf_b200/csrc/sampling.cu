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

}  // namespace

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
