// Internal (C++ linkage) interface between the translation units of the similarity stage.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

// pass 2 (per output voxel) parameters, shared by every up-sampling kernel
struct UpParams {
    const float* dots;           // (A, n_lr)
    const float* gram;           // (14, n_lr) or nullptr
    const int* class_offsets;    // (C + 1), device
    float* out;                  // (C, x1 - x0, H, z1 - z0)
    int w, h, d, A, C;
    int W, H, D, z0, z1;
    int x0, x1;                  // output x-slab [x0, x1) (slab sharding along the slowest axis); out is (C, x1 - x0, H, z1 - z0)
    int mode;
    float threshold, exponent;
};

// NS mode, uniform integer factor U in {2, 4, 8} (W = U w, H = U h, D = U d): tcgen05 cell-tile kernel (sim_up_tc.cu).
// Returns 0 when it launched, -1 when the shape is not covered (caller falls back to the generic kernel).
// dots_layout 0: dots (A, n_lr); 1: (n_lr, A4), A4 = A rounded up to a multiple of 4.
int vittf_launch_upsample_tc(const UpParams& q, int dots_layout, cudaStream_t stream);
