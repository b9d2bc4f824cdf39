// Interface of the feature-extractor kernel (feature.cu) used by planner.cu.
#pragma once
#include <cuda_runtime.h>

namespace gpp {

constexpr int FE_AGENTS_PER_TILE = 8;

struct FeArgs {
    const float* x;       // [agents][3][11][11]  (device memory, or pinned host memory mapped for the device)
    float* feat;          // [agents][128]
    int total_agents, apt, num_tiles;
    const float* w0t;     // [27][32]   k-major conv weights: row = ci*9 + ky*3 + kx, column = co
    const float* w1t;     // [288][32]
    const float* w2t;     // [288][64]
    const float* w3t;     // [576][64]
    const float* w4t;     // [576][128]
    const float* w5t;     // [128][128] compress MLP, row = input feature
    const float* sc[5];   // eval-mode BatchNorm folded to a per-channel scale ...
    const float* sh[5];   // ... and shift (conv bias included)
    const float* b5;      // compress bias
    int pdl;              // launched with programmatic stream serialization: wait before reading x
    unsigned long long* timing;  // optional [8] per-phase cycle totals of block 0 (debug, GPP_FE_TIMING); null otherwise
};

// Fills apt / num_tiles from total_agents and launches the persistent kernel on `st`.
int launch_feature_kernel(const FeArgs& fa, cudaStream_t st);

}  // namespace gpp
