// Feature extractor for sm_100a: per-agent CNN (5 x conv3x3+BN+ReLU, 3 max-pools) + compress
// MLP, all B*N agents tiled 8 at a time onto persistent 16-warp CTAs.
//
// Replaces the N sequential per-agent ConvLayers/compressMLP calls of
// DecentralPlannerNet.forward (/root/reference/graphs/models/decentralplanner.py:284-290) in
// eval mode.  Activations never leave shared memory between layers.
//
// Mapping (fp32 FMA): lanes own output channels, so every filter fetch is one coalesced 128 B
// line of the k-major re-laid-out weights (L2-resident, read once per warp-item, prefetched one
// input channel ahead in registers); activations are read from shared memory as warp-broadcast
// float4 rows.  conv3/conv4/linear keep all agents of the tile in registers (templated on the
// agent count) and split the input channels over warps; partial sums meet in shared memory.
#include "common.cuh"
#include "feature.cuh"

namespace gpp {

constexpr int FE_THREADS = 512;
constexpr int FE_WARPS = FE_THREADS / 32;
constexpr int AM = FE_AGENTS_PER_TILE;   // 8
constexpr int IN_PIX = 3 * 11 * 11;
// shared-memory map (floats)
constexpr int SZ_IN0 = AM * 3 * 144;     // [a][c][12][12]  zero border, 11x11 inside
constexpr int SZ_ACT12 = AM * 32 * 56;   // [a][c][7][8]    zero border, 5x5 inside, row stride 8
constexpr int OFF_IN0 = 0;
constexpr int OFF_ACT1 = OFF_IN0 + SZ_IN0;
// split-K partial sums, channel innermost with an odd agent stride: the producing warps (lane =
// channel) and the combining pass (agent fastest) are both bank-conflict free
constexpr int P3S = 65, P45S = 129;
constexpr int OFF_PART4 = 0;             // [4 ks][4 p][AM][129] aliases IN0+ACT1 (dead by conv4)
constexpr int SZ_PART4 = 4 * 4 * AM * P45S;
constexpr int OFF_ACT2 = OFF_ACT1 + SZ_ACT12;
constexpr int OFF_PART3 = OFF_ACT2;      // [2 ks][4 p][AM][65] aliases ACT2 (dead by conv3)
constexpr int SZ_PART3 = ((2 * 4 * AM * P3S + 3) / 4) * 4;
constexpr int OFF_PART5 = OFF_PART3 + SZ_PART3;   // [4 ks][AM][129]
constexpr int SZ_PART5 = 4 * AM * P45S;
constexpr int OFF_ACT3 = OFF_ACT2 + SZ_ACT12;     // [64][4][AM]
constexpr int OFF_ACT4 = OFF_ACT3 + 64 * 4 * AM;  // [64][4][AM]
constexpr int OFF_ACT5 = OFF_ACT4 + 64 * 4 * AM;  // [128][AM]
constexpr int FE_SMEM_FLOATS = OFF_ACT5 + 128 * AM;
static_assert(SZ_PART4 <= SZ_IN0 + SZ_ACT12, "PART4 must fit in the IN0+ACT1 region");
static_assert(SZ_PART3 + SZ_PART5 <= SZ_ACT12, "PART3+PART5 must fit in the ACT2 region");
constexpr size_t FE_SMEM_BYTES = sizeof(float) * FE_SMEM_FLOATS;

__device__ __forceinline__ float bn_relu(float v, float sc, float sh) {
    return fmaxf(fmaf(v, sc, sh), 0.f);
}

// conv1 (32->32 on 5x5): NR output rows starting at r0 of one agent; lane = output channel
template <int NR>
__device__ __forceinline__ void conv1_item(const float* __restrict__ act1, float* __restrict__ act2,
                                           const float* __restrict__ w1t, int a, int r0, int lane,
                                           float sc, float sh) {
    float acc[NR][5];
#pragma unroll
    for (int i = 0; i < NR; ++i)
#pragma unroll
        for (int j = 0; j < 5; ++j) acc[i][j] = 0.f;
    float wn[9];
#pragma unroll
    for (int t = 0; t < 9; ++t) wn[t] = __ldg(w1t + t * 32 + lane);
#pragma unroll 1
    for (int ci = 0; ci < 32; ++ci) {
        float w[9];
#pragma unroll
        for (int t = 0; t < 9; ++t) w[t] = wn[t];
        if (ci + 1 < 32) {
#pragma unroll
            for (int t = 0; t < 9; ++t) wn[t] = __ldg(w1t + ((ci + 1) * 9 + t) * 32 + lane);
        }
        const float* base = act1 + ((a * 32 + ci) * 7 + r0) * 8;
#pragma unroll
        for (int iy = 0; iy < NR + 2; ++iy) {
            const float4 ra = ld_smem4(base + iy * 8), rb = ld_smem4(base + iy * 8 + 4);
            const float row[8] = {ra.x, ra.y, ra.z, ra.w, rb.x, rb.y, rb.z, rb.w};
#pragma unroll
            for (int oy = 0; oy < NR; ++oy) {
                const int ky = iy - oy;
                if (ky < 0 || ky > 2) continue;
#pragma unroll
                for (int kx = 0; kx < 3; ++kx)
#pragma unroll
                    for (int ox = 0; ox < 5; ++ox)
                        acc[oy][ox] = fmaf(w[ky * 3 + kx], row[ox + kx], acc[oy][ox]);
            }
        }
    }
    float* out = act2 + ((a * 32 + lane) * 7 + r0 + 1) * 8 + 1;
#pragma unroll
    for (int oy = 0; oy < NR; ++oy)
#pragma unroll
        for (int ox = 0; ox < 5; ++ox) out[oy * 8 + ox] = bn_relu(acc[oy][ox], sc, sh);
}

// loads NA (<= 8, even) agent values of one [c][q] cell: agents are innermost, 8 per cell
template <int NA>
__device__ __forceinline__ void load_agents(const float* p, float (&v)[NA]) {
    const float4 i0 = ld_smem4(p);
    v[0] = i0.x; v[1] = i0.y;
    if (NA > 2) { v[2] = i0.z; v[3] = i0.w; }
    if (NA > 4) {
        const float4 i1 = ld_smem4(p + 4);
        v[4] = i1.x; v[5] = i1.y;
        if (NA > 6) { v[6] = i1.z; v[7] = i1.w; }
    }
}

// conv3 / conv4 / compress MLP for a tile whose valid agents fit NA register slots
template <int NA>
__device__ __forceinline__ void tail_layers(const FeArgs& A, float* sm, int warp, int lane, int a0, int na) {
    float* act3 = sm + OFF_ACT3;
    float* act4 = sm + OFF_ACT4;
    float* act5 = sm + OFF_ACT5;
    float* part3 = sm + OFF_PART3;
    float* part4 = sm + OFF_PART4;
    float* part5 = sm + OFF_PART5;

    // ---- conv3 64->64 on 2x2: item = (channel group, output pixel, half of the input channels);
    //      only the 4 taps that land inside the 2x2 map are read
    {
        const int cg = warp & 1, p = (warp >> 1) & 3, ks = warp >> 3;
        const int co = cg * 32 + lane, py = p >> 1, px = p & 1;
        int toff[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) toff[q] = (((q >> 1) - py + 1) * 3 + ((q & 1) - px + 1)) * 64 + co;
        float acc[NA];
#pragma unroll
        for (int a = 0; a < NA; ++a) acc[a] = 0.f;
        const int c_begin = ks * 32;
        float wn[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) wn[q] = __ldg(A.w3t + c_begin * 9 * 64 + toff[q]);
#pragma unroll 2
        for (int ci = c_begin; ci < c_begin + 32; ++ci) {
            float w[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) w[q] = wn[q];
            if (ci + 1 < c_begin + 32) {
#pragma unroll
                for (int q = 0; q < 4; ++q) wn[q] = __ldg(A.w3t + (ci + 1) * 9 * 64 + toff[q]);
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                float in[NA];
                load_agents<NA>(act3 + (ci * 4 + q) * AM, in);
#pragma unroll
                for (int a = 0; a < NA; ++a) acc[a] = fmaf(w[q], in[a], acc[a]);
            }
        }
#pragma unroll
        for (int a = 0; a < NA; ++a) part3[((ks * 4 + p) * AM + a) * P3S + co] = acc[a];
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 64 * 4 * NA; i += FE_THREADS) {
        const int a = i % NA, cell = i / NA;        // cell = co*4 + p
        const int co = cell >> 2;
        const int p = cell & 3;
        const float v = part3[(p * AM + a) * P3S + co] + part3[((4 + p) * AM + a) * P3S + co];
        act4[cell * AM + a] = bn_relu(v, __ldg(A.sc[3] + co), __ldg(A.sh[3] + co));
    }
    __syncthreads();

    // ---- conv4 64->128 on 2x2: item = (channel group, quarter of the input channels) ----------
    {
        const int cg = warp & 3, ks = warp >> 2;
        const int co = cg * 32 + lane;
        float acc[4][NA];
#pragma unroll
        for (int p = 0; p < 4; ++p)
#pragma unroll
            for (int a = 0; a < NA; ++a) acc[p][a] = 0.f;
        const int c_begin = ks * 16;
        float wn[9];
#pragma unroll
        for (int t = 0; t < 9; ++t) wn[t] = __ldg(A.w4t + (c_begin * 9 + t) * 128 + co);
#pragma unroll 1
        for (int ci = c_begin; ci < c_begin + 16; ++ci) {
            float w[9];
#pragma unroll
            for (int t = 0; t < 9; ++t) w[t] = wn[t];
            if (ci + 1 < c_begin + 16) {
#pragma unroll
                for (int t = 0; t < 9; ++t) wn[t] = __ldg(A.w4t + ((ci + 1) * 9 + t) * 128 + co);
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                float in[NA];
                load_agents<NA>(act4 + (ci * 4 + q) * AM, in);
#pragma unroll
                for (int p = 0; p < 4; ++p) {
                    const int t = ((q >> 1) - (p >> 1) + 1) * 3 + ((q & 1) - (p & 1) + 1);
#pragma unroll
                    for (int a = 0; a < NA; ++a) acc[p][a] = fmaf(w[t], in[a], acc[p][a]);
                }
            }
        }
#pragma unroll
        for (int p = 0; p < 4; ++p)
#pragma unroll
            for (int a = 0; a < NA; ++a) part4[((ks * 4 + p) * AM + a) * P45S + co] = acc[p][a];
    }
    __syncthreads();
    // combine the quarters, BN + ReLU, 2x2 maxpool -> [128][a]
    for (int i = threadIdx.x; i < 128 * NA; i += FE_THREADS) {
        const int a = i % NA, co = i / NA;
        const float sc = __ldg(A.sc[4] + co), sh = __ldg(A.sh[4] + co);
        float m = 0.f;
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            float v = 0.f;
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) v += part4[((ks * 4 + p) * AM + a) * P45S + co];
            m = fmaxf(m, bn_relu(v, sc, sh));
        }
        act5[co * AM + a] = m;
    }
    __syncthreads();

    // ---- compress MLP 128->128 + ReLU: item = (channel group, quarter of the inputs) ----------
    {
        const int cg = warp & 3, ks = warp >> 2;
        const int co = cg * 32 + lane;
        float acc[NA];
#pragma unroll
        for (int a = 0; a < NA; ++a) acc[a] = 0.f;
        const int k_begin = ks * 32;
        float wn[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) wn[j] = __ldg(A.w5t + (k_begin + j) * 128 + co);
#pragma unroll 1
        for (int k = k_begin; k < k_begin + 32; k += 8) {
            float w[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) w[j] = wn[j];
            if (k + 8 < k_begin + 32) {
#pragma unroll
                for (int j = 0; j < 8; ++j) wn[j] = __ldg(A.w5t + (k + 8 + j) * 128 + co);
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                float in[NA];
                load_agents<NA>(act5 + (k + j) * AM, in);
#pragma unroll
                for (int a = 0; a < NA; ++a) acc[a] = fmaf(w[j], in[a], acc[a]);
            }
        }
#pragma unroll
        for (int a = 0; a < NA; ++a) part5[(ks * AM + a) * P45S + co] = acc[a];
    }
    __syncthreads();
    for (int i = threadIdx.x; i < na * 128; i += FE_THREADS) {
        const int a = i >> 7, co = i & 127;
        float v = __ldg(A.b5 + co);
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) v += part5[(ks * AM + a) * P45S + co];
        A.feat[(size_t)(a0 + a) * 128 + co] = fmaxf(v, 0.f);
    }
}

__global__ void __launch_bounds__(FE_THREADS, 1) feature_kernel(const FeArgs A) {
    extern __shared__ __align__(16) float sm[];
    float* in0 = sm + OFF_IN0;
    float* act1 = sm + OFF_ACT1;
    float* act2 = sm + OFF_ACT2;
    float* act3 = sm + OFF_ACT3;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    float w0[27];
#pragma unroll
    for (int j = 0; j < 27; ++j) w0[j] = __ldg(A.w0t + j * 32 + lane);
    const float sc0 = __ldg(A.sc[0] + lane), sh0 = __ldg(A.sh[0] + lane);
    const float sc1 = __ldg(A.sc[1] + lane), sh1 = __ldg(A.sh[1] + lane);

    for (int tile = blockIdx.x; tile < A.num_tiles; tile += gridDim.x) {
        const int a0 = tile * A.apt;
        const int na = min(A.apt, A.total_agents - a0);

        // ---- zero the bordered activation buffers (their interiors / aliases are rewritten
        //      every tile) and stage the binary FOV tensors into the zero-bordered layout --------
        {
            float4* z4 = reinterpret_cast<float4*>(sm + OFF_ACT1);      // ACT1, ACT2, ACT3, ACT4, ACT5
            constexpr int n4 = (FE_SMEM_FLOATS - OFF_ACT1) / 4;
            for (int i = threadIdx.x; i < n4; i += FE_THREADS) z4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            const float* xg = A.x + (size_t)a0 * IN_PIX;
#pragma unroll 4
            for (int i = threadIdx.x; i < AM * 3 * 144; i += FE_THREADS) {
                const int a = i / 432, rem = i - a * 432;
                const int c = rem / 144, p = rem - c * 144;
                const int yy = p / 12 - 1, xx = p % 12 - 1;
                float v = 0.f;
                if (a < na && yy >= 0 && yy < 11 && xx >= 0 && xx < 11)
                    v = __ldg(xg + a * IN_PIX + c * 121 + yy * 11 + xx);
                in0[i] = v;
            }
        }
        __syncthreads();

        // ---- conv0 3->32 on 11x11 (+BN+ReLU) + maxpool2 -> 32 x 5x5 ---------------------------
        // item = (agent, pooled row): conv rows 2pr, 2pr+1, cols 0..9
        for (int item = warp; item < na * 5; item += FE_WARPS) {
            const int a = item / 5, pr = item - a * 5;
            float acc[2][10];
#pragma unroll
            for (int i = 0; i < 2; ++i)
#pragma unroll
                for (int j = 0; j < 10; ++j) acc[i][j] = 0.f;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
#pragma unroll
                for (int iy = 0; iy < 4; ++iy) {
                    const float* rp = in0 + (a * 3 + c) * 144 + (2 * pr + iy) * 12;
                    const float4 r0 = ld_smem4(rp), r1 = ld_smem4(rp + 4), r2 = ld_smem4(rp + 8);
                    const float row[12] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y,
                                           r1.z, r1.w, r2.x, r2.y, r2.z, r2.w};
#pragma unroll
                    for (int oy = 0; oy < 2; ++oy) {
                        const int ky = iy - oy;
                        if (ky < 0 || ky > 2) continue;
#pragma unroll
                        for (int kx = 0; kx < 3; ++kx)
#pragma unroll
                            for (int ox = 0; ox < 10; ++ox)
                                acc[oy][ox] = fmaf(w0[c * 9 + ky * 3 + kx], row[ox + kx], acc[oy][ox]);
                    }
                }
            }
            float* out = act1 + ((a * 32 + lane) * 7 + pr + 1) * 8 + 1;
#pragma unroll
            for (int px = 0; px < 5; ++px) {
                const float m0 = fmaxf(bn_relu(acc[0][2 * px], sc0, sh0), bn_relu(acc[0][2 * px + 1], sc0, sh0));
                const float m1 = fmaxf(bn_relu(acc[1][2 * px], sc0, sh0), bn_relu(acc[1][2 * px + 1], sc0, sh0));
                out[px] = fmaxf(m0, m1);
            }
        }
        __syncthreads();

        // ---- conv1 32->32 on 5x5 (+BN+ReLU); items: rows 0-2 of every agent, then rows 3-4 -----
        for (int item = warp; item < na * 2; item += FE_WARPS) {
            const int rg = item / na, a = item - rg * na;
            if (rg == 0)
                conv1_item<3>(act1, act2, A.w1t, a, 0, lane, sc1, sh1);
            else
                conv1_item<2>(act1, act2, A.w1t, a, 3, lane, sc1, sh1);
        }
        __syncthreads();

        // ---- conv2 32->64 on 5x5 (+BN+ReLU) + maxpool2 -> 64 x 2x2
        //      item = (agent, channel group, pooled row): conv rows 2py, 2py+1, cols 0..3 ----------
        for (int item = warp; item < na * 4; item += FE_WARPS) {
            const int a = item % na, r = item / na;
            const int cg = r & 1, py = r >> 1;
            const int co = cg * 32 + lane;
            float acc[2][4];
#pragma unroll
            for (int i = 0; i < 2; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
            float wn[9];
#pragma unroll
            for (int t = 0; t < 9; ++t) wn[t] = __ldg(A.w2t + t * 64 + co);
#pragma unroll 1
            for (int ci = 0; ci < 32; ++ci) {
                float w[9];
#pragma unroll
                for (int t = 0; t < 9; ++t) w[t] = wn[t];
                if (ci + 1 < 32) {
#pragma unroll
                    for (int t = 0; t < 9; ++t) wn[t] = __ldg(A.w2t + ((ci + 1) * 9 + t) * 64 + co);
                }
                const float* base = act2 + ((a * 32 + ci) * 7 + 2 * py) * 8;
#pragma unroll
                for (int iy = 0; iy < 4; ++iy) {
                    const float4 ra = ld_smem4(base + iy * 8), rb = ld_smem4(base + iy * 8 + 4);
                    const float row[8] = {ra.x, ra.y, ra.z, ra.w, rb.x, rb.y, rb.z, rb.w};
#pragma unroll
                    for (int oy = 0; oy < 2; ++oy) {
                        const int ky = iy - oy;
                        if (ky < 0 || ky > 2) continue;
#pragma unroll
                        for (int kx = 0; kx < 3; ++kx)
#pragma unroll
                            for (int ox = 0; ox < 4; ++ox)
                                acc[oy][ox] = fmaf(w[ky * 3 + kx], row[ox + kx], acc[oy][ox]);
                    }
                }
            }
            const float sc = __ldg(A.sc[2] + co), sh = __ldg(A.sh[2] + co);
#pragma unroll
            for (int px = 0; px < 2; ++px) {
                const float m0 = fmaxf(bn_relu(acc[0][2 * px], sc, sh), bn_relu(acc[0][2 * px + 1], sc, sh));
                const float m1 = fmaxf(bn_relu(acc[1][2 * px], sc, sh), bn_relu(acc[1][2 * px + 1], sc, sh));
                act3[(co * 4 + py * 2 + px) * AM + a] = fmaxf(m0, m1);
            }
        }
        __syncthreads();

        if (na <= 2)
            tail_layers<2>(A, sm, warp, lane, a0, na);
        else if (na <= 4)
            tail_layers<4>(A, sm, warp, lane, a0, na);
        else if (na <= 6)
            tail_layers<6>(A, sm, warp, lane, a0, na);
        else
            tail_layers<8>(A, sm, warp, lane, a0, na);
        __syncthreads();
    }
}

int launch_feature_kernel(const FeArgs& fa_in, cudaStream_t st) {
    FeArgs fa = fa_in;
    static bool configured = false;
    if (!configured) {
        GPP_CUDA_OK(cudaFuncSetAttribute(feature_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)FE_SMEM_BYTES));
        configured = true;
    }
    int apt = (fa.total_agents + sm_count() - 1) / sm_count();
    if (apt > AM) apt = AM;
    if (apt < 1) apt = 1;
    fa.apt = apt;
    fa.num_tiles = (fa.total_agents + apt - 1) / apt;
    const int grid = fa.num_tiles < sm_count() ? fa.num_tiles : sm_count();
    feature_kernel<<<grid, FE_THREADS, FE_SMEM_BYTES, st>>>(fa);
    GPP_LAUNCH_CHECK();
    return GPP_OK;
}

}  // namespace gpp
