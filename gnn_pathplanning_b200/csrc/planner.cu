// Whole-planner inference for sm_100a: per-agent CNN + compress MLP ("feature extractor")
// followed by the fused graph filter + ReLU + action MLP (graph_filter.cu).
//
// Replaces DecentralPlannerNet.forward in eval mode
// (/root/reference/graphs/models/decentralplanner.py:278-318).  Where the reference runs
// the CNN N times, once per agent, on [B,3,11,11] slices (:284-290), all B*N agents are
// independent here and are tiled 8 at a time onto persistent CTAs; activations never leave
// shared memory between the five conv layers, the pools and the compress MLP.
//
// Feature-extractor mapping (fp32 FMA): lanes own output channels, so every weight fetch is
// a coalesced 128 B line from the k-major re-laid-out filters (streamed from L2, each element
// used by one thread for all agents/positions in its register tile), while the activations
// are read from shared memory as warp-broadcast float4 rows.
#include "common.cuh"

#include <stdarg.h>
#include <string.h>

#include <vector>

namespace gpp {

// ---- error / bookkeeping ----------------------------------------------------------------
static thread_local char g_err[512] = "";
thread_local unsigned long long g_launches = 0;

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int sm_count() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess ||
            cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
            n = 148;
    }
    return n;
}

// internal launchers from graph_filter.cu
int launch_transpose_taps(const float* w, float* wt, int F, int KG, cudaStream_t st);
int launch_gf_forward_fast(const float* x, const void* S, int s_is_f64, const float* wt,
                           const float* bias, float* y, const float* wa, const float* ba,
                           float* logits, int B, int N, int K, int x_layout, int y_layout,
                           int relu, cudaStream_t st);

// ---- feature extractor ------------------------------------------------------------------
constexpr int FE_THREADS = 256;
constexpr int FE_WARPS = FE_THREADS / 32;
constexpr int AM = 8;                    // agents per tile
constexpr int IN_PIX = 3 * 11 * 11;      // 363 floats per agent
// shared-memory float offsets
constexpr int SZ_IN0 = AM * 3 * 144;     // [a][c][12][12]  (1-pixel zero border, 11x11 inside)
constexpr int SZ_ACT12 = AM * 32 * 56;   // [a][c][7][8]    (5x5 inside a zero border, row stride 8)
constexpr int SZ_ACT34 = 64 * 4 * AM;    // [c][q][a]
constexpr int SZ_PART4 = 2 * 128 * 4 * AM;
constexpr int SZ_ACT5 = 128 * AM;        // [c][a]
constexpr int SZ_PART5 = 2 * 128 * AM;
constexpr int OFF_IN0 = 0;
constexpr int OFF_ACT1 = OFF_IN0 + SZ_IN0;
constexpr int OFF_ACT2 = OFF_ACT1 + SZ_ACT12;
constexpr int OFF_ACT3 = OFF_ACT2 + SZ_ACT12;
constexpr int OFF_ACT4 = OFF_ACT3 + SZ_ACT34;
constexpr int OFF_PART4 = OFF_ACT4 + SZ_ACT34;
constexpr int OFF_ACT5 = OFF_PART4 + SZ_PART4;
constexpr int OFF_PART5 = OFF_ACT5 + SZ_ACT5;
constexpr int FE_SMEM_FLOATS = OFF_PART5 + SZ_PART5;
constexpr size_t FE_SMEM_BYTES = sizeof(float) * FE_SMEM_FLOATS;

struct FeArgs {
    const float* x;       // [agents][3][11][11]
    float* feat;          // [agents][128]
    int total_agents, apt, num_tiles;
    const float* w0t;     // [27][32]
    const float* w1t;     // [288][32]
    const float* w2t;     // [288][64]
    const float* w3t;     // [576][64]
    const float* w4t;     // [576][128]
    const float* w5t;     // [128][128]
    const float* sc[5];   // folded BatchNorm scale per conv layer
    const float* sh[5];   // folded conv-bias + BatchNorm shift
    const float* b5;      // compress bias
};

__device__ __forceinline__ float bn_relu(float v, float sc, float sh) {
    return fmaxf(fmaf(v, sc, sh), 0.f);
}

// conv1 (32->32 on 5x5): NR output rows starting at r0 for one agent; lane = output channel
template <int NR>
__device__ __forceinline__ void conv1_item(const float* __restrict__ act1, float* __restrict__ act2,
                                           const float* __restrict__ w1t, int a, int r0, int lane,
                                           float sc, float sh) {
    float acc[NR][5];
#pragma unroll
    for (int i = 0; i < NR; ++i)
#pragma unroll
        for (int j = 0; j < 5; ++j) acc[i][j] = 0.f;
#pragma unroll 2
    for (int ci = 0; ci < 32; ++ci) {
        float w[9];
#pragma unroll
        for (int t = 0; t < 9; ++t) w[t] = __ldg(w1t + (ci * 9 + t) * 32 + lane);
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

__global__ void __launch_bounds__(FE_THREADS, 1) feature_kernel(const FeArgs A) {
    extern __shared__ __align__(16) float sm[];
    float* in0 = sm + OFF_IN0;
    float* act1 = sm + OFF_ACT1;
    float* act2 = sm + OFF_ACT2;
    float* act3 = sm + OFF_ACT3;
    float* act4 = sm + OFF_ACT4;
    float* part4 = sm + OFF_PART4;
    float* act5 = sm + OFF_ACT5;
    float* part5 = sm + OFF_PART5;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    // zero everything once: the padded borders of act1/act2 stay zero for the whole kernel
    for (int i = threadIdx.x; i < FE_SMEM_FLOATS; i += FE_THREADS) sm[i] = 0.f;

    float w0[27];
#pragma unroll
    for (int j = 0; j < 27; ++j) w0[j] = __ldg(A.w0t + j * 32 + lane);
    const float sc0 = __ldg(A.sc[0] + lane), sh0 = __ldg(A.sh[0] + lane);
    const float sc1 = __ldg(A.sc[1] + lane), sh1 = __ldg(A.sh[1] + lane);
    __syncthreads();

    for (int tile = blockIdx.x; tile < A.num_tiles; tile += gridDim.x) {
        const int a0 = tile * A.apt;
        const int na = min(A.apt, A.total_agents - a0);

        // ---- stage the binary FOV tensors into the zero-bordered layout ---------------
        for (int i = threadIdx.x; i < AM * 3 * 144; i += FE_THREADS) {
            const int a = i / 432, rem = i - a * 432;
            const int c = rem / 144, p = rem - c * 144;
            const int yy = p / 12 - 1, xx = p % 12 - 1;
            float v = 0.f;
            if (a < na && yy >= 0 && yy < 11 && xx >= 0 && xx < 11)
                v = __ldg(A.x + (size_t)(a0 + a) * IN_PIX + c * 121 + yy * 11 + xx);
            in0[i] = v;
        }
        __syncthreads();

        // ---- conv0 3->32 on 11x11 (+BN+ReLU) + maxpool2 -> 32 x 5x5 ---------------------
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

        // ---- conv1 32->32 on 5x5 (+BN+ReLU); items: rows 0-2 of every agent, then rows 3-4 ---
        for (int item = warp; item < na * 2; item += FE_WARPS) {
            const int rg = item / na, a = item - rg * na;
            if (rg == 0)
                conv1_item<3>(act1, act2, A.w1t, a, 0, lane, sc1, sh1);
            else
                conv1_item<2>(act1, act2, A.w1t, a, 3, lane, sc1, sh1);
        }
        __syncthreads();

        // ---- conv2 32->64 on 5x5 (+BN+ReLU) + maxpool2 -> 64 x 2x2 (conv rows/cols 0..3) ---
        for (int item = warp; item < na * 2; item += FE_WARPS) {
            const int cg = item / na, a = item - cg * na;
            const int co = cg * 32 + lane;
            float acc[4][4];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
#pragma unroll 2
            for (int ci = 0; ci < 32; ++ci) {
                float w[9];
#pragma unroll
                for (int t = 0; t < 9; ++t) w[t] = __ldg(A.w2t + (ci * 9 + t) * 64 + co);
                const float* base = act2 + (a * 32 + ci) * 56;
#pragma unroll
                for (int iy = 0; iy < 6; ++iy) {
                    const float4 ra = ld_smem4(base + iy * 8), rb = ld_smem4(base + iy * 8 + 4);
                    const float row[8] = {ra.x, ra.y, ra.z, ra.w, rb.x, rb.y, rb.z, rb.w};
#pragma unroll
                    for (int oy = 0; oy < 4; ++oy) {
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
            for (int py = 0; py < 2; ++py)
#pragma unroll
                for (int px = 0; px < 2; ++px) {
                    const float m0 = fmaxf(bn_relu(acc[2 * py][2 * px], sc, sh), bn_relu(acc[2 * py][2 * px + 1], sc, sh));
                    const float m1 = fmaxf(bn_relu(acc[2 * py + 1][2 * px], sc, sh), bn_relu(acc[2 * py + 1][2 * px + 1], sc, sh));
                    act3[(co * 4 + py * 2 + px) * AM + a] = fmaxf(m0, m1);
                }
        }
        __syncthreads();

        // ---- conv3 64->64 on 2x2 (+BN+ReLU): item = (channel group, output pixel), all agents
        //      of the tile in registers; only the 4 taps that land inside the 2x2 map are read
        {
            const int cg = warp & 1, p = warp >> 1;
            const int co = cg * 32 + lane, py = p >> 1, px = p & 1;
            float acc[AM];
#pragma unroll
            for (int a = 0; a < AM; ++a) acc[a] = 0.f;
#pragma unroll 2
            for (int ci = 0; ci < 64; ++ci) {
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int t = ((q >> 1) - py + 1) * 3 + ((q & 1) - px + 1);
                    const float w = __ldg(A.w3t + (ci * 9 + t) * 64 + co);
                    const float4 i0 = ld_smem4(act3 + (ci * 4 + q) * AM), i1 = ld_smem4(act3 + (ci * 4 + q) * AM + 4);
                    acc[0] = fmaf(w, i0.x, acc[0]); acc[1] = fmaf(w, i0.y, acc[1]);
                    acc[2] = fmaf(w, i0.z, acc[2]); acc[3] = fmaf(w, i0.w, acc[3]);
                    acc[4] = fmaf(w, i1.x, acc[4]); acc[5] = fmaf(w, i1.y, acc[5]);
                    acc[6] = fmaf(w, i1.z, acc[6]); acc[7] = fmaf(w, i1.w, acc[7]);
                }
            }
            const float sc = __ldg(A.sc[3] + co), sh = __ldg(A.sh[3] + co);
            float* out = act4 + (co * 4 + p) * AM;
            *reinterpret_cast<float4*>(out) = make_float4(bn_relu(acc[0], sc, sh), bn_relu(acc[1], sc, sh),
                                                          bn_relu(acc[2], sc, sh), bn_relu(acc[3], sc, sh));
            *reinterpret_cast<float4*>(out + 4) = make_float4(bn_relu(acc[4], sc, sh), bn_relu(acc[5], sc, sh),
                                                              bn_relu(acc[6], sc, sh), bn_relu(acc[7], sc, sh));
        }
        __syncthreads();

        // ---- conv4 64->128 on 2x2: item = (channel group, half of the input channels) ----
        {
            const int cg = warp & 3, ks = warp >> 2;
            const int co = cg * 32 + lane;
            float acc[4][AM];
#pragma unroll
            for (int p = 0; p < 4; ++p)
#pragma unroll
                for (int a = 0; a < AM; ++a) acc[p][a] = 0.f;
            for (int ci = ks * 32; ci < ks * 32 + 32; ++ci) {
                float w[9];
#pragma unroll
                for (int t = 0; t < 9; ++t) w[t] = __ldg(A.w4t + (ci * 9 + t) * 128 + co);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float4 i0 = ld_smem4(act4 + (ci * 4 + q) * AM), i1 = ld_smem4(act4 + (ci * 4 + q) * AM + 4);
                    const float in[AM] = {i0.x, i0.y, i0.z, i0.w, i1.x, i1.y, i1.z, i1.w};
#pragma unroll
                    for (int p = 0; p < 4; ++p) {
                        const int t = ((q >> 1) - (p >> 1) + 1) * 3 + ((q & 1) - (p & 1) + 1);
#pragma unroll
                        for (int a = 0; a < AM; ++a) acc[p][a] = fmaf(w[t], in[a], acc[p][a]);
                    }
                }
            }
#pragma unroll
            for (int p = 0; p < 4; ++p) {
                float* out = part4 + ((ks * 128 + co) * 4 + p) * AM;
                *reinterpret_cast<float4*>(out) = make_float4(acc[p][0], acc[p][1], acc[p][2], acc[p][3]);
                *reinterpret_cast<float4*>(out + 4) = make_float4(acc[p][4], acc[p][5], acc[p][6], acc[p][7]);
            }
        }
        __syncthreads();
        // combine the two halves, BN + ReLU, 2x2 maxpool -> [128][a]
        for (int i = threadIdx.x; i < 128 * AM; i += FE_THREADS) {
            const int co = i / AM, a = i - co * AM;
            const float sc = __ldg(A.sc[4] + co), sh = __ldg(A.sh[4] + co);
            float m = 0.f;
#pragma unroll
            for (int p = 0; p < 4; ++p)
                m = fmaxf(m, bn_relu(part4[(co * 4 + p) * AM + a] + part4[((128 + co) * 4 + p) * AM + a], sc, sh));
            act5[i] = m;
        }
        __syncthreads();

        // ---- compress MLP 128->128 + ReLU: item = (channel group, half of the inputs) -------
        {
            const int cg = warp & 3, ks = warp >> 2;
            const int co = cg * 32 + lane;
            float acc[AM];
#pragma unroll
            for (int a = 0; a < AM; ++a) acc[a] = 0.f;
#pragma unroll 8
            for (int k = ks * 64; k < ks * 64 + 64; ++k) {
                const float w = __ldg(A.w5t + k * 128 + co);
                const float4 i0 = ld_smem4(act5 + k * AM), i1 = ld_smem4(act5 + k * AM + 4);
                acc[0] = fmaf(w, i0.x, acc[0]); acc[1] = fmaf(w, i0.y, acc[1]);
                acc[2] = fmaf(w, i0.z, acc[2]); acc[3] = fmaf(w, i0.w, acc[3]);
                acc[4] = fmaf(w, i1.x, acc[4]); acc[5] = fmaf(w, i1.y, acc[5]);
                acc[6] = fmaf(w, i1.z, acc[6]); acc[7] = fmaf(w, i1.w, acc[7]);
            }
            float* out = part5 + (ks * 128 + co) * AM;
            *reinterpret_cast<float4*>(out) = make_float4(acc[0], acc[1], acc[2], acc[3]);
            *reinterpret_cast<float4*>(out + 4) = make_float4(acc[4], acc[5], acc[6], acc[7]);
        }
        __syncthreads();
        for (int i = threadIdx.x; i < na * 128; i += FE_THREADS) {
            const int a = i >> 7, co = i & 127;
            const float v = part5[co * AM + a] + part5[(128 + co) * AM + a] + __ldg(A.b5 + co);
            A.feat[(size_t)(a0 + a) * 128 + co] = fmaxf(v, 0.f);
        }
        __syncthreads();
    }
}

// scale = gamma / sqrt(var + eps);  shift = (conv_bias - mean) * scale + beta
__global__ void fold_bn_kernel(const float* conv_b, const float* g, const float* b, const float* mean,
                               const float* var, float* sc, float* sh, int C) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= C) return;
    const float s = g[i] / sqrtf(var[i] + 1e-5f);
    sc[i] = s;
    sh[i] = (conv_b[i] - mean[i]) * s + b[i];
}

}  // namespace gpp

using namespace gpp;

static const int kConvC[6] = {3, 32, 32, 64, 64, 128};

struct gpp_planner {
    int K;
    float* arena;        // prepared weights
    size_t off_w[6];     // conv0..4 k-major, compress k-major
    size_t off_sc[5], off_sh[5];
    size_t off_b5, off_gfw, off_gfb, off_wa, off_ba, arena_floats;
    bool weights_set;
    float* raw;          // device staging for host-provided parameters
    size_t raw_floats;
    float* feat;         // [rows][128] workspace
    size_t feat_rows;
    // host-buffer path
    cudaStream_t stream;
    float* d_x;
    size_t d_x_floats;
    void* d_S;
    size_t d_S_bytes;
    float* d_logits;
    size_t d_logits_floats;
    // per-kernel event log (roofline report)
    bool profiling;
    std::vector<cudaEvent_t>* events;   // triples: start, after feature kernel, after filter kernel
    size_t events_used;
};

extern "C" const char* gpp_last_error(void) { return gpp::g_err; }
extern "C" int gpp_abi_version(void) { return 1; }
extern "C" unsigned long long gpp_launch_count(void) { return gpp::g_launches; }
extern "C" void gpp_reset_launch_count(void) { gpp::g_launches = 0; }

extern "C" int gpp_device_info(int* sms, int* cc_major, int* cc_minor) {
    int dev = 0;
    GPP_CUDA_OK(cudaGetDevice(&dev));
    cudaDeviceProp p;
    GPP_CUDA_OK(cudaGetDeviceProperties(&p, dev));
    if (sms) *sms = p.multiProcessorCount;
    if (cc_major) *cc_major = p.major;
    if (cc_minor) *cc_minor = p.minor;
    return GPP_OK;
}

extern "C" int gpp_planner_create(gpp_planner** out, int K) {
    GPP_REQUIRE(out, GPP_ERR_INVALID, "planner_create: null out pointer");
    GPP_REQUIRE(K >= 1 && K <= 8, GPP_ERR_UNSUPPORTED, "planner_create: K=%d outside [1,8]", K);
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        set_error("planner_create: no CUDA device visible");
        return GPP_ERR_NODEVICE;
    }
    gpp_planner* p = new gpp_planner();
    memset(p, 0, sizeof(*p));
    p->K = K;
    size_t off = 0;
    auto take = [&](size_t n) { size_t o = off; off += (n + 63) / 64 * 64; return o; };
    for (int l = 0; l < 5; ++l) p->off_w[l] = take((size_t)kConvC[l] * 9 * kConvC[l + 1]);
    p->off_w[5] = take(128 * 128);
    for (int l = 0; l < 5; ++l) { p->off_sc[l] = take(kConvC[l + 1]); p->off_sh[l] = take(kConvC[l + 1]); }
    p->off_b5 = take(128);
    p->off_gfw = take((size_t)K * 128 * 128);
    p->off_gfb = take(128);
    p->off_wa = take(5 * 128);
    p->off_ba = take(64);
    p->arena_floats = off;
    if (cudaMalloc(&p->arena, sizeof(float) * off) != cudaSuccess) {
        set_error("planner_create: cudaMalloc(%zu) failed", sizeof(float) * off);
        delete p;
        return GPP_ERR_CUDA;
    }
    if (cudaStreamCreateWithFlags(&p->stream, cudaStreamNonBlocking) != cudaSuccess) {
        set_error("planner_create: stream creation failed");
        cudaFree(p->arena);
        delete p;
        return GPP_ERR_CUDA;
    }
    cudaFuncSetAttribute(feature_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FE_SMEM_BYTES);
    *out = p;
    return GPP_OK;
}

extern "C" void gpp_planner_destroy(gpp_planner* p) {
    if (!p) return;
    cudaFree(p->arena);
    cudaFree(p->raw);
    cudaFree(p->feat);
    cudaFree(p->d_x);
    cudaFree(p->d_S);
    cudaFree(p->d_logits);
    if (p->stream) cudaStreamDestroy(p->stream);
    if (p->events) {
        for (cudaEvent_t e : *p->events) cudaEventDestroy(e);
        delete p->events;
    }
    delete p;
}

extern "C" int gpp_planner_set_profiling(gpp_planner* p, int enable) {
    GPP_REQUIRE(p, GPP_ERR_INVALID, "planner_set_profiling: null planner");
    p->profiling = enable != 0;
    if (enable && !p->events) p->events = new std::vector<cudaEvent_t>();
    p->events_used = 0;
    return GPP_OK;
}

extern "C" int gpp_planner_get_profile(gpp_planner* p, double* feature_ms, double* graph_filter_ms, int* steps) {
    GPP_REQUIRE(p, GPP_ERR_INVALID, "planner_get_profile: null planner");
    double fe = 0.0, gf = 0.0;
    const size_t n = p->events_used / 3;
    for (size_t i = 0; i < n; ++i) {
        float a = 0.f, b = 0.f;
        GPP_CUDA_OK(cudaEventSynchronize((*p->events)[3 * i + 2]));
        GPP_CUDA_OK(cudaEventElapsedTime(&a, (*p->events)[3 * i], (*p->events)[3 * i + 1]));
        GPP_CUDA_OK(cudaEventElapsedTime(&b, (*p->events)[3 * i + 1], (*p->events)[3 * i + 2]));
        fe += a;
        gf += b;
    }
    if (feature_ms) *feature_ms = fe;
    if (graph_filter_ms) *graph_filter_ms = gf;
    if (steps) *steps = (int)n;
    p->events_used = 0;
    return GPP_OK;
}

static cudaEvent_t next_event(gpp_planner* p) {
    if (p->events_used == p->events->size()) {
        cudaEvent_t e;
        if (cudaEventCreate(&e) != cudaSuccess) return nullptr;
        p->events->push_back(e);
    }
    return (*p->events)[p->events_used++];
}

extern "C" int gpp_planner_set_weights(gpp_planner* p, const gpp_planner_weights* w, int on_device,
                                       void* stream) {
    GPP_REQUIRE(p && w, GPP_ERR_INVALID, "planner_set_weights: null pointer");
    cudaStream_t st = on_device ? reinterpret_cast<cudaStream_t>(stream) : p->stream;
    gpp_planner_weights d = *w;
    const int K = p->K;
    if (!on_device) {
        // stage every host array into one device block, then run the same re-layout kernels
        size_t total = 0;
        for (int l = 0; l < 5; ++l) total += (size_t)kConvC[l] * 9 * kConvC[l + 1] + 5 * (size_t)kConvC[l + 1];
        total += 128 * 128 + 128 + (size_t)K * 128 * 128 + 128 + 5 * 128 + 5;
        if (p->raw_floats < total) {
            cudaFree(p->raw);
            p->raw = nullptr;
            GPP_CUDA_OK(cudaMalloc(&p->raw, sizeof(float) * total));
            p->raw_floats = total;
        }
        float* cur = p->raw;
        auto up = [&](const float*& ptr, size_t n) -> cudaError_t {
            cudaError_t e = cudaMemcpyAsync(cur, ptr, sizeof(float) * n, cudaMemcpyHostToDevice, st);
            ptr = cur;
            cur += n;
            return e;
        };
        for (int l = 0; l < 5; ++l) {
            const size_t c = kConvC[l + 1];
            GPP_CUDA_OK(up(d.conv_w[l], (size_t)kConvC[l] * 9 * c));
            GPP_CUDA_OK(up(d.conv_b[l], c));
            GPP_CUDA_OK(up(d.bn_w[l], c));
            GPP_CUDA_OK(up(d.bn_b[l], c));
            GPP_CUDA_OK(up(d.bn_mean[l], c));
            GPP_CUDA_OK(up(d.bn_var[l], c));
        }
        GPP_CUDA_OK(up(d.compress_w, 128 * 128));
        GPP_CUDA_OK(up(d.compress_b, 128));
        GPP_CUDA_OK(up(d.gf_w, (size_t)K * 128 * 128));
        GPP_CUDA_OK(up(d.gf_b, 128));
        GPP_CUDA_OK(up(d.action_w, 5 * 128));
        GPP_CUDA_OK(up(d.action_b, 5));
    }
    float* A = p->arena;
    for (int l = 0; l < 5; ++l) {
        // [Cout][Cin*9] -> [Cin*9][Cout]
        int rc = launch_transpose_taps(d.conv_w[l], A + p->off_w[l], kConvC[l + 1], kConvC[l] * 9, st);
        if (rc) return rc;
        const int C = kConvC[l + 1];
        fold_bn_kernel<<<(C + 127) / 128, 128, 0, st>>>(d.conv_b[l], d.bn_w[l], d.bn_b[l], d.bn_mean[l],
                                                         d.bn_var[l], A + p->off_sc[l], A + p->off_sh[l], C);
        GPP_LAUNCH_CHECK();
    }
    int rc = launch_transpose_taps(d.compress_w, A + p->off_w[5], 128, 128, st);
    if (rc) return rc;
    rc = launch_transpose_taps(d.gf_w, A + p->off_gfw, 128, K * 128, st);
    if (rc) return rc;
    GPP_CUDA_OK(cudaMemcpyAsync(A + p->off_b5, d.compress_b, sizeof(float) * 128, cudaMemcpyDeviceToDevice, st));
    GPP_CUDA_OK(cudaMemcpyAsync(A + p->off_gfb, d.gf_b, sizeof(float) * 128, cudaMemcpyDeviceToDevice, st));
    GPP_CUDA_OK(cudaMemcpyAsync(A + p->off_wa, d.action_w, sizeof(float) * 5 * 128, cudaMemcpyDeviceToDevice, st));
    GPP_CUDA_OK(cudaMemcpyAsync(A + p->off_ba, d.action_b, sizeof(float) * 5, cudaMemcpyDeviceToDevice, st));
    if (!on_device) GPP_CUDA_OK(cudaStreamSynchronize(st));
    p->weights_set = true;
    return GPP_OK;
}

static int ensure_feat(gpp_planner* p, size_t rows) {
    if (p->feat_rows >= rows) return GPP_OK;
    // growing the workspace must not race with kernels still using the old one
    GPP_CUDA_OK(cudaDeviceSynchronize());
    cudaFree(p->feat);
    p->feat = nullptr;
    p->feat_rows = 0;
    GPP_CUDA_OK(cudaMalloc(&p->feat, sizeof(float) * 128 * rows));
    p->feat_rows = rows;
    return GPP_OK;
}

extern "C" int gpp_planner_forward(gpp_planner* p, const float* x, const void* S, int s_is_f64,
                                   float* logits, float* features_out, int B, int N, void* stream) {
    GPP_REQUIRE(p && x && S && logits, GPP_ERR_INVALID, "planner_forward: null pointer");
    GPP_REQUIRE(p->weights_set, GPP_ERR_INVALID, "planner_forward: gpp_planner_set_weights not called");
    GPP_REQUIRE(B >= 0 && N >= 1, GPP_ERR_INVALID, "planner_forward: bad sizes B=%d N=%d", B, N);
    GPP_REQUIRE(N <= 64, GPP_ERR_UNSUPPORTED, "planner_forward: N=%d > 64 agents is outside the fused kernel's tile", N);
    if (B == 0) return GPP_OK;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const size_t rows = (size_t)B * N;
    float* feat = features_out;
    if (!feat) {
        int rc = ensure_feat(p, rows);
        if (rc) return rc;
        feat = p->feat;
    }
    const float* A = p->arena;
    FeArgs fa;
    fa.x = x; fa.feat = feat; fa.total_agents = (int)rows;
    int apt = (int)((rows + sm_count() - 1) / sm_count());
    if (apt > AM) apt = AM;
    if (apt < 1) apt = 1;
    fa.apt = apt;
    fa.num_tiles = (int)((rows + apt - 1) / apt);
    fa.w0t = A + p->off_w[0]; fa.w1t = A + p->off_w[1]; fa.w2t = A + p->off_w[2];
    fa.w3t = A + p->off_w[3]; fa.w4t = A + p->off_w[4]; fa.w5t = A + p->off_w[5];
    for (int l = 0; l < 5; ++l) { fa.sc[l] = A + p->off_sc[l]; fa.sh[l] = A + p->off_sh[l]; }
    fa.b5 = A + p->off_b5;
    const int grid = fa.num_tiles < sm_count() ? fa.num_tiles : sm_count();
    const bool prof = p->profiling && p->events && p->events_used + 3 <= 3 * 8192;
    cudaEvent_t e0 = nullptr, e1 = nullptr, e2 = nullptr;
    if (prof) {
        e0 = next_event(p); e1 = next_event(p); e2 = next_event(p);
        GPP_REQUIRE(e0 && e1 && e2, GPP_ERR_CUDA, "planner_forward: cudaEventCreate failed");
        GPP_CUDA_OK(cudaEventRecord(e0, st));
    }
    feature_kernel<<<grid, FE_THREADS, FE_SMEM_BYTES, st>>>(fa);
    GPP_LAUNCH_CHECK();
    if (prof) GPP_CUDA_OK(cudaEventRecord(e1, st));
    int rc = launch_gf_forward_fast(feat, S, s_is_f64, A + p->off_gfw, A + p->off_gfb, nullptr,
                                    A + p->off_wa, A + p->off_ba, logits, B, N, p->K, GPP_NODE_MAJOR,
                                    GPP_NODE_MAJOR, 1, st);
    if (rc) return rc;
    if (prof) GPP_CUDA_OK(cudaEventRecord(e2, st));
    return GPP_OK;
}

extern "C" int gpp_planner_forward_host(gpp_planner* p, const float* x_host, const void* S_host,
                                        int s_is_f64, float* logits_host, int B, int N) {
    GPP_REQUIRE(p && x_host && S_host && logits_host, GPP_ERR_INVALID, "planner_forward_host: null pointer");
    GPP_REQUIRE(B >= 0 && N >= 1, GPP_ERR_INVALID, "planner_forward_host: bad sizes B=%d N=%d", B, N);
    if (B == 0) return GPP_OK;
    const size_t nx = (size_t)B * N * IN_PIX;
    const size_t sb = (size_t)B * N * N * (s_is_f64 ? 8 : 4);
    const size_t nl = (size_t)B * N * 5;
    if (p->d_x_floats < nx) {
        GPP_CUDA_OK(cudaStreamSynchronize(p->stream));
        cudaFree(p->d_x); p->d_x = nullptr; p->d_x_floats = 0;
        GPP_CUDA_OK(cudaMalloc(&p->d_x, sizeof(float) * nx));
        p->d_x_floats = nx;
    }
    if (p->d_S_bytes < sb) {
        GPP_CUDA_OK(cudaStreamSynchronize(p->stream));
        cudaFree(p->d_S); p->d_S = nullptr; p->d_S_bytes = 0;
        GPP_CUDA_OK(cudaMalloc(&p->d_S, sb));
        p->d_S_bytes = sb;
    }
    if (p->d_logits_floats < nl) {
        GPP_CUDA_OK(cudaStreamSynchronize(p->stream));
        cudaFree(p->d_logits); p->d_logits = nullptr; p->d_logits_floats = 0;
        GPP_CUDA_OK(cudaMalloc(&p->d_logits, sizeof(float) * nl));
        p->d_logits_floats = nl;
    }
    GPP_CUDA_OK(cudaMemcpyAsync(p->d_x, x_host, sizeof(float) * nx, cudaMemcpyHostToDevice, p->stream));
    GPP_CUDA_OK(cudaMemcpyAsync(p->d_S, S_host, sb, cudaMemcpyHostToDevice, p->stream));
    int rc = gpp_planner_forward(p, p->d_x, p->d_S, s_is_f64, p->d_logits, nullptr, B, N, p->stream);
    if (rc) return rc;
    GPP_CUDA_OK(cudaMemcpyAsync(logits_host, p->d_logits, sizeof(float) * nl, cudaMemcpyDeviceToHost, p->stream));
    GPP_CUDA_OK(cudaStreamSynchronize(p->stream));
    return GPP_OK;
}
