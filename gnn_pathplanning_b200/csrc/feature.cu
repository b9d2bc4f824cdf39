// Feature extractor for sm_100a: per-agent CNN (5 x conv3x3+BN+ReLU, 3 max-pools) + compress
// MLP, all B*N agents tiled 8 at a time onto persistent 16-warp CTAs.
//
// Replaces the N sequential per-agent ConvLayers/compressMLP calls of
// DecentralPlannerNet.forward (/root/reference/graphs/models/decentralplanner.py:284-290) in
// eval mode.  Activations never leave shared memory between layers.
//
// Mapping (fp32 FMA): lanes own output channels; activations are read from shared memory as
// warp-broadcast float4 rows.  The filters of conv1..conv4 and the compress MLP (615 KB per tile,
// L2-resident) stream through a 4-slot shared-memory ring: one thread issues 1-D bulk async copies
// (cp.async.bulk -> mbarrier complete_tx) three chunks ahead of the 16 warps that consume them, so the
// L2 latency of the filter fetch is off the critical path of every layer.  conv3/conv4/linear keep all
// agents of the tile in registers (templated on the agent count) and split the input channels over
// warps; partial sums meet in shared memory.
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
constexpr int FE_ACT_FLOATS = OFF_ACT5 + 128 * AM;
// filter ring: every chunk is <= 18 KB of k-major filter rows (see WStream::issue)
constexpr int RING_SLOTS = 4;
constexpr int SLOT_FLOATS = 4608;
constexpr int OFF_RING = FE_ACT_FLOATS;
constexpr int OFF_BARS = OFF_RING + RING_SLOTS * SLOT_FLOATS;   // RING_SLOTS x uint64 "slot filled" barriers
constexpr int FE_SMEM_FLOATS = OFF_BARS + 2 * RING_SLOTS;
static_assert((OFF_RING % 4) == 0 && (OFF_BARS % 2) == 0, "ring / barrier alignment");
static_assert(SZ_PART4 <= SZ_IN0 + SZ_ACT12, "PART4 must fit in the IN0+ACT1 region");
static_assert(SZ_PART3 + SZ_PART5 <= SZ_ACT12, "PART3+PART5 must fit in the ACT2 region");
constexpr size_t FE_SMEM_BYTES = sizeof(float) * FE_SMEM_FLOATS;

// debug phase timer: thread 0 of block 0 adds the cycles since the previous mark to timing[i]
#define FE_MARK(i)                                                          \
    if (timed) {                                                            \
        const long long tn = clock64();                                     \
        atomicAdd(&A.timing[i], (unsigned long long)(tn - tprev));          \
        tprev = tn;                                                         \
    }

__device__ __forceinline__ float bn_relu(float v, float sc, float sh) {
    return fmaxf(fmaf(v, sc, sh), 0.f);
}

// ---------------------------------------------------------------------------------------------------
// Filter stream of one agent tile.  Chunk sequence: conv1 (4 chunks of 8 input channels), conv2 (8
// chunks of 4 input channels per pass over the items), conv3 (8 chunks: 4 input channels of each half),
// conv4 (16 chunks: 1 input channel of each quarter), compress MLP (4 chunks: 8 inputs of each quarter).
//
// Every thread tracks the (uniform) counters; thread 0 issues the copies.
// (A cluster-multicast variant of this stream - one L2 read per 2 or 4 CTAs - was measured 20 % slower than
// per-CTA copies at every batch size, see profiles/README.md, and is not kept.)
// ---------------------------------------------------------------------------------------------------
struct WStream {
    float* ring;
    uint64_t* full;
    uint32_t cons;      // chunks consumed since kernel start: slot = cons % RING_SLOTS, parity = (cons / RING_SLOTS) & 1
    uint32_t prod;      // chunks issued since kernel start
    int seq, seq_len;   // next chunk to issue / number of chunks of the current tile
    int n2;             // conv2 chunks of the current tile (8 per pass)

    __device__ __forceinline__ void init(float* sm_base) {
        ring = sm_base + OFF_RING;
        full = reinterpret_cast<uint64_t*>(sm_base + OFF_BARS);
        cons = 0; prod = 0; seq = 0; seq_len = 0; n2 = 0;
        if (threadIdx.x == 0) {
            for (int i = 0; i < RING_SLOTS; ++i) mbar_init(full + i, 1);
            fence_mbar_init();
        }
        __syncthreads();
    }
    __device__ __forceinline__ void issue(const FeArgs& A) {
        if (threadIdx.x == 0) {
            const int slot = prod % RING_SLOTS;
            float* dst = ring + slot * SLOT_FLOATS;
            uint64_t* bar = full + slot;
            int i = seq;
            if (i < 4) {
                mbar_arrive_expect_tx(bar, 9216u);
                bulk_g2s(dst, A.w1t + i * 2304, 9216u, bar);
            } else if ((i -= 4) < n2) {
                mbar_arrive_expect_tx(bar, 9216u);
                bulk_g2s(dst, A.w2t + (i & 7) * 2304, 9216u, bar);
            } else if ((i -= n2) < 8) {
                mbar_arrive_expect_tx(bar, 18432u);
#pragma unroll
                for (int ks = 0; ks < 2; ++ks)
                    bulk_g2s(dst + ks * 2304, A.w3t + (ks * 32 + 4 * i) * 576, 9216u, bar);
            } else if ((i -= 8) < 16) {
                mbar_arrive_expect_tx(bar, 18432u);
#pragma unroll
                for (int ks = 0; ks < 4; ++ks)
                    bulk_g2s(dst + ks * 1152, A.w4t + (ks * 16 + i) * 1152, 4608u, bar);
            } else {
                i -= 16;
                mbar_arrive_expect_tx(bar, 16384u);
#pragma unroll
                for (int ks = 0; ks < 4; ++ks)
                    bulk_g2s(dst + ks * 1024, A.w5t + (ks * 32 + 8 * i) * 128, 4096u, bar);
            }
        }
        ++prod;
        ++seq;
    }
    // start of a tile: the ring is empty (every chunk of the previous tile was consumed)
    __device__ __forceinline__ void begin_tile(const FeArgs& A, int conv2_passes) {
        n2 = 8 * conv2_passes;
        seq = 0;
        seq_len = 4 + n2 + 8 + 16 + 4;
#pragma unroll 1
        for (int i = 0; i < RING_SLOTS; ++i) issue(A);
    }
    // blocks until the next chunk has landed; returns its slot
    __device__ __forceinline__ const float* acquire() const {
        const uint32_t slot = cons % RING_SLOTS;
        mbar_wait(full + slot, (cons / RING_SLOTS) & 1u);
        return ring + slot * SLOT_FLOATS;
    }
    // all warps are done with the chunk: its slot is refilled with the chunk RING_SLOTS further on
    __device__ __forceinline__ void release(const FeArgs& A) {
        __syncthreads();
        ++cons;
        if (seq < seq_len) issue(A);
    }
};

// conv1 (32->32 on 5x5), 8 input channels of one chunk: NR output rows starting at r0 of agent a
template <int NR>
__device__ __forceinline__ void conv1_chunk(float (&acc)[3][5], const float* __restrict__ act1,
                                            const float* __restrict__ ws, int a, int r0, int ci0, int lane) {
#pragma unroll 2
    for (int cl = 0; cl < 8; ++cl) {
        float w[9];
#pragma unroll
        for (int t = 0; t < 9; ++t) w[t] = ws[(cl * 9 + t) * 32 + lane];
        const float* base = act1 + ((a * 32 + ci0 + cl) * 7 + r0) * 8;
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
}

// conv2 (32->64 on 5x5, pooled to 2x2), 4 input channels of one chunk: conv rows 2py, 2py+1 and either all 4
// conv columns (HALF = false, px ignored) or columns 2px, 2px+1 (HALF = true) of agent a, output channel co
template <bool HALF>
__device__ __forceinline__ void conv2_chunk(float (&acc)[2][4], const float* __restrict__ act2,
                                            const float* __restrict__ ws, int a, int py, int px, int ci0, int co) {
#pragma unroll
    for (int cl = 0; cl < 4; ++cl) {
        float w[9];
#pragma unroll
        for (int t = 0; t < 9; ++t) w[t] = ws[(cl * 9 + t) * 64 + co];
        const float* base = act2 + ((a * 32 + ci0 + cl) * 7 + 2 * py) * 8;
#pragma unroll
        for (int iy = 0; iy < 4; ++iy) {
            if (HALF) {
                const float2 u = *reinterpret_cast<const float2*>(base + iy * 8 + 2 * px);
                const float2 v = *reinterpret_cast<const float2*>(base + iy * 8 + 2 * px + 2);
                const float row[4] = {u.x, u.y, v.x, v.y};
#pragma unroll
                for (int oy = 0; oy < 2; ++oy) {
                    const int ky = iy - oy;
                    if (ky < 0 || ky > 2) continue;
#pragma unroll
                    for (int kx = 0; kx < 3; ++kx)
#pragma unroll
                        for (int ox = 0; ox < 2; ++ox)
                            acc[oy][ox] = fmaf(w[ky * 3 + kx], row[ox + kx], acc[oy][ox]);
                }
            } else {
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
    }
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
__device__ __forceinline__ void tail_layers(const FeArgs& A, WStream& wsm, float* sm, int warp, int lane, int a0,
                                            int na, bool timed, long long& tprev) {
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
        for (int j = 0; j < 8; ++j) {
            const float* ws = wsm.acquire() + ks * 2304;
#pragma unroll
            for (int cl = 0; cl < 4; ++cl) {
                const int ci = ks * 32 + 4 * j + cl;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float w = ws[cl * 576 + toff[q]];
                    float in[NA];
                    load_agents<NA>(act3 + (ci * 4 + q) * AM, in);
#pragma unroll
                    for (int a = 0; a < NA; ++a) acc[a] = fmaf(w, in[a], acc[a]);
                }
            }
            wsm.release(A);
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
    FE_MARK(4)

    // ---- conv4 64->128 on 2x2: item = (channel group, quarter of the input channels) ----------
    {
        const int cg = warp & 3, ks = warp >> 2;
        const int co = cg * 32 + lane;
        float acc[4][NA];
#pragma unroll
        for (int p = 0; p < 4; ++p)
#pragma unroll
            for (int a = 0; a < NA; ++a) acc[p][a] = 0.f;
        for (int j = 0; j < 16; ++j) {
            const float* ws = wsm.acquire() + ks * 1152 + co;
            const int ci = ks * 16 + j;
            float w[9];
#pragma unroll
            for (int t = 0; t < 9; ++t) w[t] = ws[t * 128];
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
            wsm.release(A);
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
    FE_MARK(5)

    // ---- compress MLP 128->128 + ReLU: item = (channel group, quarter of the inputs) ----------
    {
        const int cg = warp & 3, ks = warp >> 2;
        const int co = cg * 32 + lane;
        float acc[NA];
#pragma unroll
        for (int a = 0; a < NA; ++a) acc[a] = 0.f;
        for (int j = 0; j < 4; ++j) {
            const float* ws = wsm.acquire() + ks * 1024 + co;
#pragma unroll
            for (int jl = 0; jl < 8; ++jl) {
                const float w = ws[jl * 128];
                float in[NA];
                load_agents<NA>(act5 + (ks * 32 + 8 * j + jl) * AM, in);
#pragma unroll
                for (int a = 0; a < NA; ++a) acc[a] = fmaf(w, in[a], acc[a]);
            }
            wsm.release(A);
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
    const bool timed = A.timing && blockIdx.x == 0 && threadIdx.x == 0;
    long long tprev = timed ? clock64() : 0;

    WStream wsm;
    wsm.init(sm);

    for (int tile = blockIdx.x; tile < A.num_tiles; tile += gridDim.x) {
        const int a0 = tile * A.apt;
        const int na = min(A.apt, A.total_agents - a0);
        // conv2 work items: (agent, channel group, pooled row) or, when that fills the 16 warps badly, also
        // split by pooled column
        const bool c2_half = (A.apt <= 2) || A.apt == 5 || A.apt == 6;
        const int c2_items = na * (c2_half ? 8 : 4);
        const int c2_passes = (A.apt * (c2_half ? 8 : 4) + FE_WARPS - 1) / FE_WARPS;
        wsm.begin_tile(A, c2_passes);   // the first filter chunks land while the inputs are staged

        // ---- zero the bordered activation buffers (their interiors / aliases are rewritten
        //      every tile) and stage the binary FOV tensors into the zero-bordered layout --------
        {
            float4* z4 = reinterpret_cast<float4*>(sm + OFF_ACT1);      // ACT1, ACT2, ACT3, ACT4, ACT5
            constexpr int n4 = (FE_ACT_FLOATS - OFF_ACT1) / 4;
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
        FE_MARK(0)

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
        FE_MARK(1)

        // ---- conv1 32->32 on 5x5 (+BN+ReLU): one item per warp = some output rows of one agent, split
        //      so that na agents give at most 16 items (rows {0-2, 3-4}, {0-1, 2-3, 4} or single rows) ----
        {
            int a = 0, r0 = 0, nr = 0;
            if (na >= 6) {
                if (warp < na * 2) { a = warp >> 1; r0 = (warp & 1) * 3; nr = 3 - (warp & 1); }
            } else if (na >= 4) {
                if (warp < na * 3) { a = warp / 3; const int g = warp - a * 3; r0 = g * 2; nr = g == 2 ? 1 : 2; }
            } else {
                if (warp < na * 5) { a = warp / 5; r0 = warp - a * 5; nr = 1; }
            }
            float acc[3][5];
#pragma unroll
            for (int i = 0; i < 3; ++i)
#pragma unroll
                for (int j = 0; j < 5; ++j) acc[i][j] = 0.f;
            for (int j = 0; j < 4; ++j) {
                const float* ws = wsm.acquire();
                if (nr == 3) conv1_chunk<3>(acc, act1, ws, a, r0, j * 8, lane);
                else if (nr == 2) conv1_chunk<2>(acc, act1, ws, a, r0, j * 8, lane);
                else if (nr == 1) conv1_chunk<1>(acc, act1, ws, a, r0, j * 8, lane);
                wsm.release(A);
            }
            float* out = act2 + ((a * 32 + lane) * 7 + r0 + 1) * 8 + 1;
#pragma unroll
            for (int oy = 0; oy < 3; ++oy)
                if (oy < nr) {
#pragma unroll
                    for (int ox = 0; ox < 5; ++ox) out[oy * 8 + ox] = bn_relu(acc[oy][ox], sc1, sh1);
                }
        }
        __syncthreads();
        FE_MARK(2)

        // ---- conv2 32->64 on 5x5 (+BN+ReLU) + maxpool2 -> 64 x 2x2 ------------------------------
        for (int pass = 0; pass < c2_passes; ++pass) {
            const int item = pass * FE_WARPS + warp;
            const bool live = item < c2_items;
            const int nad = na > 0 ? na : 1;
            const int a = item % nad, r = item / nad;
            const int cg = r & 1, py = (r >> 1) & 1, px = r >> 2;
            const int co = cg * 32 + lane;
            float acc[2][4];
#pragma unroll
            for (int i = 0; i < 2; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
            for (int j = 0; j < 8; ++j) {
                const float* ws = wsm.acquire();
                if (live) {
                    if (c2_half) conv2_chunk<true>(acc, act2, ws, a, py, px, j * 4, co);
                    else conv2_chunk<false>(acc, act2, ws, a, py, 0, j * 4, co);
                }
                wsm.release(A);
            }
            if (live) {
                const float sc = __ldg(A.sc[2] + co), sh = __ldg(A.sh[2] + co);
                if (c2_half) {
                    const float m0 = fmaxf(bn_relu(acc[0][0], sc, sh), bn_relu(acc[0][1], sc, sh));
                    const float m1 = fmaxf(bn_relu(acc[1][0], sc, sh), bn_relu(acc[1][1], sc, sh));
                    act3[(co * 4 + py * 2 + px) * AM + a] = fmaxf(m0, m1);
                } else {
#pragma unroll
                    for (int qx = 0; qx < 2; ++qx) {
                        const float m0 = fmaxf(bn_relu(acc[0][2 * qx], sc, sh), bn_relu(acc[0][2 * qx + 1], sc, sh));
                        const float m1 = fmaxf(bn_relu(acc[1][2 * qx], sc, sh), bn_relu(acc[1][2 * qx + 1], sc, sh));
                        act3[(co * 4 + py * 2 + qx) * AM + a] = fmaxf(m0, m1);
                    }
                }
            }
        }
        __syncthreads();
        FE_MARK(3)

        if (na <= 2)
            tail_layers<2>(A, wsm, sm, warp, lane, a0, na, timed, tprev);
        else if (na <= 4)
            tail_layers<4>(A, wsm, sm, warp, lane, a0, na, timed, tprev);
        else if (na <= 6)
            tail_layers<6>(A, wsm, sm, warp, lane, a0, na, timed, tprev);
        else
            tail_layers<8>(A, wsm, sm, warp, lane, a0, na, timed, tprev);
        __syncthreads();
        FE_MARK(6)
    }
}

static unsigned long long* g_fe_timing = nullptr;  // GPP_FE_TIMING debug counters
int debug_feature_timing(unsigned long long* out7) {
    if (!g_fe_timing) return GPP_ERR_INVALID;
    cudaDeviceSynchronize();
    cudaMemcpy(out7, g_fe_timing, 56, cudaMemcpyDeviceToHost);
    cudaMemset(g_fe_timing, 0, 64);
    return GPP_OK;
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
    fa.timing = nullptr;
    if (getenv("GPP_FE_TIMING")) {
        if (!g_fe_timing) {
            GPP_CUDA_OK(cudaMalloc(&g_fe_timing, 64));
            GPP_CUDA_OK(cudaMemset(g_fe_timing, 0, 64));
        }
        fa.timing = g_fe_timing;
    }
    const int grid = fa.num_tiles < sm_count() ? fa.num_tiles : sm_count();
    feature_kernel<<<grid, FE_THREADS, FE_SMEM_BYTES, st>>>(fa);
    GPP_LAUNCH_CHECK();
    return GPP_OK;
}

}  // namespace gpp
