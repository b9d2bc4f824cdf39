// Tensor-core (tcgen05 / TMEM) fused graph filter for sm_100a.
//
// Same contract as gf_fwd_kernel (graph_filter.cu; BatchLSIGF, /root/reference/utils/graphUtils/
// graphML.py:2273-2367, + ReLU + the 128->5 action MLP), node-major input, G = F = 128.
//
// Algebra: the graph shift acts on the node index and the taps on the feature index, so they commute:
//     y = sum_k (S^k)^T-shift( x . W_k^T )        instead of        sum_k ( (S^k)^T-shift x ) . W_k^T
// "filter, then propagate".  The dense part  U = X[128 rows,128] . [W_0^T | ... | W_{K-1}^T]  is ONE
// GEMM whose A operand is the raw x tile (no per-tap operand to build) and whose K accumulators
// (128 TMEM columns each) are the per-tap products; the N-term neighbour sums run afterwards on the
// outputs, straight out of TMEM/shared memory, on the CUDA cores.
//
//   * fp32 parity through split precision (3xTF32): operands are split into hi = tf32(v), lo = tf32(v-hi);
//     U += A_hi.B_hi + A_lo.B_hi + A_hi.B_lo, fp32 accumulation in TMEM (dropped lo.lo term: 2^-22).
//   * B operand (taps): pre-split, pre-swizzled 32 KB chunk images (hi | lo) prepared once per weight
//     update, so a linear copy lands them in the canonical SWIZZLE_128B K-major layout; a 4-stage ring
//     of 16-byte cp.async (LDGSTS) copies runs one chunk ahead of the MMAs.
//   * A operand: each 128 x 32 chunk of x is read once from global memory (coalesced float4), split and
//     stored swizzled into a 2-stage ring; it feeds K chunk-MMAs (one per tap accumulator).
//   * warp-specialised: 8 producer warps stage operands and arrive on "full" mbarriers; a 9th warp's
//     elected lane issues tcgen05.mma and releases ring stages with tcgen05.commit -> "done" mbarriers.
//   * epilogue: U_1.. are copied TMEM -> shared memory (over the idle rings), then every thread owns one
//     node row x 64 features: U_0 from TMEM + bias + sum_k sum_m S^k[m,n] U_k[m,:], ReLU, optional y
//     store, and the in-thread 128->5 action reduction (two half-rows combined through shared memory).
#include "common.cuh"
#include "tc_common.cuh"

#include <stdlib.h>

namespace gpp {

constexpr int TC_THREADS = 256;
constexpr int TC_M = 128;                  // node rows per tile (UMMA M)
constexpr int TC_C = 128;                  // G = F = 128 (UMMA N per tap)
constexpr int TC_CHUNK_K = 32;             // reduction elements per chunk (128 bytes of tf32)
constexpr int TC_NCC = TC_C / TC_CHUNK_K;  // 4 A chunks per tile
constexpr int TC_OP_BYTES = TC_M * 128;    // one operand half (hi or lo) of one chunk: 16 KB
constexpr int TC_STAGE_BYTES = 2 * TC_OP_BYTES;   // hi | lo
constexpr int TC_NA = 2, TC_NB = 4;        // ring depths
constexpr int TC_RING_BYTES = (TC_NA + TC_NB) * TC_STAGE_BYTES;   // 192 KB
constexpr int TC_MAX_K = 4;                // K accumulators x 128 columns <= 512 TMEM columns
constexpr int TC_ACT = 5;
constexpr uint32_t TC_IDESC = umma_idesc_tf32(TC_M, TC_C);

// ---------------------------------------------------------------------------------------
// B-operand images: img[c] = { hi[128 x 32] , lo[128 x 32] } in SWIZZLE_128B K-major layout; chunk
// c = k*4 + cc holds w[f][k][g] for g in [32cc, 32cc+32)  (w = the module's [F,1,K,G] taps).
// ---------------------------------------------------------------------------------------
__global__ void prep_umma_taps_kernel(const float* __restrict__ w, float* __restrict__ img, int KG) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;     // one (f, kg) element each
    if (idx >= TC_C * KG) return;
    const int f = idx / KG, kg = idx - f * KG;
    const int c = kg / TC_CHUNK_K, kk = kg - c * TC_CHUNK_K;
    float hi, lo;
    split_tf32(w[idx], hi, lo);
    const uint32_t off = sw128_offset(f, kk >> 2) / 4 + (kk & 3);
    float* base = img + (size_t)c * (TC_STAGE_BYTES / 4);
    base[off] = hi;
    base[TC_OP_BYTES / 4 + off] = lo;
}

struct GfTcArgs {
    const float* x;       // [B*N][128] node-major
    const void* S;
    const float* wimg;    // chunk images
    const float* bias;    // [128] or null
    float* y;             // [B*N][128] node-major or null
    const float* wa;      // [5][128] or null
    const float* ba;
    float* logits;        // [N][B][5]
    int B, N, K, TS, num_tiles;
    int s_is_f64, relu;
    unsigned long long* timing;   // optional [8] cycle counters (debug), null in production
};

struct GfTcSmem {
    uint32_t sk_floats;
    __host__ __device__ GfTcSmem(int N, int K, int TS) {
        sk_floats = (uint32_t)(((K > 1 ? (K - 1) : 0) * TS * N * N + 3) / 4 * 4);
    }
    __host__ __device__ uint32_t a_stage(int s) const { return (uint32_t)s * TC_STAGE_BYTES; }
    __host__ __device__ uint32_t b_stage(int s) const { return (uint32_t)(TC_NA + s) * TC_STAGE_BYTES; }
    __host__ __device__ uint32_t sk_off() const { return TC_RING_BYTES; }
    __host__ __device__ uint32_t misc_off() const { return sk_off() + sk_floats * 4; }   // bias, wa, ba, plog
    __host__ __device__ uint32_t bar_off() const {
        return misc_off() + (TC_C + TC_ACT * TC_C + 8 + TC_M * TC_ACT) * 4;
    }
    __host__ __device__ uint32_t total() const { return bar_off() + 256 + 1024; }          // + alignment slack
};

// word offset of float4 `f4` of row `row` in a [128][128] fp32 tile whose float4 columns are XOR-swizzled by
// the row (conflict-free for row-per-lane writes and for same-row broadcast reads)
__device__ __forceinline__ int usm_off(int row, int f4) { return row * TC_C + ((f4 ^ (row & 31)) << 2); }

// One arrival per producer WARP (barrier count = 8) and one polling lane per warp: 256 threads arriving on /
// polling the same mbarrier serialise in the shared-memory atomic unit.
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    __syncwarp();
    if ((threadIdx.x & 31) == 0)
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait_warp(uint64_t* bar, uint32_t parity) {
    if ((threadIdx.x & 31) == 0) mbar_wait(bar, parity);
    __syncwarp();
}
__device__ __forceinline__ void producers_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

// Warp roles: warps 0-7 (256 threads) = producers (B cp.async, A split/store) and epilogue;
// warp 8 = MMA issuer (one elected lane).  Producers and the issuer only meet on mbarriers, so operand
// staging for the next chunks overlaps the tensor-core work of the current one.
__global__ void __launch_bounds__(TC_THREADS + 32, 1) gf_fwd_tc_kernel(const GfTcArgs a) {
    extern __shared__ unsigned char smem_raw_tc[];
    const uint32_t raw = smem_u32(smem_raw_tc);
    unsigned char* sm = smem_raw_tc + (((raw + 1023u) & ~1023u) - raw);      // 1024-byte aligned base
    const GfTcSmem L(a.N, a.K, a.TS);
    float* usm = reinterpret_cast<float*>(sm);                 // epilogue alias of the rings
    float* sk = reinterpret_cast<float*>(sm + L.sk_off());
    float* bias_s = reinterpret_cast<float*>(sm + L.misc_off());
    float* wa_s = bias_s + TC_C;
    float* ba_s = wa_s + TC_ACT * TC_C;
    float* plog = ba_s + 8;                                    // [128][5] partial logits of the upper half-rows
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + L.bar_off());
    uint64_t* doneA = bars;                       // [TC_NA] MMAs that read the A stage have completed
    uint64_t* doneB = bars + TC_NA;               // [TC_NB] MMAs that read the B stage have completed
    uint64_t* fullA = bars + TC_NA + TC_NB;       // [TC_NA] all producers stored their part of the A chunk
    uint64_t* fullB = fullA + TC_NA;              // [TC_NB] all producers' cp.async data of the B chunk landed
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(fullB + TC_NB);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int N = a.N, K = a.K;
    const int nitems = TC_NCC * K;

    if (tid == 0) {
        for (int i = 0; i < TC_NA + TC_NB; ++i) mbar_init(&bars[i], 1);
        for (int i = 0; i < TC_NA + TC_NB; ++i) mbar_init(&fullA[i], TC_THREADS / 32);
        fence_mbar_init();
    }
    if (warp == 0) tmem_alloc<512>(tmem_slot);
    for (int i = tid; i < TC_C; i += blockDim.x) bias_s[i] = a.bias ? a.bias[i] : 0.f;
    if (a.wa) {
        for (int i = tid; i < TC_ACT * TC_C; i += blockDim.x) wa_s[i] = a.wa[i];
        if (tid < TC_ACT) ba_s[tid] = a.ba[tid];
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_acc = *tmem_slot;

    if (warp == TC_THREADS / 32) {
        // =============================== MMA issuer ===============================
        if (lane == 0) {
            uint32_t gq = 0, ga = 0;
            for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x) {
                for (int it = 0; it < nitems; ++it, ++gq) {
                    const int cc = it / K, k = it - cc * K;
                    const int as = ga & 1, bs = gq % TC_NB;
                    if (k == 0) mbar_wait(&fullA[as], (ga >> 1) & 1);
                    mbar_wait(&fullB[bs], (gq / TC_NB) & 1);
                    tcgen05_fence_after();
                    const uint32_t sa = smem_u32(sm + L.a_stage(as)), sb = smem_u32(sm + L.b_stage(bs));
                    const uint64_t a_hi = umma_desc_sw128(sa), a_lo = umma_desc_sw128(sa + TC_OP_BYTES);
                    const uint64_t b_hi = umma_desc_sw128(sb), b_lo = umma_desc_sw128(sb + TC_OP_BYTES);
                    const uint32_t acc = tmem_acc + (uint32_t)k * TC_C;
#pragma unroll
                    for (int ks = 0; ks < TC_CHUNK_K / 8; ++ks)     // +32 bytes (2 x 16 B units) per K = 8 step
                        umma_tf32(acc, a_hi + 2 * ks, b_hi + 2 * ks, TC_IDESC, (cc | ks) != 0);
#pragma unroll
                    for (int ks = 0; ks < TC_CHUNK_K / 8; ++ks)
                        umma_tf32(acc, a_lo + 2 * ks, b_hi + 2 * ks, TC_IDESC, 1u);
#pragma unroll
                    for (int ks = 0; ks < TC_CHUNK_K / 8; ++ks)
                        umma_tf32(acc, a_hi + 2 * ks, b_lo + 2 * ks, TC_IDESC, 1u);
                    umma_commit(&doneB[bs]);
                    if (k == K - 1) {
                        umma_commit(&doneA[as]);
                        ++ga;
                    }
                }
            }
        }
        __syncwarp();
    } else {
        // =============================== producers + epilogue ===============================
        uint32_t gq = 0;    // B chunks of earlier tiles
        uint32_t ga = 0;    // A chunks produced so far
        const int j = tid & 7;      // 16-byte column of this thread inside a 128-byte operand row

        auto issue_b = [&](uint32_t q, int c) {
            const int bs = q % TC_NB;
            const uint32_t use = q / TC_NB;
            if (use >= 1) mbar_wait_warp(&doneB[bs], (use - 1) & 1);
            const char* src = reinterpret_cast<const char*>(a.wimg) + (size_t)c * TC_STAGE_BYTES;
            const uint32_t dst = smem_u32(sm + L.b_stage(bs));
#pragma unroll
            for (int i = 0; i < TC_STAGE_BYTES / (16 * TC_THREADS); ++i) {
                const int u = tid + i * TC_THREADS;
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + u * 16), "l"(src + (size_t)u * 16)
                             : "memory");
            }
        };

        unsigned long long tacc[6] = {0, 0, 0, 0, 0, 0};
        for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x) {
            const int s0 = tile * a.TS;
            const int ns = min(a.TS, a.B - s0);
            const int R = ns * N;
            const size_t row0 = (size_t)s0 * N;
            const long long t0 = clock64();

            // the first two B chunks and the first x chunk of this tile go out before anything else
            issue_b(gq, 0);
            asm volatile("cp.async.commit_group;" ::: "memory");
            if (nitems > 1) issue_b(gq + 1, (K > 1) ? TC_NCC : 1);      // item 1 = (cc 0, k 1) or (cc 1, k 0)
            asm volatile("cp.async.commit_group;" ::: "memory");
            float4 xn[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int r = (tid >> 3) + 32 * i;
                xn[i] = (r < R) ? __ldg(reinterpret_cast<const float4*>(a.x + (row0 + r) * TC_C + j * 4))
                                : make_float4(0.f, 0.f, 0.f, 0.f);
            }

            // ---- 4 A chunks x K taps: stage operands; the issuer warp runs U_k += X[:,chunk] . W_k[:,chunk]^T
            for (int it = 0; it < nitems; ++it) {
                const int cc = it / K, k = it - cc * K;
                if (it + 2 < nitems) {                       // B runs two chunks ahead
                    const int cc2 = (it + 2) / K, k2 = (it + 2) - cc2 * K;
                    issue_b(gq + it + 2, k2 * TC_NCC + cc2);
                }
                asm volatile("cp.async.commit_group;" ::: "memory");   // (possibly empty) keeps the group count uniform
                if (k == 0) {
                    const int as = ga & 1;
                    const uint32_t ause = ga >> 1;
                    if (ause >= 1) mbar_wait_warp(&doneA[as], (ause - 1) & 1);
                    float4 xv[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) xv[i] = xn[i];
                    if (cc + 1 < TC_NCC) {
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const int r = (tid >> 3) + 32 * i;
                            xn[i] = (r < R) ? __ldg(reinterpret_cast<const float4*>(a.x + (row0 + r) * TC_C +
                                                                                     (cc + 1) * TC_CHUNK_K + j * 4))
                                            : make_float4(0.f, 0.f, 0.f, 0.f);
                        }
                    }
                    unsigned char* stage = sm + L.a_stage(as);
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int r = (tid >> 3) + 32 * i;
                        float4 hi, lo;
                        split_tf32(xv[i].x, hi.x, lo.x);
                        split_tf32(xv[i].y, hi.y, lo.y);
                        split_tf32(xv[i].z, hi.z, lo.z);
                        split_tf32(xv[i].w, hi.w, lo.w);
                        const uint32_t off = sw128_offset(r, j);
                        *reinterpret_cast<float4*>(stage + off) = hi;
                        *reinterpret_cast<float4*>(stage + TC_OP_BYTES + off) = lo;
                    }
                    fence_proxy_async_smem();          // st.shared data -> visible to the tensor core
                    mbar_arrive(&fullA[as]);
                    ++ga;
                }
                asm volatile("cp.async.wait_group 2;" ::: "memory");   // this item's B chunk (3 groups back) landed
                fence_proxy_async_smem();
                mbar_arrive(&fullB[(gq + it) % TC_NB]);
                if (it == 1) {
                    // GSO tile and its powers: only the epilogue needs them, so they are built here, in the
                    // shadow of the first chunk-MMAs, not in front of them
                if (K > 1) {
                    const size_t off = (size_t)s0 * N * N;
                    const int cnt = ns * N * N;
                    if (a.s_is_f64) {
                        const double* Sd = reinterpret_cast<const double*>(a.S) + off;
                        for (int i = tid; i < cnt; i += TC_THREADS) sk[i] = static_cast<float>(Sd[i]);
                    } else {
                        const float* Sf = reinterpret_cast<const float*>(a.S) + off;
                        for (int i = tid; i < cnt; i += TC_THREADS) sk[i] = Sf[i];
                    }
                    for (int p = 1; p < K - 1; ++p) {           // sk[p] = S^(p+1) = S^p . S
                        producers_sync();
                        const float* prev = sk + (size_t)(p - 1) * a.TS * N * N;
                        float* cur = sk + (size_t)p * a.TS * N * N;
                        for (int i = tid; i < cnt; i += TC_THREADS) {
                            const int b = i / (N * N), rem = i - b * N * N;
                            const int rr = rem / N, c = rem - rr * N;
                            float acc = 0.f;
                            for (int m = 0; m < N; ++m)
                                acc = fmaf(prev[b * N * N + rr * N + m], sk[b * N * N + m * N + c], acc);
                            cur[i] = acc;
                        }
                    }
                }
                }
            }
            gq += nitems;
            const long long t1 = clock64();

            // ---- epilogue ---------------------------------------------------------------------------
            {
                const uint32_t last = gq - 1;
                mbar_wait_warp(&doneB[last % TC_NB], (last / TC_NB) & 1);     // every MMA of this tile has completed
                tcgen05_fence_after();
                const long long t2 = clock64();
                const int r = (warp & 3) * 32 + lane;        // TMEM lane = node row of the tile
                const int h = warp >> 2;                     // which 64-feature half of the row
                const uint32_t lane_addr = tmem_acc + ((uint32_t)((warp & 3) * 32) << 16);
                // U_1 .. U_{K-1}: TMEM -> shared memory (aliases the idle operand rings)
                for (int k = 1; k < K; ++k) {
                    float* up = usm + (size_t)(k - 1) * TC_M * TC_C;
#pragma unroll 1
                    for (int cb = 0; cb < 2; ++cb) {
                        float v[32];
                        tmem_ld_32x32(lane_addr + k * TC_C + h * 64 + cb * 32, v);
#pragma unroll
                        for (int i = 0; i < 8; ++i)
                            *reinterpret_cast<float4*>(up + usm_off(r, h * 16 + cb * 8 + i)) =
                                make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
                    }
                }
                float acc[64];
                {
                    float v[32];
                    tmem_ld_32x32(lane_addr + h * 64, v);
#pragma unroll
                    for (int i = 0; i < 32; ++i) acc[i] = v[i] + bias_s[h * 64 + i];
                    tmem_ld_32x32(lane_addr + h * 64 + 32, v);
#pragma unroll
                    for (int i = 0; i < 32; ++i) acc[32 + i] = v[i] + bias_s[h * 64 + 32 + i];
                }
                tcgen05_fence_before();
                producers_sync();
                const long long t3 = clock64();
                if (r < R) {
                    const int b = r / N, n = r - b * N;
                    for (int k = 1; k < K; ++k) {
                        const float* sp = sk + ((size_t)(k - 1) * a.TS + b) * N * N + n;
                        const float* up = usm + (size_t)(k - 1) * TC_M * TC_C;
                        for (int m = 0; m < N; ++m) {
                            const float s = sp[m * N];
                            const int row = b * N + m;
#pragma unroll
                            for (int i = 0; i < 16; ++i) {
                                const float4 u = ld_smem4(up + usm_off(row, h * 16 + i));
                                acc[4 * i] = fmaf(s, u.x, acc[4 * i]);
                                acc[4 * i + 1] = fmaf(s, u.y, acc[4 * i + 1]);
                                acc[4 * i + 2] = fmaf(s, u.z, acc[4 * i + 2]);
                                acc[4 * i + 3] = fmaf(s, u.w, acc[4 * i + 3]);
                            }
                        }
                    }
                    if (a.relu) {
#pragma unroll
                        for (int i = 0; i < 64; ++i) acc[i] = fmaxf(acc[i], 0.f);
                    }
                }
                const long long t4 = clock64();
                if (r < R) {
                    if (a.y) {
                        float4* yp = reinterpret_cast<float4*>(a.y + (row0 + r) * TC_C + h * 64);
#pragma unroll
                        for (int i = 0; i < 16; ++i)
                            yp[i] = make_float4(acc[4 * i], acc[4 * i + 1], acc[4 * i + 2], acc[4 * i + 3]);
                    }
                }
                if (a.wa) {
                    float s5[TC_ACT];
#pragma unroll
                    for (int q = 0; q < TC_ACT; ++q) {
                        float t = 0.f;
#pragma unroll
                        for (int i = 0; i < 16; ++i) {
                            const float4 w4 = ld_smem4(wa_s + q * TC_C + h * 64 + 4 * i);
                            t = fmaf(acc[4 * i], w4.x, t);
                            t = fmaf(acc[4 * i + 1], w4.y, t);
                            t = fmaf(acc[4 * i + 2], w4.z, t);
                            t = fmaf(acc[4 * i + 3], w4.w, t);
                        }
                        s5[q] = t;
                    }
                    if (h == 1) {
#pragma unroll
                        for (int q = 0; q < TC_ACT; ++q) plog[r * TC_ACT + q] = s5[q];
                    }
                    producers_sync();
                    if (h == 0 && r < R) {
                        const int b = r / N, n = r - b * N;
                        float* lp = a.logits + ((size_t)n * a.B + (s0 + b)) * TC_ACT;
#pragma unroll
                        for (int q = 0; q < TC_ACT; ++q) lp[q] = s5[q] + plog[r * TC_ACT + q] + ba_s[q];
                    }
                }
                producers_sync();      // the rings (usm alias), sk and plog are free for the next tile
                const long long t5 = clock64();
                tacc[0] += t1 - t0; tacc[1] += t2 - t1; tacc[2] += t3 - t2; tacc[3] += t4 - t3; tacc[4] += t5 - t4;
                tacc[5] += 1;
            }
        }
        if (a.timing && tid == 0)
            for (int i = 0; i < 6; ++i) atomicAdd(&a.timing[i], tacc[i]);
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    if (warp == 0) tmem_dealloc<512>(tmem_acc);
}

// ---------------------------------------------------------------------------------------
// Self-test of the tcgen05 plumbing: D[128][128] = A[128][32] . B[128][32]^T for tf32-exact inputs
// (descriptor encoding, SWIZZLE_128B addressing, TMEM lane/column mapping) -- used by tests only.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128, 1) umma_selftest_kernel(const float* __restrict__ A,
                                                               const float* __restrict__ Bm,
                                                               float* __restrict__ D) {
    extern __shared__ unsigned char smem_raw_st[];
    const uint32_t raw = smem_u32(smem_raw_st);
    unsigned char* sm = smem_raw_st + (((raw + 1023u) & ~1023u) - raw);
    uint64_t* bar = reinterpret_cast<uint64_t*>(sm + 2 * TC_OP_BYTES);
    uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        mbar_init(bar, 1);
        fence_mbar_init();
    }
    if (warp == 0) tmem_alloc<TC_C>(slot);
    for (int i = tid; i < TC_M * 8; i += 128) {
        const int r = i >> 3, jj = i & 7;
        *reinterpret_cast<float4*>(sm + sw128_offset(r, jj)) = *reinterpret_cast<const float4*>(A + r * 32 + jj * 4);
        *reinterpret_cast<float4*>(sm + TC_OP_BYTES + sw128_offset(r, jj)) =
            *reinterpret_cast<const float4*>(Bm + r * 32 + jj * 4);
    }
    fence_proxy_async_smem();
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_acc = *slot;
    if (tid == 0) {
        const uint32_t sa = smem_u32(sm);
        const uint64_t da = umma_desc_sw128(sa), db = umma_desc_sw128(sa + TC_OP_BYTES);
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) umma_tf32(tmem_acc, da + 2 * ks, db + 2 * ks, TC_IDESC, ks != 0);
        umma_commit(bar);
    }
    mbar_wait(bar, 0);
    tcgen05_fence_after();
    const int r = warp * 32 + lane;
    for (int cb = 0; cb < 4; ++cb) {
        float v[32];
        tmem_ld_32x32(tmem_acc + ((uint32_t)(warp * 32) << 16) + cb * 32, v);
#pragma unroll
        for (int i = 0; i < 32; ++i) D[r * TC_C + cb * 32 + i] = v[i];
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<TC_C>(tmem_acc);
}

// ---------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------
static unsigned long long* g_tc_timing = nullptr;
size_t gf_tc_image_floats(int K) { return (size_t)K * TC_C * TC_C * 2; }

int launch_prep_umma_taps(const float* w, float* img, int K, cudaStream_t st) {
    const int n = TC_C * K * TC_C;
    prep_umma_taps_kernel<<<(n + 255) / 256, 256, 0, st>>>(w, img, K * TC_C);
    GPP_LAUNCH_CHECK();
    return GPP_OK;
}

// Samples per 128-row tile such that the GSO powers fit next to the operand rings; 0 = does not fit.
int gf_tc_tile_samples(int N, int K) {
    if (N > TC_M || K > TC_MAX_K) return 0;
    int ts = TC_M / N;
    const size_t budget = 227 * 1024;
    while (ts > 0 && GfTcSmem(N, K, ts).total() > budget) --ts;
    return ts;
}

int launch_gf_forward_tc(const float* x, const void* S, int s_is_f64, const float* wimg, const float* bias,
                         float* y, const float* wa, const float* ba, float* logits, int B, int N, int K,
                         int relu, int allow_bulk, cudaStream_t st) {
    (void)allow_bulk;
    GfTcArgs a;
    a.TS = gf_tc_tile_samples(N, K);
    GPP_REQUIRE(a.TS > 0, GPP_ERR_UNSUPPORTED, "gf_forward_tc: N=%d K=%d does not fit the tensor-core tile", N, K);
    GPP_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15u) == 0, GPP_ERR_INVALID, "gf_forward_tc: x must be 16-byte aligned");
    a.x = x; a.S = S; a.wimg = wimg; a.bias = bias; a.y = y; a.wa = wa; a.ba = ba; a.logits = logits;
    a.B = B; a.N = N; a.K = K;
    a.num_tiles = (B + a.TS - 1) / a.TS;
    a.s_is_f64 = s_is_f64; a.relu = relu;
    a.timing = nullptr;
    if (debug_option(DBG_TC_TIMING)) {
        static unsigned long long* dbuf = nullptr;
        if (!dbuf) { cudaMalloc(&dbuf, 64); cudaMemset(dbuf, 0, 64); }
        a.timing = dbuf;
        g_tc_timing = dbuf;
    }
    const size_t smem = GfTcSmem(N, K, a.TS).total();
    static SmemConfig smem_cfg;
    GPP_CUDA_OK(ensure_dynamic_smem(gf_fwd_tc_kernel, smem_cfg, smem));
    const int grid = a.num_tiles < sm_count() ? a.num_tiles : sm_count();
    gf_fwd_tc_kernel<<<grid, TC_THREADS + 32, smem, st>>>(a);
    GPP_LAUNCH_CHECK();
    return GPP_OK;
}

}  // namespace gpp

using namespace gpp;

// debug: per-phase cycle totals of gf_fwd_tc_kernel (only filled when GPP_TC_TIMING is set): item staging
// loop, wait for the last MMA, TMEM->smem/regs, propagation, stores+logits, tiles
extern "C" int gpp_debug_tc_timing(unsigned long long* out6) {
    if (!g_tc_timing) return GPP_ERR_INVALID;
    cudaDeviceSynchronize();
    cudaMemcpy(out6, g_tc_timing, 48, cudaMemcpyDeviceToHost);
    cudaMemset(g_tc_timing, 0, 64);
    return GPP_OK;
}

extern "C" int gpp_debug_umma_selftest(const float* A, const float* B, float* D, void* stream) {
    GPP_REQUIRE(A && B && D, GPP_ERR_INVALID, "umma_selftest: null pointer");
    const size_t smem = 2 * TC_OP_BYTES + 64 + 1024;
    GPP_CUDA_OK(cudaFuncSetAttribute(umma_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    umma_selftest_kernel<<<1, 128, smem, reinterpret_cast<cudaStream_t>(stream)>>>(A, B, D);
    GPP_LAUNCH_CHECK();
    return GPP_OK;
}
