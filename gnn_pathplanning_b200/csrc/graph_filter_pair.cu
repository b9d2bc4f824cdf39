// CTA-pair tensor-core graph filter for sm_100a: propagate on the CUDA cores, contract on tcgen05 (cta_group::2).
//
// Same contract as gf_fwd_kernel / gf_fwd_tc_kernel (BatchLSIGF, /root/reference/utils/graphUtils/graphML.py:2273-2367,
// + the ReLU of decentralplanner.py:221 and optionally the 128->5 action MLP of :303-315), node-major in / out,
// G = F = 128, K <= 3 taps, small graphs (N nodes known at compile time so the propagation lives in registers).
//
// What bounded the round-1 tcgen05 kernel (profiles/r01_ncu_full_summary_tc_filter_b16k.csv) and what changed:
//   * 393 KB of split taps re-streamed from L2 per 128-row tile by every CTA  ->  the taps are RESIDENT in shared
//     memory for the whole kernel: two CTAs of a cluster (one TPC) form a pair, each keeps one 64-feature half of the
//     B operand (96 KB at K = 3) and `tcgen05.mma.cta_group::2` (M = 256: one 128-row tile per CTA) reads both halves.
//     L2 -> SM traffic per tile drops to the algorithmic bytes (x in, S in, y out).
//   * 3xTF32 (9.2 K tensor cycles per tile)  ->  2-way fp16 split with power-of-two scaling: v * 2^e = hi + lo, both
//     fp16 (11 + 11 significant bits, the same 22 bits 3xTF32 keeps), Y += A_hi.B_hi + A_lo.B_hi + A_hi.B_lo at twice
//     the TF32 rate, half the operand bytes.  fp16 has 5 exponent bits, so the scale is chosen per sample from a bound
//     on |z_k| (max|x| times powers of the largest absolute column sum of S) by a scout warp that runs one tile ahead
//     (its read of the x tile is also the L2 prefetch for the producers), per tensor for the taps; both are undone
//     exactly in the epilogue.  Values more than 2^10 below the per-sample maximum lose relative (not absolute)
//     precision; nothing can overflow.
//   * K x 128 accumulator columns + a 12 K-cycle neighbour-sum epilogue that could not overlap the next tile  ->
//     propagate-then-filter: z_k = z_{k-1}.S is formed BEFORE the GEMM by the producer warps, entirely in registers
//     (a lane owns 4 features of all N nodes of one sample, so z_2 needs no shared-memory round trip), split and stored
//     as A-operand rows; ONE 128-column accumulator per tile, double-buffered in TMEM, so the epilogue of tile i
//     (TMEM -> registers -> scale, bias, ReLU -> swizzled staging -> TMA tensor store / 128->5 logits) overlaps the
//     MMAs of tile i+1 and the propagation of tile i+2.
//
// Warp roles (448 threads per CTA): 0-7 producers (x -> z_k -> fp16 hi|lo operand rows), 8-11 epilogue (one TMEM lane
// quarter each), 12 MMA issuer (one lane; leader CTA only) + TMEM allocation + tap load, 13 scout (S staging, scales).
// Operand row layout: 128 bytes = [32 x hi | 32 x lo] fp16 of one (tap, 32-feature chunk), SWIZZLE_128B K-major; the
// three products of a K = 16 step differ only in the descriptors' start offsets (+0 / +64 bytes).
#include "common.cuh"
#include "tc_common.cuh"

#include <cuda.h>
#include <cuda_fp16.h>

namespace gpp {

constexpr int GP_THREADS = 448;
constexpr int GP_PROD_WARPS = 8;
constexpr int GP_EPI_WARP0 = 8;             // warps 8..11: (warp % 4) = TMEM lane quarter
constexpr int GP_MMA_WARP = 12;
constexpr int GP_SCOUT_WARP = 13;
constexpr int GP_M = 128;                   // node rows per CTA tile
constexpr int GP_C = 128;                   // G = F
constexpr int GP_NCHUNK = 4;                // 32-feature chunks
constexpr int GP_AUNIT = GP_M * 128;        // one (tap, chunk) A operand: 16 KB
constexpr int GP_BUNIT = (GP_C / 2) * 128;  // one (tap, chunk) B operand half: 64 rows, 8 KB
constexpr int GP_MAX_K = 3;
constexpr int GP_ACT = 5;
constexpr int GP_STAGE_BYTES = 2048;        // epilogue staging box: [32 rows][16 cols] fp32
constexpr int GP_SCALE_RING = 8;            // tiles of per-sample inverse scales kept for the epilogue
constexpr int GP_MAX_TS = 64;
constexpr uint32_t GP_IDESC = umma_idesc_f16(2 * GP_M, GP_C);
constexpr long long GP_WATCHDOG_CYCLES = 1LL << 31;   // ~1 s: a stuck pipeline traps instead of hanging the GPU

struct GpArgs {
    CUtensorMap ymap_full;    // y as [rows][128] fp32, box [32 rows][16 cols], SWIZZLE_64B
    CUtensorMap ymap_tail;    // same, box [(TS*N) % 32 rows][16 cols] (last row block of a full tile)
    const float* x;           // [B*N][128] node-major
    const void* S;            // [B][N][N] f32 / f64
    const unsigned char* wimg;   // two CTA halves of K*4 operand units (prep_pair_taps_kernel)
    const float* wscale;      // {2^e, 2^-e} of the taps
    const float* bias;        // [128] or null
    float* y;                 // [B*N][128] or null
    float* logits;            // [N][B][5] or null
    int B, K, TS, num_tiles, num_pairs;
    int s_is_f64, relu, has_act, tail_rows;
    float wa[GP_ACT * GP_C];  // action MLP (kernel-parameter constant bank: FFMA operands, no loads)
    float ba[8];
};

__host__ __device__ constexpr int gp_np(int N) { return (N + 3) & ~3; }

struct GpSmem {
    uint32_t b_off, a_off, s_off, s_bytes, stage_off, misc_off, bar_off, total;
    __host__ __device__ GpSmem(int N, int K, int TS) {
        b_off = 0;
        a_off = b_off + (uint32_t)K * GP_NCHUNK * GP_BUNIT;
        s_off = a_off + 2u * K * GP_AUNIT;
        s_bytes = (uint32_t)TS * N * gp_np(N) * 4;
        stage_off = (s_off + 2 * s_bytes + 511u) & ~511u;
        misc_off = stage_off + 4 * 2 * GP_STAGE_BYTES;     // bias[128] | scale_p[2][64] | scale_e[8][64]
        bar_off = misc_off + (GP_C + 2 * GP_MAX_TS + GP_SCALE_RING * GP_MAX_TS) * 4;
        total = bar_off + 256 + 1024;                      // + alignment slack
    }
};

// ---- mbarrier waits with a watchdog -------------------------------------------------------------------------
__device__ __noinline__ void gp_watchdog_trap(int id, uint32_t parity) {
    printf("gf_fwd_pair_kernel: watchdog -- block %d warp %d stuck on barrier %d parity %u\n", (int)blockIdx.x,
           (int)(threadIdx.x >> 5), id, parity);
    __trap();
}
__device__ __forceinline__ void gp_wait(uint64_t* bar, uint32_t parity, int id) {
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity))
        if (clock64() - t0 > GP_WATCHDOG_CYCLES) gp_watchdog_trap(id, parity);
}
__device__ __forceinline__ void gp_wait_cluster(uint64_t* bar, uint32_t parity, int id) {
    if (mbar_try_wait_cluster(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait_cluster(bar, parity))
        if (clock64() - t0 > GP_WATCHDOG_CYCLES) gp_watchdog_trap(id, parity);
}
__device__ __forceinline__ void gp_wait_warp(uint64_t* bar, uint32_t parity, int id) {
    if ((threadIdx.x & 31) == 0) gp_wait(bar, parity, id);
    __syncwarp();
}
__device__ __forceinline__ void gp_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// fp32 -> fp16 (hi, lo): hi = fp16(v), lo = fp16(v - hi)
__device__ __forceinline__ void gp_split4(const float4 v, uint2& hi, uint2& lo) {
    const __half2 h01 = __floats2half2_rn(v.x, v.y), h23 = __floats2half2_rn(v.z, v.w);
    const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
    const __half2 l01 = __floats2half2_rn(v.x - f01.x, v.y - f01.y), l23 = __floats2half2_rn(v.z - f23.x, v.w - f23.y);
    hi.x = *reinterpret_cast<const uint32_t*>(&h01);
    hi.y = *reinterpret_cast<const uint32_t*>(&h23);
    lo.x = *reinterpret_cast<const uint32_t*>(&l01);
    lo.y = *reinterpret_cast<const uint32_t*>(&l23);
}

// out[n] = sum_m S[m][n] * in[m]   (z_k = z_{k-1} . S, graphML.py:2350; 4 features per lane)
template <int N>
__device__ __forceinline__ void gp_propagate(const float* __restrict__ Ss, const float4 (&in)[N], float4 (&out)[N]) {
    constexpr int NP = gp_np(N);
#pragma unroll
    for (int n = 0; n < N; ++n) out[n] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int m = 0; m < N; ++m) {
#pragma unroll
        for (int q = 0; q < NP / 4; ++q) {
            const float4 s4 = ld_smem4(Ss + m * NP + 4 * q);
            const float sv[4] = {s4.x, s4.y, s4.z, s4.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int n = 4 * q + i;
                if (n < N) {
                    out[n].x = fmaf(sv[i], in[m].x, out[n].x);
                    out[n].y = fmaf(sv[i], in[m].y, out[n].y);
                    out[n].z = fmaf(sv[i], in[m].z, out[n].z);
                    out[n].w = fmaf(sv[i], in[m].w, out[n].w);
                }
            }
        }
    }
}

// rows (row0 + n) of one (tap, chunk) operand unit: this lane's 4 features as hi (bytes 8*l8 of the first 64) and lo
template <int N>
__device__ __forceinline__ void gp_store_tap(unsigned char* unit, int row0, int l8, const float4 (&v)[N]) {
#pragma unroll
    for (int n = 0; n < N; ++n) {
        uint2 hi, lo;
        gp_split4(v[n], hi, lo);
        const int row = row0 + n;
        const uint32_t sub = (uint32_t)(l8 & 1) * 8;
        *reinterpret_cast<uint2*>(unit + sw128_offset(row, l8 >> 1) + sub) = hi;
        *reinterpret_cast<uint2*>(unit + sw128_offset(row, 4 + (l8 >> 1)) + sub) = lo;
    }
}

template <int N>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GP_THREADS, 1)
gf_fwd_pair_kernel(const __grid_constant__ GpArgs a) {
    constexpr int NP = gp_np(N);
    extern __shared__ unsigned char gp_smem_raw[];
    const uint32_t raw = smem_u32(gp_smem_raw);
    unsigned char* sm = gp_smem_raw + (((raw + 1023u) & ~1023u) - raw);
    const int K = a.K, TS = a.TS;
    const GpSmem L(N, K, TS);
    float* bias_s = reinterpret_cast<float*>(sm + L.misc_off);
    float* scale_p = bias_s + GP_C;                       // [2][GP_MAX_TS]  2^e per sample (producers)
    float* scale_e = scale_p + 2 * GP_MAX_TS;             // [GP_SCALE_RING][GP_MAX_TS]  2^-e (epilogue)
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + L.bar_off);
    uint64_t* a_full = bars;          // [2] leader: 16 producer warps (both CTAs) stored their rows of the group
    uint64_t* a_free = bars + 2;      // [2] the MMAs that read the group have completed (multicast commit)
    uint64_t* acc_full = bars + 4;    // [2] all MMAs of the tile pair have completed (multicast commit)
    uint64_t* acc_free = bars + 6;    // [2] leader: 8 epilogue warps (both CTAs) have read the accumulator
    uint64_t* s_full = bars + 8;      // [2] scout staged S + scales of the tile
    uint64_t* s_free = bars + 10;     // [2] 8 producer warps are done with the S buffer
    uint64_t* b_full = bars + 12;     // this CTA's tap half has landed
    uint64_t* b_ready = bars + 13;    // leader: both CTAs' tap halves have landed
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 14);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t rank = cluster_ctarank();
    const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;

    if (tid == 0) {
        for (int i = 0; i < 2; ++i) {
            mbar_init(&a_full[i], 2 * GP_PROD_WARPS);
            mbar_init(&a_free[i], 1);
            mbar_init(&acc_full[i], 1);
            mbar_init(&acc_free[i], 2 * 4);
            mbar_init(&s_full[i], 1);
            mbar_init(&s_free[i], GP_PROD_WARPS);
        }
        mbar_init(b_full, 1);
        mbar_init(b_ready, 2);
        fence_mbar_init();
    }
    if (warp == GP_MMA_WARP) {
        tmem_alloc_cg2<2 * GP_C>(tmem_slot);
        tmem_relinquish_cg2();
    }
    for (int i = tid; i < GP_C; i += GP_THREADS) bias_s[i] = a.bias ? a.bias[i] : 0.f;
    tcgen05_fence_before();
    __syncthreads();
    cluster_sync_all();           // both CTAs' barriers are initialised before any remote arrive / multicast commit
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == GP_MMA_WARP) {
        // =========================== tap load + MMA issuer ===========================
        if (lane == 0) {
            const uint32_t bbytes = (uint32_t)K * GP_NCHUNK * GP_BUNIT;
            mbar_arrive_expect_tx(b_full, bbytes);
            const unsigned char* src = a.wimg + (size_t)rank * bbytes;
            for (int u = 0; u < K * GP_NCHUNK; ++u)
                bulk_g2s(sm + L.b_off + u * GP_BUNIT, src + (size_t)u * GP_BUNIT, GP_BUNIT, b_full);
            gp_wait(b_full, 0, 12);
            mbar_arrive_cluster(b_ready, 0);
            if (rank == 0) {
                gp_wait_cluster(b_ready, 0, 13);
                uint32_t g = 0;
                int t = 0;
                for (int p = cluster_id; p < a.num_pairs; p += num_clusters, ++t) {
                    const int buf = t & 1;
                    if (t >= 2) gp_wait_cluster(&acc_free[buf], ((t >> 1) - 1) & 1, 6 + buf);
                    tcgen05_fence_after();
                    const uint32_t acc = tmem_base + (uint32_t)buf * GP_C;
                    for (int c = 0; c < GP_NCHUNK; ++c, ++g) {
                        const int slot = g & 1;
                        gp_wait_cluster(&a_full[slot], (g >> 1) & 1, slot);
                        tcgen05_fence_after();
                        for (int k = 0; k < K; ++k) {
                            const uint64_t da = umma_desc_sw128(smem_u32(sm + L.a_off + (slot * K + k) * GP_AUNIT));
                            const uint64_t db = umma_desc_sw128(smem_u32(sm + L.b_off + (k * GP_NCHUNK + c) * GP_BUNIT));
                            // descriptor start offsets in 16-byte units: +2 per K = 16 step, +4 = the lo half of the row
#pragma unroll
                            for (int ks = 0; ks < 2; ++ks)
                                umma_f16_cg2(acc, da + 2 * ks, db + 2 * ks, GP_IDESC, (c | k | ks) != 0);
#pragma unroll
                            for (int ks = 0; ks < 2; ++ks) umma_f16_cg2(acc, da + 4 + 2 * ks, db + 2 * ks, GP_IDESC, 1u);
#pragma unroll
                            for (int ks = 0; ks < 2; ++ks) umma_f16_cg2(acc, da + 2 * ks, db + 4 + 2 * ks, GP_IDESC, 1u);
                        }
                        umma_commit_cg2(&a_free[slot], 3);
                    }
                    umma_commit_cg2(&acc_full[buf], 3);
                }
            }
        }
        __syncwarp();
    } else if (warp == GP_SCOUT_WARP) {
        // =========================== scout: S staging + per-sample scales, one tile ahead ===========================
        int t = 0;
        for (int p = cluster_id; p < a.num_pairs; p += num_clusters, ++t) {
            const int tile = 2 * p + (int)rank;
            const int s0 = tile * TS;
            const int ns = max(0, min(TS, a.B - s0));
            const int sb = t & 1;
            if (t >= 2) gp_wait_warp(&s_free[sb], ((t >> 1) - 1) & 1, 10 + sb);
            float* Sd = reinterpret_cast<float*>(sm + L.s_off + sb * L.s_bytes);
            const size_t soff = (size_t)s0 * N * N;
            const int cnt = ns * N * N;
            if (a.s_is_f64) {
                const double* Sg = reinterpret_cast<const double*>(a.S) + soff;
                for (int i = lane; i < cnt; i += 32) {
                    const int sl = i / (N * N), rem = i - sl * N * N, m = rem / N, n = rem - m * N;
                    Sd[(sl * N + m) * NP + n] = static_cast<float>(Sg[i]);      // S.float(), graphML.py:2350
                }
            } else {
                const float* Sg = reinterpret_cast<const float*>(a.S) + soff;
                for (int i = lane; i < cnt; i += 32) {
                    const int sl = i / (N * N), rem = i - sl * N * N, m = rem / N, n = rem - m * N;
                    Sd[(sl * N + m) * NP + n] = Sg[i];
                }
            }
            __syncwarp();
            for (int sl = 0; sl < TS; ++sl) {
                float e2 = 1.f, e2inv = 1.f;
                if (sl < ns) {
                    // largest |x| of the sample (this read is also the L2 prefetch of the tile for the producers)
                    const float* xp = a.x + ((size_t)(s0 + sl) * N) * GP_C + lane * 4;
                    float mx = 0.f;
#pragma unroll
                    for (int n = 0; n < N; ++n) {
                        const float4 v = __ldg(reinterpret_cast<const float4*>(xp + (size_t)n * GP_C));
                        mx = fmaxf(mx, fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));
                    }
                    // largest absolute column sum of S: |z_k[n]| <= max|z_{k-1}| * sum_m |S[m][n]|
                    float cs = 0.f;
                    if (lane < N) {
#pragma unroll
                        for (int m = 0; m < N; ++m) cs += fabsf(Sd[(sl * N + m) * NP + lane]);
                    }
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) {
                        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
                        cs = fmaxf(cs, __shfl_xor_sync(0xffffffffu, cs, o));
                    }
                    float bound = mx;
                    const float cm = fmaxf(cs, 1.f);
                    for (int k = 1; k < K; ++k) bound *= cm;
                    // bound in [2^q, 2^(q+1)): scale by 2^(14-q) so that every |z_k| * scale < 2^15 (fp16 max 65504)
                    const uint32_t bits = __float_as_uint(bound);
                    const int q = (int)((bits >> 23) & 0xFF) - 127;
                    if (bound > 0.f && q < 128) {
                        int e = 14 - q;
                        e = max(-100, min(100, e));
                        e2 = __uint_as_float((uint32_t)(e + 127) << 23);
                        e2inv = __uint_as_float((uint32_t)(127 - e) << 23);
                    }
                }
                if (lane == 0) {
                    scale_p[sb * GP_MAX_TS + sl] = e2;
                    scale_e[(t & (GP_SCALE_RING - 1)) * GP_MAX_TS + sl] = e2inv;
                }
            }
            __syncwarp();
            if (lane == 0) gp_arrive(&s_full[sb]);
        }
    } else if (warp >= GP_EPI_WARP0) {
        // =========================== epilogue: TMEM -> scale, bias, ReLU -> y (TMA store) / logits ===========================
        const int q = warp - GP_EPI_WARP0;
        const int r = q * 32 + lane;                       // TMEM lane = node row of the tile
        unsigned char* stage = sm + L.stage_off + q * 2 * GP_STAGE_BYTES;
        const float inv_w = a.wscale[1];
        const int full_rows = TS * N;
        uint32_t nstore = 0;
        int t = 0;
        for (int p = cluster_id; p < a.num_pairs; p += num_clusters, ++t) {
            const int tile = 2 * p + (int)rank;
            const int s0 = tile * TS;
            const int ns = max(0, min(TS, a.B - s0));
            const int R = ns * N;
            const int buf = t & 1;
            gp_wait_warp(&acc_full[buf], (t >> 1) & 1, 4 + buf);
            tcgen05_fence_after();
            const bool valid = r < R;
            const int bl = valid ? r / N : 0, n = r - bl * N;
            const float sc = scale_e[(t & (GP_SCALE_RING - 1)) * GP_MAX_TS + bl] * inv_w;
            const size_t row0 = (size_t)s0 * N;
            // rows of this warp's 32-row block that belong to the tile; the last tile of the batch is clipped by
            // the tensor map's row extent instead (rows past B*N are never written)
            const int rows_w = max(0, min(32, full_rows - 32 * q));
            const bool last_tile = tile == a.num_tiles - 1;
            const bool do_store = a.y != nullptr && R > 32 * q;
            float acc5[GP_ACT] = {0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int cb = 0; cb < 4; ++cb) {
                float v[32];
                tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * GP_C + cb * 32), v);
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    v[i] = fmaf(v[i], sc, bias_s[cb * 32 + i]);
                    if (a.relu) v[i] = fmaxf(v[i], 0.f);
                }
                if (a.has_act) {
#pragma unroll
                    for (int o = 0; o < GP_ACT; ++o)
#pragma unroll
                        for (int i = 0; i < 32; ++i) acc5[o] = fmaf(v[i], a.wa[o * GP_C + cb * 32 + i], acc5[o]);
                }
                if (a.y != nullptr) {
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        unsigned char* sbuf = stage + (nstore & 1) * GP_STAGE_BYTES;
                        if (lane == 0) bulk_wait_group_read<1>();      // the store that last read this buffer is done
                        __syncwarp();
#pragma unroll
                        for (int j = 0; j < 4; ++j)                     // SWIZZLE_64B: 16-byte chunk j of row r at j ^ ((r >> 1) & 3)
                            *reinterpret_cast<float4*>(sbuf + lane * 64 + ((j ^ ((lane >> 1) & 3)) << 4)) =
                                make_float4(v[h * 16 + 4 * j], v[h * 16 + 4 * j + 1], v[h * 16 + 4 * j + 2], v[h * 16 + 4 * j + 3]);
                        fence_proxy_async_smem();
                        __syncwarp();
                        if (lane == 0 && do_store) {
                            const void* map = (rows_w == 32 || last_tile) ? &a.ymap_full : &a.ymap_tail;
                            tma_store_2d(map, sbuf, cb * 32 + h * 16, (int)(row0 + 32 * q));
                        }
                        if (lane == 0) bulk_commit_group();
                        ++nstore;
                    }
                }
            }
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(&acc_free[buf], 0);
            if (a.has_act && valid) {
                float* lp = a.logits + ((size_t)n * a.B + (s0 + bl)) * GP_ACT;
#pragma unroll
                for (int o = 0; o < GP_ACT; ++o) lp[o] = acc5[o] + a.ba[o];
            }
        }
        if (lane == 0) bulk_wait_group<0>();
    } else {
        // =========================== producers: x -> z_k -> fp16 hi|lo operand rows ===========================
        const int ql = lane >> 3;          // sample of the item this lane works for
        const int l8 = lane & 7;           // which 4 of the chunk's 32 features
        const int ipg = (TS + 3) >> 2;     // items (4 samples each) per group
        uint32_t g = 0, item_base = 0;
        int t = 0;
        for (int p = cluster_id; p < a.num_pairs; p += num_clusters, ++t) {
            const int tile = 2 * p + (int)rank;
            const int s0 = tile * TS;
            const int ns = max(0, min(TS, a.B - s0));
            const int sb = t & 1;
            gp_wait_warp(&s_full[sb], (t >> 1) & 1, 8 + sb);
            const float* Ssm = reinterpret_cast<const float*>(sm + L.s_off + sb * L.s_bytes);
            for (int c = 0; c < GP_NCHUNK; ++c, ++g) {
                const int slot = g & 1;
                bool waited = g < 2;
                unsigned char* abase = sm + L.a_off + (size_t)slot * K * GP_AUNIT;
                for (int j = 0; j < ipg; ++j) {
                    if (((item_base + (uint32_t)(c * ipg + j)) & (GP_PROD_WARPS - 1)) != (uint32_t)warp) continue;
                    const int sl = j * 4 + ql;
                    const bool in_tile = sl < TS;
                    const bool sv = sl < ns;
                    float4 xv[N];
                    const float* xp = a.x + ((size_t)(s0 + sl) * N) * GP_C + c * 32 + l8 * 4;
#pragma unroll
                    for (int n = 0; n < N; ++n)
                        xv[n] = sv ? __ldg(reinterpret_cast<const float4*>(xp + (size_t)n * GP_C)) : make_float4(0.f, 0.f, 0.f, 0.f);
                    if (!waited) {
                        gp_wait_warp(&a_free[slot], ((g >> 1) - 1) & 1, 2 + slot);
                        waited = true;
                    }
                    const float scale = sv ? scale_p[sb * GP_MAX_TS + sl] : 1.f;
#pragma unroll
                    for (int n = 0; n < N; ++n) {
                        xv[n].x *= scale; xv[n].y *= scale; xv[n].z *= scale; xv[n].w *= scale;
                    }
                    const float* Ss = Ssm + (size_t)(in_tile ? sl : 0) * N * NP;
                    const int row0 = sl * N;
                    if (in_tile) gp_store_tap<N>(abase, row0, l8, xv);
                    if (K > 1) {
                        float4 z[N];
                        gp_propagate<N>(Ss, xv, z);
                        if (in_tile) gp_store_tap<N>(abase + GP_AUNIT, row0, l8, z);
                        if (K > 2) {
                            gp_propagate<N>(Ss, z, xv);
                            if (in_tile) gp_store_tap<N>(abase + 2 * GP_AUNIT, row0, l8, xv);
                        }
                    }
                }
                if (!waited) gp_wait_warp(&a_free[slot], ((g >> 1) - 1) & 1, 2 + slot);   // keeps the phases aligned
                fence_proxy_async_smem();       // st.shared operand rows -> visible to the tensor cores
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster(&a_full[slot], 0);
            }
            __syncwarp();
            if (lane == 0) gp_arrive(&s_free[sb]);
            item_base += (uint32_t)(GP_NCHUNK * ipg);
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    cluster_sync_all();           // no CTA leaves (or frees TMEM) while its peer can still arrive on its barriers
    tcgen05_fence_after();
    if (warp == GP_MMA_WARP) tmem_dealloc_cg2<2 * GP_C>(tmem_base);
}

// ---------------------------------------------------------------------------------------
// tap images: w [F][1][K][G] -> per CTA half (64 output features) K*4 units of 64 rows x 128 bytes [hi 32 | lo 32]
// fp16 of w * 2^e, SWIZZLE_128B K-major; e from the largest |w| so that |w| * 2^e < 2^15
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) gp_tap_scale_kernel(const float* __restrict__ w, int n, float* __restrict__ wscale) {
    __shared__ float red[32];
    float mx = 0.f;
    for (int i = threadIdx.x; i < n; i += blockDim.x) mx = fmaxf(mx, fabsf(w[i]));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
    __syncthreads();
    if (threadIdx.x < 32) {
        mx = red[threadIdx.x];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        if (threadIdx.x == 0) {
            float s = 1.f, si = 1.f;
            const int q = (int)((__float_as_uint(mx) >> 23) & 0xFF) - 127;
            if (mx > 0.f && q < 128) {
                const int e = max(-100, min(100, 14 - q));
                s = __uint_as_float((uint32_t)(e + 127) << 23);
                si = __uint_as_float((uint32_t)(127 - e) << 23);
            }
            wscale[0] = s;
            wscale[1] = si;
        }
    }
}

__global__ void gp_prep_taps_kernel(const float* __restrict__ w, unsigned char* __restrict__ img,
                                    const float* __restrict__ wscale, int K) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;     // one (f, k, g) element each
    if (idx >= GP_C * K * GP_C) return;
    const int f = idx / (K * GP_C), kg = idx - f * (K * GP_C), k = kg / GP_C, gch = kg - k * GP_C;
    const int c = gch >> 5, l = gch & 31;
    const int half = f >> 6, fr = f & 63;
    const float v = w[idx] * wscale[0];
    const __half hi = __float2half_rn(v);
    const __half lo = __float2half_rn(v - __half2float(hi));
    unsigned char* unit = img + (size_t)half * K * GP_NCHUNK * GP_BUNIT + (size_t)(k * GP_NCHUNK + c) * GP_BUNIT;
    *reinterpret_cast<__half*>(unit + sw128_offset(fr, l >> 3) + (l & 7) * 2) = hi;
    *reinterpret_cast<__half*>(unit + sw128_offset(fr, 4 + (l >> 3)) + (l & 7) * 2) = lo;
}

// ---------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------
size_t gf_pair_image_bytes(int K) { return (size_t)2 * K * GP_NCHUNK * GP_BUNIT + 64; }   // + {scale, 1/scale}

bool gf_pair_supported(int N, int K) { return N == 10 && K >= 1 && K <= GP_MAX_K; }

int launch_prep_pair_taps(const float* w, void* img, int K, cudaStream_t st) {
    unsigned char* base = reinterpret_cast<unsigned char*>(img);
    float* wscale = reinterpret_cast<float*>(base + (size_t)2 * K * GP_NCHUNK * GP_BUNIT);
    const int n = GP_C * K * GP_C;
    gp_tap_scale_kernel<<<1, 1024, 0, st>>>(w, n, wscale);
    GPP_LAUNCH_CHECK();
    gp_prep_taps_kernel<<<(n + 255) / 256, 256, 0, st>>>(w, base, wscale, K);
    GPP_LAUNCH_CHECK();
    return GPP_OK;
}

typedef CUresult (*GpEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static GpEncodeTiled gp_encode_fn() {
    static GpEncodeTiled fn = nullptr;      // a driver entry point: process-wide, not per device
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<GpEncodeTiled>(p);
    }
    return fn;
}
static int gp_encode_y(CUtensorMap* map, float* y, size_t rows, int box_rows) {
    GpEncodeTiled enc = gp_encode_fn();
    GPP_REQUIRE(enc, GPP_ERR_CUDA, "gf_forward_pair: cuTensorMapEncodeTiled is not available from this driver");
    const cuuint64_t dims[2] = {(cuuint64_t)GP_C, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)GP_C * 4};
    const cuuint32_t box[2] = {16, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, y, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    GPP_REQUIRE(r == CUDA_SUCCESS, GPP_ERR_CUDA, "gf_forward_pair: cuTensorMapEncodeTiled failed (%d)", (int)r);
    return GPP_OK;
}

// wa_host / ba_host: HOST copies of the action MLP (they travel as kernel parameters), or null
int launch_gf_forward_pair(const float* x, const void* S, int s_is_f64, const void* wimg, const float* bias, float* y,
                           const float* wa_host, const float* ba_host, float* logits, int B, int N, int K, int relu,
                           cudaStream_t st) {
    GPP_REQUIRE(gf_pair_supported(N, K), GPP_ERR_UNSUPPORTED, "gf_forward_pair: N=%d K=%d outside the pair kernel's envelope", N, K);
    GPP_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15u) == 0 && (reinterpret_cast<uintptr_t>(y) & 15u) == 0, GPP_ERR_INVALID,
                "gf_forward_pair: x and y must be 16-byte aligned");
    GPP_REQUIRE((wa_host != nullptr) == (logits != nullptr), GPP_ERR_INVALID, "gf_forward_pair: logits need the action MLP");
    static GpArgs a;       // large (kernel-parameter image); filled per call, copied by the launch
    static_assert(sizeof(GpArgs) <= 4096, "kernel parameters exceed 4 KB");
    const int TS = GP_M / N;
    a.x = x; a.S = S;
    a.wimg = reinterpret_cast<const unsigned char*>(wimg);
    a.wscale = reinterpret_cast<const float*>(a.wimg + (size_t)2 * K * GP_NCHUNK * GP_BUNIT);
    a.bias = bias; a.y = y; a.logits = logits;
    a.B = B; a.K = K; a.TS = TS;
    a.num_tiles = (B + TS - 1) / TS;
    a.num_pairs = (a.num_tiles + 1) / 2;
    a.s_is_f64 = s_is_f64; a.relu = relu; a.has_act = wa_host ? 1 : 0;
    a.tail_rows = (TS * N) % 32;
    if (wa_host) {
        for (int i = 0; i < GP_ACT * GP_C; ++i) a.wa[i] = wa_host[i];
        for (int i = 0; i < GP_ACT; ++i) a.ba[i] = ba_host[i];
    }
    if (y) {
        int rc = gp_encode_y(&a.ymap_full, y, (size_t)B * N, 32);
        if (rc) return rc;
        rc = gp_encode_y(&a.ymap_tail, y, (size_t)B * N, a.tail_rows > 0 ? a.tail_rows : 32);
        if (rc) return rc;
    }
    const size_t smem = GpSmem(N, K, TS).total;
    GPP_REQUIRE(smem <= 227 * 1024, GPP_ERR_UNSUPPORTED, "gf_forward_pair: %zu bytes of shared memory do not fit", smem);
    static SmemConfig smem_cfg;
    GPP_CUDA_OK(ensure_dynamic_smem(gf_fwd_pair_kernel<10>, smem_cfg, smem));
    const int max_clusters = sm_count() / 2;
    const int clusters = a.num_pairs < max_clusters ? a.num_pairs : max_clusters;
    gf_fwd_pair_kernel<10><<<2 * clusters, GP_THREADS, smem, st>>>(a);
    GPP_LAUNCH_CHECK();
    return GPP_OK;
}

}  // namespace gpp
