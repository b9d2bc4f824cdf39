// Tensor-core (tcgen05 / TMEM) fused graph filter for sm_100a.
//
// Same contract as gf_fwd_kernel (graph_filter.cu; BatchLSIGF, /root/reference/utils/graphUtils/
// graphML.py:2273-2367, + ReLU + the 128->5 action MLP), node-major input, G = F = 128, but the tap
// contraction  Y[128 rows, 128] = Z[128 rows, K*128] . W^T  runs on the 5th-generation tensor cores:
//
//   * fp32 parity through split precision (3xTF32): every operand is split into hi = tf32(v) and
//     lo = tf32(v - hi); D += A_hi.B_hi + A_lo.B_hi + A_hi.B_lo with fp32 accumulation in TMEM.
//     The dropped lo.lo term is 2^-22 relative -- inside the 1e-5 bar with margin.
//   * B operand (taps): pre-split, pre-swizzled 32 KB chunk images (hi | lo) prepared once per weight
//     update; one 1-D bulk async copy (UBLKCP) per K chunk lands them in the canonical SWIZZLE_128B
//     layout, completing on an mbarrier.
//   * A operand (node signals): CUDA cores build each 128 x 32 chunk of z_k = (S^k)^T x straight from the
//     fp32 x tile and the per-sample GSO powers in shared memory, split it and store it swizzled;
//     a 2-stage ring overlaps this with the MMAs of the previous chunk (tcgen05.commit -> mbarrier).
//   * accumulators: 4 x (128 lanes x 128 fp32 columns) = all 512 TMEM columns, used round-robin by K
//     chunk so that each one sees a quarter of the accumulation steps (the tensor core's fp32
//     accumulation is not round-to-nearest; measured 4e-6 relative error with a single accumulator).
//     The epilogue reads them with tcgen05.ld (one thread per node row), sums them in fp32, adds
//     bias, applies ReLU, optionally stores y and reduces the action logits in-thread.
#include "common.cuh"
#include "tc_common.cuh"

namespace gpp {

constexpr int TC_THREADS = 256;
constexpr int TC_M = 128;                  // node rows per tile (UMMA M)
constexpr int TC_C = 128;                  // G = F = 128 (UMMA N)
constexpr int TC_XS = TC_C + 4;            // row stride of the plain fp32 x tile
constexpr int TC_CHUNK_K = 32;             // K elements per chunk (128 bytes of tf32)
constexpr int TC_OP_BYTES = TC_M * 128;    // one operand half (hi or lo) of one chunk: 16 KB
constexpr int TC_STAGE_BYTES = 4 * TC_OP_BYTES;   // A_hi | A_lo | B_hi | B_lo
constexpr int TC_ACT = 5;
constexpr uint32_t TC_IDESC = umma_idesc_tf32(TC_M, TC_C);
constexpr int TC_NACC = 4;                 // TMEM accumulators (128 columns each)
constexpr int TC_TMEM_COLS = TC_NACC * TC_C;

// ---------------------------------------------------------------------------------------
// B-operand images: img[chunk] = { hi[128 x 32] , lo[128 x 32] } in SWIZZLE_128B K-major layout,
// chunk c covers reduction indices kg in [32c, 32c+32) of w[f][kg]  (w = the module's [F,1,K,G]).
// ---------------------------------------------------------------------------------------
__global__ void prep_umma_taps_kernel(const float* __restrict__ w, float* __restrict__ img, int KG) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;     // one (f, kg) element each
    if (idx >= TC_C * KG) return;
    const int f = idx / KG, kg = idx - f * KG;
    const int c = kg / TC_CHUNK_K, kk = kg - c * TC_CHUNK_K;
    float hi, lo;
    split_tf32(w[idx], hi, lo);
    const uint32_t off = sw128_offset(f, kk >> 2) / 4 + (kk & 3);
    float* base = img + (size_t)c * (2 * TC_OP_BYTES / 4);
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
    int s_is_f64, relu, bulk_x;
};

struct GfTcSmem {
    uint32_t sk_floats;
    __host__ __device__ GfTcSmem(int N, int K, int TS) {
        sk_floats = (uint32_t)(((K > 1 ? (K - 1) : 0) * TS * N * N + 3) / 4 * 4);
    }
    __host__ __device__ uint32_t stage(int s) const { return (uint32_t)s * TC_STAGE_BYTES; }
    __host__ __device__ uint32_t xs_off() const { return 2 * TC_STAGE_BYTES; }
    __host__ __device__ uint32_t sk_off() const { return xs_off() + TC_M * TC_XS * 4; }
    __host__ __device__ uint32_t misc_off() const { return sk_off() + sk_floats * 4; }   // bias, wa, ba
    __host__ __device__ uint32_t bar_off() const { return misc_off() + (TC_C + TC_ACT * TC_C + 8) * 4; }
    __host__ __device__ uint32_t total() const { return bar_off() + 64 + 1024; }           // + alignment slack
};

__global__ void __launch_bounds__(TC_THREADS, 1) gf_fwd_tc_kernel(const GfTcArgs a) {
    extern __shared__ unsigned char smem_raw_tc[];
    const uint32_t raw = smem_u32(smem_raw_tc);
    unsigned char* sm = smem_raw_tc + (((raw + 1023u) & ~1023u) - raw);      // 1024-byte aligned base
    const GfTcSmem L(a.N, a.K, a.TS);
    float* xs = reinterpret_cast<float*>(sm + L.xs_off());
    float* sk = reinterpret_cast<float*>(sm + L.sk_off());
    float* bias_s = reinterpret_cast<float*>(sm + L.misc_off());
    float* wa_s = bias_s + TC_C;
    float* ba_s = wa_s + TC_ACT * TC_C;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + L.bar_off());
    uint64_t* full = bars;          // [2] B chunk landed
    uint64_t* done = bars + 2;      // [2] MMAs that read the stage have completed
    uint64_t* xbar = bars + 4;      // x tile landed
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 5);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int N = a.N, K = a.K;
    const int nchunks = K * (TC_C / TC_CHUNK_K);

    if (tid == 0) {
        mbar_init(&full[0], 1); mbar_init(&full[1], 1);
        mbar_init(&done[0], 1); mbar_init(&done[1], 1);
        mbar_init(xbar, 1);
        fence_mbar_init();
    }
    if (warp == 0) tmem_alloc<TC_TMEM_COLS>(tmem_slot);
    for (int i = tid; i < TC_C; i += TC_THREADS) bias_s[i] = a.bias ? a.bias[i] : 0.f;
    if (a.wa) {
        for (int i = tid; i < TC_ACT * TC_C; i += TC_THREADS) wa_s[i] = a.wa[i];
        if (tid < TC_ACT) ba_s[tid] = a.ba[tid];
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_acc = *tmem_slot;

    uint32_t xphase = 0;
    uint32_t g = 0;     // chunks issued so far by this CTA (stage = g & 1, use index = g >> 1)

    for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x) {
        const int s0 = tile * a.TS;
        const int ns = min(a.TS, a.B - s0);
        const int R = ns * N;
        const size_t row0 = (size_t)s0 * N;

        // ---- stage the fp32 x tile and the GSO tile ---------------------------------------
        if (a.bulk_x) {
            fence_proxy_async_smem();
            if (warp == 0) {
                if (lane == 0) mbar_arrive_expect_tx(xbar, (uint32_t)R * TC_C * 4u);
                __syncwarp();
                for (int r = lane; r < R; r += 32) bulk_g2s(xs + r * TC_XS, a.x + (row0 + r) * TC_C, TC_C * 4u, xbar);
            }
        } else {
            const float4* xp = reinterpret_cast<const float4*>(a.x + row0 * TC_C);
            for (int i = tid; i < R * (TC_C / 4); i += TC_THREADS)
                *reinterpret_cast<float4*>(xs + (i >> 5) * TC_XS + (i & 31) * 4) = xp[i];
        }
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
        }
        __syncthreads();
        // GSO powers: sk[p] = S^(p+1) = S^p . S   (z_k = (S^k)^T x ; x.S^k of graphML.py:2350)
        for (int p = 1; p < K - 1; ++p) {
            const float* prev = sk + (size_t)(p - 1) * a.TS * N * N;
            float* cur = sk + (size_t)p * a.TS * N * N;
            for (int i = tid; i < ns * N * N; i += TC_THREADS) {
                const int b = i / (N * N), rem = i - b * N * N;
                const int r = rem / N, c = rem - r * N;
                float acc = 0.f;
                for (int m = 0; m < N; ++m) acc = fmaf(prev[b * N * N + r * N + m], sk[b * N * N + m * N + c], acc);
                cur[i] = acc;
            }
            __syncthreads();
        }
        if (a.bulk_x) {
            mbar_wait(xbar, xphase);
            xphase ^= 1;
        }

        // ---- K chunks: build A (CUDA cores) | fetch B (bulk copy) | MMA (tensor cores) ---------
        for (int c = 0; c < nchunks; ++c, ++g) {
            const int st = g & 1;
            const uint32_t use = g >> 1;
            unsigned char* stage = sm + L.stage(st);
            if (use >= 1) mbar_wait(&done[st], (use - 1) & 1);    // MMAs that read this stage are done
            if (tid == 0) {
                mbar_arrive_expect_tx(&full[st], 2u * TC_OP_BYTES);
                bulk_g2s(stage + 2 * TC_OP_BYTES, a.wimg + (size_t)c * (2 * TC_OP_BYTES / 4), 2u * TC_OP_BYTES,
                         &full[st]);
            }
            const int k = c >> 2, col0 = (c & 3) * TC_CHUNK_K;
            const int j = tid & 7;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int r = (tid >> 3) + 32 * i;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (r < R) {
                    if (k == 0) {
                        v = ld_smem4(xs + r * TC_XS + col0 + j * 4);
                    } else {
                        const int b = r / N, n = r - b * N;
                        const float* sp = sk + ((size_t)(k - 1) * a.TS + b) * N * N + n;
                        const float* xp = xs + (b * N) * TC_XS + col0 + j * 4;
                        for (int m = 0; m < N; ++m) {
                            const float s = sp[m * N];
                            const float4 xv = ld_smem4(xp + m * TC_XS);
                            v.x = fmaf(s, xv.x, v.x);
                            v.y = fmaf(s, xv.y, v.y);
                            v.z = fmaf(s, xv.z, v.z);
                            v.w = fmaf(s, xv.w, v.w);
                        }
                    }
                }
                float4 hi, lo;
                split_tf32(v.x, hi.x, lo.x);
                split_tf32(v.y, hi.y, lo.y);
                split_tf32(v.z, hi.z, lo.z);
                split_tf32(v.w, hi.w, lo.w);
                const uint32_t off = sw128_offset(r, j);
                *reinterpret_cast<float4*>(stage + off) = hi;
                *reinterpret_cast<float4*>(stage + TC_OP_BYTES + off) = lo;
            }
            fence_proxy_async_smem();      // generic-proxy stores -> visible to the tensor core
            __syncthreads();
            if (tid == 0) {
                mbar_wait(&full[st], use & 1);
                tcgen05_fence_after();
                const uint32_t sa = smem_u32(stage);
                const uint64_t a_hi = umma_desc_sw128(sa), a_lo = umma_desc_sw128(sa + TC_OP_BYTES);
                const uint64_t b_hi = umma_desc_sw128(sa + 2 * TC_OP_BYTES), b_lo = umma_desc_sw128(sa + 3 * TC_OP_BYTES);
                const uint32_t acc = tmem_acc + (uint32_t)(c % TC_NACC) * TC_C;   // round-robin accumulator
                const uint32_t fresh = (c < TC_NACC) ? 0u : 1u;                  // first chunk into it overwrites
#pragma unroll
                for (int ks = 0; ks < TC_CHUNK_K / 8; ++ks)     // +32 bytes (2 x 16 B units) per K = 8 step
                    umma_tf32(acc, a_hi + 2 * ks, b_hi + 2 * ks, TC_IDESC, (ks != 0) ? 1u : fresh);
#pragma unroll
                for (int ks = 0; ks < TC_CHUNK_K / 8; ++ks)
                    umma_tf32(acc, a_lo + 2 * ks, b_hi + 2 * ks, TC_IDESC, 1u);
#pragma unroll
                for (int ks = 0; ks < TC_CHUNK_K / 8; ++ks)
                    umma_tf32(acc, a_hi + 2 * ks, b_lo + 2 * ks, TC_IDESC, 1u);
                umma_commit(&done[st]);
            }
        }

        // ---- epilogue: TMEM -> registers, bias, ReLU, y, action logits -------------------------
        {
            const uint32_t last = g - 1;
            mbar_wait(&done[last & 1], (last >> 1) & 1);
            tcgen05_fence_after();
            if (warp < 4) {
                const int r = warp * 32 + lane;
                float s[TC_ACT];
#pragma unroll
                for (int q = 0; q < TC_ACT; ++q) s[q] = 0.f;
#pragma unroll 1
                for (int cb = 0; cb < TC_C / 32; ++cb) {
                    float v[32];
                    tmem_ld_32x32(tmem_acc + ((uint32_t)(warp * 32) << 16) + cb * 32, v);
                    const int nacc = nchunks < TC_NACC ? nchunks : TC_NACC;
                    for (int q = 1; q < nacc; ++q) {
                        float u[32];
                        tmem_ld_32x32(tmem_acc + ((uint32_t)(warp * 32) << 16) + q * TC_C + cb * 32, u);
#pragma unroll
                        for (int i = 0; i < 32; ++i) v[i] += u[i];
                    }
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        float t = v[i] + bias_s[cb * 32 + i];
                        if (a.relu) t = fmaxf(t, 0.f);
                        v[i] = t;
                    }
                    if (a.y && r < R) {
                        float4* yp = reinterpret_cast<float4*>(a.y + (row0 + r) * TC_C + cb * 32);
#pragma unroll
                        for (int i = 0; i < 8; ++i) yp[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
                    }
                    if (a.wa) {
#pragma unroll
                        for (int q = 0; q < TC_ACT; ++q) {
                            float acc = s[q];
#pragma unroll
                            for (int i = 0; i < 32; ++i) acc = fmaf(v[i], wa_s[q * TC_C + cb * 32 + i], acc);
                            s[q] = acc;
                        }
                    }
                }
                if (a.wa && r < R) {
                    const int b = r / N, n = r - b * N;
                    float* lp = a.logits + ((size_t)n * a.B + (s0 + b)) * TC_ACT;
#pragma unroll
                    for (int q = 0; q < TC_ACT; ++q) lp[q] = s[q] + ba_s[q];
                }
            }
            tcgen05_fence_before();
            __syncthreads();       // TMEM, xs and sk are free for the next tile
            tcgen05_fence_after();
        }
    }
    if (warp == 0) tmem_dealloc<TC_TMEM_COLS>(tmem_acc);
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
        const int r = i >> 3, j = i & 7;
        *reinterpret_cast<float4*>(sm + sw128_offset(r, j)) = *reinterpret_cast<const float4*>(A + r * 32 + j * 4);
        *reinterpret_cast<float4*>(sm + TC_OP_BYTES + sw128_offset(r, j)) =
            *reinterpret_cast<const float4*>(Bm + r * 32 + j * 4);
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
size_t gf_tc_image_floats(int K) { return (size_t)K * TC_C * TC_C * 2; }

int launch_prep_umma_taps(const float* w, float* img, int K, cudaStream_t st) {
    const int n = TC_C * K * TC_C;
    prep_umma_taps_kernel<<<(n + 255) / 256, 256, 0, st>>>(w, img, K * TC_C);
    GPP_LAUNCH_CHECK();
    return GPP_OK;
}

// Samples per 128-row tile such that the GSO powers fit next to the operand ring; 0 = does not fit.
int gf_tc_tile_samples(int N, int K) {
    if (N > TC_M) return 0;
    int ts = TC_M / N;
    const size_t budget = 227 * 1024;
    while (ts > 0 && GfTcSmem(N, K, ts).total() > budget) --ts;
    return ts;
}

int launch_gf_forward_tc(const float* x, const void* S, int s_is_f64, const float* wimg, const float* bias,
                         float* y, const float* wa, const float* ba, float* logits, int B, int N, int K,
                         int relu, int allow_bulk, cudaStream_t st) {
    GfTcArgs a;
    a.TS = gf_tc_tile_samples(N, K);
    GPP_REQUIRE(a.TS > 0, GPP_ERR_UNSUPPORTED, "gf_forward_tc: N=%d K=%d does not fit the tensor-core tile", N, K);
    a.x = x; a.S = S; a.wimg = wimg; a.bias = bias; a.y = y; a.wa = wa; a.ba = ba; a.logits = logits;
    a.B = B; a.N = N; a.K = K;
    a.num_tiles = (B + a.TS - 1) / a.TS;
    a.s_is_f64 = s_is_f64; a.relu = relu;
    a.bulk_x = (allow_bulk && (reinterpret_cast<uintptr_t>(x) & 15u) == 0) ? 1 : 0;
    const size_t smem = GfTcSmem(N, K, a.TS).total();
    static size_t configured = 0;
    if (smem > configured) {
        GPP_CUDA_OK(cudaFuncSetAttribute(gf_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    const int grid = a.num_tiles < sm_count() ? a.num_tiles : sm_count();
    gf_fwd_tc_kernel<<<grid, TC_THREADS, smem, st>>>(a);
    GPP_LAUNCH_CHECK();
    return GPP_OK;
}

}  // namespace gpp

using namespace gpp;

extern "C" int gpp_debug_umma_selftest(const float* A, const float* B, float* D, void* stream) {
    GPP_REQUIRE(A && B && D, GPP_ERR_INVALID, "umma_selftest: null pointer");
    const size_t smem = 2 * TC_OP_BYTES + 64 + 1024;
    GPP_CUDA_OK(cudaFuncSetAttribute(umma_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    umma_selftest_kernel<<<1, 128, smem, reinterpret_cast<cudaStream_t>(stream)>>>(A, B, D);
    GPP_LAUNCH_CHECK();
    return GPP_OK;
}
