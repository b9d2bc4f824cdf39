// im2col-free tensor-core feature extractor for sm_100a: the per-agent CNN + compress MLP of DecentralPlannerNet
// (/root/reference/graphs/models/decentralplanner.py:155-195,284-290, eval mode) on tcgen05 with fp16 2-way split
// operands (fp32 parity), 8 agents per tile, one persistent CTA per SM.
//
// Layout idea.  An activation map lives in shared memory as fp16 (hi | lo) planes of 8 channels, one 16-byte vector per
// padded pixel, pixels in ONE linear order over the whole tile:   row = y * (8 agents * Wp) + agent * Wp + x.
// That is exactly the canonical NO-SWIZZLE K-major UMMA operand layout (rows 16 B apart, 8-row groups 128 B apart,
// K planes `rows * 16` B apart), and in it the 3x3 tap (ky, kx) of EVERY output pixel is the same buffer shifted by
// (ky * 8 * Wp + kx) rows: the A operand of a tap is just the descriptor start address moved by a multiple of 16 B
// (profiles/probes/umma_nosw_probe.cu checks this addressing on the hardware).  No im2col copy exists anywhere: an
// epilogue writes a layer's output once, and the next layer's 9 taps x Cin/16 MMAs read it in place.  With the
// y-major order the rows a layer needs (y < Hout) are a prefix, so M covers needed image rows only.
//
//   layer  in grid (Hp x Wp)  rows   planes  out ch  M tiles  taps x k16   B bytes
//   conv0  13 x 16 (11+pad)   1664   1       32      10       3 x 2 (*)    6 KB    (*) two ky taps per K=16: LBO = one row;
//                                                                               W_hi and W_lo share the 8-channel K slot
//   conv1  7 x 7              392    4       32      3        9 x 2        36 KB
//   conv2  7 x 8              448    4       64      2        9 x 2        72 KB
//   conv3  4 x 4              128    8       64      1        9 x 4        144 KB
//   conv4  4 x 4              128    8       128     1        9 x 4        288 KB
//   linear 1 x 1              8      16      128     1        1 x 8        64 KB
//
// Precision: v = hi + lo with hi = fp16(v * 2^e), lo = fp16(v * 2^e - hi); products hi*hi + lo*hi + hi*lo accumulate
// in fp32 (TMEM).  e is chosen per agent and layer from the running maximum (two-pass epilogue), weights are scaled per
// layer; residual ~2^-22 per product.  For N <= 64 the filter operand is [W_hi | W_lo] side by side, so A_hi needs one
// MMA of width 2N (M=128 MMAs cost ~47-65 cycles whatever N <= 128 is: the A read bounds them) and the two halves of
// the accumulator are added in the epilogue.
//
// Roles: warp 0 = MMA issuer, warp 1 = loader (filter chunks through a 5 x 16 KB ring of bulk copies, input tiles),
// warps 2-9 = epilogue (TMEM -> BatchNorm/ReLU/max-pool -> fp16 split -> next layer's operand buffer).
#include "common.cuh"
#include "feature.cuh"
#include "tc_common.cuh"

#include <cuda_fp16.h>
#include <stdio.h>

namespace gpp {

constexpr int FM_A = 8;                        // agents per tile
constexpr int FM_EPI_WARPS = 8;
constexpr int FM_EPI_THREADS = 32 * FM_EPI_WARPS;
constexpr int FM_THREADS = 64 + FM_EPI_THREADS;
constexpr int FM_SLOTS = 5, FM_SLOT_BYTES = 16384;
constexpr int FM_NL = 6;

__host__ __device__ constexpr int fm_n(int L) { return L < 2 ? 32 : L < 4 ? 64 : 128; }
__host__ __device__ constexpr int fm_wp(int L) { return L == 0 ? 16 : L == 1 ? 7 : L == 2 ? 8 : L < 5 ? 4 : 1; }
__host__ __device__ constexpr int fm_hp(int L) { return L == 0 ? 13 : L < 3 ? 7 : L < 5 ? 4 : 1; }
__host__ __device__ constexpr int fm_aw(int L) { return FM_A * fm_wp(L); }
__host__ __device__ constexpr int fm_rows(int L) { return fm_hp(L) * fm_aw(L); }
__host__ __device__ constexpr int fm_planes(int L) { return L == 0 ? 1 : L < 3 ? 4 : L < 5 ? 8 : 16; }
__host__ __device__ constexpr int fm_lbo(int L) { return fm_rows(L) * 16; }
__host__ __device__ constexpr int fm_tiles(int L) { return L == 0 ? 10 : L == 1 ? 3 : L == 2 ? 2 : 1; }
__host__ __device__ constexpr int fm_hout(int L) { return L == 0 ? 10 : L == 1 ? 5 : L == 2 ? 4 : L < 5 ? 2 : 1; }
__host__ __device__ constexpr int fm_ks(int L) { return L < 3 ? 2 : L < 5 ? 4 : 8; }           // K=16 steps per tap
__host__ __device__ constexpr int fm_taps(int L) { return L == 5 ? 1 : 9; }
__host__ __device__ constexpr int fm_units(int L) { return L == 0 ? 1 : fm_taps(L) * fm_ks(L); }
__host__ __device__ constexpr int fm_unit_bytes(int L) { return L == 0 ? 6144 : L < 4 ? 64 * fm_n(L) : 8192; }
__host__ __device__ constexpr int fm_upc(int L) { return L == 0 ? 1 : L == 1 ? 8 : L < 4 ? 4 : 2; }   // units per chunk
__host__ __device__ constexpr int fm_chunks(int L) { return (fm_units(L) + fm_upc(L) - 1) / fm_upc(L); }
__host__ __device__ constexpr int fm_img_bytes(int L) { return fm_units(L) * fm_unit_bytes(L); }

// shared-memory map (bytes from the 1024-aligned base)
constexpr int FM_R1 = 0;                           // in0 (hi|lo) -> act2 -> act4
constexpr int FM_R1_BYTES = 57344;
constexpr int FM_R2 = FM_R1 + FM_R1_BYTES;         // act1 -> act3 -> act5
constexpr int FM_R2_BYTES = 50176;
constexpr int FM_XRAW = FM_R2 + FM_R2_BYTES;       // raw fp32 inputs of the tile (bulk copy target)
constexpr int FM_XRAW_BYTES = 11648;
constexpr int FM_STG = FM_XRAW + FM_XRAW_BYTES;    // pooling partners that live in another warp
constexpr int FM_STG_BYTES = 8192;
constexpr int FM_RING = FM_STG + FM_STG_BYTES;
constexpr int FM_CST = FM_RING + FM_SLOTS * FM_SLOT_BYTES;      // epilogue constants [L][scale | shift][128] floats
constexpr int FM_CST_BYTES = 6 * 2 * 128 * 4;
constexpr int FM_MISC = FM_CST + FM_CST_BYTES;
constexpr int FM_SMEM_BYTES = FM_MISC + 1024;
static_assert(2 * fm_planes(0) * fm_lbo(0) <= FM_R1_BYTES && 2 * fm_planes(2) * fm_lbo(2) <= FM_R1_BYTES, "R1");
static_assert(2 * fm_planes(1) * fm_lbo(1) <= FM_R2_BYTES && 2 * fm_planes(3) * fm_lbo(3) <= FM_R2_BYTES, "R2");
static_assert(FM_SMEM_BYTES <= 232448, "shared memory budget");
__host__ __device__ constexpr int fm_in_base(int L) { return (L == 0 || L == 2 || L == 4) ? FM_R1 : FM_R2; }

struct FmMisc {
    uint64_t w_full[FM_SLOTS], w_free[FM_SLOTS];
    uint64_t acc0_full[2], acc0_free[2];
    uint64_t layer_full, act_ready, xraw_full, xraw_free;
    uint32_t tmem_slot;
    uint32_t xflag[2];                 // tile has a non-zero lo part in its inputs
    uint32_t amax[FM_NL][FM_A];        // running max (float bits, values >= 0) of layer L's output per agent
    float mul0[FM_A], inv0[FM_A];      // input scale 2^e0 and its inverse
    float mul1[FM_A], inv1[FM_A];      // act1 scale (from the weight-norm bound of conv0) and its inverse
};
static_assert(sizeof(FmMisc) <= 1024, "misc block");

// constants block in global memory (floats): [L][0][c] = epilogue scale, [L][1][c] = shift; then misc
constexpr int FM_CONST_MISC = FM_NL * 2 * 128;
constexpr int FM_CONST_FLOATS = FM_CONST_MISC + 16;

struct FmArgs {
    const float* x;
    float* feat;
    int total_agents, num_tiles;
    const unsigned char* img[FM_NL];
    const float* consts;
    int x_bulk;     // the inputs may be fetched with bulk copies (device memory, 16-byte aligned)
    int pdl;
    unsigned long long* timing;
};

// ---- waits -------------------------------------------------------------------------------------------------------
constexpr long long FM_WATCHDOG_CYCLES = 2000000000LL;
__device__ __noinline__ void fm_watchdog_trap(int id, uint32_t parity) {
    printf("feature_mma_kernel: watchdog -- block %d warp %d stuck on barrier %d parity %u\n", (int)blockIdx.x,
           (int)(threadIdx.x >> 5), id, parity);
    __trap();
}
__device__ __forceinline__ bool fm_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity), "r"(1000000u)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void fm_wait(uint64_t* bar, uint32_t parity, int id) {
    if (fm_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    while (!fm_try_wait(bar, parity))
        if (clock64() - t0 > FM_WATCHDOG_CYCLES) fm_watchdog_trap(id, parity);
}
__device__ __forceinline__ void fm_wait_warp(uint64_t* bar, uint32_t parity, int id) {
    if ((threadIdx.x & 31) == 0) fm_wait(bar, parity, id);
    __syncwarp();
}
__device__ __forceinline__ void fm_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fm_arrive_warp(uint64_t* bar) {
    __syncwarp();
    if ((threadIdx.x & 31) == 0) fm_arrive(bar);
}
__device__ __forceinline__ void fm_epi_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

// ---- UMMA helpers (no-swizzle K-major descriptors, kind::f16) --------------------------------------------------------
__device__ __forceinline__ uint64_t fm_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) |
           ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46);
}
// Issued from an `if (lane == 0)` region the compiler wraps every tcgen05.mma in a lane loop (ELECT / R2UR.BROADCAST /
// BRA.U.ANY), ~50 cycles per instruction -- about the 47-65 cycles the tensor pipe needs for an M=128 MMA, so the pipe
// stays busy.  A PREDICATED second MMA doubles that (the loop runs whether or not the predicate holds), which is why
// conv0 has two separate code paths; electing a lane per instruction with all lanes running the loops (CUTLASS' way)
// measured slower here: ~70 cycles per MMA (profiles/r02_fm_phase_timing.txt).
__device__ __forceinline__ void fm_commit(uint64_t* bar) { umma_commit(bar); }
__device__ __forceinline__ void fm_mma(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
        : "memory");
}
// 32 lanes x 16 consecutive fp32 columns
__device__ __forceinline__ void fm_tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void fm_tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// scale exponent for a running maximum m >= 0 (float bits): m * 2^e in [2^9, 2^10)
__device__ __forceinline__ int fm_scale_exp(uint32_t bits) {
    if (bits == 0) return 0;
    int e = 9 - ((int)(bits >> 23) - 127);
    return e < -110 ? -110 : (e > 110 ? 110 : e);
}
__device__ __forceinline__ float fm_pow2(int e) { return __int_as_float((e + 127) << 23); }

// 8 floats -> 8 x fp16 hi (16 B) and 8 x fp16 lo
__device__ __forceinline__ void fm_split8(const float* v, uint4& hi, uint4& lo) {
    uint32_t h[4], l[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const __half2 hh = __floats2half2_rn(v[2 * i], v[2 * i + 1]);
        const float2 f = __half22float2(hh);
        const __half2 ll = __floats2half2_rn(v[2 * i] - f.x, v[2 * i + 1] - f.y);
        h[i] = *reinterpret_cast<const uint32_t*>(&hh);
        l[i] = *reinterpret_cast<const uint32_t*>(&ll);
    }
    hi = make_uint4(h[0], h[1], h[2], h[3]);
    lo = make_uint4(l[0], l[1], l[2], l[3]);
}
__device__ __forceinline__ void fm_sts16(uint32_t addr, const uint4 v) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// debug phase timers (A.timing != null): slots 0-5 MMA wait-for-input per layer, 6-11 MMA wait-for-filters, 12-17 MMA issue,
// 18 tiles, 19 convert, 20-25 epilogue wait-for-accumulator per layer, 26-31 epilogue work per layer (warp 2 of block 0)
#define FM_T0() const long long _t0 = timing ? clock64() : 0
#define FM_ACC(slot) do { if (timing) tacc[slot] += (unsigned long long)(clock64() - _t0); } while (0)

// ---- MMA issue of one layer (L >= 1) -------------------------------------------------------------------------------
// Everything but the ring slot is a compile-time constant: the loops are fully unrolled and a descriptor is one 32-bit
// add on its low word (address field; smem addresses < 256 KB never carry into the LBO field at bit 16).  A looped
// version that rebuilt the descriptors per instruction spent ~75-85 cycles per MMA in address arithmetic on the single
// issuing thread (profiles/r02_fm_phase_v1.txt) -- longer than the 47-65 cycles the tensor pipe needs for it.
constexpr uint32_t FM_DESC_HI = (128u >> 4) | (1u << 14);      // SBO = 128 B, descriptor version 1
__device__ __forceinline__ uint64_t fm_desc_w(uint32_t lo) { return ((uint64_t)FM_DESC_HI << 32) | lo; }
__host__ __device__ constexpr uint32_t fm_lbo_field(uint32_t lbo) { return ((lbo >> 4) & 0x3FFF) << 16; }

struct FmIssue {
    uint32_t sm16;        // shared-memory base >> 4
    uint32_t tmem;
    uint32_t slot, phase; // ring position of the next chunk
};

template <int L, int CC>
__device__ __forceinline__ void fm_issue_chunks(FmIssue& is, FmMisc* ms, bool timing, unsigned long long* tacc) {
    if constexpr (CC < fm_chunks(L)) {
        constexpr int N = fm_n(L), KS = fm_ks(L), UNITS = fm_units(L), UPC = fm_upc(L), UB = fm_unit_bytes(L);
        constexpr int AW = fm_aw(L), LBO = (L == 5) ? 128 : fm_lbo(L), TILES = fm_tiles(L);
        constexpr uint32_t A_HI = fm_in_base(L), A_LO = fm_in_base(L) + fm_planes(L) * LBO;
        constexpr int DCOLS = (L <= 3) ? 2 * N : N;
        constexpr int NU = (UNITS - CC * UPC) < UPC ? (UNITS - CC * UPC) : UPC;
        {
            FM_T0();
            fm_wait(&ms->w_full[is.slot], is.phase, 10 + is.slot);
            FM_ACC(6 + L);
        }
        tcgen05_fence_after();
        FM_T0();
        const uint32_t a_hi = is.sm16 + (A_HI >> 4) + fm_lbo_field(LBO), a_lo = is.sm16 + (A_LO >> 4) + fm_lbo_field(LBO);
        const uint32_t bs = is.sm16 + ((FM_RING + is.slot * FM_SLOT_BYTES) >> 4) + fm_lbo_field(L <= 3 ? 32 * N : 2048);
#pragma unroll
        for (int uu = 0; uu < NU; ++uu) {
            const int u = CC * UPC + uu;
            const int tap = u / KS, k16 = u - tap * KS;
            const int shift = (L == 5) ? 0 : ((tap / 3) * AW + (tap % 3));
#pragma unroll
            for (int t = 0; t < TILES; ++t) {
                const uint32_t aoff = (uint32_t)(((t * 128 + shift) * 16 + k16 * 2 * LBO) >> 4);
                const uint32_t d = is.tmem + t * DCOLS;
                if (L <= 3) {
                    const uint64_t db = fm_desc_w(bs + ((uu * UB) >> 4));     // planes of [W_hi | W_lo]: 2N rows of 16 B
                    fm_mma(d, fm_desc_w(a_hi + aoff), db, umma_idesc_f16(128, 2 * N), u > 0);
                    fm_mma(d, fm_desc_w(a_lo + aoff), db, umma_idesc_f16(128, N), 1);
                } else {
                    const uint64_t db_hi = fm_desc_w(bs + ((uu * UB) >> 4)), db_lo = fm_desc_w(bs + ((uu * UB + 4096) >> 4));
                    fm_mma(d, fm_desc_w(a_hi + aoff), db_hi, umma_idesc_f16(128, 128), u > 0);
                    fm_mma(d, fm_desc_w(a_lo + aoff), db_hi, umma_idesc_f16(128, 128), 1);
                    fm_mma(d, fm_desc_w(a_hi + aoff), db_lo, umma_idesc_f16(128, 128), 1);
                }
            }
        }
        fm_commit(&ms->w_free[is.slot]);
        if (++is.slot == FM_SLOTS) { is.slot = 0; is.phase ^= 1; }
        FM_ACC(12 + L);
        fm_issue_chunks<L, CC + 1>(is, ms, timing, tacc);
    }
}
template <int L>
__device__ __forceinline__ void fm_issue_layer(FmIssue& is, FmMisc* ms, bool timing, unsigned long long* tacc) {
    fm_issue_chunks<L, 0>(is, ms, timing, tacc);
    fm_commit(&ms->layer_full);
}

// ---- epilogue pieces -----------------------------------------------------------------------------------------------
__device__ __forceinline__ float4 fm_lds4(const float* p) { return *reinterpret_cast<const float4*>(p); }

// 16 output channels of this thread's accumulator row: (acc (+ the [.. | W_lo] half)) * (scale * inv) + shift, ReLU.
// taddr = TMEM address of the first column, cs = shared-memory constants of the first channel ([0..] scale, [128..] shift)
template <int L>
__device__ __forceinline__ void fm_block_vals(uint32_t taddr, const float* cs, float inv, float (&u)[16]) {
    constexpr int N = fm_n(L);
    float a[16];
    fm_tmem_ld16(taddr, a);
    if (L <= 3) {
        float b[16];
        fm_tmem_ld16(taddr + N, b);
        fm_tmem_wait_ld();
#pragma unroll
        for (int i = 0; i < 16; ++i) a[i] += b[i];
    } else {
        fm_tmem_wait_ld();
    }
#pragma unroll
    for (int i = 0; i < 16; i += 4) {
        const float4 sc = fm_lds4(cs + i), sh = fm_lds4(cs + 128 + i);
        u[i] = fmaxf(fmaf(a[i], sc.x * inv, sh.x), 0.f);
        u[i + 1] = fmaxf(fmaf(a[i + 1], sc.y * inv, sh.y), 0.f);
        u[i + 2] = fmaxf(fmaf(a[i + 2], sc.z * inv, sh.z), 0.f);
        u[i + 3] = fmaxf(fmaf(a[i + 3], sc.w * inv, sh.w), 0.f);
    }
}
__device__ __forceinline__ float fm_max16(const float (&u)[16]) {
    float m0 = fmaxf(u[0], u[1]), m1 = fmaxf(u[2], u[3]), m2 = fmaxf(u[4], u[5]), m3 = fmaxf(u[6], u[7]);
#pragma unroll
    for (int i = 8; i < 16; i += 4) {
        m0 = fmaxf(m0, u[i]); m1 = fmaxf(m1, u[i + 1]); m2 = fmaxf(m2, u[i + 2]); m3 = fmaxf(m3, u[i + 3]);
    }
    return fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
}
// 2-wide max-pool along x between the lanes of a pair (even lane = even x): afterwards the even lane holds the pooled
// channels 0-7 of the block and the odd lane channels 8-15 -- half the shuffles of a full exchange, and both lanes store
__device__ __forceinline__ void fm_xpool(const float (&u)[16], int odd, float (&p)[8]) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const float send = odd ? u[i] : u[i + 8];
        const float mine = odd ? u[i + 8] : u[i];
        p[i] = fmaxf(mine, __shfl_xor_sync(0xffffffffu, send, 1));
    }
}
// store 8 scaled channels as one (hi, lo) pair of 16-byte vectors in plane `plane` of the buffer at `base`
__device__ __forceinline__ void fm_store8(uint32_t base, int lbo, int planes, int row, int plane, const float (&p)[8], float mul) {
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = p[i] * mul;
    uint4 hi, lo;
    fm_split8(v, hi, lo);
    const uint32_t addr = base + plane * lbo + row * 16;
    fm_sts16(addr, hi);
    fm_sts16(addr + planes * lbo, lo);
}
__device__ __forceinline__ void fm_zero(uint32_t addr, int bytes, int et) {
    const uint4 z = make_uint4(0, 0, 0, 0);
    for (int i = et * 16; i < bytes; i += FM_EPI_THREADS * 16) fm_sts16(addr + i, z);
}

// dynamic-scale layers 1..4: zero the target buffer, wait for the accumulators, pass 1 = values (kept in registers) +
// per-agent maximum (+ pooling partners that live in another warp go through FM_STG), pass 2 = scale, split, store
template <int L>
__device__ __forceinline__ void fm_epilogue_dyn(unsigned char* sm, uint32_t sm_base, uint32_t tmem_lane, FmMisc* ms, int na,
                                                int q, int h, int lane, int et, uint32_t& gl, bool timing,
                                                unsigned long long* tacc) {
    constexpr int N = fm_n(L), AW = fm_aw(L), WP = fm_wp(L), HO = fm_hout(L), TILES = fm_tiles(L);
    constexpr int DCOLS = (L <= 3) ? 2 * N : N;
    constexpr bool POOL = (L == 2 || L == 4);
    constexpr int LO = L + 1;                                      // the layer that consumes the output
    constexpr uint32_t OUT = fm_in_base(LO);
    constexpr int OLBO = (LO == 5) ? 128 : fm_lbo(LO), OPL = fm_planes(LO), OAW = fm_aw(LO), OWP = fm_wp(LO);
    constexpr int NB = N / 32;                                     // 16-channel blocks per thread (channel half h)
    constexpr int CV = POOL ? 8 : 16;
    fm_zero(sm_base + OUT, (LO == 5) ? 8192 : 2 * OPL * OLBO, et);
    {
        FM_T0();
        fm_wait_warp(&ms->layer_full, gl & 1, 30 + L);
        FM_ACC(20 + L);
    }
    ++gl;
    tcgen05_fence_after();
    FM_T0();
    float* stg = reinterpret_cast<float*>(sm + FM_STG);
    const float* cs = reinterpret_cast<const float*>(sm + FM_CST) + L * 256 + h * (N / 2);
    const int l = q * 32 + lane, odd = lane & 1;
    const bool rows_here = (L == 1 || L == 2) ? true : (q < 2);     // conv3 / conv4: the valid rows are lanes 0..63
    const bool upper = (L == 2) ? (q >= 2) : (q == 1);              // pooled layers: rows whose y is odd
    float cv[TILES][NB][CV];
    if (rows_here) {
#pragma unroll
        for (int t = 0; t < TILES; ++t) {
            const int r = t * 128 + l;
            const int y = r / AW, a = (r % AW) / WP, x = r % WP;
            const bool valid = y < HO && x < HO && a < na;
            const float inv = (L == 1) ? ms->inv1[a] : fm_pow2(-fm_scale_exp(ms->amax[L - 1][a]));
            float mx = 0.f;
#pragma unroll
            for (int b = 0; b < NB; ++b) {
                float u[16];
                fm_block_vals<L>(tmem_lane + t * DCOLS + h * (N / 2) + 16 * b, cs + 16 * b, inv, u);
                mx = fmaxf(mx, fm_max16(u));
                if constexpr (POOL) {
                    fm_xpool(u, odd, cv[t][b]);
                    if (upper && valid) {        // the y partner lives 64 (conv2) / 32 (conv4) lanes down: park the values
                        float* dst = stg + (((L == 2 ? t * FM_A + a : a) * (L == 2 ? 2 : 1) + (x >> 1)) * N + h * (N / 2) + 16 * b + 8 * odd);
                        *reinterpret_cast<float4*>(dst) = make_float4(cv[t][b][0], cv[t][b][1], cv[t][b][2], cv[t][b][3]);
                        *reinterpret_cast<float4*>(dst + 4) = make_float4(cv[t][b][4], cv[t][b][5], cv[t][b][6], cv[t][b][7]);
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < 16; ++i) cv[t][b][i] = u[i];
                }
            }
            if (valid) atomicMax(&ms->amax[L][a], __float_as_uint(mx));
        }
    }
    tcgen05_fence_before();
    fm_epi_sync();
    const bool writer = POOL ? (rows_here && !upper) : rows_here;
    if (writer) {
#pragma unroll
        for (int t = 0; t < TILES; ++t) {
            const int r = t * 128 + l;
            const int y = r / AW, a = (r % AW) / WP, x = r % WP;
            const bool valid = y < HO && x < HO && a < na;
            const float mul = fm_pow2(fm_scale_exp(ms->amax[L][a]));
            int orow;
            if (LO == 5) orow = a;
            else if (POOL) orow = ((y >> 1) + 1) * OAW + a * OWP + (x >> 1) + 1;
            else orow = (y + 1) * OAW + a * OWP + x + 1;
            if (valid) {
#pragma unroll
                for (int b = 0; b < NB; ++b) {
                    const int c0 = h * (N / 2) + 16 * b;
                    if constexpr (POOL) {
                        const float* src = stg + (((L == 2 ? t * FM_A + a : a) * (L == 2 ? 2 : 1) + (x >> 1)) * N + c0 + 8 * odd);
                        const float4 p0 = fm_lds4(src), p1 = fm_lds4(src + 4);
                        float p[8];
                        p[0] = fmaxf(cv[t][b][0], p0.x); p[1] = fmaxf(cv[t][b][1], p0.y);
                        p[2] = fmaxf(cv[t][b][2], p0.z); p[3] = fmaxf(cv[t][b][3], p0.w);
                        p[4] = fmaxf(cv[t][b][4], p1.x); p[5] = fmaxf(cv[t][b][5], p1.y);
                        p[6] = fmaxf(cv[t][b][6], p1.z); p[7] = fmaxf(cv[t][b][7], p1.w);
                        fm_store8(sm_base + OUT, OLBO, OPL, orow, (c0 >> 3) + odd, p, mul);
                    } else {
                        float p[8];
#pragma unroll
                        for (int hh = 0; hh < 2; ++hh) {
#pragma unroll
                            for (int i = 0; i < 8; ++i) p[i] = cv[t][b][8 * hh + i];
                            fm_store8(sm_base + OUT, OLBO, OPL, orow, (c0 >> 3) + hh, p, mul);
                        }
                    }
                }
            }
        }
    }
    fence_proxy_async_smem();
    fm_arrive_warp(&ms->act_ready);
    FM_ACC(26 + L);
}

// =====================================================================================================================
__global__ void __launch_bounds__(FM_THREADS, 1) feature_mma_kernel(const FmArgs A) {
    extern __shared__ __align__(1024) unsigned char sm[];
    FmMisc* ms = reinterpret_cast<FmMisc*>(sm + FM_MISC);
    const uint32_t sm_base = smem_u32(sm);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < FM_SLOTS; ++s) { mbar_init(&ms->w_full[s], 1); mbar_init(&ms->w_free[s], 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(&ms->acc0_full[b], 1); mbar_init(&ms->acc0_free[b], FM_EPI_WARPS); }
        mbar_init(&ms->layer_full, 1);
        mbar_init(&ms->act_ready, FM_EPI_WARPS);
        mbar_init(&ms->xraw_full, 1);
        mbar_init(&ms->xraw_free, FM_EPI_WARPS);
        ms->xflag[0] = 0; ms->xflag[1] = 0;
        fence_mbar_init();
    }
    if (warp == 0) tmem_alloc<512>(&ms->tmem_slot);
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem = ms->tmem_slot;
    const float* __restrict__ cst = A.consts;
    const bool timing = A.timing != nullptr && blockIdx.x == 0;
    unsigned long long tacc[40];
#pragma unroll
    for (int i = 0; i < 40; ++i) tacc[i] = 0;
    griddep_launch_dependents();       // the next kernel of the stream may start its prologue; it still waits for this grid

    if (warp == 1) {
        // ================= loader =================
        if (lane == 0) {
            uint32_t g = 0;
            int it = 0;
            auto issue_x = [&](int tile) {
                const int na = min(FM_A, A.total_agents - tile * FM_A);
                if (A.x_bulk && na == FM_A) {
                    mbar_arrive_expect_tx(&ms->xraw_full, FM_A * 363 * 4);
                    bulk_g2s(sm + FM_XRAW, A.x + (size_t)tile * FM_A * 363, FM_A * 363 * 4, &ms->xraw_full);
                } else {
                    fm_arrive(&ms->xraw_full);        // the epilogue warps read this tile straight from global memory
                }
            };
            for (int tile = blockIdx.x; tile < A.num_tiles; tile += gridDim.x, ++it) {
                int c = 0;
                for (int L = 0; L < FM_NL; ++L) {
                    const int nch = fm_chunks(L), upc = fm_upc(L), ub = fm_unit_bytes(L), units = fm_units(L);
                    for (int cc = 0; cc < nch; ++cc, ++c, ++g) {
                        if (it == 0 && c == FM_SLOTS - 1) {
                            if (A.pdl) griddep_wait();
                            issue_x(tile);
                        }
                        if (c == 8) {
                            const int next = tile + gridDim.x;
                            if (next < A.num_tiles) {
                                fm_wait(&ms->xraw_free, it & 1, 40);
                                issue_x(next);
                            }
                        }
                        const uint32_t slot = g % FM_SLOTS, use = g / FM_SLOTS;
                        if (use >= 1) fm_wait(&ms->w_free[slot], (use - 1) & 1, 20 + slot);
                        const int nu = (units - cc * upc) < upc ? (units - cc * upc) : upc;
                        const uint32_t bytes = (uint32_t)(nu * ub);
                        mbar_arrive_expect_tx(&ms->w_full[slot], bytes);
                        bulk_g2s(sm + FM_RING + slot * FM_SLOT_BYTES, A.img[L] + (size_t)cc * upc * ub, bytes,
                                 &ms->w_full[slot]);
                    }
                }
            }
        }
    } else if (warp == 0) {
        // ================= MMA issuer =================
        if (lane == 0) {
            uint32_t gp = 0, gr = 0;
            FmIssue is;
            is.sm16 = sm_base >> 4; is.tmem = tmem; is.slot = 0; is.phase = 0;
            int it = 0;
            for (int tile = blockIdx.x; tile < A.num_tiles; tile += gridDim.x, ++it) {
                // ---- conv0: five pairs of image rows, accumulators double-buffered in TMEM columns [256, 512)
                {
                    FM_T0();
                    fm_wait(&ms->act_ready, gr & 1, 50);
                    FM_ACC(0);
                }
                ++gr;
                tcgen05_fence_after();
                const uint32_t has_lo = *reinterpret_cast<volatile uint32_t*>(&ms->xflag[it & 1]);
                {
                    {
                        FM_T0();
                        fm_wait(&ms->w_full[is.slot], is.phase, 10 + is.slot);
                        FM_ACC(6);
                    }
                    tcgen05_fence_after();
                    FM_T0();
                    const uint32_t a_hi = is.sm16 + (FM_R1 >> 4) + fm_lbo_field(2048);
                    const uint32_t a_lo = a_hi + (fm_lbo(0) >> 4);
                    const uint32_t bs = is.sm16 + ((FM_RING + is.slot * FM_SLOT_BYTES) >> 4) + fm_lbo_field(512);
                    for (int j = 0; j < 5; ++j, ++gp) {
                        const uint32_t b = gp & 1;
                        if (gp >= 2) {
                            const long long tw = timing ? clock64() : 0;
                            fm_wait(&ms->acc0_free[b], ((gp >> 1) - 1) & 1, 52 + b);
                            if (timing) tacc[32] += (unsigned long long)(clock64() - tw);
                            tcgen05_fence_after();
                        }
                        if (!has_lo) {
#pragma unroll
                            for (int tt = 0; tt < 2; ++tt) {
                                const uint32_t arow = (uint32_t)((2 * j + tt) * 128);       // 16-byte units = rows
                                const uint32_t d = tmem + 256 + b * 64 + tt * 32;
#pragma unroll
                                for (int kx = 0; kx < 3; ++kx)
#pragma unroll
                                    for (int pr = 0; pr < 2; ++pr)
                                        fm_mma(d, fm_desc_w(a_hi + arow + (uint32_t)(2 * pr * 128 + kx)),
                                               fm_desc_w(bs + (((kx * 2 + pr) * 1024) >> 4)), umma_idesc_f16(128, 32), (kx | pr) != 0);
                            }
                        } else {
#pragma unroll
                            for (int tt = 0; tt < 2; ++tt) {
                                const uint32_t arow = (uint32_t)((2 * j + tt) * 128);
                                const uint32_t d = tmem + 256 + b * 64 + tt * 32;
#pragma unroll
                                for (int kx = 0; kx < 3; ++kx)
#pragma unroll
                                    for (int pr = 0; pr < 2; ++pr) {
                                        const uint32_t aoff = arow + (uint32_t)(2 * pr * 128 + kx);
                                        const uint64_t db = fm_desc_w(bs + (((kx * 2 + pr) * 1024) >> 4));
                                        fm_mma(d, fm_desc_w(a_hi + aoff), db, umma_idesc_f16(128, 32), (kx | pr) != 0);
                                        fm_mma(d, fm_desc_w(a_lo + aoff), db, umma_idesc_f16(128, 32), 1);
                                    }
                            }
                        }
                        fm_commit(&ms->acc0_full[b]);
                    }
                    fm_commit(&ms->w_free[is.slot]);
                    if (++is.slot == FM_SLOTS) { is.slot = 0; is.phase ^= 1; }
                    FM_ACC(12);
                }
#define FM_ISSUE(Lx)                                   \
    {                                                  \
        FM_T0();                                       \
        fm_wait(&ms->act_ready, gr & 1, 50 + Lx);      \
        FM_ACC(Lx);                                    \
    }                                                  \
    ++gr;                                              \
    tcgen05_fence_after();                             \
    fm_issue_layer<Lx>(is, ms, timing, tacc);
                FM_ISSUE(1) FM_ISSUE(2) FM_ISSUE(3) FM_ISSUE(4) FM_ISSUE(5)
#undef FM_ISSUE
                tacc[18] += 1;
            }
            if (timing) {
                for (int i = 0; i < 19; ++i) atomicAdd(&A.timing[i], tacc[i]);
                atomicAdd(&A.timing[32], tacc[32]);
            }
        }
    } else {
        // ================= epilogue warps =================
        const int ew = warp - 2, q = warp & 3, h = ew >> 2, et = threadIdx.x - 64;
        const int l = q * 32 + lane;
        const uint32_t tmem_lane = tmem + ((uint32_t)(q * 32) << 16);
        uint32_t gpe = 0, gl = 0;
        int it = 0;
        if (A.pdl) griddep_wait();         // x (and the feature buffer's previous readers) belong to earlier kernels
        for (int tile = blockIdx.x; tile < A.num_tiles; tile += gridDim.x, ++it) {
            const int na = min(FM_A, A.total_agents - tile * FM_A);
            const bool bulk = A.x_bulk && na == FM_A;
            // ---- inputs: per-agent scale, fp16 split into the padded in0 planes
            const long long t_conv0 = timing ? clock64() : 0;
            if (et < FM_NL * FM_A) (&ms->amax[0][0])[et] = 0;
            if (et == 64) ms->xflag[(it + 1) & 1] = 0;
            fm_zero(sm_base + FM_R2, 2 * fm_planes(1) * fm_lbo(1), et);       // act1 borders
            if (it == 0) {   // epilogue constants (weights: never produced by the kernel in front of this one); their latency
                             // hides behind the first input tile's
                float* cs = reinterpret_cast<float*>(sm + FM_CST);
                for (int i = et; i < FM_NL * 2 * 128; i += FM_EPI_THREADS) cs[i] = __ldg(A.consts + i);
            }
            fm_wait_warp(&ms->xraw_full, it & 1, 41);
            const float* xs = reinterpret_cast<const float*>(sm + FM_XRAW);
            if (!bulk) {
                // ragged last tile, unaligned or host-mapped inputs: one coalesced pass of plain loads into the same buffer
                const float* xg = A.x + (size_t)tile * FM_A * 363;
                float* xw = reinterpret_cast<float*>(sm + FM_XRAW);
                for (int i = et; i < na * 363; i += FM_EPI_THREADS) xw[i] = __ldg(xg + i);
                fm_epi_sync();
            }
            {
                // one warp per agent: scale from max |x|, then the agent's 13 x 16 padded pixels (no barrier in between)
                const int a = ew;
                float m = 0.f;
                if (a < na)
                    for (int i = lane; i < 363; i += 32) m = fmaxf(m, fabsf(xs[a * 363 + i]));
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
                const int e0 = fm_scale_exp(__float_as_uint(m));
                const float mul = fm_pow2(e0);
                if (lane == 0) {
                    ms->mul0[a] = mul;
                    ms->inv0[a] = fm_pow2(-e0);
                    const float bound = m * __ldg(cst + FM_CONST_MISC) + __ldg(cst + FM_CONST_MISC + 1);
                    const int e1 = fm_scale_exp(__float_as_uint(bound));
                    ms->mul1[a] = fm_pow2(e1);
                    ms->inv1[a] = fm_pow2(-e1);
                }
                uint32_t any_lo = 0;
                for (int i = lane; i < 13 * 16; i += 32) {
                    const int y = i >> 4, x = i & 15, row = y * 128 + a * 16 + x;
                    uint4 hi = make_uint4(0, 0, 0, 0), lo = make_uint4(0, 0, 0, 0);
                    if (y >= 1 && y <= 11 && x >= 1 && x <= 11 && a < na) {
                        const float* p = xs + a * 363 + (y - 1) * 11 + (x - 1);
                        const float v0 = p[0] * mul, v1 = p[121] * mul, v2 = p[242] * mul;
                        const __half2 h01 = __floats2half2_rn(v0, v1), h2 = __floats2half2_rn(v2, 0.f);
                        const float2 f01 = __half22float2(h01), f2 = __half22float2(h2);
                        const __half2 l01 = __floats2half2_rn(v0 - f01.x, v1 - f01.y), l2 = __floats2half2_rn(v2 - f2.x, 0.f);
                        // K slot of a pixel: [x0 x1 x2 | x0 x1 x2 | 0 0] -- the filter image holds W_hi in the first triple and
                        // W_lo in the second, so one MMA accumulates x_hi * (W_hi + W_lo)
                        const uint32_t w01 = *reinterpret_cast<const uint32_t*>(&h01), w2 = *reinterpret_cast<const uint32_t*>(&h2);
                        hi.x = w01; hi.y = (w2 & 0xFFFFu) | (w01 << 16); hi.z = (w01 >> 16) | (w2 << 16);
                        lo.x = *reinterpret_cast<const uint32_t*>(&l01); lo.y = *reinterpret_cast<const uint32_t*>(&l2);
                        any_lo |= (lo.x | lo.y) & 0x7FFF7FFFu;
                    }
                    fm_sts16(sm_base + FM_R1 + row * 16, hi);
                    fm_sts16(sm_base + FM_R1 + fm_lbo(0) + row * 16, lo);
                }
                if (__any_sync(0xffffffffu, any_lo != 0) && lane == 0) atomicOr(&ms->xflag[it & 1], 1u);
            }
            fence_proxy_async_smem();
            fm_arrive_warp(&ms->xraw_free);
            fm_epi_sync();                     // xflag / act1 zeroes / in0 complete in every warp
            fm_arrive_warp(&ms->act_ready);
            if (timing) tacc[19] += (unsigned long long)(clock64() - t_conv0);

            // ---- conv0 epilogue: y pairs are the two accumulators of a pair, x pairs neighbouring lanes
            {
                const int a = l >> 4, x = l & 15, odd = lane & 1;
                const float inv = ms->inv0[a], mul = ms->mul1[a];
                const int c0 = h * 16;
                const float* cs = reinterpret_cast<const float*>(sm + FM_CST) + c0;
                float k[16], sh[16];
#pragma unroll
                for (int i = 0; i < 16; i += 4) {
                    const float4 s4 = fm_lds4(cs + i), t4 = fm_lds4(cs + 128 + i);
                    k[i] = s4.x * inv; k[i + 1] = s4.y * inv; k[i + 2] = s4.z * inv; k[i + 3] = s4.w * inv;
                    sh[i] = t4.x; sh[i + 1] = t4.y; sh[i + 2] = t4.z; sh[i + 3] = t4.w;
                }
                const bool valid = (x & ~1) < 10 && a < na;
                const int orow0 = fm_aw(1) + a * fm_wp(1) + (x >> 1) + 1;
                for (int j = 0; j < 5; ++j, ++gpe) {
                    const uint32_t b = gpe & 1;
                    {
                        FM_T0();
                        fm_wait_warp(&ms->acc0_full[b], (gpe >> 1) & 1, 60 + b);
                        FM_ACC(20);
                    }
                    tcgen05_fence_after();
                    FM_T0();
                    const uint32_t col = 256 + b * 64 + c0;
                    float p0[16], q0[16];
                    fm_tmem_ld16(tmem_lane + col, p0);
                    fm_tmem_ld16(tmem_lane + col + 32, q0);
                    fm_tmem_wait_ld();
                    tcgen05_fence_before();
                    fm_arrive_warp(&ms->acc0_free[b]);
                    if (timing) tacc[33] += (unsigned long long)(clock64() - _t0);
                    float u[16];
#pragma unroll
                    for (int i = 0; i < 16; ++i)
                        u[i] = fmaxf(fmaxf(fmaf(p0[i], k[i], sh[i]), fmaf(q0[i], k[i], sh[i])), 0.f);
                    float pp[8];
                    fm_xpool(u, odd, pp);
                    if (valid) fm_store8(sm_base + FM_R2, fm_lbo(1), fm_planes(1), orow0 + j * fm_aw(1), (c0 >> 3) + odd, pp, mul);
                    FM_ACC(26);
                }
                fence_proxy_async_smem();
                fm_arrive_warp(&ms->act_ready);
            }
            fm_epilogue_dyn<1>(sm, sm_base, tmem_lane, ms, na, q, h, lane, et, gl, timing, tacc);
            fm_epilogue_dyn<2>(sm, sm_base, tmem_lane, ms, na, q, h, lane, et, gl, timing, tacc);
            fm_epilogue_dyn<3>(sm, sm_base, tmem_lane, ms, na, q, h, lane, et, gl, timing, tacc);
            fm_epilogue_dyn<4>(sm, sm_base, tmem_lane, ms, na, q, h, lane, et, gl, timing, tacc);
            // ---- compress MLP: rows 0..7 of the tile are the agents
            {
                FM_T0();
                fm_wait_warp(&ms->layer_full, gl & 1, 35);
                FM_ACC(25);
            }
            ++gl;
            tcgen05_fence_after();
            const long long t_mlp0 = timing ? clock64() : 0;
            if (q == 0) {
                const int a = lane & 7;
                const float inv = fm_pow2(-fm_scale_exp(ms->amax[4][a]));
                const float* cs = reinterpret_cast<const float*>(sm + FM_CST) + 5 * 256 + h * 64;
#pragma unroll
                for (int b = 0; b < 4; ++b) {
                    float u[16];
                    fm_block_vals<5>(tmem_lane + h * 64 + 16 * b, cs + 16 * b, inv, u);
                    if (lane < 8 && a < na) {
                        float* dst = A.feat + (size_t)(tile * FM_A + a) * 128 + h * 64 + 16 * b;
#pragma unroll
                        for (int i = 0; i < 16; i += 4)
                            *reinterpret_cast<float4*>(dst + i) = make_float4(u[i], u[i + 1], u[i + 2], u[i + 3]);
                    }
                }
            }
            tcgen05_fence_before();
            fm_epi_sync();      // every warp is done with this tile's shared state before the next tile resets it
            if (timing) tacc[31] += (unsigned long long)(clock64() - t_mlp0);
        }
        if (timing && warp == 4 && lane == 0)          // warp 4 = quadrant 0, channel half 0: has work in every layer
            for (int i = 19; i < 40; ++i)
                if (i != 32) atomicAdd(&A.timing[i], tacc[i]);
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    if (warp == 0) tmem_dealloc<512>(tmem);
}

// =====================================================================================================================
// Weight preparation: per-layer power-of-two scale from max|w|, fp16 (hi, lo) images in chunk order, epilogue constants
// =====================================================================================================================
__global__ void fm_absmax_kernel(const float* __restrict__ w, int n, unsigned int* __restrict__ out) {
    float m = 0.f;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) m = fmaxf(m, fabsf(w[i]));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) atomicMax(out, __float_as_uint(m));
}

__device__ __forceinline__ __half fm_pick(float v, int islo) {
    const __half hi = __float2half_rn(v);
    return islo ? __float2half_rn(v - __half2float(hi)) : hi;
}

// one thread per fp16 element of layer L's image
__global__ void fm_prep_image_kernel(const float* __restrict__ w, __half* __restrict__ img, int L,
                                     const unsigned int* __restrict__ amax) {
    const int N = fm_n(L), idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= fm_img_bytes(L) / 2) return;
    const float mul = fm_pow2(fm_scale_exp(amax[L]));
    const int e = idx & 7;
    float v = 0.f;
    int islo = 0;
    if (L == 0) {
        // [kx][pr][plane][n (32)][8]: halves 0-2 = W_hi of input channel e, 3-5 = W_lo of channel e-3
        int r = idx >> 3;
        const int n = r & 31; r >>= 5;
        const int p = r & 1; r >>= 1;
        const int pr = r & 1, kx = r >> 1;
        const int ky = 2 * pr + p;
        islo = e >= 3;
        if (ky < 3 && e < 6) v = w[(n * 3 + (e % 3)) * 9 + ky * 3 + kx];
    } else if (L <= 3) {
        // [unit][plane][n2 (2N = hi | lo)][8]
        const int Cin = fm_planes(L) * 8, KS = fm_ks(L);
        int r = idx >> 3;
        const int n2 = r % (2 * N); r /= 2 * N;
        const int p = r & 1, u = r >> 1;
        const int tap = u / KS, k16 = u - tap * KS, ci = k16 * 16 + p * 8 + e, n = n2 % N;
        islo = n2 / N;
        v = w[((size_t)n * Cin + ci) * 9 + tap];
    } else {
        // [unit][hi | lo][plane][n (128)][8]
        const int Cin = fm_planes(L) * 8, KS = fm_ks(L);
        int r = idx >> 3;
        const int n = r & 127; r >>= 7;
        const int p = r & 1; r >>= 1;
        islo = r & 1;
        const int u = r >> 1;
        const int tap = u / KS, k16 = u - tap * KS, ci = k16 * 16 + p * 8 + e;
        v = (L == 5) ? w[(size_t)n * 128 + ci] : w[((size_t)n * Cin + ci) * 9 + tap];
    }
    img[idx] = fm_pick(v * mul, islo);
}

// epilogue constants: [L][0][c] = BatchNorm scale / 2^ew_L, [L][1][c] = shift; misc[0] = max_c |sc_c| * sum|w_c| of conv0,
// misc[1] = max_c |sh_c| of conv0 (bound of conv0's output for inputs of magnitude <= 1)
__global__ void fm_prep_consts_kernel(const float* const* sc, const float* const* sh, const float* b5, const float* w0,
                                      const unsigned int* __restrict__ amax, float* __restrict__ cst) {
    const int c = threadIdx.x;      // 128 threads
    for (int L = 0; L < FM_NL; ++L) {
        const float inv = fm_pow2(-fm_scale_exp(amax[L]));
        float s = 0.f, t = 0.f;
        if (c < fm_n(L)) {
            s = (L < 5 ? sc[L][c] : 1.f) * inv;
            t = L < 5 ? sh[L][c] : b5[c];
        }
        cst[(L * 2) * 128 + c] = s;
        cst[(L * 2 + 1) * 128 + c] = t;
    }
    __shared__ float red[2][128];
    float b0 = 0.f, b1 = 0.f;
    if (c < 32) {
        float sw = 0.f;
        for (int i = 0; i < 27; ++i) sw += fabsf(w0[c * 27 + i]);
        b0 = fabsf(sc[0][c]) * sw;
        b1 = fabsf(sh[0][c]);
    }
    red[0][c] = b0; red[1][c] = b1;
    __syncthreads();
    if (c == 0) {
        float m0 = 0.f, m1 = 0.f;
        for (int i = 0; i < 32; ++i) { m0 = fmaxf(m0, red[0][i]); m1 = fmaxf(m1, red[1][i]); }
        cst[FM_CONST_MISC] = m0;
        cst[FM_CONST_MISC + 1] = m1;
    }
}

size_t feature_mma_arena_floats() {
    size_t bytes = 0;
    for (int L = 0; L < FM_NL; ++L) bytes += fm_img_bytes(L);
    return bytes / 4 + FM_CONST_FLOATS + 64 /* amax + pointer table */;
}

// arena layout (floats): images L0..L5 | consts | amax[8] | pointer table (sc[5], sh[5] as 64-bit)
int launch_prep_feature_mma(const float* const* conv_w, const float* compress_w, const float* const* sc,
                            const float* const* sh, const float* b5, float* arena, cudaStream_t st) {
    size_t off = 0;
    unsigned char* base = reinterpret_cast<unsigned char*>(arena);
    size_t img_off[FM_NL];
    for (int L = 0; L < FM_NL; ++L) { img_off[L] = off; off += fm_img_bytes(L); }
    float* cst = reinterpret_cast<float*>(base + off);
    unsigned int* amax = reinterpret_cast<unsigned int*>(cst + FM_CONST_FLOATS);
    const float** ptrs = reinterpret_cast<const float**>(amax + 8);
    GPP_CUDA_OK(cudaMemsetAsync(amax, 0, 32, st));
    static const int cin[6] = {3, 32, 32, 64, 64, 128};
    for (int L = 0; L < FM_NL; ++L) {
        const float* w = L < 5 ? conv_w[L] : compress_w;
        const int n = L < 5 ? fm_n(L) * cin[L] * 9 : 128 * 128;
        fm_absmax_kernel<<<32, 256, 0, st>>>(w, n, amax + L);
        GPP_LAUNCH_CHECK();
    }
    for (int L = 0; L < FM_NL; ++L) {
        const int n = fm_img_bytes(L) / 2;
        fm_prep_image_kernel<<<(n + 255) / 256, 256, 0, st>>>(L < 5 ? conv_w[L] : compress_w,
                                                              reinterpret_cast<__half*>(base + img_off[L]), L, amax);
        GPP_LAUNCH_CHECK();
    }
    const float* host_ptrs[10];
    for (int l = 0; l < 5; ++l) { host_ptrs[l] = sc[l]; host_ptrs[5 + l] = sh[l]; }
    GPP_CUDA_OK(cudaMemcpyAsync(ptrs, host_ptrs, sizeof(host_ptrs), cudaMemcpyHostToDevice, st));
    GPP_CUDA_OK(cudaStreamSynchronize(st));          // host_ptrs is a stack array
    fm_prep_consts_kernel<<<1, 128, 0, st>>>(ptrs, ptrs + 5, b5, conv_w[0], amax, cst);
    GPP_LAUNCH_CHECK();
    return GPP_OK;
}

static unsigned long long* g_fm_timing = nullptr;
int debug_feature_mma_timing(unsigned long long* out32) {
    if (!g_fm_timing) return GPP_ERR_INVALID;
    cudaDeviceSynchronize();
    cudaMemcpy(out32, g_fm_timing, 320, cudaMemcpyDeviceToHost);
    cudaMemset(g_fm_timing, 0, 320);
    return GPP_OK;
}

int launch_feature_mma_kernel(const FeArgs& fa, const float* arena, int x_bulk, cudaStream_t st) {
    FmArgs a;
    a.x = fa.x; a.feat = fa.feat; a.total_agents = fa.total_agents;
    a.num_tiles = (fa.total_agents + FM_A - 1) / FM_A;
    const unsigned char* base = reinterpret_cast<const unsigned char*>(arena);
    size_t off = 0;
    for (int L = 0; L < FM_NL; ++L) { a.img[L] = base + off; off += fm_img_bytes(L); }
    a.consts = reinterpret_cast<const float*>(base + off);
    a.x_bulk = (x_bulk && (reinterpret_cast<uintptr_t>(fa.x) & 15) == 0) ? 1 : 0;
    a.pdl = fa.pdl;
    a.timing = nullptr;
    if (debug_option(DBG_TC_TIMING)) {
        if (!g_fm_timing) { cudaMalloc(&g_fm_timing, 320); cudaMemset(g_fm_timing, 0, 320); }
        a.timing = g_fm_timing;
    }
    static SmemConfig smem_cfg;
    GPP_CUDA_OK(ensure_dynamic_smem(feature_mma_kernel, smem_cfg, FM_SMEM_BYTES));
    const int grid = a.num_tiles < sm_count() ? a.num_tiles : sm_count();
    GPP_CUDA_OK(launch_maybe_pdl(feature_mma_kernel, grid, FM_THREADS, FM_SMEM_BYTES, st, fa.pdl, a));
    GPP_LAUNCH_CHECK();
    return GPP_OK;
}

}  // namespace gpp
