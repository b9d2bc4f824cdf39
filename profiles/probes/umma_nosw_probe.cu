// Probe (B200): tcgen05.mma kind::f16 with NO-SWIZZLE K-major operands laid out [k8 plane][row][8 halves]
// (row stride 16 B, 8-row groups 128 B apart = SBO, planes LBO apart), start address shifted by whole rows (16 B
// granularity, not 1024-aligned) -- the addressing the im2col-free convolution kernel relies on -- and the issue
// rate of M=128 x N x K=16 instructions for N = 32 / 64 / 128.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -o umma_nosw_probe umma_nosw_probe.cu
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc_nosw(uint32_t addr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((addr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}
__host__ __device__ constexpr uint32_t idesc_f16(int M, int N) {
    return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc)
                 : "memory");
}
__device__ __forceinline__ void commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile("{\n\t.reg .pred p;\n\tW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@!p bra W;\n\t}" ::"r"(smem_u32(bar)),
                 "r"(parity)
                 : "memory");
}

constexpr int ROWS = 512;      // A rows available
// smem: A planes [2][ROWS][8] halves, B planes [2][128][8], results via global
__global__ void __launch_bounds__(128) probe(const __half* a_g, const __half* b_g, float* d_g, int N, int shift, int lbo_rows,
                                             long long* cyc, int reps) {
    extern __shared__ __align__(1024) unsigned char sm[];
    __half* sa = reinterpret_cast<__half*>(sm);                       // 2 * ROWS * 16 B = 16 KB
    __half* sb = reinterpret_cast<__half*>(sm + 2 * ROWS * 16);       // 2 * 128 * 16 B = 4 KB
    uint64_t* bar = reinterpret_cast<uint64_t*>(sm + 2 * ROWS * 16 + 4096);
    uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
    for (int i = threadIdx.x; i < 2 * ROWS * 8; i += 128) sa[i] = a_g[i];
    for (int i = threadIdx.x; i < 2 * 128 * 8; i += 128) sb[i] = b_g[i];
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 128;" ::"r"(smem_u32(slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *slot;
    if (threadIdx.x == 0) {
        // A: row stride 16 B, SBO 128 B, LBO = plane stride (or lbo_rows*16 for the "two taps in one K=16" trick)
        const uint32_t lbo_a = lbo_rows > 0 ? lbo_rows * 16 : ROWS * 16;
        const uint64_t da = desc_nosw(smem_u32(sa) + shift * 16, lbo_a, 128);
        const uint64_t db = desc_nosw(smem_u32(sb), 128 * 16, 128);
        const long long t0 = clock64();
        for (int r = 0; r < reps; ++r) mma(tmem, da, db, idesc_f16(128, N), r > 0 ? 1u : 0u);
        commit(bar);
        mbar_wait(bar, 0);
        cyc[0] = clock64() - t0;
    }
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // read back: warp w -> lanes 32w..; 32 columns at a time
    const int w = threadIdx.x >> 5;
    for (int c0 = 0; c0 < N; c0 += 32) {
        uint32_t r[32];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                     "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                     "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                       "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
                       "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
                       "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                     : "r"(tmem + ((uint32_t)(32 * w) << 16) + c0)
                     : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int i = 0; i < 32; ++i) d_g[(size_t)threadIdx.x * 128 + c0 + i] = __uint_as_float(r[i]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;" ::"r"(tmem) : "memory");
}

// conv0-like issue pattern: one A plane of 1664 rows (LBO = 2048 B = the next image row of the tile), six (kx, pr) taps,
// two M tiles, N = 32, accumulators at columns 0 and 32 -- timing only
__global__ void __launch_bounds__(128) probe_conv0(long long* cyc, int reps, int lbo_bytes, int ncols) {
    extern __shared__ __align__(1024) unsigned char sm[];
    __half* sa = reinterpret_cast<__half*>(sm);                        // 1664 + 512 rows * 16 B
    __half* sb = reinterpret_cast<__half*>(sm + 2176 * 16);            // 6 * 2 * 32 * 16 B = 6 KB
    uint64_t* bar = reinterpret_cast<uint64_t*>(sm + 2176 * 16 + 6144);
    uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
    for (int i = threadIdx.x; i < 2176 * 8; i += 128) sa[i] = __float2half((float)(i % 3));
    for (int i = threadIdx.x; i < 6144 / 2; i += 128) sb[i] = __float2half((float)(i % 5));
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 128;" ::"r"(smem_u32(slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *slot;
    if (threadIdx.x == 0) {
        const long long t0 = clock64();
        for (int r = 0; r < reps; ++r)
            for (int tt = 0; tt < 2; ++tt)
#pragma unroll
                for (int kx = 0; kx < 3; ++kx)
#pragma unroll
                    for (int pr = 0; pr < 2; ++pr) {
                        const uint32_t arow = (uint32_t)(((r % 5) * 2 + tt + 2 * pr) * 128 + kx);
                        mma(tmem + tt * ncols, desc_nosw(smem_u32(sa) + arow * 16, lbo_bytes, 128),
                            desc_nosw(smem_u32(sb) + (kx * 2 + pr) * 1024, 512, 128), idesc_f16(128, 32), (kx | pr) != 0);
                    }
        commit(bar);
        mbar_wait(bar, 0);
        cyc[0] = clock64() - t0;
    }
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;" ::"r"(tmem) : "memory");
}

// M = 64: where do the 64 accumulator rows land in TMEM, and what does an instruction cost?
__global__ void __launch_bounds__(128) probe_m64(const __half* a_g, const __half* b_g, float* d_g, int N, long long* cyc, int reps) {
    extern __shared__ __align__(1024) unsigned char sm[];
    __half* sa = reinterpret_cast<__half*>(sm);
    __half* sb = reinterpret_cast<__half*>(sm + 2 * ROWS * 16);
    uint64_t* bar = reinterpret_cast<uint64_t*>(sm + 2 * ROWS * 16 + 4096);
    uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
    for (int i = threadIdx.x; i < 2 * ROWS * 8; i += 128) sa[i] = a_g[i];
    for (int i = threadIdx.x; i < 2 * 128 * 8; i += 128) sb[i] = b_g[i];
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 128;" ::"r"(smem_u32(slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *slot;
    // poison all 128 lanes first (M=128 MMA of zeros is awkward: just overwrite after reading) -- read-back marks untouched lanes
    if (threadIdx.x == 0) {
        const uint64_t da = desc_nosw(smem_u32(sa) + 3 * 16, ROWS * 16, 128);
        const uint64_t db = desc_nosw(smem_u32(sb), 128 * 16, 128);
        const long long t0 = clock64();
        for (int r = 0; r < reps; ++r) mma(tmem, da, db, idesc_f16(64, N), r > 0 ? 1u : 0u);
        commit(bar);
        mbar_wait(bar, 0);
        cyc[0] = clock64() - t0;
    }
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int w = threadIdx.x >> 5;
    {
        uint32_t r[16];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 "
                     "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                       "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                     : "r"(tmem + ((uint32_t)(32 * w) << 16))
                     : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int i = 0; i < 16; ++i) d_g[(size_t)threadIdx.x * 16 + i] = __uint_as_float(r[i]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;" ::"r"(tmem) : "memory");
}

int main() {
    std::vector<__half> a(2 * ROWS * 8), b(2 * 128 * 8);
    std::vector<float> af(a.size()), bf(b.size());
    srand(7);
    for (size_t i = 0; i < a.size(); ++i) { af[i] = (float)(rand() % 9 - 4); a[i] = __float2half(af[i]); }
    for (size_t i = 0; i < b.size(); ++i) { bf[i] = (float)(rand() % 7 - 3); b[i] = __float2half(bf[i]); }
    __half *ag, *bg; float* dg; long long* cg;
    cudaMalloc(&ag, a.size() * 2); cudaMalloc(&bg, b.size() * 2); cudaMalloc(&dg, 128 * 128 * 4); cudaMalloc(&cg, 8);
    cudaMemcpy(ag, a.data(), a.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(bg, b.data(), b.size() * 2, cudaMemcpyHostToDevice);
    const int smem = 2 * ROWS * 16 + 4096 + 64;
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    int bad_total = 0;
    const int Ns[3] = {32, 64, 128};
    const int shifts[5] = {0, 1, 5, 14, 31};
    for (int lbo_rows : {0, 14}) for (int N : Ns) for (int shift : shifts) {
        probe<<<1, 128, smem>>>(ag, bg, dg, N, shift, lbo_rows, cg, 1);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("launch failed: %s\n", cudaGetErrorString(e)); return 1; }
        std::vector<float> d(128 * 128);
        cudaMemcpy(d.data(), dg, d.size() * 4, cudaMemcpyDeviceToHost);
        int bad = 0;
        for (int m = 0; m < 128; ++m) for (int n = 0; n < N; ++n) {
            float ref = 0.f;
            for (int k = 0; k < 16; ++k) {
                const int plane = k >> 3, e8 = k & 7;
                const int arow = lbo_rows > 0 ? (m + shift + plane * lbo_rows) : (m + shift);
                const int aplane = lbo_rows > 0 ? 0 : plane;
                ref += af[((size_t)aplane * ROWS + arow) * 8 + e8] * bf[((size_t)plane * 128 + n) * 8 + e8];
            }
            if (ref != d[(size_t)m * 128 + n]) ++bad;
        }
        printf("lbo_rows=%2d N=%3d shift=%2d mismatches=%d\n", lbo_rows, N, shift, bad);
        bad_total += bad;
    }
    for (int N : Ns) for (int reps : {64, 256}) {
        probe<<<1, 128, smem>>>(ag, bg, dg, N, 3, 0, cg, reps);
        cudaDeviceSynchronize();
        long long c; cudaMemcpy(&c, cg, 8, cudaMemcpyDeviceToHost);
        printf("N=%3d reps=%3d cycles=%lld per_mma=%.1f\n", N, reps, c, (double)c / reps);
    }
    {
        const int smem2 = 2176 * 16 + 6144 + 64;
        cudaFuncSetAttribute(probe_conv0, cudaFuncAttributeMaxDynamicSharedMemorySize, smem2);
        for (int lbo : {2048, 2064, 6272, 8192})
            for (int ncols : {32, 64}) {
                probe_conv0<<<1, 128, smem2>>>(cg, 20, lbo, ncols);
                cudaError_t e = cudaDeviceSynchronize();
                long long c; cudaMemcpy(&c, cg, 8, cudaMemcpyDeviceToHost);
                printf("conv0-like: LBO=%5d acc spacing=%2d cols: %s, %.1f cycles per MMA (240 MMAs)\n", lbo, ncols,
                       cudaGetErrorString(e), (double)c / 240.0);
            }
    }
    {
        cudaFuncSetAttribute(probe_m64, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        probe_m64<<<1, 128, smem>>>(ag, bg, dg, 32, cg, 1);
        cudaError_t e = cudaDeviceSynchronize();
        printf("M=64 probe: %s\n", cudaGetErrorString(e));
        std::vector<float> d(128 * 16);
        cudaMemcpy(d.data(), dg, d.size() * 4, cudaMemcpyDeviceToHost);
        // reference rows (shift 3), first 16 columns
        int lane_of_row[64];
        for (int m = 0; m < 64; ++m) {
            lane_of_row[m] = -1;
            float ref[16];
            for (int n = 0; n < 16; ++n) {
                ref[n] = 0.f;
                for (int k = 0; k < 16; ++k) ref[n] += af[((size_t)(k >> 3) * ROWS + m + 3) * 8 + (k & 7)] * bf[((size_t)(k >> 3) * 128 + n) * 8 + (k & 7)];
            }
            for (int l = 0; l < 128; ++l) {
                bool same = true;
                for (int n = 0; n < 16; ++n) same = same && d[(size_t)l * 16 + n] == ref[n];
                if (same) { lane_of_row[m] = l; break; }
            }
        }
        printf("M=64 row -> TMEM lane:");
        for (int m = 0; m < 64; ++m) printf(" %d", lane_of_row[m]);
        printf("\n");
        for (int N : Ns) {
            probe_m64<<<1, 128, smem>>>(ag, bg, dg, N, cg, 256);
            cudaDeviceSynchronize();
            long long c; cudaMemcpy(&c, cg, 8, cudaMemcpyDeviceToHost);
            printf("M=64 N=%3d: %.1f cycles per MMA\n", N, (double)c / 256);
        }
    }
    printf(bad_total ? "PROBE FAILED\n" : "PROBE OK\n");
    return bad_total != 0;
}
