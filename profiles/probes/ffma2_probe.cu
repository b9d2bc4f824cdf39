// Hardware probe: throughput and dependent-issue latency of fp32 FFMA against the packed FFMA2 (PTX fma.rn.f32x2,
// sm_100+) on B200.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2_probe ffma2_probe.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint64_t ffma2(uint64_t a, uint64_t b, uint64_t c) {
    uint64_t d;
    asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ float ffma1(float a, float b, float c) {
    float d;
    asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
    return d;
}

template <int CHAINS>
__global__ void k_ffma(float* out, float a, float b, int iters, long long* cyc) {
    float acc[CHAINS];
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) acc[c] = (float)(threadIdx.x + c);
    __syncthreads();
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int c = 0; c < CHAINS; ++c) acc[c] = ffma1(acc[c], a, b);
    }
    const long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) s += acc[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int CHAINS>
__global__ void k_ffma2(float* out, float a, float b, int iters, long long* cyc) {
    uint64_t acc[CHAINS];
    const float2 av = make_float2(a, a), bv = make_float2(b, b);
    const uint64_t a2 = *reinterpret_cast<const uint64_t*>(&av), b2 = *reinterpret_cast<const uint64_t*>(&bv);
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) {
        float2 v = make_float2((float)(threadIdx.x + c), (float)(threadIdx.x - c));
        acc[c] = *reinterpret_cast<uint64_t*>(&v);
    }
    __syncthreads();
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int c = 0; c < CHAINS; ++c) acc[c] = ffma2(acc[c], a2, b2);
    }
    const long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) {
        float2 v = *reinterpret_cast<float2*>(&acc[c]);
        s += v.x + v.y;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <typename F>
static void run(const char* name, F launch, int threads, int chains, int lanes_per_instr) {
    float* out; long long* cyc; long long h = 0;
    cudaMalloc(&out, sizeof(float) * 148 * 1024); cudaMalloc(&cyc, 8);
    const int iters = 4096;
    launch(out, cyc, iters, threads); launch(out, cyc, iters, threads);
    cudaDeviceSynchronize();
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    const double instr_per_warp = (double)iters * chains;
    const double warps = threads / 32.0;
    printf("%-8s threads/SM %4d chains %d: %8lld cycles, %.2f cycles per instruction per warp, %.1f fp32 FMA/clk/SM\n",
           name, threads, chains, h, h / instr_per_warp, instr_per_warp * warps * 32 * lanes_per_instr / h);
    cudaFree(out); cudaFree(cyc);
}

int main() {
    for (int threads : {32, 128, 256, 512, 1024}) {
        run("FFMA", [](float* o, long long* c, int it, int th) { k_ffma<8><<<148, th>>>(o, 1.0001f, 0.5f, it, c); }, threads, 8, 1);
        run("FFMA2", [](float* o, long long* c, int it, int th) { k_ffma2<8><<<148, th>>>(o, 1.0001f, 0.5f, it, c); }, threads, 8, 2);
    }
    run("FFMA", [](float* o, long long* c, int it, int th) { k_ffma<1><<<148, th>>>(o, 1.0001f, 0.5f, it, c); }, 32, 1, 1);
    run("FFMA2", [](float* o, long long* c, int it, int th) { k_ffma2<1><<<148, th>>>(o, 1.0001f, 0.5f, it, c); }, 32, 1, 2);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
