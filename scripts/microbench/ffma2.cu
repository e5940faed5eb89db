// Microbenchmark: issue cost of packed fp32x2 FMA (SASS FFMA2) vs scalar FFMA on sm_100a, alone and mixed with
// integer / shared-memory instructions.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2 ffma2.cu
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c)
{
    u64 d;
    asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ float fma1(float a, float b, float c)
{
    float d;
    asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
    return d;
}
__device__ __forceinline__ unsigned iadd(unsigned a, unsigned b)
{
    unsigned d;
    asm volatile("add.u32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
    return d;
}
constexpr int ITERS = 4096;
// MODE 0: 16 scalar FFMA / iter; 1: 8 FFMA2 / iter (same flops); 2: 16 FFMA + 8 IADD; 3: 8 FFMA2 + 8 IADD;
// 4: 8 FFMA2 + 16 IADD; 5: 16 FFMA2 (2x flops); 6: 8 FFMA2 + 8 LDS
template <int MODE>
__global__ void k(float* out, long long* cyc)
{
    __shared__ float sm[1024];
    sm[threadIdx.x] = threadIdx.x;
    __syncthreads();
    float    a[16];
    u64      p[16];
    unsigned n[16];
    for (int i = 0; i < 16; ++i) {
        a[i] = threadIdx.x * 0.001f + i;
        p[i] = ((u64)__float_as_uint(a[i]) << 32) | __float_as_uint(a[i] + 1);
        n[i] = threadIdx.x + i;
    }
    const float b = 1.0001f, c = 0.5f;
    const u64   b2 = ((u64)__float_as_uint(b) << 32) | __float_as_uint(b), c2 = ((u64)__float_as_uint(c) << 32) | __float_as_uint(c);
    float       acc = 0;
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
        if (MODE == 0 || MODE == 2) {
#pragma unroll
            for (int i = 0; i < 16; ++i) a[i] = fma1(a[i], b, c);
        }
        if (MODE == 1 || MODE == 3 || MODE == 4 || MODE == 6) {
#pragma unroll
            for (int i = 0; i < 8; ++i) p[i] = fma2(p[i], b2, c2);
        }
        if (MODE == 5) {
#pragma unroll
            for (int i = 0; i < 16; ++i) p[i] = fma2(p[i], b2, c2);
        }
        if (MODE == 2 || MODE == 3) {
#pragma unroll
            for (int i = 0; i < 8; ++i) n[i] = iadd(n[i], 3);
        }
        if (MODE == 4) {
#pragma unroll
            for (int i = 0; i < 16; ++i) n[i] = iadd(n[i], 3);
        }
        if (MODE == 6) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                float v;
                asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"((unsigned)__cvta_generic_to_shared(&sm[(n[i] & 1023)])));
                acc += v;
            }
        }
    }
    long long t1 = clock64();
    float s = acc;
    for (int i = 0; i < 16; ++i) s += a[i] + __uint_as_float((unsigned)p[i]) + __uint_as_float((unsigned)(p[i] >> 32)) + n[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int MODE>
void run(const char* what)
{
    float*     out;
    long long* cyc;
    const int  blocks = 148 * 2, threads = 1024;  // 2048 threads / SM: 16 warps per scheduler
    cudaMalloc(&out, blocks * threads * 4);
    cudaMalloc(&cyc, blocks * 8);
    k<MODE><<<blocks, threads>>>(out, cyc);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0), cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<MODE><<<blocks, threads>>>(out, cyc);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    long long h[296];
    cudaMemcpy(h, cyc, blocks * 8, cudaMemcpyDeviceToHost);
    double avg = 0;
    for (int i = 0; i < blocks; ++i) avg += h[i];
    avg /= blocks;
    // per scheduler: 16 warps resident; cycles per iteration per scheduler = avg / ITERS; per warp-iteration /16
    printf("%-34s %8.3f ms  %8.2f cycles/iter/block  -> %6.2f issue-cycles per warp-iteration (16 warps/scheduler)\n", what, ms,
           avg / ITERS, avg / ITERS / 16.0);
}
int main()
{
    run<0>("16 FFMA");
    run<1>("8 FFMA2 (same flops)");
    run<5>("16 FFMA2 (2x flops)");
    run<2>("16 FFMA + 8 IADD");
    run<3>("8 FFMA2 + 8 IADD");
    run<4>("8 FFMA2 + 16 IADD");
    run<6>("8 FFMA2 + 8 LDS + 8 FADD");
    return 0;
}
