// microbenchmark: chains of field multiplications, 9x29 plain-IMAD.WIDE form vs the product's 8x32 carry-chained form
#include <cstdio>
#include <cuda_runtime.h>
#include "fe29.cuh"
#include "p256_fe.cuh"
using namespace fabgpu;
#define ITER 2000
__global__ void __launch_bounds__(128, 4) k29(int32_t* out, int seed)
{
    fe29 a, b;
    for (int i = 0; i < 9; i++) { a.v[i] = (threadIdx.x * 7919 + i * 104729 + seed) & F29_MASK; b.v[i] = (threadIdx.x * 31 + i * 7 + seed) & F29_MASK; }
    a.v[8] &= 0xffffff; b.v[8] &= 0xffffff;
#pragma unroll 1
    for (int it = 0; it < ITER; it++) {
        fe29 t = f29_mul(a, b);
        fe29 u = f29_sqr(b);
        a = f29_sub(t, u);      // lazy
        b = f29_mul(u, t);
    }
    int32_t s = 0; for (int i = 0; i < 9; i++) s ^= a.v[i] ^ b.v[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void __launch_bounds__(128, 4) k32(uint32_t* out, int seed)
{
    u256 a, b;
    for (int i = 0; i < 8; i++) { a.v[i] = threadIdx.x * 7919 + i * 104729 + seed; b.v[i] = threadIdx.x * 31 + i * 7 + seed; }
    a.v[7] &= 0x7fffffff; b.v[7] &= 0x7fffffff;
#pragma unroll 1
    for (int it = 0; it < ITER; it++) {
        u256 t = fe_mul_t<true>(a, b);
        u256 u = fe_sqr_t<true>(b);
        a = fe_sub(t, u);
        b = fe_mul_t<true>(u, t);
    }
    uint32_t s = 0; for (int i = 0; i < 8; i++) s ^= a.v[i] ^ b.v[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main()
{
    uint32_t* d; cudaMalloc(&d, 4 * 148 * 4 * 128);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int rep = 0; rep < 2; rep++) {
        float ms;
        cudaEventRecord(e0); k32<<<148 * 4, 128>>>(d, rep); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
        printf("8x32 carry-chained : %.3f ms  -> %.1f G mul-equivalents/s\n", ms, 148.0 * 4 * 128 * ITER * 3 / ms / 1e6);
        cudaEventRecord(e0); k29<<<148 * 4, 128>>>((int32_t*)d, rep); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
        printf("9x29 plain wide    : %.3f ms  -> %.1f G mul-equivalents/s\n", ms, 148.0 * 4 * 128 * ITER * 3 / ms / 1e6);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
