// Integer-pipe microbenchmark for sm_100a (SURVEY.md section 8d caveat: is IMAD.WIDE full rate?).
// Each kernel runs ITER iterations of U independent dependent-chains per thread; reports ops/clk/SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o int_pipe int_pipe.cu && ./int_pipe
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITER 4096
#define U 8

__global__ void k_imad32(uint32_t* out, uint32_t a, uint32_t b)
{
    uint32_t x[U];
    for (int i = 0; i < U; i++) x[i] = threadIdx.x + i;
    for (int it = 0; it < ITER; it++) {
#pragma unroll
        for (int i = 0; i < U; i++) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(a), "r"(b));
    }
    uint32_t s = 0; for (int i = 0; i < U; i++) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_imadhi(uint32_t* out, uint32_t a, uint32_t b)
{
    uint32_t x[U];
    for (int i = 0; i < U; i++) x[i] = threadIdx.x + i;
    for (int it = 0; it < ITER; it++) {
#pragma unroll
        for (int i = 0; i < U; i++) asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(a), "r"(b));
    }
    uint32_t s = 0; for (int i = 0; i < U; i++) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_imadwide(uint32_t* out, uint32_t a, uint32_t b)
{
    uint64_t x[U];
    for (int i = 0; i < U; i++) x[i] = threadIdx.x + i;
    for (int it = 0; it < ITER; it++) {
#pragma unroll
        for (int i = 0; i < U; i++) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(x[i]) : "r"(a), "r"(b));
    }
    uint64_t s = 0; for (int i = 0; i < U; i++) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = (uint32_t)(s ^ (s >> 32));
}
// the pattern fe_mul uses: lo/hi pairs with carry chains (U/2 chains of 4 pairs)
__global__ void k_imadwide_cc(uint32_t* out, uint32_t a, uint32_t b)
{
    uint32_t x[2 * U];
    for (int i = 0; i < 2 * U; i++) x[i] = threadIdx.x + i;
    for (int it = 0; it < ITER; it++) {
#pragma unroll
        for (int c = 0; c < U; c += 4) {
            asm volatile("mad.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(x[2*c]), "+r"(x[2*c+1]) : "r"(a), "r"(b));
            asm volatile("madc.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(x[2*c+2]), "+r"(x[2*c+3]) : "r"(a), "r"(b));
            asm volatile("madc.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(x[2*c+4]), "+r"(x[2*c+5]) : "r"(a), "r"(b));
            asm volatile("madc.lo.cc.u32 %0, %2, %3, %0; madc.hi.u32 %1, %2, %3, %1;" : "+r"(x[2*c+6]), "+r"(x[2*c+7]) : "r"(a), "r"(b));
        }
    }
    uint32_t s = 0; for (int i = 0; i < 2 * U; i++) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_iadd3(uint32_t* out, uint32_t a, uint32_t b)
{
    uint32_t x[U];
    for (int i = 0; i < U; i++) x[i] = threadIdx.x + i;
    for (int it = 0; it < ITER; it++) {
#pragma unroll
        for (int i = 0; i < U; i++) asm volatile("add.u32 %0, %0, %1;" : "+r"(x[i]) : "r"(a));
    }
    uint32_t s = 0; for (int i = 0; i < U; i++) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + b;
}
__global__ void k_addc_chain(uint32_t* out, uint32_t a, uint32_t b)
{
    uint32_t x[U];
    for (int i = 0; i < U; i++) x[i] = threadIdx.x + i;
    for (int it = 0; it < ITER; it++) {
        asm volatile("add.cc.u32 %0, %0, %1;" : "+r"(x[0]) : "r"(a));
#pragma unroll
        for (int i = 1; i < U - 1; i++) asm volatile("addc.cc.u32 %0, %0, %1;" : "+r"(x[i]) : "r"(a));
        asm volatile("addc.u32 %0, %0, %1;" : "+r"(x[U - 1]) : "r"(b));
    }
    uint32_t s = 0; for (int i = 0; i < U; i++) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// mixed: one IMAD.WIDE (pair) + one add per slot, to see whether fma and alu pipes overlap
__global__ void k_mixed(uint32_t* out, uint32_t a, uint32_t b)
{
    uint64_t x[U]; uint32_t y[U];
    for (int i = 0; i < U; i++) { x[i] = threadIdx.x + i; y[i] = i; }
    for (int it = 0; it < ITER; it++) {
#pragma unroll
        for (int i = 0; i < U; i++) {
            asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(x[i]) : "r"(a), "r"(b));
            asm volatile("add.u32 %0, %0, %1;" : "+r"(y[i]) : "r"(a));
        }
    }
    uint64_t s = 0; for (int i = 0; i < U; i++) s += x[i] + y[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = (uint32_t)(s ^ (s >> 32));
}

template <typename K>
static void run(const char* name, K kern, int ops_per_iter, uint32_t* d_out, int sms, int threads, int blocks_per_sm)
{
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    int blocks = sms * blocks_per_sm;
    kern<<<blocks, threads>>>(d_out, 3, 5);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    kern<<<blocks, threads>>>(d_out, 3, 5);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    int clk_khz; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    double ops = (double)blocks * threads * ITER * ops_per_iter;
    printf("%-16s threads/SM=%4d  %8.3f ms  %8.2f Gops/s  %7.2f ops/clk/SM (at max clock %d MHz)\n", name, threads * blocks_per_sm, ms,
           ops / ms / 1e6, ops / (ms * 1e-3) / ((double)clk_khz * 1e3) / sms, clk_khz / 1000);
}

int main()
{
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    uint32_t* d_out; cudaMalloc(&d_out, 4 * 148 * 8 * 1024);
    for (int bps : {1, 2, 4}) {
        int th = 256;
        run("imad32", k_imad32, U, d_out, sms, th, bps);
        run("imad.hi", k_imadhi, U, d_out, sms, th, bps);
        run("imad.wide", k_imadwide, U, d_out, sms, th, bps);
        run("imad.wide.cc", k_imadwide_cc, U, d_out, sms, th, bps);
        run("iadd", k_iadd3, U, d_out, sms, th, bps);
        run("addc chain", k_addc_chain, U, d_out, sms, th, bps);
        run("wide+add", k_mixed, 2 * U, d_out, sms, th, bps);
    }
    return 0;
}
