// Random 64-byte gathers from an HBM-resident table: how many per second can a B200 serve?  (The verify kernels gather 28 table
// points per signature from 3.2 GB + 64 MiB/key tables; this measures the memory system's ceiling for that access pattern.)
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gather gather.cu ; run: ./gather
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t mix(uint32_t x) { x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16; return x; }

template <int BYTES, int DEP>
__global__ void gather_kernel(const uint4* __restrict__ tab, uint64_t n_entries, int per_thread, uint32_t* out, uint32_t seed)
{
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t acc = 0, h = mix(t ^ seed);
    for (int k = 0; k < per_thread; k++) {
        h = mix(h + 0x9e3779b9u + (DEP ? acc & 1u : 0u));                 // DEP: the next address depends on the loaded data (a dependent chain)
        const uint64_t e = ((uint64_t)h * n_entries) >> 32;
        const uint4* p = tab + e * 4;
        uint4 a = __ldg(p), b = __ldg(p + 1);
        acc ^= a.x ^ a.y ^ a.z ^ a.w ^ b.x ^ b.y ^ b.z ^ b.w;
        if (BYTES == 64) { uint4 c = __ldg(p + 2), d = __ldg(p + 3); acc ^= c.x ^ c.y ^ c.z ^ c.w ^ d.x ^ d.y ^ d.z ^ d.w; }
    }
    out[t] = acc;
}

template <int BYTES, int DEP> float run(const uint4* tab, uint64_t n_entries, int threads_total, int per_thread, uint32_t* out)
{
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    gather_kernel<BYTES, DEP><<<threads_total / 256, 256>>>(tab, n_entries, per_thread, out, 1);
    cudaEventRecord(e0);
    for (int r = 0; r < 5; r++) gather_kernel<BYTES, DEP><<<threads_total / 256, 256>>>(tab, n_entries, per_thread, out, 7 + r);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    return ms / 5;
}

int main()
{
    const size_t max_bytes = (size_t)7 << 30;
    uint4* tab; cudaMalloc(&tab, max_bytes); cudaMemset(tab, 1, max_bytes);
    uint32_t* out; cudaMalloc(&out, 4 << 20 << 2);
    printf("%-10s %-8s %-6s %-9s %10s %12s %10s\n", "table", "threads", "bytes", "pattern", "ms", "Ggather/s", "GB/s");
    const size_t sizes[] = {(size_t)64 << 20, (size_t)1 << 30, (size_t)3200 << 20, (size_t)7 << 30};
    for (size_t sz : sizes)
        for (int threads : {65536, 262144, 1048576})
            for (int variant = 0; variant < 3; variant++) {
                const uint64_t n = sz / 64;
                const int per = 28;
                float ms = variant == 0 ? run<64, 0>(tab, n, threads, per, out) : variant == 1 ? run<32, 0>(tab, n, threads, per, out) : run<64, 1>(tab, n, threads, per, out);
                const double g = (double)threads * per / (ms * 1e-3) / 1e9;
                printf("%-10zu %-8d %-6d %-9s %10.3f %12.2f %10.1f\n", sz >> 20, threads, variant == 1 ? 32 : 64, variant == 2 ? "dependent" : "indep", ms, g, g * (variant == 1 ? 32 : 64));
            }
    return 0;
}
