// clusterread.cu — microbenchmark for the question "would table keys that differ only in their low-order characters
// be cheaper to read than keys scattered over the whole table?" (DESIGN.md §7b).  Every thread reads, per round, 16
// independent 16-byte entries (four in flight) of a 32 GB table: either anywhere (what the substituted keys of a
// rightwards search do today: they differ in the HIGH-order characters of the key), or inside one random region of
// 4 KB / 64 KB / 1 MB (what they would do if the key's character order were reversed for that search).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/clusterread tools/clusterread.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

__device__ __forceinline__ uint32_t ld16(const void* p)
{
    uint32_t a, b, c, d;
    asm volatile("ld.global.nc.L2::64B.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "l"(p));
    return a ^ b ^ c ^ d;
}

__device__ __forceinline__ uint64_t mix(uint64_t x)
{
    x ^= x >> 29; x *= 0xBF58476D1CE4E5B9ull; x ^= x >> 32; x *= 0x94D049BB133111EBull; x ^= x >> 29;
    return x;
}

// region_entries: entries per region (a power of two); n_entries: entries in the table
__global__ void __launch_bounds__(256, 4) reads(const uint8_t* __restrict__ buf, uint64_t n_entries, uint64_t region_entries, int rounds, uint32_t* sink)
{
    const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t acc = 0;
    for (int r = 0; r < rounds; ++r) {
        const uint64_t h = mix(tid * 1315423911ull + (uint64_t)r * 0x9E3779B97F4A7C15ull + acc);
        const uint64_t base = (h % (n_entries / region_entries)) * region_entries;
        for (int k = 0; k < 16; k += 4) {
            uint32_t v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) v[u] = ld16(buf + (base + (mix(h + k + u) & (region_entries - 1))) * 16);
            acc += v[0] ^ v[1] ^ v[2] ^ v[3];
        }
    }
    if (acc == 0x12345678u) sink[0] = acc;
}

int main()
{
    const uint64_t bytes = 32ull << 30, n_entries = bytes / 16;
    uint8_t* buf; uint32_t* sink;
    cudaMalloc(&buf, bytes); cudaMalloc(&sink, 4);
    cudaMemset(buf, 1, bytes);
    cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
    const int grid = prop.multiProcessorCount * 4, rounds = 64;
    printf("device %s, table %.0f GB of 16-byte entries, %d threads x %d rounds x 16 reads\n", prop.name, bytes / 1e9, grid * 256, rounds);
    printf("%-28s %10s %10s\n", "16 reads of a round within", "G reads/s", "ms");
    const uint64_t regions[] = {n_entries, 1ull << 16, 1ull << 12, 1ull << 8, 1ull << 6};
    const char* names[] = {"the whole table", "1 MB (4^8 entries)", "64 KB (4^6 entries)", "4 KB (4^4 entries)", "1 KB (4^3 entries)"};
    for (int i = 0; i < 5; ++i) {
        cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
        reads<<<grid, 256>>>(buf, n_entries, regions[i], 4, sink);
        cudaDeviceSynchronize();
        cudaEventRecord(a);
        reads<<<grid, 256>>>(buf, n_entries, regions[i], rounds, sink);
        cudaEventRecord(b); cudaEventSynchronize(b);
        float ms = 0; cudaEventElapsedTime(&ms, a, b);
        printf("%-28s %10.2f %10.2f\n", names[i], (double)grid * 256 * rounds * 16 / ms / 1e6, ms);
    }
    return 0;
}
