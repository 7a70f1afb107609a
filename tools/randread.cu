// randread.cu — microbenchmark behind the block-size / thread-mapping decision in DESIGN.md:
// dependent chains of random block reads over a multi-GB buffer (what an FM-index walk does),
// for 32/64/128-byte blocks, thread-per-chain vs 8-lane-group-per-chain, at several occupancies.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/randread tools/randread.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

__global__ void fill(uint32_t* p, uint64_t n)
{
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint64_t x = i * 0x9E3779B97F4A7C15ull;
    x ^= x >> 29; x *= 0xBF58476D1CE4E5B9ull; x ^= x >> 32;
    p[i] = (uint32_t)x;
}

__device__ __forceinline__ uint32_t ld256_sum(const void* p)
{
    uint32_t a, b, c, d, e, f, g, h;
    asm volatile("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(a), "=r"(b), "=r"(c), "=r"(d), "=r"(e), "=r"(f), "=r"(g), "=r"(h) : "l"(p));
    return a ^ b ^ c ^ d ^ e ^ f ^ g ^ h;
}

// one thread = one chain; BYTES per hop as BYTES/32 256-bit loads; MLP independent chains per thread
template <int BYTES, int MLP>
__global__ void chase_thread(const uint8_t* __restrict__ buf, uint64_t nblk, int iters, uint32_t* sink)
{
    uint64_t idx[MLP];
    uint32_t acc = 0;
    for (int m = 0; m < MLP; ++m) idx[m] = ((uint64_t)(blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + m * 40503u) % nblk;
    for (int it = 0; it < iters; ++it) {
        uint32_t v[MLP];
#pragma unroll
        for (int m = 0; m < MLP; ++m) {
            v[m] = 0;
#pragma unroll
            for (int k = 0; k < BYTES / 32; ++k) v[m] ^= ld256_sum(buf + idx[m] * BYTES + k * 32);
        }
#pragma unroll
        for (int m = 0; m < MLP; ++m) {
            acc += v[m];
            idx[m] = ((uint64_t)v[m] * 0x9E3779B1u + idx[m]) % nblk;
        }
    }
    if (acc == 0x12345678u) sink[0] = acc;
}

// GROUP lanes = one chain; each lane loads 16 bytes of a GROUP*16-byte block
template <int GROUP>
__global__ void chase_group(const uint8_t* __restrict__ buf, uint64_t nblk, int iters, uint32_t* sink)
{
    const int lane = threadIdx.x % GROUP;
    uint64_t idx = ((uint64_t)((blockIdx.x * blockDim.x + threadIdx.x) / GROUP) * 2654435761u) % nblk;
    uint32_t acc = 0;
    for (int it = 0; it < iters; ++it) {
        const uint4 q = __ldg(reinterpret_cast<const uint4*>(buf + idx * (GROUP * 16) + lane * 16));
        uint32_t v = q.x ^ q.y ^ q.z ^ q.w;
#pragma unroll
        for (int o = GROUP / 2; o > 0; o >>= 1) v ^= __shfl_xor_sync(0xffffffffu, v, o);
        acc += v;
        idx = ((uint64_t)v * 0x9E3779B1u + idx) % nblk;
    }
    if (acc == 0x12345678u) sink[0] = acc;
}

template <class F>
double time_ms(F f)
{
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    f();  // warm-up
    cudaDeviceSynchronize();
    cudaEventRecord(a);
    f();
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms = 0;
    cudaEventElapsedTime(&ms, a, b);
    return ms;
}

int main(int argc, char** argv)
{
    const uint64_t bytes = (argc > 1 ? atoll(argv[1]) : 2048ll) << 20;
    const int iters = argc > 2 ? atoi(argv[2]) : 400;
    uint8_t* buf; uint32_t* sink;
    cudaMalloc(&buf, bytes); cudaMalloc(&sink, 4);
    fill<<<(unsigned)((bytes / 4 + 255) / 256), 256>>>((uint32_t*)buf, bytes / 4);
    cudaDeviceSynchronize();
    cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
    const int sms = prop.multiProcessorCount;
    printf("device %s, %d SMs, buffer %.1f GB, %d hops per chain\n", prop.name, sms, bytes / 1e9, iters);
    printf("%-28s %8s %10s %10s %10s\n", "variant", "thr/SM", "Ghops/s", "GB/s", "ms");
    for (int per_sm : {256, 512, 1024, 2048}) {
        const int blocks = sms * per_sm / 256;
#define RUN_T(B, M)                                                                                        \
        {                                                                                                  \
            double ms = time_ms([&] { chase_thread<B, M><<<blocks, 256>>>(buf, bytes / B, iters, sink); }); \
            double hops = (double)blocks * 256 * M * iters;                                                \
            printf("thread %3dB mlp%-2d             %8d %10.2f %10.1f %10.2f\n", B, M, per_sm, hops / ms / 1e6, hops * B / ms / 1e6, ms); \
        }
        RUN_T(32, 1) RUN_T(64, 1) RUN_T(128, 1) RUN_T(64, 2) RUN_T(64, 4) RUN_T(32, 4)
#define RUN_G(G)                                                                                           \
        {                                                                                                  \
            double ms = time_ms([&] { chase_group<G><<<blocks, 256>>>(buf, bytes / (G * 16), iters, sink); }); \
            double hops = (double)blocks * 256 / G * iters;                                                \
            printf("group%-2d x16B = %3dB           %8d %10.2f %10.1f %10.2f\n", G, G * 16, per_sm, hops / ms / 1e6, hops * G * 16 / ms / 1e6, ms); \
        }
        RUN_G(4) RUN_G(8)
    }
    return 0;
}
