// ldflavor.cu — which global-load flavour fetches only the 32-byte sector it needs?
// Dependent chains of random 32-byte reads over a multi-GB buffer, one chain per thread, for several
// PTX load qualifiers.  Run under `ncu --metrics dram__bytes_read.sum,lts__t_sectors_srcunit_tex_op_read.sum,
// l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum` to see the DRAM / L2 traffic each flavour really causes.
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

#define LD256(NAME, QUAL)                                                                                   \
    __device__ __forceinline__ uint32_t NAME(const void* p)                                                 \
    {                                                                                                       \
        uint32_t a, b, c, d, e, f, g, h;                                                                    \
        asm volatile("ld.global" QUAL ".v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"                           \
                     : "=r"(a), "=r"(b), "=r"(c), "=r"(d), "=r"(e), "=r"(f), "=r"(g), "=r"(h) : "l"(p));    \
        return a ^ b ^ c ^ d ^ e ^ f ^ g ^ h;                                                               \
    }
LD256(ld_nc, ".nc")
LD256(ld_ca, "")
LD256(ld_cg, ".cg")
LD256(ld_cs, ".cs")
LD256(ld_cv, ".cv")
LD256(ld_nc_noalloc, ".nc.L1::no_allocate")
LD256(ld_nc_evict_first, ".nc.L1::evict_first")
LD256(ld_nc_l2_64, ".nc.L2::64B")
LD256(ld_nc_l2_128, ".nc.L2::128B")
LD256(ld_noalloc_l2_64, ".nc.L1::no_allocate.L2::64B")

template <int F>
__global__ void chase(const uint8_t* __restrict__ buf, uint64_t nblk, int iters, uint32_t* sink)
{
    uint64_t idx = ((uint64_t)(blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u) % nblk;
    uint32_t acc = 0;
    for (int it = 0; it < iters; ++it) {
        const void* p = buf + idx * 32;
        uint32_t v;
        if (F == 0) v = ld_nc(p);
        else if (F == 1) v = ld_ca(p);
        else if (F == 2) v = ld_cg(p);
        else if (F == 3) v = ld_cs(p);
        else if (F == 4) v = ld_cv(p);
        else if (F == 5) v = ld_nc_noalloc(p);
        else if (F == 6) v = ld_nc_evict_first(p);
        else if (F == 7) v = ld_nc_l2_64(p);
        else if (F == 8) v = ld_nc_l2_128(p);
        else v = ld_noalloc_l2_64(p);
        acc += v;
        idx = ((uint64_t)v * 0x9E3779B1u + idx) % nblk;
    }
    if (acc == 0x12345678u) sink[0] = acc;
}

template <int F>
void run(const char* name, const uint8_t* buf, uint64_t bytes, int iters, uint32_t* sink, int blocks)
{
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    chase<F><<<blocks, 256>>>(buf, bytes / 32, iters, sink);
    cudaDeviceSynchronize();
    cudaEventRecord(a);
    chase<F><<<blocks, 256>>>(buf, bytes / 32, iters, sink);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms = 0;
    cudaEventElapsedTime(&ms, a, b);
    double hops = (double)blocks * 256 * iters;
    printf("%-28s %8.2f Ghops/s %8.1f GB/s(32B) %8.2f ms\n", name, hops / ms / 1e6, hops * 32 / ms / 1e6, ms);
}

int main(int argc, char** argv)
{
    const uint64_t bytes = (argc > 1 ? atoll(argv[1]) : 3072ll) << 20;
    const int iters = argc > 2 ? atoi(argv[2]) : 200;
    uint8_t* buf; uint32_t* sink;
    cudaMalloc(&buf, bytes); cudaMalloc(&sink, 4);
    fill<<<(unsigned)((bytes / 4 + 255) / 256), 256>>>((uint32_t*)buf, bytes / 4);
    cudaDeviceSynchronize();
    cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
    const int blocks = prop.multiProcessorCount * 4;
    printf("buffer %.1f GB, %d hops per chain, %d threads\n", bytes / 1e9, iters, blocks * 256);
    run<0>("ld.global.nc", buf, bytes, iters, sink, blocks);
    run<1>("ld.global (ca)", buf, bytes, iters, sink, blocks);
    run<2>("ld.global.cg", buf, bytes, iters, sink, blocks);
    run<3>("ld.global.cs", buf, bytes, iters, sink, blocks);
    run<4>("ld.global.cv", buf, bytes, iters, sink, blocks);
    run<5>("ld.nc.L1::no_allocate", buf, bytes, iters, sink, blocks);
    run<6>("ld.nc.L1::evict_first", buf, bytes, iters, sink, blocks);
    run<7>("ld.nc.L2::64B", buf, bytes, iters, sink, blocks);
    run<8>("ld.nc.L2::128B", buf, bytes, iters, sink, blocks);
    run<9>("ld.nc.no_allocate.L2::64B", buf, bytes, iters, sink, blocks);
    return 0;
}
