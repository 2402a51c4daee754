// tools/probe_bench.cu -- random-sector microbenchmark used to choose the table-probe load flavour and to find the
// DRAM fetch granularity on B200 (see DESIGN.md "probe roofline").  Not part of the product library.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o tools/probe_bench tools/probe_bench.cu
//   tools/probe_bench [buffer MiB] [flavour | -1 = all] [l2 fetch granularity 0|32|64|128]
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ uint64_t mix(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

template <int F>
__device__ __forceinline__ uint32_t load_sector(const uint4* p) {
    uint32_t a, b, c, d, e, f, g, h;
    if (F == 0) {
        asm volatile("ld.global.nc.L1::no_allocate.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d), "=r"(e), "=r"(f), "=r"(g), "=r"(h) : "l"(p));
    } else if (F == 1) {
        asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "l"(p));
        asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(e), "=r"(f), "=r"(g), "=r"(h) : "l"(p + 1));
    } else if (F == 2) {
        asm volatile("ld.global.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d), "=r"(e), "=r"(f), "=r"(g), "=r"(h) : "l"(p));
    } else if (F == 3) {
        asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "l"(p));
        asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(e), "=r"(f), "=r"(g), "=r"(h) : "l"(p + 1));
    } else if (F == 4) {
        asm volatile("ld.global.cs.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "l"(p));
        asm volatile("ld.global.cs.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(e), "=r"(f), "=r"(g), "=r"(h) : "l"(p + 1));
    } else if (F == 5) {
        asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "l"(p));
        asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(e), "=r"(f), "=r"(g), "=r"(h) : "l"(p + 1));
    } else if (F == 6) { // only half a sector: does the request size matter?
        asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "l"(p));
        e = f = g = h = 0;
    } else if (F == 7) { // L2 evict_first policy
        uint64_t pol;
        asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
        asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8], %9;" : "=r"(a), "=r"(b), "=r"(c), "=r"(d), "=r"(e), "=r"(f), "=r"(g), "=r"(h) : "l"(p), "l"(pol));
    } else if (F == 8) {
        asm volatile("ld.global.lu.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "l"(p));
        asm volatile("ld.global.lu.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(e), "=r"(f), "=r"(g), "=r"(h) : "l"(p + 1));
    } else { // F == 9: L2 evict_no_allocate style
        uint64_t pol;
        asm volatile("createpolicy.fractional.L2::evict_unchanged.b64 %0, 1.0;" : "=l"(pol));
        asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8], %9;" : "=r"(a), "=r"(b), "=r"(c), "=r"(d), "=r"(e), "=r"(f), "=r"(g), "=r"(h) : "l"(p), "l"(pol));
    }
    return a ^ b ^ c ^ d ^ e ^ f ^ g ^ h;
}

template <int F, int U>
__global__ void k_bench(const uint4* __restrict__ buf, uint64_t n_sectors, uint32_t per_thread, uint32_t* sink) {
    const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t acc = 0;
    for (uint32_t it = 0; it < per_thread; it += U) {
        uint32_t v[U];
#pragma unroll
        for (int k = 0; k < U; k++) {
            uint64_t h = mix(tid * 0x100000001B3ull + it + k);
            uint64_t idx = ((h >> 32) * n_sectors) >> 32;
            v[k] = load_sector<F>(buf + 2 * idx);
        }
#pragma unroll
        for (int k = 0; k < U; k++) acc ^= v[k];
    }
    if (acc == 0x9E3779B9u) sink[0] = acc;
}

template <int F>
double run(const uint4* buf, uint64_t n_sectors, int tpb, uint32_t* sink) {
    const uint32_t per_thread = 32;
    const uint64_t n_loads = 1ull << 28;
    unsigned grid = (unsigned)(n_loads / per_thread / tpb);
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a));
    CK(cudaEventCreate(&b));
    k_bench<F, 4><<<grid, tpb>>>(buf, n_sectors, per_thread, sink);
    CK(cudaEventRecord(a));
    k_bench<F, 4><<<grid, tpb>>>(buf, n_sectors, per_thread, sink);
    CK(cudaEventRecord(b));
    CK(cudaEventSynchronize(b));
    float ms;
    CK(cudaEventElapsedTime(&ms, a, b));
    return (double)grid * tpb * per_thread / (ms * 1e-3);
}

int main(int argc, char** argv) {
    size_t mib = argc > 1 ? (size_t)atol(argv[1]) : 1536;
    int flavour = argc > 2 ? atoi(argv[2]) : -1;
    int gran = argc > 3 ? atoi(argv[3]) : 0;
    if (gran) CK(cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)gran));
    size_t got = 0;
    CK(cudaDeviceGetLimit(&got, cudaLimitMaxL2FetchGranularity));
    size_t bytes = mib << 20;
    uint4* buf;
    uint32_t* sink;
    CK(cudaMalloc(&buf, bytes));
    CK(cudaMalloc(&sink, 4));
    CK(cudaMemset(buf, 0x5A, bytes));
    uint64_t n_sectors = bytes / 32;
    printf("buffer %zu MiB, L2 fetch granularity limit %zu\n", mib, got);
    const char* names[10] = {"nc.no_allocate.v8", "nc.v4 x2", "plain v8", "cg.v4 x2", "cs.v4 x2", "volatile.v4 x2", "nc.no_allocate.v4 (16 B)",
                             "nc.v8 + L2 evict_first", "lu.v4 x2", "nc.v8 + L2 evict_unchanged"};
#define RUN(F) if (flavour < 0 || flavour == F) printf("  flavour %d %-28s %.3e sectors/s  (%.0f GB/s of 32-B sectors)\n", F, names[F], run<F>(buf, n_sectors, 256, sink), run<F>(buf, n_sectors, 256, sink) * 32e-9);
    RUN(0) RUN(1) RUN(2) RUN(3) RUN(4) RUN(5) RUN(6) RUN(7) RUN(8) RUN(9)
    return 0;
}
