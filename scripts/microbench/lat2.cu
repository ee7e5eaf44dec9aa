// Load-latency probes: L1 hit, L2 hit, effect of prefetch.global.L1 / L2, load after own store.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ long long clk() { long long c; asm volatile("mov.u64 %0, %%clock64;" : "=l"(c)); return c; }
__device__ __forceinline__ int ld_ca(const int* p) { int v; asm volatile("ld.global.ca.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ int ld_cg(const int* p) { int v; asm volatile("ld.global.cg.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ void pf_l1(const int* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }
__device__ __forceinline__ void pf_l2(const int* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__global__ void k(int* g, long long* out, int stride) {
    // every lane uses its own lines (stride ints apart), like one tree per lane
    int* base = g + (size_t)threadIdx.x * stride;
    long long t0, t1; int v = 0, acc = 0;
    // warm: touch line 0 twice
    v = ld_ca(base); acc += v;
    t0 = clk(); v = ld_ca(base + (acc & 1)); acc += v; t1 = clk(); if (threadIdx.x == 0) out[0] = t1 - t0;          // L1 hit
    t0 = clk(); v = ld_cg(base + 32 * 10 + (acc & 1)); acc += v; t1 = clk(); if (threadIdx.x == 0) out[1] = t1 - t0; // cold line: DRAM/L2 miss
    t0 = clk(); v = ld_cg(base + 32 * 10 + (acc & 1)); acc += v; t1 = clk(); if (threadIdx.x == 0) out[2] = t1 - t0; // L2 hit
    // prefetch L1, wait, load
    pf_l1(base + 32 * 20);
    for (int i = 0; i < 300; ++i) acc = acc * 3 + 1;
    t0 = clk(); v = ld_ca(base + 32 * 20 + (acc & 1)); acc += v; t1 = clk(); if (threadIdx.x == 0) out[3] = t1 - t0;
    // prefetch L2 (cold line), wait, load
    pf_l2(base + 32 * 30);
    for (int i = 0; i < 300; ++i) acc = acc * 3 + 1;
    t0 = clk(); v = ld_ca(base + 32 * 30 + (acc & 1)); acc += v; t1 = clk(); if (threadIdx.x == 0) out[4] = t1 - t0;
    // load after own store to the same line
    v = ld_ca(base + 32 * 40); acc += v; v = ld_ca(base + 32 * 40 + 1); acc += v;
    base[32 * 40 + 8] = acc;
    t0 = clk(); v = ld_ca(base + 32 * 40 + (acc & 1)); acc += v; t1 = clk(); if (threadIdx.x == 0) out[5] = t1 - t0;
    // again (was it re-cached in L1 by that load?)
    t0 = clk(); v = ld_ca(base + 32 * 40 + (acc & 1)); acc += v; t1 = clk(); if (threadIdx.x == 0) out[6] = t1 - t0;
    // store then prefetch.L1 then wait then load
    base[32 * 50 + 8] = acc; pf_l1(base + 32 * 50);
    for (int i = 0; i < 300; ++i) acc = acc * 3 + 1;
    t0 = clk(); v = ld_ca(base + 32 * 50 + (acc & 1)); acc += v; t1 = clk(); if (threadIdx.x == 0) out[7] = t1 - t0;
    // a real load used as prefetch (result unused until later), then a second load of the same line
    int early = ld_ca(base + 32 * 60);
    for (int i = 0; i < 300; ++i) acc = acc * 3 + 1;
    t0 = clk(); v = ld_ca(base + 32 * 60 + 1 + (acc & 1)); acc += v; t1 = clk(); if (threadIdx.x == 0) out[8] = t1 - t0;
    g[threadIdx.x] = acc + early;
}
int main() {
    int* g; long long* out; const int stride = 1 << 16;
    cudaMalloc(&g, (size_t)32 * stride * 4 + 4096); cudaMalloc(&out, 16 * 8);
    cudaMemset(g, 0, (size_t)32 * stride * 4);
    const char* names[] = {"ld.ca L1 hit", "ld.cg cold (DRAM)", "ld.cg L2 hit", "prefetch.L1 -> ld.ca", "prefetch.L2 (cold) -> ld.ca", "own store -> ld.ca", "  ... second ld.ca", "store, prefetch.L1 -> ld.ca", "early real load -> ld.ca"};
    for (int rep = 0; rep < 2; ++rep) {
        cudaMemset(out, 0, 128);
        k<<<1, 32>>>(g + rep * 7, out, stride);
        long long c[16]; cudaMemcpy(c, out, 128, cudaMemcpyDeviceToHost);
        printf("run %d (includes ~20 cycles of clock/measurement overhead)\n", rep);
        for (int i = 0; i < 9; ++i) printf("  %-30s %lld cycles\n", names[i], c[i]);
    }
    printf("err=%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
}
