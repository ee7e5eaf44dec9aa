// Does prefetch.global.L2 / .L1 shorten a later load? (cold lines from a 4 GB buffer, 32 lanes x distinct lines)
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ long long clk() { long long c; asm volatile("mov.u64 %0, %%clock64;" : "=l"(c)::"memory"); return c; }
__device__ __forceinline__ int ld_ca(const int* p) { int v; asm volatile("ld.global.ca.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ int ld_cg(const int* p) { int v; asm volatile("ld.global.cg.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ void pf_l1(const int* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p) : "memory"); }
__device__ __forceinline__ void pf_l2(const int* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p) : "memory"); }
__device__ __forceinline__ long long clk_after(int dep) { long long c; asm volatile("{ .reg .b32 t; mov.b32 t, %1; mov.u64 %0, %%clock64; }" : "=l"(c) : "r"(dep) : "memory"); return c; }
__device__ __forceinline__ void spin(long long cycles) { const long long t = clk(); while (clk() - t < cycles) {} }
__global__ void k(int* g, long long* out, size_t lane_stride, size_t test_stride) {
    int* base = g + (size_t)threadIdx.x * lane_stride;
    long long t0, t1; int v, acc = 0;
    for (int test = 0; test < 6; ++test) {
        int* p = base + (size_t)test * test_stride;
        if (test == 1) pf_l2(p);
        if (test == 2) pf_l1(p);
        if (test == 3) { v = ld_cg(p + 8); acc += v; }      // a real load of another word of the line, result consumed before the spin
        if (test == 4) { v = ld_ca(p + 8); acc += v; }
        if (test == 5) { p[9] = acc; pf_l2(p); }             // own store, then L2 prefetch
        spin(4000 + (acc & 1));
        t0 = clk();
        v = (test == 2 || test == 4) ? ld_ca(p + (acc & 1)) : ld_cg(p + (acc & 1));
        if (v == 0x12345678) out[7] = 1;  // a branch on the loaded value: the clock below is read after it arrived
        t1 = clk();
        acc += v;
        if (threadIdx.x == 0) out[test] = t1 - t0;
    }
    g[threadIdx.x] = acc;
}
int main() {
    int* g; long long* out;
    const size_t bytes = (size_t)4 << 30;
    cudaMalloc(&g, bytes); cudaMalloc(&out, 64);
    const char* names[] = {"cold ld.cg", "prefetch.L2, 4000 cyc, ld.cg", "prefetch.L1, 4000 cyc, ld.ca", "ld.cg line, 4000 cyc, ld.cg", "ld.ca line, 4000 cyc, ld.ca", "store + prefetch.L2, ld.cg"};
    for (int rep = 0; rep < 3; ++rep) {
        // lanes 32 MB apart, tests 1 MB apart, reps 8 KB apart: every access goes to a line never touched before
        k<<<1, 32>>>(g + rep * 2048, out, (size_t)8 << 20, (size_t)1 << 18);
        long long c[8]; cudaMemcpy(c, out, 64, cudaMemcpyDeviceToHost);
        printf("rep %d\n", rep);
        for (int i = 0; i < 6; ++i) printf("  %-32s %lld cycles\n", names[i], c[i]);
    }
    printf("err=%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
}
