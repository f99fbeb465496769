// Integer-pipe throughput probe for sm_100a (what bounds the Goldilocks kernels): ops/clk/SM for
// mad.wide.u32 (IMAD.WIDE.U32), mad.lo.u32 (IMAD), add.u32/add.cc chains (IADD3), lop3, and a Goldilocks modmul.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/int_peak tools/int_peak.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../zk-light-client-implementation_b200/csrc/gl.cuh"

#define ITERS 4096
#define ILP 8
template <int OP> __global__ void k(uint64_t *out, uint32_t seed) {
    uint64_t a[ILP]; uint32_t b = seed | 1, c = threadIdx.x * 2654435761u + 12345u;
#pragma unroll
    for (int i = 0; i < ILP; i++) a[i] = (uint64_t)(c + i) * 0x9E3779B97F4A7C15ULL;
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) {
            if (OP == 0) { asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(a[i]) : "r"((uint32_t)a[i]), "r"(b)); }
            if (OP == 1) { uint32_t x = (uint32_t)a[i]; asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x) : "r"(b), "r"(c)); a[i] = x; }
            if (OP == 2) { uint32_t x = (uint32_t)a[i]; asm volatile("add.u32 %0, %0, %1;" : "+r"(x) : "r"(b)); a[i] = x; }
            if (OP == 3) { uint32_t x = (uint32_t)a[i]; asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x) : "r"(b), "r"(c)); a[i] = x; }
            if (OP == 4) { a[i] = gl_mul(a[i], a[(i + 1) % ILP] | 1); }
            if (OP == 5) { a[i] = gl_add(gl_canon(a[i]), 0x1234567812345678ULL); }
        }
    }
    uint64_t s = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) s ^= a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int OP> void run(const char *name, uint64_t *d, int sms, double clk_ghz) {
    int blocks = sms * 8, threads = 256;
    k<OP><<<blocks, threads>>>(d, 3); cudaDeviceSynchronize();
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e9;
    for (int r = 0; r < 5; r++) { cudaEventRecord(e0); k<OP><<<blocks, threads>>>(d, 3); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms; }
    double ops = (double)blocks * threads * ITERS * ILP;
    printf("{\"op\":\"%s\",\"ms\":%.4f,\"Gops\":%.1f,\"ops_per_clk_per_sm_at_%.2fGHz\":%.1f}\n", name, best, ops / best / 1e6, clk_ghz, ops / (best * 1e-3) / (clk_ghz * 1e9) / sms);
}
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount; double clk = p.clockRate / 1e6;
    printf("{\"gpu\":\"%s\",\"sms\":%d,\"max_clock_ghz\":%.3f}\n", p.name, sms, clk);
    uint64_t *d; cudaMalloc(&d, (size_t)sms * 8 * 256 * 8);
    run<0>("mad.wide.u32", d, sms, clk); run<1>("mad.lo.u32", d, sms, clk); run<2>("add.u32", d, sms, clk);
    run<3>("lop3", d, sms, clk); run<4>("gl_mul(canonical)", d, sms, clk); run<5>("gl_add", d, sms, clk);
    return 0;
}
