// int_pipes.cu -- INT32 issue-rate microbenchmark for B200: the NTT butterflies
// and Blake3 are bound by integer issue, not HBM, so the roofline needs measured per-SM rates of
// IMAD / IMAD.HI / IMAD.WIDE (fma pipe) and IADD3 / LOP3 / SHF / VIMNMX (alu pipe), alone and mixed.
// Prints one JSON line per test: lane-ops per clock per SM and Gop/s for the chip.
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

constexpr int ITERS = 4096;
constexpr int CHAINS = 8;

#define KERNEL(name, BODY)                                                                  \
    __global__ void __launch_bounds__(256) name(uint32_t *out, uint32_t a0, uint32_t b0) {  \
        uint32_t x[CHAINS];                                                                 \
        uint32_t a = a0 + threadIdx.x, b = b0 | 1u;                                         \
        _Pragma("unroll") for (int c = 0; c < CHAINS; c++) x[c] = a + c * 977u;              \
        for (int it = 0; it < ITERS; it++) {                                                \
            _Pragma("unroll") for (int c = 0; c < CHAINS; c++) { BODY }                     \
        }                                                                                   \
        uint32_t s = 0;                                                                     \
        _Pragma("unroll") for (int c = 0; c < CHAINS; c++) s ^= x[c];                        \
        if (s == 0x12345u) out[blockIdx.x * blockDim.x + threadIdx.x] = s;                  \
    }

// 1 op per BODY
KERNEL(k_imad, asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[c]) : "r"(b), "r"(a));)
KERNEL(k_imad_hi, asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(x[c]) : "r"(b), "r"(a));)
KERNEL(k_iadd3, asm volatile("add.u32 %0, %0, %1;" : "+r"(x[c]) : "r"(b));)
KERNEL(k_lop3, asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[c]) : "r"(b), "r"(a));)
KERNEL(k_shf, asm volatile("shf.r.wrap.b32 %0, %0, %0, 7;" : "+r"(x[c]));)
KERNEL(k_min, asm volatile("min.u32 %0, %0, %1;" : "+r"(x[c]) : "r"(b)); b += 3;)
// 64-bit wide multiply-accumulate (2 regs): counted as 1 op
__global__ void __launch_bounds__(256) k_imad_wide(uint32_t *out, uint32_t a0, uint32_t b0) {
    unsigned long long x[CHAINS];
    uint32_t a = a0 + threadIdx.x, b = b0 | 1u;
#pragma unroll
    for (int c = 0; c < CHAINS; c++) x[c] = a + c * 977u;
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int c = 0; c < CHAINS; c++) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(x[c]) : "r"(a), "r"(b));
    }
    unsigned long long s = 0;
#pragma unroll
    for (int c = 0; c < CHAINS; c++) s ^= x[c];
    if (s == 0x12345ull) out[blockIdx.x * blockDim.x + threadIdx.x] = (uint32_t)s;
}
// mixes: 2 ops per BODY
KERNEL(k_mix_imad_iadd, asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[c]) : "r"(b), "r"(a));
       asm volatile("add.u32 %0, %0, %1;" : "+r"(x[c]) : "r"(b));)
KERNEL(k_mix_imad_lop, asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[c]) : "r"(b), "r"(a));
       asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[c]) : "r"(b), "r"(a));)
KERNEL(k_mix_lop_shf, asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[c]) : "r"(b), "r"(a));
       asm volatile("shf.r.wrap.b32 %0, %0, %0, 7;" : "+r"(x[c]));)
// 1 imad + 2 alu (the butterfly-like ratio)
KERNEL(k_mix_1imad_2alu, asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[c]) : "r"(b), "r"(a));
       asm volatile("add.u32 %0, %0, %1;" : "+r"(x[c]) : "r"(b));
       asm volatile("min.u32 %0, %0, %1;" : "+r"(x[c]) : "r"(a));)

// A full lazy DIF butterfly as the NTT issues it (C code, compiler-scheduled): counted as 1 butterfly
constexpr uint32_t P = 0x78000001u;
__device__ __forceinline__ uint32_t umin_(uint32_t a, uint32_t b) { return a < b ? a : b; }
__global__ void __launch_bounds__(256) k_butterfly(uint32_t *out, uint32_t a0, uint32_t b0) {
    uint32_t x[CHAINS], y[CHAINS];
    const uint32_t w = (b0 | 1u) % P, wp = (uint32_t)((((unsigned long long)w) << 32) / P);
#pragma unroll
    for (int c = 0; c < CHAINS; c++) { x[c] = (a0 + threadIdx.x + c * 977u) % P; y[c] = (x[c] * 3u + 1u) % P; }
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int c = 0; c < CHAINS; c++) {
            const uint32_t a = x[c], b = y[c];
            uint32_t s = a + b;
            x[c] = umin_(s, s - P);
            const uint32_t d = a - b + P;
            const uint32_t q = __umulhi(d, wp);
            const uint32_t r = d * w - q * P;
            y[c] = umin_(r, r - P);
        }
    }
    uint32_t s = 0;
#pragma unroll
    for (int c = 0; c < CHAINS; c++) s ^= x[c] ^ y[c];
    if (s == 0x12345u) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// The same butterfly with a Montgomery product built from full-rate IMAD.WIDE (t = d*wR; m = lo(t)*(-1/p);
// r = hi(t + m*p) in [0,2p)) instead of the Shoup product (IMAD.HI is half rate): 3 FMA-pipe slots of 2 cycles
// instead of 4 + 2 + 2.
__global__ void __launch_bounds__(256) k_butterfly_mont(uint32_t *out, uint32_t a0, uint32_t b0) {
    uint32_t x[CHAINS], y[CHAINS];
    const uint32_t w = (b0 | 1u) % P;
#pragma unroll
    for (int c = 0; c < CHAINS; c++) { x[c] = (a0 + threadIdx.x + c * 977u) % P; y[c] = (x[c] * 3u + 1u) % P; }
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int c = 0; c < CHAINS; c++) {
            const uint32_t a = x[c], b = y[c];
            uint32_t s = a + b;
            x[c] = umin_(s, s - P);
            const uint32_t d = a - b + P;
            unsigned long long t, u;
            asm("mul.wide.u32 %0, %1, %2;" : "=l"(t) : "r"(d), "r"(w));
            const uint32_t m = (uint32_t)t * 0x77ffffffu;
            asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(u) : "r"(m), "r"(P), "l"(t));
            const uint32_t r = (uint32_t)(u >> 32);
            y[c] = umin_(r, r - P);
        }
    }
    uint32_t s = 0;
#pragma unroll
    for (int c = 0; c < CHAINS; c++) s ^= x[c] ^ y[c];
    if (s == 0x12345u) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <class K>
void run(const char *name, K kern, double ops_per_body, int sms, uint32_t *out) {
    const int blocks = sms * 8, threads = 256;
    kern<<<blocks, threads>>>(out, 12345u, 6789u);
    cudaDeviceSynchronize();
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    float best = 1e30f;
    for (int rep = 0; rep < 5; rep++) {
        cudaEventRecord(a);
        kern<<<blocks, threads>>>(out, 12345u + rep, 6789u);
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms;
        cudaEventElapsedTime(&ms, a, b);
        if (ms < best) best = ms;
    }
    const double ops = (double)blocks * threads * ITERS * CHAINS * ops_per_body;
    int clk_khz = 0;
    cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    const double gops = ops / (best * 1e-3) / 1e9;
    printf("{\"test\": \"%s\", \"ms\": %.4f, \"Gops\": %.1f, \"lane_ops_per_clk_per_sm_at_max_clock\": %.2f}\n", name,
           best, gops, gops * 1e9 / ((double)clk_khz * 1e3) / sms);
    fflush(stdout);
}

int main() {
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, 0);
    int clk_khz = 0;
    cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    printf("{\"device\": \"%s\", \"sms\": %d, \"max_clock_mhz\": %d}\n", prop.name, prop.multiProcessorCount, clk_khz / 1000);
    uint32_t *out;
    cudaMalloc(&out, 148 * 8 * 256 * 4 * 2);
    const int sms = prop.multiProcessorCount;
    run("imad_lo", k_imad, 1, sms, out);
    run("imad_hi", k_imad_hi, 1, sms, out);
    run("imad_wide", k_imad_wide, 1, sms, out);
    run("iadd", k_iadd3, 1, sms, out);
    run("lop3", k_lop3, 1, sms, out);
    run("shf", k_shf, 1, sms, out);
    run("umin", k_min, 1, sms, out);
    run("mix_imad+iadd", k_mix_imad_iadd, 2, sms, out);
    run("mix_imad+lop3", k_mix_imad_lop, 2, sms, out);
    run("mix_lop3+shf", k_mix_lop_shf, 2, sms, out);
    run("mix_1imad+2alu", k_mix_1imad_2alu, 3, sms, out);
    run("dif_butterfly", k_butterfly, 1, sms, out);
    run("dif_butterfly_mont_wide", k_butterfly_mont, 1, sms, out);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("{\"error\": \"%s\"}\n", cudaGetErrorString(e)); return 1; }
    return 0;
}
