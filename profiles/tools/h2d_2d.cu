// h2d_2d: host->device throughput of the column-chunk copies the pipelined commit uses (cudaMemcpy2DAsync out of a
// pinned row-major 2^22 x 256 u32 matrix) as a function of the chunk width, next to a plain 1-D copy.
// build: nvcc -O2 -o h2d_2d h2d_2d.cu ; prints one JSON line per case.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("cuda error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)
int main() {
    const size_t rows = (size_t)1 << 22, w = 256;
    uint32_t *h = nullptr, *d = nullptr;
    CK(cudaMallocHost((void **)&h, rows * w * 4));
    CK(cudaMalloc((void **)&d, rows * w * 4));
    for (size_t i = 0; i < rows * w; i += 1024) h[i] = (uint32_t)i;
    cudaStream_t s;
    CK(cudaStreamCreate(&s));
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a));
    CK(cudaEventCreate(&b));
    for (int rep = 0; rep < 2; rep++) {
        CK(cudaEventRecord(a, s));
        CK(cudaMemcpyAsync(d, h, rows * w * 4, cudaMemcpyHostToDevice, s));
        CK(cudaEventRecord(b, s));
        CK(cudaEventSynchronize(b));
        float ms;
        CK(cudaEventElapsedTime(&ms, a, b));
        if (rep) printf("{\"case\": \"1d\", \"bytes\": %zu, \"ms\": %.3f, \"GBs\": %.2f}\n", rows * w * 4, ms, rows * w * 4 / ms / 1e6);
    }
    const int widths[] = {8, 16, 32, 64, 128, 256};
    for (int cw : widths) {
        for (int rep = 0; rep < 2; rep++) {
            CK(cudaEventRecord(a, s));
            for (size_t c0 = 0; c0 < w; c0 += cw)
                CK(cudaMemcpy2DAsync(d + c0 * rows, (size_t)cw * 4, h + c0, w * 4, (size_t)cw * 4, rows, cudaMemcpyHostToDevice, s));
            CK(cudaEventRecord(b, s));
            CK(cudaEventSynchronize(b));
            float ms;
            CK(cudaEventElapsedTime(&ms, a, b));
            if (rep) printf("{\"case\": \"2d\", \"chunk_cols\": %d, \"segment_bytes\": %d, \"ms\": %.3f, \"GBs\": %.2f}\n", cw, cw * 4, ms, rows * w * 4 / ms / 1e6);
        }
    }
    // the same narrow copies issued round-robin on several streams (do several copy engines fill the link?)
    cudaStream_t ss[4];
    for (int i = 0; i < 4; i++) CK(cudaStreamCreateWithFlags(&ss[i], cudaStreamNonBlocking));
    for (int ns : {2, 4}) {
        for (int cw : {8, 16, 32}) {
            for (int rep = 0; rep < 2; rep++) {
                CK(cudaDeviceSynchronize());
                CK(cudaEventRecord(a, s));
                for (int i = 0; i < ns; i++) CK(cudaStreamWaitEvent(ss[i], a, 0));
                int k = 0;
                for (size_t c0 = 0; c0 < w; c0 += cw, k++)
                    CK(cudaMemcpy2DAsync(d + c0 * rows, (size_t)cw * 4, h + c0, w * 4, (size_t)cw * 4, rows, cudaMemcpyHostToDevice, ss[k % ns]));
                for (int i = 0; i < ns; i++) {
                    CK(cudaEventRecord(b, ss[i]));
                    CK(cudaStreamWaitEvent(s, b, 0));
                }
                CK(cudaEventRecord(b, s));
                CK(cudaEventSynchronize(b));
                float ms;
                CK(cudaEventElapsedTime(&ms, a, b));
                if (rep) printf("{\"case\": \"2d\", \"streams\": %d, \"chunk_cols\": %d, \"segment_bytes\": %d, \"ms\": %.3f, \"GBs\": %.2f}\n", ns, cw, cw * 4, ms, rows * w * 4 / ms / 1e6);
            }
        }
    }
    return 0;
}
