// p2p_bw.cu -- achieved GPU-to-GPU bandwidth over NVLink for the re-shard pattern of the column-sharded LDE
// (tap-stark_b200/parallel.py): every GPU pushes one contiguous block to every other GPU at the same time.
// Compares copy-engine copies (cudaMemcpyPeerAsync on 1, 2, 4 and n-1 streams per GPU) with an SM copy kernel that
// stores straight into peer memory.  One process, all visible GPUs, one JSON line per variant.
#include <cstdint>
#include <cstdio>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("{\"error\": \"%s: %s\"}\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)

__global__ void copy_kernel(const uint4 *__restrict__ src, uint4 *__restrict__ dst, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) dst[i] = src[i];
}

int main(int argc, char **argv) {
    int G = 0;
    CK(cudaGetDeviceCount(&G));
    if (G < 2) { printf("{\"error\": \"needs >= 2 GPUs\"}\n"); return 0; }
    const size_t blk = (size_t)(argc > 1 ? atoi(argv[1]) : 64) << 20;  // bytes per (src, dst) pair
    std::vector<char *> send(G), recv(G);
    for (int d = 0; d < G; d++) {
        CK(cudaSetDevice(d));
        for (int p = 0; p < G; p++) if (p != d) { int ok = 0; cudaDeviceCanAccessPeer(&ok, d, p); if (ok) cudaDeviceEnablePeerAccess(p, 0); }
        cudaGetLastError();
        CK(cudaMalloc(&send[d], blk * G));
        CK(cudaMalloc(&recv[d], blk * G));
        CK(cudaMemset(send[d], d + 1, blk * G));
    }
    const int max_streams = 8;
    std::vector<std::vector<cudaStream_t>> st(G, std::vector<cudaStream_t>(max_streams));
    std::vector<cudaEvent_t> e0(G), e1(G);
    for (int d = 0; d < G; d++) {
        CK(cudaSetDevice(d));
        for (auto &s : st[d]) CK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
        CK(cudaEventCreate(&e0[d]));
        CK(cudaEventCreate(&e1[d]));
    }
    auto sync_all = [&]() { for (int d = 0; d < G; d++) { cudaSetDevice(d); cudaDeviceSynchronize(); } };
    // variants: ns = streams per GPU (copy engine), ns = 0: SM kernel with `blocks` CTAs per peer
    struct V { const char *name; int ns; int blocks; };
    const V vars[] = {{"ce_1_stream", 1, 0}, {"ce_2_streams", 2, 0}, {"ce_4_streams", 4, 0}, {"ce_stream_per_peer", 7, 0},
                      {"sm_kernel_8_ctas_per_peer", 0, 8}, {"sm_kernel_32_ctas_per_peer", 0, 32}};
    for (const V &v : vars) {
        float best = 1e30f;
        for (int rep = 0; rep < 4; rep++) {
            sync_all();
            for (int d = 0; d < G; d++) { cudaSetDevice(d); cudaEventRecord(e0[d], st[d][0]); for (int s = 1; s < max_streams; s++) cudaStreamWaitEvent(st[d][s], e0[d], 0); }
            for (int k = 1; k < G; k++)
                for (int d = 0; d < G; d++) {
                    const int p = (d + k) % G;
                    cudaSetDevice(d);
                    if (v.ns) {
                        cudaMemcpyPeerAsync(recv[p] + blk * d, p, send[d] + blk * p, d, blk, st[d][(k - 1) % v.ns]);
                    } else {
                        copy_kernel<<<v.blocks, 512, 0, st[d][(k - 1) % max_streams]>>>((const uint4 *)(send[d] + blk * p), (uint4 *)(recv[p] + blk * d), blk / 16);
                    }
                }
            float worst = 0;
            for (int d = 0; d < G; d++) {
                cudaSetDevice(d);
                for (int s = 1; s < max_streams; s++) { cudaEvent_t t; cudaEventCreateWithFlags(&t, cudaEventDisableTiming); cudaEventRecord(t, st[d][s]); cudaStreamWaitEvent(st[d][0], t, 0); cudaEventDestroy(t); }
                cudaEventRecord(e1[d], st[d][0]);
            }
            for (int d = 0; d < G; d++) { cudaSetDevice(d); cudaEventSynchronize(e1[d]); float ms; cudaEventElapsedTime(&ms, e0[d], e1[d]); if (ms > worst) worst = ms; }
            if (worst < best) best = worst;
        }
        const double out_bytes = (double)blk * (G - 1);
        printf("{\"variant\": \"%s\", \"gpus\": %d, \"block_MiB\": %zu, \"ms\": %.3f, \"egress_GBs_per_gpu\": %.1f}\n", v.name, G, blk >> 20, best,
               out_bytes / (best * 1e-3) / 1e9);
        fflush(stdout);
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) printf("{\"error\": \"%s\"}\n", cudaGetErrorString(e));
    return 0;
}
