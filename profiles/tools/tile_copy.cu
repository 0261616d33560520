// tile_copy.cu -- memory-system ceiling for the NTT tile access pattern: every CTA copies a tile of L rows x
// (K*4) bytes out of a row-major matrix with 1 KiB rows (256 columns), in place (read + write the same
// addresses), adjacent CTAs taking adjacent column slices -- exactly what ntt_pass does minus the arithmetic.
// Prints GB/s (read+write) per tile shape; (2048, 8) is the D = 11 shape.
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint4 ld128(const uint4 *p, int hint) {
    uint4 v;
    if (hint) asm volatile("ld.global.L2::128B.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    else v = *p;
    return v;
}
// variant: one CTA walks GROUP adjacent column slices of the same rows one after the other (temporal locality in
// L2 for the neighbouring 32-byte sectors of each 128-byte line), optionally with the L2::128B fetch hint
template <int LOGV, int GROUP>
__global__ void __launch_bounds__(256) copy_group_kernel(uint4 *data, int log_rows_per_tile, uint32_t width_v,
                                                         uint32_t n_slices, int lo_bits, int hint) {
    extern __shared__ uint4 sm[];
    const uint32_t n_groups = n_slices / GROUP;
    const uint32_t cg = blockIdx.x % n_groups, tile = blockIdx.x / n_groups;
    const uint32_t lo = tile & ((1u << lo_bits) - 1), hi = tile >> lo_bits;
    const size_t row_base = ((size_t)hi << (lo_bits + log_rows_per_tile)) + lo, row_stride = (size_t)1 << lo_bits;
    const int total = (1 << log_rows_per_tile) << LOGV;
    for (int s = 0; s < GROUP; s++) {
        const uint32_t cs = cg * GROUP + s;
        for (int it = threadIdx.x; it < total; it += 256) {
            const uint32_t v = it & ((1 << LOGV) - 1), q = it >> LOGV;
            sm[it] = ld128(data + (row_base + (size_t)q * row_stride) * width_v + (cs << LOGV) + v, hint);
        }
        __syncthreads();
        for (int it = threadIdx.x; it < total; it += 256) {
            const uint32_t v = it & ((1 << LOGV) - 1), q = it >> LOGV;
            uint4 x = sm[it ^ 1];
            x.x += 1;
            data[(row_base + (size_t)q * row_stride) * width_v + (cs << LOGV) + v] = x;
        }
        __syncthreads();
    }
}

template <int LOGV>  // uint4 vectors per row segment = 2^LOGV  (K = 4 << LOGV columns)
__global__ void __launch_bounds__(256) copy_kernel(uint4 *data, int log_rows_per_tile, uint32_t width_v, uint32_t n_slices,
                                                   int lo_bits) {
    extern __shared__ uint4 sm[];
    const uint32_t cs = blockIdx.x % n_slices, tile = blockIdx.x / n_slices;
    const uint32_t lo = tile & ((1u << lo_bits) - 1), hi = tile >> lo_bits;
    const size_t row_base = ((size_t)hi << (lo_bits + log_rows_per_tile)) + lo, row_stride = (size_t)1 << lo_bits;
    const int total = (1 << log_rows_per_tile) << LOGV;  // 4096 uint4
    for (int it = threadIdx.x; it < total; it += 256) {
        const uint32_t v = it & ((1 << LOGV) - 1), q = it >> LOGV;
        sm[it] = data[(row_base + (size_t)q * row_stride) * width_v + (cs << LOGV) + v];
    }
    __syncthreads();
    for (int it = threadIdx.x; it < total; it += 256) {
        const uint32_t v = it & ((1 << LOGV) - 1), q = it >> LOGV;
        uint4 x = sm[it ^ 1];
        x.x += 1;
        data[(row_base + (size_t)q * row_stride) * width_v + (cs << LOGV) + v] = x;
    }
}

template <int LOGV>
void run(uint4 *d, size_t rows, uint32_t width, int lo_bits_mode) {
    const int log_rows_per_tile = 12 - LOGV;  // 4096 uint4 per tile
    const uint32_t width_v = width / 4, n_slices = width_v >> LOGV;
    const size_t tiles = rows >> log_rows_per_tile;
    const int lo_bits = lo_bits_mode ? 11 : 0;
    cudaFuncSetAttribute(copy_kernel<LOGV>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    float best = 1e30f;
    for (int rep = 0; rep < 4; rep++) {
        cudaEventRecord(a);
        copy_kernel<LOGV><<<(unsigned)(tiles * n_slices), 256, 65536>>>(d, log_rows_per_tile, width_v, n_slices, lo_bits);
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms;
        cudaEventElapsedTime(&ms, a, b);
        if (rep > 0 && ms < best) best = ms;
    }
    const double bytes = 2.0 * rows * width * 4;
    printf("{\"tile_rows\": %d, \"segment_bytes\": %d, \"row_stride_rows\": %d, \"ms\": %.3f, \"GBs\": %.1f}\n",
           1 << log_rows_per_tile, 16 << LOGV, 1 << lo_bits, best, bytes / best / 1e6);
    fflush(stdout);
}

template <int GROUP>
void run_group(uint4 *d, size_t rows, uint32_t width, int lo_bits_mode, int hint) {
    const int LOGV = 1, log_rows_per_tile = 11;
    const uint32_t width_v = width / 4, n_slices = width_v >> LOGV;
    const size_t tiles = rows >> log_rows_per_tile;
    const int lo_bits = lo_bits_mode ? 11 : 0;
    cudaFuncSetAttribute(copy_group_kernel<LOGV, GROUP>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    float best = 1e30f;
    for (int rep = 0; rep < 4; rep++) {
        cudaEventRecord(a);
        copy_group_kernel<LOGV, GROUP><<<(unsigned)(tiles * n_slices / GROUP), 256, 65536>>>(d, log_rows_per_tile, width_v, n_slices, lo_bits, hint);
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms;
        cudaEventElapsedTime(&ms, a, b);
        if (rep > 0 && ms < best) best = ms;
    }
    const double bytes = 2.0 * rows * width * 4;
    printf("{\"variant\": \"group%d hint%d\", \"tile_rows\": 2048, \"segment_bytes\": 32, \"row_stride_rows\": %d, \"ms\": %.3f, \"GBs\": %.1f}\n",
           GROUP, hint, 1 << lo_bits, best, bytes / best / 1e6);
    fflush(stdout);
}

int main() {
    const size_t rows = (size_t)1 << 24;
    const uint32_t width = 256;
    uint4 *d;
    if (cudaMalloc(&d, rows * width * 4) != cudaSuccess) { printf("alloc failed\n"); return 1; }
    cudaMemset(d, 1, rows * width * 4);
    for (int mode = 0; mode < 2; mode++) {
        run<1>(d, rows, width, mode);  // 2048 rows x 32 B   (D = 11)
        run<2>(d, rows, width, mode);  // 1024 rows x 64 B   (D = 10)
        run<3>(d, rows, width, mode);  //  512 rows x 128 B  (D = 9)
        run<4>(d, rows, width, mode);  //  256 rows x 256 B
        run<6>(d, rows, width, mode);  //   64 rows x 1 KiB (full rows)
    }
    // "column-slice-major" (blocked) intermediate layout: one 8-column slice stored as its own [rows][8] array, so
    // the strided digit steps 64 KiB (not 2 MiB) between positions and contiguous tiles are one 64 KiB run
    printf("{\"note\": \"blocked layout: width 8\"}\n");
    run<1>(d, rows * 32, 8, 0);
    run<1>(d, rows * 32, 8, 1);
    for (int mode = 0; mode < 2; mode++)
        for (int hint = 0; hint < 1; hint++) {
            run_group<1>(d, rows, width, mode, hint);
            run_group<4>(d, rows, width, mode, hint);
        }
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("{\"error\": \"%s\"}\n", cudaGetErrorString(e)); return 1; }
    return 0;
}
