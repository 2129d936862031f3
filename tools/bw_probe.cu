// bw_probe.cu -- ground truth for streaming-read bandwidth on this GPU with several access styles.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bw_probe tools/bw_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int U>
__global__ void k_read_nc(const double2 *in, double *out, size_t n2) {
    double acc = 0.0;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    for (; i + (U - 1) * stride < n2; i += U * stride) {
        double2 v[U];
#pragma unroll
        for (int u = 0; u < U; ++u)
            asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(v[u].x), "=d"(v[u].y) : "l"(in + i + u * stride));
#pragma unroll
        for (int u = 0; u < U; ++u) acc += v[u].x + v[u].y;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
template <int U>
__global__ void k_read_plain(const double2 *in, double *out, size_t n2) {
    double acc = 0.0;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    for (; i + (U - 1) * stride < n2; i += U * stride) {
        double2 v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) v[u] = in[i + u * stride];
#pragma unroll
        for (int u = 0; u < U; ++u) acc += v[u].x + v[u].y;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
__global__ void k_copy(const double2 *in, double2 *o, size_t n2) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n2; i += stride) o[i] = in[i];
}
template <typename F>
static float timeit(F f, int reps = 5) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    f(); cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < reps; ++r) {
        cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms;
    }
    return best;
}
int main() {
    cudaDeviceProp pr; cudaGetDeviceProperties(&pr, 0);
    const int SM = pr.multiProcessorCount;
    size_t bytes = 4ull << 30; double2 *buf, *buf2; cudaMalloc(&buf, bytes); cudaMalloc(&buf2, bytes);
    cudaMemset(buf, 0, bytes); cudaMemset(buf2, 0, bytes);
    double *out; cudaMalloc(&out, sizeof(double) * SM * 64 * 1024);
    const size_t n2 = bytes / 16;
    printf("{\"gpu\": \"%s\", \"memclk_khz\": %d, \"l2_bytes\": %d,\n", pr.name, pr.memoryClockRate, pr.l2CacheSize);
    float ms;
    ms = timeit([&] { cudaMemcpyAsync(buf2, buf, bytes, cudaMemcpyDeviceToDevice); });
    printf(" \"memcpy_d2d_GBps(rd+wr)\": %.0f,\n", 2.0 * bytes / (ms * 1e-3) / 1e9);
    ms = timeit([&] { k_copy<<<SM * 16, 256>>>(buf, buf2, n2); });
    printf(" \"copy_kernel_GBps(rd+wr)\": %.0f,\n", 2.0 * bytes / (ms * 1e-3) / 1e9);
#define RUN(name, kern, grid, thr) ms = timeit([&] { kern<<<grid, thr>>>(buf, out, n2); }); printf(" \"%s\": %.0f,\n", name, bytes / (ms * 1e-3) / 1e9);
    RUN("read_nc_u1_g16x256", k_read_nc<1>, SM * 16, 256)
    RUN("read_nc_u4_g16x256", k_read_nc<4>, SM * 16, 256)
    RUN("read_nc_u8_g8x256", k_read_nc<8>, SM * 8, 256)
    RUN("read_nc_u4_g1x512", k_read_nc<4>, SM * 1, 512)
    RUN("read_nc_u8_g1x512", k_read_nc<8>, SM * 1, 512)
    RUN("read_nc_u4_g2x512", k_read_nc<4>, SM * 2, 512)
    RUN("read_plain_u1_g16x256", k_read_plain<1>, SM * 16, 256)
    RUN("read_plain_u4_g16x256", k_read_plain<4>, SM * 16, 256)
    RUN("read_plain_u8_g4x512", k_read_plain<8>, SM * 4, 512)
    // L2-resident working set (64 MB) re-read
    const size_t small = (64ull << 20) / 16;
    ms = timeit([&] { for (int r = 0; r < 8; ++r) k_read_nc<4><<<SM * 16, 256>>>(buf, out, small); });
    printf(" \"l2_resident_64MB_read_nc_u4\": %.0f}\n", 8.0 * small * 16 / (ms * 1e-3) / 1e9);
    return 0;
}
