// fp64_probe.cu -- measures what bounds the per-row math of the CGGibbs pass on this GPU:
// DFMA issue rate and the cost of the three families' row terms, against a plain streaming read.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/fp64_probe tools/fp64_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "../mcmcglm_b200/csrc/cgg_math.cuh"
using namespace cgg;

__global__ void k_dfma(double *out, int iters) {
    double a = threadIdx.x * 1e-9, b = 1.0000001, c = 1e-7, d = a + 1, e = a + 2, f = a + 3;
    for (int i = 0; i < iters; ++i) { a = fma(a, b, c); d = fma(d, b, c); e = fma(e, b, c); f = fma(f, b, c); }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a + d + e + f;
}
// dependent-issue latency of one warp: a single chain of DFMAs / of IMADs, clock64 around it
__global__ void k_latency(long long *out, int iters) {
    double a = threadIdx.x * 1e-9, b = 1.0000001, c = 1e-7;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) a = fma(a, b, c);
    long long t1 = clock64();
    int x = threadIdx.x, y = 3, z = 7;
    for (int i = 0; i < iters; ++i) x = x * y + z;
    long long t2 = clock64();
    if (threadIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t1; }
    if (a == 12345.0 && x == 7) out[2] = 1;
}
template <int FAM>
__global__ void k_term(double *out, int iters) {
    __shared__ double2 tab[MATH_TAB_N];
    load_l1p_table(tab);
    __syncthreads();
    const double yv = (threadIdx.x & 1), eta = 0.13 * (threadIdx.x & 31) - 2.0 + 1e-3 * (threadIdx.x >> 5), x = 1.0 + 1e-3 * blockIdx.x;
    double acc = 0.0;
    const RowPair<FAM> rp(make_double2(yv, 1.0 - yv), make_double2(eta, -eta), make_double2(x, x));
    for (int i = 0; i < iters; ++i) acc += rp.term(1e-4 * i, 1.0, tab);   // two row evaluations per call
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
__global__ void k_exp(double *out, int iters) {
    double eta = 0.001 * threadIdx.x - 0.1, acc = 0.0;
    for (int i = 0; i < iters; ++i) { acc += exp(eta + 1e-4 * i); acc += exp(-eta - 1e-4 * i); }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
__global__ void k_log1p(double *out, int iters) {
    double t = 0.001 * threadIdx.x + 0.01, acc = 0.0;
    for (int i = 0; i < iters; ++i) { acc += log1p(t + 1e-6 * i); acc += log1p(0.5 * t + 1e-6 * i); }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
__global__ void k_read(const double2 *in, double *out, size_t n2) {
    double acc = 0.0;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n2; i += (size_t)gridDim.x * blockDim.x) {
        double2 v = ld_stream2(reinterpret_cast<const double *>(in + i));
        acc += v.x + v.y;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
template <typename F>
static float timeit(F f) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    f(); cudaDeviceSynchronize();
    cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); return ms;
}
int main() {
    cudaDeviceProp pr; cudaGetDeviceProperties(&pr, 0);
    const int G = pr.multiProcessorCount * 8, T = 256, iters = 4096;
    double *out; cudaMalloc(&out, sizeof(double) * G * T);
    const double thr = (double)G * T;
    float ms = timeit([&] { k_dfma<<<G, T>>>(out, iters); });
    printf("{\"gpu\": \"%s\", \"sms\": %d,\n \"dfma_per_s\": %.4g,\n", pr.name, pr.multiProcessorCount, thr * iters * 4 / (ms * 1e-3));
    ms = timeit([&] { k_exp<<<G, T>>>(out, iters); });
    printf(" \"exp_per_s\": %.4g,\n", thr * iters * 2 / (ms * 1e-3));
    ms = timeit([&] { k_log1p<<<G, T>>>(out, iters); });
    printf(" \"log1p_per_s\": %.4g,\n", thr * iters * 2 / (ms * 1e-3));
    ms = timeit([&] { k_term<CGG_GAUSSIAN><<<G, T>>>(out, iters); });
    printf(" \"term_gaussian_per_s\": %.4g,\n", thr * iters * 2 / (ms * 1e-3));
    ms = timeit([&] { k_term<CGG_BINOMIAL><<<G, T>>>(out, iters); });
    printf(" \"term_binomial_per_s\": %.4g,\n", thr * iters * 2 / (ms * 1e-3));
    ms = timeit([&] { k_term<CGG_POISSON><<<G, T>>>(out, iters); });
    printf(" \"term_poisson_per_s\": %.4g,\n", thr * iters * 2 / (ms * 1e-3));
    // the sweep kernel's real occupancy: one 512-thread CTA per SM (4 warps per scheduler)
    const double thr1 = (double)pr.multiProcessorCount * 512;
    ms = timeit([&] { k_term<CGG_BINOMIAL><<<pr.multiProcessorCount, 512>>>(out, iters); });
    printf(" \"term_binomial_per_s_16warps\": %.4g,\n", thr1 * iters * 2 / (ms * 1e-3));
    ms = timeit([&] { k_term<CGG_POISSON><<<pr.multiProcessorCount, 512>>>(out, iters); });
    printf(" \"term_poisson_per_s_16warps\": %.4g,\n", thr1 * iters * 2 / (ms * 1e-3));
    ms = timeit([&] { k_term<CGG_BINOMIAL><<<pr.multiProcessorCount * 2, 512>>>(out, iters); });
    printf(" \"term_binomial_per_s_32warps\": %.4g,\n", thr1 * 2 * iters * 2 / (ms * 1e-3));
    {
        long long *lo; cudaMalloc(&lo, 64); cudaMemset(lo, 0, 64);
        k_latency<<<1, 32>>>(lo, 4096); cudaDeviceSynchronize();
        long long h[3]; cudaMemcpy(h, lo, sizeof h, cudaMemcpyDeviceToHost);
        printf(" \"dfma_dependent_latency_cycles\": %.2f,\n \"imad_dependent_latency_cycles\": %.2f,\n", h[0] / 4096.0, h[1] / 4096.0);
    }
    size_t bytes = 4ull << 30; double2 *buf; cudaMalloc(&buf, bytes); cudaMemset(buf, 0, bytes);
    ms = timeit([&] { k_read<<<pr.multiProcessorCount * 16, T>>>(buf, out, bytes / 16); });
    printf(" \"stream_read_GBps\": %.4g}\n", bytes / (ms * 1e-3) / 1e9);
    return 0;
}
