// FP64 throughput microbenchmarks: the roofline denominator for the evaluator kernels.
// MEASURED_PEAKS.json carries no FP64 figure, so bench.py measures one with these
// (register-resident DFMA chains / DMMA m8n8k4 chains, all SMs, CUDA-event timed).
#include "eml_common.cuh"

namespace eml {

__global__ void __launch_bounds__(256) dfma_peak_kernel(double* out, int iters, double a, double b) {
    double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
            x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}

__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(d0), "+d"(d1)
                 : "d"(a), "d"(b));
}

__global__ void __launch_bounds__(256) dmma_peak_kernel(double* out, int iters, double a, double b) {
    double c[8][2];
#pragma unroll
    for (int u = 0; u < 8; ++u) { c[u][0] = threadIdx.x + u; c[u][1] = u; }
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) dmma884(c[u][0], c[u][1], a, b);
    }
    double s = 0;
#pragma unroll
    for (int u = 0; u < 8; ++u) s += c[u][0] + c[u][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// mixed: every warp alternates independent DMMA and DFMA chains -- tells whether the two share one pipe
__global__ void __launch_bounds__(256) mixed_peak_kernel(double* out, int iters, double a, double b) {
    double c[4][2];
    double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
#pragma unroll
    for (int u = 0; u < 4; ++u) { c[u][0] = threadIdx.x + u; c[u][1] = u; }
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            dmma884(c[u][0], c[u][1], a, b);
            x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
            x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
        }
    }
    double s = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
#pragma unroll
    for (int u = 0; u < 4; ++u) s += c[u][0] + c[u][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

cudaError_t measure_fp64(int use_mma, double seconds, double* flops_per_s) {
    int dev = 0, sms = 0;
    cudaError_t e;
    if ((e = cudaGetDevice(&dev)) != cudaSuccess) return e;
    if ((e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) return e;
    const int blocks = sms * 8, threads = 256;
    double* out = nullptr;
    if ((e = cudaMalloc((void**)&out, sizeof(double) * blocks * threads)) != cudaSuccess) return e;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 4096;
    // flops per launch
    // mixed (use_mma == 2): per iteration and warp 4 DMMAs (512 flops each) + 32 warp-wide DFMAs (64 flops each)
    const double fl = use_mma == 2 ? (double)blocks * (threads / 32) * iters * (4.0 * 512.0 + 32.0 * 64.0)
                      : use_mma    ? (double)blocks * (threads / 32) * iters * 8.0 * (2.0 * 8 * 8 * 4)
                                   : (double)blocks * threads * iters * 64.0 * 2.0;
    double best = 0.0, elapsed = 0.0;
    int reps = 0;
    while ((elapsed < seconds * 1e3 || reps < 3) && reps < 10000) {
        cudaEventRecord(e0);
        if (use_mma == 2) mixed_peak_kernel<<<blocks, threads>>>(out, iters, 1.0000001, 1e-9);
        else if (use_mma) dmma_peak_kernel<<<blocks, threads>>>(out, iters, 1.0000001, 1e-9);
        else dfma_peak_kernel<<<blocks, threads>>>(out, iters, 1.0000001, 1e-9);
        cudaEventRecord(e1);
        if ((e = cudaEventSynchronize(e1)) != cudaSuccess) break;
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        elapsed += ms;
        if (reps > 0) best = fmax(best, fl / (ms * 1e-3));
        ++reps;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(out);
    if (e != cudaSuccess) return e;
    *flops_per_s = best;
    return cudaGetLastError();
}

}  // namespace eml
