// Microbenchmark: do SHFL and shared-memory accesses share one issue / data path on sm_100a?
// Prints warp-instructions per clock per SM for SHFL.32 alone, LDS.128 alone (conflict-free), both interleaved, and with DMMA.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/mio_probe tools/probes/mio_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
template <int MODE>
__global__ void __launch_bounds__(512) probe(double* out, long long* cyc, int iters) {
    __shared__ __align__(16) double sm[5120];
    const int tid = threadIdx.x, lane = tid & 31;
    for (int i = tid; i < 4096; i += blockDim.x) sm[i] = (double)i;
    __syncthreads();
    int a0 = tid, a1 = tid + 1, a2 = tid + 2, a3 = tid + 3;
    double2 v0 = {0, 0}, v1 = {0, 0}, v2 = {0, 0}, v3 = {0, 0};
    double d0 = 0, d1 = 0, d2 = 0, d3 = 0, e0 = 0, e1 = 0, e2 = 0, e3 = 0;
    const double2* p = reinterpret_cast<const double2*>(sm) + ((tid >> 5) & 3) * 128 + lane;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        if (MODE & 1) {
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                a0 = __shfl_xor_sync(0xffffffffu, a0, 9); a1 = __shfl_xor_sync(0xffffffffu, a1, 18);
                a2 = __shfl_xor_sync(0xffffffffu, a2, 4); a3 = __shfl_xor_sync(0xffffffffu, a3, 1);
            }
        }
        if (MODE & 2) {
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                double2 x;
                asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(x.x), "=d"(x.y) : "r"((unsigned)__cvta_generic_to_shared(p + 32 * r + ((it * 7) & 63) * 16)));
                if (r == 0) { v0.x += x.x; v0.y += x.y; } else if (r == 1) { v1.x += x.x; v1.y += x.y; } else if (r == 2) { v2.x += x.x; v2.y += x.y; } else { v3.x += x.x; v3.y += x.y; }
            }
        }
        if (MODE & 4) {
#pragma unroll
            for (int r = 0; r < 2; ++r) { dmma(d0, d1, 1.0, 1.0); dmma(d2, d3, 1.0, 1.0); dmma(e0, e1, 1.0, 1.0); dmma(e2, e3, 1.0, 1.0); }
        }
        if (MODE & 8) {   // STS.64
#pragma unroll
            for (int r = 0; r < 4; ++r)
                asm volatile("st.shared.f64 [%0], %1;" :: "r"((unsigned)__cvta_generic_to_shared(sm + (tid >> 5) * 128 + lane + 32 * (r & 1))), "d"(d0) : "memory");
        }
    }
    const long long t1 = clock64();
    __syncthreads(); const long long t2 = clock64(); if (tid == 0) cyc[blockIdx.x] = t2 - t0;
    out[blockIdx.x * blockDim.x + tid] = a0 + a1 + a2 + a3 + v0.x + v0.y + v1.x + v1.y + v2.x + v2.y + v3.x + v3.y + d0 + d1 + d2 + d3 + e0 + e1 + e2 + e3;
}
// dependent-chain latencies
__global__ void latency(double* out, long long* cyc) {
    __shared__ __align__(16) double sm[64];
    const int lane = threadIdx.x;
    sm[lane] = 0.0; sm[lane + 32] = 0.0;
    __syncwarp();
    double d0 = 0, d1 = 0, x = 1.0;
    int idx = lane;
    long long t0 = clock64();
    for (int i = 0; i < 256; ++i) dmma(d0, d1, 1.0, x);
    long long t1 = clock64();
    for (int i = 0; i < 256; ++i) x = fma(x, 1.0000001, d0);
    long long t2 = clock64();
    for (int i = 0; i < 256; ++i) idx = __shfl_xor_sync(0xffffffffu, idx, 1);
    long long t3 = clock64();
    unsigned a = (unsigned)__cvta_generic_to_shared(sm) + lane * 8;
    for (int i = 0; i < 256; ++i) { double v; asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a)); a += (unsigned)(long long)v; }
    long long t4 = clock64();
    for (int i = 0; i < 256; ++i) { dmma(d0, d1, x, 1.0); x = d0; }   // accumulate -> A operand dependency
    long long t5 = clock64();
    if (lane == 0) { cyc[0] = t1 - t0; cyc[1] = t2 - t1; cyc[2] = t3 - t2; cyc[3] = t4 - t3; cyc[4] = t5 - t4; }
    out[lane] = d0 + d1 + x + idx + a;
}
int main() {
    double* out; long long* cyc;
    cudaMalloc(&out, 148 * 512 * 8 * 2); cudaMalloc(&cyc, 1024 * 8);
    const int iters = 4096;
    const char* names[16] = {"", "SHFL", "LDS128", "SHFL+LDS128", "DMMA", "SHFL+DMMA", "LDS128+DMMA", "SHFL+LDS128+DMMA", "STS64", "SHFL+STS64", "LDS128+STS64", "", "DMMA+STS64", "", "", ""};
    for (int warps : {4, 8, 16}) {
#define RUN(M) { probe<M><<<148, warps * 32>>>(out, cyc, iters); cudaDeviceSynchronize(); probe<M><<<148, warps * 32>>>(out, cyc, iters); cudaDeviceSynchronize(); \
        long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost); double c = 0; for (int i = 0; i < 148; ++i) c += h[i]; c /= 148; \
        const double nsh = (M & 1) ? 16.0 : 0, nld = (M & 2) ? 4.0 : 0, ndm = (M & 4) ? 8.0 : 0, nst = (M & 8) ? 4.0 : 0; \
        printf("warps/SM %2d  %-18s cycles/iter/SM %8.2f | per clk per SM: SHFL %.3f LDS128 %.3f (wavefronts %.3f) DMMA %.3f STS64 %.3f\n", warps, names[M], c / iters, \
               nsh * warps * iters / c, nld * warps * iters / c, 4 * nld * warps * iters / c, ndm * warps * iters / c, nst * warps * iters / c); }
        RUN(1) RUN(2) RUN(3) RUN(4) RUN(5) RUN(6) RUN(7) RUN(8) RUN(10)
    }
    latency<<<1, 32>>>(out, cyc); cudaDeviceSynchronize();
    long long h[5]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    printf("dependent latency (cycles): DMMA accumulate chain %.1f, DFMA %.1f, SHFL %.1f, LDS.64 pointer chase (incl. cvt/add) %.1f, DMMA->A operand chain %.1f\n",
           h[0] / 256.0, h[1] / 256.0, h[2] / 256.0, h[3] / 256.0, h[4] / 256.0);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
