// Microbenchmark: fp64 FMA peak, DMMA (mma.sync m8n8k4 f64) peak, fp64 exp() throughput on this GPU.
#include <cstdio>
#include <cuda_runtime.h>
#include <mma.h>

__global__ void dfma_kernel(double *out, int iters)
{
    double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double b = 1.0000001, c = 1e-9;
    for (int i = 0; i < iters; ++i) {
        a0 = fma(a0, b, c); a1 = fma(a1, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c);
        a4 = fma(a4, b, c); a5 = fma(a5, b, c); a6 = fma(a6, b, c); a7 = fma(a7, b, c);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

__global__ void dmma_kernel(double *out, int iters)
{
    double a = threadIdx.x * 1e-3, b = 1.0 + threadIdx.x * 1e-6;
    double c0[2] = {0, 0}, c1[2] = {0, 0}, c2[2] = {0, 0}, c3[2] = {0, 0};
    for (int i = 0; i < iters; ++i) {
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0[0]), "+d"(c0[1]) : "d"(a), "d"(b));
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c1[0]), "+d"(c1[1]) : "d"(a), "d"(b));
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c2[0]), "+d"(c2[1]) : "d"(a), "d"(b));
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c3[0]), "+d"(c3[1]) : "d"(a), "d"(b));
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = c0[0] + c0[1] + c1[0] + c1[1] + c2[0] + c2[1] + c3[0] + c3[1];
}

__global__ void dexp_kernel(double *out, int iters)
{
    double x = -1.0 - threadIdx.x * 1e-3, s = 0;
    for (int i = 0; i < iters; ++i) { s += exp(x); x -= 1e-6; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void ffma_kernel(float *out, int iters)
{
    float a0 = threadIdx.x * 1e-9f, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const float b = 1.0000001f, c = 1e-9f;
    for (int i = 0; i < iters; ++i) {
        a0 = fmaf(a0, b, c); a1 = fmaf(a1, b, c); a2 = fmaf(a2, b, c); a3 = fmaf(a3, b, c);
        a4 = fmaf(a4, b, c); a5 = fmaf(a5, b, c); a6 = fmaf(a6, b, c); a7 = fmaf(a7, b, c);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

template <typename F> float time_ms(F f)
{
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); cudaDeviceSynchronize();
    cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}

int main()
{
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const int blocks = p.multiProcessorCount * 8, threads = 256;
    double *out; cudaMalloc(&out, sizeof(double) * blocks * threads);
    const int iters = 20000;
    float ms = time_ms([&] { dfma_kernel<<<blocks, threads>>>(out, iters); });
    printf("DFMA : %.2f TFLOP/s (%.1f FMA/clk/SM at 1.9GHz)\n", 2.0 * 8 * iters * blocks * threads / ms / 1e9, 8.0 * iters * blocks * threads / (ms * 1e-3) / p.multiProcessorCount / 1.9e9);
    ms = time_ms([&] { dmma_kernel<<<blocks, threads>>>(out, iters); });
    printf("DMMA : %.2f TFLOP/s\n", 2.0 * 4 * 256 * iters * (double)blocks * threads / 32 / ms / 1e9);
    ms = time_ms([&] { dexp_kernel<<<blocks, threads>>>(out, iters / 10); });
    printf("DEXP : %.2f Gexp/s\n", (double)(iters / 10) * blocks * threads / ms / 1e6);
    ms = time_ms([&] { ffma_kernel<<<blocks, threads>>>((float *)out, iters); });
    printf("FFMA : %.2f TFLOP/s\n", 2.0 * 8 * iters * blocks * threads / ms / 1e9);
    return 0;
}
