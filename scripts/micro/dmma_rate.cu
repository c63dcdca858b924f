// Microbenchmark: fp64 throughput of DFMA vs mma.sync m8n8k4 (DMMA) on one GPU.
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double &d0, double &d1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

template <int NACC>
__global__ void k_dmma(double *out, int iters, double a0, double b0)
{
    double c[NACC][2];
#pragma unroll
    for (int i = 0; i < NACC; ++i) c[i][0] = c[i][1] = threadIdx.x * 1e-9 + i;
    double a = a0 + threadIdx.x * 1e-12, b = b0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) dmma(c[i][0], c[i][1], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NACC>
__global__ void k_dfma(double *out, int iters, double a0, double b0)
{
    double c[NACC];
#pragma unroll
    for (int i = 0; i < NACC; ++i) c[i] = threadIdx.x * 1e-9 + i;
    double a = a0 + threadIdx.x * 1e-12;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) c[i] = fma(c[i], a, b0);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += c[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main()
{
    double *out;
    cudaMalloc(&out, sizeof(double) * 148 * 8 * 1024);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 20000;
    for (int threads : {128, 256, 512, 1024}) {
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(e0);
            k_dmma<8><<<148, threads>>>(out, iters, 1.0000001, 0.999999);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            double fma = 148.0 * (threads / 32) * iters * 8 * 256.0;
            if (rep) printf("DMMA threads/SM %4d: %.2f TFLOP/s (%.3f ms)\n", threads, 2 * fma / ms / 1e9, ms);
            cudaEventRecord(e0);
            k_dfma<8><<<148, threads>>>(out, iters * 8, 1.0000001, 0.999999);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            cudaEventElapsedTime(&ms, e0, e1);
            fma = 148.0 * threads * (double)iters * 8 * 8;
            if (rep) printf("DFMA threads/SM %4d: %.2f TFLOP/s (%.3f ms)\n", threads, 2 * fma / ms / 1e9, ms);
        }
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
