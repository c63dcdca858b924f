// Microbenchmark: latency / issue interval of mma.sync m8n8k4 f64 (DMMA) per warp as a function of the number of
// independent accumulator chains and of the warps per SM sub-partition, without and with the B operand coming
// from shared memory (one 128-bit load per two DMMAs, as in the MH kernel's quadratic form).
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double &d0, double &d1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

template <int NACC, bool SMEM>
__global__ void k(double *out, long long *clk, int iters, double a0, double b0)
{
    __shared__ double2 sb[32 * 64];
    for (int i = threadIdx.x; i < 32 * 64; i += blockDim.x) sb[i] = make_double2(b0 + i * 1e-9, b0 - i * 1e-9);
    __syncthreads();
    double c[NACC][2];
#pragma unroll
    for (int i = 0; i < NACC; ++i) c[i][0] = c[i][1] = threadIdx.x * 1e-9 + i;
    const double a = a0 + threadIdx.x * 1e-12;
    const int lane = threadIdx.x & 31;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) {
            double2 b = make_double2(b0, b0);
            if (SMEM) b = sb[((it * NACC + i) & 63) * 32 + lane];
            dmma(c[i][0], c[i][1], a, b.x);
        }
#pragma unroll
        for (int i = 0; i < NACC; ++i) {
            double2 b = make_double2(b0, b0);
            if (SMEM) b = sb[((it * NACC + i) & 63) * 32 + lane];
            dmma(c[i][0], c[i][1], a, b.y);
        }
    }
    const long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *clk = t1 - t0;
}

template <int NACC, bool SMEM>
void run(double *out, long long *clk, int threads)
{
    const int iters = 4000;
    k<NACC, SMEM><<<148, threads>>>(out, clk, iters, 1.0000001, 0.999999);
    cudaDeviceSynchronize();
    long long c;
    cudaMemcpy(&c, clk, sizeof c, cudaMemcpyDeviceToHost);
    const double per = (double)c / (iters * 2.0 * NACC);
    printf("smem %d  warps/SMSP %d  chains %d : %.1f clk per DMMA per warp, %.1f clk per DMMA per SMSP\n", (int)SMEM, threads / 128,
           NACC, per, per / (threads / 128));
}

int main()
{
    double *out;
    long long *clk;
    cudaMalloc(&out, sizeof(double) * 148 * 1024);
    cudaMalloc(&clk, 8);
    for (int threads : {128, 256, 512, 1024}) {
        run<1, false>(out, clk, threads);
        run<2, false>(out, clk, threads);
        run<4, false>(out, clk, threads);
        run<8, false>(out, clk, threads);
        run<4, true>(out, clk, threads);
        run<8, true>(out, clk, threads);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
