// Microbenchmark: fp64 DFMA dependent-issue latency and pipe throughput on B200 (sm_100a).
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP>
__global__ void k(double *out, long long *cyc, int iters, double a, double b)
{
    double x[ILP];
#pragma unroll
    for (int i = 0; i < ILP; i++) x[i] = threadIdx.x * 1e-9 + i;
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) x[i] = fma(x[i], a, b);
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

__global__ void kmufu(double *out, long long *cyc, int iters)
{
    double x = 1.0 + threadIdx.x;
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
        double y;
        asm volatile("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
        x = y + 1.5;
    }
    long long t1 = clock64();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) *cyc = t1 - t0;
}

template <int ILP>
void run(int warps, int blocks)
{
    double *out; long long *cyc, h;
    cudaMalloc(&out, sizeof(double) * warps * 32 * blocks);
    cudaMalloc(&cyc, 8);
    const int iters = 4096;
    k<ILP><<<blocks, warps * 32>>>(out, cyc, iters, 0.999999, 1e-7);
    cudaDeviceSynchronize();
    k<ILP><<<blocks, warps * 32>>>(out, cyc, iters, 0.999999, 1e-7);
    cudaDeviceSynchronize();
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("ILP=%d warps/block=%2d blocks=%d: %.2f cycles per DFMA-round, %.2f cycles per warp-DFMA per SM\n", ILP,
           warps, blocks, (double)h / iters, (double)h / iters / (ILP * warps));
    cudaFree(out); cudaFree(cyc);
}

int main()
{
    run<1>(1, 1); run<2>(1, 1); run<4>(1, 1); run<8>(1, 1);
    run<1>(4, 1); run<4>(4, 1); run<1>(16, 1); run<4>(16, 1); run<1>(32, 1); run<2>(32, 1); run<4>(32,1);
    double *out; long long *cyc, h;
    cudaMalloc(&out, 8 * 32); cudaMalloc(&cyc, 8);
    kmufu<<<1, 32>>>(out, cyc, 4096); cudaDeviceSynchronize();
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("rsqrt.approx.f64 + DADD dependent: %.2f cycles\n", (double)h / 4096);
    return 0;
}
