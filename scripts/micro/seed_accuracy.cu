// Largest relative error of the MUFU seeds (rcp.approx.ftz.f64 / rsqrt.approx.ftz.f64) and of the one-step third-order
// refinements fast_rcp / fast_rsqrt build on them (csrc/discretize_kernel.cuh), over 2^24 inputs spread over 1e-6 .. 1e6.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I mpconstellation_b200/csrc -o scripts/micro/seed_accuracy scripts/micro/seed_accuracy.cu
#include <cstdio>
#include <cstdint>
#include <cmath>
#include <cuda_runtime.h>
#include "discretize_kernel.cuh"

__global__ void probe(double *worst)
{
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    // a = 10^(-6 + 12 i / 2^24) times a mantissa jitter
    const double a = exp10(-6.0 + 12.0 * (double)i / 16777216.0) * (1.0 + 1e-3 * (double)(i % 977) / 977.0);
    double y0, y1;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y0) : "d"(a));
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y1) : "d"(a));
    const double r = 1.0 / a, s = 1.0 / sqrt(a);
    const double e[4] = {fabs(y0 - r) / r, fabs(y1 - s) / s, fabs(mpc::fast_rcp(a) - r) / r, fabs(mpc::fast_rsqrt(a) - s) / s};
    for (int k = 0; k < 4; ++k) {
        // atomic max on the bit pattern (non-negative doubles order like integers)
        atomicMax((unsigned long long *)(worst + k), (unsigned long long)__double_as_longlong(e[k]));
    }
}

int main()
{
    double *d, h[4] = {0, 0, 0, 0};
    cudaMalloc(&d, sizeof h);
    cudaMemcpy(d, h, sizeof h, cudaMemcpyHostToDevice);
    probe<<<16777216 / 256, 256>>>(d);
    cudaMemcpy(h, d, sizeof h, cudaMemcpyDeviceToHost);
    printf("seed rcp %.3e (2^%.1f)  seed rsqrt %.3e (2^%.1f)  fast_rcp %.3e  fast_rsqrt %.3e  (double epsilon 2.2e-16)\n", h[0], log2(h[0]), h[1],
           log2(h[1]), h[2], h[3]);
    return cudaGetLastError() != cudaSuccess;
}
