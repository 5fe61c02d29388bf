// Does a DFMA with three distinct register source pairs issue slower than one with constant operands? (sm_100a)
#include <cstdio>
#include <cuda_runtime.h>
// MODE 0: a = fma(a, B, C) with B, C kernel constants.  MODE 1: a_i = fma(a_i, b_i, c_i), all per-thread registers.
// MODE 2: a_i = fma(a_i, b, c_i): one shared register multiplier.  MODE 3: a_i = a_i * b_i (DMUL, 2 reg sources)
// MODE 4: a_i = fma(b_i, c_i, a_i) rotating which slots are distinct.
template <int MODE>
__global__ void k(double* out, const double* in, int iters, double B, double C) {
    constexpr int ILP = 8;
    double a[ILP], b[ILP], c[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) { a[i] = in[threadIdx.x + i]; b[i] = in[threadIdx.x + 64 + i]; c[i] = in[threadIdx.x + 128 + i]; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int i = 0; i < ILP; ++i) {
                if (MODE == 0) a[i] = fma(a[i], B, C);
                if (MODE == 1) a[i] = fma(a[i], b[i], c[i]);
                if (MODE == 2) a[i] = fma(a[i], b[0], c[i]);
                if (MODE == 3) a[i] = a[i] * b[i];
                if (MODE == 4) a[i] = fma(b[i], c[(i + r) % ILP], a[i]);
            }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += a[i] + b[i] + c[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE> void run(double* d, double* in, const char* name) {
    int iters = 4000; cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<148 * 2, 256>>>(d, in, 10, 0.999, 1e-9);
    cudaEventRecord(e0); k<MODE><<<148 * 2, 256>>>(d, in, iters, 0.999, 1e-9); cudaEventRecord(e1);
    cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1);
    double inst = 148.0 * 2 * 256 / 32 * iters * 64;      // warp instructions
    double rate = inst / (ms * 1e-3) / (148 * 4);         // per SMSP per second
    printf("%-46s %.3f ms  -> %.3f G warp-instr/s/SMSP  (2-cycle issue at 1.93 GHz = 0.965)\n", name, ms, rate / 1e9);
}
int main() {
    double *d, *in; cudaMalloc(&d, 148 * 2 * 256 * 8); cudaMalloc(&in, 4096 * 8); cudaMemset(in, 0, 4096 * 8);
    run<0>(d, in, "DFMA a=fma(a,const,const)");
    run<1>(d, in, "DFMA a_i=fma(a_i,b_i,c_i) 3 distinct reg pairs");
    run<2>(d, in, "DFMA a_i=fma(a_i,b,c_i) shared multiplier");
    run<3>(d, in, "DMUL a_i=a_i*b_i 2 reg pairs");
    run<4>(d, in, "DFMA a_i=fma(b_i,c_j,a_i) accumulate form");
    return 0;
}
