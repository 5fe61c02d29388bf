// DFMA dependent-issue latency and throughput vs (warps per SMSP, independent chains per thread) on sm_100a.
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP>
__global__ void k(double* out, int iters, double b, double c) {
    double a[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) a[i] = threadIdx.x + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 16; ++r)
#pragma unroll
            for (int i = 0; i < ILP; ++i) a[i] = fma(a[i], b, c);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int ILP>
void run(int warps_per_sm, double* d) {
    int iters = 2000;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<ILP><<<148, warps_per_sm * 32>>>(d, 10, 0.999, 1e-9);
    cudaEventRecord(e0); k<ILP><<<148, warps_per_sm * 32>>>(d, iters, 0.999, 1e-9); cudaEventRecord(e1);
    cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1);
    double cyc = ms * 1e-3 * 1.965e9;                       // assumes boost clock
    double per_warp_instr = (double)iters * 16 * ILP;
    double warps_per_smsp = warps_per_sm / 4.0;
    printf("ILP %d warps/SMSP %.2f: %.2f cycles per DFMA per warp (chain latency if ILP=1), SMSP issue interval %.2f cyc\n",
           ILP, warps_per_smsp, cyc / per_warp_instr * 1.0, cyc / (per_warp_instr * warps_per_smsp));
}
int main() {
    double* d; cudaMalloc(&d, 148 * 1024 * 8);
    for (int w : {4, 8, 16, 32}) { run<1>(w, d); run<2>(w, d); run<4>(w, d); run<8>(w, d); }
    return 0;
}
