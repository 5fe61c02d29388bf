// nvl_store.cu -- how fast can SM-issued writes go over NVLink to ONE peer, by store shape?  (experiment)
//   a) st.global.f64, one 256-B row segment per warp instruction, rows strided by a large pitch (the fused gather today)
//   b) st.global.v2.f64 (512 B per warp instruction)
//   c) cp.async.bulk shared::cta -> global (TMA bulk store), 1 KiB per instruction, issued by one thread per CTA
//   d) same, 4 KiB per instruction
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o nvl_store nvl_store.cu ; run on a box with >= 2 GPUs
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

constexpr int ROWS = 98;

// grid: one CTA per 128 columns; every thread owns one column and writes ROWS values (row-major [ROWS][pitch])
__global__ void k_scalar(double *dst, long long pitch)
{
    const long long col = (long long)blockIdx.x * 128 + threadIdx.x;
    const double v = (double)col;
#pragma unroll 7
    for (int r = 0; r < ROWS; ++r) dst[r * pitch + col] = v + r;
}

// every thread owns two adjacent columns (v2 store): a 64-thread CTA covers 128 columns
__global__ void k_v2(double *dst, long long pitch)
{
    const long long col = ((long long)blockIdx.x * 64 + threadIdx.x) * 2;
    const double v = (double)col;
#pragma unroll 7
    for (int r = 0; r < ROWS; ++r) {
        double2 w = make_double2(v + r, v + r + 0.5);
        *reinterpret_cast<double2 *>(dst + r * pitch + col) = w;
    }
}

// stage [ROWS][COLS] in shared memory, then one thread issues a bulk store per row segment
template <int COLS>
__global__ void k_bulk(double *dst, long long pitch)
{
    extern __shared__ __align__(128) double sm[];
    const long long col0 = (long long)blockIdx.x * COLS;
    for (int i = threadIdx.x; i < ROWS * COLS; i += blockDim.x) sm[i] = (double)(col0 + i);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int r = 0; r < ROWS; ++r) {
            const uint32_t s = (uint32_t)__cvta_generic_to_shared(sm + r * COLS);
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst + r * pitch + col0), "r"(s),
                         "r"((uint32_t)(COLS * 8))
                         : "memory");
        }
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
}

int main()
{
    int n = 0;
    CK(cudaGetDeviceCount(&n));
    if (n < 2) { printf("needs 2 GPUs\n"); return 0; }
    const long long cols = 148LL * 128 * 96;   // 1,818,624 columns -> 98 rows x 8 B = 1.43 GB
    const long long pitch = cols;
    double *peer, *local;
    CK(cudaSetDevice(1));
    CK(cudaMalloc(&peer, ROWS * pitch * 8));
    CK(cudaSetDevice(0));
    CK(cudaMalloc(&local, ROWS * pitch * 8));
    CK(cudaDeviceEnablePeerAccess(1, 0));
    CK(cudaFuncSetAttribute(k_bulk<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, ROWS * 128 * 8));
    CK(cudaFuncSetAttribute(k_bulk<512>, cudaFuncAttributeMaxDynamicSharedMemorySize, ROWS * 512 * 8 / 4));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    const double gb = ROWS * pitch * 8 / 1e9;
    for (int where = 0; where < 2; ++where) {
        double *dst = where ? peer : local;
        printf("== destination: %s\n", where ? "PEER GPU over NVLink" : "local HBM");
        for (int variant = 0; variant < 4; ++variant) {
            float best = 1e9f;
            for (int rep = 0; rep < 4; ++rep) {
                CK(cudaEventRecord(e0));
                if (variant == 0) k_scalar<<<(unsigned)(cols / 128), 128>>>(dst, pitch);
                if (variant == 1) k_v2<<<(unsigned)(cols / 128), 64>>>(dst, pitch);
                if (variant == 2) k_bulk<128><<<(unsigned)(cols / 128), 128, ROWS * 128 * 8>>>(dst, pitch);
                if (variant == 3) k_bulk<128><<<(unsigned)(cols / 128), 256, ROWS * 128 * 8>>>(dst, pitch);
                CK(cudaEventRecord(e1));
                CK(cudaEventSynchronize(e1));
                CK(cudaGetLastError());
                float t;
                CK(cudaEventElapsedTime(&t, e0, e1));
                if (rep > 0 && t < best) best = t;
            }
            const char *names[4] = {"st.f64, 256 B per warp instr, rows strided", "st.v2.f64, 512 B per warp instr",
                                    "TMA bulk store 1 KiB per instr (128-thread CTAs)", "TMA bulk store 1 KiB per instr (256-thread CTAs)"};
            printf("  %-52s %7.3f ms  %7.1f GB/s\n", names[variant], best, gb / (best * 1e-3));
        }
    }
    return 0;
}
