// TEST INFRASTRUCTURE ONLY (tests/hostk) -- compiles the CUDA kernel sources of mpconstellation_b200/csrc for the HOST so
// that `pytest -m "not gpu"` can check the arithmetic, the indexing and the status logic of the very code that runs on
// the B200 against the oracles without a GPU.  Every kernel here is one independent thread per work unit with private
// shared-memory slots ([slot][thread]), so running the threads one after the other is a faithful emulation.
// Nothing under mpconstellation_b200/ may include, load or call this: the product has no CPU path
// (tests/test_host_logic.py::test_product_package_never_imports_the_oracle also scans for it).
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#include <condition_variable>
#include <mutex>

#define __device__
#define __global__
#define __host__
#define __forceinline__ inline
#define __noinline__ __attribute__((noinline))
#define __launch_bounds__(...)
#define __maxnreg__(...)
#define __shared__
#define __constant__
#define __grid_constant__
#define CUDART_INF (__builtin_inf())

struct hostk_dim3 {
    unsigned x, y, z;
};
static thread_local hostk_dim3 threadIdx = {0, 0, 0}, blockIdx = {0, 0, 0}, blockDim = {1, 1, 1}, gridDim = {1, 1, 1};

static inline int min(int a, int b) { return a < b ? a : b; }
static inline int max(int a, int b) { return a > b ? a : b; }
static inline long long clock64() { return 0; }
static inline void __nanosleep(unsigned) {}
static inline void __threadfence() {}
static inline void __syncwarp() {}
static inline unsigned atomicAdd(unsigned *p, unsigned v)
{
    const unsigned o = *p;
    *p = o + v;
    return o;
}
static inline unsigned __activemask() { return 1u << (threadIdx.x & 31); }   // the threads run one after the other
static inline unsigned __match_any_sync(unsigned m, int) { return m; }
static inline int __ffs(unsigned v) { return __builtin_ffs((int)v); }
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline long long __double_as_longlong(double d)
{
    long long v;
    memcpy(&v, &d, sizeof v);
    return v;
}
static inline double __longlong_as_double(long long v)
{
    double d;
    memcpy(&d, &v, sizeof d);
    return d;
}
// Shuffles of the thread-group kernel (8 lanes per interval): the 8 lanes of ONE group run as 8 host threads that meet at
// every shuffle (all lanes of a group take the same path), exchanging values through a slot per lane.
struct hostk_group_ctx {
    std::mutex m;
    std::condition_variable cv;
    int waiting = 0;
    unsigned long long gen = 0, slot[8] = {};
    void barrier()
    {
        std::unique_lock<std::mutex> l(m);
        const unsigned long long g = gen;
        if (++waiting == 8) {
            waiting = 0;
            ++gen;
            cv.notify_all();
        } else {
            cv.wait(l, [&] { return gen != g; });
        }
    }
};
static thread_local hostk_group_ctx *hostk_grp = nullptr;
template <typename T>
static inline T hostk_exchange(T v, int src)
{
    unsigned long long bits = 0;
    memcpy(&bits, &v, sizeof v);
    hostk_grp->slot[threadIdx.x & 7] = bits;
    hostk_grp->barrier();
    const unsigned long long r = hostk_grp->slot[src & 7];
    hostk_grp->barrier();
    T out;
    memcpy(&out, &r, sizeof out);
    return out;
}
template <typename T>
static inline T __shfl_sync(unsigned, T v, int src, int = 32) { return hostk_exchange(v, src); }
template <typename T>
static inline T __shfl_xor_sync(unsigned, T v, int d, int = 32) { return hostk_exchange(v, (int)(threadIdx.x & 7) ^ d); }

// round-to-nearest intrinsics of the constraint-terms kernel (written without contraction on the device: keep the
// host compiler from fusing them as well)
__attribute__((noinline)) static double __dmul_rn(double a, double b) { return a * b; }
__attribute__((noinline)) static double __dadd_rn(double a, double b) { return a + b; }
__attribute__((noinline)) static double __ddiv_rn(double a, double b) { return a / b; }
__attribute__((noinline)) static double __dsqrt_rn(double a) { return sqrt(a); }
// Stand-ins for the MUFU seeds (rcp.approx.ftz.f64 / rsqrt.approx.ftz.f64: ~20 good bits): the exact value rounded to
// float.  The kernels refine the seed with one third-order step, so results agree with the device to rounding.
static inline double hostk_rcp_seed(double a) { return (double)(float)(1.0 / a); }
static inline double hostk_rsqrt_seed(double a) { return (double)(float)(1.0 / sqrt(a)); }
