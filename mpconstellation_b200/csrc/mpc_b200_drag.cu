// mpc_b200_drag.cu -- the drag branch of the linearisation (linearize_discretize.py:160-169) in its OWN translation unit.
//
// The drag kernels are the rarely used mode (no caller of the reference enables drag in the discretizer), built for
// correctness.  They used to be instantiated in mpc_b200.cu; ptxas's register allocation of the 254-register hot kernels
// there (discretize_pair_kernel, discretize_default_kernel, propagate_rk45_kernel) turned out to depend on what else the
// module holds -- adding the density-model parameter to the drag kernels moved propagate_rk45_kernel from 200 to 214
// registers and discretize_default_kernel from 253 registers / 296 B of stack to 255 / 536 B (3-5 % slower, same PTX
// body).  So the hot kernels keep a module of their own and everything with DRAG = true is compiled here.
//
// The kernel headers define their __constant__ tables at namespace scope; a second inclusion under the same namespace
// would define them twice in the library, so this unit includes them under another namespace name.  The interface to
// mpc_b200.cu is plain C (opaque pointers to the parameter structs, whose layout is the same in both units).
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

#define mpc mpc_dragtu
#include "discretize_kernel.cuh"
#include "discretize_drag_kernel.cuh"
#include "discretize_adaptive_kernel.cuh"
#include "discretize_default_kernel.cuh"
#include "mpc_b200_drag.h"

namespace {

template <typename Kern>
cudaError_t configure(Kern kern, size_t smem, int &configured_dev)
{
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (configured_dev != dev) {
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        if (e != cudaSuccess) return e;
        configured_dev = dev;
    }
    return cudaSuccess;
}

template <bool J2>
cudaError_t fixed(const MpcDragLaunch &a)
{
    constexpr int BLOCK = 64;
    auto kern = mpc::discretize_drag_kernel<J2, BLOCK>;
    const size_t smem = (size_t)mpc::kDragSlots * BLOCK * sizeof(double);
    static thread_local int configured_dev = -1;
    cudaError_t e = configure(kern, smem, configured_dev);
    if (e != cudaSuccess) return e;
    const long long n_int = (long long)a.n_sats * (a.K - 1);
    const unsigned grid = (unsigned)((n_int + BLOCK - 1) / BLOCK);
    kern<<<grid, BLOCK, smem, a.stream>>>(a.x, a.u, a.tf, *(const mpc::DiscParams *)a.disc_params, a.kf,
                                          *(const mpc::DragLin *)a.drag_lin, a.n_sats, a.K, a.n_sub,
                                          *(const mpc::DstTab *)a.dst_tab, a.pitch, a.offset, a.status);
    return cudaGetLastError();
}

template <bool J2, int BLOCK>
cudaError_t adaptive(const MpcDragLaunch &a)
{
    auto kern = mpc::discretize_default_drag_kernel<J2, BLOCK>;
    const size_t smem = (size_t)mpc::kDfSlotsDrag * BLOCK * sizeof(double);
    static thread_local int configured_dev = -1;
    cudaError_t e = configure(kern, smem, configured_dev);
    if (e != cudaSuccess) return e;
    const long long n_int = (long long)a.n_sats * (a.K - 1);
    const unsigned grid = (unsigned)((n_int + BLOCK - 1) / BLOCK);
    kern<<<grid, BLOCK, smem, a.stream>>>(a.x, a.u, a.tf, *(const mpc::DiscParams *)a.disc_params, a.n_sats, a.K, a.rtol,
                                          a.atol, a.max_step, *(const mpc::DstTab *)a.dst_tab, a.pitch, a.offset, a.status,
                                          a.n_nodes, a.kf, *(const mpc::DragLin *)a.drag_lin);
    return cudaGetLastError();
}

// round-1 build of the default-mode kernel (mpc_set_tuning(9), A/B only; with the drag branch: constant density).  All of
// its variants live here: it is measurement ballast, not a path any caller takes by default.
template <bool J2, bool GENU, bool DRAG>
cudaError_t adaptive_v1(const MpcDragLaunch &a)
{
    constexpr int BLOCK = 32;
    auto kern = mpc::discretize_adaptive_kernel<J2, BLOCK, 1, GENU, DRAG>;
    const size_t smem = (size_t)(DRAG ? mpc::kAdSlotsDrag : mpc::kAdSlots) * BLOCK * sizeof(double);
    static thread_local int configured_dev = -1;
    cudaError_t e = configure(kern, smem, configured_dev);
    if (e != cudaSuccess) return e;
    const long long n_int = (long long)a.n_sats * (a.K - 1);
    const unsigned grid = (unsigned)((n_int + BLOCK - 1) / BLOCK);
    const mpc::DragLin &L = *(const mpc::DragLin *)a.drag_lin;
    kern<<<grid, BLOCK, smem, a.stream>>>(a.x, a.u, a.tf, *(const mpc::DiscParams *)a.disc_params, a.n_sats, a.K, a.ucols, a.rtol,
                                          a.atol, a.max_step, *(const mpc::DstTab *)a.dst_tab, a.pitch, a.offset, a.status,
                                          a.n_nodes, DRAG ? a.kf : 0.0, DRAG ? L.kc * L.rho_c[0] : 0.0);
    return cudaGetLastError();
}

template <bool J2>
cudaError_t adaptive_v1_any(const MpcDragLaunch &a)
{
    if (a.drag) return adaptive_v1<J2, false, true>(a);
    return a.ucols > 0 ? adaptive_v1<J2, true, false>(a) : adaptive_v1<J2, false, false>(a);
}

}  // namespace

size_t mpc_drag_sizeof(int which)
{
    return which == 0 ? sizeof(mpc::DiscParams) : (which == 1 ? sizeof(mpc::DstTab) : sizeof(mpc::DragLin));
}

cudaError_t mpc_drag_launch_fixed(const MpcDragLaunch *a) { return a->include_j2 ? fixed<true>(*a) : fixed<false>(*a); }

cudaError_t mpc_drag_launch_adaptive(const MpcDragLaunch *a)
{
    if (a->variant == 1) return a->include_j2 ? adaptive_v1_any<true>(*a) : adaptive_v1_any<false>(*a);
    // 174 slots per thread: 5 warps fill the SM's shared memory; small batches keep one-warp CTAs (see launch_adaptive_k)
    if (a->block == 32) return a->include_j2 ? adaptive<true, 32>(*a) : adaptive<false, 32>(*a);
    return a->include_j2 ? adaptive<true, 160>(*a) : adaptive<false, 160>(*a);
}
