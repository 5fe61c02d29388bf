// TEST INFRASTRUCTURE ONLY -- see include/hostk_shim.h.  Wraps the host-compiled kernel sources (copied into _build/ by
// tests/hostk/__init__.py with the two inline-PTX seeds replaced) in extern "C" entry points that run the "threads" of a
// launch one after the other.
#include "hostk_shim.h"

#include <thread>
#include <vector>

#include "discretize_kernel.cuh"
#include "discretize_adaptive_kernel.cuh"
#include "discretize_default_kernel.cuh"
#include "discretize_pair_kernel.cuh"
#include "discretize_group_kernel.cuh"
#include "propagate_kernel.cuh"
#include "propagate_rk45_kernel.cuh"
#include "discretize_drag_kernel.cuh"
#include "constraint_terms_kernel.cuh"

namespace mpc {
double acc_smem[256 * 64];   // the kernels' `extern __shared__ double acc_smem[]` (<= 245 slots x BLOCK 32)
}

namespace {
constexpr int kBlock = 32;

template <typename F>
void run_grid(long long n_threads, F &&body)
{
    const long long grid = (n_threads + kBlock - 1) / kBlock;
    for (long long b = 0; b < grid; ++b)
        for (int t = 0; t < kBlock; ++t) {
            blockIdx.x = (unsigned)b;
            blockDim.x = kBlock;
            threadIdx.x = (unsigned)t;
            body();
        }
}

mpc::DiscParams disc_params(const double *c, int)
{
    mpc::DiscParams P;
    P.mu = c[0];
    P.kj2 = 1.5 * c[2] * c[0] * c[1] * c[1];
    P.inv_ve = 1.0 / (c[3] * c[4]);
    return P;
}
}  // namespace

// const8 = [MU, R_E, J2, G0, ISP, S, R0, RHO] (the order of mpc_params / OracleConstants)
extern "C" int hostk_discretize(const double *x, const double *u, const double *tf, const double *const8, int include_j2,
                                int n_sats, int K, int n_sub, int pair, int k0, int kc, double *out, long long pitch,
                                long long offset, int32_t *status, long long km_ntot, long long km_soff, int em)
{
    const mpc::DiscParams P = disc_params(const8, include_j2);
    mpc::DstTab dst{};
    dst.p[0] = out;
    dst.km_ntot = km_ntot;     // > 0: k-major layout, column = k km_ntot + km_soff + s
    dst.km_soff = km_soff;
    dst.em = em;               // the launcher's default: 1 (21-node form of the 101-node sums where the interval allows it)
    if (kc < 0) kc = K - 1;
    if (pair) {
        run_grid((long long)n_sats * kc, [&] {
            if (include_j2) mpc::discretize_pair_kernel<true, kBlock, 255, 1>(x, u, tf, P, n_sats, K, n_sub, dst, pitch, offset, status, k0, kc);
            else mpc::discretize_pair_kernel<false, kBlock, 255, 1>(x, u, tf, P, n_sats, K, n_sub, dst, pitch, offset, status, k0, kc);
        });
    } else {
        run_grid((long long)n_sats * (K - 1), [&] {
            if (include_j2) mpc::discretize_kernel<true, kBlock, 255, 1, false>(x, u, tf, P, n_sats, K, K, n_sub, dst, pitch, offset, status);
            else mpc::discretize_kernel<false, kBlock, 255, 1, false>(x, u, tf, P, n_sats, K, K, n_sub, dst, pitch, offset, status);
        });
    }
    return 0;
}

// the thread-group kernel (8 lanes per interval): the lanes of one group run as 8 host threads (see hostk_shim.h), group
// after group; threadIdx / blockIdx are what the 32-thread CTAs of the real launch would give them
extern "C" int hostk_discretize_group(const double *x, const double *u, const double *tf, const double *const8, int include_j2,
                                      int n_sats, int K, int n_sub, double *out, long long pitch, long long offset,
                                      int32_t *status, int extra_groups, int em)
{
    const mpc::DiscParams P = disc_params(const8, include_j2);
    mpc::DstTab dst{};
    dst.p[0] = out;
    dst.em = em;
    const long long n_groups = (long long)n_sats * (K - 1) + extra_groups;   // extra: the idle groups of a last partial warp
    for (long long g = 0; g < n_groups; ++g) {
        hostk_group_ctx ctx;
        std::vector<std::thread> lanes;
        for (int l = 0; l < 8; ++l)
            lanes.emplace_back([&, l] {
                hostk_grp = &ctx;
                blockDim.x = kBlock;
                blockIdx.x = (unsigned)(g / 4);
                threadIdx.x = (unsigned)((g % 4) * 8 + l);
                if (include_j2) mpc::discretize_group_kernel<true, kBlock>(x, u, tf, P, n_sats, K, n_sub, dst, pitch, offset, status);
                else mpc::discretize_group_kernel<false, kBlock>(x, u, tf, P, n_sats, K, n_sub, dst, pitch, offset, status);
            });
        for (auto &t : lanes) t.join();
    }
    return 0;
}

extern "C" int hostk_discretize_adaptive(const double *x, const double *u, const double *tf, const double *const8,
                                         int include_j2, int n_sats, int K, double rtol, double atol, double max_step,
                                         double *out, long long pitch, long long offset, int32_t *status, int32_t *n_nodes,
                                         int compact)
{
    const mpc::DiscParams P = disc_params(const8, include_j2);
    mpc::DstTab dst{};
    dst.p[0] = out;
    if (compact) {   // (parameter name kept) 1: the round-1 build, Phi in shared memory, both ends of every panel evaluated
        run_grid((long long)n_sats * (K - 1), [&] {
            if (include_j2)
                mpc::discretize_adaptive_kernel<true, kBlock, 1, false, false>(x, u, tf, P, n_sats, K, 0, rtol, atol, max_step, dst, pitch, offset, status, n_nodes);
            else
                mpc::discretize_adaptive_kernel<false, kBlock, 1, false, false>(x, u, tf, P, n_sats, K, 0, rtol, atol, max_step, dst, pitch, offset, status, n_nodes);
        });
        return 0;
    }
    run_grid((long long)n_sats * (K - 1), [&] {   // the shipped build
        if (include_j2)
            mpc::discretize_default_kernel<true, kBlock, false, false>(x, u, tf, P, n_sats, K, 0, rtol, atol, max_step, dst, pitch, offset, status, n_nodes);
        else
            mpc::discretize_default_kernel<false, kBlock, false, false>(x, u, tf, P, n_sats, K, 0, rtol, atol, max_step, dst, pitch, offset, status, n_nodes);
    });
    return 0;
}

// controller: kind 0 zero / 1 constant (t0,t1,t2) / 2 tangential (t0) / 3 sequence table [3][table_len] (shared) with end_tau
extern "C" int hostk_propagate(const double *y0, const double *tf, const double *const8, int include_j2, int include_drag,
                               double c_d, double rho_atm, int kind, const double *thrust, const double *table,
                               int table_len, int table_per_sat, double end_tau, int n_sats, int T, int n_sub, double *y,
                               double *u_out, int32_t *status, unsigned *progress, int seg_len)
{
    mpc::PropParams PP;
    PP.mu = const8[0];
    PP.kj2 = 1.5 * const8[2] * const8[0] * const8[1] * const8[1];
    PP.inv_ve = 1.0 / (const8[3] * const8[4]);
    PP.drag_k = include_drag ? 0.5 * c_d * const8[5] * (rho_atm / const8[7]) : 0.0;
    PP.include_j2 = include_j2;
    PP.include_drag = include_drag;
    mpc::CtrlParams C{};
    C.kind = kind;
    C.table_len = table_len;
    C.table_per_sat = table_per_sat;
    C.t0 = thrust[0];
    C.t1 = thrust[1];
    C.t2 = thrust[2];
    C.end_tau = end_tau;
    C.table = kind == 3 ? table : nullptr;
#define HK_PROP(KIND, DRAG, J2) \
    run_grid(n_sats, [&] { mpc::propagate_kernel<kBlock, KIND, DRAG, J2>(y0, tf, PP, C, n_sats, T, n_sub, y, u_out, status, progress, seg_len); })
#define HK_PROP_K(KIND)                                        \
    do {                                                       \
        if (include_drag) {                                    \
            if (include_j2) HK_PROP(KIND, true, true);         \
            else HK_PROP(KIND, true, false);                   \
        } else {                                               \
            if (include_j2) HK_PROP(KIND, false, true);        \
            else HK_PROP(KIND, false, false);                  \
        }                                                      \
    } while (0)
    switch (kind) {
        case 0: HK_PROP_K(0); break;
        case 1: HK_PROP_K(1); break;
        case 2: HK_PROP_K(2); break;
        default: HK_PROP_K(3); break;
    }
    return 0;
}

// the RK45 replay of the reference's integrator (propagate_rk45_kernel); spec = 1: with the speculative first stage
extern "C" int hostk_propagate_rk45(const double *y0, const double *tf, const double *const8, int include_j2, int include_drag,
                                    double c_d, double rho_atm, int kind, const double *thrust, const double *table,
                                    int table_len, int table_per_sat, double end_tau, const double *end_tau_arr, int n_sats,
                                    int T, double rtol, double atol, double max_step, int spec, int lpw, double *y,
                                    double *u_out, int32_t *status, int32_t *n_steps, unsigned *progress, int seg_len)
{
    mpc::PropParams PP;
    PP.mu = const8[0];
    PP.kj2 = 1.5 * const8[2] * const8[0] * const8[1] * const8[1];
    PP.inv_ve = 1.0 / (const8[3] * const8[4]);
    PP.drag_k = include_drag ? 0.5 * c_d * const8[5] * (rho_atm / const8[7]) : 0.0;
    PP.include_j2 = include_j2;
    PP.include_drag = include_drag;
    mpc::CtrlParams C{};
    C.kind = kind;
    C.table_len = table_len;
    C.table_per_sat = table_per_sat;
    C.t0 = thrust[0];
    C.t1 = thrust[1];
    C.t2 = thrust[2];
    C.end_tau = end_tau;
    C.table = kind == 3 ? table : nullptr;
    C.end_tau_arr = kind == 3 ? end_tau_arr : nullptr;
    const mpc::Rk45Opts O{rtol, atol, max_step};
    const long long n_threads = (long long)((n_sats + lpw - 1) / lpw) * 32;   // one warp per lpw satellites
#define HK_R45(KIND, DRAG, J2)                                                                                               \
    run_grid(n_threads, [&] {                                                                                                \
        if (spec) mpc::propagate_rk45_kernel<kBlock, KIND, DRAG, J2, true>(y0, tf, PP, C, O, n_sats, T, lpw, y, u_out, status, n_steps, progress, seg_len); \
        else mpc::propagate_rk45_kernel<kBlock, KIND, DRAG, J2, false>(y0, tf, PP, C, O, n_sats, T, lpw, y, u_out, status, n_steps, progress, seg_len);    \
    })
#define HK_R45_K(KIND)                                        \
    do {                                                      \
        if (include_drag) {                                   \
            if (include_j2) HK_R45(KIND, true, true);         \
            else HK_R45(KIND, true, false);                   \
        } else {                                              \
            if (include_j2) HK_R45(KIND, false, true);        \
            else HK_R45(KIND, false, false);                  \
        }                                                     \
    } while (0)
    switch (kind) {
        case 0: HK_R45_K(0); break;
        case 1: HK_R45_K(1); break;
        case 2: HK_R45_K(2); break;
        default: HK_R45_K(3); break;
    }
    return 0;
}

// drag branch of the linearisation: kf = 0.5 C_D S (rho_atm / RHO) (dynamics), ka = 0.5 const.CD S rho_func (Jacobian).
// model == nullptr: constant density (ka as given); else [r_mid, r_ihalf, n_rho, n_drho, rho_c[32], drho_c[32]] and
// ka = 0.5 const.CD S (the density comes from the series)
extern "C" int hostk_discretize_drag(const double *x, const double *u, const double *tf, const double *const8,
                                     int include_j2, double kf, double ka, int n_sats, int K, int n_sub, int adaptive,
                                     double rtol, double atol, double max_step, double *out, long long pitch,
                                     int32_t *status, int32_t *n_nodes, int em, const double *model)
{
    const mpc::DiscParams P = disc_params(const8, include_j2);
    mpc::DragLin L{};
    L.kc = ka;
    L.n_rho = 1;
    L.rho_c[0] = 1.0;
    if (model) {
        L.r_mid = model[0];
        L.r_ihalf = model[1];
        L.n_rho = (int)model[2];
        L.n_drho = (int)model[3];
        for (int i = 0; i < mpc::kRhoCheb; ++i) {
            L.rho_c[i] = model[4 + i];
            L.drho_c[i] = model[4 + mpc::kRhoCheb + i];
        }
    }
    mpc::DstTab dst{};
    dst.p[0] = out;
    dst.em = em;
    run_grid((long long)n_sats * (K - 1), [&] {
        if (adaptive) {
            if (include_j2)
                mpc::discretize_default_drag_kernel<true, kBlock>(x, u, tf, P, n_sats, K, rtol, atol, max_step, dst, pitch, 0, status, n_nodes, kf, L);
            else
                mpc::discretize_default_drag_kernel<false, kBlock>(x, u, tf, P, n_sats, K, rtol, atol, max_step, dst, pitch, 0, status, n_nodes, kf, L);
        } else {
            if (include_j2) mpc::discretize_drag_kernel<true, kBlock>(x, u, tf, P, kf, L, n_sats, K, n_sub, dst, pitch, 0, status);
            else mpc::discretize_drag_kernel<false, kBlock>(x, u, tf, P, kf, L, n_sats, K, n_sub, dst, pitch, 0, status);
        }
    });
    return 0;
}

extern "C" int hostk_constraint_terms(const double *x, const double *u, int n_sats, int K, int Ku, double mu,
                                      double *rbar_hat, double *ubar_hat, double *fin)
{
    run_grid((long long)n_sats * (K > Ku ? K : Ku),
             [&] { mpc::constraint_terms_kernel(x, u, n_sats, K, Ku, mu, 2.220446049250313e-16, rbar_hat, ubar_hat, fin); });
    return 0;
}

extern "C" int hostk_dynamics_jacobian(const double *soa, long long pitch, long long offset, int n_sats, int K,
                                       double *values, int64_t *indices, double *rhs)
{
    run_grid((long long)n_sats * 7 * (K - 1),
             [&] { mpc::dynamics_jacobian_kernel(soa, pitch, offset, n_sats, K, values, indices, rhs); });
    return 0;
}

// u on its own grid (u_cols columns; linearize_discretize.py:308-315): the GENU builds of the one-step-per-node kernel and
// of the adaptive kernel (what mpc_discretize_batch_ugrid launches)
extern "C" int hostk_discretize_ugrid(const double *x, const double *u, int u_cols, const double *tf, const double *const8,
                                      int include_j2, int n_sats, int K, int adaptive, int n_sub, double rtol, double atol,
                                      double max_step, double *out, long long pitch, int32_t *status, int32_t *n_nodes)
{
    const mpc::DiscParams P = disc_params(const8, include_j2);
    mpc::DstTab dst{};
    dst.p[0] = out;
    run_grid((long long)n_sats * (K - 1), [&] {
        if (adaptive) {
            if (include_j2) mpc::discretize_default_kernel<true, kBlock, true, false>(x, u, tf, P, n_sats, K, u_cols, rtol, atol, max_step, dst, pitch, 0, status, n_nodes);
            else mpc::discretize_default_kernel<false, kBlock, true, false>(x, u, tf, P, n_sats, K, u_cols, rtol, atol, max_step, dst, pitch, 0, status, n_nodes);
        } else {
            if (include_j2) mpc::discretize_kernel<true, kBlock, 255, 1, true>(x, u, tf, P, n_sats, K, u_cols, n_sub, dst, pitch, 0, status);
            else mpc::discretize_kernel<false, kBlock, 255, 1, true>(x, u, tf, P, n_sats, K, u_cols, n_sub, dst, pitch, 0, status);
        }
    });
    return 0;
}

// the end-node input lookup of the kernels (csrc/discretize_kernel.cuh: ref_node_input) for all nodes of one satellite:
// us [3][Ku] -> out [Ku][3]
extern "C" int hostk_ref_node_input(const double *us, int Ku, double f, double *out)
{
    for (int i = 0; i < Ku; ++i) mpc::ref_node_input(us, Ku, i, f, out[3 * i], out[3 * i + 1], out[3 * i + 2]);
    return 0;
}
