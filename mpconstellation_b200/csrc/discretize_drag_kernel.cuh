// discretize_drag_kernel.cuh -- fixed-step discretization WITH the drag branch of the linearisation.
//
// Discretizer(include_drag=True) (linearize_discretize.py:160-169) is unreachable with the reference's defaults
// (rho_func = None, Constants has no CD) but runs once the caller supplies const.CD, rho_func and drho_func.  This
// kernel covers that case for a density that depends on |r| (constant -- what Simulator.get_atmo_density returns,
// simulator.py:112 -- or any smooth radial model such as the power law / linear fits of simulator.py:110-111; the host
// fits rho_func and drho_func with Chebyshev series over the radii of the batch, DragLin):
//     f   = [v; a_g (+a_J2) + u/m - (kf/m) |v| v; mdot]                 kf = 1/2 C_D S rho_atm/RHO   (simulator.py:152)
//     Dxf = [[0 I 0],[G + W, V, d],[0 0 0]],   V = -(ka/m)(|v| I + v v^T/|v|),  d = -u/m^2 + (ka/m^2)|v| v,
//                                              W = Dr_aD = -(kc drho/m) |v| v r_hat^T              (:166)
//                                              ka = kc rho_func(r),  kc = 1/2 const.CD S           (:167-169)
// (the DYNAMICS keep the simulator's constant density; only the Jacobian reads rho_func / drho_func, as in the reference)
// The two structural shortcuts of discretize_kernel do not survive drag: the columns of Phi are no longer
// second-order systems in position only (V multiplies the velocity part) and Phi6 is no longer symplectic.  So this
// kernel is the plain formulation: classical RK4 on the first-order system in unscaled variables (step hs = tf h),
// and at every node a 6x6 Gauss-Jordan solve  Phi6 Z = [Duf_v | c | Sigma6 | xi'6]  (no pivoting: Phi6 over one
// interval is a small perturbation of [[I, tI],[tG, I]], pivots ~ 1).  One thread per interval; Phi and the state in
// registers (the compiler may spill: this is the rarely used mode, built for correctness); accumulators and the four
// stage linearisations in shared memory, [slot][thread].
#pragma once
#include "discretize_kernel.cuh"

namespace mpc {

constexpr int kDragStageSlots = 18;                      // G + W (9, not symmetric) V (6) d (3)
constexpr int kDragSlots = kAccSlots + 4 * kDragStageSlots;

// rho_func(|r|) and drho_func(|r|) of the drag linearisation as Chebyshev series in t = (|r| - r_mid) r_ihalf on the
// radii of the batch (fitted and verified on the host, discretizer.py; n_rho = 1, n_drho = 0: a constant density).
constexpr int kRhoCheb = 32;
struct DragLin {
    double kc;                 // 1/2 const.CD S
    double r_mid, r_ihalf;
    int n_rho, n_drho;
    double rho_c[kRhoCheb], drho_c[kRhoCheb];
};

__device__ __forceinline__ double cheb_eval(const double *c, int n, double t)
{
    // Clenshaw: sum_k c_k T_k(t)
    double b1 = 0.0, b2 = 0.0;
    const double t2 = 2.0 * t;
    for (int k = n - 1; k >= 1; --k) {
        const double b = fma(t2, b1, c[k]) - b2;
        b2 = b1;
        b1 = b;
    }
    return fma(t, b1, c[0]) - b2;
}

struct DragEval {
    double k[7];   // f / tf
    Sym3 g, v;
    double d[3];
    double im, iun, un, gr[3], dragv[3];   // dragv = (ka/m) |v| v
    double ux, uy, uz;
};

// what the drag branch adds to it: W = Dr_aD = gw r_hat^T (zero for a constant density).  A separate type so that the
// kernels without drag keep the layout (and the register allocation) they were tuned with.
struct DragEvalW : DragEval {
    double gw[3], rh[3];
};

template <bool J2>
__device__ __forceinline__ int drag_eval(const DiscParams &P, double kf, const DragLin &L, const double (&x)[7], double ux,
                                         double uy, double uz, DragEvalW &o)
{
    double ax, ay, az;
    gravity<J2>(P, x[0], x[1], x[2], ax, ay, az, o.g, o.gr);
    o.ux = ux;
    o.uy = uy;
    o.uz = uz;
    o.im = fast_rcp(x[6]);
    const double tx = ux * o.im, ty = uy * o.im, tz = uz * o.im;
    const double uu = fma(ux, ux, fma(uy, uy, uz * uz));
    o.iun = inv_norm_guarded(uu, 4.930380657631324e-32);
    o.un = uu * o.iun;
    const double vv = fma(x[3], x[3], fma(x[4], x[4], x[5] * x[5]));
    const double ivn = fast_rsqrt(vv), vn = vv * ivn;          // 1/|v|, |v|
    // density of the linearisation at this radius
    const double r2 = fma(x[0], x[0], fma(x[1], x[1], x[2] * x[2]));
    const double irn = fast_rsqrt(r2), rn = r2 * irn;
    double ka = L.kc * L.rho_c[0], kw = 0.0;
    if (L.n_rho > 1 || L.n_drho > 0) {
        const double t = (rn - L.r_mid) * L.r_ihalf;
        ka = L.kc * cheb_eval(L.rho_c, L.n_rho, t);
        if (L.n_drho > 0) kw = L.kc * cheb_eval(L.drho_c, L.n_drho, t);
    }
    const double cf = -kf * o.im * vn;                         // a_D = cf v
    o.k[0] = x[3];
    o.k[1] = x[4];
    o.k[2] = x[5];
    o.k[3] = fma(cf, x[3], ax + tx);
    o.k[4] = fma(cf, x[4], ay + ty);
    o.k[5] = fma(cf, x[5], az + tz);
    o.k[6] = -o.un * P.inv_ve;
    const double ca = ka * o.im;                               // ka/m
    const double cvn = -ca * vn, civ = -ca * ivn;
    o.v.xx = fma(civ * x[3], x[3], cvn);
    o.v.yy = fma(civ * x[4], x[4], cvn);
    o.v.zz = fma(civ * x[5], x[5], cvn);
    o.v.xy = civ * x[3] * x[4];
    o.v.xz = civ * x[3] * x[5];
    o.v.yz = civ * x[4] * x[5];
    const double cm = ca * vn;                                 // (ka/m) |v|
    o.dragv[0] = cm * x[3];
    o.dragv[1] = cm * x[4];
    o.dragv[2] = cm * x[5];
    o.d[0] = fma(o.dragv[0], o.im, -tx * o.im);
    o.d[1] = fma(o.dragv[1], o.im, -ty * o.im);
    o.d[2] = fma(o.dragv[2], o.im, -tz * o.im);
    // W = -(kc drho / m) |v| v r_hat^T;  (G + W) r = G r + gw |r|
    const double cw = -kw * o.im * vn;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        o.gw[a] = cw * x[3 + a];
        o.rh[a] = x[a] * irn;
        o.gr[a] = fma(o.gw[a], rn, o.gr[a]);
    }
    return !(x[6] > 0.0);
}

#define ACC(e) acc[(e) * BLOCK]

template <int BLOCK>
__device__ __forceinline__ void drag_store_stage(volatile double *acc, int s, const DragEvalW &e)
{
    const int b = kAccSlots + s * kDragStageSlots;
    // G + W, row-major
    ACC(b + 0) = fma(e.gw[0], e.rh[0], e.g.xx); ACC(b + 1) = fma(e.gw[0], e.rh[1], e.g.xy); ACC(b + 2) = fma(e.gw[0], e.rh[2], e.g.xz);
    ACC(b + 3) = fma(e.gw[1], e.rh[0], e.g.xy); ACC(b + 4) = fma(e.gw[1], e.rh[1], e.g.yy); ACC(b + 5) = fma(e.gw[1], e.rh[2], e.g.yz);
    ACC(b + 6) = fma(e.gw[2], e.rh[0], e.g.xz); ACC(b + 7) = fma(e.gw[2], e.rh[1], e.g.yz); ACC(b + 8) = fma(e.gw[2], e.rh[2], e.g.zz);
    ACC(b + 9) = e.v.xx; ACC(b + 10) = e.v.xy; ACC(b + 11) = e.v.xz; ACC(b + 12) = e.v.yy; ACC(b + 13) = e.v.yz; ACC(b + 14) = e.v.zz;
    ACC(b + 15) = e.d[0]; ACC(b + 16) = e.d[1]; ACC(b + 17) = e.d[2];
}

// derivative of one column y = (pr, pv) under stage s: (pv, G pr + V pv + d * mflag)
template <int BLOCK>
__device__ __forceinline__ void drag_col_rhs(volatile double *acc, int s, const double (&y)[6], double mflag, double (&k)[6])
{
    const int b = kAccSlots + s * kDragStageSlots;
    const Sym3 v = {ACC(b + 9), ACC(b + 10), ACC(b + 11), ACC(b + 12), ACC(b + 13), ACC(b + 14)};
    k[0] = y[3];
    k[1] = y[4];
    k[2] = y[5];
    // (G + W) pr + d mflag
    const double gx = fma(ACC(b + 2), y[2], fma(ACC(b + 1), y[1], fma(ACC(b + 0), y[0], ACC(b + 15) * mflag)));
    const double gy = fma(ACC(b + 5), y[2], fma(ACC(b + 4), y[1], fma(ACC(b + 3), y[0], ACC(b + 16) * mflag)));
    const double gz = fma(ACC(b + 8), y[2], fma(ACC(b + 7), y[1], fma(ACC(b + 6), y[0], ACC(b + 17) * mflag)));
    sym_mul_add(v, y[3], y[4], y[5], gx, gy, gz, k[3], k[4], k[5]);
}

// General-inverse quadrature node (unscaled variables): acc += w * Phi^-1 [Duf, Sigma, xi'], acc1 += ws * Phi^-1 Duf.
// Same slot layout as node_accumulate.  linearize_discretize.py:63-75.
// row6 != nullptr: the 8 accumulators of row 6 (slots 48..55) live at row6[0], row6[stride], ... (global memory) instead
// of the shared-memory slots.
template <int BLOCK>
__device__ __forceinline__ void node_accumulate_general(volatile double *acc, const double (&pr)[7][3], const double (&pv)[7][3],
                                                     const DiscParams &P, const DragEval &e, const double (&x)[7],
                                                     double w, double ws, double *row6 = nullptr, long long stride = 0)
{
    double M[6][6], R[6][6];
#pragma unroll
    for (int c = 0; c < 6; ++c)
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            M[a][c] = pr[c][a];
            M[a + 3][c] = pv[c][a];
        }
#pragma unroll
    for (int a = 0; a < 6; ++a) {
#pragma unroll
        for (int j = 0; j < 3; ++j) R[a][j] = (a == j + 3) ? e.im : 0.0;   // Duf rows 3..5 = I/m             (:201)
        R[a][3] = (a < 3) ? pr[6][a] : pv[6][a - 3];                        // c = Phi[0:6, 6]
        R[a][4] = e.k[a];                                                   // Sigma = f(tf=1)                (:252-253)
    }
    // xi' = -(Dxf x + Duf u): the -u/m and +u/m terms cancel;  V v + d_D m = -(ka/m)|v| v
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        R[a][5] = -x[3 + a];
        R[a + 3][5] = e.dragv[a] - e.gr[a];
    }
    // Gauss-Jordan without pivoting.  Only the columns of M right of the pivot are touched: the others are already
    // columns of the identity (and never read again), which saves 40 % of the elimination arithmetic and code.
    // (Skipping the Duf columns of R for the first three pivots -- their rows 0..2 are zero -- removes another 54
    // instructions but measured 3 % slower on the default-mode kernel, same box, r02w: ptxas schedules it worse.)
#pragma unroll
    for (int p = 0; p < 6; ++p) {
        const double ip = fast_rcp(M[p][p]);
#pragma unroll
        for (int c = p + 1; c < 6; ++c) M[p][c] *= ip;
#pragma unroll
        for (int c = 0; c < 6; ++c) R[p][c] *= ip;
#pragma unroll
        for (int r = 0; r < 6; ++r) {
            if (r == p) continue;
            const double f = M[r][p];
#pragma unroll
            for (int c = p + 1; c < 6; ++c) M[r][c] = fma(-f, M[p][c], M[r][c]);
#pragma unroll
            for (int c = 0; c < 6; ++c) R[r][c] = fma(-f, R[p][c], R[r][c]);
        }
    }
    // (requested after the solve -- 16 registers fewer while it runs -- and used last: the shared-memory updates cover
    // most of the L2 latency)
    double m6[8];
    if (row6) {
#pragma unroll
        for (int q = 0; q < 8; ++q) m6[q] = row6[(long long)q * stride];
    } else {
#pragma unroll
        for (int q = 0; q < 8; ++q) m6[q] = ACC(48 + q);
    }
    const double bs = -P.inv_ve * e.iun;
    const double b[3] = {bs * e.ux, bs * e.uy, bs * e.uz};      // last row of Duf (0 under the eps guard, :208)
    const double md = e.k[6];
    const double mdb = (e.iun != 0.0) ? md : 0.0;               // last row of Duf u
#pragma unroll
    for (int a = 0; a < 6; ++a) {
        const double ea = -R[a][3];                             // Phi^-1[0:6, 6] = -Phi6^-1 c
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            const double q = fma(ea, b[j], R[a][j]);
            ACC(a * 3 + j) = fma(w, q, ACC(a * 3 + j));
            ACC(18 + a * 3 + j) = fma(ws, q, ACC(18 + a * 3 + j));
        }
        ACC(36 + a) = fma(w, fma(ea, md, R[a][4]), ACC(36 + a));
        ACC(42 + a) = fma(w, fma(ea, -mdb, R[a][5]), ACC(42 + a));
    }
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        m6[j] = fma(w, b[j], m6[j]);
        m6[3 + j] = fma(ws, b[j], m6[3 + j]);
    }
    m6[6] = fma(w, md, m6[6]);
    m6[7] = fma(-w, mdb, m6[7]);
    if (row6) {
#pragma unroll
        for (int q = 0; q < 8; ++q) row6[(long long)q * stride] = m6[q];
    } else {
#pragma unroll
        for (int q = 0; q < 8; ++q) ACC(48 + q) = m6[q];
    }
}

template <bool J2, int BLOCK>
__global__ void __launch_bounds__(BLOCK)
discretize_drag_kernel(const double *__restrict__ x_in, const double *__restrict__ u_in, const double *__restrict__ tf_arr,
                       DiscParams P, double kf, const __grid_constant__ DragLin L, int n_sats, int K, int n_sub, DstTab dst,
                       long long pitch, long long offset, int32_t *__restrict__ status)
{
    extern __shared__ double acc_smem[];
    const long long n_int = (long long)n_sats * (K - 1);
    const long long gid = (long long)blockIdx.x * BLOCK + threadIdx.x;
    if (gid >= n_int) return;
    volatile double *acc = acc_smem + threadIdx.x;
    const int s = (int)(gid / (K - 1));
    const int k = (int)(gid - (long long)s * (K - 1));
    const double tf = tf_arr[s];
    const double *xs = x_in + ((long long)s * 7) * K + k;
    double x[7];
#pragma unroll
    for (int c = 0; c < 7; ++c) x[c] = xs[(long long)c * K];
    UHold<false> hold;
    hold.init(u_in, s, k, K, K);
    // The reference's 101 nodes: the trapezoid sums through their Euler-Maclaurin expansion (kEmW / kEmW2 in
    // discretize_kernel.cuh: 21 nodes = the ends of 20 steps, or 51 nodes / 50 steps on longer intervals), under the same
    // per-interval conditions as discretize_pair_kernel; drag is smooth in the state, so nothing else changes.
    const double *wt = nullptr;
    if (dst.em && n_sub == 100) {
        const double d2 = fma(hold.dux, hold.dux, fma(hold.duy, hold.duy, hold.duz * hold.duz));
        const double e0x = hold.u0x + hold.dux, e0y = hold.u0y + hold.duy, e0z = hold.u0z + hold.duz;
        const double a2 = fma(hold.u0x, hold.u0x, fma(hold.u0y, hold.u0y, hold.u0z * hold.u0z));
        const double b2 = fma(e0x, e0x, fma(e0y, e0y, e0z * e0z));
        const double r2 = fma(x[0], x[0], fma(x[1], x[1], x[2] * x[2]));
        const double hn = tf / (100.0 * (double)(K - 1));
        const double w2h2 = P.mu * hn * hn / (r2 * sqrt(r2));            // (omega h)^2 for the node spacing h
        if (d2 <= 0.0625 * fmax(a2, b2)) {
            if (w2h2 * 25.0 <= 1.2e-5) {
                wt = kEmW;
                n_sub = kEmSteps;
            } else if (w2h2 * 4.0 <= 1.94e-5) {
                wt = kEmW2;
                n_sub = kEmSteps2;
            }
        }
    }
    const double inv_n = 1.0 / (double)n_sub;
    const double h = inv_n / (double)(K - 1);
    const double hs = tf * h, hh = 0.5 * hs, h6 = hs * (1.0 / 6.0);

    double pr[7][3], pv[7][3];
#pragma unroll
    for (int c = 0; c < 7; ++c)
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            pr[c][a] = (c == a) ? 1.0 : 0.0;
            pv[c][a] = (c == a + 3) ? 1.0 : 0.0;
        }
#pragma unroll 1
    for (int e = 0; e < kDragSlots; ++e) ACC(e) = 0.0;

    int bad = 0;
    for (int n = 0; n <= n_sub; ++n) {
        double ux, uy, uz;
        if (n == 0 || n == n_sub) ref_node_input(u_in + (long long)s * 3 * K, K, k + (n == n_sub), 1.0, ux, uy, uz);   // end nodes
        else hold.at((double)n * inv_n, 0.0, ux, uy, uz);
        DragEvalW e1;
        bad |= drag_eval<J2>(P, kf, L, x, ux, uy, uz, e1);
        {
            const double w = wt ? wt[n] : ((n == 0 || n == n_sub) ? 0.5 : 1.0);
            node_accumulate_general<BLOCK>(acc, pr, pv, P, e1, x, w, w * ((double)n * inv_n));
        }
        if (n == n_sub) break;
        drag_store_stage<BLOCK>(acc, 0, e1);
        // ---- state: classical RK4 on the first-order system ------------------------------------------------
        double xt[7], ksum[7];
        double umx, umy, umz, uex, uey, uez;
        hold.at(((double)n + 0.5) * inv_n, 0.0, umx, umy, umz);
        hold.at((double)(n + 1) * inv_n, 0.0, uex, uey, uez);
        DragEvalW es;
#pragma unroll
        for (int i = 0; i < 7; ++i) {
            ksum[i] = e1.k[i];
            xt[i] = fma(hh, e1.k[i], x[i]);
        }
        bad |= drag_eval<J2>(P, kf, L, xt, umx, umy, umz, es);
        drag_store_stage<BLOCK>(acc, 1, es);
#pragma unroll
        for (int i = 0; i < 7; ++i) {
            ksum[i] = fma(2.0, es.k[i], ksum[i]);
            xt[i] = fma(hh, es.k[i], x[i]);
        }
        bad |= drag_eval<J2>(P, kf, L, xt, umx, umy, umz, es);
        drag_store_stage<BLOCK>(acc, 2, es);
#pragma unroll
        for (int i = 0; i < 7; ++i) {
            ksum[i] = fma(2.0, es.k[i], ksum[i]);
            xt[i] = fma(hs, es.k[i], x[i]);
        }
        bad |= drag_eval<J2>(P, kf, L, xt, uex, uey, uez, es);
        drag_store_stage<BLOCK>(acc, 3, es);
#pragma unroll
        for (int i = 0; i < 7; ++i) x[i] = fma(h6, ksum[i] + es.k[i], x[i]);
        // ---- columns of Phi ----------------------------------------------------------------------------------
#pragma unroll
        for (int c = 0; c < 7; ++c) {
            const double mflag = (c == 6) ? 1.0 : 0.0;
            double y[6] = {pr[c][0], pr[c][1], pr[c][2], pv[c][0], pv[c][1], pv[c][2]};
            double k1[6], k2[6], k3[6], k4[6], yt[6];
            drag_col_rhs<BLOCK>(acc, 0, y, mflag, k1);
#pragma unroll
            for (int i = 0; i < 6; ++i) yt[i] = fma(hh, k1[i], y[i]);
            drag_col_rhs<BLOCK>(acc, 1, yt, mflag, k2);
#pragma unroll
            for (int i = 0; i < 6; ++i) yt[i] = fma(hh, k2[i], y[i]);
            drag_col_rhs<BLOCK>(acc, 2, yt, mflag, k3);
#pragma unroll
            for (int i = 0; i < 6; ++i) yt[i] = fma(hs, k3[i], y[i]);
            drag_col_rhs<BLOCK>(acc, 3, yt, mflag, k4);
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                pr[c][i] = fma(h6, (k1[i] + k4[i]) + 2.0 * (k2[i] + k3[i]), y[i]);
                pv[c][i] = fma(h6, (k1[3 + i] + k4[3 + i]) + 2.0 * (k2[3 + i] + k3[3 + i]), y[3 + i]);
            }
        }
    }
    // B and xi carry tf (tf h = hs), Sigma does not (h)                               (:77-80, :182, :214, :252)
    const int nonfinite = epilogue_store<BLOCK, 1>(acc, pr, pv, hs, h, hs, dst, pitch, offset + gid);
    if (status) status[gid] = bad ? 1 : (nonfinite ? 2 : 0);
}
#undef ACC

}  // namespace mpc
