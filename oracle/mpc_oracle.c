/*
 * TEST INFRASTRUCTURE ONLY -- plain-C CPU oracle for the batched SCvx discretize / propagate
 * hot path.  It checks the CUDA kernels (tests/, __graft_entry__.smoke) and is the timed
 * "port" CPU baseline in bench.py; the product library never links or calls it.
 *
 * What it restates (file:line under /root/reference):
 *   - dynamics f                      simulator.py:115-161
 *   - Jacobians A = tf*Dxf, B = tf*Duf linearize_discretize.py:119-183, :186-215
 *   - residual xi, Sigma              linearize_discretize.py:218-254
 *   - augmented IVP  [Phi;x]' = [A Phi; f]    linearize_discretize.py:262-290
 *   - FOH input interpolation         linearize_discretize.py:294-315
 *   - quadrature: inverse of Phi at every node, lambda weights, trapezoid rule, final
 *     left-multiply by Phi(tau_{k+1})  linearize_discretize.py:52-80
 *   - controller laws                 control.py:20-29, 47-53, 66-84, 104-143
 *   - propagation over tau in [0,1]   simulator.py:164-189
 *
 * Deliberately the *naive dense* formulation (56-vector classical RK4, dense 7x7 products,
 * Gauss-Jordan inverse with partial pivoting, the reference's literal J2 Jacobian) so that it
 * is independent of the structure-exploiting CUDA kernels it checks.
 *
 * The time integrator is fixed-step classical RK4 with n_sub steps per interval and the
 * trapezoid rule on the n_sub+1 step nodes -- the reference's `use_uniform_steps=True,
 * integrator_steps=n_sub+1` node set (linearize_discretize.py:27-28,47-48), with scipy's
 * adaptive RK45 + dense output replaced by RK4 on those nodes.  Pinned against the reference
 * through tests/golden (tests/test_oracle_golden.py): <= 1e-8 norm-relative.
 */
#include <math.h>
#include <string.h>
#include <stdlib.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct {
    double MU, R_E, J2, G0, ISP, S, R0, RHO; /* satellite_scale.py:42-44 */
    double C_D, RHO_ATM;                     /* constants.py:7, simulator.py:112 */
    double CD_A, RHO_A;                      /* const.CD and rho_func(r) of the discretizer's drag branch
                                                (linearize_discretize.py:162-168; constant density, drho = 0) */
    int include_J2, include_drag;
} orc_params;

enum { ORC_CTRL_ZERO = 0, ORC_CTRL_CONSTANT = 1, ORC_CTRL_TANGENTIAL = 2, ORC_CTRL_SEQUENCE = 3 };

static double norm3(const double *a) { return sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]); }

/* f / tf : simulator.py:130-160 */
static int dyn(const double *y, const double *u, const orc_params *p, int drag, int j2, double *dy)
{
    const double *r = y, *v = y + 3;
    double m = y[6];
    if (!(m > 0.0)) return 1;
    double rn = norm3(r);
    double g = -p->MU / (rn * rn * rn);
    for (int i = 0; i < 3; ++i) {
        dy[i] = v[i];
        dy[3 + i] = g * r[i] + u[i] / m;
    }
    if (drag) {
        double vn = norm3(v);
        double c = -0.5 * p->C_D * p->S * (1.0 / m) * (p->RHO_ATM / p->RHO) * vn;
        for (int i = 0; i < 3; ++i) dy[3 + i] += c * v[i];
    }
    if (j2) {
        double q = 5.0 * (r[2] / rn) * (r[2] / rn);
        double k = 1.5 * p->J2 * p->MU * p->R_E * p->R_E / pow(rn, 5);
        dy[3] += k * (q - 1.0) * r[0];
        dy[4] += k * (q - 1.0) * r[1];
        dy[5] += k * (q - 3.0) * r[2];
    }
    dy[6] = -norm3(u) / (p->G0 * p->ISP);
    return 0;
}

/* Dxf (7x7 row-major), linearize_discretize.py:134-179; drag != 0 adds the branch :160-169 with a constant
 * density (drho_func = 0, so Dr_aD vanishes) */
static void dxf_drag(const double *x, const orc_params *p, double *D)
{
    const double *v = x + 3;
    double m = x[6], vn = norm3(v);
    double kv = -p->RHO_A * p->CD_A * p->S / (2.0 * m);        /* :166 */
    double km = p->RHO_A * p->CD_A * p->S / (2.0 * m * m);     /* :168 */
    for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) D[(3 + i) * 7 + 3 + j] = kv * ((i == j ? vn : 0.0) + v[i] * v[j] / vn);
        D[(3 + i) * 7 + 6] += km * vn * v[i];
    }
}

static void dxf(const double *x, const double *u, const orc_params *p, int j2, double *D)
{
    memset(D, 0, 49 * sizeof(double));
    const double *r = x;
    double m = x[6];
    double rn = norm3(r), rn2 = rn * rn;
    double c3 = -p->MU / (rn2 * rn), c5 = 3.0 * p->MU / (rn2 * rn2 * rn);
    for (int i = 0; i < 3; ++i) {
        D[i * 7 + 3 + i] = 1.0;
        for (int j = 0; j < 3; ++j) D[(3 + i) * 7 + j] = (i == j ? c3 : 0.0) + c5 * r[i] * r[j];
        D[(3 + i) * 7 + 6] = -u[i] / (m * m);
    }
    if (j2) {
        double kJ2 = 1.5 * p->J2 * p->MU * p->R_E * p->R_E;
        double zz = (r[2] / rn) * (r[2] / rn);
        double gd[3] = {5 * zz - 1, 5 * zz - 1, 5 * zz - 3};
        double ddr[3];
        for (int j = 0; j < 3; ++j) ddr[j] = 5 * r[2] * r[2] * (-2 * r[j] / (rn2 * rn2));
        ddr[2] += (5 / rn2) * 2 * r[2];
        double rn5 = rn2 * rn2 * rn, rn7 = rn5 * rn2;
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j)
                D[(3 + i) * 7 + j] += kJ2 * gd[i] * r[i] * (-5 * r[j] / rn7) + kJ2 / rn5 * r[i] * ddr[j]
                                      + (i == j ? kJ2 / rn5 * gd[i] : 0.0);
    }
}

/* Duf (7x3 row-major), linearize_discretize.py:200-212 */
static void duf(const double *x, const double *u, const orc_params *p, double *D)
{
    memset(D, 0, 21 * sizeof(double));
    double m = x[6];
    for (int i = 0; i < 3; ++i) D[(3 + i) * 3 + i] = 1.0 / m;
    double nT = norm3(u);
    if (nT > 2.220446049250313e-16)
        for (int j = 0; j < 3; ++j) D[6 * 3 + j] = -u[j] / (p->G0 * p->ISP * nT);
}

/* augmented RHS, y = [Phi row-major (49), x (7)]; linearize_discretize.py:262-290 */
static int aug_rhs(const double *y, const double *u, double tf, const orc_params *p, int j2, double *dy)
{
    double D[49], f[7];
    dxf(y + 49, u, p, j2, D);
    if (p->include_drag) dxf_drag(y + 49, p, D);
    if (dyn(y + 49, u, p, p->include_drag, j2, f)) return 1;
    for (int i = 0; i < 7; ++i)
        for (int j = 0; j < 7; ++j) {
            double s = 0.0;
            for (int l = 0; l < 7; ++l) s += D[i * 7 + l] * y[l * 7 + j];
            dy[i * 7 + j] = tf * s;
        }
    for (int i = 0; i < 7; ++i) dy[49 + i] = tf * f[i];
    return 0;
}

static int inv7(const double *M, double *Inv)
{
    double a[7][14];
    for (int i = 0; i < 7; ++i)
        for (int j = 0; j < 7; ++j) {
            a[i][j] = M[i * 7 + j];
            a[i][7 + j] = (i == j);
        }
    for (int c = 0; c < 7; ++c) {
        int piv = c;
        for (int r = c + 1; r < 7; ++r)
            if (fabs(a[r][c]) > fabs(a[piv][c])) piv = r;
        if (a[piv][c] == 0.0) return 1;
        if (piv != c)
            for (int j = 0; j < 14; ++j) {
                double t = a[c][j];
                a[c][j] = a[piv][j];
                a[piv][j] = t;
            }
        double d = 1.0 / a[c][c];
        for (int j = 0; j < 14; ++j) a[c][j] *= d;
        for (int r = 0; r < 7; ++r)
            if (r != c) {
                double f = a[r][c];
                if (f != 0.0)
                    for (int j = 0; j < 14; ++j) a[r][j] -= f * a[c][j];
            }
    }
    for (int i = 0; i < 7; ++i)
        for (int j = 0; j < 7; ++j) Inv[i * 7 + j] = a[i][7 + j];
    return 0;
}

/* Integration scheme of the fixed-step mode.  2 (default): scheme 1 with steps spanning two quadrature nodes where that
 * is accurate (see interval()).  1: Nystrom's 3-stage fourth-order Runge-Kutta method for
 * y'' = f(t, y) -- what the CUDA kernel runs: positions / velocities of the state and of every Phi column are a
 * second-order system (no drag in the discretizer), the mass is a quadrature of mdot(tau).  0: the classical RK4 on
 * the first-order 56-vector (also used whenever drag is on: the acceleration then depends on the velocity). */
static int g_scheme = 2;
void orc_set_scheme(int scheme) { g_scheme = scheme; }

/* accelerations of the second-order part at the stage vector ys (only its positions, mass and Phi row 6 matter) */
static int accel(const double *ys, const double *u, const orc_params *p, int j2, double *ax, double *aP)
{
    double f[7], D[49];
    if (dyn(ys + 49, u, p, 0, j2, f)) return 1;
    dxf(ys + 49, u, p, j2, D);
    for (int i = 0; i < 3; ++i) {
        ax[i] = f[3 + i];
        for (int j = 0; j < 7; ++j) {
            double s = 0.0;
            for (int l = 0; l < 7; ++l) s += D[(3 + i) * 7 + l] * ys[l * 7 + j];
            aP[i * 7 + j] = s;
        }
    }
    return 0;
}

/* one RKN4 step of size hs (unscaled time) from y; inputs at the start / middle / end of the step */
static int rkn4_step(double *y, const double *u1, const double *um, const double *ue, double hs, const orc_params *p, int j2)
{
    double a1[3], a2[3], a3[3], P1[21], P2[21], P3[21], ys[56];
    const double ve = p->G0 * p->ISP;
    const double md1 = -norm3(u1) / ve, mdm = -norm3(um) / ve, mde = -norm3(ue) / ve;
    const double m = y[55];
    const double m2 = m + hs * (5.0 * md1 + 8.0 * mdm - mde) / 24.0;   /* integral of the quadratic through the 3 values */
    const double m3 = m + hs * (md1 + 4.0 * mdm + mde) / 6.0;          /* Simpson */
    if (accel(y, u1, p, j2, a1, P1)) return 1;
    memcpy(ys, y, sizeof ys);
    for (int i = 0; i < 3; ++i) {
        ys[49 + i] = y[49 + i] + 0.5 * hs * y[52 + i] + hs * hs / 8.0 * a1[i];
        for (int j = 0; j < 7; ++j) ys[i * 7 + j] = y[i * 7 + j] + 0.5 * hs * y[(3 + i) * 7 + j] + hs * hs / 8.0 * P1[i * 7 + j];
    }
    ys[55] = m2;
    if (accel(ys, um, p, j2, a2, P2)) return 1;
    for (int i = 0; i < 3; ++i) {
        ys[49 + i] = y[49 + i] + hs * y[52 + i] + hs * hs / 2.0 * a2[i];
        for (int j = 0; j < 7; ++j) ys[i * 7 + j] = y[i * 7 + j] + hs * y[(3 + i) * 7 + j] + hs * hs / 2.0 * P2[i * 7 + j];
    }
    ys[55] = m3;
    if (accel(ys, ue, p, j2, a3, P3)) return 1;
    for (int i = 0; i < 3; ++i) {
        y[49 + i] += hs * y[52 + i] + hs * hs / 6.0 * (a1[i] + 2.0 * a2[i]);
        y[52 + i] += hs / 6.0 * (a1[i] + 4.0 * a2[i] + a3[i]);
        for (int j = 0; j < 7; ++j) {
            y[i * 7 + j] += hs * y[(3 + i) * 7 + j] + hs * hs / 6.0 * (P1[i * 7 + j] + 2.0 * P2[i * 7 + j]);
            y[(3 + i) * 7 + j] += hs / 6.0 * (P1[i * 7 + j] + 4.0 * P2[i * 7 + j] + P3[i * 7 + j]);
        }
    }
    y[55] = m3;
    return !(m3 > 0.0);
}

/* One interval; out = A(49) B_kp(21) B_kn(21) Sigma(7) xi(7); linearize_discretize.py:8-82.
 * u0/u1 are the FOH end points of this interval (u is linear inside one interval, :305-315). */
/* Python's float floor division for v >= 0, w > 0 (CPython float_divmod); differs from floor(v / w): 0.5 // 0.1 == 4 */
static double py_floordiv(double v, double w)
{
    double mod = fmod(v, w), div = (v - mod) / w;
    if (div == 0.0) return 0.0;
    double fl = floor(div);
    if (div - fl > 0.5) fl += 1.0;
    return fl;
}

/* Discretizer.u_FOH(tau_i, u) for node i of tau = np.linspace(0, 1, K), statement for statement
 * (linearize_discretize.py:294-315): the lookup k = int(tau // dtau) is on the GLOBAL grid, so a node may be looked up
 * in the interval next to it -- decisive where u is exactly 0 at the node (|u| <= eps guard of B_func, :208).
 * us: [3][K] of one satellite. */
static void ref_node_input(const double *us, int K, int i, double *un)
{
    if (i >= K - 1) {
        for (int c = 0; c < 3; ++c) un[c] = us[c * K + K - 1];
        return;
    }
    double dtau = 1.0 / (K - 1), tau = i * dtau;
    int k = (int)py_floordiv(tau, dtau);
    if (k > K - 2) k = K - 2;
    double tk = (double)k / (K - 1), tk1 = (double)(k + 1) / (K - 1);
    double ln = (tk1 - tau) / (tk1 - tk), lp = (tau - tk) / (tk1 - tk);
    for (int c = 0; c < 3; ++c) un[c] = ln * us[c * K + k] + lp * us[c * K + k + 1];
}

static int interval(const double *xk, const double *u0, const double *u1, const double *e0, const double *e1, double tf,
                    double dtau, int n_sub, const orc_params *p, int j2, double *out)
{
    double y[56], k1[56], k2[56], k3[56], k4[56], yt[56], ypend[56];
    memset(y, 0, sizeof y);
    for (int i = 0; i < 7; ++i) {
        y[i * 7 + i] = 1.0;
        y[49 + i] = xk[i];
    }
    double accBp[21] = {0}, accBn[21] = {0}, accS[7] = {0}, accX[7] = {0};
    double h = dtau / n_sub;
    /* scheme 2 (default, what the CUDA kernels do): integrator steps spanning two nodes with a Hermite midpoint, where
     * the step is short against the local orbital rate (omega H <= 3.2e-3 rad) and the number of panels is even;
     * one step per node otherwise */
    int use_pair = 0;
    if (g_scheme == 2 && !p->include_drag && n_sub % 2 == 0) {
        double rn = norm3(xk), H = 2.0 * tf * h;
        use_pair = (p->MU * H * H / (rn * rn * rn) <= 1.0e-5);
    }
    for (int n = 0; n <= n_sub; ++n) {
        double lam_p = (double)n / n_sub, lam_n = 1.0 - lam_p;
        double un[3];
        for (int i = 0; i < 3; ++i) un[i] = lam_n * u0[i] + lam_p * u1[i];
        /* the two end nodes: the input the reference looks up there (ref_node_input) */
        if (n == 0) memcpy(un, e0, sizeof un);
        if (n == n_sub) memcpy(un, e1, sizeof un);
        /* node terms, :63-75 */
        double Pinv[49], Bm[21], Dx[49], sig[7], xi[7];
        const double *xs = y + 49;
        if (inv7(y, Pinv)) return 3;
        duf(xs, un, p, Bm);
        for (int i = 0; i < 21; ++i) Bm[i] *= tf;
        if (dyn(xs, un, p, p->include_drag, j2, sig)) return 1;
        dxf(xs, un, p, j2, Dx);
        if (p->include_drag) dxf_drag(xs, p, Dx);
        for (int i = 0; i < 7; ++i) {
            double s = 0.0;
            for (int l = 0; l < 7; ++l) s += tf * Dx[i * 7 + l] * xs[l];
            for (int l = 0; l < 3; ++l) s += Bm[i * 3 + l] * un[l];
            xi[i] = -s;
        }
        double w = (n == 0 || n == n_sub) ? 0.5 * h : h; /* uniform-node trapezoid, :77-80 */
        for (int i = 0; i < 7; ++i) {
            double ss = 0.0, sx = 0.0;
            for (int l = 0; l < 7; ++l) {
                ss += Pinv[i * 7 + l] * sig[l];
                sx += Pinv[i * 7 + l] * xi[l];
            }
            accS[i] += w * ss;
            accX[i] += w * sx;
            for (int j = 0; j < 3; ++j) {
                double sb = 0.0;
                for (int l = 0; l < 7; ++l) sb += Pinv[i * 7 + l] * Bm[l * 3 + j];
                accBp[i * 3 + j] += w * lam_p * sb;
                accBn[i * 3 + j] += w * lam_n * sb;
            }
        }
        if (n == n_sub) break;
        /* classical RK4 step on the 56-vector */
        double um[3], ue[3];
        double lm = (n + 0.5) / n_sub, le = (n + 1.0) / n_sub;
        for (int i = 0; i < 3; ++i) {
            um[i] = (1.0 - lm) * u0[i] + lm * u1[i];
            ue[i] = (1.0 - le) * u0[i] + le * u1[i];
        }
        if (use_pair) {
            /* RKN4 steps spanning TWO quadrature nodes; the odd node from the cubic Hermite interpolant of the step
             * (positions and velocities at both ends), the mass there from the quadratic through the three mdot values */
            if (n % 2 == 1) {   /* odd node done: move on to the end of the pending step */
                memcpy(y, ypend, sizeof y);
                continue;
            }
            double u2m[3], u2e[3];
            double l1 = (n + 1.0) / n_sub, l2 = (n + 2.0) / n_sub;
            for (int i = 0; i < 3; ++i) {
                u2m[i] = (1.0 - l1) * u0[i] + l1 * u1[i];
                u2e[i] = (1.0 - l2) * u0[i] + l2 * u1[i];
            }
            const double H = 2.0 * tf * h, ve = p->G0 * p->ISP;
            const double md1 = -norm3(un) / ve, mdm = -norm3(u2m) / ve, mde = -norm3(u2e) / ve;
            memcpy(ypend, y, sizeof y);
            if (rkn4_step(ypend, un, u2m, u2e, H, p, j2)) return 1;
            double ymid[56];
            memcpy(ymid, y, sizeof y);
            for (int i = 0; i < 3; ++i) {
                for (int j = -1; j < 7; ++j) {   /* j = -1: the state, j >= 0: column j of Phi */
                    int ip = (j < 0) ? 49 + i : i * 7 + j, iv = (j < 0) ? 52 + i : (3 + i) * 7 + j;
                    double p0 = y[ip], p1 = ypend[ip], v0 = y[iv], v1 = ypend[iv];
                    ymid[ip] = 0.5 * (p0 + p1) + H / 8.0 * (v0 - v1);
                    ymid[iv] = 1.5 * (p1 - p0) / H - 0.25 * (v0 + v1);
                }
            }
            ymid[55] = y[55] + H * (5.0 * md1 + 8.0 * mdm - mde) / 24.0;
            memcpy(y, ymid, sizeof y);
            continue;
        }
        if (g_scheme >= 1 && !p->include_drag) {
            if (rkn4_step(y, un, um, ue, tf * h, p, j2)) return 1;
            continue;
        }
        if (aug_rhs(y, un, tf, p, j2, k1)) return 1;
        for (int i = 0; i < 56; ++i) yt[i] = y[i] + 0.5 * h * k1[i];
        if (aug_rhs(yt, um, tf, p, j2, k2)) return 1;
        for (int i = 0; i < 56; ++i) yt[i] = y[i] + 0.5 * h * k2[i];
        if (aug_rhs(yt, um, tf, p, j2, k3)) return 1;
        for (int i = 0; i < 56; ++i) yt[i] = y[i] + h * k3[i];
        if (aug_rhs(yt, ue, tf, p, j2, k4)) return 1;
        for (int i = 0; i < 56; ++i) y[i] += h / 6.0 * (k1[i] + 2.0 * k2[i] + 2.0 * k3[i] + k4[i]);
    }
    memcpy(out, y, 49 * sizeof(double));
    for (int i = 0; i < 7; ++i) {
        double ss = 0.0, sx = 0.0;
        for (int l = 0; l < 7; ++l) {
            ss += y[i * 7 + l] * accS[l];
            sx += y[i * 7 + l] * accX[l];
        }
        out[91 + i] = ss;
        out[98 + i] = sx;
        for (int j = 0; j < 3; ++j) {
            double sp = 0.0, sn = 0.0;
            for (int l = 0; l < 7; ++l) {
                sp += y[i * 7 + l] * accBp[l * 3 + j];
                sn += y[i * 7 + l] * accBn[l * 3 + j];
            }
            out[49 + i * 3 + j] = sp;
            out[70 + i * 3 + j] = sn;
        }
    }
    for (int i = 0; i < 105; ++i)
        if (!isfinite(out[i])) return 2;
    return 0;
}

/*
 * Batched discretization.  x [N][7][K], u [N][3][K], tf [N]; out [N][K-1][105]
 * (A row-major 49 | B_kp 21 | B_kn 21 | Sigma 7 | xi 7); status [N*(K-1)].
 * nthreads <= 0: all OpenMP threads.  Returns the number of failed intervals.
 */
int orc_discretize_rk4(const double *x, const double *u, const double *tf, const orc_params *p, int N, int K,
                       int n_sub, double *out, int *status, int nthreads)
{
    long total = (long)N * (K - 1);
    int bad = 0;
    if (K < 2 || N < 1) return 0;
    double dtau = 1.0 / (K - 1);
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel for schedule(static) reduction(+ : bad)
    for (long i = 0; i < total; ++i) {
        int s = (int)(i / (K - 1)), k = (int)(i % (K - 1));
        double xk[7], u0[3], u1[3];
        for (int c = 0; c < 7; ++c) xk[c] = x[((long)s * 7 + c) * K + k];
        for (int c = 0; c < 3; ++c) {
            u0[c] = u[((long)s * 3 + c) * K + k];
            u1[c] = u[((long)s * 3 + c) * K + k + 1];
        }
        double e0[3], e1[3];
        ref_node_input(u + (long)s * 3 * K, K, k, e0);
        ref_node_input(u + (long)s * 3 * K, K, k + 1, e1);
        int st = interval(xk, u0, u1, e0, e1, tf[s], dtau, n_sub, p, p->include_J2, out + i * 105);
        if (status) status[i] = st;
        bad += (st != 0);
    }
    return bad;
}

/* ------------------------------------------------------------------------------------------------
 * Adaptive mode: the reference's DEFAULT quadrature (use_uniform_steps=False, linearize_discretize.py:
 * 29-30,49-50): nodes are the accepted steps of scipy.integrate.solve_ivp(method='RK45', max_step=1e-2,
 * rtol=1e-3, atol=1e-6).  scipy is a third-party dependency the reference neither vendors nor pins; this
 * restates its published algorithm (scipy 1.18.1, integrate/_ivp/rk.py: rk_step, RungeKutta._step_impl,
 * RK45 tableau; integrate/_ivp/common.py: select_initial_step, norm) -- Dormand-Prince 5(4), local
 * extrapolation, error norm = RMS of err/(atol + rtol*max(|y|,|y_new|)), SAFETY 0.9, factor clamp
 * [0.2, 10], first step from Hairer's heuristic.
 */
#define ORC_MAX_NODES 512

static const double DP_C[6] = {0.0, 1.0 / 5, 3.0 / 10, 4.0 / 5, 8.0 / 9, 1.0};
static const double DP_A[6][5] = {{0, 0, 0, 0, 0},
                                  {1.0 / 5, 0, 0, 0, 0},
                                  {3.0 / 40, 9.0 / 40, 0, 0, 0},
                                  {44.0 / 45, -56.0 / 15, 32.0 / 9, 0, 0},
                                  {19372.0 / 6561, -25360.0 / 2187, 64448.0 / 6561, -212.0 / 729, 0},
                                  {9017.0 / 3168, -355.0 / 33, 46732.0 / 5247, 49.0 / 176, -5103.0 / 18656}};
static const double DP_B[6] = {35.0 / 384, 0, 500.0 / 1113, 125.0 / 192, -2187.0 / 6784, 11.0 / 84};
static const double DP_E[7] = {-71.0 / 57600, 0, 71.0 / 16695, -71.0 / 1920, 17253.0 / 339200, -22.0 / 525, 1.0 / 40};

static double rms56(const double *v)
{
    double s = 0.0;
    for (int i = 0; i < 56; ++i) s += v[i] * v[i];
    return sqrt(s) / sqrt(56.0);
}

typedef struct {
    const double *u0, *u1;
    double t0, t1, tf;
    const orc_params *p;
    int j2;
    const double *e0, *e1; /* inputs at the two end nodes as the reference looks them up (ref_node_input) */
} aug_ctx;

static int aug_fun(const aug_ctx *c, double t, const double *y, double *dy)
{
    /* FOH inside one interval (linearize_discretize.py:305-315) */
    double lp = (t - c->t0) / (c->t1 - c->t0), ln = (c->t1 - t) / (c->t1 - c->t0);
    double u[3];
    for (int i = 0; i < 3; ++i) u[i] = ln * c->u0[i] + lp * c->u1[i];
    return aug_rhs(y, u, c->tf, c->p, c->j2, dy);
}

/* quadrature-node integrands (linearize_discretize.py:63-75) at state y (Phi, x) and time t:
 * g[0..20] = Phi^-1 B lam+, g[21..41] = Phi^-1 B lam-, g[42..48] = Phi^-1 Sigma, g[49..55] = Phi^-1 xi */
static int node_integrands(const aug_ctx *c, double t, const double *y, double *g)
{
    double lp = (t - c->t0) / (c->t1 - c->t0), ln = (c->t1 - t) / (c->t1 - c->t0);
    double un[3], Pinv[49], Bm[21], Dx[49], sig[7], xi[7];
    const double *xs = y + 49;
    for (int i = 0; i < 3; ++i) un[i] = ln * c->u0[i] + lp * c->u1[i];
    if (t == c->t0) memcpy(un, c->e0, sizeof un);
    if (t == c->t1) memcpy(un, c->e1, sizeof un);
    if (inv7(y, Pinv)) return 3;
    duf(xs, un, c->p, Bm);
    for (int i = 0; i < 21; ++i) Bm[i] *= c->tf;
    if (dyn(xs, un, c->p, c->p->include_drag, c->j2, sig)) return 1;
    dxf(xs, un, c->p, c->j2, Dx);
    if (c->p->include_drag) dxf_drag(xs, c->p, Dx);
    for (int i = 0; i < 7; ++i) {
        double s = 0.0;
        for (int l = 0; l < 7; ++l) s += c->tf * Dx[i * 7 + l] * xs[l];
        for (int l = 0; l < 3; ++l) s += Bm[i * 3 + l] * un[l];
        xi[i] = -s;
    }
    for (int i = 0; i < 7; ++i) {
        double ss = 0.0, sx = 0.0;
        for (int l = 0; l < 7; ++l) {
            ss += Pinv[i * 7 + l] * sig[l];
            sx += Pinv[i * 7 + l] * xi[l];
        }
        g[42 + i] = ss;
        g[49 + i] = sx;
        for (int j = 0; j < 3; ++j) {
            double sb = 0.0;
            for (int l = 0; l < 7; ++l) sb += Pinv[i * 7 + l] * Bm[l * 3 + j];
            g[i * 3 + j] = sb * lp;
            g[21 + i * 3 + j] = sb * ln;
        }
    }
    return 0;
}

static int interval_rk45(const double *xk, const double *u0, const double *u1, const double *e0, const double *e1, double tf,
                         double t0, double t1, double rtol, double atol, double max_step, const orc_params *p, int j2,
                         double *out, int *n_nodes)
{
    aug_ctx c = {u0, u1, t0, t1, tf, p, j2, e0, e1};
    double y[56], f[56], K[7][56], ynew[56], fnew[56], tmp[56], sc[56];
    double acc[56] = {0}, gprev[56], gcur[56];
    memset(y, 0, sizeof y);
    for (int i = 0; i < 7; ++i) {
        y[i * 7 + i] = 1.0;
        y[49 + i] = xk[i];
    }
    double t = t0;
    if (aug_fun(&c, t, y, f)) return 1;
    /* select_initial_step (common.py) */
    double h_abs;
    {
        double interval_length = fabs(t1 - t0);
        for (int i = 0; i < 56; ++i) sc[i] = atol + fabs(y[i]) * rtol;
        for (int i = 0; i < 56; ++i) tmp[i] = y[i] / sc[i];
        double d0 = rms56(tmp);
        for (int i = 0; i < 56; ++i) tmp[i] = f[i] / sc[i];
        double d1 = rms56(tmp);
        double h0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : 0.01 * d0 / d1;
        if (h0 > interval_length) h0 = interval_length;
        for (int i = 0; i < 56; ++i) ynew[i] = y[i] + h0 * f[i];
        if (aug_fun(&c, t + h0, ynew, fnew)) return 1;
        for (int i = 0; i < 56; ++i) tmp[i] = (fnew[i] - f[i]) / sc[i];
        double d2 = rms56(tmp) / h0;
        double h1 = (d1 <= 1e-15 && d2 <= 1e-15) ? fmax(1e-6, h0 * 1e-3) : pow(0.01 / fmax(d1, d2), 1.0 / 5.0);
        h_abs = fmin(fmin(100 * h0, h1), fmin(interval_length, max_step));
    }
    int st = node_integrands(&c, t, y, gprev);
    if (st) return st;
    int nodes = 1;
    while (t < t1) {
        double min_step = 10 * fabs(nextafter(t, INFINITY) - t);
        if (h_abs > max_step) h_abs = max_step;
        else if (h_abs < min_step) h_abs = min_step;
        int accepted = 0, rejected = 0;
        double t_new = t, h = 0.0;
        while (!accepted) {
            if (h_abs < min_step) return 4;
            h = h_abs;
            t_new = t + h;
            if (t_new - t1 > 0) t_new = t1;
            h = t_new - t;
            h_abs = fabs(h);
            /* rk_step (rk.py) */
            memcpy(K[0], f, sizeof f);
            for (int s = 1; s < 6; ++s) {
                for (int i = 0; i < 56; ++i) {
                    double dy = 0.0;
                    for (int l = 0; l < s; ++l) dy += K[l][i] * DP_A[s][l];
                    tmp[i] = y[i] + dy * h;
                }
                if (aug_fun(&c, t + DP_C[s] * h, tmp, K[s])) return 1;
            }
            for (int i = 0; i < 56; ++i) {
                double d = 0.0;
                for (int l = 0; l < 6; ++l) d += K[l][i] * DP_B[l];
                ynew[i] = y[i] + h * d;
            }
            if (aug_fun(&c, t + h, ynew, fnew)) return 1;
            memcpy(K[6], fnew, sizeof fnew);
            for (int i = 0; i < 56; ++i) {
                double e = 0.0;
                for (int l = 0; l < 7; ++l) e += K[l][i] * DP_E[l];
                double scale = atol + fmax(fabs(y[i]), fabs(ynew[i])) * rtol;
                tmp[i] = e * h / scale;
            }
            double err = rms56(tmp);
            if (err < 1.0) {
                double factor = (err == 0.0) ? 10.0 : fmin(10.0, 0.9 * pow(err, -0.2));
                if (rejected && factor > 1.0) factor = 1.0;
                h_abs *= factor;
                accepted = 1;
            } else {
                h_abs *= fmax(0.2, 0.9 * pow(err, -0.2));
                rejected = 1;
            }
        }
        /* node t_new: trapezoid panel [t, t_new] (np.trapz with x = sol.t, :77-80) */
        st = node_integrands(&c, t_new, ynew, gcur);
        if (st) return st;
        double w = 0.5 * (t_new - t);
        for (int i = 0; i < 56; ++i) acc[i] += w * (gprev[i] + gcur[i]);
        memcpy(gprev, gcur, sizeof gcur);
        memcpy(y, ynew, sizeof y);
        memcpy(f, fnew, sizeof f);
        t = t_new;
        if (++nodes > ORC_MAX_NODES) return 4;
    }
    if (n_nodes) *n_nodes = nodes;
    memcpy(out, y, 49 * sizeof(double));
    for (int i = 0; i < 7; ++i) {
        double ss = 0.0, sx = 0.0;
        for (int l = 0; l < 7; ++l) {
            ss += y[i * 7 + l] * acc[42 + l];
            sx += y[i * 7 + l] * acc[49 + l];
        }
        out[91 + i] = ss;
        out[98 + i] = sx;
        for (int j = 0; j < 3; ++j) {
            double sp = 0.0, sn = 0.0;
            for (int l = 0; l < 7; ++l) {
                sp += y[i * 7 + l] * acc[l * 3 + j];
                sn += y[i * 7 + l] * acc[21 + l * 3 + j];
            }
            out[49 + i * 3 + j] = sp;
            out[70 + i * 3 + j] = sn;
        }
    }
    for (int i = 0; i < 105; ++i)
        if (!isfinite(out[i])) return 2;
    return 0;
}

/* Batched adaptive-mode discretization; same layouts as orc_discretize_rk4; n_nodes [N*(K-1)] may be NULL. */
int orc_discretize_rk45(const double *x, const double *u, const double *tf, const orc_params *p, int N, int K,
                        double rtol, double atol, double max_step, double *out, int *status, int *n_nodes,
                        int nthreads)
{
    long total = (long)N * (K - 1);
    int bad = 0;
    if (K < 2 || N < 1) return 0;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel for schedule(dynamic, 16) reduction(+ : bad)
    for (long i = 0; i < total; ++i) {
        int s = (int)(i / (K - 1)), k = (int)(i % (K - 1));
        double xk[7], u0[3], u1[3];
        for (int c = 0; c < 7; ++c) xk[c] = x[((long)s * 7 + c) * K + k];
        for (int c = 0; c < 3; ++c) {
            u0[c] = u[((long)s * 3 + c) * K + k];
            u1[c] = u[((long)s * 3 + c) * K + k + 1];
        }
        /* tau = np.linspace(0, 1, K): start + i*step, last point exactly 1 (linearize_discretize.py:356) */
        double step = 1.0 / (K - 1);
        double t0 = k * step, t1 = (k + 1 == K - 1) ? 1.0 : (k + 1) * step;
        int nn = 0;
        double e0[3], e1[3];
        ref_node_input(u + (long)s * 3 * K, K, k, e0);
        ref_node_input(u + (long)s * 3 * K, K, k + 1, e1);
        int st = interval_rk45(xk, u0, u1, e0, e1, tf[s], t0, t1, rtol, atol, max_step, p, p->include_J2, out + i * 105, &nn);
        if (status) status[i] = st;
        if (n_nodes) n_nodes[i] = nn;
        bad += (st != 0);
    }
    return bad;
}

/* controller laws, control.py:20-29,47-53,66-84,104-143 */
static void ctrl_eval(int kind, const double *cp, const double *tab, int Ku, double end_tau, const double *x,
                      double tau, double *u)
{
    u[0] = u[1] = u[2] = 0.0;
    if (kind == ORC_CTRL_CONSTANT) {
        u[0] = cp[0];
        u[1] = cp[1];
        u[2] = cp[2];
    } else if (kind == ORC_CTRL_TANGENTIAL) {
        const double *r = x, *v = x + 3;
        double rn = norm3(r);
        double rh[3] = {r[0] / rn, r[1] / rn, r[2] / rn};
        double h[3] = {r[1] * v[2] - r[2] * v[1], r[2] * v[0] - r[0] * v[2], r[0] * v[1] - r[1] * v[0]};
        double hn = norm3(h);
        double hh[3] = {h[0] / hn, h[1] / hn, h[2] / hn};
        u[0] = cp[0] * (hh[1] * rh[2] - hh[2] * rh[1]);
        u[1] = cp[0] * (hh[2] * rh[0] - hh[0] * rh[2]);
        u[2] = cp[0] * (hh[0] * rh[1] - hh[1] * rh[0]);
    } else if (kind == ORC_CTRL_SEQUENCE) {
        if (tau <= end_tau) {
            double t = tau / end_tau;
            if (t == 1.0) {
                for (int c = 0; c < 3; ++c) u[c] = tab[c * Ku + Ku - 1];
            } else {
                double dt = 1.0 / (Ku - 1);
                int k = (int)floor(t / dt);
                if (k > Ku - 2) k = Ku - 2;
                double lo = (double)k / (Ku - 1), hi = (double)(k + 1) / (Ku - 1);
                double ln = (hi - t) / (hi - lo), lp = (t - lo) / (hi - lo);
                for (int c = 0; c < 3; ++c) u[c] = ln * tab[c * Ku + k] + lp * tab[c * Ku + k + 1];
            }
        }
    }
}

static int prop_rhs(const double *y, double tau, double tf, const orc_params *p, int kind, const double *cp,
                    const double *tab, int Ku, double end_tau, double *dy)
{
    double u[3];
    ctrl_eval(kind, cp, tab, Ku, end_tau, y, tau, u);
    if (dyn(y, u, p, p->include_drag, p->include_J2, dy)) return 1;
    for (int i = 0; i < 7; ++i) dy[i] *= tf;
    return 0;
}

/*
 * Batched propagation.  y0 [N][7], tf [N]; samples at tau_j = j/(T-1) with n_sub RK4 steps between
 * samples; y [N][7][T], u_out [N][3][T] (= extract_uk, linearize_discretize.py:393-411).
 * ctrl_tab: [N][3][Ku] when tab_per_sat, else [3][Ku].
 */
int orc_propagate_rk4(const double *y0, const double *tf, const orc_params *p, int kind, const double *cp,
                      const double *ctrl_tab, int Ku, int tab_per_sat, double end_tau, int N, int T, int n_sub,
                      double *y, double *u_out, int *status, int nthreads)
{
    int bad = 0;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel for schedule(static) reduction(+ : bad)
    for (int s = 0; s < N; ++s) {
        const double *tab = ctrl_tab ? ctrl_tab + (tab_per_sat ? (long)s * 3 * Ku : 0) : 0;
        double x[7], k1[7], k2[7], k3[7], k4[7], xt[7], us[3];
        int st = 0;
        for (int c = 0; c < 7; ++c) x[c] = y0[s * 7 + c];
        double h = (T > 1) ? 1.0 / ((double)(T - 1) * n_sub) : 0.0;
        for (int j = 0; j < T && !st; ++j) {
            double tau_j = (T > 1) ? (double)j / (T - 1) : 0.0;
            for (int c = 0; c < 7; ++c) y[((long)s * 7 + c) * T + j] = x[c];
            if (u_out) {
                ctrl_eval(kind, cp, tab, Ku, end_tau, x, tau_j, us);
                for (int c = 0; c < 3; ++c) u_out[((long)s * 3 + c) * T + j] = us[c];
            }
            if (j == T - 1) break;
            for (int n = 0; n < n_sub; ++n) {
                /* step end points computed from integers so the last one is exactly tau_{j+1}
                 * (and exactly 1.0 at the end of the run, as solve_ivp clips to t_bound) */
                double tau_n = (T > 1) ? (double)(j + 1) / (T - 1) : 0.0;
                double t0 = (n == 0) ? tau_j : tau_j + n * h;
                double t1 = (n == n_sub - 1) ? tau_n : tau_j + (n + 1) * h;
                double tm = 0.5 * (t0 + t1);
                st |= prop_rhs(x, t0, tf[s], p, kind, cp, tab, Ku, end_tau, k1);
                for (int c = 0; c < 7; ++c) xt[c] = x[c] + 0.5 * h * k1[c];
                st |= prop_rhs(xt, tm, tf[s], p, kind, cp, tab, Ku, end_tau, k2);
                for (int c = 0; c < 7; ++c) xt[c] = x[c] + 0.5 * h * k2[c];
                st |= prop_rhs(xt, tm, tf[s], p, kind, cp, tab, Ku, end_tau, k3);
                for (int c = 0; c < 7; ++c) xt[c] = x[c] + h * k3[c];
                st |= prop_rhs(xt, t1, tf[s], p, kind, cp, tab, Ku, end_tau, k4);
                if (st) break;
                for (int c = 0; c < 7; ++c) x[c] += h / 6.0 * (k1[c] + 2.0 * k2[c] + 2.0 * k3[c] + k4[c]);
            }
        }
        if (status) status[s] = st;
        bad += (st != 0);
    }
    return bad;
}

/*
 * Propagation as the reference integrates it (simulator.py:185-187):
 *     solve_ivp(satellite_dynamics, [0, 1], y0, t_eval=linspace(0, 1, T), max_step=0.001)        (RK45, rtol 1e-3, atol 1e-6)
 * i.e. scipy's Dormand-Prince 5(4) with its step-size controller (integrate/_ivp/rk.py RungeKutta._step_impl, rk_step;
 * common.py select_initial_step, norm) and the samples read off the 4th-order dense output of the step that covers them
 * (rk.py RkDenseOutput._call_impl, RK45.P; ivp.py: t_eval points with t_old < t_eval <= t, searchsorted side='right').
 * scipy (1.18.1 in this image) is a third-party dependency the reference neither vendors nor pins; this restates its
 * published algorithm, statement for statement, on the 7-vector.  n_steps / n_rej [N] (may be NULL): accepted / rejected
 * step counts (nfev = 2 + 6 (n_steps + n_rej)).
 */
static const double DP_P[7][4] = {
    {1.0, -8048581381.0 / 2820520608.0, 8663915743.0 / 2820520608.0, -12715105075.0 / 11282082432.0},
    {0.0, 0.0, 0.0, 0.0},
    {0.0, 131558114200.0 / 32700410799.0, -68118460800.0 / 10900136933.0, 87487479700.0 / 32700410799.0},
    {0.0, -1754552775.0 / 470086768.0, 14199869525.0 / 1410260304.0, -10690763975.0 / 1880347072.0},
    {0.0, 127303824393.0 / 49829197408.0, -318862633887.0 / 49829197408.0, 701980252875.0 / 199316789632.0},
    {0.0, -282668133.0 / 205662961.0, 2019193451.0 / 616988883.0, -1453857185.0 / 822651844.0},
    {0.0, 40617522.0 / 29380423.0, -110615467.0 / 29380423.0, 69997945.0 / 29380423.0}};

static double rms7(const double *v)
{
    double s = 0.0;
    for (int i = 0; i < 7; ++i) s += v[i] * v[i];
    return sqrt(s) / sqrt(7.0);
}

int orc_propagate_rk45(const double *y0, const double *tf, const orc_params *p, int kind, const double *cp,
                       const double *ctrl_tab, int Ku, int tab_per_sat, double end_tau, const double *end_tau_arr, int N,
                       int T, double rtol, double atol, double max_step, double *y_out, double *u_out, int *status,
                       int *n_steps, int *n_rej, int nthreads)
{
    int bad = 0;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel for schedule(static) reduction(+ : bad)
    for (int s = 0; s < N; ++s) {
        const double *tab = ctrl_tab ? ctrl_tab + (tab_per_sat ? (long)s * 3 * Ku : 0) : 0;
        const double et = end_tau_arr ? end_tau_arr[s] : end_tau;
        const double tfs = tf[s], t_bound = 1.0;
        double y[7], f[7], K[7][7], ynew[7], fnew[7], tmp[7], sc[7], us[3];
        int st = 0, steps = 0, rej = 0;
        for (int c = 0; c < 7; ++c) y[c] = y0[s * 7 + c];
        double t = 0.0;
        /* np.linspace(0, 1, T): arange(T) * step, last point exactly 1 */
        const double lstep = (T > 1) ? 1.0 / (double)(T - 1) : 0.0;
#define TEVAL(j) (((j) == T - 1 && T > 1) ? 1.0 : (double)(j) * lstep)
        int ti = 0;
        st |= prop_rhs(y, t, tfs, p, kind, cp, tab, Ku, et, f);
        double h_abs = 0.0;
        if (!st) { /* select_initial_step */
            double interval_length = fabs(t_bound - t);
            for (int i = 0; i < 7; ++i) sc[i] = atol + fabs(y[i]) * rtol;
            for (int i = 0; i < 7; ++i) tmp[i] = y[i] / sc[i];
            double d0 = rms7(tmp);
            for (int i = 0; i < 7; ++i) tmp[i] = f[i] / sc[i];
            double d1 = rms7(tmp);
            double h0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : 0.01 * d0 / d1;
            if (h0 > interval_length) h0 = interval_length;
            for (int i = 0; i < 7; ++i) ynew[i] = y[i] + h0 * f[i];
            st |= prop_rhs(ynew, t + h0, tfs, p, kind, cp, tab, Ku, et, fnew);
            for (int i = 0; i < 7; ++i) tmp[i] = (fnew[i] - f[i]) / sc[i];
            double d2 = rms7(tmp) / h0;
            double h1 = (d1 <= 1e-15 && d2 <= 1e-15) ? fmax(1e-6, h0 * 1e-3) : pow(0.01 / fmax(d1, d2), 1.0 / 5.0);
            h_abs = fmin(fmin(100 * h0, h1), fmin(interval_length, max_step));
        }
        while (!st && t < t_bound) {
            double min_step = 10 * fabs(nextafter(t, INFINITY) - t);
            if (h_abs > max_step) h_abs = max_step;
            else if (h_abs < min_step) h_abs = min_step;
            int accepted = 0, rejected = 0;
            double t_new = t, h = 0.0;
            while (!accepted && !st) {
                if (h_abs < min_step) {
                    st = 4;
                    break;
                }
                h = h_abs;
                t_new = t + h;
                if (t_new - t_bound > 0) t_new = t_bound;
                h = t_new - t;
                h_abs = fabs(h);
                memcpy(K[0], f, sizeof f);
                for (int q = 1; q < 6 && !st; ++q) {
                    for (int i = 0; i < 7; ++i) {
                        double dy = 0.0;
                        for (int l = 0; l < q; ++l) dy += K[l][i] * DP_A[q][l];
                        tmp[i] = y[i] + dy * h;
                    }
                    st |= prop_rhs(tmp, t + DP_C[q] * h, tfs, p, kind, cp, tab, Ku, et, K[q]);
                }
                if (st) break;
                for (int i = 0; i < 7; ++i) {
                    double d = 0.0;
                    for (int l = 0; l < 6; ++l) d += K[l][i] * DP_B[l];
                    ynew[i] = y[i] + h * d;
                }
                st |= prop_rhs(ynew, t + h, tfs, p, kind, cp, tab, Ku, et, fnew);
                if (st) break;
                memcpy(K[6], fnew, sizeof fnew);
                for (int i = 0; i < 7; ++i) {
                    double e = 0.0;
                    for (int l = 0; l < 7; ++l) e += K[l][i] * DP_E[l];
                    tmp[i] = e * h / (atol + fmax(fabs(y[i]), fabs(ynew[i])) * rtol);
                }
                double err = rms7(tmp);
                if (err < 1.0) {
                    double factor = (err == 0.0) ? 10.0 : fmin(10.0, 0.9 * pow(err, -0.2));
                    if (rejected && factor > 1.0) factor = 1.0;
                    h_abs *= factor;
                    accepted = 1;
                } else {
                    h_abs *= fmax(0.2, 0.9 * pow(err, -0.2));
                    rejected = 1;
                    ++rej;
                }
            }
            if (st) break;
            ++steps;
            /* samples with t_eval <= t_new, off the dense output of this step */
            if (ti < T && TEVAL(ti) <= t_new) {
                double Q[7][4];
                for (int i = 0; i < 7; ++i)
                    for (int m = 0; m < 4; ++m) {
                        double q = 0.0;
                        for (int l = 0; l < 7; ++l) q += K[l][i] * DP_P[l][m];
                        Q[i][m] = q;
                    }
                while (ti < T && TEVAL(ti) <= t_new) {
                    double te = TEVAL(ti), xx = (te - t) / h, pw[4], ys[7];
                    pw[0] = xx;
                    for (int m = 1; m < 4; ++m) pw[m] = pw[m - 1] * xx;
                    for (int i = 0; i < 7; ++i) {
                        double q = 0.0;
                        for (int m = 0; m < 4; ++m) q += Q[i][m] * pw[m];
                        ys[i] = h * q + y[i];
                        y_out[((long)s * 7 + i) * T + ti] = ys[i];
                    }
                    if (u_out) { /* extract_uk, linearize_discretize.py:393-411 */
                        ctrl_eval(kind, cp, tab, Ku, et, ys, te, us);
                        for (int c = 0; c < 3; ++c) u_out[((long)s * 3 + c) * T + ti] = us[c];
                    }
                    ++ti;
                }
            }
            memcpy(y, ynew, sizeof y);
            memcpy(f, fnew, sizeof f);
            t = t_new;
            if (steps > 50000000) st = 4;
        }
#undef TEVAL
        if (st) { /* the reference raises: nothing is returned for this satellite */
            for (int j = ti; j < T; ++j) {
                for (int i = 0; i < 7; ++i) y_out[((long)s * 7 + i) * T + j] = NAN;
                if (u_out)
                    for (int c = 0; c < 3; ++c) u_out[((long)s * 3 + c) * T + j] = NAN;
            }
        }
        if (status) status[s] = st;
        if (n_steps) n_steps[s] = steps;
        if (n_rej) n_rej[s] = rej;
        bad += (st != 0);
    }
    return bad;
}

int orc_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
