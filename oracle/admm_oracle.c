/*
 * oracle/admm_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU (fp64, OpenMP) restatement of the decentralized TV-ADMM tomography hot path of
 * prsinha1/Distributed-Inverse-Problem-Admm.  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py may load this library.  The product
 * (distributed-inverse-problem-admm_b200/) never links or calls it.
 *
 * PARITY STATUS: "parity unpinned" at the two third-party boundaries of the reference
 * (odl.tomo.RayTransform and cvxpy/SCS are neither vendored under /root/reference nor
 * installable here, and the reference holds no golden vectors).  Everything the reference
 * itself states in NumPy (block_4 TV helpers, block_3 make_precisions, the block_6 z/y/residual
 * loop, the block_2 angle split, the phantoms, psnr) IS pinned: tests/golden/make_golden.py
 * imports those reference modules (third-party imports stubbed) and the fixtures it wrote are
 * checked against this restatement in tests/test_oracle_golden.py.
 *
 * Conventions (reference call sites block_2_load_odl_data.py:23-28,51-54,
 * Gen_Sino_Partitioned.py:126-134):
 *   image  X[ix*N + iy]  <->  point (x,y) = (-1 + (ix+.5)h, -1 + (iy+.5)h), h = 2/N, axis 0 = x
 *   detector bin j centre s_j = -det_w/2 + (j+.5)*ds, ds = det_w/D
 *   sinogram out[a*D + j]  (angle-major, detector fastest)
 *   detector axis u = (cos t, sin t); line integral p(t,s) = int f(s*u + l*(-sin t, cos t)) dl
 *   discretisation = Joseph: step along the dominant axis, linear interpolation on the other,
 *   weight h/|cos| or h/|sin|; tie |cos| == |sin| goes to the sin-dominant branch.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

void orc_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

/* ---- a2: forward projection, literal ray-driven Joseph ------------------- */
void orc_forward(const double* X, int N, int D, double det_w, const double* cs, const double* sn,
                 int nang, double* out) {
    const double h = 2.0 / N, ds = det_w / D, smin = -0.5 * det_w, x0 = -1.0 + 0.5 * h;
#pragma omp parallel for collapse(2) schedule(dynamic, 16) if (N >= 48)
    for (int a = 0; a < nang; ++a) {
        for (int j = 0; j < D; ++j) {
            const double c = cs[a], s = sn[a];
            const double sj = smin + (j + 0.5) * ds;
            double acc = 0.0;
            if (fabs(c) > fabs(s)) {
                const double w = h / fabs(c);
                for (int iy = 0; iy < N; ++iy) {
                    const double y = x0 + iy * h;
                    const double x = (sj - y * s) / c;
                    const double t = (x - x0) / h;
                    const double fl = floor(t);
                    const long i0 = (long)fl;
                    const double f = t - fl;
                    if (i0 >= 0 && i0 < N) acc += (1.0 - f) * w * X[i0 * N + iy];
                    if (i0 + 1 >= 0 && i0 + 1 < N) acc += f * w * X[(i0 + 1) * N + iy];
                }
            } else {
                const double w = h / fabs(s);
                for (int ix = 0; ix < N; ++ix) {
                    const double x = x0 + ix * h;
                    const double y = (sj - x * c) / s;
                    const double t = (y - x0) / h;
                    const double fl = floor(t);
                    const long i0 = (long)fl;
                    const double f = t - fl;
                    if (i0 >= 0 && i0 < N) acc += (1.0 - f) * w * X[(long)ix * N + i0];
                    if (i0 + 1 >= 0 && i0 + 1 < N) acc += f * w * X[(long)ix * N + i0 + 1];
                }
            }
            out[(long)a * D + j] = acc;
        }
    }
}

/* ---- a3 / block_6_admm_loop_ver2.py:145 `Ai.T @ r`: exact transpose as an atomics-free gather.
 * power = 1 -> A^T q ; power = 2 -> column norms^2 with q == NULL (block_3:22). ------------ */
static void adjoint_impl(const double* q, int N, int D, double det_w, const double* cs,
                         const double* sn, int nang, int power, double* out) {
    const double h = 2.0 / N, ds = det_w / D, smin = -0.5 * det_w, x0 = -1.0 + 0.5 * h;
#pragma omp parallel for schedule(dynamic, 1) if (N >= 48)
    for (int ix = 0; ix < N; ++ix) {
        for (int iy = 0; iy < N; ++iy) {
            const double x = x0 + ix * h, y = x0 + iy * h;
            double acc = 0.0;
            for (int a = 0; a < nang; ++a) {
                const double c = cs[a], s = sn[a];
                const double am = (fabs(c) > fabs(s)) ? fabs(c) : fabs(s);
                const double tau = (x * c + y * s - smin) / ds - 0.5;
                const double om = am * h / ds;
                const double wa = h / am;
                long jlo = (long)ceil(tau - om), jhi = (long)floor(tau + om);
                if (jlo < 0) jlo = 0;
                if (jhi > D - 1) jhi = D - 1;
                for (long j = jlo; j <= jhi; ++j) {
                    double wgt = 1.0 - fabs(tau - (double)j) / om;
                    if (wgt <= 0.0) continue;
                    wgt *= wa;
                    acc += (power == 2) ? wgt * wgt : wgt * q[(long)a * D + j];
                }
            }
            out[(long)ix * N + iy] = acc;
        }
    }
}

void orc_adjoint(const double* q, int N, int D, double det_w, const double* cs, const double* sn,
                 int nang, double* out) {
    adjoint_impl(q, N, D, det_w, cs, sn, nang, 1, out);
}

void orc_colnorm2(int N, int D, double det_w, const double* cs, const double* sn, int nang,
                  double* out) {
    adjoint_impl(NULL, N, D, det_w, cs, sn, nang, 2, out);
}

/* ---- a9: forward-difference gradient, block_4_tv_helpers.py:17-23 ------------------------- */
void orc_grad(const double* X, int N, double* gx, double* gy) {
#pragma omp parallel for schedule(static) if (N >= 48)
    for (int r = 0; r < N; ++r)
        for (int c = 0; c < N; ++c) {
            const long k = (long)r * N + c;
            gx[k] = (r < N - 1) ? X[k + N] - X[k] : 0.0;
            gy[k] = (c < N - 1) ? X[k + 1] - X[k] : 0.0;
        }
}

/* exact K^T (the adjoint block_4_tv_helpers.py:25-35 intends) */
void orc_gradT(const double* px, const double* py, int N, double* out) {
#pragma omp parallel for schedule(static) if (N >= 48)
    for (int r = 0; r < N; ++r)
        for (int c = 0; c < N; ++c) {
            const long k = (long)r * N + c;
            double v = 0.0;
            if (r >= 1) v += px[k - N];
            if (r < N - 1) v -= px[k];
            if (c >= 1) v += py[k - 1];
            if (c < N - 1) v -= py[k];
            out[k] = v;
        }
}

/* a10 as shipped: block_4_tv_helpers.py:25-35 (sign-flipped on the first/last row and column) */
void orc_div_reference(const double* px, const double* py, int N, double* out) {
#pragma omp parallel for schedule(static) if (N >= 48)
    for (int r = 0; r < N; ++r)
        for (int c = 0; c < N; ++c) {
            const long k = (long)r * N + c;
            double div = 0.0;
            if (N >= 2) {
                if (r == 0) div -= px[k];
                else if (r == N - 1) div += px[k - N];
                else div += px[k] - px[k - N];
                if (c == 0) div -= py[k];
                else if (c == N - 1) div += py[k - 1];
                else div += py[k] - py[k - 1];
            }
            out[k] = -div;
        }
}

static double dot_n(const double* a, const double* b, long n) {
    double acc = 0.0;
#pragma omp parallel for reduction(+ : acc) schedule(static) if (n >= 2304)
    for (long k = 0; k < n; ++k) acc += a[k] * b[k];
    return acc;
}

/* H v = A^T(prec * A v) + rhoD .* v + mu K^T K v ; also returns A v in Av */
static void apply_H(const double* v, int N, int D, double det_w, const double* cs, const double* sn,
                    int nang, double prec, const double* rhoD_vec, double rhoD_scalar, double mu,
                    double* Av, double* tmp_m, double* Hv) {
    const long n = (long)N * N, m = (long)nang * D;
    orc_forward(v, N, D, det_w, cs, sn, nang, Av);
    for (long k = 0; k < m; ++k) tmp_m[k] = prec * Av[k];
    orc_adjoint(tmp_m, N, D, det_w, cs, sn, nang, Hv);
#pragma omp parallel for schedule(static) if (N >= 48)
    for (int r = 0; r < N; ++r)
        for (int c = 0; c < N; ++c) {
            const long k = (long)r * N + c;
            double lap = 0.0;
            if (r >= 1) lap += v[k] - v[k - N];
            if (r < N - 1) lap += v[k] - v[k + N];
            if (c >= 1) lap += v[k] - v[k - 1];
            if (c < N - 1) lap += v[k] - v[k + 1];
            const double dd = rhoD_vec ? rhoD_vec[k] : rhoD_scalar;
            Hv[k] += dd * v[k] + mu * lap;
        }
    (void)n;
}

/*
 * a13/a14 replacement: the node x-update of eq. (1) (block_5_node_problem.py:21-29,
 * ADMM_Algo.pdf eq. (1)) by S split-Bregman sweeps, each = C warm-started CG iterations on
 *   (A^T P A + rho D + mu K^T K) x = A^T P b + rho sum_j Q_ij (z_ij - y_ij,i) + mu K^T (d - w)
 * followed by d = shrink2(Kx + w, lam/mu), w += Kx - d.   (north_star.)
 *   rhs0  = A^T P b + cons   (assembled by the caller, neighbour order of block_6_ver2:87-95)
 *   x, w (2n: wx then wy), d (2n) in/out (warm start).
 *   Ax_out (m), r_out (n) = final CG recurrence residual of the last sweep, tvrhs_out (n) =
 *   mu K^T(d - w) used in the last sweep's rhs (all needed for the a14 stationarity identity).
 */
void orc_x_update(int N, int D, double det_w, const double* cs, const double* sn, int nang,
                  double prec, const double* rhs0, const double* rhoD_vec, double rhoD_scalar,
                  double mu, double lam, int S, int C, double* x, double* d, double* w,
                  double* Ax_out, double* r_out, double* tvrhs_out) {
    const long n = (long)N * N, m = (long)nang * D;
    double* r = (double*)malloc(sizeof(double) * n);
    double* p = (double*)malloc(sizeof(double) * n);
    double* Hp = (double*)malloc(sizeof(double) * n);
    double* rhs = (double*)malloc(sizeof(double) * n);
    double* t1 = (double*)malloc(sizeof(double) * n);
    double* t2 = (double*)malloc(sizeof(double) * n);
    double* Av = (double*)malloc(sizeof(double) * m);
    double* tm = (double*)malloc(sizeof(double) * m);
    double* wx = w;
    double* wy = w + n;
    double* dx = d;
    double* dy = d + n;
    for (int sw = 0; sw < S; ++sw) {
        /* rhs = rhs0 + mu K^T (d - w) */
#pragma omp parallel for schedule(static) if (n >= 2304)
        for (long k = 0; k < n; ++k) {
            t1[k] = dx[k] - wx[k];
            t2[k] = dy[k] - wy[k];
        }
        orc_gradT(t1, t2, N, rhs);
#pragma omp parallel for schedule(static) if (n >= 2304)
        for (long k = 0; k < n; ++k) {
            rhs[k] = mu * rhs[k];
            if (tvrhs_out) tvrhs_out[k] = rhs[k];
            rhs[k] += rhs0[k];
        }
        /* warm start residual */
        apply_H(x, N, D, det_w, cs, sn, nang, prec, rhoD_vec, rhoD_scalar, mu, Ax_out, tm, Hp);
#pragma omp parallel for schedule(static) if (n >= 2304)
        for (long k = 0; k < n; ++k) {
            r[k] = rhs[k] - Hp[k];
            p[k] = r[k];
        }
        double rr = dot_n(r, r, n);
        for (int it = 0; it < C; ++it) {
            apply_H(p, N, D, det_w, cs, sn, nang, prec, rhoD_vec, rhoD_scalar, mu, Av, tm, Hp);
            const double pHp = dot_n(p, Hp, n);
            const double alpha = (pHp > 0.0) ? rr / pHp : 0.0;
#pragma omp parallel for schedule(static) if (n >= 2304)
            for (long k = 0; k < n; ++k) {
                x[k] += alpha * p[k];
                r[k] -= alpha * Hp[k];
            }
            for (long k = 0; k < m; ++k) Ax_out[k] += alpha * Av[k];
            const double rr_new = dot_n(r, r, n);
            const double beta = (rr > 0.0) ? rr_new / rr : 0.0;
#pragma omp parallel for schedule(static) if (n >= 2304)
            for (long k = 0; k < n; ++k) p[k] = r[k] + beta * p[k];
            rr = rr_new;
        }
        /* d = shrink2(Kx + w, lam/mu); w = Kx + w - d */
        orc_grad(x, N, t1, t2);
        const double kappa = lam / mu;
#pragma omp parallel for schedule(static) if (n >= 2304)
        for (long k = 0; k < n; ++k) {
            const double g1 = t1[k] + wx[k], g2 = t2[k] + wy[k];
            const double nrm = sqrt(g1 * g1 + g2 * g2);
            const double sc = (nrm > kappa) ? (1.0 - kappa / nrm) : 0.0;
            dx[k] = sc * g1;
            dy[k] = sc * g2;
            wx[k] = g1 - dx[k];
            wy[k] = g2 - dy[k];
        }
    }
    if (r_out) memcpy(r_out, r, sizeof(double) * n);
    free(r); free(p); free(Hp); free(rhs); free(t1); free(t2); free(Av); free(tm);
}

/* canonical isotropic TV value (a9 pairing) */
double orc_tv_value(const double* X, int N) {
    double acc = 0.0;
#pragma omp parallel for reduction(+ : acc) schedule(static) if (N >= 48)
    for (int r = 0; r < N; ++r)
        for (int c = 0; c < N; ++c) {
            const long k = (long)r * N + c;
            const double gx = (r < N - 1) ? X[k + N] - X[k] : 0.0;
            const double gy = (c < N - 1) ? X[k + 1] - X[k] : 0.0;
            acc += sqrt(gx * gx + gy * gy);
        }
    return acc;
}

/*
 * a15-a17: one edge of block_6_admm_loop_ver2.py:210-264.
 *   z_new = (a_i + a_j)/2            (W == NULL; :221-223 as shipped)
 *   z_new = (Wi a_i + Wj a_j)/(Wi+Wj) (PDF eq. (2), the commented-out form :221-222)
 *   y_i += x_i - z_new ; y_j += x_j - z_new      (:229-230)
 *   sums[0]=|x_i-z|^2 sums[1]=|x_j-z|^2 sums[2]=|z_new-z_old|^2   (:240-249, without rho^2)
 *   sums[3]=sum q_ij (x_i - (z_old - y_i_old))^2, sums[4] same for j (block_5:24-27 penalty,
 *   without rho/2).  q == NULL means q = qs (scalar).
 */
void orc_edge_update(long n, const double* xi, const double* xj, double* yi, double* yj, double* z,
                     const double* Wi, const double* Wj, const double* qij, const double* qji,
                     double qs, double* sums) {
    double s0 = 0, s1 = 0, s2 = 0, s3 = 0, s4 = 0;
#pragma omp parallel for reduction(+ : s0, s1, s2, s3, s4) schedule(static) if (n >= 2304)
    for (long k = 0; k < n; ++k) {
        const double ai = xi[k] + yi[k], aj = xj[k] + yj[k];
        const double zo = z[k];
        const double ei = xi[k] - (zo - yi[k]), ej = xj[k] - (zo - yj[k]);
        s3 += (qij ? qij[k] : qs) * ei * ei;
        s4 += (qji ? qji[k] : qs) * ej * ej;
        double zn;
        if (Wi) zn = (Wi[k] * ai + Wj[k] * aj) / (Wi[k] + Wj[k]);
        else zn = (ai + aj) / 2.0;
        const double ri = xi[k] - zn, rj = xj[k] - zn, dz = zn - zo;
        yi[k] = yi[k] + xi[k] - zn;
        yj[k] = yj[k] + xj[k] - zn;
        z[k] = zn;
        s0 += ri * ri;
        s1 += rj * rj;
        s2 += dz * dz;
    }
    sums[0] = s0; sums[1] = s1; sums[2] = s2; sums[3] = s3; sums[4] = s4;
}

/* cons += rho * q .* (z - y)   (block_6_admm_loop_ver2.py:93, block_5:24-27 normal equations) */
void orc_accum_cons(long n, double rho, const double* q, double qs, const double* z,
                    const double* y, double* cons) {
#pragma omp parallel for schedule(static) if (n >= 2304)
    for (long k = 0; k < n; ++k) cons[k] += rho * (q ? q[k] : qs) * (z[k] - y[k]);
}

/* ==== (f)-2: "skimage-flavoured" rotate-and-sum projector (Gen_Sino_Partitioned.py:133 pins impl='skimage') ===========
 * skimage.transform.radon(circle=False) rotates the sqrt(2)-padded image bilinearly and sums columns; ODL rescales to
 * physical units.  Restated as ray marching on the rotated pixel grid: for detector bin j (centre s_j) and angle t the
 * samples sit at  r_k = s_j (cos t, sin t) + t_k (-sin t, cos t),  t_k = (k - (P-1)/2) h,  k = 0..P-1,  P = ceil(sqrt2 N),
 * the image is interpolated BILINEARLY there (zero outside), and the bin gets h * sum_k.  SURVEY App. C: this backend is
 * "parity unpinned" (scikit-image / ODL are not installable here); the restatement brackets the discretisation
 * ambiguity next to the Joseph contract.  power = 2: column norms^2. */
static int rs_steps(int N) { return (int)ceil(sqrt(2.0) * N); }

void orc_forward_rs(const double* X, int N, int D, double det_w, const double* cs, const double* sn, int nang,
                    double* out) {
    const double h = 2.0 / N, ds = det_w / D, smin = -0.5 * det_w, x0 = -1.0 + 0.5 * h;
    const int P = rs_steps(N);
#pragma omp parallel for collapse(2) schedule(dynamic, 16) if (N >= 48)
    for (int a = 0; a < nang; ++a) {
        for (int j = 0; j < D; ++j) {
            const double c = cs[a], s = sn[a], sj = smin + (j + 0.5) * ds;
            double acc = 0.0;
            for (int k = 0; k < P; ++k) {
                const double t = (k - 0.5 * (P - 1)) * h;
                const double px = ((sj * c - t * s) - x0) / h, py = ((sj * s + t * c) - x0) / h;
                const double fx0 = floor(px), fy0 = floor(py);
                const long i0 = (long)fx0, j0 = (long)fy0;
                const double fx = px - fx0, fy = py - fy0;
                for (int di = 0; di < 2; ++di)
                    for (int dj = 0; dj < 2; ++dj) {
                        const long ii = i0 + di, jj = j0 + dj;
                        if (ii < 0 || ii >= N || jj < 0 || jj >= N) continue;
                        acc += (di ? fx : 1.0 - fx) * (dj ? fy : 1.0 - fy) * X[ii * N + jj];
                    }
            }
            out[(long)a * D + j] = acc * h;
        }
    }
}

static void adjoint_rs_impl(const double* q, int N, int D, double det_w, const double* cs, const double* sn, int nang,
                            int power, double* out) {
    const double h = 2.0 / N, ds = det_w / D, smin = -0.5 * det_w, x0 = -1.0 + 0.5 * h, r = ds / h;
    const int P = rs_steps(N);
#pragma omp parallel for schedule(dynamic, 1) if (N >= 48)
    for (int ix = 0; ix < N; ++ix)
        for (int iy = 0; iy < N; ++iy) {
            const double x = x0 + ix * h, y = x0 + iy * h;
            double acc = 0.0;
            for (int a = 0; a < nang; ++a) {
                const double c = cs[a], s = sn[a];
                /* (cos, sin) are fp32-rounded, so c^2 + s^2 = 1 + O(1e-7): the EXACT inverse of the forward's sample map
                 * divides by it -- otherwise the gather is the transpose only to 1e-7 */
                const double nrm2 = c * c + s * s;
                const double tau = ((x * c + y * s) / nrm2 - smin) / ds - 0.5;     /* bin coordinate of the pixel centre */
                const double kap = ((-x * s + y * c) / nrm2) / h + 0.5 * (P - 1);  /* step coordinate */
                const double rad = 1.4142135623730951 + 1e-6;
                long jlo = (long)ceil(tau - rad / r), jhi = (long)floor(tau + rad / r);
                long klo = (long)ceil(kap - rad), khi = (long)floor(kap + rad);
                if (jlo < 0) jlo = 0;
                if (jhi > D - 1) jhi = D - 1;
                if (klo < 0) klo = 0;
                if (khi > P - 1) khi = P - 1;
                for (long j = jlo; j <= jhi; ++j) {
                    double wsum = 0.0;      /* A[(a, j), pixel] = h * sum_k (bilinear weight of sample (j, k) on the pixel) */
                    for (long k = klo; k <= khi; ++k) {
                        const double du = (j - tau) * r, dv = (double)k - kap;      /* offsets along u, u_perp in pixels */
                        const double dx = du * c - dv * s, dy = du * s + dv * c;
                        const double wx = 1.0 - fabs(dx), wy = 1.0 - fabs(dy);
                        if (wx <= 0.0 || wy <= 0.0) continue;
                        wsum += wx * wy;
                    }
                    wsum *= h;
                    acc += (power == 2) ? wsum * wsum : wsum * q[(long)a * D + j];
                }
            }
            out[(long)ix * N + iy] = acc;
        }
}

void orc_adjoint_rs(const double* q, int N, int D, double det_w, const double* cs, const double* sn, int nang, double* out) {
    adjoint_rs_impl(q, N, D, det_w, cs, sn, nang, 1, out);
}
void orc_colnorm2_rs(int N, int D, double det_w, const double* cs, const double* sn, int nang, double* out) {
    adjoint_rs_impl(NULL, N, D, det_w, cs, sn, nang, 2, out);
}
