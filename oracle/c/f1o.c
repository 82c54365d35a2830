/*
 * f1o.c -- CPU ORACLE (test infrastructure, NOT product code).  See f1o.h for
 * the parity status of every stage.  float64, no FMA contraction
 * (-ffp-contract=off), OpenMP only in the *_batch / lut_build drivers.
 *
 * Citations are into /root/reference/f1tenth_planning/ unless noted.
 */
#include "f1o.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define PI 3.14159265358979323846

int f1o_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

void f1o_default_config(f1o_config* c) {
    memset(c, 0, sizeof(*c));
    c->n_samples = 100; /* lattice_planner.py:197 */
    c->n_newton = 8;
    c->window = 128;
    c->n_shift = 5;
    c->n_cull = 10;
    c->literal_tracker = 0;
    c->use_goal_kappa = 0;
    c->weights[0] = 0.1;
    c->weights[1] = 0.1;
    c->weights[2] = 0.1;
    c->weights[3] = 0.2;
    c->weights[4] = 0.5;
    c->kappa_max = tan(0.4189) / 0.33; /* kinematic_mpc.py:62-64 */
    c->car_length = 0.58;              /* kinematic_mpc.py:60 */
    c->car_width = 0.31;               /* kinematic_mpc.py:61 */
    c->converge_tol = 1e-4;
    c->tracker_lookahead = 0.8; /* lattice_planner.py:211 */
    c->wheelbase = 0.33;
    c->max_reacquire = 20.0; /* pure_pursuit.py:52 */
}

/* ------------------------------------------------------------------------- */
/* utils/utils.py:37-67 nearest_point                                         */
/* ------------------------------------------------------------------------- */
void f1o_nearest_point(const double* point, const double* traj, int n, int stride,
                       double* proj, double* dist, double* t, int32_t* idx) {
    double best = 0.0, bt = 0.0, bpx = 0.0, bpy = 0.0;
    int bi = -1;
    for (int i = 0; i < n - 1; ++i) {
        const double ax = traj[(size_t)i * stride], ay = traj[(size_t)i * stride + 1];
        const double dx = traj[(size_t)(i + 1) * stride] - ax;     /* :53 diffs */
        const double dy = traj[(size_t)(i + 1) * stride + 1] - ay;
        const double l2 = dx * dx + dy * dy;                       /* :54 */
        const double dot = (point[0] - ax) * dx + (point[1] - ay) * dy; /* :57 */
        double tt = dot / l2;                                      /* :58 */
        if (tt < 0.0) tt = 0.0;                                    /* :59 */
        if (tt > 1.0) tt = 1.0;                                    /* :60 */
        const double px = ax + tt * dx, py = ay + tt * dy;         /* :61 */
        const double ex = point[0] - px, ey = point[1] - py;       /* :64 */
        const double d = sqrt(ex * ex + ey * ey);                  /* :65 */
        /* :66 np.argmin: first minimum; a NaN wins over everything after it */
        const int take = bi < 0 ? 1 : (best != best ? 0 : (d != d || d < best));
        if (take) { best = d; bt = tt; bpx = px; bpy = py; bi = i; }
    }
    proj[0] = bpx; proj[1] = bpy; *dist = best; *t = bt; *idx = bi;
}

/* ------------------------------------------------------------------------- */
/* utils/utils.py:69-151 intersect_point                                      */
/* ------------------------------------------------------------------------- */
static int pymod(int a, int n) { int r = a % n; return r < 0 ? r + n : r; }

int f1o_intersect_point(const double* point, double radius, const double* traj, int n,
                        int stride, double t, int wrap, double* out_p, int32_t* out_i,
                        double* out_t) {
    const int start_i = (int)t;            /* :78 */
    const double start_t = fmod(t, 1.0);   /* :79 (t >= 0) */
    for (int i = start_i; i < n - 1; ++i) { /* :84 */
        const double sx = traj[(size_t)i * stride], sy = traj[(size_t)i * stride + 1];
        const double ex = traj[(size_t)(i + 1) * stride] + 1e-6;      /* :86 */
        const double ey = traj[(size_t)(i + 1) * stride + 1] + 1e-6;
        const double vx = ex - sx, vy = ey - sy;
        const double a = vx * vx + vy * vy;                            /* :89 */
        const double b = 2.0 * (vx * (sx - point[0]) + vy * (sy - point[1])); /* :90 */
        const double c = (sx * sx + sy * sy) + (point[0] * point[0] + point[1] * point[1]) -
                         2.0 * (sx * point[0] + sy * point[1]) - radius * radius; /* :91 */
        double disc = b * b - 4 * a * c;                               /* :92 */
        if (disc < 0) continue;                                        /* :94 */
        disc = sqrt(disc);
        const double t1 = (-b - disc) / (2.0 * a);                     /* :100 */
        const double t2 = (-b + disc) / (2.0 * a);
        if (i == start_i) {                                            /* :102-112 */
            if (t1 >= 0.0 && t1 <= 1.0 && t1 >= start_t) {
                *out_t = t1; *out_i = i; out_p[0] = sx + t1 * vx; out_p[1] = sy + t1 * vy;
                return 1;
            }
            if (t2 >= 0.0 && t2 <= 1.0 && t2 >= start_t) {
                *out_t = t2; *out_i = i; out_p[0] = sx + t2 * vx; out_p[1] = sy + t2 * vy;
                return 1;
            }
        } else if (t1 >= 0.0 && t1 <= 1.0) {                           /* :113-117 */
            *out_t = t1; *out_i = i; out_p[0] = sx + t1 * vx; out_p[1] = sy + t1 * vy;
            return 1;
        } else if (t2 >= 0.0 && t2 <= 1.0) {                           /* :118-122 */
            *out_t = t2; *out_i = i; out_p[0] = sx + t2 * vx; out_p[1] = sy + t2 * vy;
            return 1;
        }
    }
    if (wrap) {                                                        /* :124 */
        for (int i = -1; i < start_i; ++i) {                           /* :125 */
            const int i0 = pymod(i, n), i1 = pymod(i + 1, n);          /* :126-127 */
            const double sx = traj[(size_t)i0 * stride], sy = traj[(size_t)i0 * stride + 1];
            const double ex = traj[(size_t)i1 * stride] + 1e-6;
            const double ey = traj[(size_t)i1 * stride + 1] + 1e-6;
            const double vx = ex - sx, vy = ey - sy;
            const double a = vx * vx + vy * vy;
            const double b = 2.0 * (vx * (sx - point[0]) + vy * (sy - point[1]));
            const double c = (sx * sx + sy * sy) + (point[0] * point[0] + point[1] * point[1]) -
                             2.0 * (sx * point[0] + sy * point[1]) - radius * radius;
            double disc = b * b - 4 * a * c;
            if (disc < 0) continue;
            disc = sqrt(disc);
            const double t1 = (-b - disc) / (2.0 * a);
            const double t2 = (-b + disc) / (2.0 * a);
            if (t1 >= 0.0 && t1 <= 1.0) {                              /* :140-144 */
                *out_t = t1; *out_i = i; out_p[0] = sx + t1 * vx; out_p[1] = sy + t1 * vy;
                return 1;
            } else if (t2 >= 0.0 && t2 <= 1.0) {                       /* :145-149 */
                *out_t = t2; *out_i = i; out_p[0] = sx + t2 * vx; out_p[1] = sy + t2 * vy;
                return 1;
            }
        }
    }
    return 0;
}

/* utils/utils.py:153-161 */
void f1o_get_actuation(double pose_theta, const double* lp, const double* pos,
                       double lookahead_distance, double wheelbase, double* out) {
    const double wy = sin(-pose_theta) * (lp[0] - pos[0]) + cos(-pose_theta) * (lp[1] - pos[1]);
    const double speed = lp[2];
    if (fabs(wy) < 1e-6) { out[0] = speed; out[1] = 0.0; return; }
    const double radius = 1.0 / (2.0 * wy / (lookahead_distance * lookahead_distance));
    out[0] = speed;
    out[1] = atan(wheelbase / radius);
}

/* pure_pursuit.py:56-122 */
int f1o_pure_pursuit(const double* wpts, int n, int ncols, double px, double py, double theta,
                     double lookahead, double wheelbase, double max_reacquire, double* nearest4,
                     int32_t* nearest_i, double* look4, int32_t* look_i, double* act2) {
    const double pos[2] = {px, py};
    double proj[2], dist, t, lp[3], ip[2] = {0, 0}, t2 = 0.0, act[2];
    int32_t i, i2 = 0;
    int status = 0;
    f1o_nearest_point(pos, wpts, n, ncols, proj, &dist, &t, &i);   /* :69 */
    if (nearest4) { nearest4[0] = proj[0]; nearest4[1] = proj[1]; nearest4[2] = dist; nearest4[3] = t; }
    if (nearest_i) *nearest_i = i;
    int found = 0;
    if (dist < lookahead) {                                         /* :70 */
        found = f1o_intersect_point(pos, lookahead, wpts, n, ncols, (double)i + t, 1, ip, &i2, &t2);
        if (found) {                                                /* :78 */
            const int r = pymod(i2, n); /* python negative index */
            lp[0] = wpts[(size_t)r * ncols]; lp[1] = wpts[(size_t)r * ncols + 1];
            lp[2] = wpts[(size_t)i * ncols + 2];
            status = 1;
        }
    } else if (dist < max_reacquire) {                              /* :80-81 */
        lp[0] = wpts[(size_t)i * ncols]; lp[1] = wpts[(size_t)i * ncols + 1];
        lp[2] = wpts[(size_t)i * ncols + 2];
        status = 2;
    }
    if (look4) { look4[0] = ip[0]; look4[1] = ip[1]; look4[2] = t2; look4[3] = found ? 1.0 : 0.0; }
    if (look_i) *look_i = found ? i2 : 0;
    if (!status) {                                                  /* :112-114 */
        if (act2) { act2[0] = 0.0; act2[1] = 0.0; }
        return 0;
    }
    f1o_get_actuation(theta, lp, pos, lookahead, wheelbase, act);
    if (act2) { act2[0] = act[1]; act2[1] = act[0]; }              /* :122 (steer, speed) */
    return status;
}

/* stanley.py:57-112 / lqr.py:60-102: out6 = theta_e, ef, theta_raceline, kappa_ref, goal_velocity,
 * delta */
static double pi_2_pi(double a) { /* utils/utils.py:276-283 */
    if (a > PI) return a - 2.0 * PI;
    if (a < -PI) return a + 2.0 * PI;
    return a;
}

void f1o_front_axle(const double* wpts, int n, int ncols, const double state[4], double wheelbase,
                    double k_path, double* out6, int32_t* target_index) {
    const double fx = state[0] + wheelbase * cos(state[2]);      /* stanley.py:66 */
    const double fy = state[1] + wheelbase * sin(state[2]);      /* :67 */
    const double f[2] = {fx, fy};
    double proj[2], dist, t;
    int32_t i;
    f1o_nearest_point(f, wpts, n, ncols, proj, &dist, &t, &i);   /* :69 */
    const double vx = fx - proj[0], vy = fy - proj[1];           /* :70 */
    const double ef = vx * cos(state[2] - PI / 2.0) + vy * sin(state[2] - PI / 2.0); /* :73-75 */
    const double th_ref = ncols > 3 ? wpts[(size_t)i * ncols + 3] : 0.0;             /* :79 */
    const double th_e = pi_2_pi(th_ref - state[2]);                                   /* :80 */
    out6[0] = th_e; out6[1] = ef; out6[2] = th_ref;
    out6[3] = ncols > 4 ? wpts[(size_t)i * ncols + 4] : 0.0;     /* lqr.py:96 */
    out6[4] = ncols > 2 ? wpts[(size_t)i * ncols + 2] : 0.0;     /* stanley.py:83 */
    out6[5] = atan2(k_path * ef, state[3]) + th_e;               /* :108-110 */
    if (target_index) *target_index = i;
}

void f1o_pure_pursuit_batch(const double* wpts, int n, int ncols, const double* poses, int b,
                            double lookahead, double wheelbase, double max_reacquire,
                            double* nearest4, int32_t* nearest_i, double* look4, int32_t* look_i,
                            double* act2, int32_t* status, int n_threads) {
    if (n_threads < 1) n_threads = 1;
#pragma omp parallel for schedule(static) num_threads(n_threads)
    for (int k = 0; k < b; ++k) {
        int s = f1o_pure_pursuit(wpts, n, ncols, poses[3 * k], poses[3 * k + 1], poses[3 * k + 2],
                                 lookahead, wheelbase, max_reacquire,
                                 nearest4 ? nearest4 + 4 * (size_t)k : 0,
                                 nearest_i ? nearest_i + k : 0, look4 ? look4 + 4 * (size_t)k : 0,
                                 look_i ? look_i + k : 0, act2 ? act2 + 2 * (size_t)k : 0);
        if (status) status[k] = s;
    }
}

/* ------------------------------------------------------------------------- */
/* cubic spiral                                              */
/* ------------------------------------------------------------------------- */
typedef struct { double b1, b2, b3; } cubic_t;

static cubic_t spiral_coeffs(double p0, double p1, double p2, double p3) {
    cubic_t c;
    c.b1 = (-11.0 * p0 + 18.0 * p1 - 9.0 * p2 + 2.0 * p3) / 2.0;
    c.b2 = (18.0 * p0 - 45.0 * p1 + 36.0 * p2 - 9.0 * p3) / 2.0;
    c.b3 = (-9.0 * p0 + 27.0 * p1 - 27.0 * p2 + 9.0 * p3) / 2.0;
    return c;
}
static double spiral_g(double p0, cubic_t c, double u) {
    return u * (p0 + u * (c.b1 / 2.0 + u * (c.b2 / 3.0 + u * (c.b3 / 4.0))));
}
static double spiral_kappa(double p0, cubic_t c, double u) {
    return p0 + u * (c.b1 + u * (c.b2 + u * c.b3));
}

#define F1O_Q 32 /* Simpson intervals of the Newton quadrature */

/* One Newton step q <- q - J^-1 r; returns 0 if the 3x3 system is singular / non-finite. */
static int newton_step(const double goal[3], double p0, double p3, double q[3]) {
    const double p1 = q[0], p2 = q[1], sf = q[2];
    const cubic_t cf = spiral_coeffs(p0, p1, p2, p3);
    double C0 = 0, S0 = 0, Cg = 0, Sg = 0, C1 = 0, S1 = 0, C2 = 0, S2 = 0;
    for (int j = 0; j <= F1O_Q; ++j) {
        const double u = (double)j / F1O_Q;
        const double w = ((j == 0 || j == F1O_Q) ? 1.0 : ((j & 1) ? 4.0 : 2.0)) / (3.0 * F1O_Q);
        const double g = spiral_g(p0, cf, u);
        const double th = sf * g;
        const double c = cos(th), s = sin(th);
        const double u2 = u * u;
        const double d1 = u2 * (4.5 + u * (-7.5 + 3.375 * u));
        const double d2 = u2 * (-2.25 + u * (6.0 - 3.375 * u));
        C0 += w * c; S0 += w * s;
        Cg += w * c * g; Sg += w * s * g;
        C1 += w * c * d1; S1 += w * s * d1;
        C2 += w * c * d2; S2 += w * s * d2;
    }
    const double g1 = (p0 + 3.0 * p1 + 3.0 * p2 + p3) / 8.0;
    const double r0 = sf * C0 - goal[0], r1 = sf * S0 - goal[1], r2 = sf * g1 - goal[2];
    const double sf2 = sf * sf;
    const double J00 = -sf2 * S1, J01 = -sf2 * S2, J02 = C0 - sf * Sg;
    const double J10 = sf2 * C1, J11 = sf2 * C2, J12 = S0 + sf * Cg;
    const double J20 = 0.375 * sf, J21 = 0.375 * sf, J22 = g1;
    const double det = J00 * (J11 * J22 - J12 * J21) - J01 * (J10 * J22 - J12 * J20) +
                       J02 * (J10 * J21 - J11 * J20);
    const double inv = 1.0 / det;
    const double dq0 = (r0 * (J11 * J22 - J12 * J21) - J01 * (r1 * J22 - J12 * r2) +
                        J02 * (r1 * J21 - J11 * r2)) * inv;
    const double dq1 = (J00 * (r1 * J22 - J12 * r2) - r0 * (J10 * J22 - J12 * J20) +
                        J02 * (J10 * r2 - r1 * J20)) * inv;
    const double dq2 = (J00 * (J11 * r2 - r1 * J21) - J01 * (J10 * r2 - r1 * J20) +
                        r0 * (J10 * J21 - J11 * J20)) * inv;
    q[0] = p1 - dq0; q[1] = p2 - dq1; q[2] = sf - dq2;
    return isfinite(q[0]) && isfinite(q[1]) && isfinite(q[2]);
}

void f1o_spiral_solve(const double goal[3], double p0, double p3, int n_newton, double q[3]) {
    for (int it = 0; it < n_newton; ++it) newton_step(goal, p0, p3, q);
}

/* residual of the Q-Simpson endpoint (used by the LUT builder only) */
static double spiral_residual(const double goal[3], double p0, double p3, const double q[3]) {
    const cubic_t cf = spiral_coeffs(p0, q[0], q[1], p3);
    double C0 = 0, S0 = 0;
    for (int j = 0; j <= F1O_Q; ++j) {
        const double u = (double)j / F1O_Q;
        const double w = ((j == 0 || j == F1O_Q) ? 1.0 : ((j & 1) ? 4.0 : 2.0)) / (3.0 * F1O_Q);
        const double th = q[2] * spiral_g(p0, cf, u);
        C0 += w * cos(th); S0 += w * sin(th);
    }
    const double g1 = (p0 + 3.0 * q[0] + 3.0 * q[1] + p3) / 8.0;
    const double r0 = fabs(q[2] * C0 - goal[0]), r1 = fabs(q[2] * S0 - goal[1]),
                 r2 = fabs(q[2] * g1 - goal[2]);
    double r = r0 > r1 ? r0 : r1;
    return r > r2 ? r : r2;
}

void f1o_spiral_sample(const double q[3], double p0, double p3, int m, double* st) {
    const cubic_t cf = spiral_coeffs(p0, q[0], q[1], p3);
    const double sf = q[2];
    const double h = sf / (double)(m - 1);
    double x = 0.0, y = 0.0;
    double th_prev = 0.0;
    st[0] = 0.0; st[1] = 0.0; st[2] = 0.0; st[3] = p0;
    for (int i = 1; i < m; ++i) {
        const double u1 = (double)i / (double)(m - 1);
        const double um = ((double)i - 0.5) / (double)(m - 1);
        const double th1 = sf * spiral_g(p0, cf, u1);
        const double thm = sf * spiral_g(p0, cf, um);
        x += h / 6.0 * (cos(th_prev) + 4.0 * cos(thm) + cos(th1));
        y += h / 6.0 * (sin(th_prev) + 4.0 * sin(thm) + sin(th1));
        st[4 * i] = x; st[4 * i + 1] = y; st[4 * i + 2] = th1;
        st[4 * i + 3] = spiral_kappa(p0, cf, u1);
        th_prev = th1;
    }
}

/* ------------------------------------------------------------------------- */
/* G1 Hermite clothoid (the generator lattice_planner.py:196 calls) */
/* ------------------------------------------------------------------------- */
/* Bertolazzi & Frego, "G1 fitting with clothoids" (2015): with the chord as x axis,
 * phi0 = theta0 - phi, phi1 = theta1 - phi, delta = phi1 - phi0, find A with
 *   g(A) = int_0^1 sin(A t^2 + (delta - A) t + phi0) dt = 0          (1-D Newton from 3(phi0+phi1))
 * then L = r / int_0^1 cos(...), kappa0 = (delta - A)/L, dkappa = 2A/L^2.  pyclothoids (absent
 * here) evaluates these Fresnel-type integrals in closed form, so the oracle integrates them to
 * float64 accuracy -- composite Simpson on 4096 intervals, error ~1e-15 -- and the tests pin the
 * result to the closed-form Fresnel solution (scipy.special.fresnel) at 1e-12.  (The device uses
 * 32 intervals in FP32: ~4e-7 relative, inside the 1e-4 tolerance.) */
#define F1O_QC 4096
static double normalize_angle(double a) {
    while (a > PI) a -= 2.0 * PI;
    while (a <= -PI) a += 2.0 * PI;
    return a;
}

int f1o_clothoid_g1(const double goal[3], int n_newton, double out[3]) {
    const double r = sqrt(goal[0] * goal[0] + goal[1] * goal[1]);
    const double phi = atan2(goal[1], goal[0]);
    const double phi0 = normalize_angle(-phi), phi1 = normalize_angle(goal[2] - phi);
    const double delta = phi1 - phi0;
    double A = 3.0 * (phi0 + phi1);
    double X = 1.0;
    int ok = 0;
    for (int it = 0; it <= n_newton; ++it) {
        double g = 0.0, dg = 0.0;
        X = 0.0;
        for (int j = 0; j <= F1O_QC; ++j) {
            const double t = (double)j / F1O_QC;
            const double w = ((j == 0 || j == F1O_QC) ? 1.0 : ((j & 1) ? 4.0 : 2.0)) / (3.0 * F1O_QC);
            const double ph = A * t * t + (delta - A) * t + phi0;
            const double c = cos(ph), s = sin(ph);
            g += w * s; X += w * c; dg += w * c * (t * t - t);
        }
        if (fabs(g) < 1e-15) { ok = 1; break; }
        if (it == n_newton) { ok = fabs(g) < 1e-12; break; }
        A -= g / dg;
    }
    const double L = r / X;
    out[0] = (delta - A) / L;
    out[1] = 2.0 * A / (L * L);
    out[2] = L;
    return ok && isfinite(L) && L > 0.0;
}

void f1o_clothoid_sample(const double kdl[3], int m, double* st) {
    /* theta(s) = kappa0 s + dkappa s^2 / 2; x, y by Simpson on 32 sub-intervals of every sample
     * interval (error ~1e-16 per interval: float64-exact against the Fresnel closed form) */
    const double L = kdl[2], k0 = kdl[0], dk = kdl[1];
    const double h = L / (double)(m > 1 ? m - 1 : 1);
    const int sub = 32;
    double x = 0.0, y = 0.0;
    for (int i = 0; i < m; ++i) {
        if (i > 0) {
            const double s0 = (i - 1) * h, hs = h / sub;
            double ax = 0.0, ay = 0.0;
            for (int j = 0; j <= sub; ++j) {
                const double sj = s0 + j * hs, th = k0 * sj + 0.5 * dk * sj * sj;
                const double w = (j == 0 || j == sub) ? 1.0 : ((j & 1) ? 4.0 : 2.0);
                ax += w * cos(th); ay += w * sin(th);
            }
            x += ax * hs / 3.0; y += ay * hs / 3.0;
        }
        const double si = i * h;
        st[4 * i] = x; st[4 * i + 1] = y;
        st[4 * i + 2] = k0 * si + 0.5 * dk * si * si;
        st[4 * i + 3] = k0 + dk * si;
    }
}

/* heuristic seed */
static void heuristic_seed(const double goal[3], double q[3]) {
    const double d = sqrt(goal[0] * goal[0] + goal[1] * goal[1]);
    const double th = goal[2];
    q[0] = 0.0; q[1] = 0.0;
    q[2] = d * (th * th / 5.0 + 1.0) + 2.0 * fabs(th) / 5.0;
}

#define LUT_STEPS 16
#define LUT_ITERS_PER_STEP 4
#define LUT_FINAL_ITERS 8

static void lut_cell(const double goal[3], float out[4]) {
    /* continuation: straight line (gx,0,0) -> (gx, gy, gth) */
    double q[3] = {0.0, 0.0, goal[0]};
    int ok = goal[0] > 0.0;
    for (int s = 1; s <= LUT_STEPS && ok; ++s) {
        const double lam = (double)s / LUT_STEPS;
        const double gs[3] = {goal[0], lam * goal[1], lam * goal[2]};
        for (int it = 0; it < LUT_ITERS_PER_STEP && ok; ++it) ok = newton_step(gs, 0.0, 0.0, q);
        if (!(q[2] > 0.0)) ok = 0;
    }
    for (int it = 0; it < LUT_FINAL_ITERS && ok; ++it) ok = newton_step(goal, 0.0, 0.0, q);
    if (ok && q[2] > 0.0 && spiral_residual(goal, 0.0, 0.0, q) < 1e-8) {
        out[0] = (float)q[0]; out[1] = (float)q[1]; out[2] = (float)q[2]; out[3] = 1.0f;
    } else {
        heuristic_seed(goal, q);
        out[0] = (float)q[0]; out[1] = (float)q[1]; out[2] = (float)q[2]; out[3] = 0.0f;
    }
}

static double lut_axis(double lo, double hi, int n, int i) {
    return n > 1 ? lo + (hi - lo) * (double)i / (double)(n - 1) : lo;
}

void f1o_lut_build(const int32_t dims[3], const double r[6], float* lut, int n_threads) {
    const int nx = dims[0], ny = dims[1], nt = dims[2];
    if (n_threads < 1) n_threads = 1;
#pragma omp parallel for schedule(dynamic, 8) num_threads(n_threads)
    for (int cell = 0; cell < nx * ny * nt; ++cell) {
        const int it = cell % nt, iy = (cell / nt) % ny, ix = cell / (nt * ny);
        const double goal[3] = {lut_axis(r[0], r[1], nx, ix), lut_axis(r[2], r[3], ny, iy),
                                lut_axis(r[4], r[5], nt, it)};
        lut_cell(goal, lut + 4 * (size_t)cell);
    }
}

static int lut_index(double v, double lo, double hi, int n) {
    if (n <= 1) return 0;
    const double f = (v - lo) / (hi - lo) * (double)(n - 1);
    int i = (int)floor(f + 0.5);
    if (i < 0) i = 0;
    if (i > n - 1) i = n - 1;
    return i;
}

static void lut_seed(const f1o_world* w, const double goal[3], double q[3]) {
    if (!w->lut) { heuristic_seed(goal, q); return; }
    const int ix = lut_index(goal[0], w->lut_ranges[0], w->lut_ranges[1], w->lut_dims[0]);
    const int iy = lut_index(goal[1], w->lut_ranges[2], w->lut_ranges[3], w->lut_dims[1]);
    const int it = lut_index(goal[2], w->lut_ranges[4], w->lut_ranges[5], w->lut_dims[2]);
    const float* c = w->lut + 4 * (((size_t)ix * w->lut_dims[1] + iy) * w->lut_dims[2] + it);
    q[0] = c[0]; q[1] = c[1]; q[2] = c[2];
}

/* ------------------------------------------------------------------------- */
/* sampler (intent of lattice_planner.py:223-260)                 */
/* ------------------------------------------------------------------------- */
static double wrap_to_pi(double a) {
    a = fmod(a + PI, 2.0 * PI);
    if (a < 0.0) a += 2.0 * PI;
    return a - PI;
}

int f1o_sample_goals(const f1o_world* w, const double pose[4], double* goals, int32_t* centre_i,
                     uint8_t* centre_ok, int32_t* i_ego) {
    const double pos[2] = {pose[0], pose[1]};
    double proj[2], dist, t;
    int32_t i;
    f1o_nearest_point(pos, w->wpts, w->n, w->ncols, proj, &dist, &t, &i); /* :247 */
    *i_ego = i;
    const double ct = cos(pose[2]), st = sin(pose[2]);
    for (int j = 0; j < w->n_lookaheads; ++j) {
        double ip[2], t2;
        int32_t i2 = 0;
        const int found = f1o_intersect_point(pos, w->lookaheads[j], w->wpts, w->n, w->ncols,
                                              (double)i + t, 1, ip, &i2, &t2); /* :250 */
        const int r = found ? pymod(i2, w->n) : 0;
        centre_i[j] = r;
        centre_ok[j] = (uint8_t)found;
        const double cx = w->wpts[(size_t)r * w->ncols], cy = w->wpts[(size_t)r * w->ncols + 1];
        const double psi = w->ncols > 3 ? w->wpts[(size_t)r * w->ncols + 3] : 0.0; /* :251 */
        for (int k = 0; k < w->n_widths; ++k) {
            double* g = goals + 3 * ((size_t)j * w->n_widths + k);
            const double gx = cx - w->widths[k] * sin(psi), gy = cy + w->widths[k] * cos(psi);
            const double dx = gx - pos[0], dy = gy - pos[1];
            g[0] = ct * dx + st * dy;
            g[1] = -st * dx + ct * dy;
            g[2] = wrap_to_pi(psi - pose[2]);
            if (!found) { g[0] = 0.0; g[1] = 0.0; g[2] = 0.0; }
        }
    }
    return w->n_lookaheads * w->n_widths;
}

/* ------------------------------------------------------------------------- */
/* raceline deviation: nearest_point semantics on a cyclic window */
/* ------------------------------------------------------------------------- */
static double window_nearest_dist(const f1o_world* w, int seg0, int nseg, double X, double Y) {
    const int ns = w->n - 1;
    double best = INFINITY;
    for (int q = 0; q < nseg; ++q) {
        const int k = (seg0 + q) % ns;
        const double ax = w->wpts[(size_t)k * w->ncols], ay = w->wpts[(size_t)k * w->ncols + 1];
        const double dx = w->wpts[(size_t)(k + 1) * w->ncols] - ax;
        const double dy = w->wpts[(size_t)(k + 1) * w->ncols + 1] - ay;
        const double l2 = dx * dx + dy * dy;
        double tt = ((X - ax) * dx + (Y - ay) * dy) / l2;
        if (tt < 0.0) tt = 0.0;
        if (tt > 1.0) tt = 1.0;
        const double ex = X - (ax + tt * dx), ey = Y - (ay + tt * dy);
        const double d = sqrt(ex * ex + ey * ey);
        if (d < best) best = d;
    }
    return best;
}

/* ------------------------------------------------------------------------- */
/* collision                                                     */
/* ------------------------------------------------------------------------- */
/* SAT separation of two equal rectangles (half extents hl, hw); > 0 separated by that much,
 * <= 0 overlapping (touching counts as separated: collide iff sep < 0). */
static double sat_separation(double tx, double ty, double th_a, double th_b, double hl, double hw) {
    const double ca = cos(th_a), sa = sin(th_a), cb = cos(th_b), sb = sin(th_b);
    const double c = fabs(ca * cb + sa * sb), s = fabs(sa * cb - ca * sb);
    const double e0 = fabs(tx * ca + ty * sa) - (hl + hl * c + hw * s);
    const double e1 = fabs(-tx * sa + ty * ca) - (hw + hl * s + hw * c);
    const double e2 = fabs(tx * cb + ty * sb) - (hl + hl * c + hw * s);
    const double e3 = fabs(-tx * sb + ty * cb) - (hw + hl * s + hw * c);
    double m = e0 > e1 ? e0 : e1;
    if (e2 > m) m = e2;
    if (e3 > m) m = e3;
    return m;
}

static const double PROBE[9][2] = {{1, 1}, {1, -1}, {-1, 1}, {-1, -1}, {1, 0},
                                   {-1, 0}, {0, 1},  {0, -1},  {0, 0}};

static int grid_occ(const f1o_world* w, long col, long row) {
    if (col < 0 || row < 0 || col >= w->gw || row >= w->gh) return 1;
    return w->grid[(size_t)row * w->gw + col] != 0;
}

/* returns hit; *margin = distance (m) from the probe to the nearest cell of different occupancy */
static int grid_probe(const f1o_world* w, double X, double Y, double* margin) {
    const double fx = (X - w->gox) / w->gres, fy = (Y - w->goy) / w->gres;
    const long col = (long)floor(fx), row = (long)floor(fy);
    const int occ = grid_occ(w, col, row);
    if (margin) {
        const double ax = fx - (double)col, ay = fy - (double)row;
        double best = INFINITY;
        for (int dy = -1; dy <= 1; ++dy)
            for (int dx = -1; dx <= 1; ++dx) {
                if (!dx && !dy) continue;
                if (grid_occ(w, col + dx, row + dy) == occ) continue;
                const double ex = dx < 0 ? ax : (dx > 0 ? 1.0 - ax : 0.0);
                const double ey = dy < 0 ? ay : (dy > 0 ? 1.0 - ay : 0.0);
                const double d = sqrt(ex * ex + ey * ey) * w->gres;
                if (d < best) best = d;
            }
        *margin = best;
    }
    return occ;
}

/* collision_mode 1: does the disc centred at (X, Y) collide?  status of a cell = out of bounds or
 * edt2 < t2; *margin = distance (m) from the centre to the nearest neighbouring cell of different
 * status (the only place where the FP32 device path may legitimately decide differently) */
static int edt_status(const f1o_world* w, long col, long row, int t2) {
    if (col < 0 || row < 0 || col >= w->gw || row >= w->gh) return 1;
    return (int)w->edt2[(size_t)row * w->gw + col] < t2;
}

static int edt_probe(const f1o_world* w, double X, double Y, int t2, double* margin) {
    const double fx = (X - w->gox) / w->gres, fy = (Y - w->goy) / w->gres;
    const long col = (long)floor(fx), row = (long)floor(fy);
    const int st = edt_status(w, col, row, t2);
    if (margin) {
        const double ax = fx - (double)col, ay = fy - (double)row;
        double best = INFINITY;
        for (int dy = -1; dy <= 1; ++dy)
            for (int dx = -1; dx <= 1; ++dx) {
                if (!dx && !dy) continue;
                if (edt_status(w, col + dx, row + dy, t2) == st) continue;
                const double ex = dx < 0 ? ax : (dx > 0 ? 1.0 - ax : 0.0);
                const double ey = dy < 0 ? ay : (dy > 0 ? 1.0 - ay : 0.0);
                const double d = sqrt(ex * ex + ey * ey) * w->gres;
                if (d < best) best = d;
            }
        *margin = best;
    }
    return st;
}

/* ------------------------------------------------------------------------- */
/* one query                                                                  */
/* ------------------------------------------------------------------------- */
int f1o_plan(const f1o_config* cfg, const f1o_world* w, const double pose[4], const double* opp,
             int n_opp, const double* goals_in, int n_goals_in, int c_begin, int c_end,
             f1o_result* out) {
    const int M = cfg->n_samples;
    const int C = goals_in ? n_goals_in : w->n_lookaheads * w->n_widths;
    if (c_begin < 0) c_begin = 0;
    if (c_end <= 0 || c_end > C) c_end = C;
    double* goals = (double*)malloc(sizeof(double) * 3 * (size_t)C);
    int32_t* centre_i = (int32_t*)calloc((size_t)(w->n_lookaheads > 0 ? w->n_lookaheads : 1), 4);
    uint8_t* centre_ok = (uint8_t*)calloc((size_t)(w->n_lookaheads > 0 ? w->n_lookaheads : 1), 1);
    double* st = (double*)malloc(sizeof(double) * 4 * (size_t)M);
    double* best_st = (double*)malloc(sizeof(double) * 4 * (size_t)M);
    int32_t i_ego = 0;
    const double pos[2] = {pose[0], pose[1]};
    if (goals_in) {
        memcpy(goals, goals_in, sizeof(double) * 3 * (size_t)C);
        double proj[2], dist, t;
        f1o_nearest_point(pos, w->wpts, w->n, w->ncols, proj, &dist, &t, &i_ego);
    } else {
        f1o_sample_goals(w, pose, goals, centre_i, centre_ok, &i_ego);
    }
    /* raceline window (B.5) */
    const int ns = w->n - 1;
    int nseg = cfg->window;
    if (nseg <= 0 || nseg > ns) nseg = ns;
    const int seg0 = pymod(i_ego - nseg / 4, ns);

    const double ct = cos(pose[2]), stn = sin(pose[2]);
    const double hl = 0.5 * cfg->car_length, hw = 0.5 * cfg->car_width;
    const double rc2 = 4.0 * (hl * hl + hw * hw); /* (2 r_circ)^2 */
    double best_cost = INFINITY;
    int best_idx = -1;
    int best_row = 0;

    for (int c = c_begin; c < c_end; ++c) {
        const double* g = goals + 3 * (size_t)c;
        const int row = goals_in ? 0 : c / w->n_widths;
        unsigned flags = 0;
        const int have_centre = goals_in ? 1 : centre_ok[row];
        if (!have_centre) flags |= F1O_FLAG_NO_CENTRE;
        double p3 = 0.0;
        if (cfg->use_goal_kappa && !goals_in && w->ncols > 4)
            p3 = w->wpts[(size_t)centre_i[row] * w->ncols + 4];
        double q[3];
        if (cfg->generator == 1) {
            double kdl[3];
            f1o_clothoid_g1(g, cfg->n_newton, kdl);
            f1o_clothoid_sample(kdl, M, st);
            q[0] = kdl[0]; q[1] = kdl[1]; q[2] = kdl[2];   /* params = (kappa0, dkappa, L) */
            p3 = kdl[0] + kdl[1] * kdl[2];
        } else {
            lut_seed(w, g, q);
            f1o_spiral_solve(g, 0.0, p3, cfg->n_newton, q);
            f1o_spiral_sample(q, 0.0, p3, M, st);
        }
        /* validity (B.2) */
        const double gn = sqrt(g[0] * g[0] + g[1] * g[1] + g[2] * g[2]);
        const double tol = cfg->converge_tol * (gn > 1.0 ? gn : 1.0);
        const double* e = st + 4 * (size_t)(M - 1);
        double maxk = 0.0, sumk = 0.0;
        for (int i = 0; i < M; ++i) {
            const double ak = fabs(st[4 * i + 3]);
            if (ak > maxk) maxk = ak;
            sumk += ak;
        }
        int valid = have_centre && isfinite(q[0]) && isfinite(q[1]) && isfinite(q[2]) &&
                    q[2] > 0.0 && fabs(e[0] - g[0]) < tol && fabs(e[1] - g[1]) < tol &&
                    fabs(e[2] - g[2]) < tol;
        if (valid && cfg->kappa_max > 0.0 && !(maxk <= cfg->kappa_max)) valid = 0;
        if (valid) flags |= F1O_FLAG_VALID;
        /* costs (B.4) */
        double terms[F1O_N_TERMS] = {0, 0, 0, 0, 0};
        double m_opp = INFINITY, m_map = INFINITY;
        if (valid) {
            terms[0] = 1.0 / q[2];          /* lattice_planner.py:271 */
            terms[1] = maxk;                /* :277 */
            terms[2] = sumk / (double)M;    /* :284 */
            if (w->prev_theta) {            /* :290-295 */
                const int lim = M - cfg->n_shift - cfg->n_cull;
                double acc = 0.0;
                for (int i = 0; i < lim; ++i) {
                    const double d = st[4 * i + 2] - (double)w->prev_theta[i + cfg->n_shift];
                    acc += d * d;
                }
                terms[3] = acc;
            }
            double dev = 0.0;
            int hit_opp = 0, hit_map = 0;
            for (int i = 0; i < M; ++i) {
                const double x = st[4 * i], y = st[4 * i + 1], th = st[4 * i + 2];
                const double X = pose[0] + ct * x - stn * y, Y = pose[1] + stn * x + ct * y;
                const double TH = th + pose[2];
                dev += window_nearest_dist(w, seg0, nseg, X, Y);
                for (int k = 0; k < n_opp; ++k) {
                    const double tx = opp[3 * k] - X, ty = opp[3 * k + 1] - Y;
                    const double sep = sat_separation(tx, ty, TH, opp[3 * k + 2], hl, hw);
                    if (fabs(sep) < m_opp) m_opp = fabs(sep);
                    if (tx * tx + ty * ty > rc2) continue; /* broad phase */
                    if (sep < 0.0) hit_opp = 1;
                }
                if (w->grid && cfg->collision_mode == 1 && w->edt2) {
                    /* three covering discs on the body axis at -L/3, 0, +L/3 (radius
                     * sqrt((L/6)^2 + (W/2)^2)), one distance-transform lookup each: a disc
                     * collides iff the squared cell distance at its centre's cell is at most
                     * (radius / res + sqrt 2)^2; out of bounds = collision */
                    const double cth = cos(TH), sth = sin(TH);
                    const double rc = sqrt(cfg->car_length * cfg->car_length / 36.0 +
                                           cfg->car_width * cfg->car_width / 4.0) / w->gres;
                    /* cell-centre distances: + half a cell diagonal for the disc centre within its
                     * cell, + half for the extent of the occupied cell: strictly conservative */
                    const double tc = rc + 1.4142135623730951;
                    const int t2 = (int)floor(tc * tc) + 1;
                    for (int d = -1; d <= 1; ++d) {
                        const double off = d * cfg->car_length / 3.0;
                        double mg;
                        if (edt_probe(w, X + cth * off, Y + sth * off, t2, &mg)) hit_map = 1;
                        if (mg < m_map) m_map = mg;
                    }
                } else if (w->grid) {
                    const double cth = cos(TH), sth = sin(TH);
                    for (int p = 0; p < 9; ++p) {
                        const double bx = PROBE[p][0] * hl, by = PROBE[p][1] * hw;
                        double mg;
                        if (grid_probe(w, X + cth * bx - sth * by, Y + sth * bx + cth * by, &mg))
                            hit_map = 1;
                        if (mg < m_map) m_map = mg;
                    }
                }
            }
            terms[4] = dev / (double)M;
            if (hit_opp) flags |= F1O_FLAG_COLLIDE_OPP;
            if (hit_map) flags |= F1O_FLAG_COLLIDE_MAP;
        }
        double cost = INFINITY;
        if (valid && !(flags & (F1O_FLAG_COLLIDE_OPP | F1O_FLAG_COLLIDE_MAP))) {
            cost = 0.0;
            for (int j = 0; j < F1O_N_TERMS; ++j) cost += cfg->weights[j] * terms[j];
            if (!isfinite(cost)) cost = INFINITY;
        }
        if (out->costs) out->costs[c] = cost;
        if (out->terms) memcpy(out->terms + F1O_N_TERMS * (size_t)c, terms, sizeof(terms));
        if (out->flags) out->flags[c] = (uint8_t)flags;
        if (out->goals) memcpy(out->goals + 3 * (size_t)c, g, sizeof(double) * 3);
        if (out->params) {
            double* p = out->params + 4 * (size_t)c;
            p[0] = q[0]; p[1] = q[1]; p[2] = q[2]; p[3] = p3;
        }
        if (out->states) memcpy(out->states + 4 * (size_t)M * c, st, sizeof(double) * 4 * (size_t)M);
        if (out->margins) { out->margins[2 * (size_t)c] = m_opp; out->margins[2 * (size_t)c + 1] = m_map; }
        if (cost < best_cost) { /* first minimum: lattice_planner.py:169-171 */
            best_cost = cost; best_idx = c; best_row = row;
            memcpy(best_st, st, sizeof(double) * 4 * (size_t)M);
        }
    }
    out->n_candidates = C;
    out->no_feasible = best_idx < 0;
    if (best_idx < 0) {
        /* all +inf: np.argmin -> first index of the evaluated range; regenerate it */
        best_idx = c_begin;
        best_row = goals_in ? 0 : best_idx / w->n_widths;
        const double* g = goals + 3 * (size_t)best_idx;
        double p3 = 0.0, q[3];
        if (cfg->use_goal_kappa && !goals_in && w->ncols > 4)
            p3 = w->wpts[(size_t)centre_i[best_row] * w->ncols + 4];
        if (cfg->generator == 1) {
            double kdl[3];
            f1o_clothoid_g1(g, cfg->n_newton, kdl);
            f1o_clothoid_sample(kdl, M, best_st);
        } else {
            lut_seed(w, g, q);
            f1o_spiral_solve(g, 0.0, p3, cfg->n_newton, q);
            f1o_spiral_sample(q, 0.0, p3, M, best_st);
        }
    }
    out->best_idx = best_idx;
    out->best_cost = best_cost;
    if (out->best_traj) memcpy(out->best_traj, best_st, sizeof(double) * 4 * (size_t)M);
    /* tracker (B.8; lattice_planner.py:208-212) */
    {
        double act[2];
        int status;
        if (cfg->literal_tracker) {
            /* map-frame pose against the vehicle-frame trajectory, col 2 (theta) as speed */
            status = f1o_pure_pursuit(best_st, M, 4, pose[0], pose[1], pose[2],
                                      cfg->tracker_lookahead, 0.33 /* :55 default wheelbase */,
                                      cfg->max_reacquire, 0, 0, 0, 0, act);
        } else {
            /* vehicle frame: pose (0,0,0); speed = raceline speed at the goal centre */
            double* wp = (double*)malloc(sizeof(double) * 3 * (size_t)M);
            double v = pose[3];
            if (!goals_in && w->ncols > 2) v = w->wpts[(size_t)centre_i[best_row] * w->ncols + 2];
            for (int i = 0; i < M; ++i) {
                wp[3 * i] = best_st[4 * i]; wp[3 * i + 1] = best_st[4 * i + 1]; wp[3 * i + 2] = v;
            }
            status = f1o_pure_pursuit(wp, M, 3, 0.0, 0.0, 0.0, cfg->tracker_lookahead,
                                      cfg->wheelbase, cfg->max_reacquire, 0, 0, 0, 0, act);
            free(wp);
        }
        out->tracker_found = status != 0;
        out->steer = act[0];
        out->speed = act[1];
    }
    if (out->best_traj_map) {
        /* the best trajectory in the map frame with a speed column [X, Y, v, Theta],
         * X = pose + R(theta_pose) (x, y) (B.3), v = raceline speed at the goal centre */
        double v = pose[3];
        if (!goals_in && w->ncols > 2) v = w->wpts[(size_t)centre_i[best_row] * w->ncols + 2];
        for (int i = 0; i < M; ++i) {
            const double x = best_st[4 * i], y = best_st[4 * i + 1];
            out->best_traj_map[4 * i] = pose[0] + (ct * x - stn * y);
            out->best_traj_map[4 * i + 1] = pose[1] + (stn * x + ct * y);
            out->best_traj_map[4 * i + 2] = v;
            out->best_traj_map[4 * i + 3] = best_st[4 * i + 2] + pose[2];
        }
    }
    free(goals); free(centre_i); free(centre_ok); free(st); free(best_st);
    return C;
}

int64_t f1o_plan_batch(const f1o_config* cfg, const f1o_world* w, const double* poses,
                       const double* opp, const int32_t* n_opp, int s, int max_opp,
                       int32_t* best_idx, double* best_cost, double* best_traj, double* costs,
                       uint8_t* flags, double* steer_speed, int n_threads) {
    const int C = w->n_lookaheads * w->n_widths;
    const int M = cfg->n_samples;
    if (n_threads < 1) n_threads = 1;
#pragma omp parallel for schedule(dynamic, 1) num_threads(n_threads)
    for (int k = 0; k < s; ++k) {
        f1o_result r;
        memset(&r, 0, sizeof(r));
        if (best_traj) r.best_traj = best_traj + 4 * (size_t)M * k;
        if (costs) r.costs = costs + (size_t)C * k;
        if (flags) r.flags = flags + (size_t)C * k;
        const int ko = n_opp ? n_opp[k] : max_opp;
        f1o_plan(cfg, w, poses + 4 * (size_t)k, opp ? opp + 3 * (size_t)max_opp * k : 0, ko, 0, 0, 0,
                 0, &r);
        if (best_idx) best_idx[k] = r.best_idx;
        if (best_cost) best_cost[k] = r.best_cost;
        if (steer_speed) { steer_speed[2 * k] = r.steer; steer_speed[2 * k + 1] = r.speed; }
    }
    return (int64_t)s * C;
}

/* ------------------------------------------------------------------------- */
/* float32 mirror of the device collision predicate (teacher-forced checks)    */
/* ------------------------------------------------------------------------- */
/* Same operations in the same order and rounding as eval_kernel's collision block
 * (f1tenth_planning_b200/csrc/f1l_lattice.cuh: sat_collide / grid_hit and the explicitly
 * rounded fm/fa/fs arithmetic and the explicitly fused sample -> grid-cell transform); compiled
 * with -ffp-contract=off so that no other FMA is formed.
 * Inputs are the device's own float32 states, footprint headings and per-query constants, so
 * the flags must agree bit for bit. */
static int sat_collide_f32(float tx, float ty, float c, float s, float oc, float os, float hl,
                           float hw) {
    const float cc = fabsf(c * oc + s * os);
    const float ss = fabsf(s * oc - c * os);
    const float rl = hl + (hl * cc + hw * ss);
    const float rw = hw + (hl * ss + hw * cc);
    const float e0 = fabsf(tx * c + ty * s);
    const float e1 = fabsf(ty * c - tx * s);
    const float e2 = fabsf(tx * oc + ty * os);
    const float e3 = fabsf(ty * oc - tx * os);
    return e0 < rl && e1 < rw && e2 < rl && e3 < rw;
}

static int grid_hit_f32(const uint8_t* occ, int gw, int gh, int ix0, int iy0, float cx, float cy) {
    const int col = ix0 + (int)floorf(cx), row = iy0 + (int)floorf(cy);
    if (col < 0 || row < 0 || col >= gw || row >= gh) return 1;
    return occ[(size_t)row * gw + col] != 0;
}

void f1o_collide_f32(const float* states, const float* headings, int c, int m,
                     const float* opp_local, int n_opp, const float* grid_xf,
                     const int32_t* grid_i0, const uint8_t* grid, int gh, int gw, float half_l,
                     float half_w, float rc2, uint8_t* flags_out) {
    const float hl = half_l, hw = half_w;
    for (int k = 0; k < c; ++k) {
        int hit_opp = 0, hit_map = 0;
        for (int i = 0; i < m; ++i) {
            const float x = states[4 * ((size_t)k * m + i)], y = states[4 * ((size_t)k * m + i) + 1];
            const float cs = headings[2 * ((size_t)k * m + i)], sn = headings[2 * ((size_t)k * m + i) + 1];
            for (int o = 0; o < n_opp; ++o) {
                const float tx = opp_local[4 * o] - x, ty = opp_local[4 * o + 1] - y;
                const float d2 = tx * tx + ty * ty;
                if (d2 <= rc2 && sat_collide_f32(tx, ty, cs, sn, opp_local[4 * o + 2],
                                                 opp_local[4 * o + 3], hl, hw))
                    hit_opp = 1;
            }
            if (grid) {
                const float A00 = grid_xf[0], A01 = grid_xf[1], A10 = grid_xf[2], A11 = grid_xf[3];
                const float ccx = fmaf(A00, x, fmaf(A01, y, grid_xf[4]));   /* fused like the device */
                const float ccy = fmaf(A10, x, fmaf(A11, y, grid_xf[5]));
                const float lx = cs * hl, ly = sn * hl;
                const float wx = -sn * hw, wy = cs * hw;
                const float elx = A00 * lx + A01 * ly, ely = A10 * lx + A11 * ly;
                const float ewx = A00 * wx + A01 * wy, ewy = A10 * wx + A11 * wy;
                const int ix0 = grid_i0[0], iy0 = grid_i0[1];
                int h = 0;
                h |= grid_hit_f32(grid, gw, gh, ix0, iy0, (ccx + elx) + ewx, (ccy + ely) + ewy);
                h |= grid_hit_f32(grid, gw, gh, ix0, iy0, (ccx + elx) - ewx, (ccy + ely) - ewy);
                h |= grid_hit_f32(grid, gw, gh, ix0, iy0, (ccx - elx) + ewx, (ccy - ely) + ewy);
                h |= grid_hit_f32(grid, gw, gh, ix0, iy0, (ccx - elx) - ewx, (ccy - ely) - ewy);
                h |= grid_hit_f32(grid, gw, gh, ix0, iy0, ccx + elx, ccy + ely);
                h |= grid_hit_f32(grid, gw, gh, ix0, iy0, ccx - elx, ccy - ely);
                h |= grid_hit_f32(grid, gw, gh, ix0, iy0, ccx + ewx, ccy + ewy);
                h |= grid_hit_f32(grid, gw, gh, ix0, iy0, ccx - ewx, ccy - ewy);
                h |= grid_hit_f32(grid, gw, gh, ix0, iy0, ccx, ccy);
                hit_map |= h;
            }
        }
        flags_out[k] = (uint8_t)((hit_opp ? F1O_FLAG_COLLIDE_OPP : 0) | (hit_map ? F1O_FLAG_COLLIDE_MAP : 0));
    }
}
