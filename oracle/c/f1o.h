/*
 * f1o.h -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * Plain-C float64 restatement of the f1tenth_planning lattice-planner hot path.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load this library; the product (f1tenth_planning_b200)
 * never does.
 *
 * Parity status per stage:
 *   nearest_point / intersect_point / get_actuation / pure pursuit: PINNED --
 *     checked bit-for-bit (index) / 1e-12 (values) against the imported
 *     reference functions (tests/golden/ npz files, made by
 *     tests/golden/make_golden.py from /root/reference).
 *   sampler: intent pinned (reference code raises IndexError), restated per
 *     SURVEY B.1.
 *   cubic-spiral generation, raceline deviation, collision: PARITY UNPINNED --
 *     the reference has no such code (it calls pyclothoids==0.1.4
 *     Clothoid.G1Hermite, absent here, and a map_collision stub); this file
 *     is the definition.
 *   cost terms: formulas pinned by lattice_planner.py:268-296; constants
 *     N_SHIFT/N_CULL are undefined in the reference (defaults 5/10 here).
 *   argmin: pinned (np.argmin, first minimum).
 */
#ifndef F1O_H
#define F1O_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define F1O_N_TERMS 5
#define F1O_FLAG_VALID 1u
#define F1O_FLAG_COLLIDE_OPP 2u
#define F1O_FLAG_COLLIDE_MAP 4u
#define F1O_FLAG_NO_CENTRE 8u

typedef struct {
    int32_t n_samples, n_newton, window, n_shift, n_cull, literal_tracker, use_goal_kappa, generator;
    int32_t prune_window; /* layout parity with f1l_config only: the oracle always scans the whole window */
    int32_t collision_mode; /* 0: nine probes; 1: three discs on the Euclidean distance transform (world.edt2) */
    double weights[F1O_N_TERMS];
    double kappa_max, car_length, car_width, converge_tol, tracker_lookahead, wheelbase,
        max_reacquire;
} f1o_config;

typedef struct {
    const double* wpts; /* [n, ncols] */
    int32_t n, ncols;
    const uint8_t* grid; /* [gh, gw] or NULL */
    int32_t gh, gw;
    double gox, goy, gres;
    const float* lut; /* [nx, ny, nt, 4] */
    int32_t lut_dims[3];
    int32_t pad0;
    double lut_ranges[6];
    const double* lookaheads;
    const double* widths;
    int32_t n_lookaheads, n_widths;
    const float* prev_theta; /* [M] or NULL */
    const uint16_t* edt2;    /* [gh, gw] squared cell distance to the nearest occupied / out-of-bounds
                                cell (exact to 576, 577 = farther), collision_mode 1; or NULL */
} f1o_world;

typedef struct {
    double steer, speed;
    int32_t best_idx, no_feasible, tracker_found, n_candidates;
    double best_cost;
    double* best_traj; /* [M,4] */
    double* costs;     /* [C] */
    double* terms;     /* [C,5] */
    uint8_t* flags;    /* [C] */
    double* goals;     /* [C,3] */
    double* params;    /* [C,4] */
    double* states;    /* [C,M,4] */
    double* margins;   /* [C,2] opponent / map collision margins (metres, see f1o.c) */
    double* best_traj_map; /* [M,4] (X, Y, v, Theta) map frame, SURVEY B.8 */
} f1o_result;

void f1o_default_config(f1o_config* cfg);

/* utils/utils.py:37-67 */
void f1o_nearest_point(const double* point, const double* traj, int n, int stride,
                       double* proj, double* dist, double* t, int32_t* idx);
/* utils/utils.py:69-151; returns 1 if found */
int f1o_intersect_point(const double* point, double radius, const double* traj, int n,
                        int stride, double t, int wrap, double* out_p, int32_t* out_i,
                        double* out_t);
/* utils/utils.py:153-161; out = (speed, steer) */
void f1o_get_actuation(double pose_theta, const double* lookahead_point, const double* position,
                       double lookahead_distance, double wheelbase, double* out);
/* pure_pursuit.py:56-122; returns status 1 intersect, 2 reacquire, 0 none */
int f1o_pure_pursuit(const double* wpts, int n, int ncols, double px, double py, double theta,
                     double lookahead, double wheelbase, double max_reacquire, double* nearest4,
                     int32_t* nearest_i, double* look4, int32_t* look_i, double* act2);
void f1o_pure_pursuit_batch(const double* wpts, int n, int ncols, const double* poses, int b,
                            double lookahead, double wheelbase, double max_reacquire,
                            double* nearest4, int32_t* nearest_i, double* look4, int32_t* look_i,
                            double* act2, int32_t* status, int n_threads);

/* stanley.py:57-112, lqr.py:60-102: front-axle errors; out6 = theta_e, ef, theta_raceline,
 * kappa_ref, goal_velocity, delta */
void f1o_front_axle(const double* wpts, int n, int ncols, const double state[4], double wheelbase,
                    double k_path, double* out6, int32_t* target_index);

/* LUT build (continuation from the straight line), [nx,ny,nt,4] float */
void f1o_lut_build(const int32_t dims[3], const double ranges[6], float* lut, int n_threads);

/* cubic spiral: Newton from a seed, then M-sample Simpson integration.
 * q = (p1, p2, s_f) in/out; states [M,4] = x, y, theta, kappa(signed). */
void f1o_spiral_solve(const double goal[3], double p0, double p3, int n_newton, double q[3]);
void f1o_spiral_sample(const double q[3], double p0, double p3, int m, double* states);
/* G1 Hermite clothoid (0,0,0) -> goal (Bertolazzi & Frego 2015; what pyclothoids'
 * Clothoid.G1Hermite solves): out = (kappa0, dkappa, L); returns 1 if converged */
int f1o_clothoid_g1(const double goal[3], int n_newton, double out[3]);
void f1o_clothoid_sample(const double kdl[3], int m, double* states);

/* sampler B.1: goals [C,3] vehicle frame, centre waypoint index per lookahead row,
 * nearest index i_ego; returns C */
int f1o_sample_goals(const f1o_world* w, const double pose[4], double* goals, int32_t* centre_i,
                     uint8_t* centre_ok, int32_t* i_ego);

/* full plan: goals_in NULL -> built-in sampler */
int f1o_plan(const f1o_config* cfg, const f1o_world* w, const double pose[4], const double* opp,
             int n_opp, const double* goals_in, int n_goals_in, int c_begin, int c_end,
             f1o_result* out);
/* batch: poses [S,4], opp [S,max_opp,3], n_opp [S] or NULL; outputs [S]...; returns candidates
 * evaluated */
int64_t f1o_plan_batch(const f1o_config* cfg, const f1o_world* w, const double* poses,
                       const double* opp, const int32_t* n_opp, int s, int max_opp,
                       int32_t* best_idx, double* best_cost, double* best_traj, double* costs,
                       uint8_t* flags, double* steer_speed, int n_threads);

/* collision predicates on given vehicle-frame states (teacher-forced checks):
 * float32 mirror of the device predicate, bit-for-bit on identical inputs. */
void f1o_collide_f32(const float* states, const float* headings, int c, int m,
                     const float* opp_local, int n_opp, const float* grid_xf,
                     const int32_t* grid_i0, const uint8_t* grid, int gh, int gw,
                     float half_l, float half_w, float rc2, uint8_t* flags_out);

int f1o_max_threads(void);

#ifdef __cplusplus
}
#endif
#endif
