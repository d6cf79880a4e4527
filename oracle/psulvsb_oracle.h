/*
 * psulvsb_oracle.h -- C interface of the CPU parity oracle.
 *
 * TEST INFRASTRUCTURE ONLY.  This is a dependency-free CPU restatement of the PSULVSB solve path
 * of the reference (teaser/src/registration.cc, "REG" below; teaser/include/teaser/utils.h;
 * teaser/include/teaser/registration.h, "REGH").  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load it.  The product library
 * (libpsulvsb_b200.so) never links, includes or calls anything in this directory.
 *
 * Parity status: the reference cannot be compiled in this image (no Eigen3 / Boost / PCL / PMC,
 * no network), so the restatement is pinned against the reference's own test fixtures
 * (tests/golden/, see tests/test_oracle_golden.py):
 *   - length-consistency mask: fixed_scale_inliers.csv, bit-exact (28056 booleans)
 *   - GNC-TLS + svdRot: rotation-solver-test.cc:221-250 expected_R, 1e-5 rad
 *   - ScaleInliersSelector: scale-solver-test.cc:71-130
 *   - TLS translation: translation-solver-test.cc:21-113 (loose pin, estimator was rewritten)
 *   - end-to-end: registration-test.cc:229-308 (0.2 rad / 0.1 m)
 * The RANSAC schedule / sampling / self-update / refinement have no reference fixture:
 * "parity unpinned" for those; parity there is defined as product-vs-oracle on a replayed
 * Philox sample stream.
 */
#ifndef PSULVSB_ORACLE_H_
#define PSULVSB_ORACLE_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Mirrors RobustRegistrationSolver::Params (REGH:378-473) plus the compile-time constants and
 * in-loop overrides of REG lifted into fields (defaults = reference values). */
typedef struct {
  double noise_bound;             /* REGH:383  caller's noise bound                            */
  double cbar2;                   /* REGH:388                                                   */
  int estimate_scaling;           /* REGH:396                                                   */
  int rotation_max_iterations;    /* REGH:416                                                   */
  double rotation_gnc_factor;     /* REGH:411                                                   */
  double rotation_cost_threshold; /* REGH:426                                                   */
  int inlier_selection_mode;      /* REGH:365-370: 0 PMC_EXACT 1 PMC_HEU 2 KCORE_HEU 3 NONE     */
  double kcore_heuristic_threshold;
  /* lifted constants */
  double score_noise_bound;       /* REG:33  NOISE_BOUND macro (PrNoise = 2*this, REG:36)       */
  double inloop_noise_bound;      /* REG:938                                                    */
  double inloop_cbar2;            /* REG:939                                                    */
  int inloop_max_iterations;      /* REG:941                                                    */
  double inloop_gnc_factor;       /* REG:942                                                    */
  double inloop_cost_threshold;   /* REG:945                                                    */
  double rotation_similar;        /* REG:48                                                     */
  int local_max_iter;             /* REG:49                                                     */
  double tpro_host;               /* REG:772                                                    */
  double tpro_local;              /* REG:898                                                    */
  int host_round_limit;           /* REG:781                                                    */
  double wallclock_cap_s;         /* REG:1475 (60 s); <= 0 disables the rule                    */
  int self_update;                /* 1 = REG:786-832 enabled (canonical tree)                   */
  uint64_t seed;                  /* Philox key of the replayable sample stream                 */
} oracle_params_t;

typedef struct {
  int valid;
  double scale;
  int final_inlier_count;
  double translation[3];
  double rotation[9];             /* column-major, as Eigen::Matrix3d::data()                   */
  /* diagnostics */
  int host_rounds;
  int local_iters;
  long long n_line_vectors;       /* L  = C(C-1)/2                                              */
  long long n_reduced;            /* |L_reduced| after the one-time consistency pass            */
  int final_C;                    /* working-set size after self-update appends                 */
  int refined;                    /* 1 if REG:1516 accepted the weighted-SVD refinement         */
  int escalations;                /* number of rate escalations (REG:1377-1388)                 */
} oracle_solution_t;

typedef struct {
  int host_round;
  int local_iter;                 /* global index over the whole solve (keys the basic draw)    */
  int n_sampled_lines;
  int n_sampled_points;
  int basic_choose;
  int gnc_iterations;
  int rot_inliers;
  int n_rot_points;
  int similar;
  int curr_count;
  int best_count;
  int local_r;
  double p_local;
  double l_rate, b_rate;
  double scale;
  double R[9];                    /* column-major                                               */
  double t[3];
} oracle_local_trace_t;

typedef struct {
  int host_round;
  int curr_count;
  int best_host;
  int new_corr_count;
  int inlier_map_size;
  int host_r;
  double p_host;
} oracle_host_trace_t;

typedef struct {
  oracle_local_trace_t* local;
  int local_cap;
  int local_n;
  oracle_host_trace_t* host;
  int host_cap;
  int host_n;
  int* final_inliers;             /* optional, length M                                         */
  int* inlier_counter;            /* optional, length M                                         */
} oracle_trace_t;

void oracle_default_params(oracle_params_t* p);

/* REG:622-1535.  src/dst: column-major 3xC (the reduced set); ori_*: 3xM; keep_mask[M] in
 * {-1,0,1}; reduce_map[M] = reduced column of original j or -1.  src/dst are NOT modified; the
 * grown working set is reported through final_C only. */
int oracle_solve(const oracle_params_t* p, const double* src, const double* dst, int C,
                 const double* ori_src, const double* ori_dst, int M, const int* keep_mask,
                 const int* reduce_map, oracle_solution_t* out, oracle_trace_t* trace);

/* REG:418-434 applied to all ordered pairs: mask[i*n+j] = |  |s_i-s_j| - |t_i-t_j|  | <= beta,
 * diagonal 0.  margin (optional, n*n doubles) receives | |s|-|t| | - beta. */
void oracle_consistency_mask(const double* src, const double* dst, int n, double beta,
                             uint8_t* mask, double* margin);
/* Same test on explicit line vectors (3xK column-major). */
void oracle_scale_inliers(const double* sv, const double* tv, long long K, double beta,
                          uint8_t* mask);
/* Reduced set in reference order (REG:693-767, known scale): pairs (i<j) row-major. */
long long oracle_reduced_set(const double* src, const double* dst, int n, double beta,
                             int* pair_i, int* pair_j, long long cap);

/* utils.h:121-136 */
void oracle_svd_rot(const double* X, const double* Y, const double* W, long long K, double* R_colmajor);
/* 3x3 SVD used by the oracle (A = U diag(S) V^T), all column-major. */
void oracle_svd3(const double* A, double* U, double* S, double* V);

/* REG:1563-1692.  R_init: column-major warm start or NULL (first_time).  inliers[K] out. */
int oracle_gnc_tls(const double* sv, const double* tv, long long K, double noise_bound,
                   int max_iterations, double gnc_factor, double cost_threshold,
                   const double* R_init, double* R_colmajor, uint8_t* inliers, double* cost);

/* REG:436-463 + REG:121-203.  last_best: NULL (first_time) or 3 doubles. */
void oracle_tls_translation(const double* src, const double* dst, int N, double noise_bound,
                            double cbar2, const double* last_best, double* t_out,
                            uint8_t* inliers);

/* REG:66-120 via REG:397-415 (unknown scale): 1-D RANSAC consensus on line-vector length ratios. */
int oracle_tls_scale(const double* sv, const double* tv, long long K, double noise_bound,
                     double cbar2, const double* last_best, uint64_t seed, uint32_t event,
                     double* scale, uint8_t* inliers);

/* Scoring REG:1303-1336 / REG:1417-1444: residual_j = | q_j - s (R p_j + t) |. */
int oracle_score(const double* P, const double* Q, int N, double scale, const double* R_colmajor,
                 const double* t, double tau, uint8_t* inliers, double* residuals);

/* REG:526-602 */
void oracle_weighted_svd(const double* src, const double* tgt, const int* w, int M,
                         const double* T_init_colmajor4, double* T_out_colmajor4);
double oracle_rmse(const double* src, const double* tgt, const int* mask, int M,
                   const double* T_colmajor4);
/* REG:611-619: 1 - gamma_p(3/2, r^2 / (2 sigma^2)) */
double oracle_inlier_probability(double r, double sigma);

/* Replayable sample stream (Philox4x32-10, see DESIGN.md "Sample stream"). */
void oracle_philox4x32(uint64_t seed, uint32_t domain, uint32_t event, uint64_t block,
                       uint32_t out[4]);
uint32_t oracle_rand31(uint64_t seed, uint32_t domain, uint32_t event, uint64_t k);
double oracle_uniform01(uint64_t seed, uint32_t domain, uint32_t event, uint64_t k);
/* REG:852-861 / REG:916-932: `count` distinct draws of rand31 % n with rejection. */
long long oracle_sample_without_replacement(uint64_t seed, uint32_t domain, uint32_t event,
                                            long long n, long long count, int64_t* out);

/* Maximum clique (stands in for PMC, teaser/src/graph.cc:12-125): exact branch and bound. */
int oracle_max_clique(int n_vertices, const int* edge_u, const int* edge_v, long long n_edges,
                      int* clique_out);

#ifdef __cplusplus
}
#endif
#endif
