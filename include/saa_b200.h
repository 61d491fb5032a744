/*
 * saa_b200.h -- C ABI of libsaa_b200.so
 *
 * B200 (sm_100a) implementation of the sample-average-approximation (SAA)
 * "linearize + assemble" step of StanfordASL/RiskAverseTrajOpt.  Every entry
 * point below replaces a Python/JAX method of the reference's per-script
 * `Model` class; the citation after "replaces:" is the reference file:line.
 *
 * Conventions
 *   - plain C: pointers + sizes, no C++/torch types.  `stream` is a
 *     cudaStream_t passed as void* (NULL = the legacy default stream).
 *   - pointers named *_dev are device pointers owned by the CALLER (e.g. the
 *     data_ptr() of a torch CUDA tensor); *_host are host pointers.  The handle
 *     owns only its packed copy of the sample set, small scratch and tables.
 *   - every function returns 0 (SAA_OK) or a negative saa_status;
 *     saa_last_error() gives the message.  No C++ exception crosses the ABI.
 *   - a handle is not re-entrant: one host thread per handle; one process per
 *     GPU for multi-GPU runs.  All device work is stream-ordered on `stream`
 *     and asynchronous unless stated otherwise.
 *   - precision: 64 -> every output buffer (Ax, l, u, Z, Xs, factored record, hopper g / jac)
 *     is double; 32 -> float: the STORAGE precision of the outputs (half the HBM / NVLink / PCIe
 *     bytes).  The arithmetic is FP64 in both modes (the packed samples stay double), so FP32
 *     results are the FP64 results rounded once: per-entry relative error <= 6e-8, inside the
 *     1e-4 the FP32 mode promises even for entries that are small by cancellation.  Inputs to
 *     saa_set_samples_*, us_host, mean sums and CVaR / Hessian sums are always double.
 *
 * QP variable order (reference drone/drone_risk.py:226,381): z = (u[0..n_u*S),
 * y[0..M), slack, t).  Constraint-row order (drone_risk.py:282-374, :401-423):
 *   [final rows (n_fin)] [CVaR row] [-y_i rows (M)] [sample rows (M*rows_per_sample)]
 *   [-slack row] [control-bound rows (n_u*S)]            (method = SAA)
 *   [final rows] [sample rows] [control-bound rows]       (method = BASELINE)
 */
#ifndef SAA_B200_H
#define SAA_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct saa_handle saa_handle;

typedef enum {
  SAA_OK = 0,
  SAA_ERR_ARG = -1,        /* bad argument / unsupported configuration        */
  SAA_ERR_CUDA = -2,       /* a CUDA runtime call failed                      */
  SAA_ERR_STATE = -3,      /* call order violated (e.g. samples not set)      */
  SAA_ERR_NO_DEVICE = -4,  /* no usable CUDA device (there is NO CPU fallback)*/
  SAA_ERR_NONFINITE = -5   /* saa_check_finite: a rollout produced NaN / Inf  */
} saa_status;

typedef enum { SAA_DRONE = 0, SAA_CAR = 1, SAA_HOPPER = 2 } saa_problem;
typedef enum { SAA_METHOD_SAA = 0, SAA_METHOD_BASELINE = 1 } saa_method;
/* drone only: drone_risk.py vs drone_times.py relaxation / baseline constants */
typedef enum { SAA_VARIANT_RISK = 0, SAA_VARIANT_TIMES = 1 } saa_variant;

int saa_version(void);
/* message of the last failure on `h`; h == NULL -> last failure of saa_create */
const char *saa_last_error(const saa_handle *h);

/*
 * Create a handle for `M_local` samples living on CUDA device `device`, being
 * samples [sample_offset, sample_offset + M_local) of a global set of
 * `M_global` (single GPU: M_local == M_global, sample_offset == 0).
 * replaces: Model.__init__ (drone/drone_risk.py:70-93, car/driving.py:84-120,
 *           hopper/hopper.py:90-104).
 */
int saa_create(saa_handle **out, int problem, int method, int variant,
               int64_t M_local, int64_t M_global, int64_t sample_offset,
               int S, double alpha, int precision, int device);
int saa_destroy(saa_handle *h);

/* ---- problem constants (the reference's params modules) ------------------- */
typedef struct {               /* replaces: drone/drone_params.py:1-45        */
  double dt, u_max, beta, drag_coefficient;
  double gain_p, gain_v;       /* feedback_gain = -[gain_p I, gain_v I]       */
  double x_init[6], x_final[6];
  int32_t n_obs;               /* must be 3                                   */
  double obs_positions[3][3];
  double osqp_tol;             /* subtracted from Z_i in saa_cvar_terms       */
} saa_drone_params;

typedef struct {               /* replaces: car/driving_params.py:1-42        */
  double dt, u_max, beta;      /* beta: car/driving.py:94                     */
  double speed_ped_des, min_separation_distance;
  double goal[4];              /* (position_ego_goal, velocity_ego_goal)      */
  double osqp_tol;
} saa_car_params;

int saa_set_params_drone(saa_handle *h, const saa_drone_params *p);
int saa_set_params_car(saa_handle *h, const saa_car_params *p);

/* ---- sample sets (device pointers, reference array layouts, float64) ------
 * The handle repacks them into its own structure-of-arrays layout (only the
 * fields the path reads), so the caller may free its copies afterwards.       */
/* replaces: drone_utils.sample_uncertain_parameters outputs handed to Model
 * (drone/drone_utils.py:61-93): masses (M), DWs (M,S,6), obs_Qs (M,3,3,3);
 * obs_Qs must be diagonal (only [.,o,0,0] and [.,o,1,1] are read).            */
int saa_set_samples_drone(saa_handle *h, const double *masses_dev,
                          const double *DWs_dev, const double *obs_Qs_dev,
                          void *stream);
/* replaces: car Model fields (car/driving.py:95-120): states_init (M,8),
 * omegas_speed (M), omegas_repulsive (M), DWs (M,S,8).                        */
int saa_set_samples_car(saa_handle *h, const double *states_init_dev,
                        const double *omegas_speed_dev,
                        const double *omegas_repulsive_dev,
                        const double *DWs_dev, void *stream);

/* ---- static CSC pattern of A for M_global samples (host, no GPU work) ------
 * replaces: the pattern SciPy derives in sp.csr_matrix(dense) + sp.vstack(...,
 * format='csc') (drone/drone_risk.py:419-420, car/driving.py:417-418).  It is
 * structural (explicit zeros are kept).  `relaxed_pattern` != 0 gives the car's
 * scp_iter == 0 pattern, where rows >= n_x were multiplied by exactly 0 and
 * vanish (car/driving.py:411-415); ignored for the drone.                     */
int saa_pattern_sizes(const saa_handle *h, int relaxed_pattern,
                      int64_t *n_rows, int64_t *n_cols, int64_t *nnz);
int saa_pattern_i32(const saa_handle *h, int relaxed_pattern,
                    int32_t *indptr_host, int32_t *indices_host);
/* indices_host may be NULL in the _i64 variants: only the column pointers are written (the row
 * indices of a 10^6-sample matrix are 10+ GB; the runs are closed-form).        */
int saa_pattern_i64(const saa_handle *h, int relaxed_pattern,
                    int64_t *indptr_host, int64_t *indices_host);

/* Same pattern without a handle (pure host arithmetic; usable on a machine
 * without a GPU, e.g. by the process that owns the QP solver).               */
int saa_static_pattern_sizes(int problem, int method, int S, int64_t M, int relaxed_pattern,
                             int64_t *n_rows, int64_t *n_cols, int64_t *nnz);
int saa_static_pattern_i32(int problem, int method, int S, int64_t M, int relaxed_pattern,
                           int32_t *indptr_host, int32_t *indices_host);
int saa_static_pattern_i64(int problem, int method, int S, int64_t M, int relaxed_pattern,
                           int64_t *indptr_host, int64_t *indices_host);

/*
 * Where this rank's values go.  By default the destination is the GLOBAL
 * matrix (M_out = M_global, first_out = sample_offset).  A rank may instead
 * write a compact matrix of its own samples (M_out = M_local, first_out = 0).
 * Ax/l/u pointers given to the calls below are the base of a matrix with M_out
 * samples (they may be peer-mapped pointers to another GPU's buffer).
 */
int saa_set_output_geometry(saa_handle *h, int64_t M_out, int64_t first_out);

/*
 * Entries that do not depend on the iterate: y / slack / t columns, control
 * identity, every lower bound, the constant upper bounds.  They depend only on
 * whether the first-iteration relaxation is active (scp_iter below the
 * problem's threshold), so call it once per (buffer, relaxation state).  Only
 * this rank's sample slice of the per-sample constants is written, plus the
 * O(1) shared entries if `write_shared` != 0 (one rank should do that).
 * replaces: drone/drone_risk.py:221-237, :331-348, :360-368, :413-417;
 *           car/driving.py:243-258, :334-366, :411-415.
 */
int saa_write_constants(saa_handle *h, int scp_iter, int write_shared,
                        void *Ax_dev, void *l_dev, void *u_dev, void *stream);

/*
 * THE HOT PATH.  For every local sample: roll out the dynamics under the
 * controls `us_host` (S*n_u values, row-major (S, n_u), always double),
 * propagate the control sensitivities, evaluate the risk constraints, and
 * write (a) the u-column entries of the sample rows straight into the CSC
 * value array `Ax_dev`, (b) the sample rows' upper bounds into `u_dev`,
 * (c) optionally Z_i = max_k g_ik into `Z_dev` (M_local values, may be NULL),
 * (d) this rank's partial SUMS for the sample-mean rows into
 * `mean_sums_dev` (saa_mean_len() doubles; always double).
 * If `finalize` != 0 the sums are treated as global (single GPU), divided by
 * M_global and scattered into Ax/l/u (same as saa_finalize_means).
 * replaces: Model.get_all_constraints_coeffs_all + the relaxation of
 * get_constraints_coeffs (drone/drone_risk.py:239-374, :413-417;
 * car/driving.py:261-373, :411-415).
 */
int saa_linearize_assemble(saa_handle *h, const double *us_host, int scp_iter,
                           void *Ax_dev, void *l_dev, void *u_dev, void *Z_dev,
                           double *mean_sums_dev, int finalize, void *stream);
int64_t saa_mean_len(const saa_handle *h);
/* after an all-reduce(sum) of mean_sums over ranks: mean = sums / M_global ->
 * final rows of Ax, l[0..n_fin), u[0..n_fin).
 * replaces: jnp.mean(...) drone/drone_risk.py:294-300, car/driving.py:311-317 */
int saa_finalize_means(saa_handle *h, const double *mean_sums_dev, int scp_iter,
                       void *Ax_dev, void *l_dev, void *u_dev, void *stream);

/*
 * Fused all-reduce + finalize over NVLink peer memory (alternative to ncclAllReduce +
 * saa_finalize_means; <= 16 ranks of one box).  Every rank allocates an inbox of
 * saa_peer_inbox_bytes() with saa_shared_alloc (zero-initialised) and maps the others' with
 * saa_shared_open; inboxes_host[q] is rank q's inbox as seen from this rank.  One launch per
 * iteration on every rank with the same `epoch` (1, 2, 3, ...): the kernel stores this rank's
 * mean_sums into all inboxes, waits for all ranks' contributions, adds them IN RANK ORDER and
 * scatters mean = sum / M_global into Ax / l / u -- a few microseconds of NVLink latency, and
 * bitwise identical expectation rows on every rank.
 * replaces: jnp.mean(...) drone/drone_risk.py:294-300, car/driving.py:311-317 across ranks.    */
int64_t saa_peer_inbox_bytes(void);
int saa_peer_allreduce_finalize(saa_handle *h, const double *mean_sums_dev, int scp_iter,
                                void *Ax_dev, void *l_dev, void *u_dev, int rank, int world,
                                void *const *inboxes_host, uint64_t epoch, void *stream);

/*
 * Multi-GPU gather, NCCL variant: after the ranks' compact value blocks (a matrix of
 * M_shard samples each, i.e. what saa_linearize_assemble writes under
 * saa_set_output_geometry(M_shard, 0)) have been gathered into rank 0's memory,
 * copy one shard's per-iteration values -- the sub-run of every u column and the
 * sample rows' upper bounds -- to their place in the destination matrix of `h`
 * (M_out samples), starting at sample `first`.  Pure device-to-device copies.
 * (The fused alternative needs no call: pass peer-mapped Ax/u pointers of rank
 * 0's buffers to saa_linearize_assemble under the global geometry.)
 */
int saa_merge_shard(saa_handle *h, const void *shard_Ax_dev, const void *shard_u_dev,
                    int64_t M_shard, int64_t first, void *Ax_dev, void *u_dev, void *stream);

/*
 * Multi-GPU gather, factored variant (drone).  What crosses NVLink is not the 1140 Jacobian
 * entries of a sample but the two factors they are products of: the 380 control
 * sensitivities d p_k / d u_{j,a} and the 46-value trajectory record (p_1..p_S and the scaled
 * obstacle matrices per axis) -- 3.4 KB instead of 9.1 KB per sample.  A rank calls
 * saa_linearize_factored with (peer-mapped) fsp / fp / u pointers of the matrix owner, laid out
 * for the M_out samples of the output geometry (saa_factored_sizes elements); after a barrier
 * the owner calls saa_expand_factored for the sample range of the other ranks, which forms the
 * same products with the same instructions as saa_linearize_assemble (bitwise identical values).
 * The upper bounds, Z and the mean sums are produced by saa_linearize_factored as usual.
 */
int saa_factored_sizes(const saa_handle *h, int64_t *n_sp, int64_t *n_p);
int saa_linearize_factored(saa_handle *h, const double *us_host, int scp_iter, void *fsp_dev,
                           void *fp_dev, void *u_dev, void *Z_dev, double *mean_sums_dev,
                           void *stream);
int saa_expand_factored(saa_handle *h, int scp_iter, const void *fsp_dev, const void *fp_dev,
                        int64_t sample_begin, int64_t sample_count, void *Ax_dev, void *stream);

/*
 * Peer-mapped buffers for the fused multi-GPU gather (one process per GPU).  The
 * owner allocates with saa_shared_alloc (plain cudaMalloc on `device`) and sends the
 * 64-byte handle to the other processes (any host channel, e.g. torch.distributed);
 * they map it with saa_shared_open while THEIR device is current, which enables
 * NVLink peer access lazily, and pass the returned pointer as Ax/l/u to
 * saa_write_constants / saa_linearize_assemble under the global output geometry.
 */
int saa_shared_alloc(int device, int64_t bytes, void **ptr_out, unsigned char handle_out[64]);
int saa_shared_open(int device, const unsigned char handle[64], void **ptr_out);
int saa_shared_close(int device, void *ptr);
int saa_shared_free(int device, void *ptr);

/* Rollout only: Xs_dev (M_local, S+1, n_x) row-major.
 * replaces: Model.us_to_state_trajectories (drone/drone_risk.py:157-162,
 *           car/driving.py:207-214).                                          */
int saa_rollout(saa_handle *h, const double *us_host, void *Xs_dev, void *stream);

/*
 * CVaR / Monte-Carlo terms.  Z_i = max_k g_ik - osqp_tol per local sample
 * (written to Z_dev if not NULL), and into out3_dev (3 doubles, SUMS over the
 * local samples): [ sum_i max(Z_i - t_risk, 0),  #{i : Z_i <= sat_tol},
 * max_i Z_i ].  AV@R = t + out3[0] / (M_global * alpha) after an all-reduce.
 * replaces: monte_carlo_*_verification and the closed form of monte_carlo_avar
 * (drone/drone_risk.py:656-662, :694; car/driving.py:630-638, :670).          */
int saa_cvar_terms(saa_handle *h, const double *us_host, double t_risk,
                   double sat_tol, void *Z_dev, double *out3_dev, void *stream);

/* The assemble kernels are persistent (one grid fills every SM and its register file), so a
 * kernel launched on another stream -- e.g. the NCCL all-reduce of the mean sums -- cannot start
 * before they finish.  Leaving `n_sms` SMs out of the persistent grids lets it run concurrently
 * (multi-GPU steps reserve 1 of 148: 0.7 % of the kernel for ~75 us of exposed collective).   */
int saa_reserve_sms(saa_handle *h, int n_sms);

/* Measured FP64 FMA throughput (FMA / s) of `device`: 8 independent dependent-FMA chains per thread,
 * 32 warps per SM, best of 5 launches.  The denominator of the FP64 rooflines bench.py reports for
 * the car and hopper kernels (no reference counterpart; synchronous, takes a few ms).            */
int saa_measure_fp64_peak(int device, double *fma_per_s);

/*
 * Non-finite guard.  The car divides by |p_ego - p_ped| (car/driving.py:154); the reference
 * silently propagates NaN when the two coincide.  Every assemble / rollout launch counts the
 * samples whose rollout met a zero or non-finite distance (one atomicAdd per warp, on the
 * device).  saa_check_finite synchronises `stream`, reads and clears the counter, stores it in
 * *count_out (may be NULL) and returns SAA_ERR_NONFINITE if it is non-zero.                  */
int saa_check_finite(saa_handle *h, int64_t *count_out, void *stream);

/* ---- hopper slip-risk (hopper/hopper.py:300-367 and its derivatives) ------ */
/* replaces: intensities/thetas/taus (hopper/hopper.py:68-74), each (M, n_feat)*/
int saa_set_samples_hopper(saa_handle *h, int32_t n_features, double mu_nom,
                           const double *intensities_dev, const double *thetas_dev,
                           const double *taus_dev, void *stream);
/*
 * n_c contact instants with end-effector x positions px_host[c] and contact
 * forces (fx_host[c], fz_host[c]).  Writes, per local sample i and contact c
 * (index i*n_c + c):  mu_dev = mu_i(px_c),  dmu_dev = mu_i'(px_c).
 * If lambda_dev != NULL (M_local*n_c multipliers) also the per-contact sums
 * hess_sums_dev[2*c] = sum_i lambda_ic mu_i'(px_c), [2*c+1] = sum_i lambda_ic
 * mu_i''(px_c)  (2*n_c doubles, local sums; all-reduce across ranks).
 * replaces: friction_at_px under slip_risk_constraints and its jacrev /
 * hessian (hopper/hopper.py:75-81, :300-367, :569-580).                       */
int saa_hopper_friction(saa_handle *h, int32_t n_c, const double *px_host,
                        void *mu_dev, void *dmu_dev, const double *lambda_dev,
                        double *hess_sums_dev, void *stream);

/*
 * Device-side assembly of the slip-risk block of the hopper NLP.  The evaluation point is
 * given by the O(n_c) contact geometry the host extracts from the decision vector
 * (hopper/hopper.py:106-112, :166-171, :305-311) plus the M risk variables y (device):
 */
typedef struct {
  int32_t n_c;                /* contact instants (<= 32); reference: 20           */
  double px[32];              /* end-effector x = x0 + x3 sin x2 at the contact    */
  double fx[32], fz[32];      /* contact forces us[t, 2], us[t, 3]                 */
  double x2[32], x3[32];      /* leg angle and length (for d px / d(x0, x2, x3))   */
  double t_risk, slack;       /* Z[-1], Z[-2]                                      */
} saa_hopper_point;
/* g_dev[n_rows]: saa      [M alpha t + sum y] [-y_i (M)] [f_x - mu_i(p) f_z - t - y_i - slack (i-major,
 *                          contact-minor)] [0]          n_rows = 1 + M + M n_c + 1
 *                baseline [f_x - mu_i(p) f_z - slack]   n_rows = M n_c          (y_dev may be NULL)
 * replaces: slip_risk_constraints as consumed by eval_g (hopper/hopper.py:300-367, :591-593). */
int saa_hopper_g(saa_handle *h, const saa_hopper_point *pt, const double *y_dev, void *g_dev,
                 void *stream);
/* jac_dev[4 * M n_c]: the entries of the slip-risk rows that depend on the iterate, as four
 * arrays of M n_c values (index i n_c + c): d row / d x0, d x2, d x3 (= -f_z mu_i'(p) (1, x3 cos x2,
 * sin x2)) and d row / d f_z (= -mu_i(p)).  The other entries are the constants 1 (f_x), -1
 * (y_i, slack, t), 1 and M alpha (row 0), -1 (rows -y_i).
 * replaces: the slip-risk slice of jacrev(g) in eval_jac_g (hopper/hopper.py:569, :594-596). */
int saa_hopper_jac(saa_handle *h, const saa_hopper_point *pt, void *jac_dev, void *stream);
/* saa_hopper_g and saa_hopper_jac at the same point in ONE pass over the features (IPOPT asks for
 * eval_g and eval_jac_g at the same iterate): the 600 sincos per sample are evaluated once.      */
int saa_hopper_g_jac(saa_handle *h, const saa_hopper_point *pt, const double *y_dev, void *g_dev,
                     void *jac_dev, void *stream);
/* hess_dev[10 * n_c] (always double): per contact the lower triangle of the symmetric block of
 * hessian(lambda . g) on (x0, x2, x3, f_z)_t, as ten arrays of n_c values in the order (0,0) (1,0)
 * (1,1) (2,0) (2,1) (2,2) (3,0) (3,1) (3,2) (3,3); lambda_dev: the M n_c multipliers of the sample
 * rows.  Local samples only: all-reduce(sum) across ranks.
 * replaces: the slip-risk part of hess_lagrange_dot_g in eval_h (hopper/hopper.py:575-580, :622-628). */
int saa_hopper_hess(saa_handle *h, const saa_hopper_point *pt, const double *lambda_dev,
                    double *hess_dev, void *stream);
/* Monte-Carlo terms: Z_i = max_c (f_x - mu_i(p) f_z) into Z_dev (may be NULL) and out3_dev =
 * [sum_i max(Z_i - t_risk, 0), #{Z_i <= sat_tol}, max_i Z_i] with t_risk = pt->t_risk.
 * replaces: no_slip_constraints_verification and the closed form of avar
 * (hopper/hopper.py:910-925, :957).                                                           */
int saa_hopper_cvar_terms(saa_handle *h, const saa_hopper_point *pt, double sat_tol, void *Z_dev,
                          double *out3_dev, void *stream);

/* ---- tail-reduced subproblem (no reference counterpart) ----
 * At the solution of the Rockafellar-Uryasev program (drone/drone_risk.py:327-368,
 * car/driving.py:331-372) only the samples in the upper alpha-tail of
 * Z_i = max g_i have y_i > 0, so the host QP solver can be fed the K samples with
 * the largest Z_i at the current iterate instead of all M: a handle created with
 * M_local = K, M_global = M (the CVaR row keeps M*alpha*t, the expectation rows
 * keep the mean over all M samples) and output geometry (K, 0).  Per iteration:
 *   saa_linearize_means (full handle)  -> mean sums over ALL samples and Z_i, no matrix
 *   saa_select_tail     (full handle)  -> indices of the K largest Z_i, ascending
 *   saa_gather_samples  (K handle)     -> packed samples of those indices
 *   saa_linearize_assemble(K handle, finalize = 0) + saa_finalize_means(K handle, full sums)
 * The result equals the full matrix with the other samples' rows and y columns deleted. */
/* Sample-mean sums (layout of saa_linearize_assemble's mean_sums_dev) and, if Z_dev
 * is not NULL, Z_i = max_{o,k} g_i[o,k] of every local sample, without assembling.  */
int saa_linearize_means(saa_handle *h, const double *us_host, void *Z_dev,
                        double *mean_sums_dev, void *stream);
/* idx_out_dev[0..K): indices (ascending) of the K largest of the M_local values
 * Z_dev (storage precision of the handle); ties at the threshold go to the smaller index.
 * Exact and deterministic (radix select + ordered compaction).                  */
int saa_select_tail(saa_handle *h, const void *Z_dev, int64_t K,
                    int64_t *idx_out_dev, void *stream);
/*
 * The same selection over a SHARDED sample set (exact global top K, not per-rank): the radix select
 * with its 256-bin histogram exposed between the passes so that the caller can sum it over the ranks.
 *   saa_select_begin(h, K_global)
 *   for pass in 0 .. saa_select_passes(h)-1:
 *       saa_select_pass_hist(h, Z, pass, hist_dev)      local histogram of this digit (256 x uint32)
 *       all-reduce(sum) hist_dev over the ranks          (e.g. torch.distributed on an int32 tensor)
 *       saa_select_pass_pick(h, hist_dev, pass)          every rank picks the same digit
 *   saa_select_counts(h, Z, gt_eq_dev)                   local #{Z > threshold}, #{Z == threshold}
 *   all-gather the counts; ties are granted in rank (= global index) order:
 *       take_r = min(eq_r, max(0, K_global - sum_r' gt_r' - sum_{r' < r} eq_r'))
 *   saa_select_finish(h, Z, take_r, idx_out_dev)         local indices, ascending: gt_r + take_r of them
 * With one rank this reproduces saa_select_tail.                                                  */
int saa_select_passes(const saa_handle *h);
int saa_select_begin(saa_handle *h, int64_t K_global, void *stream);
int saa_select_pass_hist(saa_handle *h, const void *Z_dev, int pass, uint32_t *hist_dev, void *stream);
int saa_select_pass_pick(saa_handle *h, uint32_t *hist_dev, int pass, void *stream);
int saa_select_counts(saa_handle *h, const void *Z_dev, int64_t *gt_eq_dev, void *stream);
int saa_select_finish(saa_handle *h, const void *Z_dev, int64_t take_ties, int64_t *idx_out_dev,
                      void *stream);
/* A handle created for M_local samples may work on fewer: M_active <= that capacity of them, placed
 * at samples [first_out, first_out + M_active) of a matrix with M_out samples (the number a rank
 * contributes to a globally selected tail changes from iteration to iteration; 0 is allowed: the
 * launches become no-ops).                                                                        */
int saa_set_active(saa_handle *h, int64_t M_active, int64_t M_out, int64_t first_out);
/* dst's M_local samples := samples idx_dev[0..dst.M_local) of src (packed inputs
 * are copied device to device; dst needs saa_set_params_* but no saa_set_samples_*). */
int saa_gather_samples(saa_handle *dst, const saa_handle *src,
                       const int64_t *idx_dev, void *stream);


/* ----------------------------------------------------------------------------------------------
 * Device-resident ADMM for the CVaR program (second option).
 *
 * The QP the reference hands to OSQP every SCP iteration (drone/drone_risk.py:327-399 and 425-452,
 * car/driving.py:331-397 and 423-441) is solved where its matrix already is.  The algorithm is OSQP's
 * (ADMM, Ruiz equilibration, per-row rho, over-relaxation; Stellato et al. 2020); the linear system
 * (P + sigma I + A' R A) x = rhs is solved through the arrow structure of A: Sherman-Morrison for the
 * CVaR row, a diagonal y block, a Schur complement on the nu + 2 dense variables w = (u, slack, t).
 * These entry points are the sample-sized passes (one warp per sample, values read in place from the
 * Ax / l / u buffers of saa_linearize_assemble) and the one-block step for the dense variables; the
 * O(nu^2) host logic (inverting the Schur complement, adapting rho, the termination test) is in
 * riskaversetrajopt_b200/device_qp.py.  Sharded samples: all-reduce the reduced partials between a
 * pass and saa_qp_dense_step (sum, or max where stated).  FP64, method 'saa', not the car's scp_iter 0.
 * -------------------------------------------------------------------------------------------- */
typedef struct saa_qp_sample_state {   /* device arrays over this handle's M_local samples */
  double *Dy, *Ey, *Es;   /* Ruiz scalings: y_i column (M), "-y_i" row (M), sample rows (M * R)        */
  double *xy, *rloc;      /* scaled y variables; sample-local part of the right-hand side (M each)    */
  double *zy, *ly;        /* ADMM z and multiplier of the "-y_i" rows (M each)                        */
  double *zs, *ls;        /* ... of the sample rows (M * R each)                                      */
} saa_qp_sample_state;
/* Sizes, row / column offsets, offsets into the packed state G of saa_qp_dense_step, per-u-column
 * (offset, final rows, run length), the active columns and the pair table of saa_qp_gram_pass, as a
 * flat int64 array (order: riskaversetrajopt_b200/device_qp.py:_Layout). */
int saa_qp_layout(saa_handle *h, int64_t *out, int64_t cap);
/* kind: 0 scale, 1 gram, 2 admm pass, 3 check -> grid size and partial-vector length of that pass */
int saa_qp_partials(saa_handle *h, int kind, int64_t *nblocks, int64_t *plen);
/* One Ruiz sweep (norms with the old scalings, sample-local scalings updated in place); partials are
 * maxima: [0, nu) u-column norms, slack column, t column, max D_y. */
int saa_qp_scale_pass(saa_handle *h, const void *Ax, const void *l, const void *u, const double *Dw,
                      double Ec, const saa_qp_sample_state *st, double *partials, void *stream);
/* Sample share of the Schur complement and of the Sherman-Morrison vectors (sums). */
int saa_qp_gram_pass(saa_handle *h, const void *Ax, const void *l, const void *u, const double *Dw,
                     double Ec, double rho, double sigma, const saa_qp_sample_state *st,
                     double *partials, void *stream);
/* One ADMM iteration's sample part: finishes the iteration whose x~_w = xt[0..nu+2), gamma' =
 * xt[nu+2] the dense step produced (skipped when first != 0) and accumulates the next right-hand
 * side: partials (sums) = [R_u (nu), R_slack, R_t, sigma_1, cv]. */
int saa_qp_admm_pass(saa_handle *h, const void *Ax, const void *l, const void *u, const double *Dw,
                     double Ec, double rho, double sigma, double alpha, const saa_qp_sample_state *st,
                     const double *xt, int first, double *partials, void *stream);
/* Termination-test pieces of the sample rows at x_w = xw_lamc[0..nu+2), CVaR multiplier
 * xw_lamc[nu+2]: partials = [max |(Ax-z)/E|, max |Ax/E|, max |z/E|, max |(A'lam)_y/D_y| | sums:
 * (A'lam)_u (nu), (A'lam)_slack, (A'lam)_t]. */
int saa_qp_check_pass(saa_handle *h, const void *Ax, const void *l, const void *u, const double *Dw,
                      double Ec, const saa_qp_sample_state *st, const double *xw_lamc,
                      double *partials, void *stream);
/* out[e] = max (e < n_max) or sum over the nblocks partial vectors, in a fixed order. */
int saa_qp_reduce(saa_handle *h, const double *partials, int64_t nblocks, int64_t plen, int64_t n_max,
                  double *out, void *stream);
/* Dense variables and sample-independent rows: finishes the iteration on them (skipped when
 * first != 0), completes the right-hand side from red = reduced partials of saa_qp_admm_pass and
 * solves for the next x~_w, gamma' (written into G). */
int saa_qp_dense_step(saa_handle *h, double *G, const double *red, int first, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* SAA_B200_H */
