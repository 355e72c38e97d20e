/* chi_b200 -- C ABI of the B200-native likelihood engine for chi.
 *
 * Plain pointers and sizes only; every buffer is owned by the caller, is
 * C-contiguous, and is never retained past the call (static population data
 * is copied to the device by chi_b200_problem_upload).  Functions return 0 on
 * success and a negative code on error (message: chi_b200_last_error()).
 * Numerical failure of a single ODE solve is NOT an error: it is reported in
 * the per-system `status` word and mapped by the caller to the reference's
 * convention (score -inf, gradient +inf, RuntimeWarning;
 * chi/_log_pdfs.py:812-821, 990-996).
 *
 * Reference interfaces replaced (paths relative to the chi repository):
 *   chi_b200_simulate        SBMLModel.simulate          chi/_mechanistic_models.py:490-533
 *                            (-> myokit.Simulation.run / CVODES, un-vendored)
 *   chi_b200_problem_upload  PKPDModel.set_dosing_regimen chi/_mechanistic_models.py:795-857
 *                            LogLikelihood.__init__       chi/_log_pdfs.py:702-800, 845-883
 *   chi_b200_problem_set_sensitivity_columns
 *                            SBMLModel.enable_sensitivities chi/_mechanistic_models.py:244-312
 *   chi_b200_loglik          LogLikelihood.__call__/evaluateS1 chi/_log_pdfs.py:802-843, 973-1027
 *                            ErrorModel.compute_log_likelihood / compute_sensitivities
 *                                                         chi/_error_models.py:229-331, 581-664,
 *                                                         925-1021, 1276-1375
 *   chi_b200_error_model     the same error-model methods called stand-alone
 *   chi_b200_population_*    PopulationModel.compute_individual_parameters /
 *                            compute_log_likelihood / compute_sensitivities
 *                                                         chi/_population_models.py:386-449, 567-666,
 *                                                         1034-1194, 1575-1757, 1948-2073,
 *                                                         2394-2599, 2832-2965
 *                            LinearCovariateModel         chi/_covariate_models.py:257-320
 *   chi_b200_hier_logpdf     HierarchicalLogLikelihood.__call__/evaluateS1
 *                                                         chi/_log_pdfs.py:104-145, 255-300
 */
#ifndef CHI_B200_H
#define CHI_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CHI_B200_ABI_VERSION 1

/* error-model kinds (chi/_error_models.py) */
#define CHI_B200_ERR_GAUSSIAN 0
#define CHI_B200_ERR_LOGNORMAL 1
#define CHI_B200_ERR_CAM 2
#define CHI_B200_ERR_MULTIPLICATIVE 3

/* population block kinds (chi/_population_models.py) */
#define CHI_B200_POP_POOLED 0
#define CHI_B200_POP_HETEROGENEOUS 1
#define CHI_B200_POP_GAUSSIAN 2
#define CHI_B200_POP_LOGNORMAL 3
/* TruncatedGaussianModel (chi/_population_models.py:3500-3939), always centered */
#define CHI_B200_POP_TRUNCATED_GAUSSIAN 4

/* per-system status words */
#define CHI_B200_ST_OK 0
#define CHI_B200_ST_MAX_STEPS 1
#define CHI_B200_ST_STEP_UNDERFLOW 2
#define CHI_B200_ST_NON_FINITE 3
/* hierarchical evaluation with individuals sharded over several GPUs: all of
 * this rank's systems succeeded, a system of another rank failed */
#define CHI_B200_ST_REMOTE_FAILURE 4

typedef struct chi_b200_problem chi_b200_problem;

/* ---- library ---------------------------------------------------------- */
int chi_b200_abi_version(void);
const char* chi_b200_last_error(void);
int chi_b200_device_count(void);
/* pinned host memory for callers that want asynchronous copies */
int chi_b200_host_alloc(void** ptr, int64_t bytes);
int chi_b200_host_free(void* ptr);

/* ---- model registry ------------------------------------------------------
 * Models are compiled from SBML by chi_b200/_codegen.py (one translation unit
 * per model) and register themselves when their shared object is loaded. */
int chi_b200_n_models(void);
int chi_b200_model_find(const char* key);
/* info: n_states, n_consts, n_outputs, linear (bit 0: rhs linear in the
 * states; bit 1: every constant's column term has a single non-zero column,
 * the mass-action case), and op counts of the generated
 * code: rhs, Jacobian, the column-specialised sensitivity terms summed over
 * the constants that enter the rhs, and the column-independent second-order
 * term (zero for linear models) */
int chi_b200_model_info(int model, int32_t info[8]);
const char* chi_b200_model_key(int model);
/* bit k set: constant k does not enter the rhs (its sensitivity column of the
 * states is identically zero and is not integrated) */
int chi_b200_model_inert_mask(int model, uint32_t* mask);
/* sensitivity kernel variants: lanes per system, columns per lane */
int chi_b200_model_variants(int model, int32_t* n, int32_t* lanes,
                            int32_t* cols, int32_t capacity);
/* registers/thread, block size, resident blocks per SM, local bytes of a
 * variant (-1: forward-only kernel) */
int chi_b200_model_describe(int model, int device, int variant,
                            int32_t out[4]);

/* ---- problem: one model + one population resident on one GPU ----------- */
/* model = -1 creates a population-only problem (stand-alone population
 * model evaluation); its size is set with chi_b200_problem_set_n_ids. */
int chi_b200_problem_create(int model, int device, chi_b200_problem** out);
int chi_b200_problem_set_n_ids(chi_b200_problem* p, int64_t n_ids);
void chi_b200_problem_destroy(chi_b200_problem* p);

/* Which model outputs are observed, and the error model of each. */
int chi_b200_problem_set_outputs(chi_b200_problem* p, int32_t n_out,
                                 const int32_t* output_ids,
                                 const int32_t* err_kinds);
/* Sensitivity columns: indices into the mechanistic parameter vector
 * (initial values first, then constants; chi order). */
int chi_b200_problem_set_sensitivity_columns(chi_b200_problem* p,
                                             int32_t n_cols,
                                             const int32_t* param_idx);
int chi_b200_problem_set_tolerances(chi_b200_problem* p, double rtol,
                                    double atol, int32_t max_steps);
int chi_b200_problem_set_variant(chi_b200_problem* p, int32_t variant);
/* Entries of the per-system parameter vector [psi | sigma] held at fixed
 * values inside chi_b200_hier_logpdf (LogLikelihood.fix_parameters,
 * chi/_log_pdfs.py:1029-1073): the population model then spans the free
 * entries only.  Call after set_outputs and before population_set; indices
 * increasing. */
int chi_b200_problem_set_fixed_parameters(chi_b200_problem* p,
                                          int32_t n_fixed,
                                          const int32_t* idx,
                                          const double* values);

/* Static population data.
 *   rec_off[n_ids+1]   observation records of individual i are
 *                      rec_off[i]..rec_off[i+1]-1, sorted by time
 *   rec_t, rec_obs     time and observed value (rec_obs may be NULL)
 *   rec_out            index into the problem's output list
 *   dose_off[n_ids+1]  dose events of individual i
 *   dose_events        5 doubles per event: level, start, duration, period,
 *                      multiplier (myokit.ProtocolEvent fields; period 0 =
 *                      single pulse, multiplier 0 = indefinitely)
 */
int chi_b200_problem_upload(chi_b200_problem* p, int64_t n_ids,
                            const int64_t* rec_off, const double* rec_t,
                            const double* rec_obs, const int32_t* rec_out,
                            const int64_t* dose_off,
                            const double* dose_events);

/* ---- hot path, host buffers (copies inside the call) -------------------- */
/* psi: [B][P].  y_out: one value per record of each system, systems
 * concatenated; s_out: [rows][n_cols].  ids may be NULL (b % n_ids). */
int chi_b200_simulate(chi_b200_problem* p, int64_t B, const int32_t* ids,
                      const double* psi, int32_t with_sens, double* y_out,
                      double* s_out, int32_t* status);
/* x: [B][P + n_sigma]; ll: [B]; grad: [B][P + n_sigma] */
int chi_b200_loglik(chi_b200_problem* p, int64_t B, const int32_t* ids,
                    const double* x, int32_t with_grad, double* ll,
                    double* grad, int32_t* status, int32_t* n_steps,
                    int32_t* n_rej);

/* ---- hot path, device buffers (no copies; `stream` is a cudaStream_t) ---- */
int chi_b200_loglik_dev(chi_b200_problem* p, int64_t B, const int32_t* d_ids,
                        const double* d_x, int32_t with_grad, double* d_ll,
                        double* d_grad, int32_t* d_status, int32_t* d_n_steps,
                        int32_t* d_n_rej, void* stream);

/* ---- stand-alone error model (ErrorModel.compute_* on one individual) ---- */
/* out: ll, then (with_grad) P + n_sigma gradient entries; pointwise (may be
 * NULL): n_obs values. */
int chi_b200_error_model(int32_t device, int32_t kind, const double* sigma,
                         int64_t n_obs, int32_t n_par, const double* ybar,
                         const double* sens, const double* obs,
                         int32_t with_grad, double* out, double* pointwise);

/* ---- stand-alone linear covariate model -----------------------------------
 * LinearCovariateModel.compute_population_parameters / compute_sensitivities
 * (chi/_covariate_models.py:257-320): vartheta_i = pop + X_i beta on the
 * n_sel selected (param, dim) entries; adjoint d pop = sum_i d vartheta_i,
 * d beta[s][c] = sum_i d vartheta_i[sel s] X_ic.
 *   beta [n_sel][n_cov], pop [n_pop][n_dim], covariates [n_ids][n_cov]
 *   vartheta (may be NULL) [n_ids][n_pop][n_dim]
 *   dlogp_dvartheta (may be NULL) [n_ids][n_pop][n_dim] -> dpop
 *   [n_pop * n_dim], dbeta [n_sel * n_cov] */
int chi_b200_covariate_linear(int32_t device, int64_t n_ids, int32_t n_pop,
                              int32_t n_dim, int32_t n_cov, int32_t n_sel,
                              const int32_t* pidx, const int32_t* didx,
                              const double* beta, const double* pop,
                              const double* covariates,
                              const double* dlogp_dvartheta,
                              double* vartheta, double* dpop, double* dbeta);

/* ---- population model ------------------------------------------------------
 * A composed population model is a list of blocks over consecutive
 * dimensions.  Per block: kind, n_dim, centered flag, n_cov (0 = none) and,
 * for covariate blocks, n_sel selected (param, dim) pairs. */
int chi_b200_population_set(chi_b200_problem* p, int32_t n_blocks,
                            const int32_t* kind, const int32_t* n_dim,
                            const int32_t* centered, const int32_t* n_cov,
                            const int32_t* n_sel, const int32_t* sel_pidx,
                            const int32_t* sel_didx,
                            const double* covariates /* [n_ids][n_cov_total] */);
/* sizes implied by the population model: n_bottom, n_top, n_dim, n_hdim */
int chi_b200_population_sizes(chi_b200_problem* p, int64_t out[4]);
/* eta/psi transforms and population score/sensitivities stand-alone:
 *   top [n_top], bottom [n_ids][n_hdim] -> eta, psi [n_ids][n_dim]
 *   dlogp_dpsi [n_ids][n_dim] (may be NULL) -> score, grad[n_bottom+n_top] */
int chi_b200_population_eval(chi_b200_problem* p, const double* top,
                             const double* bottom, const double* dlogp_dpsi,
                             int32_t with_grad, double* eta, double* psi,
                             double* score, double* grad);

/* ---- hierarchical log-likelihood over C parameter vectors --------------- */
/* x: [C][n_bottom + n_top]; score: [C]; grad: [C][n_bottom + n_top];
 * status: [C] (max over the chain's systems). id range [id_begin, id_end) is
 * the shard of individuals evaluated by this call (multi-GPU).  With a shard
 * set, only the shard's block of bottom-level entries and the top-level
 * entries of x are read, and only those two column blocks of grad are
 * written (partial sums for the top level, or -- with a communicator set,
 * chi_b200_problem_comm_init -- the sums over all ranks); the rest of grad is
 * untouched.
 * The host-buffer entry point evaluates large batches in chunks on two
 * streams so that the copies overlap the solves (environment
 * CHI_B200_CHUNKS forces the number of chunks, 1 = one pass); page-locked
 * host buffers are needed for the overlap, not for correctness.  The call
 * returns when all results are in the caller's buffers.
 * With at least 32 vectors the persistent solve kernel hands the systems out
 * individual-major, one individual per warp (CHI_B200_ORDER = smallest C for
 * which it does, 0 = never; CHI_B200_REFILL_WAIT = iterations an idle lane
 * waits for its warp before it refills alone); results do not depend on it. */
int chi_b200_hier_logpdf(chi_b200_problem* p, int32_t C, const double* x,
                         int32_t with_grad, double* score, double* grad,
                         int32_t* status);
/* The same without the final wait: everything is queued (copies included, the
 * buffers have to be page-locked and must stay untouched) and the call
 * returns; chi_b200_problem_synchronize waits for the results.  Lets the
 * caller evaluate the (host-side) log-prior of chi.HierarchicalLogPosterior
 * (chi/_log_pdfs.py:474-498) while the GPU works. */
int chi_b200_hier_logpdf_begin(chi_b200_problem* p, int32_t C, const double* x,
                               int32_t with_grad, double* score, double* grad,
                               int32_t* status);
int chi_b200_problem_synchronize(chi_b200_problem* p);
int chi_b200_hier_logpdf_dev(chi_b200_problem* p, int32_t C,
                             const double* d_x, int32_t with_grad,
                             double* d_score, double* d_grad,
                             int32_t* d_status, void* stream);
/* shard of individuals this problem evaluates inside hier_logpdf
 * (default: all); without a communicator the partial sums are left for the
 * caller to reduce */
int chi_b200_problem_set_shard(chi_b200_problem* p, int64_t id_begin,
                               int64_t id_end);
/* Individuals sharded over the GPUs of one node (one process per GPU).  The
 * reference sums the individuals' contributions in a Python loop
 * (chi/_log_pdfs.py:289-292); with a communicator set, chi_b200_hier_logpdf
 * and chi_b200_hier_logpdf_dev finish every evaluation with ONE all-reduce
 * (sum, FP64, NCCL over NVLink) of [score | reduced top-level gradient |
 * invalid-population flag | failure flag] on the compute stream, before any
 * copy to the host: every rank returns the complete score and top-level
 * gradient plus its own individuals' block of the bottom-level gradient.
 *   chi_b200_comm_unique_id     rank 0 calls it and hands the
 *                               CHI_B200_COMM_ID_BYTES bytes to the other ranks
 *   chi_b200_problem_comm_init  collective over all ranks; the communicator is
 *                               kept for the life of the problem
 * All ranks must then evaluate the same number of vectors per call.  NCCL is
 * bound at run time (libnccl.so.2; CHI_B200_NCCL_LIB overrides the name). */
#define CHI_B200_COMM_ID_BYTES 128
int chi_b200_comm_unique_id(uint8_t* id, int32_t capacity);
int chi_b200_problem_comm_init(chi_b200_problem* p, int32_t world,
                               int32_t rank, const uint8_t* id);
int chi_b200_problem_comm_destroy(chi_b200_problem* p);
/* work counters of the last hot-path call: systems, accepted steps,
 * rejected steps, kernel launches */
int chi_b200_problem_counters(chi_b200_problem* p, int64_t out[4]);

/* ---- measurement ----------------------------------------------------------- */
/* When enabled the solve kernel of every hot-path call is bracketed by CUDA
 * events on its launch stream; timings returns {sum of kernel ms, launches}
 * since timing was enabled. */
int chi_b200_problem_set_timing(chi_b200_problem* p, int32_t enabled);
int chi_b200_problem_timings(chi_b200_problem* p, double out[2]);
/* Deferred observation.  By default the streamed solve kernel only stores
 * the state and the integrated sensitivity columns at every observation
 * record (B x max records x N (1 + columns) doubles of device memory) and a
 * second kernel evaluates the output map, the error model
 * (chi/_error_models.py:288-330, 631-674, 983-1026, 1342-1385) and the
 * gradient from these snapshots; inside the solve that work runs with one or
 * two active lanes per warp.  The snapshots are used when they fit `bytes`
 * (default: half of the free device memory at the first evaluation, or
 * CHI_B200_SNAP_MB); 0 keeps the observation fused in the solve kernel,
 * -1 restores the default. */
int chi_b200_problem_set_snapshot_budget(chi_b200_problem* p, int64_t bytes);
/* How a batch of parameter vectors is worked through (see
 * chi_b200_hier_logpdf): number of chunks of the host-buffer entry point,
 * smallest number of vectors for which the solve kernel hands the systems out
 * individual-major (0: never), and the iterations an idle lane waits for its
 * warp before it refills alone.  -1 for any of them: the default (which the
 * CHI_B200_CHUNKS / CHI_B200_ORDER / CHI_B200_REFILL_WAIT environment
 * variables override).  Results do not depend on these settings. */
int chi_b200_problem_set_batching(chi_b200_problem* p, int32_t n_chunks,
                                  int32_t order_min_vectors,
                                  int32_t refill_wait);
/* Small launches -- single parameter vectors, the way pints' samplers and
 * optimisers call LogPDF.evaluateS1 (chi/_inference.py:729-747) -- run one
 * warp per system, the sensitivity columns side by side on its lanes, while
 * the launch has at most `max_systems` systems (-1: the default, one resident
 * warp per system: 12 x the SM count; 0: never; CHI_B200_SPLIT_MAX overrides
 * the default).  Each column is then integrated with the state under its own
 * step-size control: results agree with the batched form within the
 * integrator's tolerance, not bit for bit.  ROSL kernels of linear models. */
int chi_b200_problem_set_small_batch(chi_b200_problem* p,
                                     int64_t max_systems);
/* measured FP64 FMA throughput of the device (DFMA micro-kernel) */
int chi_b200_fp64_peak(int32_t device, double* tflops, double* sm_clock_mhz);

#ifdef __cplusplus
}
#endif
#endif /* CHI_B200_H */
