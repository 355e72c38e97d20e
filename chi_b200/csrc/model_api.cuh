// chi_b200 -- interface between generated model code, the integrator template
// and the core library.  sm_100a only; host code is C++17.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#ifdef __CUDACC__
#define CHI_HD __host__ __device__ __forceinline__
#else
#define CHI_HD inline __attribute__((always_inline))
#endif

namespace chi_b200 {

constexpr int MAX_OUT = 8;    // outputs selected per problem
constexpr int MAX_COLS = 32;  // sensitivity columns (mechanistic parameters)
constexpr int MAX_SIGMA = 4;     // error-model parameters per problem

// Error-model kinds (chi/_error_models.py): parameters per kind 1,1,2,1.
enum ErrKind : int32_t {
  ERR_GAUSSIAN = 0,        // GaussianErrorModel                :539-869
  ERR_LOGNORMAL = 1,       // LogNormalErrorModel               :872-1228
  ERR_CAM = 2,             // ConstantAndMultiplicativeGaussian :182-536
  ERR_MULTIPLICATIVE = 3,  // MultiplicativeGaussianErrorModel  :1231-1580
};

// Per-system status words (0 = ok).  The host maps any non-zero status to the
// reference's failure convention: score -inf, gradient +inf, RuntimeWarning
// (chi/_log_pdfs.py:812-821, 990-996).
enum Status : int32_t {
  ST_OK = 0,
  ST_MAX_STEPS = 1,
  ST_STEP_UNDERFLOW = 2,
  ST_NON_FINITE = 3,
};

// Everything one launch of the solve kernel needs.  All pointers are device
// pointers.  Passed by value (fits the 4 KB kernel parameter space).
struct SolveArgs {
  int64_t B;                 // number of (individual x chain) systems
  int32_t n_ids;             // individuals in the uploaded population
  const int32_t* ids;        // [B] individual of each system, or null: b % n_ids
  const double* x;           // [B][x_stride]: psi (P) then sigma parameters
  int32_t x_stride;

  // static per-individual data (uploaded once, L2 resident)
  const int64_t* rec_off;    // [n_ids+1] observation records, sorted by time
  const double* rec_t;
  const double* rec_obs;
  const double* rec_logobs;
  const int32_t* rec_out;    // index into out_id[] below
  const int64_t* seg_off;    // [n_ids+1] piecewise-constant dose-rate segments
  const double* seg_t;       // segment start times (first is 0)
  const double* seg_rate;

  int32_t n_out;
  int32_t out_id[MAX_OUT];   // model output ids
  int32_t err_kind[MAX_OUT];
  int32_t err_off[MAX_OUT];  // offset of the output's sigmas inside the sigma block
  int32_t n_sigma;

  int32_t n_cols;            // sensitivity columns
  int32_t col_param[MAX_COLS];  // mechanistic parameter index of each column
  int32_t param_col[MAX_COLS];  // inverse map: column of parameter p, or -1
  // the columns that are integrated (all but those of constants that do not
  // enter the rhs), in column order: column index and parameter index
  int32_t n_int;
  int32_t int_col[MAX_COLS];
  int32_t int_param[MAX_COLS];

  int32_t mode;              // 0: simulate (write trajectories); 1: log-likelihood
  int32_t with_sens;
  double rtol, atol;
  int32_t max_steps;

  // outputs
  const int64_t* traj_off;   // [B] first trajectory row of each system (mode 0)
  double* traj_y;            // [rows]
  double* traj_s;            // [rows][n_cols]
  double* ll;                // [B]
  double* grad;              // [B][grad_stride] (psi columns then sigmas)
  int32_t grad_stride;
  int32_t* status;           // [B]
  int32_t* n_steps;          // [B] accepted steps (may be null)
  int32_t* n_rej;            // [B] rejected steps (may be null)

  // Deferred observation (streamed kernel, mode 1).  When `snap` is set the
  // solve kernel only stores the state and the integrated sensitivity columns
  // at every observation record, [B][snap_rows][snap_width] doubles, and a
  // second, fully convergent kernel evaluates output map, error model and
  // gradient from these snapshots.  Null: everything is fused in the solve.
  double* snap;
  int32_t snap_rows;         // rows per system (max records per individual)
  int32_t snap_width;        // N * (1 + integrated columns)

  // Order in which the persistent kernel hands out the systems.  With
  // order_inner = C > 1 and order_stride = n (systems laid out chain-major,
  // b = chain * n + individual) work item w is system
  //     b = (w % C) * n + w / C,
  // i.e. the lanes of a warp integrate the SAME individual under different
  // parameter vectors: same dose edges and observation times, so the lanes
  // reach their stops in (nearly) the same iterations and the event handler
  // runs with many active lanes instead of one or two.  0: b = w.
  int32_t order_inner, order_stride;
  // iterations an idle lane waits for the rest of its warp to finish before
  // it refills on its own (a warp whose lanes are all idle always refills as
  // a whole, 32 consecutive work items); 0: no waiting
  int32_t refill_wait;
  // work counter of the persistent kernel (one per launch that may be in
  // flight at a time); null: the launcher's own per-device counter
  unsigned long long* work_counter;
  // > 0: small launches, one work item per (system, integrated column): item
  // w integrates system w / split with the single sensitivity column
  // w % split (split = n_int) under its own step-size control -- the columns
  // of a system advance side by side on neighbouring lanes instead of one
  // after the other in one thread (log-likelihood mode, fused observation).
  // Every item integrates the state; item 0 writes the log-likelihood, the
  // error-model gradients and the gradient entries that need no sensitivity
  // column, item j the entry of its column; the status is the maximum.
  int32_t split;
  // Batched launches of the warp-synchronous kernels: scratch for handing
  // the vectors of an individual to the warps in the order of a stiffness
  // key (-trace J) instead of their order in the batch -- vectors that need
  // similar step sequences then share a warp, and fewer lanes wait at the
  // stops.  perm[i * order_inner + r] = vector at rank r of individual i
  // (written by the launcher, null: batch order), order_keys: the keys.
  int32_t* perm;
  float* order_keys;
};

// What a compiled model contributes to the registry of the core library.
struct ModelVTable {
  const char* key;
  int32_t n_states, n_consts, n_outputs, linear;
  int32_t flops_rhs, flops_jac, flops_dfdc, flops_hess;
  int32_t flops_col_sum, flops_hess_common;
  uint32_t inert_mask;  // constants that do not enter the rhs
  // group geometry of the sensitivity kernel: lanes per system and
  // sensitivity columns per lane (see integrator.cuh)
  int32_t n_variants;
  const int32_t* variant_lanes;
  const int32_t* variant_cols;
  // launches the solve on `stream`; variant < 0 picks the default.
  cudaError_t (*launch)(const SolveArgs& args, int variant,
                        cudaStream_t stream);
  // occupancy / register report for a variant (-1: forward kernel)
  cudaError_t (*describe)(int variant, int* regs, int* block,
                          int* blocks_per_sm, int* local_bytes);
};

}  // namespace chi_b200

extern "C" int chi_b200_register_model(const chi_b200::ModelVTable* vt);
