// chi_b200 -- core library: C ABI, model registry, problem handles, static
// population data, population-model kernels (K3) and the hierarchical
// log-pdf pipeline.  See include/chi_b200.h for the contract of every entry
// point and the reference interface it replaces.
#include <algorithm>
#include <climits>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include <dlfcn.h>
#include <nccl.h>

#include "../../include/chi_b200.h"
#include "model_api.cuh"

using namespace chi_b200;

namespace {

thread_local std::string g_error;

int fail(int code, const std::string& msg) {
  g_error = msg;
  return code;
}

#define CUDA_TRY(expr)                                                       \
  do {                                                                        \
    cudaError_t err__ = (expr);                                               \
    if (err__ != cudaSuccess)                                                 \
      return fail(-2, std::string(#expr) + ": " +                             \
                          cudaGetErrorString(err__));                         \
  } while (0)

// NCCL is bound at run time (dlopen) the first time a communicator is asked
// for: the library has no link-time dependency on it, and inside a process
// that has torch loaded the call resolves to the libnccl torch already mapped
// (same soname), so both see one NCCL.
struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t,
                            ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  std::string error;
  bool ok = false;
};

NcclApi& nccl_api() {
  static NcclApi api = [] {
    NcclApi a;
    const char* env = std::getenv("CHI_B200_NCCL_LIB");
    const char* names[] = {env, "libnccl.so.2", "libnccl.so"};
    for (const char* name : names) {
      if (!name || !*name) continue;
      a.handle = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
      if (a.handle) break;
    }
    if (!a.handle) {
      a.error = std::string("libnccl.so.2 not found: ") + dlerror();
      return a;
    }
    auto sym = [&](const char* n) { return dlsym(a.handle, n); };
    a.GetUniqueId =
        reinterpret_cast<decltype(a.GetUniqueId)>(sym("ncclGetUniqueId"));
    a.CommInitRank =
        reinterpret_cast<decltype(a.CommInitRank)>(sym("ncclCommInitRank"));
    a.CommDestroy =
        reinterpret_cast<decltype(a.CommDestroy)>(sym("ncclCommDestroy"));
    a.AllReduce =
        reinterpret_cast<decltype(a.AllReduce)>(sym("ncclAllReduce"));
    a.GetErrorString = reinterpret_cast<decltype(a.GetErrorString)>(
        sym("ncclGetErrorString"));
    a.ok = a.GetUniqueId && a.CommInitRank && a.CommDestroy && a.AllReduce &&
           a.GetErrorString;
    if (!a.ok) a.error = "libnccl.so.2 lacks an expected symbol";
    return a;
  }();
  return api;
}

#define NCCL_TRY(expr)                                                       \
  do {                                                                        \
    ncclResult_t res__ = (expr);                                              \
    if (res__ != ncclSuccess)                                                 \
      return fail(-4, std::string(#expr) + ": " +                             \
                          nccl_api().GetErrorString(res__));                  \
  } while (0)

std::vector<const ModelVTable*>& registry() {
  static std::vector<const ModelVTable*> r;
  return r;
}

// growable device buffer
struct DevBuf {
  void* ptr = nullptr;
  size_t cap = 0;
  cudaError_t reserve(size_t bytes) {
    if (bytes <= cap) return cudaSuccess;
    if (ptr) cudaFree(ptr);
    ptr = nullptr;
    cap = 0;
    size_t want = std::max(bytes, size_t(256));
    cudaError_t err = cudaMalloc(&ptr, want);
    if (err == cudaSuccess) cap = want;
    return err;
  }
  void release() {
    if (ptr) cudaFree(ptr);
    ptr = nullptr;
    cap = 0;
  }
  template <class T>
  T* as() const {
    return static_cast<T*>(ptr);
  }
};

// chunks of the host-buffer entry point (chi_b200_hier_logpdf): one work
// counter per chunk, one more for the un-chunked launches
constexpr int MAX_CHUNKS = 16;
constexpr int MAX_CHUNKS_P1 = MAX_CHUNKS + 1;
constexpr int MAX_POP_DIM = 32;
constexpr int MAX_POP_COV = 8;
constexpr int POP_BLOCK = 256;

// One dimension of the composed population model, flattened for the device.
struct PopDim {
  int32_t kind;
  int32_t centered;
  int32_t hcol;        // column among the hierarchical dims, -1 for special
  int32_t top0;        // pooled: theta idx; hetero: base; else mu idx
  int32_t top1;        // sigma idx
  int32_t het_stride;  // hetero: n_dim of the block
  int32_t het_off;     // hetero: dim inside the block
  int32_t n_cov;
  int32_t cov0;
  int32_t beta_mu;     // theta idx of the beta row of (mu, d), -1 none
  int32_t beta_sigma;
  int32_t red_mu;      // slot of mu / pooled theta in the reduced vector
  int32_t red_sigma;
  int32_t red_beta_mu;
  int32_t red_beta_sigma;
};

struct PopDesc {
  int32_t n_dim;
  int32_t n_hdim;
  int32_t n_cov_total;
  int32_t n_red;       // reduced top-level slots (all but heterogeneous)
  // layout of one system's parameter vector: the population dims are the
  // free entries, fixed entries carry constants (fix_parameters)
  int32_t n_sys;                    // length of the system vector
  int32_t sys_pos[MAX_POP_DIM];     // position of population dim d
  int32_t n_fixed;
  int32_t fixed_idx[MAX_POP_DIM];
  double fixed_val[MAX_POP_DIM];
  PopDim dim[MAX_POP_DIM];
};

}  // namespace

struct chi_b200_problem {
  const ModelVTable* vt = nullptr;
  int model = -1;
  int device = 0;
  cudaStream_t stream = nullptr;
  // second stream of the chunked host entry point (chi_b200_hier_logpdf) and
  // the work counters of the solve launches that may be in flight together
  cudaStream_t stream2 = nullptr;
  // copy streams and per-chunk events of the host-buffer entry point
  cudaStream_t stream_in = nullptr, stream_out = nullptr;
  std::vector<cudaEvent_t> ev_in, ev_comp;
  DevBuf work_counters;
  // chi_b200_problem_set_batching: -1 = default (environment, then built-in)
  int32_t cfg_chunks = -1, cfg_order = -1, cfg_wait = -1;
  int64_t cfg_split_max = -1;   // chi_b200_problem_set_small_batch

  // configuration
  int32_t n_out = 0;
  int32_t out_id[MAX_OUT];
  int32_t err_kind[MAX_OUT];
  int32_t err_off[MAX_OUT];
  int32_t n_sigma = 0;
  int32_t n_cols = 0;
  int32_t col_param[MAX_COLS];
  double rtol = 1e-8, atol = 1e-10;
  int32_t max_steps = 100000;
  int32_t variant = -1;

  // static data
  int64_t n_ids = 0;
  int64_t n_rec = 0;
  std::vector<int64_t> h_rec_off;
  DevBuf rec_off, rec_t, rec_obs, rec_logobs, rec_out;
  DevBuf seg_off, seg_t, seg_rate;

  // parameters of the system vector held at fixed values
  std::vector<int32_t> fixed_idx;
  std::vector<double> fixed_val;

  // population model
  bool has_pop = false;
  PopDesc pop;
  int64_t n_top = 0;
  std::vector<int32_t> red_to_top;  // reduced slot -> theta index
  DevBuf covariates;
  int64_t id_begin = 0, id_end = 0;

  // scratch
  DevBuf x, ll, grad, status, nst, nrj, ids, traj_off, traj_y, traj_s;
  // deferred observation (SolveArgs::snap): snapshots of state and
  // sensitivities at the records; used when they fit the budget
  DevBuf snap;
  int64_t max_rec = 0;          // largest number of records of an individual
  int64_t snap_budget = -1;     // bytes; -1: not determined yet
  bool snap_budget_auto = false;  // determined by the default policy
  DevBuf hx, hscore, hgrad, hstatus, partial, popflag, r2t;
  DevBuf order_perm, order_keys;
  DevBuf step_sums;             // chi_b200_problem_counters
  // individuals sharded over the GPUs of a node: communicator for the
  // all-reduce of [score | d theta] and the packed partials it works on
  ncclComm_t comm = nullptr;
  int32_t comm_world = 1, comm_rank = 0;
  DevBuf pack;
  int64_t counters[4] = {0, 0, 0, 0};
  bool counters_stale = false;

  // optional device timing of the solve kernel (bench / roofline)
  bool timing = false;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  double solve_ms_sum = 0.0;
  int64_t solve_launches = 0;
};

// ============================================================================
// Population kernels (K3)
// ============================================================================
namespace {

constexpr double LOG_2PI = 1.8378770664093454836;
constexpr double LOG_2PI_HALF_C = 0.91893853320467274178;

// theta_i(mu, sigma) of individual i for dimension d (covariate shift,
// chi/_covariate_models.py:257-284)
__device__ __forceinline__ void pop_mu_sigma(const PopDim& D,
                                             const double* top,
                                             const double* cov, int n_cov_tot,
                                             int64_t i, double& mu,
                                             double& sigma) {
  mu = top[D.top0];
  sigma = top[D.top1];
  if (D.n_cov > 0) {
    const double* xi = cov + i * n_cov_tot + D.cov0;
    if (D.beta_mu >= 0)
      for (int k = 0; k < D.n_cov; ++k) mu = fma(xi[k], top[D.beta_mu + k], mu);
    if (D.beta_sigma >= 0)
      for (int k = 0; k < D.n_cov; ++k)
        sigma = fma(xi[k], top[D.beta_sigma + k], sigma);
  }
}

// K3a: x[c] = [bottom | top] -> psi of every (chain, individual) system.
// Also writes the system -> individual map and flags chains whose population
// parameters are invalid (sigma < 0; chi/_population_models.py:1594-1595,
// 1639-1641, 2438-2439, 2480-2482).
__global__ void pop_forward_kernel(PopDesc P, int32_t C, int64_t n_ids,
                                   int64_t id_begin, int64_t id_end,
                                   int64_t n_params, int64_t n_bottom,
                                   const double* __restrict__ x,
                                   const double* __restrict__ cov,
                                   double* __restrict__ xsys,
                                   double* __restrict__ eta_out,
                                   int32_t* __restrict__ ids,
                                   int32_t* __restrict__ popflag) {
  const int64_t n_sh = id_end - id_begin;
  const int64_t tid = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (tid >= n_sh * C) return;
  const int64_t c = tid / n_sh;
  const int64_t i = id_begin + (tid - c * n_sh);
  const double* xc = x + c * n_params;
  const double* top = xc + n_bottom;
  const double* bot = xc + i * P.n_hdim;
  double* out = xsys + tid * P.n_sys;
  if (ids) ids[tid] = static_cast<int32_t>(i);
  for (int f = 0; f < P.n_fixed; ++f) out[P.fixed_idx[f]] = P.fixed_val[f];
  bool bad = false;
  for (int d = 0; d < P.n_dim; ++d) {
    const PopDim& D = P.dim[d];
    double eta, psi;
    if (D.kind == CHI_B200_POP_POOLED) {
      eta = psi = top[D.top0];
    } else if (D.kind == CHI_B200_POP_HETEROGENEOUS) {
      eta = psi = top[D.top0 + i * D.het_stride + D.het_off];
    } else {
      eta = bot[D.hcol];
      double mu, sigma;
      pop_mu_sigma(D, top, cov, P.n_cov_total, i, mu, sigma);
      if (sigma < 0.0) bad = true;
      // truncated Gaussian: sigma <= 0 or a negative individual parameter
      // is -inf (chi/_population_models.py:3557-3558, 3656-3658)
      if (D.kind == CHI_B200_POP_TRUNCATED_GAUSSIAN &&
          (sigma <= 0.0 || eta < 0.0))
        bad = true;
      if (D.centered || D.kind == CHI_B200_POP_TRUNCATED_GAUSSIAN) {
        psi = eta;
      } else if (sigma < 0.0) {
        psi = NAN;
      } else if (D.kind == CHI_B200_POP_GAUSSIAN) {
        psi = fma(sigma, eta, mu);
      } else {
        psi = exp(fma(sigma, eta, mu));
      }
    }
    out[P.sys_pos[d]] = psi;
    if (eta_out) eta_out[tid * P.n_dim + d] = eta;
  }
  if (bad && popflag) popflag[c] = 1;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// K3b: population densities and back-propagation of d ll / d psi to
// (eta, theta).  One thread per (chain, individual); top-level sums are
// reduced per warp, then per block in a fixed order, and written as partials
// [c][block][1 + n_red] (slot 0 = score) for the deterministic final sum.
__global__ void __launch_bounds__(POP_BLOCK)
pop_backward_kernel(PopDesc P, int64_t n_ids, int64_t id_begin,
                    int64_t id_end, int64_t n_params, int64_t n_bottom,
                    int32_t with_grad, const double* __restrict__ x,
                    const double* __restrict__ cov,
                    const double* __restrict__ ll,     // [C][n_sh] or null
                    const double* __restrict__ dlp,    // [C][n_sh][n_dim] or null
                    double* __restrict__ grad,         // [C][n_params]
                    double* __restrict__ partial) {
  extern __shared__ double sacc[];  // [warps][1 + n_red]
  const int n_slots = 1 + P.n_red;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_warps = blockDim.x >> 5;
  const int64_t n_sh = id_end - id_begin;
  const int64_t c = blockIdx.y;
  const int64_t k = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  const bool live = k < n_sh;
  const int64_t i = id_begin + (live ? k : 0);
  const double* xc = x + c * n_params;
  const double* top = xc + n_bottom;
  const double* bot = xc + i * P.n_hdim;
  const double* g_in =
      dlp ? dlp + (c * n_sh + (live ? k : 0)) * P.n_sys : nullptr;
  double* gc = grad ? grad + c * n_params : nullptr;

  for (int s = lane; s < n_slots; s += 32) sacc[warp * n_slots + s] = 0.0;
  __syncwarp();

  double score = (live && ll) ? ll[c * n_sh + k] : 0.0;
  for (int d = 0; d < P.n_dim; ++d) {
    const PopDim& D = P.dim[d];
    const double dl = (live && g_in) ? g_in[P.sys_pos[d]] : 0.0;
    if (D.kind == CHI_B200_POP_POOLED) {
      // PooledModel._shape(reduce=True): d theta = sum_i d ll/d psi_i
      // (chi/_population_models.py:2792-2810)
      if (with_grad) {
        const double s = warp_sum(dl);
        if (lane == 0) sacc[warp * n_slots + 1 + D.red_mu] += s;
      }
    } else if (D.kind == CHI_B200_POP_HETEROGENEOUS) {
      // HeterogeneousModel._shape(reduce=True): d theta = d ll/d psi flat
      // (chi/_population_models.py:1907-1926)
      if (with_grad && live)
        gc[n_bottom + D.top0 + i * D.het_stride + D.het_off] = dl;
    } else {
      double mu = 0.0, sigma = 1.0, eta = 0.0;
      if (live) {
        pop_mu_sigma(D, top, cov, P.n_cov_total, i, mu, sigma);
        eta = bot[D.hcol];
      }
      double s_i, deta, dmu, dsg;
      if (D.kind == CHI_B200_POP_TRUNCATED_GAUSSIAN) {
        // chi/_population_models.py:3560-3565, 3616-3627:
        // 1 - Phi(-mu/sigma) with Phi(x) = (1 + erf(x / sqrt 2)) / 2
        const double var = sigma * sigma;
        const double r = eta - mu;
        const double z = mu / sigma;
        const double tail = 1.0 - 0.5 * (1.0 + erf(-z * M_SQRT1_2));
        const double pdf = exp(-0.5 * z * z) * 0.39894228040143267794;
        s_i = -(0.5 * log(2.0 * M_PI * var) + r * r / (2.0 * var) + log(tail));
        deta = -r / var + dl;
        dmu = (r / sigma - pdf / tail) / sigma;
        dsg = (-1.0 + r * r / var + pdf * z / tail) / sigma;
      } else if (!D.centered) {
        // chi/_population_models.py:1474-1493 (Gaussian),
        // 2306-2339 (LogNormal)
        s_i = -(0.5 * LOG_2PI + 0.5 * eta * eta);
        if (D.kind == CHI_B200_POP_GAUSSIAN) {
          deta = dl * sigma - eta;
          dmu = dl;
          dsg = dl * eta;
        } else {
          const double psi = exp(fma(sigma, eta, mu));
          deta = dl * sigma * psi - eta;
          dmu = dl * psi;
          dsg = dl * eta * psi;
        }
      } else {
        const double var = sigma * sigma;
        if (D.kind == CHI_B200_POP_GAUSSIAN) {
          // chi/_population_models.py:1454-1472, 1513-1549
          const double r = eta - mu;
          s_i = -(0.5 * log(2.0 * M_PI * var) + r * r / (2.0 * var));
          deta = -r / var + dl;
          dmu = r / var;
          dsg = (-1.0 + r * r / var) / sigma;
        } else {
          // chi/_population_models.py:2285-2303, 2362-2392
          const double lo = log(eta);
          const double r = lo - mu;
          s_i = -(0.5 * log(2.0 * M_PI * var) + lo + r * r / 2.0 / var);
          deta = -(r / var + 1.0) / eta + dl;
          dmu = r / var;
          dsg = (-1.0 + r * r / var) / sigma;
        }
      }
      if (live) {
        score += s_i;
        if (with_grad) gc[i * P.n_hdim + D.hcol] = deta;
      } else {
        dmu = dsg = 0.0;
      }
      if (with_grad) {
        double s = warp_sum(dmu);
        if (lane == 0) sacc[warp * n_slots + 1 + D.red_mu] += s;
        s = warp_sum(dsg);
        if (lane == 0) sacc[warp * n_slots + 1 + D.red_sigma] += s;
        if (D.n_cov > 0) {
          // covariate adjoint, chi/_covariate_models.py:315-318
          const double* xi = cov + i * P.n_cov_total + D.cov0;
          for (int q = 0; q < D.n_cov; ++q) {
            const double xq = live ? xi[q] : 0.0;
            if (D.beta_mu >= 0) {
              s = warp_sum(dmu * xq);
              if (lane == 0) sacc[warp * n_slots + 1 + D.red_beta_mu + q] += s;
            }
            if (D.beta_sigma >= 0) {
              s = warp_sum(dsg * xq);
              if (lane == 0)
                sacc[warp * n_slots + 1 + D.red_beta_sigma + q] += s;
            }
          }
        }
      }
    }
  }
  {
    const double s = warp_sum(live ? score : 0.0);
    if (lane == 0) sacc[warp * n_slots] += s;
  }
  __syncthreads();
  double* pout = partial + (c * gridDim.x + blockIdx.x) * n_slots;
  for (int s = threadIdx.x; s < n_slots; s += blockDim.x) {
    double v = 0.0;
    for (int w = 0; w < n_warps; ++w) v += sacc[w * n_slots + s];
    pout[s] = v;
  }
}

// K3c: fixed-order sum of the block partials; scatters the reduced slots into
// the top-level gradient; applies the NaN -> -inf rule of the population
// likelihoods (chi/_population_models.py:1469-1470, 2300-2301) and the
// invalid-population flag.
// With `pack` set (individuals sharded over several GPUs) the sums are this
// rank's partials: they are written as [c][n_slots + 2] = [score | d theta
// (reduced slots) | invalid flag | failed systems] for the all-reduce, and
// pop_unpack_kernel finishes the job afterwards.
__global__ void pop_final_kernel(int32_t n_red, int64_t n_blocks,
                                 int64_t n_params, int64_t n_bottom,
                                 int32_t with_grad,
                                 const int32_t* __restrict__ red_to_top,
                                 const double* __restrict__ partial,
                                 const int32_t* __restrict__ popflag,
                                 const int32_t* __restrict__ status,
                                 double* __restrict__ score,
                                 double* __restrict__ grad,
                                 double* __restrict__ pack) {
  const int64_t c = blockIdx.x;
  const int n_slots = 1 + n_red;
  const bool bad = popflag && popflag[c];
  for (int s = threadIdx.x; s < n_slots; s += blockDim.x) {
    double v = 0.0;
    if (s == 0 || with_grad)
      for (int64_t b = 0; b < n_blocks; ++b)
        v += partial[(c * n_blocks + b) * n_slots + s];
    if (pack) {
      pack[c * (n_slots + 2) + s] = v;
    } else if (s == 0) {
      if (bad || v != v) v = -INFINITY;
      score[c] = v;
    } else if (with_grad) {
      grad[c * n_params + n_bottom + red_to_top[s - 1]] = v;
    }
  }
  if (pack && threadIdx.x == 0) {
    pack[c * (n_slots + 2) + n_slots] = bad ? 1.0 : 0.0;
    pack[c * (n_slots + 2) + n_slots + 1] =
        (status && status[c] != 0) ? 1.0 : 0.0;
  }
}

// After the all-reduce of the packed partials: the rules of K3c on the global
// sums.  A vector whose own systems all succeeded but that failed on another
// rank gets status CHI_B200_ST_REMOTE_FAILURE.
__global__ void pop_unpack_kernel(int32_t n_red, int64_t n_params,
                                  int64_t n_bottom, int32_t with_grad,
                                  const int32_t* __restrict__ red_to_top,
                                  const double* __restrict__ pack,
                                  double* __restrict__ score,
                                  double* __restrict__ grad,
                                  int32_t* __restrict__ status) {
  const int64_t c = blockIdx.x;
  const int n_slots = 1 + n_red;
  const double* pc = pack + c * (n_slots + 2);
  for (int s = threadIdx.x; s < n_slots; s += blockDim.x) {
    double v = pc[s];
    if (s == 0) {
      if (pc[n_slots] > 0.0 || v != v) v = -INFINITY;
      score[c] = v;
    } else if (with_grad) {
      grad[c * n_params + n_bottom + red_to_top[s - 1]] = v;
    }
  }
  if (status && threadIdx.x == 0 && status[c] == 0 && pc[n_slots + 1] > 0.0)
    status[c] = CHI_B200_ST_REMOTE_FAILURE;
}

__global__ void fill_kernel(double* p, int64_t n, double v) {
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i < n) p[i] = v;
}

// chain status = max over the chain's systems
__global__ void chain_status_kernel(int64_t n_sh, const int32_t* st,
                                    int32_t* out) {
  const int64_t c = blockIdx.x;
  int m = 0;
  for (int64_t k = threadIdx.x; k < n_sh; k += blockDim.x)
    m = max(m, st[c * n_sh + k]);
  m = __reduce_max_sync(0xffffffffu, m);
  __shared__ int sm[32];
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    int r = 0;
    for (int w = 0; w < (blockDim.x >> 5); ++w) r = max(r, sm[w]);
    out[c] = r;
  }
}

// sum of step counters
__global__ void count_kernel(int64_t n, const int32_t* a, const int32_t* b,
                             unsigned long long* out) {
  unsigned long long sa = 0, sb = 0;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
       i < n; i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    sa += a[i];
    sb += b[i];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    sa += __shfl_xor_sync(0xffffffffu, sa, o);
    sb += __shfl_xor_sync(0xffffffffu, sb, o);
  }
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(out, sa);
    atomicAdd(out + 1, sb);
  }
}

// Stand-alone linear covariate model (chi/_covariate_models.py:257-320).
// forward: vartheta[i][k] = pop[k] + sum_c X[i][c] beta[sel(k)][c] for the
// selected entries k = pidx * n_dim + didx (sel_of[k] = row of beta or -1).
__global__ void covariate_forward_kernel(int64_t n_ids, int32_t n_flat,
                                         int32_t n_cov,
                                         const int32_t* __restrict__ sel_of,
                                         const double* __restrict__ beta,
                                         const double* __restrict__ pop,
                                         const double* __restrict__ X,
                                         double* __restrict__ vartheta) {
  const int64_t t = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (t >= n_ids * n_flat) return;
  const int64_t i = t / n_flat;
  const int k = static_cast<int>(t - i * n_flat);
  double v = pop[k];
  const int s = sel_of[k];
  if (s >= 0)
    for (int c = 0; c < n_cov; ++c)
      v = fma(X[i * n_cov + c], beta[s * n_cov + c], v);
  vartheta[t] = v;
}

// adjoint: slot < n_flat: d pop[slot] = sum_i dl[i][slot]; else the beta
// entry (s, c): sum_i dl[i][flat_of[s]] X[i][c].  One block per slot, fixed
// summation order.
__global__ void __launch_bounds__(256)
covariate_backward_kernel(int64_t n_ids, int32_t n_flat, int32_t n_cov,
                          const int32_t* __restrict__ flat_of,
                          const double* __restrict__ X,
                          const double* __restrict__ dl,
                          double* __restrict__ out) {
  __shared__ double part[8];
  const int slot = blockIdx.x;
  int k = slot, c = -1;
  if (slot >= n_flat) {
    const int e = slot - n_flat;
    k = flat_of[e / n_cov];
    c = e % n_cov;
  }
  double acc = 0.0;
  for (int64_t i = threadIdx.x; i < n_ids; i += blockDim.x) {
    const double g = dl[i * n_flat + k];
    acc += c < 0 ? g : g * X[i * n_cov + c];
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double v = 0.0;
    for (int w = 0; w < 8; ++w) v += part[w];
    out[slot] = v;
  }
}

// Stand-alone error model (one individual): one warp.
__global__ void error_model_kernel(int32_t kind, double s0, double s1,
                                   int64_t n_obs, int32_t n_par,
                                   const double* __restrict__ ybar,
                                   const double* __restrict__ sens,
                                   const double* __restrict__ obs,
                                   int32_t with_grad, double* out,
                                   double* pointwise) {
  // out: [ll | grad (n_par + n_sigma)]
  const int lane = threadIdx.x;
  const int n_sig = kind == ERR_CAM ? 2 : 1;
  const bool invalid = (s0 <= 0.0) || (kind == ERR_CAM && s1 <= 0.0);
  bool bad_y = false;
  if (kind == ERR_LOGNORMAL)
    for (int64_t j = lane; j < n_obs; j += 32) bad_y |= ybar[j] <= 0.0;
  bad_y = __any_sync(0xffffffffu, bad_y);
  if (invalid || bad_y) {
    if (lane == 0) out[0] = -INFINITY;
    if (with_grad)
      for (int k = lane; k < n_par + n_sig; k += 32) out[1 + k] = INFINITY;
    if (pointwise)
      for (int64_t j = lane; j < n_obs; j += 32) pointwise[j] = -INFINITY;
    return;
  }
  double ll = 0.0, d0 = 0.0, d1 = 0.0;
  // gradient w.r.t. psi: lane-strided over observations, then reduced
  for (int k = 0; k < (with_grad ? n_par : 0); k += 1) {
    (void)k;
  }
  extern __shared__ double wbuf[];  // weights per observation
  for (int64_t j = lane; j < n_obs; j += 32) {
    const double yb = ybar[j], yo = obs[j];
    double w, l;
    if (kind == ERR_GAUSSIAN) {
      const double is = 1.0 / s0, q = (yo - yb) * is;
      l = -(LOG_2PI_HALF_C + log(s0) + 0.5 * q * q);
      w = q * is;
      d0 += (q * q - 1.0) * is;
    } else if (kind == ERR_LOGNORMAL) {
      const double is = 1.0 / s0, lo = log(yo);
      const double q = (lo - log(yb) + 0.5 * s0 * s0) * is;
      l = -(LOG_2PI_HALF_C + log(s0) + lo + 0.5 * q * q);
      w = q * is / yb;
      d0 += -q + (q * q - 1.0) * is;
    } else {
      const bool cam = kind == ERR_CAM;
      const double sb = cam ? s0 : 0.0, sr = cam ? s1 : s0;
      const double st = fma(sr, yb, sb), ist = 1.0 / st;
      const double q = (yo - yb) * ist, q2i = q * q * ist;
      l = -(LOG_2PI_HALF_C + log(st) + 0.5 * q * q);
      w = q * ist - sr * ist + sr * q2i;
      if (cam) {
        d0 += q2i - ist;
        d1 += (q2i - ist) * yb;
      } else {
        d0 += (q2i - ist) * yb;
      }
    }
    ll += l;
    if (pointwise) pointwise[j] = l;
    wbuf[j] = w;
  }
  __syncwarp();
  ll = warp_sum(ll);
  d0 = warp_sum(d0);
  d1 = warp_sum(d1);
  if (lane == 0) out[0] = ll;
  if (with_grad) {
    for (int k = 0; k < n_par; ++k) {
      double g = 0.0;
      for (int64_t j = lane; j < n_obs; j += 32)
        g = fma(wbuf[j], sens[j * n_par + k], g);
      g = warp_sum(g);
      if (lane == 0) out[1 + k] = g;
    }
    if (lane == 0) {
      out[1 + n_par] = d0;
      if (n_sig == 2) out[2 + n_par] = d1;
    }
  }
}

}  // namespace

// ============================================================================
// C ABI
// ============================================================================

extern "C" {

int chi_b200_register_model(const ModelVTable* vt) {
  for (auto* m : registry())
    if (std::strcmp(m->key, vt->key) == 0) return 0;
  registry().push_back(vt);
  return 0;
}

int chi_b200_abi_version(void) { return CHI_B200_ABI_VERSION; }

const char* chi_b200_last_error(void) { return g_error.c_str(); }

int chi_b200_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

int chi_b200_host_alloc(void** ptr, int64_t bytes) {
  CUDA_TRY(cudaHostAlloc(ptr, static_cast<size_t>(bytes), cudaHostAllocDefault));
  return 0;
}

int chi_b200_host_free(void* ptr) {
  CUDA_TRY(cudaFreeHost(ptr));
  return 0;
}

int chi_b200_n_models(void) { return static_cast<int>(registry().size()); }

int chi_b200_model_find(const char* key) {
  auto& r = registry();
  for (size_t i = 0; i < r.size(); ++i)
    if (std::strcmp(r[i]->key, key) == 0) return static_cast<int>(i);
  return -1;
}

const char* chi_b200_model_key(int model) {
  if (model < 0 || model >= chi_b200_n_models()) return "";
  return registry()[model]->key;
}

int chi_b200_model_info(int model, int32_t info[8]) {
  if (model < 0 || model >= chi_b200_n_models())
    return fail(-1, "model id out of range");
  const ModelVTable* vt = registry()[model];
  info[0] = vt->n_states;
  info[1] = vt->n_consts;
  info[2] = vt->n_outputs;
  info[3] = vt->linear;
  info[4] = vt->flops_rhs;
  info[5] = vt->flops_jac;
  info[6] = vt->flops_col_sum;
  info[7] = vt->flops_hess_common;
  return 0;
}

int chi_b200_model_inert_mask(int model, uint32_t* mask) {
  if (model < 0 || model >= chi_b200_n_models())
    return fail(-1, "model id out of range");
  *mask = registry()[model]->inert_mask;
  return 0;
}

int chi_b200_model_variants(int model, int32_t* n, int32_t* lanes,
                            int32_t* cols, int32_t capacity) {
  if (model < 0 || model >= chi_b200_n_models())
    return fail(-1, "model id out of range");
  const ModelVTable* vt = registry()[model];
  *n = vt->n_variants;
  for (int i = 0; i < vt->n_variants && i < capacity; ++i) {
    lanes[i] = vt->variant_lanes[i];
    cols[i] = vt->variant_cols[i];
  }
  return 0;
}

int chi_b200_model_describe(int model, int device, int variant,
                            int32_t out[4]) {
  if (model < 0 || model >= chi_b200_n_models())
    return fail(-1, "model id out of range");
  CUDA_TRY(cudaSetDevice(device));
  int regs = 0, block = 0, bpsm = 0, local = 0;
  CUDA_TRY(registry()[model]->describe(variant, &regs, &block, &bpsm, &local));
  out[0] = regs;
  out[1] = block;
  out[2] = bpsm;
  out[3] = local;
  return 0;
}

int chi_b200_problem_create(int model, int device, chi_b200_problem** out) {
  // model == -1: population-only problem (stand-alone population models)
  if (model < -1 || model >= chi_b200_n_models())
    return fail(-1, "model id out of range");
  int n_dev = chi_b200_device_count();
  if (n_dev <= 0)
    return fail(-3, "no CUDA device available: chi_b200 has no CPU fallback");
  if (device < 0 || device >= n_dev) return fail(-1, "device out of range");
  CUDA_TRY(cudaSetDevice(device));
  auto* p = new chi_b200_problem();
  p->vt = model >= 0 ? registry()[model] : nullptr;
  p->model = model;
  p->device = device;
  cudaError_t err = cudaStreamCreateWithFlags(&p->stream, cudaStreamNonBlocking);
  if (err != cudaSuccess) {
    delete p;
    return fail(-2, cudaGetErrorString(err));
  }
  if (!p->vt) {
    *out = p;
    return 0;
  }
  // defaults: all states as outputs are not assumed; caller sets outputs.
  p->n_cols = p->vt->n_states + p->vt->n_consts;
  if (p->n_cols > MAX_COLS) {
    delete p;
    return fail(-1, "too many mechanistic parameters");
  }
  for (int i = 0; i < p->n_cols; ++i) p->col_param[i] = i;
  *out = p;
  return 0;
}

void chi_b200_problem_destroy(chi_b200_problem* p) {
  if (!p) return;
  cudaSetDevice(p->device);
  for (DevBuf* b :
       {&p->snap, &p->rec_off, &p->rec_t, &p->rec_obs, &p->rec_logobs, &p->rec_out,
        &p->seg_off, &p->seg_t, &p->seg_rate, &p->covariates, &p->x, &p->ll,
        &p->grad, &p->status, &p->nst, &p->nrj, &p->ids, &p->traj_off,
        &p->traj_y, &p->traj_s, &p->hx, &p->hscore, &p->hgrad, &p->hstatus,
        &p->partial, &p->popflag, &p->r2t, &p->order_perm, &p->order_keys})
    b->release();
  if (p->comm) nccl_api().CommDestroy(p->comm);
  p->pack.release();
  p->step_sums.release();
  if (p->ev0) cudaEventDestroy(p->ev0);
  if (p->ev1) cudaEventDestroy(p->ev1);
  if (p->stream) cudaStreamDestroy(p->stream);
  if (p->stream2) cudaStreamDestroy(p->stream2);
  if (p->stream_in) cudaStreamDestroy(p->stream_in);
  if (p->stream_out) cudaStreamDestroy(p->stream_out);
  for (auto& e : p->ev_in) cudaEventDestroy(e);
  for (auto& e : p->ev_comp) cudaEventDestroy(e);
  p->work_counters.release();
  delete p;
}

int chi_b200_problem_set_outputs(chi_b200_problem* p, int32_t n_out,
                                 const int32_t* output_ids,
                                 const int32_t* err_kinds) {
  if (!p->vt) return fail(-1, "problem has no mechanistic model");
  if (n_out < 0 || n_out > MAX_OUT) return fail(-1, "too many outputs");
  int off = 0;
  for (int o = 0; o < n_out; ++o) {
    if (output_ids[o] < 0 || output_ids[o] >= p->vt->n_outputs)
      return fail(-1, "output id out of range");
    const int kind = err_kinds ? err_kinds[o] : ERR_GAUSSIAN;
    if (kind < 0 || kind > ERR_MULTIPLICATIVE)
      return fail(-1, "unknown error model kind");
    p->out_id[o] = output_ids[o];
    p->err_kind[o] = kind;
    p->err_off[o] = off;
    off += (kind == ERR_CAM) ? 2 : 1;
  }
  if (err_kinds && off > MAX_SIGMA)
    return fail(-1, "too many error-model parameters");
  p->n_out = n_out;
  p->n_sigma = err_kinds ? off : 0;
  return 0;
}

int chi_b200_problem_set_sensitivity_columns(chi_b200_problem* p,
                                             int32_t n_cols,
                                             const int32_t* param_idx) {
  if (!p->vt) return fail(-1, "problem has no mechanistic model");
  const int P = p->vt->n_states + p->vt->n_consts;
  if (n_cols < 0 || n_cols > P || n_cols > MAX_COLS)
    return fail(-1, "invalid number of sensitivity columns");
  for (int i = 0; i < n_cols; ++i) {
    if (param_idx[i] < 0 || param_idx[i] >= P)
      return fail(-1, "sensitivity parameter index out of range");
    p->col_param[i] = param_idx[i];
  }
  p->n_cols = n_cols;
  return 0;
}

int chi_b200_problem_set_tolerances(chi_b200_problem* p, double rtol,
                                    double atol, int32_t max_steps) {
  if (!(rtol > 0.0) || !(atol > 0.0) || max_steps <= 0)
    return fail(-1, "tolerances must be positive");
  p->rtol = rtol;
  p->atol = atol;
  p->max_steps = max_steps;
  return 0;
}

int chi_b200_problem_set_n_ids(chi_b200_problem* p, int64_t n_ids) {
  if (p->vt) return fail(-1, "n_ids follows from the uploaded data");
  if (n_ids <= 0) return fail(-1, "n_ids must be positive");
  p->n_ids = n_ids;
  p->id_begin = 0;
  p->id_end = n_ids;
  return 0;
}

int chi_b200_problem_set_fixed_parameters(chi_b200_problem* p,
                                          int32_t n_fixed,
                                          const int32_t* idx,
                                          const double* values) {
  if (!p->vt) return fail(-1, "problem has no mechanistic model");
  const int nd = p->vt->n_states + p->vt->n_consts + p->n_sigma;
  if (n_fixed < 0 || n_fixed >= nd || n_fixed > MAX_POP_DIM)
    return fail(-1, "invalid number of fixed parameters");
  std::vector<int32_t> fi(idx, idx + n_fixed);
  for (int f = 0; f < n_fixed; ++f) {
    if (fi[f] < 0 || fi[f] >= nd)
      return fail(-1, "fixed parameter index out of range");
    if (f > 0 && fi[f] <= fi[f - 1])
      return fail(-1, "fixed parameter indices must be increasing");
  }
  p->fixed_idx = fi;
  p->fixed_val.assign(values, values + n_fixed);
  p->has_pop = false;  // the population layout depends on the fixed set
  return 0;
}

int chi_b200_problem_set_variant(chi_b200_problem* p, int32_t variant) {
  if (!p->vt) return fail(-1, "problem has no mechanistic model");
  if (variant >= p->vt->n_variants) return fail(-1, "variant out of range");
  p->variant = variant;
  return 0;
}

int chi_b200_problem_set_shard(chi_b200_problem* p, int64_t id_begin,
                               int64_t id_end) {
  if (id_begin < 0 || id_end > p->n_ids || id_begin > id_end)
    return fail(-1, "invalid shard");
  p->id_begin = id_begin;
  p->id_end = id_end;
  return 0;
}

int chi_b200_comm_unique_id(uint8_t* id, int32_t capacity) {
  NcclApi& api = nccl_api();
  if (!api.ok) return fail(-4, api.error);
  if (capacity < static_cast<int32_t>(sizeof(ncclUniqueId)))
    return fail(-1, "unique id buffer too small");
  ncclUniqueId uid;
  NCCL_TRY(api.GetUniqueId(&uid));
  std::memcpy(id, &uid, sizeof(uid));
  return 0;
}

int chi_b200_problem_comm_init(chi_b200_problem* p, int32_t world,
                               int32_t rank, const uint8_t* id) {
  NcclApi& api = nccl_api();
  if (!api.ok) return fail(-4, api.error);
  if (world < 1 || rank < 0 || rank >= world)
    return fail(-1, "invalid rank / world size");
  CUDA_TRY(cudaSetDevice(p->device));
  if (p->comm) {
    NCCL_TRY(api.CommDestroy(p->comm));
    p->comm = nullptr;
  }
  ncclUniqueId uid;
  std::memcpy(&uid, id, sizeof(uid));
  NCCL_TRY(api.CommInitRank(&p->comm, world, uid, rank));
  p->comm_world = world;
  p->comm_rank = rank;
  return 0;
}

int chi_b200_problem_comm_destroy(chi_b200_problem* p) {
  if (p->comm) {
    CUDA_TRY(cudaSetDevice(p->device));
    NCCL_TRY(nccl_api().CommDestroy(p->comm));
    p->comm = nullptr;
  }
  p->comm_world = 1;
  p->comm_rank = 0;
  return 0;
}

int chi_b200_problem_set_timing(chi_b200_problem* p, int32_t enabled) {
  CUDA_TRY(cudaSetDevice(p->device));
  if (enabled && !p->ev0) {
    CUDA_TRY(cudaEventCreate(&p->ev0));
    CUDA_TRY(cudaEventCreate(&p->ev1));
  }
  p->timing = enabled != 0;
  p->solve_ms_sum = 0.0;
  p->solve_launches = 0;
  return 0;
}

int chi_b200_problem_set_snapshot_budget(chi_b200_problem* p, int64_t bytes) {
  p->snap_budget = bytes;
  p->snap_budget_auto = false;
  if (bytes == 0) p->snap.release();
  return 0;
}

int chi_b200_problem_set_batching(chi_b200_problem* p, int32_t n_chunks,
                                  int32_t order_min_vectors,
                                  int32_t refill_wait) {
  p->cfg_chunks = n_chunks;
  p->cfg_order = order_min_vectors;
  p->cfg_wait = refill_wait;
  return 0;
}

int chi_b200_problem_set_small_batch(chi_b200_problem* p,
                                     int64_t max_systems) {
  p->cfg_split_max = max_systems;
  return 0;
}

int chi_b200_problem_timings(chi_b200_problem* p, double out[2]) {
  out[0] = p->solve_ms_sum;
  out[1] = static_cast<double>(p->solve_launches);
  return 0;
}

int chi_b200_problem_counters(chi_b200_problem* p, int64_t out[4]) {
  if (p->counters_stale && p->nst.ptr && p->nrj.ptr && p->counters[0] > 0) {
    // accepted / rejected steps of the last hierarchical evaluation, summed
    // over its systems (for the flop model of the bench)
    CUDA_TRY(cudaSetDevice(p->device));
    cudaStream_t s = p->stream;
    CUDA_TRY(p->step_sums.reserve(2 * sizeof(unsigned long long)));
    unsigned long long* d_cnt = p->step_sums.as<unsigned long long>();
    CUDA_TRY(cudaMemsetAsync(d_cnt, 0, 2 * sizeof(unsigned long long), s));
    count_kernel<<<148, 256, 0, s>>>(p->counters[0], p->nst.as<int32_t>(),
                                     p->nrj.as<int32_t>(), d_cnt);
    unsigned long long h_cnt[2] = {0, 0};
    CUDA_TRY(cudaMemcpyAsync(h_cnt, d_cnt, sizeof(h_cnt),
                             cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    p->counters[1] = static_cast<int64_t>(h_cnt[0]);
    p->counters[2] = static_cast<int64_t>(h_cnt[1]);
    p->counters_stale = false;
  }
  for (int i = 0; i < 4; ++i) out[i] = p->counters[i];
  return 0;
}

}  // extern "C"

// Expands (level, start, duration, period, multiplier) events into the
// piecewise-constant dose rate of one individual on [0, t_end]
// (semantics of myokit.pacing.blocktrain as used at
// chi/_mechanistic_models.py:853-855).
static void expand_events(const double* ev, int64_t n_ev, double t_end,
                          std::vector<double>& seg_t,
                          std::vector<double>& seg_rate) {
  struct Pulse {
    double on, off, level;
  };
  std::vector<Pulse> pulses;
  for (int64_t k = 0; k < n_ev; ++k) {
    const double level = ev[5 * k], start = ev[5 * k + 1],
                 dur = ev[5 * k + 2], period = ev[5 * k + 3],
                 mult = ev[5 * k + 4];
    if (period == 0.0) {
      pulses.push_back({start, start + dur, level});
      continue;
    }
    for (int64_t j = 0;; ++j) {
      const double on = start + static_cast<double>(j) * period;
      if (on >= t_end || (mult > 0 && j >= static_cast<int64_t>(mult))) break;
      pulses.push_back({on, on + dur, level});
    }
  }
  std::vector<double> edges;
  edges.push_back(0.0);
  for (auto& q : pulses) {
    if (q.on < t_end) edges.push_back(std::max(q.on, 0.0));
    if (q.off < t_end) edges.push_back(std::max(q.off, 0.0));
  }
  std::sort(edges.begin(), edges.end());
  edges.erase(std::unique(edges.begin(), edges.end()), edges.end());
  for (size_t k = 0; k < edges.size(); ++k) {
    const double a = edges[k];
    const double b = (k + 1 < edges.size()) ? edges[k + 1] : t_end;
    const double mid = 0.5 * (a + std::max(b, a));
    double rate = 0.0;
    for (auto& q : pulses)
      if (q.on <= mid && mid < q.off) rate += q.level;
    seg_t.push_back(a);
    seg_rate.push_back(rate);
  }
}

template <class T>
static cudaError_t upload(DevBuf& buf, const std::vector<T>& v) {
  cudaError_t err = buf.reserve(std::max<size_t>(v.size(), 1) * sizeof(T));
  if (err != cudaSuccess) return err;
  if (v.empty()) return cudaSuccess;
  return cudaMemcpy(buf.ptr, v.data(), v.size() * sizeof(T),
                    cudaMemcpyHostToDevice);
}

extern "C" {

int chi_b200_problem_upload(chi_b200_problem* p, int64_t n_ids,
                            const int64_t* rec_off, const double* rec_t,
                            const double* rec_obs, const int32_t* rec_out,
                            const int64_t* dose_off,
                            const double* dose_events) {
  if (!p->vt) return fail(-1, "problem has no mechanistic model");
  if (n_ids <= 0) return fail(-1, "n_ids must be positive");
  if (p->n_out <= 0) return fail(-1, "set outputs before uploading data");
  CUDA_TRY(cudaSetDevice(p->device));
  const int64_t n_rec = rec_off[n_ids];
  std::vector<int64_t> off(rec_off, rec_off + n_ids + 1);
  std::vector<double> t(rec_t, rec_t + n_rec);
  std::vector<double> obs(n_rec, 0.0), logobs(n_rec, 0.0);
  std::vector<int32_t> out(n_rec, 0);
  for (int64_t i = 0; i < n_ids; ++i) {
    if (off[i + 1] < off[i]) return fail(-1, "rec_off must be non-decreasing");
    for (int64_t r = off[i]; r < off[i + 1]; ++r) {
      if (t[r] < 0.0) return fail(-1, "Times cannot be negative.");
      if (r > off[i] && t[r] < t[r - 1])
        return fail(-1, "Times must be increasing.");
      const int32_t o = rec_out ? rec_out[r] : 0;
      if (o < 0 || o >= p->n_out) return fail(-1, "rec_out out of range");
      out[r] = o;
      if (rec_obs) {
        obs[r] = rec_obs[r];
        logobs[r] = std::log(rec_obs[r]);
      }
    }
  }
  std::vector<int64_t> seg_off(n_ids + 1, 0);
  std::vector<double> seg_t, seg_rate;
  for (int64_t i = 0; i < n_ids; ++i) {
    const double t_end = off[i + 1] > off[i] ? t[off[i + 1] - 1] : 0.0;
    const int64_t e0 = dose_off ? dose_off[i] : 0;
    const int64_t e1 = dose_off ? dose_off[i + 1] : 0;
    expand_events(dose_events ? dose_events + 5 * e0 : nullptr, e1 - e0,
                  t_end, seg_t, seg_rate);
    seg_off[i + 1] = static_cast<int64_t>(seg_t.size());
  }
  // the streamed kernel keeps record / segment indices as 32-bit integers
  if (n_rec > INT32_MAX || static_cast<int64_t>(seg_t.size()) > INT32_MAX)
    return fail(-1, "more than 2^31 - 1 observation records or dose segments");
  CUDA_TRY(upload(p->rec_off, off));
  CUDA_TRY(upload(p->rec_t, t));
  CUDA_TRY(upload(p->rec_obs, obs));
  CUDA_TRY(upload(p->rec_logobs, logobs));
  CUDA_TRY(upload(p->rec_out, out));
  CUDA_TRY(upload(p->seg_off, seg_off));
  CUDA_TRY(upload(p->seg_t, seg_t));
  CUDA_TRY(upload(p->seg_rate, seg_rate));
  p->n_ids = n_ids;
  p->n_rec = n_rec;
  p->h_rec_off = off;
  p->max_rec = 0;
  for (int64_t i = 0; i < n_ids; ++i)
    p->max_rec = std::max(p->max_rec, off[i + 1] - off[i]);
  p->id_begin = 0;
  p->id_end = n_ids;
  return 0;
}

}  // extern "C"

// ---- launch helpers ---------------------------------------------------------
namespace {

void fill_args(const chi_b200_problem* p, SolveArgs& a) {
  std::memset(&a, 0, sizeof(a));
  a.n_ids = static_cast<int32_t>(p->n_ids);
  a.rec_off = p->rec_off.as<int64_t>();
  a.rec_t = p->rec_t.as<double>();
  a.rec_obs = p->rec_obs.as<double>();
  a.rec_logobs = p->rec_logobs.as<double>();
  a.rec_out = p->rec_out.as<int32_t>();
  a.seg_off = p->seg_off.as<int64_t>();
  a.seg_t = p->seg_t.as<double>();
  a.seg_rate = p->seg_rate.as<double>();
  a.n_out = p->n_out;
  for (int o = 0; o < p->n_out; ++o) {
    a.out_id[o] = p->out_id[o];
    a.err_kind[o] = p->err_kind[o];
    a.err_off[o] = p->err_off[o];
  }
  a.n_sigma = p->n_sigma;
  a.n_cols = p->n_cols;
  for (int i = 0; i < MAX_COLS; ++i) a.param_col[i] = -1;
  a.n_int = 0;
  for (int i = 0; i < p->n_cols; ++i) {
    const int pk = p->col_param[i];
    a.col_param[i] = pk;
    if (pk >= 0 && pk < MAX_COLS) a.param_col[pk] = i;
    const int N = p->vt->n_states;
    if (!(pk >= N && ((p->vt->inert_mask >> (pk - N)) & 1u))) {
      a.int_col[a.n_int] = i;
      a.int_param[a.n_int] = pk;
      ++a.n_int;
    }
  }
  a.rtol = p->rtol;
  a.atol = p->atol;
  a.max_steps = p->max_steps;
}

int launch_loglik(chi_b200_problem* p, int64_t B, const int32_t* d_ids,
                  const double* d_x, int with_grad, double* d_ll,
                  double* d_grad, int32_t* d_status, int32_t* d_nst,
                  int32_t* d_nrj, cudaStream_t stream,
                  int32_t order_inner = 0, int32_t order_stride = 0,
                  int64_t b0 = 0, int64_t B_res = 0,
                  unsigned long long* work_counter = nullptr) {
  // b0 / B_res: this launch covers systems [b0, b0 + B) of a batch of B_res
  // systems whose scratch (the snapshots here) is shared by several launches
  if (B_res < b0 + B) B_res = b0 + B;
  if (!p->vt) return fail(-1, "problem has no mechanistic model");
  if (p->n_ids <= 0) return fail(-1, "no population data uploaded");
  const int P = p->vt->n_states + p->vt->n_consts;
  SolveArgs a;
  fill_args(p, a);
  a.B = B;
  a.ids = d_ids;
  a.x = d_x;
  a.x_stride = P + p->n_sigma;
  a.mode = 1;
  a.with_sens = with_grad ? 1 : 0;
  a.ll = d_ll;
  a.grad = d_grad;
  a.grad_stride = P + p->n_sigma;
  a.status = d_status;
  a.n_steps = d_nst;
  a.n_rej = d_nrj;
  a.snap = nullptr;
  a.snap_rows = 0;
  a.snap_width = 0;
  // the persistent kernel's work counter belongs to the problem (the last
  // slot: the un-chunked launches; the others: the chunks of
  // chi_b200_hier_logpdf, which
  // may be in flight together)
  CUDA_TRY(p->work_counters.reserve(MAX_CHUNKS_P1 * sizeof(unsigned long long)));
  a.work_counter = work_counter
                       ? work_counter
                       : p->work_counters.as<unsigned long long>() + MAX_CHUNKS;
  bool ordered = false;
  {
    // Work order of the persistent kernel (SolveArgs::order_inner): with at
    // least a warp's worth of parameter vectors the lanes of a warp take the
    // same individual under different vectors and refill as a whole warp, so
    // they reach their stops in phase.  Measured on the bench workload
    // (1024 vectors x 1000 individuals) at relative parameter spreads of
    // 1 / 10 / 20 %: +8.0 / +7.3 / +5.2 % evaluations per second over
    // system order with per-lane refill; waiting a bounded number of
    // iterations (4, 12) is worse than either extreme once the spread
    // exceeds a few percent.  Tuning knobs: CHI_B200_ORDER = smallest number
    // of vectors for which the order applies (0: never),
    // CHI_B200_REFILL_WAIT = iterations an idle lane waits for its warp.
    static const int min_inner = [] {
      const char* e = std::getenv("CHI_B200_ORDER");
      return e ? std::atoi(e) : 32;
    }();
    static const int wait_env = [] {
      const char* e = std::getenv("CHI_B200_REFILL_WAIT");
      return e ? std::atoi(e) : -1;
    }();
    const int min_vectors = p->cfg_order >= 0 ? p->cfg_order : min_inner;
    const int wait = p->cfg_wait >= 0 ? p->cfg_wait : wait_env;
    // (the map is a permutation of [0, B) only for a full C x n layout)
    ordered = min_vectors > 0 && order_inner >= min_vectors &&
              static_cast<int64_t>(order_inner) * order_stride == B;
    if (ordered) {
      a.order_inner = order_inner;
      a.order_stride = order_stride;
    }
    a.refill_wait = wait >= 0 ? wait : (ordered ? INT32_MAX : 0);
    if (ordered) {
      // scratch of the stiffness order of the warp-synchronous kernels
      // (SolveArgs::perm)
      CUDA_TRY(p->order_perm.reserve(B_res * sizeof(int32_t)));
      CUDA_TRY(p->order_keys.reserve(B_res * sizeof(float)));
      a.perm = p->order_perm.as<int32_t>() + b0;
      a.order_keys = p->order_keys.as<float>() + b0;
    }
  }
  // Small launches (single parameter vectors: what pints' samplers and
  // optimisers ask for, chi/_inference.py:729-747): one WARP per system
  // (SolveArgs::split).  The sensitivity columns of the system advance side
  // by side on the first lanes of the warp, each lane with the state and one
  // column under its own step-size control, and the warp walks through the
  // system's stops like a warp of the batched launch does -- a third of the
  // instructions per step attempt, no lane waits for another system's events,
  // on a GPU that is mostly idle at this size.  ROSL kernels of linear
  // models; applies while every system finds a resident warp -- 12 per SM
  // at the 165 registers this form is compiled for
  // (chi_b200_problem_set_small_batch / CHI_B200_SPLIT_MAX: largest number of
  // systems, 0: never).
  bool split = false;
  {
    static const int64_t split_env = [] {
      const char* e = std::getenv("CHI_B200_SPLIT_MAX");
      return e ? std::atoll(e) : -1ll;
    }();
    int64_t split_max = p->cfg_split_max >= 0 ? p->cfg_split_max : split_env;
    if (split_max < 0) {
      int sms = 0;
      CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount,
                                      p->device));
      // (three resident blocks of four warps per SM: one wave)
      split_max = 12ll * sms;
    }
    const int v = p->variant >= 0 ? p->variant : 0;
    const bool rosl = v < p->vt->n_variants &&
                      p->vt->variant_lanes[v] == 0 &&
                      (p->vt->variant_cols[v] == 9 ||
                       p->vt->variant_cols[v] == 11) &&
                      (p->vt->linear & 1);
    split = rosl && !ordered && B <= split_max && (!with_grad || d_grad) &&
            (!with_grad || a.n_int <= 32);
    if (split) {
      a.split = with_grad && a.n_int > 1 ? a.n_int : 1;
      CUDA_TRY(cudaMemsetAsync(d_status, 0, B * sizeof(int32_t), stream));
    }
  }
  if (!split) {
    // Observation (output map, error model, gradient) fused in the solve
    // kernel or deferred to observe_kernel via snapshots.  Default policy:
    // fused when the lanes of a warp share the individual (they reach the
    // records together, the handler runs with about half a warp: on par with
    // the deferred form, without 0.98 GB of snapshots per 1.02 M systems and
    // the second kernel); deferred otherwise (the handler would run with one
    // or two active lanes), when the snapshots fit half of the free memory.
    // An explicit budget (chi_b200_problem_set_snapshot_budget /
    // CHI_B200_SNAP_MB) overrides the policy: 0 = always fused, > 0 = deferred
    // whenever the snapshots fit.
    const bool streamed =
        p->variant < 0 || (p->variant < p->vt->n_variants &&
                           p->vt->variant_lanes[p->variant] == 0);
    bool explicit_budget = p->snap_budget >= 0 && !p->snap_budget_auto;
    if (p->snap_budget < 0) {
      const char* env = std::getenv("CHI_B200_SNAP_MB");
      if (env) {
        p->snap_budget = static_cast<int64_t>(std::atof(env) * 1048576.0);
        explicit_budget = true;
      } else {
        size_t free_b = 0, total_b = 0;
        CUDA_TRY(cudaMemGetInfo(&free_b, &total_b));
        p->snap_budget = static_cast<int64_t>(free_b / 2);
        p->snap_budget_auto = true;
      }
    }
    const bool want_snap = explicit_budget || !ordered;
    const int N = p->vt->n_states;
    const int integrated = with_grad ? a.n_int : 0;
    const int64_t width = static_cast<int64_t>(N) * (1 + integrated);
    const int64_t bytes = B_res * p->max_rec * width * 8;
    if (streamed && want_snap && p->max_rec > 0 && p->max_rec <= INT32_MAX &&
        bytes <= p->snap_budget) {
      CUDA_TRY(p->snap.reserve(static_cast<size_t>(bytes)));
      a.snap = p->snap.as<double>() + b0 * p->max_rec * width;
      a.snap_rows = static_cast<int32_t>(p->max_rec);
      a.snap_width = static_cast<int32_t>(width);
    }
  }
  if (with_grad && p->n_cols < P) {
    // columns without sensitivities read as zero gradient
    const int64_t n = B * a.grad_stride;
    fill_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, stream>>>(
        d_grad, n, 0.0);
  }
  if (p->timing) CUDA_TRY(cudaEventRecord(p->ev0, stream));
  CUDA_TRY(p->vt->launch(a, p->variant, stream));
  if (p->timing) {
    CUDA_TRY(cudaEventRecord(p->ev1, stream));
    CUDA_TRY(cudaEventSynchronize(p->ev1));
    float ms = 0.f;
    CUDA_TRY(cudaEventElapsedTime(&ms, p->ev0, p->ev1));
    p->solve_ms_sum += ms;
    p->solve_launches += 1;
  }
  p->counters[3] += a.snap ? 2 : 1;
  {
    // the two kernels of the stiffness order (Registrar::launch_stream runs
    // them under exactly these conditions)
    static const bool sort_enabled = [] {
      const char* e = std::getenv("CHI_B200_SORT");
      return !(e && std::atoi(e) == 0);
    }();
    static const bool warp_enabled = [] {
      const char* e = std::getenv("CHI_B200_WARP_SYNC");
      return !(e && std::atoi(e) == 0);
    }();
    const int v = p->variant >= 0 ? p->variant : 0;
    const bool rosl = v < p->vt->n_variants &&
                      p->vt->variant_lanes[v] == 0 &&
                      (p->vt->variant_cols[v] == 9 ||
                       p->vt->variant_cols[v] == 11) &&
                      (p->vt->linear & 1);
    const int64_t padded = (static_cast<int64_t>(a.order_inner) + 31) / 32 * 32;
    if (rosl && sort_enabled && warp_enabled && a.split == 0 && !a.snap &&
        a.perm && a.refill_wait == INT32_MAX && a.order_inner > 32 &&
        a.order_inner <= 4096 &&
        (padded - a.order_inner) * 8 <= a.order_inner &&
        static_cast<int64_t>(a.order_inner) * a.order_stride == a.B)
      p->counters[3] += 2;
  }
  return 0;
}

// DFMA micro-benchmark: 16 independent FMA chains per thread, fully unrolled
// (8 resident warps per scheduler x 16 chains cover the 8-cycle latency of
// the pipe many times over: the loop is bound by DFMA issue alone).
__global__ void __launch_bounds__(256)
fp64_peak_kernel(double* out, int iters, double a, double b) {
  double x[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) x[k] = threadIdx.x + k;
#pragma unroll 1
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
#pragma unroll
      for (int k = 0; k < 16; ++k) x[k] = fma(x[k], a, b);
    }
  }
  double s = 0.0;
#pragma unroll
  for (int k = 0; k < 16; ++k) s += x[k];
  if (s == 123.456) out[0] = s;
}

}  // namespace

extern "C" {

int chi_b200_loglik_dev(chi_b200_problem* p, int64_t B, const int32_t* d_ids,
                        const double* d_x, int32_t with_grad, double* d_ll,
                        double* d_grad, int32_t* d_status, int32_t* d_n_steps,
                        int32_t* d_n_rej, void* stream) {
  CUDA_TRY(cudaSetDevice(p->device));
  p->counters[3] = 0;
  return launch_loglik(p, B, d_ids, d_x, with_grad, d_ll, d_grad, d_status,
                       d_n_steps, d_n_rej, static_cast<cudaStream_t>(stream));
}

int chi_b200_loglik(chi_b200_problem* p, int64_t B, const int32_t* ids,
                    const double* x, int32_t with_grad, double* ll,
                    double* grad, int32_t* status, int32_t* n_steps,
                    int32_t* n_rej) {
  if (B <= 0) return 0;
  if (!p->vt) return fail(-1, "problem has no mechanistic model");
  CUDA_TRY(cudaSetDevice(p->device));
  const int nd = p->vt->n_states + p->vt->n_consts + p->n_sigma;
  cudaStream_t s = p->stream;
  CUDA_TRY(p->x.reserve(B * nd * sizeof(double)));
  CUDA_TRY(p->ll.reserve(B * sizeof(double)));
  CUDA_TRY(p->status.reserve(B * sizeof(int32_t)));
  CUDA_TRY(p->nst.reserve(B * sizeof(int32_t)));
  CUDA_TRY(p->nrj.reserve(B * sizeof(int32_t)));
  if (with_grad) CUDA_TRY(p->grad.reserve(B * nd * sizeof(double)));
  if (ids) {
    CUDA_TRY(p->ids.reserve(B * sizeof(int32_t)));
    CUDA_TRY(cudaMemcpyAsync(p->ids.ptr, ids, B * sizeof(int32_t),
                             cudaMemcpyHostToDevice, s));
  }
  CUDA_TRY(cudaMemcpyAsync(p->x.ptr, x, B * nd * sizeof(double),
                           cudaMemcpyHostToDevice, s));
  p->counters[3] = 0;
  int rc = launch_loglik(p, B, ids ? p->ids.as<int32_t>() : nullptr,
                         p->x.as<double>(), with_grad, p->ll.as<double>(),
                         with_grad ? p->grad.as<double>() : nullptr,
                         p->status.as<int32_t>(), p->nst.as<int32_t>(),
                         p->nrj.as<int32_t>(), s);
  if (rc) return rc;
  CUDA_TRY(cudaMemcpyAsync(ll, p->ll.ptr, B * sizeof(double),
                           cudaMemcpyDeviceToHost, s));
  if (with_grad && grad)
    CUDA_TRY(cudaMemcpyAsync(grad, p->grad.ptr, B * nd * sizeof(double),
                             cudaMemcpyDeviceToHost, s));
  if (status)
    CUDA_TRY(cudaMemcpyAsync(status, p->status.ptr, B * sizeof(int32_t),
                             cudaMemcpyDeviceToHost, s));
  if (n_steps)
    CUDA_TRY(cudaMemcpyAsync(n_steps, p->nst.ptr, B * sizeof(int32_t),
                             cudaMemcpyDeviceToHost, s));
  if (n_rej)
    CUDA_TRY(cudaMemcpyAsync(n_rej, p->nrj.ptr, B * sizeof(int32_t),
                             cudaMemcpyDeviceToHost, s));
  CUDA_TRY(cudaStreamSynchronize(s));
  p->counters[0] = B;
  p->counters_stale = true;
  return 0;
}

int chi_b200_simulate(chi_b200_problem* p, int64_t B, const int32_t* ids,
                      const double* psi, int32_t with_sens, double* y_out,
                      double* s_out, int32_t* status) {
  if (B <= 0) return 0;
  if (!p->vt) return fail(-1, "problem has no mechanistic model");
  if (p->n_ids <= 0) return fail(-1, "no population data uploaded");
  CUDA_TRY(cudaSetDevice(p->device));
  const int P = p->vt->n_states + p->vt->n_consts;
  cudaStream_t s = p->stream;
  std::vector<int64_t> toff(B + 1, 0);
  for (int64_t b = 0; b < B; ++b) {
    const int64_t id = ids ? ids[b] : b % p->n_ids;
    if (id < 0 || id >= p->n_ids) return fail(-1, "individual id out of range");
    toff[b + 1] = toff[b] + (p->h_rec_off[id + 1] - p->h_rec_off[id]);
  }
  const int64_t rows = toff[B];
  CUDA_TRY(p->x.reserve(B * P * sizeof(double)));
  CUDA_TRY(p->status.reserve(B * sizeof(int32_t)));
  CUDA_TRY(p->traj_off.reserve((B + 1) * sizeof(int64_t)));
  CUDA_TRY(p->traj_y.reserve(std::max<int64_t>(rows, 1) * sizeof(double)));
  if (with_sens)
    CUDA_TRY(p->traj_s.reserve(std::max<int64_t>(rows * p->n_cols, 1) *
                               sizeof(double)));
  if (ids) {
    CUDA_TRY(p->ids.reserve(B * sizeof(int32_t)));
    CUDA_TRY(cudaMemcpyAsync(p->ids.ptr, ids, B * sizeof(int32_t),
                             cudaMemcpyHostToDevice, s));
  }
  CUDA_TRY(cudaMemcpyAsync(p->x.ptr, psi, B * P * sizeof(double),
                           cudaMemcpyHostToDevice, s));
  CUDA_TRY(cudaMemcpyAsync(p->traj_off.ptr, toff.data(),
                           (B + 1) * sizeof(int64_t), cudaMemcpyHostToDevice,
                           s));
  SolveArgs a;
  fill_args(p, a);
  a.B = B;
  a.ids = ids ? p->ids.as<int32_t>() : nullptr;
  a.x = p->x.as<double>();
  a.x_stride = P;
  a.mode = 0;
  a.with_sens = with_sens ? 1 : 0;
  a.traj_off = p->traj_off.as<int64_t>();
  a.traj_y = p->traj_y.as<double>();
  a.traj_s = p->traj_s.as<double>();
  a.status = p->status.as<int32_t>();
  CUDA_TRY(p->nst.reserve(B * sizeof(int32_t)));
  CUDA_TRY(p->nrj.reserve(B * sizeof(int32_t)));
  a.n_steps = p->nst.as<int32_t>();
  a.n_rej = p->nrj.as<int32_t>();
  CUDA_TRY(p->work_counters.reserve(MAX_CHUNKS_P1 * sizeof(unsigned long long)));
  a.work_counter = p->work_counters.as<unsigned long long>() + MAX_CHUNKS;
  if (p->timing) CUDA_TRY(cudaEventRecord(p->ev0, s));
  CUDA_TRY(p->vt->launch(a, p->variant, s));
  if (p->timing) {
    CUDA_TRY(cudaEventRecord(p->ev1, s));
    CUDA_TRY(cudaEventSynchronize(p->ev1));
    float ms = 0.f;
    CUDA_TRY(cudaEventElapsedTime(&ms, p->ev0, p->ev1));
    p->solve_ms_sum += ms;
    p->solve_launches += 1;
  }
  if (rows > 0) {
    CUDA_TRY(cudaMemcpyAsync(y_out, p->traj_y.ptr, rows * sizeof(double),
                             cudaMemcpyDeviceToHost, s));
    if (with_sens && s_out)
      CUDA_TRY(cudaMemcpyAsync(s_out, p->traj_s.ptr,
                               rows * p->n_cols * sizeof(double),
                               cudaMemcpyDeviceToHost, s));
  }
  if (status)
    CUDA_TRY(cudaMemcpyAsync(status, p->status.ptr, B * sizeof(int32_t),
                             cudaMemcpyDeviceToHost, s));
  CUDA_TRY(cudaStreamSynchronize(s));
  p->counters[0] = B;
  p->counters[3] = 1;
  p->counters_stale = true;
  return 0;
}

int chi_b200_error_model(int32_t device, int32_t kind, const double* sigma,
                         int64_t n_obs, int32_t n_par, const double* ybar,
                         const double* sens, const double* obs,
                         int32_t with_grad, double* out, double* pointwise) {
  if (chi_b200_device_count() <= 0)
    return fail(-3, "no CUDA device available: chi_b200 has no CPU fallback");
  if (kind < 0 || kind > ERR_MULTIPLICATIVE)
    return fail(-1, "unknown error model kind");
  if (n_obs * sizeof(double) > 200 * 1024)
    return fail(-1, "too many observations for the stand-alone error model");
  CUDA_TRY(cudaSetDevice(device));
  const int n_sig = kind == ERR_CAM ? 2 : 1;
  const int n_out = 1 + (with_grad ? n_par + n_sig : 0);
  // scratch kept per device across calls (grown on demand, never shrunk)
  struct Scratch {
    DevBuf y, s, o, out, pw;
  };
  static std::mutex mtx;
  static Scratch cache[16];
  if (device < 0 || device >= 16) return fail(-1, "device out of range");
  std::lock_guard<std::mutex> lock(mtx);
  Scratch& sc = cache[device];
  DevBuf &d_y = sc.y, &d_s = sc.s, &d_o = sc.o, &d_out = sc.out, &d_pw = sc.pw;
  auto cleanup = [&]() {};
  const size_t nb = std::max<int64_t>(n_obs, 1) * sizeof(double);
  cudaError_t err = d_y.reserve(nb);
  if (err == cudaSuccess) err = d_o.reserve(nb);
  if (err == cudaSuccess) err = d_out.reserve(n_out * sizeof(double));
  if (err == cudaSuccess && with_grad)
    err = d_s.reserve(std::max<size_t>(nb * n_par, 8));
  if (err == cudaSuccess && pointwise) err = d_pw.reserve(nb);
  if (err == cudaSuccess && n_obs > 0) {
    err = cudaMemcpy(d_y.ptr, ybar, n_obs * sizeof(double),
                     cudaMemcpyHostToDevice);
    if (err == cudaSuccess)
      err = cudaMemcpy(d_o.ptr, obs, n_obs * sizeof(double),
                       cudaMemcpyHostToDevice);
    if (err == cudaSuccess && with_grad && n_par > 0)
      err = cudaMemcpy(d_s.ptr, sens, n_obs * n_par * sizeof(double),
                       cudaMemcpyHostToDevice);
  }
  if (err == cudaSuccess) {
    const size_t smem = std::max<int64_t>(n_obs, 1) * sizeof(double);
    if (smem > 48 * 1024)
      err = cudaFuncSetAttribute(error_model_kernel,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 static_cast<int>(smem));
    if (err == cudaSuccess) {
      error_model_kernel<<<1, 32, smem>>>(
          kind, sigma[0], n_sig == 2 ? sigma[1] : 0.0, n_obs, n_par,
          d_y.as<double>(), d_s.as<double>(), d_o.as<double>(), with_grad,
          d_out.as<double>(), pointwise ? d_pw.as<double>() : nullptr);
      err = cudaGetLastError();
    }
  }
  if (err == cudaSuccess)
    err = cudaMemcpy(out, d_out.ptr, n_out * sizeof(double),
                     cudaMemcpyDeviceToHost);
  if (err == cudaSuccess && pointwise && n_obs > 0)
    err = cudaMemcpy(pointwise, d_pw.ptr, n_obs * sizeof(double),
                     cudaMemcpyDeviceToHost);
  cleanup();
  if (err != cudaSuccess) return fail(-2, cudaGetErrorString(err));
  return 0;
}

int chi_b200_covariate_linear(int32_t device, int64_t n_ids, int32_t n_pop,
                              int32_t n_dim, int32_t n_cov, int32_t n_sel,
                              const int32_t* pidx, const int32_t* didx,
                              const double* beta, const double* pop,
                              const double* covariates,
                              const double* dlogp_dvartheta,
                              double* vartheta, double* dpop, double* dbeta) {
  if (chi_b200_device_count() <= 0)
    return fail(-3, "no CUDA device available: chi_b200 has no CPU fallback");
  if (n_ids <= 0 || n_pop <= 0 || n_dim <= 0 || n_cov <= 0 || n_sel < 0)
    return fail(-1, "invalid covariate model sizes");
  const int n_flat = n_pop * n_dim;
  std::vector<int32_t> sel_of(n_flat, -1), flat_of(std::max(n_sel, 1), 0);
  for (int s = 0; s < n_sel; ++s) {
    if (pidx[s] < 0 || pidx[s] >= n_pop || didx[s] < 0 || didx[s] >= n_dim)
      return fail(-1, "covariate selection out of range");
    flat_of[s] = pidx[s] * n_dim + didx[s];
    sel_of[flat_of[s]] = s;
  }
  CUDA_TRY(cudaSetDevice(device));
  DevBuf d_sel, d_flat, d_beta, d_pop, d_x, d_dl, d_out;
  auto done = [&](int rc) {
    for (DevBuf* b : {&d_sel, &d_flat, &d_beta, &d_pop, &d_x, &d_dl, &d_out})
      b->release();
    return rc;
  };
  auto put = [&](DevBuf& b, const void* src, size_t bytes) {
    cudaError_t err = b.reserve(std::max<size_t>(bytes, 8));
    if (err == cudaSuccess && bytes)
      err = cudaMemcpy(b.ptr, src, bytes, cudaMemcpyHostToDevice);
    return err;
  };
  cudaError_t err = put(d_sel, sel_of.data(), n_flat * sizeof(int32_t));
  if (err == cudaSuccess)
    err = put(d_flat, flat_of.data(), flat_of.size() * sizeof(int32_t));
  if (err == cudaSuccess)
    err = put(d_beta, beta, static_cast<size_t>(n_sel) * n_cov * 8);
  if (err == cudaSuccess) err = put(d_pop, pop, n_flat * 8);
  if (err == cudaSuccess)
    err = put(d_x, covariates, static_cast<size_t>(n_ids) * n_cov * 8);
  if (err != cudaSuccess) return done(fail(-2, cudaGetErrorString(err)));
  if (vartheta) {
    const int64_t n = n_ids * n_flat;
    err = d_out.reserve(n * 8);
    if (err != cudaSuccess) return done(fail(-2, cudaGetErrorString(err)));
    covariate_forward_kernel<<<static_cast<unsigned>((n + 255) / 256), 256>>>(
        n_ids, n_flat, n_cov, d_sel.as<int32_t>(), d_beta.as<double>(),
        d_pop.as<double>(), d_x.as<double>(), d_out.as<double>());
    err = cudaGetLastError();
    if (err == cudaSuccess)
      err = cudaMemcpy(vartheta, d_out.ptr, n * 8, cudaMemcpyDeviceToHost);
    if (err != cudaSuccess) return done(fail(-2, cudaGetErrorString(err)));
  }
  if (dlogp_dvartheta && dpop && dbeta) {
    const int n_slots = n_flat + n_sel * n_cov;
    err = put(d_dl, dlogp_dvartheta, static_cast<size_t>(n_ids) * n_flat * 8);
    if (err == cudaSuccess) err = d_out.reserve(n_slots * 8);
    if (err != cudaSuccess) return done(fail(-2, cudaGetErrorString(err)));
    covariate_backward_kernel<<<n_slots, 256>>>(
        n_ids, n_flat, n_cov, d_flat.as<int32_t>(), d_x.as<double>(),
        d_dl.as<double>(), d_out.as<double>());
    err = cudaGetLastError();
    std::vector<double> h(n_slots);
    if (err == cudaSuccess)
      err = cudaMemcpy(h.data(), d_out.ptr, n_slots * 8,
                       cudaMemcpyDeviceToHost);
    if (err != cudaSuccess) return done(fail(-2, cudaGetErrorString(err)));
    std::copy(h.begin(), h.begin() + n_flat, dpop);
    std::copy(h.begin() + n_flat, h.end(), dbeta);
  }
  return done(0);
}

// ---- population ---------------------------------------------------------------

int chi_b200_population_set(chi_b200_problem* p, int32_t n_blocks,
                            const int32_t* kind, const int32_t* n_dim,
                            const int32_t* centered, const int32_t* n_cov,
                            const int32_t* n_sel, const int32_t* sel_pidx,
                            const int32_t* sel_didx,
                            const double* covariates) {
  if (p->n_ids <= 0) return fail(-1, "upload population data first");
  CUDA_TRY(cudaSetDevice(p->device));
  PopDesc& D = p->pop;
  std::memset(&D, 0, sizeof(D));
  p->red_to_top.clear();
  int d0 = 0, h0 = 0, c0 = 0, sel0 = 0;
  int64_t top = 0;
  for (int b = 0; b < n_blocks; ++b) {
    const int nd = n_dim[b];
    if (d0 + nd > MAX_POP_DIM) return fail(-1, "too many population dims");
    const int nc = n_cov ? n_cov[b] : 0;
    const int ns = (n_sel && nc > 0) ? n_sel[b] : 0;
    if (nc > MAX_POP_COV) return fail(-1, "too many covariates in a block");
    if (nc > 0 && kind[b] != CHI_B200_POP_GAUSSIAN &&
        kind[b] != CHI_B200_POP_LOGNORMAL)
      return fail(-1, "covariate blocks must be Gaussian or LogNormal");
    for (int j = 0; j < nd; ++j) {
      PopDim& q = D.dim[d0 + j];
      q.kind = kind[b];
      q.centered = centered ? centered[b] : 1;
      q.hcol = -1;
      q.n_cov = nc;
      q.cov0 = c0;
      q.beta_mu = q.beta_sigma = -1;
      q.red_mu = q.red_sigma = q.red_beta_mu = q.red_beta_sigma = -1;
      if (kind[b] == CHI_B200_POP_POOLED) {
        q.top0 = static_cast<int32_t>(top + j);
        q.red_mu = static_cast<int32_t>(p->red_to_top.size());
        p->red_to_top.push_back(q.top0);
      } else if (kind[b] == CHI_B200_POP_HETEROGENEOUS) {
        q.top0 = static_cast<int32_t>(top);
        q.het_stride = nd;
        q.het_off = j;
      } else if (kind[b] == CHI_B200_POP_GAUSSIAN ||
                 kind[b] == CHI_B200_POP_LOGNORMAL ||
                 kind[b] == CHI_B200_POP_TRUNCATED_GAUSSIAN) {
        if (kind[b] == CHI_B200_POP_TRUNCATED_GAUSSIAN) q.centered = 1;
        q.hcol = h0 + j;
        q.top0 = static_cast<int32_t>(top + j);
        q.top1 = static_cast<int32_t>(top + nd + j);
      } else {
        return fail(-1, "unknown population block kind");
      }
    }
    if (kind[b] == CHI_B200_POP_GAUSSIAN || kind[b] == CHI_B200_POP_LOGNORMAL ||
        kind[b] == CHI_B200_POP_TRUNCATED_GAUSSIAN) {
      // reduced slots in theta order: mu_1..mu_d, sigma_1..sigma_d, betas
      for (int j = 0; j < nd; ++j) {
        D.dim[d0 + j].red_mu = static_cast<int32_t>(p->red_to_top.size());
        p->red_to_top.push_back(D.dim[d0 + j].top0);
      }
      for (int j = 0; j < nd; ++j) {
        D.dim[d0 + j].red_sigma = static_cast<int32_t>(p->red_to_top.size());
        p->red_to_top.push_back(D.dim[d0 + j].top1);
      }
      for (int s = 0; s < ns; ++s) {
        const int pi = sel_pidx[sel0 + s], di = sel_didx[sel0 + s];
        if (pi < 0 || pi > 1 || di < 0 || di >= nd)
          return fail(-1, "covariate selection out of range");
        PopDim& q = D.dim[d0 + di];
        const int32_t t = static_cast<int32_t>(top + 2 * nd + s * nc);
        const int32_t rslot = static_cast<int32_t>(p->red_to_top.size());
        for (int k = 0; k < nc; ++k) p->red_to_top.push_back(t + k);
        if (pi == 0) {
          q.beta_mu = t;
          q.red_beta_mu = rslot;
        } else {
          q.beta_sigma = t;
          q.red_beta_sigma = rslot;
        }
      }
      h0 += nd;
      top += 2 * nd + static_cast<int64_t>(ns) * nc;
    } else if (kind[b] == CHI_B200_POP_POOLED) {
      top += nd;
    } else {
      top += p->n_ids * nd;
    }
    sel0 += ns;
    c0 += nc;
    d0 += nd;
  }
  D.n_dim = d0;
  D.n_hdim = h0;
  D.n_cov_total = c0;
  D.n_red = static_cast<int32_t>(p->red_to_top.size());
  p->n_top = top;
  const int n_fixed = p->vt ? static_cast<int>(p->fixed_idx.size()) : 0;
  const int nd_sys =
      p->vt ? p->vt->n_states + p->vt->n_consts + p->n_sigma : D.n_dim;
  const int nd_ll = nd_sys - n_fixed;
  if (nd_sys > MAX_POP_DIM) return fail(-1, "too many system parameters");
  D.n_sys = nd_sys;
  D.n_fixed = n_fixed;
  {
    std::vector<char> is_fixed(nd_sys, 0);
    for (int f = 0; f < n_fixed; ++f) {
      D.fixed_idx[f] = p->fixed_idx[f];
      D.fixed_val[f] = p->fixed_val[f];
      is_fixed[p->fixed_idx[f]] = 1;
    }
    int d = 0;
    for (int j = 0; j < nd_sys; ++j)
      if (!is_fixed[j] && d < MAX_POP_DIM) D.sys_pos[d++] = j;
  }
  if (D.n_dim != nd_ll)
    return fail(-1,
                "The dimension of the population model does not match the "
                "dimensions of the log-likelihood parameter space.");
  if (c0 > 0) {
    if (!covariates)
      return fail(-1,
                  "The population model needs covariates, but no covariates "
                  "have been provided.");
    std::vector<double> cv(covariates, covariates + p->n_ids * c0);
    CUDA_TRY(upload(p->covariates, cv));
  }
  CUDA_TRY(upload(p->r2t, p->red_to_top));
  p->has_pop = true;
  return 0;
}

int chi_b200_population_sizes(chi_b200_problem* p, int64_t out[4]) {
  if (!p->has_pop) return fail(-1, "no population model set");
  out[0] = p->n_ids * p->pop.n_hdim;
  out[1] = p->n_top;
  out[2] = p->pop.n_dim;
  out[3] = p->pop.n_hdim;
  return 0;
}

}  // extern "C"

namespace {

// Runs K3a -> (K1+K2) -> K3b -> K3c for C parameter vectors resident on the
// device.  `d_dlp_in` / `skip_solve` serve the stand-alone population entry.
int run_hier(chi_b200_problem* p, int32_t C, const double* d_x,
             int32_t with_grad, double* d_score, double* d_grad,
             int32_t* d_status, bool skip_solve, const double* d_dlp_in,
             double* d_eta_out, double* d_psi_out, cudaStream_t s,
             int32_t c0 = 0, int32_t C_res = 0,
             unsigned long long* work_counter = nullptr) {
  // c0 / C_res: the C vectors are vectors [c0, c0 + C) of a batch of C_res
  // that is evaluated in chunks (chi_b200_hier_logpdf); the caller's pointers
  // already point at the chunk, the scratch buffers are shared by the chunks
  // and indexed by the position in the batch.  They are sized for the whole
  // batch by the first chunk, before anything is in flight.
  if (!p->has_pop) return fail(-1, "no population model set");
  const PopDesc& D = p->pop;
  const int64_t n_sh = p->id_end - p->id_begin;
  const int64_t n_bottom = p->n_ids * D.n_hdim;
  const int64_t n_params = n_bottom + p->n_top;
  const int64_t B = n_sh * C;
  const int nd = D.n_sys;
  if (B <= 0) return 0;
  if (C_res < c0 + C) C_res = c0 + C;
  const int64_t B_res = n_sh * C_res;
  const int64_t b0 = n_sh * c0;
  CUDA_TRY(p->x.reserve(B_res * nd * sizeof(double)));
  CUDA_TRY(p->ids.reserve(B_res * sizeof(int32_t)));
  CUDA_TRY(p->ll.reserve(B_res * sizeof(double)));
  CUDA_TRY(p->status.reserve(B_res * sizeof(int32_t)));
  CUDA_TRY(p->nst.reserve(B_res * sizeof(int32_t)));
  CUDA_TRY(p->nrj.reserve(B_res * sizeof(int32_t)));
  if (with_grad) CUDA_TRY(p->grad.reserve(B_res * nd * sizeof(double)));
  const int64_t n_blk = (n_sh + POP_BLOCK - 1) / POP_BLOCK;
  const int n_slots = 1 + D.n_red;
  // partial buffer: [C][n_blk][n_slots] doubles
  const size_t partial_bytes =
      static_cast<size_t>(C_res) * n_blk * n_slots * sizeof(double);
  CUDA_TRY(p->partial.reserve(partial_bytes));
  const int32_t* d_r2t = p->r2t.as<int32_t>();
  CUDA_TRY(p->popflag.reserve(C_res * sizeof(int32_t)));
  int32_t* d_popflag = p->popflag.as<int32_t>() + c0;
  double* d_partial =
      p->partial.as<double>() + static_cast<int64_t>(c0) * n_blk * n_slots;
  int32_t* d_ids = p->ids.as<int32_t>() + b0;
  double* d_sys_ll = p->ll.as<double>() + b0;
  double* d_sys_grad = with_grad ? p->grad.as<double>() + b0 * nd : nullptr;
  int32_t* d_sys_status = p->status.as<int32_t>() + b0;
  CUDA_TRY(cudaMemsetAsync(d_popflag, 0, C * sizeof(int32_t), s));

  double* xs = d_psi_out ? d_psi_out : p->x.as<double>() + b0 * nd;
  pop_forward_kernel<<<static_cast<unsigned>((B + 255) / 256), 256, 0, s>>>(
      D, C, p->n_ids, p->id_begin, p->id_end, n_params, n_bottom, d_x,
      p->covariates.as<double>(), xs, d_eta_out, d_ids, d_popflag);
  CUDA_TRY(cudaGetLastError());
  p->counters[3] += 1;

  const double* d_ll = nullptr;
  const double* d_dlp = d_dlp_in;
  if (!skip_solve) {
    int rc = launch_loglik(p, B, d_ids, xs, with_grad, d_sys_ll, d_sys_grad,
                           d_sys_status, p->nst.as<int32_t>() + b0,
                           p->nrj.as<int32_t>() + b0, s, C,
                           static_cast<int32_t>(n_sh), b0, B_res,
                           work_counter);
    if (rc) return rc;
    d_ll = d_sys_ll;
    d_dlp = d_sys_grad;
    if (d_status) {
      chain_status_kernel<<<C, 128, 0, s>>>(n_sh, d_sys_status, d_status);
      CUDA_TRY(cudaGetLastError());
      p->counters[3] += 1;
    }
  }
  dim3 grid(static_cast<unsigned>(n_blk), static_cast<unsigned>(C));
  const size_t smem = (POP_BLOCK / 32) * n_slots * sizeof(double);
  pop_backward_kernel<<<grid, POP_BLOCK, smem, s>>>(
      D, p->n_ids, p->id_begin, p->id_end, n_params, n_bottom, with_grad,
      d_x, p->covariates.as<double>(), d_ll, d_dlp, d_grad, d_partial);
  CUDA_TRY(cudaGetLastError());
  // individuals sharded over several GPUs: one all-reduce (sum) of
  // [score | d theta | flags] per evaluation, on the compute stream, before
  // anything travels to the host (chi/_log_pdfs.py:289-292 partitions over
  // individuals; the bottom-level blocks are disjoint and stay where they are)
  const bool reduce = p->comm != nullptr && !skip_solve;
  double* d_pack = nullptr;
  if (reduce) {
    CUDA_TRY(p->pack.reserve(static_cast<size_t>(C_res) * (n_slots + 2) *
                             sizeof(double)));
    d_pack = p->pack.as<double>() + static_cast<int64_t>(c0) * (n_slots + 2);
  }
  pop_final_kernel<<<C, 64, 0, s>>>(D.n_red, n_blk, n_params, n_bottom,
                                    with_grad, d_r2t, d_partial, d_popflag,
                                    d_status, d_score, d_grad, d_pack);
  CUDA_TRY(cudaGetLastError());
  if (reduce) {
    NCCL_TRY(nccl_api().AllReduce(
        d_pack, d_pack, static_cast<size_t>(C) * (n_slots + 2), ncclDouble,
        ncclSum, p->comm, s));
    pop_unpack_kernel<<<C, 64, 0, s>>>(D.n_red, n_params, n_bottom, with_grad,
                                       d_r2t, d_pack, d_score, d_grad,
                                       d_status);
    CUDA_TRY(cudaGetLastError());
    p->counters[3] += 1;
  }
  p->counters[3] += 2;
  p->counters[0] = b0 + B;
  return 0;
}

}  // namespace

extern "C" {

int chi_b200_hier_logpdf_dev(chi_b200_problem* p, int32_t C,
                             const double* d_x, int32_t with_grad,
                             double* d_score, double* d_grad,
                             int32_t* d_status, void* stream) {
  CUDA_TRY(cudaSetDevice(p->device));
  p->counters[3] = 0;
  return run_hier(p, C, d_x, with_grad, d_score, d_grad, d_status, false,
                  nullptr, nullptr, nullptr,
                  static_cast<cudaStream_t>(stream));
}

static int hier_logpdf_host(chi_b200_problem* p, int32_t C, const double* x,
                            int32_t with_grad, double* score, double* grad,
                            int32_t* status, const bool wait) {
  if (!p->has_pop) return fail(-1, "no population model set");
  if (!p->vt) return fail(-1, "problem has no mechanistic model");
  if (C <= 0) return 0;
  CUDA_TRY(cudaSetDevice(p->device));
  cudaStream_t s = p->stream;
  const int64_t n_params = p->n_ids * p->pop.n_hdim + p->n_top;
  CUDA_TRY(p->hx.reserve(C * n_params * sizeof(double)));
  CUDA_TRY(p->hscore.reserve(C * sizeof(double)));
  CUDA_TRY(p->hstatus.reserve(C * sizeof(int32_t)));
  if (with_grad) CUDA_TRY(p->hgrad.reserve(C * n_params * sizeof(double)));
  // With a shard set only the shard's block of bottom-level parameters and
  // the top-level parameters travel (the same two column blocks of the
  // gradient come back): an individuals-sharded rank of an 8-GPU job moves an
  // eighth of the bytes.
  const bool sharded = p->id_end - p->id_begin != p->n_ids;
  const int64_t n_bottom = p->n_ids * p->pop.n_hdim;
  const size_t pitch = static_cast<size_t>(n_params) * sizeof(double);
  const size_t off_b = static_cast<size_t>(p->id_begin) * p->pop.n_hdim;
  const size_t w_b = static_cast<size_t>(p->id_end - p->id_begin) *
                     p->pop.n_hdim * sizeof(double);
  const size_t w_t = static_cast<size_t>(p->n_top) * sizeof(double);
  // The vectors are evaluated in chunks.  Copies travel on their own streams
  // -- every chunk's parameters on `stream_in`, its results on `stream_out`
  // -- and the chunks' kernels alternate between two compute streams, tied
  // together by one event per chunk and direction: the copy engines run ahead
  // of / behind the SMs, the persistent solve kernel of one chunk fills the
  // SMs as the one before it drains, and only the first chunk's copy in and
  // the last chunk's copy out are exposed.  (Copies queued on the compute
  // streams themselves do not overlap: the kernels of the two streams share
  // the SMs, finish together, and both streams then copy while the SMs
  // idle -- measured with CHI_B200_TIMELINE.)  The chunks' scratch is
  // disjoint (run_hier indexes it by the position in the batch).  One chunk
  // when there is too little to move for that to pay.
  static const int chunks_env = [] {
    const char* e = std::getenv("CHI_B200_CHUNKS");
    return e ? std::atoi(e) : 0;
  }();
  int n_chunks = 1;
  if (p->cfg_chunks > 0) {
    n_chunks = p->cfg_chunks;
  } else if (chunks_env > 0) {
    n_chunks = chunks_env;
  } else if (!p->timing && C >= 128) {
    // about 32 MB per chunk and direction (the first chunk's copy in and the
    // last chunk's copy out are the part that nothing overlaps), at least
    // four chunks once there are 8 MB to move, and no chunk below ~200 k
    // systems (the persistent solve kernel wants a few rounds of work)
    // (with a communicator every rank has to cut the batch into the SAME
    // chunks -- each chunk ends with an all-reduce --, so the count follows
    // from the largest shard of a balanced split, not from this rank's own)
    int64_t shard_ids = p->id_end - p->id_begin;
    if (p->comm && p->comm_world > 1)
      shard_ids = (p->n_ids + p->comm_world - 1) / p->comm_world;
    const size_t bytes =
        static_cast<size_t>(C) *
        (sharded ? static_cast<size_t>(shard_ids) * p->pop.n_hdim *
                           sizeof(double) + w_t
                 : pitch);
    const int64_t systems = static_cast<int64_t>(C) * shard_ids;
    if (bytes >= (size_t(8) << 20)) {
      n_chunks = std::max<int>(4, static_cast<int>(bytes >> 25));
      n_chunks = std::min<int>(
          n_chunks, std::max<int>(4, static_cast<int>(systems / 200000)));
    }
  }
  n_chunks =
      std::max(1, std::min<int>(n_chunks, std::min<int32_t>(C, MAX_CHUNKS)));
  if (n_chunks > 1) {
    if (!p->stream2)
      CUDA_TRY(cudaStreamCreateWithFlags(&p->stream2, cudaStreamNonBlocking));
    if (!p->stream_in)
      CUDA_TRY(cudaStreamCreateWithFlags(&p->stream_in, cudaStreamNonBlocking));
    if (!p->stream_out)
      CUDA_TRY(
          cudaStreamCreateWithFlags(&p->stream_out, cudaStreamNonBlocking));
    while (static_cast<int>(p->ev_in.size()) < n_chunks) {
      cudaEvent_t e0, e1;
      CUDA_TRY(cudaEventCreateWithFlags(&e0, cudaEventDisableTiming));
      CUDA_TRY(cudaEventCreateWithFlags(&e1, cudaEventDisableTiming));
      p->ev_in.push_back(e0);
      p->ev_comp.push_back(e1);
    }
  }
  const bool piped = n_chunks > 1;
  // (work a caller may have queued on the problem's stream comes first)
  if (piped) CUDA_TRY(cudaStreamSynchronize(s));
  CUDA_TRY(p->work_counters.reserve(MAX_CHUNKS_P1 * sizeof(unsigned long long)));
  p->counters[3] = 0;
  // chunk boundaries at multiples of 32 vectors when the chunks are large
  // enough: the warps of the solve kernel then hold 32 vectors of one
  // individual without padding
  const auto chunk_begin = [&](int k) {
    int64_t c = static_cast<int64_t>(C) * k / n_chunks;
    if (k < n_chunks && C >= 64 * n_chunks) c = c / 32 * 32;
    return static_cast<int32_t>(k >= n_chunks ? C : c);
  };
  // CHI_B200_TIMELINE=1: per-chunk event times on stderr (development)
  static const bool timeline = std::getenv("CHI_B200_TIMELINE") != nullptr;
  std::vector<cudaEvent_t> tl;
  if (timeline) {
    tl.resize(static_cast<size_t>(n_chunks) * 4 + 1);
    for (auto& e : tl) CUDA_TRY(cudaEventCreate(&e));
    CUDA_TRY(cudaEventRecord(tl.back(), s));
    if (piped) {
      CUDA_TRY(cudaStreamWaitEvent(p->stream2, tl.back(), 0));
      CUDA_TRY(cudaStreamWaitEvent(p->stream_in, tl.back(), 0));
    }
  }
  for (int k = 0; k < n_chunks; ++k) {
    const int32_t c0 = chunk_begin(k);
    const int32_t c1 = chunk_begin(k + 1);
    const int32_t Ck = c1 - c0;
    if (Ck <= 0) continue;
    cudaStream_t sk = (k & 1) ? p->stream2 : s;
    cudaStream_t s_in = piped ? p->stream_in : sk;
    cudaStream_t s_out = piped ? p->stream_out : sk;
    if (timeline) CUDA_TRY(cudaEventRecord(tl[k * 4], s_in));
    const double* xk = x + static_cast<int64_t>(c0) * n_params;
    double* d_xk = p->hx.as<double>() + static_cast<int64_t>(c0) * n_params;
    double* d_gk = with_grad ? p->hgrad.as<double>() +
                                   static_cast<int64_t>(c0) * n_params
                             : nullptr;
    if (!sharded) {
      CUDA_TRY(cudaMemcpyAsync(d_xk, xk, Ck * pitch, cudaMemcpyHostToDevice,
                               s_in));
    } else {
      if (w_b)
        CUDA_TRY(cudaMemcpy2DAsync(d_xk + off_b, pitch, xk + off_b, pitch,
                                   w_b, Ck, cudaMemcpyHostToDevice, s_in));
      if (w_t)
        CUDA_TRY(cudaMemcpy2DAsync(d_xk + n_bottom, pitch, xk + n_bottom,
                                   pitch, w_t, Ck, cudaMemcpyHostToDevice,
                                   s_in));
    }
    if (timeline) CUDA_TRY(cudaEventRecord(tl[k * 4 + 1], s_in));
    if (piped) {
      CUDA_TRY(cudaEventRecord(p->ev_in[k], s_in));
      CUDA_TRY(cudaStreamWaitEvent(sk, p->ev_in[k], 0));
    }
    // top-level entries that belong to individuals of other shards
    // (heterogeneous dimensions) read as zero
    if (sharded && with_grad && w_t)
      CUDA_TRY(cudaMemset2DAsync(d_gk + n_bottom, pitch, 0, w_t, Ck, sk));
    int rc = run_hier(
        p, Ck, d_xk, with_grad, p->hscore.as<double>() + c0, d_gk,
        p->hstatus.as<int32_t>() + c0, false, nullptr, nullptr, nullptr, sk,
        c0, C,
        n_chunks > 1 ? p->work_counters.as<unsigned long long>() + k
                     : nullptr);
    if (rc) return rc;
    if (timeline) CUDA_TRY(cudaEventRecord(tl[k * 4 + 2], sk));
    if (piped) {
      CUDA_TRY(cudaEventRecord(p->ev_comp[k], sk));
      CUDA_TRY(cudaStreamWaitEvent(s_out, p->ev_comp[k], 0));
    }
    CUDA_TRY(cudaMemcpyAsync(score + c0, p->hscore.as<double>() + c0,
                             Ck * sizeof(double), cudaMemcpyDeviceToHost,
                             s_out));
    if (with_grad && grad) {
      double* gk = grad + static_cast<int64_t>(c0) * n_params;
      if (!sharded) {
        CUDA_TRY(cudaMemcpyAsync(gk, d_gk, Ck * pitch, cudaMemcpyDeviceToHost,
                                 s_out));
      } else {
        if (w_b)
          CUDA_TRY(cudaMemcpy2DAsync(gk + off_b, pitch, d_gk + off_b, pitch,
                                     w_b, Ck, cudaMemcpyDeviceToHost, s_out));
        if (w_t)
          CUDA_TRY(cudaMemcpy2DAsync(gk + n_bottom, pitch, d_gk + n_bottom,
                                     pitch, w_t, Ck, cudaMemcpyDeviceToHost,
                                     s_out));
      }
    }
    if (status)
      CUDA_TRY(cudaMemcpyAsync(status + c0, p->hstatus.as<int32_t>() + c0,
                               Ck * sizeof(int32_t), cudaMemcpyDeviceToHost,
                               s_out));
    if (timeline) CUDA_TRY(cudaEventRecord(tl[k * 4 + 3], s_out));
  }
  p->counters_stale = p->nst.ptr != nullptr;
  if (!wait && !timeline) return 0;   // chi_b200_problem_synchronize
  CUDA_TRY(cudaStreamSynchronize(s));
  if (piped) {
    CUDA_TRY(cudaStreamSynchronize(p->stream2));
    CUDA_TRY(cudaStreamSynchronize(p->stream_out));
    CUDA_TRY(cudaStreamSynchronize(p->stream_in));
  }
  if (timeline) {
    for (int k = 0; k < n_chunks; ++k) {
      float t[4] = {0, 0, 0, 0};
      for (int j = 0; j < 4; ++j)
        cudaEventElapsedTime(&t[j], tl.back(), tl[k * 4 + j]);
      std::fprintf(stderr,
                   "chunk %2d [%4d, %4d): start %7.3f  h2d done %7.3f  "
                   "kernels done %7.3f  d2h done %7.3f ms\n",
                   k, chunk_begin(k), chunk_begin(k + 1), t[0], t[1], t[2],
                   t[3]);
    }
    for (auto& e : tl) cudaEventDestroy(e);
  }
  // the step counters (counters[1], [2]) are aggregated on demand
  // (chi_b200_problem_counters), not on the evaluation path
  return 0;
}

int chi_b200_hier_logpdf(chi_b200_problem* p, int32_t C, const double* x,
                         int32_t with_grad, double* score, double* grad,
                         int32_t* status) {
  return hier_logpdf_host(p, C, x, with_grad, score, grad, status, true);
}

int chi_b200_hier_logpdf_begin(chi_b200_problem* p, int32_t C, const double* x,
                               int32_t with_grad, double* score, double* grad,
                               int32_t* status) {
  return hier_logpdf_host(p, C, x, with_grad, score, grad, status, false);
}

int chi_b200_problem_synchronize(chi_b200_problem* p) {
  CUDA_TRY(cudaSetDevice(p->device));
  CUDA_TRY(cudaStreamSynchronize(p->stream));
  if (p->stream2) CUDA_TRY(cudaStreamSynchronize(p->stream2));
  if (p->stream_out) CUDA_TRY(cudaStreamSynchronize(p->stream_out));
  if (p->stream_in) CUDA_TRY(cudaStreamSynchronize(p->stream_in));
  return 0;
}

int chi_b200_population_eval(chi_b200_problem* p, const double* top,
                             const double* bottom, const double* dlogp_dpsi,
                             int32_t with_grad, double* eta, double* psi,
                             double* score, double* grad) {
  if (!p->has_pop) return fail(-1, "no population model set");
  if (p->pop.n_sys != p->pop.n_dim)
    return fail(-1,
                "stand-alone population evaluation is not available on a "
                "problem with fixed parameters");
  CUDA_TRY(cudaSetDevice(p->device));
  cudaStream_t s = p->stream;
  const PopDesc& D = p->pop;
  const int64_t n_bottom = p->n_ids * D.n_hdim;
  const int64_t n_params = n_bottom + p->n_top;
  const int64_t nd = D.n_dim;
  const int64_t id0 = p->id_begin, id1 = p->id_end;
  p->id_begin = 0;
  p->id_end = p->n_ids;
  DevBuf d_eta, d_psi, d_dlp;
  auto done = [&](int rc) {
    d_eta.release(); d_psi.release(); d_dlp.release();
    p->id_begin = id0;
    p->id_end = id1;
    return rc;
  };
  std::vector<double> hx(n_params);
  std::copy(bottom, bottom + n_bottom, hx.begin());
  std::copy(top, top + p->n_top, hx.begin() + n_bottom);
  cudaError_t err = p->hx.reserve(n_params * sizeof(double));
  if (err == cudaSuccess) err = p->hscore.reserve(sizeof(double));
  if (err == cudaSuccess) err = p->hgrad.reserve(n_params * sizeof(double));
  if (err == cudaSuccess) err = d_eta.reserve(p->n_ids * nd * sizeof(double));
  if (err == cudaSuccess) err = d_psi.reserve(p->n_ids * nd * sizeof(double));
  if (err == cudaSuccess && dlogp_dpsi) {
    err = d_dlp.reserve(p->n_ids * nd * sizeof(double));
    if (err == cudaSuccess)
      err = cudaMemcpyAsync(d_dlp.ptr, dlogp_dpsi,
                            p->n_ids * nd * sizeof(double),
                            cudaMemcpyHostToDevice, s);
  }
  if (err == cudaSuccess)
    err = cudaMemcpyAsync(p->hx.ptr, hx.data(), n_params * sizeof(double),
                          cudaMemcpyHostToDevice, s);
  if (err != cudaSuccess) return done(fail(-2, cudaGetErrorString(err)));
  int rc = run_hier(p, 1, p->hx.as<double>(), with_grad,
                    p->hscore.as<double>(), p->hgrad.as<double>(), nullptr,
                    true, dlogp_dpsi ? d_dlp.as<double>() : nullptr,
                    d_eta.as<double>(), d_psi.as<double>(), s);
  if (rc) return done(rc);
  if (eta)
    cudaMemcpyAsync(eta, d_eta.ptr, p->n_ids * nd * sizeof(double),
                    cudaMemcpyDeviceToHost, s);
  if (psi)
    cudaMemcpyAsync(psi, d_psi.ptr, p->n_ids * nd * sizeof(double),
                    cudaMemcpyDeviceToHost, s);
  if (score)
    cudaMemcpyAsync(score, p->hscore.ptr, sizeof(double),
                    cudaMemcpyDeviceToHost, s);
  if (with_grad && grad)
    cudaMemcpyAsync(grad, p->hgrad.ptr, n_params * sizeof(double),
                    cudaMemcpyDeviceToHost, s);
  err = cudaStreamSynchronize(s);
  if (err != cudaSuccess) return done(fail(-2, cudaGetErrorString(err)));
  return done(0);
}

int chi_b200_fp64_peak(int32_t device, double* tflops, double* sm_clock_mhz) {
  if (chi_b200_device_count() <= 0)
    return fail(-3, "no CUDA device available: chi_b200 has no CPU fallback");
  CUDA_TRY(cudaSetDevice(device));
  cudaDeviceProp prop;
  CUDA_TRY(cudaGetDeviceProperties(&prop, device));
  double* d_out = nullptr;
  CUDA_TRY(cudaMalloc(&d_out, sizeof(double)));
  cudaEvent_t e0, e1;
  CUDA_TRY(cudaEventCreate(&e0));
  CUDA_TRY(cudaEventCreate(&e1));
  // 8 blocks of 256 threads per SM = 16 warps per scheduler; about 10 ms per
  // launch and 24 launches, so that a clock sampler running beside it sees
  // the SM clock under this load
  const int blocks = prop.multiProcessorCount * 8, threads = 256;
  const int iters = 1 << 12;
  double best = 0.0;
  for (int rep = 0; rep < 24; ++rep) {
    CUDA_TRY(cudaEventRecord(e0));
    fp64_peak_kernel<<<blocks, threads>>>(d_out, iters, 0.999999, 1e-7);
    CUDA_TRY(cudaEventRecord(e1));
    CUDA_TRY(cudaEventSynchronize(e1));
    float ms = 0.f;
    CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
    const double flops =
        2.0 * 64.0 * iters * static_cast<double>(blocks) * threads;
    if (rep > 0) best = std::max(best, flops / (ms * 1e-3) * 1e-12);
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(d_out);
  *tflops = best;
  if (sm_clock_mhz) *sm_clock_mhz = prop.clockRate * 1e-3;
  return 0;
}

}  // extern "C"
