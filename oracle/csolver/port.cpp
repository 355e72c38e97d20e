// CPU port of the chi_b200 solve for baseline timing and debugging.
//
// TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  This translation unit
// compiles the *same* per-system routine the CUDA kernel runs
// (chi_b200/csrc/integrator.cuh: solve_system<M, 1, P>) for the host with
// g++ -O3, one system per call, so that
//   * bench.py can time "the reference's cost structure" on the host cores:
//     one native adaptive stiff solve with forward sensitivities per
//     individual (what chi gets from myokit/CVODES,
//     chi/_mechanistic_models.py:490-533), driven either by a Python loop
//     over individuals (chi/_log_pdfs.py:289-292) or by an OpenMP loop;
//   * the integrator's step control can be debugged without a GPU.
// It is a port of the product algorithm, not an independent check: parity is
// established against oracle/ode_exact.py, never against this file.
#include <cstdint>
#include <cstring>
#include <math.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "../../chi_b200/csrc/integrator.cuh"
#ifdef CHI_PORT_MODELS_FILE
#include CHI_PORT_MODELS_FILE   // a test's own model list (build.build_models)
#else
#include "port_models.inc"   // generated: includes + dispatch table
#endif

using namespace chi_b200;

// stream_cb: 0 = lanes routine (one lane, all columns in "registers"),
// 1 = streamed-columns routine with RODAS4, 5 = with RODAS5, 9 / 11 = with
// ROSL(9) / ROSL(11) (linear models), 109 / 111 = the same through
// StreamSolver::solve_uniform.
template <class M>
static void run_all(const SolveArgs& a, int n_threads, int stream_cb) {
  (void)n_threads;
  // split launches (SolveArgs::split): one work item per (system, column)
  const int64_t items = a.B * (a.split > 0 ? a.split : 1);
  if (a.split > 0)
    for (int64_t b = 0; b < a.B; ++b) a.status[b] = 0;
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 16) num_threads(a.split > 0 ? 1 : n_threads)
#endif
  for (int64_t b = 0; b < items; ++b) {
    if (!a.with_sens && stream_cb == 0) {
      solve_system<M, 1, 0>(a, b, 0, 0, 0, 1u);
    } else if (stream_cb == 1) {
      double sm[stream_slots<M>()];
      solve_system_stream<M, 1>(a, b, sm, 1);
    } else if (stream_cb == 5) {
      double sm[stream_slots<M>()];
      solve_system_stream<M, 5>(a, b, sm, 1);
    } else if (stream_cb == 109 || stream_cb == 111) {
      // the warp-synchronous form of the ROSL kernels, one lane
      double sm[stream_slots<M>()];
      if constexpr (M::LINEAR) {
        if (stream_cb == 109) solve_system_uniform<M, 9>(a, b, sm, 1);
        else solve_system_uniform<M, 11>(a, b, sm, 1);
      } else {
        solve_system_stream<M, 5>(a, b, sm, 1);
      }
    } else if (stream_cb == 9 || stream_cb == 11) {
      // ROSL(9) / ROSL(11): linear models only (RODAS5 otherwise)
      double sm[stream_slots<M>()];
      if constexpr (M::LINEAR) {
        if (stream_cb == 9) solve_system_stream<M, 9>(a, b, sm, 1);
        else solve_system_stream<M, 11>(a, b, sm, 1);
      } else {
        solve_system_stream<M, 5>(a, b, sm, 1);
      }
    } else {
      solve_system<M, 1, M::P>(a, b, 0, 0, 0, 1u);
    }
  }
  // deferred observation (after the solve: the items of a split launch share
  // their system's snapshot rows)
  if (stream_cb != 0 && a.mode == 1 && a.snap) {
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 16) num_threads(n_threads)
#endif
    for (int64_t b = 0; b < a.B; ++b) observe_system<M>(a, b);
  }
}

// fills SolveArgs::n_int / int_col / int_param from col_param
template <class M>
static int fill_integrated(SolveArgs& a) {
  a.n_int = 0;
  for (int k = 0; k < a.n_cols; ++k) {
    const int pk = a.col_param[k];
    if (!(pk >= M::N && ((M::INERT_MASK >> (pk - M::N)) & 1u))) {
      a.int_col[a.n_int] = k;
      a.int_param[a.n_int] = pk;
      ++a.n_int;
    }
  }
  return 0;
}

extern "C" int chi_port_fill_integrated(const char* key, SolveArgs* a) {
#define CHI_PORT_MODEL(KEY, TYPE)                                            \
  if (std::strcmp(key, KEY) == 0) return fill_integrated<TYPE>(*a);
  CHI_PORT_MODELS
#undef CHI_PORT_MODEL
  return -1;
}

// doubles per snapshot row for these arguments (SolveArgs::snap_width)
template <class M>
static int snap_width(const SolveArgs& a) {
  return M::N * (1 + (a.with_sens ? a.n_int : 0));
}

extern "C" int chi_port_snap_width(const char* key, const SolveArgs* a) {
#define CHI_PORT_MODEL(KEY, TYPE)                                            \
  if (std::strcmp(key, KEY) == 0) return snap_width<TYPE>(*a);
  CHI_PORT_MODELS
#undef CHI_PORT_MODEL
  return -1;
}

extern "C" int chi_port_solve2(const char* key, const SolveArgs* a,
                               int n_threads, int stream_cb) {
#define CHI_PORT_MODEL(KEY, TYPE)                                            \
  if (std::strcmp(key, KEY) == 0) {                                          \
    run_all<TYPE>(*a, n_threads, stream_cb);                                 \
    return 0;                                                                \
  }
  CHI_PORT_MODELS
#undef CHI_PORT_MODEL
  return -1;
}

extern "C" int chi_port_solve(const char* key, const SolveArgs* a,
                              int n_threads) {
  return chi_port_solve2(key, a, n_threads, 0);
}

extern "C" int chi_port_sizeof_args(void) { return (int)sizeof(SolveArgs); }

extern "C" int chi_port_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
