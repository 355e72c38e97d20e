// chi_b200 -- "streamed columns" variant of the sensitivity solve.
//
// Same method and semantics as solve_system<M, G, MC> (integrator.cuh): RODAS4
// on [y | S] with the exact block-triangular Jacobian, exact hits of dose
// edges and observation times, fused error model.  What differs is the
// mapping:
//
//   * ONE thread per system (no redundant state column, no shuffles);
//   * per step the state column runs first and keeps its six stage values
//     and increments in registers; the sensitivity columns are then streamed
//     through registers CB at a time (a uniform, non-unrolled loop: every
//     thread of the block handles column k at the same time, so the loop
//     adds no divergence and the hot loop stays inside the instruction
//     cache);
//   * the sensitivity matrix S (current and candidate) and the gradient
//     accumulators live in shared memory, slot-major (conflict free):
//         sm[slot * stride + tid],  slots = 2*P*N + P doubles per thread.
//
// Compiles for the host as well (stride = 1, tid = 0) for the CPU port.
#pragma once
#include "integrator.cuh"

namespace chi_b200 {

template <class M>
constexpr int stream_slots() {
  return 2 * M::P * M::N + M::P;
}

// (err / (atol + rtol * max(|z0|, |z1|)))^2 in single precision: the error
// norm only steers the step size, three digits are plenty, and this keeps two
// double-precision divisions per column and step off the FP64 pipe.  Values
// beyond the float range saturate to inf (-> rejection) or flush to zero.
CHI_HD float err_term(const double err, const double z0, const double z1,
                      const float atol_f, const float rtol_f) {
  const float m = fmaxf(fabsf(static_cast<float>(z0)),
                        fabsf(static_cast<float>(z1)));
  const float sc = fmaf(rtol_f, m, atol_f);
#ifdef __CUDA_ARCH__
  const float q = __fdividef(static_cast<float>(err), sc);
#else
  const float q = static_cast<float>(err) / sc;
#endif
  return q * q;
}

// State of one system while it is being integrated.  `init` loads a system,
// `advance` handles the pending events (observation records, dose edges) and
// then attempts ONE step, returning true once the system is finished,
// `finish` writes the results.  Splitting the solve this way lets a
// persistent kernel refill a lane with the next system as soon as its current
// one is done (solve_stream_persistent below).
template <class M, int CB>
struct StreamSolver {
  static constexpr int N = M::N;
  static constexpr int NC = M::NC;
  static constexpr int NCS = M::NCS;
  static constexpr int P = M::P;

  const SolveArgs& a;
  double* sm;
  const int stride;

  int64_t b;
  const double* sg;
  int n_cols;
  double y[N], c[NCS];
  int cur;
  bool loglik, invalid, after_reject;
  double ll;
  double dsig[MAX_SIGMA];
  int64_t r, r_end, r_begin, seg, seg_end, trow;
  double u, t_edge, t, h, h_full;
  int n_acc, n_rej, status;
  double J0[N * N];

  CHI_HD StreamSolver(const SolveArgs& args, double* smem, int smem_stride)
      : a(args), sm(smem), stride(smem_stride) {}

#define CHI_S(buf, k, n) sm[(((buf) * P + (k)) * N + (n)) * stride]
#define CHI_G(k) sm[(2 * P * N + (k)) * stride]

  CHI_HD void init(const int64_t system) {
    b = system;
    const int id = a.ids ? a.ids[b] : static_cast<int>(b % a.n_ids);
    const double* xb = a.x + b * a.x_stride;
    sg = xb + P;
    static_assert(CB == 1, "columns are streamed one at a time");
    n_cols = a.with_sens ? a.n_cols : 0;
#pragma unroll
    for (int i = 0; i < N; ++i) y[i] = xb[i];
#pragma unroll
    for (int i = 0; i < NC; ++i) c[i] = xb[N + i];
    if (NC == 0) c[0] = 0.0;
    cur = 0;
    for (int k = 0; k < P; ++k) {
      const int pk = k < n_cols ? a.col_param[k] : -1;
#pragma unroll
      for (int n = 0; n < N; ++n) {
        CHI_S(0, k, n) = (pk == n) ? 1.0 : 0.0;
        CHI_S(1, k, n) = 0.0;
      }
      CHI_G(k) = 0.0;
    }
    loglik = a.mode == 1;
    ll = 0.0;
#pragma unroll
    for (int k = 0; k < MAX_SIGMA; ++k) dsig[k] = 0.0;
    invalid = false;
    if (loglik) {
      for (int o = 0; o < a.n_out; ++o) {
        const int off = a.err_off[o];
        if (sg[off] <= 0.0) invalid = true;
        if (a.err_kind[o] == ERR_CAM && sg[off + 1] <= 0.0) invalid = true;
      }
    }
    r = a.rec_off[id];
    r_end = a.rec_off[id + 1];
    r_begin = r;
    seg = a.seg_off[id];
    seg_end = a.seg_off[id + 1];
    u = (seg < seg_end) ? a.seg_rate[seg] : 0.0;
    t_edge = (seg + 1 < seg_end) ? a.seg_t[seg + 1] : INFINITY;
    trow = (a.mode == 0) ? a.traj_off[b] : 0;
    t = 0.0;
    h = 0.0;
    h_full = 0.0;
    after_reject = false;
    n_acc = 0;
    n_rej = 0;
    status = ST_OK;
    if (M::LINEAR) M::jac(y, c, J0);
  }

  // register-resident accumulation into dsig[off] (no dynamic indexing, so
  // the solver state stays in registers)
  CHI_HD void add_dsig(const int off, const double v) {
#pragma unroll
    for (int k = 0; k < MAX_SIGMA; ++k) dsig[k] += (k == off) ? v : 0.0;
  }

  CHI_HD double get_sigma(const int off) const {
    return sg[off];
  }

  // observation record r at the current time
  CHI_HD void observe() {
    const int o = a.rec_out[r];
    const int oid = a.out_id[o];
    const double yb = M::out(oid, y, c);
    double w = 0.0;
    const int64_t row = trow + (r - r_begin);
    if (!loglik) {
      a.traj_y[row] = yb;
    } else {
      const int kind = a.err_kind[o];
      const int off = a.err_off[o];
      const double yo = a.rec_obs[r];
      if (kind == ERR_GAUSSIAN) {
        const double s = sg[off];
        const double is = 1.0 / s;
        const double q = (yo - yb) * is;
        ll -= LOG_2PI_HALF + log(s) + 0.5 * q * q;
        w = q * is;
        add_dsig(off, (q * q - 1.0) * is);
      } else if (kind == ERR_LOGNORMAL) {
        const double s = sg[off];
        if (yb <= 0.0) invalid = true;
        const double is = 1.0 / s;
        const double lo = a.rec_logobs[r];
        const double q = (lo - log(yb) + 0.5 * s * s) * is;
        ll -= LOG_2PI_HALF + log(s) + lo + 0.5 * q * q;
        w = q * is / yb;
        add_dsig(off, -q + (q * q - 1.0) * is);
      } else {
        const bool cam = kind == ERR_CAM;
        const double sb = cam ? sg[off] : 0.0;
        const double sr = cam ? sg[off + 1] : sg[off];
        const double st = fma(sr, yb, sb);
        const double ist = 1.0 / st;
        const double q = (yo - yb) * ist;
        ll -= LOG_2PI_HALF + log(st) + 0.5 * q * q;
        const double q2i = q * q * ist;
        w = q * ist - sr * ist + sr * q2i;
        if (cam) {
          add_dsig(off, q2i - ist);
          add_dsig(off + 1, (q2i - ist) * yb);
        } else {
          add_dsig(off, (q2i - ist) * yb);
        }
      }
    }
    for (int k = 0; k < n_cols; ++k) {
      const int pk = a.col_param[k];
      double s[N], e[NCS];
#pragma unroll
      for (int n = 0; n < N; ++n) s[n] = CHI_S(cur, k, n);
#pragma unroll
      for (int j = 0; j < NCS; ++j) e[j] = (pk == N + j) ? 1.0 : 0.0;
      const double dyb = M::dout(oid, y, c, s, e);
      if (loglik) {
        CHI_G(k) = fma(w, dyb, CHI_G(k));
      } else {
        a.traj_s[row * n_cols + k] = dyb;
      }
    }
  }

  // Handles the events pending at the current time (observation records,
  // dose-rate edges).  Returns true when the system is finished.
  CHI_HD bool events() {
    if (invalid) return true;
    for (;;) {
      if (r >= r_end) return true;
      if (a.rec_t[r] <= t) {
        observe();
        ++r;
        continue;
      }
      if (t_edge <= t) {
        // dose-rate edge: switch the input, restart the step size
        ++seg;
        u = a.seg_rate[seg];
        t_edge = (seg + 1 < seg_end) ? a.seg_t[seg + 1] : INFINITY;
        h = CHI_EDGE_KEEP_H ? h_full * CHI_EDGE_SHRINK : 0.0;
        continue;
      }
      return false;
    }
  }

  // Six RODAS4 stages of sensitivity column k (constant index KC, -1 for an
  // initial-value column) given the state column's stage values Zy and
  // increments Uy.  Returns the column's squared scaled error.
  template <int KC>
  CHI_HD float column(const int k, const int nxt, const double (&Zy)[6][N],
                      const double (&Uy)[6][N], const double (&Wi)[N * N],
                      const double hinv, const float atol_f,
                      const float rtol_f) {
    using namespace rodas4;
    double S0[N], Us[6][N];
#pragma unroll
    for (int n = 0; n < N; ++n) S0[n] = CHI_S(cur, k, n);
#define CHI_SSTAGE(I, CNT, a0, a1, a2, a3, a4, c0, c1, c2, c3, c4)           \
  {                                                                           \
    double Ji[N * N];                                                         \
    if (!M::LINEAR) M::jac(Zy[I], c, Ji);                                     \
    double Zs[N], cs[N], rs[N];                                               \
    lincomb<CNT, N>(Zs, S0, Us, a0, a1, a2, a3, a4);                          \
    lincomb<CNT, N>(cs, nullptr, Us, c0, c1, c2, c3, c4);                     \
    matvec<N>(M::LINEAR ? J0 : Ji, Zs, rs);                                   \
    M::template dfdc_col_add<KC>(Zy[I], c, rs);                               \
    M::template hess_col_add<KC>(y, c, S0, Uy[I], rs);                        \
    _Pragma("unroll") for (int n = 0; n < N; ++n) rs[n] =                     \
        fma(hinv, cs[n], rs[n]);                                              \
    matvec<N>(Wi, rs, Us[I]);                                                 \
  }
    CHI_SSTAGE(0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0)
    CHI_SSTAGE(1, 1, A21, 0, 0, 0, 0, C21, 0, 0, 0, 0)
    CHI_SSTAGE(2, 2, A31, A32, 0, 0, 0, C31, C32, 0, 0, 0)
    CHI_SSTAGE(3, 3, A41, A42, A43, 0, 0, C41, C42, C43, 0, 0)
    CHI_SSTAGE(4, 4, A51, A52, A53, A54, 0, C51, C52, C53, C54, 0)
    CHI_SSTAGE(5, 5, A51, A52, A53, A54, 1.0, C61, C62, C63, C64, C65)
#undef CHI_SSTAGE
    double Sn[N];
    lincomb<5, N>(Sn, S0, Us, A51, A52, A53, A54, 1.0);
    float s2 = 0.0f;
#pragma unroll
    for (int n = 0; n < N; ++n) {
      Sn[n] += Us[5][n];
      s2 += err_term(Us[5][n], S0[n], Sn[n], atol_f, rtol_f);
      CHI_S(nxt, k, n) = Sn[n];
    }
    return s2 * (1.0f / N);
  }

  // Uniform dispatch to the code specialised for constant index kc.  Columns
  // of constants that do not enter the rhs stay identically zero: nothing to
  // integrate.
  template <int K>
  CHI_HD float dispatch_column(const int kc, const int k, const int nxt,
                               const double (&Zy)[6][N],
                               const double (&Uy)[6][N],
                               const double (&Wi)[N * N], const double hinv,
                               const float atol_f, const float rtol_f) {
    if (kc == K) {
      if constexpr (K >= 0) {
        if constexpr ((M::INERT_MASK >> K) & 1u) return 0.0f;
      }
      return column<K>(k, nxt, Zy, Uy, Wi, hinv, atol_f, rtol_f);
    }
    if constexpr (K + 1 < NC) {
      return dispatch_column<K + 1>(kc, k, nxt, Zy, Uy, Wi, hinv, atol_f,
                                    rtol_f);
    } else {
      return 0.0f;
    }
  }

  // One step attempt towards the next stop.  Returns true on failure.
  CHI_HD bool step() {
    using namespace rodas4;
    const double t_stop = fmin(a.rec_t[r], t_edge);

    if (h == 0.0) {
      double f0[N];
      M::rhs(y, c, u, f0);
      double d0 = 0.0, d1 = 0.0;
#pragma unroll
      for (int i = 0; i < N; ++i) {
        const double sc = a.atol + a.rtol * fabs(y[i]);
        d0 = fma(y[i] / sc, y[i] / sc, d0);
        d1 = fma(f0[i] / sc, f0[i] / sc, d1);
      }
      d0 = sqrt(d0 / N);
      d1 = sqrt(d1 / N);
      h = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : 0.01 * d0 / d1;
      h = fmin(100.0 * h, t_stop - t);
      after_reject = false;
    }

    const double rem = t_stop - t;
    h_full = h;
    double hs = h;
    bool clipped = false;
    if (rem <= 1.1 * h) {
      hs = rem;
      clipped = true;
    } else if (rem < 2.0 * h) {
      hs = 0.5 * rem;
    }

    // ---- state column: six stages, values and increments kept -------------
    if (!M::LINEAR) M::jac(y, c, J0);
    double Wi[N * N];
    {
      double W[N * N];
      const double fac = 1.0 / (GAMMA * hs);
#pragma unroll
      for (int i = 0; i < N; ++i)
#pragma unroll
        for (int j = 0; j < N; ++j)
          W[i * N + j] = (i == j ? fac : 0.0) - J0[i * N + j];
      invert<N>(W, Wi);
    }
    const double hinv = 1.0 / hs;
    double Zy[6][N];
    double Uy[6][N];
#define CHI_YSTAGE(I, CNT, a0, a1, a2, a3, a4, c0, c1, c2, c3, c4)           \
  {                                                                           \
    double f[N], cy[N], ry[N];                                                \
    lincomb<CNT, N>(Zy[I], y, Uy, a0, a1, a2, a3, a4);                        \
    M::rhs(Zy[I], c, u, f);                                                   \
    lincomb<CNT, N>(cy, nullptr, Uy, c0, c1, c2, c3, c4);                     \
    _Pragma("unroll") for (int n = 0; n < N; ++n) ry[n] =                     \
        fma(hinv, cy[n], f[n]);                                               \
    matvec<N>(Wi, ry, Uy[I]);                                                 \
  }
    CHI_YSTAGE(0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0)
    CHI_YSTAGE(1, 1, A21, 0, 0, 0, 0, C21, 0, 0, 0, 0)
    CHI_YSTAGE(2, 2, A31, A32, 0, 0, 0, C31, C32, 0, 0, 0)
    CHI_YSTAGE(3, 3, A41, A42, A43, 0, 0, C41, C42, C43, 0, 0)
    CHI_YSTAGE(4, 4, A51, A52, A53, A54, 0, C51, C52, C53, C54, 0)
    CHI_YSTAGE(5, 5, A51, A52, A53, A54, 1.0, C61, C62, C63, C64, C65)
#undef CHI_YSTAGE
    double yn[N];
    float errsq;
    const float atol_f = static_cast<float>(a.atol);
    const float rtol_f = static_cast<float>(a.rtol);
    {
      float s2 = 0.0f;
#pragma unroll
      for (int n = 0; n < N; ++n) {
        yn[n] = Zy[5][n] + Uy[5][n];
        s2 += err_term(Uy[5][n], y[n], yn[n], atol_f, rtol_f);
      }
      errsq = s2 * (1.0f / N);
    }

    // ---- sensitivity columns, streamed one at a time ----------------------
    // The column's constant index is uniform over the grid, so the dispatch
    // below is a uniform branch to code specialised for that column.
    const int nxt = 1 - cur;
    for (int k = 0; k < n_cols; ++k) {
      const int pk = a.col_param[k];
      const int kc = pk >= N ? pk - N : -1;
      errsq = fmaxf(errsq, dispatch_column<-1>(kc, k, nxt, Zy, Uy, Wi, hinv,
                                               atol_f, rtol_f));
    }

    const bool accept = errsq <= 1.0f;
    float fac;
    if (!(errsq == errsq) || isinf(errsq)) {
      fac = 0.2f;
    } else {
      fac = CHI_SAFETY * fast_powf(fmaxf(errsq, 1e-30f), -0.125f);
      fac = fminf(fmaxf(fac, 0.2f), CHI_FACMAX);
    }
    if (accept) {
      t = clipped ? t_stop : t + hs;
#pragma unroll
      for (int n = 0; n < N; ++n) y[n] = yn[n];
      cur = nxt;
      if (after_reject) fac = fminf(fac, 1.0f);
      const double hn = hs * static_cast<double>(fac);
      h = (hs < h) ? fmax(hn, h) : hn;
      after_reject = false;
      ++n_acc;
    } else {
      h = hs * static_cast<double>(fminf(fac, 0.9f));
      after_reject = true;
      ++n_rej;
      if (h <= 1e-14 * fmax(fabs(t), 1.0)) {
        status = (errsq == errsq) ? ST_STEP_UNDERFLOW : ST_NON_FINITE;
        return true;
      }
    }
    if (n_acc + n_rej >= a.max_steps) {
      status = ST_MAX_STEPS;
      return true;
    }
    return false;
  }

  CHI_HD void finish() {
    a.status[b] = status;
    if (a.n_steps) a.n_steps[b] = n_acc;
    if (a.n_rej) a.n_rej[b] = n_rej;
    if (loglik) {
      const bool fail = invalid || status != ST_OK;
      a.ll[b] = fail ? -INFINITY : ll;
      if (a.with_sens && a.grad) {
        double* gb = a.grad + b * a.grad_stride;
        for (int k = 0; k < n_cols; ++k)
          gb[a.col_param[k]] = fail ? INFINITY : CHI_G(k);
#pragma unroll
        for (int k = 0; k < MAX_SIGMA; ++k)
          if (k < a.n_sigma) gb[P + k] = fail ? INFINITY : dsig[k];
      }
    } else if (status != ST_OK) {
      for (int64_t q = r; q < r_end; ++q)
        a.traj_y[trow + (q - r_begin)] = NAN;
    }
  }
#undef CHI_S
#undef CHI_G
};

template <class M, int CB>
CHI_HD void solve_system_stream(const SolveArgs& a, const int64_t b,
                                double* sm, const int stride) {
  StreamSolver<M, CB> s(a, sm, stride);
  s.init(b);
  while (!s.events()) {
    if (s.step()) break;
  }
  s.finish();
}

#ifdef __CUDACC__

// Persistent kernel: every lane pulls systems from a global counter and
// refills itself as soon as its system is finished, so lanes whose systems
// need fewer steps do not idle and there is no partial last wave.
template <class M, int CB, int BLOCK>
__global__ void __launch_bounds__(BLOCK, M::STREAM_MIN_BLOCKS)
solve_stream_kernel(const SolveArgs a, unsigned long long* counter) {
  extern __shared__ double chi_sm[];
  StreamSolver<M, CB> s(a, chi_sm + threadIdx.x, BLOCK);
  bool active = false;
  bool exhausted = false;
  for (;;) {
    if (!active && !exhausted) {
      const unsigned long long b = atomicAdd(counter, 1ull);
      if (b >= static_cast<unsigned long long>(a.B)) {
        exhausted = true;
      } else {
        s.init(static_cast<int64_t>(b));
        active = true;
      }
    }
    // all 32 lanes stay in the loop until the warp has no work left
    if (!__any_sync(0xffffffffu, active)) break;
    bool done = active ? s.events() : true;
    // the vote is the convergence point: lanes that refilled or handled an
    // event re-join the others here, so the (expensive) step below is
    // executed once per iteration by all lanes that need it
    __syncwarp(0xffffffffu);
    if (active && !done) done = s.step();
    if (active && done) {
      s.finish();
      active = false;
    }
  }
}

#endif  // __CUDACC__

}  // namespace chi_b200
