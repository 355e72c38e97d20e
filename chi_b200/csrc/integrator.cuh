// chi_b200 -- adaptive stiff integrator with forward sensitivities, fused with
// the error-model log-likelihood and gradient (kernels K1/K1f + K2 of
// SURVEY.md section 2.1).
//
// Replaces, for every (individual x chain) system at once,
//   chi.SBMLModel.simulate -> myokit.Simulation.run -> CVODES
//       (chi/_mechanistic_models.py:490-533)
//   chi.ErrorModel.compute_log_likelihood / compute_sensitivities
//       (chi/_error_models.py:229-331, 581-664, 925-1021, 1276-1375)
//   the per-output aggregation of chi.LogLikelihood.__call__/evaluateS1
//       (chi/_log_pdfs.py:802-843, 973-1027)
//
// Method: RODAS4 (Hairer & Wanner), a 6-stage, order 4(3), stiffly accurate,
// L-stable Rosenbrock method, applied to the augmented system z = [y | S]
// with the *exact* block-triangular Jacobian
//     d/dy [ J(y) S_k + (df/dc) e_k ]   (generated `hess`)
// so the sensitivities keep the full order and take part in the error test.
// Dose-rate edges and observation times are hit exactly (step clipping), the
// rate is switched at an edge and the step size re-initialised.
//
// Thread mapping: G lanes cooperate on one system; every lane carries the
// state column (redundantly) plus MC sensitivity columns in registers.  The
// only cross-lane traffic is one float max-reduction of the error norm per
// step.  (G, MC) is a compile-time variant chosen per model so that
// G * MC >= P with registers in budget; G = 1, MC = 0 is the forward-only
// kernel (thread per system).
#pragma once
#include <math.h>
#include <initializer_list>
#include <cstdlib>
#include <cstdint>
#include "model_api.cuh"

namespace chi_b200 {

namespace rodas4 {
// Hairer & Wanner, "Solving ODEs II", RODAS (METH=1), transformed form.
constexpr double GAMMA = 0.25;
constexpr double A21 = 0.1544000000000000e+01;
constexpr double A31 = 0.9466785280815826e+00;
constexpr double A32 = 0.2557011698983284e+00;
constexpr double A41 = 0.3314825187068521e+01;
constexpr double A42 = 0.2896124015972201e+01;
constexpr double A43 = 0.9986419139977817e+00;
constexpr double A51 = 0.1221224509226641e+01;
constexpr double A52 = 0.6019134481288629e+01;
constexpr double A53 = 0.1253708332932087e+02;
constexpr double A54 = -0.6878860361058950e+00;
constexpr double C21 = -0.5668800000000000e+01;
constexpr double C31 = -0.2430093356833875e+01;
constexpr double C32 = -0.2063599157091915e+00;
constexpr double C41 = -0.1073529058151375e+00;
constexpr double C42 = -0.9594562251023355e+01;
constexpr double C43 = -0.2047028614809616e+02;
constexpr double C51 = 0.7496443313967647e+01;
constexpr double C52 = -0.1024680431464352e+02;
constexpr double C53 = -0.3399990352819905e+02;
constexpr double C54 = 0.1170890893206160e+02;
constexpr double C61 = 0.8083246795921522e+01;
constexpr double C62 = -0.7981132988064893e+01;
constexpr double C63 = -0.3152159432874371e+02;
constexpr double C64 = 0.1631930543123136e+02;
constexpr double C65 = -0.6058818238834054e+01;
// K_ij = (alpha_ij + gamma_ij) / gamma of the method in its original k-stage
// form (tools/derive_kform.py derives them from the a, C tables above): the
// stages of a LINEAR system need one combination each (integrator_stream.cuh)
constexpr double K21 = 0.1268;
constexpr double K31 = 0.04988880902896939;
constexpr double K32 = 0.20411119097103052;
constexpr double K41 = 4.7841506773549481;
constexpr double K42 = 0.7099789456713099;
constexpr double K43 = -4.1189296230262583;
constexpr double K51 = 9.7145350618679512;
constexpr double K52 = -1.5309949350591384;
constexpr double K53 = -7.4228813237183214;
constexpr double K54 = 2.239341196909505;
constexpr double K61 = 1.3937770851442061;
constexpr double K62 = 0.8520544876475948;
constexpr double K63 = -0.61641013064927383;
constexpr double K64 = 1.8852831175659833;
constexpr double K65 = -0.5147045597085135;
}  // namespace rodas4

constexpr double LOG_2PI_HALF = 0.91893853320467274178;

// step-size controller constants
#ifndef CHI_SAFETY
#define CHI_SAFETY 0.95f
#endif
#ifndef CHI_FACMAX
#define CHI_FACMAX 5.0f
#endif
// a step that would end within this factor of the proposed size from the
// next stop (observation record, dose edge) is stretched to hit it; has to
// stay below 1/0.9 (a rejected step shrinks by at least 0.9, and the retry
// must not be stretched back to the same size)
#ifndef CHI_CLIP_STRETCH
#define CHI_CLIP_STRETCH 1.1
#endif
// first step size of the streamed solver, as a fraction of the distance to
// the first stop
#ifndef CHI_H0_FRACTION
#define CHI_H0_FRACTION 0.01
#endif
#ifndef CHI_EDGE_KEEP_H
#define CHI_EDGE_KEEP_H 1
#endif
#ifndef CHI_EDGE_SHRINK
#define CHI_EDGE_SHRINK 0.3
#endif

// ---- small dense helpers (fully unrolled, registers only) -----------------

// 1/x for the step size and the determinant of W: the hardware's 20-bit
// reciprocal seed and two Newton steps, without the special-case branch of
// the IEEE division (x is a step size or a determinant of order 1/h^N: finite,
// non-zero and far from the exponent range's ends; a non-finite x produces a
// non-finite result, which the error test rejects).  Accurate to ~1 ulp.
CHI_HD double rcp_newton(const double x) {
#ifdef __CUDA_ARCH__
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  r = fma(r, fma(-x, r, 1.0), r);
  r = fma(r, fma(-x, r, 1.0), r);
  return r;
#else
  return 1.0 / x;
#endif
}

template <int N>
CHI_HD void invert(const double* A, double* Ai);

template <>
CHI_HD void invert<1>(const double* A, double* Ai) {
  Ai[0] = rcp_newton(A[0]);
}

template <>
CHI_HD void invert<2>(const double* A, double* Ai) {
  const double r = rcp_newton(A[0] * A[3] - A[1] * A[2]);
  Ai[0] = A[3] * r;
  Ai[1] = -A[1] * r;
  Ai[2] = -A[2] * r;
  Ai[3] = A[0] * r;
}

template <>
CHI_HD void invert<3>(const double* A, double* Ai) {
  const double c00 = A[4] * A[8] - A[5] * A[7];
  const double c01 = A[5] * A[6] - A[3] * A[8];
  const double c02 = A[3] * A[7] - A[4] * A[6];
  const double r = rcp_newton(A[0] * c00 + A[1] * c01 + A[2] * c02);
  Ai[0] = c00 * r;
  Ai[1] = (A[2] * A[7] - A[1] * A[8]) * r;
  Ai[2] = (A[1] * A[5] - A[2] * A[4]) * r;
  Ai[3] = c01 * r;
  Ai[4] = (A[0] * A[8] - A[2] * A[6]) * r;
  Ai[5] = (A[2] * A[3] - A[0] * A[5]) * r;
  Ai[6] = c02 * r;
  Ai[7] = (A[1] * A[6] - A[0] * A[7]) * r;
  Ai[8] = (A[0] * A[4] - A[1] * A[3]) * r;
}

// Gauss-Jordan with partial pivoting for N >= 4 (select based, unrolled).
template <int N>
CHI_HD void invert(const double* A, double* Ai) {
  double a[N][N], b[N][N];
#pragma unroll
  for (int i = 0; i < N; ++i)
#pragma unroll
    for (int j = 0; j < N; ++j) {
      a[i][j] = A[i * N + j];
      b[i][j] = (i == j) ? 1.0 : 0.0;
    }
#pragma unroll
  for (int k = 0; k < N; ++k) {
#pragma unroll
    for (int i = k + 1; i < N; ++i) {
      const bool sw = fabs(a[i][k]) > fabs(a[k][k]);
#pragma unroll
      for (int j = 0; j < N; ++j) {
        const double t0 = a[k][j], t1 = a[i][j];
        a[k][j] = sw ? t1 : t0;
        a[i][j] = sw ? t0 : t1;
        const double u0 = b[k][j], u1 = b[i][j];
        b[k][j] = sw ? u1 : u0;
        b[i][j] = sw ? u0 : u1;
      }
    }
    const double r = 1.0 / a[k][k];
#pragma unroll
    for (int j = 0; j < N; ++j) {
      a[k][j] *= r;
      b[k][j] *= r;
    }
#pragma unroll
    for (int i = 0; i < N; ++i) {
      if (i == k) continue;
      const double m = a[i][k];
#pragma unroll
      for (int j = 0; j < N; ++j) {
        a[i][j] = fma(-m, a[k][j], a[i][j]);
        b[i][j] = fma(-m, b[k][j], b[i][j]);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < N; ++i)
#pragma unroll
    for (int j = 0; j < N; ++j) Ai[i * N + j] = b[i][j];
}

template <int N>
CHI_HD void matvec(const double* A, const double* x, double* r) {
#pragma unroll
  for (int i = 0; i < N; ++i) {
    double s = A[i * N] * x[0];
#pragma unroll
    for (int j = 1; j < N; ++j) s = fma(A[i * N + j], x[j], s);
    r[i] = s;
  }
}

// out = base + sum_{j<CNT} k_j U[j]   (base may be null)
template <int CNT, int N>
CHI_HD void lincomb(double* out, const double* base, const double (*U)[N],
                    double k0, double k1, double k2, double k3, double k4) {
#pragma unroll
  for (int n = 0; n < N; ++n) {
    double v = base ? base[n] : 0.0;
    if (CNT > 0) v = fma(k0, U[0][n], v);
    if (CNT > 1) v = fma(k1, U[1][n], v);
    if (CNT > 2) v = fma(k2, U[2][n], v);
    if (CNT > 3) v = fma(k3, U[3][n], v);
    if (CNT > 4) v = fma(k4, U[4][n], v);
    out[n] = v;
  }
}

template <int G_, int MC_>
struct Variant {
  static constexpr int G = G_;
  static constexpr int MC = MC_;
};

template <int G>
CHI_HD float group_max(float v, unsigned mask, int lane, int base) {
  (void)mask; (void)lane; (void)base;
  if (G == 1) return v;
#ifdef __CUDA_ARCH__
  if ((G & (G - 1)) == 0) {
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1)
      v = fmaxf(v, __shfl_xor_sync(mask, v, o));
    return v;
  }
  float r = v;
#pragma unroll
  for (int j = 1; j < G; ++j) {
    int src = lane - base + j;
    src = base + (src >= G ? src - G : src);
    r = fmaxf(r, __shfl_sync(mask, v, src));
  }
  return r;
#else
  return v;  // host build: one lane per system only
#endif
}

CHI_HD float fast_powf(float x, float y) {
#ifdef __CUDA_ARCH__
  return __powf(x, y);
#else
  return powf(x, y);
#endif
}

// Solves system `b` (one individual x one parameter vector).  `sub` is the
// lane's index inside its group of G lanes; (lane, base, gmask) describe the
// group inside the warp.  Compiles for the host too (G = 1), which is how the
// CPU port used for the baseline timing is built (oracle/csolver).
template <class M, int G, int MC>
CHI_HD void solve_system(const SolveArgs& a, const int64_t b, const int sub,
                         const int lane, const int base,
                         const unsigned gmask) {
  using namespace rodas4;
  constexpr int N = M::N;
  constexpr int NC = M::NC;
  constexpr int NCS = M::NCS;
  constexpr int P = M::P;
  constexpr int MCS = MC > 0 ? MC : 1;

  const int id = a.ids ? a.ids[b] : static_cast<int>(b % a.n_ids);
  const double* xb = a.x + b * a.x_stride;
  const double* sg = xb + P;

  double y[N], c[NCS];
#pragma unroll
  for (int i = 0; i < N; ++i) y[i] = xb[i];
#pragma unroll
  for (int i = 0; i < NC; ++i) c[i] = xb[N + i];
  if (NC == 0) c[0] = 0.0;

  // sensitivity columns owned by this lane
  int cp[MCS];
  double S[MCS][N];
  double e[MCS][NCS];
#pragma unroll
  for (int m = 0; m < MCS; ++m) {
    const int j = sub * MC + m;
    cp[m] = (MC > 0 && a.with_sens && j < a.n_cols) ? a.col_param[j] : -1;
#pragma unroll
    for (int i = 0; i < N; ++i) S[m][i] = (cp[m] == i) ? 1.0 : 0.0;
#pragma unroll
    for (int k = 0; k < NCS; ++k) e[m][k] = (cp[m] == N + k) ? 1.0 : 0.0;
  }

  const bool loglik = a.mode == 1;
  double ll = 0.0;
  double gacc[MCS];
  double dsig[MAX_SIGMA];
#pragma unroll
  for (int m = 0; m < MCS; ++m) gacc[m] = 0.0;
#pragma unroll
  for (int k = 0; k < MAX_SIGMA; ++k) dsig[k] = 0.0;
  bool invalid = false;  // -> score -inf, gradient +inf (reference rules)

  if (loglik) {
    for (int o = 0; o < a.n_out; ++o) {
      const int off = a.err_off[o];
      if (sg[off] <= 0.0) invalid = true;
      if (a.err_kind[o] == ERR_CAM && sg[off + 1] <= 0.0) invalid = true;
    }
  }

  int64_t r = a.rec_off[id];
  const int64_t r_end = a.rec_off[id + 1];
  const int64_t r_begin = r;
  int64_t seg = a.seg_off[id];
  const int64_t seg_end = a.seg_off[id + 1];
  double u = (seg < seg_end) ? a.seg_rate[seg] : 0.0;
  double t_edge = (seg + 1 < seg_end) ? a.seg_t[seg + 1] : INFINITY;
  const int64_t trow = (a.mode == 0) ? a.traj_off[b] : 0;

  double t = 0.0;
  double h = 0.0;
  bool after_reject = false;
  int edge_kind = 0;            // dose edge whose first step is pending
  float h_mem_up = 0.0f, h_mem_down = 0.0f;
  int n_acc = 0, n_rej = 0;
  int status = ST_OK;
  double J0[N * N];
  if (M::LINEAR) M::jac(y, c, J0);

  if (!invalid) {
    while (r < r_end) {
      const double t_rec = a.rec_t[r];
      if (t_rec <= t) {
        // ---- observation record at the current time --------------------
        const int o = a.rec_out[r];
        const int oid = a.out_id[o];
        const double yb = M::out(oid, y, c);
        double dyb[MCS];
#pragma unroll
        for (int m = 0; m < MCS; ++m)
          dyb[m] = (MC > 0) ? M::dout(oid, y, c, S[m], e[m]) : 0.0;
        if (!loglik) {
          const int64_t row = trow + (r - r_begin);
          if (sub == 0) a.traj_y[row] = yb;
          if (a.with_sens) {
#pragma unroll
            for (int m = 0; m < MC; ++m) {
              const int j = sub * MC + m;
              if (j < a.n_cols) a.traj_s[row * a.n_cols + j] = dyb[m];
            }
          }
        } else {
          const int kind = a.err_kind[o];
          const int off = a.err_off[o];
          const double yo = a.rec_obs[r];
          double w;  // d ll / d ybar
          if (kind == ERR_GAUSSIAN) {
            const double s = sg[off];
            const double is = 1.0 / s;
            const double er = yo - yb;
            const double q = er * is;
            ll -= LOG_2PI_HALF + log(s) + 0.5 * q * q;
            w = q * is;
            dsig[off] += (q * q - 1.0) * is;
          } else if (kind == ERR_LOGNORMAL) {
            const double s = sg[off];
            if (yb <= 0.0) invalid = true;
            const double is = 1.0 / s;
            const double lo = a.rec_logobs[r];
            const double er = lo - log(yb) + 0.5 * s * s;
            const double q = er * is;
            ll -= LOG_2PI_HALF + log(s) + lo + 0.5 * q * q;
            w = q * is / yb;
            dsig[off] += -q + (q * q - 1.0) * is;
          } else {
            const bool cam = kind == ERR_CAM;
            const double sb = cam ? sg[off] : 0.0;
            const double sr = cam ? sg[off + 1] : sg[off];
            const double st = fma(sr, yb, sb);
            const double ist = 1.0 / st;
            const double er = yo - yb;
            const double q = er * ist;
            ll -= LOG_2PI_HALF + log(st) + 0.5 * q * q;
            const double q2i = q * q * ist;   // e^2 / st^3
            w = q * ist - sr * ist + sr * q2i;
            if (cam) {
              dsig[off] += q2i - ist;
              dsig[off + 1] += (q2i - ist) * yb;
            } else {
              dsig[off] += (q2i - ist) * yb;
            }
          }
#pragma unroll
          for (int m = 0; m < MCS; ++m) gacc[m] = fma(w, dyb[m], gacc[m]);
        }
        ++r;
        continue;
      }
      if (t_edge <= t) {
        // ---- dose-rate edge: switch the input, restart the step size -----
        ++seg;
        const double u_old = u;
        u = a.seg_rate[seg];
        t_edge = (seg + 1 < seg_end) ? a.seg_t[seg + 1] : INFINITY;
        {
          // restart from the first accepted step after the previous edge of
          // this kind, if there was one (same policy as the streamed solver)
          edge_kind = (fabs(u) > fabs(u_old)) ? 1 : 2;
          const float hm = edge_kind == 1 ? h_mem_up : h_mem_down;
          h = (hm > 0.0f) ? fmin(static_cast<double>(hm), h)
                          : (CHI_EDGE_KEEP_H ? h * CHI_EDGE_SHRINK : 0.0);
        }
        continue;
      }
      const double t_stop = fmin(t_rec, t_edge);

      if (h == 0.0) {
        // first step: a fraction of the distance to the first stop (same
        // policy as the streamed solver, integrator_stream.cuh)
        h = CHI_H0_FRACTION * (t_stop - t);
        after_reject = false;
      }

      // ---- choose the step: hit t_stop exactly ----------------------------
      const double rem = t_stop - t;
      double hs = h;
      bool clipped = false;
      if (rem <= CHI_CLIP_STRETCH * h) {
        hs = rem;
        clipped = true;
      } else if (rem < 2.0 * h) {
        hs = 0.5 * rem;
      }

      // ---- one RODAS4 step of size hs on [y | S] --------------------------
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
      double Uy[6][N];
      double Us[MCS][6][N];

#define CHI_STAGE(I, CNT, a0, a1, a2, a3, a4, c0, c1, c2, c3, c4)            \
  {                                                                           \
    double Zy[N], f[N], cy[N], ry[N];                                         \
    lincomb<CNT, N>(Zy, y, Uy, a0, a1, a2, a3, a4);                           \
    M::rhs(Zy, c, u, f);                                                      \
    lincomb<CNT, N>(cy, nullptr, Uy, c0, c1, c2, c3, c4);                     \
    _Pragma("unroll") for (int n = 0; n < N; ++n) ry[n] =                     \
        fma(hinv, cy[n], f[n]);                                               \
    matvec<N>(Wi, ry, Uy[I]);                                                 \
    if (MC > 0) {                                                             \
      double Ji[N * N];                                                       \
      if (!M::LINEAR) M::jac(Zy, c, Ji);                                      \
      _Pragma("unroll") for (int m = 0; m < MC; ++m) {                        \
        double Zs[N], r1[N], r2[N], cs[N], rs[N];                             \
        lincomb<CNT, N>(Zs, S[m], Us[m], a0, a1, a2, a3, a4);                 \
        M::dfdc(Zy, c, e[m], r1);                                             \
        M::hess(y, c, S[m], e[m], Uy[I], r2);                                 \
        lincomb<CNT, N>(cs, nullptr, Us[m], c0, c1, c2, c3, c4);              \
        matvec<N>(M::LINEAR ? J0 : Ji, Zs, rs);                               \
        _Pragma("unroll") for (int n = 0; n < N; ++n) rs[n] =                 \
            fma(hinv, cs[n], rs[n] + r1[n] + r2[n]);                          \
        matvec<N>(Wi, rs, Us[m][I]);                                          \
      }                                                                       \
    }                                                                         \
  }
      CHI_STAGE(0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0)
      CHI_STAGE(1, 1, A21, 0, 0, 0, 0, C21, 0, 0, 0, 0)
      CHI_STAGE(2, 2, A31, A32, 0, 0, 0, C31, C32, 0, 0, 0)
      CHI_STAGE(3, 3, A41, A42, A43, 0, 0, C41, C42, C43, 0, 0)
      CHI_STAGE(4, 4, A51, A52, A53, A54, 0, C51, C52, C53, C54, 0)
      CHI_STAGE(5, 5, A51, A52, A53, A54, 1.0, C61, C62, C63, C64, C65)
#undef CHI_STAGE

      // new values and error estimate (= last stage increment)
      double yn[N];
      lincomb<5, N>(yn, y, Uy, A51, A52, A53, A54, 1.0);
      float errsq = 0.0f;
      {
        double s2 = 0.0;
#pragma unroll
        for (int n = 0; n < N; ++n) {
          yn[n] += Uy[5][n];
          const double sc =
              a.atol + a.rtol * fmax(fabs(y[n]), fabs(yn[n]));
          const double q = Uy[5][n] / sc;
          s2 = fma(q, q, s2);
        }
        errsq = static_cast<float>(s2 * (1.0 / N));
      }
      double Sn[MCS][N];
#pragma unroll
      for (int m = 0; m < MC; ++m) {
        lincomb<5, N>(Sn[m], S[m], Us[m], A51, A52, A53, A54, 1.0);
        double s2 = 0.0;
#pragma unroll
        for (int n = 0; n < N; ++n) {
          Sn[m][n] += Us[m][5][n];
          const double sc =
              a.atol + a.rtol * fmax(fabs(S[m][n]), fabs(Sn[m][n]));
          const double q = Us[m][5][n] / sc;
          s2 = fma(q, q, s2);
        }
        errsq = fmaxf(errsq, static_cast<float>(s2 * (1.0 / N)));
      }
      errsq = group_max<G>(errsq, gmask, lane, base);

      // errsq = err^2 ; err^(-1/4) = errsq^(-1/8)
      const bool accept = errsq <= 1.0f;  // false for NaN
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
#pragma unroll
        for (int m = 0; m < MC; ++m)
#pragma unroll
          for (int n = 0; n < N; ++n) S[m][n] = Sn[m][n];
        if (edge_kind == 1) h_mem_up = static_cast<float>(hs);
        if (edge_kind == 2) h_mem_down = static_cast<float>(hs);
        edge_kind = 0;
        if (after_reject) fac = fminf(fac, 1.0f);
        const double hn = hs * static_cast<double>(fac);
        // a step shortened to hit a stop must not shrink the working step
        h = (hs < h) ? fmax(hn, h) : hn;
        after_reject = false;
        ++n_acc;
      } else {
        h = hs * static_cast<double>(fminf(fac, 0.9f));
        after_reject = true;
        ++n_rej;
        if (h <= 1e-14 * fmax(fabs(t), 1.0)) {
          status = (errsq == errsq) ? ST_STEP_UNDERFLOW : ST_NON_FINITE;
          break;
        }
      }
      if (n_acc + n_rej >= a.max_steps) {
        status = ST_MAX_STEPS;
        break;
      }
    }
  }

  // ---- epilogue -------------------------------------------------------------
  if (sub == 0) {
    a.status[b] = status;
    if (a.n_steps) a.n_steps[b] = n_acc;
    if (a.n_rej) a.n_rej[b] = n_rej;
  }
  if (loglik) {
    const bool fail = invalid || status != ST_OK;
    if (sub == 0) {
      a.ll[b] = fail ? -INFINITY : ll;
    }
    if (a.with_sens && a.grad) {
      double* gb = a.grad + b * a.grad_stride;
#pragma unroll
      for (int m = 0; m < MC; ++m)
        if (cp[m] >= 0) gb[cp[m]] = fail ? INFINITY : gacc[m];
      if (sub == 0) {
        for (int k = 0; k < a.n_sigma; ++k)
          gb[P + k] = fail ? INFINITY : dsig[k];
      }
    }
  } else if (status != ST_OK) {
    // failed simulate: poison the remaining rows so nobody reads stale data
    for (int64_t q = r; q < r_end; ++q) {
      const int64_t row = trow + (q - r_begin);
      if (sub == 0) a.traj_y[row] = NAN;
    }
  }
}

}  // namespace chi_b200
#include "integrator_stream.cuh"
namespace chi_b200 {

#ifdef __CUDACC__

constexpr int SOLVE_BLOCK = 128;
constexpr int STREAM_BLOCK = 128;

template <class M, int G, int MC>
__global__ void __launch_bounds__(SOLVE_BLOCK)
solve_kernel(const SolveArgs a) {
  constexpr int SPW = 32 / G;
  const int lane = threadIdx.x & 31;
  const int grp = lane / G;
  const int base = grp * G;
  const int sub = lane - base;
  const int64_t warp =
      (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int64_t b = warp * SPW + grp;
  if (grp >= SPW || b >= a.B) return;
  const unsigned gmask =
      (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << base);
  solve_system<M, G, MC>(a, b, sub, lane, base, gmask);
}

// ---- registration -----------------------------------------------------------

template <class M, class... Vs>
struct Registrar {
  static constexpr int NV = sizeof...(Vs);
  static inline int32_t lanes[NV > 0 ? NV : 1] = {};
  static inline int32_t cols[NV > 0 ? NV : 1] = {};
  static inline ModelVTable vt = {};

  // Variant<0, CB>: streamed-columns persistent kernel
  // (integrator_stream.cuh)
  // resident blocks of a persistent kernel on the current device (cached)
  template <class K>
  static cudaError_t resident_blocks(K kernel, const int block,
                                     const size_t smem, int* cache,
                                     int* out) {
    constexpr int MAX_DEV = 16;
    int dev = 0;
    cudaError_t err = cudaGetDevice(&dev);
    if (err != cudaSuccess) return err;
    if (dev < 0 || dev >= MAX_DEV) return cudaErrorInvalidDevice;
    if (!cache[dev]) {
      if (smem > 48 * 1024) {
        err = cudaFuncSetAttribute(
            kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
            static_cast<int>(smem));
        if (err != cudaSuccess) return err;
      }
      int bpsm = 0, sms = 0;
      err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(
          &bpsm, kernel, block, smem);
      if (err != cudaSuccess) return err;
      err = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
      if (err != cudaSuccess) return err;
      // CHI_B200_MAX_BLOCKS_PER_SM: occupancy experiments (same binary, fewer
      // resident warps)
      if (const char* e = std::getenv("CHI_B200_MAX_BLOCKS_PER_SM")) {
        const int cap = std::atoi(e);
        if (cap > 0 && cap < bpsm) bpsm = cap;
      }
      cache[dev] = bpsm * sms;
    }
    *out = cache[dev];
    return cudaSuccess;
  }

  // the lanes of every warp of this launch integrate ONE individual under 32
  // parameter vectors: the warp-synchronous kernel applies
  // (CHI_B200_WARP_SYNC=0 keeps the general kernel, for comparisons)
  template <int CB>
  static bool warp_uniform(const SolveArgs& args) {
    static const bool enabled = [] {
      const char* e = std::getenv("CHI_B200_WARP_SYNC");
      return !(e && std::atoi(e) == 0);
    }();
    if (!(stream_is_rosl(CB) && M::LINEAR && args.mode == 1 && !args.snap))
      return false;
    // warp-per-system launches (SolveArgs::split) always run on it
    if (args.split > 0) return true;
    // (at most an eighth of the lanes padding when the number of vectors is
    // not a multiple of 32)
    const int64_t padded = (static_cast<int64_t>(args.order_inner) + 31) /
                           32 * 32;
    return enabled && args.refill_wait == INT32_MAX &&
           args.order_inner >= 32 &&
           (padded - args.order_inner) * 8 <= args.order_inner &&
           static_cast<int64_t>(args.order_inner) * args.order_stride ==
               args.B;
  }

  // Variant<0, CB>: streamed-columns persistent kernel
  // (integrator_stream.cuh)
  template <int CB>
  static cudaError_t launch_stream(const SolveArgs& args,
                                   cudaStream_t stream) {
    constexpr size_t smem =
        sizeof(double) * stream_slots<M, CB>() * STREAM_BLOCK;
    static int resident[16] = {};
    static int resident_warp[16] = {};
    static int resident_split[16] = {};
    static int resident_fwd[16] = {};
    static int resident_sub[16] = {};
    if (!args.work_counter) return cudaErrorInvalidValue;
    bool uniform = false;
    if constexpr (stream_is_rosl(CB) && M::LINEAR)
      uniform = warp_uniform<CB>(args);
    int res = 0;
    cudaError_t err;
    constexpr size_t smem_warp =
        sizeof(double) * stream_slots<M, CB, false>() * WARP_BLOCK;
    constexpr size_t smem_fwd =
        sizeof(double) * stream_slots<M, CB, false, true>() * WARP_BLOCK;
    // batched forward-only launches have an instantiation of their own
    const bool fwd = uniform && args.split == 0 && !args.with_sens;
    // ... and so have launches that integrate only some of the columns
    bool subset = false;
    if constexpr (stream_is_rosl(CB) && M::LINEAR) {
      using S1 = StreamSolver<M, CB, 1>;
      if (uniform && args.split == 0 && args.with_sens) {
        subset = args.n_int != S1::PI;
        for (int m = 0; m < S1::PI && !subset; ++m)
          subset = args.int_param[m] != S1::full_param(m);
      }
    }
    if constexpr (stream_is_rosl(CB) && M::LINEAR) {
      if (uniform && args.split > 0)
        err = resident_blocks(solve_warp_kernel<M, CB, 2>, WARP_BLOCK,
                              smem_warp, resident_split, &res);
      else if (fwd)
        err = resident_blocks(solve_warp_kernel<M, CB, 3>, WARP_BLOCK,
                              smem_fwd, resident_fwd, &res);
      else if (subset)
        err = resident_blocks(solve_warp_kernel<M, CB, 4>, WARP_BLOCK,
                              smem_warp, resident_sub, &res);
      else if (uniform)
        err = resident_blocks(solve_warp_kernel<M, CB, 1>, WARP_BLOCK,
                              smem_warp, resident_warp, &res);
    }
    if (!uniform)
      err = resident_blocks(solve_stream_kernel<M, CB, STREAM_BLOCK>,
                            STREAM_BLOCK, smem, resident, &res);
    if (err != cudaSuccess) return err;
    if (args.B == 0) return cudaSuccess;
    unsigned long long* counter = args.work_counter;
    err = cudaMemsetAsync(counter, 0, sizeof(unsigned long long), stream);
    if (err != cudaSuccess) return err;
    SolveArgs sorted = args;
    if constexpr (stream_is_rosl(CB) && M::LINEAR) {
      // batched launch: the vectors of an individual go to the warps in the
      // order of a stiffness key (order_keys_kernel / order_sort_kernel);
      // CHI_B200_SORT=0 keeps the batch order (comparisons)
      static const bool sort_enabled = [] {
        const char* e = std::getenv("CHI_B200_SORT");
        return !(e && std::atoi(e) == 0);
      }();
      sorted.perm = nullptr;
      if (uniform && args.split == 0 && sort_enabled && args.perm &&
          args.order_keys && args.order_inner > 32 &&
          args.order_inner <= ORDER_MAX_VECTORS) {
        int p2 = 64;
        while (p2 < args.order_inner) p2 <<= 1;
        order_keys_kernel<M>
            <<<static_cast<unsigned>((args.B + 255) / 256), 256, 0, stream>>>(
                args);
        const int threads = p2 / 2 < 1024 ? p2 / 2 : 1024;
        order_sort_kernel<0>
            <<<static_cast<unsigned>(args.order_stride), threads,
               static_cast<size_t>(p2) * 8, stream>>>(
                args.order_inner, p2, args.order_keys, args.perm);
        err = cudaGetLastError();
        if (err != cudaSuccess) return err;
        sorted.perm = args.perm;
      }
    }
    if constexpr (stream_is_rosl(CB) && M::LINEAR) {
      if (uniform) {
        // (warp-per-system launches: one warp per system)
        const int64_t threads =
            args.split > 0
                ? args.B * 32
                : (static_cast<int64_t>(args.order_inner) + 31) / 32 * 32 *
                      args.order_stride;
        int64_t blocks = (threads + WARP_BLOCK - 1) / WARP_BLOCK;
        if (blocks > res) blocks = res;
        if (args.split > 0)
          solve_warp_kernel<M, CB, 2>
              <<<static_cast<unsigned>(blocks), WARP_BLOCK, smem_warp,
                 stream>>>(sorted, counter);
        else if (fwd)
          solve_warp_kernel<M, CB, 3>
              <<<static_cast<unsigned>(blocks), WARP_BLOCK, smem_fwd,
                 stream>>>(sorted, counter);
        else if (subset)
          solve_warp_kernel<M, CB, 4>
              <<<static_cast<unsigned>(blocks), WARP_BLOCK, smem_warp,
                 stream>>>(sorted, counter);
        else
          solve_warp_kernel<M, CB, 1>
              <<<static_cast<unsigned>(blocks), WARP_BLOCK, smem_warp,
                 stream>>>(sorted, counter);
      }
    }
    if (!uniform) {
      const int64_t items = args.B;
      int64_t blocks = (items + STREAM_BLOCK - 1) / STREAM_BLOCK;
      if (blocks > res) blocks = res;
      solve_stream_kernel<M, CB, STREAM_BLOCK>
          <<<static_cast<unsigned>(blocks), STREAM_BLOCK, smem, stream>>>(
              args, counter);
    }
    err = cudaGetLastError();
    if (err != cudaSuccess) return err;
    if (args.mode == 1 && args.snap) {
      // deferred observation: error model and gradient from the snapshots
      observe_kernel<M><<<static_cast<unsigned>((args.B + 127) / 128), 128, 0,
                          stream>>>(args);
      err = cudaGetLastError();
    }
    return err;
  }

  template <int G, int MC>
  static cudaError_t launch_one(const SolveArgs& args, cudaStream_t stream) {
    if constexpr (G == 0) {
      return launch_stream<MC>(args, stream);
    } else {
      return launch_lanes<G, MC>(args, stream);
    }
  }

  template <int G, int MC>
  static cudaError_t launch_lanes(const SolveArgs& args,
                                  cudaStream_t stream) {
    const int64_t spw = 32 / G;
    const int64_t warps = (args.B + spw - 1) / spw;
    const int64_t wpb = SOLVE_BLOCK / 32;
    const int64_t blocks = (warps + wpb - 1) / wpb;
    if (blocks == 0) return cudaSuccess;
    solve_kernel<M, G, MC><<<static_cast<unsigned>(blocks), SOLVE_BLOCK, 0,
                             stream>>>(args);
    return cudaGetLastError();
  }

  static int pick(const SolveArgs& args, int variant) {
    auto fits = [&](int v) {
      return lanes[v] == 0 || lanes[v] * cols[v] >= args.n_cols;
    };
    if (variant >= 0 && variant < NV && fits(variant)) return variant;
    // default: first variant with enough column slots
    for (int v = 0; v < NV; ++v)
      if (fits(v)) return v;
    return -1;
  }

  static cudaError_t launch(const SolveArgs& args, int variant,
                            cudaStream_t stream) {
    if (!args.with_sens) {
      // forward-only solves run on the same streamed kernel (no columns)
      // unless a lane-parallel variant was asked for explicitly
      const bool lanes_forward =
          variant >= 0 && variant < NV && lanes[variant] != 0;
      if (lanes_forward) return launch_lanes<1, 0>(args, stream);
      cudaError_t err = cudaErrorInvalidValue;
      int idx = 0;
      int first_stream = -1;
      for (int v = 0; v < NV; ++v)
        if (lanes[v] == 0) {
          first_stream = v;
          break;
        }
      if (variant >= 0 && variant < NV && lanes[variant] == 0)
        first_stream = variant;
      if (first_stream < 0) return launch_lanes<1, 0>(args, stream);
      (void)std::initializer_list<int>{
          (idx++ == first_stream
               ? (err = launch_one<Vs::G, Vs::MC>(args, stream), 0)
               : 0)...};
      return err;
    }
    const int v = pick(args, variant);
    if (v < 0) return cudaErrorInvalidValue;
    cudaError_t err = cudaErrorInvalidValue;
    int idx = 0;
    // fold over the variant list
    (void)std::initializer_list<int>{
        (idx++ == v ? (err = launch_one<Vs::G, Vs::MC>(args, stream), 0)
                    : 0)...};
    return err;
  }

  template <int G, int MC>
  static cudaError_t describe_one(int* regs, int* block, int* bpsm,
                                  int* local_bytes) {
    cudaFuncAttributes at;
    if constexpr (G == 0 && stream_is_rosl(MC) && M::LINEAR) {
      // the kernel of the batched launches (warp-synchronous form)
      constexpr size_t smem =
          sizeof(double) * stream_slots<M, MC, false>() * WARP_BLOCK;
      cudaError_t err =
          cudaFuncGetAttributes(&at, solve_warp_kernel<M, MC, 1>);
      if (err != cudaSuccess) return err;
      *regs = at.numRegs;
      *block = WARP_BLOCK;
      *local_bytes = static_cast<int>(at.localSizeBytes);
      if (smem > 48 * 1024) {
        err = cudaFuncSetAttribute(
            solve_warp_kernel<M, MC, 1>,
            cudaFuncAttributeMaxDynamicSharedMemorySize,
            static_cast<int>(smem));
        if (err != cudaSuccess) return err;
      }
      return cudaOccupancyMaxActiveBlocksPerMultiprocessor(
          bpsm, solve_warp_kernel<M, MC, 1>, WARP_BLOCK, smem);
    } else if constexpr (G == 0) {
      constexpr size_t smem =
          sizeof(double) * stream_slots<M, MC>() * STREAM_BLOCK;
      cudaError_t err = cudaFuncGetAttributes(
          &at, solve_stream_kernel<M, MC, STREAM_BLOCK>);
      if (err != cudaSuccess) return err;
      *regs = at.numRegs;
      *block = STREAM_BLOCK;
      *local_bytes = static_cast<int>(at.localSizeBytes);
      if (smem > 48 * 1024) {
        err = cudaFuncSetAttribute(
            solve_stream_kernel<M, MC, STREAM_BLOCK>,
            cudaFuncAttributeMaxDynamicSharedMemorySize,
            static_cast<int>(smem));
        if (err != cudaSuccess) return err;
      }
      return cudaOccupancyMaxActiveBlocksPerMultiprocessor(
          bpsm, solve_stream_kernel<M, MC, STREAM_BLOCK>, STREAM_BLOCK,
          smem);
    } else {
      cudaError_t err = cudaFuncGetAttributes(&at, solve_kernel<M, G, MC>);
      if (err != cudaSuccess) return err;
      *regs = at.numRegs;
      *block = SOLVE_BLOCK;
      *local_bytes = static_cast<int>(at.localSizeBytes);
      return cudaOccupancyMaxActiveBlocksPerMultiprocessor(
          bpsm, solve_kernel<M, G, MC>, SOLVE_BLOCK, 0);
    }
  }

  static cudaError_t describe(int variant, int* regs, int* block, int* bpsm,
                              int* local_bytes) {
    if (variant < 0) return describe_one<1, 0>(regs, block, bpsm, local_bytes);
    cudaError_t err = cudaErrorInvalidValue;
    int idx = 0;
    (void)std::initializer_list<int>{
        (idx++ == variant
             ? (err = describe_one<Vs::G, Vs::MC>(regs, block, bpsm,
                                                  local_bytes),
                0)
             : 0)...};
    return err;
  }

  Registrar() {
    int i = 0;
    (void)std::initializer_list<int>{
        (lanes[i] = Vs::G, cols[i] = Vs::MC, ++i, 0)...};
    vt.key = M::key();
    vt.n_states = M::N;
    vt.n_consts = M::NC;
    vt.n_outputs = M::NOUT;
    // bit 0: linear in the states; bit 1: rank-one column terms (COL_RANK1)
    vt.linear = (M::LINEAR ? 1 : 0) | ((M::LINEAR && M::COL_RANK1) ? 2 : 0);
    vt.flops_rhs = M::FLOPS_RHS;
    vt.flops_jac = M::FLOPS_JAC;
    vt.flops_dfdc = M::FLOPS_DFDC;
    vt.flops_hess = M::FLOPS_HESS;
    vt.inert_mask = M::INERT_MASK;
    vt.flops_col_sum = M::FLOPS_COL_SUM;
    vt.flops_hess_common = M::FLOPS_HESS_COMMON;
    vt.n_variants = NV;
    vt.variant_lanes = lanes;
    vt.variant_cols = cols;
    vt.launch = &launch;
    vt.describe = &describe;
    chi_b200_register_model(&vt);
  }
};

#define CHI_B200_REGISTER_MODEL(MODEL, ...)                                   \
  static chi_b200::Registrar<MODEL, __VA_ARGS__> chi_b200_registrar_##MODEL;

#endif  // __CUDACC__

}  // namespace chi_b200
