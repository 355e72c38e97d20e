// chi_b200 -- the thread-per-system solve kernels (the default kernels).
//
// StreamSolver<M, METHOD, FORM> integrates ONE system -- an individual under
// one parameter vector -- per thread: a stiffly accurate Rosenbrock method on
// [y | S] with the exact block-triangular Jacobian, exact hits of dose edges
// and observation times, per-system scratch in shared memory (slot-major,
// conflict free: sm[slot * stride + tid]).
//
//   METHOD  9 / 11 = ROSL(9) / ROSL(11), the methods for LINEAR models: state
//           and all sensitivity columns advance together in registers, one
//           N x N product and one accumulation per stage (step_rosl);
//           1 / 5 = RODAS4 / RODAS5 (nonlinear models): the state column
//           first, then the sensitivity columns streamed through registers
//           one at a time in a uniform loop (step_rodas, column).
//   FORM    0  general event loop (solve_stream_kernel): persistent grid, the
//              lanes pull work from a global counter and handle their own
//              events; observation fused, or deferred to observe_kernel via
//              snapshots of (y, S) at the records;
//           1 / 3 / 4  warp-synchronous batched form (solve_warp_kernel,
//              solve_uniform): the 32 lanes of a warp integrate one
//              individual under 32 parameter vectors and walk through its
//              stops together, observation fused (1: all columns of the
//              model, 3: none, 4: a run-time subset);
//           2  warp-per-system form of small launches: the columns of a
//              system side by side on the lanes of one warp
//              (SolveArgs::split).
//
// Compiles for the host as well (stride = 1, tid = 0) for the CPU port.
#pragma once
#include "integrator.cuh"
#include <type_traits>

// unroll factor of the stage loop of step_rosl (stages 1 .. S-3)
#ifndef CHI_ROSL_UNROLL
#define CHI_ROSL_UNROLL 4
#endif
#define CHI_STR2(x) #x
#define CHI_STR(x) CHI_STR2(x)
#define CHI_UNROLL_N(n) _Pragma(CHI_STR(unroll n))

// vote over the lanes of a warp (the host build has one "lane")
#ifdef __CUDA_ARCH__
#define CHI_WARP_ANY(p) __any_sync(0xffffffffu, (p))
#else
#define CHI_WARP_ANY(p) (p)
#endif

namespace chi_b200 {

// shared-memory doubles per thread: S of the integrated columns (current +
// candidate), gradient accumulators, the state column's stage values 2..8
// (the first is y), eight cold per-system scalars, the model constants and
// the state
template <class M>
constexpr int stream_integrated() {
  // parameters whose sensitivity column is integrated: all but the constants
  // that do not enter the rhs (their columns are identically zero)
  int n = M::P;
  for (int k = 0; k < M::NC; ++k) n -= (M::INERT_MASK >> k) & 1u;
  return n;
}
// METHOD: 1 / 5 = RODAS4 / RODAS5, 9 / 11 = ROSL(9) / ROSL(11).  The ROSL
// step keeps everything of a step in registers: no candidate copy of S and
// no stage values in shared memory.
constexpr bool stream_is_rosl(const int method) {
  return method == 9 || method == 11;
}
// ... and it keeps a one-record buffer of (y, S) per thread, so that the
// lanes of a warp evaluate their observation records together (see
// StreamSolver::events), and the weights of its error norm (floats).
// (WITH_BUF = false: the warp-synchronous kernel, which has no use for the
// record buffer)
// (FWD = true: forward-only instantiation, no sensitivity columns at all)
template <class M, int METHOD, bool WITH_BUF = stream_is_rosl(METHOD),
          bool FWD = false>
constexpr int stream_slots() {
  const int pi = FWD ? 0 : stream_integrated<M>();
  return (stream_is_rosl(METHOD)
              ? pi * M::N + (WITH_BUF ? (1 + pi) * M::N : 0) +
                    ((1 + pi) * M::N + 1) / 2
              : 2 * pi * M::N + 7 * M::N) +
         M::P + 8 + M::NCS + M::N;
}
// largest of the methods (host buffers of the CPU port)
template <class M>
constexpr int stream_slots() {
  return stream_slots<M, 5>();
}

namespace rodas5 {
// RODAS5 (Di Marzo 1993; Hairer & Wanner, "Solving ODEs II", sec. VI.4):
// 8 stages, order 5(4), stiffly accurate, transformed form.  Typed from
// memory and verified by a convergence-order test (observed order 5.00,
// embedded estimate ~ h^5; tests/test_host.py).
constexpr double GAMMA = 0.19;
constexpr double A21 = 2.0;
constexpr double A31 = 3.040894194418781;
constexpr double A32 = 1.041747909077569;
constexpr double A41 = 2.576417536461461;
constexpr double A42 = 1.622083060776640;
constexpr double A43 = -0.9089668560264532;
constexpr double A51 = 2.760842080225597;
constexpr double A52 = 1.446624659844071;
constexpr double A53 = -0.3036980084553738;
constexpr double A54 = 0.2877498600325443;
constexpr double A61 = -14.09640773051259;
constexpr double A62 = 6.925207756232704;
constexpr double A63 = -41.47510893210728;
constexpr double A64 = 2.343771018586405;
constexpr double A65 = 24.13215229196062;
constexpr double C21 = -10.31323885133993;
constexpr double C31 = -21.04823117650003;
constexpr double C32 = -7.234992135176716;
constexpr double C41 = 32.22751541853323;
constexpr double C42 = -4.943732386540191;
constexpr double C43 = 19.44922031041879;
constexpr double C51 = -20.69865579590063;
constexpr double C52 = -8.816374604402768;
constexpr double C53 = 1.260436877740897;
constexpr double C54 = -0.7495647613787146;
constexpr double C61 = -46.22004352711257;
constexpr double C62 = -17.49534862857472;
constexpr double C63 = -289.6389582892057;
constexpr double C64 = 93.60855400400906;
constexpr double C65 = 318.3822534212147;
constexpr double C71 = 34.20013733472935;
constexpr double C72 = -14.15535402717690;
constexpr double C73 = 57.82335640988400;
constexpr double C74 = 25.83362985412365;
constexpr double C75 = 1.408950972071624;
constexpr double C76 = -6.551835421242162;
constexpr double C81 = 42.57076742291101;
constexpr double C82 = -13.80770672017997;
constexpr double C83 = 93.98938432427124;
constexpr double C84 = 18.77919633714503;
constexpr double C85 = -31.58359187223370;
constexpr double C86 = -6.685968952921985;
constexpr double C87 = -5.810979938412932;
// K_ij = (alpha_ij + gamma_ij) / gamma of the method in its original k-stage
// form, derived from the a, C tables above (tools/derive_kform.py).  For a
// LINEAR system z' = J z + g the stage equations collapse to
//     kappa_i = E arg_i + W^-1 g_i,   arg_i = z + sum_j K_ij kappa_j,
//     E = W^-1 J,  W = I / (gamma h) - J,  kappa_i = gamma k_i,
// i.e. ONE combination and one product per stage; arg_i + kappa_i equals
// Z_i + U_i of the transformed form, the new solution is arg_8 + kappa_8 and
// the embedded one arg_7 + kappa_7.  K72 = K82 = 0 (to rounding of the
// tables), and row 8 repeats row 7 in its first five entries:
//     arg_8 = arg_7 + (K86 - K76) kappa_6 + K87 kappa_7.
constexpr double K21 = 0.0404846182454133;
constexpr double K31 = -0.30594588947147515;
constexpr double K32 = -0.33290059660600704;
constexpr double K41 = 3.7240608400705508;
constexpr double K42 = -3.147526073234625;
constexpr double K43 = 2.7863850029531169;
constexpr double K51 = -0.18407908723123721;
constexpr double K52 = -1.014987413588082;
constexpr double K53 = 0.47283992697435224;
constexpr double K54 = 0.14533255537058853;
constexpr double K61 = 39.922429989489724;
constexpr double K62 = -81.969800437891473;
constexpr double K63 = -46.391615240042746;
constexpr double K64 = 8.0773631411899189;
constexpr double K65 = 84.624780441991413;
constexpr double K71 = 2.4452730732036673;
constexpr double K73 = -9.0574079412504045;
constexpr double K74 = 1.5318316952410377;
constexpr double K75 = 9.5883097975785527;
constexpr double K76 = -0.24484873003601078;
constexpr double K8D6 = 0.10408618829845701;   // K86 - K76
constexpr double K87 = -0.10408618829845708;
}  // namespace rodas5


namespace rosl {
// ROSL(s): s-stage Rosenbrock methods for LINEAR systems z' = J z + g (g
// constant over a step; all PK models of the library, and their augmented
// sensitivity systems, are of this form between two dose-rate edges).  For a
// linear system every Rosenbrock method with a single gamma reduces to
//     z_1 = R(hJ) z_0 + h Q(hJ) g,  R(x) = 1 + x Q(x) = P(x) / (1 - gamma x)^s,
// and only the linear order conditions R(x) = exp(x) + O(x^(p+1)) remain: with
// P the degree s-1 section of exp(x) (1 - gamma x)^s the method has order
// p = s - 1 (restricted Pade approximant, Norsett) and R(inf) = 0.  Written
// in powers of E = T - I = gamma h J T, T = (I - gamma h J)^-1 (the kernel's
// E = W^-1 J):
//     z_1 = z_0 + sum_k c_k w_k,   w_0 = T (h f(z_0)),   w_k = E w_{k-1},
// ONE N x N product and one accumulation per stage, nothing of the previous
// stages kept.  E^k = O(h^k): the terms decay, the c_k are O(1), and the last
// two terms are the difference to the embedded method of order s - 2:
//     err = c_{s-2} w_{s-2} + c_{s-1} w_{s-1}.
// Tables and checks (order, A-stability: max |R(iy)| = 1): tools/derive_rosl.py.
// Against RODAS5 on the bench workload: 72 instead of 150 step attempts per
// system at a 40 times smaller global error (tests/test_host.py).
template <int S>
struct Table;
template <>
struct Table<9> {   // order 8(7), A-stable for gamma in [0.171, 0.261]
  static constexpr double GAMMA = 0.2;
  static constexpr double C[9] = {
      1.0, 1.5, 0.16666666666666666667, -0.79166666666666666667, 0.375,
      0.21527777777777777778, -0.48313492063492063492,
      0.38206845238095238095, 0.44714781746031746032};
};
template <>
struct Table<11> {  // order 10(9), A-stable for gamma in [0.147, 0.165]
  static constexpr double GAMMA = 0.16;
  static constexpr double C[11] = {
      1.0, 2.125, 1.2604166666666666667, -0.98372395833333333333,
      -0.41194661458333333333, 0.91328260633680555556,
      -0.45594884478856646825, -0.24326151136368040055,
      0.6181601832573165759, -0.53172900623640265625,
      -0.57111326014869427555};
};
}  // namespace rosl

// Device copies of the coefficients in constant memory: the FP64 instructions
// then take them as constant-bank operands instead of materialising 64-bit
// immediates with two UMOVs per use (19 percent of the issued instructions of
// the hot loop before this change, ncu source view).
#ifdef __CUDACC__
namespace rodas4d {
static __constant__ double GAMMA = 0.25;
static __constant__ double A21 = 0.1544000000000000e+01;
static __constant__ double A31 = 0.9466785280815826e+00;
static __constant__ double A32 = 0.2557011698983284e+00;
static __constant__ double A41 = 0.3314825187068521e+01;
static __constant__ double A42 = 0.2896124015972201e+01;
static __constant__ double A43 = 0.9986419139977817e+00;
static __constant__ double A51 = 0.1221224509226641e+01;
static __constant__ double A52 = 0.6019134481288629e+01;
static __constant__ double A53 = 0.1253708332932087e+02;
static __constant__ double A54 = -0.6878860361058950e+00;
static __constant__ double C21 = -0.5668800000000000e+01;
static __constant__ double C31 = -0.2430093356833875e+01;
static __constant__ double C32 = -0.2063599157091915e+00;
static __constant__ double C41 = -0.1073529058151375e+00;
static __constant__ double C42 = -0.9594562251023355e+01;
static __constant__ double C43 = -0.2047028614809616e+02;
static __constant__ double C51 = 0.7496443313967647e+01;
static __constant__ double C52 = -0.1024680431464352e+02;
static __constant__ double C53 = -0.3399990352819905e+02;
static __constant__ double C54 = 0.1170890893206160e+02;
static __constant__ double C61 = 0.8083246795921522e+01;
static __constant__ double C62 = -0.7981132988064893e+01;
static __constant__ double C63 = -0.3152159432874371e+02;
static __constant__ double C64 = 0.1631930543123136e+02;
static __constant__ double C65 = -0.6058818238834054e+01;
static __constant__ double K21 = rodas4::K21;
static __constant__ double K31 = rodas4::K31;
static __constant__ double K32 = rodas4::K32;
static __constant__ double K41 = rodas4::K41;
static __constant__ double K42 = rodas4::K42;
static __constant__ double K43 = rodas4::K43;
static __constant__ double K51 = rodas4::K51;
static __constant__ double K52 = rodas4::K52;
static __constant__ double K53 = rodas4::K53;
static __constant__ double K54 = rodas4::K54;
static __constant__ double K61 = rodas4::K61;
static __constant__ double K62 = rodas4::K62;
static __constant__ double K63 = rodas4::K63;
static __constant__ double K64 = rodas4::K64;
static __constant__ double K65 = rodas4::K65;
}  // namespace rodas4d
namespace rodas5d {
static __constant__ double GAMMA = 0.19;
static __constant__ double A21 = 2.0;
static __constant__ double A31 = 3.040894194418781;
static __constant__ double A32 = 1.041747909077569;
static __constant__ double A41 = 2.576417536461461;
static __constant__ double A42 = 1.622083060776640;
static __constant__ double A43 = -0.9089668560264532;
static __constant__ double A51 = 2.760842080225597;
static __constant__ double A52 = 1.446624659844071;
static __constant__ double A53 = -0.3036980084553738;
static __constant__ double A54 = 0.2877498600325443;
static __constant__ double A61 = -14.09640773051259;
static __constant__ double A62 = 6.925207756232704;
static __constant__ double A63 = -41.47510893210728;
static __constant__ double A64 = 2.343771018586405;
static __constant__ double A65 = 24.13215229196062;
static __constant__ double C21 = -10.31323885133993;
static __constant__ double C31 = -21.04823117650003;
static __constant__ double C32 = -7.234992135176716;
static __constant__ double C41 = 32.22751541853323;
static __constant__ double C42 = -4.943732386540191;
static __constant__ double C43 = 19.44922031041879;
static __constant__ double C51 = -20.69865579590063;
static __constant__ double C52 = -8.816374604402768;
static __constant__ double C53 = 1.260436877740897;
static __constant__ double C54 = -0.7495647613787146;
static __constant__ double C61 = -46.22004352711257;
static __constant__ double C62 = -17.49534862857472;
static __constant__ double C63 = -289.6389582892057;
static __constant__ double C64 = 93.60855400400906;
static __constant__ double C65 = 318.3822534212147;
static __constant__ double C71 = 34.20013733472935;
static __constant__ double C72 = -14.15535402717690;
static __constant__ double C73 = 57.82335640988400;
static __constant__ double C74 = 25.83362985412365;
static __constant__ double C75 = 1.408950972071624;
static __constant__ double C76 = -6.551835421242162;
static __constant__ double C81 = 42.57076742291101;
static __constant__ double C82 = -13.80770672017997;
static __constant__ double C83 = 93.98938432427124;
static __constant__ double C84 = 18.77919633714503;
static __constant__ double C85 = -31.58359187223370;
static __constant__ double C86 = -6.685968952921985;
static __constant__ double C87 = -5.810979938412932;
static __constant__ double K21 = rodas5::K21;
static __constant__ double K31 = rodas5::K31;
static __constant__ double K32 = rodas5::K32;
static __constant__ double K41 = rodas5::K41;
static __constant__ double K42 = rodas5::K42;
static __constant__ double K43 = rodas5::K43;
static __constant__ double K51 = rodas5::K51;
static __constant__ double K52 = rodas5::K52;
static __constant__ double K53 = rodas5::K53;
static __constant__ double K54 = rodas5::K54;
static __constant__ double K61 = rodas5::K61;
static __constant__ double K62 = rodas5::K62;
static __constant__ double K63 = rodas5::K63;
static __constant__ double K64 = rodas5::K64;
static __constant__ double K65 = rodas5::K65;
static __constant__ double K71 = rodas5::K71;
static __constant__ double K73 = rodas5::K73;
static __constant__ double K74 = rodas5::K74;
static __constant__ double K75 = rodas5::K75;
static __constant__ double K76 = rodas5::K76;
static __constant__ double K8D6 = rodas5::K8D6;
static __constant__ double K87 = rodas5::K87;
}  // namespace rodas5d
#endif
#ifdef __CUDACC__
namespace rosld {
static __constant__ double C9[9] = {
    1.0, 1.5, 0.16666666666666666667, -0.79166666666666666667, 0.375,
    0.21527777777777777778, -0.48313492063492063492, 0.38206845238095238095,
    0.44714781746031746032};
static __constant__ double C11[11] = {
    1.0, 2.125, 1.2604166666666666667, -0.98372395833333333333,
    -0.41194661458333333333, 0.91328260633680555556, -0.45594884478856646825,
    -0.24326151136368040055, 0.6181601832573165759, -0.53172900623640265625,
    -0.57111326014869427555};
}  // namespace rosld
#endif
// coefficient k of ROSL(S): a constant-bank operand on the device
template <int S>
CHI_HD double rosl_c(const int k) {
#ifdef __CUDA_ARCH__
  return S == 9 ? rosld::C9[k] : rosld::C11[k];
#else
  return rosl::Table<S>::C[k];
#endif
}
#ifdef __CUDA_ARCH__
#define CHI_R4(x) rodas4d::x
#define CHI_R5(x) rodas5d::x
#else
#define CHI_R4(x) rodas4::x
#define CHI_R5(x) rodas5::x
#endif

// Stage tables: X(stage, n_previous, incremental, a_1..a_7, c_1..c_7).
// "incremental": the stage value is the previous stage value plus the previous
// increment (the last rows of the stiffly accurate methods repeat the row
// before them with one more unit entry), so it costs N additions instead of a
// full combination.
#define CHI_RODAS4_TABLE(X)                                                   \
  X(0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0)                        \
  X(1, 1, 0, CHI_R4(A21), 0, 0, 0, 0, 0, 0, CHI_R4(C21), 0, 0, 0, 0, 0, 0)    \
  X(2, 2, 0, CHI_R4(A31), CHI_R4(A32), 0, 0, 0, 0, 0, CHI_R4(C31),            \
    CHI_R4(C32), 0, 0, 0, 0, 0)                                               \
  X(3, 3, 0, CHI_R4(A41), CHI_R4(A42), CHI_R4(A43), 0, 0, 0, 0, CHI_R4(C41),  \
    CHI_R4(C42), CHI_R4(C43), 0, 0, 0, 0)                                     \
  X(4, 4, 0, CHI_R4(A51), CHI_R4(A52), CHI_R4(A53), CHI_R4(A54), 0, 0, 0,     \
    CHI_R4(C51), CHI_R4(C52), CHI_R4(C53), CHI_R4(C54), 0, 0, 0)              \
  X(5, 5, 1, 0, 0, 0, 0, 0, 0, 0, CHI_R4(C61), CHI_R4(C62), CHI_R4(C63),      \
    CHI_R4(C64), CHI_R4(C65), 0, 0)

#define CHI_RODAS5_TABLE(X)                                                   \
  X(0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0)                        \
  X(1, 1, 0, CHI_R5(A21), 0, 0, 0, 0, 0, 0, CHI_R5(C21), 0, 0, 0, 0, 0, 0)    \
  X(2, 2, 0, CHI_R5(A31), CHI_R5(A32), 0, 0, 0, 0, 0, CHI_R5(C31),            \
    CHI_R5(C32), 0, 0, 0, 0, 0)                                               \
  X(3, 3, 0, CHI_R5(A41), CHI_R5(A42), CHI_R5(A43), 0, 0, 0, 0, CHI_R5(C41),  \
    CHI_R5(C42), CHI_R5(C43), 0, 0, 0, 0)                                     \
  X(4, 4, 0, CHI_R5(A51), CHI_R5(A52), CHI_R5(A53), CHI_R5(A54), 0, 0, 0,     \
    CHI_R5(C51), CHI_R5(C52), CHI_R5(C53), CHI_R5(C54), 0, 0, 0)              \
  X(5, 5, 0, CHI_R5(A61), CHI_R5(A62), CHI_R5(A63), CHI_R5(A64), CHI_R5(A65), \
    0, 0, CHI_R5(C61), CHI_R5(C62), CHI_R5(C63), CHI_R5(C64), CHI_R5(C65), 0, \
    0)                                                                        \
  X(6, 6, 1, 0, 0, 0, 0, 0, 0, 0, CHI_R5(C71), CHI_R5(C72), CHI_R5(C73),      \
    CHI_R5(C74), CHI_R5(C75), CHI_R5(C76), 0)                                 \
  X(7, 7, 1, 0, 0, 0, 0, 0, 0, 0, CHI_R5(C81), CHI_R5(C82), CHI_R5(C83),      \
    CHI_R5(C84), CHI_R5(C85), CHI_R5(C86), CHI_R5(C87))

// Tables of the linear-model bodies (k-stage form, see the K constants):
// X(stage, first, count, skip, k_1..k_7):
//     arg = (first == 0 ? z : arg of the previous stage)
//           + sum_{first <= j < count, bit j of skip clear} k_j kappa_j.
#define CHI_RODAS4_K_TABLE(X)                                                 \
  X(0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0)                                          \
  X(1, 0, 1, 0, CHI_R4(K21), 0, 0, 0, 0, 0, 0)                                \
  X(2, 0, 2, 0, CHI_R4(K31), CHI_R4(K32), 0, 0, 0, 0, 0)                      \
  X(3, 0, 3, 0, CHI_R4(K41), CHI_R4(K42), CHI_R4(K43), 0, 0, 0, 0)            \
  X(4, 0, 4, 0, CHI_R4(K51), CHI_R4(K52), CHI_R4(K53), CHI_R4(K54), 0, 0, 0)  \
  X(5, 0, 5, 0, CHI_R4(K61), CHI_R4(K62), CHI_R4(K63), CHI_R4(K64),           \
    CHI_R4(K65), 0, 0)

#define CHI_RODAS5_K_TABLE(X)                                                 \
  X(0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0)                                          \
  X(1, 0, 1, 0, CHI_R5(K21), 0, 0, 0, 0, 0, 0)                                \
  X(2, 0, 2, 0, CHI_R5(K31), CHI_R5(K32), 0, 0, 0, 0, 0)                      \
  X(3, 0, 3, 0, CHI_R5(K41), CHI_R5(K42), CHI_R5(K43), 0, 0, 0, 0)            \
  X(4, 0, 4, 0, CHI_R5(K51), CHI_R5(K52), CHI_R5(K53), CHI_R5(K54), 0, 0, 0)  \
  X(5, 0, 5, 0, CHI_R5(K61), CHI_R5(K62), CHI_R5(K63), CHI_R5(K64),           \
    CHI_R5(K65), 0, 0)                                                        \
  X(6, 0, 6, 2, CHI_R5(K71), 0, CHI_R5(K73), CHI_R5(K74), CHI_R5(K75),        \
    CHI_R5(K76), 0)                                                           \
  X(7, 5, 7, 0, 0, 0, 0, 0, 0, CHI_R5(K8D6), CHI_R5(K87))

// out = base + sum_{FIRST <= j < CNT, bit j of SKIP clear} k_j K[j]; `out`
// may alias `base`
template <int FIRST, int CNT, int SKIP, int N>
CHI_HD void lincombk(double* out, const double* base, const double (*K)[N],
                     double k0, double k1, double k2, double k3, double k4,
                     double k5, double k6) {
#pragma unroll
  for (int n = 0; n < N; ++n) {
    double v = base[n];
    if (FIRST <= 0 && CNT > 0 && !(SKIP & 1)) v = fma(k0, K[0][n], v);
    if (FIRST <= 1 && CNT > 1 && !(SKIP & 2)) v = fma(k1, K[1][n], v);
    if (FIRST <= 2 && CNT > 2 && !(SKIP & 4)) v = fma(k2, K[2][n], v);
    if (FIRST <= 3 && CNT > 3 && !(SKIP & 8)) v = fma(k3, K[3][n], v);
    if (FIRST <= 4 && CNT > 4 && !(SKIP & 16)) v = fma(k4, K[4][n], v);
    if (FIRST <= 5 && CNT > 5 && !(SKIP & 32)) v = fma(k5, K[5][n], v);
    if (FIRST <= 6 && CNT > 6 && !(SKIP & 64)) v = fma(k6, K[6][n], v);
    out[n] = v;
  }
}

// out = base + sum_{j<CNT} k_j U[j]   (base may be null), up to 7 terms
template <int CNT, int N>
CHI_HD void lincomb7(double* out, const double* base, const double (*U)[N],
                     double k0, double k1, double k2, double k3, double k4,
                     double k5, double k6) {
#pragma unroll
  for (int n = 0; n < N; ++n) {
    double v = base ? base[n] : 0.0;
    if (CNT > 0) v = fma(k0, U[0][n], v);
    if (CNT > 1) v = fma(k1, U[1][n], v);
    if (CNT > 2) v = fma(k2, U[2][n], v);
    if (CNT > 3) v = fma(k3, U[3][n], v);
    if (CNT > 4) v = fma(k4, U[4][n], v);
    if (CNT > 5) v = fma(k5, U[5][n], v);
    if (CNT > 6) v = fma(k6, U[6][n], v);
    out[n] = v;
  }
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

// What one observation contributes under the error model of its output
// (chi/_error_models.py: compute_log_likelihood / compute_sensitivities of the
// four models), written once for the fused handler and the deferred kernel.
struct ErrTerms {
  double ll;      // log-likelihood without the -log(scale) term
  double scale;   // the term -log(scale) is accumulated as a product
  double w;       // d ll / d ybar
  double ds0, ds1;  // d ll / d sigma (second entry: ConstantAndMultiplicative)
  bool invalid;   // log-normal model with ybar <= 0
};

CHI_HD ErrTerms error_terms(const int kind, const double* sg, const double yb,
                            const double yo, const double lo) {
  ErrTerms e;
  e.ds1 = 0.0;
  e.invalid = false;
  const double s0 = sg[0];
  if (kind == ERR_GAUSSIAN) {
    const double is = 1.0 / s0;
    const double q = (yo - yb) * is;
    e.ll = -(LOG_2PI_HALF + 0.5 * q * q);
    e.scale = s0;
    e.w = q * is;
    e.ds0 = (q * q - 1.0) * is;
  } else if (kind == ERR_LOGNORMAL) {
    e.invalid = yb <= 0.0;
    const double is = 1.0 / s0;
    const double q = (lo - log(yb) + 0.5 * s0 * s0) * is;
    e.ll = -(LOG_2PI_HALF + lo + 0.5 * q * q);
    e.scale = s0;
    e.w = q * is / yb;
    e.ds0 = (q * q - 1.0) * is - q;
  } else {
    const bool cam = kind == ERR_CAM;
    const double sb = cam ? s0 : 0.0;
    const double sr = cam ? sg[1] : s0;
    const double st = fma(sr, yb, sb);
    const double ist = 1.0 / st;
    const double q = (yo - yb) * ist;
    e.ll = -(LOG_2PI_HALF + 0.5 * q * q);
    e.scale = st;
    const double q2i = q * q * ist;
    e.w = q * ist - sr * ist + sr * q2i;
    if (cam) {
      e.ds0 = q2i - ist;
      e.ds1 = (q2i - ist) * yb;
    } else {
      e.ds0 = (q2i - ist) * yb;
    }
  }
  return e;
}

// State of one system while it is being integrated.  `init` loads a system,
// `events` handles what is pending at the current time (observation records,
// dose edges) and returns true once the system is finished, `step` attempts
// ONE step, `finish` writes the results.  Splitting the solve this way lets
// the persistent kernel refill a lane with the next system as soon as its
// current one is done, and re-join the warp between the (divergent) event
// handling and the (expensive) step (solve_stream_kernel below).
// METHOD_CODE: 1 = RODAS4 (6 stages, order 4(3)), 5 = RODAS5 (8 stages,
// order 5(4)).
// FORM: 0 = general event loop (per-lane event cursor, one-record buffer for
// the ROSL methods), 1 = warp-synchronous batched form (solve_uniform), 2 =
// warp-per-system form of small launches (solve_uniform over the split work
// items of one system, SolveArgs::split), 3 = FORM 1 without sensitivity
// columns, 4 = FORM 1 for a subset of the columns chosen at run time (FORM 1
// itself: all columns of the model).  Only FORM 2 carries the code of split
// work items.
template <class M, int METHOD_CODE, int FORM = 0>
struct StreamSolver {
  static constexpr bool WITH_BUF = FORM == 0;
  static constexpr bool SPLIT = FORM == 2;
  // FORM 3: FORM 1 compiled for forward-only solves (no sensitivity columns
  // at compile time: +10 ... 18 % on the forward-only bench step over the
  // general instantiation with zero columns; more resident warps at 64 or 80
  // registers do NOT help, 4 blocks per SM without a register cap is best)
  static constexpr bool FWD = FORM == 3;
  static constexpr bool R5 = METHOD_CODE == 5;
  // ROSL(9), ROSL(11): the linear-model methods (namespace rosl)
  static constexpr bool ROSL = METHOD_CODE == 9 || METHOD_CODE == 11;
  static constexpr int NS = ROSL ? METHOD_CODE : (R5 ? 8 : 6);   // stages
  static constexpr double GAMMA =
      ROSL ? rosl::Table<ROSL ? METHOD_CODE : 9>::GAMMA
           : (R5 ? rodas5::GAMMA : rodas4::GAMMA);
  // step-size exponent on err^2: -1/(2 (q+1)), q = order of the estimate
  static constexpr float EXPO =
      ROSL ? -0.5f / (METHOD_CODE - 1) : (R5 ? -0.1f : -0.125f);
  static constexpr int N = M::N;
  static constexpr int NC = M::NC;
  static constexpr int NCS = M::NCS;
  static constexpr int P = M::P;

  const SolveArgs& a;
  double* sm;
  const int stride;

  // hot state (registers)
  double u, t, h, t_stop;
  double J0[N * N];
  int cur, n_att;
  int flags;        // bits 0-7 status, 8-27 rejected steps, 28-29 edge kind,
                    // 30 all columns integrated (ROSL)
  bool is_ghost = false;   // see solve_uniform
  int col0_;        // split launches: the one integrated column of this
                    // work item (index into the integrated list), else -1
  CHI_HD int col() const { return SPLIT ? col0_ : -1; }
  int pend;         // record held in the buffer, -1: none
  bool invalid, after_reject;
  bool buffering;   // records go through the buffer (see events())

  CHI_HD StreamSolver(const SolveArgs& args, double* smem, int smem_stride)
      : a(args), sm(smem), stride(smem_stride) {}

  static constexpr int PI = FWD ? 0 : stream_integrated<M>();
// S is stored for the integrated columns only; `m` is the column's rank
// among them (the loops over the columns carry it along)
  static constexpr int OFF_G = (ROSL ? 1 : 2) * PI * N;
  static constexpr int OFF_ZY = OFF_G + P;
  static constexpr int OFF_COLD = OFF_ZY + (ROSL ? 0 : 7 * N);
  static constexpr int OFF_C = OFF_COLD + 8;
  static constexpr int OFF_Y = OFF_C + NCS;
  // record buffer (ROSL): y, then S of the integrated columns
  static constexpr bool BUFFERED = ROSL && WITH_BUF;
  static constexpr int OFF_BUF = OFF_Y + N;
  // error weights (ROSL): floats, two per slot
  static constexpr int OFF_WT = OFF_BUF + (BUFFERED ? (1 + PI) * N : 0);
  static_assert(OFF_WT + (ROSL ? ((1 + PI) * N + 1) / 2 : 0) ==
                    stream_slots<M, METHOD_CODE, WITH_BUF, FWD>(),
                "slot layout");
#define CHI_S(buf, m, n) sm[(((buf) * PI + (m)) * N + (n)) * stride]
#define CHI_G(k) sm[(OFF_G + (k)) * stride]
#define CHI_ZY(i, n) sm[(OFF_ZY + ((i) - 1) * N + (n)) * stride]
// Cold per-system scalars, kept out of the register file because they are
// only touched when an event is pending:
//   0 system index     1 next dose-rate edge     2 time of the next record
//   3 fused: log-likelihood; deferred: snapshot write position (two ints)
//   4 fused: running product of the error scales (its log is taken once)
//   5 ints: next record, last record     6 ints: segment, last segment
//   7 two floats: the first accepted step size after the last dose-rate
//     increase / decrease.  Dosing is usually periodic: the next edge of the
//     same kind restarts from that size instead of a guess (halves the
//     rejected steps, -2.6 % step attempts on the bench workload)
#define CHI_COLD(k) sm[(OFF_COLD + (k)) * stride]
// the model constants: needed by the state column and the event handlers,
// not by the sensitivity columns of a linear model -- kept here and loaded
// where used so that they do not occupy registers across the column loop
#define CHI_C(j) sm[(OFF_C + (j)) * stride]
// ... and the state itself
#define CHI_Y(n) sm[(OFF_Y + (n)) * stride]
#define CHI_BUF_Y(n) sm[(OFF_BUF + (n)) * stride]
#define CHI_BUF_S(m, n) sm[(OFF_BUF + N + (m) * N + (n)) * stride]

  CHI_HD static bool inert(const int pk) {
    return pk >= N && ((M::INERT_MASK >> (pk - N)) & 1u);
  }
  CHI_HD void load_constants(double (&c)[NCS]) const {
#pragma unroll
    for (int j = 0; j < NCS; ++j) c[j] = CHI_C(j);
  }
  CHI_HD void load_state(double (&y)[N]) const {
#pragma unroll
    for (int n = 0; n < N; ++n) y[n] = CHI_Y(n);
  }

  CHI_HD float slot_float(const int slot, const int half) const {
#ifdef __CUDA_ARCH__
    return reinterpret_cast<const float*>(&sm[slot * stride])[half];
#else
    float v[2];
    __builtin_memcpy(v, &sm[slot * stride], 8);
    return v[half];
#endif
  }
  CHI_HD void set_slot_float(const int slot, const int half, const float x) {
#ifdef __CUDA_ARCH__
    reinterpret_cast<float*>(&sm[slot * stride])[half] = x;
#else
    float v[2];
    __builtin_memcpy(v, &sm[slot * stride], 8);
    v[half] = x;
    __builtin_memcpy(&sm[slot * stride], v, 8);
#endif
  }

  CHI_HD int cold_int(const int k, const int half) const {
#ifdef __CUDA_ARCH__
    return reinterpret_cast<const int*>(&CHI_COLD(k))[half];
#else
    int v[2];
    __builtin_memcpy(v, &CHI_COLD(k), 8);
    return v[half];
#endif
  }
  CHI_HD void set_cold_int(const int k, const int half, const int value) {
#ifdef __CUDA_ARCH__
    reinterpret_cast<int*>(&CHI_COLD(k))[half] = value;
#else
    int v[2];
    __builtin_memcpy(v, &CHI_COLD(k), 8);
    v[half] = value;
    __builtin_memcpy(&CHI_COLD(k), v, 8);
#endif
  }

  CHI_HD float cold_float(const int k, const int half) const {
    const int bits = cold_int(k, half);
    float v;
    __builtin_memcpy(&v, &bits, 4);
    return v;
  }
  CHI_HD void set_cold_float(const int k, const int half, const float v) {
    int bits;
    __builtin_memcpy(&bits, &v, 4);
    set_cold_int(k, half, bits);
  }
  // kind of the dose edge whose first step is still to be accepted:
  // bits 28-29 of `flags` (1 rate increase, 2 rate decrease)
  CHI_HD int edge_kind() const { return (flags >> 28) & 3; }

  CHI_HD int n_cols() const { return (!FWD && a.with_sens) ? a.n_cols : 0; }
  CHI_HD int n_int() const {
    if (FWD) return 0;
    return a.with_sens ? (col() >= 0 ? (a.n_int > 0 ? 1 : 0) : a.n_int) : 0;
  }
  // m-th integrated column of this work item: mechanistic parameter / column
  CHI_HD int ip(const int m) const {
    return a.int_param[(col() >= 0 ? col() : 0) + m];
  }
  CHI_HD int ic(const int m) const {
    return a.int_col[(col() >= 0 ? col() : 0) + m];
  }
  // does this work item write gradient column k?  (split launches: the item
  // of the column; item 0 also owns the columns that are not integrated)
  CHI_HD bool owns(const int k) const {
    if (col() < 0) return true;
    if (a.n_int > 0 && k == ic(0)) return true;
    return col() == 0 && inert(a.col_param[k]);
  }
  // (the warp-synchronous forms only run log-likelihood launches with the
  // observation fused: known at compile time there)
  CHI_HD bool loglik() const { return FORM != 0 || a.mode == 1; }
  CHI_HD bool deferred() const {
    return FORM == 0 && a.mode == 1 && a.snap != nullptr;
  }
  CHI_HD bool want_dsig() const {
    return !FWD && a.mode == 1 && a.with_sens && a.grad != nullptr;
  }
  CHI_HD int status() const { return flags & 0xff; }
  CHI_HD int64_t system() const { return static_cast<int64_t>(CHI_COLD(0)); }
  CHI_HD int individual(const int64_t b) const {
    return a.ids ? a.ids[b] : static_cast<int>(b % a.n_ids);
  }

  CHI_HD void init(const int64_t item) {
    static_assert(METHOD_CODE == 1 || METHOD_CODE == 5 || ROSL,
                  "unknown method");
    int64_t b = item;
    col0_ = -1;
    if (SPLIT && a.split > 0) {
      b = item / a.split;
      col0_ = static_cast<int>(item - b * a.split);
    }
    static_assert(P <= MAX_COLS, "too many mechanistic parameters");
    const int id = individual(b);
    const double* xb = a.x + b * a.x_stride;
    const double* sg = xb + P;
    double y[N];
#pragma unroll
    for (int i = 0; i < N; ++i) y[i] = xb[i];
#pragma unroll
    for (int i = 0; i < N; ++i) CHI_Y(i) = y[i];
    double c[NCS];
#pragma unroll
    for (int i = 0; i < NC; ++i) c[i] = xb[N + i];
    if (NC == 0) c[0] = 0.0;
#pragma unroll
    for (int j = 0; j < NCS; ++j) CHI_C(j) = c[j];
    cur = 0;
    const int nc = n_cols();
    for (int k = 0; k < P; ++k) CHI_G(k) = 0.0;
    const int ni = n_int();
    for (int m = 0; m < ni; ++m) {
      const int pk = ip(m);
#pragma unroll
      for (int n = 0; n < N; ++n) {
        CHI_S(0, m, n) = (pk == n) ? 1.0 : 0.0;
        if (!ROSL) CHI_S(1, m, n) = 0.0;
      }
    }
    invalid = false;
    pend = -1;
    // Fused observation with the lanes of a warp on the same individual
    // (whole-warp refill): a lane that reaches a record parks (y, S) in its
    // buffer and goes on; the records are evaluated when every lane of the
    // warp holds one -- ONE convergent pass of the handler per record instead
    // of one per group of lanes that happen to arrive in the same iteration
    // (the handler ran with ~9 active lanes in 42 % of the iterations: 29 %
    // of the stall samples).
    buffering = BUFFERED && a.mode == 1 && a.snap == nullptr &&
                a.refill_wait == INT32_MAX;
    if (loglik()) {
      for (int o = 0; o < a.n_out; ++o) {
        const int off = a.err_off[o];
        if (sg[off] <= 0.0) invalid = true;
        if (a.err_kind[o] == ERR_CAM && sg[off + 1] <= 0.0) invalid = true;
      }
      if (want_dsig() && !deferred() && col() <= 0 && !is_ghost) {
        // the error-model gradients accumulate in the output row itself
        double* gb = a.grad + b * a.grad_stride + P;
        for (int k = 0; k < a.n_sigma; ++k) gb[k] = 0.0;
      }
    }
    const int64_t r = a.rec_off[id];
    const int64_t r_end = a.rec_off[id + 1];
    const int64_t seg = a.seg_off[id];
    const int64_t seg_end = a.seg_off[id + 1];
    u = (seg < seg_end) ? a.seg_rate[seg] : 0.0;
    CHI_COLD(0) = static_cast<double>(b);
    CHI_COLD(1) = (seg + 1 < seg_end) ? a.seg_t[seg + 1] : INFINITY;
    CHI_COLD(2) = (r < r_end) ? a.rec_t[r] : INFINITY;
    if (deferred()) {
      const int64_t pos = b * a.snap_rows * static_cast<int64_t>(a.snap_width);
      set_cold_int(3, 0, static_cast<int>(pos & 0xffffffff));
      set_cold_int(3, 1, static_cast<int>(pos >> 32));
    } else {
      CHI_COLD(3) = 0.0;
      CHI_COLD(4) = 1.0;
    }
    set_cold_int(5, 0, static_cast<int>(r));
    set_cold_int(5, 1, static_cast<int>(r_end));
    set_cold_int(6, 0, static_cast<int>(seg));
    set_cold_int(6, 1, static_cast<int>(seg_end));
    CHI_COLD(7) = 0.0;
    t = 0.0;
    t_stop = 0.0;   // forces the first pass through the event handler
    h = 0.0;
    after_reject = false;
    n_att = 0;
    flags = ST_OK;
    if (ROSL) {
      // bit 30: the integrated columns are ALL columns of the model, in
      // their natural order (step_rosl<1>)
      bool full = ni == PI;
#pragma unroll
      for (int m = 0; m < PI; ++m)
        if (full && ip(m) != full_param(m)) full = false;
      if (full) flags |= 1 << 30;
    }
    if (M::LINEAR) M::jac(y, c, J0);
    if (ROSL) {
      const float atol_f = static_cast<float>(a.atol);
      const float rtol_f = static_cast<float>(a.rtol);
#pragma unroll
      for (int n = 0; n < N; ++n) set_weight(n, y[n], atol_f, rtol_f);
#pragma unroll
      for (int m = 0; m < PI; ++m) {
        const int pk = m < ni ? ip(m) : -1;
#pragma unroll
        for (int n = 0; n < N; ++n)
          set_weight(N + m * N + n, (pk == n) ? 1.0 : 0.0, atol_f, rtol_f);
      }
    }
  }

  // ll -= log(scale), deferred: the scales are multiplied up and the
  // logarithm is taken when the product leaves [1e-270, 1e270] or at the end
  CHI_HD void add_log_scale(const double scale) {
    double lp = CHI_COLD(4) * scale;
    if (!(lp > 1e-270 && lp < 1e270)) {
      CHI_COLD(3) -= log(lp);
      lp = 1.0;
    }
    CHI_COLD(4) = lp;
  }

  // Deferred observation: append the state and the integrated sensitivity
  // columns to the system's snapshot rows; observe_system() does the rest.
  CHI_HD void snapshot() {
    const int64_t pos =
        (static_cast<int64_t>(cold_int(3, 1)) << 32) |
        static_cast<int64_t>(static_cast<unsigned>(cold_int(3, 0)));
    double* row = a.snap + pos;
    // (split launches: the items of a system share its rows -- item 0 stores
    // the state, every item its own column)
    if (col() <= 0) {
#pragma unroll
      for (int n = 0; n < N; ++n) row[n] = CHI_Y(n);
    }
    const int ni = n_int();
    const int m0 = col() > 0 ? col() : 0;
#pragma unroll
    for (int m = 0; m < PI; ++m) {
      if (m < ni) {
#pragma unroll
        for (int n = 0; n < N; ++n)
          row[N + (m0 + m) * N + n] = CHI_S(cur, m, n);
      }
    }
    const int64_t next = pos + a.snap_width;
    set_cold_int(3, 0, static_cast<int>(next & 0xffffffff));
    set_cold_int(3, 1, static_cast<int>(next >> 32));
  }

  // parks the state and the sensitivities at record r in the buffer
  CHI_HD void park(const int r) {
#pragma unroll
    for (int n = 0; n < N; ++n) CHI_BUF_Y(n) = CHI_Y(n);
    const int ni = n_int();
#pragma unroll
    for (int m = 0; m < PI; ++m) {
      if (m < ni) {
#pragma unroll
        for (int n = 0; n < N; ++n) CHI_BUF_S(m, n) = CHI_S(cur, m, n);
      }
    }
    pend = r;
  }

  // observation record r (fused handling): at the current time, or from the
  // buffer
  template <bool FROM_BUFFER = false>
  CHI_HD void observe(const int r) {
    const int o = a.rec_out[r];
    const int oid = a.out_id[o];
    const int nc = n_cols();
    double y[N], c[NCS], gy[N], gc[NCS];
    if (FROM_BUFFER) {
#pragma unroll
      for (int n = 0; n < N; ++n) y[n] = CHI_BUF_Y(n);
    } else {
      load_state(y);
    }
    load_constants(c);
    const double yb = M::out_grad(oid, y, c, gy, gc);
    if (loglik()) {
      const int64_t b = system();
      const int off = a.err_off[o];
      const ErrTerms e =
          error_terms(a.err_kind[o], a.x + b * a.x_stride + P + off, yb,
                      a.rec_obs[r], a.rec_logobs[r]);
      if (e.invalid) invalid = true;
      CHI_COLD(3) += e.ll;
      add_log_scale(e.scale);
      if (want_dsig() && col() <= 0) {
        double* dsig = a.grad + b * a.grad_stride + P + off;
        dsig[0] += e.ds0;
        if (a.err_kind[o] == ERR_CAM) dsig[1] += e.ds1;
      }
      // d ybar / d psi_k = (dg/dy) S_k + dg/dc_k
      const int ni = n_int();
      for (int m = 0; m < ni; ++m) {
        const int k = ic(m);
        double d = 0.0;
#pragma unroll
        for (int n = 0; n < N; ++n)
          d = fma(gy[n], FROM_BUFFER ? CHI_BUF_S(m, n) : CHI_S(cur, m, n), d);
        CHI_G(k) = fma(e.w, d, CHI_G(k));
      }
#pragma unroll
      for (int j = 0; j < NC; ++j) {
        const int k = a.param_col[N + j];
        if (gc[j] != 0.0 && k >= 0 && k < nc && owns(k))
          CHI_G(k) = fma(e.w, gc[j], CHI_G(k));
      }
    } else {
      const int64_t b = system();
      const int64_t row = a.traj_off[b] + (r - a.rec_off[individual(b)]);
      a.traj_y[row] = yb;
      for (int k = 0, m = 0; k < nc; ++k) {
        const int pk = a.col_param[k];
        double d = 0.0;
#pragma unroll
        for (int j = 0; j < NC; ++j) d = (pk == N + j) ? gc[j] : d;
        if (!inert(pk)) {
#pragma unroll
          for (int n = 0; n < N; ++n) d = fma(gy[n], CHI_S(cur, m, n), d);
          ++m;
        }
        a.traj_s[row * nc + k] = d;
      }
    }
  }

  // dose-rate edge: the input switches to u_new
  CHI_HD void switch_rate(const double u_new) {
    // restart from the first accepted step after the previous edge of this
    // kind, if there was one; else shrink the current proposal
    const int kind = (fabs(u_new) > fabs(u)) ? 1 : 2;
    u = u_new;
    const float hm = cold_float(7, kind - 1);
    h = (hm > 0.0f) ? fmin(static_cast<double>(hm), h)
                    : (CHI_EDGE_KEEP_H ? h * CHI_EDGE_SHRINK : 0.0);
    flags = (flags & ~(3 << 28)) | (kind << 28);
  }

  // Handles the events pending at the current time (observation records,
  // dose-rate edges).  Returns true when the system is finished.  The common
  // case -- the next stop has not been reached -- is one comparison.
  // Returns EV_GO (take a step), EV_DONE (system finished) or EV_WAIT (the
  // lane holds a record in its buffer and cannot go on before the warp has
  // evaluated it: it reached the next record, or the end).
  static constexpr int EV_GO = 0, EV_DONE = 1, EV_WAIT = 2;
  CHI_HD int events() {
    if (t < t_stop) return EV_GO;
    if (invalid) return EV_DONE;
    int r = cold_int(5, 0);
    const int r_end = cold_int(5, 1);
    double t_edge = CHI_COLD(1);
    double t_rec = CHI_COLD(2);
    for (;;) {
      if (r >= r_end) {
        if (pend < 0) return EV_DONE;
        break;
      }
      if (t_rec <= t) {
        if (buffering) {
          if (pend >= 0) break;        // the buffer is taken: wait here
          park(r);
        } else if (deferred()) {
          snapshot();
        } else {
          observe(r);
          if (invalid) return EV_DONE;
        }
        ++r;
        t_rec = (r < r_end) ? a.rec_t[r] : INFINITY;
        continue;
      }
      if (t_edge <= t) {
        // dose-rate edge: switch the input; the step size proposed before
        // the edge is kept (shrunk), the controller takes it from there
        const int seg = cold_int(6, 0) + 1;
        set_cold_int(6, 0, seg);
        switch_rate(a.seg_rate[seg]);
        t_edge = (seg + 1 < cold_int(6, 1)) ? a.seg_t[seg + 1] : INFINITY;
        continue;
      }
      set_cold_int(5, 0, r);
      CHI_COLD(1) = t_edge;
      CHI_COLD(2) = t_rec;
      t_stop = fmin(t_rec, t_edge);
      return EV_GO;
    }
    // waiting (t >= t_stop still holds: the next call comes back here)
    set_cold_int(5, 0, r);
    CHI_COLD(1) = t_edge;
    CHI_COLD(2) = t_rec;
    return EV_WAIT;
  }

  // evaluates the record held in the buffer (all lanes of a warp together)
  CHI_HD void observe_parked() {
    observe<true>(pend);
    pend = -1;
  }

  // All stages of sensitivity column k (constant index kc, -1 for an
  // initial-value column) given what the state column left behind; returns
  // the column's squared scaled error.
  //   nonlinear models: stage values Zy (shared memory) and increments Uy;
  //     the column-specific terms are selected by a uniform switch per stage.
  //   linear models (k-stage form, see the K constants): the augmented system
  //     is linear with the block-triangular Jacobian [[J, 0], [D_k, J]], so
  //         kappa_i = E arg_i + W^-1 (D_k V_i + e_k),   E = W^-1 J,
  //     with V_i = arg_i^y + kappa_i^y stored by the state column (V0 for the
  //     first stage) and D_k, e_k fetched once per column and multiplied
  //     through by W^-1 once per step: a stage is one combination and two
  //     N x N products, and neither the state column's increments nor the
  //     model constants are live here.
  template <bool HAS_TERMS>
  CHI_HD float column(const int kc, const int m, const int nxt,
                      const double (&y)[N], const double (&c)[NCS],
                      const double (&Uy)[NS][N],
                      const double (&V0)[N],
                      const double (&Wi)[N * N], const double (&E)[N * N],
                      const double hinv, const float atol_f,
                      const float rtol_f) {
    double S0[N], Us[NS][N], Zl[N];
#pragma unroll
    for (int n = 0; n < N; ++n) S0[n] = CHI_S(cur, m, n);
    // HAS_TERMS = false: initial-value column of a linear model, D_k = 0 and
    // e_k = 0 (no column-specific work at all)
    // COL_RANK1: every D_k has a single non-zero column q (the constant
    // multiplies one state), D_k V = d_k V[q]: W^-1 d_k once per step, one
    // shared-memory load and N FMAs per stage instead of N and N^2
    constexpr bool R1 = M::LINEAR && M::COL_RANK1;
    double WD[R1 ? N : N * N], We[N], Vp[N];
    int q1 = 0;
    double v0 = V0[0];
    if (M::LINEAR && HAS_TERMS) {
      if constexpr (R1) {
        double dk[N], ek[N];
        q1 = M::col_rank1(kc, c, u, dk, ek);
        matvec<N>(Wi, ek, We);
        matvec<N>(Wi, dk, WD);
#pragma unroll
        for (int j = 1; j < N; ++j) v0 = (q1 == j) ? V0[j] : v0;
      } else {
        double DK[N * N], ek[N];
        M::col_linear(kc, c, u, DK, ek);
        matvec<N>(Wi, ek, We);
#pragma unroll
        for (int i = 0; i < N; ++i)
#pragma unroll
          for (int j = 0; j < N; ++j) {
            double acc = Wi[i * N] * DK[j];
#pragma unroll
            for (int q = 1; q < N; ++q)
              acc = fma(Wi[i * N + q], DK[q * N + j], acc);
            WD[i * N + j] = acc;
          }
      }
    }
#define CHI_KSTAGE(I, FIRST, CNT, SKIP, k0, k1, k2, k3, k4, k5, k6)           \
  {                                                                           \
    lincombk<FIRST, CNT, SKIP, N>(Zl, FIRST == 0 ? S0 : Zl, Us, k0, k1, k2,   \
                                  k3, k4, k5, k6);                            \
    double vq = 0.0;                                                          \
    if (HAS_TERMS && R1) vq = (I == 0) ? v0 : CHI_ZY(I, q1);                  \
    _Pragma("unroll") for (int n = 0; n < N; ++n) {                           \
      double acc;                                                             \
      if (HAS_TERMS) {                                                        \
        if (R1) {                                                             \
          acc = fma(WD[n], vq, We[n]);                                        \
        } else {                                                              \
          acc = We[n];                                                        \
          _Pragma("unroll") for (int q = 0; q < N; ++q) acc = fma(            \
              WD[(R1 ? 0 : n * N) + q], (I == 0) ? V0[q] : CHI_ZY(I, q),      \
              acc);                                                           \
        }                                                                     \
        _Pragma("unroll") for (int q = 0; q < N; ++q) acc =                   \
            fma(E[n * N + q], Zl[q], acc);                                    \
      } else {                                                                \
        acc = E[n * N] * Zl[0];                                               \
        _Pragma("unroll") for (int q = 1; q < N; ++q) acc =                   \
            fma(E[n * N + q], Zl[q], acc);                                    \
      }                                                                       \
      Us[I][n] = acc;                                                         \
    }                                                                         \
    if (I == NS - 2) {                                                        \
      _Pragma("unroll") for (int n = 0; n < N; ++n) Vp[n] =                   \
          Zl[n] + Us[I][n];                                                   \
    }                                                                         \
  }
#define CHI_SSTAGE(I, CNT, INCR, a0, a1, a2, a3, a4, a5, a6, c0, c1, c2, c3,  \
                   c4, c5, c6)                                                \
  {                                                                           \
    double cs[N], rs[N];                                                      \
    double Ji[N * N], Zyi[N];                                                 \
    _Pragma("unroll") for (int n = 0; n < N; ++n) Zyi[n] =                    \
        (I == 0) ? y[n] : CHI_ZY(I, n);                                       \
    M::jac(Zyi, c, Ji);                                                       \
    if (INCR) {                                                               \
      _Pragma("unroll") for (int n = 0; n < N; ++n) Zl[n] +=                  \
          Us[I > 0 ? I - 1 : 0][n];                                           \
    } else {                                                                  \
      lincomb7<CNT, N>(Zl, S0, Us, a0, a1, a2, a3, a4, a5, a6);               \
    }                                                                         \
    lincomb7<CNT, N>(cs, nullptr, Us, c0, c1, c2, c3, c4, c5, c6);            \
    matvec<N>(Ji, Zl, rs);                                                    \
    M::dfdc_col_add(kc, Zyi, c, rs);                                          \
    M::hess_col_add(kc, y, c, S0, Uy[I], rs);                                 \
    _Pragma("unroll") for (int n = 0; n < N; ++n) rs[n] =                     \
        fma(hinv, cs[n], rs[n]);                                              \
    matvec<N>(Wi, rs, Us[I]);                                                 \
  }
    if constexpr (M::LINEAR) {
      if constexpr (R5) {
        CHI_RODAS5_K_TABLE(CHI_KSTAGE)
      } else {
        CHI_RODAS4_K_TABLE(CHI_KSTAGE)
      }
    } else {
      if constexpr (R5) {
        CHI_RODAS5_TABLE(CHI_SSTAGE)
      } else {
        CHI_RODAS4_TABLE(CHI_SSTAGE)
      }
    }
#undef CHI_KSTAGE
#undef CHI_SSTAGE
    // stiffly accurate: new value = last stage value + last increment; the
    // error estimate is the last increment (transformed form) or the
    // difference to the value after the stage before (k-stage form)
    float s2 = 0.0f;
#pragma unroll
    for (int n = 0; n < N; ++n) {
      const double Sn = Zl[n] + Us[NS - 1][n];
      const double en = M::LINEAR ? Sn - Vp[n] : Us[NS - 1][n];
      s2 += err_term(en, S0[n], Sn, atol_f, rtol_f);
      CHI_S(nxt, m, n) = Sn;
    }
    return s2 * (1.0f / N);
  }

  // One step attempt towards the next stop.  Returns true on failure.
  CHI_HD bool step() {
    if constexpr (ROSL) {
      if constexpr (M::LINEAR && FWD) {
        return step_rosl<0>();
      } else if constexpr (M::LINEAR && FORM == 1) {
        // all columns of the model, natural order (the launcher checks)
        return step_rosl<1>();
      } else if constexpr (M::LINEAR && FORM == 4) {
        return step_rosl<2>();
      } else if constexpr (M::LINEAR) {
        if (n_int() == 0) return step_rosl<0>();
        if constexpr (SPLIT) {
          if (col() >= 0) return step_rosl<3>();
        }
#ifdef CHI_ROSL_NO_SUBSET
        return step_rosl<1>();   // (register-pressure experiment only)
#else
        if (flags & (1 << 30)) return step_rosl<1>();
        return step_rosl<2>();
#endif
      } else {
        flags |= ST_NON_FINITE;   // ROSL is for linear models only
        return true;
      }
    } else {
      return step_rodas();
    }
  }

  // parameter index of the m-th integrated column when ALL columns of the
  // model are asked for (the usual case): known at compile time
  static constexpr int full_param(const int m) {
    int seen = 0;
    for (int pk = 0; pk < P; ++pk) {
      const bool inert_pk =
          pk >= N && ((M::INERT_MASK >> (pk - N)) & 1u);
      if (inert_pk) continue;
      if (seen == m) return pk;
      ++seen;
    }
    return 0;
  }

  // What step_rodas and step_rosl do with the error norm of an attempt:
  // accept or reject, next step size, failure bookkeeping.  Returns true on
  // failure.
  CHI_HD bool control(const float errsq, const double hs, const bool clipped,
                      bool& accept) {
    accept = errsq <= 1.0f;
    float fac;
    if (!(errsq == errsq) || isinf(errsq)) {
      fac = 0.2f;
    } else {
      fac = CHI_SAFETY * fast_powf(fmaxf(errsq, 1e-30f), EXPO);
      fac = fminf(fmaxf(fac, 0.2f), CHI_FACMAX);
    }
    if (accept) {
      t = clipped ? t_stop : t + hs;
      if (edge_kind() != 0) {
        set_cold_float(7, edge_kind() - 1, static_cast<float>(hs));
        flags &= ~(3 << 28);
      }
      if (after_reject) fac = fminf(fac, 1.0f);
      const double hn = hs * static_cast<double>(fac);
      h = (hs < h) ? fmax(hn, h) : hn;
      after_reject = false;
    } else {
      h = hs * static_cast<double>(fminf(fac, 0.9f));
      after_reject = true;
      flags += 0x100;
      if (h <= 1e-14 * fmax(fabs(t), 1.0)) {
        flags |= (errsq == errsq) ? ST_STEP_UNDERFLOW : ST_NON_FINITE;
        return true;
      }
    }
    if (++n_att >= a.max_steps) {
      flags |= ST_MAX_STEPS;
      return true;
    }
    return false;
  }

  // One ROSL(s) step attempt of a linear model: the state column and all
  // integrated sensitivity columns advance TOGETHER, stage by stage, in
  // registers -- (1 + columns) N independent FMA chains per thread, nothing
  // of the earlier stages kept, no stage values in shared memory.  With
  // g = gamma h, T = (I - g J)^-1 (ONE reciprocal per attempt), E = T g J:
  //   w_0^y = T h f(y),                       w_k^y = E w_{k-1}^y
  //   w_0^S = (E S + G_k y) / gamma + h T e_k + G_k w_0^y,   G_k = g T D_k
  //   w_k^S = E w_{k-1}^S + G_k (w_{k-1}^y + w_k^y)
  // (the augmented Jacobian [[J, 0], [D_k, J]] is block triangular, so
  // T_aug - I = [[E, 0], [G_k T, E]], T w = w + E w and T h J S = E S / gamma).
  // The terms of the last two stages are the error estimate: stage S-2 only
  // advances w, the last stage forms e = c_{S-2} w_{S-2} + c_{S-1} w_{S-1}
  // component by component, adds it to the solution and to the error norm --
  // no second set of accumulators is ever live.  The norm is weighted by
  // 1 / (atol + rtol |z_n|) at the START of the step (as CVODES weighs its
  // error test), kept as floats in shared memory and refreshed when a step is
  // accepted: an attempt converts the twelve error terms, not three values per
  // component.
  // MODE 0: no columns (forward only); 1: all columns of the model (their
  // parameters are compile-time constants: the column terms fold); 2: a
  // subset chosen at run time (fix_parameters), unused slots carry zeros;
  // 3: the one column of a split work item (SolveArgs::split).
  template <int MODE>
  CHI_HD bool step_rosl() {
    constexpr int S = NS;
    constexpr int NCOL = MODE == 0 ? 0 : (MODE == 3 ? 1 : PI);
    constexpr int NCA = NCOL > 0 ? NCOL : 1;
    constexpr bool R1 = M::COL_RANK1;
    constexpr int WDN = R1 ? N : N * N;
    constexpr double RG = 1.0 / GAMMA;
    double y[N], c[NCS];
    load_state(y);
    load_constants(c);
    if (h == 0.0) {
      h = CHI_H0_FRACTION * (t_stop - t);
      after_reject = false;
    }
    const double rem = t_stop - t;
    double hs = h;
    bool clipped = false;
    if (rem <= CHI_CLIP_STRETCH * h) {
      hs = rem;
      clipped = true;
    } else if (rem < 2.0 * h) {
      hs = 0.5 * rem;
    }
    const double g = GAMMA * hs;
    double T[N * N], E[N * N];
    {
      double GJ[N * N], A[N * N];
#pragma unroll
      for (int i = 0; i < N; ++i)
#pragma unroll
        for (int j = 0; j < N; ++j) {
          GJ[i * N + j] = g * J0[i * N + j];
          A[i * N + j] = (i == j ? 1.0 : 0.0) - GJ[i * N + j];
        }
      invert<N>(A, T);
      // E = T g J = T - I.  The subtraction leaves an ABSOLUTE error of one
      // ulp of 1 in the diagonal of E; E only ever multiplies increments w
      // that are added to z with |w| <~ |z|, so that is an error of one ulp
      // of z -- and N^3 - N operations less per attempt than the product.
#pragma unroll
      for (int i = 0; i < N; ++i)
#pragma unroll
        for (int j = 0; j < N; ++j)
          E[i * N + j] = (i == j) ? T[i * N + j] - 1.0 : T[i * N + j];
    }
    // ---- stage 0 ----------------------------------------------------------
    double ay[N], zy[N];
    {
      double f0[N];
      M::rhs(y, c, u, f0);
#pragma unroll
      for (int n = 0; n < N; ++n) f0[n] *= hs;
#pragma unroll
      for (int n = 0; n < N; ++n) {
        double acc = f0[n];
#pragma unroll
        for (int q = 0; q < N; ++q) acc = fma(E[n * N + q], f0[q], acc);
        ay[n] = acc;
        zy[n] = fma(rosl_c<S>(0), acc, y[n]);
      }
    }
    double w[NCA][N], zs[NCA][N], WD[NCA][WDN];
    int qs[NCA];
    bool terms[NCA];
    const int ni = MODE == 2 ? n_int() : NCOL;
#pragma unroll
    for (int m = 0; m < NCOL; ++m) {
      const bool live = MODE == 1 || m < ni;
      const int pk = MODE == 1 ? full_param(m) : (live ? ip(m) : 0);
      const int kc = pk >= N ? pk - N : -1;
      // (MODE 3: the lanes of a warp hold different columns of one system;
      // an initial-value column runs the general column body with G = 0
      // instead of a branch of its own -- no divergence inside the stages)
      terms[m] = live && (MODE == 3 || kc >= 0);
      qs[m] = 0;
      double S0[N], v[N], add[N];
#pragma unroll
      for (int n = 0; n < N; ++n) S0[n] = live ? CHI_S(cur, m, n) : 0.0;
      // v = E S + G_k y (times 1 / gamma below), add = G_k w_0^y + h T e_k
      matvec<N>(E, S0, v);
#pragma unroll
      for (int j = 0; j < WDN; ++j) WD[m][j] = 0.0;
#pragma unroll
      for (int n = 0; n < N; ++n) add[n] = 0.0;
      if (terms[m]) {
        double ek[N];
        if constexpr (R1) {
          double dk[N];
          const int q1 = M::col_rank1(kc, c, u, dk, ek);
          qs[m] = q1;
          double yq = y[0], aq = ay[0];
#pragma unroll
          for (int j = 1; j < N; ++j) {
            yq = (q1 == j) ? y[j] : yq;
            aq = (q1 == j) ? ay[j] : aq;
          }
          // G_k = g T d_k (entries of d_k that are zero at compile time drop
          // out)
#pragma unroll
          for (int n = 0; n < N; ++n) {
            double acc = 0.0;
            bool first = true;
#pragma unroll
            for (int q = 0; q < N; ++q)
              if (dk[q] != 0.0) {
                acc = first ? T[n * N + q] * dk[q]
                            : fma(T[n * N + q], dk[q], acc);
                first = false;
              }
            WD[m][n] = g * acc;
          }
#pragma unroll
          for (int n = 0; n < N; ++n) {
            v[n] = fma(WD[m][n], yq, v[n]);
            add[n] = WD[m][n] * aq;
          }
        } else {
          double DK[N * N];
          M::col_linear(kc, c, u, DK, ek);
#pragma unroll
          for (int i = 0; i < N; ++i)
#pragma unroll
            for (int j = 0; j < N; ++j) {
              double acc = 0.0;
              bool first = true;
#pragma unroll
              for (int q = 0; q < N; ++q)
                if (DK[q * N + j] != 0.0) {
                  acc = first ? T[i * N + q] * DK[q * N + j]
                              : fma(T[i * N + q], DK[q * N + j], acc);
                  first = false;
                }
              WD[m][i * N + j] = g * acc;
            }
#pragma unroll
          for (int n = 0; n < N; ++n) {
#pragma unroll
            for (int q = 0; q < N; ++q) {
              v[n] = fma(WD[m][n * N + q], y[q], v[n]);
              add[n] = fma(WD[m][n * N + q], ay[q], add[n]);
            }
          }
        }
        // h T e_k (a constant that enters the rhs additively; zero for the
        // rate constants of the PK models)
#pragma unroll
        for (int n = 0; n < N; ++n)
#pragma unroll
          for (int q = 0; q < N; ++q)
            if (ek[q] != 0.0) add[n] = fma(T[n * N + q], hs * ek[q], add[n]);
      }
#pragma unroll
      for (int n = 0; n < N; ++n) {
        const double acc = fma(RG, v[n], add[n]);
        w[m][n] = acc;
        zs[m][n] = fma(rosl_c<S>(0), acc, S0[n]);
      }
    }
    // ---- stages 1 .. S-2 ----------------------------------------------------
    // One stage of all columns: w <- next power; ACC: its term is added to
    // the solution (stages 1 .. S-3: a rolled loop, one stage body in the
    // instruction cache -- fully unrolled the step is ~2000 instructions and
    // the four warps of a scheduler miss the cache, `no_instruction` 1.35
    // stall cycles per issue).
    auto stage = [&](const double ck, auto acc_tag) {
      constexpr bool ACC = decltype(acc_tag)::value;
      double an[N], sig[N];
#pragma unroll
      for (int n = 0; n < N; ++n) {
        double acc = E[n * N] * ay[0];
#pragma unroll
        for (int q = 1; q < N; ++q) acc = fma(E[n * N + q], ay[q], acc);
        an[n] = acc;
      }
#pragma unroll
      for (int n = 0; n < N; ++n) {
        sig[n] = ay[n] + an[n];
        ay[n] = an[n];
        if (ACC) zy[n] = fma(ck, an[n], zy[n]);
      }
#pragma unroll
      for (int m = 0; m < NCOL; ++m) {
        double wn[N];
        next_power<R1>(terms[m], qs[m], WD[m], E, sig, w[m], wn);
#pragma unroll
        for (int n = 0; n < N; ++n) {
          w[m][n] = wn[n];
          if (ACC) zs[m][n] = fma(ck, wn[n], zs[m][n]);
        }
      }
    };
    // (the warp-synchronous kernels unroll the loop completely: +2.4 % on
    // the all-columns batched kernel once it holds nothing but this one step
    // body, -4 % solve time for single vectors; the general kernel, which
    // carries every step body, keeps four stages per trip)
    constexpr int UNROLL = FORM != 0 ? 8 : CHI_ROSL_UNROLL;
#ifdef __CUDA_ARCH__
#pragma unroll UNROLL
#endif
    for (int k = 1; k < S - 2; ++k)
      stage(rosl_c<S>(k), std::integral_constant<bool, true>());
    stage(0.0, std::integral_constant<bool, false>());
    // ---- last stage: error estimate, new solution, error norm ---------------
    // (max over the columns of the weighted RMS)
    const double ca = rosl_c<S>(S - 2), cb = rosl_c<S>(S - 1);
    float errsq;
    {
      double an[N], sig[N];
#pragma unroll
      for (int n = 0; n < N; ++n) {
        double acc = E[n * N] * ay[0];
#pragma unroll
        for (int q = 1; q < N; ++q) acc = fma(E[n * N + q], ay[q], acc);
        an[n] = acc;
      }
      float s2 = 0.0f;
#pragma unroll
      for (int n = 0; n < N; ++n) {
        sig[n] = ay[n] + an[n];
        const double e = fma(cb, an[n], ca * ay[n]);
        zy[n] += e;
        const float q = static_cast<float>(e) * weight(n);
        s2 = fmaf(q, q, s2);
      }
      errsq = s2 * (1.0f / N);
#pragma unroll
      for (int m = 0; m < NCOL; ++m) {
        double wn[N];
        next_power<R1>(terms[m], qs[m], WD[m], E, sig, w[m], wn);
        float s2m = 0.0f;
#pragma unroll
        for (int n = 0; n < N; ++n) {
          const double e = fma(cb, wn[n], ca * w[m][n]);
          zs[m][n] += e;
          const float q = static_cast<float>(e) * weight(N + m * N + n);
          s2m = fmaf(q, q, s2m);
        }
        errsq = fmaxf(errsq, s2m * (1.0f / N));
      }
    }
    bool accept;
    const bool failed = control(errsq, hs, clipped, accept);
    if (accept) {
      const float atol_f = static_cast<float>(a.atol);
      const float rtol_f = static_cast<float>(a.rtol);
#pragma unroll
      for (int n = 0; n < N; ++n) {
        CHI_Y(n) = zy[n];
        set_weight(n, zy[n], atol_f, rtol_f);
      }
#pragma unroll
      for (int m = 0; m < NCOL; ++m) {
        if (MODE == 1 || m < ni) {
#pragma unroll
          for (int n = 0; n < N; ++n) {
            CHI_S(cur, m, n) = zs[m][n];
            set_weight(N + m * N + n, zs[m][n], atol_f, rtol_f);
          }
        }
      }
    }
    return failed;
  }

  // w_k^S of one column from w_{k-1}^S (see step_rosl): E w + G (sig[q])
  template <bool R1>
  CHI_HD static void next_power(const bool has_terms, const int q1,
                                const double* G, const double (&E)[N * N],
                                const double (&sig)[N], const double (&w)[N],
                                double (&wn)[N]) {
    if (has_terms) {
      if constexpr (R1) {
        double sq = sig[0];
#pragma unroll
        for (int j = 1; j < N; ++j) sq = (q1 == j) ? sig[j] : sq;
#pragma unroll
        for (int n = 0; n < N; ++n) wn[n] = G[n] * sq;
      } else {
#pragma unroll
        for (int n = 0; n < N; ++n) {
          double acc = G[n * N] * sig[0];
#pragma unroll
          for (int q = 1; q < N; ++q) acc = fma(G[n * N + q], sig[q], acc);
          wn[n] = acc;
        }
      }
#pragma unroll
      for (int n = 0; n < N; ++n)
#pragma unroll
        for (int q = 0; q < N; ++q) wn[n] = fma(E[n * N + q], w[q], wn[n]);
    } else {
#pragma unroll
      for (int n = 0; n < N; ++n) {
        double acc = E[n * N] * w[0];
#pragma unroll
        for (int q = 1; q < N; ++q) acc = fma(E[n * N + q], w[q], acc);
        wn[n] = acc;
      }
    }
  }

  // error weights of the ROSL step: 1 / (atol + rtol |z|) of component j
  // (0 .. N-1: the state, then the integrated columns), two floats per slot
  CHI_HD float weight(const int j) const {
    return slot_float(OFF_WT + j / 2, j & 1);
  }
  CHI_HD void set_weight(const int j, const double z, const float atol_f,
                         const float rtol_f) {
    const float sc = fmaf(rtol_f, fabsf(static_cast<float>(z)), atol_f);
#ifdef __CUDA_ARCH__
    // (approximate reciprocal: one MUFU, no slow path -- the twelve weights
    // of a step are computed side by side)
    float wt;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(wt) : "f"(fmaxf(sc, 1e-30f)));
#else
    const float wt = 1.0f / fmaxf(sc, 1e-30f);
#endif
    set_slot_float(OFF_WT + j / 2, j & 1, wt);
  }

  CHI_HD bool step_rodas() {
    double y[N], c[NCS];
    load_state(y);
    load_constants(c);
    if (h == 0.0) {
      // first step of a system: a small fraction of the distance to the
      // first stop; the controller takes it from there (at most x5 per
      // step).  Measured against Hairer's d0/d1 heuristic on the bench
      // workload: 1.7 % fewer step attempts, identical errors, and the
      // division-heavy start-up code (executed by one lane at a time
      // whenever a lane is refilled) is gone.
      h = CHI_H0_FRACTION * (t_stop - t);
      after_reject = false;
    }

    const double rem = t_stop - t;
    double hs = h;
    bool clipped = false;
    if (rem <= CHI_CLIP_STRETCH * h) {
      hs = rem;
      clipped = true;
    } else if (rem < 2.0 * h) {
      hs = 0.5 * rem;
    }

    // ---- state column: stage values and increments kept -------------------
    if (!M::LINEAR) M::jac(y, c, J0);
    const double hinv = rcp_newton(hs);
    double Wi[N * N];
    {
      double W[N * N];
#pragma unroll
      for (int i = 0; i < N; ++i)
#pragma unroll
        for (int j = 0; j < N; ++j)
          W[i * N + j] = (i == j ? hinv * (1.0 / GAMMA) : 0.0) - J0[i * N + j];
      invert<N>(W, Wi);
    }
    // E = W^-1 J (= (I - gamma h J)^-1 - I, without the cancellation), used
    // by the linear-model bodies
    double E[N * N];
#pragma unroll
    for (int i = 0; i < N; ++i)
#pragma unroll
      for (int j = 0; j < N; ++j) {
        double acc = Wi[i * N] * J0[j];
#pragma unroll
        for (int q = 1; q < N; ++q)
          acc = fma(Wi[i * N + q], J0[q * N + j], acc);
        E[i * N + j] = acc;
      }
    double Uy[NS][N];
    double Zlast[N], V0[N], Vp[N];
#define CHI_YSTAGE(I, CNT, INCR, a0, a1, a2, a3, a4, a5, a6, c0, c1, c2, c3,  \
                   c4, c5, c6)                                                \
  {                                                                           \
    double f[N], cy[N], ry[N];                                                \
    if (INCR) {                                                               \
      _Pragma("unroll") for (int n = 0; n < N; ++n) Zlast[n] +=               \
          Uy[I > 0 ? I - 1 : 0][n];                                           \
    } else {                                                                  \
      lincomb7<CNT, N>(Zlast, y, Uy, a0, a1, a2, a3, a4, a5, a6);             \
    }                                                                         \
    M::rhs(Zlast, c, u, f);                                                   \
    lincomb7<CNT, N>(cy, nullptr, Uy, c0, c1, c2, c3, c4, c5, c6);            \
    _Pragma("unroll") for (int n = 0; n < N; ++n) ry[n] =                     \
        fma(hinv, cy[n], f[n]);                                               \
    matvec<N>(Wi, ry, Uy[I]);                                                 \
    /* what the sensitivity columns need of this stage */                     \
    _Pragma("unroll") for (int n = 0; n < N; ++n) {                           \
      const double v = M::LINEAR ? Zlast[n] + Uy[I][n] : Zlast[n];            \
      if (I == 0) V0[n] = v; else CHI_ZY(I, n) = v;                           \
    }                                                                         \
  }
    // linear models, k-stage form (see the K constants): f(Z) = J Z + f(0),
    //     kappa_i = E arg_i + W^-1 f(0),   V_i = arg_i + kappa_i,
    // and V_i is exactly what the sensitivity columns need of this stage; the
    // last one is the candidate solution, the one before the embedded one
#define CHI_YKSTAGE(I, FIRST, CNT, SKIP, k0, k1, k2, k3, k4, k5, k6)          \
  {                                                                           \
    lincombk<FIRST, CNT, SKIP, N>(Zlast, FIRST == 0 ? y : Zlast, Uy, k0, k1,  \
                                  k2, k3, k4, k5, k6);                        \
    _Pragma("unroll") for (int n = 0; n < N; ++n) {                           \
      double acc = Wf0[n];                                                    \
      _Pragma("unroll") for (int q = 0; q < N; ++q) acc =                     \
          fma(E[n * N + q], Zlast[q], acc);                                   \
      Uy[I][n] = acc;                                                         \
    }                                                                         \
    _Pragma("unroll") for (int n = 0; n < N; ++n) {                           \
      const double V = Zlast[n] + Uy[I][n];                                   \
      if (I == 0) V0[n] = V;                                                  \
      else if (need_v || I == NS - 1) CHI_ZY(I, n) = V;                       \
      if (I == NS - 2) Vp[n] = V;                                             \
    }                                                                         \
  }
    if constexpr (M::LINEAR) {
      double f0[N], zero[N];
#pragma unroll
      for (int n = 0; n < N; ++n) zero[n] = 0.0;
      M::rhs(zero, c, u, f0);
      // the columns read V_i from shared memory; a forward-only solve only
      // needs the last one (the new state)
      const bool need_v = n_int() > 0;
      double Wf0[N];
      matvec<N>(Wi, f0, Wf0);
      if constexpr (R5) {
        CHI_RODAS5_K_TABLE(CHI_YKSTAGE)
      } else {
        CHI_RODAS4_K_TABLE(CHI_YKSTAGE)
      }
    } else {
      if constexpr (R5) {
        CHI_RODAS5_TABLE(CHI_YSTAGE)
      } else {
        CHI_RODAS4_TABLE(CHI_YSTAGE)
      }
    }
#undef CHI_YKSTAGE
#undef CHI_YSTAGE
    float errsq;
    const float atol_f = static_cast<float>(a.atol);
    const float rtol_f = static_cast<float>(a.rtol);
    {
      // the candidate solution yn = last stage value + last increment; for
      // linear models that is what the last stage left in CHI_ZY(NS - 1, .),
      // it is fetched from there after the column loop
      float s2 = 0.0f;
#pragma unroll
      for (int n = 0; n < N; ++n) {
        const double yn = Zlast[n] + Uy[NS - 1][n];
        const double en = M::LINEAR ? yn - Vp[n] : Uy[NS - 1][n];
        if (!M::LINEAR) Zlast[n] = yn;
        s2 += err_term(en, y[n], yn, atol_f, rtol_f);
      }
      errsq = s2 * (1.0f / N);
    }

    // ---- sensitivity columns, streamed one at a time ----------------------
    // The column's constant index kc is uniform over the grid: the switches
    // on it inside the generated column terms are uniform branches.
    const int nxt = 1 - cur;
    // (columns of constants that do not enter the rhs stay identically zero
    // and are not in the list)
    const int ni = n_int();
    int m = 0;
    if (M::LINEAR) {
      // initial-value columns first (they lead the list): no column terms
      for (; m < ni && ip(m) < N; ++m)
        errsq = fmaxf(errsq,
                      column<false>(-1, m, nxt, y, c, Uy, V0, Wi, E, hinv,
                                    atol_f, rtol_f));
    }
    for (; m < ni; ++m) {
      const int pk = ip(m);
      const int kc = pk >= N ? pk - N : -1;
      errsq = fmaxf(errsq,
                    column<true>(kc, m, nxt, y, c, Uy, V0, Wi, E, hinv,
                                 atol_f, rtol_f));
    }

    bool accept;
    const bool failed = control(errsq, hs, clipped, accept);
    if (accept) {
#pragma unroll
      for (int n = 0; n < N; ++n)
        CHI_Y(n) = M::LINEAR ? CHI_ZY(NS - 1, n) : Zlast[n];
      cur = nxt;
    }
    return failed;
  }

  // Whole solve of one system by a warp whose lanes ALL integrate the same
  // individual under different parameter vectors (solve_warp_kernel): the
  // stops -- observation records and dose-rate edges -- are the same for
  // every lane, so the warp walks through them together.  Between two stops
  // the lanes step until the last one has arrived (a lane that is there
  // early idles: +2.7 % iterations at a parameter spread of 10 %, +8 % at
  // 30 %, tools/exp/sync_cost.py), then every lane handles the stop in ONE
  // convergent pass.  The loop around the step is a vote and a branch; the
  // per-lane event cursor, the record buffer and their predicates are gone
  // (they were 15 % of the issued instructions and 37 % of the stall samples
  // of solve_stream_kernel on the bench workload).  Per lane the arithmetic
  // and its order are those of solve_system_stream: bit-identical results.
  // `ghost`: a lane without a work item of its own (warp-per-system
  // launches, SolveArgs::split): it shadows a real item's control flow, takes
  // no steps and writes nothing.
  CHI_HD void solve_uniform(const int64_t item, const bool ghost = false) {
    is_ghost = ghost;
    init(item);
    const int64_t b = system();
    const int id = individual(b);
    int r = static_cast<int>(a.rec_off[id]);
    const int r_end = static_cast<int>(a.rec_off[id + 1]);
    int seg = static_cast<int>(a.seg_off[id]);
    const int seg_end = static_cast<int>(a.seg_off[id + 1]);
    bool go = !invalid && !ghost;
    double ts = 0.0;   // the current stop: every live lane has t == ts
    for (;;) {
      while (r < r_end && a.rec_t[r] <= ts) {
        if (go) {
          observe<false>(r);
          if (invalid) go = false;
        }
        ++r;
      }
      if (r >= r_end) break;
      while (seg + 1 < seg_end && a.seg_t[seg + 1] <= ts) {
        ++seg;
        if (go) switch_rate(a.seg_rate[seg]);
      }
      ts = fmin(a.rec_t[r],
                (seg + 1 < seg_end) ? a.seg_t[seg + 1] : INFINITY);
      t_stop = ts;
      bool stepping = go;
      while (CHI_WARP_ANY(stepping)) {
        if (stepping) {
          if (step()) {
            go = false;
            stepping = false;
          } else {
            stepping = t < t_stop;
          }
        }
      }
      if (!CHI_WARP_ANY(go)) break;
    }
    if (!ghost) finish();
  }

  CHI_HD void finish() {
    const int64_t b = system();
    const int n_rej = (flags >> 8) & 0xfffff;
    if (col() >= 0) {
      // split launch: the status of a system is the maximum over its items
      // (zeroed by the launcher)
#ifdef __CUDA_ARCH__
      if (status() != ST_OK) atomicMax(a.status + b, status());
#else
      if (status() > a.status[b]) a.status[b] = status();
#endif
    } else {
      a.status[b] = status();
    }
    if (col() <= 0) {
      if (a.n_steps) a.n_steps[b] = n_att - n_rej;
      if (a.n_rej) a.n_rej[b] = n_rej;
    }
    if (deferred()) return;   // observe_system() writes ll and the gradient
    if (loglik()) {
      const bool fail = invalid || status() != ST_OK;
      if (col() <= 0)
        a.ll[b] = fail ? -INFINITY : CHI_COLD(3) - log(CHI_COLD(4));
      if (a.with_sens && a.grad) {
        const int nc = n_cols();
        double* gb = a.grad + b * a.grad_stride;
        for (int k = 0; k < nc; ++k)
          if (owns(k)) gb[a.col_param[k]] = fail ? INFINITY : CHI_G(k);
        if (fail && col() <= 0)
          for (int k = 0; k < a.n_sigma; ++k) gb[P + k] = INFINITY;
      }
    } else if (status() != ST_OK) {
      const int r_end = cold_int(5, 1);
      const int64_t shift = a.traj_off[b] - a.rec_off[individual(b)];
      for (int64_t q = cold_int(5, 0); q < r_end; ++q)
        a.traj_y[shift + q] = NAN;
    }
  }
#undef CHI_S
#undef CHI_G
#undef CHI_ZY
#undef CHI_COLD
#undef CHI_C
#undef CHI_Y
#undef CHI_BUF_Y
#undef CHI_BUF_S
};

// the warp-synchronous form with one lane (host port: CPU tests)
template <class M, int CB>
CHI_HD void solve_system_uniform(const SolveArgs& a, const int64_t b,
                                 double* sm, const int stride) {
  // the instantiation the launcher would pick (Registrar::launch_stream)
  using S1 = StreamSolver<M, CB, 1>;
  bool subset = a.n_int != S1::PI;
  for (int m = 0; m < S1::PI && !subset; ++m)
    subset = a.int_param[m] != S1::full_param(m);
  if (a.split > 0) {
    StreamSolver<M, CB, 2> s(a, sm, stride);
    s.solve_uniform(b);
  } else if (!a.with_sens) {
    StreamSolver<M, CB, 3> s(a, sm, stride);
    s.solve_uniform(b);
  } else if (subset) {
    StreamSolver<M, CB, 4> s(a, sm, stride);
    s.solve_uniform(b);
  } else {
    S1 s(a, sm, stride);
    s.solve_uniform(b);
  }
}

template <class M, int CB>
CHI_HD void solve_system_stream(const SolveArgs& a, const int64_t b,
                                double* sm, const int stride) {
  StreamSolver<M, CB> s(a, sm, stride);
  s.init(b);
  for (;;) {
    int ev = s.events();
    while (ev == s.EV_WAIT) {   // (one "lane": evaluate at once)
      s.observe_parked();
      ev = s.events();
    }
    if (ev == s.EV_DONE || s.step()) break;
  }
  if (s.pend >= 0 && !s.invalid && s.status() == ST_OK) s.observe_parked();
  s.finish();
}

// Deferred observation of one system: output map, error model, log-likelihood
// and gradient from the snapshots the solve left behind (SolveArgs::snap).
// Same arithmetic as StreamSolver::observe, but every lane of a warp is busy:
// inside the solve this work ran with one or two active lanes per warp.
template <class M>
CHI_HD void observe_system(const SolveArgs& a, const int64_t b) {
  constexpr int N = M::N, NC = M::NC, NCS = M::NCS, P = M::P;
  const int id = a.ids ? a.ids[b] : static_cast<int>(b % a.n_ids);
  const double* xb = a.x + b * a.x_stride;
  const double* sg = xb + P;
  const int nc = a.with_sens ? a.n_cols : 0;
  bool fail = a.status[b] != ST_OK;
  for (int o = 0; o < a.n_out; ++o) {
    const int off = a.err_off[o];
    if (sg[off] <= 0.0) fail = true;
    if (a.err_kind[o] == ERR_CAM && sg[off + 1] <= 0.0) fail = true;
  }
  double c[NCS];
#pragma unroll
  for (int i = 0; i < NC; ++i) c[i] = xb[N + i];
  if (NC == 0) c[0] = 0.0;
  double ll = 0.0, lp = 1.0;
  double G[P], dsig[MAX_SIGMA];
#pragma unroll
  for (int k = 0; k < P; ++k) G[k] = 0.0;
#pragma unroll
  for (int k = 0; k < MAX_SIGMA; ++k) dsig[k] = 0.0;
  if (!fail) {
    const int64_t r0 = a.rec_off[id], r1 = a.rec_off[id + 1];
    const double* row = a.snap + b * a.snap_rows *
                                     static_cast<int64_t>(a.snap_width);
    for (int64_t r = r0; r < r1; ++r, row += a.snap_width) {
      const int o = a.rec_out[r];
      const int off = a.err_off[o];
      double y[N], gy[N], gc[NCS];
#pragma unroll
      for (int n = 0; n < N; ++n) y[n] = row[n];
      const double yb = M::out_grad(a.out_id[o], y, c, gy, gc);
      const ErrTerms e = error_terms(a.err_kind[o], sg + off, yb,
                                     a.rec_obs[r], a.rec_logobs[r]);
      if (e.invalid) fail = true;
      ll += e.ll;
      lp *= e.scale;
      if (!(lp > 1e-270 && lp < 1e270)) {
        ll -= log(lp);
        lp = 1.0;
      }
#pragma unroll
      for (int k = 0; k < MAX_SIGMA; ++k) {
        dsig[k] += (k == off) ? e.ds0 : 0.0;
        dsig[k] += (k == off + 1) ? e.ds1 : 0.0;
      }
      int m = N;
#pragma unroll
      for (int k = 0; k < P; ++k) {
        if (k < nc) {
          const int pk = a.col_param[k];
          double d = 0.0;
#pragma unroll
          for (int j = 0; j < NC; ++j) d = (pk == N + j) ? gc[j] : d;
          if (!(pk >= N && ((M::INERT_MASK >> (pk - N)) & 1u))) {
#pragma unroll
            for (int n = 0; n < N; ++n) d = fma(gy[n], row[m + n], d);
            m += N;
          }
          G[k] = fma(e.w, d, G[k]);
        }
      }
    }
  }
  a.ll[b] = fail ? -INFINITY : ll - log(lp);
  if (a.with_sens && a.grad) {
    double* gb = a.grad + b * a.grad_stride;
#pragma unroll
    for (int k = 0; k < P; ++k)
      if (k < nc) gb[a.col_param[k]] = fail ? INFINITY : G[k];
#pragma unroll
    for (int k = 0; k < MAX_SIGMA; ++k)
      if (k < a.n_sigma) gb[P + k] = fail ? INFINITY : dsig[k];
  }
}

#ifdef __CUDACC__

template <class M>
__global__ void __launch_bounds__(128)
observe_kernel(const SolveArgs a) {
  const int64_t b = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (b < a.B) observe_system<M>(a, b);
}

// Persistent kernel: the lanes pull systems from a global counter, in the
// work order of SolveArgs::order_inner.  With refill_wait = 0 every lane
// refills itself as soon as its system is finished (lanes whose systems need
// fewer steps do not idle); with a warp's worth of parameter vectors per
// individual the warp refills as a whole instead, which keeps its lanes --
// same individual, same stops -- in phase.  No partial last wave either way.

// work item -> system (SolveArgs::order_inner)
__device__ __forceinline__ int64_t stream_system_of(const SolveArgs& a,
                                                    const unsigned long long w) {
  if (a.order_inner <= 1) return static_cast<int64_t>(w);
  const unsigned long long q = w / static_cast<unsigned>(a.order_inner);
  const unsigned long long r = w - q * static_cast<unsigned>(a.order_inner);
  return static_cast<int64_t>(r * static_cast<unsigned>(a.order_stride) + q);
}

// resident blocks per SM the kernel is compiled for (register cap)
#ifndef CHI_ROSL_MIN_BLOCKS
#define CHI_ROSL_MIN_BLOCKS 4
#endif
template <class M, int CB>
constexpr int stream_min_blocks() {
  return stream_is_rosl(CB) ? CHI_ROSL_MIN_BLOCKS : M::STREAM_MIN_BLOCKS;
}

template <class M, int CB, int BLOCK>
__global__ void __launch_bounds__(BLOCK, stream_min_blocks<M, CB>())
solve_stream_kernel(const SolveArgs a, unsigned long long* counter) {
  extern __shared__ double chi_sm[];
  StreamSolver<M, CB> s(a, chi_sm + threadIdx.x, BLOCK);
  bool active = false;
  bool exhausted = false;
  int waited = 0;
  const unsigned long long total = static_cast<unsigned long long>(a.B);
  for (;;) {
    // Refill.  A warp whose lanes are all done takes 32 consecutive work
    // items at once; a single idle lane waits up to `refill_wait` iterations
    // for the rest of its warp (which keeps the lanes in phase when they
    // integrate the same individual, SolveArgs::order_inner) and then takes
    // the next item on its own.  refill_wait = 0: no waiting.
    if (!__any_sync(0xffffffffu, active)) {
      unsigned long long base = 0;
      if ((threadIdx.x & 31) == 0) base = atomicAdd(counter, 32ull);
      base = __shfl_sync(0xffffffffu, base, 0);
      if (base >= total) break;
      const unsigned long long w = base + (threadIdx.x & 31);
      if (w < total) {
        s.init(stream_system_of(a, w));
        active = true;
      } else {
        exhausted = true;
      }
      waited = 0;
    } else if (!active && !exhausted &&
               (waited < a.refill_wait ? (++waited, false) : true)) {
      const unsigned long long w = atomicAdd(counter, 1ull);
      if (w >= total) {
        exhausted = true;
      } else {
        s.init(stream_system_of(a, w));
        active = true;
      }
      waited = 0;
    }
    int ev = active ? s.events() : s.EV_DONE;
    if (StreamSolver<M, CB>::BUFFERED) {
      // records parked in the lanes' buffers are evaluated when every active
      // lane holds one: the handler runs once, with all lanes
      const bool holds = active && s.pend >= 0 && ev != s.EV_DONE;
      if (__any_sync(0xffffffffu, holds) &&
          __all_sync(0xffffffffu, holds || !active || ev == s.EV_DONE)) {
        if (holds) {
          s.observe_parked();
          if (ev == s.EV_WAIT) ev = s.events();
        }
      }
    }
    bool done = ev == s.EV_DONE;
    // the vote is the convergence point: lanes that refilled or handled an
    // event re-join the others here, so the (expensive) step below is
    // executed once per iteration by all lanes that need it
    __syncwarp(0xffffffffu);
    if (active && ev == s.EV_GO) done = s.step();
    if (active && done) {
      s.finish();
      active = false;
    }
  }
}

// Warp-synchronous form (StreamSolver::solve_uniform) for launches in which
// a warp can be given one individual under (up to) 32 parameter vectors:
// B = order_inner x order_stride with at least 32 vectors, of which at most
// an eighth of the lanes end up as padding (the host checks,
// Registrar::launch_stream).  The warps of this kernel never talk to
// each other, so the block is as small as the register file's allocation
// granularity asks for: WARP_BLOCK threads, WARP_MIN_BLOCKS resident blocks
// per SM (the register cap is 65536 / (WARP_BLOCK x WARP_MIN_BLOCKS), rounded
// down to a multiple of 8).
#ifndef CHI_WARP_BLOCK
#define CHI_WARP_BLOCK 128
#endif
#ifndef CHI_WARP_MIN_BLOCKS
#define CHI_WARP_MIN_BLOCKS 4
#endif
constexpr int WARP_BLOCK = CHI_WARP_BLOCK;
constexpr int WARP_MIN_BLOCKS = CHI_WARP_MIN_BLOCKS;

// resident blocks the forward-only instantiation is compiled for
#ifndef CHI_WARP_FWD_MIN_BLOCKS
#define CHI_WARP_FWD_MIN_BLOCKS 4
#endif
// ... and the warp-per-system form: its launches are small by definition, a
// single resident warp per scheduler is the normal case, so it gets the
// registers of three blocks per SM (168) instead of four (128: 78 B spilled)
#ifndef CHI_WARP_SPLIT_MIN_BLOCKS
#define CHI_WARP_SPLIT_MIN_BLOCKS 3
#endif
constexpr int warp_min_blocks(const int form) {
  return form == 3 ? CHI_WARP_FWD_MIN_BLOCKS
                   : (form == 2 ? CHI_WARP_SPLIT_MIN_BLOCKS : WARP_MIN_BLOCKS);
}

// Work order of the batched warp-synchronous launches.  The lanes of a warp
// wait for each other at every stop, so a warp is as slow as its slowest lane
// between any two stops; vectors with similar dynamics take similar step
// sequences.  The vectors of an individual are therefore handed to the warps
// in the order of a stiffness key, -trace J (the sum of the rate constants
// of a mass-action model): measured on the host build (tools/exp/
// group_cost.py) the lane efficiency of the step rises from 0.93 to 0.97 at a
// parameter spread of 10 % and from 0.84 to 0.94 at 30 %.  Two kernels: the
// keys (model specific), then one block per individual sorts its vectors
// (bitonic, shared memory).  Only the assignment of systems to lanes changes:
// results are bit-identical.
constexpr int ORDER_MAX_VECTORS = 4096;

template <class M>
__global__ void __launch_bounds__(256)
order_keys_kernel(const SolveArgs a) {
  const int64_t b = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (b >= a.B) return;
  const double* xb = a.x + b * a.x_stride;
  double y[M::N], c[M::NCS], J[M::N * M::N];
#pragma unroll
  for (int n = 0; n < M::N; ++n) y[n] = xb[n];
#pragma unroll
  for (int j = 0; j < M::NC; ++j) c[j] = xb[M::N + j];
  if (M::NC == 0) c[0] = 0.0;
  M::jac(y, c, J);
  double k = 0.0;
#pragma unroll
  for (int n = 0; n < M::N; ++n) k -= J[n * M::N + n];
  const int64_t vec = b / a.order_stride;          // b = vector * n + individual
  const int64_t ind = b - vec * a.order_stride;
  a.order_keys[ind * a.order_inner + vec] = static_cast<float>(k);
}

// one block per individual; P2 = order_inner rounded up to a power of two
template <int UNUSED>
__global__ void __launch_bounds__(1024)
order_sort_kernel(const int32_t n_vec, const int32_t p2,
                  const float* __restrict__ keys, int32_t* __restrict__ perm) {
  extern __shared__ double chi_sm[];
  float* sk = reinterpret_cast<float*>(chi_sm);
  int32_t* si = reinterpret_cast<int32_t*>(sk + p2);
  const int64_t base = static_cast<int64_t>(blockIdx.x) * n_vec;
  for (int t = threadIdx.x; t < p2; t += blockDim.x) {
    const float k = t < n_vec ? keys[base + t] : INFINITY;
    sk[t] = (k == k) ? k : INFINITY;   // NaN keys sort last
    si[t] = t;
  }
  __syncthreads();
  for (int size = 2; size <= p2; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int t = threadIdx.x; t < (p2 >> 1); t += blockDim.x) {
        const int lo = 2 * t - (t & (stride - 1));
        const int hi = lo + stride;
        const bool up = (lo & size) == 0;
        const float kl = sk[lo], kh = sk[hi];
        const int il = si[lo], ih = si[hi];
        // (ties by index: a deterministic order)
        const bool gt = kl > kh || (kl == kh && il > ih);
        if (gt == up) {
          sk[lo] = kh; sk[hi] = kl;
          si[lo] = ih; si[hi] = il;
        }
      }
      __syncthreads();
    }
  }
  for (int t = threadIdx.x; t < n_vec; t += blockDim.x) perm[base + t] = si[t];
}

// FORM 1: batched launch, all sensitivity columns of the model; FORM 4: a
// subset of them; FORM 3: none (forward only); FORM 2: warp-per-system launch
// (small batches)
template <class M, int CB, int FORM>
__global__ void __launch_bounds__(WARP_BLOCK, warp_min_blocks(FORM))
solve_warp_kernel(const SolveArgs a, unsigned long long* counter) {
  extern __shared__ double chi_sm[];
  StreamSolver<M, CB, FORM> s(a, chi_sm + threadIdx.x, WARP_BLOCK);
  const unsigned lane = threadIdx.x & 31;
  if constexpr (FORM == 2) {
    // the warp takes the `split` work items of ONE system -- its sensitivity
    // columns side by side on the first lanes, each with the state -- and
    // walks through the system's stops like a warp of the batched launch
    // does; the other lanes are ghosts
    const unsigned long long items =
        static_cast<unsigned long long>(a.B) * a.split;
    for (;;) {
      unsigned long long base = 0;
      if (lane == 0)
        base = atomicAdd(counter, static_cast<unsigned long long>(a.split));
      base = __shfl_sync(0xffffffffu, base, 0);
      if (base >= items) break;
      const bool ghost = lane >= static_cast<unsigned>(a.split);
      s.solve_uniform(static_cast<int64_t>(base + (ghost ? 0 : lane)), ghost);
    }
  } else {
    // systems laid out vector-major (b = vector x n + individual,
    // order_inner vectors x order_stride individuals): a warp takes 32
    // consecutive vectors of ONE individual; when the number of vectors is
    // not a multiple of 32 the last warp of an individual is padded with
    // ghosts
    const unsigned long long wpi =
        (static_cast<unsigned long long>(a.order_inner) + 31ull) / 32ull;
    const unsigned long long n_warps =
        static_cast<unsigned long long>(a.order_stride) * wpi;
    for (;;) {
      unsigned long long w = 0;
      if (lane == 0) w = atomicAdd(counter, 1ull);
      w = __shfl_sync(0xffffffffu, w, 0);
      if (w >= n_warps) break;
      const unsigned long long q = w / wpi;
      const unsigned long long r0 = (w - q * wpi) * 32ull;
      const bool ghost = r0 + lane >= static_cast<unsigned>(a.order_inner);
      unsigned long long r = ghost ? r0 : r0 + lane;
      // rank r of individual q in the stiffness order -> vector
      if (a.perm) r = static_cast<unsigned>(a.perm[q * a.order_inner + r]);
      s.solve_uniform(
          static_cast<int64_t>(r * static_cast<unsigned>(a.order_stride) + q),
          ghost);
    }
  }
}

#endif  // __CUDACC__

}  // namespace chi_b200
