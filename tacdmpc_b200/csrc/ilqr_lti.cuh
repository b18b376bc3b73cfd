// ilqr_lti.cuh — whole-solve kernels for the LTI sweep (nx in {8, 16, 32}, nu in {4, 8}, T = 50, dense
// per-element or shared costs): ONE WARP PER ELEMENT, every matrix product of the Riccati step register-tiled.
//
// Why (profiles/r1_forward_generic_nx32): the run-time-dimension kernels keep every matrix of a step in
// shared memory and give each thread 1 x 4 output entries — two shared-memory instructions per four multiply-adds, nine
// barrier-separated phases per step, the nu x nu factorisation on one thread — and ran at 19 % of the issue slots (nx = 32: 6.7
// warps stalled on the CTA barrier per issue) and 8 % of the FMA pipe.  tools/bench_fvf.py settled the formulation for n_x >= 16:
// a warp per element with operands in registers beats both that scheme (2.0x) and a 3 x TF32 mma.sync batched GEMM (1.4x at
// n_x = 32, equal at 16; profiles/r2_probe_fvf.txt), so the products below are FFMA register tiles:
//
//   lanes form an (NRR x NCH) grid, rr = lane / NCH, cc = lane % NCH.  Index set W(cc) = CW consecutive columns ("wide": lanes of
//   one quarter-warp read consecutive 16-byte chunks of ONE row), D(rr) = RW consecutive rows ("deep": one address per rr,
//   broadcast).  Every product is C[D or W][W] += sum_k L[k][.] R[k][.] with both operands read as ROWS k of a shared-memory
//   matrix — 3 vector loads per 32 multiply-adds at n_x = 32 — and the left factors are stored transposed by their producers
//   (conflict-free 16-byte stores) so that no phase reads a column:
//     G  = F' V        Gx^T[j][i] (i in W, j in D),  Gu^T[j][m]                 (controller.py:551-553, (A'V) first)
//     Q  = C_t + G F   Q_xx[i][j] (i in D, j in W) stays in registers; Q_ux, Q_uu, q go through shared memory
//     gains            every lane factorises Q_uu + reg I (Cholesky, LU fallback: quirk Q8) in registers and solves ITS column of
//                      [Q_ux | q_u]: K_t, k_t (:557-566)
//     V' = Q_xx + K'Q_uu K + K'Q_ux + Q_xu K   on the same (D, W) tile, three passes in the reference's association (:573)
//   Five __syncwarp per step instead of nine CTA barriers; F = [A | B] is staged once per CTA.
//   The line search rolls out all candidates at once (lane = state row), scores them with the (N x N) cost block-partitioned over
//   the lanes (one warp reduction per PASS, not per step), and re-rolls the winner out into the alternate trajectory buffer; the
//   element's own cost (quirk Q4) is evaluated for the winner only.  X, U (two buffers), K, k live in a per-warp slice of the
//   caller's workspace (L2-resident).  Elements are pulled from a global queue.
//
// Reference lines restated: controller.py:275-315 (solve_step), 546-578 (backward_lqr), 589-620 (evaluate_alphas), 63-115 +
// 707-788 (backward, un-boxed); cost.py:37-100.  fp32 and fp64 instances (the fp64 ones exist for the parity tests).
#pragma once
#include "ilqr_generic.cuh"
#include "ilqr_small.cuh"

namespace tacd {

constexpr int kLtiNA = 5;   // line-search candidates rolled out together (n_alpha <= 5, the reference's default)

template <typename R, int W>
struct alignas((W * sizeof(R)) >= 16 ? 16 : (W * sizeof(R))) LVec { R v[W]; };
template <int W, typename R>
TACD_DEV void ldv(R (&d)[W], const R* p) {
  const LVec<R, W> t = *reinterpret_cast<const LVec<R, W>*>(p);
#pragma unroll
  for (int i = 0; i < W; ++i) d[i] = t.v[i];
}
template <int W, typename R>
TACD_DEV void stv(R* p, const R (&s)[W]) {
  LVec<R, W> t;
#pragma unroll
  for (int i = 0; i < W; ++i) t.v[i] = s[i];
  *reinterpret_cast<LVec<R, W>*>(p) = t;
}

// acc[0..W) += l * r[0..W).  fp32: packed fma.rn.f32x2 (SASS FFMA2, sm_100): two IEEE fused multiply-adds per issue slot; the
// {l, l} pair is the instruction's scalar-broadcast operand, r comes out of a 64/128-bit shared-memory load as an aligned pair
// and the accumulators live in pairs.  Same roundings as W scalar FFMAs.
template <int W, typename R>
TACD_DEV void axpyv(R (&acc)[W], R l, const R (&r)[W]) {
#pragma unroll
  for (int c = 0; c < W; ++c) acc[c] += l * r[c];
}
template <int W>
TACD_DEV void axpyv_f32x2(float (&acc)[W], float l, const float (&r)[W]) {
  static_assert(W % 2 == 0, "pairs");
#pragma unroll
  for (int h = 0; h < W / 2; ++h) {
    unsigned long long lb, rb, ab;
    asm("mov.b64 %0, {%1, %1};" : "=l"(lb) : "f"(l));
    asm("mov.b64 %0, {%1, %2};" : "=l"(rb) : "f"(r[2 * h]), "f"(r[2 * h + 1]));
    asm("mov.b64 %0, {%1, %2};" : "=l"(ab) : "f"(acc[2 * h]), "f"(acc[2 * h + 1]));
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(ab) : "l"(lb), "l"(rb));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(acc[2 * h]), "=f"(acc[2 * h + 1]) : "l"(ab));
  }
}
#ifndef TACD_LTI_NO_FFMA2
template <> TACD_DEV void axpyv<2, float>(float (&acc)[2], float l, const float (&r)[2]) { axpyv_f32x2<2>(acc, l, r); }
template <> TACD_DEV void axpyv<4, float>(float (&acc)[4], float l, const float (&r)[4]) { axpyv_f32x2<4>(acc, l, r); }
template <> TACD_DEV void axpyv<8, float>(float (&acc)[8], float l, const float (&r)[8]) { axpyv_f32x2<8>(acc, l, r); }
#endif

// a[i] for a run-time i < M with a in registers
template <int M, typename R>
TACD_DEV R lti_pick(const R* a, int i) {
  R r = a[0];
#pragma unroll
  for (int k = 1; k < M; ++k) r = (i == k) ? a[k] : r;
  return r;
}

template <int NX_, int NU_>
struct LtiCfg {
  static constexpr int NX = NX_, NU = NU_, N = NX + NU;
  static constexpr int CW = NX >= 16 ? 4 : 2;          // wide chunk
  static constexpr int NCH = NX / CW, NRR = 32 / NCH;  // lane grid
  static constexpr int RW = NX / NRR;                  // deep chunk
  static constexpr int MS = NU / NCH;                  // Q_xu columns per lane
  static constexpr int LPM = 32 / NU, UC = NX / LPM;   // (B'V): lanes per input row, columns per lane
  static constexpr int UQ = (NU * NU >= 32) ? NU * NU / 32 : 1;   // Q_uu entries per lane
  static constexpr int LDV = (NX == 16) ? 24 : NX;     // row stride of V and Gx^T (16: two rows per quarter-warp on disjoint banks)
  static_assert(NCH * MS == NU && NRR * RW == NX && LPM * UC == NX && 32 % NU == 0 && NX <= 32 && N % 4 == 0, "lane grid");
  // line-search cost blocks: (N x N) over CRG x CCG lanes when N is a multiple of 8, else one row per lane
  static constexpr bool BLK = (N % 8 == 0);
  static constexpr int CR = BLK ? N / 8 : 1, CC = BLK ? N / 4 : N;
  static constexpr int CV = BLK ? 2 : 4;               // vector width of the cost-block loads
  static_assert(CC % CV == 0 && (BLK || N <= 32), "cost blocks");
  // shared-memory words of one warp: Riccati temporaries and line-search buffers share the region
  static constexpr int oV = 0, oGx = oV + NX * LDV, oGu = oGx + NX * LDV, oQux = oGu + NX * NU, oKs = oQux + NU * NX,
                       oQK = oKs + NU * NX, oQuu = oQK + NU * NX, ov = oQuu + NU * NU, oq = ov + NX, ol = oq + N, otau = ol + N,
                       okt = otau + N, RIC = okt + NU;
  static constexpr int oz = 0, otaua = oz + 2 * kLtiNA * N, odx = otaua + kLtiNA * N, LS = odx + kLtiNA * NX;
  static constexpr int WARP_WORDS = (((RIC > LS ? RIC : LS) + 3) / 4) * 4;
  static constexpr int F_WORDS = 2 * NX * N;           // F [NX][N] and F^T [N][NX]
};

// per-warp slice of the workspace: two trajectory buffers, the gains
__host__ __device__ inline size_t lti_slice_words(int T, int nx, int nu) {
  return 2 * ((size_t)(T + 1) * nx + (size_t)T * nu) + (size_t)T * nu * nx + (size_t)T * nu + 8;
}

template <typename R, int NX, int NU>
struct LtiWarp {
  using C = LtiCfg<NX, NU>;
  static constexpr int N = C::N, CW = C::CW, RW = C::RW, MS = C::MS, UC = C::UC, UQ = C::UQ, LDV = C::LDV, NCH = C::NCH;
  static constexpr unsigned FULL = 0xffffffffu;
  const R* sF;    // [NX][N]
  const R* sFT;   // [N][NX]
  R* sm;          // this warp's region
  int lane, cc, rr, m, ju0, mq, e0, e1;
  bool has1;

  TACD_DEV void init(const R* F, R* region) {
    sF = F; sFT = F + NX * N; sm = region;
    lane = threadIdx.x & 31;
    cc = lane % NCH; rr = lane / NCH;
    m = lane % NU; ju0 = (lane / NU) * UC; mq = ((lane / NU) * UQ) % NU;
    e0 = lane < N ? lane : N - 1;
    has1 = 32 + lane < N;
    e1 = has1 ? 32 + lane : N - 1;
  }
  TACD_DEV R* Vs() const { return sm + C::oV; }
  TACD_DEV R* GxT() const { return sm + C::oGx; }
  TACD_DEV R* GuT() const { return sm + C::oGu; }
  TACD_DEV R* Qux() const { return sm + C::oQux; }
  TACD_DEV R* Ks() const { return sm + C::oKs; }
  TACD_DEV R* QKs() const { return sm + C::oQK; }
  TACD_DEV R* Quu() const { return sm + C::oQuu; }
  TACD_DEV R* vs() const { return sm + C::ov; }
  TACD_DEV R* qv() const { return sm + C::oq; }
  TACD_DEV R* lv() const { return sm + C::ol; }
  TACD_DEV R* tau() const { return sm + C::otau; }
  TACD_DEV R* kts() const { return sm + C::okt; }

  // the step's cost Hessian in the layout of the Q tiles
  struct CTile { R qxx[RW][CW]; R qxu[RW][MS]; R quu[UQ]; };
  TACD_DEV void load_ctile(CTile& c, const R* Ct) const {
#pragma unroll
    for (int a = 0; a < RW; ++a) {
      ldv<CW>(c.qxx[a], Ct + (size_t)(rr * RW + a) * N + cc * CW);
      ldv<MS>(c.qxu[a], Ct + (size_t)(rr * RW + a) * N + NX + cc * MS);
    }
    ldv<UQ>(c.quu, Ct + (size_t)(NX + m) * N + NX + mq);
  }

  // Everything one step of the reverse pass reads from global memory, in registers: the cost tile, the rows >= NX of C_t for
  // l_u (entry e0 / e1 of every row), c_t and the two operands of tau.  With PF the loads of step t-1 are issued before the
  // arithmetic of step t (small systems are a latency chain: one DRAM / L2 round trip per step was ~1 us of a ~1.6 us step at
  // n_x = 8); at n_x = 32 the 69 registers are not there and the loads are issued at the top of the step instead.
  static constexpr bool PF = NX <= 16;
  struct StepIn { CTile c; R lu0[NU], lu1[N > 32 ? NU : 1]; R cx[RW], cu, ta0, tb0, ta1, tb1; };
  // FWD: quadraticisation inputs (tau = a - b, c_t); else the KKT pass: lv = -grad
  template <bool FWD>
  TACD_DEV void load_step(StepIn& s, int t, const R* Cp, const R* cp, const R* Xn, const R* Un, const R* xr, const R* ur) const {
    const R* Ct = Cp + (size_t)t * N * N;
    load_ctile(s.c, Ct);
    if (FWD) {
#pragma unroll
      for (int mm = 0; mm < NU; ++mm) {
        s.lu0[mm] = Ct[(size_t)(NX + mm) * N + e0];
        if (N > 32) s.lu1[mm] = Ct[(size_t)(NX + mm) * N + e1];
      }
      const R* ct = cp + (size_t)t * N;
#pragma unroll
      for (int a = 0; a < RW; ++a) s.cx[a] = ct[rr * RW + a];
      s.cu = ct[NX + m];
      s.ta0 = (e0 < NX) ? Xn[(size_t)t * NX + e0] : Un[(size_t)t * NU + e0 - NX];
      s.tb0 = (e0 < NX) ? xr[(size_t)t * NX + e0] : ur[(size_t)t * NU + e0 - NX];
      if (N > 32) { s.ta1 = Un[(size_t)t * NU + e1 - NX]; s.tb1 = ur[(size_t)t * NU + e1 - NX]; }
    } else {
      // Xn = grad_X, Un = grad_U
      s.ta0 = (e0 < NX) ? Xn[(size_t)t * NX + e0] : Un[(size_t)t * NU + e0 - NX];
      if (N > 32) s.ta1 = Un[(size_t)t * NU + e1 - NX];
    }
  }
  TACD_DEV void store_tau(const StepIn& s) const {
    if (lane < N) tau()[e0] = s.ta0 - s.tb0;
    if (N > 32 && has1) tau()[e1] = s.ta1 - s.tb1;
  }
  // l = C_t tau + c_t (cost.py:86-88) -> lv[0..N): rows < NX from the tile (partial sums over the lane's columns, reduced over
  // cc), rows >= NX by the whole warp (entry e0 / e1 of the row per lane).  tau must be in shared memory.
  TACD_DEV void cost_linear(const StepIn& s) const {
    const R* t = tau();
    R tw[CW], tm[MS];
    ldv<CW>(tw, t + cc * CW);
    ldv<MS>(tm, t + NX + cc * MS);
#pragma unroll
    for (int a = 0; a < RW; ++a) {
      R acc = 0;
#pragma unroll
      for (int e = 0; e < CW; ++e) acc += s.c.qxx[a][e] * tw[e];
#pragma unroll
      for (int e = 0; e < MS; ++e) acc += s.c.qxu[a][e] * tm[e];
#pragma unroll
      for (int o = 1; o < NCH; o <<= 1) acc += __shfl_xor_sync(FULL, acc, o);
      if (cc == 0) lv()[rr * RW + a] = acc + s.cx[a];
    }
    const R t0 = t[e0], t1 = t[e1];
#pragma unroll
    for (int mm = 0; mm < NU; ++mm) {
      R acc = (lane < N) ? s.lu0[mm] * t0 : (R)0;
      if (N > 32) acc += has1 ? s.lu1[mm] * t1 : (R)0;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(FULL, acc, o);
      if (lane == mm) lv()[NX + mm] = acc + s.cu;
    }
  }

  // The gain solve.  Q_uu + reg I is factorised ONCE per warp with its rows distributed over the lanes (row = lane % NU; the
  // 32 / NU lane groups are replicas, so every shuffle source is valid): right-looking Cholesky — the same subtractions in the
  // same order as the left-looking dot products of ilqr_generic.cuh::gen_chol — and, when that fails (or in the KKT pass), LU with
  // partial pivoting (first maximum, whole rows swapped, as LAPACK getrf / the reference's linalg.solve).  The factors go through
  // shared memory; every lane then solves ITS column of [Q_ux | q_u] with the triangular factors in registers.  Round 2's first
  // version factorised redundantly on every lane in registers: 64 + 64 live values, select chains for the row swaps and 1.8 k of
  // a step's 9.5 k instructions (profiles/r2_lti_fwd_nx32_v1).  Divisions are correctly rounded (div_rn_): with the 1-ulp div_ the
  // accept / reject decisions of the unstable sweep moved (mean iterations 1.0 -> 2.5 at n_x = 32: measured).
  template <bool KKT>
  TACD_DEV bool gains(const DevProblem<R>& P, R (&Kc)[NU], R (&kt)[NU], R (&Qxi)[NU], R (&qu)[NU]) const {
    const R reg2 = KKT ? (R)1e-8 * P.reg : P.reg;
    R* Ls = GxT();                       // factors, row-major [NU][NU] (Gx^T is dead after the Q phase)
    R* LTs = Ls + NU * NU;               // Cholesky: L transposed
    int* perm = reinterpret_cast<int*>(LTs + NU * NU);
    const int li = li_();
    R a[NU];
    auto load_row = [&]() {
      ldv<NU>(a, Quu() + m * NU);
#pragma unroll
      for (int k = 0; k < NU; ++k) a[k] += (k == m) ? reg2 : (R)0;
    };
    load_row();
    bool chol = false;
    if (!KKT) {
      bool ok = true;
      R lrow[NU];
#pragma unroll
      for (int j = 0; j < NU; ++j) {
        const R d = __shfl_sync(FULL, a[j], j);
        if (!(d > 0)) ok = false;
        const R dd = sqrt_<R>(d);
        R l = div_rn_<R>(a[j], dd);
        l = (m == j) ? dd : l;
        lrow[j] = (m >= j) ? l : (R)0;
#pragma unroll
        for (int k = j + 1; k < NU; ++k) {
          const R lk = __shfl_sync(FULL, l, k);
          a[k] -= l * lk;
        }
      }
      chol = ok;
      if (ok) {
        stv<NU>(Ls + m * NU, lrow);
#pragma unroll
        for (int k = 0; k < NU; ++k) LTs[k * NU + m] = lrow[k];
      }
    }
    if (!chol) {
      if (!KKT) load_row();
      int rowid = m;
#pragma unroll
      for (int j = 0; j < NU; ++j) {
        R v = (m >= j) ? abs_<R>(a[j]) : (R)-1;
        int idx = m;
#pragma unroll
        for (int o = 1; o < NU; o <<= 1) {
          const R v2 = __shfl_xor_sync(FULL, v, o);
          const int i2 = __shfl_xor_sync(FULL, idx, o);
          const bool take = (v2 > v) || (v2 == v && i2 < idx);
          v = take ? v2 : v;
          idx = take ? i2 : idx;
        }
        const int src = (m == j) ? idx : ((m == idx) ? j : m);
#pragma unroll
        for (int k = 0; k < NU; ++k) a[k] = __shfl_sync(FULL, a[k], src);
        rowid = __shfl_sync(FULL, rowid, src);
        const R pj = __shfl_sync(FULL, a[j], j);
        const R f = div_rn_<R>(a[j], pj);
        if (m > j) a[j] = f;
#pragma unroll
        for (int k = j + 1; k < NU; ++k) {
          const R pk = __shfl_sync(FULL, a[k], j);
          if (m > j) a[k] -= f * pk;
        }
      }
      stv<NU>(Ls + m * NU, a);
      perm[m] = rowid;
    }
    __syncwarp();
#pragma unroll
    for (int mm = 0; mm < NU; ++mm) Qxi[mm] = Qux()[mm * NX + li];
    ldv<NU>(qu, qv() + NX);
    R y[NU], yk[NU];
    if (chol) {
#pragma unroll
      for (int i = 0; i < NU; ++i) {
        R row[NU];
        ldv<NU>(row, Ls + i * NU);
        R s = Qxi[i], sk = qu[i];
#pragma unroll
        for (int k = 0; k < i; ++k) { s -= row[k] * y[k]; sk -= row[k] * yk[k]; }
        y[i] = div_rn_<R>(s, row[i]);
        yk[i] = div_rn_<R>(sk, row[i]);
      }
#pragma unroll
      for (int i = NU - 1; i >= 0; --i) {
        R row[NU];
        ldv<NU>(row, LTs + i * NU);
        R s = y[i], sk = yk[i];
#pragma unroll
        for (int k = i + 1; k < NU; ++k) { s -= row[k] * y[k]; sk -= row[k] * yk[k]; }
        y[i] = div_rn_<R>(s, row[i]);
        yk[i] = div_rn_<R>(sk, row[i]);
      }
    } else {
      int pr[NU];
      ldv<NU>(pr, perm);
#pragma unroll
      for (int i = 0; i < NU; ++i) {
        R row[NU];
        ldv<NU>(row, Ls + i * NU);
        R s = Qux()[pr[i] * NX + li], sk = qv()[NX + pr[i]];
#pragma unroll
        for (int k = 0; k < i; ++k) { s -= row[k] * y[k]; sk -= row[k] * yk[k]; }
        y[i] = s;
        yk[i] = sk;
      }
#pragma unroll
      for (int i = NU - 1; i >= 0; --i) {
        R row[NU];
        ldv<NU>(row, Ls + i * NU);
        R s = y[i], sk = yk[i];
#pragma unroll
        for (int k = i + 1; k < NU; ++k) { s -= row[k] * y[k]; sk -= row[k] * yk[k]; }
        y[i] = div_rn_<R>(s, row[i]);
        yk[i] = div_rn_<R>(sk, row[i]);
      }
    }
#pragma unroll
    for (int mm = 0; mm < NU; ++mm) { Kc[mm] = -y[mm]; kt[mm] = -yk[mm]; }
    return !KKT && !chol;
  }

  // One step of the value recursion.  In: Vs, vs (value at t+1), lv (linear cost terms; KKT: the negated incoming gradients),
  // the cost tile.  Out: Vs, vs at t; K_t -> gK [NU][NX], k_t -> gk [NU] (global).  Returns true if Q_uu + reg I was not
  // positive definite (forward only).  KKT = the backward pass's variant without input bounds (controller.py:736-767):
  // LU of Q_uu + 1e-8 reg I, k_t one more right-hand side (ilqr_generic.cuh::gen_value_step explains why that is the box QP).
  template <bool KKT>
  TACD_DEV bool value_step(const DevProblem<R>& P, const CTile& c, R* gK, R* gk) const {
    // ---- G = F' V, F' v
    {
      R gx[RW][CW], gu[UC], fv0 = 0, fv1 = 0;
#pragma unroll
      for (int a = 0; a < RW; ++a)
#pragma unroll
        for (int e = 0; e < CW; ++e) gx[a][e] = 0;
#pragma unroll
      for (int e = 0; e < UC; ++e) gu[e] = 0;
#pragma unroll 4
      for (int k = 0; k < NX; ++k) {
        R aw[CW], vd[RW], vu[UC];
        ldv<CW>(aw, sF + k * N + cc * CW);
        ldv<RW>(vd, Vs() + k * LDV + rr * RW);
#pragma unroll
        for (int a = 0; a < RW; ++a) axpyv<CW, R>(gx[a], vd[a], aw);
        const R bm = sF[k * N + NX + m];
        ldv<UC>(vu, Vs() + k * LDV + ju0);
        axpyv<UC, R>(gu, bm, vu);
        const R vk = vs()[k];
        fv0 += sF[k * N + e0] * vk;
        if (N > 32) fv1 += sF[k * N + e1] * vk;
      }
#pragma unroll
      for (int a = 0; a < RW; ++a) stv<CW>(GxT() + (rr * RW + a) * LDV + cc * CW, gx[a]);
#pragma unroll
      for (int e = 0; e < UC; ++e) GuT()[(ju0 + e) * NU + m] = gu[e];
      if (lane < N) qv()[e0] = lv()[e0] + fv0;
      if (N > 32 && has1) qv()[e1] = lv()[e1] + fv1;
    }
    __syncwarp();
    // ---- Q = C_t + G F
    R qxx[RW][CW];
    {
      R axx[RW][CW], axu[RW][MS], auu[UQ];
#pragma unroll
      for (int a = 0; a < RW; ++a) {
#pragma unroll
        for (int e = 0; e < CW; ++e) axx[a][e] = 0;
#pragma unroll
        for (int e = 0; e < MS; ++e) axu[a][e] = 0;
      }
#pragma unroll
      for (int e = 0; e < UQ; ++e) auu[e] = 0;
#pragma unroll 4
      for (int k = 0; k < NX; ++k) {
        R gd[RW], aw[CW], bw[MS], bq[UQ];
        ldv<RW>(gd, GxT() + k * LDV + rr * RW);
        ldv<CW>(aw, sF + k * N + cc * CW);
        ldv<MS>(bw, sF + k * N + NX + cc * MS);
#pragma unroll
        for (int a = 0; a < RW; ++a) {
          axpyv<CW, R>(axx[a], gd[a], aw);
          axpyv<MS, R>(axu[a], gd[a], bw);
        }
        const R gum = GuT()[k * NU + m];
        ldv<UQ>(bq, sF + k * N + NX + mq);
#pragma unroll
        for (int e = 0; e < UQ; ++e) auu[e] += gum * bq[e];
      }
#pragma unroll
      for (int a = 0; a < RW; ++a) {
#pragma unroll
        for (int e = 0; e < CW; ++e) qxx[a][e] = c.qxx[a][e] + axx[a][e];
#pragma unroll
        for (int e = 0; e < MS; ++e) Qux()[(cc * MS + e) * NX + rr * RW + a] = c.qxu[a][e] + axu[a][e];
      }
#pragma unroll
      for (int e = 0; e < UQ; ++e) Quu()[m * NU + mq + e] = c.quu[e] + (auu[e] + ((m == mq + e) ? P.reg : (R)0));   // + reg I (:554)
    }
    __syncwarp();
    // ---- gains: K_t, k_t = -(Q_uu + reg I)^-1 [Q_ux | q_u]  (:557-566)
    const int li = li_();
    R Kc[NU], kt[NU], Qxi[NU], qu[NU];
    const bool nonpd = gains<KKT>(P, Kc, kt, Qxi, qu);
    {
      // Q_uu K (column li), Q_uu k, the new v_li (:574)
      R QKc[NU], Qk[NU];
#pragma unroll
      for (int mm = 0; mm < NU; ++mm) {
        R row[NU];
        ldv<NU>(row, Quu() + mm * NU);
        R s = 0, s2 = 0;
#pragma unroll
        for (int m2 = 0; m2 < NU; ++m2) { s += row[m2] * Kc[m2]; s2 += row[m2] * kt[m2]; }
        QKc[mm] = s; Qk[mm] = s2;
      }
      R a = 0, b = 0, cq = 0;
#pragma unroll
      for (int mm = 0; mm < NU; ++mm) { a += Kc[mm] * Qk[mm]; b += Kc[mm] * qu[mm]; cq += Qxi[mm] * kt[mm]; }
      const R vi = ((qv()[li] + a) + b) + cq;
      if (lane < NX) {
#pragma unroll
        for (int mm = 0; mm < NU; ++mm) {
          Ks()[mm * NX + lane] = Kc[mm];
          QKs()[mm * NX + lane] = QKc[mm];
          gK[mm * NX + lane] = Kc[mm];
        }
        vs()[lane] = vi;
      }
      if (lane < NU) { gk[lane] = lti_pick<NU, R>(kt, lane); kts()[lane] = lti_pick<NU, R>(kt, lane); }
    }
    __syncwarp();
    // ---- V = Q_xx + K'(Q_uu K) + K' Q_ux + Q_xu K on the (D, W) tile, the reference's association
    {
      R acc[RW][CW];
      auto zero = [&]() {
#pragma unroll
        for (int a = 0; a < RW; ++a)
#pragma unroll
          for (int e = 0; e < CW; ++e) acc[a][e] = 0;
      };
      auto pass = [&](const R* Lm, const R* Rm) {   // acc[a][e] = sum_m Lm[m][D_a] Rm[m][W_e]
        zero();
#pragma unroll
        for (int mm = 0; mm < NU; ++mm) {
          R ld_[RW], rw[CW];
          ldv<RW>(ld_, Lm + mm * NX + rr * RW);
          ldv<CW>(rw, Rm + mm * NX + cc * CW);
#pragma unroll
          for (int a = 0; a < RW; ++a) axpyv<CW, R>(acc[a], ld_[a], rw);
        }
#pragma unroll
        for (int a = 0; a < RW; ++a)
#pragma unroll
          for (int e = 0; e < CW; ++e) qxx[a][e] += acc[a][e];
      };
      pass(Ks(), QKs());    // K' (Q_uu K)
      pass(Ks(), Qux());    // K' Q_ux
      pass(Qux(), Ks());    // Q_xu K
#pragma unroll
      for (int a = 0; a < RW; ++a) stv<CW>(Vs() + (rr * RW + a) * LDV + cc * CW, qxx[a]);
    }
    __syncwarp();
    return nonpd;
  }

  // ---- line search / rollouts -------------------------------------------------------------------------------------------
  TACD_DEV R* za(int buf) const { return sm + C::oz + buf * kLtiNA * N; }   // [cand][N]: candidate (x, u)
  TACD_DEV R* taua() const { return sm + C::otaua; }                        // [cand][N]
  TACD_DEV R* dxa() const { return sm + C::odx; }                           // [cand][NX]

  // This lane's block of the step's cost, in registers: rows r0.., columns c0.. of C_t and entries e0 / e1 of c_t
  struct CostIn { R cb[C::CR][C::CC]; R cl0, cl1; };
  TACD_DEV void load_cost(CostIn& ci, const R* Ct, const R* ct) const {
    constexpr int CR = C::CR, CC = C::CC, CV = C::CV;
    const int r0 = C::BLK ? (lane / 4) * CR : e0, c0 = C::BLK ? (lane % 4) * CC : 0;
#pragma unroll
    for (int r = 0; r < CR; ++r)
#pragma unroll
      for (int q = 0; q < CC; q += CV) {
        R t[CV];
        ldv<CV>(t, Ct + (size_t)(r0 + r) * N + c0 + q);
#pragma unroll
        for (int e = 0; e < CV; ++e) ci.cb[r][q + e] = t[e];
      }
    ci.cl0 = ct[e0];
    ci.cl1 = ct[e1];
  }
  // quad[a] += sum over this lane's block of tau_a[r] C[r][c] tau_a[c];  lin[a] += c_t[e] tau_a[e] (entries e0, e1)
  template <int NC>
  TACD_DEV void cost_block(const CostIn& ci, R (&quad)[NC], R (&lin)[NC]) const {
    constexpr int CR = C::CR, CC = C::CC, CV = C::CV;
    const int r0 = C::BLK ? (lane / 4) * CR : e0, c0 = C::BLK ? (lane % 4) * CC : 0;
    const bool live = C::BLK || lane < N;
#pragma unroll
    for (int a = 0; a < NC; ++a) {
      const R* ta = taua() + a * N;
      R tc[CC];
#pragma unroll
      for (int q = 0; q < CC; q += CV) {
        R t[CV];
        ldv<CV>(t, ta + c0 + q);
#pragma unroll
        for (int e = 0; e < CV; ++e) tc[q + e] = t[e];
      }
#pragma unroll
      for (int r = 0; r < CR; ++r) {
        R s = 0;
#pragma unroll
        for (int q = 0; q < CC; ++q) s += ci.cb[r][q] * tc[q];
        if (live) quad[a] += ta[r0 + r] * s;
      }
      if (lane < N) lin[a] += ci.cl0 * ta[e0];
      if (N > 32 && has1) lin[a] += ci.cl1 * ta[e1];
    }
  }
  // terminal terms (cost.py:52-58): lane i < NX owns row i of C_final[:nx, :nx]; tauN in taua()[a][0..NX)
  template <int NC>
  TACD_DEV void cost_terminal(const R* Cf, const R* cf, R (&quad)[NC], R (&lin)[NC]) const {
    if (lane < NX) {
#pragma unroll
      for (int a = 0; a < NC; ++a) {
        const R* ta = taua() + a * N;
        R s = 0;
        for (int j = 0; j < NX; ++j) s += Cf[(size_t)lane * N + j] * ta[j];
        quad[a] += ta[lane] * s;
        lin[a] += cf[lane] * ta[lane];
      }
    }
  }
  static TACD_DEV R warp_sum(R v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
  }

  // Everything one step of a rollout reads from global memory (prefetched one step ahead when PF)
  struct RollIn { R xt, kr[UC], un, kk, rf0, rf1; CostIn ci; };
  template <bool COST>
  TACD_DEV void load_roll(RollIn& r, int t, const R* Xn, const R* Un, const R* gK, const R* gk, const R* xr, const R* ur, const R* Cq,
                          const R* cq) const {
    r.xt = Xn[(size_t)t * NX + li_()];
    ldv<UC>(r.kr, gK + (size_t)t * NU * NX + m * NX + ju0);
    r.un = Un[(size_t)t * NU + m];
    r.kk = gk[(size_t)t * NU + m];
    if (COST) {
      r.rf0 = (e0 < NX) ? xr[(size_t)t * NX + e0] : ur[(size_t)t * NU + (e0 - NX)];
      r.rf1 = ur[(size_t)t * NU + (e1 - NX < 0 ? 0 : e1 - NX)];
      load_cost(r.ci, Cq + (size_t)t * N * N, cq + (size_t)t * N);
    }
  }
  // Roll NC candidates out along the gains (controller.py:598-612): u = clamp(U_t + alpha k_t + K_t (x - X_t)), x+ = A x + B u,
  // cost of every candidate with (Cq, cq, Cfq, cfq) when COST.  STORE (NC = 1): the trajectory goes to Xo / Uo.
  template <int NC, bool STORE, bool COST>
  TACD_DEV void rollout(const DevProblem<R>& P, int T, const R (&alpha)[NC], const R* Xn, const R* Un, const R* gK, const R* gk, const R* xr,
                        const R* ur, const R* Cq, const R* cq, const R* Cfq, const R* cfq, R* Xo, R* Uo, R (&cost)[NC]) const {
    R quad[NC], lin[NC], qN[NC], lN[NC];
#pragma unroll
    for (int a = 0; a < NC; ++a) { quad[a] = 0; lin[a] = 0; qN[a] = 0; lN[a] = 0; }
    int cur = 0;
    if (lane < NX) {
      const R x0 = Xn[lane];
#pragma unroll
      for (int a = 0; a < NC; ++a) za(0)[a * N + lane] = x0;
      if (STORE) Xo[lane] = x0;
    }
    __syncwarp();
    const int jp = lane / NU;
    RollIn ri, rn;
    if (PF) load_roll<COST>(ri, 0, Xn, Un, gK, gk, xr, ur, Cq, cq);
    for (int t = 0; t < T; ++t) {
      if (!PF) load_roll<COST>(ri, t, Xn, Un, gK, gk, xr, ur, Cq, cq);
      else if (t + 1 < T) load_roll<COST>(rn, t + 1, Xn, Un, gK, gk, xr, ur, Cq, cq);
      const R* z = za(cur);
      R* zn = za(cur ^ 1);
      // dx_a = x_a - X_t
      if (lane < NX) {
#pragma unroll
        for (int a = 0; a < NC; ++a) dxa()[a * NX + lane] = z[a * N + lane] - ri.xt;
      }
      __syncwarp();
#pragma unroll
      for (int a = 0; a < NC; ++a) {
        R dv[UC];
        ldv<UC>(dv, dxa() + a * NX + ju0);
        R s = 0;
#pragma unroll
        for (int e = 0; e < UC; ++e) s += ri.kr[e] * dv[e];
#pragma unroll
        for (int o = NU; o < 32; o <<= 1) s += __shfl_xor_sync(FULL, s, o);
        if (jp == 0) {
          R ui = ri.un + (alpha[a] * ri.kk + s);
          if (P.has_umin) ui = tmax<R>(ui, P.umin[m]);
          if (P.has_umax) ui = tmin<R>(ui, P.umax[m]);
          za(cur)[a * N + NX + m] = ui;
          if (STORE) Uo[(size_t)t * NU + m] = ui;
        }
      }
      __syncwarp();
      if (COST) {
#pragma unroll
        for (int a = 0; a < NC; ++a) {
          if (lane < N) taua()[a * N + e0] = z[a * N + e0] - ri.rf0;
          if (N > 32 && has1) taua()[a * N + e1] = z[a * N + e1] - ri.rf1;
        }
      }
      // x+ = A x + B u: lane i owns row i (read from F^T, conflict-free)
      {
        R sx[NC], su[NC];
#pragma unroll
        for (int a = 0; a < NC; ++a) { sx[a] = 0; su[a] = 0; }
#pragma unroll 4
        for (int k = 0; k < NX; k += 4) {
          R f[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) f[e] = sFT[(k + e) * NX + li_()];
#pragma unroll
          for (int a = 0; a < NC; ++a) {
            R x4[4];
            ldv<4>(x4, z + a * N + k);
#pragma unroll
            for (int e = 0; e < 4; ++e) sx[a] += f[e] * x4[e];
          }
        }
#pragma unroll
        for (int k = 0; k < NU; k += 4) {
          R f[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) f[e] = sFT[(NX + k + e) * NX + li_()];
#pragma unroll
          for (int a = 0; a < NC; ++a) {
            R u4[4];
            ldv<4>(u4, z + a * N + NX + k);
#pragma unroll
            for (int e = 0; e < 4; ++e) su[a] += f[e] * u4[e];
          }
        }
        if (lane < NX) {
#pragma unroll
          for (int a = 0; a < NC; ++a) zn[a * N + lane] = sx[a] + su[a];
          if (STORE) Xo[(size_t)(t + 1) * NX + lane] = sx[0] + su[0];
        }
      }
      __syncwarp();
      if (COST) cost_block<NC>(ri.ci, quad, lin);
      cur ^= 1;
      if (PF && t + 1 < T) ri = rn;
    }
    if (COST) {
      __syncwarp();
      if (lane < NX) {
        const R rT = xr[(size_t)T * NX + lane];
#pragma unroll
        for (int a = 0; a < NC; ++a) taua()[a * N + lane] = za(cur)[a * N + lane] - rT;
      }
      __syncwarp();
      cost_terminal<NC>(Cfq, cfq, qN, lN);
#pragma unroll
      for (int a = 0; a < NC; ++a) {
        const R q = warp_sum(quad[a]), l = warp_sum(lin[a]), q2 = warp_sum(qN[a]), l2 = warp_sum(lN[a]);
        cost[a] = ((R)0.5 * q + l) + ((R)0.5 * q2 + l2);
      }
    }
    __syncwarp();
  }
  TACD_DEV int li_() const { return lane < NX ? lane : NX - 1; }

  // objective of a stored trajectory (cost.py:37-62)
  TACD_DEV R traj_cost(int T, const R* X, const R* U, const R* xr, const R* ur, const R* Cq, const R* cq, const R* Cfq, const R* cfq) const {
    R quad[1] = {0}, lin[1] = {0}, qN[1] = {0}, lN[1] = {0};
    for (int t = 0; t < T; ++t) {
      if (lane < N) taua()[e0] = (e0 < NX) ? X[(size_t)t * NX + e0] - xr[(size_t)t * NX + e0] : U[(size_t)t * NU + e0 - NX] - ur[(size_t)t * NU + e0 - NX];
      if (N > 32 && has1) taua()[e1] = U[(size_t)t * NU + e1 - NX] - ur[(size_t)t * NU + e1 - NX];
      __syncwarp();
      CostIn ci;
      load_cost(ci, Cq + (size_t)t * N * N, cq + (size_t)t * N);
      cost_block<1>(ci, quad, lin);
      __syncwarp();
    }
    if (lane < NX) taua()[lane] = X[(size_t)T * NX + lane] - xr[(size_t)T * NX + lane];
    __syncwarp();
    cost_terminal<1>(Cfq, cfq, qN, lN);
    const R q = warp_sum(quad[0]), l = warp_sum(lin[0]), q2 = warp_sum(qN[0]), l2 = warp_sum(lN[0]);
    __syncwarp();
    return ((R)0.5 * q + l) + ((R)0.5 * q2 + l2);
  }
};

template <int NX, int NU, typename R> constexpr int lti_warps() {
  return sizeof(R) == 8 ? 6 : (NX == 32 ? 14 : 16);
}

template <typename R, int NX, int NU>
TACD_DEV void lti_stage_F(const DevProblem<R>& P, R* sF) {
  constexpr int N = NX + NU;
  for (int i = threadIdx.x; i < NX * N; i += blockDim.x) {
    const int r = i / N, k = i % N;
    const R v = (k < NX) ? P.dyn_A[r * NX + k] : P.dyn_B[r * NU + (k - NX)];
    sF[i] = v;
    sF[NX * N + k * NX + r] = v;
  }
  __syncthreads();
}

// ---------------------------------------------------------------------------------------------------------------------------
// forward solve (solve_step, controller.py:275-315)
template <typename R, int NX, int NU>
__global__ void __launch_bounds__(lti_warps<NX, NU, R>() * 32, 1) ilqr_forward_lti(const DevProblem<R> P, const FwdPtrs<R> io, R* ws, unsigned int* counter) {
  using W = LtiWarp<R, NX, NU>;
  using C = LtiCfg<NX, NU>;
  constexpr int N = NX + NU;
  constexpr unsigned FULL = 0xffffffffu;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  R* sF = reinterpret_cast<R*>(smem_raw);
  lti_stage_F<R, NX, NU>(P, sF);
  const int T = P.T, lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  W w;
  w.init(sF, sF + C::F_WORDS + (size_t)wid * C::WARP_WORDS);
  const size_t szX = (size_t)(T + 1) * NX, szU = (size_t)T * NU;
  R* slice = ws + ((size_t)blockIdx.x * nwarp + wid) * lti_slice_words(T, NX, NU);
  R* Xb[2] = {slice, slice + szX + szU};
  R* Ub[2] = {slice + szX, slice + 2 * szX + szU};
  R* gK = slice + 2 * (szX + szU);
  R* gk = gK + (size_t)T * NU * NX;
  const bool lssep = P.ls_cost_shared != 0;
  const bool replay = io.replay_iters != nullptr;
  R alphas[kLtiNA];
#pragma unroll
  for (int a = 0; a < kLtiNA; ++a) alphas[a] = P.alphas[a < P.n_alpha ? a : P.n_alpha - 1];

  while (true) {
    unsigned nb = 0;
    if (lane == 0) nb = atomicAdd(counter, 1u);
    nb = __shfl_sync(FULL, nb, 0);
    if (nb >= (unsigned)P.B) break;
    const int b = (int)nb;
    const R* Cp = io.C + (P.C_batched ? (size_t)b * T * N * N : 0);
    const R* cp = io.c + (P.C_batched ? (size_t)b * T * N : 0);
    const R* Cfp = io.Cf + (P.Cf_batched ? (size_t)b * N * N : 0);
    const R* cfp = io.cf + (P.Cf_batched ? (size_t)b * N : 0);
    const R* Cl = lssep ? io.C_ls : Cp;
    const R* cl = lssep ? io.c_ls : cp;
    const R* Cfl = lssep ? io.Cf_ls : Cfp;
    const R* cfl = lssep ? io.cf_ls : cfp;
    const R* xr = io.xref + (P.xref_batched ? (size_t)b * szX : 0);
    const R* ur = io.uref + (P.uref_batched ? (size_t)b * szU : 0);
    int cur = 0;
    // ---- init: U = U_init (not clamped, quirk Q12), X = rollout, best = objective
    for (int i = lane; i < (int)szU; i += 32) Ub[0][i] = io.Uinit[(size_t)b * szU + i];
    if (io.alpha_trace) for (int j = lane; j < P.max_iter; j += 32) io.alpha_trace[(size_t)b * P.max_iter + j] = 0xFF;
    {
      R* z = w.za(0);
      if (lane < NX) { const R x0 = io.x0[(size_t)b * NX + lane]; z[lane] = x0; Xb[0][lane] = x0; }
      __syncwarp();
      for (int t = 0; t < T; ++t) {
        if (lane < NU) z[NX + lane] = Ub[0][(size_t)t * NU + lane];
        __syncwarp();
        R sx = 0, su = 0;
        for (int k = 0; k < NX; ++k) sx += w.sFT[k * NX + w.li_()] * z[k];
        for (int k = 0; k < NU; ++k) su += w.sFT[(NX + k) * NX + w.li_()] * z[NX + k];
        __syncwarp();
        if (lane < NX) { z[lane] = sx + su; Xb[0][(size_t)(t + 1) * NX + lane] = sx + su; }
        __syncwarp();
      }
    }
    R best = w.traj_cost(T, Xb[0], Ub[0], xr, ur, Cp, cp, Cfp, cfp);
    unsigned flg = 0;
    int it = 0;
    while (it < P.max_iter && !(replay && it >= io.replay_iters[b])) {
      const R* Xn = Xb[cur];
      const R* Un = Ub[cur];
      // ---- reverse Riccati pass: quadraticize (cost.py:64-100) + backward_lqr (controller.py:546-578)
      {
        if (lane < NX) w.tau()[lane] = Xn[(size_t)T * NX + lane] - xr[(size_t)T * NX + lane];
        for (int idx = lane; idx < NX * NX; idx += 32) w.Vs()[(idx / NX) * C::LDV + (idx % NX)] = Cfp[(size_t)(idx / NX) * N + (idx % NX)];
        __syncwarp();
        if (lane < NX) {
          R a = 0;
          for (int j = 0; j < NX; ++j) a += Cfp[(size_t)lane * N + j] * w.tau()[j];
          w.vs()[lane] = a + cfp[lane];
        }
        __syncwarp();
        typename W::StepIn si, sn;
        if (W::PF) w.template load_step<true>(si, T - 1, Cp, cp, Xn, Un, xr, ur);
        for (int t = T - 1; t >= 0; --t) {
          if (!W::PF) w.template load_step<true>(si, t, Cp, cp, Xn, Un, xr, ur);
          else if (t > 0) w.template load_step<true>(sn, t - 1, Cp, cp, Xn, Un, xr, ur);
          w.store_tau(si);
          __syncwarp();
          w.cost_linear(si);
          __syncwarp();
          if (w.template value_step<false>(P, si.c, gK + (size_t)t * NU * NX, gk + (size_t)t * NU)) flg |= TACD_FLAG_NONPD;
          if (W::PF && t > 0) si = sn;
        }
      }
      // ---- all line-search candidates (evaluate_alphas, controller.py:589-620)
      R costs[kLtiNA];
      w.template rollout<kLtiNA, false, true>(P, T, alphas, Xn, Un, gK, gk, xr, ur, Cl, cl, Cfl, cfl, nullptr, nullptr, costs);
      int ai = 0;
      {
        R cbest = costs[0];
        bool done = false;
#pragma unroll
        for (int a = 1; a < kLtiNA; ++a) {
          if (a < P.n_alpha && !done) {
            if (cbest != cbest) done = true;
            else if (costs[a] < cbest || costs[a] != costs[a]) { ai = a; cbest = costs[a]; }
          }
        }
      }
      if (replay) ai = io.replay_alpha[(size_t)b * P.max_iter + it];
      R al1[1] = {lti_pick<kLtiNA, R>(alphas, ai)};
      R newc = lti_pick<kLtiNA, R>(costs, ai);
      bool rolled = false;
      if (lssep) {   // acceptance cost = the element's OWN cost of the chosen candidate (controller.py:298)
        R c1[1];
        w.template rollout<1, true, true>(P, T, al1, Xn, Un, gK, gk, xr, ur, Cp, cp, Cfp, cfp, Xb[cur ^ 1], Ub[cur ^ 1], c1);
        newc = c1[0];
        rolled = true;
      }
      bool improved = newc < best;
      if (replay) improved = (it < io.replay_iters[b] - 1) || ((io.replay_flags[b] & TACD_FLAG_MAXITER) != 0);
      if (lane == 0) {
        if (io.alpha_trace) io.alpha_trace[(size_t)b * P.max_iter + it] = (uint8_t)ai;
        if (io.cost_trace) {
          R* ctr = io.cost_trace + ((size_t)b * P.max_iter + it) * (P.n_alpha + 2);
#pragma unroll
          for (int a = 0; a < kLtiNA; ++a)
            if (a < P.n_alpha) ctr[a] = costs[a];
          ctr[P.n_alpha] = newc;
          ctr[P.n_alpha + 1] = best;
        }
      }
      it += 1;
      if (!improved) break;
      best = newc;
      if (!rolled) {
        R c1[1];
        w.template rollout<1, true, false>(P, T, al1, Xn, Un, gK, gk, xr, ur, Cp, cp, Cfp, cfp, Xb[cur ^ 1], Ub[cur ^ 1], c1);
      }
      cur ^= 1;
      if (it >= P.max_iter) flg |= TACD_FLAG_MAXITER;
    }
    // ---- finalise: outputs, active set (controller.py:636-642), A/B of the final trajectory (LTI: the matrices themselves)
    __syncwarp();
    for (int i = lane; i < (int)szX; i += 32) io.X[(size_t)b * szX + i] = Xb[cur][i];
    for (int i = lane; i < (int)szU; i += 32) {
      const R ui = Ub[cur][i];
      io.U[(size_t)b * szU + i] = ui;
      bool mk = false;
      if (P.has_umin) mk = mk || isclose_<R>(ui, P.umin[i % NU]);
      if (P.has_umax) mk = mk || isclose_<R>(ui, P.umax[i % NU]);
      io.tight[(size_t)b * szU + i] = mk ? 1 : 0;
    }
    if (io.A) for (int i = lane; i < T * NX * NX; i += 32) io.A[(size_t)b * T * NX * NX + i] = P.dyn_A[i % (NX * NX)];
    if (io.Bm) for (int i = lane; i < T * NX * NU; i += 32) io.Bm[(size_t)b * T * NX * NU + i] = P.dyn_B[i % (NX * NU)];
    if (lane == 0) {
      io.iters[b] = it;
      io.flags[b] = (uint8_t)flg;
      if (io.cost) io.cost[b] = best;
    }
    __syncwarp();
  }
}

// ---------------------------------------------------------------------------------------------------------------------------
// backward (ILQRSolve.backward + _zero_constrained_lqr without input bounds, controller.py:63-115, 707-788)
template <typename R, int NX, int NU>
__global__ void __launch_bounds__(lti_warps<NX, NU, R>() * 32, 1) ilqr_backward_lti(const DevProblem<R> P, const BwdPtrs<R> io, R* ws, unsigned int* counter) {
  using W = LtiWarp<R, NX, NU>;
  using C = LtiCfg<NX, NU>;
  constexpr int N = NX + NU;
  constexpr unsigned FULL = 0xffffffffu;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  R* sF = reinterpret_cast<R*>(smem_raw);
  lti_stage_F<R, NX, NU>(P, sF);
  const int T = P.T, lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  W w;
  w.init(sF, sF + C::F_WORDS + (size_t)wid * C::WARP_WORDS);
  const size_t szX = (size_t)(T + 1) * NX, szU = (size_t)T * NU;
  R* slice = ws + ((size_t)blockIdx.x * nwarp + wid) * lti_slice_words(T, NX, NU);
  R* gK = slice;
  R* gk = gK + (size_t)T * NU * NX;
  // forward-sweep buffers (the Riccati temporaries are dead by then): dtau [N], tau [N]
  R* dt_ = w.sm;
  R* tt_ = w.sm + N;
  const int jp = lane / NU;
  while (true) {
    unsigned nb = 0;
    if (lane == 0) nb = atomicAdd(counter, 1u);
    nb = __shfl_sync(FULL, nb, 0);
    if (nb >= (unsigned)P.B) break;
    const int b = (int)nb;
    const R* X = io.X + (size_t)b * szX;
    const R* U = io.U + (size_t)b * szU;
    const R* Cp = io.C + (P.C_batched ? (size_t)b * T * N * N : 0);
    const R* xr = io.xref + (P.xref_batched ? (size_t)b * szX : 0);
    const R* ur = io.uref + (P.uref_batched ? (size_t)b * szU : 0);
    const R* gX = io.gX + (size_t)b * szX;
    const R* gU = io.gU + (size_t)b * szU;
    // terminal value: H_xx[T-1], -grad_X[T-1] (quirks Q1, Q2)
    for (int idx = lane; idx < NX * NX; idx += 32)
      w.Vs()[(idx / NX) * C::LDV + (idx % NX)] = Cp[(size_t)(T - 1) * N * N + (size_t)(idx / NX) * N + (idx % NX)];
    if (lane < NX) w.vs()[lane] = -gX[(size_t)(T - 1) * NX + lane];
    __syncwarp();
    typename W::StepIn si, sn;
    if (W::PF) w.template load_step<false>(si, T - 1, Cp, nullptr, gX, gU, nullptr, nullptr);
    for (int t = T - 1; t >= 0; --t) {
      if (!W::PF) w.template load_step<false>(si, t, Cp, nullptr, gX, gU, nullptr, nullptr);
      else if (t > 0) w.template load_step<false>(sn, t - 1, Cp, nullptr, gX, gU, nullptr, nullptr);
      if (lane < N) w.lv()[w.e0] = -si.ta0;
      if (N > 32 && w.has1) w.lv()[w.e1] = -si.ta1;
      __syncwarp();
      w.template value_step<true>(P, si.c, gK + (size_t)t * NU * NX, gk + (size_t)t * NU);
      if (W::PF && t > 0) si = sn;
    }
    // ---- forward sweep dx+ = A dx + B du, du = K dx + k, and the gradients (controller.py:769-788, 100-115)
    if (io.grad_x0 && lane < NX) io.grad_x0[(size_t)b * NX + lane] = (R)0;   // quirk Q3
    if (lane < N) dt_[w.e0] = 0;
    if (N > 32 && w.has1) dt_[w.e1] = 0;
    __syncwarp();
    for (int t = 0; t <= T; ++t) {
      if (t < T) {
        R kr[W::UC], dv[W::UC];
        ldv<W::UC>(kr, gK + (size_t)t * NU * NX + w.m * NX + w.ju0);
        ldv<W::UC>(dv, dt_ + w.ju0);
        R s = 0;
#pragma unroll
        for (int e = 0; e < W::UC; ++e) s += kr[e] * dv[e];
#pragma unroll
        for (int o = NU; o < 32; o <<= 1) s += __shfl_xor_sync(FULL, s, o);
        if (jp == 0) dt_[NX + w.m] = s + gk[(size_t)t * NU + w.m];
      }
      if (lane < N) tt_[w.e0] = (w.e0 < NX) ? X[(size_t)t * NX + w.e0] - xr[(size_t)t * NX + w.e0]
                                             : (t < T ? U[(size_t)t * NU + w.e0 - NX] - ur[(size_t)t * NU + w.e0 - NX] : (R)0);
      if (N > 32 && w.has1) tt_[w.e1] = t < T ? U[(size_t)t * NU + w.e1 - NX] - ur[(size_t)t * NU + w.e1 - NX] : (R)0;
      __syncwarp();
      const R tj0 = tt_[w.e0], dj0 = dt_[w.e0], tj1 = tt_[w.e1], dj1 = dt_[w.e1];
      if (t < T) {
        if (io.grad_C) {
          R* gC = P.C_batched ? io.grad_C + ((size_t)b * T + t) * N * N : io.grad_C + (size_t)t * N * N;
          for (int i = 0; i < N; ++i) {
            const R ti = tt_[i], di = dt_[i];
            const R g0 = (R)-0.5 * (di * tj0 + ti * dj0);
            if (lane < N) { if (P.C_batched) gC[(size_t)i * N + w.e0] = g0; else atomicAdd(&gC[(size_t)i * N + w.e0], g0); }
            if (N > 32 && w.has1) {
              const R g1 = (R)-0.5 * (di * tj1 + ti * dj1);
              if (P.C_batched) gC[(size_t)i * N + w.e1] = g1; else atomicAdd(&gC[(size_t)i * N + w.e1], g1);
            }
          }
        }
        if (io.grad_c) {
          R* gc = P.C_batched ? io.grad_c + ((size_t)b * T + t) * N : io.grad_c + (size_t)t * N;
          if (lane < N) { if (P.C_batched) gc[w.e0] = -dj0; else atomicAdd(&gc[w.e0], -dj0); }
          if (N > 32 && w.has1) { if (P.C_batched) gc[w.e1] = -dj1; else atomicAdd(&gc[w.e1], -dj1); }
        }
        if (lane < NU) {
          const R du = dt_[NX + lane];
          if (io.grad_uref) { if (P.uref_batched) io.grad_uref[(size_t)b * szU + (size_t)t * NU + lane] = du; else atomicAdd(&io.grad_uref[(size_t)t * NU + lane], du); }
          if (io.dU) io.dU[(size_t)b * szU + (size_t)t * NU + lane] = du;
        }
      } else {
        if (io.grad_Cf) {
          for (int i = 0; i < N; ++i) {
            const R ti = tt_[i], di = dt_[i];
            const bool in0 = i < NX && w.e0 < NX;
            const R g0 = in0 ? (R)-0.5 * (di * tj0 + ti * dj0) : (R)0;
            if (lane < N) { if (P.Cf_batched) io.grad_Cf[(size_t)b * N * N + (size_t)i * N + w.e0] = g0; else if (in0) atomicAdd(&io.grad_Cf[(size_t)i * N + w.e0], g0); }
            if (N > 32 && w.has1 && P.Cf_batched) io.grad_Cf[(size_t)b * N * N + (size_t)i * N + w.e1] = (R)0;
          }
        }
        if (io.grad_cf) {
          const R g0 = (w.e0 < NX) ? -dj0 : (R)0;
          if (lane < N) { if (P.Cf_batched) io.grad_cf[(size_t)b * N + w.e0] = g0; else if (w.e0 < NX) atomicAdd(&io.grad_cf[w.e0], g0); }
          if (N > 32 && w.has1 && P.Cf_batched) io.grad_cf[(size_t)b * N + w.e1] = (R)0;
        }
      }
      if (lane < NX) {
        const R dx = dt_[lane];
        if (io.grad_xref) { if (P.xref_batched) io.grad_xref[(size_t)b * szX + (size_t)t * NX + lane] = dx; else atomicAdd(&io.grad_xref[(size_t)t * NX + lane], dx); }
        if (io.dX) io.dX[(size_t)b * szX + (size_t)t * NX + lane] = dx;
      }
      if (t < T) {
        R a = 0, a2 = 0;
        for (int k = 0; k < NX; ++k) a += w.sFT[k * NX + w.li_()] * dt_[k];
        for (int k = 0; k < NU; ++k) a2 += w.sFT[(NX + k) * NX + w.li_()] * dt_[NX + k];
        __syncwarp();
        if (lane < NX) dt_[lane] = a + a2;
      }
      __syncwarp();
    }
  }
}

}  // namespace tacd
