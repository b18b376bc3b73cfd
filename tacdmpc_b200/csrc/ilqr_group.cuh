// ilqr_group.cuh — whole-solve forward kernel for small systems, FIVE LANES PER BATCH ELEMENT.
//
// Why a second formulation (profiles/r1_forward_v3): with one thread per element a single
// iLQR iteration is one serial chain of ~26 k instructions (~35 us), the iteration count per element spans
// 2..19, and the kernel cannot end before its longest chain: 0.68 ms of the 0.85 ms forward was that tail,
// with 20 of 32 lanes active on average.  Here an element is owned by a group of G = 5 consecutive lanes
// (6 elements per warp, 2 lanes spare), which shortens the chain 5x and gives the work queue 6x more waves:
//
//   reverse pass   lane j owns ROW j of the (n x n) Q-function  Q = C_t + F' V F,  F = [A_t | B_t]  (n = 5 for
//                  both shipped non-linear plugins), and element j of q = C_t tau + c_t + F' v.  The gains
//                  are computed redundantly by every lane from 6..12 shuffled entries (Q_xu, Q_uu, q_u), the
//                  new value-function row i by lane i; V is all-gathered through 80 bytes of shared memory
//                  per element with 128-bit accesses.  Jacobians are evaluated five time steps at a time
//                  (lane j takes step t - j) in compact form (6 numbers for the cart-pole, 4 for the
//                  unicycle) and broadcast with shuffles; the structural zeros / ones of F are skipped.
//   line search    lane a rolls out candidate alpha_a.  Its INPUTS go to its own shared-memory buffer (6 input buffers per
//                  element: nominal + 5 candidates, accepting = buffer-index swap), its STATES go to a per-warp scratch
//                  tile in global memory, [word][lane] so that every store is one coalesced 128-byte line; the tile is
//                  rewritten every iteration and stays in L2.  Only the winner's states are read back (84 words per
//                  element and iteration), so shared memory holds ONE state trajectory per element instead of six:
//                  1.3 KB per element instead of 3.0 KB, 20 resident warps per SM instead of 12 (round 2; round 1's
//                  kernel issued 70 % of its slots with 3 warps per scheduler, profiles/r1_forward_group_v6).
//   queue          groups pull elements from a global counter; the initial rollout + cost of every element
//                  is done up front by ilqr_init_rollout (thread per element), so a refill is two coalesced
//                  copies and all six groups of a warp stay in the same phase.
//
// Shared costs (C, c, C_final, c_final not batched, and the line-search cost of quirk Q4) are staged once
// per CTA in shared memory.  Reference lines restated: controller.py:275-315 (solve_step), 546-578
// (backward_lqr), 589-620 (evaluate_alphas), 636-642 (tight mask); cost.py:37-100.
#pragma once
#include "ilqr_small.cuh"

namespace tacd {

constexpr int kG = 5;              // lanes per element = line-search candidates rolled out together
constexpr int kEPW = 32 / kG;      // elements per warp
constexpr int kNBUF = kG + 1;      // INPUT (U) buffers per element (nominal + one per candidate)
// Shared-memory layout of one warp (6 element slots), all [word][slot][element-in-warp] so that consecutive lanes of
// different groups hit consecutive banks:
//   X part   (T+1) nx words x {nominal X, x_ref}             stride kXS between words
//   U part   T nu words     x {6 input buffers, u_ref}       stride kUS between words
//   V part   6 x (nx+1) x 4 words: value-function exchange
constexpr int kXS = 2 * kEPW;
constexpr int kUS = (kNBUF + 1) * kEPW;
constexpr int kURef = kNBUF * kEPW;     // slot offset of u_ref in the U part
#ifndef TACD_GROUP_MAXW
// resident warps per SM (fp32) = __launch_bounds__(32 MAXW, 1).  Shared memory would allow 28; measured at B = 65536 (forward ms,
// registers): 12 -> 0.658 (148), 16 -> 0.626 (127), 20 -> 0.672 (96), 24 -> 0.660 (80): beyond 16 the register cap costs more
// instructions (+20 % at 80 registers: constants re-loaded, values re-materialised) than the extra warps hide latency.
#define TACD_GROUP_MAXW 16
#endif
constexpr int kGroupMaxWarps = TACD_GROUP_MAXW;
// resident warps per SM by scalar type: fp64 needs two registers per value (12 warps: 168 registers, no spills)
template <typename R> constexpr int group_max_warps() { return sizeof(R) == 4 ? kGroupMaxWarps : 12; }

// ---- structured Jacobians ---------------------------------------------------------------------------------
// F = [A | B] (NX x N) of the shipped plugins in compact form: jc[NJ] holds the entries that depend on (x,u);
// col() returns column j of F for a run-time j, mulF() accumulates w' F onto o skipping structural zeros.
// col() is written as chains of single selects on lane-invariant predicates: a nested ?: over j compiles to a
// divergent branch tree (profiles/r1_forward_group_v1: 60 of 280 instructions per step, 6..20 lanes active).
template <typename R, int NX, int NU, int DYN>
struct GDyn {  // LTI: dense F, shared by the CTA in shared memory (row-major NX x N)
  static constexpr int N = NX + NU;
  static constexpr int NJ = 1;
  static constexpr bool kHasJac = false;
  static TACD_DEV void jac(const DevProblem<R>&, const R (&)[NX], const R (&)[NU], R (&jc)[NJ]) { jc[0] = 0; }
  static TACD_DEV void col(const DevProblem<R>&, const R* sF, const R (&)[NJ], int j, R (&f)[NX]) {
#pragma unroll
    for (int i = 0; i < NX; ++i) f[i] = sF[i * N + j];
  }
  static TACD_DEV void mulF(const DevProblem<R>&, const R* sF, const R (&)[NJ], const R (&w)[NX], R (&o)[N]) {
#pragma unroll
    for (int k = 0; k < N; ++k) {
      R s = o[k];
#pragma unroll
      for (int l = 0; l < NX; ++l) s += w[l] * sF[l * N + k];
      o[k] = s;
    }
  }
  // o = F d  (d = [dx; du]): one step of the linearised dynamics
  static TACD_DEV void mulFv(const DevProblem<R>&, const R* sF, const R (&)[NJ], const R (&d)[N], R (&o)[NX]) {
#pragma unroll
    for (int i = 0; i < NX; ++i) {
      R s = 0;
#pragma unroll
      for (int k = 0; k < N; ++k) s += sF[i * N + k] * d[k];
      o[i] = s;
    }
  }
};

// cart-pole: A = [[1,dt,0,0],[0,1,a,b],[0,0,1,dt],[0,0,c,d]], B = [0,e,0,f]'  (tacd_dev.cuh::cartpole_jac)
template <typename R>
struct GDyn<R, 4, 1, TACD_DYN_CARTPOLE> {
  static constexpr int NJ = 6;
  static constexpr bool kHasJac = true;
  static TACD_DEV void jac(const DevProblem<R>& P, const R (&x)[4], const R (&u)[1], R (&jc)[6]) {
    R A[16], Bm[4];
    cartpole_jac<R>(P.dynp, P.dt, x, u, A, Bm);
    jc[0] = A[6]; jc[1] = A[7]; jc[2] = A[14]; jc[3] = A[15]; jc[4] = Bm[1]; jc[5] = Bm[3];
  }
  static TACD_DEV void col(const DevProblem<R>& P, const R*, const R (&jc)[6], int j, R (&f)[4]) {
    const R dt = P.dt;
    const bool p1 = j == 1, p2 = j == 2, p3 = j == 3, p4 = j == 4;
    R f0 = (j == 0) ? (R)1 : (R)0; f0 = p1 ? dt : f0;
    R f1 = p1 ? (R)1 : (R)0; f1 = p2 ? jc[0] : f1; f1 = p3 ? jc[1] : f1; f1 = p4 ? jc[4] : f1;
    R f2 = p2 ? (R)1 : (R)0; f2 = p3 ? dt : f2;
    R f3 = p2 ? jc[2] : (R)0; f3 = p3 ? jc[3] : f3; f3 = p4 ? jc[5] : f3;
    f[0] = f0; f[1] = f1; f[2] = f2; f[3] = f3;
  }
  static TACD_DEV void mulF(const DevProblem<R>& P, const R*, const R (&jc)[6], const R (&w)[4], R (&o)[5]) {
    const R dt = P.dt;
    o[0] += w[0];
    o[1] = (o[1] + w[0] * dt) + w[1];
    o[2] = ((o[2] + w[1] * jc[0]) + w[2]) + w[3] * jc[2];
    o[3] = ((o[3] + w[1] * jc[1]) + w[2] * dt) + w[3] * jc[3];
    o[4] = (o[4] + w[1] * jc[4]) + w[3] * jc[5];
  }
  static TACD_DEV void mulFv(const DevProblem<R>& P, const R*, const R (&jc)[6], const R (&d)[5], R (&o)[4]) {
    const R dt = P.dt;
    o[0] = d[0] + dt * d[1];
    o[1] = ((d[1] + jc[0] * d[2]) + jc[1] * d[3]) + jc[4] * d[4];
    o[2] = d[2] + dt * d[3];
    o[3] = (jc[2] * d[2] + jc[3] * d[3]) + jc[5] * d[4];
  }
};

// unicycle: A = [[1,0,a],[0,1,b],[0,0,1]], B = [[p,0],[q,0],[0,dt]]  (tacd_dev.cuh::unicycle_jac)
template <typename R, int DYN>
struct GDynUnicycle {
  static constexpr int NJ = 4;
  static constexpr bool kHasJac = true;
  static TACD_DEV void jac(const DevProblem<R>& P, const R (&x)[3], const R (&u)[2], R (&jc)[4]) {
    R A[9], Bm[6];
    unicycle_jac<R>(P.dt, x, u, A, Bm);
    jc[0] = A[2]; jc[1] = A[5]; jc[2] = Bm[0]; jc[3] = Bm[2];
  }
  static TACD_DEV void col(const DevProblem<R>& P, const R*, const R (&jc)[4], int j, R (&f)[3]) {
    const bool p2 = j == 2, p3 = j == 3;
    R f0 = (j == 0) ? (R)1 : (R)0; f0 = p2 ? jc[0] : f0; f0 = p3 ? jc[2] : f0;
    R f1 = (j == 1) ? (R)1 : (R)0; f1 = p2 ? jc[1] : f1; f1 = p3 ? jc[3] : f1;
    R f2 = p2 ? (R)1 : (R)0; f2 = (j == 4) ? P.dt : f2;
    f[0] = f0; f[1] = f1; f[2] = f2;
  }
  static TACD_DEV void mulF(const DevProblem<R>& P, const R*, const R (&jc)[4], const R (&w)[3], R (&o)[5]) {
    o[0] += w[0];
    o[1] += w[1];
    o[2] = ((o[2] + w[0] * jc[0]) + w[1] * jc[1]) + w[2];
    o[3] = (o[3] + w[0] * jc[2]) + w[1] * jc[3];
    o[4] += w[2] * P.dt;
  }
  static TACD_DEV void mulFv(const DevProblem<R>& P, const R*, const R (&jc)[4], const R (&d)[5], R (&o)[3]) {
    o[0] = (d[0] + jc[0] * d[2]) + jc[2] * d[3];
    o[1] = (d[1] + jc[1] * d[2]) + jc[3] * d[3];
    o[2] = d[2] + P.dt * d[4];
  }
};
template <typename R> struct GDyn<R, 3, 2, TACD_DYN_UNICYCLE> : GDynUnicycle<R, TACD_DYN_UNICYCLE> {};
template <typename R> struct GDyn<R, 3, 2, TACD_DYN_UNICYCLE_WRAPPED> : GDynUnicycle<R, TACD_DYN_UNICYCLE_WRAPPED> {};

// a[i] for a run-time i < M with a in registers
template <int M, typename R>
TACD_DEV R pick(const R* a, int i) {
  R r = a[0];
#pragma unroll
  for (int k = 1; k < M; ++k) r = (i == k) ? a[k] : r;
  return r;
}

template <typename R> struct alignas(16) Row4 { R v[4]; };

// shared-memory words of one warp's region (6 element slots): offsets of the U part and of the value-function exchange
// block (16-byte aligned), size of the whole region
__host__ __device__ inline int group_uoff_words(int T, int nx) { return (T + 1) * nx * kXS; }
__host__ __device__ inline int group_voff_words(int T, int nx, int nu) {
  return (group_uoff_words(T, nx) + T * nu * kUS + 3) & ~3;
}
// + (stage > 0) the per-element cost blocks [C_b (T cw) | c_b (T n)] of the 6 element slots and their 6 mbarriers
__host__ __device__ inline int group_coff_words(int T, int nx, int nu) { return group_voff_words(T, nx, nu) + kEPW * (nx + 1) * 4; }
__host__ __device__ inline int group_warp_words(int T, int nx, int nu, int stage = 0) {
  return group_coff_words(T, nx, nu) + (stage > 0 ? kEPW * stage + 16 : 0);   // 16 words: 6 x 8-byte mbarriers, 16-byte aligned
}
// words of one element's staged cost block (multiple of 4), or 0 if the blocks cannot be bulk-copied: cp.async.bulk needs
// 16-byte aligned addresses and sizes, i.e. T cw and T n multiples of 16 bytes (cart-pole T = 20: 2000 + 400 bytes)
__host__ __device__ inline int group_stage_words(int T, int n, int cw, size_t es) {
  const size_t bc = (size_t)T * cw * es, bl = (size_t)T * n * es;
  if (bc % 16 != 0 || bl % 16 != 0) return 0;
  return (int)((bc + bl) / es) + ((cw + n + 3) & ~3);   // + the element's terminal cost [C_final | c_final] (plain loads)
}
// global-memory scratch of one warp: the candidates' states X_1..X_T, [word][lane]; TWO tiles, alternating with the iteration
// parity, so that the nominal states (the winner of the previous iteration) can be restored after a rejected step
__host__ __device__ inline size_t group_xtile_words(int T, int nx) { return (size_t)T * nx * 32; }
__host__ __device__ inline size_t group_xscratch_words(int T, int nx) { return 2 * group_xtile_words(T, nx); }
template <typename R> TACD_DEV R ld_cg(const R* p) { return __ldcg(p); }
// a word of the candidates' state tile: the tile is global memory (L2) or, for small batches, shared memory (generic address)
template <typename R> TACD_DEV R ld_scr(const R* p, bool in_smem) { return in_smem ? *p : __ldcg(p); }
// tiles per warp when the scratch lives in shared memory (the second tile only serves the restore of quirk Q4's layouts)
__host__ __device__ inline size_t group_xscratch_smem_words(int T, int nx, bool lssep) { return (lssep ? 2 : 1) * group_xtile_words(T, nx); }
// CTA-shared staging of the un-batched costs (+ the separate line-search cost) and the LTI matrices.
// Running cost of step t: [C_t (n*n) | c_t (n) | pad] at stride group_cost_stride(n) (multiple of 4 words, so the
// line search reads it with 128-bit loads); terminal cost: [C_final | c_final | pad].
__host__ __device__ inline int group_cost_stride(int n) { return (n * n + n + 3) & ~3; }
__host__ __device__ inline int group_cost_words(int T, int nx, int nu, int C_batched, int Cf_batched, int lssep, int lti) {
  const int n = nx + nu, run = T * group_cost_stride(n), fin = group_cost_stride(n);
  int w = 0;
  if (!C_batched) w += run;
  if (!Cf_batched) w += fin;
  if (lssep) w += run + fin;
  if (lti) w += (nx * n + 3) & ~3;
  return w;
}

// Diagonal cost: quad += sum_i tau_i (q_i tau_i), lin += sum_i c_i tau_i, bit-identical to quad_terms<> on the
// expanded matrix for finite tau (the skipped terms are exact zeros).  `chk` restores the one visible
// difference: the dense form multiplies every tau_k by a zero of C, so a non-finite tau_k makes the
// cost NaN (and NaN wins the argmin, controller.py:297); chk = sum_k 0 * tau_k is NaN exactly then.
template <typename R, int N>
TACD_DEV void quad_terms_diag(const R (&q)[N], const R (&ct)[N], const R (&tau)[N], R& quad, R& lin) {
  R chk = 0;
#pragma unroll
  for (int i = 0; i < N; ++i) chk += tau[i] * (R)0;
#pragma unroll
  for (int i = 0; i < N; ++i) {
    const R s = q[i] * tau[i] + chk;
    quad += tau[i] * s;
    lin += ct[i] * tau[i];
  }
}

// [C_t | c_t] of one step from a 16-byte aligned staged block
template <typename R, int N>
TACD_DEV void load_cost_vec(const R* p, R (&Ct)[N * N], R (&ct)[N]) {
  constexpr int NW = N * N + N, NV = (NW + 3) / 4;
  R tmp[NV * 4];
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    const Row4<R> r = *reinterpret_cast<const Row4<R>*>(p + 4 * v);
#pragma unroll
    for (int c = 0; c < 4; ++c) tmp[4 * v + c] = r.v[c];
  }
#pragma unroll
  for (int i = 0; i < N * N; ++i) Ct[i] = tmp[i];
#pragma unroll
  for (int i = 0; i < N; ++i) ct[i] = tmp[N * N + i];
}

// Initial rollout + objective of every element (controller.py:277-279, 361-398; cost.py:37-62), one thread
// per element: X_init -> io.X (re-used as the hand-over buffer), cost -> cost0.
// A warp's 32 elements are contiguous in every [B, ...] tensor, so x_ref / u_ref / U_init come in with
// cp.async as coalesced row copies into a shared-memory tile with an odd row stride (thread e then walks row
// e conflict-free), X_t is written over the slot of x_ref[t] once that has been read and leaves as
// coalesced rows; per-element costs come in the same way, `chunk` time steps at a time: the whole horizon
// in one piece when the tile fits (one DRAM latency for everything), else double-buffered one chunk ahead.
// (U_init is read straight from global memory, one step ahead: with it in the tile a cart-pole T=20 CTA needs 17 KB, 13
// fit an SM, and the 2048 tiles of B = 65 536 take 1.06 waves — the kernel lasted two passes, 45 us; without it 16 fit)
__host__ __device__ inline int init_in_words(int T, int nx, int nu) { return ((T + 1) * nx + T * nu) | 1; }
// row stride of a cost tile: chunk steps of [C_t (cw words: n*n dense, n compact diagonal)] then chunk steps of c_t
__host__ __device__ inline int init_cost_words(int chunk, int nx, int nu, int C_diag) {
  const int n = nx + nu;
  return (chunk * ((C_diag ? n : n * n) + n)) | 1;
}
__host__ __device__ inline size_t init_warp_words(int T, int nx, int nu, int C_batched, int C_diag, int chunk) {
  return (size_t)32 * init_in_words(T, nx, nu) +
         (C_batched && chunk > 0 ? (size_t)(chunk >= T ? 32 : 64) * init_cost_words(chunk, nx, nu, C_diag) : 0);
}

template <typename R, int NX, int NU, int DYN>
__global__ void __launch_bounds__(128) ilqr_init_rollout(const DevProblem<R> P, const FwdPtrs<R> io, R* cost0, int chunk) {
  constexpr int N = NX + NU;
  using D = Dyn<R, NX, NU, DYN>;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ R sAB[NX * NX + NX * NU];
  R* sA = sAB;
  R* sB = sAB + NX * NX;
  if (DYN == TACD_DYN_LTI) {
    for (int i = threadIdx.x; i < NX * NX; i += blockDim.x) sA[i] = P.dyn_A[i];
    for (int i = threadIdx.x; i < NX * NU; i += blockDim.x) sB[i] = P.dyn_B[i];
    __syncthreads();
  }
  const int T = P.T, lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int XW = (T + 1) * NX, UW = T * NU, RSI = init_in_words(T, NX, NU), RSC = init_cost_words(chunk, NX, NU, P.C_diag);
  const int cw = P.C_diag ? N : N * N;   // words of C per step (compact diagonal or dense)
  // row of one element: [x_ref -> X (XW) | u_ref (UW)]
  R* tile = reinterpret_cast<R*>(smem_raw) + (size_t)wid * init_warp_words(T, NX, NU, P.C_batched, P.C_diag, chunk);
  R* ctile = tile + (size_t)32 * RSI;   // [1 or 2][32][RSC]: [C_t0.. (chunk*cw) | c_t0.. (chunk*N)]
  const int nbuf = chunk >= T ? 1 : 2;
  const int b0 = (blockIdx.x * (blockDim.x >> 5) + wid) * 32;
  if (b0 >= P.B) return;
  const int nlive = P.B - b0 < 32 ? P.B - b0 : 32;
  // chunk == 0: per-element costs are read straight from global memory, each thread its own element's 2n or n*n + n
  // consecutive words per step (every byte of the touched sectors is used and stays in L1 for the next loads); no
  // cost tile, so the CTA keeps the 13 KB footprint of the shared-cost case
  const bool direct = chunk == 0;
  if (direct) chunk = T;
  const int nchunk = (T + chunk - 1) / chunk;
  const bool want_cost = cost0 != nullptr;   // nullptr: roll out only (the solve kernel evaluates the initial cost itself)
  auto fetch_cost = [&](int c) {
    if (want_cost && P.C_batched && !direct && c < nchunk) {
      const int t0 = c * chunk, nst = (T - t0 < chunk) ? T - t0 : chunk;
      R* buf = ctile + (size_t)(c & (nbuf - 1)) * 32 * RSC;
      for (int r = 0; r < nlive; ++r) {
        const R* Cg = io.C + ((size_t)(b0 + r) * T + t0) * cw;
        const R* cg = io.c + ((size_t)(b0 + r) * T + t0) * N;
        for (int w = lane; w < nst * cw; w += 32) cp_async<R>(&buf[(size_t)r * RSC + w], &Cg[w]);
        for (int w = lane; w < nst * N; w += 32) cp_async<R>(&buf[(size_t)r * RSC + chunk * cw + w], &cg[w]);
      }
    }
    cp_async_commit();
  };
  for (int r = 0; r < nlive; ++r) {
    R* row = tile + (size_t)r * RSI;
    const R* xr = io.xref + (P.xref_batched ? (size_t)(b0 + r) * XW : 0);
    const R* ur = io.uref + (P.uref_batched ? (size_t)(b0 + r) * UW : 0);
    for (int w = lane; w < XW; w += 32) cp_async<R>(&row[w], &xr[w]);
    for (int w = lane; w < UW; w += 32) cp_async<R>(&row[XW + w], &ur[w]);
  }
  fetch_cost(0);   // group 0: inputs + first cost chunk
  fetch_cost(1);
  const int b = b0 + (lane < nlive ? lane : nlive - 1);   // dead lanes shadow the last element, outputs masked
  const R* Cp = io.C;    // shared cost: uniform loads
  const R* cp = io.c;
  const R* Cfp = io.Cf + (P.Cf_batched ? (size_t)b * (P.C_diag ? N : N * N) : 0);
  const R* cfp = io.cf + (P.Cf_batched ? (size_t)b * N : 0);
  R* row = tile + (size_t)(lane < nlive ? lane : nlive - 1) * RSI;
  const R* ur = row + XW;
  const R* ui = io.Uinit + (size_t)b * UW;
  R x[NX], u[NU], un[NU], xn[NX];
  R quad = 0, lin = 0;
#pragma unroll
  for (int i = 0; i < NX; ++i) x[i] = io.x0[(size_t)b * NX + i];
#pragma unroll
  for (int i = 0; i < NU; ++i) un[i] = ui[i];
  for (int c = 0; c < nchunk; ++c) {
    cp_async_wait_group<1>();
    __syncwarp();
    const int t0 = c * chunk, t1 = (t0 + chunk < T) ? t0 + chunk : T;
    const R* cbuf = ctile + ((size_t)(c & (nbuf - 1)) * 32 + (lane < nlive ? lane : nlive - 1)) * RSC;
    for (int t = t0; t < t1; ++t) {
      R tau[N], Ct[N * N], ct[N];
#pragma unroll
      for (int i = 0; i < NU; ++i) { u[i] = un[i]; un[i] = ui[(t + 1 < T ? t + 1 : t) * NU + i]; }
#pragma unroll
      for (int i = 0; i < NX; ++i) { tau[i] = x[i] - row[t * NX + i]; row[t * NX + i] = x[i]; }
#pragma unroll
      for (int i = 0; i < NU; ++i) tau[NX + i] = u[i] - ur[t * NU + i];
      if (!want_cost) {
#pragma unroll
        for (int i = 0; i < N * N; ++i) Ct[i] = 0;
#pragma unroll
        for (int i = 0; i < N; ++i) ct[i] = 0;
      } else if (direct) {
        const R* Cg = io.C + ((size_t)b * T + t) * cw;
        const R* cg = io.c + ((size_t)b * T + t) * N;
        if (P.C_diag) {
#pragma unroll
          for (int i = 0; i < N * N; ++i) Ct[i] = (i / N == i % N) ? Cg[i / N] : (R)0;
        } else {
#pragma unroll
          for (int i = 0; i < N * N; ++i) Ct[i] = Cg[i];
        }
#pragma unroll
        for (int i = 0; i < N; ++i) ct[i] = cg[i];
      } else if (P.C_diag) {   // expanded to dense: same arithmetic as the dense layout (diag_embed, actor.py:74)
#pragma unroll
        for (int i = 0; i < N * N; ++i) Ct[i] = (i / N == i % N) ? cbuf[(t - t0) * N + i / N] : (R)0;
#pragma unroll
        for (int i = 0; i < N; ++i) ct[i] = cbuf[chunk * cw + (t - t0) * N + i];
      } else if (P.C_batched) {
#pragma unroll
        for (int i = 0; i < N * N; ++i) Ct[i] = cbuf[(t - t0) * N * N + i];
#pragma unroll
        for (int i = 0; i < N; ++i) ct[i] = cbuf[chunk * cw + (t - t0) * N + i];
      } else {
#pragma unroll
        for (int i = 0; i < N * N; ++i) Ct[i] = Cp[(size_t)t * N * N + i];
#pragma unroll
        for (int i = 0; i < N; ++i) ct[i] = cp[(size_t)t * N + i];
      }
      if (want_cost) quad_terms<R, N>(Ct, ct, tau, quad, lin);
      D::step(P, sA, sB, x, u, xn);
#pragma unroll
      for (int i = 0; i < NX; ++i) x[i] = xn[i];
    }
    __syncwarp();
    fetch_cost(c + 2);
  }
  R qN = 0, lN = 0;
  {
    R tauN[NX];
#pragma unroll
    for (int i = 0; i < NX; ++i) { tauN[i] = x[i] - row[T * NX + i]; row[T * NX + i] = x[i]; }
#pragma unroll
    for (int i = 0; i < NX; ++i) {
      R s = 0;
#pragma unroll
      for (int j = 0; j < NX; ++j) s += want_cost ? (P.C_diag ? ((i == j) ? Cfp[i] : (R)0) : Cfp[i * N + j]) * tauN[j] : (R)0;
      qN += tauN[i] * s;
      lN += want_cost ? cfp[i] * tauN[i] : (R)0;
    }
  }
  if (want_cost && lane < nlive) cost0[b] = ((R)0.5 * quad + lin) + ((R)0.5 * qN + lN);
  cp_async_wait_all();
  __syncwarp();
  for (int r = 0; r < nlive; ++r) {
    const R* Xs = tile + (size_t)r * RSI;
    R* Xo = io.X + (size_t)(b0 + r) * XW;
    for (int w = lane; w < XW; w += 32) Xo[w] = Xs[w];
  }
}

// CSH: the element's own running cost (C, c) is shared by the batch and staged in shared memory (layout 2a);
// otherwise it is per-element and read from global memory.  LSSEP: candidates are scored with the separate
// shared line-search cost (quirk Q4), always staged.
// DIAG (only with per-element costs, CSH = false): tacd_problem.C_diag — C, C_final (and the line-search cost)
// are compact diagonals, the ActorMPC layout (ACMPC/actor.py:58-81); 2n instead of n*n + n words per step.
// Cost mode of a shared running cost (C, c not batched), decided ON THE DEVICE by one small launch and compiled into the solve
// kernel: 0 = general, 1 = time-invariant (C_t, c_t equal for all t: both passes keep the cost in registers), 2 = time-invariant
// and diagonal (every off-diagonal entry an exact zero — Q, R = diag: the line search evaluates 5 instead of 25 products,
// quad_terms_diag, bit-identical).  The launcher enqueues all three variants; each reads the mode word and the two that do not
// match exit at once (~2 us each).  Why compiled in: every RUN-TIME form of these two tests lost — predicated-off loads that
// still take issue slots, a second code path that costs registers — while a build with the flags forced ran the
// forward 5 % (time-invariant) faster.
template <typename R>
__global__ void group_cost_mode(const R* C, const R* c, int T, int n, int* mode) {
  __shared__ int s_same, s_diag;
  if (threadIdx.x == 0) { s_same = 1; s_diag = 1; }
  __syncthreads();
  bool same = true, dg = true;
  for (int i = threadIdx.x; i < T * n * n; i += blockDim.x) {
    const int k = i % (n * n);
    const R v = C[i];
    same = same && (v == C[k]);
    if (k / n != k % n) dg = dg && (v == (R)0);
  }
  for (int i = threadIdx.x; i < T * n; i += blockDim.x) same = same && (c[i] == c[i % n]);
  if (!same) atomicAnd(&s_same, 0);
  if (!dg) atomicAnd(&s_diag, 0);
  __syncthreads();
  if (threadIdx.x == 0) *mode = s_same ? (s_diag ? 2 : 1) : 0;
}

// SSM: the candidates' state tiles live in shared memory (small batches) instead of the L2-resident workspace.  A template
// parameter, not a run-time flag: with `ssm = io.xscratch == nullptr` the tile stores of the line search became generic stores
// and the read-back a select of two loads, and the B = 65536 forward paid 3 % (shared cost) to 13 % (compact diagonal) for a
// feature it does not use (measured with the flag forced at compile time: 0.593 -> 0.572 ms, 0.868 -> 0.752 ms).
template <typename R, int NX, int NU, int DYN, bool LSSEP, bool CSH, bool DIAG, int CM = 0, bool SSM = false>
__global__ void __launch_bounds__(group_max_warps<R>() * 32, 1) ilqr_forward_group(const DevProblem<R> P, const FwdPtrs<R> io, const R* cost0) {
  static_assert(CM == 0 || (CSH && !LSSEP && !DIAG), "cost modes exist for a plain shared cost only");
  if (io.cost_mode != nullptr && *io.cost_mode != CM) return;   // another variant of this launch group runs the batch
  constexpr int N = NX + NU;
  static_assert(N <= kG, "one lane per row of the Q-function");
  static_assert(NX <= 4, "value-function rows travel as one 4-word vector");
  static_assert(!(DIAG && CSH), "compact diagonal costs are per-element");
  using D = Dyn<R, NX, NU, DYN>;
  using GD = GDyn<R, NX, NU, DYN>;
  constexpr int CW = DIAG ? N : N * N;         // words of C per step / of C_final in global memory
  constexpr int NJ = GD::NJ;
  constexpr unsigned FULL = 0xffffffffu;
  constexpr int CS = (N * N + N + 3) & ~3;     // staged cost stride per step
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int T = P.T;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const bool valid = lane < kEPW * kG;         // lanes 30, 31 are spare: they run along, own nothing
  const int e = valid ? lane / kG : 0;
  const int j = valid ? lane - e * kG : kG;    // position in the group
  const int base = e * kG;                     // first lane of the group (shuffle source)
  const int jj = j < N ? j : N - 1;            // row of the Q-function this lane computes
  const int XW = (T + 1) * NX, UW = T * NU;
  const size_t szX = (size_t)XW, szU = (size_t)T * NU;

  // ---- shared memory: CTA-wide cost staging, then one region per warp
  R* sm = reinterpret_cast<R*>(smem_raw);
  const bool ownCf_sh = !P.Cf_batched;
  R *sC = sm, *sCf = sm, *sCl = sm, *sCfl = sm, *sF = sm;
  bool same_own = true, same_ls = true;
  {
    int o = 0;
    const int NT = blockDim.x;
    // returns false if this thread saw a step whose cost differs from step 0's
    auto stage_run = [&](R* dst, const R* Cg, const R* cg) {
      bool same = true;
      for (int i = tid; i < T * CS; i += NT) {
        const int t = i / CS, k = i - t * CS;
        // (a compact diagonal source is expanded to the dense staged layout)
        const int kc = DIAG ? ((k < N * N && k / N == k % N) ? k / N : -1) : k;
        const R val = (k < N * N) ? (kc >= 0 ? Cg[t * CW + kc] : (R)0) : (k < N * N + N) ? cg[t * N + (k - N * N)] : (R)0;
        const R v0 = (k < N * N) ? (kc >= 0 ? Cg[kc] : (R)0) : (k < N * N + N) ? cg[k - N * N] : (R)0;
        same = same && (val == v0);
        dst[i] = val;
      }
      return same;
    };
    auto stage_fin = [&](R* dst, const R* Cg, const R* cg) {
      for (int k = tid; k < CS; k += NT) {
        const int kc = DIAG ? ((k < N * N && k / N == k % N) ? k / N : -1) : k;
        dst[k] = (k < N * N) ? (kc >= 0 ? Cg[kc] : (R)0) : (k < N * N + N) ? cg[k - N * N] : (R)0;
      }
    };
    if (CSH) { sC = sm + o; o += T * CS; same_own = stage_run(sC, io.C, io.c); }
    if (ownCf_sh) { sCf = sm + o; o += CS; stage_fin(sCf, io.Cf, io.cf); }
    if (LSSEP) {
      sCl = sm + o; o += T * CS; same_ls = stage_run(sCl, io.C_ls, io.c_ls);
      sCfl = sm + o; o += CS; stage_fin(sCfl, io.Cf_ls, io.cf_ls);
    }
    if (DYN == TACD_DYN_LTI) {
      sF = sm + o; o += (NX * N + 3) & ~3;
      for (int i = tid; i < NX * N; i += NT) {
        const int r = i / N, k = i % N;
        sF[i] = (k < NX) ? P.dyn_A[r * NX + k] : P.dyn_B[r * NU + (k - NX)];
      }
    }
    sm += o;
  }
  // Time-invariant running cost (C_t, c_t equal for all t — the usual regulation / tracking cost): the line
  // search keeps it in registers for the whole pass instead of re-reading 30 words per step and lane.  The
  // shared-memory / L1 data path (128 B/clk per SM, also carrying shuffles and global loads) is what bounds
  // this kernel (profiles/r1_forward_group_v2: 115 data-path cycles per warp-step against 380 issue slots
  // shared by four schedulers), and those reads were a third of its traffic.
  const bool constOwn = CM >= 1 ? true : (__syncthreads_and(same_own ? 1 : 0) != 0 && CSH);
  const bool constLs = __syncthreads_and(same_ls ? 1 : 0) != 0 && LSSEP;
  // Trip counts (0 or 1) of the per-step cost loads, made opaque with a shuffle: written as `if (!constOwn) load`, ptxas
  // PREDICATES the loads instead of branching around them, and a predicated-off load still takes its issue slot — 14 of ~390
  // slots per warp-step pair with a time-invariant cost (SASS: `@!P2 LDS.128 [UR9+..]` x 8 in the line search, `@!P2 LDS` x 6 in
  // the reverse pass).  A loop whose bound the compiler cannot see is a real (warp-uniform) branch.
  const int ldOwn = CM >= 1 ? 0 : __shfl_sync(0xffffffffu, constOwn ? 0 : 1, 0), ldLs = __shfl_sync(0xffffffffu, constLs ? 0 : 1, 0);
  // per-element costs bulk-copied into shared memory (io.stage_words > 0; never with CSH)
  const int SW = CSH ? 0 : io.stage_words;
  R* wreg = sm + (size_t)wid * group_warp_words(T, NX, NU, SW);
  R* sCe = wreg + group_coff_words(T, NX, NU) + (size_t)e * SW;                                      // this element's [C_b | c_b]
  unsigned long long* cbar = reinterpret_cast<unsigned long long*>(wreg + group_coff_words(T, NX, NU) + (size_t)kEPW * SW) + e;
  unsigned cphase = 0;
  if (SW > 0) {
    if (lane < kEPW) mbar_init(reinterpret_cast<unsigned long long*>(wreg + group_coff_words(T, NX, NU) + (size_t)kEPW * SW) + lane, 1);
    mbar_fence_init();
    __syncwarp();
  }
  R* xb = wreg + e;                                   // nominal X word w at xb[w * kXS], x_ref word w at xb[w * kXS + kEPW]
  R* ub = wreg + group_uoff_words(T, NX) + e;         // input buffer s, word w at ub[w * kUS + s * kEPW]; u_ref: slot kNBUF
  // The gains have no storage of their own: between the reverse pass that produces them and the line search
  // that consumes them the five non-nominal input buffers are dead, and candidate a overwrites U_t of its
  // buffer only at the end of step t.  K_t[m][i] lives in the U_t[m] slot of the i-th dead buffer, k_t[m] in
  // that of the NX-th one ((NX + 1) NU words needed, 5 NU there: NX <= 4); the line search reads them at the top of step t.
  // The candidates' states go to the warp's global-memory scratch tile, word w = X_{t+1}[i] of lane l at scr[(t NX + i) 32 + l].
  // (small batches — few warps per CTA, the kernel lasts as long as its longest chain of iterations — keep the tiles in shared
  // memory, behind the warps' regions: no L2 round trip per iteration; io.xscratch == nullptr)
  constexpr bool ssm = SSM;
  R* scr0 = ssm ? sm + (size_t)(blockDim.x >> 5) * group_warp_words(T, NX, NU, SW) + (size_t)wid * group_xscratch_smem_words(T, NX, LSSEP)
                : io.xscratch + ((size_t)blockIdx.x * (blockDim.x >> 5) + wid) * group_xscratch_words(T, NX);
  const size_t tileW = group_xtile_words(T, NX);
  R* myV = wreg + group_voff_words(T, NX, NU) + (size_t)e * (NX + 1) * 4;   // [e][NX+1][4]: rows of V, then v
  // Dyn<> (LTI) wants A and B separately: straight from global memory (uniform, cached)
  const R* sA = P.dyn_A;
  const R* sB = P.dyn_B;

  R ulo[NU], uhi[NU];
#pragma unroll
  for (int i = 0; i < NU; ++i) {
    ulo[i] = P.has_umin ? P.umin[i] : -inf_<R>();
    uhi[i] = P.has_umax ? P.umax[i] : inf_<R>();
  }
  const R alpha = P.alphas[(j < P.n_alpha) ? j : P.n_alpha - 1];
  const bool replay = io.replay_iters != nullptr;

  // group state, replicated in the lanes of the group
  bool has = false, exhausted = !valid;
  int b = 0, it = 0, ai_prev = -1;   // ai_prev: winner of the previous iteration (-1: the nominal states are the hand-over io.X)
  unsigned flg = 0;
  R best = 0;
  const R *Cp = io.C, *cp = io.c, *Cfp = io.Cf, *cfp = io.cf;

  // The element's OWN cost (cost.py:37-62) of the trajectory (nominal states, inputs of buffer slot wbo), the five lanes
  // splitting the evaluation by rows of C_t: lane i accumulates sum_t tau_i (C_t[i,:] tau) and sum_t c_t[i] tau_i, the partial
  // sums are combined in lane order.  Used with per-element costs scored by a separate line-search cost (quirk Q4): for the
  // acceptance cost of the winner (controller.py:298) and for the initial cost of a new element — instead of every candidate
  // lane reading the whole per-element C_t at every step of the line search (30 words per step and lane; 6 here, 2 for a
  // compact diagonal).  Every lane of the warp must call it (full-mask shuffles); all lanes of a group get the same value.
  auto own_cost_rows = [&](int wbo) -> R {
    const R* Cfo = ownCf_sh ? sCf : Cfp;
    const R* cfo = ownCf_sh ? sCf + N * N : cfp;
    // (t runs downwards, as in the reverse pass, which accumulates the same sums for the nominal trajectory)
    const R* prow = (DIAG ? Cp + jj : Cp + jj * N) + (size_t)(T - 1) * CW;
    const R* pc = cp + jj + (size_t)(T - 1) * N;
    const R* rx = xb + (T - 1) * NX * kXS;
    const R* wu = ub + (T - 1) * NU * kUS;
    R tauN[NX];
#pragma unroll
    for (int i = 0; i < NX; ++i) tauN[i] = rx[(NX + i) * kXS] - rx[(NX + i) * kXS + kEPW];
    R quad = 0, lin = 0;
    for (int t = T - 1; t >= 0; --t) {
      R tau[N];
#pragma unroll
      for (int i = 0; i < NX; ++i) tau[i] = rx[i * kXS] - rx[i * kXS + kEPW];
#pragma unroll
      for (int i = 0; i < NU; ++i) tau[NX + i] = wu[i * kUS + wbo] - wu[i * kUS + kURef];
      const R tj = pick<N, R>(tau, jj);
      R s = 0;
      if (DIAG) {
        // q_j tau_j plus the exact zeros the dense row would add (NaN if some tau_k is not finite)
#pragma unroll
        for (int k = 0; k < N; ++k) s += tau[k] * (R)0;
        s += prow[0] * tj;
      } else {
#pragma unroll
        for (int k = 0; k < N; ++k) s += prow[k] * tau[k];
      }
      quad += tj * s;
      lin += pc[0] * tj;
      rx -= NX * kXS; wu -= NU * kUS; prow -= CW; pc -= N;
    }
    R qsum = 0, lsum = 0;
#pragma unroll
    for (int i = 0; i < N; ++i) {
      qsum += __shfl_sync(FULL, quad, base + i);
      lsum += __shfl_sync(FULL, lin, base + i);
    }
    R qN2 = 0, lN2 = 0;
#pragma unroll
    for (int i = 0; i < NX; ++i) {
      R s2 = 0;
#pragma unroll
      for (int k = 0; k < NX; ++k) s2 += (DIAG ? ((i == k) ? Cfo[i] : (R)0) : Cfo[i * N + k]) * tauN[k];
      qN2 += tauN[i] * s2;
      lN2 += cfo[i] * tauN[i];
    }
    return ((R)0.5 * qsum + lsum) + ((R)0.5 * qN2 + lN2);
  };
  // with per-element costs and a separate scoring cost the initial cost of an element is evaluated here (row-split, from the
  // staged cost block) and ilqr_init_rollout only rolls the initial inputs out: a thread per element walking its own 2.4 KB of
  // costs took 176 us of the 1.06 ms forward at B = 65536 (profiles/r2_launches_per_element.csv)
  constexpr bool kOwnInit = LSSEP && !CSH;
  // States X_1..X_T of one candidate (word w at src[w st]) -> the nominal state trajectory, the group's lanes taking every
  // kG-th word.  Four loads are issued before their four stores: with the tile in shared memory a load-store-load-store loop
  // is one dependent chain (the stores may alias the next load), 17 shared-memory round trips per lane and iteration.
  auto fetch_states = [&](const R* src, size_t st, bool in_smem) {
    const int nwd = T * NX;
    for (int w = j; w < nwd; w += 4 * kG) {
      R v4[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) v4[q] = (w + q * kG < nwd) ? ld_scr<R>(src + (size_t)(w + q * kG) * st, in_smem) : (R)0;
#pragma unroll
      for (int q = 0; q < 4; ++q)
        if (w + q * kG < nwd) xb[(NX + w + q * kG) * kXS] = v4[q];
    }
  };

  while (true) {
    // ------------------------------------------------------------------ refill (queue pop by the group leader)
    bool fresh = false;
    {
      const bool need = !has && !exhausted;
      unsigned nb = 0xffffffffu;
      if (need && j == 0) nb = atomicAdd(io.counter, 1u);
      nb = __shfl_sync(FULL, nb, base);
      if (need) {
        if (nb < (io.n_dev ? *io.n_dev : (unsigned)P.B)) {
          b = io.order ? io.order[nb] : (int)nb;
          has = true;
          ai_prev = -1;
          it = io.resume ? io.iters[b] : 0;               // resume: the thread-per-element kernel ran the first iterations
          flg = io.resume ? (unsigned)io.flags[b] : 0u;
          Cp = io.C + (P.C_batched ? (size_t)b * T * CW : 0);
          cp = io.c + (P.C_batched ? (size_t)b * T * N : 0);
          if (SW > 0) {
            // The element's running cost, [T, cw] + [T, n] contiguous in global memory, comes in with two bulk asynchronous
            // copies (cp.async.bulk -> UBLKCP, completion on this slot's mbarrier) and is then read from shared memory for the
            // rest of the solve: ~10 passes over it (reverse pass + acceptance cost of every iteration) that round 1 served
            // from L2 with six scalar loads per step and lane (1.6 GB of L2 -> SM traffic per forward at B = 65536).
            if (j == 0) {
              fence_proxy_async();          // the previous element's reads of this slot (generic proxy) are ordered before the copy
              mbar_expect_tx(cbar, (unsigned)((T * CW + T * N) * sizeof(R)));
              bulk_g2s(sCe, Cp, (unsigned)(T * CW * sizeof(R)), cbar);
              bulk_g2s(sCe + T * CW, cp, (unsigned)(T * N * sizeof(R)), cbar);
            }
            Cp = sCe;
            cp = sCe + T * CW;
            if (P.Cf_batched) {   // the 30 (10) words of the terminal cost: 100-byte blocks are not 16-byte aligned, plain loads
              R* sf = sCe + T * CW + T * N;
              for (int w = j; w < CW; w += kG) sf[w] = Cfp[w];
              for (int w = j; w < N; w += kG) sf[CW + w] = cfp[w];
              Cfp = sf;
              cfp = sf + CW;
            }
          }
          Cfp = io.Cf + (P.Cf_batched ? (size_t)b * CW : 0);
          cfp = io.cf + (P.Cf_batched ? (size_t)b * N : 0);
          // nominal trajectory (initial rollout) and references: global -> shared, once per element.  The
          // references are read ten times per step and trip; from global memory (6 lines per request, a
          // ~15 KB L1 next to 217 KB of shared memory) they were 35 of 123 data-path cycles per warp-step
          // (profiles/r1_forward_group_v3).
          const R* xr = io.xref + (P.xref_batched ? (size_t)b * szX : 0);
          const R* ur = io.uref + (P.uref_batched ? (size_t)b * szU : 0);
          const R* Xi = io.X + (size_t)b * szX;
          const R* Ui = io.Uinit + (size_t)b * szU;
          for (int w = j; w < XW; w += kG) { xb[w * kXS] = Xi[w]; xb[w * kXS + kEPW] = xr[w]; }
          for (int w = j; w < UW; w += kG) { ub[w * kUS] = Ui[w]; ub[w * kUS + kURef] = ur[w]; }   // nominal inputs: buffer 0
          if (!kOwnInit || io.resume) best = cost0[b];
          else fresh = true;
          if (SW > 0) { mbar_wait(cbar, cphase); cphase ^= 1u; }   // the cost block has landed
          if (io.alpha_trace && !io.resume)
            for (int w = j; w < P.max_iter; w += kG) io.alpha_trace[(size_t)b * P.max_iter + w] = 0xFF;
        } else {
          exhausted = true;
        }
      }
    }
    __syncwarp();
    if (__ballot_sync(FULL, has) == 0u) break;

    bool run = has && it < P.max_iter;
    if (run && replay) run = it < io.replay_iters[b];
    bool improved = false;
    if constexpr (kOwnInit) {
      // a new element that runs no iteration at all (max_iter = 0) still reports its initial cost
      if (__ballot_sync(FULL, fresh && !run) != 0u) {
        const R c0 = own_cost_rows(0);
        if (fresh && !run) best = c0;
      }
    }
    if (__ballot_sync(FULL, run) != 0u) {
      // this iteration's tile of candidate states (the second tile only serves the restore after a rejected step: LSSEP)
      R* scr = LSSEP ? scr0 + (size_t)(it & 1) * tileW : scr0;
      // The nominal inputs always sit in buffer 0 (the winner's inputs are copied there on acceptance: 20 words per
      // iteration), candidate a writes buffer a + 1, gain word k lives in buffer k + 1 — every shared-memory access of the
      // two passes has a compile-time offset from one running pointer (an index swap instead of the copy cost ~25 integer
      // instructions per step in address arithmetic).
      constexpr int nomo = 0;
      const int dKj = ((j < NX ? j : NX) + 1) * kEPW;   // the dead buffer this lane stores its share of the gains into
      // terminal cost of this element: staged or global (generic pointers; read twice per pass)
      const R* Cfo = ownCf_sh ? sCf : Cfp;
      const R* cfo = ownCf_sh ? sCf + N * N : cfp;
      // ---------------------------------------------------------------- reverse Riccati pass
      // quadraticize (cost.py:64-100) + linearize (controller.py:448-471) + backward_lqr (546-578)
      {
        R V[NX * NX], v[NX];
        // (kOwnInit) the element's own cost of the nominal trajectory, accumulated row by row next to the quadraticisation —
        // the same sums own_cost_rows() forms; it is the initial `best` of an element on its first iteration
        R rq = 0, rl = 0, rqN = 0, rlN = 0;
        {
          R tauN[NX];
#pragma unroll
          for (int i = 0; i < NX; ++i) tauN[i] = xb[(T * NX + i) * kXS] - xb[(T * NX + i) * kXS + kEPW];
#pragma unroll
          for (int i = 0; i < NX; ++i) {
            R s = 0;
#pragma unroll
            for (int k = 0; k < NX; ++k) {
              const R cik = DIAG ? ((i == k) ? Cfo[i] : (R)0) : Cfo[i * N + k];
              V[i * NX + k] = cik;
              s += cik * tauN[k];
            }
            v[i] = s + cfo[i];
            if constexpr (kOwnInit) { rqN += tauN[i] * s; rlN += cfo[i] * tauN[i]; }
          }
        }
        R jcm[NJ];   // compact Jacobian of the step this lane evaluated in the current block of kG steps
#pragma unroll
        for (int m = 0; m < NJ; ++m) jcm[m] = 0;
        // running pointers of step t (all decremented once per step)
        const R* px = xb + (T - 1) * NX * kXS;   // X_t (nominal), x_ref_t at + kEPW
        R* pu = ub + (T - 1) * NU * kUS;         // U_t: word i of buffer s at pu[i * kUS + s * kEPW]
        const R* pCrow = CSH ? sC + (size_t)(T - 1) * CS + jj * N : DIAG ? Cp + (size_t)(T - 1) * N + jj : Cp + ((size_t)(T - 1) * N + jj) * N;
        const R* pcj = CSH ? sC + (size_t)(T - 1) * CS + N * N + jj : cp + (size_t)(T - 1) * N + jj;
        R Crow[N], cj;        // row jj of C_t and c_t[jj]; loaded once when the running cost is time-invariant
        constexpr bool kRowPF = !CSH && !DIAG;   // dense per-element costs (global memory): prefetched one step ahead
        R nCrow[kRowPF ? N : 1], ncj = 0;
        (void)nCrow; (void)ncj;
        if (DIAG) {
          const R qd = pCrow[0];
#pragma unroll
          for (int k = 0; k < N; ++k) Crow[k] = (k == jj) ? qd : (R)0;
        } else {
#pragma unroll
          for (int k = 0; k < N; ++k) Crow[k] = pCrow[k];
        }
        cj = pcj[0];
        int s_in_blk = 0;
        for (int t = T - 1; t >= 0; --t) {
          if (GD::kHasJac && s_in_blk == 0) {
            const int back = (j <= t) ? j : t;      // lane j evaluates step t - j (clamped at 0)
            const R* qx = px - back * NX * kXS;
            const R* qu = pu - back * NU * kUS + nomo;
            R xj[NX], uj[NU];
#pragma unroll
            for (int i = 0; i < NX; ++i) xj[i] = qx[i * kXS];
#pragma unroll
            for (int i = 0; i < NU; ++i) uj[i] = qu[i * kUS];
            GD::jac(P, xj, uj, jcm);
          }
          R jc[NJ];
          if (GD::kHasJac) {
#pragma unroll
            for (int m = 0; m < NJ; ++m) jc[m] = __shfl_sync(FULL, jcm[m], base + s_in_blk);
          } else {
#pragma unroll
            for (int m = 0; m < NJ; ++m) jc[m] = 0;
          }
          s_in_blk = (s_in_blk == kG - 1) ? 0 : s_in_blk + 1;
          // row jj of C_t, l_j = C_t[j,:] tau + c_t[j]
          R Qrow[N], qj;
          if constexpr (CM == 2) {
            // time-invariant DIAGONAL cost (cost mode 2): row jj of C has one entry, so this lane needs its OWN tau_jj only —
            // 2 shared-memory loads and 2 arithmetic instructions instead of 10 and 10.  The dense sum adds exact zeros to the same
            // product (a non-finite tau_k, k != jj, would make it NaN there: such a nominal trajectory is never improved on —
            // NaN / inf candidate costs do not pass `new < best` — and l only feeds k_t and v, not Q_uu or the flags).
            const R* pa = (jj < NX) ? px + jj * kXS : pu + (jj - NX) * kUS + nomo;   // the word of z_t = [X_t; U_t] this lane's row meets
            const R tj = pa[0] - pa[(jj < NX) ? kEPW : kURef];
#pragma unroll
            for (int k = 0; k < N; ++k) Qrow[k] = Crow[k];
            R s = 0;
            s += pick<N, R>(Crow, jj) * tj;
            qj = s + cj;
          } else {
          R tau[N];
#pragma unroll
          for (int i = 0; i < NX; ++i) tau[i] = px[i * kXS] - px[i * kXS + kEPW];
#pragma unroll
          for (int i = 0; i < NU; ++i) tau[NX + i] = pu[i * kUS + nomo] - pu[i * kUS + kURef];
            if (DIAG) {
              const R qd = pCrow[0];
#pragma unroll
              for (int k = 0; k < N; ++k) Crow[k] = (k == jj) ? qd : (R)0;
              cj = pcj[0];
            } else if constexpr (kRowPF) {
              // per-element dense costs read through L2: row jj of C_{t-1} is requested now and consumed one step later (the
              // ~200 instructions of this step hide the L2 round trip that a load-and-use paid at the top of every step)
              if (t > 0) {
#pragma unroll
                for (int k = 0; k < N; ++k) nCrow[k] = (pCrow - CW)[k];
                ncj = (pcj - N)[0];
              }
            } else {
#pragma unroll 1
              for (int r = 0; r < ldOwn; ++r) {
#pragma unroll
                for (int k = 0; k < N; ++k) Crow[k] = pCrow[k];
                cj = pcj[0];
              }
            }
            R s = 0;
#pragma unroll
            for (int k = 0; k < N; ++k) {
              Qrow[k] = Crow[k];
              s += Crow[k] * tau[k];
            }
            qj = s + cj;   // l = C tau + c (cost.py:86-88)
            if constexpr (kOwnInit) {
              const R tj = pick<N, R>(tau, jj);
              rq += tj * s;
              rl += cj * tj;
            }
          }
          // row jj of F' V F and element jj of F' v
          {
            R f[NX], w[NX];
            GD::col(P, sF, jc, jj, f);
#pragma unroll
            for (int l = 0; l < NX; ++l) {
              R s = 0;
#pragma unroll
              for (int i = 0; i < NX; ++i) s += f[i] * V[i * NX + l];
              w[l] = s;
            }
            GD::mulF(P, sF, jc, w, Qrow);          // Qrow = C_t[j,:] + w' F
#pragma unroll
            for (int k = NX; k < N; ++k) Qrow[k] = (k == jj) ? Qrow[k] + P.reg : Qrow[k];   // Q_uu + reg I (controller.py:554)
#pragma unroll
            for (int i = 0; i < NX; ++i) qj += f[i] * v[i];
          }
          // gather Q_xu, Q_uu, q_u; every lane computes the gains
          R Qxu[NX * NU], Quu[NU * NU], qu[NU];
#pragma unroll
          for (int m = 0; m < NU; ++m) {
#pragma unroll
            for (int i = 0; i < NX; ++i) Qxu[i * NU + m] = __shfl_sync(FULL, Qrow[NX + m], base + i);
#pragma unroll
            for (int m2 = 0; m2 < NU; ++m2) Quu[m * NU + m2] = __shfl_sync(FULL, Qrow[NX + m2], base + NX + m);
            qu[m] = __shfl_sync(FULL, qj, base + NX + m);
          }
          R Kt[NU * NX], kt[NU];
          const bool npd = gains<R, NX, NU>(P.reg, Qxu, Quu, qu, Kt, kt);
          if (run && npd) flg |= TACD_FLAG_NONPD;
          if (j < NX) {
#pragma unroll
            for (int m = 0; m < NU; ++m) pu[m * kUS + dKj] = pick<NX, R>(&Kt[m * NX], j);
          } else if (j < N) {
            pu[(j - NX) * kUS + dKj] = pick<NU, R>(kt, j - NX);
          }
          // row i = j of the value update (controller.py:573-574), association as value_update<>
          {
            R QK[NU * NX], Qk[NU], Ki[NU];
#pragma unroll
            for (int m = 0; m < NU; ++m) {
#pragma unroll
              for (int c = 0; c < NX; ++c) {
                R s = 0;
#pragma unroll
                for (int m2 = 0; m2 < NU; ++m2) s += Quu[m * NU + m2] * Kt[m2 * NX + c];
                QK[m * NX + c] = s;
              }
              R s = 0;
#pragma unroll
              for (int m2 = 0; m2 < NU; ++m2) s += Quu[m * NU + m2] * kt[m2];
              Qk[m] = s;
              Ki[m] = pick<NX, R>(&Kt[m * NX], jj < NX ? jj : 0);
            }
            Row4<R> vr;
#pragma unroll
            for (int c = 0; c < 4; ++c) vr.v[c] = 0;
#pragma unroll
            for (int c = 0; c < NX; ++c) {
              R acc = Qrow[c];
#pragma unroll
              for (int m = 0; m < NU; ++m) acc += Ki[m] * QK[m * NX + c];
#pragma unroll
              for (int m = 0; m < NU; ++m) acc += Ki[m] * Qxu[c * NU + m];
#pragma unroll
              for (int m = 0; m < NU; ++m) acc += Qrow[NX + m] * Kt[m * NX + c];
              vr.v[c] = acc;
            }
            R vi = qj;
#pragma unroll
            for (int m = 0; m < NU; ++m) vi += Ki[m] * Qk[m];
#pragma unroll
            for (int m = 0; m < NU; ++m) vi += Ki[m] * qu[m];
#pragma unroll
            for (int m = 0; m < NU; ++m) vi += Qrow[NX + m] * kt[m];
            if (j < NX) {
              *reinterpret_cast<Row4<R>*>(myV + j * 4) = vr;
              myV[NX * 4 + j] = vi;
            }
          }
          __syncwarp();
#pragma unroll
          for (int i = 0; i < NX; ++i) {
            const Row4<R> r = *reinterpret_cast<const Row4<R>*>(myV + i * 4);
#pragma unroll
            for (int c = 0; c < NX; ++c) V[i * NX + c] = r.v[c];
          }
          {
            const Row4<R> r = *reinterpret_cast<const Row4<R>*>(myV + NX * 4);
#pragma unroll
            for (int c = 0; c < NX; ++c) v[c] = r.v[c];
          }
          __syncwarp();
          px -= NX * kXS; pu -= NU * kUS;
          pCrow -= CSH ? CS : CW; pcj -= CSH ? CS : N;
          if constexpr (kRowPF) {
#pragma unroll
            for (int k = 0; k < N; ++k) Crow[k] = nCrow[k];
            cj = ncj;
          }
        }
        if constexpr (kOwnInit) {
          if (__ballot_sync(FULL, fresh) != 0u) {   // some group of this warp is on the first iteration of a new element
            R qsum = 0, lsum = 0;
#pragma unroll
            for (int i = 0; i < N; ++i) {
              qsum += __shfl_sync(FULL, rq, base + i);
              lsum += __shfl_sync(FULL, rl, base + i);
            }
            if (fresh) best = ((R)0.5 * qsum + lsum) + ((R)0.5 * rqN + rlN);
          }
        }
      }
      // ---------------------------------------------------------------- line search: lane a = candidate a
      // evaluate_alphas (controller.py:589-620); the candidate trajectory goes to this lane's buffer
      R cost_ls, cost_own;
      {
        const int cbo = (j + 1) * kEPW;          // this candidate's input buffer
        R x[NX], q1 = 0, l1 = 0, q2 = 0, l2 = 0;
#pragma unroll
        for (int i = 0; i < NX; ++i) x[i] = xb[i * kXS];
        const R* px = xb;
        R* pu = ub;
        R* sx = scr + lane;                      // candidate X_{t+1}[i] -> sx[(t NX + i) 32]
        const R* pCl = sCl;                      // staged line-search cost of step t
        const R* pCo = CSH ? sC : Cp;            // own cost of step t (staged block or global C_t)
        const R* pco = cp;
        R Ct[N * N], ct[N], Cl[LSSEP ? N * N : 1], cl[LSSEP ? N : 1];
        if (CSH) load_cost_vec<R, N>(pCo, Ct, ct);
        if constexpr (LSSEP) load_cost_vec<R, N>(pCl, Cl, cl);
        for (int t = 0; t < T; ++t) {
          R u[NU], tau[N], dx[NX], xn[NX];
#pragma unroll
          for (int k = 0; k < NX; ++k) dx[k] = x[k] - px[k * kXS];
#pragma unroll
          for (int i = 0; i < NU; ++i) {
            R s = 0;
#pragma unroll
            for (int k = 0; k < NX; ++k) s += pu[i * kUS + (k + 1) * kEPW] * dx[k];
            R ui = pu[i * kUS + nomo] + (alpha * pu[i * kUS + (NX + 1) * kEPW] + s);
            ui = tmin<R>(tmax<R>(ui, ulo[i]), uhi[i]);
            u[i] = ui;
            tau[NX + i] = ui - pu[i * kUS + kURef];
          }
#pragma unroll
          for (int i = 0; i < NX; ++i) tau[i] = x[i] - px[i * kXS + kEPW];
          if constexpr (LSSEP) {
#pragma unroll 1
            for (int r = 0; r < ldLs; ++r) load_cost_vec<R, N>(pCl, Cl, cl);
            if constexpr (DIAG) {
              // the line-search cost of a compact-diagonal problem is a diagonal too (element 0's, expanded while staging): 20
              // instead of 35 instructions per candidate and step, bit-identical sums (quad_terms_diag).  A RUN-TIME test for
              // "all off-diagonal entries of the staged cost are exact zeros" (the usual Q, R = diag cost) was measured too: the
              // second code path costs more (registers, code) than the 15 instructions save — forward +1..3 % on the dense
              // layouts, +8 % at B = 512 — so only the compile-time case takes the short form.
              R qd[N];
#pragma unroll
              for (int i = 0; i < N; ++i) qd[i] = Cl[i * N + i];
              quad_terms_diag<R, N>(qd, cl, tau, q1, l1);
            } else {
              quad_terms<R, N>(Cl, cl, tau, q1, l1);
            }
            pCl += CS;
          }
          if constexpr (!LSSEP) {   // (with a separate line-search cost the own cost is evaluated for the winner only)
            if (CSH) {
#pragma unroll 1
              for (int r = 0; r < ldOwn; ++r) load_cost_vec<R, N>(pCo, Ct, ct);
              pCo += CS;
            } else if (DIAG) {
#pragma unroll
              for (int i = 0; i < N; ++i) { Ct[i] = pCo[i]; ct[i] = pco[i]; }
              pCo += N; pco += N;
            } else {
#pragma unroll
              for (int i = 0; i < N * N; ++i) Ct[i] = pCo[i];
#pragma unroll
              for (int i = 0; i < N; ++i) ct[i] = pco[i];
              pCo += N * N; pco += N;
            }
            if (DIAG) {
              R qd[N];
#pragma unroll
              for (int i = 0; i < N; ++i) qd[i] = Ct[i];
              quad_terms_diag<R, N>(qd, ct, tau, q2, l2);
            } else if (CM == 2) {
              R qd[N];
#pragma unroll
              for (int i = 0; i < N; ++i) qd[i] = Ct[i * N + i];
              quad_terms_diag<R, N>(qd, ct, tau, q2, l2);
            } else {
              quad_terms<R, N>(Ct, ct, tau, q2, l2);
            }
          }
          D::step(P, sA, sB, x, u, xn);
          __syncwarp();   // every lane of the group has read K_t, k_t from the slots overwritten below
          if (valid) {
#pragma unroll
            for (int i = 0; i < NU; ++i) pu[i * kUS + cbo] = u[i];
#pragma unroll
            for (int i = 0; i < NX; ++i) sx[i * 32] = xn[i];
          }
#pragma unroll
          for (int i = 0; i < NX; ++i) x[i] = xn[i];
          px += NX * kXS; pu += NU * kUS; sx += NX * 32;
        }
        R tauN[NX];
#pragma unroll
        for (int i = 0; i < NX; ++i) tauN[i] = x[i] - px[i * kXS + kEPW];
        R qN = 0, lN = 0, qN2 = 0, lN2 = 0;
#pragma unroll
        for (int i = 0; i < NX; ++i) {
          R s = 0, s2 = 0;
#pragma unroll
          for (int k = 0; k < NX; ++k) {
            if (LSSEP) s += sCfl[i * N + k] * tauN[k];
            else s2 += (DIAG ? ((i == k) ? Cfo[i] : (R)0) : Cfo[i * N + k]) * tauN[k];
          }
          if (LSSEP) { qN += tauN[i] * s; lN += sCfl[N * N + i] * tauN[i]; }
          else { qN2 += tauN[i] * s2; lN2 += cfo[i] * tauN[i]; }
        }
        cost_own = ((R)0.5 * q2 + l2) + ((R)0.5 * qN2 + lN2);
        cost_ls = LSSEP ? ((R)0.5 * q1 + l1) + ((R)0.5 * qN + lN) : cost_own;
      }
      // ---------------------------------------------------------------- accept
      // argmin (first minimum, NaN wins) + strict improvement (controller.py:292-301)
      R costs[kG];
#pragma unroll
      for (int a = 0; a < kG; ++a) costs[a] = __shfl_sync(FULL, cost_ls, base + a);
      int ai = 0;
      {
        R cbest = costs[0];
        bool done = false;
#pragma unroll
        for (int a = 1; a < kG; ++a) {
          if (a < P.n_alpha && !done) {
            if (cbest != cbest) done = true;
            else if (costs[a] < cbest || costs[a] != costs[a]) { ai = a; cbest = costs[a]; }
          }
        }
      }
      if (run && replay) ai = io.replay_alpha[(size_t)b * P.max_iter + it];
      if constexpr (LSSEP) {
        // Acceptance cost (controller.py:298): the element's OWN cost of the chosen candidate only.  Its states come back
        // from the scratch tile into the state trajectory shared memory holds BEFORE the step is judged (one round of 17
        // independent L2 loads per lane instead of a dependent load per step of the cost loop); a rejected step — the last
        // iteration of every solve — restores the nominal states from the other tile (the previous iteration's winner).
        __syncwarp();                       // the candidates' inputs and states are complete and visible to the group
        if (run) {
          const R* wsx = scr + base + ai;
          fetch_states(wsx, 32, ssm);
        }
        __syncwarp();
        cost_own = own_cost_rows((ai + 1) * kEPW);
      }
      const R newc = LSSEP ? cost_own : __shfl_sync(FULL, cost_own, base + ai);
      improved = newc < best;
      if (run && replay) improved = (it < io.replay_iters[b] - 1) || ((io.replay_flags[b] & TACD_FLAG_MAXITER) != 0);
      if (run && j == 0) {
        if (io.alpha_trace) io.alpha_trace[(size_t)b * P.max_iter + it] = (uint8_t)ai;
        if (io.cost_trace) {
          R* ctr = io.cost_trace + ((size_t)b * P.max_iter + it) * (P.n_alpha + 2);
#pragma unroll
          for (int a = 0; a < kG; ++a)
            if (a < P.n_alpha) ctr[a] = costs[a];
          ctr[P.n_alpha] = newc;
          ctr[P.n_alpha + 1] = best;
        }
      }
      if constexpr (!LSSEP) __syncwarp();   // the candidates' states (global scratch) are visible to the whole group
      if (run) {
        it += 1;
        if (improved) {
          // accept: the winner's inputs are copied into the nominal buffer, its states come back from the scratch tile
          // into the one state trajectory shared memory holds (X_0 is common to all candidates)
          best = newc;
          const R* wu = ub + (ai + 1) * kEPW;
          if constexpr (!LSSEP) {
            const R* wsx = scr + base + ai;
            fetch_states(wsx, 32, ssm);
          }
          for (int w = j; w < UW; w += kG) ub[w * kUS] = wu[w * kUS];
          ai_prev = ai;
        } else if constexpr (LSSEP) {
          // rejected: the nominal states were overwritten by the candidate's for the acceptance cost; they are the previous
          // iteration's winner (the other scratch tile) or, on an element's first iteration, the hand-over trajectory io.X
          const R* src = ai_prev >= 0 ? scr0 + (size_t)(it & 1) * tileW + base + ai_prev : io.X + (size_t)b * szX + NX;
          const int st = ai_prev >= 0 ? 32 : 1;
          fetch_states(src, (size_t)st, ssm && ai_prev >= 0);
        }
      }
      __syncwarp();   // inputs / states written by one lane are read by the whole group from here on
    }
    // -------------------------------------------------------------------- finalise
    // outputs + final re-linearisation (controller.py:311-314) + active set (636-642)
    const bool keep_going = run && improved && it < P.max_iter;
    if (has && !keep_going) {
      if (run && improved) flg |= TACD_FLAG_MAXITER;
      const R* nomu = ub;
      R* Xo = io.X + (size_t)b * szX;
      R* Uo = io.U + (size_t)b * szU;
      for (int w = j; w < XW; w += kG) Xo[w] = xb[w * kXS];
      for (int w = j; w < UW; w += kG) {
        const R ui = nomu[w * kUS];
        Uo[w] = ui;
        const int i = w % NU;
        bool m = false;
        if (P.has_umin) m = m || isclose_<R>(ui, P.umin[i]);
        if (P.has_umax) m = m || isclose_<R>(ui, P.umax[i]);
        io.tight[(size_t)b * szU + w] = m ? 1 : 0;
      }
      if (io.A != nullptr || io.Bm != nullptr) {
        for (int t = j; t < T; t += kG) {
          R x[NX], u[NU], A[NX * NX], Bm[NX * NU];
#pragma unroll
          for (int i = 0; i < NX; ++i) x[i] = xb[(t * NX + i) * kXS];
#pragma unroll
          for (int i = 0; i < NU; ++i) u[i] = nomu[(t * NU + i) * kUS];
          D::jac(P, sA, sB, x, u, A, Bm);
          if (io.A)
#pragma unroll
            for (int i = 0; i < NX * NX; ++i) io.A[((size_t)b * T + t) * NX * NX + i] = A[i];
          if (io.Bm)
#pragma unroll
            for (int i = 0; i < NX * NU; ++i) io.Bm[((size_t)b * T + t) * NX * NU + i] = Bm[i];
        }
      }
      if (j == 0) {
        io.iters[b] = it;
        io.flags[b] = (uint8_t)flg;
        if (io.cost) io.cost[b] = best;
      }
      has = false;
    }
    __syncwarp();
  }
}

}  // namespace tacd
