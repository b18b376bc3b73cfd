// ilqr_group_bwd.cuh — implicit-differentiation backward (ILQRSolve.backward + _zero_constrained_lqr,
// controller.py:63-115, 707-788) with FIVE LANES PER BATCH ELEMENT, the layout of ilqr_group.cuh.
//
// The thread-per-element backward (ilqr_small.cuh) walks every [B, ...] tensor with one row per thread —
// 32 sectors per request — holds 246 registers, and spends half of its 1 225 instructions per element-step
// transposing per-lane gradients through shared memory to sum them over the batch
// (profiles/r1_backward_v5).  Here lane j of a group owns row j of the Q-function in the reverse pass (as in
// the forward kernel) and row j of every gradient in the forward sweep:
//   * grad_C[t][j,:] = -1/2 (dtau_j tau + tau_j dtau), grad_c[t][j] = -dtau_j, grad_ref[t][j] = dtau_j are
//     produced where they are stored, so a group writes 25 + 5 + 5 contiguous words per step;
//   * shared (un-batched) inputs: the six groups of a warp are summed with 3 shuffles per value, group 0
//     accumulates into the warp's [T x NV] accumulator, the CTA combines its warps into one partial vector,
//     reduce_shared_grads() sums the CTAs in a fixed order (deterministic, no atomics);
//   * X*, U* are staged once per element into shared memory with coalesced copies; gains K_t, k_t stay there
//     between the two passes; Jacobians are evaluated five steps at a time in compact form.
// Reference-mode only (quirks Q1-Q3, Q6 reproduced); compat="exact", external Jacobians and run-time
// dimensions stay on the thread-per-element / generic kernels.
#pragma once
#include "ilqr_group.cuh"

namespace tacd {

__host__ __device__ inline int gbwd_elem_words(int T, int nx, int nu) { return (T + 1) * nx + T * nu + T * nu * (nx + 1); }
__host__ __device__ inline int gbwd_voff_words(int T, int nx, int nu) { return (gbwd_elem_words(T, nx, nu) * kEPW + 3) & ~3; }
__host__ __device__ inline int gbwd_warp_words(int T, int nx, int nu) { return gbwd_voff_words(T, nx, nu) + kEPW * (nx + 1) * 4; }

// sum over the six groups of a warp of a value held by lane (e, j); valid in the lanes of group 0.
// Spare lanes and dead groups must hold 0.  Fixed order: ((g0 + g3) + (g1 + g4)) + (g2 + g5).
template <typename R>
TACD_DEV R sum_groups(R v) {
  v += __shfl_down_sync(0xffffffffu, v, 3 * kG);
  const R a = __shfl_down_sync(0xffffffffu, v, kG);
  const R b = __shfl_down_sync(0xffffffffu, v, 2 * kG);
  return (v + a) + b;
}

template <typename R, int NX, int NU, int DYN, bool CSH, bool DIAG>
__global__ void __launch_bounds__(512, 1) ilqr_backward_group(const DevProblem<R> P, const BwdPtrs<R> io, R* partials, int reduce) {
  constexpr int N = NX + NU;
  static_assert(N <= kG && NX <= 4, "same limits as the forward group kernel");
  static_assert(!(DIAG && CSH), "compact diagonal costs are per-element");
  using D = Dyn<R, NX, NU, DYN>;
  using GD = GDyn<R, NX, NU, DYN>;
  using L = BwdLayout<NX, NU>;
  constexpr int NJ = GD::NJ;
  constexpr unsigned FULL = 0xffffffffu;
  constexpr int CW = DIAG ? N : N * N;
  constexpr int CS = (N * N + N + 3) & ~3;     // staged shared cost: [C_t | pad] per step (c_t is not needed here)
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int T = P.T;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, nw = blockDim.x >> 5;
  const bool valid = lane < kEPW * kG;
  const int e = valid ? lane / kG : 0;
  const int j = valid ? lane - e * kG : kG;
  const int base = e * kG;
  const int jj = j < N ? j : N - 1;
  const int XW = (T + 1) * NX, W = XW + T * NU;
  const size_t szX = (size_t)XW, szU = (size_t)T * NU;
  const int NACC = L::nacc(T);

  R* sm = reinterpret_cast<R*>(smem_raw);
  R *sC = sm, *sF = sm;
  bool same = true;
  {
    int o = 0;
    const int NT = blockDim.x;
    if (CSH) {
      sC = sm;
      o += T * CS;
      for (int i = tid; i < T * CS; i += NT) {
        const int t = i / CS, k = i - t * CS;
        const R val = (k < N * N) ? io.C[t * N * N + k] : (R)0;
        same = same && (k >= N * N || val == io.C[k]);
        sC[i] = val;
      }
    }
    if (DYN == TACD_DYN_LTI) {
      sF = sm + o;
      o += (NX * N + 3) & ~3;
      for (int i = tid; i < NX * N; i += NT) {
        const int r = i / N, k = i % N;
        sF[i] = (k < NX) ? P.dyn_A[r * NX + k] : P.dyn_B[r * NU + (k - NX)];
      }
    }
    sm += o;
  }
  const bool constC = __syncthreads_and(same ? 1 : 0) != 0 && CSH;
  R* wreg = sm + (size_t)wid * gbwd_warp_words(T, NX, NU);
  R* ebuf = wreg + e;                                   // [word][e]: X (XW), U (T*NU), K (T*NU*NX), k (T*NU)
  R* eK = ebuf + (size_t)W * kEPW;
  R* ek = eK + (size_t)T * NU * NX * kEPW;
  R* myV = wreg + gbwd_voff_words(T, NX, NU) + (size_t)e * (NX + 1) * 4;
  R* acc = sm + (size_t)nw * gbwd_warp_words(T, NX, NU);   // [nw][NACC], only when reduce
  R* myacc = acc + (size_t)wid * NACC;
  if (reduce) {
    for (int i = tid; i < nw * NACC; i += blockDim.x) acc[i] = 0;
  }
  __syncthreads();
  const R* sA = P.dyn_A;
  const R* sB = P.dyn_B;
  const bool redC = reduce && io.grad_C && !P.C_batched, redc = reduce && io.grad_c && !P.C_batched;
  const bool redCf = reduce && io.grad_Cf && !P.Cf_batched, redcf = reduce && io.grad_cf && !P.Cf_batched;
  const bool redx = reduce && io.grad_xref && !P.xref_batched, redu = reduce && io.grad_uref && !P.uref_batched;

  const int ngroups = gridDim.x * nw * kEPW;
  const int g0 = (blockIdx.x * nw + wid) * kEPW;        // first group of this warp
  for (int bw = g0; bw < P.B; bw += ngroups) {          // warp-uniform loop
    const int b = bw + e;
    const bool live = valid && b < P.B;
    const int bb = (b < P.B) ? b : P.B - 1;             // dead groups shadow the last element, outputs masked
    const R* Cp = io.C + (P.C_batched ? (size_t)bb * T * CW : 0);
    const R* xr = io.xref + (P.xref_batched ? (size_t)bb * szX : 0);
    const R* ur = io.uref + (P.uref_batched ? (size_t)bb * szU : 0);
    const R* gX = io.gX + (size_t)bb * szX;
    const R* gU = io.gU + (size_t)bb * szU;
    const uint8_t* tg = io.tight + (size_t)bb * szU;
    // stage X*, U* (coalesced over the group)
    {
      const R* Xi = io.X + (size_t)bb * szX;
      const R* Ui = io.U + (size_t)bb * szU;
      if (valid) {
        for (int w = j; w < XW; w += kG) ebuf[(size_t)w * kEPW] = Xi[w];
        for (int w = j; w < W - XW; w += kG) ebuf[(size_t)(XW + w) * kEPW] = Ui[w];
      }
    }
    __syncwarp();
    // ---------------------------------------------------------------- reverse pass
    // V = H_xx[T-1], v = -grad_X[T-1] (quirks Q1, Q2), every lane
    R V[NX * NX], v[NX];
    {
      const R* CT = CSH ? sC + (size_t)(T - 1) * CS : Cp + (size_t)(T - 1) * CW;
#pragma unroll
      for (int i = 0; i < NX; ++i) {
#pragma unroll
        for (int k = 0; k < NX; ++k) V[i * NX + k] = DIAG ? ((i == k) ? CT[i] : (R)0) : CT[i * N + k];
        v[i] = -gX[(T - 1) * NX + i];
      }
    }
    R jcm[NJ];
#pragma unroll
    for (int m = 0; m < NJ; ++m) jcm[m] = 0;
    const R* px = ebuf + (size_t)(T - 1) * NX * kEPW;
    const R* pu = ebuf + (size_t)(XW + (T - 1) * NU) * kEPW;
    R* pK = eK + (size_t)(T - 1) * NU * NX * kEPW;
    R* pk = ek + (size_t)(T - 1) * NU * kEPW;
    const R* pCrow = CSH ? sC + (size_t)(T - 1) * CS + jj * N : DIAG ? Cp + (size_t)(T - 1) * N + jj : Cp + ((size_t)(T - 1) * N + jj) * N;
    // upstream gradient of this lane's row: -grad_X[t][j] or -grad_U[t][j-nx] (prefetched one step ahead)
    const R* pg = (jj < NX) ? gX + (T - 1) * NX + jj : gU + (T - 1) * NU + (jj - NX);
    const int gstep = (jj < NX) ? NX : NU;
    R gnext = pg[0];
    R Crow[N];
    if (DIAG) {
      const R qd = pCrow[0];
#pragma unroll
      for (int k = 0; k < N; ++k) Crow[k] = (k == jj) ? qd : (R)0;
    } else {
#pragma unroll
      for (int k = 0; k < N; ++k) Crow[k] = pCrow[k];
    }
    int s_in_blk = 0;
    for (int t = T - 1; t >= 0; --t) {
      if (GD::kHasJac && s_in_blk == 0) {
        const int back = (j <= t) ? j : t;
        const R* qx = px - (size_t)back * NX * kEPW;
        const R* qu = pu - (size_t)back * NU * kEPW;
        R xj[NX], uj[NU];
#pragma unroll
        for (int i = 0; i < NX; ++i) xj[i] = qx[i * kEPW];
#pragma unroll
        for (int i = 0; i < NU; ++i) uj[i] = qu[i * kEPW];
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
      R Qrow[N], qj = -gnext;
      if (t > 0) { pg -= gstep; gnext = pg[0]; }
      if (DIAG) {
        const R qd = pCrow[0];
#pragma unroll
        for (int k = 0; k < N; ++k) Crow[k] = (k == jj) ? qd : (R)0;
      } else if (!constC) {
#pragma unroll
        for (int k = 0; k < N; ++k) Crow[k] = pCrow[k];
      }
#pragma unroll
      for (int k = 0; k < N; ++k) Qrow[k] = Crow[k];
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
        GD::mulF(P, sF, jc, w, Qrow);
#pragma unroll
        for (int k = NX; k < N; ++k) Qrow[k] = (k == jj) ? Qrow[k] + P.reg : Qrow[k];   // + reg I (controller.py:731)
#pragma unroll
        for (int i = 0; i < NX; ++i) qj += f[i] * v[i];
      }
      R Qxu[NX * NU], Quu[NU * NU], qu[NU];
#pragma unroll
      for (int m = 0; m < NU; ++m) {
#pragma unroll
        for (int i = 0; i < NX; ++i) Qxu[i * NU + m] = __shfl_sync(FULL, Qrow[NX + m], base + i);
#pragma unroll
        for (int m2 = 0; m2 < NU; ++m2) Quu[m * NU + m2] = __shfl_sync(FULL, Qrow[NX + m2], base + NX + m);
        qu[m] = __shfl_sync(FULL, qj, base + NX + m);
      }
      // box QP on the free inputs (controller.py:737-763), every lane of the group
      R kt[NU];
      {
        bool fixed[NU], any_free = false;
        R H[NU * NU], q[NU], lo[NU], hi[NU];
#pragma unroll
        for (int i = 0; i < NU; ++i) { fixed[i] = tg[t * NU + i] != 0; any_free = any_free || !fixed[i]; }
#pragma unroll
        for (int i = 0; i < NU; ++i) {
#pragma unroll
          for (int c = 0; c < NU; ++c) H[i * NU + c] = (fixed[i] || fixed[c]) ? ((i == c) ? (R)1 : (R)0) : Quu[i * NU + c];
          q[i] = fixed[i] ? (R)0 : qu[i];
          const R ui = pu[i * kEPW];
          if (fixed[i]) { lo[i] = 0; hi[i] = 0; }
          else if (P.has_umin) { lo[i] = P.umin[i] - ui; hi[i] = P.umax[i] - ui; }   // quirk Q10
          else { lo[i] = -inf_<R>(); hi[i] = inf_<R>(); }
        }
        if (any_free) {
          pnqp<R, NU>(H, q, lo, hi, kt);
#pragma unroll
          for (int i = 0; i < NU; ++i) kt[i] = fixed[i] ? (R)0 : kt[i];
        } else {
#pragma unroll
          for (int i = 0; i < NU; ++i) kt[i] = 0;
        }
      }
      // K_t = -solve(Q_uu + 1e-8 reg I, Q_ux), not masked by the active set (quirk Q6)
      R Kt[NU * NX];
      {
        R M[NU * NU];
#pragma unroll
        for (int i = 0; i < NU; ++i) {
#pragma unroll
          for (int c = 0; c < NU; ++c) M[i * NU + c] = Quu[i * NU + c] + ((i == c) ? (R)1e-8 * P.reg : (R)0);
#pragma unroll
          for (int c = 0; c < NX; ++c) Kt[i * NX + c] = Qxu[c * NU + i];
        }
        lu_solve<R, NU, NX>(M, Kt);
#pragma unroll
        for (int i = 0; i < NU * NX; ++i) Kt[i] = -Kt[i];
      }
      if (j < NX) {
#pragma unroll
        for (int m = 0; m < NU; ++m) pK[(m * NX + j) * kEPW] = pick<NX, R>(&Kt[m * NX], j);
      } else if (j < N) {
        pk[(j - NX) * kEPW] = pick<NU, R>(kt, j - NX);
      }
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
          R a = Qrow[c];
#pragma unroll
          for (int m = 0; m < NU; ++m) a += Ki[m] * QK[m * NX + c];
#pragma unroll
          for (int m = 0; m < NU; ++m) a += Ki[m] * Qxu[c * NU + m];
#pragma unroll
          for (int m = 0; m < NU; ++m) a += Qrow[NX + m] * Kt[m * NX + c];
          vr.v[c] = a;
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
      px -= NX * kEPW; pu -= NU * kEPW; pK -= NU * NX * kEPW; pk -= NU * kEPW;
      pCrow -= CSH ? CS : CW;
    }
    // ---------------------------------------------------------------- forward sweep from dX_0 = 0 (quirk Q3)
    if (io.grad_x0 && live && j < NX) io.grad_x0[(size_t)b * NX + j] = 0;
    R dx[NX];
#pragma unroll
    for (int i = 0; i < NX; ++i) dx[i] = 0;
    px = ebuf;
    pu = ebuf + (size_t)XW * kEPW;
    const R* qK = eK;
    const R* qk = ek;
    s_in_blk = 0;
    for (int t = 0; t < T; ++t) {
      if (GD::kHasJac && s_in_blk == 0) {
        const int fwd = (t + j < T) ? j : T - 1 - t;     // lane j evaluates step t + j (clamped at T-1)
        const R* qx = px + (size_t)fwd * NX * kEPW;
        const R* qu = pu + (size_t)fwd * NU * kEPW;
        R xj[NX], uj[NU];
#pragma unroll
        for (int i = 0; i < NX; ++i) xj[i] = qx[i * kEPW];
#pragma unroll
        for (int i = 0; i < NU; ++i) uj[i] = qu[i * kEPW];
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
      R tau[N], dtau[N];
#pragma unroll
      for (int i = 0; i < NX; ++i) { tau[i] = px[i * kEPW] - xr[t * NX + i]; dtau[i] = dx[i]; }
#pragma unroll
      for (int i = 0; i < NU; ++i) {
        tau[NX + i] = pu[i * kEPW] - ur[t * NU + i];
        R s = 0;
#pragma unroll
        for (int k = 0; k < NX; ++k) s += qK[(i * NX + k) * kEPW] * dx[k];
        dtau[NX + i] = s + qk[i * kEPW];
      }
      // row jj of the gradients of step t
      const R tj = pick<N, R>(tau, jj), dj = pick<N, R>(dtau, jj);
      R gC[N];
#pragma unroll
      for (int k = 0; k < N; ++k) gC[k] = (R)-0.5 * (dj * tau[k] + tj * dtau[k]);
      const bool rowlive = live && j < N;
      if (rowlive) {
        if (io.grad_C && P.C_batched) {
          if (DIAG) io.grad_C[((size_t)b * T + t) * N + j] = pick<N, R>(gC, jj);
          else {
#pragma unroll
            for (int k = 0; k < N; ++k) io.grad_C[(((size_t)b * T + t) * N + j) * N + k] = gC[k];
          }
        }
        if (io.grad_c && P.C_batched) io.grad_c[((size_t)b * T + t) * N + j] = -dj;
        if (j < NX) {
          if (io.grad_xref && P.xref_batched) io.grad_xref[(size_t)b * szX + t * NX + j] = dj;
          if (io.dX) io.dX[(size_t)b * szX + t * NX + j] = dj;
        } else {
          if (io.grad_uref && P.uref_batched) io.grad_uref[(size_t)b * szU + t * NU + (j - NX)] = dj;
          if (io.dU) io.dU[(size_t)b * szU + t * NU + (j - NX)] = dj;
        }
      }
      if (redC || redc || redx || redu) {   // warp-uniform
        const R z = rowlive ? (R)1 : (R)0;
        if (redC) {
#pragma unroll
          for (int k = 0; k < N; ++k) {
            const R s = sum_groups<R>(gC[k] * z);
            if (lane < N) myacc[t * L::NV + lane * N + k] += s;
          }
        }
        if (redc) {
          const R s = sum_groups<R>(-dj * z);
          if (lane < N) myacc[t * L::NV + N * N + lane] += s;
        }
        if (redx || redu) {
          const R s = sum_groups<R>(dj * z);
          if (lane < N && ((lane < NX) ? redx : redu)) myacc[t * L::NV + N * N + N + lane] += s;
        }
      }
      R dxn[NX];
      GD::mulFv(P, sF, jc, dtau, dxn);
#pragma unroll
      for (int i = 0; i < NX; ++i) dx[i] = dxn[i];
      px += NX * kEPW; pu += NU * kEPW; qK += NU * NX * kEPW; qk += NU * kEPW;
    }
    {
      // terminal: grad_C_final[:nx,:nx], grad_c_final[:nx] (zero padding elsewhere), grad_xref[T]
      R tauN[NX];
#pragma unroll
      for (int i = 0; i < NX; ++i) tauN[i] = px[i * kEPW] - xr[T * NX + i];
      const int jx = jj < NX ? jj : 0;
      const R tj = pick<NX, R>(tauN, jx), dj = pick<NX, R>(dx, jx);
      R gF[NX];
#pragma unroll
      for (int k = 0; k < NX; ++k) gF[k] = (R)-0.5 * (dj * tauN[k] + tj * dx[k]);
      const bool rowlive = live && j < N;
      if (rowlive) {
        if (io.grad_Cf && P.Cf_batched) {
          if (DIAG) io.grad_Cf[(size_t)b * N + j] = (j < NX) ? pick<NX, R>(gF, jx) : (R)0;
          else {
#pragma unroll
            for (int k = 0; k < N; ++k) io.grad_Cf[((size_t)b * N + j) * N + k] = (j < NX && k < NX) ? gF[k < NX ? k : 0] : (R)0;
          }
        }
        if (io.grad_cf && P.Cf_batched) io.grad_cf[(size_t)b * N + j] = (j < NX) ? -dj : (R)0;
        if (j < NX) {
          if (io.grad_xref && P.xref_batched) io.grad_xref[(size_t)b * szX + T * NX + j] = dj;
          if (io.dX) io.dX[(size_t)b * szX + T * NX + j] = dj;
        }
      }
      if (redCf || redcf || redx) {
        const R z = (live && j < NX) ? (R)1 : (R)0;
        if (redCf) {
#pragma unroll
          for (int k = 0; k < NX; ++k) {
            const R s = sum_groups<R>(gF[k] * z);
            if (lane < NX) myacc[T * L::NV + lane * NX + k] += s;
          }
        }
        if (redcf) {
          const R s = sum_groups<R>(-dj * z);
          if (lane < NX) myacc[T * L::NV + NX * NX + lane] += s;
        }
        if (redx) {
          const R s = sum_groups<R>(dj * z);
          if (lane < NX) myacc[T * L::NV + NX * NX + NX + lane] += s;
        }
      }
    }
    __syncwarp();
  }
  if (reduce) {
    __syncthreads();
    for (int i = tid; i < NACC; i += blockDim.x) {
      R sum = 0;
      for (int w = 0; w < nw; ++w) sum += acc[(size_t)w * NACC + i];
      partials[(size_t)blockIdx.x * NACC + i] = sum;
    }
  }
}

}  // namespace tacd
