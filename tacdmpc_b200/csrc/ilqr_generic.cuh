// ilqr_generic.cuh — run-time-dimension kernels: one CTA (nx > 16) or one warp (nx <= 16) per batch element.
//
// Covers every (nx <= 64, nu <= 16, dynamics) the small-system kernels have no instantiation for (config 5: LTI
// nx in {8,16,32}, nu in {4,8}, T=50; Python dynamics with external Jacobians) and the stand-alone stage entry
// points.  Per-step matrices (V, Q*, A, B, C_t, gains) live in shared memory; the trajectory state that must survive
// between passes (X, U, K, k) sits right behind them when enough element slots per SM still fit, else in a
// per-slot slice of the caller's workspace (contiguous, L2-resident).  Every matrix product is a loop over output
// entries (or, for the compile-time-size instances, over 1x4 register tiles) spread over the element's threads; the
// nu x nu factorisations and the box QP run on one thread (in registers for the compile-time sizes).
#pragma once
#include "tacd_dev.cuh"

namespace tacd {

constexpr int kGenThreads = 128;

// Run-time dimensions; NXT / NUT > 0 pin nx / nu at compile time (the whole-solve kernels are instantiated for the
// sizes of the LTI sweep), which turns the index arithmetic of every matrix loop into immediates
// and lets the k-loops unroll.
template <typename R, int NXT = 0, int NUT = 0, int TPE_ = kGenThreads>
struct GenDims {
  int nx_, nu_, T, dyn;
  __host__ __device__ __forceinline__ int nx() const { return NXT > 0 ? NXT : nx_; }
  __host__ __device__ __forceinline__ int nu() const { return NUT > 0 ? NUT : nu_; }
  __host__ __device__ __forceinline__ int n() const { return nx() + nu(); }
  // Threads per element: the whole CTA (kGenThreads) or one warp.  With one warp per element a CTA carries
  // kGenThreads / 32 independent elements, each with its own shared-memory carve, and every barrier of the
  // per-step pipeline is a __syncwarp: small systems (nx <= 16) have too few matrix entries to occupy 128 threads
  // and spent most of their time at CTA barriers (profiles/r1_backward_generic: 54 % of samples).
  static constexpr int tpe = TPE_;
  static constexpr int NX = NXT, NU = NUT;
  static constexpr bool fixed = NXT > 0 && NUT > 0 && NXT % 4 == 0 && NUT % 4 == 0;  // 128-bit tiles are legal
  static constexpr int per_cta = kGenThreads / TPE_;
  static_assert(TPE_ == 32 || TPE_ == kGenThreads, "an element is worked on by one warp or by the whole CTA");
  static __device__ __forceinline__ int tid() { return TPE_ == kGenThreads ? threadIdx.x : (threadIdx.x & (TPE_ - 1)); }
  static __device__ __forceinline__ int sub() { return TPE_ == kGenThreads ? 0 : threadIdx.x / TPE_; }
  static __device__ __forceinline__ int slot() { return blockIdx.x * per_cta + sub(); }
  static __device__ __forceinline__ int nslots() { return gridDim.x * per_cta; }
  static __device__ __forceinline__ void sync() {
    if (TPE_ == 32) __syncwarp(); else __syncthreads();
  }
};

// ---- run-time-dimension dynamics on one thread (non-LTI plugins are tiny) ----
template <typename R>
__device__ void gen_dyn_step_scalar(const DevProblem<R>& P, int dyn, const R* x, const R* u, R* xn) {
  if (dyn == TACD_DYN_CARTPOLE) {
    R xx[4] = {x[0], x[1], x[2], x[3]}, uu[1] = {u[0]}, o[4];
    cartpole_step<R>(P.dynp, P.dt, xx, uu, o);
    for (int i = 0; i < 4; ++i) xn[i] = o[i];
  } else {
    R xx[3] = {x[0], x[1], x[2]}, uu[2] = {u[0], u[1]}, o[3];
    if (dyn == TACD_DYN_UNICYCLE) unicycle_step<R, false>(P.dt, xx, uu, o);
    else unicycle_step<R, true>(P.dt, xx, uu, o);
    for (int i = 0; i < 3; ++i) xn[i] = o[i];
  }
}
template <typename R>
__device__ void gen_dyn_jac_scalar(const DevProblem<R>& P, int dyn, const R* x, const R* u, R* A, R* Bm) {
  if (dyn == TACD_DYN_CARTPOLE) {
    R xx[4] = {x[0], x[1], x[2], x[3]}, uu[1] = {u[0]}, a[16], b[4];
    if (P.jac_mode == TACD_JAC_FINITE_DIFF) Dyn<R, 4, 1, TACD_DYN_CARTPOLE>::jac_fd(P, nullptr, nullptr, xx, uu, a, b);
    else cartpole_jac<R>(P.dynp, P.dt, xx, uu, a, b);
    for (int i = 0; i < 16; ++i) A[i] = a[i];
    for (int i = 0; i < 4; ++i) Bm[i] = b[i];
  } else {
    R xx[3] = {x[0], x[1], x[2]}, uu[2] = {u[0], u[1]}, a[9], b[6];
    if (P.jac_mode == TACD_JAC_FINITE_DIFF) {
      if (dyn == TACD_DYN_UNICYCLE) Dyn<R, 3, 2, TACD_DYN_UNICYCLE>::jac_fd(P, nullptr, nullptr, xx, uu, a, b);
      else Dyn<R, 3, 2, TACD_DYN_UNICYCLE_WRAPPED>::jac_fd(P, nullptr, nullptr, xx, uu, a, b);
    } else unicycle_jac<R>(P.dt, xx, uu, a, b);
    for (int i = 0; i < 9; ++i) A[i] = a[i];
    for (int i = 0; i < 6; ++i) Bm[i] = b[i];
  }
}

// CTA-parallel x+ = f(x,u) for `na` independent (x,u) pairs laid out [a][nx], [a][nu]
template <typename R, typename DT>
__device__ void gen_dyn_step(const DevProblem<R>& P, const DT& d, const R* sA, const R* sB, int na, const R* x, const R* u, R* xn) {
  const int tid = DT::tid();
  if (d.dyn == TACD_DYN_LTI) {
    for (int idx = tid; idx < na * d.nx(); idx += DT::tpe) {
      const int a = idx / d.nx(), i = idx % d.nx();
      R s = 0, s2 = 0;
      for (int j = 0; j < d.nx(); ++j) s += sA[i * d.nx() + j] * x[a * d.nx() + j];
      for (int j = 0; j < d.nu(); ++j) s2 += sB[i * d.nu() + j] * u[a * d.nu() + j];
      xn[a * d.nx() + i] = s + s2;
    }
  } else {
    if (tid < na) gen_dyn_step_scalar<R>(P, d.dyn, x + tid * d.nx(), u + tid * d.nu(), xn + tid * d.nx());
  }
}

// A_t, B_t at (x,u) into shared memory (LTI: already resident, nothing to do).
// LTI + finite differences reproduces A,B up to rounding; the reference's FD mode is
// unreachable (controller.py:468-471 raises), so LTI always uses the matrices themselves.
template <typename R, typename DT>
__device__ void gen_dyn_jac(const DevProblem<R>& P, const DT& d, const R* x, const R* u, R* sA, R* sB) {
  if (d.dyn == TACD_DYN_LTI) return;
  if (DT::tid() == 0) gen_dyn_jac_scalar<R>(P, d.dyn, x, u, sA, sB);
}

// ---- single-thread nu x nu factorizations on shared memory -------------------
template <typename R>
__device__ bool gen_chol(R* M, int m) {
  for (int j = 0; j < m; ++j) {
    R dd = M[j * m + j];
    for (int k = 0; k < j; ++k) dd -= M[j * m + k] * M[j * m + k];
    if (!(dd > 0)) return false;
    dd = sqrt_<R>(dd);
    M[j * m + j] = dd;
    for (int i = j + 1; i < m; ++i) {
      R s = M[i * m + j];
      for (int k = 0; k < j; ++k) s -= M[i * m + k] * M[j * m + k];
      M[i * m + j] = s / dd;
    }
  }
  return true;
}
// LU with partial pivoting in place; piv[j] = row swapped with j at step j
template <typename R>
__device__ void gen_lu(R* M, int m, int* piv) {
  for (int j = 0; j < m; ++j) {
    int p = j;
    R best = abs_<R>(M[j * m + j]);
    for (int i = j + 1; i < m; ++i) {
      const R a = abs_<R>(M[i * m + j]);
      if (a > best) { best = a; p = i; }
    }
    piv[j] = p;
    if (p != j)
      for (int k = 0; k < m; ++k) { const R t = M[j * m + k]; M[j * m + k] = M[p * m + k]; M[p * m + k] = t; }
    const R dd = M[j * m + j];
    for (int i = j + 1; i < m; ++i) {
      const R f = M[i * m + j] / dd;
      M[i * m + j] = f;
      for (int k = j + 1; k < m; ++k) M[i * m + k] -= f * M[j * m + k];
    }
  }
}
// solve with the factors above for column r of rhs [m, nr] (one thread per column)
// (LAPACK getrs: gen_lu swaps WHOLE rows, so the stored multipliers are in their final row order and every interchange has to
// be applied to the right-hand side before the forward substitution starts.  Round 1 interleaved the two — correct only while at
// most one interchange touches a row, i.e. always for nu <= 2 and for diagonally dominant Q_uu; found in round 2 by
// tests/test_gpu_lti.py::test_gain_solve_against_lapack: 0.2-0.9 median relative error on indefinite 4 x 4 / 8 x 8 systems.)
template <typename R>
__device__ void gen_lu_solve_col(const R* M, int m, const int* piv, R* rhs, int nr, int r) {
  for (int j = 0; j < m; ++j) {
    const int p = piv[j];
    if (p != j) { const R t = rhs[j * nr + r]; rhs[j * nr + r] = rhs[p * nr + r]; rhs[p * nr + r] = t; }
  }
  for (int j = 0; j < m; ++j) {
    for (int i = j + 1; i < m; ++i) rhs[i * nr + r] -= M[i * m + j] * rhs[j * nr + r];
  }
  for (int i = m - 1; i >= 0; --i) {
    R s = rhs[i * nr + r];
    for (int k = i + 1; k < m; ++k) s -= M[i * m + k] * rhs[k * nr + r];
    rhs[i * nr + r] = s / M[i * m + i];
  }
}
template <typename R>
__device__ void gen_chol_solve_col(const R* L, int m, R* rhs, int nr, int r) {
  for (int i = 0; i < m; ++i) {
    R s = rhs[i * nr + r];
    for (int k = 0; k < i; ++k) s -= L[i * m + k] * rhs[k * nr + r];
    rhs[i * nr + r] = s / L[i * m + i];
  }
  for (int i = m - 1; i >= 0; --i) {
    R s = rhs[i * nr + r];
    for (int k = i + 1; k < m; ++k) s -= L[k * m + i] * rhs[k * nr + r];
    rhs[i * nr + r] = s / L[i * m + i];
  }
}

// ---- fixed-size factorizations in registers --------------------------------------------------------------------
// Same operations in the same order as gen_chol / gen_lu / *_solve_col above, on a fully unrolled register copy: the
// shared-memory versions serialise ~M^3/3 load -> fma -> store round trips (~30 cycles each) on one thread, which was
// most of a Riccati step's latency for nu = 8.
template <typename R, int M>
__device__ __forceinline__ bool gen_chol_reg(R* Ms) {
  R a[M][M];
#pragma unroll
  for (int i = 0; i < M; ++i)
#pragma unroll
    for (int j = 0; j <= i; ++j) a[i][j] = Ms[i * M + j];
  bool ok = true;
#pragma unroll
  for (int j = 0; j < M; ++j) {
    R dd = a[j][j];
#pragma unroll
    for (int k = 0; k < j; ++k) dd -= a[j][k] * a[j][k];
    if (!(dd > 0)) ok = false;
    dd = sqrt_<R>(dd);
    a[j][j] = dd;
#pragma unroll
    for (int i = j + 1; i < M; ++i) {
      R s = a[i][j];
#pragma unroll
      for (int k = 0; k < j; ++k) s -= a[i][k] * a[j][k];
      a[i][j] = s / dd;
    }
  }
  if (!ok) return false;  // the caller refactorises Q_uu with LU; shared memory still holds the input
#pragma unroll
  for (int i = 0; i < M; ++i)
#pragma unroll
    for (int j = 0; j <= i; ++j) Ms[i * M + j] = a[i][j];
  return true;
}
template <typename R, int M>
__device__ __forceinline__ void gen_chol_solve_col_reg(const R* L, R* rhs, int nr, int r) {
  R x[M];
#pragma unroll
  for (int i = 0; i < M; ++i) x[i] = rhs[i * nr + r];
#pragma unroll
  for (int i = 0; i < M; ++i) {
    R s = x[i];
#pragma unroll
    for (int k = 0; k < i; ++k) s -= L[i * M + k] * x[k];
    x[i] = s / L[i * M + i];
  }
#pragma unroll
  for (int i = M - 1; i >= 0; --i) {
    R s = x[i];
#pragma unroll
    for (int k = i + 1; k < M; ++k) s -= L[k * M + i] * x[k];
    x[i] = s / L[i * M + i];
  }
#pragma unroll
  for (int i = 0; i < M; ++i) rhs[i * nr + r] = x[i];
}
template <typename R, int M>
__device__ __forceinline__ void gen_lu_reg(R* Ms, int* piv) {
  R a[M][M];
#pragma unroll
  for (int i = 0; i < M; ++i)
#pragma unroll
    for (int j = 0; j < M; ++j) a[i][j] = Ms[i * M + j];
#pragma unroll
  for (int j = 0; j < M; ++j) {
    int p = j;
    R best = abs_<R>(a[j][j]);
#pragma unroll
    for (int i = j + 1; i < M; ++i) {
      const R v = abs_<R>(a[i][j]);
      if (v > best) { best = v; p = i; }
    }
    piv[j] = p;
#pragma unroll
    for (int i = j + 1; i < M; ++i) {
      const bool sw = (i == p);
#pragma unroll
      for (int k = 0; k < M; ++k) {
        const R lo = a[j][k], hi = a[i][k];
        a[j][k] = sw ? hi : lo;
        a[i][k] = sw ? lo : hi;
      }
    }
    const R dd = a[j][j];
#pragma unroll
    for (int i = j + 1; i < M; ++i) {
      const R f = a[i][j] / dd;
      a[i][j] = f;
#pragma unroll
      for (int k = j + 1; k < M; ++k) a[i][k] -= f * a[j][k];
    }
  }
#pragma unroll
  for (int i = 0; i < M; ++i)
#pragma unroll
    for (int j = 0; j < M; ++j) Ms[i * M + j] = a[i][j];
}
template <typename R, int M>
__device__ __forceinline__ void gen_lu_solve_col_reg(const R* Ms, const int* piv, R* rhs, int nr, int r) {
  R x[M];
#pragma unroll
  for (int i = 0; i < M; ++i) x[i] = rhs[i * nr + r];
#pragma unroll
  for (int j = 0; j < M; ++j) {   // all interchanges first (see gen_lu_solve_col)
    const int p = piv[j];
#pragma unroll
    for (int i = j + 1; i < M; ++i) {
      const bool sw = (i == p);
      const R lo = x[j], hi = x[i];
      x[j] = sw ? hi : lo;
      x[i] = sw ? lo : hi;
    }
  }
#pragma unroll
  for (int j = 0; j < M; ++j) {
#pragma unroll
    for (int i = j + 1; i < M; ++i) x[i] -= Ms[i * M + j] * x[j];
  }
#pragma unroll
  for (int i = M - 1; i >= 0; --i) {
    R s = x[i];
#pragma unroll
    for (int k = i + 1; k < M; ++k) s -= Ms[i * M + k] * x[k];
    x[i] = s / Ms[i * M + i];
  }
#pragma unroll
  for (int i = 0; i < M; ++i) rhs[i * nr + r] = x[i];
}

// pnqp (utils.py:66-124) on one thread, compacted problem of size m, scratch in `w`
// (needs m*m + 4m scalars and m ints).
template <typename R>
__device__ int gen_pnqp(int m, const R* H, const R* q, const R* lo, const R* hi, R* x, R* w, int* piv) {
  R* M = w;
  R* g = M + m * m;
  R* dx = g + m;
  R* xn = dx + m;
  R* frv = xn + m;
  auto lu_solve1 = [&](R* rhs) {
    gen_lu<R>(M, m, piv);
    gen_lu_solve_col<R>(M, m, piv, rhs, 1, 0);
  };
  auto obj = [&](const R* z) {
    R quad = 0, lin = 0;
    for (int i = 0; i < m; ++i) {
      R s = 0;
      for (int j = 0; j < m; ++j) s += H[i * m + j] * z[j];
      quad += z[i] * s;
      lin += q[i] * z[i];
    }
    return (R)0.5 * quad + lin;
  };
  if (m == 1) {
    x[0] = -(q[0] / H[0]);
  } else {
    for (int i = 0; i < m * m; ++i) M[i] = H[i];
    for (int i = 0; i < m; ++i) x[i] = q[i];
    lu_solve1(x);
    for (int i = 0; i < m; ++i) x[i] = -x[i];
  }
  for (int i = 0; i < m; ++i) x[i] = tmax<R>(tmin<R>(x[i], hi[i]), lo[i]);
  int it = 0;
  for (it = 0; it < 20; ++it) {
    for (int i = 0; i < m; ++i) {
      R s = 0;
      for (int j = 0; j < m; ++j) s += H[i * m + j] * x[j];
      g[i] = s + q[i];
    }
    for (int i = 0; i < m; ++i) {
      const bool at_lo = (x[i] <= lo[i]) && (g[i] > 0);
      const bool at_hi = (x[i] >= hi[i]) && (g[i] < 0);
      frv[i] = (at_lo || at_hi) ? (R)0 : (R)1;
    }
    for (int i = 0; i < m; ++i) {
      for (int j = 0; j < m; ++j) M[i * m + j] = H[i * m + j] * (frv[i] * frv[j]) + ((i == j) ? (R)1e-10 : (R)0);
      dx[i] = g[i] * frv[i];
    }
    if (m == 1) {
      dx[0] = -dx[0] / M[0];
    } else {
      lu_solve1(dx);
      for (int i = 0; i < m; ++i) dx[i] = -dx[i];
    }
    R mx = 0;
    bool isnan_ = false;
    for (int i = 0; i < m; ++i) {
      const R a = abs_<R>(dx[i]);
      if (a != a) isnan_ = true;
      if (a > mx) mx = a;
    }
    if (!isnan_ && mx < (R)1e-5) break;
    R alpha = 1;
    const R fx = obj(x);
    for (int ls = 0; ls < 10; ++ls) {
      R gd = 0;
      for (int i = 0; i < m; ++i) {
        xn[i] = tmax<R>(tmin<R>(x[i] + alpha * dx[i], hi[i]), lo[i]);
        gd += g[i] * (xn[i] - x[i]);
      }
      if (obj(xn) <= fx + (R)0.1 * gd) break;
      alpha *= (R)0.5;
    }
    for (int i = 0; i < m; ++i) x[i] = xn[i];
  }
  return (it < 20) ? it + 1 : 20;
}

// ---- shared-memory carve-up ---------------------------------------------------
template <typename R>
struct GenSmem {
  R *V, *Vn, *v, *vn, *A, *Bm, *Ct, *ct, *Qxx, *Qxu, *Quu, *qx, *qu, *AtV, *BtV, *Kt, *kt, *L, *rhs, *QK, *Qk, *tau, *x, *u;
  R *xa, *xan, *ua, *part1, *part2, *red, *scr, *Ct2, *taua;
  int* piv;
  int* flag;
};
__host__ __device__ inline size_t gen_smem_scalars(int nx, int nu) {
  const size_t n = nx + nu, na = TACD_MAX_ALPHA;
  return 2 * nx * nx + 2 * nx + nx * nx + nx * nu + n * n + n + nx * nx + nx * nu + nu * nu + nx + nu + nx * nx + nu * nx + nu * nx + nu +
         nu * nu + nu * (nx + 1) + nu * nx + nu + n + nx + nu +
         2 * na * nx + na * nu + 2 * na * n + kGenThreads + (2 * nu * nu + 8 * nu + 8) + n * n + na * n;
}
template <typename R>
__host__ __device__ inline size_t gen_smem_bytes(int nx, int nu) {
  return gen_smem_scalars(nx, nu) * sizeof(R) + (TACD_MAX_NU + 8) * sizeof(int);
}
template <typename R>
__host__ __device__ inline size_t gen_smem_bytes_aligned(int nx, int nu) { return (gen_smem_bytes<R>(nx, nu) + 15) & ~(size_t)15; }
// bytes of one element slot: matrices, then (optionally) its X, U, K, k slice
template <typename R>
__host__ __device__ inline size_t gen_elem_bytes(int nx, int nu, size_t slice_scalars) {
  return (gen_smem_bytes_aligned<R>(nx, nu) + slice_scalars * sizeof(R) + 15) & ~(size_t)15;
}
template <typename R>
__device__ GenSmem<R> gen_carve(unsigned char* raw, int nx, int nu) {
  const int n = nx + nu, na = TACD_MAX_ALPHA;
  GenSmem<R> s;
  R* p = reinterpret_cast<R*>(raw);
  auto take = [&](size_t k) { R* q = p; p += k; return q; };
  s.V = take(nx * nx); s.Vn = take(nx * nx); s.v = take(nx); s.vn = take(nx);
  s.A = take(nx * nx); s.Bm = take(nx * nu); s.Ct = take(n * n); s.ct = take(n);
  s.Qxx = take(nx * nx); s.Qxu = take(nx * nu); s.Quu = take(nu * nu); s.qx = take(nx); s.qu = take(nu);
  s.AtV = take(nx * nx); s.BtV = take(nu * nx); s.Kt = take(nu * nx); s.kt = take(nu);
  s.L = take(nu * nu); s.rhs = take(nu * (nx + 1)); s.QK = take(nu * nx); s.Qk = take(nu);
  s.tau = take(n); s.x = take(nx); s.u = take(nu);
  s.xa = take(na * nx); s.xan = take(na * nx); s.ua = take(na * nu);
  s.part1 = take(na * n); s.part2 = take(na * n); s.red = take(kGenThreads); s.scr = take(2 * nu * nu + 8 * nu + 8);
  s.Ct2 = take(n * n); s.taua = take(na * n);
  s.piv = reinterpret_cast<int*>(p);
  s.flag = s.piv + TACD_MAX_NU;
  return s;
}

// sum over the threads of one element (all of them get the result)
template <typename R, typename DT = GenDims<R>>
__device__ R gen_block_sum(R val, R* red) {
  if (DT::tpe == 32) {
    for (int o = 16; o > 0; o >>= 1) val += __shfl_xor_sync(0xffffffffu, val, o);
    return val;
  }
  const int tid = threadIdx.x;
  red[tid] = val;
  __syncthreads();
  for (int o = kGenThreads / 2; o > 0; o >>= 1) {
    if (tid < o) red[tid] += red[tid + o];
    __syncthreads();
  }
  const R r = red[0];
  __syncthreads();
  return r;
}

// objective of one trajectory held in global memory (cost.py:37-62), CTA-parallel
template <typename R, typename DT>
__device__ R gen_objective(const DT& d, GenSmem<R>& s, const R* X, const R* U, const R* C, const R* c, const R* Cf, const R* cf,
                           const R* xr, const R* ur) {
  const int tid = DT::tid(), nx = d.nx(), nu = d.nu(), n = d.n(), T = d.T;
  R acc_q = 0, acc_l = 0;
  // work item = (t, i): tau_i * (C_t[i,:] . tau) and c_t[i] tau_i
  for (int idx = tid; idx < T * n; idx += DT::tpe) {
    const int t = idx / n, i = idx % n;
    const R* Ct = C + (size_t)t * n * n + (size_t)i * n;
    R sdot = 0;
    for (int j = 0; j < nx; ++j) sdot += Ct[j] * (X[t * nx + j] - xr[t * nx + j]);
    for (int j = 0; j < nu; ++j) sdot += Ct[nx + j] * (U[t * nu + j] - ur[t * nu + j]);
    const R ti = (i < nx) ? (X[t * nx + i] - xr[t * nx + i]) : (U[t * nu + (i - nx)] - ur[t * nu + (i - nx)]);
    acc_q += ti * sdot;
    acc_l += c[(size_t)t * n + i] * ti;
  }
  R accN_q = 0, accN_l = 0;
  for (int i = tid; i < nx; i += DT::tpe) {
    R sdot = 0;
    for (int j = 0; j < nx; ++j) sdot += Cf[i * n + j] * (X[T * nx + j] - xr[T * nx + j]);
    const R ti = X[T * nx + i] - xr[T * nx + i];
    accN_q += ti * sdot;
    accN_l += cf[i] * ti;
  }
  const R q = gen_block_sum<R, DT>(acc_q, s.red), l = gen_block_sum<R, DT>(acc_l, s.red);
  const R qN = gen_block_sum<R, DT>(accN_q, s.red), lN = gen_block_sum<R, DT>(accN_l, s.red);
  return ((R)0.5 * q + l) + ((R)0.5 * qN + lN);
}

// ---- register-tiled products for the fixed-size instances -------------------------------------------------------
// One thread owns a 1 x 4 tile of the output: per k it reads one left entry and one 128-bit (2 x 128 for double)
// row segment of the right operand, i.e. 0.5 shared-memory instructions per multiply-add where the entry-per-thread
// loops need 2.  The sum over k runs in the same (ascending) order, so results are bit-identical to those loops.
template <typename R>
struct alignas(16) Quad { R v[4]; };
// acc[0..3] += l * r[0..3].  fp32: two packed FFMA2 (sm_100's fma.rn.f32x2; ptxas folds the {l, l} pair into the
// instruction's scalar-broadcast operand) — the same two IEEE fused multiply-adds per instruction, half the issue slots.
template <typename R>
__device__ __forceinline__ void gen_axpy4(R (&acc)[4], R l, const Quad<R>& r) {
#pragma unroll
  for (int c = 0; c < 4; ++c) acc[c] += l * r.v[c];
}
template <>
__device__ __forceinline__ void gen_axpy4<float>(float (&acc)[4], float l, const Quad<float>& r) {
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    unsigned long long lb, rb, ab;
    asm("mov.b64 %0, {%1, %1};" : "=l"(lb) : "f"(l));
    asm("mov.b64 %0, {%1, %2};" : "=l"(rb) : "f"(r.v[2 * h]), "f"(r.v[2 * h + 1]));
    asm("mov.b64 %0, {%1, %2};" : "=l"(ab) : "f"(acc[2 * h]), "f"(acc[2 * h + 1]));
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(ab) : "l"(lb), "l"(rb));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(acc[2 * h]), "=f"(acc[2 * h + 1]) : "l"(ab));
  }
}
// acc[c] += sum_k L(k, i) * Rm[k * ldr + j0 + c];  LT: L is indexed [k][i] (a transposed read), else [i][k]
template <typename R, int K, bool LT>
__device__ __forceinline__ void gen_mac4(R (&acc)[4], const R* L, int ldl, int i, const R* Rm, int ldr, int j0) {
  if (LT) {
#pragma unroll 8
    for (int k = 0; k < K; ++k) {
      const R l = L[k * ldl + i];
      const Quad<R> r = *reinterpret_cast<const Quad<R>*>(Rm + k * ldr + j0);
      gen_axpy4<R>(acc, l, r);
    }
  } else {
    static_assert(K % 4 == 0, "left rows are read four entries at a time");
#pragma unroll 2
    for (int k = 0; k < K; k += 4) {
      const Quad<R> l = *reinterpret_cast<const Quad<R>*>(L + i * ldl + k);
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        const Quad<R> r = *reinterpret_cast<const Quad<R>*>(Rm + (k + kk) * ldr + j0);
        gen_axpy4<R>(acc, l.v[kk], r);
      }
    }
  }
}

// One Riccati/KKT value-recursion step on shared memory.  On entry s.Qxx/Qxu/Quu/qx/qu
// hold the cost blocks, s.A/s.Bm the Jacobians, s.V/s.v the value at t+1.  `kkt`
// selects the backward-pass variant (box QP for k, LU gain with 1e-8 reg; controller.py:736-767).
template <typename R, typename DT>
__device__ bool gen_value_step(const DevProblem<R>& P, const DT& d, GenSmem<R>& s, bool kkt, const uint8_t* tight_t, const R* U_t) {
  const int tid = DT::tid(), nx = d.nx(), nu = d.nu();
  const int nr = nx + 1;
  bool nonpd = false;
  // compat="exact" (tacd_problem.compat_exact, backward only): the KKT system proper — no reg I in Q_uu (the reference
  // adds it, controller.py:731), active inputs frozen, the free block solved alone for K_t and k_t (fixes Q6)
  const bool exact = kkt && P.exact;
  const R quu_reg = exact ? (R)0 : P.reg;
  if constexpr (DT::fixed) {
    constexpr int NX = DT::NX, NU = DT::NU;
    for (int tl = tid; tl < NX * NX / 4; tl += DT::tpe) {
      const int i = tl / (NX / 4), j0 = (tl % (NX / 4)) * 4;
      R a[4] = {0, 0, 0, 0};
      gen_mac4<R, NX, true>(a, s.A, NX, i, s.V, NX, j0);
      *reinterpret_cast<Quad<R>*>(s.AtV + i * NX + j0) = Quad<R>{{a[0], a[1], a[2], a[3]}};
    }
    for (int tl = tid; tl < NU * NX / 4; tl += DT::tpe) {
      const int i = tl / (NX / 4), j0 = (tl % (NX / 4)) * 4;
      R a[4] = {0, 0, 0, 0};
      gen_mac4<R, NX, true>(a, s.Bm, NU, i, s.V, NX, j0);
      *reinterpret_cast<Quad<R>*>(s.BtV + i * NX + j0) = Quad<R>{{a[0], a[1], a[2], a[3]}};
    }
    DT::sync();
    for (int tl = tid; tl < NX * NX / 4; tl += DT::tpe) {
      const int i = tl / (NX / 4), j0 = (tl % (NX / 4)) * 4;
      R a[4] = {0, 0, 0, 0};
      gen_mac4<R, NX, false>(a, s.AtV, NX, i, s.A, NX, j0);
      Quad<R>& q = *reinterpret_cast<Quad<R>*>(s.Qxx + i * NX + j0);
      for (int c = 0; c < 4; ++c) q.v[c] += a[c];
    }
    for (int tl = tid; tl < NX * NU / 4; tl += DT::tpe) {
      const int i = tl / (NU / 4), j0 = (tl % (NU / 4)) * 4;
      R a[4] = {0, 0, 0, 0};
      gen_mac4<R, NX, false>(a, s.AtV, NX, i, s.Bm, NU, j0);
      Quad<R>& q = *reinterpret_cast<Quad<R>*>(s.Qxu + i * NU + j0);
      for (int c = 0; c < 4; ++c) q.v[c] += a[c];
    }
    for (int tl = tid; tl < NU * NU / 4; tl += DT::tpe) {
      const int i = tl / (NU / 4), j0 = (tl % (NU / 4)) * 4;
      R a[4] = {0, 0, 0, 0};
      gen_mac4<R, NX, false>(a, s.BtV, NX, i, s.Bm, NU, j0);
      Quad<R>& q = *reinterpret_cast<Quad<R>*>(s.Quu + i * NU + j0);
      for (int c = 0; c < 4; ++c) q.v[c] += a[c] + ((i == j0 + c) ? quu_reg : (R)0);
    }
  } else {
  for (int idx = tid; idx < nx * nx; idx += DT::tpe) {
    const int i = idx / nx, j = idx % nx;
    R a = 0;
    for (int k = 0; k < nx; ++k) a += s.A[k * nx + i] * s.V[k * nx + j];
    s.AtV[idx] = a;
  }
  for (int idx = tid; idx < nu * nx; idx += DT::tpe) {
    const int i = idx / nx, j = idx % nx;
    R a = 0;
    for (int k = 0; k < nx; ++k) a += s.Bm[k * nu + i] * s.V[k * nx + j];
    s.BtV[idx] = a;
  }
  DT::sync();
  for (int idx = tid; idx < nx * nx; idx += DT::tpe) {
    const int i = idx / nx, j = idx % nx;
    R a = 0;
    for (int k = 0; k < nx; ++k) a += s.AtV[i * nx + k] * s.A[k * nx + j];
    s.Qxx[idx] += a;
  }
  for (int idx = tid; idx < nx * nu; idx += DT::tpe) {
    const int i = idx / nu, j = idx % nu;
    R a = 0;
    for (int k = 0; k < nx; ++k) a += s.AtV[i * nx + k] * s.Bm[k * nu + j];
    s.Qxu[idx] += a;
  }
  for (int idx = tid; idx < nu * nu; idx += DT::tpe) {
    const int i = idx / nu, j = idx % nu;
    R a = 0;
    for (int k = 0; k < nx; ++k) a += s.BtV[i * nx + k] * s.Bm[k * nu + j];
    s.Quu[idx] += a + ((i == j) ? quu_reg : (R)0);
  }
  }
  for (int i = tid; i < nx; i += DT::tpe) {
    R a = 0;
    for (int k = 0; k < nx; ++k) a += s.A[k * nx + i] * s.v[k];
    s.qx[i] += a;
  }
  for (int i = tid; i < nu; i += DT::tpe) {
    R a = 0;
    for (int k = 0; k < nx; ++k) a += s.Bm[k * nu + i] * s.v[k];
    s.qu[i] += a;
  }
  DT::sync();
  // rhs = [Q_ux | q_u], L = Q_uu + reg' I
  const R reg2 = kkt ? (R)1e-8 * P.reg : P.reg;
  for (int idx = tid; idx < nu * nr; idx += DT::tpe) {
    const int i = idx / nr, j = idx % nr;
    const R r = (j < nx) ? s.Qxu[j * nu + i] : s.qu[i];
    s.rhs[idx] = (exact && tight_t[i]) ? (R)0 : r;
  }
  for (int idx = tid; idx < nu * nu; idx += DT::tpe) {
    const int i = idx / nu, j = idx % nu;
    R l = s.Quu[idx] + ((i == j) ? reg2 : (R)0);
    if (exact && (tight_t[i] || tight_t[j])) l = (i == j) ? (R)1 : (R)0;
    s.L[idx] = l;
  }
  DT::sync();
  // Without input bounds nothing can be tight and the box QP of the KKT pass (utils.py:66-124 with +-inf bounds) is
  // its own first Newton step, x = -H^-1 q, accepted at the first convergence check: k_t is then one more
  // right-hand side of the gain solve below (Q_uu vs Q_uu + 1e-8 reg I: 1e-14 apart) instead of ~3 single-thread
  // LU factorisations per step, which were 55 % of this kernel's instructions (profiles/r1_backward_generic).
  const bool box = (P.has_umin || P.has_umax) && !exact;   // exact: k_t is one more column of the masked solve
  if (tid == 0) {
    if (kkt && !box) {
      if constexpr (DT::fixed) gen_lu_reg<R, DT::NU>(s.L, s.piv); else gen_lu<R>(s.L, nu, s.piv);
      s.flag[0] = 0;
    } else if (kkt) {
      // k_t: box QP over the free coordinates (compacted, as the reference does)
      int m = 0;
      int* idxs = s.piv;  // reuse: free index list, overwritten by gen_lu afterwards
      R* w = s.scr;
      R* Hf = w + nu * nu + 4 * nu;
      R* qf = Hf + nu * nu;
      R* lof = qf + nu;
      R* hif = lof + nu;
      R* duf = hif + nu;
      int freeidx[TACD_MAX_NU];
      for (int i = 0; i < nu; ++i) { s.kt[i] = 0; if (!tight_t[i]) freeidx[m++] = i; }
      if (m > 0) {
        for (int i = 0; i < m; ++i) {
          for (int j = 0; j < m; ++j) Hf[i * m + j] = s.Quu[freeidx[i] * nu + freeidx[j]];
          qf[i] = s.qu[freeidx[i]];
          if (P.has_umin) { lof[i] = P.umin[freeidx[i]] - U_t[freeidx[i]]; hif[i] = P.umax[freeidx[i]] - U_t[freeidx[i]]; }
          else { lof[i] = -inf_<R>(); hif[i] = inf_<R>(); }
        }
        gen_pnqp<R>(m, Hf, qf, lof, hif, duf, w, idxs);
        for (int i = 0; i < m; ++i) s.kt[freeidx[i]] = duf[i];
      }
      if constexpr (DT::fixed) gen_lu_reg<R, DT::NU>(s.L, s.piv); else gen_lu<R>(s.L, nu, s.piv);
      s.flag[0] = 0;
    } else {
      bool pd;
      if constexpr (DT::fixed) pd = gen_chol_reg<R, DT::NU>(s.L); else pd = gen_chol<R>(s.L, nu);
      s.flag[0] = pd ? 1 : 0;
      if (!pd) {
        for (int idx = 0; idx < nu * nu; ++idx) s.L[idx] = s.Quu[idx] + ((idx / nu == idx % nu) ? reg2 : (R)0);
        if constexpr (DT::fixed) gen_lu_reg<R, DT::NU>(s.L, s.piv); else gen_lu<R>(s.L, nu, s.piv);
      }
    }
  }
  DT::sync();
  const bool use_chol = s.flag[0] != 0;
  if (!kkt && !use_chol) nonpd = true;
  const int ncols = (kkt && box) ? nx : nr;  // with bounds the KKT pass takes k from the box QP, not from the solve
  for (int r = tid; r < ncols; r += DT::tpe) {
    if constexpr (DT::fixed) {
      if (use_chol) gen_chol_solve_col_reg<R, DT::NU>(s.L, s.rhs, nr, r);
      else gen_lu_solve_col_reg<R, DT::NU>(s.L, s.piv, s.rhs, nr, r);
    } else {
      if (use_chol) gen_chol_solve_col<R>(s.L, nu, s.rhs, nr, r);
      else gen_lu_solve_col<R>(s.L, nu, s.piv, s.rhs, nr, r);
    }
  }
  DT::sync();
  for (int idx = tid; idx < nu * nx; idx += DT::tpe) s.Kt[idx] = -s.rhs[(idx / nx) * nr + (idx % nx)];
  if (!kkt || !box) for (int i = tid; i < nu; i += DT::tpe) s.kt[i] = -s.rhs[i * nr + nx];
  DT::sync();
  for (int idx = tid; idx < nu * nx; idx += DT::tpe) {
    const int i = idx / nx, j = idx % nx;
    R a = 0;
    for (int m = 0; m < nu; ++m) a += s.Quu[i * nu + m] * s.Kt[m * nx + j];
    s.QK[idx] = a;
  }
  for (int i = tid; i < nu; i += DT::tpe) {
    R a = 0;
    for (int m = 0; m < nu; ++m) a += s.Quu[i * nu + m] * s.kt[m];
    s.Qk[i] = a;
  }
  DT::sync();
  if constexpr (DT::fixed) {
    constexpr int NX = DT::NX, NU = DT::NU;
    for (int tl = tid; tl < NX * NX / 4; tl += DT::tpe) {
      const int i = tl / (NX / 4), j0 = (tl % (NX / 4)) * 4;
      R a[4] = {0, 0, 0, 0}, b[4] = {0, 0, 0, 0}, c[4] = {0, 0, 0, 0};
#pragma unroll
      for (int m = 0; m < NU; ++m) {
        const R ki = s.Kt[m * NX + i], xi = s.Qxu[i * NU + m];
        const Quad<R> qk = *reinterpret_cast<const Quad<R>*>(s.QK + m * NX + j0);
        const Quad<R> kj = *reinterpret_cast<const Quad<R>*>(s.Kt + m * NX + j0);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          a[e] += ki * qk.v[e];
          b[e] += ki * s.Qxu[(j0 + e) * NU + m];
          c[e] += xi * kj.v[e];
        }
      }
      const Quad<R> q = *reinterpret_cast<const Quad<R>*>(s.Qxx + i * NX + j0);
      *reinterpret_cast<Quad<R>*>(s.Vn + i * NX + j0) =
          Quad<R>{{((q.v[0] + a[0]) + b[0]) + c[0], ((q.v[1] + a[1]) + b[1]) + c[1], ((q.v[2] + a[2]) + b[2]) + c[2], ((q.v[3] + a[3]) + b[3]) + c[3]}};
    }
  } else {
  for (int idx = tid; idx < nx * nx; idx += DT::tpe) {
    const int i = idx / nx, j = idx % nx;
    R a = 0, b = 0, c = 0;
    for (int m = 0; m < nu; ++m) {
      a += s.Kt[m * nx + i] * s.QK[m * nx + j];
      b += s.Kt[m * nx + i] * s.Qxu[j * nu + m];
      c += s.Qxu[i * nu + m] * s.Kt[m * nx + j];
    }
    s.Vn[idx] = ((s.Qxx[idx] + a) + b) + c;
  }
  }
  for (int i = tid; i < nx; i += DT::tpe) {
    R a = 0, b = 0, c = 0;
    for (int m = 0; m < nu; ++m) {
      a += s.Kt[m * nx + i] * s.Qk[m];
      b += s.Kt[m * nx + i] * s.qu[m];
      c += s.Qxu[i * nu + m] * s.kt[m];
    }
    s.vn[i] = ((s.qx[i] + a) + b) + c;
  }
  DT::sync();
  for (int idx = tid; idx < nx * nx; idx += DT::tpe) s.V[idx] = s.Vn[idx];
  for (int i = tid; i < nx; i += DT::tpe) s.v[i] = s.vn[i];
  DT::sync();
  return nonpd;
}

}  // namespace tacd
