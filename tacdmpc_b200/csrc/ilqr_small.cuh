// ilqr_small.cuh — whole-solve kernels for small systems (nx <= ~6, nu <= 2):
// ONE THREAD PER BATCH ELEMENT, persistent over a work queue.
//
// Layout: the per-element trajectory state that must survive
// between passes — nominal X[(T+1)nx], U[T nu] and the gains K[T nu nx], k[T nu]
// — lives in shared memory as [word][thread] (bank = thread, conflict-free);
// every per-step matrix (V, Q*, A, B, C_t) lives in registers with NX/NU known at
// compile time.  A lane that finishes its element immediately pulls the next one
// from a global counter, so the spread in iLQR iteration counts (2..40) does not
// idle lanes for a whole solve, only for the short init/finalise phases.
//
// Forward = solve_step (controller.py:275-315): per iteration a reverse Riccati
// pass (quadraticize + linearize + backward_lqr fused), one forward pass that
// rolls out ALL line-search candidates at once (evaluate_alphas) and an in-place
// re-rollout of the accepted candidate.
#pragma once
#include "tacd_dev.cuh"

namespace tacd {

template <typename R>
struct FwdPtrs {
  const R *x0, *C, *c, *Cf, *cf, *C_ls, *c_ls, *Cf_ls, *cf_ls, *xref, *uref, *Uinit;
  const int32_t* replay_iters;
  const uint8_t* replay_alpha;
  const uint8_t* replay_flags;
  R *X, *U, *A, *Bm;
  uint8_t* tight;
  int32_t* iters;
  uint8_t* alpha_trace;
  uint8_t* flags;
  R* cost;
  R* cost_trace;
  unsigned int* counter;  // work queue head; set to first_wave before launch
  const int32_t* order;   // longest-first permutation of the batch (nullable = identity)
  R* xscratch;            // five-lanes-per-element kernel: the candidates' states, one [T nx][32] tile per resident warp
  int stage_words;        // five-lanes-per-element kernel: > 0 = per-element (C, c) are bulk-copied into shared memory, this many words per element
  R* cscratch;            // per-element costs only: lane-slot copy of (C, c), [word][lane] (nullable)
  size_t cstride;         // number of lane slots (= grid * block)
  unsigned int first_wave;  // queue positions [0, first_wave) are dealt statically, one 32-segment per warp
  // Two-phase solve (ilqr_group_launch.cuh::run_forward_hybrid): the thread-per-element kernel hands a solve that is still
  // improving after cap_iters iterations over to the five-lanes-per-element kernel, which resumes it.
  int cap_iters;                 // thread kernel: > 0 = hand over after this many iterations
  unsigned int* resume_count;    // thread kernel: number of handed-over elements (device counter)
  int32_t* resume_list;          // thread kernel: their batch indices
  R* resume_cost;                // thread kernel: best cost of every handed-over element, [B]
  const unsigned int* n_dev;     // five-lane kernel: number of queue entries, read on the device (nullable = P.B)
  int resume;                    // five-lane kernel: elements continue from (io.X, io.Uinit, cost0, io.iters, io.flags)
  const int* cost_mode;          // five-lane kernel, shared costs: the mode group_cost_mode wrote (nullable); a variant whose CM differs exits
};

// ---------------------------------------------------------------------------
// Longest-first ordering of the work queue.  One element is one serial chain of iLQR iterations
// (2..40 of them), so a kernel cannot end before its longest chain; if a long chain is popped late the
// whole GPU waits for it (measured: SMs 75 % active, 20.6/32 lanes, profiles/r1_forward_v3).  The
// iteration count correlates (Spearman 0.65) with the initial stage cost tau_0' C_0 tau_0, so the queue is
// sorted by that key, descending, with a 2048-bucket counting sort on the float's exponent + 3 mantissa
// bits: histogram -> exclusive scan -> scatter.  Any permutation is a correct queue order; the key only
// affects load balance.  A negative n = -(nx+nu) marks a compact diagonal cost (tacd_problem.C_diag).
constexpr int kLptBuckets = 2048;

template <typename R>
__device__ __forceinline__ int lpt_bucket(const R* x0, const R* xr0, const R* C0, int nx, int n) {
  float q = 0.f;
  if (n < 0) {   // compact diagonal cost (tacd_problem.C_diag): C0 holds the diagonal
    for (int i = 0; i < nx; ++i) q += (float)C0[i] * (float)(x0[i] - xr0[i]) * (float)(x0[i] - xr0[i]);
  } else {
    for (int i = 0; i < nx; ++i) {
      float s = 0.f;
      for (int j = 0; j < nx; ++j) s += (float)C0[i * n + j] * (float)(x0[j] - xr0[j]);
      q += (float)(x0[i] - xr0[i]) * s;
    }
  }
  if (!(q > 0.f)) return kLptBuckets - 1;               // zero / negative / NaN keys go last
  const unsigned u = __float_as_uint(q) >> 20;           // 8 exponent + 3 mantissa bits (sign is 0)
  const int bkt = (int)(u < (unsigned)kLptBuckets ? u : kLptBuckets - 1);
  return kLptBuckets - 1 - bkt;                          // descending
}

template <typename R>
__global__ void lpt_histogram(int B, int T, int nx, int n, int C_batched, int xref_batched, const R* x0, const R* xref,
                              const R* C, unsigned int* hist) {
  __shared__ unsigned int sh[kLptBuckets];
  for (int i = threadIdx.x; i < kLptBuckets; i += blockDim.x) sh[i] = 0;
  __syncthreads();
  for (int b = blockIdx.x * blockDim.x + threadIdx.x; b < B; b += gridDim.x * blockDim.x) {
    const int k = lpt_bucket<R>(x0 + (size_t)b * nx, xref + (xref_batched ? (size_t)b * (T + 1) * nx : 0),
                                C + (C_batched ? (size_t)b * T * (n < 0 ? -n : n * n) : 0), nx, n);
    atomicAdd(&sh[k], 1u);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < kLptBuckets; i += blockDim.x)
    if (sh[i]) atomicAdd(&hist[i], sh[i]);
}
static __global__ void queue_init(unsigned int* counter, unsigned int v) { *counter = v; }
// exclusive scan of the 2048 counters by one CTA of 1024 threads (2 per thread)
static __global__ void lpt_scan(unsigned int* hist, unsigned int* counter = nullptr, unsigned int counter0 = 0) {
  __shared__ unsigned int part[1024];
  const int t = threadIdx.x;
  if (counter && t == 0) *counter = counter0;   // (the work-queue head is reset here: one launch fewer)
  const unsigned a = hist[2 * t], b = hist[2 * t + 1];
  part[t] = a + b;
  __syncthreads();
  for (int o = 1; o < 1024; o <<= 1) {
    const unsigned v = (t >= o) ? part[t - o] : 0u;
    __syncthreads();
    part[t] += v;
    __syncthreads();
  }
  const unsigned excl = part[t] - (a + b);
  hist[2 * t] = excl;
  hist[2 * t + 1] = excl + a;
}
template <typename R>
__global__ void lpt_scatter(int B, int T, int nx, int n, int C_batched, int xref_batched, const R* x0, const R* xref,
                            const R* C, unsigned int* offs, int32_t* order) {
  for (int b = blockIdx.x * blockDim.x + threadIdx.x; b < B; b += gridDim.x * blockDim.x) {
    const int k = lpt_bucket<R>(x0 + (size_t)b * nx, xref + (xref_batched ? (size_t)b * (T + 1) * nx : 0),
                                C + (C_batched ? (size_t)b * T * (n < 0 ? -n : n * n) : 0), nx, n);
    // warp-aggregated: the keys of a batch cluster in a few dozen buckets, so per-lane atomics on `offs`
    // serialise (17.5 us at B = 65536); one atomic per distinct bucket per warp instead
    const unsigned act = __activemask();
    const unsigned peers = __match_any_sync(act, k);
    const int leader = __ffs(peers) - 1, lane = threadIdx.x & 31;
    unsigned base = 0;
    if (lane == leader) base = atomicAdd(&offs[k], (unsigned)__popc(peers));
    base = __shfl_sync(peers, base, leader);
    order[base + __popc(peers & ((1u << lane) - 1u))] = b;
  }
}

template <typename R>
struct BwdPtrs {
  const R *X, *U, *C, *Cf, *xref, *uref, *A, *Bm, *gX, *gU;
  const uint8_t* tight;
  R *grad_C, *grad_c, *grad_Cf, *grad_cf, *grad_x0, *grad_xref, *grad_uref, *dX, *dU;
  // thread-per-element backward: the gains K_t, k_t of the KKT pass, [word][thread slot] in the caller's workspace (L2-resident,
  // coalesced) instead of 400 bytes of shared memory per thread; nullptr = shared memory
  R* kscratch;
};

constexpr int kNA = 5;  // line-search candidates rolled out together (the reference has 5 alphas)

// words of shared memory per thread for the forward state
__host__ __device__ inline int fwd_words(int T, int nx, int nu) { return (T + 1) * nx + T * nu + T * nu * nx + T * nu; }
__host__ __device__ inline int bwd_words(int T, int nx, int nu) { return T * nu * nx + T * nu; }

// tau^T C tau and c^T tau with C [N,N] row-major in registers
template <typename R, int N>
TACD_DEV void quad_terms(const R (&Ct)[N * N], const R (&ct)[N], const R (&tau)[N], R& quad, R& lin) {
#pragma unroll
  for (int i = 0; i < N; ++i) {
    R s = 0;
#pragma unroll
    for (int j = 0; j < N; ++j) s += Ct[i * N + j] * tau[j];
    quad += tau[i] * s;
    lin += ct[i] * tau[i];
  }
}

template <typename R, int NX, int NU, int DYN, bool LSSEP>
__global__ void __launch_bounds__(256, 1) ilqr_forward_small(const DevProblem<R> P, const FwdPtrs<R> io) {
  constexpr int N = NX + NU;
  using D = Dyn<R, NX, NU, DYN>;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int NT = blockDim.x, tid = threadIdx.x;
  const int T = P.T;
  R* sX = reinterpret_cast<R*>(smem_raw);
  R* sU = sX + (size_t)(T + 1) * NX * NT;
  R* sK = sU + (size_t)T * NU * NT;
  R* sk = sK + (size_t)T * NU * NX * NT;
  R* sA = sk + (size_t)T * NU * NT;  // LTI matrices, shared by the CTA
  R* sB = sA + NX * NX;
  if (DYN == TACD_DYN_LTI) {
    for (int i = tid; i < NX * NX; i += NT) sA[i] = P.dyn_A[i];
    for (int i = tid; i < NX * NU; i += NT) sB[i] = P.dyn_B[i];
    __syncthreads();
  }
#define SX(t, i) sX[((size_t)(t) * NX + (i)) * NT + tid]
#define SU(t, i) sU[((size_t)(t) * NU + (i)) * NT + tid]
#define SK(t, i) sK[((size_t)(t) * NU * NX + (i)) * NT + tid]
#define Sk(t, i) sk[((size_t)(t) * NU + (i)) * NT + tid]

  const size_t szX = (size_t)(T + 1) * NX, szU = (size_t)T * NU;
  // bounds with +-inf when absent: max(u, -inf) = u and min(u, +inf) = u (NaN still propagates), so the
  // inner loops carry no uniform branch
  R ulo[NU], uhi[NU];
#pragma unroll
  for (int i = 0; i < NU; ++i) {
    ulo[i] = P.has_umin ? P.umin[i] : -inf_<R>();
    uhi[i] = P.has_umax ? P.umax[i] : inf_<R>();
  }
  bool has = false, exhausted = false, first = true;
  int b = 0, it = 0;
  unsigned flg = 0;
  R best = 0;
  const R *Cp = nullptr, *cp = nullptr, *Cfp = nullptr, *cfp = nullptr, *xr = nullptr, *ur = nullptr;
  size_t cs = 1;  // element stride of Cp / cp (1 = caller's tensor, cstride = lane-slot scratch)

  while (true) {
    // ---------------------------------------------------------------- refill
    if (!has && !exhausted) {
      // first element of a lane: dealt statically so that the 32-segments of the longest-first order are
      // spread round-robin over the CTAs (the longest chains end up one warp per SM, not eight on one SM);
      // afterwards: the shared counter
      unsigned nb;
      if (first) {
        first = false;
        nb = ((unsigned)blockIdx.x + (unsigned)gridDim.x * (unsigned)(tid >> 5)) * 32u + (unsigned)(tid & 31);
        if (nb >= io.first_wave) nb = atomicAdd(io.counter, 1u);
      } else {
        nb = atomicAdd(io.counter, 1u);
      }
      if (nb < (unsigned)P.B) {
        b = io.order ? io.order[nb] : (int)nb;
        has = true;
        it = 0;
        flg = 0;
        Cp = io.C + (P.C_batched ? (size_t)b * T * N * N : 0);
        cp = io.c + (P.C_batched ? (size_t)b * T * N : 0);
        if (LSSEP && P.C_batched && io.cscratch != nullptr) {
          // Per-element costs: a thread reading its own element's C_t walks memory with a 4*T*n*n-byte
          // stride across the warp — 32 sectors per load instruction, and C is re-read twice per
          // iteration.  Copy it ONCE into this lane's slot of a [word][lane] scratch; every later read
          // is then one coalesced 128-byte wavefront (and mostly an L2 hit).
          const size_t gl = (size_t)blockIdx.x * NT + tid;
          R* dC = io.cscratch + gl;
          R* dc = io.cscratch + (size_t)T * N * N * io.cstride + gl;
          const int nC = T * N * N, nc = T * N;
#pragma unroll 8
          for (int w = 0; w < nC; ++w) dC[(size_t)w * io.cstride] = Cp[w];
#pragma unroll 5
          for (int w = 0; w < nc; ++w) dc[(size_t)w * io.cstride] = cp[w];
          Cp = dC;
          cp = dc;
          cs = io.cstride;
        } else {
          cs = 1;
        }
        Cfp = io.Cf + (P.Cf_batched ? (size_t)b * N * N : 0);
        cfp = io.cf + (P.Cf_batched ? (size_t)b * N : 0);
        xr = io.xref + (P.xref_batched ? (size_t)b * szX : 0);
        ur = io.uref + (P.uref_batched ? (size_t)b * szU : 0);
        // rollout_trajectory (controller.py:361-398, no clamping) + objective (cost.py:37-62)
        R x[NX], u[NU], xn[NX];
        R quad = 0, lin = 0;
#pragma unroll
        for (int i = 0; i < NX; ++i) { x[i] = io.x0[(size_t)b * NX + i]; SX(0, i) = x[i]; }
        for (int t = 0; t < T; ++t) {
          R tau[N], Ct[N * N], ct[N];
#pragma unroll
          for (int i = 0; i < NU; ++i) { u[i] = io.Uinit[(size_t)b * szU + t * NU + i]; SU(t, i) = u[i]; }
#pragma unroll
          for (int i = 0; i < NX; ++i) tau[i] = x[i] - xr[t * NX + i];
#pragma unroll
          for (int i = 0; i < NU; ++i) tau[NX + i] = u[i] - ur[t * NU + i];
#pragma unroll
          for (int i = 0; i < N * N; ++i) Ct[i] = Cp[((size_t)t * N * N + i) * cs];
#pragma unroll
          for (int i = 0; i < N; ++i) ct[i] = cp[((size_t)t * N + i) * cs];
          quad_terms<R, N>(Ct, ct, tau, quad, lin);
          D::step(P, sA, sB, x, u, xn);
#pragma unroll
          for (int i = 0; i < NX; ++i) { x[i] = xn[i]; SX(t + 1, i) = x[i]; }
        }
        R qN = 0, lN = 0;
        {
          R tauN[NX];
#pragma unroll
          for (int i = 0; i < NX; ++i) tauN[i] = x[i] - xr[T * NX + i];
#pragma unroll
          for (int i = 0; i < NX; ++i) {
            R s = 0;
#pragma unroll
            for (int j = 0; j < NX; ++j) s += Cfp[i * N + j] * tauN[j];
            qN += tauN[i] * s;
            lN += cfp[i] * tauN[i];
          }
        }
        best = ((R)0.5 * quad + lin) + ((R)0.5 * qN + lN);
        if (io.alpha_trace)
          for (int j = 0; j < P.max_iter; ++j) io.alpha_trace[(size_t)b * P.max_iter + j] = 0xFF;
      } else {
        exhausted = true;
      }
    }
    if (__ballot_sync(0xffffffffu, has) == 0u) break;
    if (!has) continue;

    bool stop = true;
    const bool replay = io.replay_iters != nullptr;
    if (it < P.max_iter && !(replay && it >= io.replay_iters[b])) {
      // ------------------------------------------------ reverse Riccati pass
      // quadraticize (cost.py:64-100) + linearize (controller.py:448-471) +
      // backward_lqr (controller.py:546-578), one time step at a time
      {
        R V[NX * NX], v[NX];
        {
          R tauN[NX];
#pragma unroll
          for (int i = 0; i < NX; ++i) tauN[i] = SX(T, i) - xr[T * NX + i];
#pragma unroll
          for (int i = 0; i < NX; ++i) {
            R s = 0;
#pragma unroll
            for (int j = 0; j < NX; ++j) { const R cij = Cfp[i * N + j]; V[i * NX + j] = cij; s += cij * tauN[j]; }
            v[i] = s + cfp[i];
          }
        }
        for (int t = T - 1; t >= 0; --t) {
          R x[NX], u[NU], tau[N], Ct[N * N];
          R Qxx[NX * NX], Qxu[NX * NU], Quu[NU * NU], qx[NX], qu[NU];
#pragma unroll
          for (int i = 0; i < NX; ++i) { x[i] = SX(t, i); tau[i] = x[i] - xr[t * NX + i]; }
#pragma unroll
          for (int i = 0; i < NU; ++i) { u[i] = SU(t, i); tau[NX + i] = u[i] - ur[t * NU + i]; }
#pragma unroll
          for (int i = 0; i < N * N; ++i) Ct[i] = Cp[((size_t)t * N * N + i) * cs];
#pragma unroll
          for (int i = 0; i < N; ++i) {
            R s = 0;
#pragma unroll
            for (int j = 0; j < N; ++j) s += Ct[i * N + j] * tau[j];
            s += cp[((size_t)t * N + i) * cs];
            if (i < NX) qx[i] = s; else qu[i - NX] = s;
          }
#pragma unroll
          for (int i = 0; i < NX; ++i) {
#pragma unroll
            for (int j = 0; j < NX; ++j) Qxx[i * NX + j] = Ct[i * N + j];
#pragma unroll
            for (int j = 0; j < NU; ++j) Qxu[i * NU + j] = Ct[i * N + NX + j];
          }
#pragma unroll
          for (int i = 0; i < NU; ++i)
#pragma unroll
            for (int j = 0; j < NU; ++j) Quu[i * NU + j] = Ct[(NX + i) * N + NX + j];
          R A[NX * NX], Bm[NX * NU];
          D::jac(P, sA, sB, x, u, A, Bm);
          q_blocks<R, NX, NU>(P.reg, A, Bm, V, v, Qxx, Qxu, Quu, qx, qu);
          R Kt[NU * NX], kt[NU];
          if (gains<R, NX, NU>(P.reg, Qxu, Quu, qu, Kt, kt)) flg |= TACD_FLAG_NONPD;
#pragma unroll
          for (int i = 0; i < NU * NX; ++i) SK(t, i) = Kt[i];
#pragma unroll
          for (int i = 0; i < NU; ++i) Sk(t, i) = kt[i];
          value_update<R, NX, NU>(Qxx, Qxu, Quu, qx, qu, Kt, kt, V, v);
        }
      }
      // -------------------------------------- all line-search candidates at once
      // evaluate_alphas (controller.py:589-620).  Candidate costs use the line-search
      // cost (element 0's when costs are per-element, quirk Q4); with LSSEP the
      // element's own cost of every candidate is accumulated alongside, so the
      // acceptance cost (controller.py:298) needs no extra pass.
      R costs[TACD_MAX_ALPHA], own[TACD_MAX_ALPHA];
      const R* Cl = LSSEP ? io.C_ls : Cp;
      const R* cl = LSSEP ? io.c_ls : cp;
      const R* Cfl = LSSEP ? io.Cf_ls : Cfp;
      const R* cfl = LSSEP ? io.cf_ls : cfp;
      for (int a0 = 0; a0 < P.n_alpha; a0 += kNA) {
        R xa[kNA][NX], q1[kNA], l1[kNA], q2[kNA], l2[kNA], al[kNA];
#pragma unroll
        for (int a = 0; a < kNA; ++a) {
          al[a] = P.alphas[(a0 + a < P.n_alpha) ? a0 + a : P.n_alpha - 1];
          q1[a] = l1[a] = q2[a] = l2[a] = 0;
#pragma unroll
          for (int i = 0; i < NX; ++i) xa[a][i] = SX(0, i);
        }
        for (int t = 0; t < T; ++t) {
          R Xt[NX], Ut[NU], Kt[NU * NX], kt[NU], xrt[NX], urt[NU], Ct[N * N], ct[N];
#pragma unroll
          for (int i = 0; i < NX; ++i) { Xt[i] = SX(t, i); xrt[i] = xr[t * NX + i]; }
#pragma unroll
          for (int i = 0; i < NU; ++i) { Ut[i] = SU(t, i); kt[i] = Sk(t, i); urt[i] = ur[t * NU + i]; }
#pragma unroll
          for (int i = 0; i < NU * NX; ++i) Kt[i] = SK(t, i);
#pragma unroll
          for (int i = 0; i < N * N; ++i) Ct[i] = Cl[(size_t)t * N * N + i];
#pragma unroll
          for (int i = 0; i < N; ++i) ct[i] = cl[(size_t)t * N + i];
          R ua[kNA][NU];
#pragma unroll
          for (int a = 0; a < kNA; ++a) {
            R tau[N];
#pragma unroll
            for (int i = 0; i < NU; ++i) {
              R s = 0;
#pragma unroll
              for (int j = 0; j < NX; ++j) s += Kt[i * NX + j] * (xa[a][j] - Xt[j]);
              R ui = Ut[i] + (al[a] * kt[i] + s);
              ui = tmin<R>(tmax<R>(ui, ulo[i]), uhi[i]);
              ua[a][i] = ui;
              tau[NX + i] = ui - urt[i];
            }
#pragma unroll
            for (int i = 0; i < NX; ++i) tau[i] = xa[a][i] - xrt[i];
            quad_terms<R, N>(Ct, ct, tau, q1[a], l1[a]);
          }
          if (LSSEP) {
#pragma unroll
            for (int i = 0; i < N * N; ++i) Ct[i] = Cp[((size_t)t * N * N + i) * cs];
#pragma unroll
            for (int i = 0; i < N; ++i) ct[i] = cp[((size_t)t * N + i) * cs];
#pragma unroll
            for (int a = 0; a < kNA; ++a) {
              R tau[N];
#pragma unroll
              for (int i = 0; i < NX; ++i) tau[i] = xa[a][i] - xrt[i];
#pragma unroll
              for (int i = 0; i < NU; ++i) tau[NX + i] = ua[a][i] - urt[i];
              quad_terms<R, N>(Ct, ct, tau, q2[a], l2[a]);
            }
          }
#pragma unroll
          for (int a = 0; a < kNA; ++a) {
            R xn[NX];
            D::step(P, sA, sB, xa[a], ua[a], xn);
#pragma unroll
            for (int i = 0; i < NX; ++i) xa[a][i] = xn[i];
          }
        }
#pragma unroll
        for (int a = 0; a < kNA; ++a) {
          R tauN[NX];
#pragma unroll
          for (int i = 0; i < NX; ++i) tauN[i] = xa[a][i] - xr[T * NX + i];
          R qN = 0, lN = 0, qN2 = 0, lN2 = 0;
#pragma unroll
          for (int i = 0; i < NX; ++i) {
            R s = 0, s2 = 0;
#pragma unroll
            for (int j = 0; j < NX; ++j) {
              s += Cfl[i * N + j] * tauN[j];
              if (LSSEP) s2 += Cfp[i * N + j] * tauN[j];
            }
            qN += tauN[i] * s;
            lN += cfl[i] * tauN[i];
            if (LSSEP) { qN2 += tauN[i] * s2; lN2 += cfp[i] * tauN[i]; }
          }
          if (a0 + a < P.n_alpha) {
            costs[a0 + a] = ((R)0.5 * q1[a] + l1[a]) + ((R)0.5 * qN + lN);
            own[a0 + a] = LSSEP ? ((R)0.5 * q2[a] + l2[a]) + ((R)0.5 * qN2 + lN2) : costs[a0 + a];
          }
        }
      }
      // ------------------------------------------------------------- accept
      // argmin (first minimum, NaN wins) + strict improvement (controller.py:292-301)
      int ai = 0;
      for (int a = 1; a < P.n_alpha; ++a) {
        if (costs[ai] != costs[ai]) break;
        if (costs[a] < costs[ai] || costs[a] != costs[a]) ai = a;
      }
      if (replay) ai = io.replay_alpha[(size_t)b * P.max_iter + it];
      R newc = own[0];
      for (int a = 1; a < P.n_alpha; ++a) newc = (a == ai) ? own[a] : newc;
      bool improved = newc < best;
      if (replay) improved = (it < io.replay_iters[b] - 1) || ((io.replay_flags[b] & TACD_FLAG_MAXITER) != 0);
      if (io.alpha_trace) io.alpha_trace[(size_t)b * P.max_iter + it] = (uint8_t)ai;
      if (io.cost_trace) {
        R* ctr = io.cost_trace + ((size_t)b * P.max_iter + it) * (P.n_alpha + 2);
        for (int a = 0; a < P.n_alpha; ++a) ctr[a] = costs[a];
        ctr[P.n_alpha] = newc;
        ctr[P.n_alpha + 1] = best;
      }
      it += 1;
      if (improved) {
        best = newc;
        // re-rollout of the accepted candidate, in place (X_{t+1} is read before it is overwritten)
        R alpha = P.alphas[0];
        for (int a = 1; a < P.n_alpha; ++a) alpha = (a == ai) ? P.alphas[a] : alpha;
        R x[NX], Xold[NX];
#pragma unroll
        for (int i = 0; i < NX; ++i) { x[i] = SX(0, i); Xold[i] = x[i]; }
        for (int t = 0; t < T; ++t) {
          R u[NU], xn[NX];
#pragma unroll
          for (int i = 0; i < NU; ++i) {
            R s = 0;
#pragma unroll
            for (int j = 0; j < NX; ++j) s += SK(t, i * NX + j) * (x[j] - Xold[j]);
            R ui = SU(t, i) + (alpha * Sk(t, i) + s);
            ui = tmin<R>(tmax<R>(ui, ulo[i]), uhi[i]);
            u[i] = ui;
          }
#pragma unroll
          for (int i = 0; i < NU; ++i) SU(t, i) = u[i];
          D::step(P, sA, sB, x, u, xn);
#pragma unroll
          for (int i = 0; i < NX; ++i) { Xold[i] = SX(t + 1, i); x[i] = xn[i]; SX(t + 1, i) = xn[i]; }
        }
        stop = (it >= P.max_iter);
        if (stop) flg |= TACD_FLAG_MAXITER;
        if (!stop && io.cap_iters > 0 && it >= io.cap_iters) {
          // Hand-over: this solve is in the long tail of the iteration-count distribution.  A thread's iteration is one
          // serial chain of ~26 k instructions, so a kernel that kept it could not end before max(iters) such chains
          // (round 1: 0.68 of 0.85 ms); the five-lanes-per-element kernel resumes it with a 5x shorter chain.
          for (int t = 0; t <= T; ++t)
#pragma unroll
            for (int i = 0; i < NX; ++i) io.X[(size_t)b * szX + t * NX + i] = SX(t, i);
          for (int t = 0; t < T; ++t)
#pragma unroll
            for (int i = 0; i < NU; ++i) io.U[(size_t)b * szU + t * NU + i] = SU(t, i);
          io.iters[b] = it;
          io.flags[b] = (uint8_t)flg;
          io.resume_cost[b] = best;
          io.resume_list[atomicAdd(io.resume_count, 1u)] = b;
          has = false;
        }
      }
    }
    if (stop) {
      // ------------------------------------------------------------ finalise
      // outputs + final re-linearisation (controller.py:311-314) + active set (636-642)
      for (int t = 0; t <= T; ++t)
#pragma unroll
        for (int i = 0; i < NX; ++i) io.X[(size_t)b * szX + t * NX + i] = SX(t, i);
      for (int t = 0; t < T; ++t) {
#pragma unroll
        for (int i = 0; i < NU; ++i) {
          const R ui = SU(t, i);
          io.U[(size_t)b * szU + t * NU + i] = ui;
          bool m = false;
          if (P.has_umin) m = m || isclose_<R>(ui, P.umin[i]);
          if (P.has_umax) m = m || isclose_<R>(ui, P.umax[i]);
          io.tight[(size_t)b * szU + t * NU + i] = m ? 1 : 0;
        }
        if (io.A != nullptr || io.Bm != nullptr) {
          R x[NX], u[NU], A[NX * NX], Bm[NX * NU];
#pragma unroll
          for (int i = 0; i < NX; ++i) x[i] = SX(t, i);
#pragma unroll
          for (int i = 0; i < NU; ++i) u[i] = SU(t, i);
          D::jac(P, sA, sB, x, u, A, Bm);
          if (io.A)
#pragma unroll
            for (int i = 0; i < NX * NX; ++i) io.A[((size_t)b * T + t) * NX * NX + i] = A[i];
          if (io.Bm)
#pragma unroll
            for (int i = 0; i < NX * NU; ++i) io.Bm[((size_t)b * T + t) * NX * NU + i] = Bm[i];
        }
      }
      io.iters[b] = it;
      io.flags[b] = (uint8_t)flg;
      if (io.cost) io.cost[b] = best;
      has = false;
    }
  }
#undef SX
#undef SU
#undef SK
#undef Sk
}

// ---------------------------------------------------------------------------
// Backward: ILQRSolve.backward + _zero_constrained_lqr (controller.py:63-115, 707-788).
// One thread per element (static assignment, uniform work); gains K_t,k_t in shared memory
// [word][thread]; A_t,B_t are recomputed from (X*,U*) with the embedded dynamics (or read, for
// TACD_DYN_EXTERNAL).  Per-step inputs (X,U,gX,gU,tight / refs) are prefetched one step ahead so the
// L2 latency of the strided per-element loads overlaps the Riccati arithmetic.
// Gradients of SHARED (un-batched) inputs are sums over the batch: each warp transposes its 32 x NV
// per-step values through a shared-memory tile, lanes sum columns into per-warp accumulators, the CTA
// combines its warps and writes one partial vector; reduce_shared_grads() sums the CTAs in a fixed
// order (deterministic, no atomics).
template <int NX, int NU>
struct BwdLayout {
  static constexpr int N = NX + NU;
  static constexpr int NV = N * N + N + NX + NU;    // per step: grad_C | grad_c | grad_xref | grad_uref
  static constexpr int NVF = NX * NX + NX + NX;     // terminal: grad_Cf[:nx,:nx] | grad_cf[:nx] | grad_xref[T]
  // Tile row of one element and step: grad_C is symmetric, so only its upper triangle goes through the tile —
  // NS = n(n+1)/2 columns instead of n*n; with grad_c, grad_xref, grad_uref that is NP = 25 columns for n = 5 (<= 32: every
  // column of the transposed reduction has its own lane, ONE pass of 32 row reads instead of two).  The accumulators /
  // per-CTA partials keep the dense NV layout (reduce_shared_grads and the five-lane backward share it).
  static constexpr int NS = N * (N + 1) / 2;
  static constexpr int NP = NS + N + NX + NU;
  __host__ __device__ static constexpr int pk(int i, int k) {   // packed column of entry (i, k) = (k, i)
    return (i <= k) ? i * N - i * (i - 1) / 2 + (k - i) : k * N - k * (k - 1) / 2 + (i - k);
  }
  static constexpr int TW = NV | 1;                 // odd tile row stride: conflict-free transposed access
  static constexpr int CW = (N * N) | 1;            // row stride of the two C_t prefetch buffers
  static constexpr int TILE = (32 * TW > 64 * CW) ? 32 * TW : 64 * CW;   // words per warp
  __host__ __device__ static int nacc(int T) { return T * NV + NVF; }
};

// NX consecutive words of a per-element trajectory row: one 128-bit load when NX = 4 and the row is 16-byte aligned (every
// row of a [B, T+1, 4] tensor is, if its base is), else NX scalar loads.  The backward reads four such rows per step and
// thread at a stride of one trajectory between lanes — its LSU instruction queue was a stall reason (lg_throttle 0.64 warps
// per issue, profiles/r2_backward_small_final).
template <typename R, int NX>
TACD_DEV void ld_row(const R* p, bool vec, R (&d)[NX]) {
  if constexpr (NX == 4 && sizeof(R) == 4) {
    if (vec) {
      const float4 v = *reinterpret_cast<const float4*>(p);
      d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
      return;
    }
  }
#pragma unroll
  for (int i = 0; i < NX; ++i) d[i] = p[i];
}

template <typename R, int NX, int NU, int DYN, bool DIAG, bool EXACT>
__global__ void __launch_bounds__(256, 2) ilqr_backward_small(const DevProblem<R> P, const BwdPtrs<R> io, R* partials, int reduce, int vnt) {
  constexpr int N = NX + NU;
  using L = BwdLayout<NX, NU>;
  using D = Dyn<R, NX, NU, (DYN == TACD_DYN_EXTERNAL ? TACD_DYN_LTI : DYN)>;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int NT = blockDim.x, tid = threadIdx.x;
  const int T = P.T;
  const int lane = tid & 31, wid = tid >> 5, nw = NT >> 5;
  // Gains of the KKT pass: one [word][slot] column per thread.  Round 1 kept them in shared memory (400 bytes per thread for the
  // cart-pole, T = 20), which with the reduction tiles left ONE 8-warp CTA per SM — the kernel issued 38 % of its slots and waited
  // on latencies (profiles/r1_backward_v6).  In the workspace the column is a coalesced, L2-resident stream and shared memory
  // holds only the tiles: two CTAs (16 warps) per SM at 128 registers.
  const bool kglob = io.kscratch != nullptr;
  const size_t kst = kglob ? (size_t)gridDim.x * NT : (size_t)NT;
  R* smem0 = reinterpret_cast<R*>(smem_raw);
  R* sK = kglob ? io.kscratch + (size_t)blockIdx.x * NT + tid : smem0 + tid;
  R* sk = sK + (size_t)T * NU * NX * kst;
  R* sA = kglob ? smem0 : smem0 + (size_t)T * NU * (NX + 1) * NT;
  R* sB = sA + NX * NX;
  R* tile = sB + NX * NU;                                    // [nw][32][TW]      (only when reduce)
  R* acc = tile + (size_t)nw * L::TILE;                      // [nw][nacc]
  const int NACC = L::nacc(T);
  if (DYN == TACD_DYN_LTI) {
    for (int i = tid; i < NX * NX; i += NT) sA[i] = P.dyn_A[i];
    for (int i = tid; i < NX * NU; i += NT) sB[i] = P.dyn_B[i];
  }
  if (reduce)
    for (int i = tid; i < nw * NACC; i += NT) acc[i] = 0;
  __syncthreads();
  R* mytile = tile + (size_t)wid * L::TILE;
  R* myacc = acc + (size_t)wid * NACC;
  // packed tile column j = lane + 32 q -> its dense accumulator slot(s): (i, k) and (k, i) of grad_C, or the vector entries
  // (dense per-element grad_C is stored straight from the tile rows: that layout keeps the full matrix in the tile)
  const bool packC = !(io.grad_C && P.C_batched && !DIAG);
  const int vb = packC ? L::NS : N * N;          // first column of the vector entries
  const int ncol = vb + N + NX + NU;
  constexpr int NCOLIT = (L::NV + 31) / 32;
  int cd0[NCOLIT], cd1[NCOLIT];
#pragma unroll
  for (int q = 0; q < NCOLIT; ++q) {
    const int j = lane + 32 * q;
    cd0[q] = packC ? N * N + (j - L::NS) : j;
    cd1[q] = -1;
    if (packC && j < L::NS) {
      int i = 0, rem = j;
      while (rem >= N - i) { rem -= N - i; ++i; }
      const int k = i + rem;
      cd0[q] = i * N + k;
      cd1[q] = (k != i) ? k * N + i : -1;
    }
  }
  const bool redC = reduce && io.grad_C && !P.C_batched, redc = reduce && io.grad_c && !P.C_batched;
  const bool redCf = reduce && io.grad_Cf && !P.Cf_batched, redcf = reduce && io.grad_cf && !P.Cf_batched;
  const bool redx = reduce && io.grad_xref && !P.xref_batched, redu = reduce && io.grad_uref && !P.uref_batched;
#define SK(t, i) sK[((size_t)(t) * NU * NX + (i)) * kst]
#define Sk(t, i) sk[((size_t)(t) * NU + (i)) * kst]
  const size_t szX = (size_t)(T + 1) * NX, szU = (size_t)T * NU;
  const bool vecX = NX == 4 && sizeof(R) == 4 && (((uintptr_t)io.X | (uintptr_t)io.gX | (uintptr_t)io.xref) & 15) == 0;
  // vnt (multiple of 32, <= NT) threads of every CTA take one element per wave, so that all CTAs — and with
  // them all SMs — carry the same load in every wave; the remaining warps of the CTA stay out of the loop
  const int stride = gridDim.x * vnt;
  for (int b0 = blockIdx.x * vnt + (tid - lane); b0 < P.B && tid - lane < vnt; b0 += stride) {
    const int b = b0 + lane;
    const bool live = b < P.B;
    const int bb = live ? b : P.B - 1;  // dead lanes shadow the last element, outputs masked
    const R* X = io.X + (size_t)bb * szX;
    const R* U = io.U + (size_t)bb * szU;
    // tacd_problem.C_diag: C[B,T,n] holds the diagonals (ActorMPC layout), gradients come back compact too
    const R* Cp = io.C + (P.C_batched ? (size_t)bb * T * (DIAG ? N : N * N) : 0);
    const R* xr = io.xref + (P.xref_batched ? (size_t)bb * szX : 0);
    const R* ur = io.uref + (P.uref_batched ? (size_t)bb * szU : 0);
    const R* gX = io.gX + (size_t)bb * szX;
    const R* gU = io.gU + (size_t)bb * szU;
    const uint8_t* tg = io.tight + (size_t)bb * szU;

    auto load_AB = [&](int t, const R(&x)[NX], const R(&u)[NU], R(&A)[NX * NX], R(&Bm)[NX * NU]) {
      if constexpr (DYN == TACD_DYN_EXTERNAL) {
#pragma unroll
        for (int i = 0; i < NX * NX; ++i) A[i] = io.A[((size_t)bb * T + t) * NX * NX + i];
#pragma unroll
        for (int i = 0; i < NX * NU; ++i) Bm[i] = io.Bm[((size_t)bb * T + t) * NX * NU + i];
      } else {
        D::jac(P, sA, sB, x, u, A, Bm);
      }
    };

    // ---- reverse pass: V = H_xx[T-1], v = -grad_X[T-1] (quirks Q1, Q2)
    R V[NX * NX], v[NX];
    R nx_[NX], nu_[NU], ngx[NX], ngu[NU];   // prefetched inputs of the step about to run
    bool ntg[NU];
#pragma unroll
    for (int i = 0; i < NX; ++i) {
#pragma unroll
      for (int j = 0; j < NX; ++j)
        V[i * NX + j] = DIAG ? ((i == j) ? Cp[(size_t)(T - 1) * N + i] : (R)0) : Cp[(size_t)(T - 1) * N * N + i * N + j];
      ngx[i] = gX[(T - 1) * NX + i];
      nx_[i] = X[(T - 1) * NX + i];
      v[i] = -ngx[i];
    }
    if (EXACT) {   // compat="exact": terminal value = C_final[:nx,:nx], -grad_X[T] (fixes quirks Q1, Q2)
      const R* Cfe = io.Cf + (P.Cf_batched ? (size_t)bb * (DIAG ? N : N * N) : 0);
#pragma unroll
      for (int i = 0; i < NX; ++i) {
#pragma unroll
        for (int j = 0; j < NX; ++j) V[i * NX + j] = DIAG ? ((i == j) ? Cfe[i] : (R)0) : Cfe[i * N + j];
        v[i] = -gX[T * NX + i];
      }
    }
#pragma unroll
    for (int i = 0; i < NU; ++i) { nu_[i] = U[(T - 1) * NU + i]; ngu[i] = gU[(T - 1) * NU + i]; ntg[i] = tg[(T - 1) * NU + i] != 0; }
    auto fetch_C = [&](int t) {
      const R* base = io.C + ((size_t)b0 * T + t) * N * N;
      R* buf = mytile + (t & 1) * 32 * L::CW;
      for (int k = 0; k < N * N; ++k) {
        const int idx = k * 32 + lane, r = idx / (N * N), j = idx % (N * N);
        const int rr = (b0 + r < P.B) ? r : (P.B - 1 - b0);
        cp_async<R>(&buf[r * L::CW + j], &base[(size_t)rr * T * N * N + j]);
      }
    };
    if (P.C_batched && !DIAG) { __syncwarp(); fetch_C(T - 1); }
    R nqd[N];   // C_diag: diagonal of the step about to run (prefetched like the other per-step inputs)
#pragma unroll
    for (int i = 0; i < N; ++i) nqd[i] = DIAG ? Cp[(size_t)(T - 1) * N + i] : (R)0;
    for (int t = T - 1; t >= 0; --t) {
      R x[NX], u[NU], Qxx[NX * NX], Qxu[NX * NU], Quu[NU * NU], qx[NX], qu[NU], A[NX * NX], Bm[NX * NU];
      bool fixed[NU];
#pragma unroll
      for (int i = 0; i < NX; ++i) { x[i] = nx_[i]; qx[i] = -ngx[i]; }
#pragma unroll
      for (int i = 0; i < NU; ++i) { u[i] = nu_[i]; qu[i] = -ngu[i]; fixed[i] = ntg[i]; }
      if (t > 0) {   // prefetch step t-1
        ld_row<R, NX>(X + (t - 1) * NX, vecX, nx_);
        ld_row<R, NX>(gX + (t - 1) * NX, vecX, ngx);
#pragma unroll
        for (int i = 0; i < NU; ++i) { nu_[i] = U[(t - 1) * NU + i]; ngu[i] = gU[(t - 1) * NU + i]; ntg[i] = tg[(t - 1) * NU + i] != 0; }
      }
      if (DIAG) {
#pragma unroll
        for (int i = 0; i < NX; ++i) {
#pragma unroll
          for (int j = 0; j < NX; ++j) Qxx[i * NX + j] = (i == j) ? nqd[i] : (R)0;
#pragma unroll
          for (int j = 0; j < NU; ++j) Qxu[i * NU + j] = 0;
        }
#pragma unroll
        for (int i = 0; i < NU; ++i)
#pragma unroll
          for (int j = 0; j < NU; ++j) Quu[i * NU + j] = (i == j) ? nqd[NX + i] : (R)0;
        if (t > 0) {
#pragma unroll
          for (int i = 0; i < N; ++i) nqd[i] = Cp[(size_t)(t - 1) * N + i];
        }
      } else if (P.C_batched) {
        // per-element Hessians: the warp's 32 elements are consecutive, so their C_t blocks are 32 segments
        // of n*n words at a fixed stride — fetched cooperatively (consecutive lanes read consecutive words)
        // with cp.async into one of two buffers, one step ahead of use, so the DRAM latency of streaming C
        // overlaps the Riccati arithmetic of the current step
        cp_async_wait_all();
        __syncwarp();
        const R* row = mytile + ((t & 1) * 32 + lane) * L::CW;
#pragma unroll
        for (int i = 0; i < NX; ++i) {
#pragma unroll
          for (int j = 0; j < NX; ++j) Qxx[i * NX + j] = row[i * N + j];
#pragma unroll
          for (int j = 0; j < NU; ++j) Qxu[i * NU + j] = row[i * N + NX + j];
        }
#pragma unroll
        for (int i = 0; i < NU; ++i)
#pragma unroll
          for (int j = 0; j < NU; ++j) Quu[i * NU + j] = row[(NX + i) * N + NX + j];
        if (t > 0) fetch_C(t - 1);
      } else {
        const R* Ct = Cp + (size_t)t * N * N;
#pragma unroll
        for (int i = 0; i < NX; ++i) {
#pragma unroll
          for (int j = 0; j < NX; ++j) Qxx[i * NX + j] = Ct[i * N + j];
#pragma unroll
          for (int j = 0; j < NU; ++j) Qxu[i * NU + j] = Ct[i * N + NX + j];
        }
#pragma unroll
        for (int i = 0; i < NU; ++i)
#pragma unroll
          for (int j = 0; j < NU; ++j) Quu[i * NU + j] = Ct[(NX + i) * N + NX + j];
      }
      load_AB(t, x, u, A, Bm);
      // (the reference adds reg I to Q_uu here too, controller.py:731; the exact KKT system has none)
      q_blocks<R, NX, NU>(EXACT ? (R)0 : P.reg, A, Bm, V, v, Qxx, Qxu, Quu, qx, qu);
      // box QP on the free coordinates (controller.py:737-763); tight ones pinned to 0
      R kt[NU];
      {
        bool any_free = false;
#pragma unroll
        for (int i = 0; i < NU; ++i) any_free = any_free || !fixed[i];
        R H[NU * NU], q[NU], lo[NU], hi[NU];
#pragma unroll
        for (int i = 0; i < NU; ++i) {
#pragma unroll
          for (int j = 0; j < NU; ++j)
            H[i * NU + j] = (fixed[i] || fixed[j]) ? ((i == j) ? (R)1 : (R)0) : Quu[i * NU + j];
          q[i] = fixed[i] ? (R)0 : qu[i];
          if (fixed[i]) { lo[i] = 0; hi[i] = 0; }
          else if (P.has_umin) { lo[i] = P.umin[i] - u[i]; hi[i] = P.umax[i] - u[i]; }  // quirk Q10
          else { lo[i] = -inf_<R>(); hi[i] = inf_<R>(); }
        }
        if (any_free && EXACT) {
          // exact: the active inputs are frozen, the free ones solve the unconstrained Newton system
          R M[NU * NU];
#pragma unroll
          for (int i = 0; i < NU * NU; ++i) M[i] = H[i] + (((i / NU) == (i % NU) && !fixed[i / NU]) ? (R)1e-8 * P.reg : (R)0);
#pragma unroll
          for (int i = 0; i < NU; ++i) kt[i] = q[i];
          lu_solve<R, NU, 1>(M, kt);
#pragma unroll
          for (int i = 0; i < NU; ++i) kt[i] = fixed[i] ? (R)0 : -kt[i];
        } else if (any_free) {
          pnqp<R, NU>(H, q, lo, hi, kt);
#pragma unroll
          for (int i = 0; i < NU; ++i) kt[i] = fixed[i] ? (R)0 : kt[i];
        } else {
#pragma unroll
          for (int i = 0; i < NU; ++i) kt[i] = 0;
        }
      }
      // K_t = -solve(Q_uu + 1e-8 reg I, Q_ux) — not masked by the active set (quirk Q6)
      R Kt[NU * NX];
      {
        R M[NU * NU];
#pragma unroll
        for (int i = 0; i < NU; ++i) {
#pragma unroll
          for (int j = 0; j < NU; ++j) M[i * NU + j] = Quu[i * NU + j] + ((i == j) ? (R)1e-8 * P.reg : (R)0);
#pragma unroll
          for (int j = 0; j < NX; ++j) Kt[i * NX + j] = Qxu[j * NU + i];
        }
        if (EXACT) {   // exact: rows of frozen inputs are zero, the free block is solved alone (fixes quirk Q6)
#pragma unroll
          for (int i = 0; i < NU; ++i) {
#pragma unroll
            for (int j = 0; j < NU; ++j)
              if (fixed[i] || fixed[j]) M[i * NU + j] = (i == j) ? (R)1 : (R)0;
#pragma unroll
            for (int j = 0; j < NX; ++j) Kt[i * NX + j] = fixed[i] ? (R)0 : Kt[i * NX + j];
          }
        }
        lu_solve<R, NU, NX>(M, Kt);
#pragma unroll
        for (int i = 0; i < NU * NX; ++i) Kt[i] = -Kt[i];
      }
#pragma unroll
      for (int i = 0; i < NU * NX; ++i) SK(t, i) = Kt[i];
#pragma unroll
      for (int i = 0; i < NU; ++i) Sk(t, i) = kt[i];
      value_update<R, NX, NU>(Qxx, Qxu, Quu, qx, qu, Kt, kt, V, v);
    }
    // ---- forward sweep from dX_0 = 0 (quirk Q3) + gradient formulas (controller.py:84-115)
    R dx[NX];
#pragma unroll
    for (int i = 0; i < NX; ++i) dx[i] = 0;
    if (io.grad_x0 && live)
#pragma unroll
      for (int i = 0; i < NX; ++i) io.grad_x0[(size_t)b * NX + i] = EXACT ? -v[i] : (R)0;   // exact: -lambda_0 (fixes quirk Q3)
    R nxr[NX], nur[NU];
#pragma unroll
    for (int i = 0; i < NX; ++i) { nx_[i] = X[i]; nxr[i] = xr[i]; }
#pragma unroll
    for (int i = 0; i < NU; ++i) { nu_[i] = U[i]; nur[i] = ur[i]; }
    // the gains of step t come from the workspace column (L2) one step ahead of use, like the other per-step inputs
    R nK[NU * NX], nk[NU];
#pragma unroll
    for (int i = 0; i < NU * NX; ++i) nK[i] = SK(0, i);
#pragma unroll
    for (int i = 0; i < NU; ++i) nk[i] = Sk(0, i);
    for (int t = 0; t < T; ++t) {
      R x[NX], u[NU], tau[N], dtau[N], du[NU], Kt[NU * NX], kt[NU];
#pragma unroll
      for (int i = 0; i < NU * NX; ++i) Kt[i] = nK[i];
#pragma unroll
      for (int i = 0; i < NU; ++i) kt[i] = nk[i];
#pragma unroll
      for (int i = 0; i < NX; ++i) { x[i] = nx_[i]; tau[i] = x[i] - nxr[i]; dtau[i] = dx[i]; }
#pragma unroll
      for (int i = 0; i < NU; ++i) { u[i] = nu_[i]; tau[NX + i] = u[i] - nur[i]; }
      // prefetch step t+1 (X has T+1 rows; U/u_ref only T)
      ld_row<R, NX>(X + (t + 1) * NX, vecX, nx_);
      ld_row<R, NX>(xr + (t + 1) * NX, vecX, nxr);
      if (t + 1 < T) {
#pragma unroll
        for (int i = 0; i < NU; ++i) { nu_[i] = U[(t + 1) * NU + i]; nur[i] = ur[(t + 1) * NU + i]; }
#pragma unroll
        for (int i = 0; i < NU * NX; ++i) nK[i] = SK(t + 1, i);
#pragma unroll
        for (int i = 0; i < NU; ++i) nk[i] = Sk(t + 1, i);
      }
#pragma unroll
      for (int i = 0; i < NU; ++i) {
        R s = 0;
#pragma unroll
        for (int j = 0; j < NX; ++j) s += Kt[i * NX + j] * dx[j];
        du[i] = s + kt[i];
        dtau[NX + i] = du[i];
      }
      // grad_C = -1/2 (dtau (x) tau + tau (x) dtau), grad_c = -dtau, grad_xref[t] = dx, grad_uref[t] = du.
      // Every lane writes its NV values into its tile row; shared inputs: lanes sum columns into the
      // warp's accumulators; per-element inputs: the warp stores the rows cooperatively (consecutive
      // lanes write consecutive words of the [B,T,...] outputs instead of 32 scattered segments).
      {
        R* row = mytile + lane * L::TW;
        if (packC) {
#pragma unroll
          for (int i = 0; i < N; ++i)
#pragma unroll
            for (int j = i; j < N; ++j) row[L::pk(i, j)] = live ? (R)-0.5 * (dtau[i] * tau[j] + tau[i] * dtau[j]) : (R)0;
        } else {
#pragma unroll
          for (int i = 0; i < N; ++i)
#pragma unroll
            for (int j = 0; j < N; ++j) row[i * N + j] = live ? (R)-0.5 * (dtau[i] * tau[j] + tau[i] * dtau[j]) : (R)0;
        }
#pragma unroll
        for (int i = 0; i < N; ++i) row[vb + i] = live ? -dtau[i] : (R)0;
#pragma unroll
        for (int i = 0; i < NX; ++i) row[vb + N + i] = live ? dx[i] : (R)0;
#pragma unroll
        for (int i = 0; i < NU; ++i) row[vb + N + NX + i] = live ? du[i] : (R)0;
        if (EXACT) {   // exact: d loss / d ref_t = C_t dtau_t (tau = z - ref enters the stationarity as -C ref)
          const R* Ce = Cp + (size_t)t * (DIAG ? N : N * N);
#pragma unroll
          for (int i = 0; i < N; ++i) {
            R s = 0;
            if (DIAG) s = Ce[i] * dtau[i];
            else {
#pragma unroll
              for (int j = 0; j < N; ++j) s += Ce[i * N + j] * dtau[j];
            }
            row[vb + N + i] = live ? s : (R)0;
          }
        }
        __syncwarp();
        if (redC || redc || redx || redu) {   // warp-uniform
#pragma unroll
          for (int q = 0; q < NCOLIT; ++q) {
            const int j = lane + 32 * q;
            if (j < ncol) {
              R sum = 0;
#pragma unroll 8
              for (int r = 0; r < 32; ++r) sum += mytile[r * L::TW + j];
              myacc[t * L::NV + cd0[q]] += sum;
              if (cd1[q] >= 0) myacc[t * L::NV + cd1[q]] += sum;
            }
          }
        }
        const int nlive = (P.B - b0 < 32) ? (P.B - b0) : 32;
        if (io.grad_C && DIAG) {   // compact: the diagonal of -1/2 (dtau (x) tau + tau (x) dtau)
          R* dst = io.grad_C + ((size_t)b0 * T + t) * N;
          for (int idx = lane; idx < nlive * N; idx += 32) {
            const int r = idx / N, j = idx % N;
            dst[(size_t)r * T * N + j] = mytile[r * L::TW + L::pk(j, j)];
          }
        } else if (io.grad_C && P.C_batched) {
          R* dst = io.grad_C + ((size_t)b0 * T + t) * N * N;
          for (int idx = lane; idx < nlive * N * N; idx += 32) {
            const int r = idx / (N * N), j = idx % (N * N);
            dst[(size_t)r * T * N * N + j] = mytile[r * L::TW + j];
          }
        }
        if (io.grad_c && P.C_batched) {
          R* dst = io.grad_c + ((size_t)b0 * T + t) * N;
          for (int idx = lane; idx < nlive * N; idx += 32) {
            const int r = idx / N, j = idx % N;
            dst[(size_t)r * T * N + j] = mytile[r * L::TW + vb + j];
          }
        }
        if (io.grad_xref && P.xref_batched) {
          R* dst = io.grad_xref + (size_t)b0 * szX + t * NX;
          for (int idx = lane; idx < nlive * NX; idx += 32) {
            const int r = idx / NX, j = idx % NX;
            dst[(size_t)r * szX + j] = mytile[r * L::TW + vb + N + j];
          }
        }
        if (io.grad_uref && P.uref_batched) {
          R* dst = io.grad_uref + (size_t)b0 * szU + t * NU;
          for (int idx = lane; idx < nlive * NU; idx += 32) {
            const int r = idx / NU, j = idx % NU;
            dst[(size_t)r * szU + j] = mytile[r * L::TW + vb + N + NX + j];
          }
        }
        if (live) {
          if (io.dU)
#pragma unroll
            for (int i = 0; i < NU; ++i) io.dU[(size_t)b * szU + t * NU + i] = du[i];
          if (io.dX)
#pragma unroll
            for (int i = 0; i < NX; ++i) io.dX[(size_t)b * szX + t * NX + i] = dx[i];
        }
        __syncwarp();
      }
      R A[NX * NX], Bm[NX * NU], dxn[NX];
      load_AB(t, x, u, A, Bm);
#pragma unroll
      for (int i = 0; i < NX; ++i) {
        R s = 0, s2 = 0;
#pragma unroll
        for (int j = 0; j < NX; ++j) s += A[i * NX + j] * dx[j];
#pragma unroll
        for (int j = 0; j < NU; ++j) s2 += Bm[i * NU + j] * du[j];
        dxn[i] = s + s2;
      }
#pragma unroll
      for (int i = 0; i < NX; ++i) dx[i] = dxn[i];
    }
    {
      // terminal: grad_C_final[:nx,:nx], grad_c_final[:nx] (zero padding elsewhere), grad_xref[T]
      R tauN[NX], gxT[NX];   // gxT = grad_xref[T]: dx_T (reference) or C_final[:nx,:nx] dx_T (exact)
#pragma unroll
      for (int i = 0; i < NX; ++i) { tauN[i] = nx_[i] - nxr[i]; gxT[i] = dx[i]; }
      if (EXACT) {
        const R* Cfe = io.Cf + (P.Cf_batched ? (size_t)bb * (DIAG ? N : N * N) : 0);
#pragma unroll
        for (int i = 0; i < NX; ++i) {
          R s = 0;
          if (DIAG) s = Cfe[i] * dx[i];
          else {
#pragma unroll
            for (int j = 0; j < NX; ++j) s += Cfe[i * N + j] * dx[j];
          }
          gxT[i] = s;
        }
      }
      if (live) {
        if (io.grad_Cf && DIAG) {
#pragma unroll
          for (int i = 0; i < N; ++i)
            io.grad_Cf[(size_t)b * N + i] = (i < NX) ? (R)-0.5 * (dx[i < NX ? i : 0] * tauN[i < NX ? i : 0] + tauN[i < NX ? i : 0] * dx[i < NX ? i : 0]) : (R)0;
        } else if (io.grad_Cf && P.Cf_batched)
#pragma unroll
          for (int i = 0; i < N; ++i)
#pragma unroll
            for (int j = 0; j < N; ++j)
              io.grad_Cf[(size_t)b * N * N + i * N + j] =
                  (i < NX && j < NX) ? (R)-0.5 * (dx[i < NX ? i : 0] * tauN[j < NX ? j : 0] + tauN[i < NX ? i : 0] * dx[j < NX ? j : 0]) : (R)0;
        if (io.grad_cf && P.Cf_batched)
#pragma unroll
          for (int i = 0; i < N; ++i) io.grad_cf[(size_t)b * N + i] = (i < NX) ? -dx[i < NX ? i : 0] : (R)0;
        if (io.grad_xref && P.xref_batched)
#pragma unroll
          for (int i = 0; i < NX; ++i) io.grad_xref[(size_t)b * szX + T * NX + i] = gxT[i];
        if (io.dX)
#pragma unroll
          for (int i = 0; i < NX; ++i) io.dX[(size_t)b * szX + T * NX + i] = dx[i];
      }
      if (redCf || redcf || redx) {
        R* row = mytile + lane * L::TW;
#pragma unroll
        for (int i = 0; i < NX; ++i)
#pragma unroll
          for (int j = 0; j < NX; ++j) row[i * NX + j] = live ? (R)-0.5 * (dx[i] * tauN[j] + tauN[i] * dx[j]) : (R)0;
#pragma unroll
        for (int i = 0; i < NX; ++i) { row[NX * NX + i] = live ? -dx[i] : (R)0; row[NX * NX + NX + i] = live ? gxT[i] : (R)0; }
        __syncwarp();
        for (int j = lane; j < L::NVF; j += 32) {
          R sum = 0;
#pragma unroll 8
          for (int r = 0; r < 32; ++r) sum += mytile[r * L::TW + j];
          myacc[T * L::NV + j] += sum;
        }
        __syncwarp();
      }
    }
  }
  if (reduce) {
    __syncthreads();
    for (int i = tid; i < NACC; i += NT) {
      R sum = 0;
      for (int w = 0; w < nw; ++w) sum += acc[(size_t)w * NACC + i];
      partials[(size_t)blockIdx.x * NACC + i] = sum;
    }
  }
#undef SK
#undef Sk
}

// Sum the per-CTA partial vectors in a fixed order and scatter them to the shared gradient outputs.
constexpr int kReduceSeg = 8;
template <typename R, int NX, int NU>
__global__ void reduce_shared_grads(const DevProblem<R> P, const BwdPtrs<R> io, const R* partials, int nparts) {
  constexpr int N = NX + NU;
  using L = BwdLayout<NX, NU>;
  const int T = P.T, NACC = L::nacc(T);
  // kReduceSeg adjacent lanes share one output: each sums every kReduceSeg-th partial (independent loads, all in
  // flight), then the group combines in a fixed butterfly order — deterministic, and ~nparts / kReduceSeg L2 latencies
  // shorter than one thread walking all CTAs (21 us -> 4 us at 148 partials, 2 % of the cart-pole step).
  const int gid = blockIdx.x * blockDim.x + threadIdx.x;
  const int i = gid / kReduceSeg, seg = gid % kReduceSeg;
  R sum = 0;
  if (i < NACC) {
#pragma unroll 4
    for (int c = seg; c < nparts; c += kReduceSeg) sum += partials[(size_t)c * NACC + i];
  }
#pragma unroll
  for (int o = kReduceSeg / 2; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  if (i < NACC && seg == 0) {
    if (i < T * L::NV) {
      const int t = i / L::NV, j = i % L::NV;
      if (j < N * N) { if (io.grad_C && !P.C_batched) io.grad_C[(size_t)t * N * N + j] = sum; }
      else if (j < N * N + N) { if (io.grad_c && !P.C_batched) io.grad_c[(size_t)t * N + (j - N * N)] = sum; }
      else if (j < N * N + N + NX) { if (io.grad_xref && !P.xref_batched) io.grad_xref[t * NX + (j - N * N - N)] = sum; }
      else { if (io.grad_uref && !P.uref_batched) io.grad_uref[t * NU + (j - N * N - N - NX)] = sum; }
    } else {
      const int j = i - T * L::NV;
      if (j < NX * NX) { if (io.grad_Cf && !P.Cf_batched) io.grad_Cf[(j / NX) * N + (j % NX)] = sum; }
      else if (j < NX * NX + NX) { if (io.grad_cf && !P.Cf_batched) io.grad_cf[j - NX * NX] = sum; }
      else { if (io.grad_xref && !P.xref_batched) io.grad_xref[T * NX + (j - NX * NX - NX)] = sum; }
    }
  }
}

}  // namespace tacd
