/*
 * tacd_ilqr.h — C ABI of the B200-native batched differentiable iLQR solve.
 *
 * This is the drop-in boundary for ONE hot path of TACDMPC/TACDMPC: the solve
 * behind DifferentialMPC.ILQRSolve / DifferentiableMPCController.  Every entry
 * point takes plain device pointers and sizes (no torch types); the Python host
 * mirror in tacdmpc_b200/DifferentialMPC binds them with ctypes.
 *
 * Reference interface replaced (paths relative to the reference checkout):
 *   tacd_ilqr_forward   <- ILQRSolve.forward / solve_step   DifferentialMPC/controller.py:27-60, 275-315
 *   tacd_ilqr_backward  <- ILQRSolve.backward + _zero_constrained_lqr  controller.py:63-115, 707-788
 *   tacd_linearize      <- linearize_dynamics                controller.py:448-471
 *   tacd_riccati        <- backward_lqr (live fallback loop) controller.py:546-578
 *   tacd_linesearch     <- evaluate_alphas                   controller.py:589-620
 *   tacd_rollout        <- rollout_trajectory                controller.py:361-398
 *   tacd_objective      <- GeneralQuadCost.objective         DifferentialMPC/cost.py:37-62
 *   tacd_pnqp           <- pnqp                              DifferentialMPC/utils.py:66-124
 *
 * Conventions
 *   - All tensors are dense, row-major, in the scalar type named by
 *     tacd_problem.dtype (TACD_F32 / TACD_F64); n = nx + nu.
 *   - Every function returns 0 on success or a negative TACD_E* code; nothing
 *     throws or exits.  tacd_last_error() returns a thread-local message.
 *   - Kernels are enqueued on `stream` (a cudaStream_t passed as void*); no
 *     entry point synchronises the host.  No entry point allocates or frees
 *     user-visible memory; scratch comes from the caller (`workspace`).
 *   - The oracle (oracle/ilqr_oracle.c, test infrastructure only) exports the
 *     same signatures with the prefix tacd_oracle_ and HOST pointers.
 */
#ifndef TACD_ILQR_H
#define TACD_ILQR_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TACD_ABI_VERSION 4

#define TACD_MAX_ALPHA 8
#define TACD_MAX_NU 16
#define TACD_MAX_NX 64

/* error codes */
#define TACD_OK 0
#define TACD_EINVAL (-1)      /* bad argument / unsupported shape */
#define TACD_EUNSUPPORTED (-2)/* valid request this build has no kernel for */
#define TACD_ECUDA (-3)       /* CUDA launch / runtime error (see tacd_last_error) */
#define TACD_EWORKSPACE (-4)  /* workspace too small */

enum tacd_dtype { TACD_F32 = 0, TACD_F64 = 1 };

/* Shipped dynamics.  f(x,u,dt) -> x+ and its Jacobians are
 * embedded in device code; EXTERNAL means "A_t,B_t are supplied as tensors"
 * (stand-alone Riccati / backward only). */
enum tacd_dyn {
  TACD_DYN_LTI = 0,               /* x+ = A x + B u, A[nx,nx], B[nx,nu] device ptrs (double integrator: examples/01_demo_linear.py:12-26) */
  TACD_DYN_CARTPOLE = 1,          /* explicit-Euler gym cart-pole, params {m_c,m_p,l,g} (examples/02_demo_cartpole_regulation.py:29-80) */
  TACD_DYN_UNICYCLE = 2,          /* explicit-Euler unicycle (examples/03_demo_unicycle_tracking.py:8-49) */
  TACD_DYN_UNICYCLE_WRAPPED = 3,  /* same with theta wrapped to [-pi,pi) (examples/03_demo_unicycle_AC.py:107-125) */
  TACD_DYN_EXTERNAL = 4
};

/* GradMethod (controller.py:119-129) */
enum tacd_jac { TACD_JAC_ANALYTIC = 0, TACD_JAC_AUTO_DIFF = 1, TACD_JAC_FINITE_DIFF = 2 };

/* flags[b] bits written by the solve */
#define TACD_FLAG_NONPD 1u    /* some Q_uu + reg*I was not positive definite; LU fallback used (quirk Q8) */
#define TACD_FLAG_MAXITER 2u  /* stopped by max_iter while still improving */

typedef struct tacd_problem {
  int32_t B, T, nx, nu;
  int32_t dtype;            /* enum tacd_dtype */
  /* broadcast flags (cost.py:29-35): 1 = per-element leading batch dim, 0 = shared */
  int32_t C_batched;        /* C[B,T,n,n], c[B,T,n]   vs  C[T,n,n], c[T,n] */
  int32_t Cf_batched;       /* C_final[B,n,n], c_final[B,n]  vs  [n,n],[n] */
  int32_t xref_batched;     /* x_ref[B,T+1,nx] vs [1,T+1,nx] */
  int32_t uref_batched;     /* u_ref[B,T,nu]   vs [1,T,nu] */
  /* line-search cost (quirk Q4): 1 = candidate costs use the separate, shared
   * (C_ls,c_ls,Cf_ls,cf_ls) of shapes [T,n,n],[T,n],[n,n],[n]; 0 = same cost as acceptance */
  int32_t ls_cost_shared;
  int32_t dyn_id;           /* enum tacd_dyn */
  int32_t jac_mode;         /* enum tacd_jac */
  double dyn_params[8];
  const void* dyn_A;        /* LTI only: [nx,nx], dtype as above */
  const void* dyn_B;        /* LTI only: [nx,nu] */
  double dt;
  double reg_eps;
  double fd_eps;            /* finite-difference step actually used (reference: always 1e-4, quirk Q9) */
  int32_t max_iter;
  int32_t n_alpha;
  double alphas[TACD_MAX_ALPHA];
  int32_t has_umin, has_umax;
  double umin[TACD_MAX_NU], umax[TACD_MAX_NU];
  /* Diagonal per-element cost in compact form (what ACMPC/actor.py:58-81 builds with diag_embed):
   * 1 = C[B,T,n] and C_final[B,n] hold the DIAGONALS (c, c_final unchanged); grad_C / grad_Cf come back
   * in the same compact shape; the shared line-search cost (C_ls[T,n], Cf_ls[n]) is compact too.
   * Requires C_batched = Cf_batched = 1.  Supported by tacd_ilqr_forward (five-lanes-per-element kernel)
   * and tacd_ilqr_backward (thread-per-element kernel); otherwise TACD_EUNSUPPORTED — expand to dense. */
  int32_t C_diag;
  /* 0 = reproduce the reference's backward with its quirks Q1-Q3, Q6 (default).  1 = exact implicit
   * differentiation (compat="exact"; Amos et al. 2018): terminal value C_final / -grad_X[T],
   * gains and offsets solved on the free inputs only (no box), grad_x0 = -v_0, grad_xref/uref = C_t dtau_t.
   * tacd_ilqr_backward (thread-per-element and run-time-dimension kernels); needs tacd_backward_io.Cf. */
  int32_t compat_exact;
} tacd_problem;

/* ---- whole solve ------------------------------------------------------- */
typedef struct tacd_forward_io {
  /* inputs */
  const void* x0;     /* [B,nx] */
  const void* C;      /* see C_batched */
  const void* c;
  const void* Cf;
  const void* cf;
  const void* C_ls;   /* only read when ls_cost_shared */
  const void* c_ls;
  const void* Cf_ls;
  const void* cf_ls;
  const void* xref;
  const void* uref;
  const void* Uinit;  /* [B,T,nu] */
  /* lock-step replay of a recorded decision trace (tests only): when
   * replay_iters != NULL the argmin / improved decisions are taken from the
   * trace instead of being computed. */
  const int32_t* replay_iters;   /* [B] executed iterations */
  const uint8_t* replay_alpha;   /* [B,max_iter] */
  const uint8_t* replay_flags;   /* [B] (TACD_FLAG_MAXITER decides whether the last iteration improved) */
  /* outputs */
  void* X;            /* [B,T+1,nx] */
  void* U;            /* [B,T,nu] */
  void* A;            /* [B,T,nx,nx] Jacobians at the solution, nullable */
  void* Bm;           /* [B,T,nx,nu] nullable */
  uint8_t* tight;     /* [B,T,nu] active-set mask (controller.py:636-642) */
  int32_t* iters;     /* [B] executed iLQR iterations (= improving + 1, capped by max_iter) */
  uint8_t* alpha_trace; /* [B,max_iter] argmin alpha index per executed iteration, 0xFF after; nullable */
  uint8_t* flags;     /* [B] TACD_FLAG_* */
  void* cost;         /* [B] final objective; nullable */
  void* cost_trace;   /* [B,max_iter,n_alpha+2] candidate costs, new cost, best cost before; nullable (debug) */
} tacd_forward_io;

typedef struct tacd_backward_io {
  /* inputs (saved by forward) */
  const void* X;       /* [B,T+1,nx] */
  const void* U;       /* [B,T,nu] */
  const void* C;       /* Hessian blocks are views of C (cost.py:95-97) */
  const void* Cf;      /* unused by the reference backward (quirk Q1); kept for the ABI */
  const void* xref;
  const void* uref;
  const uint8_t* tight;/* [B,T,nu] */
  const void* A;       /* [B,T,nx,nx], required iff dyn_id == EXTERNAL */
  const void* Bm;      /* [B,T,nx,nu] */
  const void* gX;      /* upstream grad wrt X* [B,T+1,nx] */
  const void* gU;      /* upstream grad wrt U* [B,T,nu] */
  /* outputs, each nullable = not needed.  When C_batched==0 grad_C/grad_c are
   * the SUM over the batch ([T,n,n]/[T,n]); same for Cf_batched / xref / uref. */
  void* grad_C;
  void* grad_c;
  void* grad_Cf;
  void* grad_cf;
  void* grad_x0;       /* [B,nx]; identically zero in the reference (quirk Q3) */
  void* grad_xref;
  void* grad_uref;
  void* dX;            /* [B,T+1,nx] KKT sensitivities, nullable (debug / tests) */
  void* dU;            /* [B,T,nu] */
} tacd_backward_io;

int tacd_abi_version(void);
const char* tacd_last_error(void);
/* diagnostic: kernels of this library enqueued by the calling process so far (all entry points) */
unsigned long long tacd_kernel_launches(void);
/* The library caches its TACD_* experiment knobs (environment variables) on first use; call this after changing one. */
void tacd_refresh_env(void);

/* bytes of device scratch tacd_ilqr_forward / tacd_ilqr_backward need */
size_t tacd_forward_workspace_bytes(const tacd_problem* p);
size_t tacd_backward_workspace_bytes(const tacd_problem* p);

int tacd_ilqr_forward(const tacd_problem* p, const tacd_forward_io* io,
                      void* workspace, size_t workspace_bytes, void* stream);
int tacd_ilqr_backward(const tacd_problem* p, const tacd_backward_io* io,
                       void* workspace, size_t workspace_bytes, void* stream);

/* ---- stand-alone stages (unit parity + generic-callable path) ---------- */
/* X[B,T+1,nx] from x0[B,nx], U[B,T,nu]; no clamping (controller.py:361-398) */
int tacd_rollout(const tacd_problem* p, const void* x0, const void* U, void* X, void* stream);
/* A[B,T,nx,nx], Bm[B,T,nx,nu] at (X[:, :-1], U) (controller.py:448-471) */
int tacd_linearize(const tacd_problem* p, const void* X, const void* U, void* A, void* Bm, void* stream);
/* cost[B] (cost.py:37-62) */
int tacd_objective(const tacd_problem* p, const void* X, const void* U,
                   const void* C, const void* c, const void* Cf, const void* cf,
                   const void* xref, const void* uref, void* cost, void* stream);
/* backward_lqr inputs in the reference's layout: A[B,T,nx,nx] B[B,T,nx,nu]
 * lx[B,T,nx] lu[B,T,nu] lxx[B,T,nx,nx] lxu[B,T,nx,nu] luu[B,T,nu,nu] lxN[B,nx] lxxN[B,nx,nx]
 * -> K[B,T,nu,nx], k[B,T,nu], flags[B] (nullable) */
int tacd_riccati(const tacd_problem* p, const void* A, const void* Bm,
                 const void* lx, const void* lu, const void* lxx, const void* lxu, const void* luu,
                 const void* lxN, const void* lxxN, void* K, void* k, uint8_t* flags, void* stream);
/* evaluate_alphas: Xc[B,n_alpha,T+1,nx], Uc[B,n_alpha,T,nu], costs[B,n_alpha];
 * cost tensors follow C_batched/Cf_batched (NOT the ls_* split). */
int tacd_linesearch(const tacd_problem* p, const void* x0, const void* X, const void* U,
                    const void* K, const void* k,
                    const void* C, const void* c, const void* Cf, const void* cf,
                    const void* xref, const void* uref,
                    void* Xc, void* Uc, void* costs, void* stream);
/* One receding-horizon step of the closed loop the reference's demos and (stale) tests run in Python
 * (examples/03_demo_unicycle_tracking.py:131-141, tests/test_batch_tracking.py:76-84), after the solve of
 * simulation step k:  u_k = U_opt[:,0];  x <- f(x, u_k);  Xs[:,k+1] = x;  Us[:,k] = u_k;
 * U_warm = roll(U_opt, -1) with the last input zeroed;  and, if k+1 < n_sim, the reference windows of the next
 * solve  xref_win = xref_full[:, k+1 : k+T+2],  uref_win = uref_full[:, k+1 : k+T+1]  (contiguous copies).
 * x[B,nx] in/out; Xs[B,n_sim+1,nx]; Us[B,n_sim,nu]; xref_full[B or 1, ref_len, nx]; uref_full[B or 1,
 * ref_len-1, nu] (batched as tacd_problem.xref_batched / uref_batched say); U_warm, xref_win, uref_win nullable. */
int tacd_closed_loop_advance(const tacd_problem* p, int32_t k, int32_t n_sim, void* x, const void* U_opt,
                             void* Xs, void* Us, void* U_warm,
                             const void* xref_full, const void* uref_full, int32_t ref_len,
                             void* xref_win, void* uref_win, void* stream);
/* Augmented-Lagrangian update for state box constraints x_min <= x_t <= x_max (t = 0..T) — the constraint
 * of DifferentialMPC/AugmentedLagrangianManager.py:33-62 (g = [x - x_max; x_min - x], utils.py:31-35).  That file
 * is un-importable at reference HEAD and its manager class is missing, so the multiplier / penalty schedule is
 * DEFINED here: Powell-Hestenes-Rockafellar, per (element, step, state, side):
 *     lam <- max(0, lam + rho * g(x)),  active = lam > 0      (update_lambda != 0)
 *     active = lam + rho_next * g(x) > 0                      (update_lambda == 0: lam unchanged)
 * and the quadratic model of the penalty of the active constraints around x_ref, which the next inner iLQR
 * solve adds to its cost:  dq[b,t,i] (added to C_t[i,i], or to C_final[i,i] for t = T) = rho_next * (#active
 * sides),  dp[b,t,i] (added to c_t[i] / c_final[i]) = act_hi (lam_hi + rho_next (x_ref - x_max)) -
 * act_lo (lam_lo + rho_next (x_min - x_ref)).   viol[b] = max_t,i max(0, g).   X, lam_*, dq, dp: [B,T+1,nx];
 * x_ref [B or 1, T+1, nx]; x_min, x_max [nx] (+-inf = unbounded side). */
int tacd_al_update(int32_t dtype, int32_t B, int32_t T, int32_t nx, const void* X, const void* xref, int32_t xref_batched,
                   const void* x_min, const void* x_max, void* lam_lo, void* lam_hi, double rho, double rho_next,
                   int32_t update_lambda, void* dq, void* dp, void* viol, void* stream);
/* ---- PPO glue: what ACMPC/training_loop.py does around the MPC solve with Python loops over
 * environments and time steps, as three batched kernels.  dtype as tacd_dtype; all arrays on the device. ---- */

/* One step of B independent environments whose transition is the problem's dynamics plugin (p->dyn_id, dyn_params,
 * dt; p->T is ignored) — ACMPC/parallel_env.py:34-54 (the per-environment Python loop), the environments of
 * examples/02_demo_cartpole_AC.py:78-96 and examples/01_demo_linear_AC.py:30-44, and the reset-on-done of
 * ACMPC/training_loop.py:103-107:
 *     a  = clip_action ? clamp(action, p->umin, p->umax) : action
 *     x' = f(x, a);   steps += 1
 *     terminated = exists i: |x'_i| > term_limit[i]          (term_limit[i] = +inf: state i is not checked)
 *     truncated  = steps >= max_steps                        (max_steps <= 0: never)
 *     reward = -(sum_i w_x[i] s_i^2 + sum_j w_u[j] action_j^2),  s = x' if reward_on_next else x
 *     obs = x'                                               (what env.step returns)
 *     next_state = (terminated or truncated) and reset_state ? reset_state : x';  steps = 0 where that happened
 * The reset states are drawn by the caller (the reference draws them on the host, per environment), which keeps the
 * random stream the caller's. */
typedef struct tacd_env_io {
  const void* state;        /* [B,nx] */
  const void* action;       /* [B,nu] */
  const void* reset_state;  /* [B,nx] or NULL */
  int32_t* steps;           /* [B] in/out, or NULL (then truncated is always 0) */
  void* obs;                /* [B,nx] out, nullable */
  void* next_state;         /* [B,nx] out; may alias state */
  void* reward;             /* [B] out */
  uint8_t* terminated;      /* [B] out */
  uint8_t* truncated;       /* [B] out */
  int32_t max_steps;
  int32_t clip_action;
  int32_t reward_on_next;
  int32_t reserved_;
  double term_limit[TACD_MAX_NX];
  double w_x[TACD_MAX_NX];
  double w_u[TACD_MAX_NU];
} tacd_env_io;
int tacd_env_step(const tacd_problem* p, const tacd_env_io* io, void* stream);

/* Generalised advantage estimation over N environments (ACMPC/training_loop.py:30-41 under vmap, :119):
 *     v_T = values[:, T-1]                                   (the reference bootstraps with the last value)
 *     delta_t = rewards_t + gamma v_{t+1} - v_t;   adv_t = delta_t + (gamma lam) adv_{t+1};   returns = adv + values
 * rewards, values, returns, advantages: [N,T].  Evaluated in the reference's operation order without fused
 * multiply-adds, so fp32 / fp64 results are bit-identical to the reference on the CPU. */
int tacd_gae(int32_t dtype, int32_t N, int32_t T, const void* rewards, const void* values, double gamma, double lam,
             void* returns, void* advantages, void* stream);
/* Value-expansion targets (ACMPC/training_loop.py:162-170):  target_t = rewards_t + gamma target_{t+1},
 * target_H = bootstrap.   rewards, targets: [N,H];  bootstrap: [N] (NULL = 0).  Same rounding contract as tacd_gae. */
int tacd_discounted_targets(int32_t dtype, int32_t N, int32_t H, const void* rewards, const void* bootstrap, double gamma,
                            void* targets, void* stream);

/* FP32-pipe microbenchmark for the roofline denominator of the forward solve (SURVEY 8d asks for an FMA probe beside the nominal
 * 148 SM x 128 lanes x 2 x f_SM): launches `blocks` CTAs of 256 threads, each thread running 16 independent dependent-FMA chains of
 * `iters` steps.  kind 0 = scalar FFMA with three register sources (what the solve kernels issue), kind 1 = packed fma.rn.f32x2
 * (FFMA2).  seed: 3 floats on the device {multiplier, addend, start}; out: blocks * 256 floats; *flops (host, nullable) receives the
 * floating-point operations of the launch.  The caller times the launch with CUDA events.  No reference counterpart. */
int tacd_fp32_probe(int32_t kind, int32_t iters, int32_t blocks, const void* seed, void* out, double* flops, void* stream);

/* Tensor-core decision probe for n_x >= 16 (BASELINE north_star: "tensor cores only if ncu shows a batched-GEMM formulation beats
 * the register-resident path"): evaluates the Riccati step's dominant product  Q_b = C_b + F' V_b F  (controller.py:551-558;
 * F = [A | B] shared by the batch, as for the LTI plugin of config 5) `reps` times per element, fp32.
 * kind 0 = CTA per element, 1 x 4 register tiles from shared memory (the round-1 generic kernel's scheme);
 * kind 1 = warp per element, lane-per-row operands in registers; kind 2 = warp per element, mma.sync TF32 with the 3-term split.
 * V [B,nx,nx], F [nx,nx+nu], C [B,n,n] in; Q [B,n,n] out (device pointers).  (nx, nu) in {(32, 8), (16, 8)}. */
int tacd_probe_fvf(int32_t kind, int32_t nx, int32_t nu, int32_t B, int32_t reps, const void* V, const void* F, const void* C,
                   void* Q, void* stream);

/* The gain solve of the warp-per-element LTI kernels (csrc/ilqr_lti.cuh::gains) in isolation, for unit parity against LAPACK:
 * x = -(Q + reg' I)^-1 r for n problems, fp32, nu in {4, 8}.  kkt = 0: the forward pass's solve (controller.py:557-566: Cholesky of
 * Q + reg I, LU with partial pivoting when that fails -> nonpd[i] = 1); kkt = 1: the backward pass's (LU of Q + 1e-8 reg I, :766).
 * Quu [n,nu,nu], rhs [n,nu,2] (two right-hand sides) in; K [n,nu], k [n,nu] = the two solutions, nonpd [n] out. */
int tacd_probe_lti_gains(int32_t dtype, int32_t nu, int32_t kkt, int32_t n, double reg, const void* Quu, const void* rhs, void* K,
                         void* k, uint8_t* nonpd, void* stream);

/* batched projected-Newton box QP: H[N,m,m] q[N,m] lo[N,m] hi[N,m] -> x[N,m], iters[N] (nullable) */
int tacd_pnqp(int32_t dtype, int32_t N, int32_t m, const void* H, const void* q,
              const void* lo, const void* hi, void* x, int32_t* iters, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* TACD_ILQR_H */
