// ppo_glue.cu — the batched pieces of ACMPC/training_loop.py around the MPC solve:
// environment step with reset-on-done, generalised advantage estimation, value-expansion targets.
// All three are HBM-bound streaming kernels; one thread per environment / sample.
#include "host_common.h"
#include "ilqr_generic.cuh"

namespace tacd {

// no fused multiply-add: the reference evaluates these recursions with separate torch multiplies and adds, and the
// contract of tacd_gae / tacd_discounted_targets is bit-identity with it
template <typename R> __device__ __forceinline__ R mul_rn(R a, R b);
template <> __device__ __forceinline__ float mul_rn<float>(float a, float b) { return __fmul_rn(a, b); }
template <> __device__ __forceinline__ double mul_rn<double>(double a, double b) { return __dmul_rn(a, b); }
template <typename R> __device__ __forceinline__ R add_rn(R a, R b);
template <> __device__ __forceinline__ float add_rn<float>(float a, float b) { return __fadd_rn(a, b); }
template <> __device__ __forceinline__ double add_rn<double>(double a, double b) { return __dadd_rn(a, b); }

struct EnvParams {
  int max_steps, clip_action, reward_on_next;
};
template <typename R>
struct EnvVecs {
  R term_limit[TACD_MAX_NX], w_x[TACD_MAX_NX], w_u[TACD_MAX_NU];
};

// ACMPC/parallel_env.py:34-54 + the environments of examples/02_demo_cartpole_AC.py:78-96 / 01_demo_linear_AC.py:30-44
template <typename R>
__global__ void __launch_bounds__(128) env_step(const DevProblem<R> P, const GenDims<R> d, const EnvParams e, const EnvVecs<R> v,
                                                const R* __restrict__ state, const R* __restrict__ action, const R* __restrict__ reset_state,
                                                int32_t* steps, R* obs, R* next_state, R* reward, uint8_t* terminated, uint8_t* truncated) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= P.B) return;
  const int nx = d.nx(), nu = d.nu();
  R xc[TACD_MAX_NX], xn[TACD_MAX_NX], a[TACD_MAX_NU];
  for (int i = 0; i < nx; ++i) xc[i] = state[(size_t)b * nx + i];
  R act_cost = 0;
  for (int j = 0; j < nu; ++j) {
    const R raw = action[(size_t)b * nu + j];
    act_cost += v.w_u[j] * (raw * raw);  // the reward charges the action the policy asked for (01_demo_linear_AC.py:39)
    R aj = raw;
    if (e.clip_action) {
      if (P.has_umin) aj = tmax<R>(aj, P.umin[j]);
      if (P.has_umax) aj = tmin<R>(aj, P.umax[j]);
    }
    a[j] = aj;
  }
  if (d.dyn == TACD_DYN_LTI) {
    for (int i = 0; i < nx; ++i) {
      R s = 0, s2 = 0;
      for (int j = 0; j < nx; ++j) s += P.dyn_A[i * nx + j] * xc[j];
      for (int j = 0; j < nu; ++j) s2 += P.dyn_B[i * nu + j] * a[j];
      xn[i] = s + s2;
    }
  } else {
    gen_dyn_step_scalar<R>(P, d.dyn, xc, a, xn);
  }
  bool term = false;
  R st_cost = 0;
  for (int i = 0; i < nx; ++i) {
    term = term || (abs_<R>(xn[i]) > v.term_limit[i]);
    const R s = e.reward_on_next ? xn[i] : xc[i];
    st_cost += v.w_x[i] * (s * s);
  }
  int st = steps ? steps[b] + 1 : 0;
  const bool trunc = steps != nullptr && e.max_steps > 0 && st >= e.max_steps;
  const bool reset = (term || trunc) && reset_state != nullptr;
  if (reset) st = 0;
  if (steps) steps[b] = st;
  reward[b] = -(st_cost + act_cost);
  terminated[b] = term ? 1 : 0;
  truncated[b] = trunc ? 1 : 0;
  for (int i = 0; i < nx; ++i) {
    if (obs) obs[(size_t)b * nx + i] = xn[i];
    next_state[(size_t)b * nx + i] = reset ? reset_state[(size_t)b * nx + i] : xn[i];
  }
}

// ACMPC/training_loop.py:30-41: one thread per environment, reverse scan over its row.  A warp reads 32 rows at
// stride T; consecutive t of one row share sectors, which L1 keeps, so DRAM traffic stays at the algorithmic 16 B
// per (environment, step).
template <typename R>
__global__ void __launch_bounds__(128) gae_scan(int N, int T, const R* __restrict__ rewards, const R* __restrict__ values, R gamma, R gl,
                                                R* __restrict__ returns, R* __restrict__ adv) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  const R* r = rewards + (size_t)n * T;
  const R* v = values + (size_t)n * T;
  R vnext = v[T - 1], last = 0;
  // the loads do not depend on the recurrence: unrolled, eight steps of them are in flight ahead of the adds
#pragma unroll 8
  for (int t = T - 1; t >= 0; --t) {
    const R vt = v[t];
    const R delta = add_rn<R>(add_rn<R>(r[t], mul_rn<R>(gamma, vnext)), -vt);
    last = add_rn<R>(delta, mul_rn<R>(gl, last));
    adv[(size_t)n * T + t] = last;
    returns[(size_t)n * T + t] = add_rn<R>(last, vt);
    vnext = vt;
  }
}

// ACMPC/training_loop.py:162-170
template <typename R>
__global__ void __launch_bounds__(128) discounted_scan(int N, int H, const R* __restrict__ rewards, const R* __restrict__ bootstrap, R gamma,
                                                       R* __restrict__ targets) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  R next = bootstrap ? bootstrap[n] : (R)0;
#pragma unroll 8
  for (int t = H - 1; t >= 0; --t) {
    next = add_rn<R>(rewards[(size_t)n * H + t], mul_rn<R>(gamma, next));
    targets[(size_t)n * H + t] = next;
  }
}

template <typename R>
static int run_env_step(const tacd_problem* p, const tacd_env_io* io, cudaStream_t s) {
  EnvParams e{io->max_steps, io->clip_action, io->reward_on_next};
  EnvVecs<R> v;
  for (int i = 0; i < TACD_MAX_NX; ++i) { v.term_limit[i] = (R)io->term_limit[i]; v.w_x[i] = (R)io->w_x[i]; }
  for (int j = 0; j < TACD_MAX_NU; ++j) v.w_u[j] = (R)io->w_u[j];
  GenDims<R> d;
  d.nx_ = p->nx; d.nu_ = p->nu; d.T = p->T; d.dyn = p->dyn_id;
  env_step<R><<<(p->B + 127) / 128, 128, 0, s>>>(to_dev<R>(p), d, e, v, (const R*)io->state, (const R*)io->action, (const R*)io->reset_state, io->steps,
                                                 (R*)io->obs, (R*)io->next_state, (R*)io->reward, io->terminated, io->truncated);
  note_launches(1);
  return check_cuda(cudaPeekAtLastError(), "env_step launch");
}

}  // namespace tacd

using namespace tacd;

extern "C" {

int tacd_env_step(const tacd_problem* p, const tacd_env_io* io, void* stream) {
  if (!p || !io) { set_error("tacd_env_step: null problem / io"); return TACD_EINVAL; }
  if (p->nx < 1 || p->nu < 1 || p->nx > TACD_MAX_NX || p->nu > TACD_MAX_NU || p->B < 0 || (p->dtype != TACD_F32 && p->dtype != TACD_F64)) {
    set_error("tacd_env_step: bad sizes / dtype"); return TACD_EINVAL;
  }
  if (p->dyn_id == TACD_DYN_EXTERNAL) { set_error("tacd_env_step needs embedded dynamics"); return TACD_EINVAL; }
  if (p->dyn_id == TACD_DYN_LTI && (!p->dyn_A || !p->dyn_B)) { set_error("tacd_env_step: LTI dynamics without A / B"); return TACD_EINVAL; }
  if (p->B == 0) return TACD_OK;
  if (!io->state || !io->action || !io->next_state || !io->reward || !io->terminated || !io->truncated) {
    set_error("tacd_env_step: required pointer is null"); return TACD_EINVAL;
  }
  return p->dtype == TACD_F32 ? run_env_step<float>(p, io, (cudaStream_t)stream) : run_env_step<double>(p, io, (cudaStream_t)stream);
}

int tacd_gae(int32_t dtype, int32_t N, int32_t T, const void* rewards, const void* values, double gamma, double lam, void* returns,
             void* advantages, void* stream) {
  if (N < 0 || T < 1 || (dtype != TACD_F32 && dtype != TACD_F64)) { set_error("tacd_gae: bad sizes / dtype"); return TACD_EINVAL; }
  if (N == 0) return TACD_OK;
  if (!rewards || !values || !returns || !advantages) { set_error("tacd_gae: required pointer is null"); return TACD_EINVAL; }
  cudaStream_t s = (cudaStream_t)stream;
  // one sequential scan per thread: PPO-sized batches (a few thousand rows) are spread one warp per CTA over the SMs
  const int block = N >= 148 * 128 ? 128 : 32;
  const int grid = (N + block - 1) / block;
  // gamma * lam is formed in double and rounded once, as the reference's Python float product is (training_loop.py:39)
  if (dtype == TACD_F32)
    gae_scan<float><<<grid, block, 0, s>>>(N, T, (const float*)rewards, (const float*)values, (float)gamma, (float)(gamma * lam), (float*)returns,
                                         (float*)advantages);
  else
    gae_scan<double><<<grid, block, 0, s>>>(N, T, (const double*)rewards, (const double*)values, gamma, gamma * lam, (double*)returns,
                                          (double*)advantages);
  note_launches(1);
  return check_cuda(cudaPeekAtLastError(), "gae_scan launch");
}

int tacd_discounted_targets(int32_t dtype, int32_t N, int32_t H, const void* rewards, const void* bootstrap, double gamma, void* targets,
                            void* stream) {
  if (N < 0 || H < 1 || (dtype != TACD_F32 && dtype != TACD_F64)) { set_error("tacd_discounted_targets: bad sizes / dtype"); return TACD_EINVAL; }
  if (N == 0) return TACD_OK;
  if (!rewards || !targets) { set_error("tacd_discounted_targets: required pointer is null"); return TACD_EINVAL; }
  cudaStream_t s = (cudaStream_t)stream;
  const int block = N >= 148 * 128 ? 128 : 32;
  const int grid = (N + block - 1) / block;
  if (dtype == TACD_F32)
    discounted_scan<float><<<grid, block, 0, s>>>(N, H, (const float*)rewards, (const float*)bootstrap, (float)gamma, (float*)targets);
  else
    discounted_scan<double><<<grid, block, 0, s>>>(N, H, (const double*)rewards, (const double*)bootstrap, gamma, (double*)targets);
  note_launches(1);
  return check_cuda(cudaPeekAtLastError(), "discounted_scan launch");
}

}  // extern "C"
