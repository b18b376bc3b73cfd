// host_common.h — host-side helpers shared by the C-ABI translation units.
#pragma once
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>

#include "tacd_dev.cuh"
#include "ilqr_small.cuh"

namespace tacd {

void set_error(const char* fmt, ...);
int check_cuda(cudaError_t e, const char* what);
void note_launches(int n);

// Experiment knobs (TACD_* environment variables: A/B measurements and the alternative-kernel parity runs).  Each is read
// from the environment ONCE and cached at its call site; tacd_refresh_env() invalidates the caches (the tests flip
// knobs inside one process).  Production leaves them unset: a launch costs no getenv() calls.
struct EnvKnob {
  const char* name;
  int gen;
  bool set;
  int ival;
  char sval[24];
};
int env_generation();
void env_read(EnvKnob& k);
inline const EnvKnob& env_get(EnvKnob& k) { if (k.gen != env_generation()) env_read(k); return k; }
#define TACD_KNOB(var, NAME) static ::tacd::EnvKnob var = {NAME, -1, false, 0, {0}}
inline int env_int(EnvKnob& k, int dflt) { const EnvKnob& e = env_get(k); return e.set ? e.ival : dflt; }
inline bool env_set(EnvKnob& k) { return env_get(k).set; }
bool env_is(EnvKnob& k, const char* value);

struct DeviceInfo { int sms; int smem_optin; };
DeviceInfo device_info();

// A non-blocking helper stream + fork/join events per (host thread, device), created on first use and kept for
// the life of the process: lets two independent pre-passes of one call overlap (fork from the caller's stream,
// join back before the dependent kernel; the pattern is legal inside CUDA-graph capture).
struct SideStream { cudaStream_t stream; cudaEvent_t fork, join; bool ok; };
SideStream side_stream();

template <typename R>
inline DevProblem<R> to_dev(const tacd_problem* p) {
  DevProblem<R> d;
  d.B = p->B; d.T = p->T; d.max_iter = p->max_iter; d.n_alpha = p->n_alpha;
  d.C_batched = p->C_batched; d.Cf_batched = p->Cf_batched;
  d.xref_batched = p->xref_batched; d.uref_batched = p->uref_batched;
  d.ls_cost_shared = p->ls_cost_shared;
  d.C_diag = p->C_diag;
  d.exact = p->compat_exact;
  d.jac_mode = p->jac_mode; d.has_umin = p->has_umin; d.has_umax = p->has_umax;
  d.dt = (R)p->dt; d.reg = (R)p->reg_eps; d.fd_eps = (R)p->fd_eps;
  for (int i = 0; i < TACD_MAX_ALPHA; ++i) d.alphas[i] = (R)p->alphas[i];
  for (int i = 0; i < TACD_MAX_NU; ++i) { d.umin[i] = (R)p->umin[i]; d.umax[i] = (R)p->umax[i]; }
  for (int i = 0; i < 8; ++i) d.dynp[i] = (R)p->dyn_params[i];
  if (p->dyn_id == TACD_DYN_CARTPOLE) {
    // derived constants, formed in the compute type exactly as the reference forms them per call
    const R M = (R)p->dyn_params[0] + (R)p->dyn_params[1];
    d.dynp[4] = (R)1 / M;
    d.dynp[5] = (R)p->dyn_params[1] * (R)p->dyn_params[2];
  }
  d.dyn_A = (const R*)p->dyn_A; d.dyn_B = (const R*)p->dyn_B;
  return d;
}

// small-system (thread-per-element) launchers; return TACD_EUNSUPPORTED when
// (nx, nu, dyn) has no instantiation or the per-thread state does not fit shared memory
// hand-over buffers of the two-phase solve (all in the caller's workspace); nullptr = the thread kernel runs every solve to its end
struct FwdHandover { int cap_iters; unsigned int* count; int32_t* list; void* cost; };
template <typename R> int launch_forward_small(const tacd_problem* p, const tacd_forward_io* io, void* ws, size_t ws_bytes, cudaStream_t s,
                                               const FwdHandover* ho = nullptr);
template <typename R> int launch_backward_small(const tacd_problem* p, const tacd_backward_io* io, void* ws, size_t ws_bytes, cudaStream_t s);
size_t small_backward_ws_bytes(const tacd_problem* p);
// five-lanes-per-element forward kernel (ilqr_group.cuh); TACD_EUNSUPPORTED when not eligible
template <typename R> int launch_forward_group(const tacd_problem* p, const tacd_forward_io* io, void* ws, size_t ws_bytes, cudaStream_t s);
template <typename R> int launch_backward_group(const tacd_problem* p, const tacd_backward_io* io, void* ws, size_t ws_bytes, cudaStream_t s);
size_t group_forward_scratch_bytes(const tacd_problem* p);

// generic (CTA-per-element, run-time dimensions) launchers
template <typename R> int launch_forward_generic(const tacd_problem* p, const tacd_forward_io* io, void* ws, size_t ws_bytes, cudaStream_t s);
template <typename R> int launch_backward_generic(const tacd_problem* p, const tacd_backward_io* io, void* ws, size_t ws_bytes, cudaStream_t s);
size_t generic_forward_ws_bytes(const tacd_problem* p);
// warp-per-element kernels of the LTI sweep (ilqr_lti.cuh); TACD_EUNSUPPORTED when not eligible
template <typename R> int launch_forward_lti(const tacd_problem* p, const tacd_forward_io* io, void* ws, size_t ws_bytes, cudaStream_t s);
template <typename R> int launch_backward_lti(const tacd_problem* p, const tacd_backward_io* io, void* ws, size_t ws_bytes, cudaStream_t s);
size_t lti_ws_bytes(const tacd_problem* p);
size_t generic_backward_ws_bytes(const tacd_problem* p);

}  // namespace tacd
