// ilqr_lti_launch.cuh — host side of the warp-per-element LTI kernels (ilqr_lti.cuh): eligibility, workspace, launch.
#pragma once
#include "host_common.h"
#include "ilqr_lti.cuh"

namespace tacd {

#define TACD_LTI_SIZES(X) X(8, 4) X(16, 4) X(16, 8) X(32, 8)

static bool lti_enabled() {
  TACD_KNOB(k, "TACD_LTI_KERNEL");     // TACD_LTI_KERNEL=0: run the LTI sweep on the run-time-dimension kernels (A/B, parity of both)
  return !(env_set(k) && env_int(k, 1) == 0);
}
static bool lti_size(const tacd_problem* p) {
#define X(NX_, NU_) if (p->nx == NX_ && p->nu == NU_) return true;
  TACD_LTI_SIZES(X)
#undef X
  return false;
}

template <typename K>
static int lti_prep_smem(K kernel, size_t bytes) {
  const DeviceInfo di = device_info();
  if (bytes > (size_t)di.smem_optin) { set_error("LTI kernel needs %zu B shared memory (> %d)", bytes, di.smem_optin); return TACD_EUNSUPPORTED; }
  return check_cuda(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes), "smem attr");
}

template <typename R, int NX, int NU>
static int run_forward_lti(const tacd_problem* p, const tacd_forward_io* io, void* ws, size_t ws_bytes, cudaStream_t s) {
  using C = LtiCfg<NX, NU>;
  constexpr int warps = lti_warps<NX, NU, R>();
  const DeviceInfo di = device_info();
  int grid = (p->B + warps - 1) / warps;
  if (grid > di.sms) grid = di.sms;
  const size_t need = 256 + (size_t)grid * warps * lti_slice_words(p->T, NX, NU) * sizeof(R);
  if (ws_bytes < need) { set_error("forward workspace too small for the LTI kernel: %zu < %zu", ws_bytes, need); return TACD_EWORKSPACE; }
  const size_t smem = ((size_t)C::F_WORDS + (size_t)warps * C::WARP_WORDS) * sizeof(R);
  auto kernel = ilqr_forward_lti<R, NX, NU>;
  if (int rc = lti_prep_smem(kernel, smem)) return rc;
  if (int rc = check_cuda(cudaMemsetAsync(ws, 0, 16, s), "memset queue counter")) return rc;
  FwdPtrs<R> f;
  f.x0 = (const R*)io->x0; f.C = (const R*)io->C; f.c = (const R*)io->c; f.Cf = (const R*)io->Cf; f.cf = (const R*)io->cf;
  f.C_ls = (const R*)io->C_ls; f.c_ls = (const R*)io->c_ls; f.Cf_ls = (const R*)io->Cf_ls; f.cf_ls = (const R*)io->cf_ls;
  f.xref = (const R*)io->xref; f.uref = (const R*)io->uref; f.Uinit = (const R*)io->Uinit;
  f.replay_iters = io->replay_iters; f.replay_alpha = io->replay_alpha; f.replay_flags = io->replay_flags;
  f.X = (R*)io->X; f.U = (R*)io->U; f.A = (R*)io->A; f.Bm = (R*)io->Bm;
  f.tight = io->tight; f.iters = io->iters; f.alpha_trace = io->alpha_trace; f.flags = io->flags;
  f.cost = (R*)io->cost; f.cost_trace = (R*)io->cost_trace; f.counter = nullptr;
  kernel<<<grid, warps * 32, smem, s>>>(to_dev<R>(p), f, (R*)((char*)ws + 256), (unsigned int*)ws);
  note_launches(1);
  return check_cuda(cudaPeekAtLastError(), "ilqr_forward_lti launch");
}

template <typename R>
int launch_forward_lti(const tacd_problem* p, const tacd_forward_io* io, void* ws, size_t ws_bytes, cudaStream_t s) {
  if (!lti_enabled() || p->dyn_id != TACD_DYN_LTI || !lti_size(p) || p->C_diag || p->n_alpha > kLtiNA || p->n_alpha < 1) return TACD_EUNSUPPORTED;
#define X(NX_, NU_) if (p->nx == NX_ && p->nu == NU_) return run_forward_lti<R, NX_, NU_>(p, io, ws, ws_bytes, s);
  TACD_LTI_SIZES(X)
#undef X
  return TACD_EUNSUPPORTED;
}

template <typename R, int NX, int NU>
static int run_backward_lti(const tacd_problem* p, const tacd_backward_io* io, void* ws, size_t ws_bytes, cudaStream_t s) {
  using C = LtiCfg<NX, NU>;
  constexpr int warps = lti_warps<NX, NU, R>();
  const DeviceInfo di = device_info();
  int grid = (p->B + warps - 1) / warps;
  if (grid > di.sms) grid = di.sms;
  const size_t need = 256 + (size_t)grid * warps * lti_slice_words(p->T, NX, NU) * sizeof(R);
  if (ws_bytes < need) { set_error("backward workspace too small for the LTI kernel: %zu < %zu", ws_bytes, need); return TACD_EWORKSPACE; }
  const size_t smem = ((size_t)C::F_WORDS + (size_t)warps * C::WARP_WORDS) * sizeof(R);
  auto kernel = ilqr_backward_lti<R, NX, NU>;
  if (int rc = lti_prep_smem(kernel, smem)) return rc;
  if (int rc = check_cuda(cudaMemsetAsync(ws, 0, 16, s), "memset queue counter")) return rc;
  const size_t T = p->T, n = NX + NU, szX = (T + 1) * NX, szU = T * NU;
  if (io->grad_C && !p->C_batched) if (int rc = check_cuda(cudaMemsetAsync(io->grad_C, 0, T * n * n * sizeof(R), s), "memset")) return rc;
  if (io->grad_c && !p->C_batched) if (int rc = check_cuda(cudaMemsetAsync(io->grad_c, 0, T * n * sizeof(R), s), "memset")) return rc;
  if (io->grad_Cf && !p->Cf_batched) if (int rc = check_cuda(cudaMemsetAsync(io->grad_Cf, 0, n * n * sizeof(R), s), "memset")) return rc;
  if (io->grad_cf && !p->Cf_batched) if (int rc = check_cuda(cudaMemsetAsync(io->grad_cf, 0, n * sizeof(R), s), "memset")) return rc;
  if (io->grad_xref && !p->xref_batched) if (int rc = check_cuda(cudaMemsetAsync(io->grad_xref, 0, szX * sizeof(R), s), "memset")) return rc;
  if (io->grad_uref && !p->uref_batched) if (int rc = check_cuda(cudaMemsetAsync(io->grad_uref, 0, szU * sizeof(R), s), "memset")) return rc;
  BwdPtrs<R> f;
  f.kscratch = nullptr;
  f.X = (const R*)io->X; f.U = (const R*)io->U; f.C = (const R*)io->C; f.xref = (const R*)io->xref; f.uref = (const R*)io->uref;
  f.A = (const R*)io->A; f.Bm = (const R*)io->Bm; f.gX = (const R*)io->gX; f.gU = (const R*)io->gU; f.tight = io->tight;
  f.Cf = (const R*)io->Cf;
  f.grad_C = (R*)io->grad_C; f.grad_c = (R*)io->grad_c; f.grad_Cf = (R*)io->grad_Cf; f.grad_cf = (R*)io->grad_cf;
  f.grad_x0 = (R*)io->grad_x0; f.grad_xref = (R*)io->grad_xref; f.grad_uref = (R*)io->grad_uref; f.dX = (R*)io->dX; f.dU = (R*)io->dU;
  kernel<<<grid, warps * 32, smem, s>>>(to_dev<R>(p), f, (R*)((char*)ws + 256), (unsigned int*)ws);
  note_launches(1);
  return check_cuda(cudaPeekAtLastError(), "ilqr_backward_lti launch");
}

// the un-boxed reference backward only: input bounds (box QP) and compat="exact" stay on the run-time-dimension kernel
template <typename R>
int launch_backward_lti(const tacd_problem* p, const tacd_backward_io* io, void* ws, size_t ws_bytes, cudaStream_t s) {
  if (!lti_enabled() || p->dyn_id != TACD_DYN_LTI || !lti_size(p) || p->C_diag || p->has_umin || p->has_umax || p->compat_exact) return TACD_EUNSUPPORTED;
#define X(NX_, NU_) if (p->nx == NX_ && p->nu == NU_) return run_backward_lti<R, NX_, NU_>(p, io, ws, ws_bytes, s);
  TACD_LTI_SIZES(X)
#undef X
  return TACD_EUNSUPPORTED;
}

}  // namespace tacd
