// capi.cu — the C ABI of include/tacd_ilqr.h: argument validation, dispatch
// between the thread-per-element kernels (shipped small plugins) and the generic
// CTA-per-element kernels, error reporting.  Nothing here synchronises the host.
#include <stdlib.h>
#include <string.h>

#include <atomic>

#include "host_common.h"

namespace tacd {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

// number of kernels of this library enqueued by the process (diagnostic; read by bench.py for gpu_launches)
static std::atomic<unsigned long long> g_launches{0};
void note_launches(int n) { g_launches.fetch_add((unsigned long long)n, std::memory_order_relaxed); }

static std::atomic<int> g_env_gen{0};
int env_generation() { return g_env_gen.load(std::memory_order_relaxed); }
void env_read(EnvKnob& k) {
  const char* e = getenv(k.name);
  k.set = e != nullptr;
  k.ival = e ? atoi(e) : 0;
  k.sval[0] = 0;
  if (e) { strncpy(k.sval, e, sizeof(k.sval) - 1); k.sval[sizeof(k.sval) - 1] = 0; }
  k.gen = env_generation();
}
bool env_is(EnvKnob& k, const char* value) { const EnvKnob& e = env_get(k); return e.set && !strcmp(e.sval, value); }

int check_cuda(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return TACD_OK;
  set_error("%s: %s", what, cudaGetErrorString(e));
  return TACD_ECUDA;
}

DeviceInfo device_info() {
  static thread_local int cached_dev = -1;
  static thread_local DeviceInfo di;
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev != cached_dev) {
    cudaDeviceGetAttribute(&di.sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&di.smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    cached_dev = dev;
  }
  return di;
}

SideStream side_stream() {
  static thread_local int cached_dev = -1;
  static thread_local SideStream ss = {nullptr, nullptr, nullptr, false};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev != cached_dev) {
    ss.ok = cudaStreamCreateWithFlags(&ss.stream, cudaStreamNonBlocking) == cudaSuccess &&
            cudaEventCreateWithFlags(&ss.fork, cudaEventDisableTiming) == cudaSuccess &&
            cudaEventCreateWithFlags(&ss.join, cudaEventDisableTiming) == cudaSuccess;
    if (!ss.ok) cudaGetLastError();
    cached_dev = dev;
  }
  return ss;
}

static int validate(const tacd_problem* p) {
  if (!p) { set_error("null problem"); return TACD_EINVAL; }
  if (p->B < 0 || p->T < 1 || p->nx < 1 || p->nu < 1) { set_error("bad sizes B=%d T=%d nx=%d nu=%d", p->B, p->T, p->nx, p->nu); return TACD_EINVAL; }
  if (p->nx > TACD_MAX_NX || p->nu > TACD_MAX_NU) { set_error("nx<=%d, nu<=%d supported", TACD_MAX_NX, TACD_MAX_NU); return TACD_EINVAL; }
  if (p->n_alpha < 1 || p->n_alpha > TACD_MAX_ALPHA) { set_error("n_alpha must be in 1..%d", TACD_MAX_ALPHA); return TACD_EINVAL; }
  if (p->dtype != TACD_F32 && p->dtype != TACD_F64) { set_error("dtype must be TACD_F32 or TACD_F64"); return TACD_EINVAL; }
  if (p->max_iter < 0) { set_error("max_iter < 0"); return TACD_EINVAL; }
  if (p->dyn_id == TACD_DYN_LTI && (!p->dyn_A || !p->dyn_B)) { set_error("LTI dynamics need dyn_A, dyn_B"); return TACD_EINVAL; }
  if (p->dyn_id == TACD_DYN_CARTPOLE && !(p->nx == 4 && p->nu == 1)) { set_error("cartpole is nx=4, nu=1"); return TACD_EINVAL; }
  if ((p->dyn_id == TACD_DYN_UNICYCLE || p->dyn_id == TACD_DYN_UNICYCLE_WRAPPED) && !(p->nx == 3 && p->nu == 2)) { set_error("unicycle is nx=3, nu=2"); return TACD_EINVAL; }
  if (p->dyn_id < 0 || p->dyn_id > TACD_DYN_EXTERNAL) { set_error("unknown dyn_id %d", p->dyn_id); return TACD_EINVAL; }
  if (p->C_diag && !(p->C_batched && p->Cf_batched)) { set_error("C_diag needs per-element costs (C_batched = Cf_batched = 1)"); return TACD_EINVAL; }
  return TACD_OK;
}

}  // namespace tacd

using namespace tacd;

extern "C" {

int tacd_abi_version(void) { return TACD_ABI_VERSION; }
const char* tacd_last_error(void) { return g_err; }
unsigned long long tacd_kernel_launches(void) { return g_launches.load(std::memory_order_relaxed); }
void tacd_refresh_env(void) { g_env_gen.fetch_add(1, std::memory_order_relaxed); }

// forward workspace: [0, kFwdHead) queue counter + bucket counters + longest-first order of the
// thread-per-element kernel; the generic kernel's per-CTA slices follow
static size_t fwd_head_bytes(const tacd_problem* p) {
  size_t b = 16384 + (((size_t)p->B * 4 + 255) / 256) * 256;
  b += (((size_t)p->B * 8 + 255) / 256) * 256;   // initial cost of every element (group kernel)
  b += ((group_forward_scratch_bytes(p) + 255) / 256) * 256;   // the line-search candidates' state tiles (group kernel)
  b += (((size_t)p->B * 4 + 255) / 256) * 256;   // hand-over list of the two-phase solve
  if (p->C_batched) {
    // lane-slot copy of the per-element (C, c): one slot per resident thread, bounded by shared memory
    const size_t es = p->dtype == TACD_F32 ? 4 : 8, n = p->nx + p->nu, T = p->T;
    const DeviceInfo di = device_info();
    const size_t sms = di.sms > 0 ? di.sms : 148, smem = di.smem_optin > 0 ? di.smem_optin : 232448;
    size_t per_sm = smem / ((size_t)fwd_words(p->T, p->nx, p->nu) * es);
    if (per_sm > 2048) per_sm = 2048;
    b += sms * per_sm * (T * n * n + T * n) * es + 256;
  }
  return b;
}
size_t tacd_forward_workspace_bytes(const tacd_problem* p) {
  if (validate(p)) return 0;
  const size_t g = generic_forward_ws_bytes(p), l = lti_ws_bytes(p);
  return fwd_head_bytes(p) + (g > l ? g : l);
}
// backward workspace: per-CTA partial sums of the shared-input gradients (thread-per-element kernel)
// or the generic kernel's per-CTA slices, whichever is larger
static size_t bwd_partials_bytes(const tacd_problem* p) {
  const size_t nx = p->nx, nu = p->nu, n = nx + nu, T = p->T;
  const size_t nacc = T * (n * n + n + nx + nu) + nx * nx + 2 * nx;
  const DeviceInfo di = device_info();
  const size_t sms = di.sms > 0 ? di.sms : 148;
  // + the gains of the thread-per-element KKT pass: T nu (nx + 1) words for each of 2 x 256 thread slots per SM
  return sms * 32 * nacc * 8 + 256 + sms * 2 * 256 * (T * nu * (nx + 1)) * 8;
}
size_t tacd_backward_workspace_bytes(const tacd_problem* p) {
  if (validate(p)) return 0;
  const size_t a = bwd_partials_bytes(p), g = generic_backward_ws_bytes(p), l = lti_ws_bytes(p), b = g > l ? g : l;
  return 256 + (a > b ? a : b);
}

int tacd_ilqr_forward(const tacd_problem* p, const tacd_forward_io* io, void* workspace, size_t workspace_bytes, void* stream) {
  if (int rc = validate(p)) return rc;
  if (p->B == 0) return TACD_OK;  // empty batch: nothing to read or write (empty tensors have null data)
  if (!io || !io->x0 || !io->C || !io->c || !io->Cf || !io->cf || !io->xref || !io->uref || !io->Uinit || !io->X || !io->U ||
      !io->tight || !io->iters || !io->flags) { set_error("tacd_ilqr_forward: required pointer is null"); return TACD_EINVAL; }
  if (p->ls_cost_shared && (!io->C_ls || !io->c_ls || !io->Cf_ls || !io->cf_ls)) { set_error("ls_cost_shared needs C_ls/c_ls/Cf_ls/cf_ls"); return TACD_EINVAL; }
  if (p->dyn_id == TACD_DYN_EXTERNAL) { set_error("the whole solve needs embedded dynamics (dyn_id != EXTERNAL)"); return TACD_EINVAL; }
  if (io->replay_iters && (!io->replay_alpha || !io->replay_flags)) { set_error("replay needs iters, alpha and flags"); return TACD_EINVAL; }
  const size_t head = fwd_head_bytes(p);
  if (!workspace || workspace_bytes < head) { set_error("workspace too small (need tacd_forward_workspace_bytes)"); return TACD_EWORKSPACE; }
  cudaStream_t s = (cudaStream_t)stream;
  int rc = (p->dtype == TACD_F32) ? launch_forward_group<float>(p, io, workspace, head, s) : launch_forward_group<double>(p, io, workspace, head, s);
  if (rc != TACD_EUNSUPPORTED) return rc;
  if (p->C_diag) { set_error("C_diag: this problem is not eligible for the five-lanes-per-element kernel; pass dense costs"); return TACD_EUNSUPPORTED; }
  rc = (p->dtype == TACD_F32) ? launch_forward_small<float>(p, io, workspace, head, s) : launch_forward_small<double>(p, io, workspace, head, s);
  if (rc != TACD_EUNSUPPORTED) return rc;
  char* ws = (char*)workspace + head;
  rc = (p->dtype == TACD_F32) ? launch_forward_lti<float>(p, io, ws, workspace_bytes - head, s)
                              : launch_forward_lti<double>(p, io, ws, workspace_bytes - head, s);
  if (rc != TACD_EUNSUPPORTED) return rc;
  rc = (p->dtype == TACD_F32) ? launch_forward_generic<float>(p, io, ws, workspace_bytes - head, s)
                              : launch_forward_generic<double>(p, io, ws, workspace_bytes - head, s);
  return rc;
}

int tacd_ilqr_backward(const tacd_problem* p, const tacd_backward_io* io, void* workspace, size_t workspace_bytes, void* stream) {
  if (int rc = validate(p)) return rc;
  if (p->B == 0) return TACD_OK;
  if (!io || !io->X || !io->U || !io->C || !io->xref || !io->uref || !io->tight || !io->gX || !io->gU) { set_error("tacd_ilqr_backward: required pointer is null"); return TACD_EINVAL; }
  if (p->dyn_id == TACD_DYN_EXTERNAL && (!io->A || !io->Bm)) { set_error("EXTERNAL dynamics need A and Bm"); return TACD_EINVAL; }
  if (p->has_umin && !p->has_umax) { set_error("backward assumes u_max whenever u_min is set (reference quirk Q10)"); return TACD_EINVAL; }
  if (p->compat_exact && !io->Cf) { set_error("compat_exact needs tacd_backward_io.Cf (the terminal Hessian)"); return TACD_EINVAL; }
  if (p->B == 0) return TACD_OK;
  cudaStream_t s = (cudaStream_t)stream;
  if (!workspace || workspace_bytes < 256) { set_error("workspace too small (need tacd_backward_workspace_bytes)"); return TACD_EWORKSPACE; }
  char* ws = (char*)workspace + 256;
  int rc = (p->dtype == TACD_F32) ? launch_backward_group<float>(p, io, ws, workspace_bytes - 256, s)
                                  : launch_backward_group<double>(p, io, ws, workspace_bytes - 256, s);
  if (rc != TACD_EUNSUPPORTED) return rc;
  rc = (p->dtype == TACD_F32) ? launch_backward_small<float>(p, io, ws, workspace_bytes - 256, s)
                              : launch_backward_small<double>(p, io, ws, workspace_bytes - 256, s);
  if (rc != TACD_EUNSUPPORTED) return rc;
  if (p->C_diag) { set_error("C_diag: no thread-per-element backward kernel for this (nx, nu, dyn); pass dense costs"); return TACD_EUNSUPPORTED; }
  rc = (p->dtype == TACD_F32) ? launch_backward_lti<float>(p, io, ws, workspace_bytes - 256, s)
                              : launch_backward_lti<double>(p, io, ws, workspace_bytes - 256, s);
  if (rc != TACD_EUNSUPPORTED) return rc;
  rc = (p->dtype == TACD_F32) ? launch_backward_generic<float>(p, io, ws, workspace_bytes - 256, s)
                              : launch_backward_generic<double>(p, io, ws, workspace_bytes - 256, s);
  return rc;
}

}  // extern "C"
