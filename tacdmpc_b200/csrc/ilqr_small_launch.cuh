// ilqr_small_launch.cuh — dispatch + launch of the thread-per-element kernels.
// Included by ilqr_small_f32.cu / ilqr_small_f64.cu, which instantiate it for one scalar type each.
#pragma once
#include <stdlib.h>
#include <string.h>

#include "host_common.h"

namespace tacd {

// Pick the block size (multiple of 32, <= 256) that maximises resident threads per SM
// for a kernel whose dynamic shared memory is words_per_thread * NT + extra scalars.
template <typename K, typename SmemFn>
static int pick_block(K kernel, SmemFn smem_of, int* blocks_per_sm, size_t* smem_out, int max_blocks = 0) {
  const DeviceInfo di = device_info();
  int best_nt = 0, best_blocks = 0;
  size_t best_smem = 0;
  // experiment knobs (profiling only): cap the block size / resident blocks
  int nt_max = 256;
  { TACD_KNOB(k, "TACD_NT_MAX"); const int v = env_int(k, 0); if (v >= 32 && v <= 256) nt_max = v - v % 32; }
  for (int nt = nt_max; nt >= 32; nt -= 32) {
    const size_t smem = smem_of(nt);
    if (smem > (size_t)di.smem_optin) continue;
    if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) { cudaGetLastError(); continue; }
    int nb = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kernel, nt, smem) != cudaSuccess) { cudaGetLastError(); continue; }
    if (max_blocks > 0 && nb > max_blocks) nb = max_blocks;
    if (nb * nt > best_blocks * best_nt) { best_nt = nt; best_blocks = nb; best_smem = smem; }
  }
  { TACD_KNOB(k, "TACD_BLOCKS_MAX"); const int v = env_int(k, 0); if (v >= 1 && v < best_blocks) best_blocks = v; }
  *blocks_per_sm = best_blocks;
  *smem_out = best_smem;
  return best_nt;
}

template <typename R, int NX, int NU, int DYN, bool LSSEP>
static int run_forward_small(const tacd_problem* p, const tacd_forward_io* io, void* ws, size_t ws_bytes, cudaStream_t s, const FwdHandover* ho) {
  auto kernel = ilqr_forward_small<R, NX, NU, DYN, LSSEP>;
  int bps = 0;
  size_t smem = 0;
  const size_t wpt = (size_t)fwd_words(p->T, NX, NU);
  const int nt = pick_block(kernel, [&](int n) { return (wpt * n + NX * NX + NX * NU) * sizeof(R); }, &bps, &smem);
  if (nt == 0 || bps == 0) return TACD_EUNSUPPORTED;
  if (int rc = check_cuda(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "smem attr")) return rc;
  const DeviceInfo di = device_info();
  int grid = di.sms * bps;
  const int need = (p->B + nt - 1) / nt;
  if (grid > need) grid = need;
  if (grid < 1) grid = 1;
  FwdPtrs<R> f;
  memset(&f, 0, sizeof(f));
  if (ho) { f.cap_iters = ho->cap_iters; f.resume_count = ho->count; f.resume_list = ho->list; f.resume_cost = (R*)ho->cost; }
  f.x0 = (const R*)io->x0; f.C = (const R*)io->C; f.c = (const R*)io->c; f.Cf = (const R*)io->Cf; f.cf = (const R*)io->cf;
  f.C_ls = (const R*)io->C_ls; f.c_ls = (const R*)io->c_ls; f.Cf_ls = (const R*)io->Cf_ls; f.cf_ls = (const R*)io->cf_ls;
  f.xref = (const R*)io->xref; f.uref = (const R*)io->uref; f.Uinit = (const R*)io->Uinit;
  f.replay_iters = io->replay_iters; f.replay_alpha = io->replay_alpha; f.replay_flags = io->replay_flags;
  f.X = (R*)io->X; f.U = (R*)io->U; f.A = (R*)io->A; f.Bm = (R*)io->Bm;
  f.tight = io->tight; f.iters = io->iters; f.alpha_trace = io->alpha_trace; f.flags = io->flags;
  f.cost = (R*)io->cost; f.cost_trace = (R*)io->cost_trace;
  // workspace: [0,4) queue counter | [256, 256+8K) bucket counters | [16K, 16K+4B) longest-first order
  f.counter = (unsigned int*)ws;
  const unsigned lanes = (unsigned)grid * (unsigned)nt;
  const unsigned first_wave = lanes < (unsigned)p->B ? lanes : (unsigned)p->B;
  f.first_wave = first_wave;
  f.order = nullptr;
  f.cscratch = nullptr;
  f.cstride = (size_t)lanes;
  const size_t order_bytes = 16384 + (((size_t)p->B * 4 + 255) / 256) * 256;
  TACD_KNOB(k_nocs, "TACD_NO_CSCRATCH");
  if (LSSEP && p->C_batched && !env_set(k_nocs)) {
    const size_t need = order_bytes + (size_t)lanes * ((size_t)p->T * (NX + NU) * (NX + NU) + (size_t)p->T * (NX + NU)) * sizeof(R);
    if (ws_bytes >= need) f.cscratch = (R*)((char*)ws + order_bytes);
  }
  queue_init<<<1, 1, 0, s>>>(f.counter, first_wave);
  note_launches(1);
  TACD_KNOB(k_nolpt, "TACD_NO_LPT");
  const bool lpt = (unsigned)p->B > lanes && !env_set(k_nolpt);   // more than one wave: order matters
  if (lpt) {
    unsigned int* hist = (unsigned int*)((char*)ws + 256);
    int32_t* order = (int32_t*)((char*)ws + 16384);
    if (int rc = check_cuda(cudaMemsetAsync(hist, 0, kLptBuckets * sizeof(unsigned int), s), "memset hist")) return rc;
    const int g2 = di.sms * 2 < (p->B + 255) / 256 ? di.sms * 2 : (p->B + 255) / 256;
    lpt_histogram<R><<<g2, 256, 0, s>>>(p->B, p->T, NX, NX + NU, p->C_batched, p->xref_batched, (const R*)io->x0, (const R*)io->xref, (const R*)io->C, hist);
    note_launches(1);
    lpt_scan<<<1, 1024, 0, s>>>(hist);
    note_launches(1);
    lpt_scatter<R><<<g2, 256, 0, s>>>(p->B, p->T, NX, NX + NU, p->C_batched, p->xref_batched, (const R*)io->x0, (const R*)io->xref, (const R*)io->C, hist, order);
    note_launches(1);
    f.order = order;
  }
  kernel<<<grid, nt, smem, s>>>(to_dev<R>(p), f);
  note_launches(1);
  return check_cuda(cudaPeekAtLastError(), "ilqr_forward_small launch");
}

template <typename R, int NX, int NU, int DYN, bool DIAG, bool EXACT>
static int run_backward_small(const tacd_problem* p, const tacd_backward_io* io, void* ws, size_t ws_bytes, cudaStream_t s) {
  auto kernel = ilqr_backward_small<R, NX, NU, DYN, DIAG, EXACT>;
  using L = BwdLayout<NX, NU>;
  const int nacc = L::nacc(p->T);
  const bool reduce = (io->grad_C && !p->C_batched) || (io->grad_c && !p->C_batched) || (io->grad_Cf && !p->Cf_batched) ||
                      (io->grad_cf && !p->Cf_batched) || (io->grad_xref && !p->xref_batched) || (io->grad_uref && !p->uref_batched);
  int bps = 0;
  size_t smem = 0;
  const size_t wpt = (size_t)bwd_words(p->T, NX, NU);
  // the gains go to the workspace (coalesced [word][thread slot] columns) when it has room for two 256-thread CTAs per SM
  TACD_KNOB(k_ksmem, "TACD_BWD_K_SMEM");
  const size_t partial_cap = (size_t)device_info().sms * 2 * nacc * sizeof(R);
  const size_t kscratch_need = (size_t)device_info().sms * 2 * 256 * wpt * sizeof(R);
  // Measured at B = 65536, cart-pole (backward ms, gains in the workspace + 2 CTAs per SM / gains in shared memory + 1 CTA): shared
  // cost 0.123 / 0.138 (round 1) — taken; per-element dense 0.231 / 0.215 and compact diagonal 0.163 / 0.144 — not taken: those
  // layouts stream 2-6 KB of gradients per element to HBM and the extra 26 MB of gain traffic costs more than the second CTA hides.
  const bool kglob = reduce && !env_set(k_ksmem) && ws && ws_bytes >= ((partial_cap + 255) / 256) * 256 + kscratch_need;
  const int nt = pick_block(kernel, [&](int n) {
    size_t w = (kglob ? 0 : wpt * n) + NX * NX + NX * NU + (size_t)(n / 32) * L::TILE;   // [gains +] LTI matrices + one tile per warp
    if (reduce) w += (size_t)(n / 32) * nacc;                                   // per-warp accumulators
    return w * sizeof(R);
  }, &bps, &smem, kglob ? 2 : 1);
  if (nt == 0 || bps == 0) return TACD_EUNSUPPORTED;
  if (int rc = check_cuda(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "smem attr")) return rc;
  const DeviceInfo di = device_info();
  // balanced waves: work per element is uniform, so every resident CTA gets the same number of elements per
  // wave (vnt of its nt threads busy) instead of packing full CTAs onto a subset of the SMs
  int grid = di.sms * bps;
  const int need_ctas = (p->B + 31) / 32;
  if (grid > need_ctas) grid = need_ctas;
  if (grid < 1) grid = 1;
  const long waves = ((long)p->B + (long)grid * nt - 1) / ((long)grid * nt);
  const long per_cta = ((long)p->B + waves * grid - 1) / (waves * grid);
  int vnt = (int)((per_cta + 31) / 32) * 32;
  if (vnt > nt) vnt = nt;
  const size_t T = p->T, n = NX + NU;
  R* partials = nullptr;
  if (reduce) {
    const size_t need = (size_t)grid * nacc * sizeof(R);
    if (!ws || ws_bytes < need) { set_error("backward workspace too small: %zu < %zu", ws_bytes, need); return TACD_EWORKSPACE; }
    partials = (R*)ws;
    // the zero padding of the terminal gradients is not touched by the reduction
    if (io->grad_Cf && !p->Cf_batched) if (int rc = check_cuda(cudaMemsetAsync(io->grad_Cf, 0, n * n * sizeof(R), s), "memset")) return rc;
    if (io->grad_cf && !p->Cf_batched) if (int rc = check_cuda(cudaMemsetAsync(io->grad_cf, 0, n * sizeof(R), s), "memset")) return rc;
  }
  (void)T;
  BwdPtrs<R> f;
  f.X = (const R*)io->X; f.U = (const R*)io->U; f.C = (const R*)io->C; f.Cf = (const R*)io->Cf; f.xref = (const R*)io->xref; f.uref = (const R*)io->uref;
  f.A = (const R*)io->A; f.Bm = (const R*)io->Bm; f.gX = (const R*)io->gX; f.gU = (const R*)io->gU; f.tight = io->tight;
  f.grad_C = (R*)io->grad_C; f.grad_c = (R*)io->grad_c; f.grad_Cf = (R*)io->grad_Cf; f.grad_cf = (R*)io->grad_cf;
  f.grad_x0 = (R*)io->grad_x0; f.grad_xref = (R*)io->grad_xref; f.grad_uref = (R*)io->grad_uref; f.dX = (R*)io->dX; f.dU = (R*)io->dU;
  f.kscratch = kglob ? (R*)((char*)ws + ((partial_cap + 255) / 256) * 256) : nullptr;
  const DevProblem<R> dp = to_dev<R>(p);
  kernel<<<grid, nt, smem, s>>>(dp, f, partials, reduce ? 1 : 0, vnt);
  note_launches(1);
  if (int rc = check_cuda(cudaPeekAtLastError(), "ilqr_backward_small launch")) return rc;
  if (reduce) {
    reduce_shared_grads<R, NX, NU><<<(nacc * kReduceSeg + 127) / 128, 128, 0, s>>>(dp, f, partials, grid);
    note_launches(1);
    return check_cuda(cudaPeekAtLastError(), "reduce_shared_grads launch");
  }
  return TACD_OK;
}

// (nx, nu, dyn) combinations with a thread-per-element instantiation: the shipped plugins
#define TACD_SMALL_FWD(X) \
  X(1, 1, TACD_DYN_LTI) X(2, 1, TACD_DYN_LTI) X(4, 1, TACD_DYN_CARTPOLE) X(3, 2, TACD_DYN_UNICYCLE) X(3, 2, TACD_DYN_UNICYCLE_WRAPPED)
#define TACD_SMALL_BWD(X) \
  TACD_SMALL_FWD(X) X(1, 1, TACD_DYN_EXTERNAL) X(2, 1, TACD_DYN_EXTERNAL) X(4, 1, TACD_DYN_EXTERNAL) X(3, 2, TACD_DYN_EXTERNAL)

template <typename R>
int launch_forward_small(const tacd_problem* p, const tacd_forward_io* io, void* ws, size_t ws_bytes, cudaStream_t s, const FwdHandover* ho) {
#define X(NX_, NU_, DYN_)                                                                      \
  if (p->nx == NX_ && p->nu == NU_ && p->dyn_id == DYN_)                                       \
    return p->ls_cost_shared ? run_forward_small<R, NX_, NU_, DYN_, true>(p, io, ws, ws_bytes, s, ho)       \
                             : run_forward_small<R, NX_, NU_, DYN_, false>(p, io, ws, ws_bytes, s, ho);
  TACD_SMALL_FWD(X)
#undef X
  return TACD_EUNSUPPORTED;
}

template <typename R>
int launch_backward_small(const tacd_problem* p, const tacd_backward_io* io, void* ws, size_t ws_bytes, cudaStream_t s) {
#define X(NX_, NU_, DYN_) \
  if (p->nx == NX_ && p->nu == NU_ && p->dyn_id == DYN_) {                                    \
    if (p->compat_exact)                                                                      \
      return p->C_diag ? run_backward_small<R, NX_, NU_, DYN_, true, true>(p, io, ws, ws_bytes, s)    \
                       : run_backward_small<R, NX_, NU_, DYN_, false, true>(p, io, ws, ws_bytes, s);  \
    return p->C_diag ? run_backward_small<R, NX_, NU_, DYN_, true, false>(p, io, ws, ws_bytes, s)     \
                     : run_backward_small<R, NX_, NU_, DYN_, false, false>(p, io, ws, ws_bytes, s);   \
  }
  TACD_SMALL_BWD(X)
#undef X
  return TACD_EUNSUPPORTED;
}

}  // namespace tacd
