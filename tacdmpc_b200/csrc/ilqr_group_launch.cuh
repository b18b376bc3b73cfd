// ilqr_group_launch.cuh — dispatch + launch of the five-lanes-per-element forward kernel.
// Included by ilqr_group_f32.cu / ilqr_group_f64.cu, which instantiate it for one scalar type each.
#pragma once
#include <stdlib.h>
#include <string.h>

#include "host_common.h"
#include "ilqr_group.cuh"
#include "ilqr_group_bwd.cuh"

namespace tacd {

// workspace offsets shared by the forward kernels of the small plugins (capi.cu::fwd_head_bytes sizes them):
// [0,4) queue counter | [8,12) hand-over count | [16,20) queue counter of the resumed solves | [256, +8K) bucket counters |
// [16K, +4B) longest-first order | initial / hand-over cost of every element (8 B each) | state tiles | hand-over list (4 B each)
struct GroupWs { size_t order, cost0, scratch, list, end; };
template <typename R>
static GroupWs group_ws(const tacd_problem* p) {
  GroupWs w;
  w.order = 16384;
  w.cost0 = w.order + (((size_t)p->B * 4 + 255) / 256) * 256;
  w.scratch = w.cost0 + (((size_t)p->B * 8 + 255) / 256) * 256;
  w.list = w.scratch + ((group_forward_scratch_bytes(p) + 255) / 256) * 256;
  w.end = w.list + (((size_t)p->B * 4 + 255) / 256) * 256;
  return w;
}

// resume = the second phase of the two-phase solve: the queue holds the solves the thread-per-element kernel handed over
template <typename R, int NX, int NU, int DYN, bool LSSEP, bool CSH, bool DIAG>
static int run_forward_group(const tacd_problem* p, const tacd_forward_io* io, void* ws, size_t ws_bytes, cudaStream_t s, bool resume = false) {
  using KFn = void (*)(const DevProblem<R>, const FwdPtrs<R>, const R*);
  const DeviceInfo di = device_info();
  const size_t cost_b = (size_t)group_cost_words(p->T, NX, NU, p->C_batched, p->Cf_batched, LSSEP ? 1 : 0, DYN == TACD_DYN_LTI) * sizeof(R);
  // Per-element costs are staged in shared memory by bulk asynchronous copies when the blocks are 16-byte aligned AND the staged
  // blocks do not cost resident warps.  Measured at B = 65536, cart-pole T = 20 (forward ms, staged / read through L2): compact
  // diagonals 0.780 / 0.800 at 16 warps either way; dense [T,5,5] blocks 1.066 / 1.043 — 2.4 KB per element leave 10 warps per SM
  // instead of 16 and the lost latency hiding outweighs the saved L2 reads, so dense blocks stay in L2 (TACD_STAGE=1 forces staging).
  int stage = 0;
  if (!CSH && p->C_batched) {
    TACD_KNOB(k_nostage, "TACD_NO_STAGE");
    TACD_KNOB(k_stage, "TACD_STAGE");
    stage = group_stage_words(p->T, NX + NU, DIAG ? NX + NU : (NX + NU) * (NX + NU), sizeof(R));
    if (((uintptr_t)io->C | (uintptr_t)io->c) % 16 != 0 || env_set(k_nostage)) stage = 0;
    if (stage > 0) {
      const size_t avail = (size_t)di.smem_optin - cost_b;
      size_t w0 = avail / ((size_t)group_warp_words(p->T, NX, NU, 0) * sizeof(R)), w1 = avail / ((size_t)group_warp_words(p->T, NX, NU, stage) * sizeof(R));
      if (w0 > (size_t)group_max_warps<R>()) w0 = group_max_warps<R>();
      if (w1 > (size_t)group_max_warps<R>()) w1 = group_max_warps<R>();
      if (w1 < 1 || (w1 < w0 && !env_set(k_stage))) stage = 0;
    }
  }
  const size_t warp_b = (size_t)group_warp_words(p->T, NX, NU, stage) * sizeof(R);
  if (cost_b + warp_b > (size_t)di.smem_optin) return TACD_EUNSUPPORTED;   // horizon too long for the per-element buffers
  int nw = (int)(((size_t)di.smem_optin - cost_b) / warp_b);
  if (nw > group_max_warps<R>()) nw = group_max_warps<R>();   // __launch_bounds__(group_max_warps<R>() * 32, 1)
  { TACD_KNOB(k, "TACD_GROUP_WARPS"); const int v = env_int(k, 0); if (v >= 1 && v < nw) nw = v; }
  // small batches: spread the elements over all SMs before stacking warps on one (resumed solves: the long tail of the
  // iteration-count distribution, a few percent of the batch; their number is only known on the device)
  const int expect = resume ? (p->B + 15) / 16 : p->B;
  const int spread = (expect + di.sms * kEPW - 1) / (di.sms * kEPW);
  if (nw > spread) nw = spread < 1 ? 1 : spread;
  // Small batches: the candidates' state tiles move from the L2-resident workspace into shared memory when the batch is at most
  // ~1.7 waves of the warps that then fit (cart-pole T = 20 fp32: 18 KB per warp, 12 warps).  Such a launch lasts as long as its
  // longest chain of iterations, and every iteration of the L2 variant pays two dependent L2 round trips (tile write-out, winner
  // read-back).  Measured B = 4096 ActorMPC-shaped costs (37 iterations): 0.608 ms with L2 tiles, 0.513 ms with round 1's
  // all-shared layout; B = 65536 keeps the L2 tiles and 16 warps (0.626 vs 0.658 ms).
  bool ssm = false;
  {
    TACD_KNOB(k_ssm, "TACD_GROUP_SSM_SPREAD");
    const int lim = env_int(k_ssm, 20);
    const size_t warp_s = warp_b + group_xscratch_smem_words(p->T, NX, LSSEP) * sizeof(R);
    int nws = (int)(((size_t)di.smem_optin - cost_b) / warp_s);
    if (nws > group_max_warps<R>()) nws = group_max_warps<R>();
    if (!resume && nws >= 1 && spread <= lim) {
      ssm = true;
      if (nw > nws) nw = nws;
    }
  }
  const size_t smem = cost_b + (warp_b + (ssm ? group_xscratch_smem_words(p->T, NX, LSSEP) * sizeof(R) : 0)) * nw;
  const KFn kernel = ssm ? (KFn)ilqr_forward_group<R, NX, NU, DYN, LSSEP, CSH, DIAG, 0, true> : (KFn)ilqr_forward_group<R, NX, NU, DYN, LSSEP, CSH, DIAG, 0, false>;
  if (int rc = check_cuda(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "smem attr")) return rc;
  int grid = (expect + nw * kEPW - 1) / (nw * kEPW);
  if (grid > di.sms) grid = di.sms;
  if (grid < 1) grid = 1;
  // cost-mode variants (ilqr_group.cuh::group_cost_mode): a plain shared cost, full solves of more than one wave — there the
  // mode is decided on the helper stream, off the critical path, and the two variants that exit cost ~5 us against ~6 % of the
  // forward (0.640 -> 0.602 ms at B = 65536); a single-wave launch (B <= 14 k) would pay the three extra launches in line
  // (measured +4 us at B = 1024)
  constexpr bool kModes = CSH && !LSSEP && !DIAG;
  TACD_KNOB(k_nomodes, "TACD_NO_COST_MODES");
  TACD_KNOB(k_forcemodes, "TACD_FORCE_COST_MODES");   // tests: the variants on single-wave launches too
  const bool modes = kModes && !resume && ((long)p->B > (long)grid * nw * kEPW || env_set(k_forcemodes)) && !env_set(k_nomodes);
  if constexpr (kModes) {
    if (modes) {
      const KFn k1 = ssm ? (KFn)ilqr_forward_group<R, NX, NU, DYN, LSSEP, CSH, DIAG, 1, true> : (KFn)ilqr_forward_group<R, NX, NU, DYN, LSSEP, CSH, DIAG, 1, false>;
      const KFn k2 = ssm ? (KFn)ilqr_forward_group<R, NX, NU, DYN, LSSEP, CSH, DIAG, 2, true> : (KFn)ilqr_forward_group<R, NX, NU, DYN, LSSEP, CSH, DIAG, 2, false>;
      if (int rc = check_cuda(cudaFuncSetAttribute(k1, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "smem attr")) return rc;
      if (int rc = check_cuda(cudaFuncSetAttribute(k2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "smem attr")) return rc;
    }
  }
  const GroupWs wo = group_ws<R>(p);
  const size_t scratch_bytes = (size_t)grid * nw * group_xscratch_words(p->T, NX) * sizeof(R);
  if (ws_bytes < wo.end || wo.scratch + scratch_bytes > wo.list) { set_error("forward workspace too small for the initial costs / state tiles"); return TACD_EWORKSPACE; }
  R* cost0 = (R*)((char*)ws + wo.cost0);
  FwdPtrs<R> f;
  memset(&f, 0, sizeof(f));
  f.x0 = (const R*)io->x0; f.C = (const R*)io->C; f.c = (const R*)io->c; f.Cf = (const R*)io->Cf; f.cf = (const R*)io->cf;
  f.C_ls = (const R*)io->C_ls; f.c_ls = (const R*)io->c_ls; f.Cf_ls = (const R*)io->Cf_ls; f.cf_ls = (const R*)io->cf_ls;
  f.xref = (const R*)io->xref; f.uref = (const R*)io->uref; f.Uinit = (const R*)io->Uinit;
  f.replay_iters = io->replay_iters; f.replay_alpha = io->replay_alpha; f.replay_flags = io->replay_flags;
  f.X = (R*)io->X; f.U = (R*)io->U; f.A = (R*)io->A; f.Bm = (R*)io->Bm;
  f.tight = io->tight; f.iters = io->iters; f.alpha_trace = io->alpha_trace; f.flags = io->flags;
  f.cost = (R*)io->cost; f.cost_trace = (R*)io->cost_trace;
  f.counter = (unsigned int*)ws;
  f.order = nullptr;
  f.xscratch = ssm ? nullptr : (R*)((char*)ws + wo.scratch);
  f.stage_words = stage;
  f.cost_mode = modes ? (const int*)((char*)ws + 32) : nullptr;
  const DevProblem<R> dp = to_dev<R>(p);
  if (resume) {
    // the nominal trajectories, costs, iteration counts and flags of the handed-over solves are where the first phase left
    // them: io->X, io->U, cost0, io->iters, io->flags
    f.Uinit = (const R*)io->U;
    f.order = (const int32_t*)((char*)ws + wo.list);
    f.counter = (unsigned int*)((char*)ws + 16);
    f.n_dev = (const unsigned int*)((char*)ws + 8);
    f.resume = 1;
    kernel<<<grid, nw * 32, smem, s>>>(dp, f, cost0);
    note_launches(1);
    return check_cuda(cudaPeekAtLastError(), "ilqr_forward_group (resume) launch");
  }
  // longest-first queue order (ilqr_small.cuh): with more than one wave of elements the long chains must
  // not be popped last (profiles/r1_forward_group_v4: schedulers idle 13 % of the kernel in the tail)
  TACD_KNOB(k_nolpt, "TACD_NO_LPT");
  TACD_KNOB(k_nofork, "TACD_NO_FORK");
  const bool lpt = (long)p->B > (long)grid * nw * kEPW && !env_set(k_nolpt);
  // the queue order (3 small launches) and the initial rollout are independent: they run side by side, the
  // order on a helper stream that forks from / joins back into the caller's stream before the solve kernel
  SideStream ss = side_stream();
  const bool fork = lpt && ss.ok && !env_set(k_nofork);
  cudaStream_t s2 = s;
  if (fork) {
    if (int rc = check_cuda(cudaEventRecord(ss.fork, s), "fork record")) return rc;
    if (int rc = check_cuda(cudaStreamWaitEvent(ss.stream, ss.fork, 0), "fork wait")) return rc;
    s2 = ss.stream;
  }
  if (modes) {   // on the helper stream when there is one: off the critical path, next to the queue ordering
    group_cost_mode<R><<<1, 256, 0, s2>>>((const R*)io->C, (const R*)io->c, p->T, NX + NU, (int*)((char*)ws + 32));
    note_launches(1);
  }
  if (!lpt) {
    queue_init<<<1, 1, 0, s>>>(f.counter, 0u);
    note_launches(1);
  } else {
    unsigned int* hist = (unsigned int*)((char*)ws + 256);
    int32_t* order = (int32_t*)((char*)ws + 16384);
    if (int rc = check_cuda(cudaMemsetAsync(hist, 0, kLptBuckets * sizeof(unsigned int), s2), "memset hist")) return rc;
    const int g2 = di.sms * 2 < (p->B + 255) / 256 ? di.sms * 2 : (p->B + 255) / 256;
    const int nkey = p->C_diag ? -(NX + NU) : NX + NU;
    lpt_histogram<R><<<g2, 256, 0, s2>>>(p->B, p->T, NX, nkey, p->C_batched, p->xref_batched, (const R*)io->x0, (const R*)io->xref, (const R*)io->C, hist);
    note_launches(1);
    lpt_scan<<<1, 1024, 0, s2>>>(hist, f.counter, 0u);
    note_launches(1);
    lpt_scatter<R><<<g2, 256, 0, s2>>>(p->B, p->T, NX, nkey, p->C_batched, p->xref_batched, (const R*)io->x0, (const R*)io->xref, (const R*)io->C, hist, order);
    note_launches(1);
    f.order = order;
  }
  {
    auto init = ilqr_init_rollout<R, NX, NU, DYN>;
    // a horizon whose tile does not fit goes to the thread-per-element kernel
    // per-element costs: the whole horizon in one tile when one warp's tile stays under ~1/2 of shared memory
    // (two resident warps per SM keep enough bytes in flight), else four steps at a time, double-buffered
    // per-element costs scored by a separate line-search cost: the solve kernel evaluates the initial cost itself (row-split
    // over five lanes, from its staged cost block); here the inputs are only rolled out
    constexpr bool own_init = LSSEP && !CSH;
    int chunk = p->T;
    if (p->C_batched && init_warp_words(p->T, NX, NU, 1, p->C_diag, chunk) * sizeof(R) > (size_t)di.smem_optin / 2) chunk = p->T < 4 ? p->T : 4;
    // dense per-element costs: no cost tile, each thread reads its element's rows straight from global memory (chunk 0;
    // measured 2 % faster end to end than the whole-horizon tile, which leaves 2 CTAs per SM; compact diagonals: 1.5 % slower)
    if (p->C_batched && !p->C_diag) chunk = 0;
    { TACD_KNOB(k, "TACD_INIT_CHUNK"); const int v = env_int(k, -1); if (p->C_batched && v >= 0 && v <= p->T) chunk = v; }
    if (own_init) chunk = 0;
    const size_t tile_b = init_warp_words(p->T, NX, NU, p->C_batched, p->C_diag, chunk) * sizeof(R);
    // one warp per CTA: the finest granularity for the last wave (2048 warps over 148 SMs at B = 65536)
    if (tile_b > (size_t)di.smem_optin) return TACD_EUNSUPPORTED;
    int iw = 1;
    { TACD_KNOB(k, "TACD_INIT_WARPS"); const int v = env_int(k, 0); if (v >= 1 && v <= 4 && tile_b * v <= (size_t)di.smem_optin) iw = v; }
    if (int rc = check_cuda(cudaFuncSetAttribute(init, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(tile_b * iw)), "smem attr")) return rc;
    const int nwarp = (p->B + 31) / 32;
    init<<<(nwarp + iw - 1) / iw, iw * 32, tile_b * iw, s>>>(dp, f, own_init ? (R*)nullptr : cost0, chunk);
    note_launches(1);
  }
  if (int rc = check_cuda(cudaPeekAtLastError(), "ilqr_init_rollout launch")) return rc;
  if (fork) {
    if (int rc = check_cuda(cudaEventRecord(ss.join, ss.stream), "join record")) return rc;
    if (int rc = check_cuda(cudaStreamWaitEvent(s, ss.join, 0), "join wait")) return rc;
  }
  if constexpr (kModes) {
    if (modes) {   // the two variants whose mode does not match exit at once
      const KFn k1 = ssm ? (KFn)ilqr_forward_group<R, NX, NU, DYN, LSSEP, CSH, DIAG, 1, true> : (KFn)ilqr_forward_group<R, NX, NU, DYN, LSSEP, CSH, DIAG, 1, false>;
      const KFn k2 = ssm ? (KFn)ilqr_forward_group<R, NX, NU, DYN, LSSEP, CSH, DIAG, 2, true> : (KFn)ilqr_forward_group<R, NX, NU, DYN, LSSEP, CSH, DIAG, 2, false>;
      k2<<<grid, nw * 32, smem, s>>>(dp, f, cost0);
      k1<<<grid, nw * 32, smem, s>>>(dp, f, cost0);
      note_launches(2);
    }
  }
  kernel<<<grid, nw * 32, smem, s>>>(dp, f, cost0);
  note_launches(1);
  return check_cuda(cudaPeekAtLastError(), "ilqr_forward_group launch");
}

// Two-phase solve (EXPERIMENT, off by default: TACD_HYBRID_CAP=k).  Phase 1: the thread-per-element kernel (1.7x cheaper in issue
// slots per element-step: no redundant gains / tau / shuffles) runs every solve for at most k iterations; phase 2: this file's
// kernel resumes the few percent that are still improving (2.4 % of the bench workload at k = 8).  Round 1 attributed the thread
// kernel's 0.81 ms to its tail (iteration counts span 2..19, a thread's iteration is a 5x longer serial chain).  MEASURED in round
// 2 (B = 65536): the forward takes 0.82-0.85 ms for every k in 5..10, the same as the thread kernel alone — that kernel is
// latency-bound at 8 warps per SM (28 % of the issue slots), not tail-bound, so capping its chains buys nothing.  Kept, with the
// hand-over / resume plumbing, as the starting point for a thread kernel with twice the resident warps.
template <typename R, int NX, int NU, int DYN, bool LSSEP, bool CSH>
static int run_forward_hybrid(const tacd_problem* p, const tacd_forward_io* io, void* ws, size_t ws_bytes, cudaStream_t s, int cap) {
  const GroupWs wo = group_ws<R>(p);
  if (ws_bytes < wo.end) { set_error("forward workspace too small for the two-phase solve"); return TACD_EWORKSPACE; }
  if (int rc = check_cuda(cudaMemsetAsync((char*)ws + 8, 0, 16, s), "memset hand-over counters")) return rc;
  FwdHandover ho;
  ho.cap_iters = cap;
  ho.count = (unsigned int*)((char*)ws + 8);
  ho.list = (int32_t*)((char*)ws + wo.list);
  ho.cost = (char*)ws + wo.cost0;
  if (int rc = launch_forward_small<R>(p, io, ws, wo.cost0, s, &ho)) return rc;
  return run_forward_group<R, NX, NU, DYN, LSSEP, CSH, false>(p, io, ws, ws_bytes, s, true);
}

// (nx, nu, dyn) combinations with an instantiation: the shipped plugins (n = nx + nu <= kG)
#define TACD_GROUP_FWD(X) \
  X(1, 1, TACD_DYN_LTI) X(2, 1, TACD_DYN_LTI) X(4, 1, TACD_DYN_CARTPOLE) X(3, 2, TACD_DYN_UNICYCLE) X(3, 2, TACD_DYN_UNICYCLE_WRAPPED)

// Eligibility: analytic / auto-diff Jacobians of a shipped plugin (finite differences stay on the
// thread-per-element kernel), at most kG line-search candidates.
template <typename R>
int launch_forward_group(const tacd_problem* p, const tacd_forward_io* io, void* ws, size_t ws_bytes, cudaStream_t s) {
  if (p->jac_mode == TACD_JAC_FINITE_DIFF || p->n_alpha > kG) return TACD_EUNSUPPORTED;
  { TACD_KNOB(k, "TACD_FWD_KERNEL"); if (env_is(k, "thread")) return TACD_EUNSUPPORTED; }
  // two-phase solve (experiment, see run_forward_hybrid): shared running cost, batch of at least TACD_HYBRID_MIN_B solves,
  // TACD_HYBRID_CAP iterations in the first phase (default 0 = off)
  int cap = 0, min_b = 32768;
  { TACD_KNOB(k, "TACD_HYBRID_CAP"); cap = env_int(k, cap); }
  { TACD_KNOB(k, "TACD_HYBRID_MIN_B"); min_b = env_int(k, min_b); }
  const bool hybrid = cap > 0 && cap < p->max_iter && p->B >= min_b && !p->C_diag && !p->C_batched && !io->replay_iters && !io->cost_trace;
#define X(NX_, NU_, DYN_)                                                                      \
  if (p->nx == NX_ && p->nu == NU_ && p->dyn_id == DYN_) {                                     \
    if (hybrid)                                                                                \
      return p->ls_cost_shared ? run_forward_hybrid<R, NX_, NU_, DYN_, true, true>(p, io, ws, ws_bytes, s, cap)    \
                               : run_forward_hybrid<R, NX_, NU_, DYN_, false, true>(p, io, ws, ws_bytes, s, cap);  \
    if (p->C_diag)                                                                             \
      return p->ls_cost_shared ? run_forward_group<R, NX_, NU_, DYN_, true, false, true>(p, io, ws, ws_bytes, s)   \
                               : run_forward_group<R, NX_, NU_, DYN_, false, false, true>(p, io, ws, ws_bytes, s); \
    if (p->ls_cost_shared)                                                                     \
      return p->C_batched ? run_forward_group<R, NX_, NU_, DYN_, true, false, false>(p, io, ws, ws_bytes, s)   \
                          : run_forward_group<R, NX_, NU_, DYN_, true, true, false>(p, io, ws, ws_bytes, s);   \
    return p->C_batched ? run_forward_group<R, NX_, NU_, DYN_, false, false, false>(p, io, ws, ws_bytes, s)    \
                        : run_forward_group<R, NX_, NU_, DYN_, false, true, false>(p, io, ws, ws_bytes, s);    \
  }
  TACD_GROUP_FWD(X)
#undef X
  return TACD_EUNSUPPORTED;
}

template <typename R, int NX, int NU, int DYN, bool CSH, bool DIAG>
static int run_backward_group(const tacd_problem* p, const tacd_backward_io* io, void* ws, size_t ws_bytes, cudaStream_t s) {
  auto kernel = ilqr_backward_group<R, NX, NU, DYN, CSH, DIAG>;
  using L = BwdLayout<NX, NU>;
  const DeviceInfo di = device_info();
  const int nacc = L::nacc(p->T);
  const bool reduce = (io->grad_C && !p->C_batched) || (io->grad_c && !p->C_batched) || (io->grad_Cf && !p->Cf_batched) ||
                      (io->grad_cf && !p->Cf_batched) || (io->grad_xref && !p->xref_batched) || (io->grad_uref && !p->uref_batched);
  const size_t n = NX + NU;
  const size_t stage_b = ((CSH ? (size_t)p->T * ((n * n + n + 3) & ~(size_t)3) : 0) + (DYN == TACD_DYN_LTI ? ((NX * n + 3) & ~(size_t)3) : 0)) * sizeof(R);
  const size_t warp_b = ((size_t)gbwd_warp_words(p->T, NX, NU) + (reduce ? (size_t)nacc : 0)) * sizeof(R);
  if (stage_b + warp_b > (size_t)di.smem_optin) return TACD_EUNSUPPORTED;
  int nw = (int)(((size_t)di.smem_optin - stage_b) / warp_b);
  if (nw > 16) nw = 16;
  { TACD_KNOB(k, "TACD_GROUP_WARPS"); const int v = env_int(k, 0); if (v >= 1 && v < nw) nw = v; }
  const int spread = (p->B + di.sms * kEPW - 1) / (di.sms * kEPW);
  if (nw > spread) nw = spread < 1 ? 1 : spread;
  const size_t smem = stage_b + warp_b * nw;
  if (int rc = check_cuda(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "smem attr")) return rc;
  int grid = (p->B + nw * kEPW - 1) / (nw * kEPW);
  if (grid > di.sms) grid = di.sms;
  if (grid < 1) grid = 1;
  R* partials = nullptr;
  if (reduce) {
    const size_t need = (size_t)grid * nacc * sizeof(R);
    if (!ws || ws_bytes < need) { set_error("backward workspace too small: %zu < %zu", ws_bytes, need); return TACD_EWORKSPACE; }
    partials = (R*)ws;
    if (io->grad_Cf && !p->Cf_batched) if (int rc = check_cuda(cudaMemsetAsync(io->grad_Cf, 0, n * n * sizeof(R), s), "memset")) return rc;
    if (io->grad_cf && !p->Cf_batched) if (int rc = check_cuda(cudaMemsetAsync(io->grad_cf, 0, n * sizeof(R), s), "memset")) return rc;
  }
  BwdPtrs<R> f;
  f.kscratch = nullptr;
  f.X = (const R*)io->X; f.U = (const R*)io->U; f.C = (const R*)io->C; f.Cf = (const R*)io->Cf; f.xref = (const R*)io->xref; f.uref = (const R*)io->uref;
  f.A = (const R*)io->A; f.Bm = (const R*)io->Bm; f.gX = (const R*)io->gX; f.gU = (const R*)io->gU; f.tight = io->tight;
  f.grad_C = (R*)io->grad_C; f.grad_c = (R*)io->grad_c; f.grad_Cf = (R*)io->grad_Cf; f.grad_cf = (R*)io->grad_cf;
  f.grad_x0 = (R*)io->grad_x0; f.grad_xref = (R*)io->grad_xref; f.grad_uref = (R*)io->grad_uref; f.dX = (R*)io->dX; f.dU = (R*)io->dU;
  const DevProblem<R> dp = to_dev<R>(p);
  kernel<<<grid, nw * 32, smem, s>>>(dp, f, partials, reduce ? 1 : 0);
  note_launches(1);
  if (int rc = check_cuda(cudaPeekAtLastError(), "ilqr_backward_group launch")) return rc;
  if (reduce) {
    reduce_shared_grads<R, NX, NU><<<(nacc * kReduceSeg + 127) / 128, 128, 0, s>>>(dp, f, partials, grid);
    note_launches(1);
    return check_cuda(cudaPeekAtLastError(), "reduce_shared_grads launch");
  }
  return TACD_OK;
}

// Eligibility: embedded dynamics of a shipped plugin, reference-mode gradients (compat="exact" and external
// Jacobians stay on the thread-per-element kernel).  OPT-IN (TACD_BWD_KERNEL=group): unlike the forward solve
// the backward does uniform work per element, so it has no tail to remove, and the redundant gains / box QP on
// five lanes plus the cross-group reduction cost more than coalescing saves — measured at B = 65536, cart-pole:
// 0.217 ms vs 0.146 ms (shared cost), 0.181 vs 0.135 (compact diagonal), 0.220 vs 0.210 (per-element dense);
// equal at B <= 4096.  Kept for the parity tests and as the starting point for a leaner row split.
template <typename R>
int launch_backward_group(const tacd_problem* p, const tacd_backward_io* io, void* ws, size_t ws_bytes, cudaStream_t s) {
  if (p->compat_exact || p->dyn_id == TACD_DYN_EXTERNAL) return TACD_EUNSUPPORTED;
  { TACD_KNOB(k, "TACD_BWD_KERNEL"); if (!env_is(k, "group")) return TACD_EUNSUPPORTED; }
#define X(NX_, NU_, DYN_)                                                                      \
  if (p->nx == NX_ && p->nu == NU_ && p->dyn_id == DYN_) {                                     \
    if (p->C_diag) return run_backward_group<R, NX_, NU_, DYN_, false, true>(p, io, ws, ws_bytes, s);   \
    return p->C_batched ? run_backward_group<R, NX_, NU_, DYN_, false, false>(p, io, ws, ws_bytes, s)   \
                        : run_backward_group<R, NX_, NU_, DYN_, true, false>(p, io, ws, ws_bytes, s);   \
  }
  TACD_GROUP_FWD(X)
#undef X
  return TACD_EUNSUPPORTED;
}

}  // namespace tacd
