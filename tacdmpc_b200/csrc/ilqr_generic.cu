// ilqr_generic.cu — CTA- / warp-per-element whole-solve kernels for run-time dimensions,
// plus the stand-alone stage entry points of the C ABI (rollout, linearize,
// objective, riccati, linesearch, pnqp).  See ilqr_generic.cuh for the layout.
#include <stdlib.h>

#include "host_common.h"
#include "ilqr_generic.cuh"

// The whole-solve kernels are instantiated for 5 sizes x 2 element mappings x 2 dtypes; to keep the build parallel the
// file is compiled four times (build.py): part 0 = fp32 forward + the stand-alone stages + the C entry points (this
// file as is), parts 1-3 = fp32 backward / fp64 forward / fp64 backward through ilqr_generic_part{1,2,3}.cu.
#ifndef TACD_GEN_PART
#define TACD_GEN_PART 0
#endif

namespace tacd {

// per-CTA slice of the workspace: X[(T+1)nx] U[T nu] K[T nu nx] k[T nu]
static inline size_t gen_slice_scalars(const tacd_problem* p) {
  const size_t nx = p->nx, nu = p->nu, T = p->T;
  return (T + 1) * nx + T * nu + T * nu * nx + T * nu;
}
// Threads per element (GenDims::tpe): one warp for systems up to nx = 16, the whole CTA above.
static int gen_tpe(const tacd_problem* p) {
  TACD_KNOB(k, "TACD_GENERIC_TPE");
  if (env_set(k)) return env_int(k, 0) == 32 ? 32 : kGenThreads;
  return p->nx <= 16 ? 32 : kGenThreads;
}
// element slots in flight: four CTAs per SM, kGenThreads / tpe elements per CTA
static int gen_ctas_per_sm() {
  TACD_KNOB(k, "TACD_GENERIC_CTAS");
  const int v = env_int(k, 4);
  return v >= 1 && v <= 16 ? v : 4;
}
static int gen_slots(const tacd_problem* p, int tpe) {
  const DeviceInfo di = device_info();
  int g = di.sms * gen_ctas_per_sm() * (kGenThreads / tpe);
  if (g > p->B) g = p->B;
  return g < 1 ? 1 : g;
}
static int gen_grid(const tacd_problem* p) { return gen_slots(p, kGenThreads); }
#if TACD_GEN_PART == 0
size_t generic_forward_ws_bytes(const tacd_problem* p) {
  const size_t es = (p->dtype == TACD_F32) ? 4 : 8;
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const size_t g = (size_t)sms * gen_ctas_per_sm() * (kGenThreads / gen_tpe(p));
  return g * gen_slice_scalars(p) * es + 256;
}
size_t generic_backward_ws_bytes(const tacd_problem* p) { return generic_forward_ws_bytes(p); }
#endif

template <typename R, int NXT = 0, int NUT = 0, int TPE = kGenThreads>
static GenDims<R, NXT, NUT, TPE> dims_of(const tacd_problem* p) {
  GenDims<R, NXT, NUT, TPE> d;
  d.nx_ = p->nx; d.nu_ = p->nu; d.T = p->T; d.dyn = p->dyn_id;
  return d;
}

// ---------------------------------------------------------------------------
// forward solve (solve_step, controller.py:275-315)
template <typename R, int NXT, int NUT, int TPE>
__global__ void __launch_bounds__(kGenThreads, sizeof(R) == 4 ? 4 : 2) ilqr_forward_generic(const DevProblem<R> P, const GenDims<R, NXT, NUT, TPE> d, const FwdPtrs<R> io, R* ws, size_t slice) {
  using DT = GenDims<R, NXT, NUT, TPE>;
  extern __shared__ __align__(16) unsigned char smem_cta[];
  // each element slot of the CTA owns [matrices | (X U K k when ws == nullptr)]
  unsigned char* smem_raw = smem_cta + (size_t)DT::sub() * gen_elem_bytes<R>(d.nx(), d.nu(), ws ? 0 : slice);
  GenSmem<R> s = gen_carve<R>(smem_raw, d.nx(), d.nu());
  const int tid = DT::tid(), nx = d.nx(), nu = d.nu(), n = d.n(), T = d.T;
  const size_t szX = (size_t)(T + 1) * nx, szU = (size_t)T * nu;
  // X, U, K, k of the element in flight: in shared memory right behind the per-step matrices when they fit
  // (ws == nullptr), else in this CTA's slice of the caller's workspace
  R* gXs = ws ? ws + (size_t)DT::slot() * slice : reinterpret_cast<R*>(smem_raw + gen_smem_bytes_aligned<R>(nx, nu));
  R* gUs = gXs + szX;
  R* gK = gUs + szU;
  R* gk = gK + (size_t)T * nu * nx;
  if (d.dyn == TACD_DYN_LTI) {
    for (int i = tid; i < nx * nx; i += DT::tpe) s.A[i] = P.dyn_A[i];
    for (int i = tid; i < nx * nu; i += DT::tpe) s.Bm[i] = P.dyn_B[i];
  }
  DT::sync();
  const bool lssep = P.ls_cost_shared != 0;
  for (int b = DT::slot(); b < P.B; b += DT::nslots()) {
    const R* Cp = io.C + (P.C_batched ? (size_t)b * T * n * n : 0);
    const R* cp = io.c + (P.C_batched ? (size_t)b * T * n : 0);
    const R* Cfp = io.Cf + (P.Cf_batched ? (size_t)b * n * n : 0);
    const R* cfp = io.cf + (P.Cf_batched ? (size_t)b * n : 0);
    const R* Cl = lssep ? io.C_ls : Cp;
    const R* cl = lssep ? io.c_ls : cp;
    const R* Cfl = lssep ? io.Cf_ls : Cfp;
    const R* cfl = lssep ? io.cf_ls : cfp;
    const R* xr = io.xref + (P.xref_batched ? (size_t)b * szX : 0);
    const R* ur = io.uref + (P.uref_batched ? (size_t)b * szU : 0);
    // ---- init: U = U_init, X = rollout (no clamping), best = objective
    for (int i = tid; i < (int)szU; i += DT::tpe) gUs[i] = io.Uinit[(size_t)b * szU + i];
    for (int i = tid; i < nx; i += DT::tpe) gXs[i] = io.x0[(size_t)b * nx + i];
    if (io.alpha_trace) for (int j = tid; j < P.max_iter; j += DT::tpe) io.alpha_trace[(size_t)b * P.max_iter + j] = 0xFF;
    DT::sync();
    for (int t = 0; t < T; ++t) {
      gen_dyn_step<R>(P, d, s.A, s.Bm, 1, gXs + (size_t)t * nx, gUs + (size_t)t * nu, gXs + (size_t)(t + 1) * nx);
      DT::sync();
    }
    R best = gen_objective<R>(d, s, gXs, gUs, Cp, cp, Cfp, cfp, xr, ur);
    unsigned flg = 0;
    int it = 0;
    const bool replay = io.replay_iters != nullptr;
    while (it < P.max_iter && !(replay && it >= io.replay_iters[b])) {
      // ---- reverse Riccati pass
      for (int i = tid; i < nx; i += DT::tpe) s.tau[i] = gXs[(size_t)T * nx + i] - xr[T * nx + i];
      DT::sync();
      for (int idx = tid; idx < nx * nx; idx += DT::tpe) s.V[idx] = Cfp[(idx / nx) * n + (idx % nx)];
      for (int i = tid; i < nx; i += DT::tpe) {
        R a = 0;
        for (int j = 0; j < nx; ++j) a += Cfp[i * n + j] * s.tau[j];
        s.v[i] = a + cfp[i];
      }
      DT::sync();
      for (int t = T - 1; t >= 0; --t) {
        for (int i = tid; i < nx; i += DT::tpe) { s.x[i] = gXs[(size_t)t * nx + i]; s.tau[i] = s.x[i] - xr[t * nx + i]; }
        for (int i = tid; i < nu; i += DT::tpe) { s.u[i] = gUs[(size_t)t * nu + i]; s.tau[nx + i] = s.u[i] - ur[t * nu + i]; }
        for (int idx = tid; idx < n * n; idx += DT::tpe) s.Ct[idx] = Cp[(size_t)t * n * n + idx];
        DT::sync();
        for (int i = tid; i < n; i += DT::tpe) {
          R a = 0;
          for (int j = 0; j < n; ++j) a += s.Ct[i * n + j] * s.tau[j];
          a += cp[(size_t)t * n + i];
          if (i < nx) s.qx[i] = a; else s.qu[i - nx] = a;
        }
        for (int idx = tid; idx < nx * nx; idx += DT::tpe) s.Qxx[idx] = s.Ct[(idx / nx) * n + (idx % nx)];
        for (int idx = tid; idx < nx * nu; idx += DT::tpe) s.Qxu[idx] = s.Ct[(idx / nu) * n + nx + (idx % nu)];
        for (int idx = tid; idx < nu * nu; idx += DT::tpe) s.Quu[idx] = s.Ct[(nx + idx / nu) * n + nx + (idx % nu)];
        gen_dyn_jac<R>(P, d, s.x, s.u, s.A, s.Bm);
        DT::sync();
        if (gen_value_step<R>(P, d, s, false, nullptr, nullptr)) flg |= TACD_FLAG_NONPD;
        for (int idx = tid; idx < nu * nx; idx += DT::tpe) gK[(size_t)t * nu * nx + idx] = s.Kt[idx];
        for (int i = tid; i < nu; i += DT::tpe) gk[(size_t)t * nu + i] = s.kt[i];
        DT::sync();
      }
      // ---- all line-search candidates (evaluate_alphas, controller.py:589-620)
      const int na = P.n_alpha;
      for (int idx = tid; idx < na * nx; idx += DT::tpe) s.xa[idx] = gXs[idx % nx];
      for (int idx = tid; idx < na * n; idx += DT::tpe) { s.part1[idx] = 0; s.part2[idx] = 0; }
      DT::sync();
      for (int t = 0; t < T; ++t) {
        for (int idx = tid; idx < na * nu; idx += DT::tpe) {
          const int a = idx / nu, i = idx % nu;
          R acc = 0;
          for (int j = 0; j < nx; ++j) acc += gK[(size_t)t * nu * nx + i * nx + j] * (s.xa[a * nx + j] - gXs[(size_t)t * nx + j]);
          R ui = gUs[(size_t)t * nu + i] + (P.alphas[a] * gk[(size_t)t * nu + i] + acc);
          if (P.has_umin) ui = tmax<R>(ui, P.umin[i]);
          if (P.has_umax) ui = tmin<R>(ui, P.umax[i]);
          s.ua[idx] = ui;
        }
        DT::sync();
        // the step's cost rows (scoring cost, and the element's own when they differ) and every candidate's tau go
        // through shared memory: one coalesced read of C_t instead of one row read per (candidate, row) thread
        for (int idx = tid; idx < n * n; idx += DT::tpe) {
          s.Ct[idx] = Cl[(size_t)t * n * n + idx];
          if (lssep) s.Ct2[idx] = Cp[(size_t)t * n * n + idx];
        }
        for (int idx = tid; idx < na * n; idx += DT::tpe) {
          const int a = idx / n, i = idx % n;
          s.taua[idx] = (i < nx) ? (s.xa[a * nx + i] - xr[t * nx + i]) : (s.ua[a * nu + (i - nx)] - ur[t * nu + (i - nx)]);
        }
        DT::sync();
        for (int idx = tid; idx < na * n; idx += DT::tpe) {
          const int a = idx / n, i = idx % n;
          const R* Ct = s.Ct + i * n;
          const R* ta = s.taua + a * n;
          R sd = 0;
          for (int j = 0; j < n; ++j) sd += Ct[j] * ta[j];
          const R ti = ta[i];
          s.part1[idx] += (R)0.5 * (ti * sd) + cl[(size_t)t * n + i] * ti;
          if (lssep) {
            const R* Co = s.Ct2 + i * n;
            R so = 0;
            for (int j = 0; j < n; ++j) so += Co[j] * ta[j];
            s.part2[idx] += (R)0.5 * (ti * so) + cp[(size_t)t * n + i] * ti;
          }
        }
        gen_dyn_step<R>(P, d, s.A, s.Bm, na, s.xa, s.ua, s.xan);
        DT::sync();
        for (int idx = tid; idx < na * nx; idx += DT::tpe) s.xa[idx] = s.xan[idx];
        DT::sync();
      }
      // terminal terms + per-candidate totals (thread a sums its candidate)
      if (tid < na) {
        const int a = tid;
        R tot = 0, tot2 = 0;
        for (int i = 0; i < n; ++i) { tot += s.part1[a * n + i]; tot2 += s.part2[a * n + i]; }
        R qN = 0, lN = 0, qN2 = 0, lN2 = 0;
        for (int i = 0; i < nx; ++i) {
          const R ti = s.xa[a * nx + i] - xr[T * nx + i];
          R sd = 0, so = 0;
          for (int j = 0; j < nx; ++j) {
            const R tj = s.xa[a * nx + j] - xr[T * nx + j];
            sd += Cfl[i * n + j] * tj;
            if (lssep) so += Cfp[i * n + j] * tj;
          }
          qN += ti * sd; lN += cfl[i] * ti;
          if (lssep) { qN2 += ti * so; lN2 += cfp[i] * ti; }
        }
        s.red[a] = tot + ((R)0.5 * qN + lN);
        s.red[TACD_MAX_ALPHA + a] = lssep ? tot2 + ((R)0.5 * qN2 + lN2) : s.red[a];
      }
      DT::sync();
      int ai = 0;
      for (int a = 1; a < na; ++a) {
        if (s.red[ai] != s.red[ai]) break;
        if (s.red[a] < s.red[ai] || s.red[a] != s.red[a]) ai = a;
      }
      if (replay) ai = io.replay_alpha[(size_t)b * P.max_iter + it];
      const R newc = s.red[TACD_MAX_ALPHA + ai];
      bool improved = newc < best;
      if (replay) improved = (it < io.replay_iters[b] - 1) || ((io.replay_flags[b] & TACD_FLAG_MAXITER) != 0);
      if (tid == 0) {
        if (io.alpha_trace) io.alpha_trace[(size_t)b * P.max_iter + it] = (uint8_t)ai;
        if (io.cost_trace) {
          R* ctr = io.cost_trace + ((size_t)b * P.max_iter + it) * (na + 2);
          for (int a = 0; a < na; ++a) ctr[a] = s.red[a];
          ctr[na] = newc;
          ctr[na + 1] = best;
        }
      }
      DT::sync();
      it += 1;
      if (!improved) break;
      best = newc;
      // ---- re-rollout of the accepted candidate in place
      const R alpha = P.alphas[ai];
      for (int i = tid; i < nx; i += DT::tpe) { s.xa[i] = gXs[i]; s.xan[nx + i] = gXs[i]; }  // xan[nx..] = old X_t
      DT::sync();
      for (int t = 0; t < T; ++t) {
        for (int i = tid; i < nu; i += DT::tpe) {
          R acc = 0;
          for (int j = 0; j < nx; ++j) acc += gK[(size_t)t * nu * nx + i * nx + j] * (s.xa[j] - s.xan[nx + j]);
          R ui = gUs[(size_t)t * nu + i] + (alpha * gk[(size_t)t * nu + i] + acc);
          if (P.has_umin) ui = tmax<R>(ui, P.umin[i]);
          if (P.has_umax) ui = tmin<R>(ui, P.umax[i]);
          s.ua[i] = ui;
        }
        DT::sync();
        for (int i = tid; i < nu; i += DT::tpe) gUs[(size_t)t * nu + i] = s.ua[i];
        gen_dyn_step<R>(P, d, s.A, s.Bm, 1, s.xa, s.ua, s.xan);
        DT::sync();
        for (int i = tid; i < nx; i += DT::tpe) {
          s.xan[nx + i] = gXs[(size_t)(t + 1) * nx + i];
          s.xa[i] = s.xan[i];
          gXs[(size_t)(t + 1) * nx + i] = s.xan[i];
        }
        DT::sync();
      }
      if (it >= P.max_iter) flg |= TACD_FLAG_MAXITER;
    }
    // ---- finalise
    for (int i = tid; i < (int)szX; i += DT::tpe) io.X[(size_t)b * szX + i] = gXs[i];
    for (int i = tid; i < (int)szU; i += DT::tpe) {
      const R ui = gUs[i];
      io.U[(size_t)b * szU + i] = ui;
      bool m = false;
      if (P.has_umin) m = m || isclose_<R>(ui, P.umin[i % nu]);
      if (P.has_umax) m = m || isclose_<R>(ui, P.umax[i % nu]);
      io.tight[(size_t)b * szU + i] = m ? 1 : 0;
    }
    if (io.A != nullptr || io.Bm != nullptr) {
      for (int t = 0; t < T; ++t) {
        DT::sync();
        for (int i = tid; i < nx; i += DT::tpe) s.x[i] = gXs[(size_t)t * nx + i];
        for (int i = tid; i < nu; i += DT::tpe) s.u[i] = gUs[(size_t)t * nu + i];
        DT::sync();
        gen_dyn_jac<R>(P, d, s.x, s.u, s.A, s.Bm);
        DT::sync();
        if (io.A) for (int i = tid; i < nx * nx; i += DT::tpe) io.A[((size_t)b * T + t) * nx * nx + i] = s.A[i];
        if (io.Bm) for (int i = tid; i < nx * nu; i += DT::tpe) io.Bm[((size_t)b * T + t) * nx * nu + i] = s.Bm[i];
      }
    }
    if (tid == 0) {
      io.iters[b] = it;
      io.flags[b] = (uint8_t)flg;
      if (io.cost) io.cost[b] = best;
    }
    DT::sync();
  }
}

// ---------------------------------------------------------------------------
// backward (ILQRSolve.backward + _zero_constrained_lqr, controller.py:63-115, 707-788)
template <typename R, int NXT, int NUT, int TPE>
__global__ void __launch_bounds__(kGenThreads, sizeof(R) == 4 ? 4 : 2) ilqr_backward_generic(const DevProblem<R> P, const GenDims<R, NXT, NUT, TPE> d, const BwdPtrs<R> io, R* ws, size_t slice) {
  using DT = GenDims<R, NXT, NUT, TPE>;
  extern __shared__ __align__(16) unsigned char smem_cta[];
  // each element slot of the CTA owns [matrices | (X U K k when ws == nullptr)]
  unsigned char* smem_raw = smem_cta + (size_t)DT::sub() * gen_elem_bytes<R>(d.nx(), d.nu(), ws ? 0 : slice);
  GenSmem<R> s = gen_carve<R>(smem_raw, d.nx(), d.nu());
  const int tid = DT::tid(), nx = d.nx(), nu = d.nu(), n = d.n(), T = d.T;
  const size_t szX = (size_t)(T + 1) * nx, szU = (size_t)T * nu;
  R* gK = ws ? ws + (size_t)DT::slot() * slice : reinterpret_cast<R*>(smem_raw + gen_smem_bytes_aligned<R>(nx, nu));
  R* gk = gK + (size_t)T * nu * nx;
  if (d.dyn == TACD_DYN_LTI) {
    for (int i = tid; i < nx * nx; i += DT::tpe) s.A[i] = P.dyn_A[i];
    for (int i = tid; i < nx * nu; i += DT::tpe) s.Bm[i] = P.dyn_B[i];
  }
  DT::sync();
  for (int b = DT::slot(); b < P.B; b += DT::nslots()) {
    const R* X = io.X + (size_t)b * szX;
    const R* U = io.U + (size_t)b * szU;
    const R* Cp = io.C + (P.C_batched ? (size_t)b * T * n * n : 0);
    const R* xr = io.xref + (P.xref_batched ? (size_t)b * szX : 0);
    const R* ur = io.uref + (P.uref_batched ? (size_t)b * szU : 0);
    const R* gX = io.gX + (size_t)b * szX;
    const R* gU = io.gU + (size_t)b * szU;
    const uint8_t* tg = io.tight + (size_t)b * szU;
    auto load_AB = [&](int t) {
      if (d.dyn == TACD_DYN_EXTERNAL) {
        for (int i = tid; i < nx * nx; i += DT::tpe) s.A[i] = io.A[((size_t)b * T + t) * nx * nx + i];
        for (int i = tid; i < nx * nu; i += DT::tpe) s.Bm[i] = io.Bm[((size_t)b * T + t) * nx * nu + i];
      } else if (d.dyn != TACD_DYN_LTI) {
        for (int i = tid; i < nx; i += DT::tpe) s.x[i] = X[t * nx + i];
        for (int i = tid; i < nu; i += DT::tpe) s.u[i] = U[t * nu + i];
        DT::sync();
        gen_dyn_jac<R>(P, d, s.x, s.u, s.A, s.Bm);
      }
      DT::sync();
    };
    const bool exact = P.exact != 0;   // compat="exact" (SURVEY 8f N3): see gen_value_step and tests/exact_ref.py
    const R* Cfe = exact ? io.Cf + (P.Cf_batched ? (size_t)b * n * n : 0) : nullptr;
    if (exact) {   // terminal value C_final[:nx,:nx], -grad_X[T] (fixes Q1, Q2)
      for (int idx = tid; idx < nx * nx; idx += DT::tpe) s.V[idx] = Cfe[(idx / nx) * n + (idx % nx)];
      for (int i = tid; i < nx; i += DT::tpe) s.v[i] = -gX[T * nx + i];
    } else {
      for (int idx = tid; idx < nx * nx; idx += DT::tpe) s.V[idx] = Cp[(size_t)(T - 1) * n * n + (idx / nx) * n + (idx % nx)];
      for (int i = tid; i < nx; i += DT::tpe) s.v[i] = -gX[(T - 1) * nx + i];
    }
    DT::sync();
    for (int t = T - 1; t >= 0; --t) {
      const R* Ct = Cp + (size_t)t * n * n;
      for (int idx = tid; idx < nx * nx; idx += DT::tpe) s.Qxx[idx] = Ct[(idx / nx) * n + (idx % nx)];
      for (int idx = tid; idx < nx * nu; idx += DT::tpe) s.Qxu[idx] = Ct[(idx / nu) * n + nx + (idx % nu)];
      for (int idx = tid; idx < nu * nu; idx += DT::tpe) s.Quu[idx] = Ct[(nx + idx / nu) * n + nx + (idx % nu)];
      for (int i = tid; i < nx; i += DT::tpe) s.qx[i] = -gX[t * nx + i];
      for (int i = tid; i < nu; i += DT::tpe) s.qu[i] = -gU[t * nu + i];
      load_AB(t);
      gen_value_step<R>(P, d, s, true, tg + (size_t)t * nu, U + (size_t)t * nu);
      for (int idx = tid; idx < nu * nx; idx += DT::tpe) gK[(size_t)t * nu * nx + idx] = s.Kt[idx];
      for (int i = tid; i < nu; i += DT::tpe) gk[(size_t)t * nu + i] = s.kt[i];
      DT::sync();
    }
    // forward sweep + gradients; s.x = dx_t, s.u = du_t, s.tau = tau_t, s.qx (n entries via ct) = dtau_t
    R* dtau = s.ct;
    for (int i = tid; i < nx; i += DT::tpe) s.x[i] = 0;
    if (io.grad_x0) for (int i = tid; i < nx; i += DT::tpe) io.grad_x0[(size_t)b * nx + i] = exact ? -s.v[i] : (R)0;   // exact: -lambda_0 (Q3)
    DT::sync();
    for (int t = 0; t <= T; ++t) {
      if (t < T) {
        for (int i = tid; i < nu; i += DT::tpe) {
          R a = 0;
          for (int j = 0; j < nx; ++j) a += gK[(size_t)t * nu * nx + i * nx + j] * s.x[j];
          s.u[i] = a + gk[(size_t)t * nu + i];
        }
      }
      DT::sync();
      for (int i = tid; i < nx; i += DT::tpe) { s.tau[i] = X[t * nx + i] - xr[t * nx + i]; dtau[i] = s.x[i]; }
      if (t < T) for (int i = tid; i < nu; i += DT::tpe) { s.tau[nx + i] = U[t * nu + i] - ur[t * nu + i]; dtau[nx + i] = s.u[i]; }
      DT::sync();
      if (t < T) {
        if (io.grad_C)
          for (int idx = tid; idx < n * n; idx += DT::tpe) {
            const int i = idx / n, j = idx % n;
            const R g = (R)-0.5 * (dtau[i] * s.tau[j] + s.tau[i] * dtau[j]);
            if (P.C_batched) io.grad_C[((size_t)b * T + t) * n * n + idx] = g; else atomicAdd(&io.grad_C[(size_t)t * n * n + idx], g);
          }
        if (io.grad_c)
          for (int i = tid; i < n; i += DT::tpe) {
            if (P.C_batched) io.grad_c[((size_t)b * T + t) * n + i] = -dtau[i]; else atomicAdd(&io.grad_c[(size_t)t * n + i], -dtau[i]);
          }
        if (io.grad_uref)
          for (int i = tid; i < nu; i += DT::tpe) {
            R g = s.u[i];
            if (exact) {   // d loss / d ref_t = C_t dtau_t
              g = 0;
              for (int j = 0; j < n; ++j) g += Cp[(size_t)t * n * n + (size_t)(nx + i) * n + j] * dtau[j];
            }
            if (P.uref_batched) io.grad_uref[(size_t)b * szU + t * nu + i] = g; else atomicAdd(&io.grad_uref[t * nu + i], g);
          }
        if (io.dU) for (int i = tid; i < nu; i += DT::tpe) io.dU[(size_t)b * szU + t * nu + i] = s.u[i];
      } else {
        if (io.grad_Cf)
          for (int idx = tid; idx < n * n; idx += DT::tpe) {
            const int i = idx / n, j = idx % n;
            const bool in = (i < nx && j < nx);
            const R g = in ? (R)-0.5 * (dtau[i] * s.tau[j] + s.tau[i] * dtau[j]) : (R)0;
            if (P.Cf_batched) io.grad_Cf[(size_t)b * n * n + idx] = g; else if (in) atomicAdd(&io.grad_Cf[idx], g);
          }
        if (io.grad_cf)
          for (int i = tid; i < n; i += DT::tpe) {
            const R g = (i < nx) ? -dtau[i] : (R)0;
            if (P.Cf_batched) io.grad_cf[(size_t)b * n + i] = g; else if (i < nx) atomicAdd(&io.grad_cf[i], g);
          }
      }
      if (io.grad_xref)
        for (int i = tid; i < nx; i += DT::tpe) {
          R g = s.x[i];
          if (exact) {   // C_t dtau_t for t < T, C_final[:nx,:nx] dx_T at the end
            g = 0;
            if (t < T) for (int j = 0; j < n; ++j) g += Cp[(size_t)t * n * n + (size_t)i * n + j] * dtau[j];
            else for (int j = 0; j < nx; ++j) g += Cfe[i * n + j] * dtau[j];
          }
          if (P.xref_batched) io.grad_xref[(size_t)b * szX + t * nx + i] = g; else atomicAdd(&io.grad_xref[t * nx + i], g);
        }
      if (io.dX) for (int i = tid; i < nx; i += DT::tpe) io.dX[(size_t)b * szX + t * nx + i] = s.x[i];
      if (t < T) {
        DT::sync();
        // dx+ = A dx + B du; the jacobian load clobbers s.x/s.u for non-LTI embedded dynamics, so stash them
        for (int i = tid; i < nx; i += DT::tpe) s.xa[i] = s.x[i];
        for (int i = tid; i < nu; i += DT::tpe) s.ua[i] = s.u[i];
        DT::sync();
        load_AB(t);
        for (int i = tid; i < nx; i += DT::tpe) {
          R a = 0, a2 = 0;
          for (int j = 0; j < nx; ++j) a += s.A[i * nx + j] * s.xa[j];
          for (int j = 0; j < nu; ++j) a2 += s.Bm[i * nu + j] * s.ua[j];
          s.xan[i] = a + a2;
        }
        DT::sync();
        for (int i = tid; i < nx; i += DT::tpe) s.x[i] = s.xan[i];
      }
      DT::sync();
    }
  }
}

// ---------------------------------------------------------------------------
// stand-alone stages (one CTA per element)
template <typename R>
__global__ void __launch_bounds__(kGenThreads) stage_rollout(const DevProblem<R> P, const GenDims<R> d, const R* x0, const R* U, R* X) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  GenSmem<R> s = gen_carve<R>(smem_raw, d.nx(), d.nu());
  const int tid = threadIdx.x, nx = d.nx(), nu = d.nu(), T = d.T;
  if (d.dyn == TACD_DYN_LTI) {
    for (int i = tid; i < nx * nx; i += kGenThreads) s.A[i] = P.dyn_A[i];
    for (int i = tid; i < nx * nu; i += kGenThreads) s.Bm[i] = P.dyn_B[i];
  }
  __syncthreads();
  for (int b = blockIdx.x; b < P.B; b += gridDim.x) {
    R* Xb = X + (size_t)b * (T + 1) * nx;
    const R* Ub = U + (size_t)b * T * nu;
    for (int i = tid; i < nx; i += kGenThreads) Xb[i] = x0[(size_t)b * nx + i];
    __syncthreads();
    for (int t = 0; t < T; ++t) {
      gen_dyn_step<R>(P, d, s.A, s.Bm, 1, Xb + (size_t)t * nx, Ub + (size_t)t * nu, Xb + (size_t)(t + 1) * nx);
      __syncthreads();
    }
  }
}

template <typename R>
__global__ void __launch_bounds__(kGenThreads) stage_linearize(const DevProblem<R> P, const GenDims<R> d, const R* X, const R* U, R* A, R* Bm) {
  // one thread per (b, t)
  const int nx = d.nx(), nu = d.nu(), T = d.T;
  const size_t total = (size_t)P.B * T;
  for (size_t w = (size_t)blockIdx.x * blockDim.x + threadIdx.x; w < total; w += (size_t)gridDim.x * blockDim.x) {
    const size_t b = w / T, t = w % T;
    R* Ao = A + w * nx * nx;
    R* Bo = Bm + w * nx * nu;
    if (d.dyn == TACD_DYN_LTI) {
      for (int i = 0; i < nx * nx; ++i) Ao[i] = P.dyn_A[i];
      for (int i = 0; i < nx * nu; ++i) Bo[i] = P.dyn_B[i];
    } else {
      gen_dyn_jac_scalar<R>(P, d.dyn, X + (b * (T + 1) + t) * nx, U + w * nu, Ao, Bo);
    }
  }
}

template <typename R>
__global__ void __launch_bounds__(kGenThreads) stage_objective(const DevProblem<R> P, const GenDims<R> d, const R* X, const R* U, const R* C,
                                                                const R* c, const R* Cf, const R* cf, const R* xref, const R* uref, R* cost) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  GenSmem<R> s = gen_carve<R>(smem_raw, d.nx(), d.nu());
  const int nx = d.nx(), nu = d.nu(), n = d.n(), T = d.T;
  const size_t szX = (size_t)(T + 1) * nx, szU = (size_t)T * nu;
  for (int b = blockIdx.x; b < P.B; b += gridDim.x) {
    const R v = gen_objective<R>(d, s, X + b * szX, U + b * szU, C + (P.C_batched ? (size_t)b * T * n * n : 0),
                                 c + (P.C_batched ? (size_t)b * T * n : 0), Cf + (P.Cf_batched ? (size_t)b * n * n : 0),
                                 cf + (P.Cf_batched ? (size_t)b * n : 0), xref + (P.xref_batched ? b * szX : 0),
                                 uref + (P.uref_batched ? b * szU : 0));
    if (threadIdx.x == 0) cost[b] = v;
  }
}

template <typename R>
__global__ void __launch_bounds__(kGenThreads) stage_riccati(const DevProblem<R> P, const GenDims<R> d, const R* A, const R* Bm, const R* lx,
                                                              const R* lu, const R* lxx, const R* lxu, const R* luu, const R* lxN,
                                                              const R* lxxN, R* K, R* k, uint8_t* flags) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  GenSmem<R> s = gen_carve<R>(smem_raw, d.nx(), d.nu());
  const int tid = threadIdx.x, nx = d.nx(), nu = d.nu(), T = d.T;
  for (int b = blockIdx.x; b < P.B; b += gridDim.x) {
    bool nonpd = false;
    for (int i = tid; i < nx * nx; i += kGenThreads) s.V[i] = lxxN[(size_t)b * nx * nx + i];
    for (int i = tid; i < nx; i += kGenThreads) s.v[i] = lxN[(size_t)b * nx + i];
    __syncthreads();
    for (int t = T - 1; t >= 0; --t) {
      const size_t bt = (size_t)b * T + t;
      for (int i = tid; i < nx * nx; i += kGenThreads) { s.A[i] = A[bt * nx * nx + i]; s.Qxx[i] = lxx[bt * nx * nx + i]; }
      for (int i = tid; i < nx * nu; i += kGenThreads) { s.Bm[i] = Bm[bt * nx * nu + i]; s.Qxu[i] = lxu[bt * nx * nu + i]; }
      for (int i = tid; i < nu * nu; i += kGenThreads) s.Quu[i] = luu[bt * nu * nu + i];
      for (int i = tid; i < nx; i += kGenThreads) s.qx[i] = lx[bt * nx + i];
      for (int i = tid; i < nu; i += kGenThreads) s.qu[i] = lu[bt * nu + i];
      __syncthreads();
      nonpd = gen_value_step<R>(P, d, s, false, nullptr, nullptr) || nonpd;
      for (int i = tid; i < nu * nx; i += kGenThreads) K[bt * nu * nx + i] = s.Kt[i];
      for (int i = tid; i < nu; i += kGenThreads) k[bt * nu + i] = s.kt[i];
      __syncthreads();
    }
    if (flags && tid == 0) flags[b] = nonpd ? TACD_FLAG_NONPD : 0;
  }
}

template <typename R>
__global__ void __launch_bounds__(kGenThreads) stage_linesearch(const DevProblem<R> P, const GenDims<R> d, const R* x0, const R* X, const R* U,
                                                                 const R* K, const R* k, const R* C, const R* c, const R* Cf, const R* cf,
                                                                 const R* xref, const R* uref, R* Xc, R* Uc, R* costs) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  GenSmem<R> s = gen_carve<R>(smem_raw, d.nx(), d.nu());
  const int tid = threadIdx.x, nx = d.nx(), nu = d.nu(), n = d.n(), T = d.T, na = P.n_alpha;
  const size_t szX = (size_t)(T + 1) * nx, szU = (size_t)T * nu;
  if (d.dyn == TACD_DYN_LTI) {
    for (int i = tid; i < nx * nx; i += kGenThreads) s.A[i] = P.dyn_A[i];
    for (int i = tid; i < nx * nu; i += kGenThreads) s.Bm[i] = P.dyn_B[i];
  }
  __syncthreads();
  for (int b = blockIdx.x; b < P.B; b += gridDim.x) {
    const R* Xb = X + b * szX;
    const R* Ub = U + b * szU;
    const R* Kb = K + (size_t)b * T * nu * nx;
    const R* kb = k + b * szU;
    R* Xcb = Xc + (size_t)b * na * szX;
    R* Ucb = Uc + (size_t)b * na * szU;
    for (int idx = tid; idx < na * nx; idx += kGenThreads) Xcb[(idx / nx) * szX + (idx % nx)] = x0[(size_t)b * nx + idx % nx];
    __syncthreads();
    for (int t = 0; t < T; ++t) {
      for (int idx = tid; idx < na * nx; idx += kGenThreads) s.xa[idx] = Xcb[(idx / nx) * szX + (size_t)t * nx + idx % nx];
      __syncthreads();
      for (int idx = tid; idx < na * nu; idx += kGenThreads) {
        const int a = idx / nu, i = idx % nu;
        R acc = 0;
        for (int j = 0; j < nx; ++j) acc += Kb[(size_t)t * nu * nx + i * nx + j] * (s.xa[a * nx + j] - Xb[(size_t)t * nx + j]);
        R ui = Ub[(size_t)t * nu + i] + (P.alphas[a] * kb[(size_t)t * nu + i] + acc);
        if (P.has_umin) ui = tmax<R>(ui, P.umin[i]);
        if (P.has_umax) ui = tmin<R>(ui, P.umax[i]);
        s.ua[idx] = ui;
        Ucb[a * szU + (size_t)t * nu + i] = ui;
      }
      __syncthreads();
      gen_dyn_step<R>(P, d, s.A, s.Bm, na, s.xa, s.ua, s.xan);
      __syncthreads();
      for (int idx = tid; idx < na * nx; idx += kGenThreads) Xcb[(idx / nx) * szX + (size_t)(t + 1) * nx + idx % nx] = s.xan[idx];
      __syncthreads();
    }
    for (int a = 0; a < na; ++a) {
      const R v = gen_objective<R>(d, s, Xcb + a * szX, Ucb + a * szU, C + (P.C_batched ? (size_t)b * T * n * n : 0),
                                   c + (P.C_batched ? (size_t)b * T * n : 0), Cf + (P.Cf_batched ? (size_t)b * n * n : 0),
                                   cf + (P.Cf_batched ? (size_t)b * n : 0), xref + (P.xref_batched ? b * szX : 0),
                                   uref + (P.uref_batched ? b * szU : 0));
      if (tid == 0) costs[(size_t)b * na + a] = v;
    }
  }
}

template <typename R>
__global__ void stage_pnqp(int N, int m, const R* H, const R* q, const R* lo, const R* hi, R* x, int32_t* iters, R* scratch) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  R* w = scratch + (size_t)i * (m * m + 4 * m);
  int piv[TACD_MAX_NU];
  const int it = gen_pnqp<R>(m, H + (size_t)i * m * m, q + (size_t)i * m, lo + (size_t)i * m, hi + (size_t)i * m, x + (size_t)i * m, w, piv);
  if (iters) iters[i] = it;
}

// ---------------------------------------------------------------------------
// host launchers
// closed-loop advance (tacd_closed_loop_advance): one thread per element, run-time dimensions
template <typename R>
__global__ void __launch_bounds__(128) closed_loop_advance(const DevProblem<R> P, const GenDims<R> d, int k, int n_sim, R* x, const R* U_opt,
                                                           R* Xs, R* Us, R* U_warm, const R* xref_full, const R* uref_full, int ref_len,
                                                           R* xref_win, R* uref_win) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= P.B) return;
  const int nx = d.nx(), nu = d.nu(), T = d.T;
  R xc[TACD_MAX_NX], xn[TACD_MAX_NX], u[TACD_MAX_NU];
  for (int i = 0; i < nx; ++i) xc[i] = x[(size_t)b * nx + i];
  for (int i = 0; i < nu; ++i) u[i] = U_opt[(size_t)b * T * nu + i];
  if (d.dyn == TACD_DYN_LTI) {
    for (int i = 0; i < nx; ++i) {
      R s = 0, s2 = 0;
      for (int j = 0; j < nx; ++j) s += P.dyn_A[i * nx + j] * xc[j];
      for (int j = 0; j < nu; ++j) s2 += P.dyn_B[i * nu + j] * u[j];
      xn[i] = s + s2;
    }
  } else {
    gen_dyn_step_scalar<R>(P, d.dyn, xc, u, xn);
  }
  if (k == 0)
    for (int i = 0; i < nx; ++i) Xs[(size_t)b * (n_sim + 1) * nx + i] = xc[i];
  for (int i = 0; i < nx; ++i) {
    x[(size_t)b * nx + i] = xn[i];
    Xs[((size_t)b * (n_sim + 1) + k + 1) * nx + i] = xn[i];
  }
  for (int i = 0; i < nu; ++i) Us[((size_t)b * n_sim + k) * nu + i] = u[i];
  if (U_warm) {
    for (int t = 0; t + 1 < T; ++t)
      for (int i = 0; i < nu; ++i) U_warm[((size_t)b * T + t) * nu + i] = U_opt[((size_t)b * T + t + 1) * nu + i];
    for (int i = 0; i < nu; ++i) U_warm[((size_t)b * T + T - 1) * nu + i] = 0;
  }
  if (k + 1 < n_sim) {
    if (xref_win && (P.xref_batched || b == 0)) {
      const R* src = xref_full + ((size_t)(P.xref_batched ? b : 0) * ref_len + k + 1) * nx;
      R* dst = xref_win + (size_t)(P.xref_batched ? b : 0) * (T + 1) * nx;
      for (int i = 0; i < (T + 1) * nx; ++i) dst[i] = src[i];
    }
    if (uref_win && (P.uref_batched || b == 0)) {
      const R* src = uref_full + ((size_t)(P.uref_batched ? b : 0) * (ref_len - 1) + k + 1) * nu;
      R* dst = uref_win + (size_t)(P.uref_batched ? b : 0) * T * nu;
      for (int i = 0; i < T * nu; ++i) dst[i] = src[i];
    }
  }
}

// tacd_al_update: one warp per element, lanes over (t, i)
template <typename R>
__global__ void __launch_bounds__(128) al_update(int B, int T, int nx, const R* X, const R* xref, int xref_batched, const R* xmin, const R* xmax,
                                                 R* lam_lo, R* lam_hi, R rho, R rho_next, int update_lambda, R* dq, R* dp, R* viol) {
  const int lane = threadIdx.x & 31;
  const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (b >= B) return;
  const int W = (T + 1) * nx;
  const R* Xb = X + (size_t)b * W;
  const R* xr = xref + (xref_batched ? (size_t)b * W : 0);
  R vmax = 0;
  for (int w = lane; w < W; w += 32) {
    const int i = w % nx;
    const R x = Xb[w], r = xr[w];
    const R ghi = x - xmax[i], glo = xmin[i] - x;          // -inf on an unbounded side
    R lhi = lam_hi[(size_t)b * W + w], llo = lam_lo[(size_t)b * W + w];
    if (update_lambda) {
      lhi = fmax(lhi + rho * ghi, (R)0);                    // (-inf -> 0)
      llo = fmax(llo + rho * glo, (R)0);
      lam_hi[(size_t)b * W + w] = lhi;
      lam_lo[(size_t)b * W + w] = llo;
    }
    // a constraint stays in the model while its multiplier is positive: predicting the active set with
    // lam + rho_next g drops constraints that sit rho-tightly on their bound (g = (mu - lam)/rho < 0 when the
    // multiplier is decreasing) as soon as rho grows, and the outer loop oscillates
    const bool ahi = update_lambda ? lhi > 0 : lhi + rho_next * ghi > 0;
    const bool alo = update_lambda ? llo > 0 : llo + rho_next * glo > 0;
    dq[(size_t)b * W + w] = rho_next * (R)((ahi ? 1 : 0) + (alo ? 1 : 0));
    dp[(size_t)b * W + w] = (ahi ? lhi + rho_next * (r - xmax[i]) : (R)0) - (alo ? llo + rho_next * (xmin[i] - r) : (R)0);
    vmax = fmax(vmax, fmax(ghi, glo));
  }
  for (int o = 16; o > 0; o >>= 1) vmax = fmax(vmax, __shfl_xor_sync(0xffffffffu, vmax, o));
  if (lane == 0 && viol) viol[b] = vmax;
}

template <typename K>
static int prep_smem(K kernel, size_t bytes) {
  const DeviceInfo di = device_info();
  if (bytes > (size_t)di.smem_optin) { set_error("generic kernel needs %zu B shared memory (> %d)", bytes, di.smem_optin); return TACD_EUNSUPPORTED; }
  return check_cuda(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes), "smem attr");
}

template <typename R>
static FwdPtrs<R> fwd_ptrs(const tacd_forward_io* io) {
  FwdPtrs<R> f;
  f.x0 = (const R*)io->x0; f.C = (const R*)io->C; f.c = (const R*)io->c; f.Cf = (const R*)io->Cf; f.cf = (const R*)io->cf;
  f.C_ls = (const R*)io->C_ls; f.c_ls = (const R*)io->c_ls; f.Cf_ls = (const R*)io->Cf_ls; f.cf_ls = (const R*)io->cf_ls;
  f.xref = (const R*)io->xref; f.uref = (const R*)io->uref; f.Uinit = (const R*)io->Uinit;
  f.replay_iters = io->replay_iters; f.replay_alpha = io->replay_alpha; f.replay_flags = io->replay_flags;
  f.X = (R*)io->X; f.U = (R*)io->U; f.A = (R*)io->A; f.Bm = (R*)io->Bm;
  f.tight = io->tight; f.iters = io->iters; f.alpha_trace = io->alpha_trace; f.flags = io->flags;
  f.cost = (R*)io->cost; f.cost_trace = (R*)io->cost_trace; f.counter = nullptr;
  return f;
}

// Sizes of the LTI sweep get kernels with compile-time nx, nu; everything else runs the
// run-time-dimension instance (NXT = NUT = 0).  TACD_GENERIC_RUNTIME_DIMS=1 forces the latter (A/B measurements).
#define TACD_GEN_SIZES(X) X(8, 4) X(16, 4) X(16, 8) X(32, 8)
static bool gen_force_ws() { TACD_KNOB(k, "TACD_GENERIC_WS"); return env_set(k); }
static bool gen_runtime_dims() { TACD_KNOB(k, "TACD_GENERIC_RUNTIME_DIMS"); return env_set(k); }

// one warp per element only if four element slots (matrices alone) fit the CTA's shared memory
template <typename R>
static bool gen_warp_mode(const tacd_problem* p) {
  return gen_tpe(p) == 32 && gen_elem_bytes<R>(p->nx, p->nu, 0) * (kGenThreads / 32) <= (size_t)device_info().smem_optin;
}

// Shared memory of one launch: per element slot the matrices and, when enough slots per SM still fit, the element's
// X, U, K, k (global-memory round trips inside the per-step loops were the bulk of the time otherwise).
struct GenLaunch { int grid; size_t smem; bool in_smem; };
template <typename R, int TPE>
static GenLaunch gen_launch_shape(const tacd_problem* p, size_t slice) {
  constexpr int per_cta = kGenThreads / TPE;
  const size_t with_slice = gen_elem_bytes<R>(p->nx, p->nu, slice) * per_cta;
  GenLaunch L;
  // CTA-per-element: four CTAs per SM; warp-per-element: two four-element CTAs per SM are enough to cover the barriers
  L.in_smem = with_slice <= (size_t)device_info().smem_optin / (TPE == 32 ? 2 : gen_ctas_per_sm()) && !gen_force_ws();
  L.smem = L.in_smem ? with_slice : gen_elem_bytes<R>(p->nx, p->nu, 0) * per_cta;
  L.grid = (gen_slots(p, TPE) + per_cta - 1) / per_cta;
  return L;
}

template <typename R, int NXT, int NUT, int TPE>
static int run_forward_generic(const tacd_problem* p, const tacd_forward_io* io, void* ws, size_t ws_bytes, cudaStream_t s) {
  const size_t slice = gen_slice_scalars(p);
  const GenLaunch L = gen_launch_shape<R, TPE>(p, slice);
  const size_t need = (size_t)L.grid * (kGenThreads / TPE) * slice * sizeof(R);
  if (!L.in_smem && ws_bytes < need) { set_error("forward workspace too small: %zu < %zu", ws_bytes, need); return TACD_EWORKSPACE; }
  auto kernel = ilqr_forward_generic<R, NXT, NUT, TPE>;
  if (int rc = prep_smem(kernel, L.smem)) return rc;
  kernel<<<L.grid, kGenThreads, L.smem, s>>>(to_dev<R>(p), dims_of<R, NXT, NUT, TPE>(p), fwd_ptrs<R>(io), L.in_smem ? nullptr : (R*)ws, slice);
  note_launches(1);
  return check_cuda(cudaPeekAtLastError(), "ilqr_forward_generic launch");
}

template <typename R>
int launch_forward_generic(const tacd_problem* p, const tacd_forward_io* io, void* ws, size_t ws_bytes, cudaStream_t s) {
  const bool warp = gen_warp_mode<R>(p);
  if (!gen_runtime_dims()) {
#define X(NX_, NU_) if (p->nx == NX_ && p->nu == NU_) \
    return warp ? run_forward_generic<R, NX_, NU_, 32>(p, io, ws, ws_bytes, s) : run_forward_generic<R, NX_, NU_, kGenThreads>(p, io, ws, ws_bytes, s);
    TACD_GEN_SIZES(X)
#undef X
  }
  return warp ? run_forward_generic<R, 0, 0, 32>(p, io, ws, ws_bytes, s) : run_forward_generic<R, 0, 0, kGenThreads>(p, io, ws, ws_bytes, s);
}

template <typename R, int NXT, int NUT, int TPE>
static int run_backward_generic(const tacd_problem* p, const tacd_backward_io* io, void* ws, size_t ws_bytes, cudaStream_t s) {
  const size_t slice = gen_slice_scalars(p);
  const GenLaunch L = gen_launch_shape<R, TPE>(p, slice);
  const size_t need = (size_t)L.grid * (kGenThreads / TPE) * slice * sizeof(R);
  if (!L.in_smem && ws_bytes < need) { set_error("backward workspace too small: %zu < %zu", ws_bytes, need); return TACD_EWORKSPACE; }
  auto kernel = ilqr_backward_generic<R, NXT, NUT, TPE>;
  if (int rc = prep_smem(kernel, L.smem)) return rc;
  const size_t T = p->T, nx = p->nx, nu = p->nu, n = nx + nu, szX = (T + 1) * nx, szU = T * nu;
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
  kernel<<<L.grid, kGenThreads, L.smem, s>>>(to_dev<R>(p), dims_of<R, NXT, NUT, TPE>(p), f, L.in_smem ? nullptr : (R*)ws, slice);
  note_launches(1);
  return check_cuda(cudaPeekAtLastError(), "ilqr_backward_generic launch");
}

template <typename R>
int launch_backward_generic(const tacd_problem* p, const tacd_backward_io* io, void* ws, size_t ws_bytes, cudaStream_t s) {
  const bool warp = gen_warp_mode<R>(p);
  if (!gen_runtime_dims()) {
#define X(NX_, NU_) if (p->nx == NX_ && p->nu == NU_) \
    return warp ? run_backward_generic<R, NX_, NU_, 32>(p, io, ws, ws_bytes, s) : run_backward_generic<R, NX_, NU_, kGenThreads>(p, io, ws, ws_bytes, s);
    TACD_GEN_SIZES(X)
#undef X
  }
  return warp ? run_backward_generic<R, 0, 0, 32>(p, io, ws, ws_bytes, s) : run_backward_generic<R, 0, 0, kGenThreads>(p, io, ws, ws_bytes, s);
}

#if TACD_GEN_PART == 0
template int launch_forward_generic<float>(const tacd_problem*, const tacd_forward_io*, void*, size_t, cudaStream_t);
#elif TACD_GEN_PART == 1
template int launch_backward_generic<float>(const tacd_problem*, const tacd_backward_io*, void*, size_t, cudaStream_t);
#elif TACD_GEN_PART == 2
template int launch_forward_generic<double>(const tacd_problem*, const tacd_forward_io*, void*, size_t, cudaStream_t);
#else
template int launch_backward_generic<double>(const tacd_problem*, const tacd_backward_io*, void*, size_t, cudaStream_t);
#endif

#if TACD_GEN_PART == 0   // stand-alone stages and the C entry points: once

template <typename R>
static int stage_common(const tacd_problem* p) {
  if (!p) { set_error("null problem"); return TACD_EINVAL; }
  if (p->nx < 1 || p->nu < 1 || p->nx > TACD_MAX_NX || p->nu > TACD_MAX_NU || p->T < 1 || p->B < 0) { set_error("bad sizes"); return TACD_EINVAL; }
  return TACD_OK;
}

template <typename R>
static int run_rollout(const tacd_problem* p, const void* x0, const void* U, void* X, cudaStream_t s) {
  auto kernel = stage_rollout<R>;
  const size_t smem = gen_smem_bytes<R>(p->nx, p->nu);
  if (int rc = prep_smem(kernel, smem)) return rc;
  kernel<<<gen_grid(p), kGenThreads, smem, s>>>(to_dev<R>(p), dims_of<R>(p), (const R*)x0, (const R*)U, (R*)X);
  note_launches(1);
  return check_cuda(cudaPeekAtLastError(), "stage_rollout launch");
}
template <typename R>
static int run_linearize(const tacd_problem* p, const void* X, const void* U, void* A, void* Bm, cudaStream_t s) {
  const size_t total = (size_t)p->B * p->T;
  int grid = (int)((total + kGenThreads - 1) / kGenThreads);
  if (grid > 148 * 16) grid = 148 * 16;
  stage_linearize<R><<<grid, kGenThreads, 0, s>>>(to_dev<R>(p), dims_of<R>(p), (const R*)X, (const R*)U, (R*)A, (R*)Bm);
  note_launches(1);
  return check_cuda(cudaPeekAtLastError(), "stage_linearize launch");
}
template <typename R>
static int run_objective(const tacd_problem* p, const void* X, const void* U, const void* C, const void* c, const void* Cf, const void* cf,
                         const void* xref, const void* uref, void* cost, cudaStream_t s) {
  auto kernel = stage_objective<R>;
  const size_t smem = gen_smem_bytes<R>(p->nx, p->nu);
  if (int rc = prep_smem(kernel, smem)) return rc;
  kernel<<<gen_grid(p), kGenThreads, smem, s>>>(to_dev<R>(p), dims_of<R>(p), (const R*)X, (const R*)U, (const R*)C, (const R*)c, (const R*)Cf,
                                                 (const R*)cf, (const R*)xref, (const R*)uref, (R*)cost);
  note_launches(1);
  return check_cuda(cudaPeekAtLastError(), "stage_objective launch");
}
template <typename R>
static int run_riccati(const tacd_problem* p, const void* A, const void* Bm, const void* lx, const void* lu, const void* lxx, const void* lxu,
                       const void* luu, const void* lxN, const void* lxxN, void* K, void* k, uint8_t* flags, cudaStream_t s) {
  auto kernel = stage_riccati<R>;
  const size_t smem = gen_smem_bytes<R>(p->nx, p->nu);
  if (int rc = prep_smem(kernel, smem)) return rc;
  kernel<<<gen_grid(p), kGenThreads, smem, s>>>(to_dev<R>(p), dims_of<R>(p), (const R*)A, (const R*)Bm, (const R*)lx, (const R*)lu, (const R*)lxx,
                                                 (const R*)lxu, (const R*)luu, (const R*)lxN, (const R*)lxxN, (R*)K, (R*)k, flags);
  note_launches(1);
  return check_cuda(cudaPeekAtLastError(), "stage_riccati launch");
}
template <typename R>
static int run_linesearch(const tacd_problem* p, const void* x0, const void* X, const void* U, const void* K, const void* k, const void* C,
                          const void* c, const void* Cf, const void* cf, const void* xref, const void* uref, void* Xc, void* Uc, void* costs,
                          cudaStream_t s) {
  auto kernel = stage_linesearch<R>;
  const size_t smem = gen_smem_bytes<R>(p->nx, p->nu);
  if (int rc = prep_smem(kernel, smem)) return rc;
  kernel<<<gen_grid(p), kGenThreads, smem, s>>>(to_dev<R>(p), dims_of<R>(p), (const R*)x0, (const R*)X, (const R*)U, (const R*)K, (const R*)k,
                                                 (const R*)C, (const R*)c, (const R*)Cf, (const R*)cf, (const R*)xref, (const R*)uref, (R*)Xc,
                                                 (R*)Uc, (R*)costs);
  note_launches(1);
  return check_cuda(cudaPeekAtLastError(), "stage_linesearch launch");
}

#endif  // TACD_GEN_PART == 0
}  // namespace tacd

#if TACD_GEN_PART == 0
using namespace tacd;

#define TACD_DISPATCH(p, fn, ...) ((p)->dtype == TACD_F32 ? fn<float>(__VA_ARGS__) : fn<double>(__VA_ARGS__))

extern "C" {

int tacd_closed_loop_advance(const tacd_problem* p, int32_t k, int32_t n_sim, void* x, const void* U_opt, void* Xs, void* Us,
                             void* U_warm, const void* xref_full, const void* uref_full, int32_t ref_len, void* xref_win,
                             void* uref_win, void* stream) {
  if (!p || p->nx < 1 || p->nu < 1 || p->nx > TACD_MAX_NX || p->nu > TACD_MAX_NU || p->T < 1 || p->B < 0) { set_error("bad sizes"); return TACD_EINVAL; }
  if (p->dyn_id == TACD_DYN_EXTERNAL) { set_error("closed-loop advance needs embedded dynamics"); return TACD_EINVAL; }
  if (k < 0 || k >= n_sim || !x || !U_opt || !Xs || !Us) { set_error("tacd_closed_loop_advance: bad step index or null pointer"); return TACD_EINVAL; }
  if ((xref_win && !xref_full) || (uref_win && !uref_full) || ((xref_win || uref_win) && ref_len < n_sim + p->T + 1)) {
    set_error("reference trajectories must cover n_sim + T + 1 steps"); return TACD_EINVAL;
  }
  if (p->B == 0) return TACD_OK;
  cudaStream_t s = (cudaStream_t)stream;
  const int grid = (p->B + 127) / 128;
  if (p->dtype == TACD_F32)
    closed_loop_advance<float><<<grid, 128, 0, s>>>(to_dev<float>(p), dims_of<float>(p), k, n_sim, (float*)x, (const float*)U_opt, (float*)Xs,
                                                    (float*)Us, (float*)U_warm, (const float*)xref_full, (const float*)uref_full, ref_len,
                                                    (float*)xref_win, (float*)uref_win);
  else
    closed_loop_advance<double><<<grid, 128, 0, s>>>(to_dev<double>(p), dims_of<double>(p), k, n_sim, (double*)x, (const double*)U_opt,
                                                     (double*)Xs, (double*)Us, (double*)U_warm, (const double*)xref_full,
                                                     (const double*)uref_full, ref_len, (double*)xref_win, (double*)uref_win);
  note_launches(1);
  return check_cuda(cudaPeekAtLastError(), "closed_loop_advance launch");
}

int tacd_al_update(int32_t dtype, int32_t B, int32_t T, int32_t nx, const void* X, const void* xref, int32_t xref_batched, const void* x_min,
                   const void* x_max, void* lam_lo, void* lam_hi, double rho, double rho_next, int32_t update_lambda, void* dq, void* dp,
                   void* viol, void* stream) {
  if (B < 0 || T < 1 || nx < 1 || nx > TACD_MAX_NX || (dtype != TACD_F32 && dtype != TACD_F64)) { set_error("tacd_al_update: bad sizes / dtype"); return TACD_EINVAL; }
  if (!X || !xref || !x_min || !x_max || !lam_lo || !lam_hi || !dq || !dp) { set_error("tacd_al_update: required pointer is null"); return TACD_EINVAL; }
  if (!(rho > 0) || !(rho_next > 0)) { set_error("tacd_al_update: rho must be positive"); return TACD_EINVAL; }
  if (B == 0) return TACD_OK;
  cudaStream_t s = (cudaStream_t)stream;
  const int grid = (B + 3) / 4;
  if (dtype == TACD_F32)
    al_update<float><<<grid, 128, 0, s>>>(B, T, nx, (const float*)X, (const float*)xref, xref_batched, (const float*)x_min, (const float*)x_max,
                                          (float*)lam_lo, (float*)lam_hi, (float)rho, (float)rho_next, update_lambda, (float*)dq, (float*)dp, (float*)viol);
  else
    al_update<double><<<grid, 128, 0, s>>>(B, T, nx, (const double*)X, (const double*)xref, xref_batched, (const double*)x_min, (const double*)x_max,
                                           (double*)lam_lo, (double*)lam_hi, rho, rho_next, update_lambda, (double*)dq, (double*)dp, (double*)viol);
  note_launches(1);
  return check_cuda(cudaPeekAtLastError(), "al_update launch");
}

int tacd_rollout(const tacd_problem* p, const void* x0, const void* U, void* X, void* stream) {
  if (int rc = stage_common<float>(p)) return rc;
  if (p->dyn_id == TACD_DYN_EXTERNAL) { set_error("rollout needs embedded dynamics"); return TACD_EINVAL; }
  if (p->B == 0) return TACD_OK;
  return TACD_DISPATCH(p, run_rollout, p, x0, U, X, (cudaStream_t)stream);
}
int tacd_linearize(const tacd_problem* p, const void* X, const void* U, void* A, void* Bm, void* stream) {
  if (int rc = stage_common<float>(p)) return rc;
  if (p->dyn_id == TACD_DYN_EXTERNAL) { set_error("linearize needs embedded dynamics"); return TACD_EINVAL; }
  if (p->B == 0) return TACD_OK;
  return TACD_DISPATCH(p, run_linearize, p, X, U, A, Bm, (cudaStream_t)stream);
}
int tacd_objective(const tacd_problem* p, const void* X, const void* U, const void* C, const void* c, const void* Cf, const void* cf,
                   const void* xref, const void* uref, void* cost, void* stream) {
  if (int rc = stage_common<float>(p)) return rc;
  if (p->B == 0) return TACD_OK;
  return TACD_DISPATCH(p, run_objective, p, X, U, C, c, Cf, cf, xref, uref, cost, (cudaStream_t)stream);
}
int tacd_riccati(const tacd_problem* p, const void* A, const void* Bm, const void* lx, const void* lu, const void* lxx, const void* lxu,
                 const void* luu, const void* lxN, const void* lxxN, void* K, void* k, uint8_t* flags, void* stream) {
  if (int rc = stage_common<float>(p)) return rc;
  if (p->B == 0) return TACD_OK;
  return TACD_DISPATCH(p, run_riccati, p, A, Bm, lx, lu, lxx, lxu, luu, lxN, lxxN, K, k, flags, (cudaStream_t)stream);
}
int tacd_linesearch(const tacd_problem* p, const void* x0, const void* X, const void* U, const void* K, const void* k, const void* C,
                    const void* c, const void* Cf, const void* cf, const void* xref, const void* uref, void* Xc, void* Uc, void* costs,
                    void* stream) {
  if (int rc = stage_common<float>(p)) return rc;
  if (p->dyn_id == TACD_DYN_EXTERNAL) { set_error("linesearch needs embedded dynamics"); return TACD_EINVAL; }
  if (p->B == 0) return TACD_OK;
  return TACD_DISPATCH(p, run_linesearch, p, x0, X, U, K, k, C, c, Cf, cf, xref, uref, Xc, Uc, costs, (cudaStream_t)stream);
}
int tacd_pnqp(int32_t dtype, int32_t N, int32_t m, const void* H, const void* q, const void* lo, const void* hi, void* x, int32_t* iters,
              void* stream) {
  if (m < 1 || m > TACD_MAX_NU || N < 0) { set_error("pnqp: 1 <= m <= %d", TACD_MAX_NU); return TACD_EINVAL; }
  if (N == 0) return TACD_OK;
  cudaStream_t s = (cudaStream_t)stream;
  const size_t es = dtype == TACD_F32 ? 4 : 8;
  void* scratch = nullptr;
  if (int rc = check_cuda(cudaMallocAsync(&scratch, (size_t)N * (m * m + 4 * m) * es, s), "pnqp scratch")) return rc;
  const int grid = (N + 127) / 128;
  if (dtype == TACD_F32) stage_pnqp<float><<<grid, 128, 0, s>>>(N, m, (const float*)H, (const float*)q, (const float*)lo, (const float*)hi, (float*)x, iters, (float*)scratch);
  else stage_pnqp<double><<<grid, 128, 0, s>>>(N, m, (const double*)H, (const double*)q, (const double*)lo, (const double*)hi, (double*)x, iters, (double*)scratch);
  note_launches(1);
  int rc = check_cuda(cudaPeekAtLastError(), "stage_pnqp launch");
  cudaFreeAsync(scratch, s);
  return rc;
}

}  // extern "C"
#endif  // TACD_GEN_PART == 0
