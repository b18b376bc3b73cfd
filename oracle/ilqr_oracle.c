/*
 * ilqr_oracle.c — CPU oracle for the batched differentiable iLQR solve.
 *
 * TEST INFRASTRUCTURE ONLY (see ilqr_oracle_impl.h for the full header and the
 * parity status).  Exports the same entry points as include/tacd_ilqr.h with
 * the prefix tacd_oracle_ and HOST pointers.  Built by oracle/Makefile into
 * oracle/libtacd_oracle.so; only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load it.
 *
 * The batch loop is OpenMP-parallel (elements are independent, quirk Q11), so
 * the CPU baseline can use every host core; TACD_ORACLE_THREADS or
 * omp_set_num_threads controls the count.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "../include/tacd_ilqr.h"

#define CAT_(a, b) a##b
#define CAT(a, b) CAT_(a, b)

#define REAL float
#define SUF(x) CAT(x, _f32)
#include "ilqr_oracle_impl.h"
#undef REAL
#undef SUF

#define REAL double
#define SUF(x) CAT(x, _f64)
#include "ilqr_oracle_impl.h"
#undef REAL
#undef SUF

static int check_problem(const tacd_problem* p) {
  if (!p) return TACD_EINVAL;
  if (p->B < 0 || p->T < 1 || p->nx < 1 || p->nu < 1) return TACD_EINVAL;
  if (p->nx > TACD_MAX_NX || p->nu > TACD_MAX_NU) return TACD_EINVAL;
  if (p->n_alpha < 1 || p->n_alpha > TACD_MAX_ALPHA) return TACD_EINVAL;
  if (p->dtype != TACD_F32 && p->dtype != TACD_F64) return TACD_EINVAL;
  return TACD_OK;
}

int tacd_oracle_abi_version(void) { return TACD_ABI_VERSION; }

int tacd_oracle_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
void tacd_oracle_set_num_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

int tacd_oracle_ilqr_forward(const tacd_problem* p, const tacd_forward_io* io) {
  int rc = check_problem(p);
  if (rc) return rc;
  const size_t wl = (p->dtype == TACD_F32) ? solve_ws_len_f32(p) : solve_ws_len_f64(p);
#pragma omp parallel
  {
    void* ws = malloc((p->dtype == TACD_F32 ? sizeof(float) : sizeof(double)) * wl);
#pragma omp for schedule(dynamic, 16)
    for (int b = 0; b < p->B; ++b) {
      if (p->dtype == TACD_F32) solve1_f32(p, io, b, (float*)ws);
      else solve1_f64(p, io, b, (double*)ws);
    }
    free(ws);
  }
  return TACD_OK;
}

/* Backward.  Batched outputs are written per element; shared outputs are summed
 * over the batch in element order (single thread) so results are reproducible.
 * When every requested output is batched the loop is split across threads by
 * giving each thread a contiguous sub-batch through a shifted problem view. */
int tacd_oracle_ilqr_backward(const tacd_problem* p, const tacd_backward_io* io) {
  int rc = check_problem(p);
  if (rc) return rc;
  const int shared_out = (io->grad_C && !p->C_batched) || (io->grad_c && !p->C_batched) ||
                         (io->grad_Cf && !p->Cf_batched) || (io->grad_cf && !p->Cf_batched) ||
                         (io->grad_xref && !p->xref_batched) || (io->grad_uref && !p->uref_batched);
  if (shared_out || p->B < 64) {
    if (p->dtype == TACD_F32) backward_all_f32(p, io); else backward_all_f64(p, io);
    return TACD_OK;
  }
  const size_t es = (p->dtype == TACD_F32) ? sizeof(float) : sizeof(double);
  const size_t nx = p->nx, nu = p->nu, T = p->T, n = nx + nu;
  const size_t szX = (T + 1) * nx, szU = T * nu;
  const int chunk = 64;
  const int nchunks = (p->B + chunk - 1) / chunk;
#pragma omp parallel for schedule(dynamic, 1)
  for (int ci = 0; ci < nchunks; ++ci) {
    const size_t b0 = (size_t)ci * chunk;
    tacd_problem q = *p;
    q.B = (int)((b0 + chunk <= (size_t)p->B) ? chunk : (p->B - b0));
    tacd_backward_io s = *io;
#define SHIFT(f, per) if (s.f) s.f = (void*)((char*)s.f + b0 * (per) * es)
#define SHIFTC(f, per) if (s.f) s.f = (const void*)((const char*)s.f + b0 * (per) * es)
    SHIFTC(X, szX); SHIFTC(U, szU);
    if (p->C_batched) SHIFTC(C, T * n * n);
    if (p->xref_batched) SHIFTC(xref, szX);
    if (p->uref_batched) SHIFTC(uref, szU);
    s.tight = io->tight + b0 * szU;
    SHIFTC(A, T * nx * nx); SHIFTC(Bm, T * nx * nu);
    SHIFTC(gX, szX); SHIFTC(gU, szU);
    SHIFT(grad_C, T * n * n); SHIFT(grad_c, T * n); SHIFT(grad_Cf, n * n); SHIFT(grad_cf, n);
    SHIFT(grad_x0, nx); SHIFT(grad_xref, szX); SHIFT(grad_uref, szU);
    SHIFT(dX, szX); SHIFT(dU, szU);
#undef SHIFT
#undef SHIFTC
    if (p->dtype == TACD_F32) backward_all_f32(&q, &s); else backward_all_f64(&q, &s);
  }
  return TACD_OK;
}

int tacd_oracle_rollout(const tacd_problem* p, const void* x0, const void* U, void* X) {
  int rc = check_problem(p);
  if (rc) return rc;
  const size_t szX = (size_t)(p->T + 1) * p->nx, szU = (size_t)p->T * p->nu;
#pragma omp parallel for
  for (int b = 0; b < p->B; ++b) {
    if (p->dtype == TACD_F32) rollout1_f32(p, (const float*)x0 + (size_t)b * p->nx, (const float*)U + b * szU, (float*)X + b * szX);
    else rollout1_f64(p, (const double*)x0 + (size_t)b * p->nx, (const double*)U + b * szU, (double*)X + b * szX);
  }
  return TACD_OK;
}

int tacd_oracle_linearize(const tacd_problem* p, const void* X, const void* U, void* A, void* Bm) {
  int rc = check_problem(p);
  if (rc) return rc;
  const size_t nx = p->nx, nu = p->nu, T = p->T;
#pragma omp parallel for
  for (int b = 0; b < p->B; ++b)
    for (size_t t = 0; t < T; ++t) {
      const size_t ix = ((size_t)b * (T + 1) + t) * nx, iu = ((size_t)b * T + t) * nu;
      const size_t ia = ((size_t)b * T + t) * nx * nx, ib = ((size_t)b * T + t) * nx * nu;
      if (p->dtype == TACD_F32) dyn_jac_f32(p, (const float*)X + ix, (const float*)U + iu, (float*)A + ia, (float*)Bm + ib);
      else dyn_jac_f64(p, (const double*)X + ix, (const double*)U + iu, (double*)A + ia, (double*)Bm + ib);
    }
  return TACD_OK;
}

int tacd_oracle_objective(const tacd_problem* p, const void* X, const void* U,
                          const void* C, const void* c, const void* Cf, const void* cf,
                          const void* xref, const void* uref, void* cost) {
  int rc = check_problem(p);
  if (rc) return rc;
  const size_t nx = p->nx, nu = p->nu, T = p->T, n = nx + nu;
  const size_t szX = (T + 1) * nx, szU = T * nu;
#pragma omp parallel for
  for (int b = 0; b < p->B; ++b) {
    const size_t oC = p->C_batched ? b * T * n * n : 0, oc = p->C_batched ? b * T * n : 0;
    const size_t oCf = p->Cf_batched ? b * n * n : 0, ocf = p->Cf_batched ? b * n : 0;
    const size_t ox = p->xref_batched ? b * szX : 0, ou = p->uref_batched ? b * szU : 0;
    if (p->dtype == TACD_F32)
      ((float*)cost)[b] = objective1_f32(p, (const float*)X + b * szX, (const float*)U + b * szU, (const float*)C + oC,
                                         (const float*)c + oc, (const float*)Cf + oCf, (const float*)cf + ocf,
                                         (const float*)xref + ox, (const float*)uref + ou);
    else
      ((double*)cost)[b] = objective1_f64(p, (const double*)X + b * szX, (const double*)U + b * szU, (const double*)C + oC,
                                          (const double*)c + oc, (const double*)Cf + oCf, (const double*)cf + ocf,
                                          (const double*)xref + ox, (const double*)uref + ou);
  }
  return TACD_OK;
}

int tacd_oracle_riccati(const tacd_problem* p, const void* A, const void* Bm,
                        const void* lx, const void* lu, const void* lxx, const void* lxu, const void* luu,
                        const void* lxN, const void* lxxN, void* K, void* k, uint8_t* flags) {
  int rc = check_problem(p);
  if (rc) return rc;
  const size_t nx = p->nx, nu = p->nu, T = p->T;
#pragma omp parallel for
  for (int b = 0; b < p->B; ++b) {
    int f;
#define AT(ptr, type, per) ((const type*)(ptr) + (size_t)b * (per))
    if (p->dtype == TACD_F32)
      f = riccati_blocks_f32(p, AT(A, float, T * nx * nx), AT(Bm, float, T * nx * nu), AT(lx, float, T * nx),
                             AT(lu, float, T * nu), AT(lxx, float, T * nx * nx), AT(lxu, float, T * nx * nu),
                             AT(luu, float, T * nu * nu), AT(lxN, float, nx), AT(lxxN, float, nx * nx),
                             (float*)K + (size_t)b * T * nu * nx, (float*)k + (size_t)b * T * nu);
    else
      f = riccati_blocks_f64(p, AT(A, double, T * nx * nx), AT(Bm, double, T * nx * nu), AT(lx, double, T * nx),
                             AT(lu, double, T * nu), AT(lxx, double, T * nx * nx), AT(lxu, double, T * nx * nu),
                             AT(luu, double, T * nu * nu), AT(lxN, double, nx), AT(lxxN, double, nx * nx),
                             (double*)K + (size_t)b * T * nu * nx, (double*)k + (size_t)b * T * nu);
#undef AT
    if (flags) flags[b] = f ? TACD_FLAG_NONPD : 0;
  }
  return TACD_OK;
}

int tacd_oracle_linesearch(const tacd_problem* p, const void* x0, const void* X, const void* U,
                           const void* K, const void* k,
                           const void* C, const void* c, const void* Cf, const void* cf,
                           const void* xref, const void* uref, void* Xc, void* Uc, void* costs) {
  int rc = check_problem(p);
  if (rc) return rc;
  const size_t nx = p->nx, nu = p->nu, T = p->T, n = nx + nu, na = p->n_alpha;
  const size_t szX = (T + 1) * nx, szU = T * nu;
#pragma omp parallel for
  for (int b = 0; b < p->B; ++b) {
    const size_t oC = p->C_batched ? b * T * n * n : 0, oc = p->C_batched ? b * T * n : 0;
    const size_t oCf = p->Cf_batched ? b * n * n : 0, ocf = p->Cf_batched ? b * n : 0;
    const size_t ox = p->xref_batched ? b * szX : 0, ou = p->uref_batched ? b * szU : 0;
    for (size_t a = 0; a < na; ++a) {
      if (p->dtype == TACD_F32) {
        float* xc = (float*)Xc + (b * na + a) * szX;
        float* uc = (float*)Uc + (b * na + a) * szU;
        candidate_f32(p, (float)p->alphas[a], (const float*)x0 + b * nx, (const float*)X + b * szX, (const float*)U + b * szU,
                      (const float*)K + b * T * nu * nx, (const float*)k + b * szU, xc, uc);
        ((float*)costs)[b * na + a] = objective1_f32(p, xc, uc, (const float*)C + oC, (const float*)c + oc, (const float*)Cf + oCf,
                                                     (const float*)cf + ocf, (const float*)xref + ox, (const float*)uref + ou);
      } else {
        double* xc = (double*)Xc + (b * na + a) * szX;
        double* uc = (double*)Uc + (b * na + a) * szU;
        candidate_f64(p, (double)p->alphas[a], (const double*)x0 + b * nx, (const double*)X + b * szX, (const double*)U + b * szU,
                      (const double*)K + b * T * nu * nx, (const double*)k + b * szU, xc, uc);
        ((double*)costs)[b * na + a] = objective1_f64(p, xc, uc, (const double*)C + oC, (const double*)c + oc, (const double*)Cf + oCf,
                                                      (const double*)cf + ocf, (const double*)xref + ox, (const double*)uref + ou);
      }
    }
  }
  return TACD_OK;
}

int tacd_oracle_pnqp(int32_t dtype, int32_t N, int32_t m, const void* H, const void* q,
                     const void* lo, const void* hi, void* x, int32_t* iters) {
  if (m < 1 || m > TACD_MAX_NU || N < 0) return TACD_EINVAL;
  for (int i = 0; i < N; ++i) {
    int it;
    if (dtype == TACD_F32)
      it = pnqp1_f32(m, (const float*)H + (size_t)i * m * m, (const float*)q + (size_t)i * m, (const float*)lo + (size_t)i * m,
                     (const float*)hi + (size_t)i * m, (float*)x + (size_t)i * m);
    else
      it = pnqp1_f64(m, (const double*)H + (size_t)i * m * m, (const double*)q + (size_t)i * m, (const double*)lo + (size_t)i * m,
                     (const double*)hi + (size_t)i * m, (double*)x + (size_t)i * m);
    if (iters) iters[i] = it;
  }
  return TACD_OK;
}
