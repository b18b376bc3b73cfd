// warp-per-element LTI kernels (ilqr_lti.cuh), fp32 instances + the workspace size both dtypes need
#include "ilqr_lti_launch.cuh"
namespace tacd {
template int launch_forward_lti<float>(const tacd_problem*, const tacd_forward_io*, void*, size_t, cudaStream_t);
template int launch_backward_lti<float>(const tacd_problem*, const tacd_backward_io*, void*, size_t, cudaStream_t);
size_t lti_ws_bytes(const tacd_problem* p) {
  if (p->dyn_id != TACD_DYN_LTI || !lti_size(p)) return 0;
  const size_t es = (p->dtype == TACD_F32) ? 4 : 8;
  const DeviceInfo di = device_info();
  const size_t sms = di.sms > 0 ? di.sms : 148;
  return 256 + sms * 16 * lti_slice_words(p->T, p->nx, p->nu) * es;   // 16 = the most warps per CTA of any instance
}
}  // namespace tacd

// ---- the gain solve of the LTI kernels in isolation (unit parity: tests/test_gpu_lti.py) -----------------------------------
namespace tacd {
template <typename R, int NX, int NU, bool KKT, bool OLD>
__global__ void __launch_bounds__(32) lti_gains_probe(DevProblem<R> P, int n, const R* Quu, const R* rhs, R* K, R* k, uint8_t* nonpd) {
  using W = LtiWarp<R, NX, NU>;
  using C = LtiCfg<NX, NU>;
  __shared__ __align__(16) R sm[C::F_WORDS + C::WARP_WORDS];
  W w;
  w.init(sm, sm + C::F_WORDS);
  const int lane = threadIdx.x;
  for (int b = blockIdx.x; b < n; b += gridDim.x) {
    __syncwarp();
    for (int i = lane; i < NU * NU; i += 32) w.Quu()[i] = Quu[(size_t)b * NU * NU + i];
    // column i of Q_ux: rhs[b][:, i % 2] (two right-hand sides per problem, alternating over the lanes); q_u = rhs[b][:, 1]
    for (int i = lane; i < NU * NX; i += 32) w.Qux()[i] = rhs[(size_t)b * NU * 2 + (i / NX) * 2 + ((i % NX) & 1)];
    if (lane < NU) w.qv()[NX + lane] = rhs[(size_t)b * NU * 2 + lane * 2 + 1];
    __syncwarp();
    R Kc[NU], kt[NU], Qxi[NU], qu[NU];
    bool np = false;
    if (!OLD) {
      np = w.template gains<KKT>(P, Kc, kt, Qxi, qu);
    } else {
      // the run-time-dimension kernels' routines (ilqr_generic.cuh), every lane redundantly: what csrc/ilqr_generic.cu runs
      R L[NU * NU];
      int piv[NU];
      const R reg2 = KKT ? (R)1e-8 * P.reg : P.reg;
      for (int i = 0; i < NU * NU; ++i) L[i] = w.Quu()[i] + ((i / NU == i % NU) ? reg2 : (R)0);
      bool chol = false;
      if (!KKT) chol = gen_chol_reg<R, NU>(L);
      if (!chol) { np = !KKT; gen_lu_reg<R, NU>(L, piv); }
      for (int mm = 0; mm < NU; ++mm) { Kc[mm] = w.Qux()[mm * NX + w.li_()]; kt[mm] = w.qv()[NX + mm]; }
      if (chol) { gen_chol_solve_col_reg<R, NU>(L, Kc, 1, 0); gen_chol_solve_col_reg<R, NU>(L, kt, 1, 0); }
      else { gen_lu_solve_col_reg<R, NU>(L, piv, Kc, 1, 0); gen_lu_solve_col_reg<R, NU>(L, piv, kt, 1, 0); }
      for (int mm = 0; mm < NU; ++mm) { Kc[mm] = -Kc[mm]; kt[mm] = -kt[mm]; }
    }
    if (lane == 0) {
      for (int mm = 0; mm < NU; ++mm) { K[(size_t)b * NU + mm] = Kc[mm]; k[(size_t)b * NU + mm] = kt[mm]; }
      nonpd[b] = np ? 1 : 0;
    }
  }
}
}  // namespace tacd

extern "C" int tacd_probe_lti_gains(int32_t dtype, int32_t nu, int32_t kkt, int32_t n, double reg, const void* Quu, const void* rhs, void* K,
                                    void* k, uint8_t* nonpd, void* stream) {
  using namespace tacd;
  if (dtype != TACD_F32 || (nu != 4 && nu != 8) || n < 1 || !Quu || !rhs || !K || !k || !nonpd) { set_error("tacd_probe_lti_gains: bad argument"); return TACD_EINVAL; }
  DevProblem<float> P{};
  P.reg = (float)reg;
  cudaStream_t s = (cudaStream_t)stream;
  const int grid = n < 1184 ? n : 1184;
#define TACD_GP(NX_, NU_, KKT_, OLD_) lti_gains_probe<float, NX_, NU_, KKT_, OLD_><<<grid, 32, 0, s>>>(P, n, (const float*)Quu, (const float*)rhs, (float*)K, (float*)k, nonpd)
  const bool kk = (kkt & 1) != 0, old = (kkt & 2) != 0;   // bit 1: the run-time-dimension kernels' routines instead
  if (nu == 4) {
    if (old) { if (kk) TACD_GP(8, 4, true, true); else TACD_GP(8, 4, false, true); }
    else { if (kk) TACD_GP(8, 4, true, false); else TACD_GP(8, 4, false, false); }
  } else {
    if (old) { if (kk) TACD_GP(16, 8, true, true); else TACD_GP(16, 8, false, true); }
    else { if (kk) TACD_GP(16, 8, true, false); else TACD_GP(16, 8, false, false); }
  }
#undef TACD_GP
  note_launches(1);
  return check_cuda(cudaPeekAtLastError(), "lti gains probe launch");
}
