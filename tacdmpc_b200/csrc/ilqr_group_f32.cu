// fp32 instantiation of the five-lanes-per-element forward kernel
#include "ilqr_group_launch.cuh"
namespace tacd {
template int launch_forward_group<float>(const tacd_problem*, const tacd_forward_io*, void*, size_t, cudaStream_t);
template int launch_backward_group<float>(const tacd_problem*, const tacd_backward_io*, void*, size_t, cudaStream_t);
// bytes of the candidates' state tiles in the forward workspace (one [T nx][32] tile per resident warp, sized for doubles)
size_t group_forward_scratch_bytes(const tacd_problem* p) {
  if (p->nx > 4 || p->nx + p->nu > kG) return 0;
  const DeviceInfo di = device_info();
  const size_t sms = di.sms > 0 ? di.sms : 148, smem = di.smem_optin > 0 ? di.smem_optin : 232448;
  const size_t es = p->dtype == TACD_F32 ? 4 : 8;
  const size_t warp_b = (size_t)group_warp_words(p->T, p->nx, p->nu) * es;
  if (warp_b > smem) return 0;   // horizon too long for the kernel: another kernel takes the problem
  size_t nw = smem / warp_b;
  if (nw > (size_t)kGroupMaxWarps) nw = kGroupMaxWarps;
  return sms * nw * group_xscratch_words(p->T, p->nx) * es + 256;
}
}
