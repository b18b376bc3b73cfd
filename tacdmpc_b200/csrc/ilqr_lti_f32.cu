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
