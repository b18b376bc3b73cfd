// warp-per-element LTI kernels (ilqr_lti.cuh), fp64 instances (parity tests)
#include "ilqr_lti_launch.cuh"
namespace tacd {
template int launch_forward_lti<double>(const tacd_problem*, const tacd_forward_io*, void*, size_t, cudaStream_t);
template int launch_backward_lti<double>(const tacd_problem*, const tacd_backward_io*, void*, size_t, cudaStream_t);
}  // namespace tacd
