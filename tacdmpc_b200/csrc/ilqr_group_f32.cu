// fp32 instantiation of the five-lanes-per-element forward kernel
#include "ilqr_group_launch.cuh"
namespace tacd {
template int launch_forward_group<float>(const tacd_problem*, const tacd_forward_io*, void*, size_t, cudaStream_t);
template int launch_backward_group<float>(const tacd_problem*, const tacd_backward_io*, void*, size_t, cudaStream_t);
}
