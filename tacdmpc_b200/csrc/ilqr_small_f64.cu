// fp64 instantiations of the thread-per-element iLQR kernels
#include "ilqr_small_launch.cuh"
namespace tacd {
template int launch_forward_small<double>(const tacd_problem*, const tacd_forward_io*, void*, size_t, cudaStream_t, const FwdHandover*);
template int launch_backward_small<double>(const tacd_problem*, const tacd_backward_io*, void*, size_t, cudaStream_t);
}
