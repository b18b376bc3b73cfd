// probe.cu — FP32-pipe microbenchmark ("an FMA microbenchmark beside the nominal 74.45 TFLOP/s").
// The forward solve is bound by the FP32 / issue pipes, not by HBM, so its roofline denominator has to be measured on the
// same part and clocks as the solve itself: bench.py launches this probe, times it with CUDA events and prints the result
// next to the nominal 148 SM x 128 lanes x 2 x f_SM.
#include "host_common.h"

namespace tacd {

constexpr int kProbeChains = 16;   // independent accumulators per thread (latency 4, issue every 1-2 cycles: >= 8 needed)

// KIND 0: scalar FFMA, all three sources in registers (what the solve kernels issue).
// KIND 1: packed fma.rn.f32x2 (SASS FFMA2): two IEEE fp32 FMAs per lane and instruction.
template <int KIND>
__global__ void __launch_bounds__(256) fp32_probe(int iters, const float* __restrict__ seed, float* __restrict__ out) {
  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
  // multiplier / addend come from memory so that ptxas cannot fold them into immediates or constant-bank operands
  const float a = seed[0], b = seed[1];
  float x[kProbeChains];
#pragma unroll
  for (int k = 0; k < kProbeChains; ++k) x[k] = seed[2] + (float)(tid + k);
  if (KIND == 0) {
    for (int i = 0; i < iters; ++i) {
#pragma unroll
      for (int k = 0; k < kProbeChains; ++k) x[k] = fmaf(x[k], a, b);
    }
  } else {
    unsigned long long ab, bb, xb[kProbeChains / 2];
    asm("mov.b64 %0, {%1, %1};" : "=l"(ab) : "f"(a));
    asm("mov.b64 %0, {%1, %1};" : "=l"(bb) : "f"(b));
#pragma unroll
    for (int k = 0; k < kProbeChains / 2; ++k) asm("mov.b64 %0, {%1, %2};" : "=l"(xb[k]) : "f"(x[2 * k]), "f"(x[2 * k + 1]));
    for (int i = 0; i < iters; ++i) {
#pragma unroll
      for (int k = 0; k < kProbeChains / 2; ++k) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(xb[k]) : "l"(ab), "l"(bb));
    }
#pragma unroll
    for (int k = 0; k < kProbeChains / 2; ++k) asm("mov.b64 {%0, %1}, %2;" : "=f"(x[2 * k]), "=f"(x[2 * k + 1]) : "l"(xb[k]));
  }
  float s = 0;
#pragma unroll
  for (int k = 0; k < kProbeChains; ++k) s += x[k];
  out[tid] = s;
}

}  // namespace tacd

using namespace tacd;

extern "C" int tacd_fp32_probe(int32_t kind, int32_t iters, int32_t blocks, const void* seed, void* out, double* flops, void* stream) {
  if (kind < 0 || kind > 1 || iters < 1 || blocks < 1 || !seed || !out) { set_error("tacd_fp32_probe: bad argument"); return TACD_EINVAL; }
  cudaStream_t s = (cudaStream_t)stream;
  if (kind == 0) fp32_probe<0><<<blocks, 256, 0, s>>>(iters, (const float*)seed, (float*)out);
  else fp32_probe<1><<<blocks, 256, 0, s>>>(iters, (const float*)seed, (float*)out);
  note_launches(1);
  if (flops) *flops = 2.0 * kProbeChains * (double)iters * 256.0 * (double)blocks;
  return check_cuda(cudaPeekAtLastError(), "fp32_probe launch");
}
