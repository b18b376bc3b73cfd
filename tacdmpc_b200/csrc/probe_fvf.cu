// probe_fvf.cu — the tensor-core question of BASELINE.json's north star, as an experiment that can be re-run:
// "Tensor cores are used only if ncu shows a batched-GEMM formulation beats the register-resident path at the tested n_x."
//
// The dominant arithmetic of one Riccati step at n_x >= 16 is the Q-function  Q_b = C_b + F' V_b F  (controller.py:551-558:
// A'VA, A'VB, B'VB with F = [A | B] shared by the batch for the LTI plugin of config 5), 2 n_x n (n_x + n) multiply-adds per
// element and step (92 k at n_x = 32, n_u = 8).  tacd_probe_fvf() evaluates it for B elements with three formulations, REPS
// times per element on operands staged once in shared memory, so that the measured time is the product itself:
//   kind 0   the round-1 generic kernel's scheme: CTA of 128 threads per element, every thread 1 x 4 output tiles, operands read
//            from shared memory with 128-bit loads, packed FFMA2 (ilqr_generic.cuh::gen_mac4)
//   kind 1   register-resident: one WARP per element, lane i owns row i of V and of W = V F in registers, F is read with
//            broadcast 128-bit loads; the second product F' W goes through one shared-memory transpose of W
//   kind 2   tensor cores: one warp per element, mma.sync.m16n8k8 TF32 with the 3-term split (a_hi b_hi + a_hi b_lo + a_lo b_hi)
//            that keeps fp32-level accuracy (plain TF32 has a 2^-11 mantissa: 5e-4, far outside the 1e-5 parity bar)
// bench: tools/bench_fvf.py (CUDA events + error against fp64); ncu counters: profiles/r2_probe_fvf_*.
#include "host_common.h"

namespace tacd {

template <int NX, int NU>
struct FvfShape {
  static constexpr int N = NX + NU;
  static constexpr int LDF = N;          // F [NX][N] row-major
};

// ---- kind 0: CTA per element, 1 x 4 tiles from shared memory ---------------------------------------------------------------
template <int NX, int NU>
__global__ void __launch_bounds__(128) fvf_cta_tiles(int B, int reps, const float* __restrict__ Vg, const float* __restrict__ Fg,
                                                      const float* __restrict__ Cg, float* __restrict__ Qg) {
  constexpr int N = NX + NU;
  __shared__ __align__(16) float sV[NX * NX], sF[NX * N], sW[NX * N], sQ[N * N];
  const int tid = threadIdx.x;
  for (int i = tid; i < NX * N; i += 128) sF[i] = Fg[i];
  for (int b = blockIdx.x; b < B; b += gridDim.x) {
    __syncthreads();
    for (int i = tid; i < NX * NX; i += 128) sV[i] = Vg[(size_t)b * NX * NX + i];
    for (int i = tid; i < N * N; i += 128) sQ[i] = Cg[(size_t)b * N * N + i];
    __syncthreads();
    for (int r = 0; r < reps; ++r) {
      // W = V F  (NX x N): tile (i, j0..j0+3)
      for (int tl = tid; tl < NX * N / 4; tl += 128) {
        const int i = tl / (N / 4), j0 = (tl % (N / 4)) * 4;
        float a[4] = {0, 0, 0, 0};
#pragma unroll 8
        for (int k = 0; k < NX; ++k) {
          const float l = sV[i * NX + k];
          const float4 rr = *reinterpret_cast<const float4*>(sF + k * N + j0);
          a[0] += l * rr.x; a[1] += l * rr.y; a[2] += l * rr.z; a[3] += l * rr.w;
        }
        *reinterpret_cast<float4*>(sW + i * N + j0) = make_float4(a[0], a[1], a[2], a[3]);
      }
      __syncthreads();
      // Q += F' W  (N x N): tile (a, j0..j0+3), F read transposed
      for (int tl = tid; tl < N * N / 4; tl += 128) {
        const int i = tl / (N / 4), j0 = (tl % (N / 4)) * 4;
        float a[4] = {0, 0, 0, 0};
#pragma unroll 8
        for (int k = 0; k < NX; ++k) {
          const float l = sF[k * N + i];
          const float4 rr = *reinterpret_cast<const float4*>(sW + k * N + j0);
          a[0] += l * rr.x; a[1] += l * rr.y; a[2] += l * rr.z; a[3] += l * rr.w;
        }
        float4& q = *reinterpret_cast<float4*>(sQ + i * N + j0);
        q.x += a[0]; q.y += a[1]; q.z += a[2]; q.w += a[3];
      }
      __syncthreads();
      if (r + 1 < reps)   // feed the result back so that no repetition can be hoisted: V <- V + 1e-6 Q_xx
        for (int i = tid; i < NX * NX; i += 128) sV[i] += 1e-6f * sQ[(i / NX) * N + (i % NX)];
      __syncthreads();
    }
    for (int i = tid; i < N * N; i += 128) Qg[(size_t)b * N * N + i] = sQ[i];
  }
}

// ---- kind 1: warp per element, lane = row, operands in registers -----------------------------------------------------------
template <int NX, int NU>
__global__ void __launch_bounds__(128) fvf_lane_rows(int B, int reps, const float* __restrict__ Vg, const float* __restrict__ Fg,
                                                      const float* __restrict__ Cg, float* __restrict__ Qg) {
  constexpr int N = NX + NU;
  static_assert(NX <= 32 && N <= 64 && N % 4 == 0, "lane i owns row i (and row 32 + i) of the N x N result");
  constexpr int R2 = N > 32 ? N - 32 : 0;      // rows 32.. of Q are second rows of lanes 0..R2-1
  __shared__ __align__(16) float sF[NX * N];
  __shared__ __align__(16) float sWall[4][NX * N];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  float* sW = sWall[wid];
  for (int i = threadIdx.x; i < NX * N; i += 128) sF[i] = Fg[i];
  __syncthreads();
  const int li = lane < NX ? lane : NX - 1;    // (NX < 32: the spare lanes shadow the last row)
  // column `lane` (and 32 + lane) of F in registers: the left operand of the second product
  float fc0[NX], fc1[NX];
#pragma unroll
  for (int k = 0; k < NX; ++k) { fc0[k] = sF[k * N + (lane < N ? lane : N - 1)]; fc1[k] = sF[k * N + (32 + lane < N ? 32 + lane : N - 1)]; }
  for (int b = blockIdx.x * 4 + wid; b < B; b += gridDim.x * 4) {
    float v[NX];
#pragma unroll
    for (int k = 0; k < NX; ++k) v[k] = Vg[(size_t)b * NX * NX + li * NX + k];
    float q0[N], q1[N];
#pragma unroll
    for (int j = 0; j < N; ++j) {
      q0[j] = Cg[(size_t)b * N * N + (size_t)(lane < N ? lane : N - 1) * N + j];
      q1[j] = Cg[(size_t)b * N * N + (size_t)(32 + lane < N ? 32 + lane : N - 1) * N + j];
    }
    for (int r = 0; r < reps; ++r) {
      // W[i][:] = sum_k V[i][k] F[k][:]: the row of F is a broadcast 128-bit load, 4 multiply-adds per load
      float w[N];
#pragma unroll
      for (int j = 0; j < N; ++j) w[j] = 0;
#pragma unroll
      for (int k = 0; k < NX; ++k) {
#pragma unroll
        for (int j0 = 0; j0 < N; j0 += 4) {
          const float4 f4 = *reinterpret_cast<const float4*>(sF + k * N + j0);
          w[j0] += v[k] * f4.x; w[j0 + 1] += v[k] * f4.y; w[j0 + 2] += v[k] * f4.z; w[j0 + 3] += v[k] * f4.w;
        }
      }
      __syncwarp();
      if (lane < NX)
#pragma unroll
        for (int j0 = 0; j0 < N; j0 += 4) *reinterpret_cast<float4*>(sW + lane * N + j0) = make_float4(w[j0], w[j0 + 1], w[j0 + 2], w[j0 + 3]);
      __syncwarp();
      // Q[a][:] += sum_i F[i][a] W[i][:], a = lane (and 32 + lane): W rows broadcast from shared memory
#pragma unroll
      for (int i = 0; i < NX; ++i) {
#pragma unroll
        for (int j0 = 0; j0 < N; j0 += 4) {
          const float4 w4 = *reinterpret_cast<const float4*>(sW + i * N + j0);
          q0[j0] += fc0[i] * w4.x; q0[j0 + 1] += fc0[i] * w4.y; q0[j0 + 2] += fc0[i] * w4.z; q0[j0 + 3] += fc0[i] * w4.w;
          if (R2 > 0) { q1[j0] += fc1[i] * w4.x; q1[j0 + 1] += fc1[i] * w4.y; q1[j0 + 2] += fc1[i] * w4.z; q1[j0 + 3] += fc1[i] * w4.w; }
        }
      }
      if (r + 1 < reps)
#pragma unroll
        for (int k = 0; k < NX; ++k) v[k] += 1e-6f * q0[k];     // V <- V + 1e-6 Q_xx (row `lane`)
    }
    if (lane < N)
#pragma unroll
      for (int j = 0; j < N; ++j) Qg[(size_t)b * N * N + (size_t)lane * N + j] = q0[j];
    if (32 + lane < N)
#pragma unroll
      for (int j = 0; j < N; ++j) Qg[(size_t)b * N * N + (size_t)(32 + lane) * N + j] = q1[j];
  }
}

// ---- kind 2: warp per element, mma.sync m16n8k8 TF32 x 3 -------------------------------------------------------------------
__device__ __forceinline__ void split_tf32(float x, unsigned& hi, unsigned& lo) {
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hi) : "f"(x));
  const float rest = x - __uint_as_float(hi);
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(lo) : "f"(rest));
}
__device__ __forceinline__ void mma_tf32(float (&d)[4], const unsigned (&a)[4], const unsigned (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
// D (16 x 8 tile at (m0, n0)) += A[m0.., :K] B[:K, n0..]   with A(m, k) = Am[m * lda + k] or, AT, Am[k * lda + m]; B(k, n) = Bm[k * ldb + n];
// rows m >= MA of A read as zero (padding of the last 16-row tile)
template <int K, bool AT>
__device__ __forceinline__ void mma_tile_3x(float (&d)[4], const float* Am, int lda, int MA, int m0, const float* Bm, int ldb, int n0, int lane) {
  const int g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int k0 = 0; k0 < K; k0 += 8) {
    float af[4], bf[2];
    const int r0 = m0 + g, r1 = m0 + g + 8;
    af[0] = r0 < MA ? (AT ? Am[(k0 + t) * lda + r0] : Am[r0 * lda + k0 + t]) : 0.f;
    af[1] = r1 < MA ? (AT ? Am[(k0 + t) * lda + r1] : Am[r1 * lda + k0 + t]) : 0.f;
    af[2] = r0 < MA ? (AT ? Am[(k0 + t + 4) * lda + r0] : Am[r0 * lda + k0 + t + 4]) : 0.f;
    af[3] = r1 < MA ? (AT ? Am[(k0 + t + 4) * lda + r1] : Am[r1 * lda + k0 + t + 4]) : 0.f;
    bf[0] = Bm[(k0 + t) * ldb + n0 + g];
    bf[1] = Bm[(k0 + t + 4) * ldb + n0 + g];
    unsigned ah[4], al[4], bh[2], bl[2];
#pragma unroll
    for (int i = 0; i < 4; ++i) split_tf32(af[i], ah[i], al[i]);
#pragma unroll
    for (int i = 0; i < 2; ++i) split_tf32(bf[i], bh[i], bl[i]);
    mma_tf32(d, al, bh);     // small terms first
    mma_tf32(d, ah, bl);
    mma_tf32(d, ah, bh);
  }
}
template <int NX, int NU>
__global__ void __launch_bounds__(128) fvf_mma_tf32x3(int B, int reps, const float* __restrict__ Vg, const float* __restrict__ Fg,
                                                       const float* __restrict__ Cg, float* __restrict__ Qg) {
  constexpr int N = NX + NU;
  static_assert(NX % 16 == 0 && N % 8 == 0, "16 x 8 output tiles");
  extern __shared__ __align__(16) float fvf_smem[];      // F, then per warp V, W, Q (66 KB at n_x = 32: dynamic)
  float* sF = fvf_smem;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
  float* sV = fvf_smem + NX * N + wid * (NX * NX + NX * N + N * N);
  float *sW = sV + NX * NX, *sQ = sW + NX * N;
  for (int i = threadIdx.x; i < NX * N; i += 128) sF[i] = Fg[i];
  __syncthreads();
  for (int b = blockIdx.x * 4 + wid; b < B; b += gridDim.x * 4) {
    __syncwarp();
    for (int i = lane; i < NX * NX; i += 32) sV[i] = Vg[(size_t)b * NX * NX + i];
    for (int i = lane; i < N * N; i += 32) sQ[i] = Cg[(size_t)b * N * N + i];
    __syncwarp();
    for (int r = 0; r < reps; ++r) {
      // W = V F: (NX/16) x (N/8) tiles
#pragma unroll
      for (int m0 = 0; m0 < NX; m0 += 16)
#pragma unroll
        for (int n0 = 0; n0 < N; n0 += 8) {
          float d[4] = {0, 0, 0, 0};
          mma_tile_3x<NX, false>(d, sV, NX, NX, m0, sF, N, n0, lane);
          sW[(m0 + g) * N + n0 + 2 * t] = d[0]; sW[(m0 + g) * N + n0 + 2 * t + 1] = d[1];
          sW[(m0 + g + 8) * N + n0 + 2 * t] = d[2]; sW[(m0 + g + 8) * N + n0 + 2 * t + 1] = d[3];
        }
      __syncwarp();
      // Q += F' W: ceil(N/16) x (N/8) tiles, A = F transposed, rows >= N padded with zeros
#pragma unroll
      for (int m0 = 0; m0 < N; m0 += 16)
#pragma unroll
        for (int n0 = 0; n0 < N; n0 += 8) {
          float d[4];
          const int r0 = m0 + g, r1 = m0 + g + 8;
          d[0] = r0 < N ? sQ[r0 * N + n0 + 2 * t] : 0.f; d[1] = r0 < N ? sQ[r0 * N + n0 + 2 * t + 1] : 0.f;
          d[2] = r1 < N ? sQ[r1 * N + n0 + 2 * t] : 0.f; d[3] = r1 < N ? sQ[r1 * N + n0 + 2 * t + 1] : 0.f;
          mma_tile_3x<NX, true>(d, sF, N, N, m0, sW, N, n0, lane);
          if (r0 < N) { sQ[r0 * N + n0 + 2 * t] = d[0]; sQ[r0 * N + n0 + 2 * t + 1] = d[1]; }
          if (r1 < N) { sQ[r1 * N + n0 + 2 * t] = d[2]; sQ[r1 * N + n0 + 2 * t + 1] = d[3]; }
        }
      __syncwarp();
      if (r + 1 < reps)
        for (int i = lane; i < NX * NX; i += 32) sV[i] += 1e-6f * sQ[(i / NX) * N + (i % NX)];
      __syncwarp();
    }
    for (int i = lane; i < N * N; i += 32) Qg[(size_t)b * N * N + i] = sQ[i];
  }
}

template <int NX, int NU>
static int launch_fvf(int kind, int B, int reps, const float* V, const float* F, const float* C, float* Q, cudaStream_t s) {
  const DeviceInfo di = device_info();
  if (kind == 0) fvf_cta_tiles<NX, NU><<<di.sms * 4, 128, 0, s>>>(B, reps, V, F, C, Q);
  else if (kind == 1) fvf_lane_rows<NX, NU><<<di.sms * 4, 128, 0, s>>>(B, reps, V, F, C, Q);
  else {
    constexpr int N = NX + NU;
    const size_t smem = sizeof(float) * (NX * N + 4 * (NX * NX + NX * N + N * N));
    int rc = check_cuda(cudaFuncSetAttribute(fvf_mma_tf32x3<NX, NU>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "fvf probe smem");
    if (rc) return rc;
    fvf_mma_tf32x3<NX, NU><<<di.sms * 2, 128, smem, s>>>(B, reps, V, F, C, Q);
  }
  note_launches(1);
  return check_cuda(cudaPeekAtLastError(), "fvf probe launch");
}

}  // namespace tacd

using namespace tacd;

extern "C" int tacd_probe_fvf(int32_t kind, int32_t nx, int32_t nu, int32_t B, int32_t reps, const void* V, const void* F, const void* C,
                              void* Q, void* stream) {
  if (kind < 0 || kind > 2 || B < 1 || reps < 1 || !V || !F || !C || !Q) { set_error("tacd_probe_fvf: bad argument"); return TACD_EINVAL; }
  cudaStream_t s = (cudaStream_t)stream;
  if (nx == 32 && nu == 8) return launch_fvf<32, 8>(kind, B, reps, (const float*)V, (const float*)F, (const float*)C, (float*)Q, s);
  if (nx == 16 && nu == 8) return launch_fvf<16, 8>(kind, B, reps, (const float*)V, (const float*)F, (const float*)C, (float*)Q, s);
  set_error("tacd_probe_fvf: instantiated for (nx, nu) in {(32, 8), (16, 8)}");
  return TACD_EUNSUPPORTED;
}
