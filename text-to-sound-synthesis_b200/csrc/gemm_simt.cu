// Exact fp32 (FFMA) GEMM: out = epi(A W^T + bias) (+ residual).  A (M,K), W (N,K) both K-contiguous.
// Not the throughput path -- used for set-up time tables (AdaLN timestep MLP) and as the "fp32-exact" mode that proves
// free-running token parity against the fp32 reference.  64x64 tiles, 4x4 per thread.
#include "common.cuh"
#include "diffsound_b200.h"

namespace dsb {
constexpr int TS = 64, TK = 16;

__global__ void __launch_bounds__(256)
gemm_f32_kernel(const float* __restrict__ A, const float* __restrict__ W, const float* __restrict__ bias, const float* residual, float* out,
                int M, int N, int K, long long lda, long long ldw, long long ldo, long long ld_res, int flags) {
  __shared__ float As[TK][TS + 1];
  __shared__ float Ws[TK][TS + 1];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int m0 = blockIdx.y * TS, n0 = blockIdx.x * TS;
  float acc[4][4] = {};
  const int lr = threadIdx.x >> 2;        // 0..63 tile row
  const int lk = (threadIdx.x & 3) * 4;   // 0,4,8,12
  for (int k0 = 0; k0 < K; k0 += TK) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int k = k0 + lk + j;
      As[lk + j][lr] = (m0 + lr < M && k < K) ? A[(long long)(m0 + lr) * lda + k] : 0.f;
      Ws[lk + j][lr] = (n0 + lr < N && k < K) ? W[(long long)(n0 + lr) * ldw + k] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < TK; ++k) {
      float a[4], w[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { a[i] = As[k][ty * 4 + i]; w[i] = Ws[k][tx * 4 + i]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], w[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= N) continue;
      float v = acc[i][j];
      if (bias) v += bias[n];
      if (flags & DSB_GEMM_GELU2) v = v / (1.0f + expf(-1.702f * v));
      if (flags & DSB_GEMM_LRELU) v = v > 0.f ? v : 0.2f * v;
      if (flags & DSB_GEMM_TANH) v = tanhf(v);
      if (residual) v += residual[(long long)m * ld_res + n];
      if (flags & DSB_GEMM_ROUND_TF32) v = round_tf32(v);
      out[(long long)m * ldo + n] = v;
    }
  }
}
}  // namespace dsb
using namespace dsb;

extern "C" int dsb_gemm_f32(const float* A, const float* W, const float* bias, const float* residual, float* out, int M, int N, int K,
                            long long lda, long long ldw, long long ldo, long long ld_res, int flags, void* stream) {
  DSB_REQUIRE(M > 0 && N > 0 && K > 0, "dsb_gemm_f32: bad shape");
  DSB_REQUIRE((flags & DSB_GEMM_OUT_BF16) == 0, "dsb_gemm_f32: fp32 output only");
  dim3 grid((N + TS - 1) / TS, (M + TS - 1) / TS);
  gemm_f32_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(A, W, bias, residual, out, M, N, K, lda, ldw, ldo, ld_res, flags);
  DSB_CHECK_CUDA(cudaGetLastError());
  return 0;
}
