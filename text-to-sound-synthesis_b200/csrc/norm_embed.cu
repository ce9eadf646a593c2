// Token embedding gather and LayerNorm / AdaLayerNorm kernels (HBM-bound, one warp per row, float4 accesses).
//   reference: sound_synthesis/modeling/embeddings/dalle_mask_image_embedding.py:36-58
//              sound_synthesis/modeling/transformers/transformer_utils.py:134-149 (AdaLayerNorm), :197, :345 (LayerNorm)
#include "common.cuh"
#include "diffsound_b200.h"
#include <cuda_bf16.h>
#include <cuda_fp16.h>

namespace dsb {

__global__ void embed_tokens_kernel(const int64_t* __restrict__ ids, const float* __restrict__ emb, const float* __restrict__ hemb,
                                    const float* __restrict__ wemb, float* __restrict__ out, int rows, int L, int D, int W, int num_embed,
                                    int* err_flag) {
  const int warps_per_block = blockDim.x >> 5;
  const int row = blockIdx.x * warps_per_block + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const int l = row % L;
  long long id = ids[row];
  if (id < 0) id = 0;  // index[index < 0] = 0  (dalle_mask_image_embedding.py:40)
  if (id >= num_embed) {
    if (lane == 0 && err_flag) atomicExch(err_flag, 1);
    id = num_embed - 1;
  }
  const float4* e = reinterpret_cast<const float4*>(emb + id * D);
  const float4* h = reinterpret_cast<const float4*>(hemb + (long long)(l / W) * D);
  const float4* w = reinterpret_cast<const float4*>(wemb + (long long)(l % W) * D);
  float4* o = reinterpret_cast<float4*>(out + (long long)row * D);
  for (int i = lane; i < D / 4; i += 32) {
    const float4 a = __ldg(e + i), b = __ldg(h + i), c = __ldg(w + i);
    // reference order: emb + (height + width)
    o[i] = make_float4(a.x + (b.x + c.x), a.y + (b.y + c.y), a.z + (b.z + c.z), a.w + (b.w + c.w));
  }
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// MODE 0: plain affine (gamma, beta); MODE 1: AdaLN (table row = scale | shift selected by t[b])
template <int MODE>
__global__ void layernorm_kernel(const float* __restrict__ x, void* __restrict__ out, const float* __restrict__ p0, const float* __restrict__ p1,
                                 const int64_t* __restrict__ t, int rows, int L, int D, int T, float eps, int flags) {
  const int warps_per_block = blockDim.x >> 5;
  const int row = blockIdx.x * warps_per_block + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const float4* xr = reinterpret_cast<const float4*>(x + (long long)row * D);
  const int nv = D / 4;
  float s = 0.f;
  for (int i = lane; i < nv; i += 32) {
    const float4 v = xr[i];
    s += (v.x + v.y) + (v.z + v.w);
  }
  const float mean = warp_sum(s) / (float)D;
  float q = 0.f;
  for (int i = lane; i < nv; i += 32) {
    const float4 v = xr[i];
    const float a = v.x - mean, b = v.y - mean, c = v.z - mean, d = v.w - mean;
    q += (a * a + b * b) + (c * c + d * d);
  }
  const float rstd = rsqrtf(warp_sum(q) / (float)D + eps);
  const float4 *g4, *b4;
  if (MODE == 0) {
    g4 = reinterpret_cast<const float4*>(p0);
    b4 = reinterpret_cast<const float4*>(p1);
  } else {
    long long ti = t[row / L];
    ti = ti < 0 ? 0 : (ti >= T ? T - 1 : ti);
    g4 = reinterpret_cast<const float4*>(p0 + ti * 2LL * D);      // scale
    b4 = reinterpret_cast<const float4*>(p0 + ti * 2LL * D + D);  // shift
  }
  const bool rnd = (flags & DSB_GEMM_ROUND_TF32) != 0;
  const bool obf = (flags & DSB_GEMM_OUT_BF16) != 0;
  const bool of16 = (flags & DSB_GEMM_OUT_F16) != 0;
  for (int i = lane; i < nv; i += 32) {
    const float4 v = xr[i];
    const float4 g = __ldg(g4 + i), b = __ldg(b4 + i);
    float4 y;
    if (MODE == 0) {
      y.x = (v.x - mean) * rstd * g.x + b.x; y.y = (v.y - mean) * rstd * g.y + b.y;
      y.z = (v.z - mean) * rstd * g.z + b.z; y.w = (v.w - mean) * rstd * g.w + b.w;
    } else {
      y.x = (v.x - mean) * rstd * (1.f + g.x) + b.x; y.y = (v.y - mean) * rstd * (1.f + g.y) + b.y;
      y.z = (v.z - mean) * rstd * (1.f + g.z) + b.z; y.w = (v.w - mean) * rstd * (1.f + g.w) + b.w;
    }
    if (of16) {
      __half2 h0 = __floats2half2_rn(y.x, y.y), h1 = __floats2half2_rn(y.z, y.w);
      uint2 u;
      u.x = *reinterpret_cast<uint32_t*>(&h0); u.y = *reinterpret_cast<uint32_t*>(&h1);
      reinterpret_cast<uint2*>(reinterpret_cast<__half*>(out) + (long long)row * D)[i] = u;
    } else if (obf) {
      __nv_bfloat162 h0 = __floats2bfloat162_rn(y.x, y.y), h1 = __floats2bfloat162_rn(y.z, y.w);
      uint2 u;
      u.x = *reinterpret_cast<uint32_t*>(&h0); u.y = *reinterpret_cast<uint32_t*>(&h1);
      reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(out) + (long long)row * D)[i] = u;
    } else {
      if (rnd) { y.x = round_tf32(y.x); y.y = round_tf32(y.y); y.z = round_tf32(y.z); y.w = round_tf32(y.w); }
      reinterpret_cast<float4*>(reinterpret_cast<float*>(out) + (long long)row * D)[i] = y;
    }
  }
}
}  // namespace dsb
using namespace dsb;

extern "C" int dsb_embed_tokens(const int64_t* ids, const float* emb, const float* hemb, const float* wemb, float* out, int B, int L, int D,
                                int H, int W, int num_embed, int* err_flag, void* stream) {
  DSB_REQUIRE(D % 4 == 0, "dsb_embed_tokens: D must be a multiple of 4");
  DSB_REQUIRE(L <= H * W, "dsb_embed_tokens: L=%d exceeds the %dx%d grid", L, H, W);
  const int rows = B * L;
  embed_tokens_kernel<<<(rows + 7) / 8, 256, 0, (cudaStream_t)stream>>>(ids, emb, hemb, wemb, out, rows, L, D, W, num_embed, err_flag);
  DSB_CHECK_CUDA(cudaGetLastError());
  return 0;
}
extern "C" int dsb_layernorm(const float* x, void* out, const float* gamma, const float* beta, int rows, int D, float eps, int flags, void* stream) {
  DSB_REQUIRE(D % 4 == 0, "dsb_layernorm: D must be a multiple of 4");
  layernorm_kernel<0><<<(rows + 7) / 8, 256, 0, (cudaStream_t)stream>>>(x, out, gamma, beta, nullptr, rows, 1, D, 0, eps, flags);
  DSB_CHECK_CUDA(cudaGetLastError());
  return 0;
}
extern "C" int dsb_ada_layernorm(const float* x, void* out, const float* table, const int64_t* t, int B, int L, int D, int T, float eps, int flags,
                                 void* stream) {
  DSB_REQUIRE(D % 4 == 0, "dsb_ada_layernorm: D must be a multiple of 4");
  const int rows = B * L;
  layernorm_kernel<1><<<(rows + 7) / 8, 256, 0, (cudaStream_t)stream>>>(x, out, table, nullptr, t, rows, L, D, T, eps, flags);
  DSB_CHECK_CUDA(cudaGetLastError());
  return 0;
}
