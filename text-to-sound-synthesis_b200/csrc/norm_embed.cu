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
  pdl_wait();
  pdl_trigger();
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

// MODE 0: plain affine (gamma, beta); MODE 1: AdaLN (table row = scale | shift selected by t[b]).
// NV = D / 128 float4 vectors per lane, the whole row lives in registers: x is read from HBM exactly once.
template <int MODE, int NV>
__global__ void __launch_bounds__(256)
layernorm_kernel(const float* __restrict__ x, void* __restrict__ out, const float* __restrict__ p0, const float* __restrict__ p1,
                 const int64_t* __restrict__ t, int rows, int L, int D, int T, float eps, int flags) {
  const int warps_per_block = blockDim.x >> 5;
  const int row = blockIdx.x * warps_per_block + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const float4* xr = reinterpret_cast<const float4*>(x + (long long)row * D);
  pdl_wait();
  pdl_trigger();
  float4 v[NV];
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    v[j] = xr[lane + 32 * j];
    s += (v[j].x + v[j].y) + (v[j].z + v[j].w);
  }
  const float mean = warp_sum(s) / (float)D;
  float q = 0.f;
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    const float a = v[j].x - mean, b = v[j].y - mean, c = v[j].z - mean, d = v[j].w - mean;
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
  const int omode = (flags & DSB_GEMM_OUT_F16_SPLIT) ? 3 : ((flags & DSB_GEMM_OUT_F16) ? 1 : ((flags & DSB_GEMM_OUT_BF16) ? 2 : 0));
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    const int i = lane + 32 * j;
    const float4 g = __ldg(g4 + i), b = __ldg(b4 + i);
    float4 y;
    if (MODE == 0) {
      y.x = (v[j].x - mean) * rstd * g.x + b.x; y.y = (v[j].y - mean) * rstd * g.y + b.y;
      y.z = (v[j].z - mean) * rstd * g.z + b.z; y.w = (v[j].w - mean) * rstd * g.w + b.w;
    } else {
      y.x = (v[j].x - mean) * rstd * (1.f + g.x) + b.x; y.y = (v[j].y - mean) * rstd * (1.f + g.y) + b.y;
      y.z = (v[j].z - mean) * rstd * (1.f + g.z) + b.z; y.w = (v[j].w - mean) * rstd * (1.f + g.w) + b.w;
    }
    if (omode == 1) {
      __half2 h0 = __floats2half2_rn(y.x, y.y), h1 = __floats2half2_rn(y.z, y.w);
      uint2 u;
      u.x = *reinterpret_cast<uint32_t*>(&h0); u.y = *reinterpret_cast<uint32_t*>(&h1);
      reinterpret_cast<uint2*>(reinterpret_cast<__half*>(out) + (long long)row * D)[i] = u;
    } else if (omode == 3) {  // fp16 (hi | lo) pair, 2*D columns per row: the A operand of a split-fp16 GEMM
      const __half2 h0 = __floats2half2_rn(y.x, y.y), h1 = __floats2half2_rn(y.z, y.w);
      const __half2 l0 = __floats2half2_rn(y.x - __low2float(h0), y.y - __high2float(h0));
      const __half2 l1 = __floats2half2_rn(y.z - __low2float(h1), y.w - __high2float(h1));
      uint2 u, w;
      u.x = *reinterpret_cast<const uint32_t*>(&h0); u.y = *reinterpret_cast<const uint32_t*>(&h1);
      w.x = *reinterpret_cast<const uint32_t*>(&l0); w.y = *reinterpret_cast<const uint32_t*>(&l1);
      __half* orow = reinterpret_cast<__half*>(out) + (long long)row * 2 * D;
      reinterpret_cast<uint2*>(orow)[i] = u;
      reinterpret_cast<uint2*>(orow + D)[i] = w;
    } else if (omode == 2) {
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

// x[r, :] /= ||x[r, :]||_2  (CLIPTextEmbedding's per-token normalisation, clip_text_embedding.py:78-79); one warp per row, in place
__global__ void __launch_bounds__(256)
l2_normalize_rows_kernel(float* __restrict__ x, long long rows, int D) {
  const long long row = blockIdx.x * 8LL + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  float* r = x + row * D;
  float ss = 0.f;
  for (int i = lane; i < D; i += 32) ss += r[i] * r[i];
  const float inv = 1.0f / sqrtf(warp_sum(ss));
  for (int i = lane; i < D; i += 32) r[i] *= inv;
}

template <int MODE>
static int launch_ln(const float* x, void* out, const float* p0, const float* p1, const int64_t* t, int rows, int L, int D, int T, float eps,
                     int flags, cudaStream_t st) {
  const int grid = (rows + 7) / 8;
  switch (D / 128) {
#define DSB_LN_CASE(N) case N: DSB_CHECK_CUDA(launch_pdl(layernorm_kernel<MODE, N>, dim3(grid), dim3(256), 0, st, x, out, p0, p1, t, rows, L, D, T, eps, flags)); break;
    DSB_LN_CASE(1) DSB_LN_CASE(2) DSB_LN_CASE(3) DSB_LN_CASE(4) DSB_LN_CASE(5) DSB_LN_CASE(6) DSB_LN_CASE(7) DSB_LN_CASE(8)
    DSB_LN_CASE(12) DSB_LN_CASE(16)
#undef DSB_LN_CASE
    default: set_error("layernorm: D=%d unsupported (need D %% 128 == 0 and D/128 in {1..8,12,16})", D); return 2;
  }
  DSB_CHECK_CUDA(cudaGetLastError());
  return 0;
}
}  // namespace dsb
using namespace dsb;

extern "C" int dsb_embed_tokens(const int64_t* ids, const float* emb, const float* hemb, const float* wemb, float* out, int B, int L, int D,
                                int H, int W, int num_embed, int* err_flag, void* stream) {
  DSB_REQUIRE(D % 4 == 0, "dsb_embed_tokens: D must be a multiple of 4");
  DSB_REQUIRE(L <= H * W, "dsb_embed_tokens: L=%d exceeds the %dx%d grid", L, H, W);
  const int rows = B * L;
  DSB_CHECK_CUDA(launch_pdl(embed_tokens_kernel, dim3((rows + 7) / 8), dim3(256), 0, (cudaStream_t)stream, ids, emb, hemb, wemb, out, rows, L, D, W,
                            num_embed, err_flag));
  return 0;
}
extern "C" int dsb_l2_normalize_rows(float* x, long long rows, int D, void* stream) {
  DSB_REQUIRE(rows > 0 && D > 0, "dsb_l2_normalize_rows: bad shape");
  l2_normalize_rows_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, (cudaStream_t)stream>>>(x, rows, D);
  DSB_CHECK_CUDA(cudaGetLastError());
  return 0;
}
extern "C" int dsb_layernorm(const float* x, void* out, const float* gamma, const float* beta, int rows, int D, float eps, int flags, void* stream) {
  DSB_REQUIRE(D % 128 == 0, "dsb_layernorm: D must be a multiple of 128");
  return launch_ln<0>(x, out, gamma, beta, nullptr, rows, 1, D, 0, eps, flags, (cudaStream_t)stream);
}
extern "C" int dsb_ada_layernorm(const float* x, void* out, const float* table, const int64_t* t, int B, int L, int D, int T, float eps, int flags,
                                 void* stream) {
  DSB_REQUIRE(D % 128 == 0, "dsb_ada_layernorm: D must be a multiple of 128");
  return launch_ln<1>(x, out, table, nullptr, t, B * L, L, D, T, eps, flags, (cudaStream_t)stream);
}
