// SpecVQGAN decoder support kernels on zero-padded channels-last buffers (B, H+2, W+2, C):
//   codebook gather + ColumnMajor un-permute (reference dalle_spec.py:80-91, permuter.py:46-49, quantize.py:88-103),
//   GroupNorm(32, eps 1e-6) statistics / apply (+swish) (model.py:29-35), nearest x2 upsample (model.py:48-52),
//   AttnBlock plumbing (model.py:202-226): token compaction, masked row softmax, scatter-add back into the padded image.
// The 3x3 / 1x1 convolutions themselves run on the tcgen05 GEMM with 9 / 1 taps (gemm_tcgen05.cu); a zero border makes a
// 3x3 tap a pure row shift of the flattened (b, y, x) index.  All kernels here are HBM-bound elementwise/reduction passes.
#include "common.cuh"
#include "diffsound_b200.h"
#include <cuda_fp16.h>

namespace dsb {

// fp16 (hi | lo) pair of four fp32 values: hi at o[0..3], lo at o[lo_off..lo_off+3]  (8-byte stores)
__device__ __forceinline__ void store_pair_f16(__half* o, long long lo_off, float4 v) {
  const __half2 h0 = __floats2half2_rn(v.x, v.y), h1 = __floats2half2_rn(v.z, v.w);
  const __half2 l0 = __floats2half2_rn(v.x - __low2float(h0), v.y - __high2float(h0));
  const __half2 l1 = __floats2half2_rn(v.z - __low2float(h1), v.w - __high2float(h1));
  uint2 u, w;
  u.x = *reinterpret_cast<const uint32_t*>(&h0); u.y = *reinterpret_cast<const uint32_t*>(&h1);
  w.x = *reinterpret_cast<const uint32_t*>(&l0); w.y = *reinterpret_cast<const uint32_t*>(&l1);
  *reinterpret_cast<uint2*>(o) = u;
  *reinterpret_cast<uint2*>(o + lo_off) = w;
}

__global__ void codebook_gather_padded_kernel(const int64_t* __restrict__ ids, const float* __restrict__ codebook, float* __restrict__ out,
                                              int B, int H, int W, int E, int n_codes, int flags, int* err_flag) {
  const int Hp = H + 2, Wp = W + 2;
  const long long row = blockIdx.x * (long long)(blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= (long long)B * Hp * Wp) return;
  const int lane = threadIdx.x & 31;
  const int b = row / (Hp * Wp);
  const int p = row % (Hp * Wp);
  const int y = p / Wp - 1, x = p % Wp - 1;
  const bool split = flags & DSB_SPLIT_OUT;
  if (flags & DSB_SPLIT_OUT_F16) {  // rows of 2E halves: [hi | lo]
    __half* oh = reinterpret_cast<__half*>(out) + row * 2 * E;
    const bool inside = !(y < 0 || y >= H || x < 0 || x >= W);
    long long id = 0;
    if (inside) {
      id = ids[(long long)b * H * W + (long long)x * H + y];
      if (id < 0 || id >= n_codes) {
        if (lane == 0 && err_flag) atomicExch(err_flag, 1);
        id = 0;
      }
    }
    const float4* c = reinterpret_cast<const float4*>(codebook + id * E);
    for (int i = lane; i < E / 4; i += 32) store_pair_f16(oh + i * 4, E, inside ? __ldg(c + i) : make_float4(0.f, 0.f, 0.f, 0.f));
    return;
  }
  float4* o = reinterpret_cast<float4*>(out + row * (split ? 2 * E : E));
  if (y < 0 || y >= H || x < 0 || x >= W) {
    for (int i = lane; i < (split ? E / 2 : E / 4); i += 32) o[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    return;
  }
  // ColumnMajor reverse: row-major position (y, x) holds the token at column-major index x*H + y
  long long id = ids[(long long)b * H * W + (long long)x * H + y];
  if (id < 0 || id >= n_codes) {
    if (lane == 0 && err_flag) atomicExch(err_flag, 1);
    id = 0;
  }
  const float4* c = reinterpret_cast<const float4*>(codebook + id * E);
  const bool rnd = flags & DSB_GEMM_ROUND_TF32;
  for (int i = lane; i < E / 4; i += 32) {
    float4 v = __ldg(c + i);
    if (split) {
      const float4 hi = make_float4(round_tf32(v.x), round_tf32(v.y), round_tf32(v.z), round_tf32(v.w));
      o[i] = hi;
      o[E / 4 + i] = make_float4(round_tf32(v.x - hi.x), round_tf32(v.y - hi.y), round_tf32(v.z - hi.z), round_tf32(v.w - hi.w));
      continue;
    }
    if (rnd) { v.x = round_tf32(v.x); v.y = round_tf32(v.y); v.z = round_tf32(v.z); v.w = round_tf32(v.w); }
    o[i] = v;
  }
}

// stats[b][g] = (sum, sumsq) in fp64 over all rows of image b (border rows are exact zeros, so they do not contribute)
__global__ void __launch_bounds__(256)
groupnorm_stats_kernel(const float* __restrict__ x, double* __restrict__ stats, int P, int C, int groups, int rows_per_block) {
  extern __shared__ double sacc[];  // [groups][2]
  const int b = blockIdx.y;
  const int c4n = C / 4;
  const int cg = C / groups;
  for (int i = threadIdx.x; i < groups * 2; i += blockDim.x) sacc[i] = 0.0;
  __syncthreads();
  const int r0 = blockIdx.x * rows_per_block;
  const int r1 = min(P, r0 + rows_per_block);
  const float* xb = x + (long long)b * P * C;
  const int tpr = blockDim.x / c4n > 0 ? blockDim.x / c4n : 1;  // threads per row-slot
  const int c4 = threadIdx.x % c4n;
  const int rs = threadIdx.x / c4n;
  if (rs < tpr) {
    float s[4] = {0.f, 0.f, 0.f, 0.f}, q[4] = {0.f, 0.f, 0.f, 0.f};
    int r = r0 + rs;
    for (; r + 3 * tpr < r1; r += 4 * tpr) {  // four independent 16-byte loads in flight
      float4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) v[u] = *reinterpret_cast<const float4*>(xb + (long long)(r + u * tpr) * C + c4 * 4);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        s[0] += v[u].x; s[1] += v[u].y; s[2] += v[u].z; s[3] += v[u].w;
        q[0] += v[u].x * v[u].x; q[1] += v[u].y * v[u].y; q[2] += v[u].z * v[u].z; q[3] += v[u].w * v[u].w;
      }
    }
    for (; r < r1; r += tpr) {
      const float4 v = *reinterpret_cast<const float4*>(xb + (long long)r * C + c4 * 4);
      s[0] += v.x; s[1] += v.y; s[2] += v.z; s[3] += v.w;
      q[0] += v.x * v.x; q[1] += v.y * v.y; q[2] += v.z * v.z; q[3] += v.w * v.w;
    }
    if (cg % 4 == 0 && blockDim.x == 256 && 256 % c4n == 0) {
      // a thread's four channels share one group: stage one (sum, sumsq) pair per thread, then ONE thread per group adds its members in a fixed
      // order (no contended shared-memory fp64 atomics: they are CAS loops and dominated the small low-resolution launches)
      __shared__ double part[256][2];
      part[threadIdx.x][0] = (double)s[0] + (double)s[1] + (double)s[2] + (double)s[3];
      part[threadIdx.x][1] = (double)q[0] + (double)q[1] + (double)q[2] + (double)q[3];
      __syncthreads();
      if (threadIdx.x < groups) {
        const int g = threadIdx.x, per = cg / 4;
        double a = 0.0, c = 0.0;
        for (int rr = 0; rr < tpr; ++rr)
          for (int k = 0; k < per; ++k) {
            a += part[rr * c4n + g * per + k][0];
            c += part[rr * c4n + g * per + k][1];
          }
        sacc[g * 2] = a;
        sacc[g * 2 + 1] = c;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int g = (c4 * 4 + j) / cg;
        atomicAdd(&sacc[g * 2], (double)s[j]);
        atomicAdd(&sacc[g * 2 + 1], (double)q[j]);
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < groups * 2; i += blockDim.x) atomicAdd(&stats[(long long)b * groups * 2 + i], sacc[i]);
}

// out = [swish]( (x - mean) * rstd * gamma + beta ) on interior pixels, 0 on the border.
// COMPACT mode writes tokens (B, Lp, C) (row-major pixel order, rows >= H*W zero) instead of the padded image.
__global__ void groupnorm_apply_kernel(const float* __restrict__ x, const double* __restrict__ stats, const float* __restrict__ gamma,
                                       const float* __restrict__ beta, float* __restrict__ out, int B, int H, int W, int C, int groups, float eps,
                                       int flags, int Lp) {
  const int Hp = H + 2, Wp = W + 2, c4n = C / 4;
  const bool compact = flags & DSB_GN_COMPACT;
  const long long total = (compact ? (long long)B * Lp : (long long)B * Hp * Wp) * c4n;
  const int cg = C / groups;
  const double cnt = (double)H * W * cg;
  // per-(image, group) mean / rstd once per CTA (fp64 -> fp32), not per element: the element loop below is pure fp32 and HBM-bound
  extern __shared__ float gn_ms[];  // [B * groups][2]
  const bool cached = B * groups <= 2048;
  if (cached) {
    for (int i = threadIdx.x; i < B * groups; i += blockDim.x) {
      const double mean = stats[(long long)i * 2] / cnt;
      double var = stats[(long long)i * 2 + 1] / cnt - mean * mean;
      var = var < 0.0 ? 0.0 : var;
      gn_ms[2 * i] = (float)mean;
      gn_ms[2 * i + 1] = (float)(1.0 / sqrt(var + (double)eps));
    }
    __syncthreads();
  }
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c4 = i % c4n;
    const long long orow = i / c4n;
    int b, y, xx;
    bool inside;
    if (compact) {
      b = orow / Lp;
      const int tkn = orow % Lp;
      inside = tkn < H * W;
      y = tkn / W; xx = tkn % W;
    } else {
      b = orow / (Hp * Wp);
      const int p = orow % (Hp * Wp);
      y = p / Wp - 1; xx = p % Wp - 1;
      inside = y >= 0 && y < H && xx >= 0 && xx < W;
    }
    float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
    if (inside) {
      const float4 v = *reinterpret_cast<const float4*>(x + (((long long)b * Hp + y + 1) * Wp + xx + 1) * C + c4 * 4);
      const float vv[4] = {v.x, v.y, v.z, v.w};
      float o[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int c = c4 * 4 + j;
        const int g = c / cg;
        float meanf, rstd;
        if (cached) {
          meanf = gn_ms[2 * (b * groups + g)];
          rstd = gn_ms[2 * (b * groups + g) + 1];
        } else {
          const double mean = stats[((long long)b * groups + g) * 2] / cnt;
          double var = stats[((long long)b * groups + g) * 2 + 1] / cnt - mean * mean;
          var = var < 0.0 ? 0.0 : var;
          meanf = (float)mean;
          rstd = (float)(1.0 / sqrt(var + (double)eps));
        }
        float t = (vv[j] - meanf) * rstd * __ldg(gamma + c) + __ldg(beta + c);
        if (flags & DSB_GN_SWISH) t = __fdividef(t, 1.0f + __expf(-t));
        if (flags & DSB_GEMM_ROUND_TF32) t = round_tf32(t);
        o[j] = t;
      }
      r = make_float4(o[0], o[1], o[2], o[3]);
    }
    if (flags & DSB_SPLIT_OUT_F16) {  // rows of 2*C halves: [hi | lo]
      store_pair_f16(reinterpret_cast<__half*>(out) + orow * 2 * C + c4 * 4, C, r);
      continue;
    }
    if (flags & DSB_SPLIT_OUT) {  // rows of 2*C floats: [hi | lo]
      const float4 hi = make_float4(round_tf32(r.x), round_tf32(r.y), round_tf32(r.z), round_tf32(r.w));
      *reinterpret_cast<float4*>(out + orow * 2 * C + c4 * 4) = hi;
      *reinterpret_cast<float4*>(out + orow * 2 * C + C + c4 * 4) =
          make_float4(round_tf32(r.x - hi.x), round_tf32(r.y - hi.y), round_tf32(r.z - hi.z), round_tf32(r.w - hi.w));
      continue;
    }
    *reinterpret_cast<float4*>(out + orow * C + c4 * 4) = r;
  }
}

// Padded-image form of the same op, one CTA per (image, padded row): with 256 % (C/4) == 0 a thread keeps ONE float4 channel slot for the whole
// row, so its mean / rstd / gamma / beta are loop invariants and the element loop has no integer division at all (the flat-index kernel above
// spends ~5 64-bit div/mod per 16 bytes and measured 2 TB/s on the 80 x 848 level); four independent 16-byte loads in flight per thread.
__global__ void __launch_bounds__(256) groupnorm_apply_rows_kernel(const float* __restrict__ x, const double* __restrict__ stats,
                                                                   const float* __restrict__ gamma, const float* __restrict__ beta,
                                                                   float* __restrict__ out, int H, int W, int C, int groups, float eps, int flags) {
  const int Hp = H + 2, Wp = W + 2, c4n = C / 4;
  const int b = blockIdx.x / Hp, yp = blockIdx.x - b * Hp;
  const int c4 = threadIdx.x % c4n, x0 = threadIdx.x / c4n, xstep = 256 / c4n;
  const long long rowbase = ((long long)b * Hp + yp) * Wp;
  const bool row_inside = yp >= 1 && yp <= H;
  const int cg = C / groups;
  float meanf[4], rstd[4], ga[4], be[4];
  if (row_inside) {
    const double cnt = (double)H * W * cg;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = c4 * 4 + j;
      const int g = c / cg;
      if (j > 0 && g == (c - 1) / cg) {
        meanf[j] = meanf[j - 1];
        rstd[j] = rstd[j - 1];
      } else {
        const double mean = stats[((long long)b * groups + g) * 2] / cnt;
        double var = stats[((long long)b * groups + g) * 2 + 1] / cnt - mean * mean;
        var = var < 0.0 ? 0.0 : var;
        meanf[j] = (float)mean;
        rstd[j] = (float)(1.0 / sqrt(var + (double)eps));
      }
      ga[j] = __ldg(gamma + c);
      be[j] = __ldg(beta + c);
    }
  }
  constexpr int U = 4;
  for (int xb = x0; xb < Wp; xb += U * xstep) {
    float4 v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int xx = xb + u * xstep;
      v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (row_inside && xx >= 1 && xx <= W) v[u] = *reinterpret_cast<const float4*>(x + (rowbase + xx) * C + c4 * 4);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int xx = xb + u * xstep;
      if (xx >= Wp) break;
      float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
      if (row_inside && xx >= 1 && xx <= W) {
        const float vv[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
        float o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float t = (vv[j] - meanf[j]) * rstd[j] * ga[j] + be[j];
          if (flags & DSB_GN_SWISH) t = __fdividef(t, 1.0f + __expf(-t));
          if (flags & DSB_GEMM_ROUND_TF32) t = round_tf32(t);
          o[j] = t;
        }
        r = make_float4(o[0], o[1], o[2], o[3]);
      }
      const long long orow = rowbase + xx;
      if (flags & DSB_SPLIT_OUT_F16) {
        store_pair_f16(reinterpret_cast<__half*>(out) + orow * 2 * C + c4 * 4, C, r);
      } else if (flags & DSB_SPLIT_OUT) {
        const float4 hi = make_float4(round_tf32(r.x), round_tf32(r.y), round_tf32(r.z), round_tf32(r.w));
        *reinterpret_cast<float4*>(out + orow * 2 * C + c4 * 4) = hi;
        *reinterpret_cast<float4*>(out + orow * 2 * C + C + c4 * 4) =
            make_float4(round_tf32(r.x - hi.x), round_tf32(r.y - hi.y), round_tf32(r.z - hi.z), round_tf32(r.w - hi.w));
      } else {
        *reinterpret_cast<float4*>(out + orow * C + c4 * 4) = r;
      }
    }
  }
}

__global__ void upsample2x_padded_kernel(const float* __restrict__ in, float* __restrict__ out, int B, int H, int W, int C, int flags) {
  const int Hi = H + 2, Wi = W + 2, Ho = 2 * H + 2, Wo = 2 * W + 2, c4n = C / 4;
  const long long total = (long long)B * Ho * Wo * c4n;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c4 = i % c4n;
    const long long orow = i / c4n;
    const int b = orow / (Ho * Wo);
    const int p = orow % (Ho * Wo);
    const int y = p / Wo - 1, x = p % Wo - 1;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (y >= 0 && y < 2 * H && x >= 0 && x < 2 * W) {
      v = *reinterpret_cast<const float4*>(in + (((long long)b * Hi + (y >> 1) + 1) * Wi + (x >> 1) + 1) * C + c4 * 4);
      if (flags & DSB_GEMM_ROUND_TF32) { v.x = round_tf32(v.x); v.y = round_tf32(v.y); v.z = round_tf32(v.z); v.w = round_tf32(v.w); }
    }
    if (flags & DSB_SPLIT_OUT_F16) {
      store_pair_f16(reinterpret_cast<__half*>(out) + orow * 2 * C + c4 * 4, C, v);
      continue;
    }
    if (flags & DSB_SPLIT_OUT) {
      const float4 hi = make_float4(round_tf32(v.x), round_tf32(v.y), round_tf32(v.z), round_tf32(v.w));
      *reinterpret_cast<float4*>(out + orow * 2 * C + c4 * 4) = hi;
      *reinterpret_cast<float4*>(out + orow * 2 * C + C + c4 * 4) =
          make_float4(round_tf32(v.x - hi.x), round_tf32(v.y - hi.y), round_tf32(v.z - hi.z), round_tf32(v.w - hi.w));
      continue;
    }
    *reinterpret_cast<float4*>(out + orow * C + c4 * 4) = v;
  }
}

// Downsample (model.py:55-75: zero pad (0,1,0,1) + 3x3 stride-2 conv) as a tap GEMM: the padded input (B, H+2, W+2, C) is rearranged into
// its four stride-2 phases side by side, on the OUTPUT's padded grid (B, H/2+2, W/2+2, 4C):
//   out[b, i+1, j+1, (2p+q) C + c] = in_unpadded[b, 2i+p, 2j+q, c]   (zero beyond the image: that is the reference's right / bottom pad)
// so that tap (dy, dx) of the strided conv becomes the constant row shift (dy/2) (W/2+2) + dx/2 with A column offset (2 (dy%2) + dx%2) C.
// DSB_SPLIT_OUT writes the split-TF32 operand [hi (4C) | lo (4C)].
__global__ void space_to_depth_padded_kernel(const float* __restrict__ in, float* __restrict__ out, int B, int H, int W, int C, int flags) {
  const int Hi = H + 2, Wi = W + 2, Ho = H / 2 + 2, Wo = W / 2 + 2, c4n = C / 4;
  const long long total = (long long)B * Ho * Wo * 4 * c4n;
  const int ocols = (flags & DSB_SPLIT_OUT) ? 8 * C : 4 * C;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int c4 = idx % c4n;
    const int ph = (idx / c4n) % 4;
    const long long orow = idx / (4 * c4n);
    const int b = orow / (Ho * Wo);
    const int pp = orow % (Ho * Wo);
    const int i = pp / Wo - 1, j = pp % Wo - 1;
    const int r = 2 * i + (ph >> 1), s_ = 2 * j + (ph & 1);  // unpadded source pixel
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (i >= 0 && j >= 0 && r < H && s_ < W) v = *reinterpret_cast<const float4*>(in + (((long long)b * Hi + r + 1) * Wi + s_ + 1) * C + c4 * 4);
    float* o = out + orow * ocols + ph * C + c4 * 4;
    if (flags & DSB_SPLIT_OUT) {
      const float4 hi = make_float4(round_tf32(v.x), round_tf32(v.y), round_tf32(v.z), round_tf32(v.w));
      *reinterpret_cast<float4*>(o) = hi;
      *reinterpret_cast<float4*>(o + 4 * C) = make_float4(round_tf32(v.x - hi.x), round_tf32(v.y - hi.y), round_tf32(v.z - hi.z), round_tf32(v.w - hi.w));
    } else {
      if (flags & DSB_GEMM_ROUND_TF32) { v.x = round_tf32(v.x); v.y = round_tf32(v.y); v.z = round_tf32(v.z); v.w = round_tf32(v.w); }
      *reinterpret_cast<float4*>(o) = v;
    }
  }
}

// out[r] = argmin_k x[r, k] over the first n columns, first index on ties (VectorQuantizer.forward, quantize.py:56-63); one warp per row
__global__ void row_argmin_kernel(const float* __restrict__ x, long long ld, long long rows, int n, int64_t* __restrict__ out) {
  const long long row = blockIdx.x * (long long)(blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const float* r = x + row * ld;
  float best = INFINITY;
  int bi = 0x7fffffff;
  for (int k = lane; k < n; k += 32) {
    const float v = r[k];
    if (v < best) { best = v; bi = k; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ob = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ob < best || (ob == best && oi < bi)) { best = ob; bi = oi; }
  }
  if (lane == 0) out[row] = bi;
}

// in-place softmax over the first n_valid columns of each row (one warp per row); columns [n_valid, ld) are set to zero
__global__ void softmax_rows_kernel(float* __restrict__ x, long long rows, int n_valid, int ld, int flags) {
  const long long row = blockIdx.x * (long long)(blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  float* r = x + row * ld;
  float m = -INFINITY;
  for (int i = lane; i < n_valid; i += 32) m = fmaxf(m, r[i]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  float s = 0.f;
  for (int i = lane; i < n_valid; i += 32) s += expf(r[i] - m);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float inv = 1.0f / s;
  for (int i = lane; i < ld; i += 32) {
    float v = i < n_valid ? expf(r[i] - m) * inv : 0.f;
    if (flags & DSB_GEMM_ROUND_TF32) v = round_tf32(v);
    r[i] = v;
  }
}

// x_pad[b, y+1, x+1, :] += tok[b, y*W + x, :]
__global__ void tokens_add_to_padded_kernel(const float* __restrict__ tok, float* __restrict__ xpad, int B, int H, int W, int C, int Lp) {
  const int Hp = H + 2, Wp = W + 2, c4n = C / 4;
  const long long total = (long long)B * H * W * c4n;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c4 = i % c4n;
    const long long t = i / c4n;
    const int b = t / (H * W);
    const int p = t % (H * W);
    const int y = p / W, x = p % W;
    const float4 a = *reinterpret_cast<const float4*>(tok + ((long long)b * Lp + p) * C + c4 * 4);
    float4* d = reinterpret_cast<float4*>(xpad + (((long long)b * Hp + y + 1) * Wp + x + 1) * C + c4 * 4);
    float4 v = *d;
    v.x += a.x; v.y += a.y; v.z += a.z; v.w += a.w;
    *d = v;
  }
}
}  // namespace dsb
using namespace dsb;

static int ew_grid(long long n) {
  long long g = (n + 255) / 256;
  const long long cap = (long long)sm_count() * 8;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

extern "C" int dsb_codebook_gather_padded(const int64_t* ids, const float* codebook, float* out, int B, int H, int W, int E, int n_codes, int flags,
                                          int* err_flag, void* stream) {
  DSB_REQUIRE(E % 4 == 0, "dsb_codebook_gather_padded: embed dim must be a multiple of 4");
  const long long rows = (long long)B * (H + 2) * (W + 2);
  codebook_gather_padded_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, (cudaStream_t)stream>>>(ids, codebook, out, B, H, W, E, n_codes, flags, err_flag);
  DSB_CHECK_CUDA(cudaGetLastError());
  return 0;
}
extern "C" int dsb_groupnorm_stats(const float* x, double* stats, int B, int P, int C, int groups, void* stream) {
  DSB_REQUIRE(C % 4 == 0 && C % groups == 0 && C / 4 <= 256, "dsb_groupnorm_stats: unsupported channel count %d", C);
  cudaStream_t st = (cudaStream_t)stream;
  DSB_CHECK_CUDA(cudaMemsetAsync(stats, 0, sizeof(double) * 2 * B * groups, st));
  // enough CTAs to fill the machine on the low-resolution levels too (P = 385 padded pixels at 5 x 53)
  long long rpb = ((long long)P * B) / ((long long)sm_count() * 4);
  const int rows_per_block = rpb < 16 ? 16 : (rpb > 256 ? 256 : (int)rpb);
  dim3 grid((P + rows_per_block - 1) / rows_per_block, B);
  groupnorm_stats_kernel<<<grid, 256, sizeof(double) * 2 * groups, st>>>(x, stats, P, C, groups, rows_per_block);
  DSB_CHECK_CUDA(cudaGetLastError());
  return 0;
}
extern "C" int dsb_groupnorm_apply(const float* x, const double* stats, const float* gamma, const float* beta, float* out, int B, int H, int W, int C,
                                   int groups, float eps, int flags, int Lp, void* stream) {
  DSB_REQUIRE(C % 4 == 0 && C % groups == 0, "dsb_groupnorm_apply: unsupported channel count %d", C);
  DSB_REQUIRE(!(flags & DSB_SPLIT_OUT) || C % 32 == 0, "dsb_groupnorm_apply: split output needs C %% 32 == 0");
  DSB_REQUIRE(!(flags & DSB_GN_COMPACT) || Lp >= H * W, "dsb_groupnorm_apply: Lp too small");
  const long long total = ((flags & DSB_GN_COMPACT) ? (long long)B * Lp : (long long)B * (H + 2) * (W + 2)) * (C / 4);
  if (!(flags & DSB_GN_COMPACT) && C / 4 <= 256 && 256 % (C / 4) == 0) {
    groupnorm_apply_rows_kernel<<<B * (H + 2), 256, 0, (cudaStream_t)stream>>>(x, stats, gamma, beta, out, H, W, C, groups, eps, flags);
    DSB_CHECK_CUDA(cudaGetLastError());
    return 0;
  }
  const size_t gn_smem = B * groups <= 2048 ? sizeof(float) * 2 * B * groups : 0;
  groupnorm_apply_kernel<<<ew_grid(total), 256, gn_smem, (cudaStream_t)stream>>>(x, stats, gamma, beta, out, B, H, W, C, groups, eps, flags, Lp);
  DSB_CHECK_CUDA(cudaGetLastError());
  return 0;
}
extern "C" int dsb_upsample2x_padded(const float* in, float* out, int B, int H, int W, int C, int flags, void* stream) {
  DSB_REQUIRE(C % 4 == 0, "dsb_upsample2x_padded: C must be a multiple of 4");
  const long long total = (long long)B * (2 * H + 2) * (2 * W + 2) * (C / 4);
  upsample2x_padded_kernel<<<ew_grid(total), 256, 0, (cudaStream_t)stream>>>(in, out, B, H, W, C, flags);
  DSB_CHECK_CUDA(cudaGetLastError());
  return 0;
}
extern "C" int dsb_space_to_depth_padded(const float* in, float* out, int B, int H, int W, int C, int flags, void* stream) {
  DSB_REQUIRE(C % 4 == 0 && H % 2 == 0 && W % 2 == 0, "dsb_space_to_depth_padded: C must be a multiple of 4 and H, W even (H=%d W=%d C=%d)", H, W, C);
  const long long total = (long long)B * (H / 2 + 2) * (W / 2 + 2) * C;
  space_to_depth_padded_kernel<<<ew_grid(total), 256, 0, (cudaStream_t)stream>>>(in, out, B, H, W, C, flags);
  DSB_CHECK_CUDA(cudaGetLastError());
  return 0;
}
extern "C" int dsb_row_argmin(const float* x, long long ld, long long rows, int n, int64_t* out, void* stream) {
  DSB_REQUIRE(rows > 0 && n > 0 && n <= ld, "dsb_row_argmin: bad shape");
  row_argmin_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, (cudaStream_t)stream>>>(x, ld, rows, n, out);
  DSB_CHECK_CUDA(cudaGetLastError());
  return 0;
}
extern "C" int dsb_softmax_rows(float* x, long long rows, int n_valid, int ld, int flags, void* stream) {
  DSB_REQUIRE(n_valid > 0 && n_valid <= ld, "dsb_softmax_rows: bad n_valid");
  softmax_rows_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, (cudaStream_t)stream>>>(x, rows, n_valid, ld, flags);
  DSB_CHECK_CUDA(cudaGetLastError());
  return 0;
}
extern "C" int dsb_tokens_add_to_padded(const float* tok, float* xpad, int B, int H, int W, int C, int Lp, void* stream) {
  DSB_REQUIRE(C % 4 == 0, "dsb_tokens_add_to_padded: C must be a multiple of 4");
  tokens_add_to_padded_kernel<<<ew_grid((long long)B * H * W * (C / 4)), 256, 0, (cudaStream_t)stream>>>(tok, xpad, B, H, W, C, Lp);
  DSB_CHECK_CUDA(cudaGetLastError());
  return 0;
}
