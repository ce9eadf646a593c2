// MelGAN generator support kernel (reference vocoder/modules.py:72-130): LeakyReLU + (reflection | zero) padding into a
// channels-last (B, T + 2*pad, C) buffer, optionally reading a channel-major (B, C, T) input (the mel).  All convolutions
// (k=7, dilated k=3, 1x1, and the polyphase form of ConvTranspose1d) run on the tcgen05 GEMM with row-shift taps.
#include "common.cuh"
#include "diffsound_b200.h"
#include <cuda_fp16.h>

namespace dsb {
__global__ void lrelu_pad_kernel(const float* __restrict__ in, float* __restrict__ out, int B, int T, int C, int pad, float slope, int reflect,
                                 int in_channel_major, int flags) {
  const int Tp = T + 2 * pad;
  const bool split = flags & DSB_SPLIT_OUT;
  const int Cp = split ? (C + 31) / 32 * 32 : C;  // split rows: [hi (Cp) | lo (Cp)], padding columns zero
  const long long total = (long long)B * Tp * Cp;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = i % Cp;
    const long long r = i / Cp;
    const int b = r / Tp;
    int t = (int)(r % Tp) - pad;
    float v = 0.f;
    bool valid = true;
    if (t < 0) { if (reflect) t = -t; else valid = false; }
    else if (t >= T) { if (reflect) t = 2 * (T - 1) - t; else valid = false; }
    if (c >= C) valid = false;
    if (valid) {
      v = in_channel_major ? in[((long long)b * C + c) * T + t] : in[((long long)b * T + t) * C + c];
      v = v > 0.f ? v : v * slope;
      if (flags & DSB_GEMM_ROUND_TF32) v = round_tf32(v);
    }
    if (split) {
      const float hi = round_tf32(v);
      out[r * 2 * Cp + c] = hi;
      out[r * 2 * Cp + Cp + c] = round_tf32(v - hi);
    } else {
      out[i] = v;
    }
  }
}

// ---- split-fp16 ("f16x3") MelGAN path: activations live in per-stage STATE buffers (B, P + T + P, 4C) fp16 whose rows are
// [raw_hi | raw_lo | act_hi | act_lo] (act = LeakyReLU(0.2)(raw)); every conv is a dsb_gemm_ex call over them.
// mel (B, Cm, T) fp32 channel-major -> (B, T + 2 pad, 2 Kp) fp16 [hi (Kp) | lo (Kp)], reflection-padded in time, channel columns >= Cm zero
__global__ void mel_pack_f16_kernel(const float* __restrict__ mel, __half* __restrict__ out, int B, int Cm, int T, int pad, int Kp) {
  const int Tp = T + 2 * pad;
  const long long total = (long long)B * Tp * Kp;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % Kp);
    const long long r = i / Kp;
    const int b = (int)(r / Tp);
    int t = (int)(r % Tp) - pad;
    if (t < 0) t = -t;
    else if (t >= T) t = 2 * (T - 1) - t;
    const float v = c < Cm ? mel[((long long)b * Cm + c) * T + t] : 0.f;
    const __half h = __float2half_rn(v);
    out[r * 2 * Kp + c] = h;
    out[r * 2 * Kp + Kp + c] = __float2half_rn(v - __half2float(h));
  }
}
// rows P-j and P+T-1+j (j = 1..d) of columns [c0, c0 + ncols) of every clip: reflection of rows P+j / P+T-1-j (ReflectionPad1d, reference
// vocoder/modules.py:77) or zeros (the implicit zero padding of ConvTranspose1d's polyphase taps)
__global__ void edge_pad_f16_kernel(__half* __restrict__ s, long long ld, long long bstride, int B, int T, int P, int d, int c0, int ncols, int reflect) {
  const int n8 = ncols / 8;
  const long long total = (long long)B * 2 * d * n8;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % n8) * 8;
    long long r = i / n8;
    const int j = (int)(r % d) + 1;
    r /= d;
    const int side = (int)(r & 1);
    const int b = (int)(r >> 1);
    const int dst = side == 0 ? P - j : P + T - 1 + j;
    const int src = side == 0 ? P + j : P + T - 1 - j;
    __half* base = s + (long long)b * bstride + c0 + c;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (reflect) v = *reinterpret_cast<const uint4*>(base + (long long)src * ld);
    *reinterpret_cast<uint4*>(base + (long long)dst * ld) = v;
  }
}

// MelGAN's output layer (reference vocoder/modules.py:121-126: LeakyReLU -> ReflectionPad1d(3) -> Conv1d(ngf, 1, k=7) -> tanh) straight off the
// last stage's state buffer: one output channel is no tensor-core shape (the GEMM form stages 7 x 16 KB boxes per 128 samples and measured
// 764 us per 16 clips against 70 us of HBM time), so this runs on the FMA pipe: a block converts TILE + KT - 1 rows of the activated (hi | lo)
// pair to fp32 in shared memory (chunk-swizzled: conflict-free 16-byte reads at a stride of four rows), each thread accumulates four
// consecutive samples so that a staged row is read 2.5x instead of 7x, weights are fp32 (exact), accumulation fp32.
template <int CS, int KT>
__global__ void __launch_bounds__(128) conv_out_pair_kernel(const __half* __restrict__ s, long long ld, long long bstride, int T, int row0, int col0,
                                                            const float* __restrict__ w, const float* __restrict__ bias, float scale,
                                                            float* __restrict__ out) {
  constexpr int TILE = 512, ROWS = TILE + KT - 1, C8 = CS / 8, C4 = CS / 4;
  static_assert(CS == 32, "the chunk swizzle below assumes one 128-byte shared-memory row per time step");
  extern __shared__ __align__(16) float cx_smem[];
  float* xs = cx_smem;              // ROWS x CS
  float* ws = cx_smem + ROWS * CS;  // KT x CS
  const int b = blockIdx.y;
  const int t0 = blockIdx.x * TILE;
  const __half* sb = s + (long long)b * bstride + col0;
  for (int i = threadIdx.x; i < KT * CS; i += blockDim.x) ws[i] = w[i];
  // staging loop: four independent 2 x 16-byte loads in flight per thread before any conversion (a one-item-per-iteration loop measured
  // latency-bound: 1.9 TB/s)
  constexpr int ITEMS = ROWS * C8, U = 4;
  for (int i0 = threadIdx.x; i0 < ITEMS; i0 += U * 128) {
    uint4 h[U], l[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int i = i0 + u * 128;
      const int r = i / C8, k = i - r * C8;
      h[u] = make_uint4(0u, 0u, 0u, 0u);
      l[u] = h[u];
      if (i < ITEMS && t0 + r < T + KT - 1) {
        const __half* rp = sb + (long long)(row0 + t0 + r) * ld + 8 * k;
        h[u] = __ldg(reinterpret_cast<const uint4*>(rp));
        l[u] = __ldg(reinterpret_cast<const uint4*>(rp + CS));
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int i = i0 + u * 128;
      if (i >= ITEMS) break;
      const int r = i / C8, k = i - r * C8;
      const __half2* h2 = reinterpret_cast<const __half2*>(&h[u]);
      const __half2* l2 = reinterpret_cast<const __half2*>(&l[u]);
      float v[8];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 a = __half22float2(h2[e]), c = __half22float2(l2[e]);
        v[2 * e] = a.x + c.x;
        v[2 * e + 1] = a.y + c.y;
      }
      const int sw = (r >> 2) & 7;
      *reinterpret_cast<float4*>(xs + r * CS + (((2 * k) ^ sw) << 2)) = make_float4(v[0], v[1], v[2], v[3]);
      *reinterpret_cast<float4*>(xs + r * CS + (((2 * k + 1) ^ sw) << 2)) = make_float4(v[4], v[5], v[6], v[7]);
    }
  }
  __syncthreads();
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  const int rb = threadIdx.x * 4;
#pragma unroll 2
  for (int c4 = 0; c4 < C4; ++c4) {
    float4 xr[KT + 3];
#pragma unroll
    for (int r = 0; r < KT + 3; ++r) xr[r] = *reinterpret_cast<const float4*>(xs + (rb + r) * CS + ((c4 ^ (((rb + r) >> 2) & 7)) << 2));
#pragma unroll
    for (int j = 0; j < KT; ++j) {
      const float4 wj = *reinterpret_cast<const float4*>(ws + j * CS + c4 * 4);
#pragma unroll
      for (int o = 0; o < 4; ++o) {
        acc[o] = fmaf(xr[o + j].x, wj.x, acc[o]);
        acc[o] = fmaf(xr[o + j].y, wj.y, acc[o]);
        acc[o] = fmaf(xr[o + j].z, wj.z, acc[o]);
        acc[o] = fmaf(xr[o + j].w, wj.w, acc[o]);
      }
    }
  }
  const float bz = bias ? __ldg(bias) : 0.f;
  float y[4];
#pragma unroll
  for (int o = 0; o < 4; ++o) {  // the GEMM epilogue's tanh: 1 - 2 / (1 + exp(2x)), clamped so exp stays finite
    const float z = fminf(fmaxf(fmaf(acc[o], scale, bz), -15.f), 15.f);
    y[o] = 1.0f - __fdividef(2.0f, 1.0f + __expf(2.0f * z));
  }
  const int t = t0 + rb;
  float* op = out + (long long)b * T + t;
  if (t + 3 < T && ((reinterpret_cast<uintptr_t>(op) & 15) == 0)) {
    *reinterpret_cast<float4*>(op) = make_float4(y[0], y[1], y[2], y[3]);
  } else {
#pragma unroll
    for (int o = 0; o < 4; ++o)
      if (t + o < T) op[o] = y[o];
  }
}
}  // namespace dsb
using namespace dsb;

extern "C" int dsb_conv_out_pair(const void* state, long long ld, long long batch_stride, int B, int T, int row0, int col0, int cs, int kt, const float* w,
                                 const float* bias, float scale, float* out, void* stream) {
  DSB_REQUIRE(B > 0 && T > 0 && row0 >= 0, "dsb_conv_out_pair: bad shape");
  DSB_REQUIRE(cs == 32 && kt == 7, "dsb_conv_out_pair: built for 32-channel rows and 7 taps (MelGAN's output conv); got cs=%d kt=%d -- use dsb_gemm_ex", cs, kt);
  DSB_REQUIRE(ld % 8 == 0 && batch_stride % 8 == 0 && col0 % 8 == 0 && (reinterpret_cast<uintptr_t>(state) & 15) == 0 && (reinterpret_cast<uintptr_t>(w) & 15) == 0,
              "dsb_conv_out_pair: the state rows / weights must be 16-byte aligned");
  auto kern = conv_out_pair_kernel<32, 7>;
  const int smem = ((512 + 6) * 32 + 7 * 32) * 4;
  static bool attr_done = false;
  if (!attr_done) {
    DSB_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_done = true;
  }
  kern<<<dim3((T + 511) / 512, B), 128, smem, (cudaStream_t)stream>>>((const __half*)state, ld, batch_stride, T, row0, col0, w, bias, scale, out);
  DSB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int dsb_mel_pack_f16(const float* mel, void* out, int B, int Cm, int T, int pad, int Kp, void* stream) {
  DSB_REQUIRE(B > 0 && Cm > 0 && T > pad && pad >= 0 && Kp >= Cm, "dsb_mel_pack_f16: bad shape");
  const long long total = (long long)B * (T + 2 * pad) * Kp;
  long long g = (total + 255) / 256;
  const long long cap = (long long)sm_count() * 8;
  mel_pack_f16_kernel<<<(unsigned)(g > cap ? cap : g), 256, 0, (cudaStream_t)stream>>>(mel, (__half*)out, B, Cm, T, pad, Kp);
  DSB_CHECK_CUDA(cudaGetLastError());
  return 0;
}
extern "C" int dsb_edge_pad_f16(void* state, long long ld, long long batch_stride, int B, int T, int P, int d, int col0, int ncols, int reflect, void* stream) {
  DSB_REQUIRE(B > 0 && T > d && d >= 1 && d <= P && ncols > 0, "dsb_edge_pad_f16: bad shape (T=%d P=%d d=%d)", T, P, d);
  DSB_REQUIRE(ncols % 8 == 0 && col0 % 8 == 0 && ld % 8 == 0 && batch_stride % 8 == 0 && (reinterpret_cast<uintptr_t>(state) & 15) == 0,
              "dsb_edge_pad_f16: columns, leading dimension and batch stride must be multiples of 8 halves (16 bytes)");
  const long long total = (long long)B * 2 * d * (ncols / 8);
  edge_pad_f16_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>((__half*)state, ld, batch_stride, B, T, P, d, col0, ncols, reflect);
  DSB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int dsb_lrelu_pad(const float* in, float* out, int B, int T, int C, int pad, float slope, int reflect, int in_channel_major, int flags,
                             void* stream) {
  DSB_REQUIRE(B > 0 && T > 0 && C > 0 && pad >= 0 && (!reflect || pad < T), "dsb_lrelu_pad: bad shape");
  const long long total = (long long)B * (T + 2 * pad) * ((flags & DSB_SPLIT_OUT) ? (C + 31) / 32 * 32 : C);
  long long g = (total + 255) / 256;
  const long long cap = (long long)sm_count() * 8;
  lrelu_pad_kernel<<<(unsigned)(g > cap ? cap : g), 256, 0, (cudaStream_t)stream>>>(in, out, B, T, C, pad, slope, reflect, in_channel_major, flags);
  DSB_CHECK_CUDA(cudaGetLastError());
  return 0;
}
