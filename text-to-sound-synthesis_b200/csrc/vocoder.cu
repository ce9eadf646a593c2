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
}  // namespace dsb
using namespace dsb;

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
