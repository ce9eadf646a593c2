// MelGAN generator support kernel (reference vocoder/modules.py:72-130): LeakyReLU + (reflection | zero) padding into a
// channels-last (B, T + 2*pad, C) buffer, optionally reading a channel-major (B, C, T) input (the mel).  All convolutions
// (k=7, dilated k=3, 1x1, and the polyphase form of ConvTranspose1d) run on the tcgen05 GEMM with row-shift taps.
#include "common.cuh"
#include "diffsound_b200.h"

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
}  // namespace dsb
using namespace dsb;

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
