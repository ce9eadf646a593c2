// Error plumbing, device info and elementwise helpers of libdiffsound_b200.so.
#include "common.cuh"
#include "diffsound_b200.h"
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cstdarg>
#include <cstdlib>

namespace dsb {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
bool pdl_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("DSB_PDL");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v == 1;
}
int sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
  }
  return n;
}

__global__ void round_tf32_kernel(const float* __restrict__ in, float* __restrict__ out, long long n) {
  long long i = (blockIdx.x * (long long)blockDim.x + threadIdx.x) * 4;
  const long long stride = (long long)gridDim.x * blockDim.x * 4;
  for (; i + 3 < n; i += stride) {
    float4 v = *reinterpret_cast<const float4*>(in + i);
    v.x = round_tf32(v.x); v.y = round_tf32(v.y); v.z = round_tf32(v.z); v.w = round_tf32(v.w);
    *reinterpret_cast<float4*>(out + i) = v;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0)
    for (long long j = n & ~3LL; j < n; ++j) out[j] = round_tf32(in[j]);
}
__global__ void f32_to_bf16_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ out, long long n) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < n; i += stride) out[i] = __float2bfloat16(in[i]);
}
__global__ void f32_to_f16_kernel(const float* __restrict__ in, __half* __restrict__ out, long long n) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < n; i += stride) out[i] = __float2half_rn(in[i]);
}
__global__ void split_tf32_kernel(const float* __restrict__ in, long long ld_in, float* __restrict__ out, long long ld_out, long long rows, int C, int Cp, int wfmt) {
  const int c4n = Cp / 4;
  const bool vec = ((C & 3) == 0) && ((ld_in & 3) == 0) && ((reinterpret_cast<uintptr_t>(in) & 15) == 0);
  const long long total = rows * c4n;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / c4n;
    const int c = (int)(i - r * c4n) * 4;
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    if (vec && c + 3 < C) {
      const float4 t = *reinterpret_cast<const float4*>(in + r * ld_in + c);
      v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    } else {
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (c + k < C) v[k] = in[r * ld_in + c + k];
    }
    float4 hi = make_float4(round_tf32(v[0]), round_tf32(v[1]), round_tf32(v[2]), round_tf32(v[3]));
    float4 lo = make_float4(round_tf32(v[0] - hi.x), round_tf32(v[1] - hi.y), round_tf32(v[2] - hi.z), round_tf32(v[3] - hi.w));
    float* o = out + r * ld_out + c;  // ld_out and Cp are multiples of 4 -> 16-byte aligned
    *reinterpret_cast<float4*>(o) = hi;
    if (wfmt) {  // weight-side layout [hi | hi | lo]
      *reinterpret_cast<float4*>(o + Cp) = hi;
      *reinterpret_cast<float4*>(o + 2 * Cp) = lo;
    } else {     // activation-side layout [hi | lo]
      *reinterpret_cast<float4*>(o + Cp) = lo;
    }
  }
}
// fp16 (hi | lo) split of scale * in: hi = f16(x), lo = f16(x - hi)  (x - hi is exact in fp32)
__global__ void split_f16_kernel(const float* __restrict__ in, long long ld_in, __half* __restrict__ out, long long ld_out, long long lo_off,
                                 long long rows, int C, float scale) {
  const long long total = rows * C;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / C;
    const int c = (int)(i - r * C);
    const float x = in[r * ld_in + c] * scale;
    const __half h = __float2half_rn(x);
    out[r * ld_out + c] = h;
    out[r * ld_out + lo_off + c] = __float2half_rn(x - __half2float(h));
  }
}
__global__ void silu_kernel(const float* __restrict__ in, float* __restrict__ out, long long n) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    const float x = in[i];
    out[i] = x / (1.0f + expf(-x));
  }
}
}  // namespace dsb
using namespace dsb;

extern "C" const char* dsb_last_error(void) { return g_err; }
extern "C" int dsb_version(void) { return DSB_VERSION; }
extern "C" int dsb_device_info(int* sms, int* major, int* minor) {
  int dev = 0;
  DSB_CHECK_CUDA(cudaGetDevice(&dev));
  cudaDeviceProp prop;
  DSB_CHECK_CUDA(cudaGetDeviceProperties(&prop, dev));
  if (sms) *sms = prop.multiProcessorCount;
  if (major) *major = prop.major;
  if (minor) *minor = prop.minor;
  return 0;
}
static int grid_for(long long n, int per_block) {
  long long g = (n + per_block - 1) / per_block;
  const long long cap = (long long)sm_count() * 16;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}
extern "C" int dsb_round_tf32(const float* in, float* out, long long n, void* stream) {
  DSB_REQUIRE(((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 15) == 0, "dsb_round_tf32: pointers must be 16-byte aligned");
  round_tf32_kernel<<<grid_for(n, 1024), 256, 0, (cudaStream_t)stream>>>(in, out, n);
  DSB_CHECK_CUDA(cudaGetLastError());
  return 0;
}
extern "C" int dsb_f32_to_bf16(const float* in, void* out, long long n, void* stream) {
  f32_to_bf16_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(in, (__nv_bfloat16*)out, n);
  DSB_CHECK_CUDA(cudaGetLastError());
  return 0;
}
extern "C" int dsb_f32_to_f16(const float* in, void* out, long long n, void* stream) {
  f32_to_f16_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(in, (__half*)out, n);
  DSB_CHECK_CUDA(cudaGetLastError());
  return 0;
}
extern "C" int dsb_split_tf32(const float* in, long long ld_in, float* out, long long ld_out, long long rows, int C, int Cp, int wfmt, void* stream) {
  DSB_REQUIRE(rows > 0 && C > 0 && Cp >= C && ld_out >= (wfmt ? 3LL : 2LL) * Cp, "dsb_split_tf32: bad shape");
  DSB_REQUIRE(Cp % 4 == 0 && ld_out % 4 == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0, "dsb_split_tf32: output must be 16-byte aligned with Cp %% 4 == 0");
  split_tf32_kernel<<<grid_for(rows * (Cp / 4), 256), 256, 0, (cudaStream_t)stream>>>(in, ld_in, out, ld_out, rows, C, Cp, wfmt);
  DSB_CHECK_CUDA(cudaGetLastError());
  return 0;
}
extern "C" int dsb_split_f16(const float* in, long long ld_in, void* out, long long ld_out, long long lo_off, long long rows, int C, float scale,
                             void* stream) {
  DSB_REQUIRE(rows > 0 && C > 0 && lo_off >= C && ld_out >= lo_off + C, "dsb_split_f16: bad shape (rows=%lld C=%d lo_off=%lld ld_out=%lld)", rows, C, lo_off, ld_out);
  split_f16_kernel<<<grid_for(rows * C, 256), 256, 0, (cudaStream_t)stream>>>(in, ld_in, (__half*)out, ld_out, lo_off, rows, C, scale == 0.f ? 1.f : scale);
  DSB_CHECK_CUDA(cudaGetLastError());
  return 0;
}
extern "C" int dsb_silu(const float* in, float* out, long long n, void* stream) {
  silu_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(in, out, n);
  DSB_CHECK_CUDA(cudaGetLastError());
  return 0;
}
