// tcgen05 / TMEM attention core (same math as attention_f16.cu; reference transformer_utils.py:48-54, :99-105):
//   S = Q K^T  : tcgen05.mma kind::f16, A = Q tile (128 x 64, smem, K-major SW128), B = K (Lk_pad x 64, smem, K-major SW128),
//                accumulator S (128 lanes x Lk_pad fp32 columns) in TMEM;
//   softmax    : 4 warps, one query row per thread, read S with tcgen05.ld, write P = exp2((S - max) c) as packed fp16 back
//                into TMEM with tcgen05.st (two keys per 32-bit column);
//   O = P V    : tcgen05.mma with the A operand taken FROM TMEM (P) and B = V (Lk_pad x 64, smem, MN-major SW128: the rows TMA
//                writes for V[key][d] are already that layout), accumulator O (64 fp32 columns) in TMEM;
//   epilogue   : O / rowsum -> fp16 -> HBM.
// One CTA per (batch, head): K and V are staged once and reused by every 128-row query tile of that head.
#include "common.cuh"
#include "diffsound_b200.h"
#include <cuda_fp16.h>

namespace dsb {
namespace {
constexpr int TC_HD = 64;        // head dim
constexpr int TC_QM = 128;       // query rows per tile (UMMA M)
constexpr int TC_KMAX = 272;     // padded key capacity (multiple of 16; S needs TC_KMAX TMEM columns)
constexpr int TC_THREADS = 192;  // warps 0-3: softmax + epilogue (TMEM lane quadrant = warp), warp 4: loader + MMA issuer, warp 5: TMEM owner

__device__ __forceinline__ void tmem_st_32x32_x16(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]),
        "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// D[tmem] (+)= A[tmem] * B[smem]   (A operand from tensor memory)
__device__ __forceinline__ void umma_ts_f16(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
               ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
// instruction descriptor: fp16 operands, fp32 accumulate, A K-major, B major selectable (bit 16 = MN-major)
__device__ __forceinline__ constexpr uint32_t idesc_f16(int M, int N, bool b_mn_major) {
  return (1u << 4) | (b_mn_major ? (1u << 16) : 0u) | (static_cast<uint32_t>(N >> 3) << 17) | (static_cast<uint32_t>(M >> 4) << 24);
}
// 128-byte rows, SWIZZLE_128B: element (row r, 16-byte chunk c) lives at r*128 + ((c ^ (r & 7)) << 4)
__device__ __forceinline__ uint32_t sw128_off(int r, int chunk) { return r * 128 + ((chunk ^ (r & 7)) << 4); }

__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// Copies `rows` x 64 halves (row stride ld) into a SW128 tile; rows >= valid are zero-filled.  All 192 threads participate.
__device__ __forceinline__ void fill_tile(uint8_t* dst, const __half* src, long long ld, int rows, int valid) {
  for (int idx = threadIdx.x; idx < rows * 8; idx += TC_THREADS) {
    const int r = idx >> 3, c = idx & 7;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (r < valid) v = *reinterpret_cast<const uint4*>(src + (long long)r * ld + c * 8);
    *reinterpret_cast<uint4*>(dst + sw128_off(r, c)) = v;
  }
}

__global__ void __launch_bounds__(TC_THREADS, 1)
attention_tc_kernel(const __half* __restrict__ q, long long ldq, const __half* __restrict__ k, long long ldk, const __half* __restrict__ v,
                    long long ldv, __half* __restrict__ o, long long ldo, int Lq, int Lk, float scale_log2e) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;                          // 128 x 128 B
  uint8_t* sK = sQ + TC_QM * 128;              // 272 x 128 B
  uint8_t* sV = sK + TC_KMAX * 128;            // 272 x 128 B
  uint64_t* bar_s = reinterpret_cast<uint64_t*>(sV + TC_KMAX * 128);  // S ready
  uint64_t* bar_o = bar_s + 1;                                        // O ready
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bar_o + 1);

  const int h = blockIdx.x, b = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const __half* qb = q + (long long)b * Lq * ldq + h * TC_HD;
  const __half* kb = k + (long long)b * Lk * ldk + h * TC_HD;
  const __half* vb = v + (long long)b * Lk * ldv + h * TC_HD;
  __half* ob = o + (long long)b * Lq * ldo + h * TC_HD;
  const int kpad = (Lk + 15) & ~15;            // keys rounded to the UMMA K step
  const int n_hi = kpad > 256 ? 256 : kpad;    // first S instruction covers keys [0, n_hi)
  const int n_lo = kpad - n_hi;                // second one the remaining (multiple of 16) keys

  if (threadIdx.x == 0) {
    mbar_init(bar_s, 1);
    mbar_init(bar_o, 1);
    fence_barrier_init();
  }
  if (warp == 5) {
    tmem_alloc(tmem_ptr, 512);
    tmem_relinquish();
  }
  pdl_wait();
  pdl_trigger();
  fill_tile(sK, kb, ldk, kpad, Lk);
  fill_tile(sV, vb, ldv, kpad, Lk);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  const uint32_t tS = tmem_base;               // S: columns [0, 272)
  const uint32_t tP = tmem_base + 272;         // P: columns [272, 408), two fp16 keys per column
  const uint32_t tO = tmem_base + 416;         // O: columns [416, 480)

  const int n_qtiles = (Lq + TC_QM - 1) / TC_QM;
  uint32_t phase = 0;
  for (int qt = 0; qt < n_qtiles; ++qt) {
    const int q0 = qt * TC_QM;
    fill_tile(sQ, qb + (long long)q0 * ldq, ldq, TC_QM, Lq - q0);
    fence_proxy_async_smem();                  // generic-proxy smem writes -> visible to the tensor-core (async) proxy
    __syncthreads();
    if (warp == 4 && lane == 0) {
      // ---- S = Q K^T
      tc_fence_after();
      const uint64_t dq = make_sw128_kmajor_desc(smem_u32(sQ));
      const uint64_t dk = make_sw128_kmajor_desc(smem_u32(sK));
      const uint64_t dk2 = make_sw128_kmajor_desc(smem_u32(sK) + 256 * 128);
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        umma<false>(tS, dq + 2 * ks, dk + 2 * ks, idesc_f16(TC_QM, n_hi, false), ks != 0);
        if (n_lo > 0) umma<false>(tS + 256, dq + 2 * ks, dk2 + 2 * ks, idesc_f16(TC_QM, n_lo, false), ks != 0);
      }
      umma_commit(bar_s);
    }
    if (warp < 4) {
      // ---- softmax: thread = one query row (TMEM lane 32*warp + lane)
      mbar_wait(bar_s, phase);
      tc_fence_after();
      const uint32_t lane_off = static_cast<uint32_t>(warp * 32) << 16;
      float mx = -INFINITY;
      for (int c = 0; c < kpad; c += 32) {
        uint32_t sv[32];
        tmem_ld_32x32(tS + lane_off + c, sv);  // kpad is a multiple of 16: the last chunk may read 16 stale columns, masked below
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (c + j < Lk) mx = fmaxf(mx, __uint_as_float(sv[j]));
      }
      const float ms = mx * scale_log2e;
      float sum = 0.f;
      for (int c = 0; c < kpad; c += 32) {
        uint32_t sv[32];
        tmem_ld_32x32(tS + lane_off + c, sv);
        tmem_ld_wait();
        uint32_t pk[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float p0 = (c + 2 * j < Lk) ? ex2f(fmaf(__uint_as_float(sv[2 * j]), scale_log2e, -ms)) : 0.f;
          const float p1 = (c + 2 * j + 1 < Lk) ? ex2f(fmaf(__uint_as_float(sv[2 * j + 1]), scale_log2e, -ms)) : 0.f;
          sum += p0 + p1;
          __half2 hh = __floats2half2_rn(p0, p1);  // low half = even key
          pk[j] = *reinterpret_cast<uint32_t*>(&hh);
        }
        tmem_st_32x32_x16(tP + lane_off + (c >> 1), pk);
      }
      tmem_st_wait();
      tc_fence_before();
      // hand P to the MMA thread: named barrier over the 4 softmax warps + the MMA warp (160 threads)
      asm volatile("bar.sync 1, 160;" ::: "memory");
      // ---- epilogue of this tile
      mbar_wait(bar_o, phase);
      tc_fence_after();
      const float inv = 1.0f / sum;
      const int row = q0 + warp * 32 + lane;
#pragma unroll
      for (int c = 0; c < TC_HD; c += 32) {
        uint32_t ov[32];
        tmem_ld_32x32(tO + lane_off + c, ov);
        tmem_ld_wait();
        if (row < Lq) {
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            __half2 h0 = __floats2half2_rn(__uint_as_float(ov[j]) * inv, __uint_as_float(ov[j + 1]) * inv);
            __half2 h1 = __floats2half2_rn(__uint_as_float(ov[j + 2]) * inv, __uint_as_float(ov[j + 3]) * inv);
            __half2 h2 = __floats2half2_rn(__uint_as_float(ov[j + 4]) * inv, __uint_as_float(ov[j + 5]) * inv);
            __half2 h3 = __floats2half2_rn(__uint_as_float(ov[j + 6]) * inv, __uint_as_float(ov[j + 7]) * inv);
            uint4 u;
            u.x = *reinterpret_cast<uint32_t*>(&h0); u.y = *reinterpret_cast<uint32_t*>(&h1);
            u.z = *reinterpret_cast<uint32_t*>(&h2); u.w = *reinterpret_cast<uint32_t*>(&h3);
            *reinterpret_cast<uint4*>(ob + (long long)row * ldo + c + j) = u;
          }
        }
      }
      tc_fence_before();
    } else if (warp == 4) {
      asm volatile("bar.sync 1, 160;" ::: "memory");  // P is in TMEM
      if (lane == 0) {
        // ---- O = P V : A from TMEM (8 columns per 16-key step), B = V rows (MN-major SW128, 16 keys = 2048 B per step)
        tc_fence_after();
        const uint64_t dv = make_sw128_kmajor_desc(smem_u32(sV));  // same field values; the MN-major interpretation comes from idesc bit 16
        const int ksteps = kpad >> 4;
        for (int ks = 0; ks < ksteps; ++ks)
          umma_ts_f16(tO, tP + ks * 8, dv + (uint64_t)(ks * 128), idesc_f16(TC_QM, TC_HD, true), ks != 0);
        umma_commit(bar_o);
      }
    }
    phase ^= 1;
    __syncthreads();  // sQ is refilled and S / P / O are overwritten by the next tile
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 5) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}
}  // namespace
}  // namespace dsb
using namespace dsb;

extern "C" int dsb_attention_tc(const void* q, long long ldq, const void* k, long long ldk, const void* v, long long ldv, void* o, long long ldo,
                                int B, int H, int Lq, int Lk, float scale, void* stream) {
  DSB_REQUIRE(B > 0 && H > 0 && Lq > 0 && Lk > 0 && Lk <= TC_KMAX, "dsb_attention_tc: need 0 < Lk <= %d", TC_KMAX);
  DSB_REQUIRE(ldq % 8 == 0 && ldk % 8 == 0 && ldv % 8 == 0 && ldo % 8 == 0, "dsb_attention_tc: leading dimensions must be multiples of 8 halves");
  DSB_REQUIRE(((reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(k) | reinterpret_cast<uintptr_t>(v) | reinterpret_cast<uintptr_t>(o)) & 15) == 0,
              "dsb_attention_tc: pointers must be 16-byte aligned");
  const int smem = (TC_QM + 2 * TC_KMAX) * 128 + 1024 + 64;
  static bool attr_done = false;
  if (!attr_done) {
    DSB_CHECK_CUDA(cudaFuncSetAttribute(attention_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_done = true;
  }
  DSB_CHECK_CUDA(launch_pdl(attention_tc_kernel, dim3(H, B), dim3(TC_THREADS), smem, (cudaStream_t)stream, (const __half*)q, ldq, (const __half*)k, ldk,
                            (const __half*)v, ldv, (__half*)o, ldo, Lq, Lk, scale * 1.4426950408889634f));
  return 0;
}
