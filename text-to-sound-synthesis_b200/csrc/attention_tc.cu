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

// ================================================================================================================================
// Pipelined version.  Work list = (batch, head, 128-row query tile), flattened head-major so that consecutive tiles share K/V; every
// CTA takes a contiguous, equally sized slice of it.  Roles (14 warps):
//   warps 0-7  softmax: warp w owns TMEM lanes 32*(w%4).., half (w/4) of the key columns; row max / row sum exchanged through smem
//   warps 8-11 epilogue: O / rowsum -> fp16 -> HBM
//   warp 12    TMA producer: K/V (double-buffered per (batch, head)) and Q tiles (double-buffered)
//   warp 13    tcgen05.mma issuer (+ TMEM owner): S(t) ... [softmax(t)] ... P.V(t), S(t+1) back to back
// TMEM: S columns [0,272), P (packed fp16) [272,408), O [416,480).
namespace dsb {
namespace {
constexpr int T2_THREADS = 512;   // 16 warps: 0-7 softmax, 8-11 epilogue, 12 producer, 13 MMA issuer, 14-15 remainder rows
constexpr int T2_KV_BYTES = TC_KMAX * 128;   // one K or V buffer
struct T2Params {
  int B, H, Lq, Lk, kpad, n_qt, total_tiles, box_rows, n_box;
  int rem_rows;         // > 0: the last (Lq % 128 <= 16) query rows of every head are computed by the remainder warp, not by a tensor tile
  long long ldo, ldq;
  __half* o;
  const __half* q;
  float scale_log2e;
};

// ---- mma.sync helpers for the remainder warp (same fragments as attention_f16.cu, addresses follow the SW128 tile layout)
__device__ __forceinline__ void t2_ldsm_x4(uint32_t (&r)[4], const uint8_t* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(smem_u32(p)));
}
__device__ __forceinline__ void t2_ldsm_x4_t(uint32_t (&r)[4], const uint8_t* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(smem_u32(p)));
}
__device__ __forceinline__ void t2_mma(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t t2_pack(float a, float b) {
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

// One warp: rows [row0, row0 + 16) of one head (rows >= Lq masked) against the K / V tiles in shared memory, flash-style over
// 64-key chunks, written straight to HBM.  Runs concurrently with the tensor-core tiles of the same head.
__device__ __forceinline__ void remainder_rows(const T2Params& p, const uint8_t* sKb, const uint8_t* sVb, int b, int h, int row0, int lane) {
  const int g = lane >> 2, t = lane & 3;
  const __half* qb = p.q + (long long)b * p.Lq * p.ldq + h * TC_HD;
  uint32_t a[4][4];
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) {
    const int ra = row0 + g, rb = row0 + g + 8, c = ks * 16 + 2 * t;
    a[ks][0] = ra < p.Lq ? *reinterpret_cast<const uint32_t*>(qb + (long long)ra * p.ldq + c) : 0u;
    a[ks][1] = rb < p.Lq ? *reinterpret_cast<const uint32_t*>(qb + (long long)rb * p.ldq + c) : 0u;
    a[ks][2] = ra < p.Lq ? *reinterpret_cast<const uint32_t*>(qb + (long long)ra * p.ldq + c + 8) : 0u;
    a[ks][3] = rb < p.Lq ? *reinterpret_cast<const uint32_t*>(qb + (long long)rb * p.ldq + c + 8) : 0u;
  }
  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
  float oacc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i) oacc[i][0] = oacc[i][1] = oacc[i][2] = oacc[i][3] = 0.f;
  for (int kc = 0; kc * 64 < p.kpad; ++kc) {
    const int keys_left = p.Lk - kc * 64;
    const int n_live = keys_left >= 64 ? 8 : (keys_left + 7) >> 3;
    float s[8][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
      if (nt >= n_live) continue;
      const int row = kc * 64 + nt * 8 + (lane & 7);
      uint32_t kf[4];
      t2_ldsm_x4(kf, sKb + sw128_off(row, lane >> 3));
      t2_mma(s[nt], a[0], kf[0], kf[1]);
      t2_mma(s[nt], a[1], kf[2], kf[3]);
      t2_ldsm_x4(kf, sKb + sw128_off(row, 4 + (lane >> 3)));
      t2_mma(s[nt], a[2], kf[0], kf[1]);
      t2_mma(s[nt], a[3], kf[2], kf[3]);
    }
    float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const int key = kc * 64 + nt * 8 + 2 * t;
      if (key >= p.Lk) { s[nt][0] = -INFINITY; s[nt][2] = -INFINITY; }
      if (key + 1 >= p.Lk) { s[nt][1] = -INFINITY; s[nt][3] = -INFINITY; }
      mx0 = fmaxf(mx0, fmaxf(s[nt][0], s[nt][1]));
      mx1 = fmaxf(mx1, fmaxf(s[nt][2], s[nt][3]));
    }
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1)); mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1)); mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    const float mn0 = fmaxf(m0, mx0), mn1 = fmaxf(m1, mx1);
    const float c0 = ex2f((m0 - mn0) * p.scale_log2e), c1 = ex2f((m1 - mn1) * p.scale_log2e);
    const float ms0 = mn0 * p.scale_log2e, ms1 = mn1 * p.scale_log2e;
    m0 = mn0; m1 = mn1;
    l0 *= c0; l1 *= c1;
#pragma unroll
    for (int nd = 0; nd < 8; ++nd) { oacc[nd][0] *= c0; oacc[nd][1] *= c0; oacc[nd][2] *= c1; oacc[nd][3] *= c1; }
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      s[nt][0] = ex2f(fmaf(s[nt][0], p.scale_log2e, -ms0)); s[nt][1] = ex2f(fmaf(s[nt][1], p.scale_log2e, -ms0));
      s[nt][2] = ex2f(fmaf(s[nt][2], p.scale_log2e, -ms1)); s[nt][3] = ex2f(fmaf(s[nt][3], p.scale_log2e, -ms1));
      l0 += s[nt][0] + s[nt][1];
      l1 += s[nt][2] + s[nt][3];
    }
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      if (kk * 2 >= n_live) continue;
      const uint32_t pa[4] = {t2_pack(s[2 * kk][0], s[2 * kk][1]), t2_pack(s[2 * kk][2], s[2 * kk][3]),
                              t2_pack(s[2 * kk + 1][0], s[2 * kk + 1][1]), t2_pack(s[2 * kk + 1][2], s[2 * kk + 1][3])};
      const int row = kc * 64 + kk * 16 + (lane & 7) + 8 * ((lane >> 3) & 1);
#pragma unroll
      for (int np = 0; np < 4; ++np) {
        uint32_t vf[4];
        t2_ldsm_x4_t(vf, sVb + sw128_off(row, 2 * np + (lane >> 4)));
        t2_mma(oacc[2 * np], pa, vf[0], vf[1]);
        t2_mma(oacc[2 * np + 1], pa, vf[2], vf[3]);
      }
    }
  }
  l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
  const float i0 = 1.0f / l0, i1 = 1.0f / l1;
  const int ra = row0 + g, rb = row0 + g + 8;
  __half* ob = p.o + (long long)b * p.Lq * p.ldo + h * TC_HD;
#pragma unroll
  for (int nd = 0; nd < 8; ++nd) {
    if (ra < p.Lq) *reinterpret_cast<__half2*>(ob + (long long)ra * p.ldo + nd * 8 + 2 * t) = __floats2half2_rn(oacc[nd][0] * i0, oacc[nd][1] * i0);
    if (rb < p.Lq) *reinterpret_cast<__half2*>(ob + (long long)rb * p.ldo + nd * 8 + 2 * t) = __floats2half2_rn(oacc[nd][2] * i1, oacc[nd][3] * i1);
  }
}

__global__ void __launch_bounds__(T2_THREADS, 1)
attention_tc2_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k, const __grid_constant__ CUtensorMap map_v,
                     const __grid_constant__ T2Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;                               // 2 x 16 KB
  uint8_t* sK = sQ + 2 * TC_QM * 128;               // 2 x 34 KB
  uint8_t* sV = sK + 2 * T2_KV_BYTES;               // 2 x 34 KB
  float* s_max = reinterpret_cast<float*>(sV + 2 * T2_KV_BYTES);  // [2 halves][128 rows]
  float* s_sum = s_max + 512;                       // [2 tile parities][2 halves][128 rows]   (s_max: [2 tile parities][2 halves][128])
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_sum + 512);
  uint64_t* kv_full = bars;        // [2]
  uint64_t* kv_empty = bars + 2;   // [2]
  uint64_t* q_full = bars + 4;     // [2]
  uint64_t* q_empty = bars + 6;    // [2]
  uint64_t* s_full = bars + 8;
  uint64_t* p_full = bars + 9;
  uint64_t* o_full = bars + 10;
  uint64_t* o_empty = bars + 11;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 12);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g0 = (int)((long long)p.total_tiles * blockIdx.x / gridDim.x);
  const int g1 = (int)((long long)p.total_tiles * (blockIdx.x + 1) / gridDim.x);

  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&kv_full[i], 1); mbar_init(&kv_empty[i], p.rem_rows > 0 ? 2 : 1);  // MMA commit (+ the remainder warp)
      mbar_init(&q_full[i], 1); mbar_init(&q_empty[i], 1);
    }
    mbar_init(s_full, 1); mbar_init(p_full, 8); mbar_init(o_full, 1); mbar_init(o_empty, 4);
    fence_barrier_init();
    prefetch_tmap(&map_q); prefetch_tmap(&map_k); prefetch_tmap(&map_v);
  }
  if (warp == 13) {
    tmem_alloc(tmem_ptr, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  const uint32_t tS = tmem_base, tP = tmem_base + 272, tO = tmem_base + 416;
  pdl_wait();
  pdl_trigger();

  if (warp == 12) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int prev_u = -1, kv_n = 0;
      for (int g = g0, t = 0; g < g1; ++g, ++t) {
        const int u = g / p.n_qt, qt = g - u * p.n_qt;
        const int b = u / p.H, h = u - b * p.H;
        if (u != prev_u) {
          const int kb = kv_n & 1;
          mbar_wait(&kv_empty[kb], ((kv_n >> 1) & 1) ^ 1);
          mbar_arrive_expect_tx(&kv_full[kb], 2 * p.kpad * 128);
          for (int bx = 0; bx < p.n_box; ++bx) {
            tma_load_3d(&map_k, &kv_full[kb], sK + kb * T2_KV_BYTES + bx * p.box_rows * 128, h * TC_HD, b * p.Lk + bx * p.box_rows, 0);
            tma_load_3d(&map_v, &kv_full[kb], sV + kb * T2_KV_BYTES + bx * p.box_rows * 128, h * TC_HD, b * p.Lk + bx * p.box_rows, 0);
          }
          ++kv_n;
          prev_u = u;
        }
        const int qb = t & 1;
        mbar_wait(&q_empty[qb], ((t >> 1) & 1) ^ 1);
        mbar_arrive_expect_tx(&q_full[qb], TC_QM * 128);
        tma_load_3d(&map_q, &q_full[qb], sQ + qb * TC_QM * 128, h * TC_HD, b * p.Lq + qt * TC_QM, 0);
      }
    }
  } else if (warp == 13) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      const int n_hi = p.kpad > 256 ? 256 : p.kpad, n_lo = p.kpad - n_hi;
      const uint32_t id_hi = idesc_f16(TC_QM, n_hi, false), id_lo = idesc_f16(TC_QM, n_lo > 0 ? n_lo : 16, false);
      const uint32_t id_pv = idesc_f16(TC_QM, TC_HD, true);
      const int ksteps = p.kpad >> 4;
      int prev_u = -1, kv_n = 0, kb = 0;
      // issue S for local tile t (global tile g): waits for its K/V and Q, then 4 (x2) MMAs; returns the K/V buffer it used
      auto issue_s = [&](int g, int t) {
        const int u = g / p.n_qt;
        if (u != prev_u) {
          kb = kv_n & 1;
          mbar_wait(&kv_full[kb], (kv_n >> 1) & 1);
          ++kv_n;
          prev_u = u;
        }
        const int qb = t & 1;
        mbar_wait(&q_full[qb], (t >> 1) & 1);
        tc_fence_after();
        const uint64_t dq = make_sw128_kmajor_desc(smem_u32(sQ + qb * TC_QM * 128));
        const uint64_t dk = make_sw128_kmajor_desc(smem_u32(sK + kb * T2_KV_BYTES));
        const uint64_t dk2 = make_sw128_kmajor_desc(smem_u32(sK + kb * T2_KV_BYTES) + 256 * 128);
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          umma<false>(tS, dq + 2 * ks, dk + 2 * ks, id_hi, ks != 0);
          if (n_lo > 0) umma<false>(tS + 256, dq + 2 * ks, dk2 + 2 * ks, id_lo, ks != 0);
        }
        umma_commit(s_full);
        umma_commit(&q_empty[qb]);
        return kb;
      };
      int kb_cur = (g0 < g1) ? issue_s(g0, 0) : 0;
      for (int g = g0, t = 0; g < g1; ++g, ++t) {
        const int u = g / p.n_qt;
        // P(t) is in TMEM and softmax(t) no longer reads S: the next tile's S can go first, P.V(t) then runs under softmax(t+1)
        mbar_wait(p_full, t & 1);
        int kb_next = kb_cur;
        if (g + 1 < g1) kb_next = issue_s(g + 1, t + 1);
        mbar_wait(o_empty, (t & 1) ^ 1);  // epilogue(t-1) has read O
        tc_fence_after();
        const uint64_t dv = make_sw128_kmajor_desc(smem_u32(sV + kb_cur * T2_KV_BYTES));
        for (int ks = 0; ks < ksteps; ++ks) umma_ts_f16(tO, tP + ks * 8, dv + (uint64_t)(ks * 128), id_pv, ks != 0);
        umma_commit(o_full);
        const bool last_of_unit = (g + 1 == g1) || ((g + 1) / p.n_qt != u);
        if (last_of_unit) umma_commit(&kv_empty[kb_cur]);
        kb_cur = kb_next;
      }
    }
  } else if (warp >= 14) {
    // ------------------------------------------------------------------ remainder rows (Lq % 128 <= 16) of every head whose last tile is
    // ours.  Two warps, one per K/V buffer (warp 14 serves buffer 0 = even local heads, warp 15 buffer 1), so each observes every phase
    // of "its" kv_full barrier in order and has two heads' worth of tensor time to finish one remainder.
    if (p.rem_rows > 0) {
      const int mine = warp - 14;
      int prev_u = -1, kv_n = 0;
      for (int g = g0; g < g1; ++g) {
        const int u = g / p.n_qt, qt = g - u * p.n_qt;
        bool first_of_unit = false;
        if (u != prev_u) { first_of_unit = true; prev_u = u; ++kv_n; }
        const int n = kv_n - 1, kb = n & 1;
        if (kb != mine) continue;
        if (first_of_unit) mbar_wait(&kv_full[kb], (n >> 1) & 1);
        const bool last_of_unit = (g + 1 == g1) || ((g + 1) / p.n_qt != u);
        if (qt == p.n_qt - 1) remainder_rows(p, sK + kb * T2_KV_BYTES, sV + kb * T2_KV_BYTES, u / p.H, u % p.H, p.n_qt * TC_QM, lane);
        if (last_of_unit) {
          __syncwarp();
          if (lane == 0) mbar_arrive(&kv_empty[kb]);
        }
      }
    }
  } else if (warp < 8) {
    // ------------------------------------------------------------------ softmax: 2 warps per 32-row slab, each half of the key columns
    const int quad = warp & 3, half = warp >> 2;
    const uint32_t lane_off = static_cast<uint32_t>(quad * 32) << 16;
    const int row_in_tile = quad * 32 + lane;
    const int csplit = (((p.kpad + 31) >> 5) + 1) / 2 * 32;  // columns [0, csplit) -> half 0, [csplit, kpad) -> half 1
    const int c_begin = half == 0 ? 0 : csplit, c_end = half == 0 ? (csplit < p.kpad ? csplit : p.kpad) : p.kpad;
    for (int g = g0, t = 0; g < g1; ++g, ++t) {
      const int qt = g % p.n_qt;
      // a 32-row slab that lies entirely beyond Lq (most of the last query tile of a 265-row head) does no exp work at all:
      // its P rows stay whatever they were, the rows are independent in the MMA and are never stored
      const bool live = qt * TC_QM + quad * 32 < p.Lq;
      mbar_wait(s_full, t & 1);
      tc_fence_after();
      float mx = -INFINITY, sum = 0.f;
      float* smx = s_max + (t & 1) * 256;  // double-buffered by tile parity: no second barrier needed
      if (live) {
        for (int c = c_begin; c < c_end; c += 64) {
          uint32_t sa[32], sb[32];
          const bool two = c + 32 < c_end;
          tmem_ld_32x32(tS + lane_off + c, sa);
          if (two) tmem_ld_32x32(tS + lane_off + c + 32, sb);
          tmem_ld_wait();
          if (c + 32 <= p.Lk) {
#pragma unroll
            for (int j = 0; j < 32; ++j) mx = fmaxf(mx, __uint_as_float(sa[j]));
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (c + j < p.Lk) mx = fmaxf(mx, __uint_as_float(sa[j]));
          }
          if (two) {
            if (c + 64 <= p.Lk) {
#pragma unroll
              for (int j = 0; j < 32; ++j) mx = fmaxf(mx, __uint_as_float(sb[j]));
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (c + 32 + j < p.Lk) mx = fmaxf(mx, __uint_as_float(sb[j]));
            }
          }
        }
      }
      smx[half * 128 + row_in_tile] = mx;
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (t > 0) {  // P.V(t-1) still reads the P region this pass is about to overwrite
        mbar_wait(o_full, (t - 1) & 1);
        tc_fence_after();
      }
      if (live) {
        mx = fmaxf(mx, smx[(half ^ 1) * 128 + row_in_tile]);
        const float ms = mx * p.scale_log2e;
        auto do_chunk = [&](const uint32_t (&sv)[32], int c) {
          uint32_t pk[16];
          if (c + 32 <= p.Lk) {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const float p0 = ex2f(fmaf(__uint_as_float(sv[2 * j]), p.scale_log2e, -ms));
              const float p1 = ex2f(fmaf(__uint_as_float(sv[2 * j + 1]), p.scale_log2e, -ms));
              sum += p0 + p1;
              __half2 hh = __floats2half2_rn(p0, p1);
              pk[j] = *reinterpret_cast<uint32_t*>(&hh);
            }
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const float p0 = (c + 2 * j < p.Lk) ? ex2f(fmaf(__uint_as_float(sv[2 * j]), p.scale_log2e, -ms)) : 0.f;
              const float p1 = (c + 2 * j + 1 < p.Lk) ? ex2f(fmaf(__uint_as_float(sv[2 * j + 1]), p.scale_log2e, -ms)) : 0.f;
              sum += p0 + p1;
              __half2 hh = __floats2half2_rn(p0, p1);
              pk[j] = *reinterpret_cast<uint32_t*>(&hh);
            }
          }
          tmem_st_32x32_x16(tP + lane_off + (c >> 1), pk);
        };
        for (int c = c_begin; c < c_end; c += 64) {
          uint32_t sa[32], sb[32];
          const bool two = c + 32 < c_end;
          tmem_ld_32x32(tS + lane_off + c, sa);
          if (two) tmem_ld_32x32(tS + lane_off + c + 32, sb);
          tmem_ld_wait();
          do_chunk(sa, c);
          if (two) do_chunk(sb, c + 32);
        }
        tmem_st_wait();
      }
      s_sum[((t & 1) * 2 + half) * 128 + row_in_tile] = sum;
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full);
    }
  } else {
    // ------------------------------------------------------------------ epilogue (warps 8-11)
    const int quad = warp & 3;
    const uint32_t lane_off = static_cast<uint32_t>(quad * 32) << 16;
    const int row_in_tile = quad * 32 + lane;
    for (int g = g0, t = 0; g < g1; ++g, ++t) {
      const int u = g / p.n_qt, qt = g - u * p.n_qt;
      const int b = u / p.H, h = u - b * p.H;
      mbar_wait(o_full, t & 1);
      tc_fence_after();
      const float inv = 1.0f / (s_sum[((t & 1) * 2) * 128 + row_in_tile] + s_sum[((t & 1) * 2 + 1) * 128 + row_in_tile]);
      const int row = qt * TC_QM + row_in_tile;
      __half* orow = p.o + ((long long)b * p.Lq + row) * p.ldo + h * TC_HD;
#pragma unroll
      for (int c = 0; c < TC_HD; c += 32) {
        uint32_t ov[32];
        tmem_ld_32x32(tO + lane_off + c, ov);
        tmem_ld_wait();
        if (row < p.Lq) {
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            __half2 h0 = __floats2half2_rn(__uint_as_float(ov[j]) * inv, __uint_as_float(ov[j + 1]) * inv);
            __half2 h1 = __floats2half2_rn(__uint_as_float(ov[j + 2]) * inv, __uint_as_float(ov[j + 3]) * inv);
            __half2 h2 = __floats2half2_rn(__uint_as_float(ov[j + 4]) * inv, __uint_as_float(ov[j + 5]) * inv);
            __half2 h3 = __floats2half2_rn(__uint_as_float(ov[j + 6]) * inv, __uint_as_float(ov[j + 7]) * inv);
            uint4 uu;
            uu.x = *reinterpret_cast<uint32_t*>(&h0); uu.y = *reinterpret_cast<uint32_t*>(&h1);
            uu.z = *reinterpret_cast<uint32_t*>(&h2); uu.w = *reinterpret_cast<uint32_t*>(&h3);
            *reinterpret_cast<uint4*>(orow + c + j) = uu;
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(o_empty);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 13) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}
}  // namespace
}  // namespace dsb

extern "C" int dsb_attention_tc2(const void* q, long long ldq, const void* k, long long ldk, const void* v, long long ldv, void* o, long long ldo,
                                 int B, int H, int Lq, int Lk, float scale, void* stream) {
  using namespace dsb;
  DSB_REQUIRE(B > 0 && H > 0 && Lq > 0 && Lk > 0 && Lk <= TC_KMAX, "dsb_attention_tc2: need 0 < Lk <= %d", TC_KMAX);
  DSB_REQUIRE(ldo % 8 == 0 && (reinterpret_cast<uintptr_t>(o) & 15) == 0, "dsb_attention_tc2: o must be 16-byte aligned with ldo %% 8 == 0");
  T2Params p{};
  p.B = B; p.H = H; p.Lq = Lq; p.Lk = Lk;
  p.kpad = (Lk + 15) & ~15;
  const int rem = Lq % TC_QM;
  if (Lq > TC_QM && rem > 0 && rem <= 16) {  // e.g. 265 = 2 x 128 + 9: the 9 rows go to the remainder warp, not to a third tensor tile
    p.n_qt = Lq / TC_QM;
    p.rem_rows = rem;
  } else {
    p.n_qt = (Lq + TC_QM - 1) / TC_QM;
    p.rem_rows = 0;
  }
  p.total_tiles = B * H * p.n_qt;
  p.q = (const __half*)q; p.ldq = ldq;
  p.n_box = p.kpad > 256 ? 2 : 1;
  p.box_rows = p.kpad / p.n_box;
  DSB_REQUIRE(p.box_rows % 8 == 0, "dsb_attention_tc2: internal box size");
  p.ldo = ldo; p.o = (__half*)o;
  p.scale_log2e = scale * 1.4426950408889634f;
  CUtensorMap mq, mk, mv;
  if (make_operand_map(&mq, q, DSB_DTYPE_F16, (long long)H * TC_HD, (long long)B * Lq, 1, ldq, 0, TC_QM)) return 3;
  if (make_operand_map(&mk, k, DSB_DTYPE_F16, (long long)H * TC_HD, (long long)B * Lk, 1, ldk, 0, p.box_rows)) return 3;
  if (make_operand_map(&mv, v, DSB_DTYPE_F16, (long long)H * TC_HD, (long long)B * Lk, 1, ldv, 0, p.box_rows)) return 3;
  const int smem = 2 * TC_QM * 128 + 4 * T2_KV_BYTES + (512 + 512) * 4 + 16 * 8 + 1024;
  static bool attr_done = false;
  if (!attr_done) {
    DSB_CHECK_CUDA(cudaFuncSetAttribute(attention_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_done = true;
  }
  int grid = sm_count();
  if (grid > p.total_tiles) grid = p.total_tiles;
  DSB_CHECK_CUDA(launch_pdl(attention_tc2_kernel, dim3(grid), dim3(T2_THREADS), smem, (cudaStream_t)stream, mq, mk, mv, p));
  return 0;
}
