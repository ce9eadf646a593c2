// fp16-operand variant of the attention core (same math as attention.cu, reference transformer_utils.py:48-54, :99-105):
// Q, K, V arrive as fp16 straight from the QKV / cross-KV GEMM epilogues, scores and the online softmax stay fp32 in
// registers, P is re-packed to fp16 in registers (FA2 register reuse) for the P.V product; output fp16 (A operand of the
// projection GEMM).  fp16 carries the same 11-bit significand as TF32, at twice the mma.sync rate and half the smem bytes.
// Still the warp-level mma.sync path (m16n8k16) -- the tcgen05/TMEM version is the follow-up.
#include "common.cuh"
#include "diffsound_b200.h"
#include <cuda_fp16.h>

namespace dsb {
namespace {
constexpr int HD = 64, QT = 64, KT = 64, LDS = 72;  // LDS: padded smem row stride in halves (144 B -> conflict-free ldmatrix)

__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], const __half* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(smem_u32(p)));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], const __half* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(smem_u32(p)));
}
__device__ __forceinline__ void mma_f16(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ float ex2(float x) {  // 2^x, single MUFU (inputs are <= 0 here; flush-to-zero underflow is exact enough)
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

// asynchronous 64 x 64-half tile copy (global -> padded smem) with cp.async; rows past nrows_valid are zero-filled
__device__ __forceinline__ void load_tile_async(__half* dst, const __half* src, long long ld, int row0, int nrows_valid) {
#pragma unroll
  for (int it = 0; it < 4; ++it) {
    const int idx = threadIdx.x + it * 128;
    const int r = idx >> 3, c8 = (idx & 7) * 8;
    const bool ok = row0 + r < nrows_valid;
    const __half* g = src + (long long)(ok ? row0 + r : 0) * ld + c8;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(dst + r * LDS + c8)), "l"(g), "r"(ok ? 16 : 0) : "memory");
  }
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__global__ void __launch_bounds__(128, 4)
attention_f16_kernel(const __half* __restrict__ q, long long ldq, const __half* __restrict__ k, long long ldk, const __half* __restrict__ v,
                     long long ldv, void* __restrict__ o, long long ldo, int Lq, int Lk, float scale_log2e, int flags) {
  __shared__ __align__(16) __half Qs[QT * LDS];
  __shared__ __align__(16) __half Ks[2][KT * LDS];
  __shared__ __align__(16) __half Vs[2][KT * LDS];
  const int qt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const bool causal = (flags & DSB_ATTN_CAUSAL) != 0;
  const __half* qb = q + (long long)b * Lq * ldq + h * HD;
  const __half* kb = k + (long long)b * Lk * ldk + h * HD;
  const __half* vb = v + (long long)b * Lk * ldv + h * HD;
  pdl_wait();
  pdl_trigger();

  load_tile_async(Qs, qb, ldq, qt * QT, Lq);
  load_tile_async(Ks[0], kb, ldk, 0, Lk);
  load_tile_async(Vs[0], vb, ldv, 0, Lk);
  cp_async_commit();
  uint32_t a[4][4];  // Q fragments: 4 k-steps of 16 head dims
  const bool warp_live = qt * QT + warp * 16 < Lq;  // warps whose 16 query rows are all out of range only help with the tile copies
  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
  float oacc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i) oacc[i][0] = oacc[i][1] = oacc[i][2] = oacc[i][3] = 0.f;

  const int nchunks = (Lk + KT - 1) / KT;
  for (int kc = 0; kc < nchunks; ++kc) {
    const __half* Kc = Ks[kc & 1];
    const __half* Vc = Vs[kc & 1];
    if (kc + 1 < nchunks) {  // prefetch the next K/V chunk into the other buffer while this one is consumed
      load_tile_async(Ks[(kc + 1) & 1], kb, ldk, (kc + 1) * KT, Lk);
      load_tile_async(Vs[(kc + 1) & 1], vb, ldv, (kc + 1) * KT, Lk);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    if (kc == 0) {
      const int row = warp * 16 + (lane & 7) + 8 * ((lane >> 3) & 1);
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) ldsm_x4(a[ks], Qs + row * LDS + ks * 16 + 8 * (lane >> 4));
    }

    // key tiles (8 keys) / key steps (16 keys) that lie entirely beyond Lk are skipped (warp-uniform): the last chunk of a
    // 265-key sequence holds 9 keys, of a 77-key sequence 13
    const int keys_left = Lk - kc * KT;
    const int n_live = keys_left >= KT ? 8 : (keys_left + 7) >> 3;
    float s[8][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
      if (!warp_live || nt >= n_live) continue;
      uint32_t kf[4];
      const __half* kp = Kc + (nt * 8 + (lane & 7)) * LDS + 8 * (lane >> 3);
      ldsm_x4(kf, kp);            // head dims 0..31  -> (b0,b1) of k-step 0 and 1
      mma_f16(s[nt], a[0], kf[0], kf[1]);
      mma_f16(s[nt], a[1], kf[2], kf[3]);
      ldsm_x4(kf, kp + 32);       // head dims 32..63 -> k-steps 2 and 3
      mma_f16(s[nt], a[2], kf[0], kf[1]);
      mma_f16(s[nt], a[3], kf[2], kf[3]);
    }
    float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const int key = kc * KT + nt * 8 + 2 * t;
      if (key >= Lk) { s[nt][0] = -INFINITY; s[nt][2] = -INFINITY; }
      if (key + 1 >= Lk) { s[nt][1] = -INFINITY; s[nt][3] = -INFINITY; }
      if (causal) {  // CLIP text tower: a query attends to keys at or before its own position (build_attention_mask, clip/model.py)
        const int qa = qt * QT + warp * 16 + g, qb_ = qa + 8;
        if (key > qa) s[nt][0] = -INFINITY;
        if (key + 1 > qa) s[nt][1] = -INFINITY;
        if (key > qb_) s[nt][2] = -INFINITY;
        if (key + 1 > qb_) s[nt][3] = -INFINITY;
      }
      mx0 = fmaxf(mx0, fmaxf(s[nt][0], s[nt][1]));
      mx1 = fmaxf(mx1, fmaxf(s[nt][2], s[nt][3]));
    }
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1)); mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1)); mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    const float mn0 = fmaxf(m0, mx0), mn1 = fmaxf(m1, mx1);
    const float c0 = ex2((m0 - mn0) * scale_log2e), c1 = ex2((m1 - mn1) * scale_log2e);
    const float ms0 = mn0 * scale_log2e, ms1 = mn1 * scale_log2e;
    m0 = mn0; m1 = mn1;
    l0 *= c0; l1 *= c1;
#pragma unroll
    for (int nd = 0; nd < 8; ++nd) { oacc[nd][0] *= c0; oacc[nd][1] *= c0; oacc[nd][2] *= c1; oacc[nd][3] *= c1; }
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      s[nt][0] = ex2(fmaf(s[nt][0], scale_log2e, -ms0)); s[nt][1] = ex2(fmaf(s[nt][1], scale_log2e, -ms0));
      s[nt][2] = ex2(fmaf(s[nt][2], scale_log2e, -ms1)); s[nt][3] = ex2(fmaf(s[nt][3], scale_log2e, -ms1));
      l0 += s[nt][0] + s[nt][1];
      l1 += s[nt][2] + s[nt][3];
    }
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {  // 16 keys per step
      if (!warp_live || kk * 2 >= n_live) continue;  // score tiles 2kk and 2kk+1 re-packed as the fp16 A fragment
      const uint32_t pa[4] = {pack_h2(s[2 * kk][0], s[2 * kk][1]), pack_h2(s[2 * kk][2], s[2 * kk][3]),
                              pack_h2(s[2 * kk + 1][0], s[2 * kk + 1][1]), pack_h2(s[2 * kk + 1][2], s[2 * kk + 1][3])};
      const __half* vp = Vc + (kk * 16 + (lane & 7) + 8 * ((lane >> 3) & 1)) * LDS + 8 * (lane >> 4);
#pragma unroll
      for (int np = 0; np < 4; ++np) {  // two 8-wide head-dim tiles per ldmatrix.x4.trans
        uint32_t vf[4];
        ldsm_x4_t(vf, vp + np * 16);
        mma_f16(oacc[2 * np], pa, vf[0], vf[1]);
        mma_f16(oacc[2 * np + 1], pa, vf[2], vf[3]);
      }
    }
    __syncthreads();  // everyone is done with this buffer before the prefetch two iterations ahead overwrites it
  }
  l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
  const float i0 = 1.0f / l0, i1 = 1.0f / l1;
  const int ra = qt * QT + warp * 16 + g, rb = ra + 8;
  if (flags & DSB_GEMM_OUT_F16) {
    __half* oh = reinterpret_cast<__half*>(o) + (long long)b * Lq * ldo + h * HD;
#pragma unroll
    for (int nd = 0; nd < 8; ++nd) {
      if (ra < Lq) *reinterpret_cast<__half2*>(oh + (long long)ra * ldo + nd * 8 + 2 * t) = __floats2half2_rn(oacc[nd][0] * i0, oacc[nd][1] * i0);
      if (rb < Lq) *reinterpret_cast<__half2*>(oh + (long long)rb * ldo + nd * 8 + 2 * t) = __floats2half2_rn(oacc[nd][2] * i1, oacc[nd][3] * i1);
    }
  } else {
    float* of = reinterpret_cast<float*>(o) + (long long)b * Lq * ldo + h * HD;
#pragma unroll
    for (int nd = 0; nd < 8; ++nd) {
      if (ra < Lq) *reinterpret_cast<float2*>(of + (long long)ra * ldo + nd * 8 + 2 * t) = make_float2(oacc[nd][0] * i0, oacc[nd][1] * i0);
      if (rb < Lq) *reinterpret_cast<float2*>(of + (long long)rb * ldo + nd * 8 + 2 * t) = make_float2(oacc[nd][2] * i1, oacc[nd][3] * i1);
    }
  }
}
}  // namespace
}  // namespace dsb
using namespace dsb;

extern "C" int dsb_attention_f16(const void* q, long long ldq, const void* k, long long ldk, const void* v, long long ldv, void* o, long long ldo,
                                 int B, int H, int Lq, int Lk, float scale, int flags, void* stream) {
  DSB_REQUIRE(B > 0 && H > 0 && Lq > 0 && Lk > 0, "dsb_attention_f16: bad shape");
  DSB_REQUIRE(ldq % 8 == 0 && ldk % 8 == 0 && ldv % 8 == 0 && ldo % 2 == 0, "dsb_attention_f16: ldq/ldk/ldv must be multiples of 8 halves");
  DSB_REQUIRE(((reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(k) | reinterpret_cast<uintptr_t>(v)) & 15) == 0,
              "dsb_attention_f16: q/k/v must be 16-byte aligned");
  dim3 grid((Lq + QT - 1) / QT, H, B);
  DSB_CHECK_CUDA(launch_pdl(attention_f16_kernel, grid, dim3(128), 0, (cudaStream_t)stream, (const __half*)q, ldq, (const __half*)k, ldk,
                            (const __half*)v, ldv, o, ldo, Lq, Lk, scale * 1.4426950408889634f, flags));
  return 0;
}
