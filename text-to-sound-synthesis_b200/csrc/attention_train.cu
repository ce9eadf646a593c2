// Fused attention for TRAINING (bf16 operands, fp32 softmax / accumulation), head_dim 64, no mask, no dropout
// (reference transformer_utils.py:43-58 FullAttention, :91-109 CrossAttention; the reference differentiates them with torch autograd).
//   forward : O = softmax(scale Q K^T) V, plus LSE[row] = log2(sum_k exp(scale s_k)) so that the backward can rebuild P without storing it
//   backward: two register-resident kernels in the style of the forward one (warp-level mma.sync m16n8k16, FA2 register re-packing):
//     dq kernel  (one CTA per 64 query rows) : S, P = 2^(cS - LSE), dP = dO V^T, dS = scale P (dP - Delta), dQ = dS K; also writes
//                                              Delta[row] = sum_d dO O for the second kernel
//     dkv kernel (one CTA per 64 keys)       : the transposed products S^T = K Q^T, dP^T = V dO^T -> dV = P^T dO, dK = dS^T Q
//   Q, K, V, O, dO, dQ, dK, dV are token-major bf16 with row strides (head h = columns [64h, 64h+64)): they are read from / written into
//   the QKV and gradient buffers of the surrounding GEMMs in place -- no head permutes, no transposes, no stored probabilities.
#include "common.cuh"
#include "diffsound_b200.h"
#include <cuda_bf16.h>

namespace dsb {
namespace {
typedef __nv_bfloat16 bf16;
constexpr int HD = 64, TQ = 64, TK = 64, LDS = 72;  // LDS: padded smem row stride (144 B -> conflict-free ldmatrix)
constexpr int TILE_ELEMS = 64 * LDS;

__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], const bf16* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(smem_u32(p)));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], const bf16* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(smem_u32(p)));
}
__device__ __forceinline__ void mma_bf16(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ uint32_t pack_bf2(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ float dot_bf2(uint32_t x, uint32_t y) {  // (x.lo * y.lo + x.hi * y.hi) of two packed bf16 pairs
  return __uint_as_float(x << 16) * __uint_as_float(y << 16) + __uint_as_float(x & 0xffff0000u) * __uint_as_float(y & 0xffff0000u);
}
// asynchronous 64 x 64 tile copy (global -> padded smem); rows past nrows_valid are zero-filled
__device__ __forceinline__ void load_tile_async(bf16* dst, const bf16* src, long long ld, int row0, int nrows_valid) {
#pragma unroll
  for (int it = 0; it < 4; ++it) {
    const int idx = threadIdx.x + it * 128;
    const int r = idx >> 3, c8 = (idx & 7) * 8;
    const bool ok = row0 + r < nrows_valid;
    const bf16* g = src + (long long)(ok ? row0 + r : 0) * ld + c8;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(dst + r * LDS + c8)), "l"(g), "r"(ok ? 16 : 0) : "memory");
  }
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// A fragments (4 k-steps of 16 head dims) of this warp's 16 rows of a 64 x 64 smem tile
__device__ __forceinline__ void load_a_frags(uint32_t (&a)[4][4], const bf16* tile, int warp, int lane) {
  const int row = warp * 16 + (lane & 7) + 8 * ((lane >> 3) & 1);
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) ldsm_x4(a[ks], tile + row * LDS + ks * 16 + 8 * (lane >> 4));
}
// acc[nt] (16 x 8 per n-tile, 8 n-tiles = 64 columns) += A (16 x 64, fragments) * Bt^T where Bt is a 64 x 64 smem tile whose ROWS are the
// output columns (K of Q K^T, V of dO V^T, Q of K Q^T, dO of V dO^T)
__device__ __forceinline__ void mma_a_bt(float (&acc)[8][4], const uint32_t (&a)[4][4], const bf16* bt, int lane) {
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
    uint32_t f[4];
    const bf16* p = bt + (nt * 8 + (lane & 7)) * LDS + 8 * (lane >> 3);
    ldsm_x4(f, p);
    mma_bf16(acc[nt], a[0], f[0], f[1]);
    mma_bf16(acc[nt], a[1], f[2], f[3]);
    ldsm_x4(f, p + 32);
    mma_bf16(acc[nt], a[2], f[0], f[1]);
    mma_bf16(acc[nt], a[3], f[2], f[3]);
  }
}
// acc (16 x 64) += Pm (16 x 64, fp32 C-layout values re-packed to bf16 A fragments) * Bm where Bm is a 64 x 64 smem tile whose ROWS are the
// reduction index (V of P V, K of dS K, dO of P^T dO, Q of dS^T Q)
__device__ __forceinline__ void mma_p_b(float (&acc)[8][4], const float (&pm)[8][4], const bf16* bm, int lane) {
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) {
    const uint32_t pa[4] = {pack_bf2(pm[2 * kk][0], pm[2 * kk][1]), pack_bf2(pm[2 * kk][2], pm[2 * kk][3]),
                            pack_bf2(pm[2 * kk + 1][0], pm[2 * kk + 1][1]), pack_bf2(pm[2 * kk + 1][2], pm[2 * kk + 1][3])};
    const bf16* bp = bm + (kk * 16 + (lane & 7) + 8 * ((lane >> 3) & 1)) * LDS + 8 * (lane >> 4);
#pragma unroll
    for (int np = 0; np < 4; ++np) {
      uint32_t f[4];
      ldsm_x4_t(f, bp + np * 16);
      mma_bf16(acc[2 * np], pa, f[0], f[1]);
      mma_bf16(acc[2 * np + 1], pa, f[2], f[3]);
    }
  }
}
__device__ __forceinline__ void zero_acc(float (&acc)[8][4]) {
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
}
// store this warp's 16 x 64 accumulator as bf16 rows (row_a = row of elements [0],[1]; row_b = row_a + 8 for [2],[3])
__device__ __forceinline__ void store_rows(bf16* base, long long ld, int row_a, int nrows, const float (&acc)[8][4], float s0, float s1, int t) {
  const int row_b = row_a + 8;
#pragma unroll
  for (int nd = 0; nd < 8; ++nd) {
    if (row_a < nrows) *reinterpret_cast<__nv_bfloat162*>(base + (long long)row_a * ld + nd * 8 + 2 * t) = __floats2bfloat162_rn(acc[nd][0] * s0, acc[nd][1] * s0);
    if (row_b < nrows) *reinterpret_cast<__nv_bfloat162*>(base + (long long)row_b * ld + nd * 8 + 2 * t) = __floats2bfloat162_rn(acc[nd][2] * s1, acc[nd][3] * s1);
  }
}

// ------------------------------------------------------------------------------------------------ forward (+ LSE)
__global__ void __launch_bounds__(128, 4)
attn_train_fwd_kernel(const bf16* __restrict__ q, long long ldq, const bf16* __restrict__ k, long long ldk, const bf16* __restrict__ v, long long ldv,
                      bf16* __restrict__ o, long long ldo, float* __restrict__ lse, int H, int Lq, int Lk, float scale_log2e) {
  __shared__ __align__(16) bf16 Qs[TILE_ELEMS];
  __shared__ __align__(16) bf16 Ks[2][TILE_ELEMS];
  __shared__ __align__(16) bf16 Vs[2][TILE_ELEMS];
  const int qt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const bf16* qb = q + (long long)b * Lq * ldq + h * HD;
  const bf16* kb = k + (long long)b * Lk * ldk + h * HD;
  const bf16* vb = v + (long long)b * Lk * ldv + h * HD;
  load_tile_async(Qs, qb, ldq, qt * TQ, Lq);
  load_tile_async(Ks[0], kb, ldk, 0, Lk);
  load_tile_async(Vs[0], vb, ldv, 0, Lk);
  cp_async_commit();
  uint32_t a[4][4];
  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
  float oacc[8][4];
  zero_acc(oacc);
  const int nchunks = (Lk + TK - 1) / TK;
  for (int kc = 0; kc < nchunks; ++kc) {
    const bf16* Kc = Ks[kc & 1];
    const bf16* Vc = Vs[kc & 1];
    if (kc + 1 < nchunks) {
      load_tile_async(Ks[(kc + 1) & 1], kb, ldk, (kc + 1) * TK, Lk);
      load_tile_async(Vs[(kc + 1) & 1], vb, ldv, (kc + 1) * TK, Lk);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    if (kc == 0) load_a_frags(a, Qs, warp, lane);
    float s[8][4];
    zero_acc(s);
    mma_a_bt(s, a, Kc, lane);
    float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const int key = kc * TK + nt * 8 + 2 * t;
      if (key >= Lk) { s[nt][0] = -INFINITY; s[nt][2] = -INFINITY; }
      if (key + 1 >= Lk) { s[nt][1] = -INFINITY; s[nt][3] = -INFINITY; }
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
    mma_p_b(oacc, s, Vc, lane);
    __syncthreads();
  }
  l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
  const int ra = qt * TQ + warp * 16 + g;
  store_rows(o + (long long)b * Lq * ldo + h * HD, ldo, ra, Lq, oacc, 1.0f / l0, 1.0f / l1, t);
  if (t == 0) {
    float* lb = lse + ((long long)b * H + h) * Lq;
    if (ra < Lq) lb[ra] = m0 * scale_log2e + log2f(l0);
    if (ra + 8 < Lq) lb[ra + 8] = m1 * scale_log2e + log2f(l1);
  }
}

// ------------------------------------------------------------------------------------------------ backward, dQ (and Delta)
__global__ void __launch_bounds__(128)
attn_train_dq_kernel(const bf16* __restrict__ q, long long ldq, const bf16* __restrict__ k, long long ldk, const bf16* __restrict__ v, long long ldv,
                     const bf16* __restrict__ o, long long ldo, const bf16* __restrict__ dout, long long lddo, const float* __restrict__ lse,
                     float* __restrict__ delta, bf16* __restrict__ dq, long long lddq, int H, int Lq, int Lk, float scale, float scale_log2e) {
  extern __shared__ __align__(16) unsigned char smem_dq[];
  bf16* Qs = reinterpret_cast<bf16*>(smem_dq);
  bf16* Ds = Qs + TILE_ELEMS;       // dO
  bf16* Os = Ds + TILE_ELEMS;       // O (only for Delta)
  bf16* Ks = Os + TILE_ELEMS;       // [2]
  bf16* Vs = Ks + 2 * TILE_ELEMS;   // [2]
  const int qt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const long long hoff = h * HD;
  const bf16* kb = k + (long long)b * Lk * ldk + hoff;
  const bf16* vb = v + (long long)b * Lk * ldv + hoff;
  load_tile_async(Qs, q + (long long)b * Lq * ldq + hoff, ldq, qt * TQ, Lq);
  load_tile_async(Ds, dout + (long long)b * Lq * lddo + hoff, lddo, qt * TQ, Lq);
  load_tile_async(Os, o + (long long)b * Lq * ldo + hoff, ldo, qt * TQ, Lq);
  load_tile_async(Ks, kb, ldk, 0, Lk);
  load_tile_async(Vs, vb, ldv, 0, Lk);
  cp_async_commit();
  const int ra = qt * TQ + warp * 16 + g, rb = ra + 8;
  const float* lb = lse + ((long long)b * H + h) * Lq;
  const float lse0 = ra < Lq ? lb[ra] : 0.f, lse1 = rb < Lq ? lb[rb] : 0.f;
  uint32_t aq[4][4], ad[4][4];
  float d0 = 0.f, d1 = 0.f;
  float acc[8][4];
  zero_acc(acc);
  const int nchunks = (Lk + TK - 1) / TK;
  for (int kc = 0; kc < nchunks; ++kc) {
    const bf16* Kc = Ks + (kc & 1) * TILE_ELEMS;
    const bf16* Vc = Vs + (kc & 1) * TILE_ELEMS;
    if (kc + 1 < nchunks) {
      load_tile_async(Ks + ((kc + 1) & 1) * TILE_ELEMS, kb, ldk, (kc + 1) * TK, Lk);
      load_tile_async(Vs + ((kc + 1) & 1) * TILE_ELEMS, vb, ldv, (kc + 1) * TK, Lk);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    if (kc == 0) {
      load_a_frags(aq, Qs, warp, lane);
      load_a_frags(ad, Ds, warp, lane);
      uint32_t ao[4][4];
      load_a_frags(ao, Os, warp, lane);
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {   // fragment registers 0, 2 belong to row g; 1, 3 to row g + 8
        d0 += dot_bf2(ad[ks][0], ao[ks][0]) + dot_bf2(ad[ks][2], ao[ks][2]);
        d1 += dot_bf2(ad[ks][1], ao[ks][1]) + dot_bf2(ad[ks][3], ao[ks][3]);
      }
      d0 += __shfl_xor_sync(0xffffffffu, d0, 1); d0 += __shfl_xor_sync(0xffffffffu, d0, 2);
      d1 += __shfl_xor_sync(0xffffffffu, d1, 1); d1 += __shfl_xor_sync(0xffffffffu, d1, 2);
      if (t == 0) {
        float* db = delta + ((long long)b * H + h) * Lq;
        if (ra < Lq) db[ra] = d0;
        if (rb < Lq) db[rb] = d1;
      }
    }
    float s[8][4], dp[8][4];
    zero_acc(s);
    zero_acc(dp);
    mma_a_bt(s, aq, Kc, lane);    // S  = Q K^T
    mma_a_bt(dp, ad, Vc, lane);   // dP = dO V^T
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const int key = kc * TK + nt * 8 + 2 * t;
      const bool v0 = key < Lk, v1 = key + 1 < Lk;
      const float p0 = v0 ? ex2(fmaf(s[nt][0], scale_log2e, -lse0)) : 0.f, p1 = v1 ? ex2(fmaf(s[nt][1], scale_log2e, -lse0)) : 0.f;
      const float p2 = v0 ? ex2(fmaf(s[nt][2], scale_log2e, -lse1)) : 0.f, p3 = v1 ? ex2(fmaf(s[nt][3], scale_log2e, -lse1)) : 0.f;
      s[nt][0] = scale * p0 * (dp[nt][0] - d0); s[nt][1] = scale * p1 * (dp[nt][1] - d0);
      s[nt][2] = scale * p2 * (dp[nt][2] - d1); s[nt][3] = scale * p3 * (dp[nt][3] - d1);
    }
    mma_p_b(acc, s, Kc, lane);    // dQ += dS K
    __syncthreads();
  }
  store_rows(dq + (long long)b * Lq * lddq + hoff, lddq, ra, Lq, acc, 1.f, 1.f, t);
}

// ------------------------------------------------------------------------------------------------ backward, dK and dV
__global__ void __launch_bounds__(128)
attn_train_dkv_kernel(const bf16* __restrict__ q, long long ldq, const bf16* __restrict__ k, long long ldk, const bf16* __restrict__ v, long long ldv,
                      const bf16* __restrict__ dout, long long lddo, const float* __restrict__ lse, const float* __restrict__ delta,
                      bf16* __restrict__ dk, long long lddk, bf16* __restrict__ dv, long long lddv, int H, int Lq, int Lk, float scale,
                      float scale_log2e) {
  extern __shared__ __align__(16) unsigned char smem_dkv[];
  bf16* Ks = reinterpret_cast<bf16*>(smem_dkv);
  bf16* Vs = Ks + TILE_ELEMS;
  bf16* Qs = Vs + TILE_ELEMS;       // [2]
  bf16* Ds = Qs + 2 * TILE_ELEMS;   // [2]
  float* lse_s = reinterpret_cast<float*>(Ds + 2 * TILE_ELEMS);  // [2][64]
  float* del_s = lse_s + 2 * TQ;                                  // [2][64]
  const int kt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const long long hoff = h * HD;
  const bf16* qb = q + (long long)b * Lq * ldq + hoff;
  const bf16* db = dout + (long long)b * Lq * lddo + hoff;
  const float* lb = lse + ((long long)b * H + h) * Lq;
  const float* eb = delta + ((long long)b * H + h) * Lq;
  auto load_rowstats = [&](int buf, int qc) {
    if (threadIdx.x < TQ) {
      const int r = qc * TQ + threadIdx.x;
      lse_s[buf * TQ + threadIdx.x] = r < Lq ? lb[r] : 0.f;
      del_s[buf * TQ + threadIdx.x] = r < Lq ? eb[r] : 0.f;
    }
  };
  load_tile_async(Ks, k + (long long)b * Lk * ldk + hoff, ldk, kt * TK, Lk);
  load_tile_async(Vs, v + (long long)b * Lk * ldv + hoff, ldv, kt * TK, Lk);
  load_tile_async(Qs, qb, ldq, 0, Lq);
  load_tile_async(Ds, db, lddo, 0, Lq);
  cp_async_commit();
  load_rowstats(0, 0);
  const int ka = kt * TK + warp * 16 + g, kb2 = ka + 8;  // this thread's key rows
  uint32_t ak[4][4], av[4][4];
  float acc_dk[8][4], acc_dv[8][4];
  zero_acc(acc_dk);
  zero_acc(acc_dv);
  const int nchunks = (Lq + TQ - 1) / TQ;
  for (int qc = 0; qc < nchunks; ++qc) {
    const bf16* Qc = Qs + (qc & 1) * TILE_ELEMS;
    const bf16* Dc = Ds + (qc & 1) * TILE_ELEMS;
    const float* lsc = lse_s + (qc & 1) * TQ;
    const float* dlc = del_s + (qc & 1) * TQ;
    if (qc + 1 < nchunks) {
      load_tile_async(Qs + ((qc + 1) & 1) * TILE_ELEMS, qb, ldq, (qc + 1) * TQ, Lq);
      load_tile_async(Ds + ((qc + 1) & 1) * TILE_ELEMS, db, lddo, (qc + 1) * TQ, Lq);
      cp_async_commit();
      load_rowstats((qc + 1) & 1, qc + 1);
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    if (qc == 0) {
      load_a_frags(ak, Ks, warp, lane);
      load_a_frags(av, Vs, warp, lane);
    }
    float st[8][4], dpt[8][4];
    zero_acc(st);
    zero_acc(dpt);
    mma_a_bt(st, ak, Qc, lane);    // S^T  = K Q^T     (rows: keys, columns: queries)
    mma_a_bt(dpt, av, Dc, lane);   // dP^T = V dO^T
    float pt[8][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const int c = nt * 8 + 2 * t;                       // query column inside the tile
      const int qrow = qc * TQ + c;
      const bool q0 = qrow < Lq, q1 = qrow + 1 < Lq;
      const float l0 = lsc[c], l1 = lsc[c + 1], e0 = dlc[c], e1 = dlc[c + 1];
      pt[nt][0] = (q0 && ka < Lk) ? ex2(fmaf(st[nt][0], scale_log2e, -l0)) : 0.f;
      pt[nt][1] = (q1 && ka < Lk) ? ex2(fmaf(st[nt][1], scale_log2e, -l1)) : 0.f;
      pt[nt][2] = (q0 && kb2 < Lk) ? ex2(fmaf(st[nt][2], scale_log2e, -l0)) : 0.f;
      pt[nt][3] = (q1 && kb2 < Lk) ? ex2(fmaf(st[nt][3], scale_log2e, -l1)) : 0.f;
      st[nt][0] = scale * pt[nt][0] * (dpt[nt][0] - e0); st[nt][1] = scale * pt[nt][1] * (dpt[nt][1] - e1);
      st[nt][2] = scale * pt[nt][2] * (dpt[nt][2] - e0); st[nt][3] = scale * pt[nt][3] * (dpt[nt][3] - e1);
    }
    mma_p_b(acc_dv, pt, Dc, lane);   // dV += P^T dO
    mma_p_b(acc_dk, st, Qc, lane);   // dK += dS^T Q
    __syncthreads();
  }
  store_rows(dv + (long long)b * Lk * lddv + hoff, lddv, ka, Lk, acc_dv, 1.f, 1.f, t);
  store_rows(dk + (long long)b * Lk * lddk + hoff, lddk, ka, Lk, acc_dk, 1.f, 1.f, t);
}

constexpr int SMEM_DQ = 7 * TILE_ELEMS * 2;
constexpr int SMEM_DKV = 6 * TILE_ELEMS * 2 + 4 * TQ * 4;
}  // namespace
}  // namespace dsb
using namespace dsb;

static int check_rows(const void* p, long long ld, const char* what) {
  DSB_REQUIRE(ld % 8 == 0 && (reinterpret_cast<uintptr_t>(p) & 15) == 0, "attention_train: %s needs a 16-byte aligned base and a row stride that is a multiple of 8", what);
  return 0;
}

extern "C" int dsb_attention_train_fwd(const void* q, long long ldq, const void* k, long long ldk, const void* v, long long ldv, void* o, long long ldo,
                                       float* lse, int B, int H, int Lq, int Lk, float scale, void* stream) {
  DSB_REQUIRE(B > 0 && H > 0 && Lq > 0 && Lk > 0 && lse, "dsb_attention_train_fwd: bad arguments");
  if (check_rows(q, ldq, "q") || check_rows(k, ldk, "k") || check_rows(v, ldv, "v") || check_rows(o, ldo, "o")) return 2;
  dim3 grid((Lq + TQ - 1) / TQ, H, B);
  attn_train_fwd_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>((const bf16*)q, ldq, (const bf16*)k, ldk, (const bf16*)v, ldv, (bf16*)o, ldo, lse, H, Lq, Lk,
                                                                scale * 1.4426950408889634f);
  DSB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int dsb_attention_train_bwd(const void* q, long long ldq, const void* k, long long ldk, const void* v, long long ldv, const void* o, long long ldo,
                                       const void* dout, long long lddo, const float* lse, float* delta, void* dq, long long lddq, void* dk, long long lddk,
                                       void* dv, long long lddv, int B, int H, int Lq, int Lk, float scale, void* stream) {
  DSB_REQUIRE(B > 0 && H > 0 && Lq > 0 && Lk > 0 && lse && delta, "dsb_attention_train_bwd: bad arguments");
  if (check_rows(q, ldq, "q") || check_rows(k, ldk, "k") || check_rows(v, ldv, "v") || check_rows(o, ldo, "o") || check_rows(dout, lddo, "dout") ||
      check_rows(dq, lddq, "dq") || check_rows(dk, lddk, "dk") || check_rows(dv, lddv, "dv")) return 2;
  static bool attr_done = false;
  if (!attr_done) {
    DSB_CHECK_CUDA(cudaFuncSetAttribute(attn_train_dq_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_DQ));
    DSB_CHECK_CUDA(cudaFuncSetAttribute(attn_train_dkv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_DKV));
    attr_done = true;
  }
  cudaStream_t st = (cudaStream_t)stream;
  const float sl2 = scale * 1.4426950408889634f;
  attn_train_dq_kernel<<<dim3((Lq + TQ - 1) / TQ, H, B), 128, SMEM_DQ, st>>>((const bf16*)q, ldq, (const bf16*)k, ldk, (const bf16*)v, ldv, (const bf16*)o, ldo,
                                                                            (const bf16*)dout, lddo, lse, delta, (bf16*)dq, lddq, H, Lq, Lk, scale, sl2);
  DSB_CHECK_CUDA(cudaGetLastError());
  attn_train_dkv_kernel<<<dim3((Lk + TK - 1) / TK, H, B), 128, SMEM_DKV, st>>>((const bf16*)q, ldq, (const bf16*)k, ldk, (const bf16*)v, ldv, (const bf16*)dout, lddo,
                                                                              lse, delta, (bf16*)dk, lddk, (bf16*)dv, lddv, H, Lq, Lk, scale, sl2);
  DSB_CHECK_CUDA(cudaGetLastError());
  return 0;
}
