// Split-fp16 ("f16x3") tcgen05 / TMEM attention core -- the parity-grade twin of attention_tc.cu
// (reference transformer_utils.py:48-54 FullAttention, :99-105 CrossAttention: softmax(Q K^T / sqrt(64)) V, 16 heads x 64).
//
// Every operand is an fp16 (hi | lo) pair (x ~ hi + lo, 22 significand bits) produced by the split GEMM epilogue:
//   S = Qlo Khi^T + Qhi Klo^T + Qhi Khi^T : 3 x tcgen05.mma kind::f16 passes into one fp32 TMEM accumulator (128 lanes x kpad columns);
//   softmax: 8 warps (two per TMEM lane quadrant, each half of the 32-key chunks), one query row per thread; fp32 max / exp2 / sum;
//            P is split into (hi | lo) and written IN PLACE over the S columns it came from (chunk of 32 keys = 32 fp32 columns
//            -> 16 packed-fp16 hi columns + 16 lo columns), so S + P + O fit the 512 TMEM columns at 288 keys;
//   O = Plo Vhi + Phi Vlo + Phi Vhi       : A operand from TMEM, B = V rows as TMA wrote them (MN-major SW128), fp32 accumulator in TMEM;
//   epilogue: O / rowsum -> (hi | lo) fp16 pair -> HBM (the A operand of the output projection's split GEMM).
// One CTA per (batch, head); K/V (hi and lo) are staged once by TMA, Q tiles are double-buffered.  fp32-class accuracy: measured
// against an fp64 reference in tests/test_gpu_split.py.
#include "common.cuh"
#include "diffsound_b200.h"
#include <cuda_fp16.h>

namespace dsb {
namespace {
constexpr int SP_HD = 64;
constexpr int SP_QM = 128;
constexpr int SP_KMAX = 288;      // padded key capacity (multiple of 32)
// warps [0, 4*NSW): softmax (NSW per TMEM lane quadrant; they also run the epilogue when EPIW == 0), warp 4*NSW: control (TMA producer, MMA issuer,
// TMEM owner), warp 4*NSW+1: remainder rows, then EPIW (0 or 4) dedicated epilogue warps, one per lane quadrant
constexpr int sp_threads(int nsw, int epiw) { return (4 * nsw + 2 + epiw) * 32; }
constexpr int SP_STG = 16 * 33;  // floats of one epilogue staging tile: 16 rows x 32 columns, padded (bank-conflict free both ways)
constexpr int SP_QTILE = SP_QM * 128;  // bytes of one 128 x 64 fp16 SW128 tile

struct SpParams {
  int B, H, Lq, Lk, kpad, n_qt, box_rows, n_box;
  int n_heads;          // B * H (batch, head) units; CTA c owns the contiguous slice [c * n / grid, (c + 1) * n / grid)
  int q_bufs;           // Q tile buffers: 2 (prefetch one tile ahead) or 1 (short key sequences: leaves room for two CTAs per SM)
  int rem_rows;         // > 0: the last (Lq % 128 <= 16) query rows are computed by the remainder warp with mma.sync, not by a third tensor tile
  const __half* q;      // global Q (hi half; lo at + q_lo_col) for the remainder warp's register fragments
  long long ldq;
  int q_lo_col, k_lo_col, v_lo_col;  // column (element) offset of the lo halves inside the tensor maps
  long long ldo, o_lo_off;
  __half* o;
  float scale_log2e;
  uint32_t tmem_cols;
};

__device__ __forceinline__ void sp_tmem_st_x16(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]),
        "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}
__device__ __forceinline__ void sp_tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// D[tmem] (+)= A[tmem] * B[smem]
__device__ __forceinline__ void sp_umma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
               ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ constexpr uint32_t sp_idesc(int M, int N, bool b_mn_major) {  // fp16 operands, fp32 accumulate
  return (1u << 4) | (b_mn_major ? (1u << 16) : 0u) | (static_cast<uint32_t>(N >> 3) << 17) | (static_cast<uint32_t>(M >> 4) << 24);
}
// optional phase timestamps of CTA 0 (tools/attn_split_timing.py): [tile][0..7] softmax warp 0, [tile][8..15] control thread; SM clock cycles
__device__ long long sp_dbg_times[8 * 16];
__device__ int sp_dbg_on = 0;
__device__ __forceinline__ void sp_stamp(int tile, int slot) {
  if (sp_dbg_on && blockIdx.x == 0 && tile < 8) sp_dbg_times[tile * 16 + slot] = clock64();
}
// wait used by the many softmax / epilogue threads: back off between polls so that they do not take issue slots from the one control thread
__device__ __forceinline__ void sp_wait_backoff(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    __nanosleep(40);
    if (++spins > (1u << 24)) { printf("dsb: attention mbarrier wait timeout (block %d thread %d)\n", blockIdx.x, threadIdx.x); __trap(); }
  }
}
// epilogue of one 32-row x 32-column block of O (ov = this thread's row): scale, stage 16 rows at a time through padded shared memory, and store
// (hi | lo) pairs with lanes 0-15 / 16-31 each covering one row's 32 columns = 64 contiguous bytes
__device__ __forceinline__ void sp_store_block(const SpParams& p, float* stg, const uint32_t (&ov)[32], float inv, __half* obase, int row0, int lane) {
#pragma unroll
  for (int hr = 0; hr < 2; ++hr) {
    if ((lane >> 4) == hr) {
#pragma unroll
      for (int j = 0; j < 32; ++j) stg[(lane & 15) * 33 + j] = __uint_as_float(ov[j]) * inv;
    }
    __syncwarp();
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      const int r = it + 8 * (lane >> 4), c = 2 * (lane & 15);
      const float x0 = stg[r * 33 + c], x1 = stg[r * 33 + c + 1];
      const int row = row0 + hr * 16 + r;
      if (row < p.Lq) {
        uint32_t hi, lo;
        const __half2 hh = __floats2half2_rn(x0, x1);
        const __half2 ll = __floats2half2_rn(x0 - __low2float(hh), x1 - __high2float(hh));
        hi = *reinterpret_cast<const uint32_t*>(&hh);
        lo = *reinterpret_cast<const uint32_t*>(&ll);
        __half* o = obase + (long long)row * p.ldo + c;
        *reinterpret_cast<uint32_t*>(o) = hi;
        *reinterpret_cast<uint32_t*>(o + p.o_lo_off) = lo;
      }
    }
    __syncwarp();
  }
}
__device__ __forceinline__ float sp_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// ---- remainder rows (265 = 2 x 128 + 9): one warp, legacy mma.sync m16n8k16, same split arithmetic (3 MMAs per product), straight from the
// K / V tiles the TMA staged in shared memory, concurrently with the tensor tiles -- instead of a third latency-bound 128-row tile.
__device__ __forceinline__ uint32_t sp_sw128(int r, int chunk) { return r * 128 + ((chunk ^ (r & 7)) << 4); }
__device__ __forceinline__ void sp_ldsm_x4(uint32_t (&r)[4], const uint8_t* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(smem_u32(p)));
}
__device__ __forceinline__ void sp_ldsm_x4_t(uint32_t (&r)[4], const uint8_t* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(smem_u32(p)));
}
__device__ __forceinline__ void sp_mma(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void sp_pack_pair(float a, float b, uint32_t& hi, uint32_t& lo) {
  const __half2 h = __floats2half2_rn(a, b);
  const __half2 l = __floats2half2_rn(a - __low2float(h), b - __high2float(h));
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}
__device__ __forceinline__ void sp_remainder_rows(const SpParams& p, const uint8_t* sKh, const uint8_t* sKl, const uint8_t* sVh, const uint8_t* sVl, int b, int h,
                                                  int row0, int lane) {
  const int g = lane >> 2, t = lane & 3;
  const __half* qb = p.q + (long long)b * p.Lq * p.ldq + h * SP_HD;
  uint32_t ah[4][4], al[4][4];
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) {
    const int ra = row0 + g, rb = row0 + g + 8, c = ks * 16 + 2 * t;
    const __half* pa = qb + (long long)ra * p.ldq + c;
    const __half* pb = qb + (long long)rb * p.ldq + c;
    ah[ks][0] = ra < p.Lq ? *reinterpret_cast<const uint32_t*>(pa) : 0u;
    ah[ks][1] = rb < p.Lq ? *reinterpret_cast<const uint32_t*>(pb) : 0u;
    ah[ks][2] = ra < p.Lq ? *reinterpret_cast<const uint32_t*>(pa + 8) : 0u;
    ah[ks][3] = rb < p.Lq ? *reinterpret_cast<const uint32_t*>(pb + 8) : 0u;
    al[ks][0] = ra < p.Lq ? *reinterpret_cast<const uint32_t*>(pa + p.q_lo_col) : 0u;
    al[ks][1] = rb < p.Lq ? *reinterpret_cast<const uint32_t*>(pb + p.q_lo_col) : 0u;
    al[ks][2] = ra < p.Lq ? *reinterpret_cast<const uint32_t*>(pa + p.q_lo_col + 8) : 0u;
    al[ks][3] = rb < p.Lq ? *reinterpret_cast<const uint32_t*>(pb + p.q_lo_col + 8) : 0u;
  }
  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
  float oacc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i) oacc[i][0] = oacc[i][1] = oacc[i][2] = oacc[i][3] = 0.f;
  for (int kc = 0; kc * 64 < p.kpad; ++kc) {
    const int keys_left = p.Lk - kc * 64;
    if (keys_left <= 0) break;
    const int n_live = keys_left >= 64 ? 8 : (keys_left + 7) >> 3;
    float s[8][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
      if (nt >= n_live) continue;
      const int row = kc * 64 + nt * 8 + (lane & 7);
      uint32_t kh[4], kl[4];
#pragma unroll
      for (int half = 0; half < 2; ++half) {  // head-dim columns [0,32) then [32,64)
        sp_ldsm_x4(kh, sKh + sp_sw128(row, 4 * half + (lane >> 3)));
        sp_ldsm_x4(kl, sKl + sp_sw128(row, 4 * half + (lane >> 3)));
        sp_mma(s[nt], al[2 * half], kh[0], kh[1]);
        sp_mma(s[nt], al[2 * half + 1], kh[2], kh[3]);
        sp_mma(s[nt], ah[2 * half], kl[0], kl[1]);
        sp_mma(s[nt], ah[2 * half + 1], kl[2], kl[3]);
        sp_mma(s[nt], ah[2 * half], kh[0], kh[1]);
        sp_mma(s[nt], ah[2 * half + 1], kh[2], kh[3]);
      }
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
    const float c0 = sp_ex2((m0 - mn0) * p.scale_log2e), c1 = sp_ex2((m1 - mn1) * p.scale_log2e);
    const float ms0 = mn0 * p.scale_log2e, ms1 = mn1 * p.scale_log2e;
    m0 = mn0; m1 = mn1;
    l0 *= c0; l1 *= c1;
#pragma unroll
    for (int nd = 0; nd < 8; ++nd) { oacc[nd][0] *= c0; oacc[nd][1] *= c0; oacc[nd][2] *= c1; oacc[nd][3] *= c1; }
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      s[nt][0] = sp_ex2(fmaf(s[nt][0], p.scale_log2e, -ms0)); s[nt][1] = sp_ex2(fmaf(s[nt][1], p.scale_log2e, -ms0));
      s[nt][2] = sp_ex2(fmaf(s[nt][2], p.scale_log2e, -ms1)); s[nt][3] = sp_ex2(fmaf(s[nt][3], p.scale_log2e, -ms1));
      l0 += s[nt][0] + s[nt][1];
      l1 += s[nt][2] + s[nt][3];
    }
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      if (kk * 2 >= n_live) continue;
      uint32_t ph[4], pl[4];
      sp_pack_pair(s[2 * kk][0], s[2 * kk][1], ph[0], pl[0]);
      sp_pack_pair(s[2 * kk][2], s[2 * kk][3], ph[1], pl[1]);
      sp_pack_pair(s[2 * kk + 1][0], s[2 * kk + 1][1], ph[2], pl[2]);
      sp_pack_pair(s[2 * kk + 1][2], s[2 * kk + 1][3], ph[3], pl[3]);
      const int row = kc * 64 + kk * 16 + (lane & 7) + 8 * ((lane >> 3) & 1);
#pragma unroll
      for (int np = 0; np < 4; ++np) {
        uint32_t vh[4], vl[4];
        sp_ldsm_x4_t(vh, sVh + sp_sw128(row, 2 * np + (lane >> 4)));
        sp_ldsm_x4_t(vl, sVl + sp_sw128(row, 2 * np + (lane >> 4)));
        sp_mma(oacc[2 * np], pl, vh[0], vh[1]);
        sp_mma(oacc[2 * np + 1], pl, vh[2], vh[3]);
        sp_mma(oacc[2 * np], ph, vl[0], vl[1]);
        sp_mma(oacc[2 * np + 1], ph, vl[2], vl[3]);
        sp_mma(oacc[2 * np], ph, vh[0], vh[1]);
        sp_mma(oacc[2 * np + 1], ph, vh[2], vh[3]);
      }
    }
  }
  l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
  const float i0 = 1.0f / l0, i1 = 1.0f / l1;
  const int ra = row0 + g, rb = row0 + g + 8;
  __half* ob = p.o + (long long)b * p.Lq * p.ldo + h * SP_HD;
#pragma unroll
  for (int nd = 0; nd < 8; ++nd) {
    uint32_t hi, lo;
    if (ra < p.Lq) {
      sp_pack_pair(oacc[nd][0] * i0, oacc[nd][1] * i0, hi, lo);
      *reinterpret_cast<uint32_t*>(ob + (long long)ra * p.ldo + nd * 8 + 2 * t) = hi;
      *reinterpret_cast<uint32_t*>(ob + (long long)ra * p.ldo + p.o_lo_off + nd * 8 + 2 * t) = lo;
    }
    if (rb < p.Lq) {
      sp_pack_pair(oacc[nd][2] * i1, oacc[nd][3] * i1, hi, lo);
      *reinterpret_cast<uint32_t*>(ob + (long long)rb * p.ldo + nd * 8 + 2 * t) = hi;
      *reinterpret_cast<uint32_t*>(ob + (long long)rb * p.ldo + p.o_lo_off + nd * 8 + 2 * t) = lo;
    }
  }
}

// MINB = CTAs per SM the register budget must allow: 2 for short key sequences (cross-attention, 96 keys: 113 KB of smem and 256 TMEM columns
// per CTA, so two heads' softmax / MMA chains interleave on one SM), 1 for the 288-key self-attention tiles (208 KB of smem).
// NSW = softmax warps per lane quadrant, MAXC = 32-key chunks fetched per TMEM round trip.  Measured at B=16 (tools/attn_split_bench.py): the softmax
// is bound by the MUFU / conversion pipe (288 exp2 + fp16 pair conversions per row), not by TMEM latency -- <1,3,2> (12 warps, two chunks per
// round trip) ran 47.5 us against 45.2 us for <1,2,1>, so the simplest configuration is used.
// EPIW = 4: dedicated epilogue warps, so the softmax warps go straight on to the next tile (self-attention: the chain is what bounds a head).
// Measured with the phase timestamps (tools/attn_split_timing.py, B=16): one thread-per-row epilogue store pass cost 4 160 cycles per tile (every
// warp-level store touched 32 different 128-byte lines) and ran on the softmax warps, serialised with the next tile; spinning waiters starved the
// single MMA-issuing thread (65 cycles per tcgen05.mma issued).  Now: stores go through a padded shared-memory transpose (two rows = 2 x 64
// contiguous bytes per warp-level store), waiting warps back off with nanosleep, and the epilogue has its own warps where registers allow.
template <int MINB, int NSW, int MAXC, int EPIW>
__global__ void __launch_bounds__(sp_threads(NSW, EPIW), MINB)
attention_tc_split_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k, const __grid_constant__ CUtensorMap map_v,
                          const __grid_constant__ SpParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int kv_bytes = p.kpad * 128;
  uint8_t* sQ = smem;                         // [q_bufs][hi, lo] x 16 KB
  uint8_t* sKh = sQ + p.q_bufs * 2 * SP_QTILE;
  uint8_t* sKl = sKh + kv_bytes;
  uint8_t* sVh = sKl + kv_bytes;
  uint8_t* sVl = sVh + kv_bytes;
  float* s_max = reinterpret_cast<float*>(sVl + kv_bytes);   // [2 tile parities][NSW][128 rows]
  float* s_sum = s_max + 2 * NSW * 128;                      // same shape
  float* s_stage = s_sum + 2 * NSW * 128;                    // [8 epilogue slots][SP_STG]
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_stage + 8 * SP_STG);
  constexpr int CTRL = 4 * NSW;
  uint64_t* k_full = bars;
  uint64_t* q_full = bars + 1;   // [2]
  uint64_t* s_full = bars + 3;
  uint64_t* p_full = bars + 4;
  uint64_t* o_full = bars + 5;
  uint64_t* o_empty = bars + 6;
  uint64_t* v_full = bars + 7;
  uint64_t* rem_done = bars + 8;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 10);

  // persistent over a contiguous slice of (batch, head) units: the next head's K is fetched as soon as the last S of the current head has been
  // computed, its V as soon as the last P.V has, so that only the first head's loads are exposed
  const int u0 = (int)((long long)p.n_heads * blockIdx.x / gridDim.x), u1 = (int)((long long)p.n_heads * (blockIdx.x + 1) / gridDim.x);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    mbar_init(k_full, 1); mbar_init(v_full, 1); mbar_init(rem_done, 1);
    mbar_init(&q_full[0], 1); mbar_init(&q_full[1], 1);
    mbar_init(s_full, 1); mbar_init(p_full, 4 * NSW); mbar_init(o_full, 1); mbar_init(o_empty, EPIW > 0 ? EPIW : 8);
    fence_barrier_init();
    prefetch_tmap(&map_q); prefetch_tmap(&map_k); prefetch_tmap(&map_v);
  }
  if (warp == CTRL) {
    tmem_alloc(tmem_ptr, p.tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  const uint32_t tS = tmem_base, tO = tmem_base + p.kpad;
  pdl_wait();
  pdl_trigger();

  if (warp == CTRL) {
    if (lane == 0) {
      // ------------------------------------------------------------------ control thread: TMA producer + MMA issuer
      auto load_k = [&](int u) {
        const int b = u / p.H, h = u - b * p.H;
        mbar_arrive_expect_tx(k_full, 2 * kv_bytes);
        for (int bx = 0; bx < p.n_box; ++bx) {
          const int r = b * p.Lk + bx * p.box_rows, off = bx * p.box_rows * 128;
          tma_load_3d(&map_k, k_full, sKh + off, h * SP_HD, r, 0);
          tma_load_3d(&map_k, k_full, sKl + off, p.k_lo_col + h * SP_HD, r, 0);
        }
      };
      auto load_v = [&](int u) {
        const int b = u / p.H, h = u - b * p.H;
        mbar_arrive_expect_tx(v_full, 2 * kv_bytes);
        for (int bx = 0; bx < p.n_box; ++bx) {
          const int r = b * p.Lk + bx * p.box_rows, off = bx * p.box_rows * 128;
          tma_load_3d(&map_v, v_full, sVh + off, h * SP_HD, r, 0);
          tma_load_3d(&map_v, v_full, sVl + off, p.v_lo_col + h * SP_HD, r, 0);
        }
      };
      const int nb = p.q_bufs;
      const int n_tiles = (u1 - u0) * p.n_qt;
      auto issue_q = [&](int tg) {  // tg = tile index inside this CTA's slice
        const int u = u0 + tg / p.n_qt, qt = tg % p.n_qt;
        const int b = u / p.H, h = u - b * p.H;
        const int qb = tg % nb;
        mbar_arrive_expect_tx(&q_full[qb], 2 * SP_QTILE);
        tma_load_3d(&map_q, &q_full[qb], sQ + qb * 2 * SP_QTILE, h * SP_HD, b * p.Lq + qt * SP_QM, 0);
        tma_load_3d(&map_q, &q_full[qb], sQ + qb * 2 * SP_QTILE + SP_QTILE, p.q_lo_col + h * SP_HD, b * p.Lq + qt * SP_QM, 0);
      };
      if (u0 < u1) {
        load_k(u0);
        load_v(u0);
        issue_q(0);
        if (nb > 1 && n_tiles > 1) issue_q(1);
      }
      const int n_hi = p.kpad > 256 ? 256 : p.kpad, n_lo = p.kpad - n_hi;
      const uint32_t id_hi = sp_idesc(SP_QM, n_hi, false), id_lo = sp_idesc(SP_QM, n_lo > 0 ? n_lo : 16, false);
      const uint32_t id_pv = sp_idesc(SP_QM, SP_HD, true);
      const int ksteps = p.kpad >> 4;
      for (int tg = 0; tg < n_tiles; ++tg) {
        const int hu = tg / p.n_qt, qt = tg - hu * p.n_qt;     // head index inside the slice, query tile inside the head
        const bool last_of_head = qt == p.n_qt - 1, more_heads = u0 + hu + 1 < u1;
        const int qb = tg % nb;
        sp_stamp(tg, 12);
        if (qt == 0) mbar_wait(k_full, hu & 1);
        mbar_wait(&q_full[qb], (tg / nb) & 1);
        sp_stamp(tg, 13);
        tc_fence_after();
        sp_stamp(tg, 14);
        // ---- S = Qlo Khi^T + Qhi Klo^T + Qhi Khi^T  (tcgen05.mma from one thread execute in issue order: this also follows P.V(tg-1))
        const uint32_t qh = smem_u32(sQ + qb * 2 * SP_QTILE), ql = qh + SP_QTILE;
        const uint32_t kh = smem_u32(sKh), kl = smem_u32(sKl);
#pragma unroll
        for (int pass = 0; pass < 3; ++pass) {
          const uint32_t qa = pass == 0 ? ql : qh, ka = pass == 1 ? kl : kh;
          const uint64_t dq = make_sw128_kmajor_desc(qa), dk = make_sw128_kmajor_desc(ka), dk2 = make_sw128_kmajor_desc(ka + 256 * 128);
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            const uint32_t acc = (pass | ks) != 0 ? 1u : 0u;
            umma<false>(tS, dq + 2 * ks, dk + 2 * ks, id_hi, acc);
            if (n_lo > 0) umma<false>(tS + 256, dq + 2 * ks, dk2 + 2 * ks, id_lo, acc);
          }
        }
        umma_commit(s_full);
        sp_stamp(tg, 8);
        // the Q buffer (and, after a head's last tile, the K tiles) are free once S has been computed
        const bool refill_q = tg + nb < n_tiles, refill_k = last_of_head && more_heads;
        if (refill_q || refill_k) mbar_wait(s_full, tg & 1);
        if (refill_q) issue_q(tg + nb);
        if (refill_k) {
          if (p.rem_rows > 0) mbar_wait(rem_done, hu & 1);   // the remainder warp reads K / V of this head through ldmatrix
          load_k(u0 + hu + 1);
        }
        sp_stamp(tg, 9);
        mbar_wait(p_full, tg & 1);                       // P(tg) is in TMEM
        sp_stamp(tg, 10);
        if (tg > 0) mbar_wait(o_empty, (tg - 1) & 1);    // epilogue(tg-1) has read O
        if (qt == 0) mbar_wait(v_full, hu & 1);
        tc_fence_after();
        // ---- O = Plo Vhi + Phi Vlo + Phi Vhi : per 16-key step, A = 8 packed-fp16 TMEM columns, B = 16 V rows (2048 B)
        const uint64_t dvh = make_sw128_kmajor_desc(smem_u32(sVh)), dvl = make_sw128_kmajor_desc(smem_u32(sVl));
        for (int ks = 0; ks < ksteps; ++ks) {
          const uint32_t a_hi = tS + (ks >> 1) * 32 + (ks & 1) * 8, a_lo = a_hi + 16;
          sp_umma_ts(tO, a_lo, dvh + (uint64_t)(ks * 128), id_pv, ks != 0 ? 1u : 0u);
          sp_umma_ts(tO, a_hi, dvl + (uint64_t)(ks * 128), id_pv, 1u);
          sp_umma_ts(tO, a_hi, dvh + (uint64_t)(ks * 128), id_pv, 1u);
        }
        umma_commit(o_full);
        sp_stamp(tg, 11);
        if (refill_k) {  // V is free once this head's last P.V has completed
          mbar_wait(o_full, tg & 1);
          load_v(u0 + hu + 1);
        }
      }
    }
  } else if (warp == CTRL + 1) {
    // ------------------------------------------------------------------ remainder rows (after the full 128-row tiles), concurrent with them
    if (p.rem_rows > 0) {
      for (int u = u0; u < u1; ++u) {
        const int hu = u - u0;
        mbar_wait(k_full, hu & 1);
        mbar_wait(v_full, hu & 1);
        sp_remainder_rows(p, sKh, sKl, sVh, sVl, u / p.H, u % p.H, p.n_qt * SP_QM, lane);
        __syncwarp();
        if (lane == 0) mbar_arrive(rem_done);
      }
    }
  } else if (warp < CTRL) {
    // ------------------------------------------------------------------ softmax (+ epilogue when EPIW == 0) warps
    const int quad = warp & 3, half = warp >> 2;  // half = which of the NSW column slices of this quadrant
    const uint32_t lane_off = static_cast<uint32_t>(quad * 32) << 16;
    const int row_in_tile = quad * 32 + lane;
    const int n_chunks = p.kpad >> 5;
    // slice j owns chunks [j * n / NSW, (j + 1) * n / NSW): 9 chunks -> 3,3,3 with NSW = 3
    const int c_begin = (half * n_chunks / NSW) * 32, c_end = ((half + 1) * n_chunks / NSW) * 32;
    // MAXC = chunks held in registers at once: one TMEM round trip per pass covers up to MAXC chunks; longer slices go in groups
    const int n_tiles = (u1 - u0) * p.n_qt;
    for (int tg = 0; tg < n_tiles; ++tg) {
      const int hu = tg / p.n_qt, qt = tg - hu * p.n_qt;
      const int b = (u0 + hu) / p.H, h = (u0 + hu) - b * p.H;
      const bool live = qt * SP_QM + quad * 32 < p.Lq;  // a 32-row slab entirely beyond Lq does no exp work; its rows are never stored
      sp_wait_backoff(s_full, tg & 1);
      tc_fence_after();
      if (threadIdx.x == 0) sp_stamp(tg, 0);
      float mx = -INFINITY, sum = 0.f;
      float* smx = s_max + (tg & 1) * (NSW * 128);
      auto chunk_max = [&](const uint32_t (&sv)[32], int c) {
        if (c + 32 <= p.Lk) {
#pragma unroll
          for (int j = 0; j < 32; ++j) mx = fmaxf(mx, __uint_as_float(sv[j]));
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (c + j < p.Lk) mx = fmaxf(mx, __uint_as_float(sv[j]));
        }
      };
      if (live) {
        for (int c = c_begin; c < c_end; c += 32 * MAXC) {
          uint32_t sa[32], sb[32];
          const bool two = MAXC > 1 && c + 32 < c_end;
          tmem_ld_32x32(tS + lane_off + c, sa);
          if (two) tmem_ld_32x32(tS + lane_off + c + 32, sb);
          tmem_ld_wait();
          chunk_max(sa, c);
          if (two) chunk_max(sb, c + 32);
        }
      }
      smx[half * 128 + row_in_tile] = mx;
      if (threadIdx.x == 0) sp_stamp(tg, 1);
      asm volatile("bar.sync 1, %0;" ::"n"(32 * 4 * NSW) : "memory");
      if (threadIdx.x == 0) sp_stamp(tg, 2);
      if (live) {
#pragma unroll
        for (int j = 0; j < NSW; ++j) mx = fmaxf(mx, smx[j * 128 + row_in_tile]);
        const float ms = mx * p.scale_log2e;
        auto chunk_p = [&](const uint32_t (&sv)[32], int c) {
          uint32_t ph[16], pl[16];
          const bool full = c + 32 <= p.Lk;
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            float p0 = sp_ex2(fmaf(__uint_as_float(sv[2 * j]), p.scale_log2e, -ms));
            float p1 = sp_ex2(fmaf(__uint_as_float(sv[2 * j + 1]), p.scale_log2e, -ms));
            if (!full) {
              if (c + 2 * j >= p.Lk) p0 = 0.f;
              if (c + 2 * j + 1 >= p.Lk) p1 = 0.f;
            }
            sum += p0 + p1;
            const __half2 hh = __floats2half2_rn(p0, p1);  // low half = even key
            const __half2 ll = __floats2half2_rn(p0 - __low2float(hh), p1 - __high2float(hh));
            ph[j] = *reinterpret_cast<const uint32_t*>(&hh);
            pl[j] = *reinterpret_cast<const uint32_t*>(&ll);
          }
          sp_tmem_st_x16(tS + lane_off + c, ph);        // in place: this thread has consumed these 32 S columns
          sp_tmem_st_x16(tS + lane_off + c + 16, pl);
        };
        for (int c = c_begin; c < c_end; c += 32 * MAXC) {
          uint32_t sa[32], sb[32];
          const bool two = MAXC > 1 && c + 32 < c_end;
          tmem_ld_32x32(tS + lane_off + c, sa);
          if (two) tmem_ld_32x32(tS + lane_off + c + 32, sb);
          tmem_ld_wait();
          chunk_p(sa, c);
          if (two) chunk_p(sb, c + 32);
        }
        sp_tmem_st_wait();
      }
      s_sum[((tg & 1) * NSW + half) * 128 + row_in_tile] = sum;
      tc_fence_before();
      __syncwarp();
      if (threadIdx.x == 0) sp_stamp(tg, 3);
      if (lane == 0) mbar_arrive(p_full);
      if (EPIW > 0) continue;  // dedicated epilogue warps take it from here
      // ---- epilogue on the softmax warps (slices 0 and 1 of every quadrant): 32 rows x 32 of the 64 output columns
      if (half >= 2) continue;
      sp_wait_backoff(o_full, tg & 1);
      tc_fence_after();
      if (threadIdx.x == 0) sp_stamp(tg, 4);
      if (live) {
        float tot = 0.f;
#pragma unroll
        for (int j = 0; j < NSW; ++j) tot += s_sum[((tg & 1) * NSW + j) * 128 + row_in_tile];
        uint32_t ov[32];
        tmem_ld_32x32(tO + lane_off + half * 32, ov);
        tmem_ld_wait();
        sp_store_block(p, s_stage + warp * SP_STG, ov, 1.0f / tot, p.o + (long long)b * p.Lq * p.ldo + h * SP_HD + half * 32, qt * SP_QM + quad * 32, lane);
      }
      tc_fence_before();
      __syncwarp();
      if (threadIdx.x == 0) sp_stamp(tg, 5);
      if (lane == 0) mbar_arrive(o_empty);
    }
  }
  if (EPIW > 0 && warp >= CTRL + 2) {
    // ------------------------------------------------------------------ dedicated epilogue warps: quadrant = warp % 4, all 64 output columns
    const int quad = warp & 3;
    const uint32_t lane_off = static_cast<uint32_t>(quad * 32) << 16;
    const int row_in_tile = quad * 32 + lane;
    const int n_tiles = (u1 - u0) * p.n_qt;
    float* stg = s_stage + (warp - (CTRL + 2)) * SP_STG;
    for (int tg = 0; tg < n_tiles; ++tg) {
      const int hu = tg / p.n_qt, qt = tg - hu * p.n_qt;
      const int b = (u0 + hu) / p.H, h = (u0 + hu) - b * p.H;
      sp_wait_backoff(o_full, tg & 1);
      tc_fence_after();
      if (warp == CTRL + 2 && lane == 0) sp_stamp(tg, 6);
      if (qt * SP_QM + quad * 32 < p.Lq) {
        float tot = 0.f;
#pragma unroll
        for (int j = 0; j < NSW; ++j) tot += s_sum[((tg & 1) * NSW + j) * 128 + row_in_tile];
        const float inv = 1.0f / tot;
        __half* obase = p.o + (long long)b * p.Lq * p.ldo + h * SP_HD;
#pragma unroll
        for (int cb = 0; cb < 2; ++cb) {
          uint32_t ov[32];
          tmem_ld_32x32(tO + lane_off + cb * 32, ov);
          tmem_ld_wait();
          if (cb == 1) {  // O has been read completely: the MMA thread may start the next P.V
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(o_empty);
          }
          sp_store_block(p, stg, ov, inv, obase + cb * 32, qt * SP_QM + quad * 32, lane);
        }
        if (warp == CTRL + 2 && lane == 0) sp_stamp(tg, 7);
      } else {
        __syncwarp();
        if (lane == 0) mbar_arrive(o_empty);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == CTRL) {
    tc_fence_after();
    tmem_dealloc(tmem_base, p.tmem_cols);
  }
}
}  // namespace
}  // namespace dsb

extern "C" int dsb_attention_tc_split(const void* q, long long ldq, long long q_lo_off, const void* k, long long ldk, long long k_lo_off, const void* v,
                                      long long ldv, long long v_lo_off, void* o, long long ldo, long long o_lo_off, int B, int H, int Lq, int Lk,
                                      float scale, void* stream) {
  using namespace dsb;
  DSB_REQUIRE(B > 0 && H > 0 && Lq > 0 && Lk > 0 && Lk <= SP_KMAX, "dsb_attention_tc_split: need 0 < Lk <= %d", SP_KMAX);
  DSB_REQUIRE(ldo % 8 == 0 && o_lo_off % 8 == 0 && (reinterpret_cast<uintptr_t>(o) & 15) == 0, "dsb_attention_tc_split: o must be 16-byte aligned, ldo / o_lo_off %% 8 == 0");
  DSB_REQUIRE(q_lo_off >= (long long)H * SP_HD && k_lo_off >= (long long)H * SP_HD && v_lo_off >= (long long)H * SP_HD && o_lo_off >= (long long)H * SP_HD,
              "dsb_attention_tc_split: the lo halves must not overlap the hi halves");
  DSB_REQUIRE(q_lo_off + (long long)H * SP_HD <= ldq && k_lo_off + (long long)H * SP_HD <= ldk && v_lo_off + (long long)H * SP_HD <= ldv,
              "dsb_attention_tc_split: lo halves must lie inside a row (lo_off + H*64 <= ld)");
  SpParams p{};
  p.B = B; p.H = H; p.Lq = Lq; p.Lk = Lk;
  p.kpad = (Lk + 31) & ~31;
  const int rem = Lq % SP_QM;
  if (Lq > SP_QM && rem > 0 && rem <= 16) {  // e.g. 265 = 2 x 128 + 9: the 9 rows go to the remainder warp, not to a third tensor tile
    p.n_qt = Lq / SP_QM;
    p.rem_rows = rem;
  } else {
    p.n_qt = (Lq + SP_QM - 1) / SP_QM;
    p.rem_rows = 0;
  }
  p.q = (const __half*)q; p.ldq = ldq;
  p.n_box = p.kpad > 256 ? 2 : 1;
  p.box_rows = p.kpad / p.n_box;
  DSB_REQUIRE(p.box_rows % 8 == 0, "dsb_attention_tc_split: internal box size");
  p.q_lo_col = (int)q_lo_off; p.k_lo_col = (int)k_lo_off; p.v_lo_col = (int)v_lo_off;
  p.ldo = ldo; p.o_lo_off = o_lo_off; p.o = (__half*)o;
  p.scale_log2e = scale * 1.4426950408889634f;
  uint32_t cols = 32;
  while ((int)cols < p.kpad + SP_HD) cols <<= 1;
  p.tmem_cols = cols;
  CUtensorMap mq, mk, mv;
  if (make_operand_map(&mq, q, DSB_DTYPE_F16, q_lo_off + (long long)H * SP_HD, (long long)B * Lq, 1, ldq, 0, SP_QM)) return 3;
  if (make_operand_map(&mk, k, DSB_DTYPE_F16, k_lo_off + (long long)H * SP_HD, (long long)B * Lk, 1, ldk, 0, p.box_rows)) return 3;
  if (make_operand_map(&mv, v, DSB_DTYPE_F16, v_lo_off + (long long)H * SP_HD, (long long)B * Lk, 1, ldv, 0, p.box_rows)) return 3;
  const int fixed = (2 * 2 * 2 * 128 + 8 * SP_STG) * 4 + 16 * 8 + 1024;  // s_max + s_sum (NSW = 2), epilogue staging, 9 barriers + TMEM address, alignment
  const bool two = cols <= 256 && 2 * SP_QTILE + 4 * p.kpad * 128 + fixed <= 112 * 1024;  // two CTAs per SM fit
  p.q_bufs = 1;  // the next Q tile is requested as soon as S has been computed and lands long before the chain needs it
  const int smem = p.q_bufs * 2 * SP_QTILE + 4 * p.kpad * 128 + fixed;
  static int attr_smem[2] = {0, 0};
  if (smem > attr_smem[two]) {
    if (two) DSB_CHECK_CUDA(cudaFuncSetAttribute(attention_tc_split_kernel<2, 2, 1, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    else DSB_CHECK_CUDA(cudaFuncSetAttribute(attention_tc_split_kernel<1, 2, 1, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_smem[two] = smem;
  }
  p.n_heads = B * H;
  int grid = sm_count() * (two ? 2 : 1);
  if (grid > p.n_heads) grid = p.n_heads;
  if (two) DSB_CHECK_CUDA(launch_pdl(attention_tc_split_kernel<2, 2, 1, 0>, dim3(grid), dim3(sp_threads(2, 0)), smem, (cudaStream_t)stream, mq, mk, mv, p));
  else DSB_CHECK_CUDA(launch_pdl(attention_tc_split_kernel<1, 2, 1, 4>, dim3(grid), dim3(sp_threads(2, 4)), smem, (cudaStream_t)stream, mq, mk, mv, p));
  return 0;
}

/* debug: enable / read the phase timestamps of CTA 0 (not part of the product path) */
extern "C" int dsb_attention_split_timing(int enable, long long* host_out_128) {
  using namespace dsb;
  int v = enable;
  DSB_CHECK_CUDA(cudaMemcpyToSymbol(sp_dbg_on, &v, sizeof(int)));
  if (host_out_128) DSB_CHECK_CUDA(cudaMemcpyFromSymbol(host_out_128, sp_dbg_times, sizeof(long long) * 128));
  return 0;
}
