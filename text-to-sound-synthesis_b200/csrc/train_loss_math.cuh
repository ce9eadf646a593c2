// Per-column math of DiffusionTransformer._train_loss and its analytic gradient with respect to the denoiser logits
// (reference diffusion_transformer.py:408-476 with predict_start :285-289, q_posterior :293-339, multinomial_kl :236-238,
// log_categorical :24).  One "column" = one (batch, position): K logits in, K+1 log-probabilities inside.
//
// The code is written against a lane context `Ctx` (lane(), lanes(), sumf/sumd/maxf/mini reductions) so that the SAME source
// runs as a warp (32 lanes, shuffles) inside train.cu and as a single serial lane when tests/native/loss_math_host.cpp compiles
// this header with g++ -- that host build is how the gradient formulas are checked against torch autograd without a GPU.
//
// Forward (N = K+1, class K = [MASK]):
//   lp[k<K] = clamp(float(log_softmax_fp64(z))[k], -70, 0), lp[K] = -70
//   post(v) : q = v - log_qt ; s = logsumexp(q) ; qn = q - s ; r = lae(qn + A', B') (class K: lae(qn + C1', C')) ;
//             e = r + log_q1 + s ; o = clamp(e, -70, 0)
//   o = post(lp) (model), T = post(X) (true, X = log one-hot of x0 with floor log 1e-30)
//   kl = sum_k exp(T)(T - o) ; nll = -sum_k exp(X) o ; aux = sum_{k<K} exp(X)(X - lp)
//   main = t==0 ? nll : w kl ; auxc = t==0 ? nll : w aux        (w = mask_weight[x_t is MASK ? 0 : 1])
// Backward, with G = d loss / d main and Ga = d loss / d auxc of this column:
//   do[k]  = t==0 ? -(G+Ga) exp(X[k]) : -G w exp(T[k]) ;  de = do * [-70 <= e <= 0]
//   sig[k] = exp(u - r) (u = qn + A' or C1') ;  S1 = sum de ; S2 = sum de sig
//   dq[k]  = de sig + exp(qn[k]) (S1 - S2)
//   dlp[k<K] = dq[k] + (t==0 ? 0 : -Ga w exp(X[k])) ;  dl = dlp * [-70 <= lsm <= 0] ;  dz[k] = dl[k] - softmax(z)[k] sum_j dl[j]
#pragma once
#include <math.h>
#include <stdint.h>

#ifdef __CUDACC__
#define DSB_LM __device__ __forceinline__
#else
#define DSB_LM inline
#endif

namespace dsb_loss {

constexpr float LOGZ = -69.07755279f;  // log(1e-30) in fp32 (index_to_log_onehot floor, diffusion_transformer.py:54)

struct Sched {          // schedule scalars of one batch element
  float la, lb, lc;     // log_at, log_bt, log_ct at t
  float cA, cB, cC;     // log_cumprod_{at,bt,ct} at t
  float pA, pB, pC, pC1;  // log_cumprod_{at,bt,ct}, log_1_min_cumprod_ct at (t-1) mod (T+1)
};

// sched: (8, T+1) rows log_at, log_bt, log_ct, log_1_min_ct, log_cumprod_at, log_cumprod_bt, log_cumprod_ct, log_1_min_cumprod_ct
DSB_LM Sched load_sched(const float* sched, int T, long long t) {
  const int S1 = T + 1;
  const long long tp = t < 0 ? 0 : (t >= T ? T - 1 : t);
  const int tm1 = (int)((tp - 1 + S1) % S1);
  Sched s;
  s.la = sched[0 * S1 + tp]; s.lb = sched[1 * S1 + tp]; s.lc = sched[2 * S1 + tp];
  s.cA = sched[4 * S1 + tp]; s.cB = sched[5 * S1 + tp]; s.cC = sched[6 * S1 + tp];
  s.pA = sched[4 * S1 + tm1]; s.pB = sched[5 * S1 + tm1]; s.pC = sched[6 * S1 + tm1]; s.pC1 = sched[7 * S1 + tm1];
  return s;
}

DSB_LM float lae(float a, float b) {  // log_add_exp, diffusion_transformer.py:28-30
  const float m = fmaxf(a, b);
  return m + logf(expf(a - m) + expf(b - m));
}

struct ColumnIn {
  int K;
  int x0, xt;
  int is0;            // t == 0
  float g_main, g_aux;  // d loss / d main, d loss / d auxc
  float mw0, mw1;     // mask_weight
};
struct ColumnOut {
  float main, aux;    // the column's contribution to kl_loss[b] and kl_aux_loss[b]
  int x0_hit;         // argmax(log_x0_recon) == x0      (diffusion_acc_list bookkeeping, :424-433)
  int keep_hit;       // argmax(log_model_prob) == x_t   (diffusion_keep_list)
};

// closed-form posterior of one column; v = log p(x0) over N classes.  Outputs o (clamped), and for the backward pass qn, sig, egate.
template <class Ctx, int CAP>
DSB_LM void posterior_column(const Ctx& c, const float (&v)[CAP], int K, int xt, const Sched& s, float (&o)[CAP], float (&qn)[CAP],
                             float (&sig)[CAP], float (&egate)[CAP]) {
  const int N = K + 1, lane = c.lane(), NL = c.lanes();
  const bool masked = (xt == K);
  float lq1[CAP];
  float qmax = -INFINITY;
#pragma unroll
  for (int j = 0; j < CAP; ++j) {
    const int k = lane + NL * j;
    float lqt, l1;
    if (k < K) {
      if (masked) { lqt = s.cC; l1 = s.lc; }
      else {
        const float oh = (k == xt) ? 0.f : LOGZ;
        lqt = lae(oh + s.cA, s.cB);
        l1 = lae(oh + s.la, s.lb);
      }
    } else {
      lqt = masked ? 0.f : LOGZ;
      l1 = lqt;
    }
    lq1[j] = l1;
    qn[j] = (k < N) ? v[j] - lqt : -INFINITY;
    qmax = fmaxf(qmax, qn[j]);
  }
  qmax = c.maxf(qmax);
  float ssum = 0.f;
#pragma unroll
  for (int j = 0; j < CAP; ++j)
    if (lane + NL * j < N) ssum += expf(qn[j] - qmax);
  ssum = c.sumf(ssum);
  const float slse = logf(ssum) + qmax;
#pragma unroll
  for (int j = 0; j < CAP; ++j) {
    const int k = lane + NL * j;
    if (k < N) {
      const float q = qn[j] - slse;
      const float u = q + (k < K ? s.pA : s.pC1);
      const float r = lae(u, k < K ? s.pB : s.pC);
      const float e = r + lq1[j] + slse;
      qn[j] = q;
      sig[j] = expf(u - r);
      egate[j] = (e >= -70.f && e <= 0.f) ? 1.f : 0.f;
      o[j] = fminf(fmaxf(e, -70.f), 0.f);
    } else {
      qn[j] = -INFINITY; sig[j] = 0.f; egate[j] = 0.f; o[j] = -70.f;
    }
  }
}

// argmax over the N valid entries (first index wins)
template <class Ctx, int CAP>
DSB_LM int argmax_column(const Ctx& c, const float (&v)[CAP], int N) {
  const int lane = c.lane(), NL = c.lanes();
  float best = -INFINITY;
#pragma unroll
  for (int j = 0; j < CAP; ++j)
    if (lane + NL * j < N) best = fmaxf(best, v[j]);
  best = c.maxf(best);
  int bi = 0x7fffffff;
#pragma unroll
  for (int j = 0; j < CAP; ++j) {
    const int k = lane + NL * j;
    if (k < N && v[j] == best && k < bi) bi = k;
  }
  return c.mini(bi);
}

// z: K logits of the column (global memory); dz: K gradients out (may be null); prob_out: optional log_model_prob, element k at
// prob_out[k * prob_stride] (exp() of it when prob_exp: the 'logits' entry of forward()'s output, :573-574).
template <class Ctx, int CAP>
DSB_LM ColumnOut column_loss(const Ctx& c, const float* z, float* dz, float* prob_out, long long prob_stride, bool prob_exp,
                             const ColumnIn& in, const Sched& s) {
  const int K = in.K, N = K + 1, lane = c.lane(), NL = c.lanes();
  float lp[CAP], pk[CAP], gate[CAP];
  {
    float mx = -INFINITY;
#pragma unroll
    for (int j = 0; j < CAP; ++j) {
      const int k = lane + NL * j;
      lp[j] = k < K ? z[k] : -INFINITY;
      mx = fmaxf(mx, lp[j]);
    }
    mx = c.maxf(mx);
    double se = 0.0;
#pragma unroll
    for (int j = 0; j < CAP; ++j)
      if (lane + NL * j < K) se += exp((double)lp[j] - (double)mx);
    se = c.sumd(se);
    const double lse = log(se);
#pragma unroll
    for (int j = 0; j < CAP; ++j) {
      const int k = lane + NL * j;
      if (k < K) {
        const float l = (float)(((double)lp[j] - (double)mx) - lse);
        pk[j] = expf(l);
        gate[j] = (l >= -70.f && l <= 0.f) ? 1.f : 0.f;
        lp[j] = fminf(fmaxf(l, -70.f), 0.f);
      } else {
        pk[j] = 0.f; gate[j] = 0.f; lp[j] = -70.f;
      }
    }
  }
  ColumnOut out;
  out.x0_hit = argmax_column<Ctx, CAP>(c, lp, N) == in.x0;

  float o[CAP], qn[CAP], sig[CAP], egate[CAP];
  float T[CAP];
  {
    float X[CAP], tq[CAP], ts[CAP], tg[CAP];
#pragma unroll
    for (int j = 0; j < CAP; ++j) X[j] = (lane + NL * j == in.x0) ? 0.f : LOGZ;
    posterior_column<Ctx, CAP>(c, X, K, in.xt, s, T, tq, ts, tg);
  }
  posterior_column<Ctx, CAP>(c, lp, K, in.xt, s, o, qn, sig, egate);
  out.keep_hit = argmax_column<Ctx, CAP>(c, o, N) == in.xt;
  if (prob_out) {
#pragma unroll
    for (int j = 0; j < CAP; ++j) {
      const int k = lane + NL * j;
      if (k < N) prob_out[(long long)k * prob_stride] = prob_exp ? expf(o[j]) : o[j];
    }
  }

  const float w = (in.xt == K) ? in.mw0 : in.mw1;
  float kl = 0.f, nll = 0.f, aux = 0.f;
#pragma unroll
  for (int j = 0; j < CAP; ++j) {
    const int k = lane + NL * j;
    if (k < N) {
      const float Xk = (k == in.x0) ? 0.f : LOGZ;
      const float eX = expf(Xk);
      kl += expf(T[j]) * (T[j] - o[j]);
      nll -= eX * o[j];
      if (k < K) aux += eX * (Xk - lp[j]);
    }
  }
  kl = c.sumf(kl); nll = c.sumf(nll); aux = c.sumf(aux);
  out.main = in.is0 ? nll : w * kl;
  out.aux = in.is0 ? nll : w * aux;
  if (!dz) return out;

  // ---- backward ----
  float S1 = 0.f, S2 = 0.f;
#pragma unroll
  for (int j = 0; j < CAP; ++j) {
    const int k = lane + NL * j;
    float de = 0.f;
    if (k < N) {
      const float Xk = (k == in.x0) ? 0.f : LOGZ;
      const float d_o = in.is0 ? -(in.g_main + in.g_aux) * expf(Xk) : -in.g_main * w * expf(T[j]);
      de = d_o * egate[j];
    }
    T[j] = de;  // reuse
    S1 += de;
    S2 += de * sig[j];
  }
  S1 = c.sumf(S1); S2 = c.sumf(S2);
  float S3 = 0.f;
#pragma unroll
  for (int j = 0; j < CAP; ++j) {
    const int k = lane + NL * j;
    float dl = 0.f;
    if (k < K) {
      const float Xk = (k == in.x0) ? 0.f : LOGZ;
      const float dq = T[j] * sig[j] + expf(qn[j]) * (S1 - S2);
      const float dlp = dq + (in.is0 ? 0.f : -in.g_aux * w * expf(Xk));
      dl = dlp * gate[j];
    }
    T[j] = dl;
    S3 += dl;
  }
  S3 = c.sumf(S3);
#pragma unroll
  for (int j = 0; j < CAP; ++j) {
    const int k = lane + NL * j;
    if (k < K) dz[k] = T[j] - pk[j] * S3;
  }
  return out;
}

}  // namespace dsb_loss
