// Host build of csrc/train_loss_math.cuh with a single serial lane -- TEST INFRASTRUCTURE.  tests/test_cpu_train_math.py compiles this
// with g++ and compares the column loss and its analytic logit gradient with torch autograd through the oracle.
#include "train_loss_math.cuh"

namespace {
struct SerialCtx {
  int lane() const { return 0; }
  int lanes() const { return 1; }
  float sumf(float v) const { return v; }
  double sumd(double v) const { return v; }
  float maxf(float v) const { return v; }
  int mini(int v) const { return v; }
};
constexpr int CAP = 1056;
}  // namespace

// logits (B, L, K); x0, xt (B, L); t (B); g_main, g_aux (B); sched (8, T+1).  Outputs: dz (B, L, K), prob (B, K+1, L), col (B, L, 2), hits (B, L, 2)
extern "C" int loss_columns_host(const float* logits, const long long* x0, const long long* xt, const long long* t, const float* g_main,
                                 const float* g_aux, const float* sched, int B, int K, int L, int T, float mw0, float mw1, float* dz,
                                 float* prob, float* col, int* hits) {
  if (K + 1 > CAP) return 1;
  SerialCtx c;
  for (int b = 0; b < B; ++b) {
    const dsb_loss::Sched s = dsb_loss::load_sched(sched, T, t[b]);
    for (int l = 0; l < L; ++l) {
      dsb_loss::ColumnIn in;
      in.K = K; in.x0 = (int)x0[b * L + l]; in.xt = (int)xt[b * L + l]; in.is0 = t[b] == 0;
      in.g_main = g_main[b]; in.g_aux = g_aux[b]; in.mw0 = mw0; in.mw1 = mw1;
      const long long ci = (long long)b * L + l;
      dsb_loss::ColumnOut o = dsb_loss::column_loss<SerialCtx, CAP>(c, logits + ci * K, dz + ci * K, prob + (long long)b * (K + 1) * L + l, L, false, in, s);
      col[ci * 2] = o.main; col[ci * 2 + 1] = o.aux;
      hits[ci * 2] = o.x0_hit; hits[ci * 2 + 1] = o.keep_hit;
    }
  }
  return 0;
}
