/*
 * ORACLE - test infrastructure, not product code.
 *
 * Plain-C restatement (fp64 accumulation, no libraries) of the reference's
 * TCN / GCN eval-mode forward, written independently of the ATen-based
 * restatement in nasr_oracle.py so the two triangulate each other.
 * Only tests/ may load it (oracle/c_oracle.py); the product never does.
 * Parity pin: tests/test_oracle.py checks it against the golden vectors the
 * reference itself produced (tests/golden/, made by tests/golden/make_golden.py).
 *
 * Reference lines restated (relative to src/neural_audio_spring_reverb/):
 *   causal dilated conv      networks/custom_layers.py:85-88
 *   FiLM + eval BatchNorm    networks/custom_layers.py:32-42
 *   PReLU / residual (TCN)   networks/tcn.py:73-86
 *   gate / residual (GCN)    networks/gcn.py:53-61, custom_layers.py:103-111
 *   out_net (+ tanh)         networks/tcn.py:154, networks/gcn.py:145-146
 *
 * Weight blob order: per block conv.w[W][Cin][k], conv.b[W], (film: adaptor.w[2W][cd],
 * adaptor.b[2W], bn.w[W], bn.b[W], bn.mean[W], bn.var[W]), (TCN: prelu[1]),
 * res.w[C][Cin]; then out_net.w[out_ch][C].
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
  int arch; /* 0 TCN, 1 GCN */
  int n_blocks, in_ch, out_ch, n_channels, kernel_size, cond_dim, has_film;
  const int* dilations;
} oracle_desc;

/* x [B][in_ch][T] -> y [B][out_ch][T]; returns 0, or -1 on allocation failure */
int nasr_oracle_forward(const oracle_desc* d, const float* w, const float* x, const float* cond,
                        float* y, int B, int64_t T) {
  const int C = d->n_channels, k = d->kernel_size, cd = d->cond_dim;
  const int W = d->arch ? 2 * C : C;
  const int maxc = C > d->in_ch ? C : d->in_ch;
  double* cur = (double*)malloc(sizeof(double) * (size_t)maxc * T);
  double* nxt = (double*)malloc(sizeof(double) * (size_t)maxc * T);
  double* z = (double*)malloc(sizeof(double) * (size_t)W);
  double* g = (double*)malloc(sizeof(double) * (size_t)2 * W);
  if (!cur || !nxt || !z || !g) { free(cur); free(nxt); free(z); free(g); return -1; }
  for (int b = 0; b < B; ++b) {
    const float* p = w;
    for (int64_t i = 0; i < (int64_t)d->in_ch * T; ++i) cur[i] = x[(int64_t)b * d->in_ch * T + i];
    for (int blk = 0; blk < d->n_blocks; ++blk) {
      const int Cin = blk == 0 ? d->in_ch : C;
      const int64_t dil = d->dilations[blk];
      const float* cw = p; p += (size_t)W * Cin * k;
      const float* cb = p; p += W;
      const float *aw = 0, *ab = 0, *bw = 0, *bb = 0, *bm = 0, *bv = 0;
      if (d->has_film) {
        aw = p; p += (size_t)2 * W * cd;
        ab = p; p += 2 * W;
        bw = p; p += W; bb = p; p += W; bm = p; p += W; bv = p; p += W;
        for (int o = 0; o < 2 * W; ++o) { /* adaptor Linear */
          double s = ab[o];
          for (int q = 0; q < cd; ++q) s += (double)aw[(size_t)o * cd + q] * cond[(size_t)b * cd + q];
          g[o] = s;
        }
      }
      double slope = 0.0;
      if (!d->arch) { slope = *p; p += 1; }
      const float* rw = p; p += (size_t)C * Cin;
      for (int64_t t = 0; t < T; ++t) {
        for (int o = 0; o < W; ++o) {
          double s = cb[o];
          for (int j = 0; j < k; ++j) {
            const int64_t ts = t - (int64_t)(k - 1 - j) * dil; /* left zero padding */
            if (ts < 0) continue;
            for (int ci = 0; ci < Cin; ++ci) s += (double)cw[((size_t)o * Cin + ci) * k + j] * cur[(size_t)ci * T + ts];
          }
          if (d->has_film) {
            s = (s - bm[o]) / sqrt((double)bv[o] + 1e-5) * bw[o] + bb[o]; /* eval BatchNorm */
            s = s * g[o] + g[W + o];                                      /* FiLM: g first, then b */
          }
          z[o] = s;
        }
        for (int o = 0; o < C; ++o) {
          double a;
          if (!d->arch) a = z[o] > 0 ? z[o] : slope * z[o];
          else a = tanh(z[o]) * (1.0 / (1.0 + exp(-z[C + o])));
          double r = 0.0;
          for (int ci = 0; ci < Cin; ++ci) r += (double)rw[(size_t)o * Cin + ci] * cur[(size_t)ci * T + t];
          nxt[(size_t)o * T + t] = a + r;
        }
      }
      double* tmp = cur; cur = nxt; nxt = tmp;
    }
    for (int64_t t = 0; t < T; ++t)
      for (int o = 0; o < d->out_ch; ++o) {
        double s = 0.0;
        for (int c = 0; c < C; ++c) s += (double)p[(size_t)o * C + c] * cur[(size_t)c * T + t];
        y[((size_t)b * d->out_ch + o) * T + t] = (float)(d->arch ? tanh(s) : s);
      }
  }
  free(cur); free(nxt); free(z); free(g);
  return 0;
}
