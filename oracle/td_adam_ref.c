/* Plain-C restatement of the two fp32 / index pieces of the Q-learning step.  TEST INFRASTRUCTURE ONLY:
 * built by oracle/Makefile (and __graft_entry__.build()) into oracle/_ref/libtdref.so and called from
 * tests/ only; the product path never loads it.
 *
 *   td_ref    process_batch's loss arithmetic (train_q_network.py:134-180): gather Q(s) at the taken
 *             action, first-maximum arg-max of the online Q(s'), gather of the target Q(s'), terminal
 *             masking, Bellman target (or the LINEAR form), rect clamp, 0.5 (Q_b - y)^2 (* valid_mask),
 *             mean; and the compare_ground_truth branches (:170-178).  Also the closed-form gradient
 *             dLoss/dQ(s) = (Q_b - y) * mask / (B*C) at the taken action.
 *   adam_ref  torch.optim.Adam's update with default hyper-parameters (train_q_network.py:124,227):
 *             m = b1 m + (1-b1) g; v = b2 v + (1-b2) g^2; p -= lr/(1-b1^t) * m / (sqrt(v)/sqrt(1-b2^t) + eps)
 *
 * Pinned by tests/test_oracle_golden.py against tests/golden/td_branches.npz (vectors produced by the
 * reference's own process_batch closure, oracle/make_td_goldens.py) and against torch.optim.Adam.
 * Arithmetic is fp32 in the order torch evaluates it (every intermediate is a float tensor there).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

/* labels (rew / term / valid_mask) as fp32: what `.float()` makes of the loader's int64 detections or, with
 * CONFIDENCE_REWARD, of its float64 detector scores (train_q_network.py:158-160, 166-167) */
int td_ref_f32(const float* q_s, const float* q_no, const float* q_nt, const int64_t* act, const float* rew,
               const float* term, const float* valid, const double* gt, int B, int C, int A, float gamma,
               int double_dqn, int clip_rect, int linear, int use_valid, int ground_truth, int value_learning,
               float* dq, float* y_out, int64_t* best_out, float* loss_out) {
  if (B < 0 || C < 1 || A < 1) return -2;
  const long total = (long)B * C;
  const float inv = total > 0 ? 1.0f / (float)total : 0.f;
  double sum = 0.0;
  for (long i = 0; i < total; ++i) {
    const long b = i / C;
    const int a_taken = (int)act[b];
    const float q_b = q_s[i * A + a_taken];                       /* .gather(2, action_indices) (:134-137) */
    float diff, mask = 1.f, y;
    int best = 0;
    if (ground_truth) {
      const double g = gt[i];
      if (value_learning) {                                       /* :172-176 */
        const int nan = isnan(g);
        mask = nan ? 0.f : 1.f;
        y = nan ? 0.f : (float)g;
        diff = q_b * mask - y;
      } else {                                                    /* :177-178 (NaNs propagate) */
        y = (float)g;
        diff = q_b - y;
      }
    } else {
      const float* sel = double_dqn ? q_no + i * A : q_nt + i * A;
      float bv = sel[0];
      for (int a = 1; a < A; ++a)
        if (sel[a] > bv) { bv = sel[a]; best = a; }               /* argmax(-1): first maximum (:147) */
      float q_a = q_nt[i * A + best];                             /* :153-154 */
      q_a = q_a * (1.f - term[i]);                                /* :158 */
      y = linear ? rew[i] + (q_a - 0.1f) : rew[i] + gamma * q_a;  /* :159-162 */
      if (clip_rect) y = fminf(fmaxf(y, 0.f), 1.f);               /* :163-164 */
      diff = q_b - y;
    }
    float l = 0.5f * (diff * diff);                               /* :165 */
    if (use_valid && !ground_truth) {                             /* :166-167 */
      mask = valid[i];
      l = l * mask;
    }
    sum += (double)l;
    if (dq != 0)
      for (int a = 0; a < A; ++a) dq[i * A + a] = (a == a_taken) ? diff * mask * inv : 0.f;
    if (y_out != 0) y_out[i] = y;
    if (best_out != 0) best_out[i] = best;
  }
  if (loss_out != 0) *loss_out = (float)(sum * (double)inv);      /* .mean() (:180) */
  return 0;
}

/* the loader's usual int64 labels (dataloaders/q_learning_real.py:78-84) */
int td_ref(const float* q_s, const float* q_no, const float* q_nt, const int64_t* act, const int64_t* rew,
           const int64_t* term, const int64_t* valid, const double* gt, int B, int C, int A, float gamma,
           int double_dqn, int clip_rect, int linear, int use_valid, int ground_truth, int value_learning,
           float* dq, float* y_out, int64_t* best_out, float* loss_out) {
  if (B < 0 || C < 1 || A < 1) return -2;
  const long total = (long)B * C;
  float* f = (float*)malloc(sizeof(float) * 3 * (size_t)(total > 0 ? total : 1));
  if (f == 0) return -3;
  for (long i = 0; i < total; ++i) {
    f[i] = rew != 0 ? (float)rew[i] : 0.f;
    f[total + i] = term != 0 ? (float)term[i] : 0.f;
    f[2 * total + i] = valid != 0 ? (float)valid[i] : 1.f;
  }
  const int rc = td_ref_f32(q_s, q_no, q_nt, act, f, f + total, f + 2 * total, gt, B, C, A, gamma, double_dqn,
                            clip_rect, linear, use_valid, ground_truth, value_learning, dq, y_out, best_out, loss_out);
  free(f);
  return rc;
}

int adam_ref(float* p, const float* g, float* m, float* v, long n, double lr, double b1, double b2, double eps,
             int step) {
  if (n < 0 || step < 1) return -2;
  const double bc1 = 1.0 - pow(b1, (double)step), bc2 = 1.0 - pow(b2, (double)step);
  const float step_size = (float)(lr / bc1);
  const float inv_sqrt_bc2 = (float)(1.0 / sqrt(bc2));
  for (long i = 0; i < n; ++i) {
    const float gi = g[i];
    m[i] = (float)b1 * m[i] + (1.f - (float)b1) * gi;
    v[i] = (float)b2 * v[i] + (1.f - (float)b2) * gi * gi;
    const float denom = sqrtf(v[i]) * inv_sqrt_bc2 + (float)eps;
    p[i] -= step_size * (m[i] / denom);
  }
  return 0;
}
