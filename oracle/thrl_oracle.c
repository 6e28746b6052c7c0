/* thrl_oracle.c — CPU restatement of the th_rl QTable training hot path.
 *
 * TEST INFRASTRUCTURE, NOT PRODUCT.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this file's shared object; the product
 * (th_rl_b200/) never does and has no CPU fallback.
 *
 * Parity status: PINNED.  tests/test_oracle_golden.py checks this restatement bit-for-bit
 * (actions, rewards, prices, f64 tables, counters, epsilon, per-epoch logs) against streams
 * recorded from the unmodified reference by oracle/make_goldens.py (the .npz files under tests/golden).
 * The reference ships no tests or golden vectors of its own.
 *
 * Everything here is scalar f64/f32 IEEE arithmetic in the reference's operation order; build
 * with -ffp-contract=off (see oracle/Makefile) so no FMA is formed.
 *
 * Reference lines restated (paths relative to /root/reference):
 *   scan_run()      th_rl/trainer.py:45-70      epoch / step loop, log accumulation :65-66
 *   act_row()       th_rl/agents.py:47-49,84-88 float32 encode used when acting (trainer.py:53)
 *   upd_row()       th_rl/agents.py:47-49,62,66 float64 encode used when updating
 *   scale_action()  th_rl/agents.py:51-57
 *   env step        th_rl/environments.py:22-39
 *   buffer          th_rl/buffers.py:12,18-19,28-41 (deque(maxlen=capacity), replay in order, empty)
 *   train_net()     th_rl/agents.py:59-78
 *   mlp_*()         th_rl/agents.py:119-194 (Reinforce: pi, sample_action, scale, train_net) + torch.optim.Adam,
 *                   torch.nn.utils.clip_grad_norm_, torch.distributions.Categorical (math restated; pinned against the
 *                   reference's torch results to a stated tolerance, tests/test_oracle_golden.py)
 * Two modes have no reference counterpart and are specified in DESIGN.md ("fp32 storage",
 * "Philox streams"); the CUDA kernels must match this file bit-for-bit in those modes too.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../include/thrl.h"
#include <pthread.h>
#include <unistd.h>

/* ---------------------------------------------------------------- Philox4x32-10 (DESIGN.md "Philox streams") */
#define PHILOX_M0 0xD2511F53u
#define PHILOX_M1 0xCD9E8D57u
#define PHILOX_W0 0x9E3779B9u
#define PHILOX_W1 0xBB67AE85u
enum { STREAM_ACT = 0, STREAM_ENV = 1, STREAM_INIT_Q = 2, STREAM_INIT_P = 3 };

static void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                          uint32_t out[4]) {
  for (int r = 0; r < 10; ++r) {
    uint64_t p0 = (uint64_t)PHILOX_M0 * c0, p1 = (uint64_t)PHILOX_M1 * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += PHILOX_W0; k1 += PHILOX_W1;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
/* 53-bit uniform in [0,1): 32 bits of hi, top 21 bits of lo */
static double u53(uint32_t hi, uint32_t lo) {
  uint64_t m = ((uint64_t)hi << 21) | (lo >> 11);
  return (double)m * (1.0 / 9007199254740992.0);
}

/* ---------------------------------------------------------------- deterministic N(0,1) from a 53-bit uniform
 * (used by thrl_oracle_qtable_init; DESIGN.md "Device init").  Only + - * / and sqrt, in a fixed order, so the
 * CUDA version (explicit _rn intrinsics) reproduces it bit-for-bit.  log() is our own: frexp + atanh series. */
#define THRL_ORACLE_MAX_ACTIONS 1024 /* scratch for one agent's action probabilities */
static double det_log(double x) {
  int e;
  double m = frexp(x, &e); /* m in [0.5,1) */
  if (m < 0.70710678118654752) { m = m * 2.0; e -= 1; }
  double s = (m - 1.0) / (m + 1.0), s2 = s * s;
  double p = 1.0 / 27.0;
  for (int k = 25; k >= 1; k -= 2) p = p * s2 + 1.0 / (double)k;
  return (double)e * 0.6931471805599453094 + 2.0 * s * p;
}
/* Acklam's rational approximation of the inverse normal CDF (|rel err| < 1.2e-9), p in (0,1) */
static double det_norminv(double p) {
  static const double a[6] = {-3.969683028665376e+01, 2.209460984245205e+02, -2.759285104469687e+02,
                              1.383577518672690e+02, -3.066479806614716e+01, 2.506628277459239e+00};
  static const double b[5] = {-5.447609879822406e+01, 1.615858368580409e+02, -1.556989798598866e+02,
                              6.680131188771972e+01, -1.328068155288572e+01};
  static const double c[6] = {-7.784894002430293e-03, -3.223964580411365e-01, -2.400758277161838e+00,
                              -2.549732539343734e+00, 4.374664141464968e+00, 2.938163982698783e+00};
  static const double d[4] = {7.784695709041462e-03, 3.224671290700398e-01, 2.445134137142996e+00,
                              3.754408661907416e+00};
  const double plow = 0.02425;
  if (p < plow || p > 1.0 - plow) {
    double pp = p < plow ? p : 1.0 - p;
    double q = sqrt(-2.0 * det_log(pp));
    double num = ((((c[0] * q + c[1]) * q + c[2]) * q + c[3]) * q + c[4]) * q + c[5];
    double den = (((d[0] * q + d[1]) * q + d[2]) * q + d[3]) * q + 1.0;
    double x = num / den;
    return p < plow ? x : -x;
  }
  double q = p - 0.5, r = q * q;
  double num = (((((a[0] * r + a[1]) * r + a[2]) * r + a[3]) * r + a[4]) * r + a[5]) * q;
  double den = ((((b[0] * r + b[1]) * r + b[2]) * r + b[3]) * r + b[4]) * r + 1.0;
  return num / den;
}

/* ---------------------------------------------------------------- table access by dtype */
typedef struct Table {
  int dtype;
  void* base; /* this agent's table of this run */
  int rows, cols;
  int stride; /* elements between rows (ThrlAgentSpec.row_stride) */
} Table;
static double tget(const Table* t, int64_t row, int64_t col) {
  int64_t i = row * t->stride + col;
  return t->dtype == THRL_F64 ? ((const double*)t->base)[i] : (double)((const float*)t->base)[i];
}
static void tset(Table* t, int64_t row, int64_t col, double v) {
  int64_t i = row * t->stride + col;
  if (t->dtype == THRL_F64) ((double*)t->base)[i] = v;
  else ((float*)t->base)[i] = (float)v; /* fp32 storage: one rounding, to nearest even */
}
/* numpy.argmax(table[[row]]): first maximal index (agents.py:88) */
static int row_argmax(const Table* t, int64_t row) {
  int best = 0;
  double bv = tget(t, row, 0);
  for (int k = 1; k < t->cols; ++k) {
    double v = tget(t, row, k);
    if (v > bv) { bv = v; best = k; }
  }
  return best;
}
/* numpy.max(table[ns]) (agents.py:71) */
static double row_max(const Table* t, int64_t row) {
  double bv = tget(t, row, 0);
  for (int k = 1; k < t->cols; ++k) {
    double v = tget(t, row, k);
    if (v > bv) bv = v;
  }
  return bv;
}

/* agents.py:47-49 on the float32 state trainer.py:53 hands to sample_action */
static int64_t act_row(double price, const ThrlAgentSpec* s) {
  float st = (float)price;
  float x = st / (float)s->max_state;
  x = x * (float)s->states;
  return (int64_t)rintf(x); /* numpy.round: half to even */
}
/* agents.py:47-49 on the float64 states stored in the buffer (agents.py:62,66) */
static int64_t upd_row(double price, const ThrlAgentSpec* s) {
  double x = price / s->max_state;
  x = x * (double)s->states;
  return (int64_t)rint(x);
}
/* agents.py:51-57 */
static double scale_action(int k, const ThrlAgentSpec* s) {
  return (double)k / ((double)s->actions - 1.0) * (s->action_hi - s->action_lo) + s->action_lo;
}

/* environments.py:27 `Q = sum(A)` as CPython >= 3.12 evaluates it (Python/bltinmodule.c builtin_sum): while the running
 * result and the items are exact Python floats the sum is Neumaier-compensated, the compensation is folded in when the
 * first other item (or the end) is reached, and everything after that is plain left-to-right addition.  Scaled actions
 * of Reinforce / ActorCritic / CAC are Python floats (`.item()`, agents.py:160-163,270-273,374-378); a QTable's are
 * numpy.float64 (numpy.int64 index, agents.py:80-89), which is not an exact float.  So: `lead` = number of leading MLP
 * agents; for lead <= 2 this equals the naive sum.  (Interpreters before 3.12 sum naively; the goldens are recorded with
 * the 3.12 of this image.)  Checked against builtin sum() on 3e5 random mixed lists by tests/test_oracle_golden.py. */
static int lead_exact_floats(const ThrlGame* G) {
  int m = 0;
  while (m < G->n_agents && G->agent[m].kind != THRL_AGENT_QTABLE) ++m;
  return m;
}
static double py_sum_quantities(const double* aq, int n, int lead) {
  if (lead == 0) {
    double q = 0.0;
    for (int i = 0; i < n; ++i) q = q + aq[i];
    return q;
  }
  double f = 0.0 + aq[0], c = 0.0;
  int i = 1;
  for (; i < lead; ++i) {
    const double x = aq[i], t = f + x;
    if (fabs(f) >= fabs(x)) { double d = f - t; d = d + x; c = c + d; }
    else { double d = x - t; d = d + f; c = c + d; }
    f = t;
  }
  if (c != 0.0 && isfinite(c)) f = f + c;
  for (; i < n; ++i) f = f + aq[i];
  return f;
}

/* ---------------------------------------------------------------- Reinforce agent (th_rl/agents.py:119-194), float32
 * Every operation is a separate IEEE f32 operation in a fixed order (no FMA), so the CUDA kernel reproduces it exactly.
 * expf is our own (Cody-Waite reduction + Cephes polynomial, ~2 ulp): libm and libdevice expf differ in the last bit. */
static float det_expf(float x) {
  if (x < -87.0f) return 0.0f;
  float kf = rintf(x * 1.44269504f);
  float r = x - kf * 0.693359375f;
  r = r - kf * -2.12194440e-4f;
  float p = 1.9875691500e-4f;
  p = p * r + 1.3981999507e-3f;
  p = p * r + 8.3334519073e-3f;
  p = p * r + 4.1665795894e-2f;
  p = p * r + 1.6666665459e-1f;
  p = p * r + 5.0000001201e-1f;
  p = p * r;
  p = p * r;
  p = p + r;
  p = p + 1.0f;
  return ldexpf(p, (int)kf);
}
static int mlp_P(const ThrlAgentSpec* s) {
  if (s->kind == THRL_AGENT_CAC) return 5 * s->hidden + 3;
  const int p = 2 * s->hidden + s->actions * s->hidden + s->actions;
  return s->kind == THRL_AGENT_ACTORCRITIC ? p + s->hidden + 1 : p;
}
static int mlp_entry_words(const ThrlAgentSpec* s) { return s->kind == THRL_AGENT_REINFORCE ? 3 : 4; }
/* pi(x) (agents.py:148-152): h = relu(fc1(x)), logits = fc_pi(h), softmax.  par: w1[H] b1[H] W[A][H] bp[A]. */
static void mlp_forward(const float* par, int H, int A, float s, float* h, float* prob) {
  const float *w1 = par, *b1 = par + H, *W = par + 2 * H, *bp = par + 2 * H + (size_t)A * H;
  for (int j = 0; j < H; ++j) {
    float v = s * w1[j];
    v = v + b1[j];
    h[j] = v > 0.0f ? v : 0.0f;
  }
  float mx = -INFINITY;
  for (int k = 0; k < A; ++k) {
    float acc = 0.0f;
    for (int j = 0; j < H; ++j) { float t = h[j] * W[(size_t)k * H + j]; acc = acc + t; }
    acc = acc + bp[k];
    prob[k] = acc;
    if (acc > mx) mx = acc;
  }
  float sum = 0.0f;
  for (int k = 0; k < A; ++k) { prob[k] = det_expf(prob[k] - mx); sum = sum + prob[k]; }
  for (int k = 0; k < A; ++k) prob[k] = prob[k] / sum;
}
/* Categorical(prob).sample() from one 24-bit uniform: first k with cumsum > u (free-running mode only) */
static int mlp_sample(const float* prob, int A, uint32_t x) {
  const float u = (float)(x >> 8) * (1.0f / 16777216.0f);
  float c = 0.0f;
  for (int k = 0; k < A; ++k) { c = c + prob[k]; if (c > u) return k; }
  return A - 1;
}
/* xor-butterfly sum of 32 lane partials, the order the device's shuffle reduction uses */
static float butterfly_sum(float* p) {
  for (int off = 16; off >= 1; off >>= 1) {
    float q[32];
    for (int l = 0; l < 32; ++l) q[l] = p[l] + p[l ^ off];
    for (int l = 0; l < 32; ++l) p[l] = q[l];
  }
  return p[0];
}
/* v(x) (agents.py:259-262) with h = relu(fc1(x)) already computed: lane-strided partial dot products, butterfly sum, + bias */
static float ac_value(const float* wv, float bv, const float* h, int H) {
  float part[32];
  for (int l = 0; l < 32; ++l) part[l] = 0.0f;
  for (int j = 0; j < H; ++j) { float t = h[j] * wv[j]; part[j & 31] = part[j & 31] + t; }
  float v = butterfly_sum(part);
  return v + bv;
}
static void mlp_hidden(const float* par, int H, float s, float* h) {
  const float *w1 = par, *b1 = par + H;
  for (int j = 0; j < H; ++j) { float v = s * w1[j]; v = v + b1[j]; h[j] = v > 0.0f ? v : 0.0f; }
}

/* clip_grad_norm_(parameters, 1.0) then one Adam step (torch.optim.Adam defaults, single-tensor formulas); g in state_dict order */
static void mlp_clip_adam(float* blk, const ThrlAgentSpec* sp, const float* g) {
  const int P = mlp_P(sp);
  float *par = blk, *am = blk + P, *av = blk + 2 * (size_t)P;
  int32_t* hdr = (int32_t*)(blk + 3 * (size_t)P);
  double part[32];
  for (int l = 0; l < 32; ++l) part[l] = 0.0;
  for (int i = 0; i < P; ++i) part[i & 31] += (double)g[i] * (double)g[i];
  double tot = 0.0;
  for (int l = 0; l < 32; ++l) tot += part[l];
  const float total_norm = (float)sqrt(tot);
  float coef = 1.0f / (total_norm + 1e-6f);
  if (coef > 1.0f) coef = 1.0f;
  const int step = hdr[0] + 1;
  hdr[0] = step;
  const double b1 = 0.9, b2 = 0.999;
  double pw1 = 1.0, pw2 = 1.0; /* beta^step by repeated multiplication: reproducible on the device, unlike pow() */
  for (int q2 = 0; q2 < step; ++q2) { pw1 *= b1; pw2 *= b2; }
  const double bc1 = 1.0 - pw1, bc2 = 1.0 - pw2;
  const float neg_step_size = (float)(-(sp->lr / bc1));
  const float bc2_sqrt = (float)sqrt(bc2);
  const float w1m = (float)(1.0 - b1), fb2 = (float)b2, w2 = (float)(1.0 - b2), eps = 1e-8f;
  for (int i = 0; i < P; ++i) {
    const float gi = g[i] * coef;
    float m = am[i], v = av[i];
    float d = gi - m;
    d = d * w1m;
    m = m + d; /* exp_avg.lerp_(grad, 1 - beta1) */
    v = v * fb2;
    float q = w2 * gi;
    q = q * gi;
    v = v + q; /* exp_avg_sq.mul_(beta2).addcmul_(grad, grad, value = 1 - beta2) */
    am[i] = m;
    av[i] = v;
    float den = sqrtf(v);
    den = den / bc2_sqrt;
    den = den + eps;
    float up = neg_step_size * m;
    up = up / den;
    par[i] = par[i] + up; /* param.addcdiv_(exp_avg, denom, value = -step_size) */
  }
}

/* accumulate the gradient of one sample through fc_pi: dl[k] = d loss / d logits_k; adds into g and into dh[H] */
static void mlp_back_pi(const float* par, int H, int A, const float* h, const float* dl, float* g, float* dh) {
  const float* W = par + 2 * H;
  float *gW = g + 2 * H, *gbp = g + 2 * H + (size_t)A * H;
  for (int k = 0; k < A; ++k) gbp[k] = gbp[k] + dl[k];
  for (int jh = 0; jh < H; ++jh) {
    float acc = 0.0f;
    const float hj = h[jh];
    for (int k = 0; k < A; ++k) {
      float t = dl[k] * W[(size_t)k * H + jh];
      acc = acc + t;
      float u = dl[k] * hj;
      gW[(size_t)k * H + jh] = gW[(size_t)k * H + jh] + u;
    }
    dh[jh] = acc;
  }
}
static void mlp_back_fc1(int H, const float* h, const float* dh, float s, float* g) {
  float *gw1 = g, *gb1 = g + H;
  for (int jh = 0; jh < H; ++jh) {
    if (h[jh] > 0.0f) {
      float t = dh[jh] * s;
      gw1[jh] = gw1[jh] + t;
      gb1[jh] = gb1[jh] + dh[jh];
    }
  }
}

/* entropy regulariser of the discrete agents (agents.py:187-189, 298-300): loss += c_e * (-mean_n H(pi(.|s_n))), Categorical(probs):
 * H = -sum_k p_k log p_k.  d/dlogits_k of -H_n / N is p_k (log p_k + H_n) / N (softmax Jacobian; the Categorical's renormalisation
 * of probabilities that already sum to one has no first-order effect).  Adds ce_n = c_e / N times that to dl[k]; p = pi(.|s_n).
 * A probability that underflowed to zero contributes nothing (torch clamps it before the log). */
static void mlp_entropy_grad(const float* p, int A, float ce_n, float* dl) {
  float lp[THRL_ORACLE_MAX_ACTIONS];
  float Hn = 0.0f;
  for (int k = 0; k < A; ++k) {
    lp[k] = p[k] > 0.0f ? (float)det_log((double)p[k]) : 0.0f;
    float t = p[k] * lp[k];
    Hn = Hn - t;
  }
  for (int k = 0; k < A; ++k) {
    float t = lp[k] + Hn;
    t = p[k] * t;
    t = ce_n * t;
    dl[k] = dl[k] + t;
  }
}

/* Reinforce.train_net (agents.py:170-194) on the N buffered transitions buf[(head+j) % cap] = (state, action, reward),
 * then clip_grad_norm_(1.0) and one Adam step.  blk = this agent's block of the MLP slab (include/thrl.h).
 * scratch: P + 2H + A + N floats. */
static void mlp_train(float* blk, const ThrlAgentSpec* sp, int cap, int head, int N, float* scratch) {
  const int H = sp->hidden, A = sp->actions;
  const int P = mlp_P(sp), EW = mlp_entry_words(sp);
  float* par = blk;
  float* buf = blk + 3 * (size_t)P + THRL_MLP_HEADER_WORDS;
  float *g = scratch, *h = scratch + P, *dh = h + H, *prob = dh + H, *disc = prob + A;
  for (int i = 0; i < P; ++i) g[i] = 0.0f;
  /* discounted returns over the whole buffer, newest to oldest (:177-180), float32 */
  const float gam = (float)sp->gamma;
  for (int j = N - 1; j >= 0; --j) {
    const float r = buf[(size_t)((head + j) % cap) * EW + 2];
    if (j == N - 1) disc[j] = r;
    else { float t = gam * disc[j + 1]; disc[j] = r + t; }
    buf[(size_t)((head + j) % cap) * EW + 2] = disc[j]; /* kept in the (about to be emptied) buffer, as the device does */
  }
  /* (discounted - mean) / std, unbiased std (:181) */
  double sum = 0.0;
  for (int j = 0; j < N; ++j) sum += (double)disc[j];
  const float mean = (float)(sum / (double)N);
  double ss = 0.0;
  for (int j = 0; j < N; ++j) { const double d = (double)disc[j] - (double)mean; ss += d * d; }
  const float sd = (float)sqrt(ss / (double)(N - 1));
  const float invN = 1.0f / (float)N;
  const float ce = (float)sp->entropy;
  /* loss = -mean(log_prob(a) * G) (:185); d loss / d logits_k = (p_k - [k == a]) * G / N, back through fc_pi, relu, fc1 */
  for (int j = 0; j < N; ++j) {
    const float* tr = buf + (size_t)((head + j) % cap) * EW;
    const float s = tr[0];
    int32_t a;
    memcpy(&a, &tr[1], 4);
    float G = disc[j] - mean;
    G = G / sd;
    const float c = G * invN;
    mlp_forward(par, H, A, s, h, prob);
    if (ce != 0.0f) {
      float pk[THRL_ORACLE_MAX_ACTIONS];
      for (int k = 0; k < A; ++k) pk[k] = prob[k];
      for (int k = 0; k < A; ++k) { float dl = prob[k] - (k == a ? 1.0f : 0.0f); prob[k] = dl * c; }
      mlp_entropy_grad(pk, A, ce * invN, prob);
    } else {
      for (int k = 0; k < A; ++k) { float dl = prob[k] - (k == a ? 1.0f : 0.0f); prob[k] = dl * c; }
    }
    mlp_back_pi(par, H, A, h, prob, g, dh);
    mlp_back_fc1(H, h, dh, s, g);
  }
  mlp_clip_adam(blk, sp, g);
}

/* ActorCritic.train_net (agents.py:280-305).  The reference's `rewards` is reshaped to [N] while v, v' stay [N,1], so
 * advantage = rewards + gamma*v' - v broadcasts to [N,N]: adv[i][j] = r_j + d_i with d_i = gamma*v'(s'_i) - v(s_i); critic
 * loss adv^2, actor loss -log_prob(a_j)*adv[i][j].detach(), loss = mean over the N^2 entries (entropy coefficient 0).
 * The N^2 sums collapse:  dL/dd_i = 2 (R + N d_i) / N^2  (R = sum r_j; v' is NOT detached, so dL/dv'_i = gamma * that,
 * dL/dv_i = - that), and dL/dlogits_jk = (p_jk - [k == a_j]) (N r_j + D) / N^2  (D = sum d_i). */
static void ac_train(float* blk, const ThrlAgentSpec* sp, int cap, int head, int N, float* scratch) {
  const int H = sp->hidden, A = sp->actions;
  const int P = mlp_P(sp), EW = mlp_entry_words(sp);
  float* par = blk;
  const float* buf = blk + 3 * (size_t)P + THRL_MLP_HEADER_WORDS;
  const float *wv = par + 2 * H + (size_t)A * H + A, bv = par[2 * H + (size_t)A * H + A + H];
  float *g = scratch, *h = scratch + P, *dh = h + H, *prob = dh + H, *dval = prob + A;
  float *gwv = g + 2 * H + (size_t)A * H + A, *gbv = gwv + H;
  for (int i = 0; i < P; ++i) g[i] = 0.0f;
  const float gam = (float)sp->gamma;
  double R = 0.0, D = 0.0;
  for (int i = 0; i < N; ++i) { /* d_i = gamma * v(s'_i) - v(s_i) (:289) */
    const float* tr = buf + (size_t)((head + i) % cap) * EW;
    mlp_hidden(par, H, tr[0], h);
    const float v = ac_value(wv, bv, h, H);
    mlp_hidden(par, H, tr[3], h);
    const float vp = ac_value(wv, bv, h, H);
    float t = gam * vp;
    dval[i] = t - v;
    R += (double)tr[2];
    D += (double)dval[i];
  }
  const float fN = (float)N, fR = (float)R, fD = (float)D;
  const float invN2 = 1.0f / (fN * fN);
  const float ce = (float)sp->entropy, invN = 1.0f / fN;
  for (int j = 0; j < N; ++j) {
    const float* tr = buf + (size_t)((head + j) % cap) * EW;
    const float s = tr[0], s2 = tr[3];
    int32_t a;
    memcpy(&a, &tr[1], 4);
    float ca = fN * tr[2];
    ca = ca + fD;
    ca = ca * invN2; /* actor weight (N r_j + D) / N^2 */
    float cv = fN * dval[j];
    cv = fR + cv;
    cv = cv * invN2;
    cv = -2.0f * cv; /* dL/dv_j = -2 (R + N d_j) / N^2 */
    const float cvp = (-gam) * cv; /* dL/dv'_j */
    mlp_forward(par, H, A, s, h, prob);
    if (ce != 0.0f) {
      float pk[THRL_ORACLE_MAX_ACTIONS];
      for (int k = 0; k < A; ++k) pk[k] = prob[k];
      for (int k = 0; k < A; ++k) { float dl = prob[k] - (k == a ? 1.0f : 0.0f); prob[k] = dl * ca; }
      mlp_entropy_grad(pk, A, ce * invN, prob);
    } else {
      for (int k = 0; k < A; ++k) { float dl = prob[k] - (k == a ? 1.0f : 0.0f); prob[k] = dl * ca; }
    }
    mlp_back_pi(par, H, A, h, prob, g, dh);
    for (int jh = 0; jh < H; ++jh) { /* value head at s_j shares h with the policy head */
      float t = cv * h[jh];
      gwv[jh] = gwv[jh] + t;
      float u = cv * wv[jh];
      dh[jh] = dh[jh] + u;
    }
    gbv[0] = gbv[0] + cv;
    mlp_back_fc1(H, h, dh, s, g);
    mlp_hidden(par, H, s2, h); /* value head at s'_j */
    for (int jh = 0; jh < H; ++jh) {
      float t = cvp * h[jh];
      gwv[jh] = gwv[jh] + t;
      dh[jh] = cvp * wv[jh];
    }
    gbv[0] = gbv[0] + cvp;
    mlp_back_fc1(H, h, dh, s2, g);
  }
  mlp_clip_adam(blk, sp, g);
}

/* ---------------------------------------------------------------- CAC agent (th_rl/agents.py:333-417), float32 / f64 mix
 * Transcendentals are our own deterministic sequences (exp: det_expf; log: det_log in f64) so host and device agree. */
static float det_tanhf(float x) {
  const float ax = fabsf(x);
  float t;
  if (ax < 0.1f) { /* odd series: x - x^3/3 + 2x^5/15 - 17x^7/315 */
    const float x2 = ax * ax;
    float p = -17.0f / 315.0f;
    p = p * x2 + 2.0f / 15.0f;
    p = p * x2 - 1.0f / 3.0f;
    p = p * x2;
    p = p * ax;
    t = ax + p;
  } else {
    const float e = det_expf(-2.0f * ax);
    t = (1.0f - e) / (1.0f + e);
  }
  return x < 0.0f ? -t : t;
}
static float det_sigmoidf(float x) { /* 1 / (1 + exp(-x)) */
  if (x >= 0.0f) { const float e = det_expf(-x); return 1.0f / (1.0f + e); }
  const float e = det_expf(x);
  return e / (1.0f + e);
}
static float det_softplusf(float x) { /* F.softplus: x if x > 20 else log1p(exp(x)) */
  if (x > 20.0f) return x;
  const float e = det_expf(-fabsf(x));
  const float l = (float)det_log(1.0 + (double)e);
  return (x > 0.0f ? x : 0.0f) + l;
}
/* CAC.pi (agents.py:362-366) and v at one state; h must hold relu(fc1(s)).  par: w1 b1 | wmu bmu | wsd bsd | wv bv */
static void cac_heads(const float* par, int H, const float* h, float* zmu, float* zsd, float* v) {
  const float *wmu = par + 2 * H, *wsd = wmu + H + 1, *wv = wsd + H + 1;
  *zmu = ac_value(wmu, wmu[H], h, H);
  *zsd = ac_value(wsd, wsd[H], h, H);
  *v = ac_value(wv, wv[H], h, H);
}
/* sample_action (agents.py:374-378): sigmoid(Normal(mu, std).sample()); z = standard normal deviate */
static float cac_action(const float* par, int H, float s, float* h, double z) {
  float zmu, zsd, v;
  mlp_hidden(par, H, s, h);
  cac_heads(par, H, h, &zmu, &zsd, &v);
  const float mu = 4.0f * det_tanhf(zmu), sd = det_softplusf(zsd);
  float raw = sd * (float)z;
  raw = mu + raw;
  return det_sigmoidf(raw);
}
/* CAC.train_net (agents.py:391-417).  As in ActorCritic, `rewards` is [N] while mu, std, v, v' are [N,1]: adv[i][j] = r_j + d_i,
 * log_prob[i][j] = logN(l_j; mu_i, sd_i) with l_j = logit(5e-5 + (1 - 1e-4) a_j); loss = mean_ij(adv^2 - log_prob * adv.detach()).
 * With the moments Sr = sum r, Sl = sum l, Sl2 = sum l^2, Srl = sum r l, Srl2 = sum r l^2 (f64):
 *   dL/dd_i  = 2 (Sr + N d_i) / N^2
 *   dL/dmu_i = -[Srl - mu Sr + d (Sl - N mu)] / (N^2 sd^2)
 *   dL/dsd_i = -[(Srl2 - 2 mu Srl + mu^2 Sr + d (Sl2 - 2 mu Sl + N mu^2)) / sd^3 - (Sr + N d) / sd] / N^2 */
static void cac_train(float* blk, const ThrlAgentSpec* sp, int cap, int head, int N, float* scratch) {
  const int H = sp->hidden;
  const int P = mlp_P(sp), EW = mlp_entry_words(sp);
  float* par = blk;
  const float* buf = blk + 3 * (size_t)P + THRL_MLP_HEADER_WORDS;
  const float *wmu = par + 2 * H, *wsd = wmu + H + 1, *wv = wsd + H + 1;
  float *g = scratch, *h = scratch + P, *dh = h + H;
  float *gwmu = g + 2 * H, *gwsd = gwmu + H + 1, *gwv = gwsd + H + 1;
  for (int i = 0; i < P; ++i) g[i] = 0.0f;
  const float gam = (float)sp->gamma;
  double Sr = 0.0, Sl = 0.0, Sl2 = 0.0, Srl = 0.0, Srl2 = 0.0;
  for (int j = 0; j < N; ++j) {
    const float* tr = buf + (size_t)((head + j) % cap) * EW;
    float a_ = (1.0f - 1e-4f) * tr[1];
    a_ = 5e-5f + a_;
    const float ratio = a_ / (1.0f - a_);
    const double l = (double)(float)det_log((double)ratio), r = (double)tr[2];
    Sr += r; Sl += l; Sl2 += l * l; Srl += r * l; Srl2 += r * l * l;
  }
  const double dN = (double)N, invN2 = 1.0 / (dN * dN);
  for (int i = 0; i < N; ++i) {
    const float* tr = buf + (size_t)((head + i) % cap) * EW;
    const float s = tr[0], s2 = tr[3];
    float zmu, zsd, v, z2, z3, vp;
    mlp_hidden(par, H, s2, h);
    cac_heads(par, H, h, &z2, &z3, &vp);
    mlp_hidden(par, H, s, h);
    cac_heads(par, H, h, &zmu, &zsd, &v);
    const float t = det_tanhf(zmu), mu = 4.0f * t, sd = det_softplusf(zsd);
    float d = gam * vp;
    d = d - v;
    const double dd = (double)d, dmu = (double)mu, dsd = (double)sd;
    const double A0 = Sr + dN * dd;
    const double A1 = Srl - dmu * Sr + dd * (Sl - dN * dmu);
    const double A2 = Srl2 - 2.0 * dmu * Srl + dmu * dmu * Sr + dd * (Sl2 - 2.0 * dmu * Sl + dN * dmu * dmu);
    const float gmu = (float)(-(A1 / (dsd * dsd)) * invN2);
    float gsd = (float)(-(A2 / (dsd * dsd * dsd) - A0 / dsd) * invN2);
    if (sp->entropy != 0.0) { /* + c_e * (-mean Normal(mu, sd).entropy()) (:410-412): d/dsd = -c_e / (N sd) */
      const float ge = (float)(-(sp->entropy / (dN * dsd)));
      gsd = gsd + ge;
    }
    const float cv = (float)(-2.0 * A0 * invN2); /* dL/dv_i */
    const float cvp = (-gam) * cv;               /* dL/dv'_i */
    float dzmu = 1.0f - t * t;
    dzmu = 4.0f * dzmu;
    dzmu = gmu * dzmu;                            /* through mu = 4 tanh(z) */
    const float dzsd = gsd * det_sigmoidf(zsd);   /* through softplus */
    for (int jh = 0; jh < H; ++jh) {
      const float hj = h[jh];
      float u = dzmu * hj; gwmu[jh] = gwmu[jh] + u;
      u = dzsd * hj; gwsd[jh] = gwsd[jh] + u;
      u = cv * hj; gwv[jh] = gwv[jh] + u;
      float acc = dzmu * wmu[jh];
      u = dzsd * wsd[jh]; acc = acc + u;
      u = cv * wv[jh]; acc = acc + u;
      dh[jh] = acc;
    }
    gwmu[H] = gwmu[H] + dzmu;
    gwsd[H] = gwsd[H] + dzsd;
    gwv[H] = gwv[H] + cv;
    mlp_back_fc1(H, h, dh, s, g);
    mlp_hidden(par, H, s2, h); /* value head at s'_i */
    for (int jh = 0; jh < H; ++jh) {
      float u = cvp * h[jh];
      gwv[jh] = gwv[jh] + u;
      dh[jh] = cvp * wv[jh];
    }
    gwv[H] = gwv[H] + cvp;
    mlp_back_fc1(H, h, dh, s2, g);
  }
  mlp_clip_adam(blk, sp, g);
}

typedef struct Transition { /* buffers.py Experience(state, action, reward, done, new_state); `done` is never read */
  double state, reward, new_state;
  int action;
} Transition;

typedef struct Buffer { /* deque(maxlen=capacity) */
  Transition* item;
  int cap, len, head; /* head = index of the oldest element */
} Buffer;
static void buf_append(Buffer* b, Transition t) {
  if (b->cap <= 0) return;
  if (b->len < b->cap) {
    b->item[(b->head + b->len) % b->cap] = t;
    b->len++;
  } else { /* full: drop the oldest (deque maxlen) */
    b->item[b->head] = t;
    b->head = (b->head + 1) % b->cap;
  }
}

static int64_t fx_round(double x) { return (int64_t)llrint(x); }

/* One run, epochs [e0,e1).  Returns 0, or -1 on a row index outside the table (the reference's IndexError). */
static int scan_run(const ThrlScanArgs* A, int64_t r) {
  const ThrlGame* G = A->game;
  const int n = G->n_agents, T = G->max_steps, E = A->epoch_end - A->epoch_begin;
  Table tab[THRL_MAX_AGENTS];
  Buffer buf[THRL_MAX_AGENTS];
  double alpha[THRL_MAX_AGENTS], gamma[THRL_MAX_AGENTS], eps_end[THRL_MAX_AGENTS], eps_step[THRL_MAX_AGENTS];
  const size_t esz = A->table_dtype == THRL_F64 ? 8 : 4;
  int rc = 0;
  float* mlp_blk[THRL_MAX_AGENTS];
  float* mlp_scratch = NULL;
  size_t mlp_scratch_n = 0;
  for (int i = 0; i < n; ++i) {
    const ThrlAgentSpec* s = &G->agent[i];
    mlp_blk[i] = NULL;
    buf[i].item = NULL;
    if (s->kind != THRL_AGENT_QTABLE) {
      mlp_blk[i] = A->mlp + (size_t)r * G->mlp_stride + s->mlp_offset;
      const size_t P = (size_t)mlp_P(s);
      const size_t need = 2 * P + 2 * (size_t)s->hidden + s->actions + (size_t)G->mlp_buffer_len[i] + 8;
      if (need > mlp_scratch_n) mlp_scratch_n = need;
      buf[i].cap = buf[i].len = buf[i].head = 0;
      alpha[i] = gamma[i] = eps_end[i] = eps_step[i] = 0.0;
      continue;
    }
    tab[i].dtype = A->table_dtype;
    tab[i].base = (char*)A->q + ((size_t)r * G->run_stride + s->table_offset) * esz;
    tab[i].rows = s->states + 1;
    tab[i].cols = s->actions;
    tab[i].stride = s->row_stride;
    buf[i].cap = s->capacity;
    buf[i].len = buf[i].head = 0;
    buf[i].item = (Transition*)malloc(sizeof(Transition) * (size_t)(s->capacity > 0 ? s->capacity : 1));
    if (A->hp) {
      const double* h = A->hp + ((size_t)r * n + i) * 4;
      alpha[i] = h[0]; gamma[i] = h[1]; eps_end[i] = h[2]; eps_step[i] = h[3];
    } else {
      alpha[i] = s->alpha; gamma[i] = s->gamma; eps_end[i] = s->eps_end; eps_step[i] = s->eps_step;
    }
  }
  double* eps = A->eps + (size_t)r * n;
  double price = A->price[r]; /* trainer.py:45 state = environment.reset(), carried over all epochs */
  const uint64_t gid = (uint64_t)(A->run_id0 + r);
  const uint32_t k0 = (uint32_t)A->seed, k1 = (uint32_t)(A->seed >> 32);
  int act[THRL_MAX_AGENTS];
  double xs[THRL_MAX_AGENTS], Aq[THRL_MAX_AGENTS], rew[THRL_MAX_AGENTS];
  double* old_value = NULL;
  int64_t *st_row = NULL, *ns_row = NULL;
  int maxcap = 1;
  for (int i = 0; i < n; ++i) if (G->agent[i].kind == THRL_AGENT_QTABLE && G->agent[i].capacity > maxcap) maxcap = G->agent[i].capacity;
  if (mlp_scratch_n) mlp_scratch = (float*)malloc(sizeof(float) * mlp_scratch_n);
  old_value = (double*)malloc(sizeof(double) * (size_t)maxcap);
  st_row = (int64_t*)malloc(sizeof(int64_t) * (size_t)maxcap);
  ns_row = (int64_t*)malloc(sizeof(int64_t) * (size_t)maxcap);

  for (int e = 0; e < E && rc == 0; ++e) {
    const int eabs = A->epoch_begin + e;
    double rlog[THRL_MAX_AGENTS], alog[THRL_MAX_AGENTS];
    for (int i = 0; i < n; ++i) rlog[i] = alog[i] = 0.0; /* trainer.py:40-41 zeros */
    for (int t = 0; t < T && rc == 0; ++t) { /* trainer.py:50 while not done (episode counts to max_steps) */
      const size_t sidx = ((size_t)r * E + e) * T + t;
      /* trainer.py:52-55, agent order 0..n-1; agents.py:80-89 */
      for (int i = 0; i < n; ++i) {
        const ThrlAgentSpec* s = &G->agent[i];
        int k;
        if (s->kind != THRL_AGENT_QTABLE) {
          /* Reinforce.sample_action (agents.py:160-163): in both replay modes the recorded sample is forced (it came from
           * torch's generator); free running: inverse-CDF on one Philox word */
          if (s->kind == THRL_AGENT_CAC) {
            /* CAC.sample_action (agents.py:374-378): a float32 in (0,1); streams carry its bit pattern (never negative) */
            float af;
            if (A->rng_mode != THRL_RNG_PHILOX && A->replay_ra[sidx * n + i] >= 0) {
              memcpy(&af, &A->replay_ra[sidx * n + i], 4);
            } else {
              uint32_t x[4];
              philox4x32_10((uint32_t)gid, (uint32_t)eabs, (uint32_t)t, (uint32_t)(i >> 1) | (STREAM_ACT << 16), k0, k1, x);
              const uint64_t m = ((uint64_t)x[2 * (i & 1)] << 21) | (x[2 * (i & 1) + 1] >> 11);
              const double z = det_norminv(((double)m + 0.5) * (1.0 / 9007199254740992.0));
              af = cac_action(mlp_blk[i], s->hidden, (float)price, mlp_scratch, z);
            }
            memcpy(&act[i], &af, 4);
            xs[i] = (double)af * (s->action_hi - s->action_lo) + s->action_lo; /* CAC.scale (agents.py:368-372) */
            continue;
          }
          if (A->rng_mode != THRL_RNG_PHILOX && A->replay_ra[sidx * n + i] >= 0) {
            k = A->replay_ra[sidx * n + i];
          } else { /* free running, or a replay stream that leaves this agent's sample to the device (negative entry) */
            float* h = mlp_scratch;
            float* prob = h + s->hidden;
            uint32_t x[4];
            mlp_forward(mlp_blk[i], s->hidden, s->actions, (float)price, h, prob);
            philox4x32_10((uint32_t)gid, (uint32_t)eabs, (uint32_t)t, (uint32_t)(i >> 1) | (STREAM_ACT << 16), k0, k1, x);
            k = mlp_sample(prob, s->actions, x[2 * (i & 1)]);
          }
          if (k < 0 || k >= s->actions) { rc = -1; break; }
          act[i] = k;
          /* Reinforce.scale (agents.py:154-158): action / actions, not / (actions - 1) */
          xs[i] = (double)k / (double)s->actions * (s->action_hi - s->action_lo) + s->action_lo;
          continue;
        }
        if (A->rng_mode == THRL_RNG_REPLAY_ACTIONS) {
          k = A->replay_ra[sidx * n + i];
        } else {
          double u;
          int ra;
          if (A->rng_mode == THRL_RNG_REPLAY_DRAWS) {
            u = A->replay_u[sidx * n + i];
            ra = A->replay_ra[sidx * n + i];
          } else {
            /* one Philox call serves agents 2p and 2p+1: words (0,1) / (2,3) = (uniform, random action) */
            uint32_t x[4];
            philox4x32_10((uint32_t)gid, (uint32_t)eabs, (uint32_t)t, (uint32_t)(i >> 1) | (STREAM_ACT << 16), k0, k1, x);
            u = (double)x[2 * (i & 1)] * (1.0 / 4294967296.0);
            ra = (int)(((uint64_t)x[2 * (i & 1) + 1] * (uint64_t)s->actions) >> 32);
          }
          if (u < eps[i]) {
            k = ra;
          } else {
            int64_t row = act_row(price, s);
            if (row < 0 || row > s->states) { rc = -1; break; }
            k = row_argmax(&tab[i], row);
          }
        }
        if (k < 0 || k >= s->actions) { rc = -1; break; }
        act[i] = k;
        xs[i] = scale_action(k, s); /* trainer.py:56 */
      }
      if (rc) break;
      /* environments.py:25-39 */
      const double ab = G->a / G->b;
      for (int i = 0; i < n; ++i) Aq[i] = ab * xs[i];
      const double Q = py_sum_quantities(Aq, n, lead_exact_floats(G));
      double new_a;
      if (A->rng_mode == THRL_RNG_PHILOX) {
        new_a = G->a;
        if (G->noise_prob > 0.0) {
          uint32_t x[4];
          philox4x32_10((uint32_t)gid, (uint32_t)eabs, (uint32_t)t, (uint32_t)(STREAM_ENV << 16), k0, k1, x);
          double un = u53(x[0], x[1]);
          if (un < G->noise_prob) {
            double lo = G->a * 0.7;
            new_a = lo + (G->a - lo) * u53(x[2], x[3]); /* numpy uniform(low, high) = low + (high-low)*U */
          }
        }
      } else {
        new_a = A->replay_new_a ? A->replay_new_a[sidx] : G->a;
      }
      double pn = new_a - G->b * Q;
      double next_price = pn > 0.0 ? pn : 0.0; /* numpy.max([0, x]) */
      if (pn != pn) next_price = pn;           /* NaN propagates through numpy.max */
      for (int i = 0; i < n; ++i) rew[i] = next_price * Aq[i];
      /* trainer.py:61-62 */
      for (int i = 0; i < n; ++i) {
        if (G->agent[i].kind != THRL_AGENT_QTABLE) {
          /* memory.append; replay(cast) turns state and reward into float32 (buffers.py:28-38, agents.py:142) */
          const int cap = G->mlp_buffer_len[i];
          if (cap > 0) {
            const size_t P = (size_t)mlp_P(&G->agent[i]);
            const int EW = mlp_entry_words(&G->agent[i]);
            int32_t* hdr = (int32_t*)(mlp_blk[i] + 3 * P);
            float* mb = mlp_blk[i] + 3 * P + THRL_MLP_HEADER_WORDS;
            int len = hdr[1], head = hdr[2];
            int slot;
            if (len < cap) { slot = (head + len) % cap; len++; }
            else { slot = head; head = (head + 1) % cap; }
            const float sf = (float)price, rf = (float)rew[i];
            const int32_t ai = act[i];
            mb[(size_t)slot * EW] = sf;
            memcpy(&mb[(size_t)slot * EW + 1], &ai, 4);
            mb[(size_t)slot * EW + 2] = rf;
            if (EW == 4) mb[(size_t)slot * EW + 3] = (float)next_price;
            hdr[1] = len; hdr[2] = head;
          }
          continue;
        }
        Transition tr = {price, rew[i], next_price, act[i]};
        buf_append(&buf[i], tr);
      }
      /* trainer.py:65-66: divide, then add */
      for (int i = 0; i < n; ++i) {
        rlog[i] += rew[i] / (double)T;
        alog[i] += xs[i] / (double)T;
      }
      if (A->trace_actions) for (int i = 0; i < n; ++i) A->trace_actions[sidx * n + i] = act[i];
      if (A->trace_rewards) for (int i = 0; i < n; ++i) A->trace_rewards[sidx * n + i] = rew[i];
      if (A->trace_prices) A->trace_prices[sidx] = next_price;
      price = next_price; /* trainer.py:67 */
    }
    if (rc) break;
    /* trainer.py:70, agents.py:59-78 */
    for (int i = 0; i < n && rc == 0; ++i) {
      const ThrlAgentSpec* s = &G->agent[i];
      Buffer* b = &buf[i];
      if (s->kind != THRL_AGENT_QTABLE) {
        const size_t P = (size_t)mlp_P(s);
        int32_t* hdr = (int32_t*)(mlp_blk[i] + 3 * P);
        if (G->mlp_buffer_len[i] > 0 && hdr[1] >= s->min_memory) { /* agents.py:171 / :281 */
          if (s->kind == THRL_AGENT_CAC) cac_train(mlp_blk[i], s, G->mlp_buffer_len[i], hdr[2], hdr[1], mlp_scratch);
          else if (s->kind == THRL_AGENT_ACTORCRITIC) ac_train(mlp_blk[i], s, G->mlp_buffer_len[i], hdr[2], hdr[1], mlp_scratch);
          else mlp_train(mlp_blk[i], s, G->mlp_buffer_len[i], hdr[2], hdr[1], mlp_scratch);
          hdr[1] = 0; hdr[2] = 0; /* :194 memory.empty() */
        }
        continue;
      }
      if (b->len >= s->min_memory) {
        const int L = b->len;
        for (int j = 0; j < L; ++j) { /* :62,:66 encodes, :67 snapshot */
          const Transition* tr = &b->item[(b->head + j) % b->cap];
          st_row[j] = upd_row(tr->state, s);
          ns_row[j] = upd_row(tr->new_state, s);
          if (st_row[j] < 0 || st_row[j] > s->states || ns_row[j] < 0 || ns_row[j] > s->states) { rc = -1; break; }
          old_value[j] = tget(&tab[i], st_row[j], tr->action);
        }
        if (rc) break;
        uint32_t* cnt = A->counter ? A->counter + (size_t)r * G->run_stride + s->table_offset : NULL;
        for (int j = 0; j < L; ++j) { /* :68-76 */
          const Transition* tr = &b->item[(b->head + j) % b->cap];
          double next_max = row_max(&tab[i], ns_row[j]);
          double new_value = (1.0 - alpha[i]) * old_value[j] + alpha[i] * (tr->reward + gamma[i] * next_max);
          tset(&tab[i], st_row[j], tr->action, new_value);
          if (cnt) cnt[st_row[j] * s->row_stride + tr->action] += 1;
        }
        b->len = 0; b->head = 0; /* :77 */
      }
      eps[i] = eps_end[i] + (eps[i] - eps_end[i]) * eps_step[i]; /* :78, every epoch */
    }
    if (rc) break;
    if (r < A->n_log_runs) {
      for (int i = 0; i < n; ++i) {
        if (A->rewards_log) A->rewards_log[((size_t)r * E + e) * n + i] = rlog[i];
        if (A->actions_log) A->actions_log[((size_t)r * E + e) * n + i] = alog[i];
      }
    }
    if (A->stats) {
      for (int i = 0; i < n; ++i) {
        int64_t* s4 = A->stats + ((size_t)e * n + i) * THRL_STATS_K;
        int64_t v0 = fx_round(rlog[i] * THRL_STATS_SCALE_SUM), v1 = fx_round(rlog[i] * rlog[i] * THRL_STATS_SCALE_SQ);
        int64_t v2 = fx_round(alog[i] * THRL_STATS_SCALE_SUM), v3 = fx_round(alog[i] * alog[i] * THRL_STATS_SCALE_SQ);
        __atomic_fetch_add(&s4[0], v0, __ATOMIC_RELAXED); /* integer sums: exact in any order */
        __atomic_fetch_add(&s4[1], v1, __ATOMIC_RELAXED);
        __atomic_fetch_add(&s4[2], v2, __ATOMIC_RELAXED);
        __atomic_fetch_add(&s4[3], v3, __ATOMIC_RELAXED);
      }
    }
  }
  A->price[r] = price;
  for (int i = 0; i < n; ++i) free(buf[i].item);
  free(old_value); free(st_row); free(ns_row); free(mlp_scratch);
  return rc;
}

typedef struct Worker {
  const ThrlScanArgs* args;
  int64_t* next; /* shared run cursor */
  int bad;
} Worker;
static void* worker_main(void* p) {
  Worker* w = (Worker*)p;
  for (;;) {
    int64_t r0 = __atomic_fetch_add(w->next, 16, __ATOMIC_RELAXED);
    if (r0 >= w->args->n_runs) break;
    int64_t r1 = r0 + 16 < w->args->n_runs ? r0 + 16 : w->args->n_runs;
    for (int64_t r = r0; r < r1; ++r) w->bad |= scan_run(w->args, r) != 0;
  }
  return NULL;
}

/* Same meaning as thrl_qtable_scan (include/thrl.h) on HOST pointers.  Buffers start empty (args->ring is ignored),
 * so call it once over the whole epoch range.  Runs are independent, so they are spread over n_threads pthreads
 * (n_threads <= 0: one per online core).  Returns 0 / THRL_ERR_BAD_CONFIG. */
int thrl_oracle_qtable_scan(const ThrlScanArgs* args, int n_threads) {
  if (n_threads <= 0) n_threads = (int)sysconf(_SC_NPROCESSORS_ONLN);
  if (n_threads < 1) n_threads = 1;
  if (n_threads > 256) n_threads = 256;
  if ((int64_t)n_threads > args->n_runs) n_threads = (int)(args->n_runs > 0 ? args->n_runs : 1);
  int64_t next = 0;
  Worker w[256];
  pthread_t th[256];
  for (int i = 0; i < n_threads; ++i) { w[i].args = args; w[i].next = &next; w[i].bad = 0; }
  for (int i = 1; i < n_threads; ++i) pthread_create(&th[i], NULL, worker_main, &w[i]);
  worker_main(&w[0]);
  int bad = w[0].bad;
  for (int i = 1; i < n_threads; ++i) { pthread_join(th[i], NULL); bad |= w[i].bad; }
  return bad ? THRL_ERR_BAD_CONFIG : THRL_OK;
}
/* test hook: the restated builtin sum() (tests compare it with the interpreter's own sum) */
double thrl_oracle_py_sum(const double* items, int n, int lead) { return py_sum_quantities(items, n, lead); }

int thrl_oracle_online_cores(void) { return (int)sysconf(_SC_NPROCESSORS_ONLN); }

/* DESIGN.md "Device init": q = 12.5/(1-gamma) + N(0,1) (agents.py:29), counter = 0 (agents.py:45),
 * price ~ U(0,a) (environments.py:15-16), eps = eps0. */
int thrl_oracle_game_init(const ThrlGame* G, int64_t n_runs, int64_t run_id0, uint64_t seed, int32_t table_dtype,
                          const double* hp, const double* eps0, void* q, uint32_t* counter, double* eps, double* price,
                          float* mlp);
int thrl_oracle_qtable_init(const ThrlGame* G, int64_t n_runs, int64_t run_id0, uint64_t seed, int32_t table_dtype,
                            const double* hp, const double* eps0, void* q, uint32_t* counter, double* eps,
                            double* price) {
  return thrl_oracle_game_init(G, n_runs, run_id0, seed, table_dtype, hp, eps0, q, counter, eps, price, NULL);
}
int thrl_oracle_game_init(const ThrlGame* G, int64_t n_runs, int64_t run_id0, uint64_t seed, int32_t table_dtype,
                          const double* hp, const double* eps0, void* q, uint32_t* counter, double* eps, double* price,
                          float* mlp) {
  const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
  const int n = G->n_agents;
  for (int64_t r = 0; r < n_runs; ++r) {
    const uint64_t gid = (uint64_t)(run_id0 + r);
    for (int i = 0; i < n; ++i) {
      const ThrlAgentSpec* s = &G->agent[i];
      if (s->kind != THRL_AGENT_QTABLE) { /* nn.Linear default init, the rest of the block zero */
        const int64_t Pn = mlp_P(s);
        const int64_t words = 3 * Pn + THRL_MLP_HEADER_WORDS + (int64_t)mlp_entry_words(s) * G->mlp_buffer_len[i];
        float* blk = mlp + (size_t)r * G->mlp_stride + s->mlp_offset;
        const float b_fc1 = 1.0f, b_pi = (float)(1.0 / sqrt((double)s->hidden));
        for (int64_t w = 0; w < words; ++w) {
          float v = 0.0f;
          if (w < Pn) {
            uint32_t x[4];
            philox4x32_10((uint32_t)gid, (uint32_t)(w >> 1), (uint32_t)i, (uint32_t)(STREAM_INIT_Q << 16), k0, k1, x);
            const double u = (w & 1) ? u53(x[2], x[3]) : u53(x[0], x[1]);
            const float bound = (w < 2 * (int64_t)s->hidden) ? b_fc1 : b_pi;
            v = (float)(2.0 * u - 1.0) * bound;
            if (s->kind == THRL_AGENT_ACTORCRITIC && w == Pn - 1) v = 1000.0f; /* fc_v.bias.data.fill_(1000.0), agents.py:244 */
            /* (CAC does not touch fc_v.bias: agents.py:349-352) */
          }
          blk[w] = v;
        }
        eps[(size_t)r * n + i] = eps0[i];
        continue;
      }
      const double gamma = hp ? hp[((size_t)r * n + i) * 4 + 1] : s->gamma;
      const double base = 12.5 / (1.0 - gamma);
      const int64_t cells = (int64_t)(s->states + 1) * s->actions;
      for (int64_t c = 0; c < cells; c += 2) { /* one Philox call -> two normals */
        uint32_t x[4];
        philox4x32_10((uint32_t)gid, (uint32_t)(c >> 1), (uint32_t)i, (uint32_t)(STREAM_INIT_Q << 16), k0, k1, x);
        for (int h = 0; h < 2 && c + h < cells; ++h) {
          /* (m + 0.5) * 2^-53 lies strictly inside (0,1) */
          uint64_t m = ((uint64_t)x[2 * h] << 21) | (x[2 * h + 1] >> 11);
          double p = ((double)m + 0.5) * (1.0 / 9007199254740992.0);
          double v = base + det_norminv(p);
          /* the draw of a cell depends on its logical index row * actions + col only, not on the row padding */
          size_t idx = (size_t)r * G->run_stride + s->table_offset + (size_t)((c + h) / s->actions) * s->row_stride + (size_t)((c + h) % s->actions);
          if (table_dtype == THRL_F64) ((double*)q)[idx] = v;
          else ((float*)q)[idx] = (float)v;
          if (counter) counter[idx] = 0;
        }
      }
      eps[(size_t)r * n + i] = eps0[i];
    }
    uint32_t x[4];
    philox4x32_10((uint32_t)gid, 0, 0, (uint32_t)(STREAM_INIT_P << 16), k0, k1, x);
    price[r] = 0.0 + (G->a - 0.0) * u53(x[0], x[1]);
  }
  return THRL_OK;
}

/* utils.py:27-47 play_game with each agent's get_action: QTable agents.py:91-92 (float64 encode of the state, first argmax);
 * Reinforce / ActorCritic agents.py:165-168 / :275-278 (argmax of pi(float32 state)); CAC agents.py:380-384
 * (Normal(mu, 0).sample() == mu, so the action is sigmoid(4 tanh(fc_mu(h)))).  No exploration, no update.
 * price0[r][it] replaces environment.reset()'s draw.  rewards/actions: [R][iters*T][n]. */
int thrl_oracle_greedy_eval_noise(const ThrlGame* G, int64_t n_runs, int32_t table_dtype, const void* q, const float* mlp,
                                  int32_t iters, const double* price0, const double* new_a, double* rewards, double* actions);
int thrl_oracle_greedy_eval_mlp(const ThrlGame* G, int64_t n_runs, int32_t table_dtype, const void* q, const float* mlp,
                                int32_t iters, const double* price0, double* rewards, double* actions) {
  return thrl_oracle_greedy_eval_noise(G, n_runs, table_dtype, q, mlp, iters, price0, NULL, rewards, actions);
}
/* new_a [R][iters][T]: the demand intercept of every step (environments.py:28-31: a, or the redrawn value); NULL = a throughout */
int thrl_oracle_greedy_eval_noise(const ThrlGame* G, int64_t n_runs, int32_t table_dtype, const void* q, const float* mlp,
                                  int32_t iters, const double* price0, const double* new_a, double* rewards, double* actions) {
  const int n = G->n_agents, T = G->max_steps;
  const size_t esz = table_dtype == THRL_F64 ? 8 : 4;
  float hbuf[1024 + 512];
  for (int64_t r = 0; r < n_runs; ++r) {
    for (int it = 0; it < iters; ++it) {
      double price = price0[(size_t)r * iters + it];
      for (int t = 0; t < T; ++t) {
        double xs[THRL_MAX_AGENTS], Aq[THRL_MAX_AGENTS];
        const double ab = G->a / G->b;
        for (int i = 0; i < n; ++i) {
          const ThrlAgentSpec* s = &G->agent[i];
          if (s->kind == THRL_AGENT_QTABLE) {
            Table tb = {table_dtype, (char*)q + ((size_t)r * G->run_stride + s->table_offset) * esz, s->states + 1,
                        s->actions, s->row_stride};
            int64_t row = upd_row(price, s);
            if (row < 0 || row > s->states) return THRL_ERR_BAD_CONFIG;
            xs[i] = scale_action(row_argmax(&tb, row), s);
          } else {
            if (s->hidden > 1024 || s->actions > 256) return THRL_ERR_BAD_CONFIG;
            const float* par = mlp + (size_t)r * G->mlp_stride + s->mlp_offset;
            float* h = hbuf;
            float* prob = hbuf + 1024;
            if (s->kind == THRL_AGENT_CAC) {
              float zmu, zsd, v;
              mlp_hidden(par, s->hidden, (float)price, h);
              cac_heads(par, s->hidden, h, &zmu, &zsd, &v);
              const float af = det_sigmoidf(4.0f * det_tanhf(zmu));
              xs[i] = (double)af * (s->action_hi - s->action_lo) + s->action_lo;
            } else {
              mlp_forward(par, s->hidden, s->actions, (float)price, h, prob);
              int best = 0;
              for (int k = 1; k < s->actions; ++k) if (prob[k] > prob[best]) best = k; /* torch.argmax: first maximal index */
              xs[i] = (double)best / (double)s->actions * (s->action_hi - s->action_lo) + s->action_lo;
            }
          }
          Aq[i] = ab * xs[i];
        }
        const double Q = py_sum_quantities(Aq, n, lead_exact_floats(G));
        const double na = new_a ? new_a[((size_t)r * iters + it) * T + t] : G->a;
        double pn = na - G->b * Q;
        double next_price = pn > 0.0 ? pn : 0.0;
        size_t o = (((size_t)r * iters + it) * T + t) * n;
        for (int i = 0; i < n; ++i) { rewards[o + i] = next_price * Aq[i]; actions[o + i] = xs[i]; }
        price = next_price;
      }
    }
  }
  return THRL_OK;
}
int thrl_oracle_greedy_eval(const ThrlGame* G, int64_t n_runs, int32_t table_dtype, const void* q, int32_t iters,
                            const double* price0, double* rewards, double* actions) {
  return thrl_oracle_greedy_eval_mlp(G, n_runs, table_dtype, q, NULL, iters, price0, rewards, actions);
}

/* Same checks / layout as thrl_game_layout in the product, restated so the oracle stands alone. */
int thrl_oracle_game_layout(ThrlGame* G) {
  if (G->n_agents < 1 || G->n_agents > THRL_MAX_AGENTS || G->max_steps < 1) return THRL_ERR_BAD_CONFIG;
  int64_t off = 0, moff = 0, cells = 0;
  int ring = 0, regular = 1;
  for (int i = 0; i < G->n_agents; ++i)
    if (G->agent[i].kind == THRL_AGENT_QTABLE && G->agent[i].states >= 1 && G->agent[i].actions >= 1)
      cells += (int64_t)(G->agent[i].states + 1) * G->agent[i].actions;
  const int padded = cells * 4 >= THRL_PAD_THRESHOLD_BYTES; /* include/thrl.h: HBM-resident tables have rows of 4k elements */
  for (int i = 0; i < G->n_agents; ++i) {
    ThrlAgentSpec* s = &G->agent[i];
    G->mlp_buffer_len[i] = 0;
    s->row_stride = 0;
    s->reserved_ = 0;
    if (s->actions < 2 || s->actions > THRL_MAX_ACTIONS || s->capacity < 0 || s->min_memory < 0) return THRL_ERR_BAD_CONFIG;
    const int T = G->max_steps, mm = s->min_memory > 0 ? s->min_memory : 1;
    int64_t need = (int64_t)T * ((mm + T - 1) / T);
    if (need > s->capacity) need = s->capacity;
    if (s->kind == THRL_AGENT_REINFORCE || s->kind == THRL_AGENT_ACTORCRITIC || s->kind == THRL_AGENT_CAC) {
      if (s->states != 1 || s->hidden < 1) return THRL_ERR_BAD_CONFIG;
      const int64_t P = mlp_P(s);
      G->mlp_buffer_len[i] = s->min_memory <= s->capacity ? (int32_t)need : 0;
      s->mlp_offset = moff;
      s->table_offset = 0;
      moff += 3 * P + THRL_MLP_HEADER_WORDS + (int64_t)mlp_entry_words(s) * G->mlp_buffer_len[i];
      continue;
    }
    if (s->kind != THRL_AGENT_QTABLE || s->states < 1) return THRL_ERR_BAD_CONFIG;
    /* a row beyond `states` is an IndexError in the reference (agents.py:88); prices never exceed a (environments.py:28-32) */
    if (!(s->max_state > 0.0) || !(G->a == G->a)) return THRL_ERR_BAD_CONFIG;
    if (act_row(G->a, s) > s->states || upd_row(G->a, s) > s->states) return THRL_ERR_BAD_CONFIG;
    s->table_offset = off;
    s->mlp_offset = 0;
    s->row_stride = padded ? (s->actions + 3) / 4 * 4 : s->actions;
    off += (int64_t)(s->states + 1) * s->row_stride;
    if (s->min_memory <= s->capacity) { /* otherwise the update never fires and the buffer content is irrelevant */
      if (need > ring) ring = (int)need;
      if (s->min_memory > T) regular = 0;
    }
  }
  G->run_stride = off;
  G->mlp_stride = moff;
  G->ring_len = ring;
  G->regular = regular;
  return THRL_OK;
}
