// thrl_scan_mixed.cuh — games that contain MLP agents (Reinforce, th_rl/agents.py:119-194; ActorCritic, :222-305) next to
// QTable agents (both configs the reference ships pair a QTable with a Reinforce agent).
//
// Correctness-first tier: one warp plays one run; QTable tables stay in HBM and are addressed directly (no greedy cache);
// every MLP agent's parameters are staged in shared memory (fc_pi.weight transposed to [hidden][actions] so that both the
// forward pass, lanes = actions, and the backward pass, lanes = hidden units, read conflict-free).  All MLP arithmetic is
// float32 with the operation order of oracle/thrl_oracle.c (mlp_forward / mlp_train), so the two agree bit for bit; the
// oracle in turn is pinned against the reference's torch results.  Tensor cores are not used here yet: per-run weights
// make the update a batch of 1000x256x22 GEMMs per run, which is the next round's tcgen05 grouped-GEMM work.
#pragma once
#include "thrl_device.cuh"
#include "thrl_scan_generic.cuh"  // RingHeader
#include "thrl_aux_kernels.cuh"   // det_log, det_norminv

namespace thrl {

struct MixedParams {
  ThrlGame game;
  long long n_runs, run_id0;
  int epoch_begin, E, rng_mode;
  uint32_t k0, k1;
  void* q;
  uint32_t* counter;
  double* eps;
  double* price;
  const double* hp;
  const double* replay_u;
  const int32_t* replay_ra;
  const double* replay_new_a;
  double* rewards_log;
  double* actions_log;
  long long n_log_runs;
  long long* stats;
  int32_t* trace_actions;
  double* trace_rewards;
  double* trace_prices;
  float* mlp;
  unsigned char* ring;
  long long ring_bytes;
  // shared-memory layout (bytes)
  int cta_bytes, warp_bytes;
  int off_P, off_act, off_pre, off_newa, off_row, off_old, off_hp, off_par, off_grad, off_h;
  int par_off[THRL_MAX_AGENTS];  // float offset of MLP agent i's staged parameters inside off_par
  int lut_total, Hp, noisy;
};

__device__ __forceinline__ int mlp_P(const ThrlAgentSpec& s) {
  if (s.kind == THRL_AGENT_CAC) return 5 * s.hidden + 3;
  const int p = 2 * s.hidden + s.actions * s.hidden + s.actions;
  return s.kind == THRL_AGENT_ACTORCRITIC ? p + s.hidden + 1 : p;
}
__device__ __forceinline__ int mlp_entry_words(const ThrlAgentSpec& s) { return s.kind == THRL_AGENT_REINFORCE ? 3 : 4; }
// state_dict order -> staged order: fc_pi.weight [A][H] is staged transposed [H][A]; CAC has no matrix head (identity)
__device__ __forceinline__ int mlp_flat2st(const ThrlAgentSpec& s, int i) {
  const int H = s.hidden, A = s.actions;
  if (s.kind == THRL_AGENT_CAC || i < 2 * H || i >= 2 * H + A * H) return i;
  const int e = i - 2 * H, k = e / H, j = e - k * H;
  return 2 * H + j * A + k;
}

// expf with the oracle's operation sequence (Cody-Waite + Cephes polynomial)
__device__ __forceinline__ float det_expf(float x) {
  if (x < -87.0f) return 0.0f;
  const float kf = rintf(__fmul_rn(x, 1.44269504f));
  float r = __fsub_rn(x, __fmul_rn(kf, 0.693359375f));
  r = __fsub_rn(r, __fmul_rn(kf, -2.12194440e-4f));
  float p = 1.9875691500e-4f;
  p = __fadd_rn(__fmul_rn(p, r), 1.3981999507e-3f);
  p = __fadd_rn(__fmul_rn(p, r), 8.3334519073e-3f);
  p = __fadd_rn(__fmul_rn(p, r), 4.1665795894e-2f);
  p = __fadd_rn(__fmul_rn(p, r), 1.6666665459e-1f);
  p = __fadd_rn(__fmul_rn(p, r), 5.0000001201e-1f);
  p = __fmul_rn(p, r);
  p = __fmul_rn(p, r);
  p = __fadd_rn(p, r);
  p = __fadd_rn(p, 1.0f);
  // p * 2^k: one exact-or-correctly-rounded multiply wherever 2^k is a float (the same value ldexpf returns, subnormal
  // results included; the library routine costs ~40 instructions)
  const int k = (int)kf;
  if (k >= -126 && k <= 128) return __fmul_rn(p, __int_as_float((k + 127) << 23));
  return ldexpf(p, k);
}

// pi(x) (agents.py:148-152) with staged parameters sp = [w1 (H)] [b1 (H)] [WT (H x A)] [bp (A)].
// Leaves h[0..H) and prob[0..A) in shared memory (hs / ps); every lane returns after a __syncwarp.
__device__ __forceinline__ void mlp_forward_warp(const float* sp, int H, int A, float s, float* hs, float* ps, int lane) {
  const float *w1 = sp, *b1 = sp + H, *WT = sp + 2 * H, *bp = sp + 2 * H + H * A;
  for (int j = lane; j < H; j += 32) {
    const float v = __fadd_rn(__fmul_rn(s, w1[j]), b1[j]);
    hs[j] = v > 0.0f ? v : 0.0f;
  }
  __syncwarp();
  float mx = NegInf<float>::v();
  for (int k0 = 0; k0 < A; k0 += 32) {  // lane = action (several rounds when A > 32)
    const int k = k0 + lane;
    float acc = 0.0f;
    if (k < A) {
      for (int j = 0; j < H; ++j) acc = __fadd_rn(acc, __fmul_rn(hs[j], WT[j * A + k]));
      acc = __fadd_rn(acc, bp[k]);
      ps[k] = acc;
      mx = acc > mx ? acc : mx;
    }
  }
  mx = warp_max(mx);
  __syncwarp();
  for (int k = lane; k < A; k += 32) ps[k] = det_expf(__fsub_rn(ps[k], mx));
  __syncwarp();
  float sum = 0.0f;
  for (int k = 0; k < A; ++k) sum = __fadd_rn(sum, ps[k]);  // sequential order, every lane the same
  __syncwarp();
  for (int k = lane; k < A; k += 32) ps[k] = __fdiv_rn(ps[k], sum);
  __syncwarp();
}

// h = relu(fc1(x)) into shared memory
__device__ __forceinline__ void mlp_hidden_warp(const float* sp, int H, float s, float* hs, int lane) {
  const float *w1 = sp, *b1 = sp + H;
  __syncwarp();
  for (int j = lane; j < H; j += 32) {
    const float v = __fadd_rn(__fmul_rn(s, w1[j]), b1[j]);
    hs[j] = v > 0.0f ? v : 0.0f;
  }
  __syncwarp();
}
// v(x) (agents.py:259-262) from h in shared memory: lane-strided partial dot products, xor-butterfly sum, + bias (oracle ac_value)
__device__ __forceinline__ float ac_value_warp(const float* wv, float bv, const float* hs, int H, int lane) {
  float part = 0.0f;
  for (int j = lane; j < H; j += 32) part = __fadd_rn(part, __fmul_rn(hs[j], wv[j]));
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) part = __fadd_rn(part, __shfl_xor_sync(kFull, part, off));
  return __fadd_rn(part, bv);
}

// clip_grad_norm_(1.0) + one Adam step (oracle mlp_clip_adam).  gs: gradient in the staged layout.
__device__ inline void mlp_clip_adam_warp(float* blk, const ThrlAgentSpec& spec, float* sp, const float* gs, int lane) {
  const int P = mlp_P(spec);
  float *am = blk + P, *av = blk + 2 * (size_t)P;
  int32_t* hdr = reinterpret_cast<int32_t*>(blk + 3 * (size_t)P);
  auto flat2st = [&](int i) { return mlp_flat2st(spec, i); };  // state_dict order (oracle) -> staged layout
  __syncwarp();
  double part = 0.0;
  for (int i = lane; i < P; i += 32) { const double gd = (double)gs[flat2st(i)]; part = __dadd_rn(part, __dmul_rn(gd, gd)); }
  double tot = 0.0;
  for (int l = 0; l < 32; ++l) tot = __dadd_rn(tot, shfl_d(part, l));
  const float total_norm = (float)sqrt(tot);
  float coef = __fdiv_rn(1.0f, __fadd_rn(total_norm, 1e-6f));
  if (coef > 1.0f) coef = 1.0f;
  const int step = hdr[0] + 1;
  double pw1 = 1.0, pw2 = 1.0;
  for (int q2 = 0; q2 < step; ++q2) { pw1 = __dmul_rn(pw1, 0.9); pw2 = __dmul_rn(pw2, 0.999); }
  const double bc1 = __dsub_rn(1.0, pw1), bc2 = __dsub_rn(1.0, pw2);
  const float neg_step_size = (float)(-__ddiv_rn(spec.lr, bc1));
  const float bc2_sqrt = (float)sqrt(bc2);
  const float w1m = (float)__dsub_rn(1.0, 0.9), fb2 = (float)0.999, w2 = (float)__dsub_rn(1.0, 0.999), eps = 1e-8f;
  for (int i = lane; i < P; i += 32) {
    const int st = flat2st(i);
    const float gi = __fmul_rn(gs[st], coef);
    float m = am[i], v = av[i];
    m = __fadd_rn(m, __fmul_rn(__fsub_rn(gi, m), w1m));
    v = __fadd_rn(__fmul_rn(v, fb2), __fmul_rn(__fmul_rn(w2, gi), gi));
    am[i] = m;
    av[i] = v;
    const float den = __fadd_rn(__fdiv_rn(sqrtf(v), bc2_sqrt), eps);
    sp[st] = __fadd_rn(sp[st], __fdiv_rn(__fmul_rn(neg_step_size, m), den));
  }
  __syncwarp();
  if (lane == 0) hdr[0] = step;
}

// Entropy regulariser (agents.py:187-189, 298-300; oracle mlp_entropy_grad): H_n = -sum_k p_k log p_k summed in action order
// by every lane (the same float32 sequence as the oracle); lane k then adds ce_n * p_k * (log p_k + H_n) to its dl.
__device__ __forceinline__ float mlp_logp(float p) { return p > 0.0f ? (float)det_log((double)p) : 0.0f; }
__device__ __forceinline__ float mlp_entropy_warp(const float* ps, int A) {
  float Hn = 0.0f;
  for (int k = 0; k < A; ++k) {
    const float pk = ps[k];
    Hn = __fsub_rn(Hn, __fmul_rn(pk, mlp_logp(pk)));
  }
  return Hn;
}
__device__ __forceinline__ float mlp_entropy_term(float pk, float Hn, float ce_n) {
  return __fmul_rn(ce_n, __fmul_rn(pk, __fadd_rn(mlp_logp(pk), Hn)));
}

// Reinforce.train_net (agents.py:170-194) for one agent of one run; mirrors oracle mlp_train.
// blk: the agent's block in the global MLP slab; sp: staged parameters (updated in place); gs: gradient scratch [P] in the
// same staged layout; hs/ps: [H] / [A + 32] scratch.
__device__ inline void mlp_train_warp(float* blk, const ThrlAgentSpec& spec, int cap, int head, int N, float* sp, float* gs,
                                      float* hs, float* ps, int lane) {
  const int H = spec.hidden, A = spec.actions;
  const int P = mlp_P(spec), EW = mlp_entry_words(spec);
  float* buf = blk + 3 * (size_t)P + THRL_MLP_HEADER_WORDS;
  float *gw1 = gs, *gb1 = gs + H, *gWT = gs + 2 * H, *gbp = gs + 2 * H + H * A;
  const float* WT = sp + 2 * H;
  for (int i = lane; i < P; i += 32) gs[i] = 0.0f;
  // discounted returns, newest to oldest (:177-180); they overwrite the rewards in the buffer (it is emptied afterwards)
  const float gam = (float)spec.gamma;
  float mean = 0.0f, sd = 1.0f;
  if (lane == 0) {
    float nxt = 0.0f;
    for (int j = N - 1; j >= 0; --j) {
      int sl = head + j;
      if (sl >= cap) sl -= cap;
      const float r = buf[(size_t)sl * EW + 2];
      const float d = j == N - 1 ? r : __fadd_rn(r, __fmul_rn(gam, nxt));
      buf[(size_t)sl * EW + 2] = d;
      nxt = d;
    }
    double sum = 0.0;
    for (int j = 0; j < N; ++j) { int sl = head + j; if (sl >= cap) sl -= cap; sum = __dadd_rn(sum, (double)buf[(size_t)sl * EW + 2]); }
    mean = (float)__ddiv_rn(sum, (double)N);
    double ss = 0.0;
    for (int j = 0; j < N; ++j) {
      int sl = head + j; if (sl >= cap) sl -= cap;
      const double d = __dsub_rn((double)buf[(size_t)sl * EW + 2], (double)mean);
      ss = __dadd_rn(ss, __dmul_rn(d, d));
    }
    sd = (float)sqrt(__ddiv_rn(ss, (double)(N - 1)));
  }
  mean = __shfl_sync(kFull, mean, 0);
  sd = __shfl_sync(kFull, sd, 0);
  const float invN = __fdiv_rn(1.0f, (float)N);
  const float ce = (float)spec.entropy, cen = __fmul_rn(ce, invN);
  __syncwarp();
  for (int j = 0; j < N; ++j) {
    int sl = head + j;
    if (sl >= cap) sl -= cap;
    const float s = buf[(size_t)sl * EW];
    const int a = __float_as_int(buf[(size_t)sl * EW + 1]);
    const float G = __fdiv_rn(__fsub_rn(buf[(size_t)sl * EW + 2], mean), sd);
    const float c = __fmul_rn(G, invN);
    mlp_forward_warp(sp, H, A, s, hs, ps, lane);
    float Hn = 0.0f;
    if (ce != 0.0f) { Hn = mlp_entropy_warp(ps, A); __syncwarp(); }
    for (int k = lane; k < A; k += 32) {  // d loss / d logits (:185)
      const float pk = ps[k];
      float dl = __fmul_rn(__fsub_rn(pk, k == a ? 1.0f : 0.0f), c);
      if (ce != 0.0f) dl = __fadd_rn(dl, mlp_entropy_term(pk, Hn, cen));
      ps[k] = dl;
      gbp[k] = __fadd_rn(gbp[k], dl);
    }
    __syncwarp();
    for (int jh = lane; jh < H; jh += 32) {  // lane = hidden unit
      float dh = 0.0f;
      const float hj = hs[jh];
      for (int k = 0; k < A; ++k) {
        const float dl = ps[k];
        dh = __fadd_rn(dh, __fmul_rn(dl, WT[jh * A + k]));
        gWT[jh * A + k] = __fadd_rn(gWT[jh * A + k], __fmul_rn(dl, hj));
      }
      if (hj > 0.0f) {
        gw1[jh] = __fadd_rn(gw1[jh], __fmul_rn(dh, s));
        gb1[jh] = __fadd_rn(gb1[jh], dh);
      }
    }
    __syncwarp();
  }
  mlp_clip_adam_warp(blk, spec, sp, gs, lane);
}

// ActorCritic.train_net (agents.py:280-305); the [N,N] advantage broadcast collapsed to O(N) sums, see oracle ac_train.
__device__ inline void ac_train_warp(float* blk, const ThrlAgentSpec& spec, int cap, int head, int N, float* sp, float* gs,
                                     float* hs, float* ps, int lane) {
  const int H = spec.hidden, A = spec.actions;
  const int P = mlp_P(spec), EW = mlp_entry_words(spec);
  const float* buf = blk + 3 * (size_t)P + THRL_MLP_HEADER_WORDS;
  float *gw1 = gs, *gb1 = gs + H, *gWT = gs + 2 * H, *gbp = gs + 2 * H + H * A, *gwv = gbp + A, *gbv = gwv + H;
  const float *WT = sp + 2 * H, *wv = sp + 2 * H + H * A + A;
  const float bv = sp[2 * H + H * A + A + H];
  for (int i = lane; i < P; i += 32) gs[i] = 0.0f;
  const float gam = (float)spec.gamma;
  double R = 0.0, D = 0.0;  // every lane accumulates the same (warp-uniform) values in the same order
  for (int i = 0; i < N; ++i) {  // d_i = gamma * v(s'_i) - v(s_i) (:289)
    int sl = head + i;
    if (sl >= cap) sl -= cap;
    mlp_hidden_warp(sp, H, buf[(size_t)sl * EW], hs, lane);
    const float v = ac_value_warp(wv, bv, hs, H, lane);
    mlp_hidden_warp(sp, H, buf[(size_t)sl * EW + 3], hs, lane);
    const float vp = ac_value_warp(wv, bv, hs, H, lane);
    const float d = __fsub_rn(__fmul_rn(gam, vp), v);
    R = __dadd_rn(R, (double)buf[(size_t)sl * EW + 2]);
    D = __dadd_rn(D, (double)d);
  }
  const float fN = (float)N, fR = (float)R, fD = (float)D;
  const float invN2 = __fdiv_rn(1.0f, __fmul_rn(fN, fN));
  const float ce = (float)spec.entropy, cen = __fmul_rn(ce, __fdiv_rn(1.0f, fN));
  for (int j = 0; j < N; ++j) {
    int sl = head + j;
    if (sl >= cap) sl -= cap;
    const float s = buf[(size_t)sl * EW], r = buf[(size_t)sl * EW + 2], s2 = buf[(size_t)sl * EW + 3];
    const int a = __float_as_int(buf[(size_t)sl * EW + 1]);
    mlp_hidden_warp(sp, H, s2, hs, lane);
    const float vp = ac_value_warp(wv, bv, hs, H, lane);
    mlp_forward_warp(sp, H, A, s, hs, ps, lane);  // leaves h(s) in hs, pi(s) in ps
    const float v = ac_value_warp(wv, bv, hs, H, lane);
    const float d = __fsub_rn(__fmul_rn(gam, vp), v);
    const float ca = __fmul_rn(__fadd_rn(__fmul_rn(fN, r), fD), invN2);                        // actor weight (N r_j + D) / N^2
    const float cv = __fmul_rn(-2.0f, __fmul_rn(__fadd_rn(fR, __fmul_rn(fN, d)), invN2));      // dL/dv_j
    const float cvp = __fmul_rn(-gam, cv);                                                     // dL/dv'_j
    float Hn = 0.0f;
    if (ce != 0.0f) { Hn = mlp_entropy_warp(ps, A); __syncwarp(); }
    for (int k = lane; k < A; k += 32) {
      const float pk = ps[k];
      float dl = __fmul_rn(__fsub_rn(pk, k == a ? 1.0f : 0.0f), ca);
      if (ce != 0.0f) dl = __fadd_rn(dl, mlp_entropy_term(pk, Hn, cen));
      ps[k] = dl;
      gbp[k] = __fadd_rn(gbp[k], dl);
    }
    __syncwarp();
    for (int jh = lane; jh < H; jh += 32) {
      float dh = 0.0f;
      const float hj = hs[jh];
      for (int k = 0; k < A; ++k) {
        const float dl = ps[k];
        dh = __fadd_rn(dh, __fmul_rn(dl, WT[jh * A + k]));
        gWT[jh * A + k] = __fadd_rn(gWT[jh * A + k], __fmul_rn(dl, hj));
      }
      gwv[jh] = __fadd_rn(gwv[jh], __fmul_rn(cv, hj));  // value head at s_j shares h with the policy head
      dh = __fadd_rn(dh, __fmul_rn(cv, wv[jh]));
      if (hj > 0.0f) {
        gw1[jh] = __fadd_rn(gw1[jh], __fmul_rn(dh, s));
        gb1[jh] = __fadd_rn(gb1[jh], dh);
      }
    }
    if (lane == 0) gbv[0] = __fadd_rn(gbv[0], cv);
    mlp_hidden_warp(sp, H, s2, hs, lane);  // value head at s'_j
    for (int jh = lane; jh < H; jh += 32) {
      const float hj = hs[jh];
      gwv[jh] = __fadd_rn(gwv[jh], __fmul_rn(cvp, hj));
      const float dh = __fmul_rn(cvp, wv[jh]);
      if (hj > 0.0f) {
        gw1[jh] = __fadd_rn(gw1[jh], __fmul_rn(dh, s2));
        gb1[jh] = __fadd_rn(gb1[jh], dh);
      }
    }
    if (lane == 0) gbv[0] = __fadd_rn(gbv[0], cvp);
    __syncwarp();
  }
  mlp_clip_adam_warp(blk, spec, sp, gs, lane);
}

// ---- CAC (agents.py:333-417): deterministic transcendentals with the oracle's operation sequences
__device__ __forceinline__ float det_tanhf(float x) {
  const float ax = fabsf(x);
  float t;
  if (ax < 0.1f) {
    const float x2 = __fmul_rn(ax, ax);
    float p = -17.0f / 315.0f;
    p = __fadd_rn(__fmul_rn(p, x2), 2.0f / 15.0f);
    p = __fsub_rn(__fmul_rn(p, x2), 1.0f / 3.0f);
    p = __fmul_rn(p, x2);
    p = __fmul_rn(p, ax);
    t = __fadd_rn(ax, p);
  } else {
    const float e = det_expf(__fmul_rn(-2.0f, ax));
    t = __fdiv_rn(__fsub_rn(1.0f, e), __fadd_rn(1.0f, e));
  }
  return x < 0.0f ? -t : t;
}
__device__ __forceinline__ float det_sigmoidf(float x) {
  if (x >= 0.0f) { const float e = det_expf(-x); return __fdiv_rn(1.0f, __fadd_rn(1.0f, e)); }
  const float e = det_expf(x);
  return __fdiv_rn(e, __fadd_rn(1.0f, e));
}
__device__ __forceinline__ float det_softplusf(float x) {
  if (x > 20.0f) return x;
  const float e = det_expf(-fabsf(x));
  const float l = (float)det_log(__dadd_rn(1.0, (double)e));
  return __fadd_rn(x > 0.0f ? x : 0.0f, l);
}
// staged CAC parameters: w1 b1 | wmu bmu | wsd bsd | wv bv
__device__ __forceinline__ float cac_head(const float* w, const float* hs, int H, int lane) { return ac_value_warp(w, w[H], hs, H, lane); }

// CAC.sample_action (agents.py:374-378) for a standard normal deviate z; warp-cooperative, every lane returns the action
__device__ inline float cac_action_warp(const float* sp, int H, float s, float* hs, double z, int lane) {
  mlp_hidden_warp(sp, H, s, hs, lane);
  const float zmu = cac_head(sp + 2 * H, hs, H, lane), zsd = cac_head(sp + 3 * H + 1, hs, H, lane);
  const float mu = __fmul_rn(4.0f, det_tanhf(zmu)), sd = det_softplusf(zsd);
  return det_sigmoidf(__fadd_rn(mu, __fmul_rn(sd, (float)z)));
}

// CAC.train_net (agents.py:391-417); closed form of the [N,N] loss via five moments, see oracle cac_train.
__device__ inline void cac_train_warp(float* blk, const ThrlAgentSpec& spec, int cap, int head, int N, float* sp, float* gs,
                                      float* hs, int lane) {
  const int H = spec.hidden;
  const int P = mlp_P(spec), EW = mlp_entry_words(spec);
  const float* buf = blk + 3 * (size_t)P + THRL_MLP_HEADER_WORDS;
  const float *wmu = sp + 2 * H, *wsd = wmu + H + 1, *wv = wsd + H + 1;
  float *gw1 = gs, *gb1 = gs + H, *gwmu = gs + 2 * H, *gwsd = gwmu + H + 1, *gwv = gwsd + H + 1;
  for (int i = lane; i < P; i += 32) gs[i] = 0.0f;
  const float gam = (float)spec.gamma;
  double Sr = 0.0, Sl = 0.0, Sl2 = 0.0, Srl = 0.0, Srl2 = 0.0;  // every lane the same values, same order
  for (int j = 0; j < N; ++j) {
    int sl = head + j;
    if (sl >= cap) sl -= cap;
    const float a_ = __fadd_rn(5e-5f, __fmul_rn(__fsub_rn(1.0f, 1e-4f), buf[(size_t)sl * EW + 1]));
    const float ratio = __fdiv_rn(a_, __fsub_rn(1.0f, a_));
    const double l = (double)(float)det_log((double)ratio), r = (double)buf[(size_t)sl * EW + 2];
    Sr = __dadd_rn(Sr, r);
    Sl = __dadd_rn(Sl, l);
    Sl2 = __dadd_rn(Sl2, __dmul_rn(l, l));
    Srl = __dadd_rn(Srl, __dmul_rn(r, l));
    Srl2 = __dadd_rn(Srl2, __dmul_rn(__dmul_rn(r, l), l));
  }
  const double dN = (double)N, invN2 = __ddiv_rn(1.0, __dmul_rn(dN, dN));
  for (int i = 0; i < N; ++i) {
    int sl = head + i;
    if (sl >= cap) sl -= cap;
    const float s = buf[(size_t)sl * EW], s2 = buf[(size_t)sl * EW + 3];
    mlp_hidden_warp(sp, H, s2, hs, lane);
    const float vp = cac_head(wv, hs, H, lane);
    mlp_hidden_warp(sp, H, s, hs, lane);
    const float zmu = cac_head(wmu, hs, H, lane), zsd = cac_head(wsd, hs, H, lane), v = cac_head(wv, hs, H, lane);
    const float t = det_tanhf(zmu), mu = __fmul_rn(4.0f, t), sd = det_softplusf(zsd);
    const float d = __fsub_rn(__fmul_rn(gam, vp), v);
    const double dd = (double)d, dmu = (double)mu, dsd = (double)sd;
    const double A0 = __dadd_rn(Sr, __dmul_rn(dN, dd));
    const double A1 = __dadd_rn(__dsub_rn(Srl, __dmul_rn(dmu, Sr)), __dmul_rn(dd, __dsub_rn(Sl, __dmul_rn(dN, dmu))));
    const double A2 = __dadd_rn(__dadd_rn(__dsub_rn(Srl2, __dmul_rn(__dmul_rn(2.0, dmu), Srl)), __dmul_rn(__dmul_rn(dmu, dmu), Sr)),
                                __dmul_rn(dd, __dadd_rn(__dsub_rn(Sl2, __dmul_rn(__dmul_rn(2.0, dmu), Sl)), __dmul_rn(__dmul_rn(dN, dmu), dmu))));
    const float gmu = (float)__dmul_rn(-__ddiv_rn(A1, __dmul_rn(dsd, dsd)), invN2);
    float gsd = (float)__dmul_rn(-__dsub_rn(__ddiv_rn(A2, __dmul_rn(__dmul_rn(dsd, dsd), dsd)), __ddiv_rn(A0, dsd)), invN2);
    if (spec.entropy != 0.0)  // + c_e * (-mean Normal(mu, sd).entropy()) (agents.py:410-412): d/dsd = -c_e / (N sd)
      gsd = __fadd_rn(gsd, (float)(-__ddiv_rn(spec.entropy, __dmul_rn(dN, dsd))));
    const float cv = (float)__dmul_rn(__dmul_rn(-2.0, A0), invN2);
    const float cvp = __fmul_rn(-gam, cv);
    const float dzmu = __fmul_rn(gmu, __fmul_rn(4.0f, __fsub_rn(1.0f, __fmul_rn(t, t))));
    const float dzsd = __fmul_rn(gsd, det_sigmoidf(zsd));
    for (int jh = lane; jh < H; jh += 32) {
      const float hj = hs[jh];
      gwmu[jh] = __fadd_rn(gwmu[jh], __fmul_rn(dzmu, hj));
      gwsd[jh] = __fadd_rn(gwsd[jh], __fmul_rn(dzsd, hj));
      gwv[jh] = __fadd_rn(gwv[jh], __fmul_rn(cv, hj));
      float dh = __fmul_rn(dzmu, wmu[jh]);
      dh = __fadd_rn(dh, __fmul_rn(dzsd, wsd[jh]));
      dh = __fadd_rn(dh, __fmul_rn(cv, wv[jh]));
      if (hj > 0.0f) {
        gw1[jh] = __fadd_rn(gw1[jh], __fmul_rn(dh, s));
        gb1[jh] = __fadd_rn(gb1[jh], dh);
      }
    }
    if (lane == 0) {
      gwmu[H] = __fadd_rn(gwmu[H], dzmu);
      gwsd[H] = __fadd_rn(gwsd[H], dzsd);
      gwv[H] = __fadd_rn(gwv[H], cv);
    }
    mlp_hidden_warp(sp, H, s2, hs, lane);  // value head at s'_i
    for (int jh = lane; jh < H; jh += 32) {
      const float hj = hs[jh];
      gwv[jh] = __fadd_rn(gwv[jh], __fmul_rn(cvp, hj));
      const float dh = __fmul_rn(cvp, wv[jh]);
      if (hj > 0.0f) {
        gw1[jh] = __fadd_rn(gw1[jh], __fmul_rn(dh, s2));
        gb1[jh] = __fadd_rn(gb1[jh], dh);
      }
    }
    if (lane == 0) gwv[H] = __fadd_rn(gwv[H], cvp);
    __syncwarp();
  }
  mlp_clip_adam_warp(blk, spec, sp, gs, lane);
}

template <typename QT>
__global__ void __launch_bounds__(512, 1) qtable_scan_mixed(const __grid_constant__ MixedParams p) {
  extern __shared__ __align__(16) unsigned char smem[];
  const ThrlGame& G = p.game;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, warps_per_cta = blockDim.x >> 5;
  const int n = G.n_agents, T = G.max_steps, E = p.E, Hp = p.Hp;
  const bool is_agent = lane < n;

  // ---- CTA-shared per-action tables (QTable.scale: k/(A-1), agents.py:51-57; Reinforce.scale: k/A, agents.py:154-158)
  double* lutAQ = reinterpret_cast<double*>(smem);
  double* lutXT = lutAQ + p.lut_total;
  {
    const double ab = __ddiv_rn(G.a, G.b);
    int base = 0;
    for (int i = 0; i < n; ++i) {
      const ThrlAgentSpec& s = G.agent[i];
      for (int k = threadIdx.x; k < s.actions; k += blockDim.x) {
        const double x = s.kind == THRL_AGENT_QTABLE
                             ? scale_action(k, s.actions, s.action_lo, s.action_hi)
                             : __dadd_rn(__dmul_rn(__ddiv_rn((double)k, (double)s.actions), __dsub_rn(s.action_hi, s.action_lo)), s.action_lo);
        lutAQ[base + k] = __dmul_rn(ab, x);
        lutXT[base + k] = __ddiv_rn(x, (double)T);
      }
      base += s.actions;
    }
  }
  __syncthreads();

  unsigned char* slot = smem + p.cta_bytes + (size_t)warp * p.warp_bytes;
  double* P = reinterpret_cast<double*>(slot + p.off_P);          // [Hp] price ring (QTable batches)
  uint8_t* act = slot + p.off_act;                                // [n][Hp]
  int32_t* pre = reinterpret_cast<int32_t*>(slot + p.off_pre);    // [T][n] forced action (CAC: float32 bits), -1 greedy (QTable), -2 sample (MLP)
  double* newa = reinterpret_cast<double*>(slot + p.off_newa);    // [T]
  uint16_t* rowbuf = reinterpret_cast<uint16_t*>(slot + p.off_row);
  QT* oldv = reinterpret_cast<QT*>(slot + p.off_old);
  double* hpw = reinterpret_cast<double*>(slot + p.off_hp);       // [n][5]
  float* par = reinterpret_cast<float*>(slot + p.off_par);        // staged MLP parameters of all MLP agents
  float* grad = reinterpret_cast<float*>(slot + p.off_grad);      // gradient scratch (largest MLP agent)
  float* hs = reinterpret_cast<float*>(slot + p.off_h);           // [Hmax] hidden activations, then [Amax+32] probabilities
  int Hmax = 0;
  for (int i = 0; i < n; ++i) if (G.agent[i].kind != THRL_AGENT_QTABLE && G.agent[i].hidden > Hmax) Hmax = G.agent[i].hidden;
  float* ps = hs + Hmax;

  int my_cap = 0, my_minmem = 0x7fffffff, my_lut = 0, my_kind = 0, my_len = 0;
  float my_msf = 1.f, my_sf = 1.f;
  if (is_agent) {
    const ThrlAgentSpec& s = G.agent[lane];
    my_cap = s.capacity; my_minmem = s.min_memory; my_kind = s.kind;
    my_msf = (float)s.max_state; my_sf = (float)s.states;
    for (int j = 0; j < lane; ++j) my_lut += G.agent[j].actions;
  }
  const bool never_fires = my_minmem > my_cap;
  const int lead = lead_exact_floats(G);

  const long long total_warps = (long long)gridDim.x * warps_per_cta;
  for (long long r = (long long)blockIdx.x * warps_per_cta + warp; r < p.n_runs; r += total_warps) {
    QT* tab = reinterpret_cast<QT*>(p.q) + r * G.run_stride;
    uint32_t* cnt = p.counter ? p.counter + r * G.run_stride : nullptr;
    float* slab = p.mlp + r * G.mlp_stride;
    const uint32_t gid = (uint32_t)(p.run_id0 + r);

    if (is_agent) {
      const ThrlAgentSpec& s = G.agent[lane];
      double* h = hpw + lane * 5;
      if (p.hp) {
        const double* src = p.hp + (r * n + lane) * 4;
        h[0] = src[0]; h[1] = src[1]; h[2] = src[2]; h[3] = src[3];
      } else {
        h[0] = s.alpha; h[1] = s.gamma; h[2] = s.eps_end; h[3] = s.eps_step;
      }
      h[4] = p.eps[r * n + lane];
    }
    // stage the MLP parameters (fc_pi.weight transposed to [hidden][actions])
    for (int i = 0; i < n; ++i) {
      const ThrlAgentSpec& s = G.agent[i];
      if (s.kind == THRL_AGENT_QTABLE) continue;
      const int Pn = mlp_P(s);
      const float* src = slab + s.mlp_offset;
      float* dst = par + p.par_off[i];
      for (int e2 = lane; e2 < Pn; e2 += 32) dst[mlp_flat2st(s, e2)] = src[e2];
    }
    double price = p.price[r];
    int pos = 0;
    my_len = 0;
    if (p.ring && !G.regular) {  // QTable transitions pending from the previous call (include/thrl.h ThrlScanArgs.ring)
      const unsigned char* blob = p.ring + r * p.ring_bytes;
      const RingHeader* hd = reinterpret_cast<const RingHeader*>(blob);
      const double* bp = reinterpret_cast<const double*>(blob + sizeof(RingHeader));
      const uint8_t* ba = blob + sizeof(RingHeader) + (size_t)Hp * 8;
      pos = hd->pos;
      if (is_agent) my_len = hd->len[lane];
      for (int j = lane; j < Hp; j += 32) P[j] = bp[j];
      for (int j = lane; j < n * Hp; j += 32) act[j] = ba[j];
    }
    __syncwarp();
    if (lane == 0) P[pos] = price;
    __syncwarp();

    for (int e = 0; e < E; ++e) {
      const uint32_t eabs = (uint32_t)(p.epoch_begin + e);
      const long long step0 = (r * E + e) * (long long)T;

      // ---- per-episode draws (QTable: agents.py:81-82; MLP: forced sample in replay modes)
      for (int idx = lane; idx < T * n; idx += 32) {
        const int t = idx / n, i = idx - t * n;
        int v;
        if (G.agent[i].kind != THRL_AGENT_QTABLE) {
          v = p.rng_mode == THRL_RNG_PHILOX ? -2 : p.replay_ra[step0 * n + idx];
          if (v < 0) v = -2;  // a replay stream may leave this agent's sample to the device
        } else if (p.rng_mode == THRL_RNG_REPLAY_ACTIONS) {
          v = p.replay_ra[step0 * n + idx];
        } else if (p.rng_mode == THRL_RNG_REPLAY_DRAWS) {
          const double u = p.replay_u[step0 * n + idx];
          v = u < hpw[i * 5 + 4] ? p.replay_ra[step0 * n + idx] : -1;
        } else {
          uint32_t x[4];
          philox4x32_10(gid, eabs, (uint32_t)t, (uint32_t)(i >> 1) | (kStreamAct << 16), p.k0, p.k1, x);
          const double u = u32_unit(x[2 * (i & 1)]);
          const int ra = (int)__umulhi(x[2 * (i & 1) + 1], (uint32_t)G.agent[i].actions);
          v = u < hpw[i * 5 + 4] ? ra : -1;
        }
        pre[idx] = v;
      }
      if (p.noisy) {
        for (int t = lane; t < T; t += 32) {
          double na = G.a;
          if (p.rng_mode == THRL_RNG_PHILOX) {
            uint32_t x[4];
            philox4x32_10(gid, eabs, (uint32_t)t, kStreamEnv << 16, p.k0, p.k1, x);
            if (u53(x[0], x[1]) < G.noise_prob) {
              const double lo = __dmul_rn(G.a, 0.7);
              na = __dadd_rn(lo, __dmul_rn(__dsub_rn(G.a, lo), u53(x[2], x[3])));
            }
          } else if (p.replay_new_a) {
            na = p.replay_new_a[step0 + t];
          }
          newa[t] = na;
        }
      }
      __syncwarp();

      // ---- the episode (trainer.py:50-67)
      double rlog = 0.0, alog = 0.0;
      for (int t = 0; t < T; ++t) {
        int k = is_agent ? pre[t * n + lane] : 0;
        int arow = 0;
        if (is_agent && k == -1) arow = act_row(price, my_msf, my_sf);
        // greedy QTable actions: first argmax of the live (frozen within the episode) table row (agents.py:84-88)
        unsigned need = __ballot_sync(kFull, is_agent && k == -1);
        while (need) {
          const int i = __ffs(need) - 1;
          need &= need - 1;
          const int ri = __shfl_sync(kFull, arow, i);
          const ThrlAgentSpec& s = G.agent[i];
          const int g = row_argmax(tab + s.table_offset + (size_t)ri * s.row_stride, s.actions, lane);
          if (lane == i) k = g;
        }
        // free-running MLP agents: forward pass + inverse-CDF sample (agents.py:160-163)
        unsigned samp = __ballot_sync(kFull, is_agent && k == -2);
        while (samp) {
          const int i = __ffs(samp) - 1;
          samp &= samp - 1;
          const ThrlAgentSpec& s = G.agent[i];
          uint32_t x[4];
          philox4x32_10(gid, eabs, (uint32_t)t, (uint32_t)(i >> 1) | (kStreamAct << 16), p.k0, p.k1, x);
          if (s.kind == THRL_AGENT_CAC) {  // sigmoid(Normal(mu, std).sample()) (agents.py:374-378)
            const unsigned long long m = ((unsigned long long)x[2 * (i & 1)] << 21) | (unsigned long long)(x[2 * (i & 1) + 1] >> 11);
            const double z = det_norminv(((double)m + 0.5) * (1.0 / 9007199254740992.0));
            const float af = cac_action_warp(par + p.par_off[i], s.hidden, (float)price, hs, z, lane);
            if (lane == i) k = __float_as_int(af);
            __syncwarp();
            continue;
          }
          mlp_forward_warp(par + p.par_off[i], s.hidden, s.actions, (float)price, hs, ps, lane);
          const float u = __fmul_rn((float)(x[2 * (i & 1)] >> 8), 1.0f / 16777216.0f);
          float c = 0.0f;
          int ks = s.actions - 1;
          for (int kk = 0; kk < s.actions; ++kk) {
            c = __fadd_rn(c, ps[kk]);
            if (c > u) { ks = kk; break; }
          }
          if (lane == i) k = ks;
          __syncwarp();
        }
        double aq = 0.0, xt = 0.0;
        if (is_agent) {
          if (my_kind == THRL_AGENT_CAC) {  // CAC.scale (agents.py:368-372): action * (hi - lo) + lo, action a float32 in (0,1)
            const ThrlAgentSpec& s = G.agent[lane];
            const double x = __dadd_rn(__dmul_rn((double)__int_as_float(k), __dsub_rn(s.action_hi, s.action_lo)), s.action_lo);
            aq = __dmul_rn(__ddiv_rn(G.a, G.b), x);
            xt = __ddiv_rn(x, (double)T);
          } else {
            aq = lutAQ[my_lut + k];
            xt = lutXT[my_lut + k];
          }
        }
        const double Q = py_sum_quantities(n, lead, [&](int i) { return shfl_d(aq, i); });  // environments.py:27 sum(A)
        const double na = p.noisy ? newa[t] : G.a;
        const double pn = __dsub_rn(na, __dmul_rn(G.b, Q));
        const double next_price = pn > 0.0 ? pn : (pn != pn ? pn : 0.0);
        const double rew = __dmul_rn(next_price, aq);
        int nxt = pos + 1;
        if (nxt == Hp) nxt = 0;
        if (is_agent) {
          rlog = __dadd_rn(rlog, __ddiv_rn(rew, (double)T));
          alog = __dadd_rn(alog, xt);
          if (my_kind == THRL_AGENT_QTABLE) {
            act[lane * Hp + pos] = (uint8_t)k;
            my_len = my_len < my_cap ? my_len + 1 : my_cap;
          } else {
            // memory.append; replay(cast) makes state and reward float32 (buffers.py:28-38, agents.py:142)
            const ThrlAgentSpec& s = G.agent[lane];
            const int cap = G.mlp_buffer_len[lane];
            if (cap > 0) {
              const int Pn = mlp_P(s), EW = mlp_entry_words(s);
              float* blk = slab + s.mlp_offset;
              int32_t* hdr = reinterpret_cast<int32_t*>(blk + 3 * (size_t)Pn);
              float* mb = blk + 3 * (size_t)Pn + THRL_MLP_HEADER_WORDS;
              int len = hdr[1], head = hdr[2], sl;
              if (len < cap) { sl = head + len; if (sl >= cap) sl -= cap; len++; }
              else { sl = head; head = head + 1 == cap ? 0 : head + 1; }
              mb[(size_t)sl * EW] = (float)price;
              mb[(size_t)sl * EW + 1] = __int_as_float(k);
              mb[(size_t)sl * EW + 2] = (float)rew;
              if (EW == 4) mb[(size_t)sl * EW + 3] = (float)next_price;
              hdr[1] = len; hdr[2] = head;
            }
          }
          if (p.trace_actions) p.trace_actions[(step0 + t) * n + lane] = k;
          if (p.trace_rewards) p.trace_rewards[(step0 + t) * n + lane] = rew;
        }
        if (lane == 0) {
          P[nxt] = next_price;
          if (p.trace_prices) p.trace_prices[step0 + t] = next_price;
        }
        pos = nxt;
        price = next_price;
      }
      __syncwarp();

      // ---- train_net for every agent in order (trainer.py:70)
      for (int i = 0; i < n; ++i) {
        const ThrlAgentSpec& s = G.agent[i];
        if (s.kind != THRL_AGENT_QTABLE) {  // Reinforce.train_net (agents.py:170-194)
          const int cap = G.mlp_buffer_len[i];
          if (cap == 0) continue;
          const int Pn = mlp_P(s);
          float* blk = slab + s.mlp_offset;
          int32_t* hdr = reinterpret_cast<int32_t*>(blk + 3 * (size_t)Pn);
          const int len = hdr[1], head = hdr[2];
          if (len >= s.min_memory) {
            if (s.kind == THRL_AGENT_CAC) cac_train_warp(blk, s, cap, head, len, par + p.par_off[i], grad, hs, lane);
            else if (s.kind == THRL_AGENT_ACTORCRITIC) ac_train_warp(blk, s, cap, head, len, par + p.par_off[i], grad, hs, ps, lane);
            else mlp_train_warp(blk, s, cap, head, len, par + p.par_off[i], grad, hs, ps, lane);
            if (lane == 0) { hdr[1] = 0; hdr[2] = 0; }  // :194 memory.empty()
            __syncwarp();
          }
          continue;
        }
        const int L = __shfl_sync(kFull, my_len, i);
        const int fires = __shfl_sync(kFull, (int)(!never_fires && my_len >= my_minmem), i);
        if (!fires) continue;
        const int A = s.actions;
        QT* tb = tab + s.table_offset;
        const double alpha = hpw[i * 5 + 0], gamma = hpw[i * 5 + 1];
        const double one_m_alpha = __dsub_rn(1.0, alpha);
        int first = pos - L;
        if (first < 0) first += Hp;
        const int loff = __shfl_sync(kFull, my_lut, i);
        for (int j = lane; j <= L; j += 32) {  // encodes (agents.py:62,66)
          int sl = first + j;
          if (sl >= Hp) sl -= Hp;
          rowbuf[j] = (uint16_t)upd_row(P[sl], s.max_state, (double)s.states);
        }
        __syncwarp();
        for (int j = lane; j < L; j += 32) {  // stale snapshot (:67)
          int sl = first + j;
          if (sl >= Hp) sl -= Hp;
          oldv[j] = tb[(size_t)rowbuf[j] * s.row_stride + act[i * Hp + sl]];
        }
        __syncwarp();
        int sl = first;
        for (int j = 0; j < L; ++j) {  // the sequential pass (:68-76)
          const int st = rowbuf[j], ns = rowbuf[j + 1];
          const int k = act[i * Hp + sl];
          int sn = sl + 1;
          if (sn == Hp) sn = 0;
          const double reward = __dmul_rn(P[sn], lutAQ[loff + k]);
          const double next_max = (double)row_max(tb + (size_t)ns * s.row_stride, A, lane);
          const double nv = __dadd_rn(__dmul_rn(one_m_alpha, (double)oldv[j]),
                                      __dmul_rn(alpha, __dadd_rn(reward, __dmul_rn(gamma, next_max))));
          if ((k & 31) == lane) {  // the lane that owns column k
            tb[(size_t)st * s.row_stride + k] = (QT)nv;
            if (cnt) atomicAdd(cnt + s.table_offset + (size_t)st * s.row_stride + k, 1u);
          }
          sl = sn;
        }
        if (lane == i) my_len = 0;
        __syncwarp();
      }
      // epsilon decay (QTable, :78); logs
      if (is_agent) {
        double* h = hpw + lane * 5;
        if (my_kind == THRL_AGENT_QTABLE) h[4] = __dadd_rn(h[2], __dmul_rn(__dsub_rn(h[4], h[2]), h[3]));
        if (r < p.n_log_runs) {
          if (p.rewards_log) p.rewards_log[(r * E + e) * n + lane] = rlog;
          if (p.actions_log) p.actions_log[(r * E + e) * n + lane] = alog;
        }
        if (p.stats) {
          unsigned long long* s4 = reinterpret_cast<unsigned long long*>(p.stats) + ((size_t)e * n + lane) * THRL_STATS_K;
          atomicAdd(s4 + 0, (unsigned long long)fx_round(__dmul_rn(rlog, THRL_STATS_SCALE_SUM)));
          atomicAdd(s4 + 1, (unsigned long long)fx_round(__dmul_rn(__dmul_rn(rlog, rlog), THRL_STATS_SCALE_SQ)));
          atomicAdd(s4 + 2, (unsigned long long)fx_round(__dmul_rn(alog, THRL_STATS_SCALE_SUM)));
          atomicAdd(s4 + 3, (unsigned long long)fx_round(__dmul_rn(__dmul_rn(alog, alog), THRL_STATS_SCALE_SQ)));
        }
      }
      __syncwarp();
    }

    // ---- write the run back (MLP parameters in state_dict order)
    for (int i = 0; i < n; ++i) {
      const ThrlAgentSpec& s = G.agent[i];
      if (s.kind == THRL_AGENT_QTABLE) continue;
      const int Pn = mlp_P(s);
      float* dstg = slab + s.mlp_offset;
      const float* srcs = par + p.par_off[i];
      for (int e2 = lane; e2 < Pn; e2 += 32) dstg[e2] = srcs[mlp_flat2st(s, e2)];
    }
    if (is_agent) p.eps[r * n + lane] = hpw[lane * 5 + 4];
    if (lane == 0) p.price[r] = price;
    if (p.ring && !G.regular) {
      unsigned char* blob = p.ring + r * p.ring_bytes;
      RingHeader* hd = reinterpret_cast<RingHeader*>(blob);
      double* bp = reinterpret_cast<double*>(blob + sizeof(RingHeader));
      uint8_t* ba = blob + sizeof(RingHeader) + (size_t)Hp * 8;
      if (lane == 0) hd->pos = pos;
      if (is_agent) hd->len[lane] = my_len;
      for (int j = lane; j < Hp; j += 32) bp[j] = P[j];
      for (int j = lane; j < n * Hp; j += 32) ba[j] = act[j];
    }
    __syncwarp();
  }
}

// ---------------------------------------------------------------- greedy evaluation with MLP agents (utils.py:27-47)
struct EvalMixedParams {
  ThrlGame game;
  long long n_runs;
  int iters;
  const void* q;
  const float* mlp;
  const double* price0;
  const double* new_a;  // [R][iters][T] demand intercept per step, or NULL: the noise-free curve
  double* rewards;
  double* actions;
  int warp_bytes, off_par, off_h;
  int par_off[THRL_MAX_AGENTS];
};

// play_game with every agent's get_action: QTable first argmax on the f64 encode (agents.py:91-92); Reinforce / ActorCritic
// argmax of pi(float32 state) (agents.py:165-168, 275-278); CAC sigmoid(4 tanh(fc_mu(h))) (agents.py:380-384).  Warp per run.
template <typename QT>
__global__ void __launch_bounds__(256) greedy_eval_mixed(const __grid_constant__ EvalMixedParams p) {
  extern __shared__ __align__(16) unsigned char smem[];
  const ThrlGame& G = p.game;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int n = G.n_agents, T = G.max_steps;
  unsigned char* slot = smem + (size_t)warp * p.warp_bytes;
  float* par = reinterpret_cast<float*>(slot + p.off_par);
  float* hs = reinterpret_cast<float*>(slot + p.off_h);
  int Hmax = 0;
  for (int i = 0; i < n; ++i) if (G.agent[i].kind != THRL_AGENT_QTABLE && G.agent[i].hidden > Hmax) Hmax = G.agent[i].hidden;
  float* ps = hs + Hmax;
  const long long gw = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long total = ((long long)gridDim.x * blockDim.x) >> 5;
  const double ab = __ddiv_rn(G.a, G.b);
  for (long long r = gw; r < p.n_runs; r += total) {
    const QT* tab = reinterpret_cast<const QT*>(p.q) + r * G.run_stride;
    const float* slab = p.mlp + r * G.mlp_stride;
    __syncwarp();
    for (int i = 0; i < n; ++i) {
      const ThrlAgentSpec& s = G.agent[i];
      if (s.kind == THRL_AGENT_QTABLE) continue;
      const int Pn = mlp_P(s);
      for (int e2 = lane; e2 < Pn; e2 += 32) par[p.par_off[i] + mlp_flat2st(s, e2)] = slab[s.mlp_offset + e2];
    }
    __syncwarp();
    for (int it = 0; it < p.iters; ++it) {
      double price = p.price0[r * p.iters + it];
      for (int t = 0; t < T; ++t) {
        double my_x = 0.0, my_aq = 0.0;
        for (int i = 0; i < n; ++i) {
          const ThrlAgentSpec& s = G.agent[i];
          double x;
          if (s.kind == THRL_AGENT_QTABLE) {
            const int row = upd_row(price, s.max_state, (double)s.states);
            const int k = row_argmax(tab + s.table_offset + (size_t)row * s.row_stride, s.actions, lane);
            x = scale_action(k, s.actions, s.action_lo, s.action_hi);
          } else if (s.kind == THRL_AGENT_CAC) {
            const float* sp = par + p.par_off[i];
            mlp_hidden_warp(sp, s.hidden, (float)price, hs, lane);
            const float zmu = cac_head(sp + 2 * s.hidden, hs, s.hidden, lane);
            const float af = det_sigmoidf(__fmul_rn(4.0f, det_tanhf(zmu)));
            x = __dadd_rn(__dmul_rn((double)af, __dsub_rn(s.action_hi, s.action_lo)), s.action_lo);
          } else {
            mlp_forward_warp(par + p.par_off[i], s.hidden, s.actions, (float)price, hs, ps, lane);
            const int k = row_argmax(ps, s.actions, lane);  // torch.argmax: first maximal index
            x = __dadd_rn(__dmul_rn(__ddiv_rn((double)k, (double)s.actions), __dsub_rn(s.action_hi, s.action_lo)), s.action_lo);
            __syncwarp();
          }
          const double aq = __dmul_rn(ab, x);
          if (lane == i) { my_x = x; my_aq = aq; }
        }
        const double Q = py_sum_quantities(n, lead_exact_floats(G), [&](int i) { return shfl_d(my_aq, i); });
        const double na = p.new_a ? p.new_a[(r * p.iters + it) * T + t] : G.a;  // environments.py:28-31
        const double pn = __dsub_rn(na, __dmul_rn(G.b, Q));
        const double next_price = pn > 0.0 ? pn : (pn != pn ? pn : 0.0);
        if (lane < n) {
          const long long o = ((r * p.iters + it) * T + t) * n + lane;
          p.rewards[o] = __dmul_rn(next_price, my_aq);
          p.actions[o] = my_x;
        }
        price = next_price;
      }
    }
  }
}

}  // namespace thrl
