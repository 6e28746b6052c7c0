// thrl_scan_pwc.cuh — games with MLP agents whose network input is a CONTINUOUS price: demand noise (environments.py:28-31,
// the environment's default noise_prob is 0.05), CAC agents (continuous actions, agents.py:333-417), or QTable agents whose
// batches span episodes next to MLP agents.  These are the games the lattice kernel (thrl_scan_pwl.cuh) cannot take.
//
// The per-run "N x 256 x 22 GEMMs" do not survive here either, because fact (2) of thrl_scan_pwl.cuh does not need a
// lattice: every network of the reference is 1 -> H -> heads with a ReLU, i.e. a piecewise-linear function of the scalar
// price with one breakpoint per hidden unit.  Per (run, MLP agent) the kernel keeps
//   * th[j]: the EXACT float32 threshold of unit j -- the smallest (w1_j > 0: first active) or first inactive (w1_j < 0)
//     price under the reference's own predicate fl(fl(s*w1_j)+b1_j) > 0, found by bisection over the float bit patterns
//     (the predicate is monotone in s), so the set of active units of any price is exactly the reference's;
//   * the units ranked by threshold (ord) and an INTERVAL TABLE tab[r][c] = (S1, S0), r = 0..H: on the r-th interval between
//     consecutive thresholds every head output is z_c(s) = S1*s + S0 (f64 running sums over the ranked units, bias folded in).
// Acting: r = #{j: th[j] <= s} (one compare per unit, warp-wide add), one table row (lane = head column), softmax + inverse CDF
// (Reinforce / ActorCritic) or tanh / softplus / sigmoid (CAC): O(H/32 + A) per step instead of O(H*A).
// Updating: the buffered transitions are put in interval order (stable counting sort) and swept once; each one's head gradient
// dL/dz (the oracle's float32 per-sample coefficients) is added to running f64 sums that are stored whenever the interval
// advances.  These prefix sums give, at every unit's rank, the sums over the samples the unit is active on: M0 = sum dz, M1 = sum dz*s, hence d/dW[c][j] = w1_j*M1 + b1_j*M0, d/db1_j = sum_c W[c][j]*M0,
// d/dw1_j = sum_c W[c][j]*M1.  O(N*A + H*A) instead of O(N*H*A).  Then clip_grad_norm_ + Adam, and the tables are rebuilt.
// The arithmetic is the reference's autograd graph summed in another order (f64 accumulation): results agree with the
// order-exact kernel (thrl_scan_mixed.cuh, THRL_KERNEL=mixed) and the oracle to float32 rounding -- tests state the
// tolerance.  QTable agents of the same game are handled exactly as in thrl_scan_mixed.cuh (bit-exact).
#pragma once
#include "thrl_device.cuh"
#include "thrl_scan_mixed.cuh"
#include "thrl_scan_pwl.cuh"  // warp_sum, pwl_active, pwl_clip_adam

namespace thrl {

constexpr int kPwcMaxHidden = 256;        // thresholds are ranked with 8 units per lane
#ifndef THRL_PWC_MAXWARPS
#define THRL_PWC_MAXWARPS 16
#endif
constexpr int kPwcMaxWarps = THRL_PWC_MAXWARPS;  // resident runs per CTA (launch bound: 65,536 / (32 * warps) registers per thread)
constexpr unsigned kPwcKeyMin = 0x007fffffu;   // ukey(-inf)
constexpr unsigned kPwcKeyNone = 0xff800001u;  // above ukey(+inf): the unit never switches

struct PwcParams {
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
  unsigned char* ws;  // per resident warp: interval tables, ranked units, prefix sums, gradient, per-sample scratch, event order
  long long ws_warp_bytes, ws_bkt, ws_grad, ws_xs, ws_xe, ws_evs;
  long long ws_tab[THRL_MAX_AGENTS], ws_ord[THRL_MAX_AGENTS];
  int cta_bytes, warp_bytes;
  int off_P, off_act, off_pre, off_zf, off_newa, off_row, off_old, off_hp, off_hist;
  int off_th[THRL_MAX_AGENTS];  // MLP agent: its thresholds (ranked) in the warp's shared memory
  int ncp[THRL_MAX_AGENTS];     // MLP agent: table row length in (S1, S0) pairs (head columns rounded up to 2)
  int lut_total, Hp, noisy;
};

// order-preserving map float32 -> uint32 (every non-NaN value; -0 < +0)
__device__ __forceinline__ unsigned pwc_ukey(float s) {
  const unsigned b = (unsigned)__float_as_int(s);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float pwc_unkey(unsigned k) { return __int_as_float((int)((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k)); }

// head columns: Reinforce: the actions; ActorCritic: the actions, then v; CAC: mu, std, v
__device__ __forceinline__ int pwc_ncol(const ThrlAgentSpec& s) {
  return s.kind == THRL_AGENT_CAC ? 3 : s.actions + (s.kind == THRL_AGENT_ACTORCRITIC ? 1 : 0);
}
__device__ __forceinline__ int pwc_wrow(const ThrlAgentSpec& s, int c) {  // float offset of column c's weight row [H] (state_dict order)
  const int H = s.hidden, A = s.actions;
  if (s.kind == THRL_AGENT_CAC) return 2 * H + c * (H + 1);
  return c < A ? 2 * H + c * H : 2 * H + A * H + A;
}
__device__ __forceinline__ int pwc_bias(const ThrlAgentSpec& s, int c) {
  const int H = s.hidden, A = s.actions;
  if (s.kind == THRL_AGENT_CAC) return 2 * H + c * (H + 1) + H;
  return c < A ? 2 * H + A * H + c : 2 * H + A * H + A + H;
}

__device__ __forceinline__ double2 pwc_ld2(const double2* a) { return __ldcg(a); }
__device__ __forceinline__ float pwc_eval(double2 t, float s) { return (float)__dadd_rn(__dmul_rn(t.x, (double)s), t.y); }

// exp(x) for the softmaxes of this kernel: Cody-Waite reduction and the Cephes polynomial evaluated with fused multiply-adds
// (under 1 ulp; half the instructions of det_expf, whose unfused sequence exists to be bit-equal to the oracle's -- this kernel
// is compared to float32 rounding, not bit for bit)
__device__ __forceinline__ float pwc_expf(float x) {
  if (x < -87.0f) return 0.0f;
  const float kf = rintf(__fmul_rn(x, 1.44269504f));
  float r = __fmaf_rn(kf, -0.693359375f, x);
  r = __fmaf_rn(kf, 2.12194440e-4f, r);
  float p = 1.9875691500e-4f;
  p = __fmaf_rn(p, r, 1.3981999507e-3f);
  p = __fmaf_rn(p, r, 8.3334519073e-3f);
  p = __fmaf_rn(p, r, 4.1665795894e-2f);
  p = __fmaf_rn(p, r, 1.6666665459e-1f);
  p = __fmaf_rn(p, r, 5.0000001201e-1f);
  p = __fmaf_rn(__fmul_rn(p, r), r, r);
  p = __fadd_rn(p, 1.0f);
  const int k = (int)kf;  // -126 <= k <= 128 here for every finite x >= -87 that does not overflow
  return __fmul_rn(p, __int_as_float((min(max(k, -126), 128) + 127) << 23));
}

// x / T, correctly rounded, for a constant T with y = RN(1 / T): two residual corrections by FMA (Markstein: the second one
// starts from a faithful quotient and is then the correctly rounded quotient).  The reference divides every reward by
// max_steps before adding it to the log (trainer.py:63); div.rn.f64 costs ~35 instructions, this 7.  Outside the range where
// the products are free of overflow / underflow the plain division is used.  (/tmp-style check: 8e8 random operands, T = 1..400,
// no mismatch against the hardware division.)
__device__ __forceinline__ double pwc_div(double x, double T, double y) {
  const double ax = fabs(x);
  if (!(ax < 1e280) || (ax < 1e-280 && x != 0.0)) return __ddiv_rn(x, T);
  double q = __dmul_rn(x, y);
  double r = __fma_rn(-q, T, x);
  q = __fma_rn(r, y, q);
  r = __fma_rn(-q, T, x);
  return __fma_rn(r, y, q);
}

// r = number of thresholds <= key, every lane the same.  th: the ranked thresholds, padded with kPwcKeyNone to 32 * per
// entries; thc[l] = th[l * per + per - 1], the last threshold of block l.  One ballot finds the blocks that lie entirely at or
// below the key, a second one counts inside the next block.
__device__ __forceinline__ int pwc_rank_warp(const unsigned* th, const unsigned* thc, int H, int per, unsigned key, int lane) {
  const int full = __popc(__ballot_sync(kFull, thc[lane] <= key));
  const unsigned m = __ballot_sync(kFull, lane < per && full < 32 && th[full * per + lane] <= key);
  const int r = full * per + __popc(m);
  return r < H ? r : H;
}
// the same for one lane's own key on the RANKED thresholds (upper bound by bisection)
__device__ __forceinline__ int pwc_rank_lane(const unsigned* th, int H, unsigned key) {
  int lo = 0, hi = H;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (th[mid] <= key) lo = mid + 1; else hi = mid;
  }
  return lo;
}

// Thresholds of the current parameters, ranked: th[q] ascending (ties by unit index), ord[q] = unit | leave << 15 where
// leave = 0: the unit is active on prices >= th (intervals r > q), leave = 1: active on prices < th (intervals r <= q).
// Then the interval table tab[(H+1)][ncp].
__device__ inline void pwc_build(const float* blk, const ThrlAgentSpec& spec, unsigned* th, uint16_t* ord, double2* tab, int ncp, int lane) {
  const int H = spec.hidden, NC = pwc_ncol(spec);
  const float *w1 = blk, *b1 = blk + H;
  constexpr int U = kPwcMaxHidden / 32;
  unsigned myth[U];
  unsigned leave_bits = 0;
  __syncwarp();
#pragma unroll
  for (int u = 0; u < U; ++u) {
    const int j = lane + 32 * u;
    myth[u] = kPwcKeyNone;
    if (j < H) {
      const float w = w1[j], b = b1[j];
      unsigned key = kPwcKeyNone, leave = 0;
      if (w > 0.0f || w < 0.0f) {  // first key at which the (monotone) predicate has switched
        const bool target = w > 0.0f;  // w > 0: first active price; w < 0: first inactive price
        unsigned lo = kPwcKeyMin, hi = kPwcKeyNone;
        while (lo < hi) {
          const unsigned mid = lo + ((hi - lo) >> 1);
          if (pwl_active(pwc_unkey(mid), w, b) == target) hi = mid; else lo = mid + 1;
        }
        key = lo;
        leave = target ? 0u : 1u;
      } else {  // w == 0 (or NaN: never active): the same on every price
        leave = (w == 0.0f && b > 0.0f) ? 1u : 0u;
      }
      myth[u] = key;
      leave_bits |= leave << u;
      th[j] = key;
    }
  }
  __syncwarp();
  int pos[U];
#pragma unroll
  for (int u = 0; u < U; ++u) pos[u] = 0;
  for (int jp = 0; jp < H; ++jp) {
    const unsigned t = th[jp];
#pragma unroll
    for (int u = 0; u < U; ++u) pos[u] += (t < myth[u] || (t == myth[u] && jp < lane + 32 * u)) ? 1 : 0;
  }
  __syncwarp();
  const int per = (H + 31) >> 5;
#pragma unroll
  for (int u = 0; u < U; ++u) {
    const int j = lane + 32 * u;
    if (j < H) {
      th[pos[u]] = myth[u];
      ord[pos[u]] = (uint16_t)(j | ((leave_bits >> u & 1u) << 15));
    } else if (j < 32 * per) {
      th[j] = kPwcKeyNone;
    }
  }
  __syncwarp();
  th[32 * per + lane] = th[lane * per + per - 1];  // thc
  __syncwarp();
  // interval table, lane = head column: S(r) = sum over the units active on interval r of W[c][j] * (w1_j, b1_j)
  const bool use = lane < NC;
  const float* wrow = blk + pwc_wrow(spec, use ? lane : 0);
  // Eight ranked units per round: their (w1, b1, W[c]) are loaded together, then applied in rank order.  First round trip: the
  // sum over the units that are active from the lowest price on (they leave at their threshold); second: the running sums.
  double L1 = 0.0, L0 = 0.0;
  for (int pass = 0; pass < 2; ++pass) {
    double S1 = L1, S0 = 0.0;
    if (pass == 1) {
      S0 = __dadd_rn(L0, use ? (double)blk[pwc_bias(spec, lane)] : 0.0);
      if (use) __stcg(tab + lane, make_double2(S1, S0));
    }
    for (int q0 = 0; q0 < H; q0 += 8) {
      unsigned o[8];
      float cwf[8], wf[8], bf[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) o[u] = q0 + u < H ? (unsigned)ord[q0 + u] : 0u;
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int j = (int)(o[u] & 0x7fffu);
        cwf[u] = wrow[j]; wf[u] = w1[j]; bf[u] = b1[j];
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        if (q0 + u < H) {
          const double cw = (double)cwf[u];
          const double t1 = __dmul_rn(cw, (double)wf[u]), t0 = __dmul_rn(cw, (double)bf[u]);
          const bool leave = (o[u] & 0x8000u) != 0;
          if (pass == 0) {
            if (leave) { L1 = __dadd_rn(L1, t1); L0 = __dadd_rn(L0, t0); }
          } else {
            if (leave) { S1 = __dsub_rn(S1, t1); S0 = __dsub_rn(S0, t0); }
            else { S1 = __dadd_rn(S1, t1); S0 = __dadd_rn(S0, t0); }
            if (use) __stcg(tab + (size_t)(q0 + u + 1) * ncp + lane, make_double2(S1, S0));
          }
        }
      }
    }
  }
  __syncwarp();
}

// softmax of the lanes < A of z, then the first k with cumsum(pi)[k] > u (agents.py:160-163), the last action if none
__device__ __forceinline__ int pwc_sample(float z, int A, float u, int lane) {
  const bool col = lane < A;
  const float mx = warp_max(col ? z : NegInf<float>::v());
  const float ex = col ? pwc_expf(__fsub_rn(z, mx)) : 0.0f;
  const float sum = warp_sum(ex);
  float c = col ? __fdiv_rn(ex, sum) : 0.0f;
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    const float t = __shfl_up_sync(kFull, c, off);
    if (lane >= off) c = __fadd_rn(c, t);
  }
  const unsigned m = __ballot_sync(kFull, col && c > u);
  return m ? __ffs(m) - 1 : A - 1;
}

// ---------------------------------------------------------------------------------------------------------------- update
// One update of one MLP agent on its N buffered transitions:
//  1. per-transition coefficients (the oracle's float32 formulas) and interval ranks, lane = transition  -> xs[n]
//  2. the EVENTS -- transition n at its state s_n (all heads), and for ActorCritic / CAC transition n at its next state s'_n
//     (value head only) -- are put in interval order by a stable counting sort (hist in shared memory, evs in the workspace)
//  3. one sweep over the ranked events, lane = head column: pi(.|s) from the interval's table row, dL/dz added to running f64
//     sums (P0 = sum dz, P1 = sum dz * s); whenever the interval advances the sums are stored: pf[r] = sums over the intervals
//     < r, r = 0..H, pf[H+1] = totals.  No read-modify-write, no zero fill; the next event's table row is loaded while the
//     current one is processed, and events of one interval share the row.
//  4. lane = rank q: unit ord[q] is active on the intervals r > q (enter) or r <= q (leave), so its sums over active
//     transitions are totals - pf[q+1] or pf[q+1]; gradient of every parameter; clip_grad_norm_ + Adam.

// stable counting sort of the NE events by rank; afterwards evs[0..NE) lists the events in interval order
// (rank H + 1 marks an event that was merged into another one: those end up behind the returned count)
__device__ inline int pwc_sort_events(const float4* xs, int N, int NE, int H, int* hist, uint32_t* evs, int lane) {
  const int NB = H + 3;
  for (int i = lane; i < NB; i += 32) hist[i] = 0;
  __syncwarp();
  auto rank_of = [&](int e) {
    const int rb = __float_as_int(xs[e < N ? e : e - N].x);
    return e < N ? (rb & 0xffff) : ((rb >> 16) & 0xffff);
  };
  for (int e0 = 0; e0 < NE; e0 += 32) {
    const int e = e0 + lane;
    const unsigned kk = e < NE ? (unsigned)rank_of(e) : 0x10000u + (unsigned)lane;
    const unsigned peers = __match_any_sync(kFull, kk);
    if (e < NE && (peers & lanemask_lt()) == 0) hist[kk + 1] += __popc(peers);
    __syncwarp();
  }
  {  // inclusive scan: afterwards hist[k] = number of events of rank < k = first position of rank k
    const int per = (NB + 31) / 32;
    const int beg = lane * per, end = beg + per < NB ? beg + per : NB;
    int sum = 0;
    for (int k = beg; k < end; ++k) sum += hist[k];
    int incl = sum;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const int t = __shfl_up_sync(kFull, incl, off);
      if (lane >= off) incl += t;
    }
    int run = incl - sum;
    for (int k = beg; k < end; ++k) { run += hist[k]; hist[k] = run; }
  }
  __syncwarp();
  for (int e0 = 0; e0 < NE; e0 += 32) {
    const int e = e0 + lane;
    const unsigned kk = e < NE ? (unsigned)rank_of(e) : 0x10000u + (unsigned)lane;
    const unsigned peers = __match_any_sync(kFull, kk);
    int base = 0;
    if (e < NE) {
      base = hist[kk];
      evs[base + __popc(peers & lanemask_lt())] = (uint32_t)e;
    }
    __syncwarp();
    if (e < NE && (peers & lanemask_lt()) == 0) hist[kk] = base + __popc(peers);
    __syncwarp();
  }
  return hist[H];  // end of rank H = number of events that take part in the sweep
}

__device__ inline void pwc_nan_update(float* blk, const ThrlAgentSpec& spec, float* g, int lane) {
  const int P = mlp_P(spec);  // non-finite coefficients (e.g. zero return variance): the reference's gradient is NaN everywhere
  for (int i = lane; i < P; i += 32) g[i] = __int_as_float(0x7fc00000);
  pwl_clip_adam(blk, spec, g, lane);
}

// Reinforce.train_net (agents.py:170-194), ActorCritic.train_net (:280-305), CAC.train_net (:391-417)
__device__ inline void pwc_train(float* blk, const ThrlAgentSpec& spec, int cap, int head, int N, const unsigned* th, const uint16_t* ord,
                                 const double2* tab, double2* pf, int ncp, float* g, float4* xs, float* xe, uint32_t* evs, int* hist, int lane) {
  const int H = spec.hidden, A = spec.actions, P = mlp_P(spec), EW = mlp_entry_words(spec), NC = pwc_ncol(spec);
  const int kind = spec.kind;
  const bool ac = kind == THRL_AGENT_ACTORCRITIC, cac = kind == THRL_AGENT_CAC;
  float* buf = blk + 3 * (size_t)P + THRL_MLP_HEADER_WORDS;
  const float gam = (float)spec.gamma;
  const float *w1 = blk, *b1 = blk + H;
  auto entry = [&](int nn) {
    int sl = head + nn;
    if (sl >= cap) sl -= cap;
    return buf + (size_t)sl * EW;
  };
  __syncwarp();
  for (int i = lane * 32; i < 3 * P; i += 32 * 32) asm volatile("prefetch.global.L2 [%0];" ::"l"(blk + i));
  // ---- 1. coefficients and ranks
  bool bad = false;
  if (kind == THRL_AGENT_REINFORCE) {
    // discounted returns, newest to oldest (:177-180): the float32 recurrence itself, 32 transitions per round
    float carry = 0.0f;
    bool first = true;
    double part = 0.0;
    for (int hi = N; hi > 0; hi -= 32) {
      const int nn = hi - 32 + lane;
      float rv = 0.0f;
      if (nn >= 0) rv = entry(nn)[2];
      float dv = 0.0f;
      const int lmin = hi >= 32 ? 0 : 32 - hi;
      for (int l = 31; l >= lmin; --l) {
        const float rl = __shfl_sync(kFull, rv, l);
        const float d = first ? rl : __fadd_rn(rl, __fmul_rn(gam, carry));
        first = false;
        carry = d;
        if (lane == l) dv = d;
      }
      if (nn >= 0) { entry(nn)[2] = dv; part = __dadd_rn(part, (double)dv); }  // kept in the buffer like the order-exact kernel
    }
    __syncwarp();
    const float mean = (float)__ddiv_rn(warp_sum(part), (double)N);
    double ss = 0.0;
    for (int nn = lane; nn < N; nn += 32) {
      const double d = __dsub_rn((double)entry(nn)[2], (double)mean);
      ss = __dadd_rn(ss, __dmul_rn(d, d));
    }
    const float sd = (float)sqrt(__ddiv_rn(warp_sum(ss), (double)(N - 1)));  // unbiased std (:181)
    const float invN = __fdiv_rn(1.0f, (float)N);
    for (int nn = lane; nn < N; nn += 32) {
      const float* en = entry(nn);
      const float ca = __fmul_rn(__fdiv_rn(__fsub_rn(en[2], mean), sd), invN);  // d loss / d logits = (p - onehot) * G / N (:185)
      bad |= !isfinite(ca);
      xs[nn] = make_float4(__int_as_float(pwc_rank_lane(th, H, pwc_ukey(en[0]))), ca, 0.0f, 0.0f);
    }
  } else if (ac) {
    double Rp = 0.0, Dp = 0.0;
    for (int nn = lane; nn < N; nn += 32) {  // d_i = gamma * v(s'_i) - v(s_i) (:289)
      const float* en = entry(nn);
      const float s = en[0], s2 = en[3];
      const int r = pwc_rank_lane(th, H, pwc_ukey(s)), r2 = pwc_rank_lane(th, H, pwc_ukey(s2));
      const float v = pwc_eval(pwc_ld2(tab + (size_t)r * ncp + A), s), vp = pwc_eval(pwc_ld2(tab + (size_t)r2 * ncp + A), s2);
      const float d = __fsub_rn(__fmul_rn(gam, vp), v);
      Rp = __dadd_rn(Rp, (double)en[2]);
      Dp = __dadd_rn(Dp, (double)d);
      xs[nn] = make_float4(__int_as_float(r | (r2 << 16)), 0.0f, 0.0f, d);
    }
    const float fN = (float)N, fR = (float)warp_sum(Rp), fD = (float)warp_sum(Dp);
    const float invN2 = __fdiv_rn(1.0f, __fmul_rn(fN, fN));
    for (int nn = lane; nn < N; nn += 32) {  // the [N,N] advantage broadcast collapsed as in oracle ac_train
      float4 q = xs[nn];
      const float r = entry(nn)[2];
      q.y = __fmul_rn(__fadd_rn(__fmul_rn(fN, r), fD), invN2);                       // actor weight (N r_j + D) / N^2
      q.z = __fmul_rn(-2.0f, __fmul_rn(__fadd_rn(fR, __fmul_rn(fN, q.w)), invN2));   // dL/dv_j; dL/dv'_j = -gamma * that
      bad |= !isfinite(q.y) || !isfinite(q.z);
      xs[nn] = q;
    }
  } else {  // CAC: closed form of the [N,N] loss via five moments (oracle cac_train); the whole scalar chain per lane
    double Sr = 0.0, Sl = 0.0, Sl2 = 0.0, Srl = 0.0, Srl2 = 0.0;
    for (int nn = lane; nn < N; nn += 32) {
      const float* en = entry(nn);
      const float a_ = __fadd_rn(5e-5f, __fmul_rn(__fsub_rn(1.0f, 1e-4f), en[1]));
      const float ratio = __fdiv_rn(a_, __fsub_rn(1.0f, a_));
      const double l = (double)(float)det_log((double)ratio), r = (double)en[2];
      Sr = __dadd_rn(Sr, r);
      Sl = __dadd_rn(Sl, l);
      Sl2 = __dadd_rn(Sl2, __dmul_rn(l, l));
      Srl = __dadd_rn(Srl, __dmul_rn(r, l));
      Srl2 = __dadd_rn(Srl2, __dmul_rn(__dmul_rn(r, l), l));
    }
    Sr = warp_sum(Sr); Sl = warp_sum(Sl); Sl2 = warp_sum(Sl2); Srl = warp_sum(Srl); Srl2 = warp_sum(Srl2);
    const double dN = (double)N, invN2 = __ddiv_rn(1.0, __dmul_rn(dN, dN));
    for (int nn = lane; nn < N; nn += 32) {
      const float* en = entry(nn);
      const float s = en[0], s2 = en[3];
      const int r = pwc_rank_lane(th, H, pwc_ukey(s)), r2 = pwc_rank_lane(th, H, pwc_ukey(s2));
      const double2* row = tab + (size_t)r * ncp;
      const float zmu = pwc_eval(pwc_ld2(row), s), zsd = pwc_eval(pwc_ld2(row + 1), s), v = pwc_eval(pwc_ld2(row + 2), s);
      const float vp = pwc_eval(pwc_ld2(tab + (size_t)r2 * ncp + 2), s2);
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
      const float dzmu = __fmul_rn(gmu, __fmul_rn(4.0f, __fsub_rn(1.0f, __fmul_rn(t, t))));
      const float dzsd = __fmul_rn(gsd, det_sigmoidf(zsd));
      xs[nn] = make_float4(__int_as_float(r | (r2 << 16)), dzmu, dzsd, cv);  // dL/dv' = -gamma * cv
    }
  }
  if (__any_sync(kFull, bad)) { pwc_nan_update(blk, spec, g, lane); return; }
  __syncwarp();
  if (kind != THRL_AGENT_REINFORCE) {
    // Consecutive transitions of an episode share a state: s'_n is s_(n+1), bit for bit.  The value-head event at s'_n is then
    // carried by the event at s_(n+1) (xe = its dL/dv' = -gamma * dL/dv_n) and leaves the event list (rank H + 1).
    for (int nn = lane; nn < N; nn += 32) {
      const float* en = entry(nn);
      float ex = 0.0f;
      if (nn > 0 && __float_as_int(entry(nn - 1)[3]) == __float_as_int(en[0])) {
        const float4 qp = xs[nn - 1];
        ex = __fmul_rn(-gam, cac ? qp.w : qp.z);
      }
      xe[nn] = ex;
    }
    __syncwarp();
    for (int nn = lane; nn + 1 < N; nn += 32) {
      if (__float_as_int(entry(nn)[3]) == __float_as_int(entry(nn + 1)[0])) {
        const int rb = __float_as_int(xs[nn].x);
        xs[nn].x = __int_as_float((rb & 0xffff) | ((H + 1) << 16));
      }
    }
    __syncwarp();
  }
  // ---- 2. events in interval order
  const int NE = pwc_sort_events(xs, N, kind == THRL_AGENT_REINFORCE ? N : 2 * N, H, hist, evs, lane);
  __syncwarp();
  // ---- 3. the sweep
  {
    const bool col = !cac && lane < A, use = lane < NC;
    const int vc = NC - 1;  // value-head column of ActorCritic / CAC
    const float cen = cac ? 0.0f : __fmul_rn((float)spec.entropy, __fdiv_rn(1.0f, (float)N));
    double P0 = 0.0, P1 = 0.0;
    int rcur = 0;  // rows < rcur of pf are written
    for (int i0 = 0; i0 < NE; i0 += 32) {
      // lane l fetches event i0 + l: rank, state, coefficients
      int my_r = 0, my_a = -1;
      float my_s = 0.0f, my_c0 = 0.0f, my_c1 = 0.0f, my_c2 = 0.0f, my_c3 = 0.0f;
      bool my_next = false;  // event at the transition's next state: value head only
      if (i0 + lane < NE) {
        const int e = (int)evs[i0 + lane];
        my_next = e >= N;
        const int nn = my_next ? e - N : e;
        const float4 q = xs[nn];
        const float* en = entry(nn);
        const int rb = __float_as_int(q.x);
        if (!my_next) {
          my_r = rb & 0xffff; my_s = en[0]; my_a = __float_as_int(en[1]);
          my_c0 = q.y; my_c1 = q.z; my_c2 = q.w;
          if (kind != THRL_AGENT_REINFORCE) my_c3 = xe[nn];
        } else {
          my_r = (rb >> 16) & 0xffff; my_s = en[3];
          my_c0 = __fmul_rn(-gam, cac ? q.w : q.z);
        }
      }
      const int cnt = NE - i0 < 32 ? NE - i0 : 32;
      int r_ld = -1;  // interval whose table row t_cur holds
      double2 t_cur = make_double2(0.0, 0.0);
      {
        const int r0 = __shfl_sync(kFull, my_r, 0);
        const bool nx0 = __shfl_sync(kFull, (int)my_next, 0) != 0;
        if (!cac && !nx0) { r_ld = r0; if (use) t_cur = pwc_ld2(tab + (size_t)r0 * ncp + lane); }
      }
      for (int l = 0; l < cnt; ++l) {
        const int r = __shfl_sync(kFull, my_r, l);
        const bool nxt = __shfl_sync(kFull, (int)my_next, l) != 0;
        const float s = __shfl_sync(kFull, my_s, l), c0 = __shfl_sync(kFull, my_c0, l), c1 = __shfl_sync(kFull, my_c1, l);
        const float c2 = __shfl_sync(kFull, my_c2, l), c3 = __shfl_sync(kFull, my_c3, l);
        const int a = __shfl_sync(kFull, my_a, l);
        // the row of the event after this one is requested before this one's arithmetic
        double2 t_nx = t_cur;
        int r_nx = r_ld;
        if (!cac && l + 1 < cnt) {
          const int rn = __shfl_sync(kFull, my_r, l + 1);
          const bool nn2 = __shfl_sync(kFull, (int)my_next, l + 1) != 0;
          if (!nn2 && rn != r_ld) { r_nx = rn; if (use) t_nx = pwc_ld2(tab + (size_t)rn * ncp + lane); }
        }
        while (rcur <= r) {  // the sums over the intervals < rcur are complete
          if (use) __stcg(pf + (size_t)rcur * ncp + lane, make_double2(P0, P1));
          ++rcur;
        }
        double dz = 0.0;
        if (nxt) {
          if (lane == vc) dz = (double)c0;
        } else if (cac) {
          dz = lane == 2 ? __dadd_rn((double)c2, (double)c3) : (double)(lane == 0 ? c0 : c1);
        } else {
          const float z = use ? pwc_eval(t_cur, s) : 0.0f;
          const float mx = warp_max(col ? z : NegInf<float>::v());
          const float ex = col ? pwc_expf(__fsub_rn(z, mx)) : 0.0f;
          const float inv = __frcp_rn(warp_sum(ex));  // pi_k = e_k * (1 / sum): one reciprocal per event
          if (cen != 0.0f) {  // entropy regulariser (agents.py:187-189, 298-300): + c_e / N * p_k (log p_k + H), H = -sum p log p
            const float pk = col ? __fmul_rn(ex, inv) : 0.0f;
            const float lp = pk > 0.0f ? (float)det_log((double)pk) : 0.0f;
            const float Hn = -warp_sum(__fmul_rn(pk, lp));
            if (col) dz = (double)__fadd_rn(__fmul_rn(__fsub_rn(pk, lane == a ? 1.0f : 0.0f), c0),
                                            __fmul_rn(cen, __fmul_rn(pk, __fadd_rn(lp, Hn))));
            else if (ac && lane == A) dz = __dadd_rn((double)c1, (double)c3);
          } else if (col) {
            dz = (double)__fmul_rn(__fsub_rn(__fmul_rn(ex, inv), lane == a ? 1.0f : 0.0f), c0);
          } else if (ac && lane == A) {
            dz = __dadd_rn((double)c1, (double)c3);
          }
        }
        P0 = __dadd_rn(P0, dz);
        P1 = __dadd_rn(P1, __dmul_rn(dz, (double)s));
        t_cur = t_nx;
        r_ld = r_nx;
      }
    }
    while (rcur <= H + 1) {
      if (use) __stcg(pf + (size_t)rcur * ncp + lane, make_double2(P0, P1));
      ++rcur;
    }
    if (use) g[pwc_bias(spec, lane)] = (float)P0;
  }
  __syncwarp();
  // ---- 4. lane = rank q
  for (int q0 = 0; q0 < H; q0 += 32) {
    const int q = q0 + lane;
    if (q < H) {
      const unsigned o = ord[q];
      const int j = (int)(o & 0x7fffu);
      const bool leave = (o & 0x8000u) != 0;
      const double w = (double)w1[j], b = (double)b1[j];
      double gw = 0.0, gb = 0.0;
      // four columns per round, every load of the round issued before its arithmetic (the prefix sums sit in L2)
      for (int c0 = 0; c0 < NC; c0 += 4) {
        double2 pre[4], tt[4];
        float cwf[4];
        int wi[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int c = c0 + u < NC ? c0 + u : NC - 1;
          pre[u] = pwc_ld2(pf + (size_t)(q + 1) * ncp + c);
          tt[u] = pwc_ld2(pf + (size_t)(H + 1) * ncp + c);
          wi[u] = pwc_wrow(spec, c) + j;
          cwf[u] = blk[wi[u]];
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          if (c0 + u < NC) {
            const double M0 = leave ? pre[u].x : __dsub_rn(tt[u].x, pre[u].x), M1 = leave ? pre[u].y : __dsub_rn(tt[u].y, pre[u].y);
            g[wi[u]] = (float)__dadd_rn(__dmul_rn(w, M1), __dmul_rn(b, M0));
            const double cw = (double)cwf[u];
            gb = __dadd_rn(gb, __dmul_rn(cw, M0));
            gw = __dadd_rn(gw, __dmul_rn(cw, M1));
          }
        }
      }
      g[j] = (float)gw;
      g[H + j] = (float)gb;
    }
  }
  pwl_clip_adam(blk, spec, g, lane);
}

// kTwo: the game is two discrete-action MLP agents (Reinforce / ActorCritic) and nothing else -- the BASELINE C5 shape with
// demand noise.  Its episode loop keeps both agents' thresholds, tables and shapes in registers and evaluates the two policies
// side by side; every other game takes the general loop.
template <typename QT, bool kTwo>
__global__ void __launch_bounds__(32 * kPwcMaxWarps, 1) mlp_scan_pwc(const __grid_constant__ PwcParams p) {
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
  float* zf = reinterpret_cast<float*>(slot + p.off_zf);          // [T][n] MLP agents that sample: the uniform (discrete) / the normal deviate (CAC)
  double* newa = reinterpret_cast<double*>(slot + p.off_newa);    // [T]
  uint16_t* rowbuf = reinterpret_cast<uint16_t*>(slot + p.off_row);
  QT* oldv = reinterpret_cast<QT*>(slot + p.off_old);
  double* hpw = reinterpret_cast<double*>(slot + p.off_hp);       // [n][5]
  unsigned char* wsw = p.ws + ((size_t)blockIdx.x * warps_per_cta + warp) * p.ws_warp_bytes;
  double2* bkt = reinterpret_cast<double2*>(wsw + p.ws_bkt);
  float* gws = reinterpret_cast<float*>(wsw + p.ws_grad);
  float4* xs = reinterpret_cast<float4*>(wsw + p.ws_xs);
  float* xe = reinterpret_cast<float*>(wsw + p.ws_xe);
  uint32_t* evs = reinterpret_cast<uint32_t*>(wsw + p.ws_evs);
  int* hist = reinterpret_cast<int*>(slot + p.off_hist);  // [Hmax + 3] event counts per interval (updates only)

  int my_cap = 0, my_minmem = 0x7fffffff, my_lut = 0, my_kind = 0, my_len = 0;
  int my_mcap = 0, my_EW = 3;  // MLP agent: buffer capacity in the slab, words per entry
  long long my_boff = 0;       // MLP agent: float offset of the header inside the run slab
  float my_msf = 1.f, my_sf = 1.f;
  if (is_agent) {
    const ThrlAgentSpec& s = G.agent[lane];
    my_cap = s.capacity; my_minmem = s.min_memory; my_kind = s.kind;
    my_msf = (float)s.max_state; my_sf = (float)s.states;
    for (int j = 0; j < lane; ++j) my_lut += G.agent[j].actions;
    if (s.kind != THRL_AGENT_QTABLE) {
      my_mcap = G.mlp_buffer_len[lane];
      my_EW = mlp_entry_words(s);
      my_boff = s.mlp_offset + 3 * (long long)mlp_P(s);
    }
  }
  const bool never_fires = my_minmem > my_cap;
  const int lead = lead_exact_floats(G);
  const bool tracing = p.trace_actions || p.trace_rewards || p.trace_prices;
  const double dT = (double)T, rT = __ddiv_rn(1.0, dT);

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
    // MLP agent (lane = agent): deque(maxlen = capacity) as length + next slot to write
    int m_len = 0, m_wr = 0;
    int32_t* m_hdr = nullptr;
    float* m_buf = nullptr;
    if (is_agent && my_kind != THRL_AGENT_QTABLE) {
      m_hdr = reinterpret_cast<int32_t*>(slab + my_boff);
      m_buf = slab + my_boff + THRL_MLP_HEADER_WORDS;
      if (my_mcap > 0) {
        m_len = m_hdr[1];
        m_wr = m_hdr[2] + m_len;
        if (m_wr >= my_mcap) m_wr -= my_mcap;
      }
    }
    __syncwarp();
    if (lane == 0) P[pos] = price;
    for (int i = 0; i < n; ++i) {
      const ThrlAgentSpec& s = G.agent[i];
      if (s.kind == THRL_AGENT_QTABLE) continue;
      pwc_build(slab + s.mlp_offset, s, reinterpret_cast<unsigned*>(slot + p.off_th[i]), reinterpret_cast<uint16_t*>(wsw + p.ws_ord[i]),
                reinterpret_cast<double2*>(wsw + p.ws_tab[i]), p.ncp[i], lane);
    }
    __syncwarp();

    for (int e = 0; e < E; ++e) {
      const uint32_t eabs = (uint32_t)(p.epoch_begin + e);
      const long long step0 = (r * E + e) * (long long)T;

      // ---- per-episode draws (QTable: agents.py:81-82; MLP: forced sample in replay modes, else the deviate it is drawn from)
      for (int idx = lane; idx < T * n; idx += 32) {
        const int t = idx / n, i = idx - t * n;
        const int kind = G.agent[i].kind;
        int v;
        if (kind != THRL_AGENT_QTABLE) {
          v = p.rng_mode == THRL_RNG_PHILOX ? -2 : p.replay_ra[step0 * n + idx];
          if (v < 0) {  // a replay stream may leave this agent's sample to the device
            v = -2;
            uint32_t x[4];
            philox4x32_10(gid, eabs, (uint32_t)t, (uint32_t)(i >> 1) | (kStreamAct << 16), p.k0, p.k1, x);
            if (kind == THRL_AGENT_CAC) {
              const unsigned long long m = ((unsigned long long)x[2 * (i & 1)] << 21) | (unsigned long long)(x[2 * (i & 1) + 1] >> 11);
              zf[idx] = (float)det_norminv(((double)m + 0.5) * (1.0 / 9007199254740992.0));
            } else {
              zf[idx] = __fmul_rn((float)(x[2 * (i & 1)] >> 8), 1.0f / 16777216.0f);
            }
          }
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
      if (kTwo) {
        const int A0 = G.agent[0].actions, A1 = G.agent[1].actions, H0 = G.agent[0].hidden, H1 = G.agent[1].hidden;
        const int per0 = (H0 + 31) >> 5, per1 = (H1 + 31) >> 5;
        const unsigned* th0 = reinterpret_cast<const unsigned*>(slot + p.off_th[0]);
        const unsigned* th1 = reinterpret_cast<const unsigned*>(slot + p.off_th[1]);
        const double2* tb0 = reinterpret_cast<const double2*>(wsw + p.ws_tab[0]) + (lane < A0 ? lane : 0);
        const double2* tb1 = reinterpret_cast<const double2*>(wsw + p.ws_tab[1]) + (lane < A1 ? lane : 0);
        const int ncp0 = p.ncp[0], ncp1 = p.ncp[1];
        const bool in0 = lane < A0, in1 = lane < A1;
        const unsigned last0 = 1u << (A0 - 1), last1 = 1u << (A1 - 1);
        const double* aq1p = lutAQ + A0;
        const double* xtp = lutXT + (lane == 1 ? A0 : 0);
        const bool append = is_agent && my_mcap > 0;
        for (int t = 0; t < T; ++t) {
          const int2 v = *reinterpret_cast<const int2*>(pre + 2 * t);  // forced action, or -2: sample
          int k0 = v.x, k1 = v.y;
          if ((v.x | v.y) < 0) {
            const float2 dev = *reinterpret_cast<const float2*>(zf + 2 * t);
            const float sf = (float)price;
            const unsigned key = pwc_ukey(sf);
            const int r0 = pwc_rank_warp(th0, th0 + 32 * per0, H0, per0, key, lane);
            const int r1 = pwc_rank_warp(th1, th1 + 32 * per1, H1, per1, key, lane);
            const double2 t0 = pwc_ld2(tb0 + (size_t)r0 * ncp0), t1 = pwc_ld2(tb1 + (size_t)r1 * ncp1);
            const float z0 = pwc_eval(t0, sf), z1 = pwc_eval(t1, sf);
            // both softmaxes side by side.  pi_k = e_k / sum and `cumsum(pi)[k] > u` (agents.py:160-163) are evaluated as
            // cumsum(e)[k] > u * sum: one scan yields both the running sums and (last column) the total
            const float m0 = warp_max(in0 ? z0 : NegInf<float>::v()), m1 = warp_max(in1 ? z1 : NegInf<float>::v());
            float c0 = in0 ? pwc_expf(__fsub_rn(z0, m0)) : 0.0f, c1 = in1 ? pwc_expf(__fsub_rn(z1, m1)) : 0.0f;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
              const float u0 = __shfl_up_sync(kFull, c0, off), u1 = __shfl_up_sync(kFull, c1, off);
              if (lane >= off) { c0 = __fadd_rn(c0, u0); c1 = __fadd_rn(c1, u1); }
            }
            const float tot0 = __shfl_sync(kFull, c0, 31), tot1 = __shfl_sync(kFull, c1, 31);
            const unsigned b0 = __ballot_sync(kFull, in0 && c0 > __fmul_rn(dev.x, tot0)) | last0;
            const unsigned b1 = __ballot_sync(kFull, in1 && c1 > __fmul_rn(dev.y, tot1)) | last1;
            if (v.x < 0) k0 = __ffs(b0) - 1;
            if (v.y < 0) k1 = __ffs(b1) - 1;
          }
          const int kmine = lane == 0 ? k0 : k1;
          const double aq0 = lutAQ[k0], aq1 = aq1p[k1];  // environments.py:27 sum(A): 0 + A[0] + A[1]
          const double Q = __dadd_rn(__dadd_rn(0.0, aq0), aq1);
          const double na = p.noisy ? newa[t] : G.a;
          const double pn = __dsub_rn(na, __dmul_rn(G.b, Q));
          const double next_price = pn > 0.0 ? pn : (pn != pn ? pn : 0.0);
          const double rew = __dmul_rn(next_price, lane == 0 ? aq0 : aq1);
          if (is_agent) {
            rlog = __dadd_rn(rlog, pwc_div(rew, dT, rT));
            alog = __dadd_rn(alog, xtp[kmine]);
            if (append) {  // memory.append; replay(cast) makes state and reward float32 (buffers.py:28-38, agents.py:142)
              float* en = m_buf + (size_t)m_wr * my_EW;
              en[0] = (float)price;
              en[1] = __int_as_float(kmine);
              en[2] = (float)rew;
              if (my_EW == 4) en[3] = (float)next_price;
              m_wr = m_wr + 1 == my_mcap ? 0 : m_wr + 1;
              m_len = m_len < my_mcap ? m_len + 1 : my_mcap;
            }
            if (tracing) {
              if (p.trace_actions) p.trace_actions[(step0 + t) * n + lane] = kmine;
              if (p.trace_rewards) p.trace_rewards[(step0 + t) * n + lane] = rew;
            }
          }
          if (tracing && lane == 0 && p.trace_prices) p.trace_prices[step0 + t] = next_price;
          price = next_price;
        }
      } else
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
        // MLP agents that sample: the interval of the price, one table row, softmax + inverse CDF (agents.py:160-163) or
        // sigmoid(Normal(mu, std).sample()) (agents.py:374-378)
        unsigned samp = __ballot_sync(kFull, is_agent && k == -2);
        if (samp) {
          const float sf = (float)price;
          const unsigned key = pwc_ukey(sf);
          while (samp) {
            // two agents per round: both ranks and both table-row loads are issued before either policy is evaluated
            int ids[2];
            double2 rows[2];
#pragma unroll
            for (int u = 0; u < 2; ++u) {
              ids[u] = -1;
              rows[u] = make_double2(0.0, 0.0);
              if (samp) {
                const int i = __ffs(samp) - 1;
                samp &= samp - 1;
                ids[u] = i;
                const ThrlAgentSpec& s = G.agent[i];
                const int per = (s.hidden + 31) >> 5;
                const unsigned* thi = reinterpret_cast<const unsigned*>(slot + p.off_th[i]);
                const int rk = pwc_rank_warp(thi, thi + 32 * per, s.hidden, per, key, lane);
                const int nc = s.kind == THRL_AGENT_CAC ? 2 : s.actions;
                if (lane < nc) rows[u] = pwc_ld2(reinterpret_cast<const double2*>(wsw + p.ws_tab[i]) + (size_t)rk * p.ncp[i] + lane);
              }
            }
#pragma unroll
            for (int u = 0; u < 2; ++u) {
              const int i = ids[u];
              if (i < 0) continue;
              const ThrlAgentSpec& s = G.agent[i];
              const float dev = zf[t * n + i];
              const float z = pwc_eval(rows[u], sf);
              int ks;
              if (s.kind == THRL_AGENT_CAC) {
                const float zmu = __shfl_sync(kFull, z, 0), zsd = __shfl_sync(kFull, z, 1);
                const float mu = __fmul_rn(4.0f, det_tanhf(zmu)), sd = det_softplusf(zsd);
                ks = __float_as_int(det_sigmoidf(__fadd_rn(mu, __fmul_rn(sd, dev))));
              } else {
                ks = pwc_sample(z, s.actions, dev, lane);
              }
              if (lane == i) k = ks;
            }
          }
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
          rlog = __dadd_rn(rlog, pwc_div(rew, dT, rT));
          alog = __dadd_rn(alog, xt);
          if (my_kind == THRL_AGENT_QTABLE) {
            act[lane * Hp + pos] = (uint8_t)k;
            my_len = my_len < my_cap ? my_len + 1 : my_cap;
          } else if (my_mcap > 0) {
            // memory.append; replay(cast) makes state and reward float32 (buffers.py:28-38, agents.py:142)
            float* en = m_buf + (size_t)m_wr * my_EW;
            en[0] = (float)price;
            en[1] = __int_as_float(k);
            en[2] = (float)rew;
            if (my_EW == 4) en[3] = (float)next_price;
            m_wr = m_wr + 1 == my_mcap ? 0 : m_wr + 1;
            m_len = m_len < my_mcap ? m_len + 1 : my_mcap;
          }
          if (tracing) {
            if (p.trace_actions) p.trace_actions[(step0 + t) * n + lane] = k;
            if (p.trace_rewards) p.trace_rewards[(step0 + t) * n + lane] = rew;
          }
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
        if (s.kind != THRL_AGENT_QTABLE) {
          const int cap = G.mlp_buffer_len[i];
          if (cap == 0) continue;
          const int len = __shfl_sync(kFull, m_len, i);
          if (len < s.min_memory) continue;
          int hd = __shfl_sync(kFull, m_wr, i) - len;  // oldest buffered transition
          if (hd < 0) hd += cap;
          float* blk = slab + s.mlp_offset;
          unsigned* th = reinterpret_cast<unsigned*>(slot + p.off_th[i]);
          uint16_t* ord = reinterpret_cast<uint16_t*>(wsw + p.ws_ord[i]);
          double2* itab = reinterpret_cast<double2*>(wsw + p.ws_tab[i]);
          __syncwarp();  // the episode's buffer stores are visible to every lane
          pwc_train(blk, s, cap, hd, len, th, ord, itab, bkt, p.ncp[i], gws, xs, xe, evs, hist, lane);
          if (lane == i) { m_len = 0; m_wr = 0; }  // :194 memory.empty()
          pwc_build(blk, s, th, ord, itab, p.ncp[i], lane);
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

    // ---- write the run back (the MLP parameters were updated in place)
    if (is_agent) p.eps[r * n + lane] = hpw[lane * 5 + 4];
    if (lane == 0) p.price[r] = price;
    if (is_agent && my_kind != THRL_AGENT_QTABLE && my_mcap > 0) {
      int hd = m_wr - m_len;
      if (hd < 0) hd += my_mcap;
      m_hdr[1] = m_len;
      m_hdr[2] = m_len ? hd : 0;
    }
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

}  // namespace thrl
