// thrl_aux_kernels.cuh — device init (QTable.__init__ / NoisyPriceState.reset) and the greedy evaluation rollout.
#pragma once
#include "thrl_device.cuh"

namespace thrl {

// ---------------------------------------------------------------- deterministic N(0,1) (DESIGN.md "Device init")
// Only + - * / sqrt and frexp, in a fixed order; compiled with --fmad=false this is bit-identical to the oracle's
// det_log / det_norminv (oracle/thrl_oracle.c), so device-initialised tables can be reproduced on the host.
__device__ inline double det_log(double x) {
  int e;
  double m = frexp(x, &e);
  if (m < 0.70710678118654752) { m = m * 2.0; e -= 1; }
  const double s = (m - 1.0) / (m + 1.0), s2 = s * s;
  double p = 1.0 / 27.0;
  for (int k = 25; k >= 1; k -= 2) p = p * s2 + 1.0 / (double)k;
  return (double)e * 0.6931471805599453094 + 2.0 * s * p;
}
__device__ inline double det_norminv(double p) {
  const double a0 = -3.969683028665376e+01, a1 = 2.209460984245205e+02, a2 = -2.759285104469687e+02,
               a3 = 1.383577518672690e+02, a4 = -3.066479806614716e+01, a5 = 2.506628277459239e+00;
  const double b0 = -5.447609879822406e+01, b1 = 1.615858368580409e+02, b2 = -1.556989798598866e+02,
               b3 = 6.680131188771972e+01, b4 = -1.328068155288572e+01;
  const double c0 = -7.784894002430293e-03, c1 = -3.223964580411365e-01, c2 = -2.400758277161838e+00,
               c3 = -2.549732539343734e+00, c4 = 4.374664141464968e+00, c5 = 2.938163982698783e+00;
  const double d0 = 7.784695709041462e-03, d1 = 3.224671290700398e-01, d2 = 2.445134137142996e+00,
               d3 = 3.754408661907416e+00;
  const double plow = 0.02425;
  if (p < plow || p > 1.0 - plow) {
    const double pp = p < plow ? p : 1.0 - p;
    const double q = sqrt(-2.0 * det_log(pp));
    const double num = ((((c0 * q + c1) * q + c2) * q + c3) * q + c4) * q + c5;
    const double den = (((d0 * q + d1) * q + d2) * q + d3) * q + 1.0;
    const double x = num / den;
    return p < plow ? x : -x;
  }
  const double q = p - 0.5, r = q * q;
  const double num = (((((a0 * r + a1) * r + a2) * r + a3) * r + a4) * r + a5) * q;
  const double den = ((((b0 * r + b1) * r + b2) * r + b3) * r + b4) * r + 1.0;
  return num / den;
}

struct InitParams {
  ThrlGame game;
  long long n_runs, run_id0;
  uint32_t k0, k1;
  int f64;
  const double* hp;
  double eps0[THRL_MAX_AGENTS];
  void* q;
  uint32_t* counter;
  double* eps;
  double* price;
  float* mlp;
};

// grid.y = run, threads stride over the cell pairs of all agents: coalesced stores of the run slab.
__global__ void __launch_bounds__(256) qtable_init(const __grid_constant__ InitParams p) {
  const ThrlGame& G = p.game;
  const int n = G.n_agents;
  for (long long r = blockIdx.y; r < p.n_runs; r += gridDim.y) {
    const uint32_t gid = (uint32_t)(p.run_id0 + r);
    for (int i = 0; i < n; ++i) {
      const ThrlAgentSpec& s = G.agent[i];
      if (s.kind != THRL_AGENT_QTABLE) {
        // nn.Linear default init (kaiming_uniform(a=sqrt(5)) == U(-1/sqrt(fan_in), 1/sqrt(fan_in)) for weight and bias,
        // agents.py:137-138); Adam state, header and transition buffer zeroed
        const bool ac = s.kind == THRL_AGENT_ACTORCRITIC;
        const long long Pn = s.kind == THRL_AGENT_CAC ? 5LL * s.hidden + 3
                                                      : 2LL * s.hidden + (long long)s.actions * s.hidden + s.actions + (ac ? s.hidden + 1 : 0);
        const long long words = 3 * Pn + THRL_MLP_HEADER_WORDS + (s.kind == THRL_AGENT_REINFORCE ? 3LL : 4LL) * G.mlp_buffer_len[i];
        float* blk = p.mlp + r * G.mlp_stride + s.mlp_offset;
        const float b_fc1 = 1.0f, b_pi = (float)__ddiv_rn(1.0, sqrt((double)s.hidden));
        for (long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x; w < words; w += (long long)gridDim.x * blockDim.x) {
          float v = 0.0f;
          if (w < Pn) {
            uint32_t x[4];
            philox4x32_10(gid, (uint32_t)(w >> 1), (uint32_t)i, kStreamInitQ << 16, p.k0, p.k1, x);
            const double u = (w & 1) ? u53(x[2], x[3]) : u53(x[0], x[1]);
            const float bound = (w < 2LL * s.hidden) ? b_fc1 : b_pi;
            v = __fmul_rn((float)__dsub_rn(__dmul_rn(2.0, u), 1.0), bound);
            if (ac && w == Pn - 1) v = 1000.0f;  // fc_v.bias.data.fill_(1000.0), agents.py:244
          }
          blk[w] = v;
        }
        continue;
      }
      const double gamma = p.hp ? p.hp[(r * n + i) * 4 + 1] : s.gamma;
      const double base = 12.5 / (1.0 - gamma);  // agents.py:29
      const long long cells = (long long)(s.states + 1) * s.actions;
      for (long long pr = (long long)blockIdx.x * blockDim.x + threadIdx.x; pr * 2 < cells;
           pr += (long long)gridDim.x * blockDim.x) {
        uint32_t x[4];
        philox4x32_10(gid, (uint32_t)pr, (uint32_t)i, kStreamInitQ << 16, p.k0, p.k1, x);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const long long c = pr * 2 + h;
          if (c >= cells) break;
          const unsigned long long m = ((unsigned long long)x[2 * h] << 21) | (unsigned long long)(x[2 * h + 1] >> 11);
          const double u = ((double)m + 0.5) * (1.0 / 9007199254740992.0);
          const double v = base + det_norminv(u);
          // the draw of a cell depends on its logical index row * actions + col only, not on the row padding
          const long long idx = r * G.run_stride + s.table_offset + (c / s.actions) * s.row_stride + (c % s.actions);
          if (p.f64) reinterpret_cast<double*>(p.q)[idx] = v;
          else reinterpret_cast<float*>(p.q)[idx] = (float)v;
          if (p.counter) p.counter[idx] = 0u;  // agents.py:45
        }
      }
      const int pad = s.row_stride - s.actions;  // padded layout (include/thrl.h): padding cells are zeroed once, here
      for (long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x; c < (long long)(s.states + 1) * pad;
           c += (long long)gridDim.x * blockDim.x) {
        const long long idx = r * G.run_stride + s.table_offset + (c / pad) * s.row_stride + s.actions + (c % pad);
        if (p.f64) reinterpret_cast<double*>(p.q)[idx] = 0.0;
        else reinterpret_cast<float*>(p.q)[idx] = 0.0f;
        if (p.counter) p.counter[idx] = 0u;
      }
    }
    if (blockIdx.x == 0 && threadIdx.x < n) p.eps[r * n + threadIdx.x] = p.eps0[threadIdx.x];
    if (blockIdx.x == 0 && threadIdx.x == 0) {  // environments.py:15-16 uniform(0, a)
      uint32_t x[4];
      philox4x32_10(gid, 0u, 0u, kStreamInitP << 16, p.k0, p.k1, x);
      p.price[r] = 0.0 + (G.a - 0.0) * u53(x[0], x[1]);
    }
  }
}

struct EvalParams {
  ThrlGame game;
  long long n_runs;
  int iters;
  const void* q;
  const double* price0;
  const double* new_a;  // [R][iters][T] demand intercept per step, or NULL: the noise-free curve
  double* rewards;
  double* actions;
};

// utils.py:27-47 play_game + agents.py:91-92 get_action: warp per run, greedy on the f64 encode, no update.
template <typename QT>
__global__ void __launch_bounds__(256) greedy_eval(const __grid_constant__ EvalParams p) {
  const ThrlGame& G = p.game;
  const int lane = threadIdx.x & 31;
  const int n = G.n_agents, T = G.max_steps;
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long total = ((long long)gridDim.x * blockDim.x) >> 5;
  const double ab = __ddiv_rn(G.a, G.b);
  for (long long r = warp; r < p.n_runs; r += total) {
    const QT* tab = reinterpret_cast<const QT*>(p.q) + r * G.run_stride;
    for (int it = 0; it < p.iters; ++it) {
      double price = p.price0[r * p.iters + it];
      for (int t = 0; t < T; ++t) {
        double Q = 0.0, my_x = 0.0, my_aq = 0.0;
        for (int i = 0; i < n; ++i) {
          const ThrlAgentSpec& s = G.agent[i];
          const int row = upd_row(price, s.max_state, (double)s.states);
          const int k = row_argmax(tab + s.table_offset + (size_t)row * s.row_stride, s.actions, lane);
          const double x = scale_action(k, s.actions, s.action_lo, s.action_hi);
          const double aq = __dmul_rn(ab, x);
          Q = __dadd_rn(Q, aq);
          if (lane == i) { my_x = x; my_aq = aq; }
        }
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

// ---------------------------------------------------------------- learning-curve quantile statistics (utils.py:132-145)
// plot_learning_curve_conf: per run, pandas' ewm(halflife).mean() (adjust=True) of the per-epoch mean reward, summed over the
// agents; per epoch the median / quartiles of that over the runs.  Here: thread = run carries the run's EWM numerator
// num_t = num_{t-1} * (1 - alpha) + x_t across epochs (and calls), x_t = ((0 + r_0) + r_1) + ... the agents' log entries;
// the value num_t / den_t (den_t = sum_{i<=t} (1-alpha)^i is the same for every run: a host-computed array) goes into a
// fixed-bin histogram per epoch.  Counts are exact integers, so sharded histograms add up (NCCL all-reduce).
struct CurveHistParams {
  const double* rewards_log;  // [n_runs][E][n]
  long long n_runs;
  int E, n;
  double decay;               // 1 - alpha = 0.5^(1/halflife)
  const double* den;          // [E]
  double* ewm_num;            // [n_runs] in/out
  double lo, inv_width;       // bin = floor((v - lo) * inv_width), clamped to [0, n_bins)
  int n_bins;
  unsigned long long* hist;   // [E][n_bins] +=
};

__global__ void __launch_bounds__(256) curve_hist(const CurveHistParams p) {
  const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = r < p.n_runs;
  const int lane = threadIdx.x & 31;
  double num = live ? p.ewm_num[r] : 0.0;
  const double* row = p.rewards_log + (live ? r : 0) * (long long)p.E * p.n;
  for (int e = 0; e < p.E; ++e) {
    int bin = -1 - lane;  // idle lanes match nobody
    if (live) {
      double x = 0.0;
      for (int i = 0; i < p.n; ++i) x = __dadd_rn(x, row[(long long)e * p.n + i]);
      num = __dadd_rn(__dmul_rn(num, p.decay), x);
      const double v = __ddiv_rn(num, p.den[e]);
      const double b = floor(__dmul_rn(__dsub_rn(v, p.lo), p.inv_width));
      bin = b < 0.0 ? 0 : (b >= (double)p.n_bins ? p.n_bins - 1 : (int)b);
      if (!(b == b)) bin = p.n_bins - 1;  // NaN rewards land in the last bin
    }
    const unsigned same = __match_any_sync(kFull, bin);  // one atomic per distinct bin of the warp
    if (live && (same & ((1u << lane) - 1u)) == 0u) atomicAdd(p.hist + (long long)e * p.n_bins + bin, (unsigned long long)__popc(same));
  }
  if (live) p.ewm_num[r] = num;
}

}  // namespace thrl
