// thrl_scan_pwl.cuh — games made only of discrete-action MLP agents (Reinforce, th_rl/agents.py:119-194; ActorCritic,
// :222-305) on the noise-free demand curve: the BASELINE C5 family.
//
// Two facts make the per-run "1000 x 256 x 22 GEMMs" of such a game collapse:
//  (1) the network input is the price, and with discrete actions and no demand noise the price only takes the finitely
//      many values a - b*sum(A) of the joint actions (the LATTICE, <= 1024 distinct float32 states; plus the run's
//      arbitrary initial price, an "extra" state).  pi(.|s) and v(s) are therefore tabulated once per parameter update
//      (policy LUT, stored as a CDF so that acting is one shared-memory row read + ballot), and a batch of N
//      transitions is equivalent to at most NS weighted states: the per-sample loss coefficients are summed per state
//      (and per state x action) with fixed-point integer atomics, which makes the sums independent of the order.
//  (2) a 1 -> H -> A ReLU network of a scalar is piecewise linear with one breakpoint per hidden unit: unit j is active
//      on a prefix or a suffix of the sorted lattice (the float32 predicate fl(fl(s*w1)+b1) > 0 is monotone in s).
//      Sorting the units by breakpoint, one sweep over the lattice evaluates every logit as S1*s + S0 with running f64
//      sums (forward), and one sweep of f64 prefix sums of the per-state gradient coefficients yields every
//      d loss / d fc_pi.weight[k][j] = w1_j*M1[k][j] + b1_j*M0[k][j] and the fc1 gradients (backward).
// Work per update: O(NS*A + H*A) instead of O(N*H*A); no dense contraction is left, so there is nothing to put on the
// tensor cores.  The arithmetic is the same real-number computation as the reference's autograd graph, summed in a
// different order (f64 accumulation): results agree with oracle/thrl_oracle.c (which follows the reference's order and
// is pinned against torch) to float32 rounding, not bit for bit; tests/test_gpu_pwl.py states the tolerance.
// THRL_KERNEL=mixed selects the order-exact kernel (thrl_scan_mixed.cuh) instead.
#pragma once
#include "thrl_device.cuh"
#include "thrl_scan_mixed.cuh"  // mlp_P, mlp_entry_words, det_expf

namespace thrl {

#ifndef THRL_PWL_MAXWARPS
#define THRL_PWL_MAXWARPS 16
#endif
constexpr int kPwlMaxWarps = THRL_PWL_MAXWARPS;  // resident runs per CTA (launch bound)
constexpr int kPwlMaxJoint = 1024;   // joint actions (price table in the kernel parameters)
constexpr int kPwlMaxLattice = 1024; // distinct float32 lattice prices (<= joint actions)
constexpr int kPwlExtras = 4;        // off-lattice states one run may hold at a time (its initial price)

struct PwlParams {
  ThrlGame game;
  long long n_runs, run_id0;
  long long run_lo;      // first run of this launch (a call of several rounds is launched round by round: runs [run_lo, n_runs))
  int epoch_begin, E, rng_mode;
  uint32_t k0, k1;
  double* price;
  void* q;               // QTable agents of the same game (NULL when there are none)
  uint32_t* counter;
  double* eps;
  const double* hp;
  const double* replay_u;
  const int32_t* replay_ra;
  double* rewards_log;
  double* actions_log;
  long long n_log_runs;
  long long* stats;
  int32_t* trace_actions;
  double* trace_rewards;
  double* trace_prices;
  float* mlp;
  unsigned char* ws;  // per resident warp: accumulators, probability LUT, gradient, per-sample scratch
  long long ws_warp_bytes, ws_acc, ws_pf, ws_p, ws_cdf, ws_grad, ws_xs;
  int cdf_global;  // 1: the CDF LUT lives in the workspace (L2) instead of shared memory (large lattices: more resident runs)
  int J, NS, lut_total;
  int a_off[THRL_MAX_AGENTS];    // agent's offset in the per-action tables
  int jmul[THRL_MAX_AGENTS];     // joint index = sum_i action_i * jmul[i]
  int cdf_off[THRL_MAX_AGENTS];  // float offset of the agent's CDF LUT [NS + extras][A] (shared memory and ws_p alike)
  int val_off[THRL_MAX_AGENTS];  // float offset of the agent's v(s) LUT [NS + extras]
  int nq, dwords;                // QTable agents; words of the dirty-row bitmap (largest table)
  int qidx[THRL_MAX_AGENTS];     // ordinal among the QTable agents, -1 for MLP agents
  int L[THRL_MAX_AGENTS];        // QTable agent: transitions per episode-end update (0: it never fires)
  int off_tab[THRL_MAX_AGENTS];  // byte offset of the staged table in the warp's shared memory
  int cta_bytes, off_priceJ, off_rT, off_rF, off_slotof, off_urowJ;
  int warp_bytes, off_sv, off_cdf, off_val, off_pre, off_ev, off_ord, off_bkt, off_hpw, off_jrec, off_oldv, off_gq, off_arow, off_dirty;
  float slot_val[kPwlMaxLattice];   // lattice states, ascending
  uint16_t slot_of[kPwlMaxJoint];   // joint action -> lattice state
  double priceJ[kPwlMaxJoint];      // joint action -> next price (environments.py:25-33)
};

// shared-memory loads by 32-bit shared address (the episode loop of the 2-agent all-MLP kernel: one address add + LDS)
__device__ __forceinline__ uint32_t smem_u32(const void* ptr) { return (uint32_t)__cvta_generic_to_shared(ptr); }
__device__ __forceinline__ float lds_f32(uint32_t a) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ double lds_f64(uint32_t a) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ int lds_u16(uint32_t a) {
  unsigned short v;
  asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(a) : "memory");
  return (int)v;
}
__device__ __forceinline__ int2 lds_v2s32(uint32_t a) {
  int2 v;
  asm volatile("ld.shared.v2.s32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a) : "memory");
  return v;
}

__device__ __forceinline__ unsigned lanemask_lt() {
  unsigned m;
  asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
  return m;
}
__device__ __forceinline__ double warp_sum(double v) {  // xor butterfly: every lane ends with the same bits
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) v = __dadd_rn(v, shfl_xor_t(v, off));
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) v = __fadd_rn(v, __shfl_xor_sync(kFull, v, off));
  return v;
}

// index of state s among the lattice states sv[0..NS) (ascending) or the extras sv[NS..NS+nx); -1 when absent
__device__ __forceinline__ int pwl_find(const float* sv, int NS, int nx, float s) {
  int lo = 0, hi = NS;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (sv[mid] < s) lo = mid + 1; else hi = mid;
  }
  if (lo < NS && sv[lo] == s) return lo;
  for (int e = 0; e < nx; ++e)
    if (__float_as_int(sv[NS + e]) == __float_as_int(s)) return NS + e;
  return -1;
}

__device__ __forceinline__ bool pwl_active(float s, float w, float b) { return __fadd_rn(__fmul_rn(s, w), b) > 0.0f; }

// Breakpoint of every hidden unit on the sorted lattice and the units in breakpoint order.
// ev[j] = key | leave << 15:  leave = 0: unit j is active on ranks [key, NS);  leave = 1: active on ranks [0, key).
// ord[0..H): unit indices sorted by key (stable in j, so the result does not depend on the schedule).
__device__ inline void pwl_unit_events(const float* w1, const float* b1, int H, const float* sv, int NS, uint16_t* ev,
                                       uint16_t* ord, uint16_t* bkt, int lane, bool sort = true) {
  const int NB = NS + 2;
  __syncwarp();
  for (int b = lane; b < NB; b += 32) bkt[b] = 0;
  __syncwarp();
  for (int j0 = 0; j0 < H; j0 += 32) {
    const int j = j0 + lane;
    unsigned kk = 0x10000u + (unsigned)lane;  // lanes past H: unique, no bucket
    if (j < H) {
      const float w = w1[j], b = b1[j];
      int key;
      unsigned leave;
      if (w > 0.0f) {  // the predicate is non-decreasing in the rank: first active rank
        int lo = 0, hi = NS;
        while (lo < hi) {
          const int mid = (lo + hi) >> 1;
          if (pwl_active(sv[mid], w, b)) hi = mid; else lo = mid + 1;
        }
        key = lo; leave = 0;
      } else if (w < 0.0f) {  // non-increasing: first inactive rank
        int lo = 0, hi = NS;
        while (lo < hi) {
          const int mid = (lo + hi) >> 1;
          if (pwl_active(sv[mid], w, b)) lo = mid + 1; else hi = mid;
        }
        key = lo; leave = 1;
      } else {  // w == 0 (or NaN: never active): the same for every state
        key = (w == 0.0f && b > 0.0f) ? NS : 0; leave = 1;
      }
      ev[j] = (uint16_t)(key | (leave << 15));
      kk = (unsigned)key;
    }
    if (sort) {
      const unsigned peers = __match_any_sync(kFull, kk);
      if (j < H && (peers & lanemask_lt()) == 0) bkt[kk + 1] += (uint16_t)__popc(peers);
    }
    __syncwarp();
  }
  if (!sort) return;
  {  // inclusive scan: afterwards bkt[k] = first position of bucket k
    const int per = (NB + 31) / 32;
    const int beg = lane * per, end = beg + per < NB ? beg + per : NB;
    int s = 0;
    for (int b = beg; b < end; ++b) s += bkt[b];
    int incl = s;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const int t = __shfl_up_sync(kFull, incl, off);
      if (lane >= off) incl += t;
    }
    int run = incl - s;
    for (int b = beg; b < end; ++b) { run += bkt[b]; bkt[b] = (uint16_t)run; }
  }
  __syncwarp();
  for (int j0 = 0; j0 < H; j0 += 32) {
    const int j = j0 + lane;
    const unsigned kk = j < H ? (unsigned)(ev[j] & 0x7fff) : 0x10000u + (unsigned)lane;
    const unsigned peers = __match_any_sync(kFull, kk);
    int base = 0;
    if (j < H) {
      base = bkt[kk];
      ord[base + __popc(peers & lanemask_lt())] = (uint16_t)j;
    }
    __syncwarp();
    if (j < H && (peers & lanemask_lt()) == 0) bkt[kk] = (uint16_t)(base + __popc(peers));
    __syncwarp();
  }
}

// softmax over the lanes < A of one state's logits -> probabilities (ws) and their running sum (shared CDF row);
// the lane A of an ActorCritic agent carries v(s).
__device__ __forceinline__ void pwl_emit(float z, bool col, bool vcol, int A, int x, float* cdf, float* val, float* pws, int lane) {
  const float mx = warp_max(col ? z : NegInf<float>::v());
  const float ex = col ? det_expf(__fsub_rn(z, mx)) : 0.0f;
  const float sum = warp_sum(ex);
  const float pk = col ? __fdiv_rn(ex, sum) : 0.0f;
  float c = pk;
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    const float t = __shfl_up_sync(kFull, c, off);
    if (lane >= off) c = __fadd_rn(c, t);
  }
  if (col) { cdf[x * A + lane] = c; pws[x * A + lane] = pk; }
  if (vcol) val[x] = z;
}

// The same for four states at a time: the four softmax / running-sum chains are independent, so their shuffle and
// special-function latencies overlap (one chain is ~450 cycles long).  zs: the logits of state x at zs[x * A + lane] (the
// probabilities overwrite them); states x0 .. x0+3 that are >= NX are skipped.  Per state the arithmetic is pwl_emit's.
__device__ __forceinline__ void pwl_emit4(int x0, int NX, bool col, int A, float* cdf, float* zs, int lane) {
  float z[4], mx[4], ex[4], sum[4], pk[4], c[4];
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int x = x0 + u < NX ? x0 + u : NX - 1;
    z[u] = col ? zs[x * A + lane] : 0.0f;
  }
#pragma unroll
  for (int u = 0; u < 4; ++u) mx[u] = warp_max(col ? z[u] : NegInf<float>::v());
#pragma unroll
  for (int u = 0; u < 4; ++u) ex[u] = col ? det_expf(__fsub_rn(z[u], mx[u])) : 0.0f;
#pragma unroll
  for (int u = 0; u < 4; ++u) sum[u] = ex[u];
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
#pragma unroll
    for (int u = 0; u < 4; ++u) sum[u] = __fadd_rn(sum[u], __shfl_xor_sync(kFull, sum[u], off));
  }
#pragma unroll
  for (int u = 0; u < 4; ++u) { pk[u] = col ? __fdiv_rn(ex[u], sum[u]) : 0.0f; c[u] = pk[u]; }
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const float t = __shfl_up_sync(kFull, c[u], off);
      if (lane >= off) c[u] = __fadd_rn(c[u], t);
    }
  }
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int x = x0 + u;
    if (col && x < NX) { cdf[x * A + lane] = c[u]; zs[x * A + lane] = pk[u]; }
  }
}

constexpr int kPwlTile = 8;  // hidden units staged per round of the forward sweep

// pi(.|s) and v(s) for every lattice state (one sweep, see the header) and for the extras (direct evaluation).
// Needs ev / ord of the CURRENT parameters (pwl_unit_events).  blk: the agent's parameters in state_dict order.
// tile: shared scratch of kPwlTile * ((A + 1) * 4 + 12) bytes (the episode's draw buffer, idle between episodes).
__device__ inline void pwl_build_lut(const float* blk, const ThrlAgentSpec& spec, const float* sv, int NS, int nx,
                                     const uint16_t* ev, const uint16_t* ord, float* cdf, float* val, float* pws,
                                     unsigned char* tile, int lane) {
  const int H = spec.hidden, A = spec.actions;
  const bool ac = spec.kind == THRL_AGENT_ACTORCRITIC;
  const float *w1 = blk, *b1 = blk + H, *W = blk + 2 * H, *bp = W + (size_t)A * H, *wv = bp + A;
  const bool col = lane < A, vcol = ac && lane == A, use = col || vcol;
  const float* crow = col ? W + (size_t)lane * H : wv;  // this lane's row of fc_pi.weight, or fc_v.weight
  const double bias = col ? (double)bp[lane] : (vcol ? (double)wv[H] : 0.0);
  const int NC = ac ? A + 1 : A;  // columns: the actions, then the value head
  double S1 = 0.0, S0 = 0.0;
  __syncwarp();
  // units active from rank 0 on (they leave at their breakpoint): lane = unit, coalesced rows of the weights, one
  // butterfly sum per column; lane c keeps column c
  for (int c = 0; c < NC; ++c) {
    const float* row = c < A ? W + (size_t)c * H : wv;
    double p1 = 0.0, p0 = 0.0;
    for (int j0 = lane; j0 < H; j0 += 128) {  // four units per lane and round: the (predicated) weight loads go out together
      float cwf[4];
      bool lv[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int j = j0 + 32 * u;
        lv[u] = j < H && (ev[j] & 0x8000);
        cwf[u] = lv[u] ? row[j] : 0.0f;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (lv[u]) {
          const int j = j0 + 32 * u;
          const double cw = (double)cwf[u];
          p1 = __dadd_rn(p1, __dmul_rn(cw, (double)w1[j]));
          p0 = __dadd_rn(p0, __dmul_rn(cw, (double)b1[j]));
        }
      }
    }
    p1 = warp_sum(p1);
    p0 = warp_sum(p0);
    if (lane == c) { S1 = p1; S0 = p0; }
  }
  // the sweep: kPwlTile units per round are gathered with independent loads (lane = element) one round ahead, parked in
  // shared memory, then applied in breakpoint order (lane = column)
  const int NCP = A + 1;
  float* tW = reinterpret_cast<float*>(tile);                  // [kPwlTile][NCP] weights of the staged units
  float* tw1 = tW + NCP * kPwlTile;                              // [kPwlTile]
  float* tb1 = tw1 + kPwlTile;
  uint16_t* tev = reinterpret_cast<uint16_t*>(tb1 + kPwlTile);   // [kPwlTile]
  constexpr int kRegs = (32 * kPwlTile + 31) / 32;               // elements per lane and round (A + 1 <= 32)
  float rW[kRegs], rw1 = 0.0f, rb1 = 0.0f;
  unsigned rev = 0;
  auto gather = [&](int e0) {
    const int ne = H - e0 < kPwlTile ? H - e0 : kPwlTile;
#pragma unroll
    for (int q = 0; q < kRegs; ++q) {
      const int idx = lane + 32 * q, e = idx / NC, c = idx - e * NC;
      rW[q] = 0.0f;
      if (e < ne) {
        const int j = ord[e0 + e];
        rW[q] = c < A ? W[(size_t)c * H + j] : wv[j];
      }
    }
    if (lane < ne) {
      const int j = ord[e0 + lane];
      rw1 = w1[j]; rb1 = b1[j]; rev = ev[j];
    }
  };
  int r = 0;
  auto emit_row = [&]() {  // the logits of rank r are parked (pws, val); the softmaxes follow the sweep, four states at a time
    const float z = (float)__dadd_rn(__dadd_rn(__dmul_rn(S1, (double)sv[r]), S0), bias);
    if (col) pws[r * A + lane] = z;
    if (vcol) val[r] = z;
    ++r;
  };
  gather(0);
  for (int e0 = 0; e0 < H; e0 += kPwlTile) {
    const int ne = H - e0 < kPwlTile ? H - e0 : kPwlTile;
    __syncwarp();
#pragma unroll
    for (int q = 0; q < kRegs; ++q) {
      const int idx = lane + 32 * q, e = idx / NC, c = idx - e * NC;
      if (e < ne) tW[e * NCP + c] = rW[q];
    }
    if (lane < ne) { tw1[lane] = rw1; tb1[lane] = rb1; tev[lane] = (uint16_t)rev; }
    __syncwarp();
    if (e0 + kPwlTile < H) gather(e0 + kPwlTile);  // in flight while this round is applied
    for (int e = 0; e < ne; ++e) {
      const unsigned v = tev[e];
      const int key = (int)(v & 0x7fff);
      while (r < key && r < NS) emit_row();  // every rank below the breakpoint has all its events
      const double cw = use ? (double)tW[e * NCP + lane] : 0.0;
      const double t1 = __dmul_rn(cw, (double)tw1[e]), t0 = __dmul_rn(cw, (double)tb1[e]);
      if (v & 0x8000) { S1 = __dsub_rn(S1, t1); S0 = __dsub_rn(S0, t0); }
      else { S1 = __dadd_rn(S1, t1); S0 = __dadd_rn(S0, t0); }
    }
  }
  while (r < NS) emit_row();
  for (int x = NS; x < NS + nx; ++x) {  // off-lattice states: the plain sum over the hidden units
    const float s = sv[x];
    double acc = 0.0;
    for (int j = 0; j < H; ++j) {
      const float hv = __fadd_rn(__fmul_rn(s, w1[j]), b1[j]);
      if (hv > 0.0f) acc = __dadd_rn(acc, __dmul_rn((double)hv, use ? (double)crow[j] : 0.0));
    }
    const float z = (float)__dadd_rn(acc, bias);
    if (col) pws[x * A + lane] = z;
    if (vcol) val[x] = z;
  }
  __syncwarp();
  for (int x0 = 0; x0 < NS + nx; x0 += 4) pwl_emit4(x0, NS + nx, col, A, cdf, pws, lane);
  __syncwarp();
}

// 2^k with N * cmax * 2^k < 2^62: the fixed-point scale of one update's coefficient sums
__device__ __forceinline__ double pwl_scale(float cmax, int N) {
  if (!(cmax > 0.0f)) return 1.0;
  int e;
  frexpf(cmax, &e);  // cmax < 2^e
  int k = 62 - e - (32 - __clz(N));
  k = k > 1000 ? 1000 : (k < -1000 ? -1000 : k);
  return ldexp(1.0, k);
}

// clip_grad_norm_(1.0) + one Adam step (oracle mlp_clip_adam) on parameters that stay in the global slab; g: state_dict order
__device__ inline void pwl_clip_adam(float* blk, const ThrlAgentSpec& spec, const float* g, int lane) {
  const int P = mlp_P(spec);
  float *am = blk + P, *av = blk + 2 * (size_t)P;
  int32_t* hdr = reinterpret_cast<int32_t*>(blk + 3 * (size_t)P);
  __syncwarp();
  double part = 0.0;
  for (int i = lane; i < P; i += 32) { const double gd = (double)g[i]; part = __dadd_rn(part, __dmul_rn(gd, gd)); }
  const float total_norm = (float)sqrt(warp_sum(part));
  float coef = __fdiv_rn(1.0f, __fadd_rn(total_norm, 1e-6f));
  if (coef > 1.0f) coef = 1.0f;
  const int step = hdr[0] + 1;
  double pw1 = 1.0, pw2 = 1.0;
  for (int q2 = 0; q2 < step; ++q2) { pw1 = __dmul_rn(pw1, 0.9); pw2 = __dmul_rn(pw2, 0.999); }
  const double bc1 = __dsub_rn(1.0, pw1), bc2 = __dsub_rn(1.0, pw2);
  const float neg_step_size = (float)(-__ddiv_rn(spec.lr, bc1));
  const float bc2_sqrt = (float)sqrt(bc2);
  const float w1m = (float)__dsub_rn(1.0, 0.9), fb2 = (float)0.999, w2 = (float)__dsub_rn(1.0, 0.999), eps = 1e-8f;
  // eight parameters per lane and round, every load issued before the arithmetic (the moments live in HBM: the loop is
  // latency-bound otherwise; pwl_train prefetches the block into L2 while the gradient is being formed)
  constexpr int kW = 8;  // parameters per lane and round
  for (int i0 = lane; i0 < P; i0 += 32 * kW) {
    float gi[kW], m[kW], v[kW], w[kW];
#pragma unroll
    for (int u = 0; u < kW; ++u) {
      const int i = i0 + 32 * u;
      const bool in = i < P;
      gi[u] = in ? g[i] : 0.0f;
      m[u] = in ? am[i] : 0.0f;
      v[u] = in ? av[i] : 0.0f;
      w[u] = in ? blk[i] : 0.0f;
    }
#pragma unroll
    for (int u = 0; u < kW; ++u) {
      const int i = i0 + 32 * u;
      if (i < P) {
        const float gc = __fmul_rn(gi[u], coef);
        const float mm = __fadd_rn(m[u], __fmul_rn(__fsub_rn(gc, m[u]), w1m));
        const float vv = __fadd_rn(__fmul_rn(v[u], fb2), __fmul_rn(__fmul_rn(w2, gc), gc));
        am[i] = mm;
        av[i] = vv;
        // entries whose gradient has been exactly zero so far (units that are inactive on the whole lattice) keep m = v = 0:
        // same result as the general formula, without sending the warp through the slow paths of sqrt and division
        const float den = vv != 0.0f ? __fadd_rn(__fdiv_rn(sqrtf(vv), bc2_sqrt), eps) : eps;
        const float num = __fmul_rn(neg_step_size, mm);
        blk[i] = __fadd_rn(w[u], mm != 0.0f ? __fdiv_rn(num, den) : num);
      }
    }
  }
  __syncwarp();
  if (lane == 0) hdr[0] = step;
  __syncwarp();
}

// Reinforce.train_net (agents.py:170-194) / ActorCritic.train_net (:280-305) on the N buffered transitions, by states.
// The per-sample coefficients are the oracle's (mlp_train / ac_train, float32, same operation order); only the sums over
// samples and hidden units are re-associated.  val: v(s) LUT of the current parameters; pws: their pi(.|s) LUT.
__device__ inline void pwl_train(float* blk, const ThrlAgentSpec& spec, int cap, int head, int N, const float* sv, int NS,
                                 int nx, uint16_t* ev, uint16_t* ord, uint16_t* bkt, const float* val, const float* pws,
                                 long long* acc, double2* pf, float* g, float4* xs, int lane) {
  const int H = spec.hidden, A = spec.actions, P = mlp_P(spec), EW = mlp_entry_words(spec);
  const bool ac = spec.kind == THRL_AGENT_ACTORCRITIC;
  float* buf = blk + 3 * (size_t)P + THRL_MLP_HEADER_WORDS;
  const float *w1 = blk, *b1 = blk + H;
  const int C = A + 2;  // accumulator columns per state: [0, A) state x action, A: value head, A + 1: state total
  const int NX = NS + nx;
  const float gam = (float)spec.gamma;
  auto entry = [&](int nn) {
    int sl = head + nn;
    if (sl >= cap) sl -= cap;
    return buf + (size_t)sl * EW;
  };
  __syncwarp();
  for (int i = lane * 32; i < 3 * P; i += 32 * 32)  // parameters and Adam moments towards L2 (one line per lane and round)
    asm volatile("prefetch.global.L2 [%0];" ::"l"(blk + i));
  for (int i = lane; i < NX * C; i += 32) acc[i] = 0;
  bool bad = false;
  float camax = 0.0f, cvmax = 0.0f;
  if (!ac) {
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
      const int x = pwl_find(sv, NS, nx, en[0]);
      bad |= !isfinite(ca) || x < 0;
      camax = fmaxf(camax, fabsf(ca));
      xs[nn] = make_float4(__int_as_float(x & 0xffff), ca, 0.0f, 0.0f);
    }
  } else {
    double Rp = 0.0, Dp = 0.0;
    for (int nn = lane; nn < N; nn += 32) {  // d_i = gamma * v(s'_i) - v(s_i) (:289)
      const float* en = entry(nn);
      const int x = pwl_find(sv, NS, nx, en[0]), x2 = pwl_find(sv, NS, nx, en[3]);
      bad |= x < 0 || x2 < 0;
      const float v = val[x < 0 ? 0 : x], vp = val[x2 < 0 ? 0 : x2];
      const float d = __fsub_rn(__fmul_rn(gam, vp), v);
      Rp = __dadd_rn(Rp, (double)en[2]);
      Dp = __dadd_rn(Dp, (double)d);
      xs[nn] = make_float4(__int_as_float((x & 0xffff) | ((x2 & 0xffff) << 16)), 0.0f, 0.0f, d);
    }
    const float fN = (float)N, fR = (float)warp_sum(Rp), fD = (float)warp_sum(Dp);
    const float invN2 = __fdiv_rn(1.0f, __fmul_rn(fN, fN));
    for (int nn = lane; nn < N; nn += 32) {  // the [N,N] advantage broadcast collapsed as in oracle ac_train
      float4 q = xs[nn];
      const float r = entry(nn)[2];
      q.y = __fmul_rn(__fadd_rn(__fmul_rn(fN, r), fD), invN2);                       // actor weight (N r_j + D) / N^2
      q.z = __fmul_rn(-2.0f, __fmul_rn(__fadd_rn(fR, __fmul_rn(fN, q.w)), invN2));   // dL/dv_j; dL/dv'_j = -gamma * that
      bad |= !isfinite(q.y) || !isfinite(q.z);
      camax = fmaxf(camax, fabsf(q.y));
      cvmax = fmaxf(cvmax, fmaxf(fabsf(q.z), fabsf(__fmul_rn(-gam, q.z))));
      xs[nn] = q;
    }
  }
  bad = __any_sync(kFull, bad);
  camax = warp_max(camax);
  cvmax = warp_max(cvmax);
  if (bad) {  // non-finite coefficients (e.g. zero return variance): the reference's gradient is NaN everywhere
    for (int i = lane; i < P; i += 32) g[i] = __int_as_float(0x7fc00000);
    pwl_clip_adam(blk, spec, g, lane);
    return;
  }
  const double sa = pwl_scale(camax, N), sc = pwl_scale(cvmax, 2 * N);
  __syncwarp();
  {
    unsigned long long* uacc = reinterpret_cast<unsigned long long*>(acc);
    for (int nn = lane; nn < N; nn += 32) {
      const float4 q = xs[nn];
      const int xb = __float_as_int(q.x), x = xb & 0xffff, x2 = (xb >> 16) & 0xffff;
      const int a = __float_as_int(entry(nn)[1]);
      const unsigned long long fa = (unsigned long long)__double2ll_rn(__dmul_rn((double)q.y, sa));
      atomicAdd(uacc + x * C + a, fa);
      atomicAdd(uacc + x * C + A + 1, fa);
      if (ac) {
        atomicAdd(uacc + x * C + A, (unsigned long long)__double2ll_rn(__dmul_rn((double)q.z, sc)));
        atomicAdd(uacc + x2 * C + A, (unsigned long long)__double2ll_rn(__dmul_rn((double)__fmul_rn(-gam, q.z), sc)));
      }
    }
  }
  __syncwarp();
  pwl_unit_events(w1, b1, H, sv, NS, ev, ord, bkt, lane, /*sort=*/false);
  // ---- lane = column (action k, or the value head): per-state gradient coefficients
  //   DL[x][k] = pi(k|x) * CA[x] - CAa[x][k],  DL[x][A] = CV[x]
  // and their running sums over the sorted lattice, pf[r][c] = (sum_{x<r} DL[x][c], sum_{x<r} DL[x][c] * s_x), r = 0..NS
  // (row NS = totals); rows NS+1.. hold (DL, DL * s) of the extra states.
  const bool col = lane < A, vcol = ac && lane == A, use = col || vcol;
  const int CW = A + 1;
  const double isa = __ddiv_rn(1.0, sa), isc = __ddiv_rn(1.0, sc);
  {
    double P0 = 0.0, P1 = 0.0, tot = 0.0;
    for (int x = 0; x < NX; ++x) {
      double dl = 0.0;
      if (use) {
        const double raw = (double)__ldcg(acc + x * C + lane);
        if (col) {
          const double cax = __dmul_rn((double)__ldcg(acc + x * C + A + 1), isa);
          dl = __dsub_rn(__dmul_rn((double)pws[x * A + lane], cax), __dmul_rn(raw, isa));
        } else {
          dl = __dmul_rn(raw, isc);
        }
        const double ds = __dmul_rn(dl, (double)sv[x]);
        if (x < NS) {
          pf[x * CW + lane] = make_double2(P0, P1);
          P0 = __dadd_rn(P0, dl);
          P1 = __dadd_rn(P1, ds);
        } else {
          pf[(x + 1) * CW + lane] = make_double2(dl, ds);
        }
        tot = __dadd_rn(tot, dl);
      }
    }
    if (use) pf[NS * CW + lane] = make_double2(P0, P1);
    if (col) g[2 * H + A * H + lane] = (float)tot;      // fc_pi.bias
    if (vcol) g[2 * H + A * H + A + H] = (float)tot;    // fc_v.bias
  }
  __syncwarp();
  // ---- lane = hidden unit: unit j is active on the ranks [key, NS) or [0, key), so with M0/M1 = the sums of DL, DL * s
  // over its active states:  d/dW[k][j] = w1_j * M1 + b1_j * M0 (= sum_x DL[x][k] * h_x[j]),
  // d/db1[j] = sum_c W[c][j] * M0[c], d/dw1[j] = sum_c W[c][j] * M1[c]   (c runs over the actions and the value head)
  const int NC = ac ? A + 1 : A;
  for (int j0 = 0; j0 < H; j0 += 32) {
    const int j = j0 + lane;
    if (j < H) {
      const float w = w1[j], b = b1[j];
      const unsigned evj = ev[j];
      const int key = (int)(evj & 0x7fff);
      const bool leave = (evj & 0x8000) != 0;
      unsigned xm = 0;
      for (int e2 = 0; e2 < nx; ++e2) xm |= pwl_active(sv[NS + e2], w, b) ? 1u << e2 : 0u;
      double gw = 0.0, gb = 0.0;
      // four columns per round, every load of the round issued before its arithmetic (the prefix sums sit in L2: the loop is
      // bound by their latency otherwise)
      for (int c0 = 0; c0 < NC; c0 += 4) {
        double2 pre[4], tt[4];
        float cwf[4];
        int wi[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int c = c0 + u < NC ? c0 + u : NC - 1;
          pre[u] = pf[key * CW + c];
          tt[u] = pf[NS * CW + c];
          wi[u] = c < A ? 2 * H + c * H + j : 2 * H + A * H + A + j;  // fc_pi.weight[c][j] / fc_v.weight[j]
          cwf[u] = blk[wi[u]];
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int c = c0 + u;
          if (c < NC) {
            double M0 = leave ? pre[u].x : __dsub_rn(tt[u].x, pre[u].x), M1 = leave ? pre[u].y : __dsub_rn(tt[u].y, pre[u].y);
            for (int e2 = 0; e2 < nx; ++e2) {
              if (xm >> e2 & 1) {
                const double2 dx = pf[(NS + 1 + e2) * CW + c];
                M0 = __dadd_rn(M0, dx.x);
                M1 = __dadd_rn(M1, dx.y);
              }
            }
            g[wi[u]] = (float)__dadd_rn(__dmul_rn((double)w, M1), __dmul_rn((double)b, M0));
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

// QT: storage type of the Q-tables of QTable agents playing in the same game (unused when there are none).
// kN: number of agents when it is 2 (the agent loop of the episode is unrolled), else 0.
//
// QTable agents (agents.py:14-112) in a lattice game: the table is staged in shared memory for the whole call (the game
// must be regular, include/thrl.h), the greedy action is cached per lattice state and agent (the table is frozen within an
// episode; entries of rows the episode-end update wrote are dropped), the update rows are tabulated per joint action (the
// float64 encode depends on the exact price, not on its float32 state), and the episode-end update is the sequential pass
// of the other kernels (stale snapshot, live next_max, visit counters, epsilon decay).
// kQ: the game has QTable agents.  kCdfG: the CDF LUT lives in the workspace (PwlParams.cdf_global).
template <typename QT, int kN, bool kQ, bool kCdfG>
__global__ void __launch_bounds__(32 * kPwlMaxWarps, 1) mlp_scan_pwl(const __grid_constant__ PwlParams p) {
  extern __shared__ __align__(16) unsigned char smem[];
  const ThrlGame& G = p.game;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpc = blockDim.x >> 5;
  const int n = G.n_agents, T = G.max_steps, E = p.E, NS = p.NS, J = p.J, NSX = NS + kPwlExtras;
  const bool is_agent = lane < n;

  // ---- CTA-shared tables: per-action quantities, joint action -> price / lattice state / reward share / update rows
  double* lutAQ = reinterpret_cast<double*>(smem);
  double* lutXT = lutAQ + p.lut_total;
  double* priceJ = reinterpret_cast<double*>(smem + p.off_priceJ);
  double* rT = reinterpret_cast<double*>(smem + p.off_rT);  // [J][n] reward / max_steps (trainer.py:63)
  float* rF = reinterpret_cast<float*>(smem + p.off_rF);    // [J][n] reward as the float32 the buffers hold (agents.py:142)
  uint16_t* slot_of = reinterpret_cast<uint16_t*>(smem + p.off_slotof);
  uint16_t* urowJ = reinterpret_cast<uint16_t*>(smem + p.off_urowJ);  // [QTable agent][J] float64 encode of the price (agents.py:62,66)
  {
    const double ab = __ddiv_rn(G.a, G.b);
    for (int i = 0; i < n; ++i) {
      const ThrlAgentSpec& s = G.agent[i];
      for (int k = threadIdx.x; k < s.actions; k += blockDim.x) {  // QTable.scale: k/(A-1) (agents.py:51-57); Reinforce.scale: k/A (:154-158)
        const double x = s.kind == THRL_AGENT_QTABLE
                             ? scale_action(k, s.actions, s.action_lo, s.action_hi)
                             : __dadd_rn(__dmul_rn(__ddiv_rn((double)k, (double)s.actions), __dsub_rn(s.action_hi, s.action_lo)), s.action_lo);
        lutAQ[p.a_off[i] + k] = __dmul_rn(ab, x);
        lutXT[p.a_off[i] + k] = __ddiv_rn(x, (double)T);
      }
    }
    for (int j = threadIdx.x; j < J; j += blockDim.x) { priceJ[j] = p.priceJ[j]; slot_of[j] = p.slot_of[j]; }
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < J * n; idx += blockDim.x) {
    const int j = idx / n, i = idx - j * n;
    const int k = (j / p.jmul[i]) % G.agent[i].actions;
    const double rew = __dmul_rn(priceJ[j], lutAQ[p.a_off[i] + k]);
    rT[idx] = __ddiv_rn(rew, (double)T);
    rF[idx] = (float)rew;
  }
  for (int i = 0; i < n; ++i) {
    const int qi = p.qidx[i];
    if (qi < 0) continue;
    const ThrlAgentSpec& s = G.agent[i];
    for (int j = threadIdx.x; j < J; j += blockDim.x) urowJ[qi * J + j] = (uint16_t)upd_row(priceJ[j], s.max_state, (double)s.states);
  }
  __syncthreads();

  unsigned char* slot = smem + p.cta_bytes + (size_t)warp * p.warp_bytes;
  float* sv = reinterpret_cast<float*>(slot + p.off_sv);      // [NS + extras] state values
  // per MLP agent [NS + extras][A] running sums of pi(.|s): shared memory, or the workspace when the lattice is large
  float* cdfb = kCdfG ? reinterpret_cast<float*>(p.ws + ((size_t)blockIdx.x * wpc + warp) * p.ws_warp_bytes + p.ws_cdf)
                      : reinterpret_cast<float*>(slot + p.off_cdf);
  float* valb = reinterpret_cast<float*>(slot + p.off_val);   // per agent [NS + extras] v(s)
  int32_t* pre = reinterpret_cast<int32_t*>(slot + p.off_pre);  // [T][n] draws of the episode, see below
  uint16_t* ev = reinterpret_cast<uint16_t*>(slot + p.off_ev);
  uint16_t* ord = reinterpret_cast<uint16_t*>(slot + p.off_ord);
  uint16_t* bkt = reinterpret_cast<uint16_t*>(slot + p.off_bkt);
  // QTable agents only:
  double* hpw = reinterpret_cast<double*>(slot + p.off_hpw);          // [n][5] alpha, gamma, eps_end, eps_step, epsilon
  uint16_t* jrec = reinterpret_cast<uint16_t*>(slot + p.off_jrec);    // [T] joint action of every step of the episode
  QT* oldv = reinterpret_cast<QT*>(slot + p.off_oldv);                // [T] stale snapshot (agents.py:67)
  uint8_t* gq = slot + p.off_gq;                                      // [nq][NS + extras] greedy action of the state, 0xFF unknown
  uint16_t* arow = reinterpret_cast<uint16_t*>(slot + p.off_arow);    // [nq][NS + extras] float32 encode of the state (acting row)
  uint32_t* dirty = reinterpret_cast<uint32_t*>(slot + p.off_dirty);  // [dwords] rows written by the running update
  unsigned char* wsw = p.ws + ((size_t)blockIdx.x * wpc + warp) * p.ws_warp_bytes;
  long long* acc = reinterpret_cast<long long*>(wsw + p.ws_acc);
  float* pws = reinterpret_cast<float*>(wsw + p.ws_p);
  double2* pfw = reinterpret_cast<double2*>(wsw + p.ws_pf);
  float* gws = reinterpret_cast<float*>(wsw + p.ws_grad);
  float4* xs = reinterpret_cast<float4*>(wsw + p.ws_xs);
  for (int x = lane; x < NS; x += 32) sv[x] = p.slot_val[x];

  int my_cap = 0, my_aoff = 0, my_EW = 3, my_P = 0, my_kind = THRL_AGENT_REINFORCE;
  long long my_off = 0;
  if (is_agent) {
    const ThrlAgentSpec& s = G.agent[lane];
    my_kind = s.kind;
    my_cap = s.kind == THRL_AGENT_QTABLE ? 0 : G.mlp_buffer_len[lane];
    my_aoff = p.a_off[lane];
    my_EW = mlp_entry_words(s);
    my_P = mlp_P(s);
    my_off = s.mlp_offset;
  }

  const bool tracing = p.trace_actions || p.trace_rewards || p.trace_prices;
  // the two agents of a 2-agent game: action counts and CDF tables in registers
  const int A0 = G.agent[0].actions, A1 = G.agent[kN == 2 ? 1 : 0].actions;
  const float *cdf0 = cdfb + p.cdf_off[0], *cdf1 = cdfb + p.cdf_off[kN == 2 ? 1 : 0];
  constexpr int kGreedy = 0x7fffffff;  // draw of a QTable agent that acts greedily (not a uniform < 1: those are < 0x3f800000)

  const long long total_warps = (long long)gridDim.x * wpc;
  for (long long r = p.run_lo + (long long)blockIdx.x * wpc + warp; r < p.n_runs; r += total_warps) {
    float* slab = p.mlp + r * G.mlp_stride;
    QT* tabg = reinterpret_cast<QT*>(p.q) + r * G.run_stride;
    uint32_t* cnt = p.counter ? p.counter + r * G.run_stride : nullptr;
    const uint32_t gid = (uint32_t)(p.run_id0 + r);
    const double price = p.price[r];
    int jlast = -1;  // joint action of the latest step: the run's price is priceJ[jlast]
    int nx = 0;
    bool overflow = false;
    auto find_or_insert = [&](float s) {  // warp-uniform s
      int x = pwl_find(sv, NS, nx, s);
      if (x < 0) {
        if (nx < kPwlExtras) {
          __syncwarp();
          if (lane == 0) sv[NS + nx] = s;
          __syncwarp();
          x = NS + nx;
          ++nx;
        } else {
          overflow = true;
          x = NS + kPwlExtras - 1;
        }
      }
      return x;
    };
    __syncwarp();
    int x = find_or_insert((float)price);
    // transitions pending from the previous call may hold off-lattice states too
    int my_len = 0, my_head = 0, my_wr = 0;  // deque(maxlen = capacity): length, oldest slot (at entry), next slot to write
    float* my_buf = nullptr;
    int32_t* my_hdr = nullptr;
    if (is_agent && my_kind != THRL_AGENT_QTABLE) {
      float* blk = slab + my_off;
      my_hdr = reinterpret_cast<int32_t*>(blk + 3 * (size_t)my_P);
      my_buf = blk + 3 * (size_t)my_P + THRL_MLP_HEADER_WORDS;
      if (my_cap > 0) {
        my_len = my_hdr[1]; my_head = my_hdr[2];
        my_wr = my_head + my_len;
        if (my_wr >= my_cap) my_wr -= my_cap;
      }
    }
    for (int i = 0; i < n; ++i) {
      const ThrlAgentSpec& s = G.agent[i];
      if (s.kind == THRL_AGENT_QTABLE) continue;
      const int L = __shfl_sync(kFull, my_len, i), hd = __shfl_sync(kFull, my_head, i), cap = G.mlp_buffer_len[i];
      const int EW = mlp_entry_words(s);
      const float* buf = slab + s.mlp_offset + 3 * (size_t)mlp_P(s) + THRL_MLP_HEADER_WORDS;
      for (int b0 = 0; b0 < L; b0 += 32) {
        const int nn = b0 + lane;
        float s0 = 0.0f, s1 = 0.0f;
        bool m0 = false, m1 = false;
        if (nn < L) {
          int sl = hd + nn;
          if (sl >= cap) sl -= cap;
          s0 = buf[(size_t)sl * EW];
          m0 = pwl_find(sv, NS, nx, s0) < 0;
          if (EW == 4) { s1 = buf[(size_t)sl * EW + 3]; m1 = pwl_find(sv, NS, nx, s1) < 0; }
        }
        unsigned mm = __ballot_sync(kFull, m0);
        while (mm) { const int l = __ffs(mm) - 1; mm &= mm - 1; (void)find_or_insert(__shfl_sync(kFull, s0, l)); }
        mm = __ballot_sync(kFull, m1);
        while (mm) { const int l = __ffs(mm) - 1; mm &= mm - 1; (void)find_or_insert(__shfl_sync(kFull, s1, l)); }
      }
    }
    for (int i = 0; i < n; ++i) {
      const ThrlAgentSpec& s = G.agent[i];
      if (s.kind == THRL_AGENT_QTABLE) {  // stage the table; acting rows of all states; nothing known about greedy actions
        const int qi = p.qidx[i], cells = (s.states + 1) * s.row_stride;  // staged with its row padding, if any
        QT* tb = reinterpret_cast<QT*>(slot + p.off_tab[i]);
        const QT* src = tabg + s.table_offset;
        for (int c = lane; c < cells; c += 32) tb[c] = src[c];
        for (int xx = lane; xx < NS + nx; xx += 32) {
          arow[qi * NSX + xx] = (uint16_t)act_row((double)sv[xx], (float)s.max_state, (float)s.states);
          gq[qi * NSX + xx] = 0xFF;
        }
        if (lane == 0) {
          double* h = hpw + i * 5;
          if (p.hp) {
            const double* hs = p.hp + (r * n + i) * 4;
            h[0] = hs[0]; h[1] = hs[1]; h[2] = hs[2]; h[3] = hs[3];
          } else {
            h[0] = s.alpha; h[1] = s.gamma; h[2] = s.eps_end; h[3] = s.eps_step;
          }
          h[4] = p.eps[r * n + i];
        }
        continue;
      }
      const float* blk = slab + s.mlp_offset;
      pwl_unit_events(blk, blk + s.hidden, s.hidden, sv, NS, ev, ord, bkt, lane);
      pwl_build_lut(blk, s, sv, NS, nx, ev, ord, cdfb + p.cdf_off[i], valb + p.val_off[i], pws + p.cdf_off[i], reinterpret_cast<unsigned char*>(pre), lane);
    }
    for (int w = lane; w < p.dwords; w += 32) dirty[w] = 0;
    int urow_cur = 0;  // lane i (a QTable agent): update row of the state the next episode starts from
    if (is_agent && my_kind == THRL_AGENT_QTABLE) urow_cur = upd_row(price, G.agent[lane].max_state, (double)G.agent[lane].states);
    __syncwarp();

    for (int e = 0; e < E; ++e) {
      const uint32_t eabs = (uint32_t)(p.epoch_begin + e);
      const long long step0 = (r * E + e) * (long long)T;
      // ---- per-episode draws.  MLP agent: -1 - action when the action is forced (replay modes), else the bits of the
      // uniform u = 24 random bits * 2^-24 that Categorical.sample() is emulated with.  QTable agent (agents.py:80-89):
      // -1 - action for an exploring (or forced) step, kGreedy otherwise.
      for (int idx = lane; idx < T * n; idx += 32) {
        const int t = idx / n, i = idx - t * n;
        int v;
        if (G.agent[i].kind == THRL_AGENT_QTABLE) {
          if (p.rng_mode == THRL_RNG_REPLAY_ACTIONS) {
            v = -1 - p.replay_ra[step0 * n + idx];
          } else if (p.rng_mode == THRL_RNG_REPLAY_DRAWS) {
            v = p.replay_u[step0 * n + idx] < hpw[i * 5 + 4] ? -1 - p.replay_ra[step0 * n + idx] : kGreedy;
          } else {
            uint32_t xr[4];
            philox4x32_10(gid, eabs, (uint32_t)t, (uint32_t)(i >> 1) | (kStreamAct << 16), p.k0, p.k1, xr);
            const int ra = (int)__umulhi(xr[2 * (i & 1) + 1], (uint32_t)G.agent[i].actions);
            v = u32_unit(xr[2 * (i & 1)]) < hpw[i * 5 + 4] ? -1 - ra : kGreedy;
          }
        } else {
          v = p.rng_mode == THRL_RNG_PHILOX ? -1 : p.replay_ra[step0 * n + idx];
          if (v < 0) {
            uint32_t xr[4];
            philox4x32_10(gid, eabs, (uint32_t)t, (uint32_t)(i >> 1) | (kStreamAct << 16), p.k0, p.k1, xr);
            v = __float_as_int(__fmul_rn((float)(xr[2 * (i & 1)] >> 8), 1.0f / 16777216.0f));
          } else {
            v = -1 - v;
          }
        }
        pre[idx] = v;
      }
      __syncwarp();

      // ---- the episode (trainer.py:50-67): pi(.|s) is a row of the CDF LUT, the environment a table of the joint action
      double rlog = 0.0, alog = 0.0;
      auto pick = [&](int v, int i) {
        if (v < 0) return -1 - v;
        const ThrlAgentSpec& s = G.agent[i];
        const int Ai = kN == 2 ? (i == 0 ? A0 : A1) : s.actions;
        if (kQ && s.kind == THRL_AGENT_QTABLE) {  // first argmax of the (frozen) table row of the state (agents.py:84-88)
          const int qi = p.qidx[i];
          int g = gq[qi * NSX + x];
          if (g == 0xFF) {
            g = row_argmax(reinterpret_cast<const QT*>(slot + p.off_tab[i]) + (size_t)arow[qi * NSX + x] * G.agent[i].row_stride, Ai, lane);
            __syncwarp();
            if (lane == 0) gq[qi * NSX + x] = (uint8_t)g;
            __syncwarp();
          }
          return g;
        }
        // first k with cumsum(pi)[k] > u (agents.py:160-163), the last action if there is none
        const float* cdf = kN == 2 ? (i == 0 ? cdf0 : cdf1) : cdfb + p.cdf_off[i];
        const float c = lane < Ai ? cdf[x * Ai + lane] : 0.0f;
        const unsigned m = __ballot_sync(kFull, lane < Ai && c > __int_as_float(v));
        return m ? __ffs(m) - 1 : Ai - 1;
      };
      if (kN == 2 && !kQ && !kCdfG && !tracing) {
        // two MLP agents, LUT in shared memory: branch-free steps on 32-bit shared addresses.  A forced action overrides
        // the sampled one; bit A-1 is or-ed into the ballot so that "no k with cdf > u" yields the last action.
        const int c0l = lane < A0 ? lane : A0 - 1, c1l = lane < A1 ? lane : A1 - 1;
        const uint32_t pre_s = smem_u32(pre), cdf0_s = smem_u32(cdf0) + 4 * c0l, cdf1_s = smem_u32(cdf1) + 4 * c1l;
        const uint32_t slot_s = smem_u32(slot_of), sv_s = smem_u32(sv);
        const uint32_t rT_s = smem_u32(rT) + 8 * (lane & 1), rF_s = smem_u32(rF) + 4 * (lane & 1), xt_s = smem_u32(lutXT + my_aoff);
        const unsigned last0 = 1u << (A0 - 1), last1 = 1u << (A1 - 1);
        const bool in0 = lane < A0, in1 = lane < A1, first = lane == 0, append = is_agent && my_cap > 0;
        for (int t = 0; t < T; ++t) {
          const int2 v = lds_v2s32(pre_s + 8 * t);
          const float c0 = lds_f32(cdf0_s + 4 * A0 * x), c1 = lds_f32(cdf1_s + 4 * A1 * x);
          const unsigned m0 = __ballot_sync(kFull, in0 && c0 > __int_as_float(v.x)) | last0;
          const unsigned m1 = __ballot_sync(kFull, in1 && c1 > __int_as_float(v.y)) | last1;
          const int k0 = v.x < 0 ? -1 - v.x : __ffs(m0) - 1, k1 = v.y < 0 ? -1 - v.y : __ffs(m1) - 1;
          const int joint = k0 * A1 + k1, kmine = first ? k0 : k1;
          const int xn = lds_u16(slot_s + 2 * joint);
          if (is_agent) {
            rlog = __dadd_rn(rlog, lds_f64(rT_s + 16 * joint));
            alog = __dadd_rn(alog, lds_f64(xt_s + 8 * kmine));
          }
          if (append) {  // memory.append; replay(cast) makes state and reward float32 (buffers.py:28-38, agents.py:142)
            float* en = my_buf + (size_t)my_wr * my_EW;
            en[0] = lds_f32(sv_s + 4 * x);
            en[1] = __int_as_float(kmine);
            en[2] = lds_f32(rF_s + 8 * joint);
            if (my_EW == 4) en[3] = lds_f32(sv_s + 4 * xn);
            my_wr = my_wr + 1 == my_cap ? 0 : my_wr + 1;
            my_len = my_len < my_cap ? my_len + 1 : my_cap;
          }
          x = xn;
          jlast = joint;
        }
      } else
      for (int t = 0; t < T; ++t) {
        int joint = 0, kmine = 0;
        if (kN == 2) {
          const int2 v = *reinterpret_cast<const int2*>(pre + 2 * t);
          const int k0 = pick(v.x, 0), k1 = pick(v.y, 1);
          joint = k0 * A1 + k1;
          kmine = lane == 0 ? k0 : k1;
        } else {
          for (int i = 0; i < n; ++i) {
            const int k = pick(pre[t * n + i], i);
            joint += k * p.jmul[i];
            if (lane == i) kmine = k;
          }
        }
        const int xn = slot_of[joint];
        if (kQ && lane == 0) jrec[t] = (uint16_t)joint;
        if (is_agent) {
          rlog = __dadd_rn(rlog, rT[joint * n + lane]);
          alog = __dadd_rn(alog, lutXT[my_aoff + kmine]);
          if (my_cap > 0) {  // memory.append; replay(cast) makes state and reward float32 (buffers.py:28-38, agents.py:142)
            float* en = my_buf + (size_t)my_wr * my_EW;
            en[0] = sv[x];
            en[1] = __int_as_float(kmine);
            en[2] = rF[joint * n + lane];
            if (my_EW == 4) en[3] = sv[xn];
            my_wr = my_wr + 1 == my_cap ? 0 : my_wr + 1;
            my_len = my_len < my_cap ? my_len + 1 : my_cap;
          }
        }
        if (tracing) {
          const double next_price = priceJ[joint];
          if (is_agent) {
            if (p.trace_actions) p.trace_actions[(step0 + t) * n + lane] = kmine;
            if (p.trace_rewards) p.trace_rewards[(step0 + t) * n + lane] = __dmul_rn(next_price, lutAQ[my_aoff + kmine]);
          }
          if (lane == 0 && p.trace_prices) p.trace_prices[step0 + t] = next_price;
        }
        x = xn;
        jlast = joint;
      }
      __syncwarp();

      // ---- train_net for every agent in order (trainer.py:70)
      for (int i = 0; i < n; ++i) {
        const ThrlAgentSpec& s = G.agent[i];
        if (s.kind == THRL_AGENT_QTABLE) {  // QTable.train_net (agents.py:59-78) on the newest L transitions of the episode
          const int qi = p.qidx[i], L = p.L[i], A = s.actions, jm = p.jmul[i];
          const int ucur = __shfl_sync(kFull, urow_cur, i);
          if (L > 0) {
            QT* tb = reinterpret_cast<QT*>(slot + p.off_tab[i]);
            const uint16_t* ur = urowJ + qi * J;
            const int t0 = T - L;
            const double alpha = hpw[i * 5 + 0], gamma = hpw[i * 5 + 1], one_m_alpha = __dsub_rn(1.0, alpha);
            for (int t = t0 + lane; t < T; t += 32) {  // stale snapshot (:67)
              const int st = t == 0 ? ucur : ur[jrec[t - 1]];
              oldv[t] = tb[(size_t)st * s.row_stride + (jrec[t] / jm) % A];
            }
            __syncwarp();
            for (int t = t0; t < T; ++t) {  // the sequential pass (:68-76)
              const int jt = jrec[t];
              const int st = t == 0 ? ucur : ur[jrec[t - 1]], ns = ur[jt], k = (jt / jm) % A;
              const double reward = __dmul_rn(priceJ[jt], lutAQ[p.a_off[i] + k]);
              const double next_max = (double)row_max(tb + (size_t)ns * s.row_stride, A, lane);
              const double nv = __dadd_rn(__dmul_rn(one_m_alpha, (double)oldv[t]),
                                          __dmul_rn(alpha, __dadd_rn(reward, __dmul_rn(gamma, next_max))));
              if ((k & 31) == lane) {  // the lane that owns column k
                tb[(size_t)st * s.row_stride + k] = (QT)nv;
                if (cnt) atomicAdd(cnt + s.table_offset + (size_t)st * s.row_stride + k, 1u);
                dirty[st >> 5] |= 1u << (st & 31);
              }
            }
            __syncwarp();
            for (int xx = lane; xx < NS + nx; xx += 32) {  // greedy actions of rewritten rows are unknown again
              const int rw = arow[qi * NSX + xx];
              if (dirty[rw >> 5] >> (rw & 31) & 1) gq[qi * NSX + xx] = 0xFF;
            }
            __syncwarp();
            for (int w = lane; w < p.dwords; w += 32) dirty[w] = 0;
          }
          if (lane == i) urow_cur = urowJ[qi * J + jlast];
          if (lane == 0) {  // epsilon decay (:78), every episode
            double* h = hpw + i * 5;
            h[4] = __dadd_rn(h[2], __dmul_rn(__dsub_rn(h[4], h[2]), h[3]));
          }
          __syncwarp();
          continue;
        }
        const int cap = G.mlp_buffer_len[i];
        if (cap == 0) continue;
        const int L = __shfl_sync(kFull, my_len, i);
        if (L < s.min_memory) continue;
        int hd = __shfl_sync(kFull, my_wr, i) - L;  // oldest buffered transition
        if (hd < 0) hd += cap;
        float* blk = slab + s.mlp_offset;
        pwl_train(blk, s, cap, hd, L, sv, NS, nx, ev, ord, bkt, valb + p.val_off[i], pws + p.cdf_off[i], acc, pfw, gws, xs, lane);
        if (lane == i) { my_len = 0; my_wr = 0; }  // :194 memory.empty()
        pwl_unit_events(blk, blk + s.hidden, s.hidden, sv, NS, ev, ord, bkt, lane);
        pwl_build_lut(blk, s, sv, NS, nx, ev, ord, cdfb + p.cdf_off[i], valb + p.val_off[i], pws + p.cdf_off[i], reinterpret_cast<unsigned char*>(pre), lane);
      }
      if (is_agent) {
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

    // ---- write the run back
    for (int i = 0; i < n; ++i) {
      const ThrlAgentSpec& s = G.agent[i];
      if (s.kind != THRL_AGENT_QTABLE) continue;
      const int cells = (s.states + 1) * s.row_stride;
      const QT* tb = reinterpret_cast<const QT*>(slot + p.off_tab[i]);
      QT* dst = tabg + s.table_offset;
      for (int c = lane; c < cells; c += 32) dst[c] = tb[c];
      if (lane == 0) p.eps[r * n + i] = hpw[i * 5 + 4];
    }
    if (is_agent && my_cap > 0) {
      int hd = my_wr - my_len;
      if (hd < 0) hd += my_cap;
      my_hdr[1] = my_len;
      my_hdr[2] = my_len ? hd : 0;
    }
    if (overflow && lane == 0) {  // more off-lattice states than kPwlExtras: fail loudly (NaN weights)
      for (int i = 0; i < n; ++i)
        if (G.agent[i].kind != THRL_AGENT_QTABLE) { slab[G.agent[i].mlp_offset] = __int_as_float(0x7fc00000); break; }
    }
    if (lane == 0 && jlast >= 0) p.price[r] = priceJ[jlast];
    __syncwarp();
  }
}

}  // namespace thrl
