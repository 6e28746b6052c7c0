// thrl_scan_hbm.cuh — Q-table games whose tables stay in HBM (BASELINE C4: 8 agents x 1001 x 101 fp32 = 3.2 MB per run).
//
// One warp plays one run, persistent grid.  The rollout (trainer.py:50-67) acts from an exact greedy-action cache in
// shared memory, so it touches HBM only on the first visit of a table row.  QTable.train_net (agents.py:59-78: stale
// snapshot, live row max, writes in batch order) is NOT executed as a load -> max -> store chain through memory.  Per
// episode and agent the kernel
//   1. snapshots old[j] = Q[s_j, k_j] of every transition (scattered 4-byte loads, all independent) and counts the visit,
//   2. TAGS every cell the batch will write: an atomic max leaves a NaN-boxed marker in the cell that carries the smallest
//      transition index writing it -- the table itself now says which cells are being rewritten and by whom,
//   3. gathers the rows of all states of the episode with 1-D bulk copies (cp.async.bulk, one 16-byte-aligned padded row
//      = one copy, completion on an mbarrier) into a ring of staging slots: every load is independent of every other, a
//      warp keeps `nb` x n_agents rows in flight,
//   4. walks the states in order entirely on chip: the staged row with its tagged cells replaced by their current values
//      cur[canonical transition] IS the live row of agents.py:71, so next_max is one warp reduction; the new value goes
//      to cur[] (shared memory),
//   5. stores cur[] of the canonical transitions over the tags, and refreshes the greedy-action cache of every touched
//      row exactly: (max, first argmax) over the row's untagged cells -- which no write of this batch can change -- merged
//      with the final values of its tagged cells.
// scripts/model_hbm_update.py replays steps 1-5 on random batches against the plain sequential form.
// HBM traffic per agent-step: the bootstrap row (row_stride * 4 B, once), one sector for the cell and one for the
// counter; the greedy row read of SURVEY 8(d)'s algorithmic count is served by the cache.
//
// Applies to: Q-table agents only, regular games (every agent's batch is the newest min(T, capacity) transitions of the
// episode), max_steps <= 254, <= 128 actions, padded slab layout (include/thrl.h ThrlAgentSpec.row_stride).  Everything
// else runs on thrl_scan_generic.cuh, which is also the checker this kernel is tested against (THRL_KERNEL=generic).
#pragma once
#include "thrl_device.cuh"

namespace thrl {

constexpr int kHbmMaxT = 254;   // transition / state indices fit one byte next to 0xFF = none
constexpr int kHbmMaxNb = 8;    // staging ring depth (batches of n rows)

struct HbmParams {
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
  // per agent: rows with a greedy-cache slot, offsets into the cache / the action LUT, batch length (0 = never updates)
  int gcap[THRL_MAX_AGENTS], goff[THRL_MAX_AGENTS], loff[THRL_MAX_AGENTS], L[THRL_MAX_AGENTS];
  int lut_total, rows_total;
  int noisy;     // new_a varies per step
  int bulk;      // how rows are staged: 1 = cp.async.bulk (TMA unit), 2 = cp.async 16-byte copies (LSU, no registers),
                 // 0 = 16-byte vector loads + shared stores (comparison)
  int nb;        // staging ring depth
  int slot_bytes;  // one staged row
  int Tp, Sp;    // padded strides of the [n][T] / [n][T+1] per-agent arrays (elements)
  // shared memory (bytes): [cta_bytes][warp 0][warp 1]...
  int cta_bytes, warp_bytes;
  int off_bar, off_g, off_P, off_act, off_hp, off_srow, off_cur, off_canon, off_bm, off_ba, off_stage;
  // regions that alias the staging ring: the rollout's draws (before the update), the greedy merge (after the walk)
  int off_pre, off_newa, off_rs, off_vkey, off_cmin;
};

// ---------------------------------------------------------------- PTX: mbarrier + 1-D bulk copy (TMA unit, no descriptor)
__device__ __forceinline__ uint32_t hbm_smem_addr(const void* ptr) { return (uint32_t)__cvta_generic_to_shared(ptr); }
__device__ __forceinline__ void hbm_mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(hbm_smem_addr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void hbm_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(hbm_smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void hbm_mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(hbm_smem_addr(bar)) : "memory");
}
__device__ __forceinline__ void hbm_mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(hbm_smem_addr(bar)), "r"(parity)
        : "memory");
  } while (!done);
}
// per-thread 16-byte asynchronous copy global -> shared that bypasses L1 (it must see the tags), and the mbarrier arrival
// that fires once all of this thread's earlier copies have landed
__device__ __forceinline__ void hbm_cp_async16(void* dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(hbm_smem_addr(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void hbm_cp_async_arrive(uint64_t* bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(hbm_smem_addr(bar)) : "memory");
}
__device__ __forceinline__ void hbm_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(hbm_smem_addr(dst)),
               "l"(src), "r"(bytes), "r"(hbm_smem_addr(bar))
               : "memory");
}
// generic-proxy accesses (the tags in global memory, this warp's use of the staging ring in shared memory) are ordered
// before the async-proxy accesses of the bulk copies issued after the fence
__device__ __forceinline__ void hbm_fence_proxy_async() {
  asm volatile("fence.proxy.async.global;\n\tfence.proxy.async.shared::cta;" ::: "memory");
}

// ---------------------------------------------------------------- tags: NaN-boxed transition indices
// A tagged cell holds kBase | (1023 - j): a negative quiet NaN, above every number (and -inf) as an unsigned integer, so
// one unsigned atomic max both installs the tag over the value and keeps the SMALLEST j among the transitions that write
// the cell.  The update arithmetic never produces such a bit pattern (its NaNs, should inputs be NaN, are positive).
template <typename QT> struct HbmBits;
template <> struct HbmBits<float> {
  using U = uint32_t;
  static constexpr U kBase = 0xFFC00000u;
  __device__ static U of(float v) { return __float_as_uint(v); }
  __device__ static U key(float v) {  // order-preserving, -0 folded onto +0 (numpy compares them equal)
    const U b = __float_as_uint(v == 0.0f ? 0.0f : v);
    return (b >> 31) ? ~b : (b | 0x80000000u);
  }
};
template <> struct HbmBits<double> {
  using U = unsigned long long;
  static constexpr U kBase = 0xFFF8000000000000ull;
  __device__ static U of(double v) { return (U)__double_as_longlong(v); }
  __device__ static U key(double v) { return dkey(v == 0.0 ? 0.0 : v); }
};

// this lane's four columns 4*lane .. 4*lane+3 of a 16-byte aligned row (shared or global memory)
__device__ __forceinline__ void hbm_load4(const float* row, int lane, float (&v)[4]) {
  const float4 x = *reinterpret_cast<const float4*>(row + 4 * lane);
  v[0] = x.x; v[1] = x.y; v[2] = x.z; v[3] = x.w;
}
__device__ __forceinline__ void hbm_load4(const double* row, int lane, double (&v)[4]) {
  const double2 x = *reinterpret_cast<const double2*>(row + 4 * lane), y = *reinterpret_cast<const double2*>(row + 4 * lane + 2);
  v[0] = x.x; v[1] = x.y; v[2] = y.x; v[3] = y.y;
}

template <typename QT>
__global__ void __launch_bounds__(512, 1) qtable_scan_hbm(const __grid_constant__ HbmParams p) {
  using B = HbmBits<QT>;
  using U = typename B::U;
  extern __shared__ __align__(128) unsigned char smem_hbm[];
  unsigned char* const smem = smem_hbm;
  const ThrlGame& G = p.game;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, warps_per_cta = blockDim.x >> 5;
  const int n = G.n_agents, T = G.max_steps, E = p.E, Tp = p.Tp, Sp = p.Sp, nb = p.nb;
  const bool is_agent = lane < n;
  constexpr unsigned kNone = 0xffffffffu;

  // ---- CTA-shared: AQ[k] = (a/b)*scale(k), XT[k] = scale(k)/max_steps per agent, and the per-agent constants of the update
  double* lutAQ = reinterpret_cast<double*>(smem);
  double* lutXT = lutAQ + p.lut_total;
  int* agc = reinterpret_cast<int*>(lutXT + p.lut_total);  // [n][8]: table offset, row stride, actions, lut offset, cache offset, cache rows, t0, L
  {
    const double ab = __ddiv_rn(G.a, G.b);  // environments.py:23 self.a/self.b
    for (int i = 0; i < n; ++i) {
      const ThrlAgentSpec& s = G.agent[i];
      for (int k = threadIdx.x; k < s.actions; k += blockDim.x) {
        const double x = scale_action(k, s.actions, s.action_lo, s.action_hi);
        lutAQ[p.loff[i] + k] = __dmul_rn(ab, x);
        lutXT[p.loff[i] + k] = __ddiv_rn(x, (double)T);  // trainer.py:66 scaled_acts / max_steps
      }
      if (threadIdx.x == 0) {
        int* a = agc + i * 8;
        a[0] = (int)s.table_offset; a[1] = s.row_stride; a[2] = s.actions; a[3] = p.loff[i];
        a[4] = p.goff[i]; a[5] = p.gcap[i]; a[6] = T - p.L[i]; a[7] = p.L[i];
      }
    }
  }
  __syncthreads();

  // ---- this warp's slot
  unsigned char* slot = smem + p.cta_bytes + (size_t)warp * p.warp_bytes;
  uint64_t* bar = reinterpret_cast<uint64_t*>(slot + p.off_bar);          // [nb] one mbarrier per staging batch
  uint8_t* Gc = slot + p.off_g;                                            // greedy action per (agent,row); 0xFF = unknown
  double* P = reinterpret_cast<double*>(slot + p.off_P);                   // [T+1] prices of the episode, P[0] = state it starts from
  uint8_t* act = slot + p.off_act;                                         // [n][Tp] chosen actions
  double* hpw = reinterpret_cast<double*>(slot + p.off_hp);                // [n][5] alpha,gamma,eps_end,eps_step,eps
  uint16_t* srow = reinterpret_cast<uint16_t*>(slot + p.off_srow);         // [n][Sp] float64 encode of every state (agents.py:62,66)
  QT* cur = reinterpret_cast<QT*>(slot + p.off_cur);                       // [n][Tp] snapshot, then live values of the canonical cells
  uint8_t* canon = slot + p.off_canon;                                     // [n][Tp] first transition of the batch that writes the same cell
  QT* bm = reinterpret_cast<QT*>(slot + p.off_bm);                         // [n][Sp] max over the untagged cells of the state's row
  uint8_t* ba = slot + p.off_ba;                                           // [n][Sp] its first column (0xFF: every column is tagged)
  unsigned char* stage = slot + p.off_stage;                               // [nb][n][slot_bytes] staged rows
  int16_t* pre = reinterpret_cast<int16_t*>(slot + p.off_pre);             // [T][n] forced action or -1 (= greedy)       } alias
  double* newa = reinterpret_cast<double*>(slot + p.off_newa);             // [T] demand intercept (noisy only)            } the
  uint8_t* rs = slot + p.off_rs;                                           // [n][Sp] state that represents the row        } staging
  U* vkey = reinterpret_cast<U*>(slot + p.off_vkey);                       // [n][Sp] ordered key of the row's final max   } ring
  uint32_t* cmin = reinterpret_cast<uint32_t*>(slot + p.off_cmin);         // [n][Sp] its first column                     }

  if (lane == 0)
    for (int b = 0; b < nb; ++b) hbm_mbar_init(bar + b, p.bulk == 1 ? 1u : 32u);  // bulk: one arrival + byte count; else every lane arrives
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  hbm_fence_proxy_async();
  __syncwarp();
  uint32_t parity = 0;  // bit b: phase the next wait on bar[b] completes

  // ---- lane i < n keeps agent i's constants in registers
  int my_A = 2, my_lut = 0, my_goff = 0, my_gcap = 0;
  float my_msf = 1.f, my_sf = 1.f;
  if (is_agent) {
    const ThrlAgentSpec& s = G.agent[lane];
    my_A = s.actions;
    my_msf = (float)s.max_state;
    my_sf = (float)s.states;
    my_lut = p.loff[lane]; my_goff = p.goff[lane]; my_gcap = p.gcap[lane];
  }
  int tmin = T;  // first state any agent's batch needs
  for (int i = 0; i < n; ++i)
    if (p.L[i] > 0 && T - p.L[i] < tmin) tmin = T - p.L[i];

  const long long total_warps = (long long)gridDim.x * warps_per_cta;
  for (long long r = (long long)blockIdx.x * warps_per_cta + warp; r < p.n_runs; r += total_warps) {
    QT* qg = reinterpret_cast<QT*>(p.q) + r * G.run_stride;
    uint32_t* cnt = p.counter ? p.counter + r * G.run_stride : nullptr;
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
    for (int c = lane; c < p.rows_total; c += 32) Gc[c] = 0xFF;
    double price = p.price[r];
    __syncwarp();

    for (int e = 0; e < E; ++e) {
      const uint32_t eabs = (uint32_t)(p.epoch_begin + e);
      const long long step0 = (r * E + e) * (long long)T;

      // ---- per-episode draws, lane-parallel (epsilon is frozen within an episode: trainer.py:50-70 only calls train_net
      //      after the episode).  pre[t][i] = action forced by exploration / replay, or -1 = greedy.
      if (p.rng_mode == THRL_RNG_PHILOX) {
        const int npair = (n + 1) >> 1;  // one Philox call serves agents 2p and 2p+1 (DESIGN.md "Philox streams")
        for (int idx = lane; idx < T * npair; idx += 32) {
          const int t = idx / npair, pr = idx - t * npair;
          uint32_t x[4];
          philox4x32_10(gid, eabs, (uint32_t)t, (uint32_t)pr | (kStreamAct << 16), p.k0, p.k1, x);
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int i = 2 * pr + h;
            if (i < n) {
              const double u = u32_unit(x[2 * h]);
              const int ra = (int)__umulhi(x[2 * h + 1], (uint32_t)agc[i * 8 + 2]);
              pre[t * n + i] = (int16_t)(u < hpw[i * 5 + 4] ? ra : -1);  // agents.py:81-82
            }
          }
        }
      } else {
        for (int idx = lane; idx < T * n; idx += 32) {
          const int i = idx % n;
          int v;
          if (p.rng_mode == THRL_RNG_REPLAY_ACTIONS) {
            v = p.replay_ra[step0 * n + idx];
          } else {
            const double u = p.replay_u[step0 * n + idx];
            v = u < hpw[i * 5 + 4] ? p.replay_ra[step0 * n + idx] : -1;  // agents.py:81
          }
          pre[idx] = (int16_t)v;
        }
      }
      if (p.noisy) {
        for (int t = lane; t < T; t += 32) {
          double na = G.a;
          if (p.rng_mode == THRL_RNG_PHILOX) {
            uint32_t x[4];
            philox4x32_10(gid, eabs, (uint32_t)t, kStreamEnv << 16, p.k0, p.k1, x);
            if (u53(x[0], x[1]) < G.noise_prob) {  // environments.py:28
              const double lo = __dmul_rn(G.a, 0.7);
              na = __dadd_rn(lo, __dmul_rn(__dsub_rn(G.a, lo), u53(x[2], x[3])));
            }
          } else if (p.replay_new_a) {
            na = p.replay_new_a[step0 + t];
          }
          newa[t] = na;
        }
      }
      if (lane == 0) P[0] = price;
      __syncwarp();

      // ---- the episode (trainer.py:50-67); lane i < n acts for agent i
      double rlog = 0.0, alog = 0.0;  // trainer.py:40-41
      for (int t = 0; t < T; ++t) {
        int k = 0, arow = 0;
        if (is_agent) {
          k = pre[t * n + lane];
          if (k < 0) {  // agents.py:84-88 on the frozen table
            arow = act_row(price, my_msf, my_sf);
            const int g = arow < my_gcap ? (int)Gc[my_goff + arow] : 0xFF;  // rows beyond the cache: always recomputed
            k = g == 0xFF ? -1 : g;
          }
        }
        unsigned need = __ballot_sync(kFull, is_agent && k < 0);  // first visit of a row: its greedy action is not known yet
        while (need) {
          const int i = __ffs(need) - 1;
          need &= need - 1;
          const int ri = __shfl_sync(kFull, arow, i);
          const int* a = agc + i * 8;
          const QT* row = qg + a[0] + (size_t)ri * a[1];
          QT v[4];
          if (4 * lane < a[1]) hbm_load4(row, lane, v);
          QT best = NegInf<QT>::v();
          unsigned bidx = kNone;
#pragma unroll
          for (int c = 0; c < 4; ++c)  // ascending columns, strict >: first maximal index (agents.py:88)
            if (4 * lane + c < a[2] && (bidx == kNone || v[c] > best)) { best = v[c]; bidx = 4 * lane + c; }
          const QT wm = warp_max(best);
          const int g = (int)__reduce_min_sync(kFull, (bidx != kNone && best == wm) ? bidx : kNone);
          if (lane == 0 && ri < a[5]) Gc[a[4] + ri] = (uint8_t)g;
          if (lane == i) k = g;
        }
        double aq = 0.0;
        if (is_agent) aq = lutAQ[my_lut + k];
        double Q = 0.0;  // environments.py:27 sum(A): ((0 + A0) + A1) + ...
        for (int i = 0; i < n; ++i) Q = __dadd_rn(Q, shfl_d(aq, i));
        const double na = p.noisy ? newa[t] : G.a;
        const double pn = __dsub_rn(na, __dmul_rn(G.b, Q));
        const double next_price = pn > 0.0 ? pn : (pn != pn ? pn : 0.0);  // numpy.max([0, x])
        const double rew = __dmul_rn(next_price, aq);                       // environments.py:34
        if (is_agent) {
          rlog = __dadd_rn(rlog, __ddiv_rn(rew, (double)T));  // trainer.py:65
          alog = __dadd_rn(alog, lutXT[my_lut + k]);          // trainer.py:66
          act[lane * Tp + t] = (uint8_t)k;                    // trainer.py:61-62 memory.append
          if (p.trace_actions) p.trace_actions[(step0 + t) * n + lane] = k;
          if (p.trace_rewards) p.trace_rewards[(step0 + t) * n + lane] = rew;
        }
        if (lane == 0) {
          P[t + 1] = next_price;
          if (p.trace_prices) p.trace_prices[step0 + t] = next_price;
        }
        price = next_price;  // trainer.py:67
      }
      __syncwarp();

      // ================================================================ train_net for every agent (trainer.py:70)
      // ---- 1. encodes (agents.py:62,66), stale snapshot (:67) and visit counters (:76), lane-parallel
      for (int i = 0; i < n; ++i) {
        const int* a = agc + i * 8;
        if (a[7] == 0) continue;
        const double ms = G.agent[i].max_state, st = (double)G.agent[i].states;
        for (int t = a[6] + lane; t <= T; t += 32) srow[i * Sp + t] = (uint16_t)upd_row(P[t], ms, st);
      }
      __syncwarp();
      for (int i = 0; i < n; ++i) {
        const int* a = agc + i * 8;
        if (a[7] == 0) continue;
        for (int t = a[6] + lane; t < T; t += 32) {
          const int cell = a[0] + (int)srow[i * Sp + t] * a[1] + act[i * Tp + t];  // a run's slab has < 2^31 elements (checked at layout)
          cur[i * Tp + t] = qg[cell];
          if (cnt) atomicAdd(cnt + cell, 1u);  // fire-and-forget RED
        }
      }
      __syncwarp();  // every snapshot value has arrived (it was stored to cur[]) before any cell is tagged
      // ---- 2. tag the cells of the batch
      for (int i = 0; i < n; ++i) {
        const int* a = agc + i * 8;
        if (a[7] == 0) continue;
        for (int t = a[6] + lane; t < T; t += 32) {
          const int cell = a[0] + (int)srow[i * Sp + t] * a[1] + act[i * Tp + t];
          atomicMax(reinterpret_cast<U*>(qg + cell), (U)(B::kBase | (U)(1023 - t)));
        }
      }
      __threadfence();          // the tags are performed at L2 ...
      hbm_fence_proxy_async();  // ... and ordered before the bulk copies (async proxy) that read them; also orders this
      __syncwarp();             // warp's generic accesses to the staging ring (pre / newa / merge) before its reuse

      // ---- 3 + 4. gather the rows of states tmin..T through the staging ring and walk them in order
      auto issue = [&](int b, int t) {
        uint32_t bytes = 0;
        const QT* src = nullptr;
        if (is_agent) {
          const int* a = agc + lane * 8;
          if (a[7] > 0 && t >= a[6]) {
            bytes = (uint32_t)a[1] * (uint32_t)sizeof(QT);
            src = qg + a[0] + (size_t)srow[lane * Sp + t] * a[1];
          }
        }
        unsigned char* dst = stage + ((size_t)b * n + lane) * p.slot_bytes;
        if (p.bulk == 1) {
          const uint32_t total = __reduce_add_sync(kFull, bytes);
          if (lane == 0) hbm_mbar_expect_tx(bar + b, total);
          if (bytes) hbm_bulk_g2s(dst, src, bytes, bar + b);
        } else {  // the same rows in 16-byte pieces: lane l moves bytes [16 l, 16 l + 16) (and [16 (l + 32), ...) of 8-byte tables)
          for (int i = 0; i < n; ++i) {
            const uint32_t bi = __shfl_sync(kFull, bytes, i);
            const unsigned long long si = __shfl_sync(kFull, (unsigned long long)src, i);
            int4* di = reinterpret_cast<int4*>(stage + ((size_t)b * n + i) * p.slot_bytes);
            if (p.bulk == 2) {
              if (16u * lane < bi) hbm_cp_async16(di + lane, reinterpret_cast<const int4*>(si) + lane);
              if (sizeof(QT) == 8 && 16u * (lane + 32) < bi) hbm_cp_async16(di + lane + 32, reinterpret_cast<const int4*>(si) + lane + 32);
            } else {  // loads that bypass L1 (they must see the tags)
              if (16u * lane < bi) di[lane] = __ldcg(reinterpret_cast<const int4*>(si) + lane);
              if (sizeof(QT) == 8 && 16u * (lane + 32) < bi) di[lane + 32] = __ldcg(reinterpret_cast<const int4*>(si) + lane + 32);
            }
          }
          if (p.bulk == 2) hbm_cp_async_arrive(bar + b);
          else hbm_mbar_arrive(bar + b);
        }
      };
      if (tmin < T) {
        for (int b = 0; b < nb && tmin + b <= T; ++b) issue(b, tmin + b);
        for (int t = tmin; t <= T; ++t) {
          const int b = (t - tmin) % nb;
          hbm_mbar_wait(bar + b, (parity >> b) & 1u);
          parity ^= 1u << b;
          for (int i = 0; i < n; ++i) {
            const int* a = agc + i * 8;
            if (a[7] == 0 || t < a[6]) continue;
            const int RS = a[1], A = a[2];
            const QT* srow_s = reinterpret_cast<const QT*>(stage + ((size_t)b * n + i) * p.slot_bytes);
            QT v[4];
            if (4 * lane < RS) hbm_load4(srow_s, lane, v);
            U bits[4];
            bool tg[4];
            QT bmv = NegInf<QT>::v();
            unsigned bidx = kNone;
            bool anytag = false;
#pragma unroll
            for (int c = 0; c < 4; ++c) {  // (max, first argmax) over this lane's untagged columns
              const bool valid = 4 * lane + c < A;
              bits[c] = valid ? B::of(v[c]) : (U)0;
              tg[c] = bits[c] >= B::kBase;
              anytag |= tg[c];
              if (valid && !tg[c] && (bidx == kNone || v[c] > bmv)) { bmv = v[c]; bidx = 4 * lane + c; }
            }
            const QT wbm = warp_max(bmv);
            const unsigned wba = __reduce_min_sync(kFull, (bidx != kNone && bmv == wbm) ? bidx : kNone);
            QT wlm = wbm;  // live row max (agents.py:71): tagged cells read through to their current values
            if (__any_sync(kFull, anytag)) {
              QT lm = bmv;
#pragma unroll
              for (int c = 0; c < 4; ++c)
                if (tg[c]) {
                  int ci = 1023 - (int)(bits[c] & (U)1023);
                  ci = ci < T ? ci : T - 1;
                  const QT pv = cur[i * Tp + ci];
                  lm = pv > lm ? pv : lm;
                }
              wlm = warp_max(lm);
            }
            if (t < T) {  // canonical transition of the cell transition t writes: read from the tag it left in this row
              const int kt = act[i * Tp + t];
              const int kc = kt & 3;
              const U bsel = kc == 0 ? bits[0] : (kc == 1 ? bits[1] : (kc == 2 ? bits[2] : bits[3]));
              const uint32_t mine = (uint32_t)(bsel & (U)1023);
              const int ct = 1023 - (int)__shfl_sync(kFull, mine, kt >> 2);
              if (lane == 0) canon[i * Tp + t] = (uint8_t)ct;
            }
            if (t > a[6]) {  // transition j = t - 1 bootstraps from this row (:71-75)
              const int j = t - 1, kj = act[i * Tp + j];
              const double alpha = hpw[i * 5 + 0], gamma = hpw[i * 5 + 1];
              const double reward = __dmul_rn(P[t], lutAQ[a[3] + kj]);
              const double nv = __dadd_rn(__dmul_rn(__dsub_rn(1.0, alpha), (double)cur[i * Tp + j]),
                                          __dmul_rn(alpha, __dadd_rn(reward, __dmul_rn(gamma, (double)wlm))));
              if (lane == 0) cur[i * Tp + canon[i * Tp + j]] = (QT)nv;  // cur[j] itself is still the snapshot: the canonical slot is the first writer's
            }
            if (lane == 0) {
              bm[i * Sp + t] = wbm;
              ba[i * Sp + t] = wba == kNone ? (uint8_t)0xFF : (uint8_t)wba;
            }
          }
          __syncwarp();  // cur[] / canon[] of this state are visible; every lane is done with staging batch b
          if (t + nb <= T) issue(b, t + nb);
        }
      }

      // ---- 5. write back over the tags (:75; the last write of the batch per cell), exact greedy refresh of every row touched
      for (int i = 0; i < n; ++i) {
        const int* a = agc + i * 8;
        if (a[7] == 0) continue;
        for (int t = a[6] + lane; t < T; t += 32)
          if (canon[i * Tp + t] == t) qg[a[0] + (int)srow[i * Sp + t] * a[1] + act[i * Tp + t]] = cur[i * Tp + t];
        for (int t = a[6] + lane; t <= T; t += 32) {  // A: one representative state per distinct row (any winner)
          const int row = srow[i * Sp + t];
          if (row < a[5]) Gc[a[4] + row] = (uint8_t)t;
        }
      }
      __syncwarp();
      for (int i = 0; i < n; ++i) {
        const int* a = agc + i * 8;
        if (a[7] == 0) continue;
        for (int t = a[6] + lane; t <= T; t += 32) {  // B: the representative starts from the untagged part of its row
          const int row = srow[i * Sp + t];
          uint8_t rep = 0xFF;
          if (row < a[5]) {
            rep = Gc[a[4] + row];
            if (rep == t) { vkey[i * Sp + t] = B::key(bm[i * Sp + t]); cmin[i * Sp + t] = kNone; }
          }
          rs[i * Sp + t] = rep;
        }
      }
      __syncwarp();
      for (int i = 0; i < n; ++i) {
        const int* a = agc + i * 8;
        if (a[7] == 0) continue;
        for (int t = a[6] + lane; t < T; t += 32) {  // C1: final values of the rewritten cells
          const int rep = rs[i * Sp + t];
          if (rep != 0xFF && canon[i * Tp + t] == t) atomicMax(&vkey[i * Sp + rep], B::key(cur[i * Tp + t]));
        }
      }
      __syncwarp();
      for (int i = 0; i < n; ++i) {
        const int* a = agc + i * 8;
        if (a[7] == 0) continue;
        for (int t = a[6] + lane; t <= T; t += 32) {  // C2: first column that holds the maximum
          const int rep = rs[i * Sp + t];
          if (rep == 0xFF) continue;
          if (rep == t && ba[i * Sp + t] != 0xFF && B::key(bm[i * Sp + t]) == vkey[i * Sp + t]) atomicMin(&cmin[i * Sp + t], (uint32_t)ba[i * Sp + t]);
          if (t < T && canon[i * Tp + t] == t && B::key(cur[i * Tp + t]) == vkey[i * Sp + rep])
            atomicMin(&cmin[i * Sp + rep], (uint32_t)act[i * Tp + t]);
        }
      }
      __syncwarp();
      for (int i = 0; i < n; ++i) {
        const int* a = agc + i * 8;
        if (a[7] == 0) continue;
        for (int t = a[6] + lane; t <= T; t += 32)  // D
          if (rs[i * Sp + t] == t) Gc[a[4] + srow[i * Sp + t]] = (uint8_t)cmin[i * Sp + t];  // (kNone -> 0xFF: cannot happen, every row has a column)
      }
      __syncwarp();

      // epsilon decay, every epoch (:78); logs
      if (is_agent) {
        double* h = hpw + lane * 5;
        h[4] = __dadd_rn(h[2], __dmul_rn(__dsub_rn(h[4], h[2]), h[3]));
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

    // ---- write the run back (tables are already in place)
    if (is_agent) p.eps[r * n + lane] = hpw[lane * 5 + 4];
    if (lane == 0) p.price[r] = price;
    __syncwarp();
  }
}

}  // namespace thrl
