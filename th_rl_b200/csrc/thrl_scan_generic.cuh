// thrl_scan_generic.cuh — the general warp-per-run scan kernel.
//
// One warp plays one run from epoch_begin to epoch_end, then picks up the next run (persistent grid).  It covers
// every configuration the reference's QTable path accepts: 1..16 heterogeneous agents, demand noise, any
// min_memory / capacity / max_steps regime, all three RNG modes, fp32 or f64 tables, tables staged in shared
// memory (kSmemTables) or left in HBM when they do not fit.  The 2-agent noise-free configuration of the headline
// benchmark has a specialised kernel (thrl_scan_lut2.cuh); this one is its general fallback and the C4 (HBM) path.
//
// Restated reference lines: see include/thrl.h; the per-step order is the one in SURVEY.md Appendix A.
#pragma once
#include <type_traits>

#include "thrl_device.cuh"

namespace thrl {

struct ScanParams {
  ThrlGame game;
  long long n_runs, run_id0;
  int epoch_begin, E, rng_mode;
  uint32_t k0, k1;
  void* q;
  uint32_t* counter;
  double* eps;
  double* price;
  const double* hp;
  unsigned char* ring;
  long long ring_bytes;
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
  // shared-memory layout (bytes): [cta_bytes shared by the CTA][warp 0 slot][warp 1 slot]...
  int cta_bytes, warp_bytes;
  int off_tab, off_g, off_pre, off_newa, off_P, off_act, off_row, off_old, off_hp, off_des;
  int row_stride, old_stride;  // per-agent strides (elements) of the rowbuf / oldv scratch
  int lut_total;   // sum of actions_i
  int rows_total;  // sum of gcap_i
  int gcap[THRL_MAX_AGENTS];  // rows 0..gcap_i-1 of agent i have a greedy-cache slot (rows the price can reach); others are uncached
  int Hp;          // ring slots = ring_len + 1
  int noisy;       // new_a varies per step (noise_prob > 0 or replay_new_a given)
  int quarter;     // use the quarter-warp update path (n <= 8 and every agent has <= 128 actions)
  int qfull;       // every agent has >= 8*(NC-1) actions, NC = the dispatched column count: only the last column needs a bounds test
  int qchunks;     // ceil(max actions / 8): columns per lane of a quarter warp (dispatched to 4 / 8 / 13 / 16 at compile time)
};

// Per-run ring blob carried between calls when the game is not regular (include/thrl.h ThrlScanArgs.ring).
struct RingHeader {
  int32_t pos;
  int32_t pad;
  int32_t len[THRL_MAX_AGENTS];
  int32_t pad2[2];
};
static_assert(sizeof(RingHeader) == 80, "ring header");

constexpr int kPrefetchAhead = 4;  // update pass, HBM-resident tables: rows of transition j + kPrefetchAhead are pulled into L2

template <typename QT, bool kSmemTables>
__global__ void __launch_bounds__(kSmemTables ? 1024 : 512, 1) qtable_scan_generic(const __grid_constant__ ScanParams p) {
  extern __shared__ __align__(16) unsigned char smem[];
  const ThrlGame& G = p.game;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, warps_per_cta = blockDim.x >> 5;
  const int n = G.n_agents, T = G.max_steps, E = p.E, Hp = p.Hp;
  const bool is_agent = lane < n;

  // ---- CTA-shared per-action tables: AQ[k] = (a/b)*scale(k), XT[k] = scale(k)/max_steps
  double* lutAQ = reinterpret_cast<double*>(smem);
  double* lutXT = lutAQ + p.lut_total;
  {
    const double ab = __ddiv_rn(G.a, G.b);  // environments.py:23 self.a/self.b
    int base = 0;
    for (int i = 0; i < n; ++i) {
      const ThrlAgentSpec& s = G.agent[i];
      for (int k = threadIdx.x; k < s.actions; k += blockDim.x) {
        const double x = scale_action(k, s.actions, s.action_lo, s.action_hi);
        lutAQ[base + k] = __dmul_rn(ab, x);
        lutXT[base + k] = __ddiv_rn(x, (double)T);  // trainer.py:66 scaled_acts / max_steps
      }
      base += s.actions;
    }
  }
  __syncthreads();

  // ---- this warp's slot
  unsigned char* slot = smem + p.cta_bytes + (size_t)warp * p.warp_bytes;
  QT* tab_s = reinterpret_cast<QT*>(slot + p.off_tab);
  uint8_t* Gc = slot + p.off_g;                                    // greedy action per (agent,row)
  int16_t* pre = reinterpret_cast<int16_t*>(slot + p.off_pre);     // [T][n] forced action or -1 (= greedy)
  double* newa = reinterpret_cast<double*>(slot + p.off_newa);     // [T] demand intercept (noisy only)
  double* P = reinterpret_cast<double*>(slot + p.off_P);           // [Hp] price ring
  uint8_t* act = slot + p.off_act;                                 // [n][Hp] action ring
  uint16_t* rowbuf_all = reinterpret_cast<uint16_t*>(slot + p.off_row);  // [n][row_stride] update-encode rows of the batch
  QT* oldv_all = reinterpret_cast<QT*>(slot + p.off_old);                // [n][old_stride] snapshot (agents.py:67)
  int* des = reinterpret_cast<int*>(slot + p.off_des);                   // [n][4] per-epoch batch descriptor: L, first slot, fires, j offset
  double* hpw = reinterpret_cast<double*>(slot + p.off_hp);        // [n][5] alpha,gamma,eps_end,eps_step,eps

  // ---- lane i < n keeps agent i's constants in registers
  int my_states = 1, my_actions = 2, my_cap = 0, my_minmem = 0x7fffffff, my_lut = 0, my_goff = 0, my_gcap = 0;
  long long my_toff = 0;
  float my_msf = 1.f, my_sf = 1.f;
  if (is_agent) {
    const ThrlAgentSpec& s = G.agent[lane];
    my_states = s.states; my_actions = s.actions; my_cap = s.capacity; my_minmem = s.min_memory;
    my_toff = s.table_offset;
    my_msf = (float)s.max_state;
    my_sf = (float)s.states;
    my_gcap = p.gcap[lane];
    for (int j = 0; j < lane; ++j) { my_lut += G.agent[j].actions; my_goff += p.gcap[j]; }
  }
  const bool never_fires = my_minmem > my_cap;  // buffer can never reach min_memory

  const long long total_warps = (long long)gridDim.x * warps_per_cta;
  for (long long r = (long long)blockIdx.x * warps_per_cta + warp; r < p.n_runs; r += total_warps) {
    QT* qg = reinterpret_cast<QT*>(p.q) + r * G.run_stride;
    QT* tab = kSmemTables ? tab_s : qg;
    uint32_t* cnt = p.counter ? p.counter + r * G.run_stride : nullptr;
    const uint32_t gid = (uint32_t)(p.run_id0 + r);

    // ---- stage the run: tables, hyper-parameters, greedy cache, ring
    if (kSmemTables) {
      for (long long c = lane; c < G.run_stride; c += 32) tab_s[c] = qg[c];
    }
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
    __syncwarp();
    // Quarter-warp mapping for the update (n <= 8, every agent <= 128 actions): lanes 8q..8q+7 serve agents q and 4+q; the
    // scalar work of four agents' update steps then runs SIMD across the quarters instead of once per agent on all 32 lanes.
    const int ql = lane & 7, qq = lane >> 3;
    int qa_A[2], qa_RS[2], qa_toff[2], qa_loff[2], qa_goff[2], qa_gcap[2];
    double qa_alpha[2], qa_gamma[2], qa_oma[2];
    bool qa_ok[2];
#pragma unroll
    for (int g2 = 0; g2 < 2; ++g2) {
      const int i = 4 * g2 + qq;
      qa_ok[g2] = i < n;
      const int ii = qa_ok[g2] ? i : 0;
      const ThrlAgentSpec& s = G.agent[ii];
      qa_A[g2] = s.actions;
      qa_RS[g2] = s.row_stride;
      qa_toff[g2] = (int)s.table_offset;
      qa_loff[g2] = 0; qa_goff[g2] = 0;
      for (int j2 = 0; j2 < ii; ++j2) { qa_loff[g2] += G.agent[j2].actions; qa_goff[g2] += p.gcap[j2]; }
      qa_gcap[g2] = p.gcap[ii];
      qa_alpha[g2] = hpw[ii * 5 + 0];
      qa_gamma[g2] = hpw[ii * 5 + 1];
      qa_oma[g2] = __dsub_rn(1.0, qa_alpha[g2]);
    }
    // greedy-action cache: 0xFF = not computed; filled on first visit, invalidated when the row is written
    for (int c = lane; c < p.rows_total; c += 32) Gc[c] = 0xFF;
    double price = p.price[r];
    int pos = 0, my_len = 0;
    if (p.ring && !G.regular) {
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

      // ---- per-episode draws, lane-parallel (epsilon is frozen within an episode: trainer.py:50-70 only calls
      //      train_net after the episode).  pre[t][i] = action forced by exploration / replay, or -1 = greedy.
      for (int idx = lane; idx < T * n; idx += 32) {
        const int t = idx / n, i = idx - t * n;
        int v;
        if (p.rng_mode == THRL_RNG_REPLAY_ACTIONS) {
          v = p.replay_ra[step0 * n + idx];
        } else if (p.rng_mode == THRL_RNG_REPLAY_DRAWS) {
          const double u = p.replay_u[step0 * n + idx];
          v = u < hpw[i * 5 + 4] ? p.replay_ra[step0 * n + idx] : -1;  // agents.py:81
        } else {
          uint32_t x[4];
          philox4x32_10(gid, eabs, (uint32_t)t, (uint32_t)(i >> 1) | (kStreamAct << 16), p.k0, p.k1, x);
          const double u = u32_unit(x[2 * (i & 1)]);
          const int ra = (int)__umulhi(x[2 * (i & 1) + 1], (uint32_t)G.agent[i].actions);
          v = u < hpw[i * 5 + 4] ? ra : -1;
        }
        pre[idx] = (int16_t)v;
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
      __syncwarp();

      // ---- the episode (trainer.py:50-67); lane i < n acts for agent i
      double rlog = 0.0, alog = 0.0;  // trainer.py:40-41
      for (int t = 0; t < T; ++t) {
        int k = 0;
        double aq = 0.0;
        int arow = 0;
        if (is_agent) {
          k = pre[t * n + lane];
          if (k < 0) {  // agents.py:84-88 on the frozen table
            arow = act_row(price, my_msf, my_sf);
            const int g = arow < my_gcap ? (int)Gc[my_goff + arow] : 0xFF;  // rows beyond the cache: always recomputed
            k = g == 0xFF ? -1 : g;
          }
        }
        unsigned need = __ballot_sync(kFull, is_agent && k < 0);  // greedy action of that row not cached yet
        if (need && p.quarter) {
          // quarter-warp fill: quarter q computes the first argmax of agent 4*g2+q's row (8 lanes x qchunks columns); the loads
          // of both passes are issued before any reduction
          QT bv[2];
          int bi[2], rr[2];
          auto fill_rows = [&](auto ncc, auto fullc) {
            constexpr int NC = decltype(ncc)::value;
            constexpr bool kFullRows = decltype(fullc)::value;  // every agent has >= 8*(NC-1) actions: only the last column can be missing
            QT v[2][NC];
            int left[2];  // columns of the row at or beyond this lane's first one (<= 0: the lane has none)
#pragma unroll
            for (int g2 = 0; g2 < 2; ++g2) {
              const int i = 4 * g2 + qq;
              const bool mine = qa_ok[g2] && ((need >> i) & 1u);
              rr[g2] = __shfl_sync(kFull, arow, qa_ok[g2] ? i : 0);
              left[g2] = (mine ? qa_A[g2] : 0) - ql;
              const QT* row = tab + (qa_toff[g2] + (mine ? rr[g2] : 0) * qa_RS[g2]);
#pragma unroll
              for (int c = 0; c < NC; ++c)  // (rows of quarters that are not served are read too -- row 0, valid memory -- and ignored)
                v[g2][c] = ((kFullRows && c < NC - 1) || 8 * c < left[g2]) ? row[ql + 8 * c] : NegInf<QT>::v();
            }
#pragma unroll
            for (int g2 = 0; g2 < 2; ++g2) {
              // ascending columns, strict >: lowest index of this lane's maximum.  A lane without a first column has no
              // column at all, and a missing column reads as -inf, which is never > anything: no validity tests needed.
              // pairwise tree (depth log2 NC instead of a chain of NC dependent selects); the left operand has the lower
              // columns and wins ties, so the result is the one of the left-to-right scan
              int ci[NC];
#pragma unroll
              for (int c = 0; c < NC; ++c) ci[c] = c;
#pragma unroll
              for (int st = 1; st < NC; st *= 2) {
#pragma unroll
                for (int c = 0; c + st < NC; c += 2 * st)
                  if (v[g2][c + st] > v[g2][c]) { v[g2][c] = v[g2][c + st]; ci[c] = ci[c + st]; }
              }
              bv[g2] = v[g2][0];
              bi[g2] = left[g2] > 0 ? ql + 8 * ci[0] : 0x7fffffff;
            }
          };
          if (p.qfull) {
            if (p.qchunks <= 4) fill_rows(std::integral_constant<int, 4>{}, std::true_type{});
            else if (p.qchunks <= 8) fill_rows(std::integral_constant<int, 8>{}, std::true_type{});
            else if (p.qchunks <= 13) fill_rows(std::integral_constant<int, 13>{}, std::true_type{});  // C4: 101 actions = 13 columns per lane
            else fill_rows(std::integral_constant<int, 16>{}, std::true_type{});
          } else {
            if (p.qchunks <= 4) fill_rows(std::integral_constant<int, 4>{}, std::false_type{});
            else if (p.qchunks <= 8) fill_rows(std::integral_constant<int, 8>{}, std::false_type{});
            else if (p.qchunks <= 13) fill_rows(std::integral_constant<int, 13>{}, std::false_type{});
            else fill_rows(std::integral_constant<int, 16>{}, std::false_type{});
          }
#pragma unroll
          for (int g2 = 0; g2 < 2; ++g2) {
            if (!((need >> (4 * g2)) & 0xFu)) continue;
            QT v = bv[g2];
            int idx = bi[g2];
#pragma unroll
            for (int off = 4; off >= 1; off >>= 1) {  // numpy.argmax: first maximal index (agents.py:88)
              const QT ov = shfl_xor_t(v, off);
              const int oi = __shfl_xor_sync(kFull, idx, off);
              if (oi != 0x7fffffff && (idx == 0x7fffffff || ov > v || (ov == v && oi < idx))) { v = ov; idx = oi; }
            }
            const int i = 4 * g2 + qq;
            const bool mine = qa_ok[g2] && ((need >> i) & 1u);
            if (mine && ql == 0 && rr[g2] < qa_gcap[g2]) Gc[qa_goff[g2] + rr[g2]] = (uint8_t)idx;
            const int got = __shfl_sync(kFull, idx, 8 * (lane & 3));  // agent lane l is served by quarter l & 3 in pass l >> 2
            if (is_agent && (lane >> 2) == g2 && k < 0) k = got;
          }
          __syncwarp();
          need = 0;
        }
        if (need) {
          do {  // up to four agents per round: all their row loads are issued before any comparison
            int ia[4], ra[4], bidx[4];
            QT bval[4];
            bool wide = false;
#pragma unroll
            for (int a = 0; a < 4; ++a) {
              ia[a] = need ? __ffs(need) - 1 : -1;
              if (need) need &= need - 1;
              ra[a] = ia[a] >= 0 ? __shfl_sync(kFull, arow, ia[a]) : 0;
              if (ia[a] >= 0 && G.agent[ia[a]].actions > 128) wide = true;
            }
            if (!wide) {
              QT v[4][4];
#pragma unroll
              for (int a = 0; a < 4; ++a) {
                const ThrlAgentSpec& s = G.agent[ia[a] >= 0 ? ia[a] : 0];
                const QT* row = tab + ((int)s.table_offset + ra[a] * s.row_stride);
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                  const int kk = lane + 32 * c;
                  v[a][c] = (ia[a] >= 0 && kk < s.actions) ? row[kk] : NegInf<QT>::v();
                }
              }
#pragma unroll
              for (int a = 0; a < 4; ++a) {  // first maximal column of this lane: strict > keeps the lowest index
                bval[a] = v[a][0];
                bidx[a] = lane;
#pragma unroll
                for (int c = 1; c < 4; ++c)
                  if (v[a][c] > bval[a]) { bval[a] = v[a][c]; bidx[a] = lane + 32 * c; }
                if (ia[a] < 0 || lane >= G.agent[ia[a] >= 0 ? ia[a] : 0].actions) bidx[a] = 0x7fffffff;
              }
            } else {
#pragma unroll
              for (int a = 0; a < 4; ++a) {
                bval[a] = NegInf<QT>::v();
                bidx[a] = 0x7fffffff;
                if (ia[a] >= 0) {
                  const ThrlAgentSpec& s = G.agent[ia[a]];
                  const QT* row = tab + ((int)s.table_offset + ra[a] * s.row_stride);
                  for (int kk = lane; kk < s.actions; kk += 32) {
                    const QT v = row[kk];
                    if (v > bval[a] || bidx[a] == 0x7fffffff) { bval[a] = v; bidx[a] = kk; }
                  }
                }
              }
            }
#pragma unroll
            for (int a = 0; a < 4; ++a) {
              if (ia[a] < 0) continue;
              const QT wm = warp_max(bval[a]);  // numpy.argmax: first maximal index (agents.py:88)
              const int g = (int)__reduce_min_sync(kFull, (bidx[a] != 0x7fffffff && bval[a] == wm) ? (unsigned)bidx[a] : 0xffffffffu);
              const int gi = __shfl_sync(kFull, my_goff, ia[a]);
              if (lane == 0 && ra[a] < p.gcap[ia[a]]) Gc[gi + ra[a]] = (uint8_t)g;
              if (lane == ia[a]) k = g;
            }
          } while (need);
          __syncwarp();
        }
        if (is_agent) aq = lutAQ[my_lut + k];
        double Q = 0.0;  // environments.py:27 sum(A): ((0 + A0) + A1) + ...
        for (int i = 0; i < n; ++i) Q = __dadd_rn(Q, shfl_d(aq, i));
        const double na = p.noisy ? newa[t] : G.a;
        const double pn = __dsub_rn(na, __dmul_rn(G.b, Q));
        const double next_price = pn > 0.0 ? pn : (pn != pn ? pn : 0.0);  // numpy.max([0, x])
        const double rew = __dmul_rn(next_price, aq);                       // environments.py:34
        int nxt = pos + 1;
        if (nxt == Hp) nxt = 0;
        if (is_agent) {
          rlog = __dadd_rn(rlog, __ddiv_rn(rew, (double)T));  // trainer.py:65
          alog = __dadd_rn(alog, lutXT[my_lut + k]);          // trainer.py:66
          act[lane * Hp + pos] = (uint8_t)k;                  // trainer.py:61-62 memory.append
          my_len = my_len < my_cap ? my_len + 1 : my_cap;     // deque(maxlen=capacity)
          if (p.trace_actions) p.trace_actions[(step0 + t) * n + lane] = k;
          if (p.trace_rewards) p.trace_rewards[(step0 + t) * n + lane] = rew;
        }
        if (lane == 0) {
          P[nxt] = next_price;
          if (p.trace_prices) p.trace_prices[step0 + t] = next_price;
        }
        pos = nxt;
        price = next_price;  // trainer.py:67
      }
      __syncwarp();

      // ---- train_net for every agent (trainer.py:70, agents.py:59-78).  Agents only touch their own tables, so their
      //      sequential passes are independent chains: they are advanced together, up to eight agents' row loads in flight at once.
      int Lmax = 0;
      for (int i = 0; i < n; ++i) {
        const int L = __shfl_sync(kFull, my_len, i);
        const int fires = __shfl_sync(kFull, (int)(!never_fires && my_len >= my_minmem), i);
        if (fires && L > Lmax) Lmax = L;
        if (lane == 0) {
          int first = pos - L;
          if (first < 0) first += Hp;
          des[i * 4 + 0] = L; des[i * 4 + 1] = first; des[i * 4 + 2] = fires;
        }
      }
      __syncwarp();
      for (int i = 0; i < n; ++i) {  // encodes (:62,:66) and the stale snapshot (:67), lane-parallel
        if (!des[i * 4 + 2]) continue;
        const ThrlAgentSpec& s = G.agent[i];
        const int L = des[i * 4 + 0], first = des[i * 4 + 1], A = s.actions, RS = s.row_stride;
        uint16_t* rowbuf = rowbuf_all + i * p.row_stride;
        QT* oldv = oldv_all + i * p.old_stride;
        const QT* tb = tab + s.table_offset;
        if (lane == 0) des[i * 4 + 3] = Lmax - L;
        for (int j = lane; j <= L; j += 32) {
          int sl = first + j;
          if (sl >= Hp) sl -= Hp;
          rowbuf[j] = (uint16_t)upd_row(P[sl], s.max_state, (double)s.states);
        }
        __syncwarp();
        if (!kSmemTables && lane < kPrefetchAhead && lane < L) {  // rows the first transitions of the pass will read: into L2 now
          const QT* row = tb + (size_t)rowbuf[lane + 1] * RS;
          for (int b = 0; b < A * (int)sizeof(QT); b += 128)
            asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char*>(row) + b));
        }
        {  // snapshot: up to four scattered cells per lane are in flight before the first one is stored
          QT v4[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int j = lane + 32 * u;
            int sl = first + j;
            if (sl >= Hp) sl -= Hp;
            v4[u] = j < L ? tb[(int)rowbuf[j] * RS + act[i * Hp + sl]] : (QT)0;
          }
#pragma unroll
          for (int u = 0; u < 4; ++u)
            if (lane + 32 * u < L) oldv[lane + 32 * u] = v4[u];
        }
        for (int j = lane + 128; j < L; j += 32) {
          int sl = first + j;
          if (sl >= Hp) sl -= Hp;
          oldv[j] = tb[(int)rowbuf[j] * RS + act[i * Hp + sl]];
        }
      }
      __syncwarp();
      // Quarter path: the row a transition reads for its bootstrap is the row the agent's NEXT transition writes (s'_j =
      // s_{j+1}), and the column that write will hit is known (the action ring).  The row scan therefore yields the
      // (max, first argmax) over all OTHER columns plus the value of that column; the pair is carried to the next
      // iteration, where one comparison against the value just written gives the row's greedy action after the write.
      // The greedy cache survives the update instead of being invalidated by it, and later episodes act from the cache
      // without touching HBM.  cok = the carried pair is valid.
      QT cm[2] = {NegInf<QT>::v(), NegInf<QT>::v()};
      int ca[2] = {0, 0};
      bool cok[2] = {false, false};
      for (int j = 0; j < Lmax; ++j) {  // the sequential pass (:68-76)
        if (!kSmemTables && lane < n && des[lane * 4 + 2]) {  // HBM tables: pull the rows of step j + kPrefetchAhead into L2 early
          const int jp = j + kPrefetchAhead - des[lane * 4 + 3];
          if (jp >= 0 && jp < des[lane * 4 + 0]) {
            const ThrlAgentSpec& s = G.agent[lane];
            const QT* row = tab + s.table_offset + (size_t)rowbuf_all[lane * p.row_stride + jp + 1] * s.row_stride;
            for (int b = 0; b < s.actions * (int)sizeof(QT); b += 128)
              asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char*>(row) + b));
          }
        }
        auto group = [&](auto g0c) {
          constexpr int g0 = decltype(g0c)::value;  // compile-time base: per-agent constants become immediate operands
          QT loc[8];
          bool on[8];
          bool wide = false;  // some agent has more than 128 actions: general loop below
#pragma unroll
          for (int a = 0; a < 8; ++a) {
            const int i = g0 + a;
            on[a] = i < n && des[i * 4 + 2] && j >= des[i * 4 + 3];
            if (on[a] && G.agent[i].actions > 128) wide = true;
          }
          if (!wide) {
            // live row max (:71): every active agent's row is loaded into registers first (up to 8 x 4 loads in flight
            // per lane), only then reduced -- the agents' chains are independent, so their memory latencies overlap
            QT v[8][4];
#pragma unroll
            for (int a = 0; a < 8; ++a) {
              const int i = g0 + a;
              const ThrlAgentSpec& s = G.agent[on[a] ? i : 0];
              const int A = s.actions;
              const int ns = on[a] ? (int)rowbuf_all[i * p.row_stride + (j - des[i * 4 + 3]) + 1] : 0;
              const QT* row = tab + ((int)s.table_offset + ns * s.row_stride);
#pragma unroll
              for (int c = 0; c < 4; ++c) {
                const int kk = lane + 32 * c;
                v[a][c] = (on[a] && kk < A) ? row[kk] : NegInf<QT>::v();
              }
            }
#pragma unroll
            for (int a = 0; a < 8; ++a) {
              const QT m01 = v[a][0] > v[a][1] ? v[a][0] : v[a][1], m23 = v[a][2] > v[a][3] ? v[a][2] : v[a][3];
              loc[a] = m01 > m23 ? m01 : m23;
            }
          } else {
#pragma unroll
            for (int a = 0; a < 8; ++a) {
              const int i = g0 + a;
              loc[a] = NegInf<QT>::v();
              if (on[a]) {
                const ThrlAgentSpec& s = G.agent[i];
                const int ns = rowbuf_all[i * p.row_stride + (j - des[i * 4 + 3]) + 1];
                const QT* row = tab + ((int)s.table_offset + ns * s.row_stride);
                for (int k = lane; k < s.actions; k += 32) { const QT v = row[k]; loc[a] = v > loc[a] ? v : loc[a]; }
              }
            }
          }
#pragma unroll
          for (int a = 0; a < 8; ++a) {
            if (!on[a]) continue;
            const int i = g0 + a;
            const ThrlAgentSpec& s = G.agent[i];
            const int jj = j - des[i * 4 + 3];
            int sl = des[i * 4 + 1] + jj;
            if (sl >= Hp) sl -= Hp;
            int sn = sl + 1;
            if (sn == Hp) sn = 0;
            const int st = rowbuf_all[i * p.row_stride + jj], k = act[i * Hp + sl];
            const int loff = __shfl_sync(kFull, my_lut, i), goff = __shfl_sync(kFull, my_goff, i);
            const double alpha = hpw[i * 5 + 0], gamma = hpw[i * 5 + 1];
            const double reward = __dmul_rn(P[sn], lutAQ[loff + k]);
            const double next_max = (double)warp_max(loc[a]);
            const double nv = __dadd_rn(__dmul_rn(__dsub_rn(1.0, alpha), (double)oldv_all[i * p.old_stride + jj]),
                                        __dmul_rn(alpha, __dadd_rn(reward, __dmul_rn(gamma, next_max))));
            if ((k & 31) == lane) {  // the lane that owns column k: it alone ever reads or writes that column
              const int cell = (int)s.table_offset + st * s.row_stride + k;  // a run's slab has < 2^31 elements (checked at layout)
              tab[cell] = (QT)nv;
              if (cnt) atomicAdd(cnt + cell, 1u);  // :76, fire-and-forget RED
            }
            if (lane == 0 && st < p.gcap[i]) Gc[goff + st] = 0xFF;  // the row changed: its greedy action is recomputed on the next visit
          }
        };
        if (p.quarter) {
          // ---- quarter-warp path: phase 1 loads + row max for up to 8 agents (two passes of four), phase 2 finishes + stores
          QT qm[2];
          bool qon[2];
          int qjj[2], qidx[2], qns[2], qleft[2], qknx[2];
          QT qvk[2];
          auto load_rows = [&](auto ncc, auto fullc) {
            constexpr int NC = decltype(ncc)::value;  // columns per lane, compile time: every load is issued before any compare
            constexpr bool kFullRows = decltype(fullc)::value;
            QT v[2][NC];
#pragma unroll
            for (int g2 = 0; g2 < 2; ++g2) {
              const int i = qa_ok[g2] ? 4 * g2 + qq : 0;
              qon[g2] = qa_ok[g2] && des[i * 4 + 2] && j >= des[i * 4 + 3];
              qjj[g2] = qon[g2] ? j - des[i * 4 + 3] : 0;
              const int ns = qon[g2] ? (int)rowbuf_all[i * p.row_stride + qjj[g2] + 1] : 0;  // idle quarters read row 0 and are ignored
              const QT* row = tab + (qa_toff[g2] + ns * qa_RS[g2]);
              const int left = (qon[g2] ? qa_A[g2] : 0) - ql;  // columns at or beyond this lane's first one
              qns[g2] = ns;
              qleft[g2] = left;
              {  // column the agent's next transition writes in this row (-1: there is none)
                int sn2 = des[i * 4 + 1] + qjj[g2] + 1;
                if (sn2 >= Hp) sn2 -= Hp;
                qknx[g2] = (qon[g2] && qjj[g2] + 1 < des[i * 4 + 0]) ? (int)act[i * Hp + sn2] : -1;
              }
#pragma unroll
              for (int c = 0; c < NC; ++c)  // live table (:71): this lane's columns ql, ql+8, ...
                v[g2][c] = ((kFullRows && c < NC - 1) || 8 * c < left) ? row[ql + 8 * c] : NegInf<QT>::v();
            }
#pragma unroll
            for (int g2 = 0; g2 < 2; ++g2) {
              int ci[NC];
              const int cst = (qknx[g2] >= 0 && (qknx[g2] & 7) == ql) ? (qknx[g2] >> 3) : -1;  // this lane holds the excluded column
              qvk[g2] = NegInf<QT>::v();
#pragma unroll
              for (int c = 0; c < NC; ++c) {
                ci[c] = c;
                if (c == cst) { qvk[g2] = v[g2][c]; v[g2][c] = NegInf<QT>::v(); }
              }
#pragma unroll
              for (int st = 1; st < NC; st *= 2) {  // pairwise tree: depth log2 NC; the left operand (lower columns) wins ties
#pragma unroll
                for (int c = 0; c + st < NC; c += 2 * st)
                  if (v[g2][c + st] > v[g2][c]) { v[g2][c] = v[g2][c + st]; ci[c] = ci[c + st]; }
              }
              qm[g2] = v[g2][0];
              qidx[g2] = qleft[g2] > 0 ? ql + 8 * ci[0] : 0x7fffffff;
            }
          };
          if (p.qfull) {
            if (p.qchunks <= 4) load_rows(std::integral_constant<int, 4>{}, std::true_type{});
            else if (p.qchunks <= 8) load_rows(std::integral_constant<int, 8>{}, std::true_type{});
            else if (p.qchunks <= 13) load_rows(std::integral_constant<int, 13>{}, std::true_type{});
            else load_rows(std::integral_constant<int, 16>{}, std::true_type{});
          } else {
            if (p.qchunks <= 4) load_rows(std::integral_constant<int, 4>{}, std::false_type{});
            else if (p.qchunks <= 8) load_rows(std::integral_constant<int, 8>{}, std::false_type{});
            else if (p.qchunks <= 13) load_rows(std::integral_constant<int, 13>{}, std::false_type{});
            else load_rows(std::integral_constant<int, 16>{}, std::false_type{});
          }
#pragma unroll
          for (int g2 = 0; g2 < 2; ++g2) {
            QT mex = qm[g2];  // (max, first maximal column) over the quarter's 8 lanes, the next write's column excluded
            int am = qidx[g2];
#pragma unroll
            for (int off = 4; off >= 1; off >>= 1) {
              const QT o = shfl_xor_t(mex, off);
              const int oi = __shfl_xor_sync(kFull, am, off);
              if (oi != 0x7fffffff && (am == 0x7fffffff || o > mex || (o == mex && oi < am))) { mex = o; am = oi; }
            }
            const QT vkq = shfl_t(qvk[g2], (lane & 24) | (qknx[g2] & 7));  // value of the excluded column (-inf: none)
            const QT m = vkq > mex ? vkq : mex;                              // live row max (:71)
            const int i = qa_ok[g2] ? 4 * g2 + qq : 0;
            const int jj = qjj[g2];
            int sl = des[i * 4 + 1] + jj;
            if (sl >= Hp) sl -= Hp;
            int sn = sl + 1;
            if (sn == Hp) sn = 0;
            const int st = rowbuf_all[i * p.row_stride + jj], k = act[i * Hp + sl];
            const double reward = __dmul_rn(P[sn], lutAQ[qa_loff[g2] + k]);
            const double nv = __dadd_rn(__dmul_rn(qa_oma[g2], (double)oldv_all[i * p.old_stride + jj]),
                                        __dmul_rn(qa_alpha[g2], __dadd_rn(reward, __dmul_rn(qa_gamma[g2], (double)m))));
            if (qon[g2] && (k & 7) == ql) {  // the lane of this quarter that owns column k: it alone reads or writes it
              const int cell = qa_toff[g2] + st * qa_RS[g2] + k;
              tab[cell] = (QT)nv;
              if (cnt) atomicAdd(cnt + cell, 1u);  // :76, fire-and-forget RED
            }
            // (max, argmax) of a row after its cell k became x.  Unknown when the holder of the maximum went down.
            auto patch = [](QT& rm, int& ra, bool& ok, int kk, QT x) {
              if (kk != ra) {
                if (x > rm || (x == rm && kk < ra)) { rm = x; ra = kk; }
              } else if (x >= rm) {
                rm = x;
              } else {
                ok = false;
              }
            };
            const QT nvq = (QT)nv;
            // row st as this write leaves it: the carried pair covers every column but k, the one written now
            if (qon[g2] && ql == 0 && st < qa_gcap[g2]) {
              const bool mine = nvq > cm[g2] || (nvq == cm[g2] && k < ca[g2]);
              Gc[qa_goff[g2] + st] = cok[g2] ? (uint8_t)(mine ? k : ca[g2]) : (uint8_t)0xFF;
            }
            // the row just read is the one the next transition writes (its column excluded); if it is also the row written
            // just now, apply that write to the pair -- unless it hit the excluded column itself
            cm[g2] = mex;
            ca[g2] = am;
            cok[g2] = qon[g2] && qknx[g2] >= 0 && am != 0x7fffffff;
            if (qns[g2] == st && k != qknx[g2]) patch(cm[g2], ca[g2], cok[g2], k, nvq);
          }
        } else {
          group(std::integral_constant<int, 0>{});
          if (n > 8) group(std::integral_constant<int, 8>{});
        }
      }
      for (int i = 0; i < n; ++i)
        if (des[i * 4 + 2] && lane == i) my_len = 0;  // :77 memory.empty()
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

    // ---- write the run back
    if (kSmemTables) {
      for (long long c = lane; c < G.run_stride; c += 32) qg[c] = tab_s[c];
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

}  // namespace thrl
