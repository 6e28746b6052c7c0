// thrl_scan_hbm.cuh — Q-table games whose tables stay in HBM (BASELINE C4: 8 agents x 1001 x 101 fp32 = 3.2 MB per run).
//
// One warp plays one run, persistent grid.  The rollout (trainer.py:50-67) acts from an exact greedy-action cache in
// shared memory, so it touches HBM only on the first visit of a table row.  QTable.train_net (agents.py:59-78: stale
// snapshot, live row max, writes in batch order) is NOT executed as a load -> max -> store chain through memory:
//
//   U1  float64 encodes of the episode's states (agents.py:62,66), once per *state class* (agents that share max_state,
//       states and batch length see the same rows), lane = state.
//   U2  the states are grouped by table row: a chain through the states of one row in ascending order, its first state
//       is the row's representative.  Only DISTINCT rows are gathered.
//   U3  the distinct rows of every agent are fetched from HBM with 1-D bulk copies (cp.async.bulk: one 16-byte aligned
//       padded row = one copy issued by one lane, completion counted in bytes on an mbarrier) into a ring of staging
//       batches in shared memory; every copy is independent of every other, `nb` batches are in flight per warp.
//       When a batch lands, one lane per row walks the row's transitions in order: the staged cell is the stale
//       snapshot (agents.py:67) of the first transition that writes it, which then TAGS the staged cell (a NaN-boxed
//       transition index) -- later writers of the cell find the tag and share the first writer's slot.  Then `lpr`
//       lanes per row scan the staged row for (max, first argmax) over its UNTAGGED cells: a hardware float compare
//       never selects a NaN, so the tags drop out for free.  That pair cannot be changed by any write of this batch.
//   U4  lane = agent walks its transitions in order entirely on chip: the live row max of agents.py:71 is the untagged
//       maximum of the next state's row merged with the current values of that row's rewritten cells (a short list);
//       the new value goes to the cell's slot in shared memory.
//   U5  the final value of every rewritten cell is stored to HBM (one 4-byte store per distinct cell), visit counters
//       get fire-and-forget REDs, and the greedy-action cache of every touched row is refreshed exactly from the
//       untagged (max, argmax) and the final values of the row's rewritten cells.
// scripts/model_hbm_update.py replays U2-U5 on random batches against the plain sequential form.
// HBM traffic per agent-step: at most one padded row (row_stride * 4 B; less when states repeat), one sector for the
// cell store and one for the counter; the greedy row read of SURVEY 8(d)'s algorithmic count is served by the cache.
//
// Applies to: Q-table agents only, regular games (every agent's batch is the newest min(T, capacity) transitions of the
// episode), max_steps <= 254, <= 128 actions, padded slab layout (include/thrl.h ThrlAgentSpec.row_stride).  Everything
// else runs on thrl_scan_generic.cuh, which is also the checker this kernel is tested against (THRL_KERNEL=generic).
#pragma once
#include "thrl_device.cuh"

namespace thrl {

constexpr int kHbmMaxT = 254;   // transition / state indices fit one byte next to 0xFF = none
constexpr int kHbmMaxNb = 8;    // staging ring depth (batches)
constexpr int kHbmAgc = 17;     // ints per agent in the CTA-shared constant block (odd: lanes that read different agents' entries hit different banks)
constexpr int kHbmMaxWarps = 14;  // resident runs per CTA: 448 threads leave 146 registers per thread
constexpr int kHbmRegWarps = 12;  // register landing: 12 resident runs fit the 196 KB carve-out (plan_hbm); 384 threads leave 170 registers

struct HbmParams {
  ThrlGame game;
  long long n_runs, run_id0, per_round;  // per_round: runs played at a time (balanced over the rounds that are needed anyway)
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
  // state classes: agents with equal (max_state, states, batch length) share encodes, row groups and gather order
  int ncls;
  int cls_of[THRL_MAX_AGENTS];     // class of agent i (-1: the agent never updates)
  int cls_beg[THRL_MAX_AGENTS];    // class c's agents are cls_agent[cls_beg[c] .. cls_beg[c] + cls_na[c])
  int cls_na[THRL_MAX_AGENTS];
  int cls_agent[THRL_MAX_AGENTS];
  int lut_owner[THRL_MAX_AGENTS];  // first agent with the same action grid: they share one action LUT
  int lut_total, rows_total;
  int noisy;       // new_a varies per step
  int staged;      // 1: rows are staged in shared memory (bulk copies); 0: rows land in registers (16-byte vector loads, L2 prefetch)
  int bulk;        // staged rows arrive by: 1 = cp.async.bulk (TMA unit); 0 = 16-byte vector loads + shared stores (comparison)
  int pf_dist;     // register landing: batches prefetched into L2 ahead of the one being loaded
  int nch_max;     // register landing: 16-byte chunks of the longest row
  int seg;         // steps whose reward / max_steps are divided at a time (U4)
  int nb;          // staging ring depth (batches)
  int lpr_shift;   // lanes per row = 1 << lpr_shift; a batch holds 32 >> lpr_shift rows
  int slot_bytes;  // one staged row (padded so that the lanes of a quarter-warp hit distinct banks)
  int Tp, Sp;      // padded strides of the [.][T] / [.][T+1] arrays (elements)
  // shared memory (bytes): [cta_bytes][warp 0][warp 1]...
  int cta_bytes, warp_bytes;
  // Dynamic scheduling (nchunk > 1): a run is played as nchunk tasks of echunk epochs; the warps of a CTA take the tasks of the
  // CTA's runs from a shared-memory queue (off_q: next task, then one done-counter per local run), see the kernel.
  int nchunk, echunk, off_q;
  int off_bar, off_g, off_P, off_hp, off_miss, off_mask, off_srow, off_rs, off_next, off_dl, off_act, off_cur, off_canon, off_nextc,
      off_bm, off_stage;
  // regions that alias the staging ring: the rollout's draws (before the update), reward / max_steps (after the gather)
  int off_pre, off_newa, off_rq;
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
__device__ __forceinline__ void hbm_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(hbm_smem_addr(dst)),
               "l"(src), "r"(bytes), "r"(hbm_smem_addr(bar))
               : "memory");
}
// generic-proxy accesses (this warp's table stores in global memory, its use of the staging ring in shared memory) are
// ordered before the async-proxy accesses of the bulk copies issued after the fence
__device__ __forceinline__ void hbm_fence_proxy_async() {
  asm volatile("fence.proxy.async.global;\n\tfence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void hbm_fence_proxy_async_shared() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---------------------------------------------------------------- tags: NaN-boxed transition indices
// A tagged staged cell holds kBase | t: a negative quiet NaN.  `v > best` is false for it, so the row scan skips tagged
// cells without testing for them; as an unsigned integer it lies above every number (and -inf).  The update arithmetic
// never produces such a bit pattern (its NaNs, should inputs be NaN, are positive).
template <typename QT> struct HbmBits;
template <> struct HbmBits<float> {
  using U = uint32_t;
  static constexpr U kBase = 0xFFC00000u;
  __device__ static U of(float v) { return __float_as_uint(v); }
  __device__ static float val(U b) { return __uint_as_float(b); }
};
template <> struct HbmBits<double> {
  using U = unsigned long long;
  static constexpr U kBase = 0xFFF8000000000000ull;
  __device__ static U of(double v) { return (U)__double_as_longlong(v); }
  __device__ static double val(U b) { return __longlong_as_double((long long)b); }
};

// one 16-byte chunk of a row: 4 floats / 2 doubles
template <bool kGlobal> __device__ __forceinline__ void hbm_chunk(const float* row, int c, float (&v)[4]) {
  const float4* p = reinterpret_cast<const float4*>(row) + c;
  const float4 x = kGlobal ? __ldcg(p) : *p;
  v[0] = x.x; v[1] = x.y; v[2] = x.z; v[3] = x.w;
}
template <bool kGlobal> __device__ __forceinline__ void hbm_chunk(const double* row, int c, double (&v)[2]) {
  const double2* p = reinterpret_cast<const double2*>(row) + c;
  const double2 x = kGlobal ? __ldcg(p) : *p;
  v[0] = x.x; v[1] = x.y;
}

// (max, first argmax) over columns [0, A) of a 16-byte aligned row, NaNs (= tags) skipped.  The lanes q = 0 .. lpr-1 of a
// group take the 16-byte chunks q, q + lpr, ... (a quarter-warp reads consecutive chunks: conflict-free on the padded
// staging slots); two independent compare chains per lane, merged by (value, then smaller column); then a butterfly over
// the group, after which every lane of the group holds the result.  bidx = 0xffffffff: no column holds a value > -inf.
template <typename QT, bool kGlobal>
__device__ __forceinline__ void hbm_scan_row(const QT* row, int A, int q, int lpr, bool active, QT& best, unsigned& bidx) {
  constexpr int EPC = 16 / (int)sizeof(QT);
  constexpr unsigned kNone = 0xffffffffu;
  QT b0 = NegInf<QT>::v(), b1 = NegInf<QT>::v();
  unsigned i0 = kNone, i1 = kNone;
  if (active) {
    const int nfull = A / EPC;
    int c = q;
    for (; c + lpr < nfull; c += 2 * lpr) {
      QT v[EPC], w[EPC];
      hbm_chunk<kGlobal>(row, c, v);
      hbm_chunk<kGlobal>(row, c + lpr, w);
#pragma unroll
      for (int e = 0; e < EPC; ++e) {  // ascending columns, strict >: first maximal index (agents.py:88)
        if (v[e] > b0) { b0 = v[e]; i0 = (unsigned)(c * EPC + e); }
        if (w[e] > b1) { b1 = w[e]; i1 = (unsigned)((c + lpr) * EPC + e); }
      }
    }
    if (c < nfull) {
      QT v[EPC];
      hbm_chunk<kGlobal>(row, c, v);
#pragma unroll
      for (int e = 0; e < EPC; ++e)
        if (v[e] > b0) { b0 = v[e]; i0 = (unsigned)(c * EPC + e); }
    }
    if (b1 > b0 || (b1 == b0 && i1 < i0)) { b0 = b1; i0 = i1; }
    if (nfull * EPC < A && (nfull & (lpr - 1)) == q) {  // the partial last chunk: padding cells are not values
      QT v[EPC];
      hbm_chunk<kGlobal>(row, nfull, v);
#pragma unroll
      for (int e = 0; e < EPC; ++e)
        if (nfull * EPC + e < A && v[e] > b0) { b0 = v[e]; i0 = (unsigned)(nfull * EPC + e); }  // highest columns: after the merge
    }
  }
  for (int off = 1; off < lpr; off <<= 1) {
    const QT ob = shfl_xor_t(b0, off);
    const unsigned oi = __shfl_xor_sync(kFull, i0, off);
    if (ob > b0 || (ob == b0 && oi < i0)) { b0 = ob; i0 = oi; }
  }
  best = b0;
  bidx = i0;
}

// ---------------------------------------------------------------- register landing (kStaged = false)
// lanes per row and 16-byte chunk slots per lane are compile-time so that a row lives in registers: fp32 rows (<= 128 columns =
// 32 chunks) take 2 lanes x 16 slots, f64 rows (64 chunks) 4 lanes x 16 slots; lanes-per-row x elements-per-chunk = 8 either way
template <typename QT> struct HbmReg {
  static constexpr int kLprShift = sizeof(QT) == 4 ? 1 : 2;
  static constexpr int kLpr = 1 << kLprShift;
  static constexpr int kRows = 32 >> kLprShift;  // rows per batch
  static constexpr int kEpc = 16 / (int)sizeof(QT);
  static constexpr int kSlots = 16;
};
__device__ __forceinline__ void hbm_prefetch_l2(const void* ptr) { asm volatile("prefetch.global.L2 [%0];" ::"l"(ptr)); }

// kDyn: dynamic schedule (HbmParams.nchunk > 1), compiled separately so that the static schedule keeps its code
template <typename QT, bool kStaged, bool kDyn = false>
__global__ void __launch_bounds__(kStaged ? 32 * kHbmMaxWarps : 32 * kHbmRegWarps, 1) qtable_scan_hbm(const __grid_constant__ HbmParams p) {
  using B = HbmBits<QT>;
  using U = typename B::U;
  using R = HbmReg<QT>;
  extern __shared__ __align__(128) unsigned char smem_hbm[];
  unsigned char* const smem = smem_hbm;
  const ThrlGame& G = p.game;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, warps_per_cta = blockDim.x >> 5;
  const int n = G.n_agents, T = G.max_steps, E = p.E, Tp = p.Tp, Sp = p.Sp, nb = p.nb;
  const int lpr_shift = kStaged ? p.lpr_shift : R::kLprShift;
  const int lpr = 1 << lpr_shift, rb = 32 >> lpr_shift;     // lanes per gathered row, rows per batch
  const int grp = lane >> lpr_shift, qq = lane & (lpr - 1);  // this lane's row of the batch and its part of the row
  const bool is_agent = lane < n;
  constexpr unsigned kNone = 0xffffffffu;

  // ---- CTA-shared: AQ[k] = (a/b)*scale(k), XT[k] = scale(k)/max_steps per agent, and the per-agent constants of the update
  double* lutAQ = reinterpret_cast<double*>(smem);
  double* lutXT = lutAQ + p.lut_total;
  int* agc = reinterpret_cast<int*>(lutXT + p.lut_total);  // [n][kHbmAgc]: table offset, row stride, actions, lut offset, cache offset,
                                                           // cache rows, t0, L, class, -, -, -, 4 words: columns >= actions
  {
    const double ab = __ddiv_rn(G.a, G.b);  // environments.py:23 self.a/self.b
    for (int i = 0; i < n; ++i) {
      const ThrlAgentSpec& s = G.agent[i];
      if (p.lut_owner[i] == i)
        for (int k = threadIdx.x; k < s.actions; k += blockDim.x) {
          const double x = scale_action(k, s.actions, s.action_lo, s.action_hi);
          lutAQ[p.loff[i] + k] = __dmul_rn(ab, x);
          lutXT[p.loff[i] + k] = __ddiv_rn(x, (double)T);  // trainer.py:66 scaled_acts / max_steps
        }
      if (threadIdx.x == 0) {
        int* a = agc + i * kHbmAgc;
        a[0] = (int)s.table_offset; a[1] = s.row_stride; a[2] = s.actions; a[3] = p.loff[i];
        a[4] = p.goff[i]; a[5] = p.gcap[i]; a[6] = T - p.L[i]; a[7] = p.L[i]; a[8] = p.cls_of[i];
        for (int w = 0; w < 4; ++w)
          a[12 + w] = (int)(s.actions >= 32 * w + 32 ? 0u : (s.actions <= 32 * w ? 0xffffffffu : (0xffffffffu << (s.actions - 32 * w))));
      }
    }
  }
  __syncthreads();

  // ---- this warp's slot
  unsigned char* slot = smem + p.cta_bytes + (size_t)warp * p.warp_bytes;
  uint64_t* bar = reinterpret_cast<uint64_t*>(slot + p.off_bar);     // [nb] one mbarrier per staging batch
  uint8_t* Gc = slot + p.off_g;                                       // greedy action per (agent,row); 0xFF = unknown
  double* P = reinterpret_cast<double*>(slot + p.off_P);              // [T+1] prices of the episode, P[0] = state it starts from
  double* hpw = reinterpret_cast<double*>(slot + p.off_hp);           // [n][5] alpha,gamma,eps_end,eps_step,eps
  uint8_t* missk = slot + p.off_miss;                                 // [16] greedy action found by a cache miss
  uint8_t* ndc = missk + THRL_MAX_AGENTS;                             // [16] distinct rows of the episode per state class
  uint16_t* srow = reinterpret_cast<uint16_t*>(slot + p.off_srow);    // [ncls][Sp] float64 encode of every state (agents.py:62,66)
  uint8_t* rs = slot + p.off_rs;                                      // [ncls][Sp] first state of the episode with the same row
  uint8_t* nexts = slot + p.off_next;                                 // [ncls][Sp] next state with the same row (ascending), 0xFF = last
  uint8_t* dl = slot + p.off_dl;                                      // [ncls][Sp] the representatives = distinct rows, ascending
  uint8_t* act = slot + p.off_act;                                    // [n][Tp] chosen actions
  QT* cur = reinterpret_cast<QT*>(slot + p.off_cur);                  // [n][Tp] snapshot; for a cell's first writer: the cell's live value
  uint8_t* canon = slot + p.off_canon;                                // [n][Tp] first transition of the batch that writes the same cell
  uint8_t* nextc = slot + p.off_nextc;                                // [n][Tp] a row's first writers (= its rewritten cells), ascending from the
                                                                      //         representative transition itself; 0xFF = last
  QT* bm = reinterpret_cast<QT*>(slot + p.off_bm);                    // [n][Sp] max over the untagged cells of the representative's row; its
                                                                      //         first column waits in the row's greedy-cache byte (0xFF: none)
  unsigned char* stage = slot + p.off_stage;                          // [nb][rb][slot_bytes] staged rows (kStaged)
  int8_t* pre = reinterpret_cast<int8_t*>(slot + p.off_pre);          // [T][n] forced action or -1 (= greedy)          } alias the
  uint32_t* maskS = reinterpret_cast<uint32_t*>(slot + p.off_mask);   // [rows per batch][4] rewritten / padding columns } (register landing)
  double* newa = reinterpret_cast<double*>(slot + p.off_newa);        // [T] demand intercept (noisy only)               } staging
  double* rq = reinterpret_cast<double*>(slot + p.off_rq);            // [n][seg] reward / max_steps (trainer.py:65)     } ring
  const int seg = p.seg;

  uint32_t parity = 0;  // bit b: phase the next wait on bar[b] completes
  if (kStaged) {
    if (lane == 0)
      for (int b = 0; b < nb; ++b) hbm_mbar_init(bar + b, p.bulk ? 1u : 32u);  // bulk: one arrival + byte count; else every lane arrives
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    hbm_fence_proxy_async();
    __syncwarp();
  }

  // ---- lane i < n keeps agent i's constants in registers
  int my_lut = 0, my_goff = 0, my_gcap = 0, my_t0 = T, my_cls = -1;
  float my_msf = 1.f, my_sf = 1.f;
  if (is_agent) {
    const ThrlAgentSpec& s = G.agent[lane];
    my_msf = (float)s.max_state;
    my_sf = (float)s.states;
    my_lut = p.loff[lane]; my_goff = p.goff[lane]; my_gcap = p.gcap[lane];
    my_cls = p.cls_of[lane];
    if (p.L[lane] > 0) my_t0 = T - p.L[lane];
  }
  const double Td = (double)T;
  const bool tracing = p.trace_actions || p.trace_rewards || p.trace_prices;

  // Static schedule (nchunk <= 1): slot w * grid + b plays runs slot, slot + per_round, ...: a round that does not fill every
  // warp leaves the idle warps spread over all SMs.
  // Dynamic schedule (nchunk > 1), for launches whose run count is not a whole number of rounds (4,096 runs = 2.3 rounds of
  // 1,776 leave two thirds of the slots idle for a third of the launch): CTA b owns the runs b, b + grid, ...; every run is cut
  // into nchunk tasks of echunk epochs, ordered chunk by chunk; a warp that is free takes the next task of its CTA from a
  // shared-memory counter.  A run's state between two of its tasks lives where it always does (tables in HBM, epsilon and
  // price in their arrays); only the greedy cache starts cold.  Task (run, c) waits until (run, c - 1) is done -- a counter per
  // local run in shared memory; the producer is a resident warp of the same CTA that never waits on a later task, so the wait
  // terminates (and traps instead of hanging should that invariant ever be broken).
  const long long slot_id = (long long)warp * gridDim.x + blockIdx.x;
  int* qnext = reinterpret_cast<int*>(smem + p.off_q);
  volatile int* qdone = qnext + 4;
  const int nloc = kDyn && blockIdx.x < p.n_runs ? (int)((p.n_runs - blockIdx.x + gridDim.x - 1) / gridDim.x) : 0;
  if (kDyn) {
    if (threadIdx.x == 0) *qnext = 0;
    for (int i = threadIdx.x; i < nloc; i += blockDim.x) qdone[i] = 0;
    __syncthreads();
  }
  for (long long it = 0;; ++it) {
    long long r;
    int e_lo = 0, e_hi = E, iloc = 0, cidx = 0;
    if (!kDyn) {
      r = slot_id + it * p.per_round;
      if (!(slot_id < p.per_round && r < p.n_runs)) break;
    } else {
      int t = 0;
      if (lane == 0) t = atomicAdd(qnext, 1);
      t = __shfl_sync(kFull, t, 0);
      if (t >= nloc * p.nchunk) break;
      cidx = t / nloc;
      iloc = t - cidx * nloc;
      r = blockIdx.x + (long long)gridDim.x * iloc;
      e_lo = cidx * p.echunk;
      e_hi = e_lo + p.echunk < E ? e_lo + p.echunk : E;
      if (cidx > 0) {
        if (lane == 0) {
          unsigned spins = 0;
          while (qdone[iloc] < cidx) {
            __nanosleep(256);
            if (++spins > (1u << 31)) __trap();  // ~10 minutes: far beyond any predecessor task
          }
        }
        __syncwarp();
        __threadfence_block();
      }
    }
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

    for (int e = e_lo; e < e_hi; ++e) {
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
              const int ra = (int)__umulhi(x[2 * h + 1], (uint32_t)agc[i * kHbmAgc + 2]);
              pre[t * n + i] = (int8_t)(u < hpw[i * 5 + 4] ? ra : -1);  // agents.py:81-82
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
          pre[idx] = (int8_t)v;
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
      int kpre = is_agent ? (int)pre[lane] : 0;
      for (int t = 0; t < T; ++t) {
        int k = kpre, arow = 0;
        if (is_agent && t + 1 < T) kpre = pre[(t + 1) * n + lane];  // next step's draw does not depend on this step
        if (is_agent && k < 0) {  // agents.py:84-88 on the frozen table
          arow = act_row(price, my_msf, my_sf);
          const int g = arow < my_gcap ? (int)Gc[my_goff + arow] : 0xFF;  // rows beyond the cache: always recomputed
          k = g == 0xFF ? -1 : g;
        }
        const unsigned need = __ballot_sync(kFull, is_agent && k < 0);  // first visit of a row: its greedy action is not known yet
        if (need) {  // lane group g reads the row of the g-th agent that missed, straight from HBM
          unsigned left = need;
          while (left) {
            const unsigned tgt = __fns(left, 0, grp + 1);  // lane of this group's agent, 0xffffffff: none left for it
            const bool has = tgt < 32u;
            const int ri = __shfl_sync(kFull, arow, has ? (int)tgt : 0);
            const int* a = agc + (has ? (int)tgt : 0) * kHbmAgc;
            QT best;
            unsigned bidx;
            hbm_scan_row<QT, true>(qg + a[0] + (size_t)ri * a[1], a[2], qq, lpr, has, best, bidx);
            if (has && qq == 0) {
              const uint8_t g = bidx == kNone ? (uint8_t)0 : (uint8_t)bidx;  // numpy.argmax of a row without a value > -inf: column 0
              if (ri < a[5]) Gc[a[4] + ri] = g;
              missk[tgt] = g;
            }
            for (int z = 0; z < rb && left; ++z) left &= left - 1;
          }
          __syncwarp();
          if (is_agent && k < 0) k = missk[lane];
        }
        double aq = 0.0;
        if (is_agent) aq = lutAQ[my_lut + k];
        double Q = 0.0;  // environments.py:27 sum(A): ((0 + A0) + A1) + ...
        if (n <= 8) {    // all shuffles first, then the dependent additions; lanes >= n hold +0.0, which leaves the sum unchanged
          double v[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) v[i] = shfl_d(aq, i);
#pragma unroll
          for (int i = 0; i < 8; ++i) Q = __dadd_rn(Q, v[i]);
        } else {
          for (int i = 0; i < n; ++i) Q = __dadd_rn(Q, shfl_d(aq, i));
        }
        const double na = p.noisy ? newa[t] : G.a;
        const double pn = __dsub_rn(na, __dmul_rn(G.b, Q));
        const double next_price = pn > 0.0 ? pn : (pn != pn ? pn : 0.0);  // numpy.max([0, x])
        if (is_agent) act[lane * Tp + t] = (uint8_t)k;  // trainer.py:61-62 memory.append
        if (lane == 0) P[t + 1] = next_price;
        if (tracing) {
          if (is_agent) {
            if (p.trace_actions) p.trace_actions[(step0 + t) * n + lane] = k;
            if (p.trace_rewards) p.trace_rewards[(step0 + t) * n + lane] = __dmul_rn(next_price, aq);  // environments.py:34
          }
          if (lane == 0 && p.trace_prices) p.trace_prices[step0 + t] = next_price;
        }
        price = next_price;  // trainer.py:67
      }
      __syncwarp();

      // ================================================================ train_net for every agent (trainer.py:70)
      // ---- U1. float64 encodes (agents.py:62,66), once per state class, lane = state
      for (int c = 0; c < p.ncls; ++c) {
        const int i0 = p.cls_agent[p.cls_beg[c]];
        const int t0 = agc[i0 * kHbmAgc + 6];
        const double ms = G.agent[i0].max_state, st = (double)G.agent[i0].states;
        for (int t = t0 + lane; t <= T; t += 32) srow[c * Sp + t] = (uint16_t)upd_row(P[t], ms, st);
      }
      __syncwarp();
      // ---- U2. group the states by row: a chain through the states of one row in ascending order.  The greedy-cache byte of a
      //      touched row serves as the chain head while the groups are formed (it is refreshed in U5 anyway); rows beyond the
      //      cache (only a call's initial price can lie there) are searched linearly.
      for (int c = 0; c < p.ncls; ++c) {
        const int* a = agc + p.cls_agent[p.cls_beg[c]] * kHbmAgc;
        const uint16_t* sr = srow + c * Sp;
        const int t0 = a[6], gcap = a[5];
        uint8_t* Gh = Gc + a[4];
        for (int t = t0 + lane; t <= T; t += 32) {
          const int row = sr[t];
          if (row < gcap) Gh[row] = 0xFF;
        }
        __syncwarp();
        for (int tb = t0 + ((T - t0) & ~31); tb >= t0; tb -= 32) {  // 32 states at a time, last chunk first
          const int t = tb + lane;
          const bool valid = t <= T;
          const int row = valid ? (int)sr[t] : 0x10000 + lane;  // idle lanes match nobody
          const unsigned same = __match_any_sync(kFull, row);
          const unsigned above = same & ~((2u << lane) - 1u);
          uint8_t nx = 0xFF;
          if (above) {
            nx = (uint8_t)(tb + __ffs(above) - 1);
          } else if (valid) {
            if (row < gcap) {
              nx = Gh[row];  // first state of the later chunks with this row
            } else {
              for (int u = tb + 32; u <= T; ++u)
                if (sr[u] == row) { nx = (uint8_t)u; break; }
            }
          }
          if (valid) nexts[c * Sp + t] = nx;
          __syncwarp();
          if (valid && row < gcap && (same & ((1u << lane) - 1u)) == 0u) Gh[row] = (uint8_t)t;
          __syncwarp();
        }
        int nd = 0;  // representatives = distinct rows, compacted in ascending order
        for (int tb = t0; tb <= T; tb += 32) {
          const int t = tb + lane;
          bool isrep = false;
          if (t <= T) {
            const int row = sr[t];
            int rep = t;
            if (row < gcap) {
              rep = Gh[row];
            } else {
              for (int u = t0; u < t; ++u)
                if (sr[u] == row) { rep = u; break; }
            }
            rs[c * Sp + t] = (uint8_t)rep;
            isrep = rep == t;
          }
          const unsigned m = __ballot_sync(kFull, isrep);
          if (isrep) dl[c * Sp + nd + __popc(m & ((1u << lane) - 1u))] = (uint8_t)t;
          nd += __popc(m);
        }
        if (lane == 0) ndc[c] = (uint8_t)nd;  // <= T + 1 <= 255
      }
      // visit counters (agents.py:76): fire-and-forget REDs, lane = transition
      if (cnt) {
        for (int i = 0; i < n; ++i) {
          const int* a = agc + i * kHbmAgc;
          if (a[7] == 0) continue;
          const uint16_t* sr = srow + a[8] * Sp;
          for (int t = a[6] + lane; t < T; t += 32) atomicAdd(cnt + a[0] + (int)sr[t] * a[1] + act[i * Tp + t], 1u);
        }
      }
      if (kStaged) hbm_fence_proxy_async();  // this warp's generic accesses to the staging ring (pre / newa / rq) and to the tables
      __syncwarp();                          // (last episode's stores) are ordered before the bulk copies that follow

      // ---- U3. gather the distinct rows; per row: stale snapshots + first writers of its rewritten cells, then (max, first
      //      argmax) over the cells no transition of this batch writes
      for (int c = 0; c < p.ncls; ++c) {
        const int na = p.cls_na[c], abeg = p.cls_beg[c];
        const int nd = ndc[c];
        const int items = nd * na;  // item g = (distinct row g / na, agent g % na): the agents of one state land together
        const int nbatch = (items + rb - 1) / rb;
        const int dstep = rb / na, astep = rb - dstep * na;  // item index advances by rb per batch, without a division
        if (kStaged) {
          int di = grp / na, ai = grp - di * na;    // item of the batch being issued
          int dp = di, ap = ai;                     // item of the batch being processed
          auto issue = [&](int b) {
            uint32_t bytes = 0;
            const QT* src = nullptr;
            if (di < nd) {
              const int* a = agc + p.cls_agent[abeg + ai] * kHbmAgc;
              bytes = (uint32_t)a[1] * (uint32_t)sizeof(QT);
              src = qg + a[0] + (size_t)srow[c * Sp + dl[c * Sp + di]] * a[1];
            }
            unsigned char* dst = stage + ((size_t)b * rb + grp) * p.slot_bytes;
            if (p.bulk) {
              const uint32_t total = __reduce_add_sync(kFull, qq == 0 ? bytes : 0u);
              if (lane == 0) hbm_mbar_expect_tx(bar + b, total);
              if (qq == 0 && bytes) hbm_bulk_g2s(dst, src, bytes, bar + b);
            } else {  // comparison path: the same rows with 16-byte vector loads that bypass L1
              for (uint32_t o = 16u * qq; o < bytes; o += 16u * lpr)
                *reinterpret_cast<int4*>(dst + o) = __ldcg(reinterpret_cast<const int4*>(reinterpret_cast<const unsigned char*>(src) + o));
              hbm_mbar_arrive(bar + b);
            }
            di += dstep; ai += astep;
            if (ai >= na) { ai -= na; ++di; }
          };
          for (int k = 0; k < nb && k < nbatch; ++k) issue(k);
          for (int k = 0; k < nbatch; ++k) {
            const int b = k % nb;
            hbm_mbar_wait(bar + b, (parity >> b) & 1u);
            parity ^= 1u << b;
            const bool has = dp < nd;
            const int i = p.cls_agent[abeg + ap];
            const int rep = has ? (int)dl[c * Sp + dp] : 0;
            const int* a = agc + i * kHbmAgc;
            QT* srow_s = reinterpret_cast<QT*>(stage + ((size_t)b * rb + grp) * p.slot_bytes);
            if (has && qq == 0) {
              // the row's transitions in ascending order: the first writer of a cell takes the staged value as everybody's
              // stale snapshot (agents.py:67) and leaves its index; later writers of the cell share its slot
              U* cells = reinterpret_cast<U*>(srow_s);
              int tail = rep;
              for (int t = rep; t != 0xFF && t < T; t = nexts[c * Sp + t]) {
                const int kt = act[i * Tp + t];
                const U bits = cells[kt];
                if (bits >= B::kBase) {
                  const int f = (int)(bits & (U)0xFF);
                  cur[i * Tp + t] = cur[i * Tp + f];
                  canon[i * Tp + t] = (uint8_t)f;
                } else {
                  cur[i * Tp + t] = B::val(bits);
                  canon[i * Tp + t] = (uint8_t)t;
                  cells[kt] = B::kBase | (U)t;
                  nextc[i * Tp + t] = 0xFF;
                  if (t != rep) nextc[i * Tp + tail] = (uint8_t)t;  // the representative transition is always a first writer
                  tail = t;
                }
              }
            }
            __syncwarp();
            QT best;
            unsigned bidx;
            hbm_scan_row<QT, false>(srow_s, a[2], qq, lpr, has, best, bidx);
            if (has && qq == 0) {
              if (bidx == kNone) {  // no untagged value above -inf: the first untagged column, if any, holds the maximum (-inf)
                const U* cells = reinterpret_cast<const U*>(srow_s);
                for (int col = 0; col < a[2]; ++col)
                  if (cells[col] < B::kBase) { bidx = (unsigned)col; break; }
              }
              bm[i * Sp + rep] = best;
              const int row = srow[c * Sp + rep];
              if (row < a[5]) Gc[a[4] + row] = bidx == kNone ? (uint8_t)0xFF : (uint8_t)bidx;
            }
            dp += dstep; ap += astep;
            if (ap >= na) { ap -= na; ++dp; }
            hbm_fence_proxy_async_shared();  // the tags (generic stores) before the next bulk copy into this slot
            __syncwarp();
            if (k + nb < nbatch) issue(b);
          }
        } else {
          // rows land in registers: 16-byte vector loads (L1 bypassed), all of a lane's chunks in flight at once, rows of the
          // batches ahead prefetched into L2.  The columns a transition of this batch writes are a 128-bit mask per row (plus
          // the padding columns), tested per element in the compare's predicate.
          constexpr int EPC = R::kEpc, LPR = R::kLpr, NS = R::kSlots;
          int dp = grp / na, ap = grp - dp * na;  // item of the batch being processed
          int df = dp, af = ap;                   // item of the batch being prefetched
          auto prefetch = [&]() {
            if (df < nd) {
              const int* a = agc + p.cls_agent[abeg + af] * kHbmAgc;
              const unsigned char* src = reinterpret_cast<const unsigned char*>(qg + a[0] + (size_t)srow[c * Sp + dl[c * Sp + df]] * a[1]);
              const int bytes = a[1] * (int)sizeof(QT);
              for (int o = 32 * qq; o < bytes; o += 32 * LPR) hbm_prefetch_l2(src + o);  // one 32-byte sector per prefetch
            }
            df += dstep; af += astep;
            if (af >= na) { af -= na; ++df; }
          };
          for (int k = 0; k < p.pf_dist && k < nbatch; ++k) prefetch();
          for (int k = 0; k < nbatch; ++k) {
            const bool has = dp < nd;
            const int i = p.cls_agent[abeg + ap];
            const int rep = has ? (int)dl[c * Sp + dp] : 0;
            const int* a = agc + i * kHbmAgc;
            const QT* row = qg + a[0] + (size_t)(has ? (int)srow[c * Sp + rep] : 0) * a[1];
            const int nch = has ? a[1] / EPC : 0;  // row_stride is a multiple of a chunk
            QT v[NS][EPC];
#pragma unroll
            for (int s = 0; s < NS; ++s) {
              if (s * LPR < p.nch_max) {  // uniform: no agent's row reaches this slot otherwise
                const int c4 = qq + s * LPR;
                if (c4 < nch) {
                  hbm_chunk<true>(row, c4, v[s]);
                } else {
#pragma unroll
                  for (int e2 = 0; e2 < EPC; ++e2) v[s][e2] = NegInf<QT>::v();
                }
              }
            }
            if (k + p.pf_dist < nbatch) prefetch();
            uint32_t* M = maskS + grp * 4;
            if (qq == 0) {
              if (has) {
                M[0] = (uint32_t)a[12]; M[1] = (uint32_t)a[13]; M[2] = (uint32_t)a[14]; M[3] = (uint32_t)a[15];
                // the row's transitions in ascending order: the first writer of a cell loads everybody's stale snapshot
                // (agents.py:67; an L2 hit behind the row loads above) and joins the row's list of rewritten cells
                int tail = rep;
                for (int t = rep; t != 0xFF && t < T; t = nexts[c * Sp + t]) {
                  const int kt = act[i * Tp + t];
                  const uint32_t bit = 1u << (kt & 31);
                  if (M[kt >> 5] & bit) {
                    int f = rep;
                    while (act[i * Tp + f] != kt) f = nextc[i * Tp + f];
                    cur[i * Tp + t] = cur[i * Tp + f];
                    canon[i * Tp + t] = (uint8_t)f;
                  } else {
                    M[kt >> 5] |= bit;
                    cur[i * Tp + t] = __ldcg(row + kt);
                    canon[i * Tp + t] = (uint8_t)t;
                    nextc[i * Tp + t] = 0xFF;
                    if (t != rep) nextc[i * Tp + tail] = (uint8_t)t;  // the representative transition is always a first writer
                    tail = t;
                  }
                }
              } else {
                M[0] = M[1] = M[2] = M[3] = 0xffffffffu;
              }
            }
            __syncwarp();
            uint32_t mq[4];
            {
              const uint4 m4 = *reinterpret_cast<const uint4*>(M);
              const int sh = qq * EPC;  // element (slot s, e) of this lane is column 8 s + sh + e: bit 8 (s % 4) + e of word s / 4 after the shift
              mq[0] = m4.x >> sh; mq[1] = m4.y >> sh; mq[2] = m4.z >> sh; mq[3] = m4.w >> sh;
            }
            QT b0 = NegInf<QT>::v(), b1 = NegInf<QT>::v();
            unsigned i0 = kNone, i1 = kNone;
#pragma unroll
            for (int s = 0; s < NS; ++s) {
              if (s * LPR < p.nch_max) {
#pragma unroll
                for (int e2 = 0; e2 < EPC; ++e2) {  // ascending columns, strict >: first maximal index (agents.py:88)
                  const bool free_cell = (mq[s >> 2] & (1u << (8 * (s & 3) + e2))) == 0u;
                  const unsigned col = (unsigned)((qq + s * LPR) * EPC + e2);
                  if (s & 1) {
                    if (free_cell && v[s][e2] > b1) { b1 = v[s][e2]; i1 = col; }
                  } else {
                    if (free_cell && v[s][e2] > b0) { b0 = v[s][e2]; i0 = col; }
                  }
                }
              }
            }
            if (b1 > b0 || (b1 == b0 && i1 < i0)) { b0 = b1; i0 = i1; }
#pragma unroll
            for (int off = 1; off < LPR; off <<= 1) {
              const QT ob = shfl_xor_t(b0, off);
              const unsigned oi = __shfl_xor_sync(kFull, i0, off);
              if (ob > b0 || (ob == b0 && oi < i0)) { b0 = ob; i0 = oi; }
            }
            if (has && qq == 0) {
              if (i0 == kNone) {  // no free cell above -inf: the first free column, if any, holds the maximum (-inf)
#pragma unroll
                for (int w = 3; w >= 0; --w)
                  if (~M[w]) i0 = (unsigned)(32 * w + __ffs((int)~M[w]) - 1);
              }
              bm[i * Sp + rep] = b0;
              const int rowi = srow[c * Sp + rep];
              if (rowi < a[5]) Gc[a[4] + rowi] = i0 == kNone ? (uint8_t)0xFF : (uint8_t)i0;
            }
            dp += dstep; ap += astep;
            if (ap >= na) { ap -= na; ++dp; }
            __syncwarp();  // the masks are rewritten by the next batch
          }
        }
      }
      __syncwarp();

      // ---- U4. lane = agent: logs in step order (trainer.py:65-66) and the batch in order (agents.py:68-76), all on chip.
      //      reward / max_steps (trainer.py:65) is divided lane-parallel, `seg` steps at a time.
      double rlog = 0.0, alog = 0.0;  // trainer.py:40-41
      {
        const uint8_t* rsc = rs + (my_cls < 0 ? 0 : my_cls) * Sp;
        const double alpha = is_agent ? hpw[lane * 5 + 0] : 0.0, gamma = is_agent ? hpw[lane * 5 + 1] : 0.0;
        const double oma = __dsub_rn(1.0, alpha);
        const uint8_t* actl = act + lane * Tp;
        for (int s0 = 0; s0 < T; s0 += seg) {
          const int sl = T - s0 < seg ? T - s0 : seg;
          for (int idx = lane; idx < n * sl; idx += 32) {
            const int i = idx / sl, tt = idx - i * sl;
            const double rew = __dmul_rn(P[s0 + tt + 1], lutAQ[agc[i * kHbmAgc + 3] + act[i * Tp + s0 + tt]]);  // environments.py:34
            rq[i * seg + tt] = __ddiv_rn(rew, Td);
          }
          __syncwarp();
          if (is_agent) {
            for (int t = s0; t < s0 + sl; ++t) {
              const int kt = actl[t];
              const double pn = P[t + 1], aqk = lutAQ[my_lut + kt];
              rlog = __dadd_rn(rlog, rq[lane * seg + t - s0]);
              alog = __dadd_rn(alog, lutXT[my_lut + kt]);
              if (t >= my_t0) {
                const int rep = rsc[t + 1];
                const int cn = canon[lane * Tp + t];
                const double oldv = (double)cur[lane * Tp + t];  // still the snapshot: only t itself or later writers change cur[t]
                QT m = bm[lane * Sp + rep];  // live row max (agents.py:71): untagged cells + current values of the rewritten ones
                for (int cc = rep < T ? rep : 0xFF; cc != 0xFF; cc = nextc[lane * Tp + cc]) {
                  const QT v = cur[lane * Tp + cc];
                  m = v > m ? v : m;
                }
                const double reward = __dmul_rn(pn, aqk);
                const double nv = __dadd_rn(__dmul_rn(oma, oldv), __dmul_rn(alpha, __dadd_rn(reward, __dmul_rn(gamma, (double)m))));  // agents.py:72-75
                cur[lane * Tp + cn] = (QT)nv;  // cur[t] itself stays the snapshot unless t is the cell's first writer
              }
            }
          }
          __syncwarp();
        }
      }

      // ---- U5. final values to HBM (the last write of the batch per cell, agents.py:75); exact greedy refresh of every
      //      touched row: untagged (max, argmax) merged with the final values of the row's rewritten cells
      for (int i = 0; i < n; ++i) {
        const int* a = agc + i * kHbmAgc;
        if (a[7] == 0) continue;
        const int c = a[8];
        const uint16_t* sr = srow + c * Sp;
        for (int t = a[6] + lane; t <= T; t += 32) {
          if (t < T && canon[i * Tp + t] == t) qg[a[0] + (int)sr[t] * a[1] + act[i * Tp + t]] = cur[i * Tp + t];
          const int row = sr[t];
          if (rs[c * Sp + t] == t && row < a[5]) {
            QT best = bm[i * Sp + t];
            const uint8_t ba = Gc[a[4] + row];  // first column of the untagged maximum, left there by U3
            unsigned bidx = ba == 0xFF ? kNone : (unsigned)ba;
            for (int cc = t < T ? t : 0xFF; cc != 0xFF; cc = nextc[i * Tp + cc]) {
              const QT v = cur[i * Tp + cc];
              const unsigned col = act[i * Tp + cc];
              if (v > best || (v == best && col < bidx)) { best = v; bidx = col; }
            }
            Gc[a[4] + row] = bidx == kNone ? (uint8_t)0 : (uint8_t)bidx;
          }
        }
      }
      if (kStaged) __threadfence();  // the table stores are performed before next episode's bulk copies read the rows (fence.proxy.async there)

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
    if (kDyn) {  // the run's next task may start: its state is in memory
      __threadfence_block();
      __syncwarp();
      if (lane == 0) qdone[iloc] = cidx + 1;
    }
  }
}

}  // namespace thrl
