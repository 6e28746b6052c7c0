/* thrl.h — C ABI of the B200-native th_rl training hot path.
 *
 * One shared library (libthrl.so, built from th_rl_b200/csrc/) exports the `thrl_*` entry
 * points declared here.  They replace, batched over R independent runs, the duck-typed Python
 * protocol the reference's trainer drives once per run (the reference has no FFI; these are the
 * symbols a ctypes binding of that protocol would bind):
 *
 *   thrl_qtable_scan        <- th_rl/trainer.py:45-70   (epoch/step loop)
 *                              th_rl/agents.py:80-89    (QTable.sample_action)
 *                              th_rl/agents.py:47-57    (QTable.encode / scale)
 *                              th_rl/environments.py:18-39 (NoisyPriceState.step)
 *                              th_rl/buffers.py:12,18-19,28-41 (ReplayBuffer append/replay/empty)
 *                              th_rl/agents.py:59-78    (QTable.train_net)
 *                              th_rl/trainer.py:40-41,65-66 (rewards_log / actions_log)
 *   thrl_qtable_scan_host   <- the same call with HOST buffers (H2D + scan + D2H inside)
 *   thrl_qtable_init        <- th_rl/agents.py:29,45 (table = 12.5/(1-gamma)+randn, counter = 0)
 *                              th_rl/environments.py:15-16,50-53 (reset: price ~ U(0,a))
 *   thrl_greedy_eval        <- th_rl/utils.py:27-47 (play_game) + th_rl/agents.py:91-92 (get_action)
 *   (agents of kind THRL_AGENT_REINFORCE inside thrl_qtable_scan)
 *                           <- th_rl/agents.py:119-194 (Reinforce: pi / sample_action / scale / train_net + Adam)
 *
 * Conventions: plain C types only; every pointer in ThrlScanArgs is a DEVICE pointer owned by
 * the caller unless the entry point's name ends in `_host`; calls are asynchronous on the caller's
 * cudaStream_t (passed as void*); the return value is 0 or a negative ThrlStatus, with a message in
 * thrl_last_error().  Device memory: the Q-table kernels allocate nothing.  Two paths use
 * library-owned scratch, both stream-ordered and cached (steady-state calls allocate nothing): the
 * lattice and interval-table kernels for MLP agents take their per-warp workspace from a per-device memory pool
 * (cudaMallocFromPoolAsync on the caller's stream, release threshold = keep), and the `_host` entry
 * point keeps a per-device staging arena (three chunk slots) between calls; thrl_release_device_memory()
 * returns both to the driver.
 * The reference raises AssertionError/IndexError for bad configs (trainer.py:21-23, agents.py:88);
 * here those become THRL_ERR_BAD_CONFIG before anything is launched.
 */
#ifndef THRL_H_
#define THRL_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define THRL_ABI_VERSION 5
#define THRL_MAX_AGENTS 16
#define THRL_MAX_ACTIONS 255 /* greedy-action cache is one byte per table row, 0xFF = not cached */

typedef enum ThrlStatus {
  THRL_OK = 0,
  THRL_ERR_BAD_CONFIG = -1,  /* reference: AssertionError / IndexError at construction or first step */
  THRL_ERR_BAD_ARGS = -2,    /* null / inconsistent pointers */
  THRL_ERR_UNSUPPORTED = -3, /* shape outside what the kernels were built for */
  THRL_ERR_CUDA = -4,        /* a CUDA runtime call failed; text in thrl_last_error() */
  THRL_ERR_NO_DEVICE = -5    /* no sm_100 device: there is no CPU fallback */
} ThrlStatus;

typedef enum ThrlDtype {
  THRL_F32 = 0, /* fp32 storage, update arithmetic in f64 with one final rounding (fast path) */
  THRL_F64 = 1  /* the reference's own dtype: bit-exact tables (agents.py:29) */
} ThrlDtype;

typedef enum ThrlAgentKind {
  THRL_AGENT_QTABLE = 0,   /* th_rl/agents.py:12-116 */
  THRL_AGENT_REINFORCE = 1,  /* th_rl/agents.py:119-219: MLP 1 -> hidden -> actions, policy gradient, Adam(lr) */
  THRL_AGENT_ACTORCRITIC = 2, /* th_rl/agents.py:222-330: as Reinforce plus value head fc_v (bias 1000), [N,N] advantage loss */
  THRL_AGENT_CAC = 3          /* th_rl/agents.py:333-442: continuous action sigmoid(N(mu, std)), heads fc_mu / fc_std / fc_v.
                                 Its actions are float32 values: action streams (replay_ra, trace_actions, buffers) carry their
                                 bit patterns; `actions` is unused (set it to 2) */
} ThrlAgentKind;

typedef enum ThrlRngMode {
  THRL_RNG_PHILOX = 0,        /* free running: Philox4x32-10 keyed (seed), counter (run, epoch, step, stream) */
  THRL_RNG_REPLAY_DRAWS = 1,  /* replay recorded exploration draws u / random action; greedy picked here */
  THRL_RNG_REPLAY_ACTIONS = 2 /* teacher forcing: every action is taken from replay_ra */
} ThrlRngMode;

/* One agent.  Field names and defaults follow QTable.__init__ (th_rl/agents.py:13-27). */
typedef struct ThrlAgentSpec {
  int32_t states;     /* table has states+1 rows (agents.py:29) */
  int32_t actions;    /* columns; 1..THRL_MAX_ACTIONS */
  int32_t min_memory; /* update fires when buffered transitions >= min_memory (agents.py:60) */
  int32_t capacity;   /* deque(maxlen=capacity) (buffers.py:12) */
  double action_lo, action_hi; /* action_range */
  double max_state;
  double gamma, alpha, eps_end, eps_step; /* used when ThrlScanArgs.hp == NULL */
  int64_t table_offset;                   /* element offset inside one run's slab; set by thrl_game_layout */
  /* ABI 2: MLP agents (kind != THRL_AGENT_QTABLE).  `states` is the input width (must be 1: the price), `capacity` /
   * `min_memory` keep their meaning, gamma is the return discount; alpha / eps_* / max_state are unused. */
  int32_t kind;       /* ThrlAgentKind */
  int32_t hidden;     /* width of fc1 (reference: 256) */
  double lr;          /* Adam learning rate (reference: 2e-4); betas (0.9, 0.999), eps 1e-8 */
  double entropy;     /* entropy coefficient c_e of loss + c_e * (-mean entropy) (agents.py:187-189, 298-300, 410-412; default 0) */
  int64_t mlp_offset; /* float offset of this agent's block inside one run's MLP slab; set by thrl_game_layout */
  /* ABI 3: elements between the starts of consecutive table rows; set by thrl_game_layout.  Equal to `actions` while one run's
   * tables are small enough to be staged in shared memory; games whose tables stay in HBM (sum_i (states_i+1)*actions_i*4 B >=
   * THRL_PAD_THRESHOLD_BYTES) get rows padded to a multiple of 4 elements, table offsets and run_stride likewise, so that every
   * row starts 16-byte aligned (fp32) and can be moved with one bulk copy / 16-byte vector loads.  Padding cells are never read
   * as values and never written by the scan; thrl_game_init zeroes them. */
  int32_t row_stride;
  int32_t reserved_;
} ThrlAgentSpec;

/* Layout of one MLP agent's block inside the run's MLP slab (floats / 32-bit words), P = 2*hidden + actions*hidden + actions
 * (+ hidden + 1 for ActorCritic; CAC: P = 5*hidden + 3):
 *   [0, P)        parameters in state_dict order: fc1.weight[hidden] fc1.bias[hidden] fc_pi.weight[actions][hidden] fc_pi.bias[actions]
 *                 (ActorCritic: then fc_v.weight[hidden] fc_v.bias[1]; CAC: fc1.weight fc1.bias fc_mu.weight[hidden] fc_mu.bias[1]
 *                 fc_std.weight[hidden] fc_std.bias[1] fc_v.weight[hidden] fc_v.bias[1])
 *   [P, 2P)       Adam exp_avg          [2P, 3P)  Adam exp_avg_sq
 *   [3P, 3P+4)    int32 header: Adam step count, buffered transitions, 2 reserved
 *   [3P+4, ...)   transition buffer, `mlp_buffer_len` entries of 3 words: state (f32), action (int32), reward (f32);
 *                 ActorCritic and CAC entries have a 4th word: new_state (f32); CAC's action word is a float32 */
#define THRL_MLP_HEADER_WORDS 4
/* one run's fp32 tables at or above this size are left in HBM by every kernel (a quarter of the 227 KB of shared memory an
 * sm_100 CTA can have: fewer than four runs would fit per SM); from here on the slab layout is padded (ThrlAgentSpec.row_stride) */
#define THRL_PAD_THRESHOLD_BYTES 58112

/* One game = n agents + NoisyPriceState kwargs (th_rl/environments.py:5-13). */
typedef struct ThrlGame {
  int32_t n_agents;
  int32_t max_steps;
  double a, b, noise_prob;
  ThrlAgentSpec agent[THRL_MAX_AGENTS];
  int64_t run_stride; /* elements per run slab = sum_i (states_i+1)*row_stride_i (rounded up to 4 when padded); set by thrl_game_layout */
  int32_t ring_len;   /* transitions a run may hold across an epoch boundary; set by thrl_game_layout */
  int32_t regular;    /* 1 if every agent's buffer is empty at every epoch boundary (min_memory <= max_steps) */
  int64_t mlp_stride; /* ABI 2: 32-bit words per run in the MLP slab (0 when all agents are QTables); set by thrl_game_layout */
  int32_t mlp_buffer_len[THRL_MAX_AGENTS]; /* transitions an MLP agent can hold; set by thrl_game_layout */
} ThrlGame;

/* Fixed-point scales of the cross-run statistics (exact, order-independent integer sums). */
#define THRL_STATS_K 4 /* sum r, sum r^2, sum x, sum x^2 of the per-epoch means, per (epoch, agent) */
#define THRL_STATS_SCALE_SUM 4294967296.0 /* 2^32 */
#define THRL_STATS_SCALE_SQ 16777216.0    /* 2^24 */

typedef struct ThrlScanArgs {
  const ThrlGame* game; /* HOST pointer */
  int64_t n_runs;       /* runs held by this call (this GPU's shard) */
  int64_t run_id0;      /* global id of local run 0: Philox counters use global ids => results do not depend on the sharding.
                           The id occupies one 32-bit counter word: run_id0 + n_runs <= 2^32 (THRL_ERR_BAD_ARGS beyond) */
  int32_t epoch_begin;  /* epochs [epoch_begin, epoch_end) are played; arrays below are indexed by e - epoch_begin */
  int32_t epoch_end;
  int32_t table_dtype; /* ThrlDtype */
  int32_t rng_mode;    /* ThrlRngMode */
  uint64_t seed;

  /* per-run state, read and written in place */
  void* q;           /* [R][run_stride] f32 or f64; agent i's table at table_offset_i, row-major [states_i+1][row_stride_i],
                        columns 0..actions_i-1 of every row used; 16-byte aligned */
  uint32_t* counter; /* [R][run_stride] visit counts (agents.py:45,76); may be NULL */
  double* eps;       /* [R][n] current epsilon */
  double* price;     /* [R] current state = last price (environments.py:36) */
  const double* hp;  /* NULL or [R][n][4] = alpha, gamma, eps_end, eps_step per run (hyper-parameter sweeps) */
  void* ring;        /* [R][thrl_ring_bytes(game)] opaque transition buffer carried between calls (zero-filled = empty buffers);
                        REQUIRED when game->regular == 0 (THRL_ERR_BAD_ARGS otherwise), ignored (may be NULL) for regular games */

  /* replay streams (rng_mode != PHILOX); [R][E][T][n] / [R][E][T] */
  const double* replay_u;     /* recorded random.uniform(0,1) (agents.py:81); REPLAY_DRAWS only */
  const int32_t* replay_ra;   /* REPLAY_DRAWS: recorded random.choice result, -1 where not drawn (agents.py:82);
                                 REPLAY_ACTIONS: the action to take */
  const double* replay_new_a; /* NULL (no noise: new_a = a) or recorded demand intercept per step (environments.py:28-31) */

  /* outputs, all optional */
  double* rewards_log; /* [n_log_runs][E][n] per-epoch mean reward (trainer.py:65) for runs 0..n_log_runs-1 */
  double* actions_log; /* [n_log_runs][E][n] per-epoch mean scaled action (trainer.py:66) */
  int64_t n_log_runs;
  int64_t* stats;         /* [E][n][THRL_STATS_K] fixed-point sums over the runs of this call, accumulated (+=) */
  int32_t* trace_actions; /* [R][E][T][n] chosen action indices */
  double* trace_rewards;  /* [R][E][T][n] */
  double* trace_prices;   /* [R][E][T] */
  /* ABI 2 */
  float* mlp; /* [R][game->mlp_stride] parameters, Adam state and transition buffers of the MLP agents; NULL iff mlp_stride == 0.
                 The host initialises parameters (torch's nn.Linear init) and zeroes the rest. */
} ThrlScanArgs;

/* Host-side helpers (no device needed). */
int thrl_abi_version(void);
const char* thrl_last_error(void);
/* Validates the game like the reference's constructors would, fills table_offset / run_stride / ring_len / regular. */
int thrl_game_layout(ThrlGame* game);
int64_t thrl_ring_bytes(const ThrlGame* game);

/* Device entry points. `stream` is a cudaStream_t. */
int thrl_qtable_scan(const ThrlScanArgs* args, void* stream);
/* Same arguments but every pointer in *args is a HOST pointer.  The run range is cut into chunks (whole multiples of the
 * kernel's resident-run count); chunk c+1's host->device copies and chunk c-1's device->host copies run on their own streams
 * while chunk c is in the kernel; three chunk slots of device memory are kept in a per-device arena between calls.  Copies
 * overlap the kernel when the host buffers are page-locked (cudaHostRegister / pinned allocations); pageable buffers work,
 * serialised by the driver.  Returns after everything has been copied back. */
int thrl_qtable_scan_host(const ThrlScanArgs* args, int device);
/* Frees the library-owned device scratch of the current device (lattice-kernel pool, `_host` staging arena). */
int thrl_release_device_memory(void);
/* Fills q (12.5/(1-gamma_i) + N(0,1)), counter (0), eps (eps0[i]), price (U(0,a)) for runs of this shard from
 * Philox(seed, global run id); eps0 is a HOST array of n doubles. */
int thrl_qtable_init(const ThrlGame* game, int64_t n_runs, int64_t run_id0, uint64_t seed, int32_t table_dtype,
                     const double* hp, const double* eps0, void* q, uint32_t* counter, double* eps, double* price,
                     void* stream);
/* ABI 2: thrl_qtable_init for games that may contain MLP agents: additionally fills the MLP slab (parameters ~ nn.Linear's
 * default U(-1/sqrt(fan_in), 1/sqrt(fan_in)), Adam state / header / buffer zero).  mlp may be NULL iff mlp_stride == 0. */
int thrl_game_init(const ThrlGame* game, int64_t n_runs, int64_t run_id0, uint64_t seed, int32_t table_dtype,
                   const double* hp, const double* eps0, void* q, uint32_t* counter, double* eps, double* price, float* mlp,
                   void* stream);
/* Greedy rollout (no exploration, no update): `iters` episodes per run, each starting from price0[r][it];
 * rewards/actions [R][iters*T][n] as utils.play_game returns them per run. */
int thrl_greedy_eval(const ThrlGame* game, int64_t n_runs, int32_t table_dtype, const void* q, int32_t iters,
                     const double* price0, double* rewards, double* actions, void* stream);
/* ABI 2: the same for games with MLP agents (get_action of Reinforce / ActorCritic = argmax of pi, of CAC = sigmoid(mu)); mlp is
 * the run slab of ThrlScanArgs.mlp (read only); q / mlp may be NULL when the game has no Q-tables / no MLP agents. */
int thrl_greedy_eval_mlp(const ThrlGame* game, int64_t n_runs, int32_t table_dtype, const void* q, const float* mlp,
                         int32_t iters, const double* price0, double* rewards, double* actions, void* stream);
/* ABI 5: the same with demand noise (environments.py:28-31 inside utils.play_game's env.step).  new_a [R][iters][T] holds the demand
 * intercept of every step as the environment would have drawn it -- a where no noise fires, the redrawn value in [0.7 a, a]
 * otherwise; the Python mirror draws it from numpy's global generator in the reference's order, so a seeded evaluation equals
 * utils.play_game's.  new_a == NULL plays the noise-free curve and is refused when game->noise_prob > 0 (as the two entry points
 * above, which have no such argument, always do). */
int thrl_greedy_eval_noise(const ThrlGame* game, int64_t n_runs, int32_t table_dtype, const void* q, const float* mlp,
                           int32_t iters, const double* price0, const double* new_a, double* rewards, double* actions, void* stream);
/* ABI 4: cross-run quantile statistics of the learning curve, th_rl/utils.py:132-145 (plot_learning_curve_conf: per run
 * pandas' ewm(halflife).mean() of the per-epoch mean rewards summed over the agents, then the median / quartiles over the runs).
 * rewards_log [n_runs][epochs][n_agents] is what thrl_qtable_scan wrote (device); ewm_num [n_runs] carries every run's EWM
 * numerator num_t = num_{t-1} * decay + x_t between calls (zero before the first epoch), decay = 1 - alpha = 0.5^(1/halflife);
 * den [epochs] (device) holds the EWM denominators den_t = den_{t-1} * decay + 1 of the epochs of this call (den_{-1} = 0) -- the
 * same for every run, so the host computes them.  The value
 * num_t / den_t of every run is counted into hist [epochs][n_bins] (+=; bin = floor((v - lo) / (hi - lo) * n_bins), clamped):
 * exact integer counts, so per-shard histograms add up (NCCL all-reduce). */
int thrl_curve_hist(const double* rewards_log, int64_t n_runs, int32_t epochs, int32_t n_agents, double decay,
                    const double* den, double* ewm_num, double lo, double hi, int32_t n_bins, int64_t* hist, void* stream);
/* Number of kernels this library has launched since load (bench.py reports it as gpu_launches). */
int64_t thrl_launch_count(void);
/* Name of the scan kernel the calling thread's latest thrl_qtable_scan / thrl_qtable_scan_host launched: "lut2", "lpc",
 * "hbm", "generic" (Q-table games), "pwl" (lattice kernel: Reinforce / ActorCritic agents on the noise-free demand curve),
 * "pwc" (interval-table kernel: MLP agents on a continuous price -- demand noise, CAC, pending QTable batches), "mixed"
 * (order-exact MLP kernel).  The dispatch is a function of the game, the inputs and THRL_KERNEL (= generic | lpc | mixed | pwc
 * force a kernel where it applies); tests use this to check it. */
const char* thrl_last_kernel(void);
/* Runs the calling thread's latest scan launch kept resident at once (persistent grid x runs per CTA).  A caller that cuts a
 * batch into several launches (engine.scan_from_host overlaps them with host copies) sizes the pieces in multiples of this,
 * so no launch ends on a partly filled round. */
int64_t thrl_last_wave_runs(void);

#ifdef __cplusplus
}
#endif
#endif /* THRL_H_ */
