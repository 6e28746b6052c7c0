"""Batched runs of one game on one GPU: device-resident state + calls into the C ABI.

`RunBatch` owns the per-run state as torch CUDA tensors (torch is only the allocator / stream provider here) and
advances all runs with one `thrl_qtable_scan` launch per call.  This is what `trainer.train_many` / `train_one`
drive; it replaces the reference's one-run-at-a-time Python loop (th_rl/trainer.py:45-70, th_rl/main.py:19-21).
"""
import ctypes as C

import numpy as np
import torch

from . import abi
from ._lib import check, game_layout, lib


def _dp(t):
    return None if t is None else C.c_void_p(t.data_ptr())


class ScanOutput:
    """What one scan call produced (device tensors unless stated)."""
    rewards_log = None   # [n_log_runs, E, n] f64 per-epoch mean reward (trainer.py:65)
    actions_log = None   # [n_log_runs, E, n] f64 per-epoch mean scaled action (trainer.py:66)
    stats = None         # [E, n, 4] int64 fixed-point cross-run sums (abi.THRL_STATS_*)
    trace_actions = None
    trace_rewards = None
    trace_prices = None
    ring = None          # scan_host only: pending transitions of a non-regular game at the end of the call (host array)


class RunBatch:
    def __init__(self, config, n_runs, *, device="cuda:0", dtype=torch.float32, run_id0=0, seed=0, hp=None):
        if not torch.cuda.is_available():
            raise RuntimeError("th_rl_b200 needs a CUDA (sm_100a) device; there is no CPU fallback")
        self.config = config
        self.game = game_layout(config)
        self.n_runs = int(n_runs)
        self.device = torch.device(device)
        assert dtype in (torch.float32, torch.float64)
        self.dtype = dtype
        self.table_dtype = abi.THRL_F64 if dtype == torch.float64 else abi.THRL_F32
        self.run_id0 = int(run_id0)
        self.seed = int(seed)
        self.epoch = 0
        self.wave = 0  # runs one scan launch keeps resident (learned from the first scan)
        self._zeros = torch.zeros  # allocator of the per-call output buffers (tests substitute a guard-banded arena)
        g, R, n = self.game, self.n_runs, self.game.n_agents
        with torch.cuda.device(self.device):
            self.q = torch.empty((R, g.run_stride), dtype=dtype, device=self.device)
            self.counter = torch.zeros((R, g.run_stride), dtype=torch.int32, device=self.device)  # u32 bit pattern
            self.eps = torch.empty((R, n), dtype=torch.float64, device=self.device)
            self.price = torch.empty((R,), dtype=torch.float64, device=self.device)
            self.hp = None if hp is None else torch.as_tensor(np.ascontiguousarray(hp, np.float64)).reshape(R, n, 4).to(self.device)
            self.mlp = None  # parameters, Adam state and transition buffers of the MLP agents (include/thrl.h)
            if g.mlp_stride:
                self.mlp = torch.zeros((R, g.mlp_stride), dtype=torch.float32, device=self.device)
            self.ring = None
            if not g.regular:
                rb = lib().thrl_ring_bytes(C.byref(g))
                self.ring = torch.zeros((R, rb), dtype=torch.uint8, device=self.device)

    # ---- initial state -------------------------------------------------------------------------------------------
    def init_device(self):
        """QTable.__init__ / environment.reset() for every run, on the device, from Philox(seed, global run id)."""
        eps0 = (C.c_double * self.game.n_agents)(*abi.eps0_from_config(self.config))
        with torch.cuda.device(self.device):
            check(lib().thrl_game_init(C.byref(self.game), self.n_runs, self.run_id0, self.seed, self.table_dtype,
                                       _dp(self.hp), eps0, _dp(self.q), _dp(self.counter), _dp(self.eps),
                                       _dp(self.price), _dp(self.mlp), C.c_void_p(torch.cuda.current_stream().cuda_stream)))
        self.epoch = 0
        return self

    def load_state(self, q, eps, price, counter=None, non_blocking=False, mlp=None):
        """Initial state provided by the caller as host arrays / tensors (e.g. tables drawn by numpy like the reference)."""
        def put(dst, src, dt):
            src = (src if torch.is_tensor(src) else torch.as_tensor(np.asarray(src))).reshape(dst.shape)
            if src.dtype != dt:
                src = src.to(dt)
            dst.copy_(src, non_blocking=non_blocking)
        put(self.q, q, self.dtype)
        put(self.eps, eps, torch.float64)
        put(self.price, price, torch.float64)
        if self.mlp is not None:
            if mlp is None:
                raise ValueError("games with MLP agents need the mlp slab")
            put(self.mlp, mlp, torch.float32)
        if counter is not None:
            put(self.counter, torch.as_tensor(np.asarray(counter).astype(np.uint32).view(np.int32)) if not torch.is_tensor(counter) else counter, torch.int32)
        else:
            self.counter.zero_()
        self.epoch = 0
        return self

    # ---- the hot path --------------------------------------------------------------------------------------------
    def scan(self, epochs, *, rng_mode=abi.THRL_RNG_PHILOX, replay_u=None, replay_ra=None, replay_new_a=None,
             n_log_runs=0, stats=False, trace=False, stream=None, run_range=None, advance=True):
        """Play `epochs` more epochs for every run (or runs [b, e) with run_range=(b, e)) with one kernel launch.
        Asynchronous on the current stream.  `stats` may be True (fresh buffer) or an int64 [E, n, 4] tensor to add into."""
        g, n, T, E = self.game, self.game.n_agents, self.game.max_steps, int(epochs)
        rb, re_ = (0, self.n_runs) if run_range is None else (int(run_range[0]), int(run_range[1]))
        R = re_ - rb
        dev = self.device
        out = ScanOutput()
        keep = []

        def dev_in(a, dt, shape):
            if a is None:
                return None
            t = torch.as_tensor(a)
            if t.dtype != dt:
                t = t.to(dt)
            t = t.reshape(shape).contiguous().to(dev, non_blocking=True)
            keep.append(t)
            return t

        with torch.cuda.device(dev):
            ru = dev_in(replay_u, torch.float64, (R, E, T, n))
            rra = dev_in(replay_ra, torch.int32, (R, E, T, n))
            rna = dev_in(replay_new_a, torch.float64, (R, E, T))
            if n_log_runs:
                out.rewards_log = self._zeros((n_log_runs, E, n), dtype=torch.float64, device=dev)
                out.actions_log = self._zeros((n_log_runs, E, n), dtype=torch.float64, device=dev)
            if torch.is_tensor(stats):
                out.stats = stats
            elif stats:
                out.stats = self._zeros((E, n, abi.THRL_STATS_K), dtype=torch.int64, device=dev)
            if trace:
                out.trace_actions = self._zeros((R, E, T, n), dtype=torch.int32, device=dev)
                out.trace_rewards = self._zeros((R, E, T, n), dtype=torch.float64, device=dev)
                out.trace_prices = self._zeros((R, E, T), dtype=torch.float64, device=dev)
            a = abi.ThrlScanArgs()
            a.game = C.pointer(g)
            a.n_runs, a.run_id0 = R, self.run_id0 + rb
            a.epoch_begin, a.epoch_end = self.epoch, self.epoch + E
            a.table_dtype, a.rng_mode, a.seed = self.table_dtype, rng_mode, self.seed
            a.q, a.eps, a.price = _dp(self.q[rb:re_]), _dp(self.eps[rb:re_]), _dp(self.price[rb:re_])
            a.counter = None if self.counter is None else _dp(self.counter[rb:re_])  # NULL: visit counts are not kept
            a.hp = None if self.hp is None else _dp(self.hp[rb:re_])
            a.ring = None if self.ring is None else _dp(self.ring[rb:re_])
            a.mlp = None if self.mlp is None else _dp(self.mlp[rb:re_])
            a.replay_u, a.replay_ra, a.replay_new_a = _dp(ru), _dp(rra), _dp(rna)
            a.rewards_log, a.actions_log, a.n_log_runs = _dp(out.rewards_log), _dp(out.actions_log), int(n_log_runs)
            a.stats = _dp(out.stats)
            a.trace_actions, a.trace_rewards, a.trace_prices = _dp(out.trace_actions), _dp(out.trace_rewards), _dp(out.trace_prices)
            s = stream if stream is not None else torch.cuda.current_stream()
            check(lib().thrl_qtable_scan(C.byref(a), C.c_void_p(s.cuda_stream)))
            self.wave = max(self.wave, int(lib().thrl_last_wave_runs()))  # resident runs per launch (see scan_from_host)
            for t in keep:  # inputs must outlive the asynchronous kernel
                t.record_stream(s)
        if advance:
            self.epoch += E
        return out

    # ---- results -------------------------------------------------------------------------------------------------
    def tables(self, run=None):
        """Per-agent tables [R, states+1, actions] (views of the packed slab; rows may be padded, include/thrl.h)."""
        q = self.q if run is None else self.q[run:run + 1]
        return [abi.table_view(q, s) if s.kind == abi.THRL_AGENT_QTABLE else None
                for s in (self.game.agent[i] for i in range(self.game.n_agents))]

    def counters(self, run=None):
        c = self.counter if run is None else self.counter[run:run + 1]
        return [abi.table_view(c, s) if s.kind == abi.THRL_AGENT_QTABLE else None
                for s in (self.game.agent[i] for i in range(self.game.n_agents))]

    def mlp_state_dicts(self, run=0):
        """Per agent: parameters of `run` in torch state_dict names/shapes (None for QTable agents)."""
        out = []
        for s in (self.game.agent[i] for i in range(self.game.n_agents)):
            if s.kind == abi.THRL_AGENT_QTABLE:
                out.append(None)
                continue
            p = self.mlp[run, s.mlp_offset:s.mlp_offset + abi.mlp_param_count(s)].cpu()
            d, o = {}, 0
            for k, shp in abi.mlp_param_shapes(s).items():
                cnt = int(np.prod(shp))
                d[k] = p[o:o + cnt].reshape(shp).clone()
                o += cnt
            out.append(d)
        return out

    def greedy_eval(self, price0, new_a=None):
        """utils.play_game for every run: price0 [R, iters] -> (actions, rewards) [R, iters*T, n] f64 device tensors.
        Games with demand noise (environments.py:28-31) need the demand intercept of every step, new_a [R, iters, T]; when it
        is omitted it is drawn here from numpy's global generator exactly as the environment does inside play_game -- per step
        one uniform, and a second one only when the first fell below noise_prob -- run after run, episode after episode, so a
        seeded evaluation of one run equals the reference's."""
        g, R, n, T = self.game, self.n_runs, self.game.n_agents, self.game.max_steps
        p0 = torch.as_tensor(price0, dtype=torch.float64).reshape(R, -1).contiguous().to(self.device)
        iters = p0.shape[1]
        if new_a is None and g.noise_prob > 0:
            new_a = draw_demand_intercepts(g.a, g.noise_prob, R * iters * T).reshape(R, iters, T)
        na = None if new_a is None else torch.as_tensor(np.ascontiguousarray(new_a, np.float64)).reshape(R, iters, T).to(self.device)
        rewards = self._zeros((R, iters * T, n), dtype=torch.float64, device=self.device)
        actions = self._zeros((R, iters * T, n), dtype=torch.float64, device=self.device)
        with torch.cuda.device(self.device):
            check(lib().thrl_greedy_eval_noise(C.byref(g), R, self.table_dtype, _dp(self.q), _dp(self.mlp), iters, _dp(p0), _dp(na),
                                               _dp(rewards), _dp(actions), C.c_void_p(torch.cuda.current_stream().cuda_stream)))
        return actions, rewards


def draw_demand_intercepts(a, noise_prob, steps):
    """The demand intercepts of `steps` consecutive environment steps, drawn from numpy's global generator in the order
    NoisyPriceState.step consumes it (environments.py:28-31): u = uniform(); new_a = uniform(0.7 a, a) if u < noise_prob else a."""
    out = np.full(int(steps), float(a))
    for t in range(int(steps)):
        if np.random.uniform() < noise_prob:
            out[t] = np.random.uniform(a * 0.7, a)
    return out


def save_checkpoint(batch, path, extra=None, extra_arrays=None):
    """Writes the batch's whole state (tables, counters, epsilon, price, hyper-parameters, MLP slab, transition rings) and
    its position (epoch, seed, global run ids) as one packed file (th_rl_b200/checkpoint.py)."""
    from . import checkpoint
    arrays = {"q": batch.q.cpu().numpy(), "counter": batch.counter.cpu().numpy().view(np.uint32),
              "eps": batch.eps.cpu().numpy(), "price": batch.price.cpu().numpy()}
    for name in ("hp", "mlp", "ring"):
        t = getattr(batch, name)
        if t is not None:
            arrays[name] = t.cpu().numpy()
    for name, a in (extra_arrays or {}).items():  # caller's arrays (e.g. stats.CurveHistogram.num) travel as "x_<name>"
        arrays["x_" + name] = np.ascontiguousarray(a)
    header = dict(abi_version=abi.THRL_ABI_VERSION, config=batch.config, n_runs=batch.n_runs, run_id0=batch.run_id0,
                  seed=batch.seed, epoch=batch.epoch, table_dtype="f64" if batch.dtype == torch.float64 else "f32",
                  run_stride=int(batch.game.run_stride), mlp_stride=int(batch.game.mlp_stride), extra=extra or {})
    return checkpoint.write_pack(path, header, arrays)


def load_checkpoint(path, device="cuda:0"):
    """-> (RunBatch positioned where the checkpoint was taken, the header's `extra` dict, the caller's extra arrays).  Scanning on from it gives
    bit-identical results to a run that was never interrupted (Philox streams are keyed by global run id and epoch)."""
    from . import checkpoint
    hdr, arr = checkpoint.read_pack(path)
    dtype = torch.float64 if hdr["table_dtype"] == "f64" else torch.float32
    b = RunBatch(hdr["config"], hdr["n_runs"], device=device, dtype=dtype, run_id0=hdr["run_id0"], seed=hdr["seed"],
                 hp=arr.get("hp"))
    if int(b.game.run_stride) != hdr["run_stride"] or int(b.game.mlp_stride) != hdr["mlp_stride"]:
        raise ValueError("%s was written with another slab layout (run_stride %d / mlp_stride %d, this library: %d / %d)"
                         % (path, hdr["run_stride"], hdr["mlp_stride"], b.game.run_stride, b.game.mlp_stride))

    def put(dst, src):
        dst.copy_(torch.from_numpy(np.ascontiguousarray(src)).reshape(dst.shape))
    put(b.q, arr["q"])
    put(b.counter, np.asarray(arr["counter"]).view(np.int32))
    put(b.eps, arr["eps"])
    put(b.price, arr["price"])
    if b.mlp is not None:
        put(b.mlp, arr["mlp"])
    if b.ring is not None:
        put(b.ring, arr["ring"])
    b.epoch = int(hdr["epoch"])
    return b, hdr.get("extra", {}), {k[2:]: np.array(v) for k, v in arr.items() if k.startswith("x_")}


class HostState:
    """Per-run state in pinned host memory (what a host-side caller owns between calls)."""

    def __init__(self, q, counter, eps, price, mlp=None):
        self.q, self.counter, self.eps, self.price, self.mlp = q, counter, eps, price, mlp

    @classmethod
    def from_batch(cls, batch):
        def pin(t):
            h = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
            h.copy_(t)
            return h
        return cls(pin(batch.q), pin(batch.counter), pin(batch.eps), pin(batch.price),
                   None if batch.mlp is None else pin(batch.mlp))

    def nbytes(self):
        return sum(t.numel() * t.element_size() for t in (self.q, self.counter, self.eps, self.price, self.mlp) if t is not None)


def chunk_bounds(R, n_chunks, wave=0):
    """Run-range boundaries of the launches scan_from_host cuts a batch of R runs into.

    The first upload and the last download are the only copies no kernel hides: those two chunks get half the weight.  Once a
    scan has told us how many runs one launch keeps resident (`wave`, thrl_last_wave_runs) and the chunks span several rounds
    of the persistent grid, chunks are whole multiples of that, so no launch ends on a partly filled round."""
    n_chunks = max(1, min(int(n_chunks), R))
    w = [2] * n_chunks
    if n_chunks >= 3:
        w[0] = w[-1] = 1
    unit = wave if wave and -(-R // wave) >= 2 * n_chunks else 1
    units = -(-R // unit)
    acc, tot, bounds = 0, sum(w), [0]
    for x in w:
        acc += x
        bounds.append(min(R, (units * acc // tot) * unit))
    bounds[-1] = R
    return bounds


def scan_from_host(batch, host, epochs, n_chunks=12):
    """Host-resident state -> `epochs` more epochs -> host-resident state, plus the cross-run statistics.

    The run range is cut into n_chunks; chunk c+1's host->device copy and chunk c-1's device->host copy run on their
    own streams while chunk c is in the kernel.  Returns (stats device tensor [E, n, 4], h2d bytes, d2h bytes).
    """
    dev, R, E, n = batch.device, batch.n_runs, int(epochs), batch.game.n_agents
    n_chunks = max(1, min(int(n_chunks), R))
    bounds = chunk_bounds(R, n_chunks, batch.wave)
    with torch.cuda.device(dev):
        if not hasattr(batch, "_streams"):
            batch._streams = (torch.cuda.Stream(dev), torch.cuda.Stream(dev))
        up, down = batch._streams
        cs = torch.cuda.current_stream()
        stats = torch.zeros((E, n, abi.THRL_STATS_K), dtype=torch.int64, device=dev)
        up.wait_stream(cs)
        pairs = [(batch.q, host.q), (batch.counter, host.counter), (batch.eps, host.eps), (batch.price, host.price)]
        if batch.mlp is not None:
            pairs.append((batch.mlp, host.mlp))
        for c in range(n_chunks):
            b, e = bounds[c], bounds[c + 1]
            with torch.cuda.stream(up):
                for d, h in pairs:
                    d[b:e].copy_(h[b:e], non_blocking=True)
                ev_up = up.record_event()
            cs.wait_event(ev_up)
            batch.scan(E, stats=stats, run_range=(b, e), advance=False)
            ev_k = cs.record_event()
            down.wait_event(ev_k)
            with torch.cuda.stream(down):
                for d, h in pairs:
                    h[b:e].copy_(d[b:e], non_blocking=True)
        cs.wait_stream(down)
    batch.epoch += E
    nb = host.nbytes()
    return stats, nb, nb


def scan_host(config, q, eps, price, epochs, *, counter=None, hp=None, rng_mode=abi.THRL_RNG_PHILOX, seed=0, run_id0=0,
              epoch_begin=0, replay_u=None, replay_ra=None, replay_new_a=None, n_log_runs=0, stats=False, trace=False,
              device=0, mlp=None, ring=None):
    """thrl_qtable_scan_host: HOST buffers in and out, all copies inside the call, chunked and overlapped with the kernel by
    the library (the reference-facing boundary; bench.py's e2e leg times exactly this call on page-locked buffers).
    q / eps / price (/ counter / mlp / ring) are numpy arrays updated IN PLACE.  Games that are not regular (an agent's
    min_memory exceeds max_steps) carry their pending transitions in `ring` (uint8 [R, thrl_ring_bytes], zero = empty):
    pass the same array to consecutive calls; when omitted the call starts from empty buffers and the result's `ring`
    holds what is pending at its end."""
    g = game_layout(config)
    n, T, E = g.n_agents, g.max_steps, int(epochs)
    R = q.shape[0]
    assert q.flags.c_contiguous and q.dtype in (np.float32, np.float64) and q.shape == (R, g.run_stride)
    if g.mlp_stride:
        assert mlp is not None and mlp.dtype == np.float32 and mlp.flags.c_contiguous and mlp.shape == (R, g.mlp_stride)
    assert eps.dtype == np.float64 and price.dtype == np.float64 and eps.flags.c_contiguous and price.flags.c_contiguous
    out = ScanOutput()
    keep = []

    def inp(a, dt, shape):
        if a is None:
            return None
        a = np.ascontiguousarray(a, dtype=dt).reshape(shape)
        keep.append(a)
        return a.ctypes.data_as(C.c_void_p)

    def outp(a):
        return None if a is None else a.ctypes.data_as(C.c_void_p)

    if n_log_runs:
        out.rewards_log = np.zeros((n_log_runs, E, n), np.float64)
        out.actions_log = np.zeros((n_log_runs, E, n), np.float64)
    if stats:
        out.stats = np.zeros((E, n, abi.THRL_STATS_K), np.int64)
    if trace:
        out.trace_actions = np.zeros((R, E, T, n), np.int32)
        out.trace_rewards = np.zeros((R, E, T, n), np.float64)
        out.trace_prices = np.zeros((R, E, T), np.float64)
    a = abi.ThrlScanArgs()
    a.game = C.pointer(g)
    a.n_runs, a.run_id0 = R, run_id0
    a.epoch_begin, a.epoch_end = epoch_begin, epoch_begin + E
    a.table_dtype = abi.THRL_F64 if q.dtype == np.float64 else abi.THRL_F32
    a.rng_mode, a.seed = rng_mode, seed
    a.q, a.eps, a.price = outp(q), outp(eps), outp(price)
    if counter is not None:
        assert counter.dtype == np.uint32 and counter.shape == q.shape and counter.flags.c_contiguous
        a.counter = outp(counter)
    a.hp = inp(hp, np.float64, (R, n, 4))
    a.mlp = outp(mlp) if g.mlp_stride else None
    if not g.regular:
        rb = int(lib().thrl_ring_bytes(C.byref(g)))
        if ring is None:
            ring = np.zeros((R, rb), np.uint8)
        assert ring.dtype == np.uint8 and ring.shape == (R, rb) and ring.flags.c_contiguous
        a.ring = outp(ring)
        out.ring = ring
    a.replay_u = inp(replay_u, np.float64, (R, E, T, n))
    a.replay_ra = inp(replay_ra, np.int32, (R, E, T, n))
    a.replay_new_a = inp(replay_new_a, np.float64, (R, E, T))
    a.rewards_log, a.actions_log, a.n_log_runs = outp(out.rewards_log), outp(out.actions_log), int(n_log_runs)
    a.stats = outp(out.stats)
    a.trace_actions, a.trace_rewards, a.trace_prices = outp(out.trace_actions), outp(out.trace_rewards), outp(out.trace_prices)
    check(lib().thrl_qtable_scan_host(C.byref(a), int(device)))
    return out
