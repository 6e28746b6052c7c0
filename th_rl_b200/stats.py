"""Cross-run quantile statistics of the learning curve without keeping every run's log.

th_rl/utils.py:132-145 (plot_learning_curve_conf) reads every run's log.csv, smooths the per-epoch mean rewards with
pandas' ewm(halflife=1000).mean(), sums the agents and draws the median and the quartiles over the runs.  With 10^5-10^6
runs the logs themselves are never materialised on the host: `CurveHistogram` keeps, per run, the EWM numerator on the
device and, per epoch, a fixed-bin histogram of the smoothed total reward over the runs (thrl_curve_hist, include/thrl.h).
Bin counts are exact integers: shards add up, chunked updates equal one call, and any quantile is read off to bin width.
"""
import ctypes as C
import math

import numpy


def ewm_decay(halflife):
    """pandas: alpha = 1 - exp(ln(0.5) / halflife); the per-epoch weight decay is 1 - alpha."""
    return math.exp(math.log(0.5) / float(halflife))


def ewm_denominators(decay, epoch_begin, epochs, den_before=None):
    """den_t = den_{t-1} * decay + 1 for t in [epoch_begin, epoch_begin + epochs), den_{-1} = 0 (adjust=True weights).
    `den_before` = den_{epoch_begin-1} when the caller carries it; otherwise it is recomputed from epoch 0."""
    d = 0.0
    if den_before is not None:
        d = float(den_before)
    else:
        for _ in range(int(epoch_begin)):
            d = d * decay + 1.0
    out = numpy.empty(int(epochs), numpy.float64)
    for t in range(int(epochs)):
        d = d * decay + 1.0
        out[t] = d
    return out


def default_range(config):
    """[0, 1.2 x the cartel's total reward a^2 / (4 b)] (environments.py:41-48 get_optimal): no per-epoch mean total reward of
    the Cournot game lies above the monopoly profit."""
    env = config["environment"]
    a, b = float(env.get("a", 10.0)), float(env.get("b", 1.0))
    return 0.0, 1.2 * a * a / (4.0 * b)


def quantiles_from_hist(hist, qs, lo, hi):
    """hist [E, n_bins] counts -> [E, len(qs)] quantiles (numpy.quantile's linear rule applied to bin centres)."""
    hist = numpy.asarray(hist, numpy.int64)
    E, nb = hist.shape
    width = (hi - lo) / nb
    centres = lo + (numpy.arange(nb) + 0.5) * width
    cum = numpy.cumsum(hist, axis=1)
    total = cum[:, -1]
    out = numpy.full((E, len(qs)), numpy.nan)
    for k, q in enumerate(qs):
        pos = q * (total - 1)              # fractional rank among the runs of the epoch, as numpy.quantile places it
        lo_rank, frac = numpy.floor(pos).astype(numpy.int64), pos - numpy.floor(pos)
        for e in range(E):
            if total[e] <= 0:
                continue
            i0 = int(numpy.searchsorted(cum[e], lo_rank[e] + 1))
            i1 = int(numpy.searchsorted(cum[e], min(lo_rank[e] + 2, total[e])))
            out[e, k] = centres[i0] + frac[e] * (centres[i1] - centres[i0])
    return out


class CurveHistogram:
    """Device-side accumulator for one shard of runs.  update() consumes the rewards_log a scan wrote for ALL runs of the shard."""

    def __init__(self, config, n_runs, device, halflife=1000.0, bins=1024, value_range=None):
        import torch
        self.lo, self.hi = value_range if value_range is not None else default_range(config)
        self.bins, self.halflife, self.decay = int(bins), float(halflife), ewm_decay(halflife)
        self.n_runs, self.device = int(n_runs), torch.device(device)
        self.num = torch.zeros((self.n_runs,), dtype=torch.float64, device=self.device)
        self.den_last = 0.0  # den_{epoch-1}
        self.epoch = 0

    def update(self, rewards_log, hist):
        """rewards_log [n_runs, E, n] (device, f64) of epochs [self.epoch, self.epoch + E); hist [E, bins] int64 (device), +="""
        import torch
        from ._lib import check, lib
        R, E, n = rewards_log.shape
        assert R == self.n_runs and rewards_log.dtype == torch.float64 and rewards_log.is_contiguous()
        assert hist.shape == (E, self.bins) and hist.dtype == torch.int64 and hist.is_contiguous()
        den = ewm_denominators(self.decay, self.epoch, E, den_before=self.den_last)
        den_d = torch.from_numpy(den).to(self.device)
        with torch.cuda.device(self.device):
            s = torch.cuda.current_stream()
            check(lib().thrl_curve_hist(C.c_void_p(rewards_log.data_ptr()), R, E, n, self.decay, C.c_void_p(den_d.data_ptr()),
                                        C.c_void_p(self.num.data_ptr()), self.lo, self.hi, self.bins,
                                        C.c_void_p(hist.data_ptr()), C.c_void_p(s.cuda_stream)))
            den_d.record_stream(s)
        self.den_last = float(den[-1]) if E else self.den_last
        self.epoch += E
        return hist
