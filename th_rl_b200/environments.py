"""NoisyPriceState — host-side mirror of th_rl/environments.py:4-53.

Constructor keywords, defaults, attributes, `reset`, `encode`, `scale_actions`, `sample_state` and `get_optimal` follow the
reference (same draws from numpy's global generator).  `step` is fused into the device scan and raises on the host.
"""
import numpy


class NoisyPriceState:
    def __init__(self, nplayers, action_range=[0, 1], a=10, b=1, max_steps=1, noise_prob=0.05, **kwargs):
        self.nplayers = nplayers
        self.action_range = action_range
        self.b = b
        self.a = a
        self.max_steps = max_steps
        self.state = self.sample_state()  # environments.py:11 — one numpy uniform draw, like the reference
        self.episode = 0
        self.noise_prob = noise_prob

    def sample_state(self):
        return numpy.random.uniform(0, self.a)

    def encode(self):
        return numpy.atleast_1d(self.state)

    def scale_actions(self, actions):
        return [self.a / self.b * a for a in actions]

    def step(self, actions):
        raise NotImplementedError("NoisyPriceState.step is fused into the device scan (thrl_qtable_scan / "
                                  "thrl_greedy_eval); th_rl_b200 has no host implementation of the hot path")

    def get_optimal(self):
        """Total reward per step at the Cournot-Nash and at the cartel (joint monopoly) quantities (environments.py:41-48, used
        by the plots as reference levels): with linear demand p = a - b Q every one of n symmetric players supplies
        a / (b (n + 1)) at Nash and a / (2 b n) in the cartel, so the totals are n (a / (n + 1))^2 / b and a^2 / (4 b)."""
        n = self.nplayers
        return n * (self.a / (n + 1)) ** 2 / self.b, self.a ** 2 / (4 * self.b)

    def reset(self):
        self.episode = 0
        self.state = self.sample_state()
        return self.encode()


ENVIRONMENTS = {"NoisyPriceState": NoisyPriceState}
