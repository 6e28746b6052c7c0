"""QTable — host-side mirror of th_rl/agents.py:12-116 for the B200 path.

Same constructor keywords, defaults and attributes as the reference class, so `create_game`, saved artefacts and the
reference's analysis code (`utils.py` reads .table/.counter/.states/.actions/.action_space/.action_range/.max_state and
calls load/get_action/scale) keep working.  What differs: acting and learning do not happen per call on the host.
`sample_action` / `train_net` / `memory.append` are fused into the device scan (csrc/thrl_scan_*.cuh); calling them on the
host raises, because this package has no CPU implementation of the hot path.
"""
import numpy

from .buffers import ReplayBuffer  # noqa: F401  (the reference resolves `buffer` by name)

_FUSED = ("%s is fused into the device scan (th_rl_b200.trainer.train_one / train_many -> thrl_qtable_scan); "
          "th_rl_b200 has no per-step host implementation of the hot path")


class QTable:
    def __init__(self, states=16, actions=4, action_range=[0, 1], gamma=0.99, buffer="ReplayBuffer", capacity=500,
                 max_state=10, alpha=0.1, eps_end=2e-2, epsilon=0.5, eps_step=5e-4, min_memory=100, **kwargs):
        # agents.py:29 — same draw from numpy's global generator, so a seeded script starts from the same table
        self.table = 12.5 / (1 - gamma) + numpy.random.randn(states + 1, actions)
        self.gamma = gamma
        self.alpha = alpha
        self.action_space = numpy.arange(0, actions)
        self.action_range = action_range
        self.actions = actions
        self.epsilon = epsilon
        self.eps_step = eps_step
        self.eps_end = eps_end
        self.states = states
        self.max_state = max_state
        self.min_memory = min_memory
        self.capacity = capacity
        if buffer != "ReplayBuffer":
            raise ValueError("unknown buffer %r (the reference only ships ReplayBuffer)" % (buffer,))
        self.memory = ReplayBuffer(capacity, None)
        self.counter = 0 * self.table

    # agents.py:47-49
    def encode(self, state):
        return numpy.round(state / self.max_state * self.states).astype("int64")

    # agents.py:51-57
    def scale(self, actions):
        return actions / (self.actions - 1.0) * (self.action_range[1] - self.action_range[0]) + self.action_range[0]

    # agents.py:91-92 (evaluation helper used by utils.play_game; batched on the device by RunBatch.greedy_eval)
    def get_action(self, state):
        return numpy.argmax(self.table[self.encode(state)])

    def sample_action(self, state):
        raise NotImplementedError(_FUSED % "QTable.sample_action")

    def train_net(self):
        raise NotImplementedError(_FUSED % "QTable.train_net")

    # agents.py:110-116
    def save(self, loc):
        numpy.save(loc, self.table)
        numpy.save(loc + "_counter", self.counter)

    def load(self, loc):
        self.table = numpy.load(loc + ".npy")
        self.counter = numpy.load(loc + "_counter.npy")


class Reinforce:
    """Host-side mirror of th_rl/agents.py:119-219 (MLP 1 -> 256 -> actions, policy gradient, Adam lr 2e-4).

    The constructor builds the same two nn.Linear layers in the same order, so a script that seeds torch gets the same
    initial weights as the reference; `save` / `load` use torch.save(state_dict) with the reference's key names.  Acting and
    training are fused into the device scan (csrc/thrl_scan_mixed.cuh)."""

    def __init__(self, states=4, actions=2, action_range=[0, 1], gamma=0.98, buffer="ReplayBuffer", capacity=50000,
                 min_memory=1000, entropy=0, **kwargs):
        import torch.nn as nn
        self.gamma = gamma
        self.action_range = action_range
        self.actions = actions
        self.states = states
        self.fc1 = nn.Linear(states, 256)       # agents.py:137
        self.fc_pi = nn.Linear(256, actions)    # agents.py:138
        if buffer != "ReplayBuffer":
            raise ValueError("unknown buffer %r (the reference only ships ReplayBuffer)" % (buffer,))
        self.memory = ReplayBuffer(capacity, None)
        self.capacity = capacity
        self.min_memory = min_memory
        self.entropy = entropy

    def state_dict(self):
        return {"fc1.weight": self.fc1.weight.detach().clone(), "fc1.bias": self.fc1.bias.detach().clone(),
                "fc_pi.weight": self.fc_pi.weight.detach().clone(), "fc_pi.bias": self.fc_pi.bias.detach().clone()}

    def load_state_dict(self, sd):
        import torch
        with torch.no_grad():
            self.fc1.weight.copy_(torch.as_tensor(sd["fc1.weight"]).reshape(self.fc1.weight.shape))
            self.fc1.bias.copy_(torch.as_tensor(sd["fc1.bias"]))
            self.fc_pi.weight.copy_(torch.as_tensor(sd["fc_pi.weight"]))
            self.fc_pi.bias.copy_(torch.as_tensor(sd["fc_pi.bias"]))

    # agents.py:154-158 (note: / actions, not / (actions - 1))
    def scale(self, action):
        return action / self.actions * (self.action_range[1] - self.action_range[0]) + self.action_range[0]

    def sample_action(self, state):
        raise NotImplementedError(_FUSED % "Reinforce.sample_action")

    def train_net(self):
        raise NotImplementedError(_FUSED % "Reinforce.train_net")

    # agents.py:215-219
    def save(self, loc):
        import torch
        torch.save(self.state_dict(), loc)

    def load(self, loc):
        import torch
        self.load_state_dict(torch.load(loc))


class ActorCritic(Reinforce):
    """Host-side mirror of th_rl/agents.py:222-330: Reinforce's network plus the value head fc_v (bias filled with 1000.0,
    agents.py:244).  Constructed in the reference's order (fc1, fc_pi, fc_v) so seeded torch gives the same initial weights."""

    def __init__(self, states=4, actions=2, action_range=[0, 1], gamma=0.98, buffer="ReplayBuffer", capacity=50000,
                 min_memory=1000, entropy=0, **kwargs):
        import torch.nn as nn
        super().__init__(states=states, actions=actions, action_range=action_range, gamma=gamma, buffer=buffer,
                         capacity=capacity, min_memory=min_memory, entropy=entropy)
        self.fc_v = nn.Linear(256, 1)           # agents.py:243
        self.fc_v.bias.data.fill_(1000.0)       # agents.py:244

    def state_dict(self):
        sd = super().state_dict()
        sd["fc_v.weight"] = self.fc_v.weight.detach().clone()
        sd["fc_v.bias"] = self.fc_v.bias.detach().clone()
        return sd

    def load_state_dict(self, sd):
        import torch
        super().load_state_dict(sd)
        with torch.no_grad():
            self.fc_v.weight.copy_(torch.as_tensor(sd["fc_v.weight"]).reshape(self.fc_v.weight.shape))
            self.fc_v.bias.copy_(torch.as_tensor(sd["fc_v.bias"]).reshape(self.fc_v.bias.shape))

    def sample_action(self, state):
        raise NotImplementedError(_FUSED % "ActorCritic.sample_action")

    def train_net(self):
        raise NotImplementedError(_FUSED % "ActorCritic.train_net")


class CAC:
    """Host-side mirror of th_rl/agents.py:333-442 (continuous actor-critic: heads fc_mu, fc_std, fc_v on a 1 -> 256 trunk;
    action = sigmoid(Normal(4 tanh(mu), softplus(std)).sample())).  Layers are built in the reference's order."""

    def __init__(self, states=4, action_range=[0, 1], gamma=0.98, buffer="ReplayBuffer", capacity=50000, min_memory=1000,
                 entropy=0, **kwargs):
        import torch.nn as nn
        self.gamma = gamma
        self.action_range = action_range
        self.states = states
        self.fc1 = nn.Linear(states, 256)   # agents.py:349-352
        self.fc_mu = nn.Linear(256, 1)
        self.fc_std = nn.Linear(256, 1)
        self.fc_v = nn.Linear(256, 1)
        if buffer != "ReplayBuffer":
            raise ValueError("unknown buffer %r (the reference only ships ReplayBuffer)" % (buffer,))
        self.memory = ReplayBuffer(capacity, None)
        self.capacity = capacity
        self.min_memory = min_memory
        self.entropy = entropy

    _LAYERS = ("fc1", "fc_mu", "fc_std", "fc_v")

    def state_dict(self):
        sd = {}
        for nm in self._LAYERS:
            layer = getattr(self, nm)
            sd[nm + ".weight"] = layer.weight.detach().clone()
            sd[nm + ".bias"] = layer.bias.detach().clone()
        return sd

    def load_state_dict(self, sd):
        import torch
        with torch.no_grad():
            for nm in self._LAYERS:
                layer = getattr(self, nm)
                layer.weight.copy_(torch.as_tensor(sd[nm + ".weight"]).reshape(layer.weight.shape))
                layer.bias.copy_(torch.as_tensor(sd[nm + ".bias"]).reshape(layer.bias.shape))

    # agents.py:368-372
    def scale(self, action):
        return action * (self.action_range[1] - self.action_range[0]) + self.action_range[0]

    def sample_action(self, state):
        raise NotImplementedError(_FUSED % "CAC.sample_action")

    def train_net(self):
        raise NotImplementedError(_FUSED % "CAC.train_net")

    def save(self, loc):
        import torch
        torch.save(self.state_dict(), loc)

    def load(self, loc):
        import torch
        self.load_state_dict(torch.load(loc))


AGENTS = {"QTable": QTable, "Reinforce": Reinforce, "ActorCritic": ActorCritic, "CAC": CAC}
