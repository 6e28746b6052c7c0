"""ReplayBuffer — host-side mirror of th_rl/buffers.py:6-41.

On the B200 path the per-agent transition buffer is a ring in shared memory inside the scan kernel (the newest
min(capacity, ...) transitions, replayed in insertion order and emptied by the fused train_net).  This class only
carries the capacity so that `agent.memory` exists with the reference's constructor signature.
"""

_FUSED = ("ReplayBuffer.%s is fused into the device scan (the buffer is a shared-memory ring inside thrl_qtable_scan); "
          "there is no host-side transition store")


class ReplayBuffer:
    def __init__(self, capacity, experience):
        self.capacity = capacity
        self.experience = experience

    def __len__(self):
        return 0  # transitions never live on the host

    def append(self, *args):
        raise NotImplementedError(_FUSED % "append")

    def replay(self, cast=None, replay_size=0):
        raise NotImplementedError(_FUSED % "replay")

    def sample(self, batch_size, cast=None):
        raise NotImplementedError(_FUSED % "sample")

    def empty(self):
        pass
