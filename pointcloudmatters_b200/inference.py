"""Rollout-side helpers of the policies -- host-side mirror of
`src/utils/misc.py:88-140` (`TemporalAgg`, ACT's temporal ensembling of overlapping action chunks).

The policies' own inference entry points live with the modules: `act.ACTPCD.forward` without
`"actions"` in the batch (latent = 0, act.py:177-182) and
`diffusion.DiffusionUnetImagePolicy.predict_action` (CUDA-graphed DDPM sampling loop).
`TemporalAgg` is bookkeeping on a (chunk, chunk, action_dim) numpy buffer executed once per control step on
the host; it is mirrored verbatim in behaviour (including the reference's "row is empty iff all zeros"
population test) rather than moved to the device -- there is no hot loop in it.
"""
from __future__ import annotations

import numpy as np


class TemporalAgg:
    def __init__(self, apply=False, action_dim=8, chunk_size=20, k=0.01) -> None:
        self.apply = apply
        if self.apply:
            self.action_dim, self.chunk_size, self.k = action_dim, chunk_size, k
            self.action_buffer = np.zeros((chunk_size, chunk_size, action_dim))
            self.full_action = False

    def reset(self):
        self.action_buffer = np.zeros((self.chunk_size, self.chunk_size, self.action_dim))

    def _populated(self):
        # misc.py:118,128-131: a chunk row counts as populated iff it holds any non-zero entry
        return int(((self.action_buffer != 0).sum(1).sum(1) != 0).sum())

    def add_action(self, action):
        if not self.full_action:
            t = self._populated()
            self.action_buffer[t] = action
            if t == self.chunk_size - 1:
                self.full_action = True
        else:
            self.action_buffer = np.roll(self.action_buffer, -1, axis=0)
            self.action_buffer[-1] = action

    def get_action(self):
        n = self.chunk_size if self.full_action else self._populated()
        w = np.exp(-np.arange(n) * self.k)
        w = w / w.sum()
        # entry (i, n-1-i) of the last n anti-diagonal: the action each stored chunk predicted for "now"
        current = self.action_buffer[:n][np.eye(self.chunk_size)[::-1][-n:].astype(bool)]
        return (current * w[:, None]).sum(0)

    def __call__(self, action):
        if not self.apply:
            return action[0]
        self.add_action(action)
        return self.get_action()
