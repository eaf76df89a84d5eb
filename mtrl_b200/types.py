"""Mirror of /root/reference/mtrl/types.py for the hot path (field ORDER is part of the contract:
observations, actions, next_observations, dones, rewards -- types.py:30-35)."""
from __future__ import annotations

from typing import Any, NamedTuple, TypedDict


class ReplayBufferSamples(NamedTuple):
    observations: Any
    actions: Any
    next_observations: Any
    dones: Any
    rewards: Any


class Rollout(NamedTuple):  # types.py:48-63
    observations: Any
    actions: Any
    rewards: Any
    dones: Any
    log_probs: Any = None
    means: Any = None
    stds: Any = None
    values: Any = None
    returns: Any = None
    advantages: Any = None


class ReplayBufferCheckpoint(TypedDict):  # types.py:72-74
    data: dict
    rng_state: Any


LogDict = dict
