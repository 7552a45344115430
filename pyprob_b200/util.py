"""Enums and small helpers mirroring the reference's public names (pyprob/util.py:38-77, :88-130)."""
import datetime
import enum
import math
import random
import time

import numpy as np
import torch

_verbosity = 2
_epsilon = 1e-8
_log_epsilon = math.log(_epsilon)
_seed = 0            # Philox key of the device samplers
_draw_counter = 0    # Philox offset: one fresh stream per sample statement execution


class TraceMode(enum.Enum):
    PRIOR = 1
    POSTERIOR = 2
    PRIOR_FOR_INFERENCE_NETWORK = 3


class PriorInflation(enum.Enum):
    DISABLED = 0
    ENABLED = 1


class InferenceEngine(enum.Enum):
    IMPORTANCE_SAMPLING = 0
    IMPORTANCE_SAMPLING_WITH_INFERENCE_NETWORK = 1
    LIGHTWEIGHT_METROPOLIS_HASTINGS = 2
    RANDOM_WALK_METROPOLIS_HASTINGS = 3


class InferenceNetwork(enum.Enum):
    FEEDFORWARD = 0
    LSTM = 1


class ObserveEmbedding(enum.Enum):
    FEEDFORWARD = 0
    CNN2D5C = 1
    CNN3D5C = 2


class Optimizer(enum.Enum):
    ADAM = 0
    SGD = 1
    ADAM_LARC = 2
    SGD_LARC = 3


class LearningRateScheduler(enum.Enum):
    NONE = 0
    POLY1 = 1
    POLY2 = 2


def set_verbosity(v=2):
    global _verbosity
    _verbosity = v


def seed(s=None):
    """Seed python/numpy/torch and the Philox key of the device samplers (reference: util.py:88-97)."""
    global _seed, _draw_counter
    if s is None:
        s = int((time.time() * 1e6) % 1e8)
    _seed = int(s)
    _draw_counter = 0
    random.seed(s)
    np.random.seed(s % (2 ** 32))
    torch.manual_seed(s)


def next_draw_offset():
    global _draw_counter
    _draw_counter += 1
    return _draw_counter


def set_device(device='cuda'):
    if 'cuda' not in str(device):
        raise RuntimeError('pyprob_b200 runs the hot path on CUDA only; there is no CPU execution path')
    torch.cuda.set_device(torch.device(device))


def get_time_str():
    return datetime.datetime.fromtimestamp(time.time()).strftime('%Y-%m-%d %H:%M:%S')


def get_time_stamp():
    return datetime.datetime.fromtimestamp(time.time()).strftime('-%Y%m%d-%H%M%S')


def days_hours_mins_secs_str(total_seconds):
    d, r = divmod(total_seconds, 86400)
    h, r = divmod(r, 3600)
    m, s = divmod(r, 60)
    return '{0}d:{1:02}:{2:02}:{3:02}'.format(int(d), int(h), int(m), int(s))
