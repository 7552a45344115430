"""Particle-batched distributions backed by the CUDA scoring / sampling kernels.

Public protocol follows the reference (pyprob/distributions/distribution.py:9-95): ``name``,
``_address_suffix``, ``sample()``, ``log_prob(value, sum=False)``, ``mean/stddev/variance``.  Parameters may
be Python scalars (shared by all particles) or length-n CUDA tensors (one per particle).  There is no CPU
path: every sample/log_prob is a kernel launch over the particle axis.
"""
import math

import torch

from . import ops, util

_shard_first_index = 0  # global index of particle 0 on this rank (set by the engine for sharded runs)


def set_shard_first_index(i):
    global _shard_first_index
    _shard_first_index = int(i)


def _as_param(x):
    if torch.is_tensor(x):
        return x.to(device='cuda', dtype=torch.float32).reshape(-1) if x.numel() > 1 else float(x)
    return float(x)


def _length(*params):
    n = 1
    for p in params:
        if torch.is_tensor(p):
            n = max(n, p.numel())
    return n


def _value(v, n=None):
    if not torch.is_tensor(v):
        v = torch.tensor(v, dtype=torch.float32)
    v = v.to(device='cuda', dtype=torch.float32).reshape(-1)
    if n is not None and v.numel() == 1 and n > 1:
        v = v.expand(n).contiguous()
    return v


class Distribution:
    def __init__(self, name, address_suffix):
        self.name = name
        self._address_suffix = address_suffix

    @property
    def batch_length(self):
        return 1

    def _draw(self, n, with_log_prob):
        raise NotImplementedError()

    def sample(self, n=None, with_log_prob=False):
        n = self.batch_length if n is None else n
        return self._draw(n, with_log_prob)

    def log_prob(self, value, sum=False):
        lp = self._log_prob(value)
        return lp.sum() if sum else lp

    def prob(self, value):
        return torch.exp(self.log_prob(value))

    @property
    def stddev(self):
        return self.variance ** 0.5 if torch.is_tensor(self.variance) else math.sqrt(self.variance)

    def _seed_args(self):
        return util._seed, util.next_draw_offset(), _shard_first_index


class Normal(Distribution):
    def __init__(self, loc, scale):
        super().__init__('Normal', 'Normal')
        self.loc, self.scale = _as_param(loc), _as_param(scale)

    @property
    def batch_length(self):
        return _length(self.loc, self.scale)

    mean = property(lambda self: self.loc)
    variance = property(lambda self: self.scale ** 2)
    stddev = property(lambda self: self.scale)

    def _draw(self, n, with_log_prob):
        s, o, f = self._seed_args()
        return ops.normal_sample(self.loc, self.scale, n, s, o, f, with_log_prob)

    def _log_prob(self, value):
        return ops.normal_log_prob(_value(value, self.batch_length), self.loc, self.scale)

    def score_into(self, value, acc, scale):
        ops.normal_log_prob(value, self.loc, self.scale, acc=acc, acc_scale=scale)

    def __repr__(self):
        return 'Normal({}, {})'.format(self.loc, self.scale)


class Uniform(Distribution):
    def __init__(self, low, high):
        super().__init__('Uniform', 'Uniform')
        self.low, self.high = _as_param(low), _as_param(high)

    @property
    def batch_length(self):
        return _length(self.low, self.high)

    mean = property(lambda self: (self.low + self.high) / 2)
    variance = property(lambda self: (self.high - self.low) ** 2 / 12)

    def _draw(self, n, with_log_prob):
        s, o, f = self._seed_args()
        return ops.uniform_sample(self.low, self.high, n, s, o, f, with_log_prob)

    def _log_prob(self, value):
        return ops.uniform_log_prob(_value(value, self.batch_length), self.low, self.high)

    def score_into(self, value, acc, scale):
        ops.uniform_log_prob(value, self.low, self.high, acc=acc, acc_scale=scale)

    def __repr__(self):
        return 'Uniform(low={}, high={})'.format(self.low, self.high)


class Poisson(Distribution):
    def __init__(self, rate):
        super().__init__('Poisson', 'Poisson')
        self.rate = _as_param(rate)

    @property
    def batch_length(self):
        return _length(self.rate)

    mean = property(lambda self: self.rate)
    variance = property(lambda self: self.rate)

    def _draw(self, n, with_log_prob):
        s, o, f = self._seed_args()
        return ops.poisson_sample(self.rate, n, s, o, f, with_log_prob)

    def _log_prob(self, value):
        return ops.poisson_log_prob(_value(value, self.batch_length), self.rate)

    def score_into(self, value, acc, scale):
        ops.poisson_log_prob(value, self.rate, acc=acc, acc_scale=scale)

    def __repr__(self):
        return 'Poisson({})'.format(self.rate)


class Categorical(Distribution):
    """probs: [C] shared or [n, C] per particle (unnormalised, like the reference: categorical.py:8-21)."""

    def __init__(self, probs=None, logits=None):
        if probs is None:
            probs = torch.softmax(torch.as_tensor(logits, dtype=torch.float32), dim=-1)
        probs = torch.as_tensor(probs, dtype=torch.float32).to('cuda')
        if probs.dim() == 0:
            raise ValueError('probs cannot be a scalar.')
        self._probs = probs.contiguous()
        self._num_categories = probs.size(-1)
        super().__init__('Categorical', 'Categorical(len_probs:{})'.format(self._num_categories))

    @property
    def batch_length(self):
        return self._probs.size(0) if self._probs.dim() == 2 else 1

    num_categories = property(lambda self: self._num_categories)

    @property
    def probs(self):
        return self._probs / self._probs.sum(-1, keepdim=True)

    @property
    def mean(self):
        return (self.probs * torch.arange(self._num_categories, device='cuda')).sum(-1)

    @property
    def variance(self):
        idx = torch.arange(self._num_categories, device='cuda')
        return (self.probs * idx ** 2).sum(-1) - self.mean ** 2

    def _draw(self, n, with_log_prob):
        s, o, f = self._seed_args()
        return ops.categorical_sample(self._probs, n, s, o, f, with_log_prob)

    def _log_prob(self, value):
        return ops.categorical_log_prob(_value(value, self.batch_length), self._probs)

    def score_into(self, value, acc, scale):
        ops.categorical_log_prob(value, self._probs, acc=acc, acc_scale=scale)

    def __repr__(self):
        return 'Categorical(num_categories={})'.format(self._num_categories)


class TruncatedNormal(Distribution):
    """Scored and drawn as a one-component truncated mixture (reference: truncated_normal.py:11-112)."""

    def __init__(self, mean_non_truncated, stddev_non_truncated, low, high):
        super().__init__('TruncatedNormal', 'TruncatedNormal')
        self.mean_non_truncated, self.stddev_non_truncated = _as_param(mean_non_truncated), _as_param(stddev_non_truncated)
        self.low, self.high = _as_param(low), _as_param(high)

    @property
    def batch_length(self):
        return _length(self.mean_non_truncated, self.stddev_non_truncated, self.low, self.high)

    def _rows(self, n):
        def col(p):
            t = p if torch.is_tensor(p) else torch.full((n,), p, device='cuda')
            return t.reshape(-1, 1).expand(n, 1).contiguous()
        return col(self.mean_non_truncated), col(self.stddev_non_truncated), torch.ones(n, 1, device='cuda')

    def _draw(self, n, with_log_prob):
        s, o, f = self._seed_args()
        m, sd, p = self._rows(n)
        return ops.mixture_truncated_normal_sample(m, sd, p, self.low, self.high, n, s, o, f, with_log_prob)

    def _log_prob(self, value):
        v = _value(value, self.batch_length)
        m, sd, p = self._rows(v.numel())
        return ops.mixture_truncated_normal_log_prob(v, m, sd, p, self.low, self.high)


class Mixture(Distribution):
    """Mixture of K Normals or K TruncatedNormals with per-particle parameters.

    Either built like the reference (``Mixture([Normal(..), ...], probs)``, mixture.py:8-30) or directly from
    parameter rows with ``Mixture.from_rows``."""

    def __init__(self, distributions, probs=None):
        super().__init__('Mixture', 'Mixture({})'.format(', '.join(d._address_suffix for d in distributions)))
        K = len(distributions)
        n = max(d.batch_length for d in distributions)
        trunc = isinstance(distributions[0], TruncatedNormal)

        def stack(attr):
            cols = []
            for d in distributions:
                p = getattr(d, attr)
                cols.append(p.reshape(-1) if torch.is_tensor(p) else torch.full((n,), p, device='cuda'))
            return torch.stack(cols, dim=1).contiguous()
        self._means = stack('mean_non_truncated' if trunc else 'loc')
        self._stddevs = stack('stddev_non_truncated' if trunc else 'scale')
        if probs is None:
            probs = torch.full((K,), 1.0 / K)
        probs = torch.as_tensor(probs, dtype=torch.float32).to('cuda')
        self._probs = (probs.expand(n, K) if probs.dim() == 1 else probs).contiguous()
        self._low = distributions[0].low if trunc else None
        self._high = distributions[0].high if trunc else None
        self._trunc, self._n, self.length = trunc, n, K

    @classmethod
    def from_rows(cls, means, stddevs, probs, low=None, high=None):
        self = cls.__new__(cls)
        Distribution.__init__(self, 'Mixture', 'Mixture')
        self._means, self._stddevs, self._probs = means, stddevs, probs
        self._low, self._high = low, high
        self._trunc, self._n, self.length = low is not None, means.size(0), means.size(1)
        return self

    @property
    def batch_length(self):
        return self._n

    @property
    def probs(self):
        return self._probs / self._probs.sum(-1, keepdim=True)

    @property
    def mean(self):
        if self._trunc:
            raise NotImplementedError('mean of a truncated mixture')
        return (self.probs * self._means).sum(-1)

    @property
    def variance(self):
        m = self.mean.view(-1, 1)
        return (self.probs * ((self._means - m) ** 2 + self._stddevs ** 2)).sum(-1)

    def _draw(self, n, with_log_prob):
        s, o, f = self._seed_args()
        if self._trunc:
            return ops.mixture_truncated_normal_sample(self._means, self._stddevs, self._probs, self._low, self._high,
                                                       n, s, o, f, with_log_prob)
        return ops.mixture_normal_sample(self._means, self._stddevs, self._probs, n, s, o, f, with_log_prob)

    def _log_prob(self, value):
        v = _value(value, self._n)
        if self._trunc:
            return ops.mixture_truncated_normal_log_prob(v, self._means, self._stddevs, self._probs, self._low,
                                                         self._high)
        return ops.mixture_normal_log_prob(v, self._means, self._stddevs, self._probs)

    def score_into(self, value, acc, scale):
        if self._trunc:
            ops.mixture_truncated_normal_log_prob(value, self._means, self._stddevs, self._probs, self._low, self._high,
                                                  acc=acc, acc_scale=scale)
        else:
            ops.mixture_normal_log_prob(value, self._means, self._stddevs, self._probs, acc=acc, acc_scale=scale)
