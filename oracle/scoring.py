"""Oracle: per-family log_prob, restated from the reference (TEST INFRASTRUCTURE ONLY).

All inputs/outputs are CPU torch fp32 tensors; formulas follow

* Normal       pyprob/distributions/normal.py:8-11  -> torch/distributions/normal.py log_prob
* Uniform      pyprob/distributions/uniform.py:8-11 -> torch/distributions/uniform.py log_prob
* Poisson      pyprob/distributions/poisson.py:8-10 -> torch/distributions/poisson.py log_prob
* Categorical  pyprob/distributions/categorical.py:8-21 -> torch Categorical(probs=).log_prob
* Mixture      pyprob/distributions/mixture.py:8-45
* TruncatedNormal pyprob/distributions/truncated_normal.py:11-54
"""
import math

import torch

EPS32 = torch.finfo(torch.float32).eps
LOG_SQRT_2PI = math.log(math.sqrt(2 * math.pi))


def _t(x):
    return torch.as_tensor(x, dtype=torch.float32)


def normal_log_prob(value, mean, stddev):
    value, mean, stddev = _t(value), _t(mean), _t(stddev)
    var = stddev ** 2
    return -((value - mean) ** 2) / (2 * var) - stddev.log() - LOG_SQRT_2PI


def uniform_log_prob(value, low, high):
    value, low, high = _t(value), _t(low), _t(high)
    lb = low.le(value).type_as(low)
    ub = high.gt(value).type_as(low)
    return torch.log(lb.mul(ub)) - torch.log(high - low)


def poisson_log_prob(value, rate):
    value, rate = _t(value), _t(rate)
    return torch.xlogy(value, rate) - rate - torch.lgamma(value + 1)


def clamp_probs(probs):
    # pyprob/util.py:393-395
    return probs.clamp(min=EPS32, max=1 - EPS32)


def categorical_log_prob(value, probs):
    """probs [n, C] or [C] (unnormalised); value holds category indices."""
    value, probs = _t(value), _t(probs)
    probs = probs / probs.sum(-1, keepdim=True)
    logits = torch.log(clamp_probs(probs))
    idx = value.long()
    if logits.dim() == 1:
        return logits[idx]
    return logits.gather(-1, idx.view(-1, 1)).view(-1)


def std_normal_cdf(x):
    return 0.5 * (1 + torch.erf(x / math.sqrt(2)))


def truncated_normal_log_prob(value, mean, stddev, low, high):
    value, mean, stddev, low, high = _t(value), _t(mean), _t(stddev), _t(low), _t(high)
    alpha = (low - mean) / stddev
    beta = (high - mean) / stddev
    Z = std_normal_cdf(beta) - std_normal_cdf(alpha)
    log_stddev_Z = torch.log(stddev * Z)
    lb = value.ge(low).type_as(low)
    ub = value.le(high).type_as(low)
    z = (value - mean) / stddev
    return torch.log(lb.mul(ub)) + (-(z ** 2) / 2 - LOG_SQRT_2PI) - log_stddev_Z


def mixture_log_prob(component_log_probs, probs):
    """component_log_probs [n, K], probs [n, K] or [K] -> [n]  (mixture.py:15-16, :38-45)."""
    probs = _t(probs)
    probs = probs / probs.sum(-1, keepdim=True)
    log_w = torch.log(clamp_probs(probs))
    return torch.logsumexp(log_w + component_log_probs, dim=-1)


def mixture_normal_log_prob(value, means, stddevs, probs):
    value = _t(value).view(-1, 1)
    return mixture_log_prob(normal_log_prob(value, _t(means), _t(stddevs)), probs)


def mixture_truncated_normal_log_prob(value, means, stddevs, probs, low, high):
    value = _t(value).view(-1, 1)
    low = _t(low).view(-1, 1) if _t(low).dim() > 0 else _t(low)
    high = _t(high).view(-1, 1) if _t(high).dim() > 0 else _t(high)
    return mixture_log_prob(truncated_normal_log_prob(value, _t(means), _t(stddevs), low, high), probs)
