"""Tensor-level wrappers over the scoring / sampling / weight kernels (C-ABI sections 1-3).

Every function takes CUDA fp32 tensors and launches on the current torch stream.  Parameters may be
0-d / 1-element tensors (broadcast over particles) or length-n tensors.
"""
import torch

from . import _lib
from ._lib import call, ptr, stream


_scalar_cache = {}


def _f32(t, device):
    if not torch.is_tensor(t):
        if isinstance(t, (int, float)):
            # python scalars (distribution parameters shared by all particles): a host->device copy of a pageable scalar is a
            # full stream synchronisation; keep one device constant per value instead (filled by a kernel, no copy)
            key = (float(t), str(device))
            c = _scalar_cache.get(key)
            if c is None:
                if len(_scalar_cache) > 4096:
                    _scalar_cache.clear()
                c = _scalar_cache[key] = torch.full((1,), float(t), dtype=torch.float32, device=device)
            return c
        t = torch.tensor(t, dtype=torch.float32, device=device)
    return t.to(device=device, dtype=torch.float32).contiguous()


def _param(t, n, device):
    """-> (tensor kept alive, pointer, stride)"""
    t = _f32(t, device)
    if t.numel() == 1:
        return t, ptr(t), 0
    if t.numel() != n:
        raise ValueError('parameter has {} elements, expected 1 or {}'.format(t.numel(), n))
    t = t.reshape(n)
    return t, ptr(t), 1


def _sink(n, device, lp_out, acc):
    if lp_out is None and acc is None:
        lp_out = torch.empty(n, dtype=torch.float32, device=device)
    if acc is not None and (acc.dtype != torch.float64 or acc.numel() != n or not acc.is_contiguous()):
        raise ValueError('acc must be a contiguous float64 tensor of length n')
    return lp_out, acc


def normal_log_prob(value, mean, stddev, lp_out=None, acc=None, acc_scale=1.0):
    value = _f32(value, value.device).reshape(-1)
    n = value.numel()
    m, mp, ms = _param(mean, n, value.device)
    s, sp, ss = _param(stddev, n, value.device)
    lp_out, acc = _sink(n, value.device, lp_out, acc)
    call('ppb_normal_log_prob', ptr(value), mp, ms, sp, ss, ptr(lp_out), ptr(acc), float(acc_scale), n, stream())
    return lp_out


def uniform_log_prob(value, low, high, lp_out=None, acc=None, acc_scale=1.0):
    value = _f32(value, value.device).reshape(-1)
    n = value.numel()
    a, ap, as_ = _param(low, n, value.device)
    b, bp, bs = _param(high, n, value.device)
    lp_out, acc = _sink(n, value.device, lp_out, acc)
    call('ppb_uniform_log_prob', ptr(value), ap, as_, bp, bs, ptr(lp_out), ptr(acc), float(acc_scale), n, stream())
    return lp_out


def poisson_log_prob(value, rate, lp_out=None, acc=None, acc_scale=1.0):
    value = _f32(value, value.device).reshape(-1)
    n = value.numel()
    r, rp, rs = _param(rate, n, value.device)
    lp_out, acc = _sink(n, value.device, lp_out, acc)
    call('ppb_poisson_log_prob', ptr(value), rp, rs, ptr(lp_out), ptr(acc), float(acc_scale), n, stream())
    return lp_out


def _rows(t, n, device):
    """[C] shared or [n, C] per particle -> (tensor, row_stride, C)"""
    if torch.is_tensor(t) and t.is_cuda and t.dtype == torch.float32 and t.dim() == 2 and t.stride(1) == 1 \
            and t.size(0) == n and t.stride(0) >= t.size(1):
        return t, t.stride(0), t.size(1)  # strided rows (e.g. a column block of the head output): no copy
    t = _f32(t, device)
    if t.dim() == 1:
        return t, 0, t.size(0)
    if t.dim() == 2 and t.size(0) == n:
        return t, t.size(1), t.size(1)
    if t.dim() == 2 and t.size(0) == 1:
        return t.reshape(-1), 0, t.size(1)
    raise ValueError('expected [C] or [n, C], got {}'.format(tuple(t.shape)))


def categorical_log_prob(value, probs, lp_out=None, acc=None, acc_scale=1.0):
    value = _f32(value, value.device).reshape(-1)
    n = value.numel()
    p, stride, C = _rows(probs, n, value.device)
    lp_out, acc = _sink(n, value.device, lp_out, acc)
    call('ppb_categorical_log_prob', ptr(value), ptr(p), stride, C, ptr(lp_out), ptr(acc), float(acc_scale), n,
         stream())
    return lp_out


def _mixture_rows(means, stddevs, probs, n, device):
    m, sm, K = _rows(means, n, device)
    s, ss, K2 = _rows(stddevs, n, device)
    p, sp, K3 = _rows(probs, n, device)
    if not (K == K2 == K3) or not (sm == ss == sp):
        raise ValueError('means/stddevs/probs must have identical shapes')
    return m, s, p, sm, K


def mixture_normal_log_prob(value, means, stddevs, probs, lp_out=None, acc=None, acc_scale=1.0):
    value = _f32(value, value.device).reshape(-1)
    n = value.numel()
    m, s, p, stride, K = _mixture_rows(means, stddevs, probs, n, value.device)
    lp_out, acc = _sink(n, value.device, lp_out, acc)
    call('ppb_mixture_normal_log_prob', ptr(value), ptr(m), ptr(s), ptr(p), stride, K, ptr(lp_out), ptr(acc),
         float(acc_scale), n, stream())
    return lp_out


def mixture_truncated_normal_log_prob(value, means, stddevs, probs, low, high, lp_out=None, acc=None, acc_scale=1.0):
    value = _f32(value, value.device).reshape(-1)
    n = value.numel()
    m, s, p, stride, K = _mixture_rows(means, stddevs, probs, n, value.device)
    lo, lop, los = _param(low, n, value.device)
    hi, hip, his = _param(high, n, value.device)
    lp_out, acc = _sink(n, value.device, lp_out, acc)
    call('ppb_mixture_truncated_normal_log_prob', ptr(value), ptr(m), ptr(s), ptr(p), stride, K, lop, los, hip, his,
         ptr(lp_out), ptr(acc), float(acc_scale), n, stream())
    return lp_out


# ---- samplers ---------------------------------------------------------------------------------------

def _out(n, device, want_lp):
    v = torch.empty(n, dtype=torch.float32, device=device)
    lp = torch.empty(n, dtype=torch.float32, device=device) if want_lp else None
    return v, lp


def normal_sample(mean, stddev, n, seed, offset, first_index=0, with_log_prob=False, device='cuda'):
    m, mp, ms = _param(mean, n, device)
    s, sp, ss = _param(stddev, n, device)
    v, lp = _out(n, device, with_log_prob)
    call('ppb_normal_sample', mp, ms, sp, ss, ptr(v), ptr(lp), n, seed, offset, first_index, stream())
    return (v, lp) if with_log_prob else v


def uniform_sample(low, high, n, seed, offset, first_index=0, with_log_prob=False, device='cuda'):
    a, ap, as_ = _param(low, n, device)
    b, bp, bs = _param(high, n, device)
    v, lp = _out(n, device, with_log_prob)
    call('ppb_uniform_sample', ap, as_, bp, bs, ptr(v), ptr(lp), n, seed, offset, first_index, stream())
    return (v, lp) if with_log_prob else v


def poisson_sample(rate, n, seed, offset, first_index=0, with_log_prob=False, device='cuda'):
    r, rp, rs = _param(rate, n, device)
    v, lp = _out(n, device, with_log_prob)
    call('ppb_poisson_sample', rp, rs, ptr(v), ptr(lp), n, seed, offset, first_index, stream())
    return (v, lp) if with_log_prob else v


def categorical_sample(probs, n, seed, offset, first_index=0, with_log_prob=False, device='cuda'):
    p, stride, Cn = _rows(probs, n, device)
    v, lp = _out(n, device, with_log_prob)
    call('ppb_categorical_sample', ptr(p), stride, Cn, ptr(v), ptr(lp), n, seed, offset, first_index, stream())
    return (v, lp) if with_log_prob else v


def mixture_normal_sample(means, stddevs, probs, n, seed, offset, first_index=0, with_log_prob=False, device='cuda'):
    m, s, p, stride, K = _mixture_rows(means, stddevs, probs, n, device)
    v, lp = _out(n, device, with_log_prob)
    call('ppb_mixture_normal_sample', ptr(m), ptr(s), ptr(p), stride, K, ptr(v), ptr(lp), n, seed, offset,
         first_index, stream())
    return (v, lp) if with_log_prob else v


def mixture_truncated_normal_sample(means, stddevs, probs, low, high, n, seed, offset, first_index=0,
                                    with_log_prob=False, device='cuda'):
    m, s, p, stride, K = _mixture_rows(means, stddevs, probs, n, device)
    lo, lop, los = _param(low, n, device)
    hi, hip, his = _param(high, n, device)
    v, lp = _out(n, device, with_log_prob)
    call('ppb_mixture_truncated_normal_sample', ptr(m), ptr(s), ptr(p), stride, K, lop, los, hip, his, ptr(v),
         ptr(lp), n, seed, offset, first_index, stream())
    return (v, lp) if with_log_prob else v


# ---- importance weights -----------------------------------------------------------------------------

def weights_cast(acc):
    """fp64 accumulators -> (fp32 log weights, uint8 invalid mask)."""
    n = acc.numel()
    w = torch.empty(n, dtype=torch.float32, device=acc.device)
    bad = torch.empty(n, dtype=torch.uint8, device=acc.device)
    call('ppb_weights_cast', ptr(acc), ptr(w), ptr(bad), n, stream())
    return w, bad


def weights_partials(log_w):
    n = log_w.numel()
    nb = _lib.call('ppb_weights_num_partials', n)
    part = torch.empty(3 * nb, dtype=torch.float64, device=log_w.device)
    call('ppb_weights_partials', ptr(log_w), n, ptr(part), stream())
    return part


def weights_finalize(log_w, partials=None, want_logits=True):
    """-> (stats[4] = (logsumexp, ESS, max, sum exp(w-max)), normalised fp64 logits or None)."""
    log_w = log_w.contiguous()
    n = log_w.numel()
    if partials is None:
        partials = weights_partials(log_w)
    stats = torch.empty(4, dtype=torch.float64, device=log_w.device)
    logits = torch.empty(n, dtype=torch.float64, device=log_w.device) if want_logits else None
    call('ppb_weights_finalize', ptr(log_w), n, ptr(partials), partials.numel() // 3, ptr(stats), ptr(logits),
         stream())
    return stats, logits
