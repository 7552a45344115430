"""TEST INFRASTRUCTURE ONLY — CPU restatement of the reference's optimiser step on a flat parameter arena.

Follows pyprob/nn/inference_network.py:343-355 (`optim.Adam(lr, weight_decay)` / `optim.SGD(lr, momentum,
nesterov=True, weight_decay)`, optionally wrapped in LARC) and pyprob/nn/optimizer_larc.py:74-107 (NVIDIA apex LARC,
clip mode, trust 0.002).  The arithmetic lives in third-party torch (torch/optim/adam.py, sgd.py): the update rules
are restated here per *segment* (= one parameter tensor of the reference) of a flat fp32 arena, with the reference's
treatment of tensors whose gradient is absent from a minibatch: torch skips a parameter whose `.grad is None` —
no moment decay, no step-count increment — and LARC leaves it alone as well.

Pinned by tests/test_oracle_optim.py against torch.optim itself and against the reference's LARC class
(tests/golden/optim_golden.npz, written by tests/golden/make_optim_golden.py).
"""
import math

import torch


def adam_step(p, g, m, v, steps, present, segs, lr, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
    """One torch.optim.Adam step (no amsgrad) on the segments flagged present; in place on p, m, v, steps.

    p, g, m, v : flat fp32 tensors;  segs : list of (offset, length);  steps : int64 tensor [len(segs)] of
    per-segment step counts;  present : bool sequence [len(segs)]."""
    b1, b2 = betas
    for k, (off, n) in enumerate(segs):
        if not present[k]:
            continue
        sl = slice(off, off + n)
        steps[k] += 1
        t = int(steps[k])
        grad = g[sl]
        if weight_decay != 0:
            grad = grad + weight_decay * p[sl]
        m[sl] = m[sl] * b1 + (1 - b1) * grad                       # exp_avg.lerp_(grad, 1 - beta1)
        v[sl] = v[sl] * b2 + (1 - b2) * grad * grad                # exp_avg_sq.mul_(b2).addcmul_(grad, grad, 1 - b2)
        bc1, bc2 = 1 - b1 ** t, 1 - b2 ** t
        denom = v[sl].sqrt() / math.sqrt(bc2) + eps
        p[sl] = p[sl] - (lr / bc1) * (m[sl] / denom)
    return p


def sgd_step(p, g, buf, seen, present, segs, lr, momentum=0.9, weight_decay=0.0):
    """torch.optim.SGD(momentum, nesterov=True, dampening=0): buf = g on a tensor's first step, then
    buf = momentum*buf + g; update with g + momentum*buf.  `seen` (bool tensor) marks initialised buffers."""
    for k, (off, n) in enumerate(segs):
        if not present[k]:
            continue
        sl = slice(off, off + n)
        grad = g[sl]
        if weight_decay != 0:
            grad = grad + weight_decay * p[sl]
        if not bool(seen[k]):
            buf[sl] = grad
            seen[k] = True
        else:
            buf[sl] = momentum * buf[sl] + grad
        p[sl] = p[sl] - lr * (grad + momentum * buf[sl])
    return p


def larc_adjust(p, g, present, segs, lr, weight_decay, trust_coefficient=0.002, clip=True, eps=1e-8,
                epsilon=1.0 / 16000.0):
    """LARC.step() up to the wrapped optimiser's own step (optimizer_larc.py:80-104): per present tensor
    g <- (g + weight_decay*p) * adaptive_lr; the wrapped optimiser then runs with weight_decay = 0."""
    for k, (off, n) in enumerate(segs):
        if not present[k]:
            continue
        sl = slice(off, off + n)
        param_norm, grad_norm = torch.norm(p[sl]), torch.norm(g[sl])
        if param_norm != 0 and grad_norm != 0:
            local_lr = trust_coefficient * param_norm / (grad_norm + param_norm * weight_decay + eps)
        else:
            local_lr = epsilon
        adaptive = min(local_lr / lr, 1) if clip else local_lr
        g[sl] = (g[sl] + weight_decay * p[sl]) * adaptive
    return g
