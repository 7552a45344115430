"""Opt-in shared-memory-staged mixture scoring kernel (csrc/scoring.cu:k_mixture_staged, PPB_MIXTURE_STAGED=1) must
return the same values as the default kernel — it only changes how the parameter rows reach the registers."""
import os
import subprocess
import sys

import numpy as np
import pytest


pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

_SCRIPT = r'''
import sys, numpy as np, torch
sys.path.insert(0, %r)
from pyprob_b200 import ops
out = {}
for n, K in ((1, 10), (255, 10), (256, 10), (70001, 10), (4097, 3)):
    g = torch.Generator().manual_seed(n + K)
    means = torch.randn(n, K, generator=g).cuda()
    stddevs = (torch.rand(n, K, generator=g) + 0.2).cuda()
    probs = torch.rand(n, K, generator=g).cuda()
    value = torch.randn(n, generator=g).cuda()
    low = (value - torch.rand(n, generator=g).cuda() - 0.1)
    high = (value + torch.rand(n, generator=g).cuda() + 0.1)
    out['n%%d_k%%d_normal' %% (n, K)] = ops.mixture_normal_log_prob(value, means, stddevs, probs).cpu().numpy()
    out['n%%d_k%%d_trunc' %% (n, K)] = ops.mixture_truncated_normal_log_prob(value, means, stddevs, probs, low, high).cpu().numpy()
    acc = torch.zeros(n, dtype=torch.float64, device='cuda')
    ops.mixture_normal_log_prob(value, means, stddevs, probs, acc=acc, acc_scale=-1.0)
    out['n%%d_k%%d_acc' %% (n, K)] = acc.cpu().numpy()
np.savez(sys.argv[1], **out)
''' % ROOT


def _run(path, staged):
    env = dict(os.environ)
    env.pop('PPB_MIXTURE_STAGED', None)
    if staged:
        env['PPB_MIXTURE_STAGED'] = '1'
    subprocess.run([sys.executable, '-c', _SCRIPT, path], check=True, env=env, timeout=300)
    return dict(np.load(path))


def test_staged_mixture_kernel_matches_default(cuda, tmp_path):
    base = _run(str(tmp_path / 'base.npz'), staged=False)
    staged = _run(str(tmp_path / 'staged.npz'), staged=True)
    assert base.keys() == staged.keys() and len(base) == 15
    for k in base:   # same formulas on the same values; only FMA contraction may differ between the two kernels
        np.testing.assert_allclose(staged[k], base[k], rtol=2e-6, atol=2e-6, err_msg=k)
