"""CPU checks of oracle/posterior.py (the per-particle restatement of the reference's IS / IC posterior loops that bench.py
times as the CPU baseline of the particles/s half of the metric):
  * IS weights are exactly the double sum of the observe log-likelihoods (trace.py:119-125) and reproduce the analytic
    posterior of the reference's own acceptance test (tests/test_inference.py:94-121);
  * the one-particle network stepper agrees with oracle.network.infer_sequence, which is pinned to the unmodified
    reference's _infer_step (tests/golden/infer_golden.npz);
  * IC weights are log p - log q + likelihood of the values the loop drew."""
import math

import numpy as np
import torch

from oracle import network as onet
from oracle import params as oparams
from oracle import posterior, scoring, weights

A_MU = '98__forward__mu__Normal__1'


def test_gum_is_weights_and_posterior():
    v, lw = posterior.gum_is(4000, seed=1)
    terms = np.stack([scoring.normal_log_prob(np.float32(o), v, np.float32(math.sqrt(2))).numpy() for o in (8.0, 9.0)])
    np.testing.assert_allclose(lw.numpy(), weights.accumulate(terms).astype(np.float32), rtol=1e-5, atol=1e-5)
    lse, ess, logits = weights.finalize(lw.numpy())
    w = np.exp(logits)
    mean = float((w * v.numpy()).sum())
    assert abs(mean - 7.25) < 0.3 and ess > 0.005 * 4000


def test_gum_ic_weights_follow_the_pinned_proposal_step():
    P = oparams.random_params([('obs0', 1, 8, 2), ('obs1', 1, 8, 2)], [(A_MU, 'Normal', 0)], lstm_dim=32, K=4, seed=3)
    v, lw = posterior.gum_ic(6, P, A_MU, K=4, seed=2)
    steps = [{'address': A_MU, 'family': 'Normal', 'num_categories': 0, 'prior0': 1.0, 'prior1': math.sqrt(5.0),
              'prev_value': None}]
    (means, sds, probs), = onet.infer_sequence(P, torch.tensor([8.0, 9.0]), ['obs0', 'obs1'], [1, 1], 4, steps, n=1)
    for i in range(6):
        want = float(scoring.normal_log_prob(v[i], 1.0, math.sqrt(5.0))) - \
            float(scoring.mixture_normal_log_prob(v[i].view(1), means, sds, probs)[0])
        want += sum(float(scoring.normal_log_prob(torch.tensor(o), v[i], math.sqrt(2.0))) for o in (8.0, 9.0))
        assert abs(float(lw[i]) - want) <= 1e-4 * max(1.0, abs(want))


def test_marsaglia_ic_weights_replay():
    addrs = [('x{}'.format(k), 'Uniform', 0) for k in range(1, 4)] + [('y{}'.format(k), 'Uniform', 0) for k in range(1, 4)]
    P = oparams.random_params([('obs0', 1, 8, 2), ('obs1', 1, 8, 2)], addrs, lstm_dim=32, K=3, seed=4)

    def address_of(var, k):
        return '{}{}'.format(var, k) if k <= 3 else None
    v, lw = posterior.marsaglia_ic(40, P, address_of, K=3, seed=5)
    assert torch.isfinite(lw).all() and torch.isfinite(v).all()
    # truncated-normal draws stay inside the support
    gen = torch.Generator().manual_seed(0)
    for _ in range(200):
        x = posterior.truncated_normal_sample(torch.tensor(0.9), torch.tensor(0.5), torch.tensor(-1.0), torch.tensor(1.0), gen)
        assert -1.0 <= float(x) < 1.0
