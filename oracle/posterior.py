"""Oracle: the reference's importance-sampling posterior loops, restated particle by particle (TEST INFRASTRUCTURE ONLY —
used by tests and by bench.py's cpu_baseline legs, never by the product path).

The reference runs the user's ``forward`` once per particle in Python (pyprob/model.py:40-88) and, inside it, one torch
CPU call per sample/observe statement (pyprob/state.py:118-155 observe, :186-219 sample in the IS and IC branches), sums
the per-variable weights in Python floats (pyprob/trace.py:119-125) and normalises them in an fp64 Categorical
(pyprob/distributions/empirical.py:298-302).  The functions below do exactly that work for the BASELINE models, with the
same per-particle granularity (batch of one everywhere), so their throughput is a *lower bound on the cost* of the
reference path: they leave out the reference's address extraction (frame walks), Variable/Trace object churn and progress
printing, i.e. they are FASTER than the reference itself (survey container: reference IS 2.9 k particles/s, IC 0.26-0.37 k).

  gum_is            BASELINE configs[0]: GaussianUnknownMean, proposals from the prior
  gum_ic            north_star posterior case: GaussianUnknownMean with the LSTM proposal network (h = 512)
  marsaglia_ic      BASELINE configs[2]: GaussianUnknownMeanMarsaglia (rejection loop), TruncatedNormal-mixture proposals

Distribution arithmetic follows oracle/scoring.py (pinned to the reference's fixtures); the proposal network follows
oracle/network.py (pinned to the reference's _loss / _infer_step fixtures); samplers follow
pyprob/distributions/mixture.py:47-63 (component pick, then the component's sample) and truncated_normal.py:94-112
(inverse-CDF draw between the truncated CDF bounds, redrawn while outside the domain)."""
import math

import torch

from . import network as onet
from . import scoring

SQRT2 = math.sqrt(2.0)


def _std_normal_icdf(p):
    return SQRT2 * torch.erfinv(2 * p - 1)


def truncated_normal_sample(mean, stddev, low, high, gen):
    alpha, beta = (low - mean) / stddev, (high - mean) / stddev
    ca, cb = scoring.std_normal_cdf(alpha), scoring.std_normal_cdf(beta)
    for _ in range(10000):
        u = torch.rand((), generator=gen)
        v = _std_normal_icdf(ca + u * (cb - ca)) * stddev + mean
        if bool(torch.isfinite(v)) and bool(v >= low) and bool(v < high):
            return v
    return v


def mixture_sample(component_sample, probs, gen):
    i = int(torch.multinomial(probs / probs.sum(), 1, generator=gen))
    return component_sample(i)


def gum_is(n, obs=(8.0, 9.0), seed=0):
    """-> (values [n], log_weights [n]).  prior mu ~ Normal(1, sqrt 5); y_i ~ Normal(mu, sqrt 2) observed."""
    gen = torch.Generator().manual_seed(seed)
    values, lws = torch.empty(n), torch.empty(n)
    prior_mean, prior_sd, lik_sd = torch.tensor(1.0), torch.tensor(math.sqrt(5.0)), torch.tensor(SQRT2)
    for i in range(n):
        mu = prior_mean + prior_sd * torch.randn((), generator=gen)     # proposal == prior: weight term None (state.py:198)
        lw = 0.0
        for y in obs:
            lw += float(scoring.normal_log_prob(torch.tensor(y), mu, lik_sd))   # observed variable: state.py:147, :181
        values[i], lws[i] = mu, lw
    return values, lws


class _NetworkStepper:
    """InferenceNetworkLSTM._infer_init + _infer_step for ONE particle (inference_network_lstm.py:82-134)."""

    def __init__(self, params, observe_names, observe_in_dims, K, obs_row):
        self.P, self.K = params, K
        self.obs_emb = onet.embed_observe(params, obs_row.reshape(1, -1).float(), observe_names, observe_in_dims)
        self.H = params['_layers_lstm.weight_hh_l0'].size(1)
        self.smp_dim = next(v.size(0) for k, v in params.items()
                            if k.startswith('_layers_sample_embedding.') and k.endswith('.bias'))

    def begin(self):
        self.h, self.c = torch.zeros(1, self.H), torch.zeros(1, self.H)
        self.prev = None

    def step(self, address, family, prior0, prior1):
        P = self.P
        cur_t = P['_layers_distribution_type_embedding.' + family]
        cur_a = P['_layers_address_embedding.' + address]
        if self.prev is None:
            smp = torch.zeros(1, self.smp_dim)
            prev_t, prev_a = torch.zeros_like(cur_t), torch.zeros_like(cur_a)
        else:
            pa, pf, pv = self.prev
            smp = onet.sample_embedding(P, pa, pf, 0, pv.reshape(1))
            prev_t = P['_layers_distribution_type_embedding.' + pf]
            prev_a = P['_layers_address_embedding.' + pa]
        x = torch.cat([self.obs_emb, smp, torch.cat([prev_t, prev_a, cur_t, cur_a]).view(1, -1)], dim=1)
        H = self.H
        g = x @ P['_layers_lstm.weight_ih_l0'].t() + P['_layers_lstm.bias_ih_l0'] + \
            self.h @ P['_layers_lstm.weight_hh_l0'].t() + P['_layers_lstm.bias_hh_l0']
        i, f, gg, o = (torch.sigmoid(g[:, :H]), torch.sigmoid(g[:, H:2 * H]), torch.tanh(g[:, 2 * H:3 * H]),
                       torch.sigmoid(g[:, 3 * H:]))
        self.c = f * self.c + i * gg
        self.h = o * torch.tanh(self.c)
        return onet.head_params(P, address, family, self.K, self.h, prior0, prior1)

    def record(self, address, family, value):
        self.prev = (address, family, value)


def gum_ic(n, params, address, K=10, obs=(8.0, 9.0), seed=0):
    """IC posterior of GaussianUnknownMean: mu ~ q = mixture of K Normals from the network (state.py:203-219)."""
    gen = torch.Generator().manual_seed(seed)
    net = _NetworkStepper(params, ['obs0', 'obs1'], [1, 1], K, torch.tensor(list(obs)))
    values, lws = torch.empty(n), torch.empty(n)
    lik_sd = torch.tensor(SQRT2)
    with torch.no_grad():
        for i in range(n):
            net.begin()
            means, sds, probs = net.step(address, 'Normal', 1.0, math.sqrt(5.0))
            means, sds, probs = means[0], sds[0], probs[0]
            mu = mixture_sample(lambda k: means[k] + sds[k] * torch.randn((), generator=gen), probs, gen)
            log_q = scoring.mixture_normal_log_prob(mu.view(1), means.view(1, -1), sds.view(1, -1), probs.view(1, -1))[0]
            log_p = scoring.normal_log_prob(mu, torch.tensor(1.0), torch.tensor(math.sqrt(5.0)))
            lw = float(log_p) - float(log_q)
            for y in obs:
                lw += float(scoring.normal_log_prob(torch.tensor(y), mu, lik_sd))
            values[i], lws[i] = mu, lw
    return values, lws


def marsaglia_ic(n, params, address_of, K=10, obs=(8.0, 9.0), seed=0, max_iterations=10000):
    """IC posterior of GaussianUnknownMeanMarsaglia (tests/test_inference.py:249-275 of the reference): x, y ~ Uniform(-1,1)
    proposed from TruncatedNormal mixtures until x^2 + y^2 < 1.  ``address_of(var, k)`` -> address of variable 'x' / 'y' at
    loop iteration k (1-based) or None when the network does not know it (prior proposal, weight term zero)."""
    gen = torch.Generator().manual_seed(seed)
    net = _NetworkStepper(params, ['obs0', 'obs1'], [1, 1], K, torch.tensor(list(obs)))
    values, lws = torch.empty(n), torch.empty(n)
    lik_sd = torch.tensor(SQRT2)
    lo, hi = torch.tensor(-1.0), torch.tensor(1.0)
    with torch.no_grad():
        for i in range(n):
            net.begin()
            lw, known = 0.0, True
            for it in range(1, max_iterations + 1):
                xy = []
                for var in ('x', 'y'):
                    a = address_of(var, it) if known else None
                    if a is None:
                        known = False
                        v = lo + (hi - lo) * torch.rand((), generator=gen)
                    else:
                        means, sds, probs = net.step(a, 'Uniform', -1.0, 1.0)
                        means, sds, probs = means[0], sds[0], probs[0]
                        v = mixture_sample(lambda k: truncated_normal_sample(means[k], sds[k], lo, hi, gen), probs, gen)
                        log_q = scoring.mixture_truncated_normal_log_prob(v.view(1), means.view(1, -1), sds.view(1, -1),
                                                                           probs.view(1, -1), lo, hi)[0]
                        lw += float(scoring.uniform_log_prob(v, lo, hi)) - float(log_q)
                        net.record(a, 'Uniform', v)
                    xy.append(v)
                s = xy[0] * xy[0] + xy[1] * xy[1]
                if float(s) < 1:
                    break
            mu = 1 + math.sqrt(5.0) * (xy[0] * torch.sqrt(-2 * torch.log(s) / s))
            for y in obs:
                lw += float(scoring.normal_log_prob(torch.tensor(y), mu, lik_sd))
            values[i], lws[i] = mu, lw
    return values, lws
