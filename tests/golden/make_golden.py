"""Generate golden fixtures by running the UNMODIFIED reference (pyprob v1.5.0) in the build container.

    python tests/golden/make_golden.py            # writes tests/golden/*.npz

Needs /root/reference (read-only) and the import stubs in oracle/ref_stubs (five pure-Python,
non-arithmetic dependencies of the reference that are not installed here).
The GPU box has no /root/reference, so the outputs are committed as small fixtures.
"""
import math
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, '/root/reference')
sys.path.insert(0, os.path.join(ROOT, 'oracle', 'ref_stubs'))

import numpy as np  # noqa: E402
import torch  # noqa: E402

import pyprob  # noqa: E402  (the reference)
from pyprob.distributions import (Categorical, Empirical, Mixture, Normal, Poisson, TruncatedNormal,  # noqa: E402
                                  Uniform)


def scoring_fixture(seed=1234, n=257, K=10, C=7):
    g = torch.Generator().manual_seed(seed)

    def rnd(*shape):
        return torch.randn(*shape, generator=g)

    out = {}
    # Normal
    v, mu, sd = rnd(n) * 3, rnd(n) * 2, rnd(n).abs() + 0.1
    out['normal_value'], out['normal_mean'], out['normal_stddev'] = v, mu, sd
    out['normal_lp'] = Normal(mu, sd).log_prob(v)
    # Uniform (values inside the support: torch validates by default)
    lo = rnd(n) - 2
    hi = lo + rnd(n).abs() + 0.5
    v = lo + torch.rand(n, generator=g) * (hi - lo) * 0.999
    out['uniform_value'], out['uniform_low'], out['uniform_high'] = v, lo, hi
    out['uniform_lp'] = Uniform(lo, hi).log_prob(v)
    # Poisson
    rate = rnd(n).abs() * 6 + 0.2
    v = torch.poisson(rate, generator=g)
    out['poisson_value'], out['poisson_rate'] = v, rate
    out['poisson_lp'] = Poisson(rate).log_prob(v)
    # Categorical, per-particle probs
    probs = torch.rand(n, C, generator=g) + 0.01
    probs[::11, 2] = 0.0  # exercise the clamp
    v = torch.randint(0, C, (n,), generator=g)
    out['categorical_value'], out['categorical_probs'] = v.float(), probs
    out['categorical_lp'] = Categorical(probs).log_prob(v)
    # Mixture of Normals (batched, as the proposal heads build it)
    means, sds = rnd(n, K) * 2, rnd(n, K).abs() + 0.05
    coeffs = torch.softmax(rnd(n, K) * 2, dim=1)
    v = rnd(n) * 2
    mix = Mixture([Normal(means[:, i], sds[:, i]) for i in range(K)], coeffs)
    out['mixn_value'], out['mixn_means'], out['mixn_stddevs'], out['mixn_probs'] = v, means, sds, coeffs
    out['mixn_lp'] = mix.log_prob(v)
    # Mixture of TruncatedNormals
    lo = rnd(n) - 1.5
    hi = lo + rnd(n).abs() * 2 + 1.0
    means = lo.view(n, 1) + torch.rand(n, K, generator=g) * (hi - lo).view(n, 1)
    sds = (hi - lo).view(n, 1) * (0.001 + torch.rand(n, K, generator=g) * 2)
    coeffs = torch.softmax(rnd(n, K), dim=1)
    v = lo + torch.rand(n, generator=g) * (hi - lo)
    v[::13] = hi[::13] + 0.25  # outside the truncation domain -> -inf
    mix = Mixture([TruncatedNormal(means[:, i], sds[:, i], low=lo, high=hi) for i in range(K)], coeffs)
    out['mixt_value'], out['mixt_means'], out['mixt_stddevs'], out['mixt_probs'] = v, means, sds, coeffs
    out['mixt_low'], out['mixt_high'] = lo, hi
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        import io
        import contextlib
        with contextlib.redirect_stdout(io.StringIO()):
            out['mixt_lp'] = mix.log_prob(v)
    # Importance weights: Empirical.finalize / ESS
    lw = (rnd(1000) * 4 - 30).float()
    emp = Empirical(values=list(range(1000)), log_weights=lw)
    out['weights_log_w'] = lw
    out['weights_logits'] = emp._categorical.logits  # fp64
    out['weights_ess'] = torch.as_tensor(float(emp.effective_sample_size), dtype=torch.float64)
    return {k: (v.detach().numpy() if torch.is_tensor(v) else np.asarray(v)) for k, v in out.items()}


def gum_is_fixture(seed=7, n=64):
    """Config 1 semantics: GUM importance sampling from the prior — per-trace weights as the reference
    accumulates them (state.py:147-149, trace.py:123-125) for given latent draws."""
    from pyprob import Model

    class GUM(Model):
        def __init__(self):
            super().__init__('gum')

        def forward(self):
            mu = pyprob.sample(Normal(1, math.sqrt(5)))
            lik = Normal(mu, math.sqrt(2))
            pyprob.observe(lik, name='obs0')
            pyprob.observe(lik, name='obs1')
            return mu

    pyprob.seed(seed)
    pyprob.set_verbosity(0)
    model = GUM()
    post = model.posterior(n, observe={'obs0': 8, 'obs1': 9})
    mus = np.array([float(t.result) for t in post.values], dtype=np.float32)
    lw = np.array([float(w) for w in post.log_weights], dtype=np.float32)
    return {'gum_mu': mus, 'gum_log_w': lw,
            'gum_logits': post._categorical.logits.numpy(),
            'gum_ess': np.asarray(float(post.effective_sample_size))}


def make_scoring_fixtures():
    fx = scoring_fixture()
    fx.update(gum_is_fixture())
    np.savez_compressed(os.path.join(HERE, 'scoring_golden.npz'), **fx)
    print('wrote scoring_golden.npz with', len(fx), 'arrays')


# ------------------------------------------------------------------------------------------------------
# network fixtures: reference InferenceNetworkLSTM._loss + backward on real reference traces
# ------------------------------------------------------------------------------------------------------
def _silence():
    import contextlib
    import io
    return contextlib.redirect_stdout(io.StringIO())


def network_fixture(model, observe_embeddings, lstm_dim, K, batch_size, train_traces, seed, tag):
    from pyprob import InferenceNetwork as INType
    from pyprob.nn.dataset import Batch, OnlineDataset
    sys.path.insert(0, ROOT)
    from oracle import network as onet
    pyprob.seed(seed)
    with _silence():
        model.learn_inference_network(num_traces=train_traces, batch_size=batch_size, inference_network=INType.LSTM,
                                      observe_embeddings=observe_embeddings, lstm_dim=lstm_dim,
                                      proposal_mixture_components=K)
    net = model._inference_network
    ds = OnlineDataset(model)
    traces = [ds[i] for i in range(batch_size)]
    batch = Batch(traces)
    with _silence():
        changed = net._polymorph(batch)
    net.zero_grad()
    success, loss = net._loss(batch)
    assert success
    loss.backward()
    fx = {tag + '/loss': np.asarray(float(loss), dtype=np.float64)}
    names = list(observe_embeddings.keys())
    in_dims = [int(np.prod(traces[0].named_variables[n].value.shape)) or 1 for n in names]
    fx[tag + '/observe_names'] = np.asarray(names)
    fx[tag + '/observe_in_dims'] = np.asarray(in_dims)
    fx[tag + '/dims'] = np.asarray([lstm_dim, K, batch_size])
    for k, v in net.state_dict().items():
        fx[tag + '/param/' + k] = v.detach().numpy()
    for k, p in net.named_parameters():
        g = p.grad if p.grad is not None else torch.zeros_like(p)
        fx[tag + '/grad/' + k] = g.detach().numpy()
    # address table in the network's insertion order (= address ids)
    fx[tag + '/address_order'] = np.asarray(list(net._layers_address_embedding.keys()))
    fx[tag + '/type_order'] = np.asarray(list(net._layers_distribution_type_embedding.keys()))
    for s, sub in enumerate(batch.sub_batches):
        sb = onet.sub_batch_from_traces(sub, names)
        p = '{}/sub{}/'.format(tag, s)
        fx[p + 'addresses'] = np.asarray(sb['addresses'])
        fx[p + 'families'] = np.asarray(sb['families'])
        fx[p + 'num_categories'] = np.asarray(sb['num_categories'])
        for k in ('values', 'prior0', 'prior1', 'obs'):
            fx[p + k] = sb[k].numpy()
        # per-(t,b) reference log-probs of the proposal, recomputed through the reference layers
        with torch.no_grad():
            obs_emb = net._embed_observe(sub)
    fx[tag + '/num_sub'] = np.asarray(len(batch.sub_batches))
    return fx


def make_network_fixtures():
    from pyprob import Model

    class GUM(Model):
        def __init__(self):
            super().__init__('gum')

        def forward(self):
            mu = pyprob.sample(Normal(1, math.sqrt(5)))
            lik = Normal(mu, math.sqrt(2))
            pyprob.observe(lik, name='obs0')
            pyprob.observe(lik, name='obs1')
            return mu

    class Mixed(Model):
        """All four proposal-head families, data-dependent priors, two trace types (branching)."""

        def __init__(self):
            super().__init__('mixed')

        def forward(self):
            u = pyprob.sample(Uniform(-1, 2))
            k = pyprob.sample(Categorical([0.2, 0.3, 0.5]))
            if int(k) == 0:
                z = pyprob.sample(Normal(u, 0.5))
            else:
                z = pyprob.sample(Poisson(3.0))
                z = z * 0.25
            mu = pyprob.sample(Normal(z * 0.1, 1))
            pyprob.observe(Normal(mu, 0.3), name='y0')
            pyprob.observe(Normal(u, 0.7), name='y1')
            return mu

    pyprob.set_verbosity(0)
    fx = {}
    fx.update(network_fixture(GUM(), {'obs0': {'dim': 8}, 'obs1': {'dim': 8}}, lstm_dim=32, K=10, batch_size=16,
                              train_traces=64, seed=11, tag='gum'))
    fx.update(network_fixture(Mixed(), {'y0': {'dim': 8, 'depth': 2}, 'y1': {'dim': 4, 'depth': 1}}, lstm_dim=32, K=5,
                              batch_size=24, train_traces=96, seed=12, tag='mixed'))
    np.savez_compressed(os.path.join(HERE, 'network_golden.npz'), **fx)
    print('wrote network_golden.npz with', len(fx), 'arrays')


if __name__ == '__main__':
    make_scoring_fixtures()
    make_network_fixtures()
