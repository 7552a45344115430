"""Golden per-site proposals from the UNMODIFIED reference: InferenceNetworkLSTM._infer_init / _infer_step
(pyprob/nn/inference_network.py:141-148, pyprob/nn/inference_network_lstm.py:82-134) on real reference traces.

    python tests/golden/make_infer_golden.py       # writes tests/golden/infer_golden.npz

Build container only (needs /root/reference + oracle/ref_stubs).  A straight-line model visits all four proposal-head
families; a reference network is trained briefly (seeded), then for a few prior traces the reference's own per-site
loop is replayed: _infer_init(observation), then _infer_step(variable, prev_variable) for every controlled
variable with the trace's values as the "previous" samples.  Stored: parameters, the traces, and the parameters of
every returned proposal distribution.
"""
import contextlib
import io
import math
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, '/root/reference')
sys.path.insert(0, os.path.join(ROOT, 'oracle', 'ref_stubs'))

import numpy as np  # noqa: E402
import torch  # noqa: E402

import pyprob  # noqa: E402
from pyprob import InferenceNetwork, Model  # noqa: E402
from pyprob.distributions import Categorical, Normal, Poisson, Uniform  # noqa: E402
from pyprob.nn.dataset import OnlineDataset  # noqa: E402

K, H, N_TRACES = 5, 32, 6


class FourFamilies(Model):
    def __init__(self):
        super().__init__('four families')

    def forward(self):
        u = pyprob.sample(Uniform(-1, 2))
        k = pyprob.sample(Categorical([0.2, 0.3, 0.1, 0.4]))
        z = pyprob.sample(Normal(u, 0.5 + 0.25 * k.float()))
        r = pyprob.sample(Poisson(3.0))
        mu = pyprob.sample(Normal(z * 0.1 + r * 0.05, 1))
        pyprob.observe(Normal(mu, 0.3), name='y0')
        pyprob.observe(Normal(u, 0.7), name='y1')
        return mu


def main():
    pyprob.set_verbosity(0)
    pyprob.seed(21)
    model = FourFamilies()
    emb = {'y0': {'dim': 8, 'depth': 2}, 'y1': {'dim': 4, 'depth': 1}}
    with contextlib.redirect_stdout(io.StringIO()):
        model.learn_inference_network(num_traces=96, batch_size=24, inference_network=InferenceNetwork.LSTM,
                                      observe_embeddings=emb, lstm_dim=H, proposal_mixture_components=K)
    net = model._inference_network
    net.eval()
    fx = {'dims': np.asarray([H, K, N_TRACES]), 'observe_names': np.asarray(list(emb.keys())),
          'observe_in_dims': np.asarray([1, 1])}
    for k, v in net.state_dict().items():
        fx['param/' + k] = v.detach().numpy()
    ds = OnlineDataset(model)
    with torch.no_grad():
        for i in range(N_TRACES):
            trace = ds[i]
            obs = {n: trace.named_variables[n].value for n in emb}
            net._infer_init(obs)
            prev = None
            ctrl = trace.variables_controlled
            fx['trace{}/obs'.format(i)] = np.asarray([float(obs[n]) for n in emb], np.float32)
            fx['trace{}/addresses'.format(i)] = np.asarray([v.address for v in ctrl])
            fx['trace{}/families'.format(i)] = np.asarray([v.distribution.name for v in ctrl])
            fx['trace{}/values'.format(i)] = np.asarray([float(v.value) for v in ctrl], np.float32)
            p0, p1, cats = [], [], []
            for t, var in enumerate(ctrl):
                d = var.distribution
                if d.name == 'Normal':
                    p0.append(float(d.mean)); p1.append(float(d.stddev)); cats.append(0)
                elif d.name == 'Uniform':
                    p0.append(float(d.low)); p1.append(float(d.high)); cats.append(0)
                elif d.name == 'Categorical':
                    p0.append(0.0); p1.append(0.0); cats.append(int(d.num_categories))
                else:
                    p0.append(0.0); p1.append(0.0); cats.append(0)
                q = net._infer_step(var, prev_variable=prev, proposal_min_train_iterations=None)
                key = 'trace{}/step{}/'.format(i, t)
                if q.name == 'Categorical':
                    fx[key + 'probs'] = q._probs.reshape(-1).numpy()
                else:
                    comps = q._distributions
                    if comps[0].name == 'TruncatedNormal':
                        fx[key + 'means'] = np.asarray([float(c._mean_non_truncated) for c in comps], np.float32)
                        fx[key + 'stddevs'] = np.asarray([float(c._stddev_non_truncated) for c in comps], np.float32)
                    else:
                        fx[key + 'means'] = np.asarray([float(c.mean) for c in comps], np.float32)
                        fx[key + 'stddevs'] = np.asarray([float(c.stddev) for c in comps], np.float32)
                    fx[key + 'probs'] = q._probs.reshape(-1).numpy()
                # the reference's own log-density of the trace value under the proposal (what state.py:211-217 uses)
                fx[key + 'log_prob_of_value'] = np.asarray(float(q.log_prob(var.value, sum=True)), np.float32)
                prev = var
            fx['trace{}/prior0'.format(i)] = np.asarray(p0, np.float32)
            fx['trace{}/prior1'.format(i)] = np.asarray(p1, np.float32)
            fx['trace{}/num_categories'.format(i)] = np.asarray(cats)
    np.savez_compressed(os.path.join(HERE, 'infer_golden.npz'), **fx)
    print('wrote infer_golden.npz with', len(fx), 'arrays')


if __name__ == '__main__':
    main()
