"""CPU: the oracle's restatement of the per-site proposal step (oracle/network.py:infer_sequence, head_params) against
the UNMODIFIED reference's InferenceNetworkLSTM._infer_init/_infer_step on real reference traces
(tests/golden/infer_golden.npz, written by tests/golden/make_infer_golden.py)."""
import os

import numpy as np
import torch

from oracle import network as onet
from oracle import scoring

HERE = os.path.dirname(os.path.abspath(__file__))


def load_infer_golden():
    z = dict(np.load(os.path.join(HERE, 'golden', 'infer_golden.npz')))
    H, K, n_traces = (int(x) for x in z['dims'])
    params = {k[len('param/'):]: torch.from_numpy(v) for k, v in z.items() if k.startswith('param/')}
    traces = []
    for i in range(n_traces):
        p = 'trace{}/'.format(i)
        T = len(z[p + 'addresses'])
        steps = []
        for t in range(T):
            steps.append({'address': str(z[p + 'addresses'][t]), 'family': str(z[p + 'families'][t]),
                          'num_categories': int(z[p + 'num_categories'][t]), 'prior0': float(z[p + 'prior0'][t]),
                          'prior1': float(z[p + 'prior1'][t]),
                          'prev_value': torch.tensor([float(z[p + 'values'][t - 1])]) if t > 0 else None,
                          'value': float(z[p + 'values'][t]),
                          'want': {k: z['{}step{}/{}'.format(p, t, k)] for k in ('means', 'stddevs', 'probs', 'log_prob_of_value')
                                   if '{}step{}/{}'.format(p, t, k) in z}})
        traces.append({'obs': torch.from_numpy(z[p + 'obs']), 'steps': steps})
    return {'H': H, 'K': K, 'params': params, 'observe_names': [str(x) for x in z['observe_names']],
            'observe_in_dims': [int(x) for x in z['observe_in_dims']], 'traces': traces}


def test_fixture_visits_every_family():
    fx = load_infer_golden()
    fams = {st['family'] for tr in fx['traces'] for st in tr['steps']}
    assert fams == {'Uniform', 'Categorical', 'Normal', 'Poisson'}


def test_infer_sequence_matches_reference_infer_step():
    fx = load_infer_golden()
    for tr in fx['traces']:
        got = onet.infer_sequence(fx['params'], tr['obs'], fx['observe_names'], fx['observe_in_dims'], fx['K'], tr['steps'])
        for st, q in zip(tr['steps'], got):
            want = st['want']
            if st['family'] == 'Categorical':
                # reference Categorical normalises the probs it is given (categorical.py:8-21)
                probs = q[0] / q[0].sum(dim=1, keepdim=True)
                np.testing.assert_allclose(probs[0].numpy(), want['probs'], rtol=1e-5, atol=1e-7)
                lp = scoring.categorical_log_prob(torch.tensor([st['value']]), q[0])
            else:
                means, stddevs, coeffs = q
                np.testing.assert_allclose(means[0].numpy(), want['means'], rtol=1e-5, atol=1e-6)
                np.testing.assert_allclose(stddevs[0].numpy(), want['stddevs'], rtol=1e-5, atol=1e-7)
                np.testing.assert_allclose(coeffs[0].numpy(), want['probs'], rtol=1e-5, atol=1e-7)
                v = torch.tensor([st['value']])
                if st['family'] == 'Normal':
                    lp = scoring.mixture_normal_log_prob(v, means, stddevs, coeffs)
                elif st['family'] == 'Uniform':
                    lp = scoring.mixture_truncated_normal_log_prob(v, means, stddevs, coeffs, torch.tensor([st['prior0']]),
                                                                   torch.tensor([st['prior1']]))
                else:
                    lp = scoring.mixture_truncated_normal_log_prob(v, means, stddevs, coeffs, torch.zeros(1),
                                                                   torch.full((1,), 40.0))
            # the proposal density the IC weight uses (state.py:211-217), as the reference computed it
            np.testing.assert_allclose(float(lp), float(want['log_prob_of_value']), rtol=1e-4, atol=1e-5)
