"""Programmatic dependent launch (csrc/common.cuh: ppb_launch, ppb_pdl_wait): kernels of the critical chains may start while
their predecessor is still running and block at griddepcontrol.wait before touching its results.  A missing or misplaced wait
is a race, so the step must give the same loss and gradients with PPB_PDL=0 (plain stream order) and with the default (the
graph-replayed form of the same launches is covered by tests/test_host_step_gpu.py and bench.py), and agree with the oracle."""
import numpy as np
import pytest
import torch

from oracle import network as onet
from pyprob_b200 import synthetic

pytestmark = pytest.mark.gpu

TABLE = [('a_u', 'Uniform', 0), ('a_c', 'Categorical', 5), ('a_n', 'Normal', 0), ('a_p', 'Poisson', 0),
         ('a_n2', 'Normal', 0), ('a_c2', 'Categorical', 3)]


def _case(seed, lstm_dim, spec):
    rng = np.random.default_rng(seed)
    net = synthetic.build_network({'o0': {'dim': 12, 'depth': 2}, 'o1': {'dim': 6, 'depth': 3}}, [3, 1], TABLE,
                                  lstm_dim=lstm_dim, mixture_components=4, seed=seed, precision=0)
    subs = [synthetic.random_sub_batch(rng, [TABLE[i] for i in seq], B, 4) for seq, B in spec]
    return net, subs


def _loss_and_grad(net, subs):
    ok, loss = net._loss(synthetic.ArrayBatch(subs))
    assert ok
    loss.backward()
    return float(loss.detach()), net._arena.grad.clone()


@pytest.mark.parametrize('seed,lstm_dim,spec', [
    (11, 64, [([0, 1, 2, 3, 4, 5], 40), ([2], 3), ([0, 3], 64)]),
    (12, 128, [([0, 1, 2, 3, 4, 5, 0, 1, 2, 3, 4, 5], 260)]),      # twelve dependent LSTM steps, three row tiles
    (13, 256, [([2, 0, 4, 1], 140), ([3, 5], 20)]),
])
def test_step_is_the_same_with_and_without_pdl(cuda, monkeypatch, seed, lstm_dim, spec):
    monkeypatch.setenv('PPB_PDL', '0')
    base, subs = _case(seed, lstm_dim, spec)
    loss0, g0 = _loss_and_grad(base, subs)
    monkeypatch.setenv('PPB_PDL', '1')
    net, _ = _case(seed, lstm_dim, spec)
    assert torch.equal(net._arena.data, base._arena.data)
    scale = float(g0.abs().max())
    for _ in range(3):    # a race does not have to show on the first try
        net._arena.grad = None
        loss1, g1 = _loss_and_grad(net, subs)
        # not bit-identical: the weight-gradient GEMMs reduce with fp32 atomics whatever the launch mode
        assert abs(loss1 - loss0) <= 2e-6 * abs(loss0)
        assert float((g1 - g0).abs().max()) <= 2e-5 * scale
    params = {k: v.cpu() for k, v in net.reference_state_dict().items()}
    tsubs = [{k: (torch.from_numpy(v) if isinstance(v, np.ndarray) else v) for k, v in sb.items()} for sb in subs]
    want_loss, want, _ = onet.loss_and_grads(params, tsubs, ['o0', 'o1'], [3, 1], 4)
    assert abs(loss1 - float(want_loss)) <= 1e-4 * abs(float(want_loss))
