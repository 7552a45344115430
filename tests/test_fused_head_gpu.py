"""Opt-in kernel k_head_out_nll (csrc/net.cu, PPB_FUSE_HEAD_OUT=1): output layer of the proposal heads + NLL in one CUDA-core
kernel instead of the h2 tensor-core GEMM followed by k_head_nll.  Both forms share the per-row routine nll_row; loss, per-row
log-probs and every gradient must agree to rounding (the dot products are summed in a different order), and with the oracle."""
import numpy as np
import pytest
import torch

from oracle import network as onet
from pyprob_b200 import synthetic

pytestmark = pytest.mark.gpu

TABLE = [('a_u', 'Uniform', 0), ('a_c', 'Categorical', 5), ('a_n', 'Normal', 0), ('a_p', 'Poisson', 0),
         ('a_n2', 'Normal', 0), ('a_c2', 'Categorical', 3)]


def _case(seed, lstm_dim, spec, precision):
    rng = np.random.default_rng(seed)
    net = synthetic.build_network({'o0': {'dim': 12, 'depth': 2}, 'o1': {'dim': 6, 'depth': 3}}, [3, 1], TABLE,
                                  lstm_dim=lstm_dim, mixture_components=4, seed=seed, precision=precision)
    subs = [synthetic.random_sub_batch(rng, [TABLE[i] for i in seq], B, 4) for seq, B in spec]
    return net, subs


@pytest.mark.parametrize('precision', [0, 1])
@pytest.mark.parametrize('seed,lstm_dim,spec', [
    (21, 64, [([0, 1, 2, 3, 4, 5], 40), ([2], 3), ([0, 3], 64)]),
    (22, 128, [([5, 1, 4], 300), ([3], 129)]),          # segments with padding rows behind them
    (23, 256, [([2, 0, 4, 1], 140), ([3, 5], 20)]),
])
def test_fused_head_output_layer_matches_gemm_plus_nll(cuda, monkeypatch, seed, lstm_dim, spec, precision):
    monkeypatch.delenv('PPB_FUSE_HEAD_OUT', raising=False)
    base, subs = _case(seed, lstm_dim, spec, precision)
    ok, loss0 = base._loss(synthetic.ArrayBatch(subs))
    assert ok
    loss0.backward()
    g0 = base._arena.grad.clone()
    monkeypatch.setenv('PPB_FUSE_HEAD_OUT', '1')
    net, _ = _case(seed, lstm_dim, spec, precision)
    assert torch.equal(net._arena.data, base._arena.data)
    ok, loss1 = net._loss(synthetic.ArrayBatch(subs))
    assert ok
    loss1.backward()
    tol = 2e-6 if precision == 0 else 2e-3     # precision 1: the GEMM form rounds the products to tf32, the fused form does not
    assert abs(float(loss1.detach()) - float(loss0.detach())) <= tol * abs(float(loss0.detach()))
    scale = float(g0.abs().max())
    assert float((net._arena.grad - g0).abs().max()) <= 10 * tol * scale
    if precision == 0:
        params = {k: v.cpu() for k, v in net.reference_state_dict().items()}
        tsubs = [{k: (torch.from_numpy(v) if isinstance(v, np.ndarray) else v) for k, v in sb.items()} for sb in subs]
        want_loss, _, _ = onet.loss_and_grads(params, tsubs, ['o0', 'o1'], [3, 1], 4)
        assert abs(float(loss1.detach()) - float(want_loss)) <= 1e-4 * abs(float(want_loss))
