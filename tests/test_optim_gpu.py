"""Segment-aware optimiser kernels (csrc/optim.cu: Adam / Nesterov SGD, optionally under LARC, skipping tensors whose
gradient is absent) against the pinned oracle (oracle/optim.py, itself checked against torch.optim wrapped in the
reference's LARC class in tests/test_oracle_optim.py)."""
import ctypes as C
import math

import numpy as np
import pytest
import torch

from oracle import optim as ooptim

pytestmark = pytest.mark.gpu

SIZES = [35, 5, 12, 4, 16, 3, 70001, 130]     # parameter tensors; regions start on multiples of 4, ragged tail
KINDS = {'adam': 0, 'adam_larc': 1, 'sgd': 2, 'sgd_larc': 3}


def _layout():
    segs, off = [], 0
    for n in SIZES:
        off = (off + 3) // 4 * 4
        segs.append((off, n))
        off += n
    return segs, off


@pytest.mark.parametrize('kind', ['adam', 'adam_larc', 'sgd', 'sgd_larc'])
@pytest.mark.parametrize('wd', [0.0, 1e-2])
def test_segmented_step_matches_oracle(cuda, kind, wd):
    from pyprob_b200._lib import call, ptr
    segs, n = _layout()
    assert n % 4 != 0          # exercises the ragged last block
    S = len(segs)
    seg_of_block = np.full((n + 3) // 4, -1, np.int32)
    for k, (off, ln) in enumerate(segs):
        seg_of_block[off // 4:(off + ln + 3) // 4] = k
    gen = torch.Generator().manual_seed(3)
    p = torch.zeros(n)
    for off, ln in segs:
        p[off:off + ln] = torch.randn(ln, generator=gen)
    lr, gscale, momentum = 1e-2, 0.5, 0.9
    # oracle state (CPU) and device state
    po, mo, vo = p.clone(), torch.zeros(n), torch.zeros(n)
    steps_o, seen_o = torch.zeros(S, dtype=torch.int64), torch.zeros(S, dtype=torch.bool)
    pd, md, vd = p.cuda(), torch.zeros(n, device=cuda), torch.zeros(n, device=cuda)
    steps_d = torch.zeros(S, dtype=torch.int64, device=cuda)
    sob_d = torch.from_numpy(seg_of_block).cuda()
    scratch = torch.empty(int(call('ppb_optimizer_scratch_bytes', S)), dtype=torch.uint8, device=cuda)
    hyper = torch.tensor([lr, 0.9, 0.999, 1e-8, wd, gscale, momentum, 0.002, 1e-8, 1.0 / 16000.0], device=cuda)
    for step in range(5):
        g = torch.zeros(n)
        for off, ln in segs:
            g[off:off + ln] = torch.randn(ln, generator=gen) * (0.1 + step)
        present = [True] * S
        present[2] = step >= 2          # first appears late
        present[5] = step % 2 == 0      # rare
        present[6] = step != 3          # the big tensor misses one step
        go = g * gscale                  # the reference divides the summed gradient before the optimiser sees it
        eff_wd = wd
        if kind.endswith('larc'):
            ooptim.larc_adjust(po, go, present, segs, lr, wd)
            eff_wd = 0.0
        if kind.startswith('adam'):
            ooptim.adam_step(po, go, mo, vo, steps_o, present, segs, lr, weight_decay=eff_wd)
        else:
            ooptim.sgd_step(po, go, mo, seen_o, present, segs, lr, momentum=momentum, weight_decay=eff_wd)
        present_d = torch.tensor([int(x) for x in present], dtype=torch.int32, device=cuda)
        gd = g.cuda()
        call('ppb_optimizer_step_segmented', ptr(pd), ptr(gd), ptr(md), ptr(vd) if kind.startswith('adam') else None,
             n, ptr(sob_d), S, ptr(present_d), ptr(steps_d), ptr(scratch), scratch.numel(), KINDS[kind], ptr(hyper),
             torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        torch.testing.assert_close(pd.cpu(), po, rtol=2e-5, atol=2e-6, msg=lambda m: '{} step {}: {}'.format(kind, step, m))
        if kind.startswith('adam'):
            assert steps_d.cpu().tolist() == steps_o.tolist()
    # absent tensors were not touched on the steps they missed: segment 2 took 3 steps, not 5
    assert int(steps_d[2]) == 3 and int(steps_d[5]) == 3 and int(steps_d[6]) == 4
    # padding between regions is never written
    mask = torch.ones(n, dtype=torch.bool)
    for off, ln in segs:
        mask[off:off + ln] = False
    assert float(pd.cpu()[mask].abs().sum()) == 0.0


def test_argument_checks(cuda):
    from pyprob_b200 import _lib
    lib = _lib.load()
    t = torch.zeros(8, device=cuda)
    rc = lib.ppb_optimizer_step_segmented(t.data_ptr(), t.data_ptr(), t.data_ptr(), None, 8, t.data_ptr(), 1, t.data_ptr(),
                                          t.data_ptr(), t.data_ptr(), 1 << 20, 0, t.data_ptr(), None)
    assert rc == -1 and 'second-moment' in _lib.last_error()      # Adam without exp_avg_sq
    rc = lib.ppb_optimizer_step_segmented(t.data_ptr(), t.data_ptr(), t.data_ptr(), t.data_ptr(), 8, t.data_ptr(), 1,
                                          t.data_ptr(), t.data_ptr(), t.data_ptr(), 4, 0, t.data_ptr(), None)
    assert rc == -1 and 'scratch' in _lib.last_error()


def test_adam_larc_trains_through_the_public_api(cuda, monkeypatch):
    import pyprob_b200 as pyprob
    from pyprob_b200 import InferenceNetwork, Model, Optimizer
    from pyprob_b200.distributions import Normal

    class GaussianUnknownMean(Model):
        def forward(self):
            mu = pyprob.sample(Normal(1, math.sqrt(5)))
            pyprob.observe(Normal(mu, math.sqrt(2)), name='obs0')
            pyprob.observe(Normal(mu, math.sqrt(2)), name='obs1')
            return mu

    pyprob.seed(2)
    model = GaussianUnknownMean()
    model.learn_inference_network(num_traces=256 * 30, inference_network=InferenceNetwork.LSTM,
                                  observe_embeddings={'obs0': {'dim': 16}, 'obs1': {'dim': 16}}, batch_size=256,
                                  lstm_dim=64, optimizer_type=Optimizer.ADAM_LARC, learning_rate_init=1e-3)
    hist = model._inference_network._history_train_loss
    assert len(hist) == 30 and np.all(np.isfinite(hist)) and np.mean(hist[-5:]) < np.mean(hist[:5])


def test_flat_adam_switches_to_skipping_when_the_present_set_changes(cuda):
    """Default Optimizer.ADAM through the flat-arena kernel must follow torch.optim.Adam with grad = None for the tensors a
    minibatch does not touch (the reference on torch >= 2.0): identical while the set of touched tensors is constant, and
    from the first change on through the segment-aware step with the step counts the history implies."""
    from pyprob_b200 import synthetic
    from pyprob_b200.util import Optimizer
    table = [('a_n', 'Normal', 0), ('a_u', 'Uniform', 0), ('a_c', 'Categorical', 4)]
    net = synthetic.build_network({'o0': {'dim': 8, 'depth': 2}}, [2], table, lstm_dim=32, mixture_components=3, seed=3)
    net._optimizer_type, net._learning_rate_init, net._weight_decay = Optimizer.ADAM, 1e-3, 0.0
    net._create_optimizer()
    net._auto_skip_absent = True
    rng = np.random.default_rng(8)
    full = synthetic.ArrayBatch([synthetic.random_sub_batch(rng, table, 20, 2)])
    part = synthetic.ArrayBatch([synthetic.random_sub_batch(rng, table[:2], 12, 2)])
    names = net._segment_names()
    ref = {k: net.view(k).clone().requires_grad_(True) for k in names}
    opt = torch.optim.Adam([ref[k] for k in names], lr=1e-3)
    switched_at = None
    for it, batch in enumerate((full, full, part, full, part, full)):
        net._arena.grad = None
        ok, loss = net._loss(batch)
        assert ok
        loss.backward()
        present = net._segment_presence(net._last_enc, force=True)
        for k, name in enumerate(names):
            ref[name].grad = net.grad_view(name).clone() if present[k] else None
        opt.step()
        net.optimizer_step()
        if net._seg is not None and switched_at is None:
            switched_at = it
        for name in names:
            torch.testing.assert_close(net.view(name), ref[name].data, rtol=2e-5, atol=2e-6,
                                       msg=lambda m, name=name, it=it: 'step {} {}: {}'.format(it, name, m))
    assert switched_at == 2      # the first minibatch without a_c
