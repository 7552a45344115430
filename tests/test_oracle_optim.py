"""CPU: the flat-arena optimiser restatement (oracle/optim.py) against trajectories produced by torch.optim wrapped
in the reference's own LARC class (tests/golden/optim_golden.npz), including tensors whose gradient is absent on
some steps."""
import os

import numpy as np
import pytest
import torch

from oracle import optim as ooptim

HERE = os.path.dirname(os.path.abspath(__file__))


def _load():
    return dict(np.load(os.path.join(HERE, 'golden', 'optim_golden.npz')))


@pytest.mark.parametrize('tag', ['adam_wd0.0', 'adam_wd0.01', 'adam_larc_wd0.001', 'sgd_larc_wd0.001', 'sgd_wd0.01'])
def test_flat_optimiser_matches_torch_and_reference_larc(tag):
    z = _load()
    sizes = [int(x) for x in z['shapes']]
    offs = np.concatenate([[0], np.cumsum(sizes)])
    segs = [(int(offs[k]), sizes[k]) for k in range(len(sizes))]
    lr, wd = float(z[tag + '/lr']), float(z[tag + '/wd'])
    p = torch.from_numpy(z[tag + '/init'].copy())
    m, v = torch.zeros_like(p), torch.zeros_like(p)
    steps = torch.zeros(len(segs), dtype=torch.int64)
    seen = torch.zeros(len(segs), dtype=torch.bool)
    for step in range(int(z['steps'])):
        g = torch.from_numpy(z['{}/grad{}'.format(tag, step)].copy())
        present = [bool(x) for x in z['{}/present{}'.format(tag, step)]]
        eff_wd = wd
        if tag.split('_wd')[0].endswith('larc'):
            ooptim.larc_adjust(p, g, present, segs, lr, wd)
            eff_wd = 0.0
        if tag.startswith('adam'):
            ooptim.adam_step(p, g, m, v, steps, present, segs, lr, weight_decay=eff_wd)
        else:
            ooptim.sgd_step(p, g, m, seen, present, segs, lr, momentum=0.9, weight_decay=eff_wd)
        want = z['{}/param{}'.format(tag, step)]
        np.testing.assert_allclose(p.numpy(), want, rtol=2e-6, atol=2e-7, err_msg='{} step {}'.format(tag, step))
    # a tensor that was absent keeps its value and its step count
    assert int(steps[2]) in (0, int(z['steps']) - 2)


def test_segment_presence_matches_autograd_none_pattern():
    """Host logic of the segment-aware optimiser: which parameter tensors count as present for a minibatch must be
    exactly the tensors torch autograd populates in the reference's graph (all others keep .grad None and are skipped
    by torch.optim).  Ground truth: autograd through the oracle's restatement of _loss on reference-named parameters."""
    from oracle import network as onet
    from pyprob_b200.encoding import EncodedBatch, SubBatch
    from pyprob_b200.network import InferenceNetworkLSTM
    from tests import netfixture
    fx = netfixture.load('mixed')
    fam_of = {}
    for sb in fx['subs']:
        for a, f in zip(sb['addresses'], sb['families']):
            fam_of[a] = f
    net = InferenceNetworkLSTM(model=None, observe_embeddings={n: {} for n in fx['observe_names']})
    net._addresses = {a: {'id': i, 'type': fam_of[a]} for i, a in enumerate(fx['address_order'])}
    names = sorted(fx['params'])
    net._seg = {'names': names}
    net._skip_absent_gradients = True
    ids = {a: i for i, a in enumerate(fx['address_order'])}
    checked_absent = 0
    for chosen in ([0], [1], list(range(len(fx['subs'])))):
        subs = [fx['subs'][i] for i in chosen]
        p = {k: v.clone().float().requires_grad_(True) for k, v in fx['params'].items()}
        value, _ = onet.loss(p, subs, fx['observe_names'], fx['observe_in_dims'], fx['K'])
        value.backward()
        want = np.asarray([int(p[n].grad is not None) for n in names])
        enc = EncodedBatch([SubBatch([ids[a] for a in sb['addresses']], sb['values'].numpy(), sb['prior0'].numpy(),
                                     sb['prior1'].numpy(), sb['obs'].numpy()) for sb in subs])
        got = net._segment_presence(enc)
        assert got.tolist() == want.tolist(), [n for n, a, b in zip(names, got, want) if a != b]
        checked_absent += int((want == 0).sum())
    assert checked_absent > 0   # the single-sub-batch cases really leave some tensors without a gradient
    net._skip_absent_gradients = False
    assert net._segment_presence(enc).min() == 1


def test_flat_adam_hands_over_to_the_skipping_step_at_the_right_moment(monkeypatch):
    """Host logic of Optimizer.ADAM (network._maybe_switch_to_segmented): the flat kernel is kept exactly as long as it is
    indistinguishable from torch.optim's skipping of grad-None tensors — constant set of untouched tensors AND zero weight
    decay — and the hand-over initialises every tensor's step count with what the history implies."""
    import torch
    from pyprob_b200.network import InferenceNetworkLSTM
    from pyprob_b200.util import Optimizer
    names = ['a', 'b', 'c', 'd']

    def make(weight_decay):
        net = InferenceNetworkLSTM(model=None, observe_embeddings={'o': {}})
        net._optimizer_type, net._weight_decay, net._last_enc = Optimizer.ADAM, weight_decay, object()
        net._present_sig, net._seg, net._optimizer_step, net._auto_skip_absent = None, None, 0, True
        monkeypatch.setattr(net, '_segment_names', lambda: names)

        def create():
            net._seg = {'names': names, 'steps': torch.zeros(len(names), dtype=torch.int64)}
        monkeypatch.setattr(net, '_create_segment_state', create)
        return net

    def run(net, patterns):
        switched = None
        for it, pat in enumerate(patterns):
            monkeypatch.setattr(net, '_segment_presence', lambda enc, force=False, pat=pat: np.asarray(pat, dtype=np.int32))
            net._maybe_switch_to_segmented()
            if net._seg is not None and switched is None:
                switched = (it, net._seg['steps'].tolist())
            net._optimizer_step += 1
        return switched

    # constant pattern, no weight decay: never switches, even though tensor d is never touched
    assert run(make(0.0), [[1, 1, 1, 0]] * 5) is None
    # pattern changes at step 3 (0-based): tensors touched so far took 3 steps, the never-touched one none
    assert run(make(0.0), [[1, 1, 1, 0]] * 3 + [[1, 0, 1, 0], [1, 1, 1, 1]]) == (3, [3, 3, 3, 0])
    # weight decay with an untouched tensor: the flat kernel would decay it, torch would not -> switch before step 1
    assert run(make(1e-2), [[1, 1, 1, 0]] * 2) == (0, [0, 0, 0, 0])
    # weight decay but everything touched: no switch
    assert run(make(1e-2), [[1, 1, 1, 1]] * 3) is None
