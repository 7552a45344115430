"""Offline data end to end on the GPU: Model.save_dataset writes columnar trace files from lock-step prior traces,
OfflineDataset reads them back, learn_inference_network(dataset_dir=...) trains from disk
(reference: tests/test_train.py offline cases, pyprob/nn/dataset.py:121-137, model.py:226-231)."""
import math

import numpy as np
import pytest
import torch

import pyprob_b200 as pyprob
from pyprob_b200 import InferenceEngine, InferenceNetwork, Model
from pyprob_b200.distributions import Normal, Uniform
from pyprob_b200.offline import OfflineDataset

pytestmark = pytest.mark.gpu


class GaussianUnknownMean(Model):
    def forward(self):
        mu = pyprob.sample(Normal(1, math.sqrt(5)))
        likelihood = Normal(mu, math.sqrt(2))
        pyprob.observe(likelihood, name='obs0')
        pyprob.observe(likelihood, name='obs1')
        return mu


class Marsaglia(Model):
    def forward(self):
        def body(s):
            x = pyprob.sample(Uniform(-1, 1))
            y = pyprob.sample(Uniform(-1, 1))
            return {'x': x, 'y': y, 's': x * x + y * y}
        st = pyprob.while_loop(lambda s: s['s'] >= 1, body, {'x': 0.0, 'y': 0.0, 's': 2.0})
        mu = 1 + math.sqrt(5) * (st['x'] * torch.sqrt(-2 * torch.log(st['s']) / st['s']))
        pyprob.observe(Normal(mu, math.sqrt(2)), name='obs0')
        return mu


def test_save_dataset_then_train_from_disk(cuda, tmp_path):
    pyprob.seed(3)
    model = GaussianUnknownMean()
    files = model.save_dataset(str(tmp_path / 'train'), num_traces=2048, num_traces_per_file=512)
    assert len(files) == 4
    model.save_dataset(str(tmp_path / 'valid'), num_traces=256, num_traces_per_file=256)
    ds = OfflineDataset(str(tmp_path / 'train'))
    assert len(ds) == 2048 and ds.num_trace_types == 1 and ds.observe_names == ['obs0', 'obs1']
    assert [a[1] for a in ds.addresses] == ['Normal']
    full = ds.batch(list(range(2048))).subs[0]
    mu, obs = full['values'][0], full['obs']
    assert abs(mu.mean() - 1.0) < 0.25 and abs(mu.std() - math.sqrt(5)) < 0.25
    assert abs((obs - mu[:, None]).std() - math.sqrt(2)) < 0.15
    assert np.all(full['prior0'] == 1.0) and np.allclose(full['prior1'], math.sqrt(5))
    model.learn_inference_network(num_traces=256 * 40, dataset_dir=str(tmp_path / 'train'),
                                  dataset_valid_dir=str(tmp_path / 'valid'), inference_network=InferenceNetwork.LSTM,
                                  observe_embeddings={'obs0': {'dim': 16}, 'obs1': {'dim': 16}}, batch_size=256,
                                  lstm_dim=64, valid_every=2048, learning_rate_init=1e-3)
    net = model._inference_network
    hist = net._history_train_loss
    assert len(hist) == 40 and np.all(np.isfinite(hist)) and np.mean(hist[-5:]) < np.mean(hist[:5])
    assert len(net._history_valid_loss) >= 3 and np.all(np.isfinite(net._history_valid_loss))
    post = model.posterior_results(4096, InferenceEngine.IMPORTANCE_SAMPLING_WITH_INFERENCE_NETWORK,
                                   observe={'obs0': 8, 'obs1': 9})
    assert abs(float(post.mean) - 7.25) < 0.75


def test_save_dataset_keeps_stochastic_control_flow(cuda, tmp_path):
    pyprob.seed(4)
    model = Marsaglia()
    model.save_dataset(str(tmp_path / 'd'), num_traces=1024, num_traces_per_file=512, batch_size=256)
    ds = OfflineDataset(str(tmp_path / 'd'))
    assert len(ds) == 1024 and ds.num_trace_types > 1
    assert np.all(ds.lengths % 2 == 0) and ds.lengths.min() == 2
    assert np.all(np.diff(ds.lengths[ds._sorted_indices]) >= 0)
    batch = ds.batch(list(range(1024)))
    assert batch.size == 1024
    for sb in batch.subs:      # every stored trace is a valid run of the rejection loop: only the last pair accepted
        v = sb['values']
        s = v[0::2] ** 2 + v[1::2] ** 2
        assert np.all(s[-1] < 1) and np.all(s[:-1] >= 1)
        assert np.all(sb['prior0'] == -1.0) and np.all(sb['prior1'] == 1.0)
    model.learn_inference_network(num_traces=256 * 4, dataset_dir=str(tmp_path / 'd'), pre_generate_layers=True,
                                  inference_network=InferenceNetwork.LSTM, observe_embeddings={'obs0': {'dim': 16}},
                                  batch_size=256, lstm_dim=32)
    net = model._inference_network
    assert len(net._addresses) == int(ds.lengths.max()) and np.all(np.isfinite(net._history_train_loss))
