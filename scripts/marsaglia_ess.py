"""Spread of the Marsaglia IC acceptance statistic (ESS of 8192 proposals) over seeds and training budgets."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import torch
import pyprob_b200 as pyprob
from pyprob_b200 import InferenceEngine, InferenceNetwork
from test_model_gpu import GaussianUnknownMeanMarsaglia

pyprob.set_verbosity(0)
for name, kw in (('300k/512/h128', dict(num_traces=300000, batch_size=512, lstm_dim=128)),
                 ('600k/256/h128', dict(num_traces=600000, batch_size=256, lstm_dim=128)),
                 ('300k/128/h128', dict(num_traces=300000, batch_size=128, lstm_dim=128))):
    for seed in (5, 6, 7):
        pyprob.seed(seed)
        m = GaussianUnknownMeanMarsaglia()
        t0 = time.time()
        m.learn_inference_network(inference_network=InferenceNetwork.LSTM,
                                  observe_embeddings={'obs0': {'dim': 16}, 'obs1': {'dim': 16}}, **kw)
        t1 = time.time()
        post = m.posterior_results(8192, InferenceEngine.IMPORTANCE_SAMPLING_WITH_INFERENCE_NETWORK, observe={'obs0': 8, 'obs1': 9})
        print('%-16s seed %d: train %.1f s, ESS %.1f (floor %.1f), mean %.3f, loss %.4f' % (
            name, seed, t1 - t0, post.effective_sample_size, 0.016 * 8192, float(post.mean), m._inference_network._loss_min), flush=True)
