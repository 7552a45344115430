"""ppb_ic_train_step_host — the host-buffer training step (pinned batch image in, loss out; what bench.py's `e2e` times):
it must equal forward + backward + Adam issued separately, call after call, also when the library replays the step from its
cached CUDA graph (PPB_HOST_STEP_GRAPH=1: first call eager, second captured, later ones replayed)."""
import ctypes as C

import numpy as np
import pytest
import torch

from pyprob_b200 import synthetic
from pyprob_b200._lib import call, ptr
from pyprob_b200.util import Optimizer

pytestmark = pytest.mark.gpu


def _setup(seed):
    net = synthetic.gum_network(lstm_dim=64, precision=0, seed=seed)
    net._optimizer_type, net._learning_rate_init, net._weight_decay = Optimizer.ADAM, 1e-3, 0.0
    net._create_optimizer()
    net._sync_native()
    return net


def test_host_step_matches_separate_calls(cuda):
    rng = np.random.default_rng(3)
    batches = [synthetic.gum_batch(rng, 200) for _ in range(3)]
    ref = _setup(5)
    net = _setup(5)
    assert torch.equal(ref._arena.data, net._arena.data)
    encs = [b.encode(net) for b in batches]
    hosts = [torch.from_numpy(e.pack().copy()).pin_memory() for e in encs]
    img_dev = torch.empty_like(hosts[0], device=cuda)
    need = net._ensure_workspace(encs[0])
    grad = torch.zeros_like(net._arena.data)
    loss_host = torch.zeros(1).pin_memory()
    status_host = torch.zeros(1, dtype=torch.int32).pin_memory()
    st = torch.cuda.current_stream().cuda_stream
    for step in range(1, 7):
        i = (step - 1) % 3
        # reference: the public pieces one by one
        ref._arena.grad = None
        ok, loss = ref._loss(batches[i])
        assert ok
        loss.backward()
        ref._learning_rate = 1e-3
        ref.optimizer_step()
        # host step
        call('ppb_ic_train_step_host', net._handle, ptr(net._arena.data), ptr(grad), ptr(net._exp_avg), ptr(net._exp_avg_sq),
             net._arena.numel(), hosts[i].data_ptr(), hosts[i].numel(), ptr(img_dev), ptr(net._workspace),
             net._workspace.numel(), 0, 1e-3, 0.9, 0.999, 1e-8, 0.0, step, loss_host.data_ptr(), status_host.data_ptr(), st)
        assert int(status_host[0]) == 0
        assert abs(float(loss_host[0]) - float(loss.detach())) <= 1e-6 * abs(float(loss.detach()))
        torch.testing.assert_close(net._arena.data, ref._arena.data, rtol=1e-6, atol=1e-7,
                                   msg=lambda m, step=step: 'step {}: {}'.format(step, m))
