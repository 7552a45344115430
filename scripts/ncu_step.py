"""A few eager training steps of one BASELINE shape for `ncu` (launch lists and --set full captures of the tensor-core
kernels):  python scripts/ncu_step.py gum|s50 [steps]"""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pyprob_b200 import synthetic  # noqa: E402
from pyprob_b200._lib import call, ptr  # noqa: E402
from pyprob_b200.network import BatchStruct  # noqa: E402
from pyprob_b200.util import Optimizer  # noqa: E402

cfg = sys.argv[1] if len(sys.argv) > 1 else 'gum'
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
dev = torch.device('cuda:0')
rng = np.random.default_rng(0)
if cfg == 'gum':
    net = synthetic.gum_network(lstm_dim=512, precision=0)
    batch = synthetic.gum_batch(rng, 256)
else:
    net = synthetic.synthetic50_network(precision=0, T=50)
    batch = synthetic.synthetic50_batch(rng, 512, T=50)
net._optimizer_type, net._learning_rate_init, net._weight_decay = Optimizer.ADAM, 1e-3, 0.0
net._create_optimizer()
net._sync_native()
enc = batch.encode(net)
grad = torch.zeros_like(net._arena.data)
img = torch.from_numpy(enc.pack().copy()).pin_memory()
dimg = img.to(dev)
bs = BatchStruct()
call('ppb_batch_from_image', img.data_ptr(), dimg.data_ptr(), img.numel(), C.byref(bs))
need = net._ensure_workspace(enc)
st = torch.cuda.current_stream().cuda_stream
loss = torch.empty((), device=dev)
status = torch.zeros(1, dtype=torch.int32, device=dev)
hyper = torch.tensor([1e-3, 0.9, 0.999, 1e-8, 0.0, 1.0], dtype=torch.float32, device=dev)
state = torch.zeros(4, dtype=torch.int32, device=dev)
for _ in range(steps):
    grad.zero_()
    call('ppb_ic_loss_forward', net._handle, ptr(net._arena.data), C.byref(bs), ptr(net._workspace), need, 0, ptr(loss),
         ptr(status), None, 1, st)
    call('ppb_ic_loss_backward', net._handle, ptr(net._arena.data), ptr(grad), C.byref(bs), ptr(net._workspace), need, 0, 1.0, st)
    call('ppb_adam_step_dev', ptr(net._arena.data), ptr(grad), ptr(net._exp_avg), ptr(net._exp_avg_sq), net._arena.numel(),
         ptr(hyper), ptr(state), st)
    torch.cuda.synchronize()
print('loss', float(loss))
