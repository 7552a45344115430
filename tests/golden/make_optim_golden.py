"""Golden optimiser trajectories from torch.optim wrapped in the UNMODIFIED reference LARC class.

    python tests/golden/make_optim_golden.py      # writes tests/golden/optim_golden.npz

Build container only (needs /root/reference + oracle/ref_stubs).  Several parameter tensors, some of which receive
no gradient on some steps (as the proposal layers of addresses absent from a minibatch do), stepped with
Adam / Adam+LARC / SGD(nesterov)+LARC exactly as pyprob/nn/inference_network.py:343-355 builds them.
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, '/root/reference')
sys.path.insert(0, os.path.join(ROOT, 'oracle', 'ref_stubs'))

import numpy as np  # noqa: E402
import torch  # noqa: E402
from torch import optim  # noqa: E402

import pyprob  # noqa: E402,F401
from pyprob.nn.optimizer_larc import LARC  # noqa: E402

SHAPES = [(7, 5), (5,), (3, 4), (4,), (16,), (2, 2)]
STEPS = 6


def presence(step, k):
    if k in (2, 3):
        return step >= 2          # an address that first appears at step 2 (pre-generated layers)
    if k == 5:
        return step % 2 == 0      # a rare address
    return True


def run(kind, lr, wd, seed):
    g = torch.Generator().manual_seed(seed)
    params = [torch.nn.Parameter(torch.randn(*s, generator=g)) for s in SHAPES]
    if kind.startswith('adam'):
        opt = optim.Adam(params, lr=lr, weight_decay=wd)
    else:
        opt = optim.SGD(params, lr=lr, momentum=0.9, nesterov=True, weight_decay=wd)
    if kind.endswith('larc'):
        opt = LARC(opt)
    out = {'init': np.concatenate([p.detach().numpy().reshape(-1) for p in params])}
    for step in range(STEPS):
        opt.zero_grad()   # set_to_none on torch >= 2.0: absent tensors keep .grad None
        grads = []
        for k, p in enumerate(params):
            gk = torch.randn(*SHAPES[k], generator=g) * (0.1 + k)
            if presence(step, k):
                p.grad = gk.clone()
            else:
                p.grad = None
            grads.append(gk)
        opt.step()
        out['grad{}'.format(step)] = np.concatenate([x.numpy().reshape(-1) for x in grads])
        out['present{}'.format(step)] = np.asarray([presence(step, k) for k in range(len(SHAPES))])
        out['param{}'.format(step)] = np.concatenate([p.detach().numpy().reshape(-1) for p in params])
    return out


if __name__ == '__main__':
    fx = {'shapes': np.asarray([int(np.prod(s)) for s in SHAPES]), 'steps': np.asarray(STEPS),
          'torch_version': np.asarray(torch.__version__)}
    for kind, lr, wd in (('adam', 1e-2, 0.0), ('adam', 1e-2, 1e-2), ('adam_larc', 1e-2, 1e-3), ('sgd_larc', 5e-2, 1e-3),
                         ('sgd', 5e-2, 1e-2)):
        tag = '{}_wd{}'.format(kind, wd)
        for k, v in run(kind, lr, wd, seed=17).items():
            fx['{}/{}'.format(tag, k)] = v
        fx[tag + '/lr'] = np.asarray(lr)
        fx[tag + '/wd'] = np.asarray(wd)
    np.savez_compressed(os.path.join(HERE, 'optim_golden.npz'), **fx)
    print('wrote optim_golden.npz with', len(fx), 'arrays (torch {})'.format(torch.__version__))
