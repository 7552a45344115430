"""Helpers shared by the network tests: load the reference-generated fixture (tests/golden/network_golden.npz)."""
import os

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_cache = {}


def load(tag):
    if 'npz' not in _cache:
        _cache['npz'] = dict(np.load(os.path.join(ROOT, 'tests', 'golden', 'network_golden.npz')))
    z = _cache['npz']
    pre = tag + '/'
    fx = {'loss': float(z[pre + 'loss']),
          'observe_names': [str(x) for x in z[pre + 'observe_names']],
          'observe_in_dims': [int(x) for x in z[pre + 'observe_in_dims']],
          'lstm_dim': int(z[pre + 'dims'][0]), 'K': int(z[pre + 'dims'][1]), 'batch_size': int(z[pre + 'dims'][2]),
          'address_order': [str(x) for x in z[pre + 'address_order']],
          'type_order': [str(x) for x in z[pre + 'type_order']],
          'params': {k[len(pre + 'param/'):]: torch.from_numpy(v) for k, v in z.items() if k.startswith(pre + 'param/')},
          'grads': {k[len(pre + 'grad/'):]: torch.from_numpy(v) for k, v in z.items() if k.startswith(pre + 'grad/')},
          'subs': []}
    for s in range(int(z[pre + 'num_sub'])):
        p = '{}sub{}/'.format(pre, s)
        fx['subs'].append({'addresses': [str(x) for x in z[p + 'addresses']],
                           'families': [str(x) for x in z[p + 'families']],
                           'num_categories': [int(x) for x in z[p + 'num_categories']],
                           'values': torch.from_numpy(z[p + 'values']), 'prior0': torch.from_numpy(z[p + 'prior0']),
                           'prior1': torch.from_numpy(z[p + 'prior1']), 'obs': torch.from_numpy(z[p + 'obs'])})
    return fx
