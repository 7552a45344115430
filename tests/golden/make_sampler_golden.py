"""Golden minibatch index sequences from the UNMODIFIED reference samplers (pyprob/nn/dataset.py:312-400).

    python tests/golden/make_sampler_golden.py      # writes tests/golden/sampler_golden.json

Runs in the build container only (needs /root/reference and the stubs of oracle/ref_stubs).  The samplers only
read ``offline_dataset._sorted_indices`` and torch.distributed's world size / rank, so a bare OfflineDataset object
carrying a list of indices and a patched world size / rank are enough — no trace files are involved.
"""
import json
import os
import random
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, '/root/reference')
sys.path.insert(0, os.path.join(ROOT, 'oracle', 'ref_stubs'))

import numpy as np  # noqa: E402
import torch.distributed as dist  # noqa: E402

import pyprob  # noqa: E402,F401  (the reference)
from pyprob.nn import dataset as ref  # noqa: E402


def fake_dataset(sorted_indices):
    ds = ref.OfflineDataset.__new__(ref.OfflineDataset)
    ds._sorted_indices = list(sorted_indices)
    ds.cumulative_sizes = [len(sorted_indices)]   # ConcatDataset.__len__
    return ds


def run_case(n, batch_size, world, num_buckets, shuffle_batches, shuffle_buckets, epochs, perm_seed, np_seed):
    rng = np.random.default_rng(perm_seed)
    sorted_indices = [int(i) for i in rng.permutation(n)]
    case = {'n': n, 'batch_size': batch_size, 'world': world, 'num_buckets': num_buckets,
            'shuffle_batches': shuffle_batches, 'shuffle_buckets': shuffle_buckets, 'epochs': epochs,
            'np_seed': np_seed, 'sorted_indices': sorted_indices, 'ranks': []}
    for rank in range(world):
        dist.get_world_size = lambda *a, **k: world
        dist.get_rank = lambda *a, **k: rank
        dist.is_available = lambda: True
        random.seed(99)            # the sampler must leave Python's generator where it found it
        np.random.seed(np_seed)    # per-rank stream used by shuffle_batches
        import contextlib
        import io
        with contextlib.redirect_stdout(io.StringIO()):
            s = ref.DistributedTraceBatchSampler(fake_dataset(sorted_indices), batch_size, shuffle_batches=shuffle_batches,
                                                 num_buckets=num_buckets, shuffle_buckets=shuffle_buckets)
        out = {'num_batches': len(s), 'epochs': [], 'python_random_after': random.random()}
        for _ in range(epochs):
            out['epochs'].append([[int(i) for i in b] for b in s])
        case['ranks'].append(out)
    return case


def run_single(n, batch_size, shuffle, epochs, perm_seed, np_seed):
    rng = np.random.default_rng(perm_seed)
    sorted_indices = [int(i) for i in rng.permutation(n)]
    np.random.seed(np_seed)
    s = ref.TraceBatchSampler(fake_dataset(sorted_indices), batch_size, shuffle_batches=shuffle)
    return {'n': n, 'batch_size': batch_size, 'shuffle_batches': shuffle, 'epochs': epochs, 'np_seed': np_seed,
            'sorted_indices': sorted_indices, 'num_batches': len(s),
            'yielded': [[[int(i) for i in b] for b in s] for _ in range(epochs)]}


if __name__ == '__main__':
    cases = {'distributed': [
        run_case(1000, 16, 2, None, False, True, 3, 1, 5),
        run_case(1000, 16, 2, 4, True, True, 2, 2, 6),
        run_case(1037, 8, 4, 5, True, True, 2, 3, 7),       # ragged: dropped minibatches, merged last bucket
        run_case(999, 10, 3, 7, False, False, 2, 4, 8),
        run_case(512, 32, 8, 2, True, True, 2, 5, 9),
        run_case(300, 7, 1, 3, True, True, 2, 6, 10),
    ], 'single': [
        run_single(100, 8, True, 3, 11, 12),
        run_single(64, 64, False, 1, 13, 14),
        run_single(10, 3, True, 2, 15, 16),
    ]}
    with open(os.path.join(HERE, 'sampler_golden.json'), 'w') as f:
        json.dump(cases, f, separators=(',', ':'))
    print('wrote sampler_golden.json:', len(cases['distributed']), 'distributed and', len(cases['single']), 'single cases')
