"""Host-side throughput of the offline trace store (CPU only): write, open + sort, and minibatch assembly for the
synthetic 50-address model (BASELINE configs[3] shape).  The reference's path for the same step is one sqlite lookup
+ zlib decompress + unpickle of a Python Trace object per trace (pyprob/nn/dataset.py:140-171)."""
import json
import os
import sys
import tempfile
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pyprob_b200 import offline, synthetic  # noqa: E402


def main():
    n_files, per_file, T, bs = 8, 16384, 50, 512
    rng = np.random.default_rng(0)
    d = tempfile.mkdtemp()
    t0 = time.perf_counter()
    nbytes = 0
    for _ in range(n_files):
        sb = synthetic.synthetic50_batch(rng, per_file, T).subs[0]
        name = offline.save_columns(d, offline.TraceColumns.from_sub_batches([sb], ['obs'], [1]))
        nbytes += os.path.getsize(name)
    t_write = time.perf_counter() - t0
    t0 = time.perf_counter()
    ds = offline.OfflineDataset(d)
    t_open = time.perf_counter() - t0
    np.random.seed(0)
    for _ in range(5):
        ds.next_batch(bs)
    t0 = time.perf_counter()
    steps = 200
    for _ in range(steps):
        ds.next_batch(bs)
    t_batch = (time.perf_counter() - t0) / steps
    n = n_files * per_file
    print(json.dumps({'traces': n, 'steps_per_trace': T, 'bytes_on_disk': nbytes, 'bytes_per_trace': nbytes / n,
                      'write_traces_per_s': n / t_write, 'open_and_sort_s': t_open,
                      'minibatch_512_ms': t_batch * 1e3, 'assembly_traces_per_s': bs / t_batch}))


if __name__ == '__main__':
    main()
