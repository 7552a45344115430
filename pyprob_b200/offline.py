"""Offline trace datasets: a columnar, pre-encoded on-disk format and the batch samplers that walk it.

This is SURVEY.md 8(f) row 3 — the step *before* the training path in offline mode.  The reference stores one
pickled + zlib-compressed ``Trace`` object per key in sqlite shelves (pyprob/nn/dataset.py:121-137,
pyprob/util.py:347-355), re-hashes every trace to sort the data set (:232-247) and unpickles ``batch_size`` Python
objects per minibatch (:140-171).  Here a file holds *columns* — the same flat value arrays the encoder feeds to
the CUDA path (encoding.py:SubBatch) — so that a minibatch is a handful of numpy gathers on memory-mapped arrays:

    magic 'PPBTRC01' | uint64 header bytes | JSON header | 64-byte aligned little-endian arrays (offsets in the
    header are relative to the header end rounded up to 64 bytes)
        trace_start int64[n+1]   first step of every trace (steps = controlled sample statements)
        step_addr   int32[S]     file-local address id of every step (table in the header)
        values      float32[S]   sampled value (category index for Categorical)
        prior0/1    float32[S]   prior mean/stddev (Normal) or low/high (Uniform), zero otherwise
        obs         float32[n, sum(observe_dims)]

Semantics kept from the reference:
  * data set = concatenation of files in name order; ``pyprob_b200_traces_sorted_*`` files win over unsorted ones
    (dataset.py:174-182);
  * sorted order = by controlled trace length, traces of one trace type (address sequence) adjacent
    (dataset.py:227-247; the reference's tie-break among types of equal length is a per-process string hash);
  * ``TraceBatchSampler`` / ``DistributedTraceBatchSampler`` yield exactly the index lists the reference's samplers
    yield for the same sorted indices and the same RNG state (dataset.py:312-400) — pinned by
    tests/golden/sampler_golden.json, generated from the reference classes themselves.
"""
import glob
import json
import math
import os
import random
import uuid

import numpy as np

MAGIC = b'PPBTRC01'
_ALIGN = 64
_COLUMNS = (('trace_start', np.int64), ('step_addr', np.int32), ('values', np.float32), ('prior0', np.float32),
            ('prior1', np.float32), ('obs', np.float32))


class TraceColumns:
    """A chunk of traces in columnar form (in memory or memory-mapped)."""

    def __init__(self, addresses, observe_names, observe_dims, trace_start, step_addr, values, prior0, prior1, obs,
                 sorted_on_disk=False):
        self.addresses = [(str(a), str(f), int(c)) for a, f, c in addresses]
        self.observe_names = [str(x) for x in observe_names]
        self.observe_dims = [int(x) for x in observe_dims]
        self.trace_start, self.step_addr = trace_start, step_addr
        self.values, self.prior0, self.prior1 = values, prior0, prior1
        self.obs = obs.reshape(len(trace_start) - 1, sum(self.observe_dims))
        self.sorted_on_disk = bool(sorted_on_disk)
        n_steps = int(trace_start[-1]) if len(trace_start) else 0
        if not (len(step_addr) == len(values) == len(prior0) == len(prior1) == n_steps):
            raise ValueError('inconsistent trace columns')

    def __len__(self):
        return len(self.trace_start) - 1

    @property
    def lengths(self):
        return np.diff(self.trace_start)

    # ---- builders ----------------------------------------------------------------------------------------
    @staticmethod
    def from_sub_batches(subs, observe_names, observe_dims):
        """subs: the dicts ArrayBatch takes (addresses, families, num_categories, values[T,B], prior0, prior1,
        obs[B,D]).  Traces are laid out sub-batch after sub-batch, column after column."""
        table, index = [], {}
        starts, addr, vals, p0s, p1s, obs = [0], [], [], [], [], []
        for sb in subs:
            ids = []
            for a, f, c in zip(sb['addresses'], sb['families'], sb['num_categories']):
                key = str(a)
                if key not in index:
                    index[key] = len(table)
                    table.append((key, str(f), int(c)))
                ids.append(index[key])
            v = np.asarray(sb['values'], np.float32)
            T, B = v.shape
            if T == 0:
                raise ValueError('Trace of length zero.')
            addr.append(np.tile(np.asarray(ids, np.int32), B))
            vals.append(v.T.reshape(-1))
            p0s.append(np.asarray(sb['prior0'], np.float32).T.reshape(-1))
            p1s.append(np.asarray(sb['prior1'], np.float32).T.reshape(-1))
            obs.append(np.asarray(sb['obs'], np.float32).reshape(B, -1))
            base = starts[-1]
            starts.extend(base + T * (i + 1) for i in range(B))
        return TraceColumns(table, observe_names, observe_dims, np.asarray(starts, np.int64), np.concatenate(addr),
                            np.concatenate(vals), np.concatenate(p0s), np.concatenate(p1s), np.concatenate(obs, 0))

    @staticmethod
    def from_reference_traces(traces, observe_names):
        """Bridge from the reference's per-particle ``Trace`` objects (duck-typed: ``variables_controlled`` with
        ``address``, ``value``, ``distribution``; ``named_variables[name].value``) — what its shelves hold."""
        table, index = [], {}
        starts, addr, vals, p0s, p1s, obs = [0], [], [], [], [], []
        dims = None
        for tr in traces:
            ctrl = list(tr.variables_controlled)
            if len(ctrl) == 0:
                raise ValueError('Trace of length zero.')
            for var in ctrl:
                d = var.distribution
                name = type(d).__name__
                c = int(getattr(d, 'num_categories', 0)) if name == 'Categorical' else 0
                if var.address not in index:
                    index[var.address] = len(table)
                    table.append((var.address, name, c))
                addr.append(index[var.address])
                vals.append(float(var.value))
                if name == 'Normal':
                    p0s.append(float(d.mean)); p1s.append(float(d.stddev))
                elif name == 'Uniform':
                    p0s.append(float(d.low)); p1s.append(float(d.high))
                else:
                    p0s.append(0.0); p1s.append(0.0)
            row = [np.asarray(tr.named_variables[nm].value, np.float32).reshape(-1) for nm in observe_names]
            if dims is None:
                dims = [int(r.size) for r in row]
            obs.append(np.concatenate(row))
            starts.append(starts[-1] + len(ctrl))
        return TraceColumns(table, observe_names, dims or [], np.asarray(starts, np.int64), np.asarray(addr, np.int32),
                            np.asarray(vals, np.float32), np.asarray(p0s, np.float32), np.asarray(p1s, np.float32),
                            np.stack(obs, 0) if obs else np.zeros((0, 0), np.float32))

    def select(self, indices):
        """New in-memory chunk holding the given traces in the given order."""
        indices = np.asarray(indices, np.int64)
        lens = self.lengths[indices]
        starts = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
        steps = _expand(self.trace_start[indices], lens)
        return TraceColumns(self.addresses, self.observe_names, self.observe_dims, starts, self.step_addr[steps],
                            self.values[steps], self.prior0[steps], self.prior1[steps], self.obs[indices],
                            self.sorted_on_disk)

    # ---- file I/O ----------------------------------------------------------------------------------------
    def save(self, path, sorted_on_disk=False):
        cols = {'trace_start': self.trace_start, 'step_addr': self.step_addr, 'values': self.values,
                'prior0': self.prior0, 'prior1': self.prior1, 'obs': self.obs.reshape(-1)}
        header = {'version': 1, 'num_traces': len(self), 'num_steps': int(self.trace_start[-1]),
                  'observe_names': self.observe_names, 'observe_dims': self.observe_dims,
                  'addresses': [list(a) for a in self.addresses], 'sorted': bool(sorted_on_disk), 'arrays': {}}
        off = 0   # array offsets are relative to the data base = the header end rounded up to 64 bytes
        for name, dt in _COLUMNS:
            count = int(np.asarray(cols[name]).size)
            header['arrays'][name] = [np.dtype(dt).str, off, count]
            off = _round_up(off + count * np.dtype(dt).itemsize, _ALIGN)
        blob = json.dumps(header).encode('utf-8')
        base = _round_up(16 + len(blob), _ALIGN)
        tmp = path + '.tmp'
        with open(tmp, 'wb') as f:
            f.write(MAGIC)
            f.write(np.uint64(len(blob)).tobytes())
            f.write(blob)
            for name, dt in _COLUMNS:
                f.seek(base + header['arrays'][name][1])
                f.write(np.ascontiguousarray(cols[name], dtype=dt).tobytes())
            f.truncate(base + off)
        os.replace(tmp, path)

    @staticmethod
    def load(path):
        with open(path, 'rb') as f:
            if f.read(8) != MAGIC:
                raise ValueError('not a pyprob_b200 trace file: {}'.format(path))
            (hlen,) = np.frombuffer(f.read(8), np.uint64)
            header = json.loads(f.read(int(hlen)).decode('utf-8'))
        if header.get('version') != 1:
            raise ValueError('unsupported trace file version in {}'.format(path))
        size = os.path.getsize(path)
        base = _round_up(16 + int(hlen), _ALIGN)
        arrs = {}
        for name, dt in _COLUMNS:
            dts, off, count = header['arrays'][name]
            if np.dtype(dts) != np.dtype(dt) or base + off + count * np.dtype(dt).itemsize > size:
                raise ValueError('corrupt trace file: {}'.format(path))
            arrs[name] = np.memmap(path, dtype=dt, mode='r', offset=base + off, shape=(count,)) if count \
                else np.zeros(0, dt)
        if len(arrs['trace_start']) != header['num_traces'] + 1:
            raise ValueError('corrupt trace file: {}'.format(path))
        return TraceColumns(header['addresses'], header['observe_names'], header['observe_dims'], arrs['trace_start'],
                            arrs['step_addr'], arrs['values'], arrs['prior0'], arrs['prior1'], arrs['obs'],
                            header['sorted'])


def _round_up(x, a):
    return (x + a - 1) // a * a


def _expand(starts, lens):
    """Concatenation of arange(s, s+l) for every (s, l) — the step indices of a list of traces."""
    lens = np.asarray(lens, np.int64)
    total = int(lens.sum())
    if total == 0:
        return np.zeros(0, np.int64)
    first = np.repeat(np.asarray(starts, np.int64), lens)
    local = np.arange(total, dtype=np.int64) - np.repeat(np.concatenate([[0], np.cumsum(lens)[:-1]]), lens)
    return first + local


class _ExampleTrace:
    """Just enough of a trace for InferenceNetworkLSTM._init_layers_observe_embedding."""

    def __init__(self, names, dims):
        self.named_variables = {k: k for k in names}
        self._shapes = {k: ((d,) if d > 1 else ()) for k, d in zip(names, dims)}

    def value_shape(self, variable):
        return self._shapes[variable]


class OfflineDataset:
    """All trace files of a directory as one indexable data set (reference: dataset.py:173-247)."""
    PREFIX = 'pyprob_b200_traces_'

    def __init__(self, dataset_dir, verbose=False):
        self._dataset_dir = dataset_dir
        files = sorted(glob.glob(os.path.join(dataset_dir, self.PREFIX + 'sorted_*')))
        self._sorted_on_disk = len(files) > 0
        if not files:
            files = sorted(glob.glob(os.path.join(dataset_dir, self.PREFIX + '*')))
        files = [f for f in files if not f.endswith('.tmp')]
        if not files:
            raise RuntimeError('Cannot find any data set files at {}'.format(dataset_dir))
        self._files, self._chunks = [], []
        for f in files:
            try:
                self._chunks.append(TraceColumns.load(f))
                self._files.append(f)
            except Exception as e:  # same policy as the reference: warn and omit (dataset.py:187-193)
                print('Dataset file potentially corrupt, omitting: {} ({})'.format(f, e))
        if not self._chunks:
            raise RuntimeError('Cannot find any readable data set files at {}'.format(dataset_dir))
        first = self._chunks[0]
        self.observe_names, self.observe_dims = first.observe_names, first.observe_dims
        for c in self._chunks:
            if c.observe_names != self.observe_names or c.observe_dims != self.observe_dims:
                raise RuntimeError('data set files disagree on the observables')
        # global address table (first-seen order over files) and per-file id translation
        self.addresses, index = [], {}
        self._addr_map = []
        for c in self._chunks:
            m = np.zeros(len(c.addresses), np.int32)
            for i, a in enumerate(c.addresses):
                if a[0] not in index:
                    index[a[0]] = len(self.addresses)
                    self.addresses.append(a)
                elif self.addresses[index[a[0]]] != a:
                    raise RuntimeError('address {} has different distributions in different files'.format(a[0]))
                m[i] = index[a[0]]
            self._addr_map.append(m)
        counts = np.asarray([len(c) for c in self._chunks], np.int64)
        self._file_first = np.concatenate([[0], np.cumsum(counts)])
        self._length = int(counts.sum())
        self.lengths = np.concatenate([c.lengths for c in self._chunks]).astype(np.int64)
        self._trace_type = self._compute_trace_types()
        if self._sorted_on_disk:
            self._sorted_indices = list(range(self._length))
        else:
            order = np.lexsort((self._trace_type, self.lengths))   # by length, then trace type; stable
            self._sorted_indices = [int(i) for i in order]
        self._sampler_iter = None
        self._sampler_batch_size = None
        self._epoch = 0
        self.current_bucket_id = None
        self.num_buckets = None                      # DistributedTraceBatchSampler's num_buckets (None = its default)
        self._obs_columns = None                     # column selection of select_observables (None = all stored)
        self._stored_observe_names, self._stored_observe_dims = list(self.observe_names), list(self.observe_dims)
        if verbose:
            print('OfflineDataset at: {}'.format(dataset_dir))
            print('Num. traces      : {:,}'.format(self._length))
            print('Sorted on disk   : {}'.format(self._sorted_on_disk))
            print('Num. trace types : {:,}'.format(int(self._trace_type.max()) + 1 if self._length else 0))

    def __len__(self):
        return self._length

    def _compute_trace_types(self):
        """Trace type id (first-seen order) of every trace = id of its controlled address sequence, the key the
        reference's Batch groups by (dataset.py:32)."""
        types, out = {}, np.zeros(self._length, np.int64)
        g = 0
        for c, m in zip(self._chunks, self._addr_map):
            seq = m[np.asarray(c.step_addr)] if len(c.step_addr) else np.zeros(0, np.int32)
            ts = np.asarray(c.trace_start)
            for i in range(len(c)):
                key = seq[ts[i]:ts[i + 1]].tobytes()
                t = types.get(key)
                if t is None:
                    t = types[key] = len(types)
                out[g] = t
                g += 1
        self.num_trace_types = len(types)
        return out

    def _locate(self, indices):
        indices = np.asarray(indices, np.int64)
        if indices.size and (indices.min() < 0 or indices.max() >= self._length):
            raise IndexError('trace index out of range')
        f = np.searchsorted(self._file_first, indices, side='right') - 1
        return f, indices - self._file_first[f]

    def columns(self, indices):
        """The given traces (in the given order) as one in-memory TraceColumns with global address ids."""
        f, local = self._locate(indices)
        n = len(local)
        lens = self.lengths[np.asarray(indices, np.int64)] if n else np.zeros(0, np.int64)
        starts = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
        S = int(starts[-1])
        addr = np.zeros(S, np.int32)
        vals, p0, p1 = np.zeros(S, np.float32), np.zeros(S, np.float32), np.zeros(S, np.float32)
        obs = np.zeros((n, sum(self._stored_observe_dims)), np.float32)
        for fi in np.unique(f):
            c = self._chunks[fi]
            pos = np.nonzero(f == fi)[0]
            src = _expand(np.asarray(c.trace_start)[local[pos]], lens[pos])
            dst = _expand(starts[pos], lens[pos])
            addr[dst] = self._addr_map[fi][np.asarray(c.step_addr)[src]]
            vals[dst] = np.asarray(c.values)[src]
            p0[dst] = np.asarray(c.prior0)[src]
            p1[dst] = np.asarray(c.prior1)[src]
            obs[pos] = np.asarray(c.obs)[local[pos]]
        return TraceColumns(self.addresses, self._stored_observe_names, self._stored_observe_dims, starts, addr, vals,
                            p0, p1, obs)

    def select_observables(self, names):
        """Restrict / reorder the observation columns of every minibatch to the given observable names (the keys of
        ``observe_embeddings``): the files may hold more named variables than the network embeds."""
        offs = np.concatenate([[0], np.cumsum(self._stored_observe_dims)])
        cols, dims = [], []
        for nm in names:
            if nm not in self._stored_observe_names:
                raise KeyError('observable {} is not stored in the data set (stored: {})'.format(
                    nm, self._stored_observe_names))
            j = self._stored_observe_names.index(nm)
            cols.extend(range(int(offs[j]), int(offs[j + 1])))
            dims.append(self._stored_observe_dims[j])
        self._obs_columns = np.asarray(cols, np.int64)
        self.observe_names, self.observe_dims = [str(n) for n in names], dims

    def batch(self, indices):
        """Minibatch of the given traces -> ArrayBatch; sub-batches in order of first appearance of each trace
        type, traces inside a sub-batch in the given order (reference Batch, dataset.py:21-37)."""
        from .synthetic import ArrayBatch
        cols = self.columns(indices)
        types = self._trace_type[np.asarray(indices, np.int64)]
        _, first_pos = np.unique(types, return_index=True)
        subs = []
        for p in np.sort(first_pos):
            members = np.nonzero(types == types[p])[0]
            T = int(cols.trace_start[p + 1] - cols.trace_start[p])
            steps = np.asarray(cols.trace_start)[members][:, None] + np.arange(T)[None, :]   # [B, T]
            ids = cols.step_addr[steps[0]]
            subs.append({'addresses': [self.addresses[i][0] for i in ids],
                         'families': [self.addresses[i][1] for i in ids],
                         'num_categories': [self.addresses[i][2] for i in ids],
                         'values': np.ascontiguousarray(cols.values[steps].T),
                         'prior0': np.ascontiguousarray(cols.prior0[steps].T),
                         'prior1': np.ascontiguousarray(cols.prior1[steps].T),
                         'obs': cols.obs[members] if self._obs_columns is None
                         else cols.obs[members][:, self._obs_columns]})
        return ArrayBatch(subs)

    def example_trace(self):
        return _ExampleTrace(self.observe_names, self.observe_dims)

    def address_signature(self):
        """(address, distribution name, num_categories) of every address in the data set, in first-seen order —
        what a pass of InferenceNetwork._polymorph over the whole data set would discover
        (inference_network.py:270-288, _pre_generate_layers)."""
        return list(self.addresses)

    # ---- the training loop's interface (same as dataset.OnlineDataset) -----------------------------------------
    def _make_sampler(self, batch_size):
        from .parallel import world_info
        world, rank = world_info()
        if world > 1:
            return DistributedTraceBatchSampler(self, batch_size, num_buckets=self.num_buckets, world=world, rank=rank)
        return TraceBatchSampler(self, batch_size)

    def epoch_batches(self, batch_size):
        """One pass over the data set in sampler order (what iterating the reference's DataLoader once yields);
        used for the validation loss (inference_network.py:536-542)."""
        for indices in self._make_sampler(batch_size):
            yield self.batch(indices)

    def num_batches(self, batch_size):
        return len(self._make_sampler(batch_size))

    def next_batch(self, batch_size):
        """Next minibatch of an endless epoch loop.  Single process: TraceBatchSampler; under torch.distributed
        with world > 1: DistributedTraceBatchSampler (reference inference_network.py:398-402)."""
        if self._sampler_iter is None or self._sampler_batch_size != batch_size:
            self._sampler = self._make_sampler(batch_size)
            self._sampler_batch_size = batch_size
            self._sampler_iter = iter(self._sampler)
        try:
            indices = next(self._sampler_iter)
        except StopIteration:
            self._epoch += 1
            self._sampler_iter = iter(self._sampler)
            indices = next(self._sampler_iter)
        self.current_bucket_id = getattr(self._sampler, '_current_bucket_id', None)
        return self.batch(indices)

    def save_sorted(self, sorted_dataset_dir, num_traces_per_file=None, num_files=None):
        """Rewrite the data set in sorted order (reference dataset.py:249-297)."""
        if (num_traces_per_file is None) == (num_files is None):
            raise ValueError('Expecting either num_traces_per_file or num_files')
        if num_traces_per_file is None:
            num_traces_per_file = math.ceil(len(self) / num_files)
        os.makedirs(sorted_dataset_dir, exist_ok=True)
        chunks = [self._sorted_indices[i:i + num_traces_per_file]
                  for i in range(0, len(self._sorted_indices), num_traces_per_file)]
        digits = len(str(len(chunks)))
        names = []
        for i, idx in enumerate(chunks):
            name = os.path.join(sorted_dataset_dir, '{}sorted_{:d}_{:0{}d}'.format(self.PREFIX, num_traces_per_file, i,
                                                                                   digits))
            self.columns(idx).save(name, sorted_on_disk=True)
            names.append(name)
        return names


def concat_columns(chunks):
    """One TraceColumns holding the traces of all chunks in order (address tables merged, first seen first)."""
    table, index = [], {}
    starts, addr = [np.zeros(1, np.int64)], []
    base = 0
    for c in chunks:
        if c.observe_names != chunks[0].observe_names or c.observe_dims != chunks[0].observe_dims:
            raise ValueError('chunks disagree on the observables')
        m = np.zeros(len(c.addresses), np.int32)
        for i, a in enumerate(c.addresses):
            if a[0] not in index:
                index[a[0]] = len(table)
                table.append(a)
            elif table[index[a[0]]] != a:
                raise ValueError('address {} has different distributions in different chunks'.format(a[0]))
            m[i] = index[a[0]]
        addr.append(m[np.asarray(c.step_addr)])
        ts = np.asarray(c.trace_start, np.int64)
        starts.append(ts[1:] + base)
        base += int(ts[-1])
    cat = lambda name: np.concatenate([np.asarray(getattr(c, name)) for c in chunks])  # noqa: E731
    return TraceColumns(table, chunks[0].observe_names, chunks[0].observe_dims, np.concatenate(starts),
                        np.concatenate(addr), cat('values'), cat('prior0'), cat('prior1'), cat('obs'))


def save_columns(dataset_dir, columns):
    """Write one unsorted file (reference naming scheme, dataset.py:128)."""
    os.makedirs(dataset_dir, exist_ok=True)
    name = os.path.join(dataset_dir, '{}{}_{}'.format(OfflineDataset.PREFIX, len(columns), uuid.uuid4()))
    columns.save(name)
    return name


def _chunks(seq, n):
    return [seq[i:i + n] for i in range(0, len(seq), n)]


class TraceBatchSampler:
    """Minibatches = consecutive chunks of the sorted indices, visited in shuffled order (dataset.py:312-327;
    the shuffle draws from numpy's global generator, as the reference's does)."""

    def __init__(self, offline_dataset, batch_size, shuffle_batches=True):
        self._batches = _chunks(list(offline_dataset._sorted_indices), batch_size)
        self._shuffle_batches = shuffle_batches

    def __iter__(self):
        if self._shuffle_batches:
            np.random.shuffle(self._batches)
        return iter(self._batches)

    def __len__(self):
        return len(self._batches)


class DistributedTraceBatchSampler:
    """Bucketed, rank-strided minibatch order (dataset.py:330-400):
      * drop whole minibatches (chosen with Python's `random` seeded 0, identically on all ranks) until the number of
        minibatches is a multiple of the world size; drop a ragged last minibatch;
      * consecutive minibatches form buckets (a ragged last bucket joins its predecessor); every epoch the bucket
        order is shuffled with numpy seeded by the epoch number (same on all ranks);
      * inside a bucket rank r takes minibatches r, r+world, ... truncated to floor(len/world) so that all ranks run
        the same number of iterations, optionally shuffled with the rank's own numpy stream."""

    def __init__(self, offline_dataset, batch_size, shuffle_batches=True, num_buckets=None, shuffle_buckets=True,
                 world=None, rank=None):
        if world is None or rank is None:
            from .parallel import world_info
            world, rank = world_info()
        self._world_size, self._rank = world, rank
        indices = list(offline_dataset._sorted_indices)
        drop = (len(indices) // batch_size) % world * batch_size
        state = random.getstate()
        random.seed(0)
        if drop > len(indices):
            raise ValueError('Cannot drop more items than the list length')
        for _ in range(drop):
            del indices[random.randrange(len(indices))]
        random.setstate(state)
        self._batches = _chunks(indices, batch_size)
        if len(self._batches[-1]) < batch_size:
            del self._batches[-1]
        if num_buckets is None:
            num_buckets = len(self._batches) / world
        self._num_buckets = num_buckets
        self._bucket_size = math.ceil(len(self._batches) / num_buckets)
        if self._bucket_size < world:
            raise RuntimeError('offline_dataset:{}, batch_size:{} and num_buckets:{} imply a bucket_size:{} smaller '
                               'than world_size:{}'.format(len(offline_dataset), batch_size, num_buckets,
                                                           self._bucket_size, world))
        self._buckets = _chunks(self._batches, self._bucket_size)
        if len(self._buckets[-1]) < self._bucket_size:
            if len(self._buckets) < 2:
                raise RuntimeError('offline_dataset:{} too small for given batch_size:{} and num_buckets:{}'.format(
                    len(offline_dataset), batch_size, num_buckets))
            self._buckets[-2].extend(self._buckets[-1])
            del self._buckets[-1]
        self._shuffle_batches, self._shuffle_buckets = shuffle_batches, shuffle_buckets
        self._epoch = 0
        self._current_bucket_id = 0

    def __iter__(self):
        self._epoch += 1
        bucket_ids = list(range(len(self._buckets)))
        if self._shuffle_buckets:
            state = np.random.get_state()
            np.random.seed(self._epoch)
            np.random.shuffle(bucket_ids)
            np.random.set_state(state)
        for bucket_id in bucket_ids:
            bucket = self._buckets[bucket_id]
            self._current_bucket_id = bucket_id
            per_rank = len(bucket) // self._world_size
            mine = bucket[self._rank::self._world_size][:per_rank]
            if self._shuffle_batches:
                np.random.shuffle(mine)
            for b in mine:
                yield b

    def __len__(self):
        return len(self._batches)
