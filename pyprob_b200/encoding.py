"""Trace-minibatch encoding: the index tensors the CUDA proposal network consumes.

This is the host-side restatement of what the reference computes with Python loops:
  * sub-batching by concatenated address string       pyprob/nn/dataset.py:21-37 (Batch)
  * per-(t, b) LSTM-input assembly                     pyprob/nn/inference_network_lstm.py:146-182
turned into integer index arrays ("address/index tensors bit-exact" in north_star) plus flat value arrays.

Layout (see include/pyprob_b200.h, ppb_batch):
  traces are ordered by sub-batch; sub-batches by decreasing length T (stable w.r.t. the reference's
  dict-insertion order); rows are time-major, row(t, i) = row_off[t] + i for the traces with T > t.
"""
from collections import OrderedDict

import numpy as np

IMAGE_MAGIC = 0x5050423230304231
HEADER_WORDS = 32

_ARRAYS = [  # (name, dtype) in image order — must match ppb_batch_from_image
    ('row_off', np.int32), ('group_addr', np.int32), ('group_start', np.int32), ('trace_sub', np.int32),
    ('step_addr', np.int32), ('step_prev_addr', np.int32), ('step_row0', np.int32), ('step_nrows', np.int32),
    ('row_step', np.int32), ('row_prev', np.int32), ('values', np.float32), ('prior0', np.float32),
    ('prior1', np.float32), ('obs', np.float32), ('head_rows', np.int32),
    ('row_trace', np.int32), ('row_next', np.int32), ('step_t', np.int32), ('step_prev_row0', np.int32),
]


class SubBatch:
    """Traces that share one address sequence (one entry of reference Batch.sub_batches).

    addr_ids : int sequence [T]      address ids (insertion order of the network's address table)
    values   : float32 [T, B]        sampled value per step and trace (category index for Categorical)
    prior0/1 : float32 [T, B]        prior mean/stddev (Normal) or low/high (Uniform); zeros otherwise
    obs      : float32 [B, obs_dim]  flattened observed values in observe_embeddings order
    """

    def __init__(self, addr_ids, values, prior0, prior1, obs):
        self.addr_ids = np.asarray(addr_ids, dtype=np.int32)
        self.values = np.ascontiguousarray(values, dtype=np.float32)
        self.prior0 = np.ascontiguousarray(prior0, dtype=np.float32)
        self.prior1 = np.ascontiguousarray(prior1, dtype=np.float32)
        self.obs = np.ascontiguousarray(obs, dtype=np.float32)
        T, B = self.values.shape
        assert self.addr_ids.shape == (T,) and self.prior0.shape == (T, B) and self.prior1.shape == (T, B)
        assert self.obs.shape[0] == B and T > 0 and B > 0
        self.T, self.B = T, B


class EncodedBatch:
    """Index/value arrays of one minibatch + the packed image that crosses the C-ABI."""

    def __init__(self, sub_batches, row_align=1):
        """row_align = 1: compact rows (SIMT path).  row_align = 128: every (t, sub-batch) segment is padded to
        whole 128-row tiles so that all tensor-core GEMM tiles are aligned (padding rows carry row_trace = -1)."""
        if len(sub_batches) == 0:
            raise ValueError('empty batch')
        self.row_align = int(row_align)
        order = sorted(range(len(sub_batches)), key=lambda i: -sub_batches[i].T)  # stable
        subs = [sub_batches[i] for i in order]
        self.sub_order = np.asarray(order, dtype=np.int32)  # position -> index in the caller's list
        S = len(subs)
        Ts = np.asarray([s.T for s in subs], dtype=np.int64)
        Bs = np.asarray([s.B for s in subs], dtype=np.int64)
        B = int(Bs.sum())
        t_max = int(Ts[0])
        trace_off = np.concatenate([[0], np.cumsum(Bs)])
        ra = self.row_align
        Bpad = (Bs + ra - 1) // ra * ra                      # segment sizes incl. padding rows
        seg_off = np.concatenate([[0], np.cumsum(Bpad)])      # offset of sub-batch s inside a full step block
        n_active = np.asarray([int(Bpad[Ts > t].sum()) for t in range(t_max)], dtype=np.int64)
        row_off = np.concatenate([[0], np.cumsum(n_active)]).astype(np.int32)
        R = int(row_off[-1])
        obs_dim = subs[0].obs.shape[1]

        self.n_traces, self.n_sub, self.t_max, self.n_rows, self.obs_in_total = B, S, t_max, R, obs_dim
        a = {}
        a['row_off'] = row_off
        a['trace_sub'] = np.repeat(np.arange(S, dtype=np.int32), Bs)
        step_addr, step_prev, step_row0, step_nrows = [], [], [], []
        step_t, step_prev_row0 = [], []
        row_step = np.full(R, -1, dtype=np.int32)
        row_prev = np.full(R, -1, dtype=np.int32)
        row_next = np.full(R, -1, dtype=np.int32)
        row_trace = np.full(R, -1, dtype=np.int32)
        values = np.zeros(R, dtype=np.float32)
        prior0 = np.zeros(R, dtype=np.float32)
        prior1 = np.ones(R, dtype=np.float32)
        groups = OrderedDict()
        for t in range(t_max):
            for s in range(S):
                if Ts[s] <= t:
                    break  # sorted by decreasing T
                st = len(step_addr)
                r0 = int(row_off[t] + seg_off[s])
                nb = int(Bs[s])
                npad = int(Bpad[s])
                aid = int(subs[s].addr_ids[t])
                step_addr.append(aid)
                step_prev.append(int(subs[s].addr_ids[t - 1]) if t > 0 else -1)
                step_row0.append(r0)
                step_nrows.append(nb)
                step_t.append(t)
                rp = int(row_off[t - 1] + seg_off[s]) if t > 0 else -1
                step_prev_row0.append(rp)
                row_step[r0:r0 + npad] = st
                row_trace[r0:r0 + nb] = np.arange(trace_off[s], trace_off[s] + nb)
                if t > 0:
                    row_prev[r0:r0 + nb] = np.arange(rp, rp + nb)
                    row_next[rp:rp + nb] = np.arange(r0, r0 + nb)
                values[r0:r0 + nb] = subs[s].values[t]
                prior0[r0:r0 + nb] = subs[s].prior0[t]
                prior1[r0:r0 + nb] = subs[s].prior1[t]
                groups.setdefault(aid, []).append((r0, nb, npad))
        a['step_addr'] = np.asarray(step_addr, dtype=np.int32)
        a['step_prev_addr'] = np.asarray(step_prev, dtype=np.int32)
        a['step_row0'] = np.asarray(step_row0, dtype=np.int32)
        a['step_nrows'] = np.asarray(step_nrows, dtype=np.int32)
        a['row_step'], a['row_prev'] = row_step, row_prev
        a['row_next'], a['row_trace'] = row_next, row_trace
        a['step_t'] = np.asarray(step_t, dtype=np.int32)
        a['step_prev_row0'] = np.asarray(step_prev_row0, dtype=np.int32)
        a['values'], a['prior0'], a['prior1'] = values, prior0, prior1
        a['obs'] = np.concatenate([s.obs for s in subs], axis=0).astype(np.float32).reshape(B, obs_dim)
        g_addr, g_start, head_rows = [], [0], []
        for aid, segs in groups.items():
            g_addr.append(aid)
            for (r0, nb, npad) in segs:
                head_rows.append(np.arange(r0, r0 + nb, dtype=np.int32))
            g_start.append(g_start[-1] + sum(nb for _, nb, _ in segs))
        a['group_addr'] = np.asarray(g_addr, dtype=np.int32)
        a['group_start'] = np.asarray(g_start, dtype=np.int32)
        a['head_rows'] = np.concatenate(head_rows)
        self.n_valid_rows = int(a['head_rows'].size)
        self.n_steps = len(step_addr)
        self.n_groups = len(g_addr)
        self.arrays = a
        self.trace_off = trace_off
        self.mean_length_controlled = float((Ts * Bs).sum()) / B

    # structure key: everything except the float payload (lets callers reuse uploaded index arrays)
    def structure_key(self):
        a = self.arrays
        return (self.row_align, self.n_traces, self.n_rows, self.obs_in_total, a['step_addr'].tobytes(),
                a['step_nrows'].tobytes())

    def offsets(self):
        off = HEADER_WORDS * 8
        out = {}
        for name, _ in _ARRAYS:
            out[name] = off
            off += (self.arrays[name].nbytes + 15) // 16 * 16
            if self.arrays[name].nbytes == 0:
                off += 16
        return out, off

    def pack(self, out=None):
        """-> uint8 numpy image (header + arrays).  `out` may be a preallocated (pinned) uint8 buffer."""
        offs, total = self.offsets()
        if out is None:
            out = np.zeros(total, dtype=np.uint8)
        elif out.nbytes < total:
            raise ValueError('image buffer too small')
        hd = out[:HEADER_WORDS * 8].view(np.int64)
        hd[:] = 0
        hd[0] = IMAGE_MAGIC
        hd[1:8] = [self.n_traces, self.n_sub, self.t_max, self.n_rows, self.n_steps, self.n_groups,
                   self.obs_in_total]
        hd[8] = total
        hd[28] = self.row_align
        for k, (name, dt) in enumerate(_ARRAYS):
            arr = np.ascontiguousarray(self.arrays[name], dtype=dt)
            hd[9 + k] = offs[name]
            out[offs[name]:offs[name] + arr.nbytes] = arr.view(np.uint8).reshape(-1)
        return out[:total]
