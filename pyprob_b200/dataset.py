"""Online training data: prior traces generated on the GPU in lock-step, grouped into sub-batches.

``TraceBatch`` is the batched counterpart of the reference's ``Batch`` (pyprob/nn/dataset.py:21-47) and
``OnlineDataset`` of its ``OnlineDataset`` (:50-62): every minibatch is one execution of the user's
``forward`` over ``batch_size`` particles in TraceMode.PRIOR_FOR_INFERENCE_NETWORK (observes are sampled).
"""
import numpy as np
import torch

from .distributions import Categorical, Normal, Uniform
from .encoding import EncodedBatch, SubBatch
from .util import PriorInflation, TraceMode


class TraceBatch:
    def __init__(self, trace):
        self.trace = trace
        self.size = trace.n
        self.groups = trace.sub_batches()  # [(controlled sites, particle index tensor or None)]
        sizes = [trace.n if idx is None else int(idx.numel()) for _, idx in self.groups]
        self.mean_length_controlled = sum(len(s) * b for (s, _), b in zip(self.groups, sizes)) / self.size
        self.num_sub_batches = len(self.groups)
        self._encoded = None

    def __len__(self):
        return self.size

    @property
    def sub_batches(self):
        return self.groups

    def address_signature(self):
        seen, out = set(), []
        for sites, _ in self.groups:
            for s in sites:
                if s.address not in seen:
                    seen.add(s.address)
                    d = s.distribution
                    out.append((s.address, d.name, d.num_categories if isinstance(d, Categorical) else 0))
        return out

    def encode(self, net):
        """-> EncodedBatch (index tensors + packed image), or None if an address is unknown to the network."""
        if self._encoded is not None:
            return self._encoded
        n = self.trace.n
        cols, plan = [], []

        def push(t):
            cols.append(t.reshape(n).float())
            return len(cols) - 1

        def param(p):
            return push(p if torch.is_tensor(p) else torch.full((n,), float(p), device='cuda'))
        obs_cols = []
        for name in net._observe_names:
            v = self.trace.named_variables[name].value
            obs_cols.append(push(v))
        for sites, idx in self.groups:
            ids, vc, p0c, p1c = [], [], [], []
            for s in sites:
                if s.address not in net._addresses:
                    print('Address unknown by inference network: {}'.format(s.address))
                    return None
                ids.append(net._addresses[s.address]['id'])
                vc.append(push(s.value))
                d = s.distribution
                if isinstance(d, Normal):
                    p0c.append(param(d.loc)); p1c.append(param(d.scale))
                elif isinstance(d, Uniform):
                    p0c.append(param(d.low)); p1c.append(param(d.high))
                else:
                    p0c.append(-1); p1c.append(-1)
            plan.append((ids, vc, p0c, p1c, idx))
        host = torch.stack(cols, dim=0).cpu().numpy()  # one device->host copy for the whole batch
        zeros = np.zeros(n, dtype=np.float32)
        subs = []
        for ids, vc, p0c, p1c, idx in plan:
            sel = slice(None) if idx is None else idx.cpu().numpy()

            def rows(cs):
                return np.stack([(zeros if c < 0 else host[c])[sel] for c in cs], axis=0)
            obs = np.stack([host[c][sel] for c in obs_cols], axis=1)
            subs.append(SubBatch(ids, rows(vc), rows(p0c), rows(p1c), obs))
        self._encoded = EncodedBatch(subs, row_align=net.row_align)
        return self._encoded


class OnlineDataset:
    def __init__(self, model, length=None, prior_inflation=PriorInflation.DISABLED):
        self._model = model
        self._length = int(1e6) if length is None else length
        self._prior_inflation = prior_inflation
        self._example = None

    def __len__(self):
        return self._length

    def next_batch(self, batch_size):
        trace = self._model._run_batched(batch_size, trace_mode=TraceMode.PRIOR_FOR_INFERENCE_NETWORK,
                                         prior_inflation=self._prior_inflation)
        return TraceBatch(trace)

    def example_trace(self):
        if self._example is None:
            self._example = self._model._run_batched(1, trace_mode=TraceMode.PRIOR_FOR_INFERENCE_NETWORK,
                                                     prior_inflation=self._prior_inflation)
        return self._example
