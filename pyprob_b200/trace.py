"""Particle-batched traces.

The reference keeps one ``Trace`` of ``Variable`` objects per particle (pyprob/trace.py:9-125) and runs the
user's ``forward`` once per particle.  Here ``forward`` runs once for n particles in lock-step: a ``Site`` is
one executed sample/observe statement holding length-n device tensors, and ``BatchedTrace`` carries the
per-particle log importance weights as an fp64 accumulator (Trace.end's double sum, trace.py:123-125).
"""
import numpy as np
import torch


class Site:
    """One sample/observe statement execution over all particles (the batched ``Variable``)."""

    def __init__(self, distribution, value, address_base, address, instance, control=False, name=None,
                 observed=False, tagged=False, mask=None, log_prob=None):
        self.distribution = distribution
        self.value = value            # [n] CUDA tensor (or arbitrary payload for tagged sites)
        self.address_base = address_base
        self.address = address
        self.instance = instance
        self.control = control
        self.name = name
        self.observed = observed
        self.tagged = tagged
        self.observable = ((not tagged) and (name is not None)) or observed
        self.mask = mask              # None = all particles active, else bool [n]
        self.log_prob = log_prob      # [n] prior log-prob of the value (None if not scored)

    def __repr__(self):
        return 'Site(address:{}, name:{}, control:{}, observed:{}, distribution:{})'.format(
            self.address, self.name, self.control, self.observed, self.distribution)


class BatchedTrace:
    def __init__(self, n):
        self.n = n
        self.sites = []
        self.named_variables = {}
        self._instances = {}
        self.log_w = torch.zeros(n, dtype=torch.float64, device='cuda')  # per-particle log importance weight
        self.result = None
        self.execution_time_sec = None
        self.ic_state = None   # per-lane LSTM state and previous site of the proposal network (state._sample_from_proposal)

    def add(self, site):
        self.sites.append(site)
        if site.name is not None:
            self.named_variables[site.name] = site

    def next_instance(self, address_base):
        i = self._instances.get(address_base, 0) + 1
        self._instances[address_base] = i
        return i

    @property
    def variables_controlled(self):
        return [s for s in self.sites if s.control]

    @property
    def variables_observed(self):
        return [s for s in self.sites if s.observed]

    @property
    def length_controlled(self):
        return len(self.variables_controlled)

    @staticmethod
    def value_shape(site):
        v = site.value
        return tuple(v.shape[1:]) if torch.is_tensor(v) and v.dim() > 1 else ()

    def sub_batches(self):
        """Group particles by their sequence of active controlled sites -> list of (site list, index tensor).

        Straight-line models give one group holding every particle; masked loops give one group per
        distinct activity pattern (the batched form of dataset.py:25-36)."""
        ctrl = self.variables_controlled
        if all(s.mask is None for s in ctrl):
            return [(ctrl, None)]
        n = self.n
        bits = torch.stack([torch.ones(n, dtype=torch.bool, device='cuda') if s.mask is None else s.mask
                            for s in ctrl], dim=1).cpu().numpy()
        keys, inverse = np.unique(bits, axis=0, return_inverse=True)
        groups = []
        for g in range(keys.shape[0]):
            idx = np.nonzero(inverse.reshape(-1) == g)[0]
            sites = [s for s, on in zip(ctrl, keys[g]) if on]
            if len(sites) == 0:
                raise ValueError('Trace of length zero.')
            groups.append((sites, torch.as_tensor(idx, device='cuda')))
        # reference order: first appearance of each pattern in particle order (dict insertion order)
        groups.sort(key=lambda g: int(g[1][0]))
        return groups
