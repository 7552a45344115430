"""Test helper: replay a lock-step IC execution particle group by particle group through the oracle.

Given the BatchedTrace of one `Model._run_batched(..., IMPORTANCE_SAMPLING_WITH_INFERENCE_NETWORK)` execution, recompute for
every particle sum_sites [log p(v) - log q(v)] with q from oracle.network.infer_sequence (pinned to the unmodified reference's
_infer_step) fed the SAME values, following each particle's own sequence of executed sites (pyprob/state.py:203-219: the
previous variable of a trace is the last controlled site that trace executed)."""
import numpy as np
import torch

from oracle import network as onet
from oracle import scoring


def _cpu(x):
    return x.cpu() if torch.is_tensor(x) else torch.tensor(float(x))


def _sel(x, idx):
    x = _cpu(x)
    return x[idx] if x.dim() > 0 and x.numel() > 1 else x


def site_weight_terms(trace, net, observe_row):
    """-> (want [n] float64 of the summed site terms, covered [n] bool: particles all of whose sites the network knows)."""
    n = trace.n
    ctrl = trace.variables_controlled
    vals = torch.stack([s.value.float() for s in ctrl]).cpu()
    active = torch.stack([torch.ones(n, dtype=torch.bool) if s.mask is None else s.mask.cpu() for s in ctrl]).numpy()
    P = {k: v.cpu() for k, v in net.reference_state_dict().items()}
    K = net._proposal_mixture_components
    want = np.zeros(n, np.float64)
    covered = np.zeros(n, bool)
    patterns, inverse = np.unique(active.T, axis=0, return_inverse=True)
    for g, pat in enumerate(patterns):
        idx = torch.as_tensor(np.nonzero(inverse.reshape(-1) == g)[0])
        seq = [t for t in range(len(ctrl)) if pat[t]]
        if any(ctrl[t].address not in net._addresses for t in seq):
            continue
        steps = []
        for j, t in enumerate(seq):
            d = ctrl[t].distribution
            fam = d.name
            p0 = p1 = None
            if fam == 'Normal':
                p0, p1 = _sel(d.loc, idx), _sel(d.scale, idx)
            elif fam == 'Uniform':
                p0, p1 = _sel(d.low, idx), _sel(d.high, idx)
            steps.append({'address': ctrl[t].address, 'family': fam,
                          'num_categories': getattr(d, 'num_categories', 0) if fam == 'Categorical' else 0,
                          'prior0': p0, 'prior1': p1, 'prev_value': vals[seq[j - 1], idx] if j > 0 else None})
        props = onet.infer_sequence(P, observe_row, net._observe_names, net._observe_in_dims, K, steps, n=idx.numel())
        lw = torch.zeros(idx.numel(), dtype=torch.float64)
        for j, t in enumerate(seq):
            d = ctrl[t].distribution
            v = vals[t, idx]
            if d.name == 'Categorical':
                log_q = scoring.categorical_log_prob(v, props[j][0])
                log_p = scoring.categorical_log_prob(v, _cpu(d.probs))
            else:
                means, sds, probs = props[j]
                if d.name == 'Normal':
                    log_q = scoring.mixture_normal_log_prob(v, means, sds, probs)
                elif d.name == 'Uniform':
                    log_q = scoring.mixture_truncated_normal_log_prob(v, means, sds, probs, _sel(d.low, idx), _sel(d.high, idx))
                else:
                    log_q = scoring.mixture_truncated_normal_log_prob(v, means, sds, probs, torch.tensor(0.0), torch.tensor(40.0))
                if d.name == 'Normal':
                    log_p = scoring.normal_log_prob(v, _sel(d.loc, idx), _sel(d.scale, idx))
                elif d.name == 'Uniform':
                    log_p = scoring.uniform_log_prob(v, _sel(d.low, idx), _sel(d.high, idx))
                else:
                    log_p = scoring.poisson_log_prob(v, _sel(d.rate, idx))
            lw += log_p.double() - log_q.double()
        want[idx.numpy()] = lw.numpy()
        covered[idx.numpy()] = True
    return want, covered
