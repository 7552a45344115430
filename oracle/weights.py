"""Oracle: importance-weight accumulation and normalisation (TEST INFRASTRUCTURE ONLY).

* per-trace weight = double-precision sum of fp32 per-variable terms   pyprob/trace.py:123-125
* stored as fp32                                                        pyprob/distributions/empirical.py:326
* normalised logits = w - logsumexp(w) in fp64                          pyprob/distributions/empirical.py:298-302
* ESS = 1 / sum p^2                                                     pyprob/distributions/empirical.py:759-766
"""
import numpy as np


def accumulate(terms_fp32):
    """terms_fp32: [n_terms, n] fp32 -> fp64 [n] python-float style sum, in order."""
    acc = np.zeros(terms_fp32.shape[1], dtype=np.float64)
    for t in terms_fp32:
        acc += t.astype(np.float64)
    return acc


def finalize(log_w_fp32):
    w = np.asarray(log_w_fp32, dtype=np.float32).astype(np.float64)
    m = w.max()
    s = np.exp(w - m).sum()
    lse = m + np.log(s)
    logits = w - lse
    p = np.exp(logits)
    ess = 1.0 / np.sum(p * p)
    return lse, ess, logits
