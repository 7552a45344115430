"""CPU: the oracle restatement of the scoring path vs (a) fixtures produced by the unmodified reference
(tests/golden/make_golden.py) and (b) the known-answer constants of the reference's own
tests/test_distributions.py (:1183-1214 Normal, :1315-1350 TruncatedNormal, :1467-1525 Categorical,
:1556-1593 Uniform, :1635-1669 Poisson, :2091-2161 Mixture).  The reference pins those at atol=0.1; the
oracle is held to 1e-5 against the live reference outputs."""
import numpy as np
import torch

from oracle import philox, scoring, weights

TOL = dict(rtol=1e-5, atol=1e-5)


def _close(a, b, **kw):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    kw = kw or TOL
    fin = np.isfinite(b)
    assert np.array_equal(np.isfinite(a), fin)
    assert np.array_equal(a[~fin], b[~fin]) or np.all(np.isnan(a[~fin]) == np.isnan(b[~fin]))
    np.testing.assert_allclose(a[fin], b[fin], **kw)


def test_normal_vs_reference(golden_scoring):
    g = golden_scoring
    _close(scoring.normal_log_prob(g['normal_value'], g['normal_mean'], g['normal_stddev']), g['normal_lp'])


def test_uniform_vs_reference(golden_scoring):
    g = golden_scoring
    _close(scoring.uniform_log_prob(g['uniform_value'], g['uniform_low'], g['uniform_high']), g['uniform_lp'])


def test_poisson_vs_reference(golden_scoring):
    g = golden_scoring
    _close(scoring.poisson_log_prob(g['poisson_value'], g['poisson_rate']), g['poisson_lp'])


def test_categorical_vs_reference(golden_scoring):
    g = golden_scoring
    _close(scoring.categorical_log_prob(g['categorical_value'], g['categorical_probs']), g['categorical_lp'])


def test_mixture_normal_vs_reference(golden_scoring):
    g = golden_scoring
    _close(scoring.mixture_normal_log_prob(g['mixn_value'], g['mixn_means'], g['mixn_stddevs'], g['mixn_probs']),
           g['mixn_lp'])


def test_mixture_truncated_normal_vs_reference(golden_scoring):
    g = golden_scoring
    lp = scoring.mixture_truncated_normal_log_prob(g['mixt_value'], g['mixt_means'], g['mixt_stddevs'],
                                                   g['mixt_probs'], g['mixt_low'], g['mixt_high'])
    assert np.isinf(g['mixt_lp']).sum() > 0  # the fixture exercises the outside-domain branch
    _close(lp, g['mixt_lp'], rtol=1e-4, atol=1e-5)


def test_weights_vs_reference(golden_scoring):
    g = golden_scoring
    lse, ess, logits = weights.finalize(g['weights_log_w'])
    np.testing.assert_allclose(logits, g['weights_logits'], rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(ess, g['weights_ess'], rtol=1e-10)
    lse, ess, logits = weights.finalize(g['gum_log_w'])
    np.testing.assert_allclose(logits, g['gum_logits'], rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(ess, g['gum_ess'], rtol=1e-10)


def test_gum_is_weights_vs_reference(golden_scoring):
    """Config 1: IS weight = double sum of the two observe log-likelihoods (state.py:147-149, trace.py:123-125)."""
    g = golden_scoring
    mu = g['gum_mu']
    t0 = scoring.normal_log_prob(np.float32(8.0), mu, np.float32(np.sqrt(2.0))).numpy()
    t1 = scoring.normal_log_prob(np.float32(9.0), mu, np.float32(np.sqrt(2.0))).numpy()
    w = weights.accumulate(np.stack([t0, t1])).astype(np.float32)
    np.testing.assert_allclose(w, g['gum_log_w'], rtol=2e-6)


def test_reference_known_answers():
    assert abs(float(scoring.normal_log_prob(0., 0., 1.)) - (-0.918939)) < 1e-5
    assert abs(float(scoring.truncated_normal_log_prob(2., 2., 3., -4., 4.)) - (-1.69563)) < 1e-4
    lp = scoring.categorical_log_prob([0., 1.], [[0.1, 0.2, 0.7], [0.2, 0.5, 0.3]])
    np.testing.assert_allclose(lp.numpy(), [-2.30259, -0.693147], atol=1e-5)
    assert abs(float(scoring.uniform_log_prob(0.5, 0., 1.))) < 1e-7
    assert abs(float(scoring.poisson_log_prob(4., 4.)) - (-1.63288)) < 1e-4
    lp = scoring.mixture_normal_log_prob([0.7], [[0., 2., 3.]], [[0.1, 0.1, 0.1]], [0.7, 0.2, 0.1])
    assert abs(float(lp) - (-23.473)) < 1e-2


def test_philox_known_answer():
    # Random123 known-answer vectors for philox4x32-10
    out = philox.philox4x32_10(0, np.array([0], dtype=np.uint64), 0)[0]
    assert [hex(int(x)) for x in out] == ['0x6627e8d5', '0xe169c58d', '0xbc57ac4c', '0x9b00dbd8']
    # counter = (0xffffffff,)*4, key = (0xffffffff,)*2
    out = philox.philox4x32_10(0xffffffffffffffff, np.array([0xffffffffffffffff], dtype=np.uint64),
                               0xffffffffffffffff)[0]
    assert [hex(int(x)) for x in out] == ['0x408f276d', '0x41c83b0e', '0xa20bc7c6', '0x6d5451fd']
