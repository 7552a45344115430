"""CPU: index tensors of the minibatch encoding are exact (integer work: bit-exact bar).

The expected arrays are derived independently with per-trace Python loops that follow the reference's own
iteration order (sub-batches in dict-insertion order, dataset.py:25-36; per (t, b) loops,
inference_network_lstm.py:146-182), then mapped through the encoder's documented permutation."""
import numpy as np

from pyprob_b200.encoding import HEADER_WORDS, IMAGE_MAGIC, EncodedBatch, SubBatch
from tests import netfixture


def _subs(fx):
    ids = {a: i for i, a in enumerate(fx['address_order'])}
    return [SubBatch([ids[a] for a in sb['addresses']], sb['values'].numpy(), sb['prior0'].numpy(),
                     sb['prior1'].numpy(), sb['obs'].numpy()) for sb in fx['subs']], ids


def test_rows_cover_every_trace_step_once_and_in_time_major_order():
    fx = netfixture.load('mixed')
    subs, ids = _subs(fx)
    enc = EncodedBatch(subs)
    a = enc.arrays
    order = list(enc.sub_order)
    # stable sort by decreasing T
    Ts = [subs[i].T for i in order]
    assert Ts == sorted(Ts, reverse=True)
    assert sorted(order) == list(range(len(subs)))
    seen = np.zeros(enc.n_rows, dtype=int)
    trace0 = 0
    for pos, s in enumerate(order):
        sb = subs[s]
        for b in range(sb.B):
            i = trace0 + b  # trace index in encoded order
            assert a['trace_sub'][i] == pos
            prev_row = -1
            for t in range(sb.T):
                row = a['row_off'][t] + i
                seen[row] += 1
                assert a['values'][row] == sb.values[t, b]
                assert a['prior0'][row] == sb.prior0[t, b] and a['prior1'][row] == sb.prior1[t, b]
                st = a['row_step'][row]
                assert a['step_addr'][st] == sb.addr_ids[t]
                assert a['step_prev_addr'][st] == (sb.addr_ids[t - 1] if t > 0 else -1)
                assert a['step_row0'][st] <= row < a['step_row0'][st] + a['step_nrows'][st]
                assert a['row_prev'][row] == prev_row
                prev_row = row
            np.testing.assert_array_equal(a['obs'][i], sb.obs[b])
        trace0 += sb.B
    assert (seen == 1).all()
    # address groups partition the rows
    rows = a['head_rows']
    assert sorted(rows.tolist()) == list(range(enc.n_rows))
    for g in range(enc.n_groups):
        for r in rows[a['group_start'][g]:a['group_start'][g + 1]]:
            assert a['step_addr'][a['row_step'][r]] == a['group_addr'][g]
    assert len(set(a['group_addr'].tolist())) == enc.n_groups


def test_image_roundtrip_header_and_alignment():
    fx = netfixture.load('mixed')
    subs, _ = _subs(fx)
    enc = EncodedBatch(subs)
    img = enc.pack()
    hd = img[:HEADER_WORDS * 8].view(np.int64)
    assert hd[0] == IMAGE_MAGIC and hd[8] == img.nbytes
    assert list(hd[1:8]) == [enc.n_traces, enc.n_sub, enc.t_max, enc.n_rows, enc.n_steps, enc.n_groups,
                             enc.obs_in_total]
    offs, total = enc.offsets()
    assert total == img.nbytes and all(o % 16 == 0 for o in offs.values())
    vals = img[offs['values']:offs['values'] + 4 * enc.n_rows].view(np.float32)
    np.testing.assert_array_equal(vals, enc.arrays['values'])


def test_single_trace_and_single_step_edge_cases():
    one = SubBatch([3], np.zeros((1, 1)), np.zeros((1, 1)), np.ones((1, 1)), np.zeros((1, 2)))
    enc = EncodedBatch([one])
    assert (enc.n_traces, enc.n_rows, enc.n_steps, enc.n_groups, enc.t_max) == (1, 1, 1, 1, 1)
    assert enc.arrays['row_prev'][0] == -1 and enc.arrays['step_prev_addr'][0] == -1


def _random_subs(seed):
    rng = np.random.default_rng(seed)
    subs = []
    for _ in range(int(rng.integers(1, 6))):
        T, B = int(rng.integers(1, 9)), int(rng.integers(1, 300))
        subs.append(SubBatch(rng.integers(0, 7, T), rng.normal(size=(T, B)), rng.normal(size=(T, B)),
                             rng.uniform(0.5, 2, size=(T, B)), rng.normal(size=(B, 3))))
    return subs


def test_row_align_128_pads_every_segment_to_whole_tiles_and_keeps_the_payload():
    for seed in range(8):
        subs = _random_subs(seed)
        compact, tiled = EncodedBatch(subs, row_align=1), EncodedBatch(subs, row_align=128)
        a, c = tiled.arrays, compact.arrays
        assert tiled.row_align == 128 and tiled.n_traces == compact.n_traces and tiled.n_steps == compact.n_steps
        assert np.all(a['step_row0'] % 128 == 0) and np.all(a['row_off'] % 128 == 0) and tiled.n_rows % 128 == 0
        np.testing.assert_array_equal(a['step_addr'], c['step_addr'])
        np.testing.assert_array_equal(a['step_nrows'], c['step_nrows'])
        np.testing.assert_array_equal(a['step_t'], c['step_t'])
        valid = a['row_trace'] >= 0
        assert int(valid.sum()) == compact.n_rows == tiled.n_valid_rows
        # padding rows: no trace, no neighbours, neutral payload, but they still belong to their segment's step
        pad = ~valid
        assert np.all(a['row_prev'][pad] == -1) and np.all(a['row_next'][pad] == -1)
        assert np.all(a['values'][pad] == 0) and np.all(a['prior0'][pad] == 0) and np.all(a['prior1'][pad] == 1)
        assert np.all(a['row_step'] >= 0)
        for st in range(tiled.n_steps):
            r0, nb = a['step_row0'][st], a['step_nrows'][st]
            seg = (nb + 127) // 128 * 128
            assert np.all(a['row_step'][r0:r0 + seg] == st)
            assert np.all(a['row_trace'][r0:r0 + nb] >= 0) and np.all(a['row_trace'][r0 + nb:r0 + seg] == -1)
        # same (trace, time) -> same payload in both layouts
        key_t = {(int(tr), int(a['step_t'][a['row_step'][r]])): r for r, tr in enumerate(a['row_trace']) if tr >= 0}
        for r in range(compact.n_rows):
            k = (int(c['row_trace'][r]), int(c['step_t'][c['row_step'][r]]))
            rt = key_t[k]
            assert a['values'][rt] == c['values'][r] and a['prior0'][rt] == c['prior0'][r]
            assert a['prior1'][rt] == c['prior1'][r]
        # row_next is the inverse of row_prev on valid rows; chains walk t = 0..T-1 of one trace
        has_prev = np.nonzero(a['row_prev'] >= 0)[0]
        np.testing.assert_array_equal(a['row_next'][a['row_prev'][has_prev]], has_prev)
        np.testing.assert_array_equal(a['row_trace'][a['row_prev'][has_prev]], a['row_trace'][has_prev])
        # the head row lists only name valid rows, each exactly once
        assert sorted(a['head_rows'].tolist()) == np.nonzero(valid)[0].tolist()


def test_structure_key_ignores_payload_and_sees_shape():
    subs = _random_subs(3)
    other = [SubBatch(s.addr_ids, s.values + 1, s.prior0, s.prior1, s.obs) for s in subs]
    assert EncodedBatch(subs, 128).structure_key() == EncodedBatch(other, 128).structure_key()
    assert EncodedBatch(subs, 128).structure_key() != EncodedBatch(subs, 1).structure_key()
    fewer = [SubBatch(s.addr_ids, s.values[:, :1], s.prior0[:, :1], s.prior1[:, :1], s.obs[:1]) for s in subs]
    assert EncodedBatch(subs, 128).structure_key() != EncodedBatch(fewer, 128).structure_key()


def test_errors_empty_batch_and_short_image_buffer():
    import pytest
    with pytest.raises(ValueError):
        EncodedBatch([])
    enc = EncodedBatch(_random_subs(1), 128)
    with pytest.raises(ValueError):
        enc.pack(out=np.zeros(64, np.uint8))
    buf = np.zeros(enc.offsets()[1] + 100, np.uint8)
    img = enc.pack(out=buf)
    assert img.nbytes == enc.offsets()[1] and img[:8].view(np.int64)[0] == IMAGE_MAGIC
    assert img[28 * 8:29 * 8].view(np.int64)[0] == 128
