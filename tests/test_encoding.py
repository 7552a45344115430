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
