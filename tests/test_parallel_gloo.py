"""CPU, world_size 2 over gloo: the data-parallel host logic — flat-arena gradient all-reduce with the loss
piggy-backed, 1/world folded into the optimiser, particle-range sharding, and the exact combination of
per-rank log-sum-exp partials (checked against the single-process oracle)."""
import os

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import weights as oweights
from pyprob_b200 import parallel


def _triples(w, block):
    out = []
    for i in range(0, len(w), block):
        c = w[i:i + block].astype(np.float64)
        m = c.max()
        e = np.exp(c - m)
        out += [m, e.sum(), (e * e).sum()]
    return np.asarray(out)


def _combine(tri):
    t = tri.reshape(-1, 3)
    m = t[:, 0].max()
    s = (t[:, 1] * np.exp(t[:, 0] - m)).sum()
    s2 = (t[:, 2] * np.exp(2 * (t[:, 0] - m))).sum()
    return m + np.log(s), s * s / s2


def _worker(rank, world, port, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    g = torch.Generator().manual_seed(100 + rank)
    grad = torch.randn(1000, generator=g)
    loss = torch.tensor(1.5 + rank)
    local = grad.clone()
    mean_loss, scale = parallel.allreduce_grad_and_loss(grad, loss)
    # particle sharding + weight partials
    n = 10007
    first, count = parallel.shard_range(n, rank, world)
    wg = np.random.default_rng(7).normal(-30, 4, n).astype(np.float32)
    part = torch.from_numpy(_triples(wg[first:first + count], 2048 if rank == 0 else 1000))
    allp = parallel.gather_weight_partials(part)
    from pyprob_b200 import dataset
    q.put((rank, local.numpy(), grad.numpy(), mean_loss, scale, first, count, allp.numpy(), dataset.rank_first_index(256)))
    dist.destroy_process_group()


def test_flat_allreduce_and_weight_partials_world2():
    world, port = 2, 29731
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(world)], key=lambda r: r[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    total = res[0][1] + res[1][1]
    for r in res:
        np.testing.assert_allclose(r[2], total, rtol=1e-6)       # summed gradient on every rank
        assert abs(r[3] - 2.0) < 1e-6 and r[4] == 0.5             # mean loss, grad_scale = 1/world
    assert res[0][5] == 0 and res[0][5] + res[0][6] == res[1][5] and res[1][5] + res[1][6] == 10007
    np.testing.assert_array_equal(res[0][7], res[1][7])           # identical gathered list on both ranks
    assert res[0][8] == 0 and res[1][8] == 256                    # online minibatches: disjoint Philox index ranges per rank
    wg = np.random.default_rng(7).normal(-30, 4, 10007).astype(np.float32)
    lse, ess, _ = oweights.finalize(wg)
    got_lse, got_ess = _combine(res[0][7])
    np.testing.assert_allclose(got_lse, lse, rtol=1e-12)
    np.testing.assert_allclose(got_ess, ess, rtol=1e-9)


def test_shard_range_partitions():
    for n in (0, 1, 7, 64, 65537):
        for world in (1, 2, 3, 8):
            spans = [parallel.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and sum(c for _, c in spans) == n
            for (f0, c0), (f1, _) in zip(spans, spans[1:]):
                assert f0 + c0 == f1


def _offline_worker(rank, world, port, path, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from pyprob_b200 import offline
    ds = offline.OfflineDataset(path)
    ds.num_buckets = 3
    np.random.seed(50 + rank)
    per_epoch = None
    seen = []
    for epoch in range(2):
        batches = []
        # one epoch of this rank's share: the sampler yields the same number of minibatches on every rank
        sampler = ds._make_sampler(16)
        if per_epoch is None:
            per_epoch = sum(1 for _ in sampler)
            ds._sampler_iter = None
        for _ in range(per_epoch):
            b = ds.next_batch(16)
            batches.append((b.size, sorted(float(x) for sb in b.subs for x in sb['values'][0])))
        seen.append(batches)
    q.put((rank, per_epoch, seen))
    dist.destroy_process_group()


def test_offline_dataset_shards_minibatches_over_ranks_world2(tmp_path):
    """The training loop's data side under torch.distributed: OfflineDataset.next_batch picks the bucketed,
    rank-strided sampler (reference dataset.py:330-400); ranks see disjoint minibatches and run in lock step."""
    from pyprob_b200 import offline, synthetic
    rng = np.random.default_rng(0)
    table = [('a_n', 'Normal', 0), ('a_u', 'Uniform', 0)]
    subs = [synthetic.random_sub_batch(rng, [table[i] for i in seq], B, 2) for seq, B in (([0], 300), ([0, 1], 220))]
    offline.save_columns(str(tmp_path), offline.TraceColumns.from_sub_batches(subs, ['o'], [2]))
    world, port = 2, 29741
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_offline_worker, args=(r, world, port, str(tmp_path), q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(world)], key=lambda r: r[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0][1] == res[1][1] > 0                      # same number of iterations per epoch on both ranks
    for epoch in range(2):
        a = [tuple(v) for _, v in res[0][2][epoch]]
        b = [tuple(v) for _, v in res[1][2][epoch]]
        assert all(size == 16 for size, _ in res[0][2][epoch] + res[1][2][epoch])
        assert not set(a) & set(b)                         # disjoint minibatches
        assert len(set(a)) == len(a) and len(set(b)) == len(b)


def test_owned_optimizer_slices_partition_the_arena():
    """Checkpoint gather under the fused data-parallel step: the slice whose Adam moments live on rank r (network._owned_slice)
    must be the slice k_dp_adam gives that rank (csrc/dp.cu: ceil(n / 4 world) float4 blocks each) and together they must
    cover the arena exactly once."""
    from pyprob_b200.network import InferenceNetworkLSTM
    net = InferenceNetworkLSTM(model=None, observe_embeddings={'o': {}})
    for n in (1, 7, 100003, 1643588):
        for world in (1, 2, 3, 8):
            per = ((n + world * 4 - 1) // (world * 4)) * 4
            covered = 0
            for r in range(world):
                lo, hi = net._owned_slice(n, world, r)
                assert lo == min(n, r * per) and hi == min(n, lo + per) and (lo % 4 == 0 or lo == n)
                assert lo == covered or lo == n
                covered = max(covered, hi)
            assert covered == n
