"""N > 1 host logic on CPU: world_size-2 gloo run in which every rank scores its contiguous
shard (with the CPU oracle standing in for the device) and rank 0 reassembles the compacted
logL vector exactly as mdns_fetch does across shards."""
import os
import socket

import numpy
import pytest

from massivedatans_b200 import sharding, synth


def test_shard_ranges_cover_everything():
    for ndata in (1, 2, 7, 100, 1000003):
        for g in (1, 2, 3, 8):
            r = sharding.shard_ranges(ndata, g)
            assert r[0][0] == 0 and sum(n for _, n in r) == ndata
            assert all(r[i][0] + r[i][1] == r[i + 1][0] for i in range(len(r) - 1))
            assert max(n for _, n in r) - min(n for _, n in r) <= 1
            assert len(r) == min(g, ndata)


def test_compaction_offsets():
    mask = numpy.array([1, 0, 1, 1, 0, 0, 1, 1, 1, 0], dtype=bool)
    r = sharding.shard_ranges(10, 3)
    assert r == [(0, 4), (4, 3), (7, 3)]
    assert sharding.compaction_offsets(mask, r) == [(0, 3), (3, 1), (4, 2)]


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, N, out_path):
    import torch
    import torch.distributed as dist
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from oracle import port as oracle_port
    x, y, _ = synth.horns(N)
    mask = synth.masks(N)['half']
    pts = synth.parameter_points(3)
    ranges = sharding.shard_ranges(N, world)
    offs = sharding.compaction_offsets(mask, ranges)
    i0, n = ranges[rank]
    ys = numpy.ascontiguousarray(y[:, i0:i0 + n])          # this rank's resident shard
    ms = numpy.ascontiguousarray(mask[i0:i0 + n])
    part = numpy.array([oracle_port.clike(x, ys, p[0], p[1], p[2], 0.01, ms) for p in pts])
    # timing plumbing of bench.py: barrier + max over ranks
    t = torch.tensor([float(rank + 1)], dtype=torch.float64)
    dist.barrier()
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    assert t.item() == float(world)
    gathered = [None] * world
    dist.all_gather_object(gathered, part)
    if rank == 0:
        n_act = int(mask.sum())
        full = numpy.empty((len(pts), n_act))
        for g, (off, cnt) in enumerate(offs):
            full[:, off:off + cnt] = gathered[g]
        numpy.save(out_path, full)
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_sharded_evaluation_matches_single(tmp_path, oracle_port):
    import torch.multiprocessing as mp
    N = 301
    out = str(tmp_path / 'full.npy')
    mp.spawn(_worker, args=(2, _free_port(), N, out), nprocs=2, join=True)
    got = numpy.load(out)
    x, y, _ = synth.horns(N)
    mask = synth.masks(N)['half']
    want = numpy.array([oracle_port.clike(x, y, p[0], p[1], p[2], 0.01, mask)
                        for p in synth.parameter_points(3)])
    assert numpy.array_equal(got, want)


def _worker_accept(rank, world, port, N, out_path):
    import torch.distributed as dist
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from oracle import port as oracle_port
    x, y, _ = synth.horns(N)
    pts = synth.parameter_points(6, seed=3)
    allm = numpy.ones(N, dtype=bool)
    full = numpy.array([-0.5 * oracle_port.clike(x, y, p[0], p[1], p[2], 0.01, allm) for p in pts])
    # thresholds: candidates 0..2 rejected everywhere, candidate 3 accepted for data set N-1 only
    # (which lives on the LAST rank), candidate 4 accepted on the first rank
    Lmins = full.max(axis=0) + 1.0
    Lmins[N - 1] = full[3, N - 1] - 1e-3
    Lmins[0] = full[4, 0] - 1e-3
    i0, n = sharding.shard_ranges(N, world)[rank]
    local = (full[:, i0:i0 + n] > Lmins[i0:i0 + n]).sum(axis=1)      # what draw_counts returns
    import torch

    def allreduce_sum(v):          # the exchange run by the caller: gloo on the host
        t = torch.as_tensor(numpy.asarray(v, dtype=numpy.int64))
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return t.numpy()

    k, total = sharding.global_first_accepted(local, allreduce_sum)
    want_total = (full > Lmins).sum(axis=1)
    assert numpy.array_equal(total, want_total)
    assert k == sharding.first_accepted_from_counts(want_total)
    if rank == 0:
        numpy.save(out_path, numpy.array([k, sharding.first_accepted_from_counts(local)]))
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_first_accept_needs_the_count_exchange(tmp_path, oracle_port):
    import torch.multiprocessing as mp
    out = str(tmp_path / 'k.npy')
    mp.spawn(_worker_accept, args=(2, _free_port(), 200, out), nprocs=2, join=True)
    k_global, k_rank0 = numpy.load(out)
    # the globally first accepted candidate is not the one rank 0 would have picked alone
    assert k_global <= k_rank0 or k_rank0 == -1
    assert k_global >= 0


def _worker_rendezvous(rank, world, port, out_dir):
    # the hand-over of the NCCL id (sharding.exchange_unique_id): plain TCP, no framework
    made = []

    def make_id():
        made.append(1)
        return bytes(range(128))

    uid = sharding.exchange_unique_id(make_id, rank, world, addr='127.0.0.1', port=port, timeout=60)
    assert (len(made) == 1) == (rank == 0)
    with open(os.path.join(out_dir, 'uid%d' % rank), 'wb') as f:
        f.write(uid)


@pytest.mark.timeout(120)
def test_unique_id_rendezvous_over_tcp(tmp_path):
    import torch.multiprocessing as mp
    world = 3
    mp.spawn(_worker_rendezvous, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        assert open(str(tmp_path / ('uid%d' % r)), 'rb').read() == bytes(range(128))


def test_dist_env_and_decision_rule(monkeypatch):
    monkeypatch.setenv('RANK', '3')
    monkeypatch.setenv('WORLD_SIZE', '8')
    monkeypatch.setenv('LOCAL_RANK', '3')
    assert sharding.dist_env() == (3, 8, 3)
    assert sharding.global_first_accepted([0, 0, 2, 1]) == (2, pytest.approx([0, 0, 2, 1]))
    assert sharding.global_first_accepted([0, 0], lambda v: v + numpy.array([0, 5]))[0] == 1
    assert sharding.first_accepted_from_counts([0, 0, 0]) == -1


def test_package_does_not_import_torch():
    # north star: the shim and its Python surface carry no PyTorch dependency
    import subprocess
    import sys
    from conftest import ROOT
    code = ('import sys; import massivedatans_b200, massivedatans_b200.sharding, '
            'massivedatans_b200.likelihood, massivedatans_b200.livepoints, '
            'massivedatans_b200.hiermetriclearn, massivedatans_b200.clustering.neighbors, '
            'massivedatans_b200.clustering.radfriendsregion; '
            'assert "torch" not in sys.modules, "torch was imported"')
    subprocess.check_call([sys.executable, '-c', code], cwd=ROOT)
    src = open(os.path.join(ROOT, 'massivedatans_b200', 'sharding.py')).read()
    assert 'import torch' not in src
