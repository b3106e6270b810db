"""One process per GPU with the shim's own NCCL communicator (mdns_comm_*): the accept counts of
the speculative batch are summed over the ranks on the device, so every rank takes the same --
global -- decision (hiermetriclearn.py:181-196 on sharded data).  Needs 2 GPUs; the ranks are
plain subprocesses, the 128-byte NCCL id travels over TCP (sharding.exchange_unique_id)."""
import json
import multiprocessing
import os
import socket

import numpy
import pytest

from massivedatans_b200 import _lib, sharding, synth

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _rank_main(rank, world, port, N, out_dir):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank),
                      MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    from massivedatans_b200.likelihood import ResidentDataset
    x, y, _ = synth.horns(N, legacy=False, seed=21)
    K = 12
    pts = synth.parameter_points(K, seed=4)
    i0, n = sharding.shard_ranges(N, world)[rank]
    ds = ResidentDataset(x, numpy.ascontiguousarray(y[:, i0:i0 + n]), devices=[rank])
    assert sharding.init_comm_from_env(ds) == (rank, world)
    assert ds.comm_allreduce([rank + 1.0, 2.0])[0] == world * (world + 1) / 2
    assert ds.comm_allreduce([float(rank)], op='max')[0] == world - 1
    # the full matrix of THIS rank's shard, then thresholds that make the decision a global one:
    # everything is out of reach except ONE data set per rank, where only that row's best candidate
    # gets through -- on the last rank a row whose best candidate comes early in the batch, on rank 0
    # one whose best candidate comes late
    L = numpy.array(ds.loglike_batch(pts, None, synth.NOISE_LEVEL), copy=True)
    Lmins = L.max(axis=0) + 1.0 + numpy.abs(L.max(axis=0)) * 1e-6
    srt = numpy.sort(L, axis=0)
    clear = srt[-1] - srt[-2] > 1e-6 * numpy.abs(srt[-1])          # no near-tie for the best
    best = numpy.where(clear, numpy.argmax(L, axis=0), -1)
    if rank == world - 1:
        r = int(numpy.nonzero(best == best[best >= 0].min())[0][0])
        Lmins[r] = 0.5 * (srt[-1][r] + srt[-2][r])
    if rank == 0:
        r = int(numpy.nonzero(best == best.max())[0][0])
        Lmins[r] = 0.5 * (srt[-1][r] + srt[-2][r])
    local_counts = (L > Lmins).sum(axis=1)
    res = {'rank': rank, 'local_counts': local_counts.tolist()}
    ds.begin_draw(None, Lmins)
    k, Lk, counts = ds.draw_batch(pts, synth.NOISE_LEVEL)
    res.update(k=int(k), counts=counts.tolist(),
               vector_ok=bool(Lk is not None and numpy.allclose(Lk, L[k], rtol=1e-12, atol=0)))
    ks, js, Ljs, cs = ds.draw_batch_sparse(pts, synth.NOISE_LEVEL)
    want_j = numpy.nonzero(L[ks] > Lmins)[0] if ks >= 0 else numpy.zeros(0, dtype=int)
    res.update(ks=int(ks), sparse_ok=bool(numpy.array_equal(js, want_j) and
                                          numpy.allclose(Ljs, L[ks][want_j], rtol=1e-12, atol=0)),
               sparse_counts=cs.tolist(), counts_only=ds.draw_counts(pts, synth.NOISE_LEVEL).tolist())
    # a mask that leaves one rank without a single active data set: it still takes part
    mask = numpy.zeros(n, dtype=bool)
    if rank == 0:
        mask[:7] = True
    n_act = ds.begin_draw(mask, Lmins[mask])
    k2, L2, c2 = ds.draw_batch(pts, synth.NOISE_LEVEL)
    res.update(k_masked=int(k2), n_act_masked=int(n_act), counts_masked=c2.tolist())
    # device-consumer exchange: every rank ends up with everybody's vector of candidate 1
    ds.begin_draw(None, Lmins)
    ds.draw_counts(pts, synth.NOISE_LEVEL)
    Lall, nper = ds.allgather_candidate(1, world)
    res.update(nper=nper.tolist(), gathered_sum=float(Lall.sum()), own_sum=float(L[1].sum()),
               gathered_own_ok=bool(numpy.allclose(Lall[i0:i0 + n], L[1], rtol=1e-12, atol=0)))
    with open(os.path.join(out_dir, 'rank%d.json' % rank), 'w') as f:
        json.dump(res, f)
    ds.close()


@pytest.mark.skipif(_lib.load().mdns_device_count() < 2, reason='needs 2 GPUs')
@pytest.mark.parametrize('N', [4001, 300000])
def test_two_ranks_take_the_global_first_accept_decision(tmp_path, N):
    world = 2
    ctx = multiprocessing.get_context('spawn')
    port = _free_port()
    procs = [ctx.Process(target=_rank_main, args=(r, world, port, N, str(tmp_path))) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(240)
    hung = [p for p in procs if p.is_alive()]
    for p in hung:
        p.kill()
    assert not hung, 'a rank hung'
    assert all(p.exitcode == 0 for p in procs)
    res = [json.load(open(str(tmp_path / ('rank%d.json' % r)))) for r in range(world)]
    total = numpy.sum([r['local_counts'] for r in res], axis=0)
    want_k = int(numpy.nonzero(total)[0][0])
    alone = [int(numpy.nonzero(r['local_counts'])[0][0]) if any(r['local_counts']) else -1 for r in res]
    assert total.sum() == 2 and any(k != want_k for k in alone)     # the exchange matters
    want_masked = int(numpy.nonzero(res[0]['local_counts'])[0][0]) if any(res[0]['local_counts']) else -1
    for r in res:
        # every rank: the GLOBAL counts and the globally first accepted candidate, whoever saw it
        assert r['counts'] == total.tolist() and r['k'] == want_k and r['vector_ok']
        assert r['ks'] == want_k and r['sparse_ok'] and r['sparse_counts'] == total.tolist()
        assert r['counts_only'] == total.tolist()
        assert r['n_act_masked'] == (7 if r['rank'] == 0 else 0)
        assert sum(r['nper']) == N and r['gathered_own_ok']
    # (the masked draw keeps rank 0's first seven data sets only: same answer on both ranks)
    assert res[0]['k_masked'] == res[1]['k_masked'] and res[0]['counts_masked'] == res[1]['counts_masked']
    assert abs(res[0]['gathered_sum'] - res[1]['gathered_sum']) == 0
    assert abs(res[0]['gathered_sum'] - (res[0]['own_sum'] + res[1]['own_sum'])) < 1e-6 * abs(res[0]['gathered_sum'])
