"""Contiguous data-set sharding (host logic shared by the shim, bench.py and the tests).

Data sets are independent (clike.c:68-74, cmuselike.c:48-62): GPU g holds the columns
[i0, i0+n) of the host matrix; masks are sliced the same way and the compacted logL of
shard g lands at offset sum(n_act of the shards before it) -- exactly what
mdns_dataset_create / mdns_fetch do natively (capi.cu).
"""
import numpy


def shard_ranges(ndata, nshards):
    """[(i0, n)] -- remainder spread over the first shards (as in mdns_dataset_create)."""
    nshards = max(1, min(int(nshards), int(ndata)))
    base, extra = divmod(int(ndata), nshards)
    out, i0 = [], 0
    for g in range(nshards):
        n = base + (1 if g < extra else 0)
        out.append((i0, n))
        i0 += n
    return out


def compaction_offsets(data_mask, ranges):
    """Per shard: (offset of its first compacted output, number of active data sets)."""
    data_mask = numpy.asarray(data_mask, dtype=bool)
    counts = [int(numpy.count_nonzero(data_mask[i0:i0 + n])) for i0, n in ranges]
    offs = numpy.concatenate([[0], numpy.cumsum(counts)[:-1]]).astype(int)
    return list(zip(offs.tolist(), counts))


def first_accepted_from_counts(counts):
    """Index of the first candidate accepted for at least one data set, or -1
    (hiermetriclearn.py:181-196: candidates are tried in order until numpy.any(L > Lmins))."""
    nz = numpy.nonzero(numpy.asarray(counts) > 0)[0]
    return int(nz[0]) if len(nz) else -1


def global_first_accepted(local_counts, device=None, group=None):
    """One process per GPU: add the per-candidate accept counts of all ranks (the only exchange
    step of the sharded path -- K integers through torch.distributed, NCCL over NVLink when
    `device` is a CUDA device, gloo on the host otherwise) and return (k, global_counts); every
    rank then fetches candidate k from its own shard."""
    import torch
    import torch.distributed as dist
    t = torch.as_tensor(numpy.asarray(local_counts, dtype=numpy.int64))
    if device is not None:
        t = t.to(device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    total = t.cpu().numpy()
    return first_accepted_from_counts(total), total
