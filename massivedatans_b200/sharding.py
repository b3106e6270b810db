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
