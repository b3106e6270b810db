"""Contiguous data-set sharding (host logic shared by the shim, bench.py and the tests) and the
rendezvous of the one-process-per-GPU mode.

Data sets are independent (clike.c:68-74, cmuselike.c:48-62): GPU g holds the columns
[i0, i0+n) of the host matrix; masks are sliced the same way and the compacted logL of
shard g lands at offset sum(n_act of the shards before it) -- exactly what
mdns_dataset_create / mdns_fetch do natively (capi.cu).

One process per GPU: the only global decision of the path is `numpy.any(L > Lmins)`
(hiermetriclearn.py:193).  The shim adds the K per-candidate accept counts up over the ranks with
ncclAllReduce on the data set's stream (mdns_comm_*, include/mdns_b200.h); what Python contributes
is the hand-over of the 128-byte NCCL id from rank 0 to the others -- a few lines of TCP on
MASTER_ADDR, no framework needed.
"""
import os
import socket
import struct
import time

import numpy

UNIQUE_ID_BYTES = 128
PORT_OFFSET = 23          # rendezvous port = MASTER_PORT + PORT_OFFSET (MASTER_PORT itself belongs
                          # to whoever launched the ranks, e.g. torchrun's store)


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


def global_first_accepted(local_counts, allreduce_sum=None):
    """The decision rule of the exchange step, for callers that run the exchange themselves
    (ResidentDataset.draw_counts + fetch_candidate): add the per-candidate accept counts of all
    ranks with `allreduce_sum` (a callable numpy int64 array -> summed array; None = one rank)
    and return (k, global_counts).  With a communicator attached (ResidentDataset.init_comm)
    the shim does all of this on the device and this function is not needed."""
    total = numpy.asarray(local_counts, dtype=numpy.int64)
    if allreduce_sum is not None:
        total = numpy.asarray(allreduce_sum(total), dtype=numpy.int64)
    return first_accepted_from_counts(total), total


def dist_env():
    """(rank, world_size, local_rank) from the launcher's environment (torchrun, mpirun, srun)."""
    e = os.environ
    rank = int(e.get('RANK', e.get('OMPI_COMM_WORLD_RANK', e.get('SLURM_PROCID', '0'))))
    world = int(e.get('WORLD_SIZE', e.get('OMPI_COMM_WORLD_SIZE', e.get('SLURM_NTASKS', '1'))))
    local = int(e.get('LOCAL_RANK', e.get('OMPI_COMM_WORLD_LOCAL_RANK', e.get('SLURM_LOCALID', '0'))))
    return rank, world, local


def _recv_exact(conn, n):
    buf = b''
    while len(buf) < n:
        chunk = conn.recv(n - len(buf))
        if not chunk:
            raise ConnectionError('rendezvous peer closed the connection')
        buf += chunk
    return buf


def exchange_unique_id(make_id, rank, world, addr=None, port=None, timeout=300.0):
    """Rank 0 calls `make_id()` (-> 128 bytes) and serves it to the other ranks over TCP; every
    rank returns the same bytes.  addr / port default to MASTER_ADDR and MASTER_PORT + 23."""
    if world <= 1:
        return make_id()
    addr = addr or os.environ.get('MASTER_ADDR', '127.0.0.1')
    if port is None:
        port = int(os.environ.get('MDNS_COMM_PORT', 0)) or \
            int(os.environ.get('MASTER_PORT', '29500')) + PORT_OFFSET
    if rank == 0:
        uid = bytes(make_id())
        assert len(uid) == UNIQUE_ID_BYTES
        srv = socket.socket(socket.AF_INET, socket.SOCK_STREAM)
        srv.setsockopt(socket.SOL_SOCKET, socket.SO_REUSEADDR, 1)
        srv.bind((addr if addr in ('127.0.0.1', 'localhost') else '', port))
        srv.listen(world)
        srv.settimeout(timeout)
        served = set()
        try:
            while len(served) < world - 1:
                conn, _ = srv.accept()
                with conn:
                    conn.settimeout(timeout)
                    peer = struct.unpack('<i', _recv_exact(conn, 4))[0]
                    conn.sendall(uid)
                    served.add(peer)
        finally:
            srv.close()
        return uid
    deadline = time.time() + timeout
    while True:
        try:
            conn = socket.create_connection((addr, port), timeout=5.0)
            break
        except OSError:
            if time.time() > deadline:
                raise
            time.sleep(0.05)
    with conn:
        conn.settimeout(timeout)
        conn.sendall(struct.pack('<i', rank))
        return _recv_exact(conn, UNIQUE_ID_BYTES)


def init_comm_from_env(dataset, addr=None, port=None):
    """One process per GPU: attach a communicator over all ranks of the launcher's environment
    to `dataset` (a ResidentDataset living on this rank's GPU).  Returns (rank, world)."""
    rank, world, _ = dist_env()
    if world > 1:
        uid = exchange_unique_id(dataset.comm_unique_id, rank, world, addr=addr, port=port)
        dataset.init_comm(uid, world, rank)
    return rank, world


def bind_near_device(device_index):
    """Pin this process to the CPU cores next to its GPU (NVML's ideal affinity), so that pinned
    result buffers are allocated on the GPU's own NUMA node: with one process per GPU the D2H
    copies of all ranks then do not funnel through one socket.  Best effort; returns True if bound."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(int(device_index))
        pynvml.nvmlDeviceSetCpuAffinity(h)
        return True
    except Exception:
        return False
