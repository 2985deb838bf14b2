"""Host-side plumbing for data-parallel training: one process per GPU, torch.distributed for rendezvous only.

The engine all-reduces weight gradients itself (NCCL over NVLink, inside b2nn_train, overlapped with backprop); this
module only (1) carries the 128-byte NCCL unique id from rank 0 to the others and (2) cuts every mini-batch of the
reference's batch table (allocVarGPU, cublasNN.cpp:401, 421-432) into per-rank row slices.
"""
from __future__ import annotations

import numpy as np


def batch_table(m: int, batch_num: int):
    """Contiguous row ranges of the reference's batches: floor(m / batch_num) rows each, the last absorbs the rest."""
    bs = m // batch_num
    return [(b * bs, m if b == batch_num - 1 else (b + 1) * bs) for b in range(batch_num)]


def rank_slice(lo: int, hi: int, rank: int, world: int):
    """Rows of [lo, hi) owned by `rank`: near-equal contiguous shares, earlier ranks take the extra rows."""
    n = hi - lo
    base, extra = divmod(n, world)
    start = lo + rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def shard_training_set(x_cm: np.ndarray, y_cm: np.ndarray, m: int, n1: int, k: int, batch_num: int, rank: int, world: int):
    """Local training set of `rank`: its slice of EVERY batch, concatenated in batch order, in the same column-major
    host layout.  Returns (x_local, y_local, m_local).  The local set is again cut into `batch_num` batches by the
    engine; the slices are chosen so that this local table reproduces exactly the per-batch slices."""
    X = x_cm.reshape(n1, m)
    Y = y_cm.reshape(k, m)
    cols = []
    sizes = []
    for lo, hi in batch_table(m, batch_num):
        s, e = rank_slice(lo, hi, rank, world)
        cols.append(np.arange(s, e))
        sizes.append(e - s)
    m_local = int(sum(sizes))
    bs_local = m_local // batch_num
    want = [bs_local] * (batch_num - 1) + [m_local - bs_local * (batch_num - 1)]
    if sizes != want:
        raise ValueError(f"batch shares {sizes} are not reproducible by a local batch table {want}; "
                         f"choose m, batch_num and world so that every batch divides evenly")
    idx = np.concatenate(cols)
    return (np.ascontiguousarray(X[:, idx]).reshape(-1), np.ascontiguousarray(Y[:, idx]).reshape(-1), m_local)


def exchange_unique_id(make_id, rank: int, world: int) -> bytes:
    """Rank 0 creates the NCCL unique id (make_id() -> 128 bytes); everyone receives it over torch.distributed."""
    import torch.distributed as dist
    box = [make_id() if rank == 0 else None]
    if world > 1:
        dist.broadcast_object_list(box, src=0)
    uid = box[0]
    if not isinstance(uid, (bytes, bytearray)) or len(uid) != 128:
        raise RuntimeError("bad NCCL unique id")
    return bytes(uid)


def enable_peer_exchange(net, rank: int, world: int) -> bool:
    """Switch the large layers' gradient exchange from NCCL to the fused peer-memory path (b2nn_comm_p2p_*): every rank
    exports its three CUDA IPC handles, the blobs are all-gathered in rank order, every rank opens its peers'."""
    import torch.distributed as dist
    if world < 2 or world > 8:
        return False
    mine = net.comm_p2p_export()
    blobs = [None] * world
    dist.all_gather_object(blobs, mine)
    net.comm_p2p_open(b"".join(blobs))
    return True
