"""Multi-GPU plumbing: one process per GPU, voxels sharded in contiguous row blocks, no data-path collective.

The FID -> spectrum chain is embarrassingly parallel over voxels, so each rank owns
``rows[lo:hi]`` and never touches another rank's data.  The only exchange is the one the reference's semantics
force: ``autophase(mode="single")`` needs the GLOBAL first-occurrence argmax of |S| (``phasing.py:229-231``), i.e.
one ``(max, flat index)`` pair per rank, and the winning rank's ``(p0, p1, pivot)`` back -- a few dozen bytes through
``torch.distributed`` (NCCL on the GPU box, gloo in the CPU tests).
"""

from __future__ import annotations

import numpy as np


def shard_bounds(n_rows: int, world_size: int, rank: int):
    """Contiguous, balanced row block of ``rank`` (first ``n_rows % world_size`` ranks get one extra row)."""
    base, extra = divmod(n_rows, world_size)
    lo = rank * base + min(rank, extra)
    hi = lo + base + (1 if rank < extra else 0)
    return lo, hi


def pick_winner(values, flat_indices, row_offsets):
    """Global first-occurrence argmax from per-rank ``(max, local flat index)`` pairs.

    Ties go to the lowest global flat index (numpy ``argmax`` order; shards are contiguous row blocks, so this
    is the lowest rank holding the maximum).  Ranks with an empty shard report ``-inf``.
    Returns ``(winner_rank, global_flat_index)``.
    """
    values = np.asarray(values, dtype=np.float64)
    best = None
    for r in range(len(values)):
        if not np.isfinite(values[r]) and values[r] < 0:
            continue
        if best is None or values[r] > values[best]:
            best = r
    if best is None:
        raise ValueError("all shards are empty")
    return best, int(flat_indices[best]) + int(row_offsets[best])


def slot_elems(n_in: int) -> int:
    """complex64 elements of one candidate slot: the FID row (padded to an even length), then {max |S|, 0}, then the int64
    global row (``pack_winner_kernel`` / ``select_winner_kernel`` in ``csrc/xmris_abi.cu``)."""
    return ((n_in + 1) & ~1) + 2


def pack_slot(row_fid, vmax, global_row):
    """Host mirror of ``pack_winner_kernel`` (tests and documentation of the wire layout): one slot as complex64."""
    n_in = len(row_fid)
    slot = np.zeros(slot_elems(n_in), dtype=np.complex64)
    slot[:n_in] = row_fid
    slot[-2] = np.float32(vmax)
    slot[-1:].view(np.int64)[0] = global_row
    return slot


def select_winner(gathered):
    """Host mirror of ``select_winner_kernel``: ``(slot index, max |S|, global row)`` of the winner among ``[world, slot]``
    gathered slots; ties go to the lowest global row (numpy ``argmax`` order over the concatenated shards)."""
    best, brow, bslot = -2.0, np.iinfo(np.int64).max, 0
    for r in range(gathered.shape[0]):
        v = float(gathered[r, -2].real)
        row = int(gathered[r, -1:].view(np.int64)[0])
        if v > best or (v == best and row < brow):
            best, brow, bslot = v, row, r
    return bslot, best, brow


class SlotAllGather:
    """The ONE collective of the sharded ``mode="single"`` chain: all-gather of the ranks' candidate slots
    (``xmr_chain_single_front_c64`` -> this -> ``xmr_chain_single_back_c64``, see ``device.chain_single_dev``).

    Enqueued on the current CUDA stream by ``torch.distributed`` (NCCL): no host synchronisation, no second collective --
    every rank selects the winner on the device and runs the 0.3 ms search redundantly instead of waiting for a broadcast.
    """

    def __init__(self, dist, group=None):
        self.dist, self.group = dist, group
        self.world_size = dist.get_world_size(group)

    def __call__(self, recv, send):
        self.dist.all_gather_into_tensor(recv.view(-1), send, group=self.group)


def make_exchange(dist, device, n_out: int, row_offset_elems: int):
    """Build the ``exchange`` callable for :func:`xmris_b200.chain.chain_single` on an initialised process group.

    ``row_offset_elems`` = (first global row of this rank's shard) * n_out, so that flat indices are global.
    """
    import torch

    world = dist.get_world_size()
    rank = dist.get_rank()

    def exchange(local_max, local_flat, search_fn):
        mine = torch.tensor([float(local_max), float(rank), float(local_flat), float(row_offset_elems)],
                            dtype=torch.float64, device=device)
        flat = torch.empty(world * 4, dtype=torch.float64, device=device)
        try:
            dist.all_gather_into_tensor(flat, mine)            # one collective (NCCL; recent gloo)
        except (RuntimeError, NotImplementedError, AttributeError):
            gathered = [torch.empty_like(mine) for _ in range(world)]
            dist.all_gather(gathered, mine)
            flat = torch.cat(gathered)
        table = flat.cpu().numpy().reshape(world, 4)
        winner, _ = pick_winner(table[:, 0], table[:, 2], table[:, 3])
        res = torch.zeros(4, dtype=torch.float64, device=device)
        if rank == winner:
            p0, p1, pivot, fun = search_fn()
            res = torch.tensor([p0, p1, pivot, fun], dtype=torch.float64, device=device)
        dist.broadcast(res, src=winner)
        out = res.cpu().numpy()
        return float(out[0]), float(out[1]), float(out[2]), float(out[3])

    return exchange
