"""Multi-GPU sharding of the rx_tools hot path.  One process per GPU.

* rx_power: tuner hops are independent -> contiguous hop ranges per rank; ONE all-gather of the
  (padded) int64 spectrum rows per report collates them in hop order for csv_dbm
  (src/rtl_power.c:1047-1050).  No other collective on the data path.
* rx_fm: channels are independent -> contiguous channel ranges per rank, no collective.  A single
  stream does not shard (serial carry): replicas only.
"""
from __future__ import annotations

from typing import Tuple


def rows_per_rank(n_units: int, world: int) -> int:
    return -(-n_units // world)


def unit_range(rank: int, world: int, n_units: int) -> Tuple[int, int]:
    """Contiguous [begin, end) of hops/channels owned by `rank`; the last ranks may own fewer (or none)."""
    per = rows_per_rank(n_units, world)
    return min(rank * per, n_units), min((rank + 1) * per, n_units)


def make_comm(rank: int, world: int, device: int, group=None):
    """librxb200's own NCCL communicator for this rank (rx_tools_b200.power.Comm).  torch.distributed only carries
    the 128-byte unique id from rank 0 to the others -- plumbing; the all-gather itself runs inside the library."""
    import torch.distributed as dist
    from . import power
    box = [power.Comm.unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0, group=group)
    return power.Comm(world, rank, box[0], device)


def gather_rows(local_rows, n_units: int, row_len: int, world: int, group=None):
    """Host-logic stand-in of rxb200_power_gather for the CPU (gloo) tests: the same partition, padding and order.
    all_gather of equal-sized row blocks; local_rows: int64 tensor [(end-begin) * row_len] on the rank's
    device (cuda with nccl, cpu with gloo).  Returns the [n_units, row_len] tensor in unit order."""
    import torch
    import torch.distributed as dist
    per = rows_per_rank(n_units, world)
    send = torch.zeros(per * row_len, dtype=torch.int64, device=local_rows.device)
    send[: local_rows.numel()].copy_(local_rows.reshape(-1))
    out = torch.empty(world * per * row_len, dtype=torch.int64, device=local_rows.device)
    dist.all_gather_into_tensor(out, send, group=group)
    return out.view(world * per, row_len)[:n_units]
