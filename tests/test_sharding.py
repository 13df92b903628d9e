"""N>1 host logic on CPU: world_size-2 gloo processes shard rx_power hops / rx_fm channels exactly as
bench.py does on GPUs; the per-rank compute is stood in for by the port oracle (tests may use it), the
partition + padded all_gather + ordering is the product code under test (rx_tools_b200/sharding.py)."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_hops, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    import oracle
    from rx_tools_b200 import power, sharding, synth
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    plan = power.plan_range("24M:60M:1k", 0.285)
    assert plan.n_hops == n_hops
    n = 1 << plan.bin_e
    win = power.window_table("hamming", n)
    x = synth.power_hops(2, n_hops, plan.buf_len, seed=4000)       # every rank can rebuild the full input
    hb, he = sharding.unit_range(rank, world, n_hops)
    pp = oracle.PowerParams(bin_e=plan.bin_e, buf_len=plan.buf_len)
    local = np.zeros((max(he - hb, 0), n), dtype=np.int64)
    if he > hb:
        local, _ = oracle.port().power_scan(pp, win, np.ascontiguousarray(x[:, hb:he]), 2, he - hb)
    rows = sharding.gather_rows(torch.from_numpy(local.reshape(-1)), n_hops, n, world)
    if rank == 0:
        full, _ = oracle.port().power_scan(pp, win, x, 2, n_hops)
        q.put(bool(np.array_equal(rows.numpy(), full)))
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_hop_sharding_gathers_rows_in_order(world):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, 18, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=240)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True


def test_unit_ranges_cover_everything():
    from rx_tools_b200 import sharding
    for n in (1, 7, 18, 256, 871):
        for world in (1, 2, 3, 4, 8):
            got = []
            for r in range(world):
                b, e = sharding.unit_range(r, world, n)
                assert 0 <= b <= e <= n
                got.extend(range(b, e))
            assert got == list(range(n))
            assert sharding.rows_per_rank(n, world) * world >= n
