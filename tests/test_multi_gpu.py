"""rx_power on several GPUs: hops sharded over GPUs, ONE in-place NCCL all-gather inside librxb200
(rxb200_power_gather / rxb200_power_group_gather) collating the rows in hop order for the report loop
(src/rtl_power.c:1047-1050).  The gathered rows must be byte-identical to the single-GPU rows and to the oracle.
Skipped on a box with one GPU; the partition/ordering logic itself is covered on CPU by tests/test_sharding.py."""
import os
import sys

import numpy as np
import pytest

import oracle
from rx_tools_b200 import power, synth

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _n_gpus():
    from rx_tools_b200 import _lib
    return int(_lib.lib().rxb200_device_count())


def _case():
    plan = power.plan_range("24M:60M:1k", 0.285)           # 18 hops x 4096 bins, the cfg4 shape in small
    win = power.window_table("hamming", 1 << plan.bin_e)
    x = np.concatenate([synth.power_hops(3, 1, plan.buf_len, seed=4000 + h) for h in range(plan.n_hops)], axis=1)
    return plan, win, np.ascontiguousarray(x)


@pytest.mark.parametrize("n_dev", [2, 3, 4, 8])
def test_group_gather_matches_single_gpu_and_oracle(n_dev, port):
    if _n_gpus() < n_dev:
        pytest.skip(f"needs {n_dev} GPUs")
    plan, win, x = _case()
    one = power.PowerScanner(plan, win, device=0)
    one.scanner(x, 3)
    avg1, smp1 = one.read()
    one.close()
    want, wsmp = port.power_scan(oracle.PowerParams(bin_e=plan.bin_e, buf_len=plan.buf_len), win, x, 3, plan.n_hops)
    assert np.array_equal(avg1, want) and np.array_equal(smp1, wsmp)
    g = power.PowerGroup(plan, win, n_dev=n_dev)
    g.scanner(x, 3)
    # before the collation a member only holds its own hops
    hb, he = power.shard(plan.n_hops, n_dev, 1)
    a_before, _ = g.read(1)
    assert np.array_equal(a_before[hb:he], want[hb:he]) and not a_before[:hb].any()
    g.gather()
    for m in range(n_dev):
        a, s = g.read(m)
        assert a.tobytes() == want.tobytes(), f"member {m}: gathered rows differ"
        assert np.array_equal(s, wsmp)
    # a second report interval: reset, accumulate a sub-range in two calls, gather again
    g.reset()
    g.scanner(x[:2, :7], 2, 0, 7)
    g.scanner(x[:2, 7:], 2, 7, plan.n_hops)
    g.gather()
    want2, wsmp2 = port.power_scan(oracle.PowerParams(bin_e=plan.bin_e, buf_len=plan.buf_len), win, x[:2], 2, plan.n_hops)
    a, s = g.read(0)
    assert a.tobytes() == want2.tobytes() and np.array_equal(s, wsmp2)
    g.close()


def _rank_main(rank, world, q_id, q_out):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from rx_tools_b200 import power as pw
    plan, win, x = _case()
    if rank == 0:
        uid = pw.Comm.unique_id()
        for _ in range(world - 1):
            q_id.put(uid)
    else:
        uid = q_id.get(timeout=120)
    comm = pw.Comm(world, rank, uid, rank)                 # one process per GPU, as under torchrun
    sc = pw.PowerScanner(plan, win, device=rank)
    hb, he = pw.shard(plan.n_hops, world, rank)
    if he > hb:
        sc.scanner(x[:, hb:he], 3, hb, he)
    sc.gather(comm, sync=True)
    avg, smp = sc.read()
    q_out.put((rank, avg.tobytes(), smp.tobytes()))
    sc.close()
    comm.close()


@pytest.mark.parametrize("world", [2, 4])
def test_one_process_per_gpu_gather(world, port):
    if _n_gpus() < world:
        pytest.skip(f"needs {world} GPUs")
    import multiprocessing as mp
    ctx = mp.get_context("spawn")
    q_id, q_out = ctx.Queue(), ctx.Queue()
    procs = [ctx.Process(target=_rank_main, args=(r, world, q_id, q_out)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q_out.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    plan, win, x = _case()
    want, wsmp = port.power_scan(oracle.PowerParams(bin_e=plan.bin_e, buf_len=plan.buf_len), win, x, 3, plan.n_hops)
    for rank, a, s in got:
        assert a == want.tobytes(), f"rank {rank}: gathered rows differ from the oracle"
        assert s == wsmp.astype(np.int32).tobytes()
