"""Multi-GPU host logic on CPU: world slicing, rank-independent input generation and the final gather of rollout
states (gloo, world_size 2 and 3) — the N>1 path of bench.py without GPUs."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from box2d_sim_env_b200 import sharding


def test_world_slices_partition_the_batch():
    for n in (0, 1, 7, 4096, 1048576, 1000003):
        for ws in (1, 2, 3, 4, 8):
            edges = [sharding.world_slice(n, r, ws) for r in range(ws)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            assert all(edges[r][1] == edges[r + 1][0] for r in range(ws - 1))
            sizes = [hi - lo for lo, hi in edges]
            assert max(sizes) - min(sizes) <= 1 and sizes == sharding.slice_sizes(n, ws)
    with pytest.raises(ValueError):
        sharding.world_slice(10, 2, 2)


def test_seeded_rows_do_not_depend_on_the_gpu_count():
    draw = lambda rng, n: rng.uniform(-1, 1, (n, 7)).astype(np.float32)
    full = sharding.seeded_rows(1000, 0, 1000, 1234, draw)
    for ws in (2, 3, 8):
        parts = [sharding.seeded_rows(1000, *sharding.world_slice(1000, r, ws), 1234, draw) for r in range(ws)]
        assert np.array_equal(np.concatenate(parts), full)


def test_bench_inputs_are_slices_of_one_global_batch():
    import bench
    limits = np.zeros((11, 6), np.float32)
    limits[:, 2], limits[:, 3] = -2.0, 2.0
    cfg = bench.CONFIGS["cfg5"]
    q, qd, tg = bench.synth_inputs(limits, 64, 99, cfg)
    lo, hi = sharding.world_slice(64, 1, 3)
    q1, qd1, tg1 = bench.synth_inputs(limits, 64, 99, cfg, lo, hi)
    assert np.array_equal(q1, q[lo:hi]) and np.array_equal(tg1, tg[:, lo:hi]) and tg1.flags["C_CONTIGUOUS"]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, ws, port, n_total, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=ws)
    try:
        lo, hi = sharding.world_slice(n_total, rank, ws)
        # row w of the global state tensor is a pure function of the world index
        local = torch.arange(lo, hi, dtype=torch.float32)[:, None, None] * torch.ones((1, 2, 11)) + \
            torch.arange(22, dtype=torch.float32).view(1, 2, 11) / 100.0
        full = sharding.gather_states(local, n_total)
        expect = torch.arange(n_total, dtype=torch.float32)[:, None, None] * torch.ones((1, 2, 11)) + \
            torch.arange(22, dtype=torch.float32).view(1, 2, 11) / 100.0
        ok = torch.equal(full, expect)
        # max-over-ranks timing reduction used by bench.py
        t = torch.tensor([float(rank + 1)], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ok = ok and float(t.item()) == float(ws)
        open(os.path.join(out_dir, "rank%d" % rank), "w").write("ok" if ok else "bad")
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("ws,n_total", [(2, 64), (2, 33), (3, 100)])
def test_gather_states_gloo(tmp_path, ws, n_total):
    port = _free_port()
    mp.spawn(_worker, args=(ws, port, n_total, str(tmp_path)), nprocs=ws, join=True)
    for r in range(ws):
        assert open(os.path.join(str(tmp_path), "rank%d" % r)).read() == "ok"


def test_single_process_gather_is_identity():
    x = torch.arange(12, dtype=torch.float32).view(6, 2)
    assert sharding.gather_states(x, 6) is x
