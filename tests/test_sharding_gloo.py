"""N > 1 host logic on CPU: world_size-2 gloo processes.  (1) the partition + in-place record all-gather plumbing of
prototype2_b200/sharding.py; (2) partition independence of the static pass: each rank steps only its slice of the
characters (here with the CPU oracle standing in for the device) against replicated geometry, and the gathered result
is byte-identical to the unsharded run."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    for p in (ROOT, os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from prototype2_b200 import sharding
    import helpers as H
    n_global = 64
    first, count = sharding.partition(n_global, world, rank)
    # (1) in-place all-gather of 64-byte records
    rec = torch.zeros(n_global, sharding.RECORD_FLOATS)
    rec[first:first + count] = torch.arange(first, first + count, dtype=torch.float32)[:, None] + torch.arange(16)[None, :] / 100.0
    sharding.allgather_records(rec, first, count)
    want = torch.arange(n_global, dtype=torch.float32)[:, None] + torch.arange(16)[None, :] / 100.0
    assert torch.equal(rec, want)
    # (2) static pass: slice stepped alone == the same rows of the unsharded run
    scene = H.make_scene(32, 2, 2, n_global, seed=5, separated=True)
    sub = dict(scene, pos=scene["pos"][first:first + count], vel=scene["vel"][first:first + count], rot=scene["rot"][first:first + count],
               move=scene["move"][first:first + count], n=count)
    w = H.build_oracle(sub, "l1", order="canonical")
    for _ in range(5):
        w.step(1.0 / 30.0)
    s = w.get_state()
    rows = torch.from_numpy(np.concatenate([s["pos"], s["vel"], s["gnormal"], s["grounded"][:, None].astype(np.float32)], axis=1))
    full = sharding.gather_rows(rows)
    if rank == 0:
        np.save(os.path.join(out_dir, "sharded.npy"), full.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_partition_and_gather(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    sharded = np.load(os.path.join(str(tmp_path), "sharded.npy"))
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import helpers as H
    scene = H.make_scene(32, 2, 2, 64, seed=5, separated=True)
    w = H.build_oracle(scene, "l1", order="canonical")
    for _ in range(5):
        w.step(1.0 / 30.0)
    s = w.get_state()
    whole = np.concatenate([s["pos"], s["vel"], s["gnormal"], s["grounded"][:, None].astype(np.float32)], axis=1)
    assert sharded.shape == whole.shape and np.array_equal(sharded.view(np.uint32), whole.view(np.uint32))


def test_partition_rejects_uneven_splits():
    from prototype2_b200 import sharding
    assert sharding.partition(256000, 8, 3) == (96000, 32000)
    with pytest.raises(ValueError):
        sharding.partition(10, 4, 0)
