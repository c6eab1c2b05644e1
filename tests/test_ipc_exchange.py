"""The record exchange of the dynamic pass ACROSS PROCESSES without a collective (prtc_dyn_ipc_*): two processes, one context
each, CUDA IPC handles swapped once over gloo; every frame each rank's k_dyn_prep stores its records into the other rank's
array and rings its arrival word, the other rank's step waits for it on the device.  Bar: byte-identical to one context
stepping the whole batch.  Runs with both processes on cuda:0 and, when the box has two GPUs, on one GPU each."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu
DT = 1.0 / 30.0
FRAMES = 8


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _scene():
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import helpers as H
    return H, H.make_scene(64, 2, 2, 6000, seed=123)


def _worker(rank, world, port, out_dir, devices):
    for p in (ROOT, os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from prototype2_b200 import capi, sharding
    H, scene = _scene()
    n = scene["n"]
    first, count = sharding.partition(n, world, rank)
    dev = devices[rank]
    torch.cuda.set_device(dev)
    ctx = capi.PhysicsContext(device=dev, dyn_mode=capi.DYN_JACOBI)
    ctx.dyn_configure(first, n)
    ctx.add_model_collider(scene["xyz"], scene["mesh_first"], scene["mesh_count"])
    cid = ctx.add_capsule_collider()
    ctx.commit()
    assert sharding.ipc_connect(ctx, rank, world)
    tf = capi.make_transforms(scene["pos"][first:first + count])
    ph = capi.make_physics(scene["vel"][first:first + count], collider=np.full(count, cid, np.uint32))
    d_tf = torch.from_numpy(tf.view(np.uint8).reshape(-1)).cuda()
    d_ph = torch.from_numpy(ph.view(np.uint8).reshape(-1)).cuda()
    stream = torch.cuda.current_stream()
    ctx.set_stream(stream.cuda_stream)
    out = []
    for f in range(FRAMES):
        if rank == 1 and f == 3:
            import time
            time.sleep(0.3)                                        # one rank falls behind: the other one has to wait on the device
        ctx.update_characters_device(DT, count, d_tf.data_ptr(), d_ph.data_ptr())
        torch.cuda.synchronize()
        out.append(np.concatenate([d_tf.cpu().numpy().reshape(count, 40), d_ph.cpu().numpy().reshape(count, 44)], axis=1))
    ctx.synchronize()
    np.save(os.path.join(out_dir, "rank%d.npy" % rank), np.stack(out))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("spread", [False, True])
def test_two_processes_exchange_records_without_a_collective(tmp_path, spread):
    import torch
    import torch.multiprocessing as mp
    from prototype2_b200 import capi
    if spread and capi.device_count() < 2:
        pytest.skip("needs two GPUs")
    devices = [0, 1] if spread else [0, 0]
    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path), devices), nprocs=2, join=True)
    got = np.concatenate([np.load(os.path.join(str(tmp_path), "rank%d.npy" % r)) for r in range(2)], axis=1)
    H, scene = _scene()
    n = scene["n"]
    ctx = capi.PhysicsContext(dyn_mode=capi.DYN_JACOBI)
    ctx.add_model_collider(scene["xyz"], scene["mesh_first"], scene["mesh_count"])
    cid = ctx.add_capsule_collider()
    tf = capi.make_transforms(scene["pos"])
    ph = capi.make_physics(scene["vel"], collider=np.full(n, cid, np.uint32))
    for f in range(FRAMES):
        ctx.update_characters(DT, tf, ph)
        want = np.concatenate([tf.view(np.uint8).reshape(n, 40), ph.view(np.uint8).reshape(n, 44)], axis=1)
        assert np.array_equal(got[f], want), "frame %d" % f
    assert ctx.counters()["cc_tests"] > 1000
