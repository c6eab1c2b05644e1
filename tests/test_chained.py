"""PRTC_OPT_CHAIN_FRAMES: consecutive device-resident frames of one batch overlap on the device (programmatic dependent
launch + a word per claim that orders a character's frames).  Same bytes as one launch after the other, and as the oracle."""
import numpy as np
import pytest

import helpers as H

pytestmark = pytest.mark.gpu

DT = 1.0 / 30.0
N = 100000            # > 148 x 16 claims of 32 characters: the batch fills the machine, which is what makes a launch chain


def _scene(moving):
    scene = H.make_scene(330, 2, 2, N, seed=77)
    if moving:
        ang = np.random.RandomState(3).uniform(-np.pi, np.pi, N)
        scene["move"] = np.stack([0.1 * np.cos(ang), np.zeros(N), 0.1 * np.sin(ang)], axis=1).astype(np.float32)
    return scene


def _device_run(scene, frames, chain, interrupt_at=None):
    import torch
    from prototype2_b200 import capi
    ctx, tf, ph = H.build_gpu(scene)
    ctx.set_option(capi.OPT_CHAIN_FRAMES, 1 if chain else 0)
    d_tf = torch.from_numpy(tf.view(np.uint8).reshape(-1).copy()).cuda()
    d_ph = torch.from_numpy(ph.view(np.uint8).reshape(-1).copy()).cuda()
    torch.cuda.synchronize()
    ctx.counters(reset=True)
    for f in range(frames):
        ctx.update_characters_device(DT, scene["n"], d_tf.data_ptr(), d_ph.data_ptr())
        if interrupt_at is not None and f == interrupt_at:
            # something else of the library in between (four rays): the chain breaks and starts again
            o = np.array([[10.0, 50.0, 10.0]] * 4, np.float32)
            d = np.array([[0.0, -1.0, 0.0]] * 4, np.float32)
            ctx.raycast(o, d, np.full(4, 100.0, np.float32))
            # ... and a host-buffer step of ten other characters, which also rewrites the first ten leaf-list slots
            tf10, ph10 = tf[500:510].copy(), ph[500:510].copy()
            ph10["collider"] = np.arange(10, dtype=np.uint32)
            ctx.update_characters(DT, tf10, ph10)
            ctx.counters(reset=False)
    ctx.synchronize()
    got_tf = d_tf.cpu().numpy().view(capi.TRANSFORM_DTYPE)
    got_ph = d_ph.cpu().numpy().view(capi.PHYSICS_DTYPE)
    return H.gpu_state(got_tf, got_ph), ctx.counters()


@pytest.mark.parametrize("moving", [False, True])
def test_chained_frames_equal_unchained_and_the_oracle(moving):
    import os
    scene = _scene(moving)
    frames = 6
    orc = H.build_oracle(scene, "l1", order="canonical", threads=os.cpu_count() or 1)
    for _ in range(frames):
        orc.step(DT)
    want = orc.get_state()
    plain, c_plain = _device_run(scene, frames, chain=False)
    chained, c_chain = _device_run(scene, frames, chain=True)
    assert H.state_diff(want, plain) == []
    assert H.state_diff(want, chained) == []
    co = orc.counters()
    for k in ("substeps", "tri_tests", "contacts"):
        assert co[k] == c_plain[k] == c_chain[k], k
    # one launch per chained frame (the slot refresh is part of it); a refresh and a step otherwise
    assert c_chain["launches"] == frames and c_plain["launches"] == 2 * frames


@pytest.mark.parametrize("n", [30000, 64000])
def test_chained_frames_of_a_batch_that_does_not_fill_the_machine(n):
    """fewer claims than warp slots (one claim per block): every queued frame is on the machine at once, each block waiting
    for the same characters' claim of the frame before it"""
    import os
    scene = H.make_scene(200, 2, 2, n, seed=79)
    ang = np.random.RandomState(6).uniform(-np.pi, np.pi, n)
    scene["move"] = np.stack([0.1 * np.cos(ang), np.zeros(n), 0.1 * np.sin(ang)], axis=1).astype(np.float32)
    frames = 12
    orc = H.build_oracle(scene, "l1", order="canonical", threads=os.cpu_count() or 1)
    for _ in range(frames):
        orc.step(DT)
    chained, c_chain = _device_run(scene, frames, chain=True)
    assert H.state_diff(orc.get_state(), chained) == []
    assert c_chain["launches"] == frames
    co = orc.counters()
    for k in ("substeps", "tri_tests", "contacts"):
        assert co[k] == c_chain[k], k


def test_chain_broken_by_another_call_and_resumed():
    scene = _scene(True)
    plain, _ = _device_run(scene, 5, chain=False)
    chained, _ = _device_run(scene, 5, chain=True, interrupt_at=2)
    assert H.state_diff(plain, chained) == []


def test_chained_soak_long_run_equals_unchained():
    """150 frames of a walking crowd of 300 000 (four rounds of claims: every frame overlaps its neighbours), one shared capsule
    shape; chained and unchained runs must leave the same bytes, and so must a chained run that is synchronised every 7th frame"""
    import torch
    from prototype2_b200 import capi
    n = 300000
    scene = H.make_scene(330, 2, 2, n, seed=78)
    ang = np.random.RandomState(4).uniform(-np.pi, np.pi, n)
    move = np.stack([0.06 * np.cos(ang), np.zeros(n), 0.06 * np.sin(ang)], axis=1).astype(np.float32)

    def run(chain, sync_every=0):
        ctx = capi.PhysicsContext()
        ctx.add_model_collider(scene["xyz"], scene["mesh_first"], scene["mesh_count"])
        cid = ctx.add_capsule_collider()
        ctx.set_option(capi.OPT_CHAIN_FRAMES, 1 if chain else 0)
        tf = capi.make_transforms(scene["pos"])
        ph = capi.make_physics(scene["vel"], movement=move, collider=np.full(n, cid, np.uint32))
        d_tf = torch.from_numpy(tf.view(np.uint8).reshape(-1).copy()).cuda()
        d_ph = torch.from_numpy(ph.view(np.uint8).reshape(-1).copy()).cuda()
        torch.cuda.synchronize()
        for f in range(150):
            ctx.update_characters_device(DT, n, d_tf.data_ptr(), d_ph.data_ptr())
            if sync_every and f % sync_every == 0:
                ctx.synchronize()
        ctx.synchronize()
        return d_tf.cpu().numpy().tobytes(), d_ph.cpu().numpy().tobytes(), ctx.counters()

    a = run(False)
    b = run(True)
    c = run(True, sync_every=7)
    assert a[0] == b[0] and a[1] == b[1]
    assert a[0] == c[0] and a[1] == c[1]
    for k in ("substeps", "tri_tests", "contacts", "exact_tests"):
        assert a[2][k] == b[2][k] == c[2][k], k
    assert b[2]["launches"] == 150 and a[2]["launches"] == 300
