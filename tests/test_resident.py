"""Resident state (prtc_resident_*): the records stay on the device, movement comes in and position + isGrounded go out
through mapped pinned arrays.  Bar: the same bits as prtc_step_characters on the same records, every frame."""
import numpy as np
import pytest

import helpers as H

pytestmark = pytest.mark.gpu

DT = 1.0 / 30.0


def _pair(scene, **kw):
    a, tf, ph = H.build_gpu(scene, **kw)
    b, _, _ = H.build_gpu(scene, **kw)
    return a, b, tf, ph


@pytest.mark.parametrize("n,dyn", [(50, "off"), (6000, "off"), (70000, "off"), (300, "jacobi"), (12000, "jacobi")])
def test_resident_steps_equal_host_buffer_steps(n, dyn):
    """warp-per-character kernel (n = 50), pooled kernel with fewer and with 32 owners per warp, and the dynamic pass;
    movement input changes every frame, some characters jump or are teleported in between"""
    from prototype2_b200 import capi
    scene = H.make_scene(96, 2, 2, n, seed=300 + n % 11)
    mode = capi.DYN_JACOBI if dyn == "jacobi" else capi.DYN_OFF
    host, res, tf, ph = _pair(scene, dyn_mode=mode)
    res.resident_begin(tf, ph)
    move, result = res.resident_io()
    assert move.shape == (n, 3) and result.shape == (n, 4)
    rs = np.random.RandomState(5)
    for f in range(8):
        mv = (rs.uniform(-0.15, 0.15, (n, 3)) * (rs.uniform(0, 1, (n, 1)) < 0.5)).astype(np.float32)
        ph["movement"] = mv
        move[:] = mv
        if f in (3, 5):                                            # jumps (character_system.cpp:92) and teleports
            idx = rs.choice(n, size=max(1, n // 10), replace=False).astype(np.uint32)
            ph["velocity"][idx, 1] = np.float32(0.35)
            tf["position"][idx[: len(idx) // 2], 0] += np.float32(1.5)
            res.resident_edit(idx, tf[idx].copy(), ph[idx].copy())
        host.update_characters(DT, tf, ph)
        res.resident_step(DT)
        assert H.same(result[:, :3], tf["position"]), "frame %d positions" % f
        assert np.array_equal(result[:, 3] != 0, ph["grounded"] != 0), "frame %d grounded" % f
    rtf, rph = res.resident_fetch()
    assert H.state_diff(H.gpu_state(tf, ph), H.gpu_state(rtf, rph)) == []
    assert H.same(rph["movement"], ph["movement"])
    ch, cr = host.counters(), res.counters()
    for k in ("substeps", "tri_tests", "contacts", "cc_tests"):
        assert ch[k] == cr[k], k
    res.resident_end()


def test_resident_by_id_order_and_oracle():
    """resident state against the CPU oracle directly (by-id order, tree built on the device)"""
    from prototype2_b200 import capi
    scene = H.make_scene(64, 2, 2, 5000, seed=17)
    orc = H.build_oracle(scene, "l1", order="by_id")
    ctx, tf, ph = H.build_gpu(scene, order_mode=capi.ORDER_BY_ID)
    ctx.resident_begin(tf, ph)
    _, result = ctx.resident_io()
    for f in range(5):
        orc.step(DT)
        ctx.resident_step(DT)
        assert H.same(result[:, :3], orc.get_state()["pos"]), "frame %d" % f
    rtf, rph = ctx.resident_fetch()
    assert H.state_diff(orc.get_state(), H.gpu_state(rtf, rph)) == []


def test_resident_errors():
    from prototype2_b200 import capi
    scene = H.make_scene(24, 2, 2, 64, seed=2)
    ctx, tf, ph = H.build_gpu(scene)
    with pytest.raises(capi.PrtcError):
        ctx.resident_step(DT)                                      # no resident state yet
    ph["collider"][5] = 0x7FFFFFF0
    ctx.resident_begin(tf, ph)
    with pytest.raises(capi.PrtcError) as e:
        ctx.resident_step(DT)
    assert e.value.code == capi.PRTC_ERR_INVALID and "never added" in str(e.value)
    with pytest.raises(capi.PrtcError):
        ctx.resident_edit(np.array([64], np.uint32), tf[:1].copy(), ph[:1].copy())       # index out of range
    lit, tf2, ph2 = H.build_gpu(scene, order_mode=capi.ORDER_LITERAL)
    with pytest.raises(capi.PrtcError):
        lit.resident_begin(tf2, ph2)                               # LITERAL order keeps host records
