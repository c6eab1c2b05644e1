"""prtc_multi: several devices behind one handle, one process.  Bar: byte-identical to ONE context stepping the whole batch.
The same code runs with two contexts on cuda:0 (any box) and, when the box has them, with one context on each of two or
more GPUs (peer stores over NVLink for the dynamic pass's records)."""
import numpy as np
import pytest

import helpers as H

pytestmark = pytest.mark.gpu

DT = 1.0 / 30.0


def _device_sets():
    from prototype2_b200 import capi
    sets = [[0], [0, 0], [0, 0, 0]]
    n = capi.device_count()
    if n >= 2:
        sets.append(list(range(min(n, 8))))
    return sets


def _build_multi(scene, devices, **kw):
    from prototype2_b200 import capi
    m = capi.MultiContext(devices, **kw)
    m.add_model_collider(scene["xyz"], scene["mesh_first"], scene["mesh_count"], scene["model_transform"])
    h, r, off = scene["capsule"]
    for _ in range(scene["n"]):
        m.add_capsule_collider(h, r, off)
    return m


@pytest.mark.parametrize("dyn", ["off", "jacobi"])
@pytest.mark.parametrize("n", [45, 7001])
def test_multi_equals_one_context(dyn, n):
    from prototype2_b200 import capi
    scene = H.make_scene(64, 2, 2, n, seed=400 + n % 5)          # crowded: the dynamic pass has contacts across the slices
    mode = capi.DYN_JACOBI if dyn == "jacobi" else capi.DYN_OFF
    one, tf1, ph1 = H.build_gpu(scene, dyn_mode=mode)
    frames = 6
    ref = []
    for f in range(frames):
        one.update_characters(DT, tf1, ph1)
        ref.append(H.gpu_state(tf1, ph1))
    c1 = one.counters()
    for devices in _device_sets():
        m = _build_multi(scene, devices, dyn_mode=mode)
        _, tf, ph = H.build_gpu(scene, dyn_mode=mode)
        for f in range(frames):
            m.update_characters(DT, tf, ph)
            assert H.state_diff(ref[f], H.gpu_state(tf, ph)) == [], "devices %s frame %d" % (devices, f)
        cm = m.counters()
        for k in ("substeps", "tri_tests", "contacts", "cc_tests"):
            assert c1[k] == cm[k], (devices, k)
        m.close()


@pytest.mark.parametrize("dyn", ["off", "jacobi"])
def test_multi_resident_equals_one_context(dyn):
    from prototype2_b200 import capi
    n = 5003
    scene = H.make_scene(64, 2, 2, n, seed=77)
    mode = capi.DYN_JACOBI if dyn == "jacobi" else capi.DYN_OFF
    one, tf1, ph1 = H.build_gpu(scene, dyn_mode=mode)
    rs = np.random.RandomState(3)
    moves = [(rs.uniform(-0.1, 0.1, (n, 3))).astype(np.float32) for _ in range(5)]
    ref = []
    for mv in moves:
        ph1["movement"] = mv
        one.update_characters(DT, tf1, ph1)
        ref.append(H.gpu_state(tf1, ph1))
    for devices in _device_sets():
        m = _build_multi(scene, devices, dyn_mode=mode)
        _, tf, ph = H.build_gpu(scene, dyn_mode=mode)
        m.resident_begin(tf, ph)
        move, result = m.resident_io()
        for f, mv in enumerate(moves):
            move[:] = mv
            m.resident_step(DT)
            assert H.same(result[:, :3], ref[f]["pos"]), "devices %s frame %d" % (devices, f)
            assert np.array_equal(result[:, 3] != 0, ref[f]["grounded"] != 0)
        rtf, rph = m.resident_fetch()
        assert H.state_diff(ref[-1], H.gpu_state(rtf, rph)) == []
        m.close()


def test_multi_rejects_literal_order():
    from prototype2_b200 import capi
    with pytest.raises(capi.PrtcError):
        capi.MultiContext([0], order_mode=capi.ORDER_LITERAL)
