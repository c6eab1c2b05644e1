"""Golden parses of the `.prt` fixtures in tests/golden/scenes/.

The fixtures are our own files in the reference's scene grammar; each `<name>.dump` is what the REFERENCE's parser
(src/game/scene/scene_serialization.cpp compiled unmodified, oracle/_ref/libprt_l0_scene.so) did with `<name>.prt`,
rendered by oracle/l0_scene_driver.cpp.  Also writes cavern.dump, the reference parser's reading of the shipped level
res/scenes/cavern.prt (the .prt itself stays in /root/reference).  Run in the dev container:
    python tests/golden/make_scene_dumps.py
"""
import ctypes
import glob
import os

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))


def l0_dump(path):
    lib = ctypes.CDLL(os.path.join(ROOT, "oracle", "_ref", "libprt_l0_scene.so"))
    lib.l0s_dump.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_size_t, ctypes.POINTER(ctypes.c_size_t)]
    need = ctypes.c_size_t()
    lib.l0s_dump(path.encode(), None, 0, ctypes.byref(need))
    buf = ctypes.create_string_buffer(need.value)
    lib.l0s_dump(path.encode(), buf, need.value, None)
    return buf.value.decode()


if __name__ == "__main__":
    for prt in sorted(glob.glob(os.path.join(HERE, "scenes", "*.prt"))):
        out = prt[:-4] + ".dump"
        open(out, "w").write(l0_dump(prt))
        print("wrote", os.path.relpath(out, ROOT))
    shipped = "/root/reference/res/scenes/cavern.prt"
    if os.path.exists(shipped):
        out = os.path.join(HERE, "scenes", "cavern.dump")
        open(out, "w").write(l0_dump(shipped))
        print("wrote", os.path.relpath(out, ROOT))
