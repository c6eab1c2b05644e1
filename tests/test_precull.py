"""Soundness of the conservative pre-culls in front of the exact capsule-triangle test (host build of collide.cuh).

tests/cpp/precull_check.cpp runs the very functions the kernels compile -- capsule_triangle (the restatement of
reference collision_system.cpp:56-142), the box cull and tight_reject -- on the CPU and counts triangles a cull would
have skipped although the exact test reports a contact.  No GPU, no oracle."""
import os
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "cpp", "precull_check.cpp")
BIN = os.path.join(HERE, "cpp", "_bin", "precull_check")


@pytest.fixture(scope="module")
def checker():
    os.makedirs(os.path.dirname(BIN), exist_ok=True)
    hdr = os.path.join(HERE, "..", "prototype2_b200", "csrc", "collide.cuh")
    if not os.path.exists(BIN) or os.path.getmtime(BIN) < max(os.path.getmtime(SRC), os.path.getmtime(hdr)):
        subprocess.check_call(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-x", "c++", "-o", BIN, SRC])
    return BIN


def _run(checker, *args, mult=None):
    env = dict(os.environ)
    env.pop("PRECULL_MULT", None)
    if mult is not None:
        env["PRECULL_MULT"] = str(mult)
    r = subprocess.run([checker] + [str(a) for a in args], env=env, capture_output=True, text=True, timeout=600)
    return r.returncode, r.stdout


@pytest.mark.parametrize("shift,abs_mode", [(0, 0), (0, 1), (2000, 0), (-30000, 1)])
def test_crowd_on_terrain_never_loses_a_contact(checker, shift, abs_mode):
    """capsules landing on and sliding over the SURVEY 8d terrain, near the origin and far from it"""
    rc, out = _run(checker, "sim", 128, 4000, 30, shift, abs_mode)
    assert rc == 0, out
    # and the second stage earns its keep: it leaves well under half of what the box test lets through
    f = out.split("after box cull ")[1]
    after_box, after_tight = float(f.split(",")[0]), float(f.split("after tight cull ")[1].split(",")[0])
    if abs(shift) <= 2000:                      # (30 000 units out the margin is 0.125, a quarter of the radius)
        assert after_tight < 0.5 * after_box, out


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_random_triangles_and_capsules_never_lose_a_contact(checker, seed):
    rc, out = _run(checker, "fuzz", 1500000, seed)
    assert rc == 0, out
    assert int(out.split("contacts ")[1].split(";")[0]) > 400000, out      # the generator does produce contacts


def test_the_checker_has_teeth(checker):
    """with the margin taken away the same generator does find false rejects"""
    rc, out = _run(checker, "fuzz", 1500000, 3, mult=0)
    assert rc == 1, out
