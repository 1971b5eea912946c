#!/usr/bin/env python
"""Install the UNMODIFIED reference (caburu/gym-supplychain) into baseline/_ref -- the CPU arm of bench.py.

    python baseline/install_ref.py [--force]

Exactly the offline recipe: `pip install --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse
--target baseline/_ref <copy of /root/reference>` (a copy under /tmp because the build writes egg-info into the source
tree and /root/reference is read-only; --no-deps because its only dependency besides numpy, `gym`, is neither in the image
nor in the wheelhouse -- the reference runs here and on the GPU box through the stand-in under oracle/gym_shim).
baseline/_ref is git-ignored (no reference source enters the history) but NOT gpurun-ignored: it travels to the GPU box,
where /root/reference does not exist.  After installing, a fast subset of the reference's OWN tests (every known-answer
and golden-return test; the RNG-fixture tests need tests/data/*.npy, which its setup.py does not package) must pass
through the shim, or the install is rejected.
"""
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF_SRC = os.environ.get("SCG_REFERENCE", "/root/reference")
DEST = os.path.join(HERE, "_ref")
SHIM = os.path.join(ROOT, "oracle", "gym_shim")
STAMP = os.path.join(DEST, ".selfcheck_ok")


def installed():
    return os.path.exists(os.path.join(DEST, "gym_supplychain", "envs", "supplychain_env.py")) and os.path.exists(STAMP)


def self_check():
    """The reference's own known-answer / golden-return tests, run from the installed copy through the gym stand-in."""
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([SHIM, DEST]))
    tests = os.path.join(DEST, "gym_supplychain", "envs", "tests")
    cmd = [sys.executable, "-m", "pytest", "-q", "-x", "-p", "no:cacheprovider", tests, "-k", "not consecutive_episodes"]
    res = subprocess.run(cmd, cwd=tempfile.gettempdir(), env=env, capture_output=True, text=True, timeout=900)
    tail = (res.stdout or "").strip().splitlines()[-1:] or [""]
    return res.returncode == 0, tail[0]


def install(force=False):
    if installed() and not force:
        return DEST
    if not os.path.isdir(os.path.join(REF_SRC, "gym_supplychain")):
        raise RuntimeError("reference tree %s not present (it only exists in the build container)" % REF_SRC)
    shutil.rmtree(DEST, ignore_errors=True)
    with tempfile.TemporaryDirectory() as tmp:
        src = os.path.join(tmp, "reference")
        shutil.copytree(REF_SRC, src)
        cmd = [sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps", "--quiet",
               "--find-links", "/opt/wheelhouse", "--target", DEST, src]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError("pip install of the reference failed:\n" + (res.stderr or res.stdout)[-2000:])
    ok, tail = self_check()
    if not ok:
        raise RuntimeError("the installed reference fails its own tests through the gym stand-in: " + tail)
    with open(STAMP, "w") as f:
        f.write(tail + "\n")
    return DEST


def import_reference():
    """sys.path plumbing for the CPU arm: returns the `gym` stand-in module with gym_supplychain's ids registered."""
    if not os.path.isdir(os.path.join(DEST, "gym_supplychain")):
        raise RuntimeError("baseline/_ref is missing: run `python baseline/install_ref.py` in the build container")
    for p in (DEST, SHIM):
        if p not in sys.path:
            sys.path.insert(0, p)
    import gym
    import gym_supplychain  # noqa: F401  (registers the env ids)
    return gym


if __name__ == "__main__":
    print(install(force="--force" in sys.argv))
    print(open(STAMP).read().strip())
