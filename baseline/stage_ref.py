"""Stage the UNMODIFIED reference (getkeops/keops v2.3: ``keopscore`` + ``pykeops``, pure Python) into
``baseline/_ref/`` -- git-ignored, but it travels to the GPU box with the repository snapshot.

    python baseline/stage_ref.py            (in the build container, where /root/reference is mounted)

It is the documented offline install ``pip install --no-index --no-build-isolation --no-deps --target baseline/_ref``
of the two reference packages, from a copy under /tmp because /root/reference is read-only and setuptools writes
``build/`` and ``*.egg-info`` next to ``setup.py``.  Nothing of it is imported by the product: ``baseline/_ref`` is used

* by ``tests/test_reference_frontend_gpu.py``: the reference's own front-end (``pykeops.torch`` ``Genred`` / ``LazyTensor`` /
  ``KernelSolve``) and its own test files run on the B200 with ``keops_binder["nvrtc"] = LoadKeOpsB200``;
* by ``tests/test_reference_seam.py`` when /root/reference itself is not mounted.
"""
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE = "/root/reference"
TARGET = os.path.join(HERE, "_ref")


def stage_core_tests():
    """``keopscore/setup.py`` does not list ``keopscore.test`` among its packages, so the wheel leaves out the reference's
    per-operation harness (``keopscore/keopscore/test/test_op.py``); copy it next to the installed package, unmodified."""
    src = os.path.join(REFERENCE, "keopscore", "keopscore", "test")
    dst = os.path.join(TARGET, "keopscore", "test")
    if os.path.isdir(src) and os.path.isdir(os.path.join(TARGET, "keopscore")) and not os.path.isdir(dst):
        shutil.copytree(src, dst, ignore=shutil.ignore_patterns("__pycache__"))


def main():
    if not os.path.isdir(REFERENCE):
        print("stage_ref: %s is not mounted; keeping the existing baseline/_ref" % REFERENCE)
        return 0
    if os.path.exists(os.path.join(TARGET, "pykeops", "keops_version")) and "--force" not in sys.argv:
        stage_core_tests()
        print("stage_ref: baseline/_ref already staged (use --force to redo)")
        return 0
    tmp = tempfile.mkdtemp(prefix="kb200_refsrc_")
    try:
        for pkg in ("keopscore", "pykeops"):
            shutil.copytree(os.path.join(REFERENCE, pkg), os.path.join(tmp, pkg))
        if os.path.isdir(TARGET):
            shutil.rmtree(TARGET)
        cmd = [sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps", "--find-links",
               "/opt/wheelhouse", "--target", TARGET, os.path.join(tmp, "keopscore"), os.path.join(tmp, "pykeops")]
        subprocess.check_call(cmd, stdout=subprocess.DEVNULL)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    stage_core_tests()
    for root, dirs, _ in os.walk(TARGET):
        for d in dirs:
            if d == "__pycache__":
                shutil.rmtree(os.path.join(root, d), ignore_errors=True)
    print("stage_ref: reference staged into", TARGET)
    return 0


if __name__ == "__main__":
    sys.exit(main())
