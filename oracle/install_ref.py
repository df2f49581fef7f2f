"""TEST / MEASUREMENT INFRASTRUCTURE ONLY.  Installs the UNMODIFIED reference package into baseline/_ref/ so that it
travels to the GPU box (baseline/_ref is git-ignored but not gpurun-ignored) and can be timed there as the CPU
baseline (`cpu_baseline.kind == "reference"`, SURVEY.md 8d / BASELINE.md section 3).

    python -m pip install --no-index --no-build-isolation --find-links /opt/wheelhouse --no-deps \
           --target baseline/_ref <copy of /root/reference>

(`--no-deps`: the reference's only declared dependency, `gym`, is not in the offline wheelhouse -- it is stubbed at
import time by oracle/ref_shim.py, like matplotlib; the install runs from a copy because /root/reference is read-only
and setuptools writes build/ and egg-info into the source tree.)  Nothing is copied into the repository history.
"""
import os
import shutil
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TARGET = os.path.join(ROOT, "baseline", "_ref")
SOURCE = "/root/reference"


def installed():
    return os.path.isfile(os.path.join(TARGET, "gym_SBR", "envs", "gym_SBR_env2.py"))


def install(force=False):
    """Returns the path of the installed reference, or None when the source checkout is absent (GPU box: the
    pre-installed copy that travelled with the snapshot is used)."""
    if installed() and not force:
        return TARGET
    if not os.path.isdir(os.path.join(SOURCE, "gym_SBR")):
        return TARGET if installed() else None
    tmp = tempfile.mkdtemp(prefix="sbr_ref_")
    try:
        src = os.path.join(tmp, "reference")
        shutil.copytree(SOURCE, src)
        if os.path.isdir(TARGET):
            shutil.rmtree(TARGET)
        os.makedirs(os.path.dirname(TARGET), exist_ok=True)
        cmd = [sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--find-links",
               "/opt/wheelhouse", "--no-deps", "--quiet", "--target", TARGET, src]
        try:
            subprocess.check_call(cmd, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        except (subprocess.CalledProcessError, OSError):
            # no usable pip/setuptools: the package is pure Python, a plain copy of the tree is the same install
            shutil.copytree(os.path.join(SOURCE, "gym_SBR"), os.path.join(TARGET, "gym_SBR"))
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    return TARGET if installed() else None


if __name__ == "__main__":
    print(install(force="--force" in sys.argv))
