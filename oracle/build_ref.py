"""Recipe: install the UNMODIFIED reference package into oracle/_ref (TEST INFRASTRUCTURE / baseline only).

    python oracle/build_ref.py            # needs /root/reference (the build container)

`/root/reference` does not exist on the GPU box, and the reference is plain Python (nothing to compile), so the
only way its own code can be the `--impl reference` arm of bench.py there is to carry an installed copy:

    pip install --no-index --no-build-isolation --no-deps --target oracle/_ref <copy of /root/reference>

(from a copy under /tmp because setuptools writes build files into the source tree and /root/reference is
read-only; --no-deps because azure-* / datasets / python-dotenv are not in the offline wheelhouse -- the harness
stubs the two that are imported, see oracle/ref_harness.py).  oracle/_ref/ is git-ignored (no reference source
enters the history) but NOT gpurun-ignored, so it travels to the GPU box like the built .so files.  Nothing in the
product package imports it; `bench.py --impl reference`, the `cpu_baseline` / `reference_gpu` legs and tests/ do.
"""
from __future__ import annotations

import filecmp
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("NRB_REFERENCE_ROOT", "/root/reference")
OUT = os.path.join(HERE, "_ref")


def installed() -> bool:
    return os.path.isfile(os.path.join(OUT, "news_rec_utils", "data_model_helper.py"))


def up_to_date() -> bool:
    src = os.path.join(REF, "src", "news_rec_utils")
    if not installed() or not os.path.isdir(src):
        return installed()
    names = [f for f in os.listdir(src) if f.endswith(".py")]
    match, mismatch, errors = filecmp.cmpfiles(src, os.path.join(OUT, "news_rec_utils"), names, shallow=False)
    return not mismatch and not errors


def build(force: bool = False) -> str | None:
    """Returns the install directory, or None when the reference tree is absent (GPU box: use what travelled)."""
    if not os.path.isdir(os.path.join(REF, "src", "news_rec_utils")):
        return OUT if installed() else None
    if not force and up_to_date():
        return OUT
    with tempfile.TemporaryDirectory(prefix="nrb_ref_") as tmp:
        work = os.path.join(tmp, "reference")
        shutil.copytree(REF, work)
        if os.path.isdir(OUT):
            shutil.rmtree(OUT)
        cmd = [sys.executable, "-m", "pip", "install", "--quiet", "--no-index", "--no-build-isolation", "--no-deps",
               "--find-links", "/opt/wheelhouse", "--target", OUT, work]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("pip install of the reference failed:\n" + r.stdout + r.stderr)
    shutil.rmtree(os.path.join(OUT, "news_rec_utils", "__pycache__"), ignore_errors=True)
    if not up_to_date():
        raise RuntimeError("installed reference differs from /root/reference/src")
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
