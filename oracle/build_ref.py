"""Recipe for oracle/_ref/: byte-compile the UNMODIFIED reference modules of the hot path where they lie under
/root/reference (build container only) so that `bench.py --impl reference` and the `cpu_baseline` leg can time the
reference's own implementation on the GPU box's host cores, where /root/reference does not exist.

Outputs only (CPython bytecode, `<module>.pyc`, importable as sourceless modules by the same interpreter version the
image ships); no reference source is copied.  oracle/_ref/ is git-ignored and NOT gpurun-ignored, like the built .so.
Test infrastructure: nothing under flowtron_b200/ imports it.
"""
from __future__ import annotations

import os
import py_compile
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
REF = os.environ.get("FLOWTRON_REFERENCE", "/root/reference")
MODULES = ["flowtron", "audio_processing", "radam"]     # flowtron.py (model + loss), the STFT front-end, the optimizer


def build(verbose: bool = False) -> bool:
    """Returns True if oracle/_ref/ is complete afterwards.  A no-op (keeping what is there) when the reference tree is
    absent, i.e. on the GPU box, which uses the prebuilt files."""
    if not os.path.isfile(os.path.join(REF, "flowtron.py")):
        return all(os.path.isfile(os.path.join(OUT, m + ".pyc")) for m in MODULES)
    os.makedirs(OUT, exist_ok=True)
    for m in MODULES:
        src, dst = os.path.join(REF, m + ".py"), os.path.join(OUT, m + ".pyc")
        if not os.path.exists(dst) or os.path.getmtime(dst) < os.path.getmtime(src):
            py_compile.compile(src, cfile=dst, dfile=f"<reference>/{m}.py", doraise=True)
            if verbose:
                print("compiled", src, "->", dst)
    with open(os.path.join(OUT, "PYTHON_VERSION"), "w") as f:
        f.write(sys.version.split()[0] + "\n")
    return True


if __name__ == "__main__":
    print("oracle/_ref complete:", build(verbose=True))
