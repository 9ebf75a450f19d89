"""CPU ORACLE build recipe (test infrastructure only).

Compiles, from the sources WHERE THEY LIE under /root/reference (nothing is copied into
the repo), the pieces of the reference that are plain C:

* ``lib/csparse/instructionsTableUTHash.c`` -> ``oracle/_ref/libitable.so``
  (instruction store + the C emitter ``writeCinstructionsC``), with ``-DIGNORE_MEX`` and
  the 6-line ``oracle/stub/mex.h``;
* per problem, the generated translation unit that ``#include``s the unmodified
  ``lib/ipmPD_CSsolver.c`` -> ``oracle/_ref/<name>_ref.so`` (see ``oracle.py``).

``oracle/_ref/`` is git-ignored but travels to the GPU box with the snapshot; the GPU box
has no /root/reference, so everything the `-m gpu` tests and bench.py need from the
reference must have been built here beforehand (``__graft_entry__.build()`` does that).
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"
OUT = os.path.join(HERE, "_ref")


def build_itable(force: bool = False) -> str:
    os.makedirs(OUT, exist_ok=True)
    so = os.path.join(OUT, "libitable.so")
    src = os.path.join(REF, "lib/csparse/instructionsTableUTHash.c")
    if not os.path.exists(src):
        if os.path.exists(so):
            return so
        raise FileNotFoundError("reference sources not available and no prebuilt libitable.so")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        cmd = ["gcc", "-O1", "-w", "-shared", "-fPIC", "-DIGNORE_MEX", "-DDYNAMIC_LIBRARY",
               "-I" + os.path.join(HERE, "stub"), "-I" + os.path.join(REF, "lib/csparse"), src, "-o", so, "-lm"]
        subprocess.check_call(cmd)
    return so


if __name__ == "__main__":
    print(build_itable(force="--force" in sys.argv))
