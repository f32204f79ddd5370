#!/usr/bin/env python3
"""oracle/build_oracle.py -- TEST INFRASTRUCTURE ONLY.

Compiles oracle/oracle_c.c (the C restatement of the reference's compiled loops) into
oracle/_build/liboracle_c.so with plain gcc.  Called by __graft_entry__.build() and, on demand,
by oracle/oracle_np.py when the library is missing or older than its source.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "oracle_c.c")
OUT_DIR = os.path.join(HERE, "_build")
OUT = os.path.join(OUT_DIR, "liboracle_c.so")


def build(force=False):
    if (not force and os.path.exists(OUT)
            and os.path.getmtime(OUT) >= os.path.getmtime(SRC)):
        return OUT
    os.makedirs(OUT_DIR, exist_ok=True)
    tmp = OUT + ".tmp%d" % os.getpid()
    # -ffp-contract=off: the restatement spells out every fma() it wants; nothing implicit.
    cmd = ["gcc", "-O2", "-fPIC", "-shared", "-std=c11", "-ffp-contract=off",
           "-o", tmp, SRC, "-lm"]
    subprocess.check_call(cmd)
    os.replace(tmp, OUT)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
