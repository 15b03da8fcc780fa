"""Build recipe for the CPU oracle (TEST INFRASTRUCTURE, not product code).

    python oracle/build_oracle.py

compiles oracle/zutis_oracle.c into oracle/libzutis_oracle.so with gcc.  The reference
(NoelShin/zutis) is pure Python, so there is nothing to compile into oracle/_ref/; the
"reference" used to pin this oracle is the Python code itself, imported in the build
container by tests/golden/make_golden.py.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "zutis_oracle.c")
OUT = os.path.join(HERE, "libzutis_oracle.so")


def build(force: bool = False) -> str:
    if (not force and os.path.exists(OUT)
            and os.path.getmtime(OUT) >= os.path.getmtime(SRC)):
        return OUT
    cmd = ["gcc", "-O2", "-mfma", "-ffp-contract=off", "-fopenmp", "-fPIC", "-shared",
           "-fvisibility=hidden", "-o", OUT, SRC, "-lm"]
    subprocess.run(cmd, check=True)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
