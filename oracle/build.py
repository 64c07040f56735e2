"""Compile the plain-C oracle (oracle/serra09_c.c) into oracle/liboracle_serra09.so.

Called by __graft_entry__.build().  No reference source is compiled here: the reference is
pure Python plus the absent essentia dependency, so there is no oracle/_ref."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "serra09_c.c")
OUT = os.path.join(HERE, "liboracle_serra09.so")


def build(force: bool = False) -> str:
    if (not force and os.path.exists(OUT)
            and os.path.getmtime(OUT) >= os.path.getmtime(SRC)):
        return OUT
    cmd = ["gcc", "-O2", "-fPIC", "-shared", "-ffp-contract=off", "-fno-fast-math", "-pthread",
           "-o", OUT, SRC, "-lm"]
    subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    print(build(force=True))
