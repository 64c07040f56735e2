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


STUB_SRC = os.path.join(HERE, "abi_stub.c")
STUB_OUT = os.path.join(HERE, "libacoss_abi_stub.so")


def build_stub(force: bool = False) -> str:
    """CPU stand-in of four C-ABI entry points over the oracle (tests/test_reference_class.py only; see abi_stub.c)."""
    if (not force and os.path.exists(STUB_OUT)
            and os.path.getmtime(STUB_OUT) >= max(os.path.getmtime(STUB_SRC), os.path.getmtime(SRC))):
        return STUB_OUT
    cmd = ["gcc", "-O2", "-fPIC", "-shared", "-ffp-contract=off", "-fno-fast-math", "-pthread",
           "-o", STUB_OUT, STUB_SRC, SRC, "-lm"]
    subprocess.check_call(cmd)
    return STUB_OUT


if __name__ == "__main__":
    print(build(force=True))
