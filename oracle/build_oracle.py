"""gcc build of oracle/sn_oracle_c.c -> oracle/_build/libsn_oracle.so (test infrastructure)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "sn_oracle_c.c")
OUT = os.path.join(HERE, "_build", "libsn_oracle.so")


def build(force: bool = False) -> str:
    if not force and os.path.exists(OUT) and os.path.getmtime(OUT) >= os.path.getmtime(SRC):
        return OUT
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    # no -ffast-math and no FMA contraction: keep the numpy oracle's rounding
    subprocess.run(["gcc", "-O3", "-ffp-contract=off", "-fPIC", "-shared", "-o", OUT, SRC], check=True)
    return OUT


if __name__ == "__main__":
    print(build(force=True))
