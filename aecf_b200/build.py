"""Build csrc/libaecf_b200.so in-tree for sm_100a:  python -m aecf_b200.build [-jN] [--clean]"""
from __future__ import annotations

import os
import subprocess
import sys

CSRC = os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc")


def build(jobs: int | None = None, clean: bool = False, verbose: bool = False) -> str:
    jobs = jobs or max(1, (os.cpu_count() or 4))
    if clean:
        subprocess.run(["make", "-C", CSRC, "clean"], check=True)
    cmd = ["make", "-C", CSRC, f"-j{jobs}"]
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if verbose or res.returncode != 0:
        sys.stdout.write(res.stdout)
    if res.returncode != 0:
        raise RuntimeError(f"building libaecf_b200.so failed (exit {res.returncode})")
    return os.path.join(CSRC, "libaecf_b200.so")


if __name__ == "__main__":
    j = next((int(a[2:]) for a in sys.argv[1:] if a.startswith("-j") and a[2:].isdigit()), None)
    print(build(j, clean="--clean" in sys.argv, verbose=True))
