"""Summarise `make EXTRA="-Xptxas -v" 2> build_log.txt`: registers / spills per kernel."""
import re, subprocess, sys
txt = open(sys.argv[1] if len(sys.argv) > 1 else "build_log.txt").read()
entries = re.findall(r"Compiling entry function '(\S+)' for 'sm_100a'\nptxas info\s+: Function properties for \S+\n\s+(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\nptxas info\s+: Used (\d+) registers", txt)
names = subprocess.run(["c++filt"], input="\n".join(e[0] for e in entries), capture_output=True, text=True).stdout.split("\n")
only = sys.argv[2] if len(sys.argv) > 2 else ""
for (mangled, stack, ss, sl, regs), name in zip(entries, names):
    name = re.sub(r"\(.*", "", name).replace("aecf::", "").replace("__nv_bfloat16", "bf16")
    if only in name:
        print(f"{name:60s} regs {regs:>3s} stack {stack:>4s} spill {ss}/{sl}")
