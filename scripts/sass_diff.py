"""Per-function SASS comparison of two builds: proves that a change left kernels measured on hardware untouched.

    python scripts/sass_diff.py snapshot DIR          # cuobjdump -sass of every aecf_b200/csrc/build/*.o into DIR
    python scripts/sass_diff.py compare BEFORE AFTER   # functions whose instruction stream differs, added, removed

Instruction text only (addresses, encodings and comments are dropped), keyed by mangled function name.
"""
import glob
import hashlib
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, "aecf_b200", "csrc", "build")


def snapshot(dest):
    os.makedirs(dest, exist_ok=True)
    for obj in sorted(glob.glob(os.path.join(BUILD, "*.o"))):
        out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True, check=True).stdout
        with open(os.path.join(dest, os.path.basename(obj)[:-2] + ".sass"), "w") as f:
            f.write(out)


INSTR = re.compile(r"^\s*/\*[0-9a-f]{4,}\*/\s+(.*?)\s*/\*")


def functions(path):
    """{mangled name: (sha1 of the instruction text, instruction count)}"""
    table, name, h, n = {}, None, None, 0
    for line in open(path):
        if "Function :" in line:
            if name is not None:
                table[name] = (h.hexdigest(), n)
            name, h, n = line.split("Function :")[1].strip(), hashlib.sha1(), 0
            continue
        m = INSTR.match(line)
        if m and name is not None:
            h.update(m.group(1).encode())
            n += 1
    if name is not None:
        table[name] = (h.hexdigest(), n)
    return table


def load(directory):
    table = {}
    for path in sorted(glob.glob(os.path.join(directory, "*.sass"))):
        table.update(functions(path))
    return table


def compare(before, after):
    a, b = load(before), load(after)
    changed = sorted(k for k in a if k in b and a[k] != b[k])
    removed = sorted(k for k in a if k not in b)
    added = sorted(k for k in b if k not in a)
    print(f"{len(a)} functions before, {len(b)} after: {len(changed)} changed, {len(removed)} removed, {len(added)} added")
    for title, names in (("changed", changed), ("removed", removed), ("added", added)):
        for k in names[:40]:
            demangled = subprocess.run(["c++filt", k], capture_output=True, text=True).stdout.strip()[:150]
            extra = f"  {a[k][1]} -> {b[k][1]} instructions" if title == "changed" else ""
            print(f"  {title}: {demangled}{extra}")
        if len(names) > 40:
            print(f"  ... and {len(names) - 40} more {title}")
    return 1 if (changed or removed) else 0


if __name__ == "__main__":
    if len(sys.argv) >= 3 and sys.argv[1] == "snapshot":
        snapshot(sys.argv[2])
    elif len(sys.argv) >= 4 and sys.argv[1] == "compare":
        sys.exit(compare(sys.argv[2], sys.argv[3]))
    else:
        sys.exit(__doc__)
