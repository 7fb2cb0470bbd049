"""Registers / spills of the library's kernels from a fresh `make` (`-Xptxas -v`), one demangled kernel per line; with an older
build of the library as argument, only the kernels that build does not hold (what was added since).

  python tools/ptxas_report.py [old/libmoka_b200.so] > profiles/rNN_ptxas_new_kernels.txt"""
import os
import re
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from sass_diff import norm  # noqa: E402  (template parameters added since: same kernel, longer name)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "mpas-ocean.jl_b200")


def kernels_of(lib):
    out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    return set(re.findall(r"Function : (\S+)", out))


def main():
    old = kernels_of(sys.argv[1]) if len(sys.argv) > 1 else set()
    os.utime(os.path.join(PKG, "csrc", "moka_b200.cu"))
    log = subprocess.run(["make", "-C", PKG], capture_output=True, text=True)
    text = log.stdout + log.stderr
    rows = {}
    cur = None
    for line in text.splitlines():
        m = re.search(r"Compiling entry function '(\S+)'", line)
        if m:
            cur = m.group(1)
            rows[cur] = []
        elif cur and ("spill" in line or "Used" in line):
            rows[cur].append(re.sub(r"^ptxas info\s*:\s*", "", line.strip()))
    names = [k for k in rows if norm(k) not in old]
    dem = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    for pretty, k in sorted(zip(dem, names)):
        pretty = re.sub(r"^void ", "", pretty)
        pretty = re.sub(r"\(.*$", "", pretty)
        print(f"{pretty}\t{' '.join(rows[k])}")


if __name__ == "__main__":
    main()
