#!/usr/bin/env python
"""Which kernels of two builds of libmoka_b200.so have byte-identical SASS?  Used when device code is edited without a GPU
at hand: a kernel whose SASS did not change keeps its measured behaviour.

  python tools/sass_diff.py <old libmoka_b200.so> [<new libmoka_b200.so>]
e.g.  git archive <rev> mpas-ocean.jl_b200/csrc mpas-ocean.jl_b200/Makefile include | tar -x -C /tmp/old && make -C /tmp/old/mpas-ocean.jl_b200"""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def sass(lib):
    out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    funcs, cur = {}, None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            funcs[cur] = []
        elif cur and re.match(r"\s+/\*[0-9a-f]{4}\*/", line):
            funcs[cur].append(re.sub(r"/\*[0-9a-f]+\*/", "", line).strip())
    return funcs


def norm(name):
    # k_rk_stage gained trailing template parameters (bool PUSH = false, int TMA = 0 for every pre-existing instantiation)
    if "k_rk_stage" in name and "Lb0ELi0EEEvNS0_9StageArgs" in name:
        return name.replace("Lb0ELi0EEEvNS0_9StageArgs", "EEvNS0_9StageArgs")
    # k_fe_step gained a trailing bool LIST = false
    m = re.match(r"(_ZN5mokab5fused9k_fe_stepILi\d+ELi\d+ELb[01]E)Lb0E(EEvNS0_6FeArgsE)$", name)
    if m:
        return m.group(1) + m.group(2)
    return name


def main():
    old = sass(sys.argv[1])
    new = {norm(k): v for k, v in sass(sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, "mpas-ocean.jl_b200", "libmoka_b200.so")).items()}
    same = [k for k in old if k in new and old[k] == new[k]]
    diff = [k for k in old if k in new and old[k] != new[k]]
    gone = [k for k in old if k not in new]
    added = [k for k in new if k not in old]
    print(f"{len(old)} kernels in the old build: {len(same)} byte-identical, {len(diff)} different, {len(gone)} gone; {len(added)} new")
    for tag, names in (("different", diff), ("gone", gone), ("new", added)):
        for k in names:
            print(f"  {tag}: {k}")
    sys.exit(1 if diff or gone else 0)


if __name__ == "__main__":
    try:
        main()
    except BrokenPipeError:          # piped into head
        sys.exit(0)
