"""Build tests/sim/_build/libmoka_b200_sim.so: the library's own sources compiled for the HOST against the simulation
shim (tests/sim/include/cuda_runtime.h + sim_runtime.cpp).  TEST INFRASTRUCTURE -- see the shim's header.

The only source transformation is the kernel-launch syntax, which g++ cannot parse:
    K<<<grid, block, smem, stream>>>(args...)   ->   ::mokab_sim::launch_impl(coop, "K", K, grid, block, smem, stream, args...)
Everything else (host logic, kernels, templates) is compiled as written, with MOKAB_SIM defined
(csrc/common.cuh swaps its inline-PTX streaming loads for plain loads under that macro)."""
from __future__ import annotations

import hashlib
import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
# MOKAB_SIM_CSRC / MOKAB_SIM_BUILD: build another revision of the sources (e.g. `git show <rev>:...` extracted somewhere)
# into another directory -- how a suspected ordering bug is confirmed against the code that had it
CSRC = os.environ.get("MOKAB_SIM_CSRC", os.path.join(ROOT, "mpas-ocean.jl_b200", "csrc"))
BUILD = os.environ.get("MOKAB_SIM_BUILD", os.path.join(HERE, "_build"))
LIB = os.path.join(BUILD, "libmoka_b200_sim.so")

# Kernels whose threads cooperate (__syncthreads / warp shuffles) run their blocks as fibers; they are recognised by
# name at launch time (mokab_sim::is_coop_name in the shim header: the reduce:: kernels).


def _match_back(s: str, i: int, open_c: str, close_c: str) -> int:
    """s[i] == close_c; index of the matching open_c."""
    depth = 0
    while i >= 0:
        if s[i] == close_c:
            depth += 1
        elif s[i] == open_c:
            depth -= 1
            if depth == 0:
                return i
        i -= 1
    raise ValueError("unbalanced")


def _match_fwd(s: str, i: int) -> int:
    """s[i] == '('; index of the matching ')'."""
    depth = 0
    while i < len(s):
        if s[i] == "(":
            depth += 1
        elif s[i] == ")":
            depth -= 1
            if depth == 0:
                return i
        i += 1
    raise ValueError("unbalanced")


def _split_top(s: str):
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch in "([{":
            depth += 1
        elif ch in ")]}":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur.strip())
            cur = ""
        else:
            cur += ch
    out.append(cur.strip())
    return out


def rewrite_launches(src: str) -> tuple[str, int]:
    n = 0
    while True:
        k = src.find("<<<")
        if k < 0:
            return src, n
        # kernel expression to the left
        j = k - 1
        while src[j].isspace():
            j -= 1
        if src[j] == ")":
            start = _match_back(src, j, "(", ")")
        else:
            if src[j] == ">":
                j = _match_back(src, j, "<", ">") - 1
            while j >= 0 and (src[j].isalnum() or src[j] in "_:"):
                j -= 1
            start = j + 1
        kernel = src[start:k].strip()
        end_cfg = src.index(">>>", k)
        cfg = _split_top(src[k + 3:end_cfg].replace("\\\n", " "))
        assert len(cfg) in (2, 3, 4), cfg
        grid, block = cfg[0], cfg[1]
        smem = cfg[2] if len(cfg) >= 3 else "0"
        stream = cfg[3] if len(cfg) == 4 else "nullptr"
        a0 = end_cfg + 3
        while src[a0].isspace():
            a0 += 1
        assert src[a0] == "(", src[a0:a0 + 40]
        a1 = _match_fwd(src, a0)
        args = src[a0 + 1:a1]
        name = "#" + kernel if re.fullmatch(r"[A-Za-z_]\w*", kernel) and kernel == "kernel" else '"' + kernel.replace('"', "'") + '"'
        coop = "::mokab_sim::is_coop_name(%s)" % name
        repl = f"::mokab_sim::launch_impl({coop}, {name}, {kernel}, (unsigned)({grid}), (unsigned)({block}), (size_t)({smem}), ({stream}), {args})"
        src = src[:start] + repl + src[a1 + 1:]
        n += 1


def _sources():
    files = sorted(f for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh")))
    extra = [os.path.join(HERE, "sim_runtime.cpp"), os.path.join(HERE, "include", "cuda_runtime.h"), os.path.abspath(__file__),
             os.path.join(ROOT, "include", "moka_b200.h")]
    return [os.path.join(CSRC, f) for f in files], extra


def build(force: bool = False, verbose: bool = False) -> str:
    srcs, extra = _sources()
    h = hashlib.sha256()
    for p in srcs + extra:
        h.update(open(p, "rb").read())
    h.update(os.environ.get("MOKAB_SIM_ASAN", "").encode())
    extra_flags = os.environ.get("MOKAB_SIM_CFLAGS", "").split()          # e.g. -DMOKAB_BLOCK_CELLS=128: a tuning variant of the kernels
    h.update(" ".join(extra_flags).encode())
    stamp = os.path.join(BUILD, "stamp")
    if not force and os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read() == h.hexdigest():
        return LIB
    gen = os.path.join(BUILD, "mpas-ocean.jl_b200", "csrc")      # same depth as the real tree: "../../include/moka_b200.h"
    os.makedirs(gen, exist_ok=True)
    os.makedirs(os.path.join(BUILD, "include"), exist_ok=True)
    with open(os.path.join(BUILD, "include", "moka_b200.h"), "w") as f:
        f.write(open(os.path.join(ROOT, "include", "moka_b200.h")).read())
    total = 0
    for p in srcs:
        text, n = rewrite_launches(open(p).read())
        total += n
        out = os.path.join(gen, os.path.basename(p).replace(".cu", ".cpp") if p.endswith(".cu") else os.path.basename(p))
        with open(out, "w") as f:
            f.write(text)
    # MOKAB_SIM_ASAN=1: AddressSanitizer build -- every out-of-bounds access of a kernel or of the host code is reported (run with
    # LD_PRELOAD=$(gcc -print-file-name=libasan.so) ASAN_OPTIONS=detect_leaks=0:detect_stack_use_after_return=0)
    asan = ["-fsanitize=address", "-fno-omit-frame-pointer"] if os.environ.get("MOKAB_SIM_ASAN") == "1" else []
    if os.environ.get("MOKAB_SIM_ASAN") == "undefined":       # UndefinedBehaviorSanitizer instead (shifts, signed overflow, misaligned access)
        asan = ["-fsanitize=undefined", "-fno-sanitize-recover=undefined", "-fno-omit-frame-pointer"]
    cmd = ["g++", "-std=c++17", "-O1", "-g", "-fPIC", "-shared", *asan, *extra_flags, "-fopenmp", "-ffp-contract=off", "-fno-fast-math", "-DMOKAB_SIM", "-U_FORTIFY_SOURCE",
           "-Wno-unknown-pragmas", "-I", os.path.join(HERE, "include"), os.path.join(gen, "moka_b200.cpp"),
           os.path.join(HERE, "sim_runtime.cpp"), "-o", LIB,
           "-Wl,-Bsymbolic",      # the cuda* symbols defined here must win over a real libcudart that torch may have loaded
           "-lgomp", "-lpthread"]
    if verbose:
        print(f"[sim] {total} kernel launches rewritten; {' '.join(cmd)}", file=sys.stderr)
    subprocess.check_call(cmd)
    with open(stamp, "w") as f:
        f.write(h.hexdigest())
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
