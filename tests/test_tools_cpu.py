"""CPU: the host-side logic of tools/trace_stages.py (grouping the per-block records of the TRACE build into launches)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
import trace_stages as T  # noqa: E402


def test_block_records_are_grouped_into_launches_by_kind_and_grid():
    # two launches of the stage-2 interior kernel (3 blocks each), one boundary launch (2 blocks), a push and a wait kernel
    rec = np.zeros(11, T.REC)
    rec["kind"] = [18, 18, 18, 33, 33, 100, 101, 18, 18, 18, 103]
    rec["grid"] = [3, 3, 3, 2, 2, 1, 1, 3, 3, 3, 1]
    rec["t0"] = [0, 1, 2, 0, 1, 10, 12, 20, 21, 22, 30]
    rec["t1"] = [5, 6, 7, 4, 5, 11, 15, 25, 26, 27, 33]
    rec["t2"][6] = 14
    rng = np.random.default_rng(0)
    la = T.launches(rec[rng.permutation(rec.size)])              # the ring is not ordered
    assert [(x["kind"], x["grid"], x["t0"], x["t1"], x["complete"]) for x in la] == [
        (18, 3, 0, 7, True), (33, 2, 0, 5, True), (100, 1, 10, 11, True), (101, 1, 12, 15, True), (18, 3, 20, 27, True), (103, 1, 30, 33, True)]
    assert la[3]["gate"] == 14 and la[0]["gate"] == 0
    assert T.label(18) == "stage2 interior" and T.label(33) == "stage1 boundary" and T.label(52) == "stage4 boundary+push"
    assert T.label(100) == "halo push" and T.label(103).startswith("wait + unpack")


def test_an_incomplete_launch_is_flagged():
    rec = np.zeros(2, T.REC)
    rec["kind"], rec["grid"], rec["t0"], rec["t1"] = 17, 3, [0, 1], [2, 3]      # two of three blocks made it into the ring
    la = T.launches(rec)
    assert len(la) == 1 and not la[0]["complete"]
