#!/usr/bin/env python
"""Fuzz the CUDA path (through the C ABI) against the plain-C restatement on the seeded generators that
tools/fuzz_oracle.py pins against the compiled reference.  Needs a B200; reads nothing under /root/reference.

    python tools/fuzz_gpu.py search 1000 3000     # tests/helpers.py::constrained_problem
    python tools/fuzz_gpu.py wide 0 3000          # tools/fuzz_oracle.py::wide_problem
    python tools/fuzz_gpu.py pert 1100 1400       # tests/helpers.py::fully_perturbed_problem

Written at the end of round 1 (no GPU budget left): not run yet."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tools")]
from helpers import compare_search, constrained_problem, fully_perturbed_problem  # noqa: E402


def main():
    from catnet_b200 import _abi
    from fuzz_oracle import wide_problem
    from oracle import port
    gen = {"search": constrained_problem, "wide": wide_problem, "pert": fully_perturbed_problem}[sys.argv[1]]
    ran, bad = 0, []
    for seed in range(int(sys.argv[2]), int(sys.argv[3])):
        try:
            s, P, K, kw = gen(seed)
        except ValueError:                       # fewer samples than categories
            continue
        theirs = port.search_order(s, P, K, **kw)
        try:
            ours = _abi.search_order(s, max_parent_set=P, max_complexity=K, **kw)
        except _abi.CnbError as e:               # the library refuses what the restatement answers with nothing
            ours = []
            if theirs:
                bad.append((seed, "error %s" % e))
                continue
        ran += 1
        try:
            assert len(ours) == len(theirs), (len(ours), len(theirs))
            assert compare_search(ours, theirs) == len(theirs)
        except AssertionError as e:
            bad.append((seed, str(e)[:200]))
    print("problems %d  mismatches %s" % (ran, bad[:20]))
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
