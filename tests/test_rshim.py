"""CPU suite, part 4: the R-facing shim (catnet_b200/csrc/rshim.cpp) cannot be built here (no R in the image), so it
is syntax-checked against declaration-only stand-ins of R's headers, and its .Call table is compared with the
reference's registration (names and arity, src/ccnInit.c:29-45)."""
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIM = os.path.join(ROOT, "catnet_b200", "csrc", "rshim.cpp")


def test_rshim_syntax():
    r = subprocess.run(["g++", "-std=c++14", "-fsyntax-only", "-Wall", "-I" + os.path.join(ROOT, "tests", "rstub"),
                        "-DCNB_STANDALONE_RSHIM", SHIM], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_call_table_matches_reference_registration():
    src = open(SHIM).read()
    table = dict((m.group(1), int(m.group(2))) for m in re.finditer(r'\{"(ccn\w+)",\s*\(DL_FUNC\)&\w+,\s*(\d+)\}', src))
    # names and arities the reference registers for this path (src/ccnInit.c:30-32, 36-39, 41, 43)
    assert table == {"ccnOptimalNetsForOrder": 12, "ccnOptimalNetsSA": 23, "ccnParHistogram": 14, "ccnReleaseCache": 0,
                     "ccnSetSeed": 1, "ccnEntropyPairwise": 2, "ccnEntropyOrder": 2, "ccnKLPairwise": 2,
                     "ccnPearsonPairwise": 2, "ccnLoglik": 4, "ccnNodeLoglik": 4, "ccnSetProb": 3, "ccnSamples": 4}
    for name, arity in table.items():
        sig = re.search(r"SEXP %s\(([^)]*)\)" % name, src).group(1)
        nargs = 0 if sig.strip() == "void" else sig.count("SEXP")
        assert nargs == arity, (name, nargs, arity)
