"""CPU suite, part 2: the C-ABI library loads without a GPU, exports every symbol include/catnet_b200.h declares,
fails loudly (no fallback) when asked to compute without a device, and its host-side logic is right."""
import itertools
import math
import os
import re

import numpy as np
import pytest

from catnet_b200 import _abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "catnet_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(cnb_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = _abi.lib()
    names = declared_symbols()
    assert len(names) >= 30
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    assert sorted(names) == sorted(_abi.SYMBOLS), set(names) ^ set(_abi.SYMBOLS)


def test_no_silent_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(_abi.CnbError) as e:
        _abi.Context(0)
    assert e.value.code == _abi.CNB_ERR_CUDA
    s = np.ones((4, 3), dtype=np.int32)
    with pytest.raises(_abi.CnbError) as e:
        _abi.search_order(s, max_parent_set=1, max_complexity=10)
    assert e.value.code == _abi.CNB_ERR_CUDA


def test_product_does_not_touch_the_oracle():
    """the product path must never import, link or call anything under oracle/"""
    pat = re.compile(r"from\s+oracle|import\s+oracle|libcatnet_oracle|libcatnet_ref|catnet_oracle\.h|\borc_\w+\(|\bref_\w+\(")
    for dirpath, _, files in os.walk(os.path.join(ROOT, "catnet_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", "Makefile")):
                text = open(os.path.join(dirpath, f), errors="replace").read()
                assert not pat.search(text), (dirpath, f, pat.search(text).group(0))
    out = os.popen("ldd %s" % _abi.LIB_PATH).read()
    assert "catnet_oracle" not in out and "catnet_ref" not in out


def test_unrank_is_lexicographic():
    for n, k in [(1, 1), (5, 2), (6, 3), (7, 1), (9, 4), (12, 5), (6, 6), (8, 0)]:
        combos = list(itertools.combinations(range(n), k))
        for r, c in enumerate(combos):
            assert tuple(_abi.unrank_combination(n, k, r)) == c
    with pytest.raises(_abi.CnbError):
        _abi.unrank_combination(5, 2, 10)


def test_num_families_matches_formula():
    L = _abi.lib()
    for n, p in [(37, 2), (100, 3), (500, 2), (5, 4), (1, 2)]:
        keep = _abi.Keep()
        prm = _abi.make_params(keep, n, 10, max_parent_set=p, max_complexity=1 << 20)
        want = sum(math.comb(i, d) for i in range(n) for d in range(1, p + 1))
        assert L.cnb_num_families(prm) == want
    assert sum(math.comb(500, d + 1) for d in (1, 2)) == 20833250


def test_num_families_with_pools_and_fixed():
    L = _abi.lib()
    n = 6
    keep = _abi.Keep()
    pool = [[], [1], [1, 2], [1, 2, 3], [2, 3], [1, 2, 3, 4, 5]]
    fixed = [[], [], [1], [], [], [2]]
    prm = _abi.make_params(keep, n, 10, max_parent_set=2, max_complexity=1 << 20, parents_pool=pool, fixed_parents=fixed,
                           parent_sizes=[2, 2, 2, 1, 2, 2], num_cats=[2] * n)
    # node2: pool{1}: d=1:1 | node3: fixed{1} free{2}: d=2:1 | node4: free{1,2,3}, size<=1: 3 | node5: free{2,3}: 2+1
    # node6: fixed{2} free{1,3,4,5}: d=2: 4
    assert L.cnb_num_families(prm) == 1 + 1 + 3 + 3 + 4


def test_bad_order_is_rejected():
    L = _abi.lib()
    keep = _abi.Keep()
    prm = _abi.make_params(keep, 4, 10, max_parent_set=1, max_complexity=100, order=[1, 2, 2, 4])
    assert L.cnb_num_families(prm) == -1
    assert b"Invalid nodeOrder" in L.cnb_last_error()


def test_host_log_is_libm_log():
    rng = np.random.default_rng(0)
    x = np.concatenate([rng.random(20000), 0.9 + 0.2 * rng.random(20000), np.exp(rng.uniform(-700, 700, 20000)),
                        np.array([k * (1.0 / n) for n in range(1, 200) for k in range(1, n + 1)]),
                        np.array([5e-324, 1e-310, 2.2250738585072014e-308, 1.0, 1.7e308])])
    ours = _abi.host_log(x)
    libm = np.array([math.log(v) for v in x])
    assert np.array_equal(ours.view(np.int64), libm.view(np.int64))
