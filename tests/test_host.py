"""CPU suite, part 5: host-side logic above the C ABI -- the Python mirror of the R front-ends and the fixtures."""
import math

import numpy as np
import pytest

from catnet_b200 import _abi, search, synth


def test_default_max_complexity_truncates_like_r():
    # R: as.integer(numnodes * exp(log(maxCategories)*maxParentSet) * (maxCategories-1)); SURVEY.md section 3.1
    assert synth.default_max_complexity(37, 4, 2) == 1775
    assert synth.default_max_complexity(100, 2, 3) == 799
    assert synth.default_max_complexity(100, 3, 3) == 5400
    assert synth.default_max_complexity(500, 4, 2) == 23999


def test_alarm_fixture_and_true_order():
    net = synth.load_alarm()
    assert len(net["cats"]) == 37 and sum((net["cats"][i] - 1) * int(np.prod([net["cats"][p] for p in net["parents"][i]]))
                                         for i in range(37)) == 509
    order = [v + 1 for v in synth.order_nodes_descend(net["parents"])]
    assert order == [4, 6, 7, 8, 11, 13, 14, 15, 17, 19, 23, 25, 27, 28, 29, 30, 31, 32, 33, 1, 5, 16, 18, 20, 22, 24,
                     26, 2, 3, 21, 34, 35, 36, 37, 9, 10, 12]
    for c in net["cpts"]:
        assert np.allclose(c.sum(axis=1), 1.0)
    s = synth.sample_network(net, 500, np.random.Generator(np.random.PCG64(1)))
    assert s.shape == (500, 37) and s.min() >= 1 and (s.max(axis=0) <= net["cats"]).all()


def test_breast_fixture():
    s = synth.load_breast()
    assert s.shape == (98, 100) and set(np.unique(s)) == {1, 2}


def test_random_network_is_a_dag_with_valid_cpts():
    rng = np.random.Generator(np.random.PCG64(3))
    net = synth.random_network(40, 3, 3, rng)
    order = synth.order_nodes_descend(net["parents"])
    assert sorted(order) == list(range(40))
    pos = {v: i for i, v in enumerate(order)}
    assert all(pos[p] < pos[i] for i in range(40) for p in net["parents"][i])
    assert all(len(p) <= 3 for p in net["parents"])
    for c in net["cpts"]:
        assert np.allclose(c.sum(axis=1), 1.0) and (c >= 0).all()


def test_discretize_uniform():
    x = np.array([[0.0, 1.0, 2.0, 3.0, 4.0], [5.0, 5.0, 5.0, 5.0, 6.0]])
    d = synth.discretize_uniform(x, 2)
    assert d.tolist() == [[1, 1, 2, 2, 2], [1, 1, 1, 1, 2]]


def test_categorize_sample():
    data = [[3.0, 1.0, float("nan"), 2.0], [10.0, 20.0, 10.0, 20.0]]
    out, pert, cats, mc = search.categorize_sample(data)
    assert out.tolist() == [[3, 1, _abi.NA_INTEGER, 2], [1, 2, 1, 2]]
    assert cats == [[1.0, 2.0, 3.0], [10.0, 20.0]] and mc == 3
    with pytest.raises(ValueError):
        search.categorize_sample([[1.0, 5.0]], node_cats=[[1.0, 2.0]])


def test_cnsearchorder_marshalling(monkeypatch):
    seen = {}

    def fake(samples, **kw):
        seen.update(kw, samples=samples)
        return []
    monkeypatch.setattr(_abi, "search_order", fake)
    data = np.array([[1, 2, 1, 2, 3], [1, 1, 2, 2, 1], [2, 1, 2, 1, 1]], dtype=float)
    e = np.array([[0, 1.0, 0.3], [0.2, 0, 1.7], [np.nan, -1, 0]])
    ev = search.cnSearchOrder(data, maxParentSet=0, parentSizes=[1, 5, -2], nodeOrder=["c", "a", "b"], edgeProb=e,
                              nodeNames=["a", "b", "c"], fixedParents=[["b", "c"], [], []])
    assert ev.nets == []
    assert seen["samples"].shape == (5, 3)                       # [sample][node], as R's column-major N x M matrix
    assert seen["max_parent_set"] == 5 and seen["order"] == [3, 1, 2]
    assert list(seen["parent_sizes"]) == [1, 5, 0]
    assert seen["num_cats"] is None                               # nodeCats not given -> the library infers them
    assert seen["max_complexity"] == int(3 * math.exp(math.log(3) * 5) * 2)
    ep = seen["edge_prob"]
    assert ep[1, 2] == 1.0 and ep[2, 0] == 0.5 and ep[2, 1] == 0.0
    # edgeProb == 1 entries are appended to the fixed parents, duplicates included, exactly as R's c() does (:127-133)
    assert seen["fixed_parents"][0] == [2, 3, 2] and seen["fixed_parents"][1] == [3]


def test_cnsearchsa_argument_clamps(monkeypatch):
    seen = {}

    def fake(samples, sa, **kw):
        seen.update(kw, sa=sa)
        return {"nets": [], "order": np.arange(1, 4)}
    monkeypatch.setattr(_abi, "search_sa", fake)
    data = np.array([[1, 2, 1, 2], [1, 1, 2, 2], [2, 1, 2, 1]], dtype=float)
    with pytest.warns(UserWarning):
        search.cnSearchSA(data, maxParentSet=2, tempStart=-1, tempCoolFact=3, tempCheckOrders=0, maxIter=0,
                          selectMode="AIC", seed=5, dirProb=np.array([[0, 2.0, 1], [1, 0, 1], [1, 1, 0.0]]))
    sa = seen["sa"]
    assert sa["temp_start"] == 0 and sa["temp_cool_fact"] == 1 and sa["temp_check_orders"] == 1 and sa["max_iter"] == 1
    assert sa["select_mode"] == 1 and sorted(sa["start_order"].tolist()) == [1, 2, 3]
    d = sa["dir_prob"]
    assert abs(d[0, 1] - 2.0 / 3.0) < 1e-15 and d[0, 0] == 0.5


@pytest.mark.parametrize("n", [3, 4, 9, 29, 63, 64, 65, 66, 127, 128, 129, 200, 500])
def test_tensor_tiling_is_a_bijection(n):
    """cnb_tiles.h: every two-parent family of the unrestricted group owns exactly one slot of one tile (F tiles:
    pairs (a,b) x children; D tiles: pairs (b,c) x first parents), and the slot decodes back to the family"""
    from catnet_b200 import _abi
    ntiles = _abi.lib().cnb_tile_selftest(n)
    assert ntiles > 0, ntiles
    fam = n * (n - 1) * (n - 2) // 6
    assert ntiles * 28 * 64 >= fam
    if n == 500:
        assert ntiles < 12500          # 15,762 before the diagonal tiles were transposed


@pytest.mark.parametrize("n", [3, 4, 9, 29, 47, 48, 49, 50, 95, 96, 97, 145, 200, 500])
def test_full_table_tiling_is_a_bijection(n):
    """the same for the full-table geometry (16 pairs x 48 column nodes per tile: missing values / perturbation masks)"""
    from catnet_b200 import _abi
    ntiles = _abi.lib().cnb_tile_selftest_full(n)
    assert ntiles > 0, ntiles
    assert ntiles * 16 * 48 >= n * (n - 1) * (n - 2) // 6
    if n == 500:
        assert ntiles < 29500          # 27,127 slots-worth + diagonal overhead; 2.33 x the 12,279 inclusion-exclusion tiles


@pytest.mark.parametrize("fr,tp,tc", [(3, 28, 64), (2, 64, 96), (1, 256, 192)])
@pytest.mark.parametrize("n", [3, 4, 9, 65, 96, 97, 100, 192, 193, 200, 400])
def test_reduced_tiling_is_a_bijection(n, fr, tp, tc):
    """data with fewer than four categories per node: a pair takes fr^2 rows and a column node fr columns, so a tile
    holds 64 x 96 (three categories) or 256 x 192 (binary) family slots; same bijection"""
    from catnet_b200 import _abi
    ntiles = _abi.lib().cnb_tile_selftest_geo(n, fr)
    assert ntiles > 0, ntiles
    assert ntiles * tp * tc >= n * (n - 1) * (n - 2) // 6
    if fr == 3:
        assert ntiles == _abi.lib().cnb_tile_selftest(n)


@pytest.mark.parametrize("n,f1", [(4, 1), (5, 2), (30, 3), (64, 2), (65, 2), (70, 3), (100, 2), (129, 1), (97, 2), (100, -2), (50, -1), (193, 1)])
def test_three_parent_tiling_is_consistent(n, f1):
    """cnb_tiles.h, T tiles: every three-parent family (a < b < c | x) and free value i of a owns one slot in the column
    tile of x, and the table kernel's decode of that tile puts (virtual node of (a, b, i), c) there"""
    from catnet_b200 import _abi
    ntiles = _abi.lib().cnb_tile3_selftest(n, f1)
    assert ntiles > 0, ntiles
    if (n, f1) == (100, -2):
        assert ntiles == 11672             # C3 in 28 x 64 tiles: column tiles [0, 36) and [36, 100)
    if (n, f1) == (100, 2):
        assert ntiles == 4903              # 64 x 96 tiles: column tiles [0, 4) and [4, 100), 2.4 x fewer MMAs than 11,672


@pytest.mark.parametrize("seed", range(6))
def test_entropy_order_host_tail(seed):
    """cnb_entropy_order's host half (ranks from the entropy matrix, the reference's _order with its holes at tied ranks)
    against the oracle's order for the same data, fed with the oracle's entropy matrix"""
    from oracle import port
    rng = np.random.default_rng(seed)
    m, n = int(rng.integers(5, 200)), int(rng.integers(2, 9))
    s = rng.integers(1, 4, size=(m, n)).astype(np.int32)
    if seed % 2 and n > 2:
        s[:, 2] = s[:, 1]                                   # tied ranks
    p = (rng.random((m, n)) < 0.3).astype(np.int32) if seed % 3 == 0 else None
    ent = port.pairwise("entropy", s, p)                    # [n1][n2] = mat[n2 * N + n1]
    mat = np.ascontiguousarray(ent.T)
    out = np.zeros(n, dtype=np.int32)
    _abi.lib().cnb_host_entropy_order(_abi.ptr(mat), n, _abi.ptr(out))
    assert out.tolist() == port.pairwise("order", s, p).tolist()
