"""GPU parity tests proper: the CUDA path, through the C ABI, against
  (1) the committed golden fixtures = outputs of the reference's own C++ (tools/make_golden.py),
  (2) the oracle restatement (oracle/catnet_oracle.c) on further seeded inputs,
  (3) the reference core itself (oracle/_ref) when its library travelled to this box.

Bars (BASELINE.json north_star): family counts bit-exact; family log-lik within 1e-9 relative in fp64 -- the device
evaluates glibc's own log algorithm, so these tests require BIT equality; cnSearchOrder selects identical parent
sets for every complexity level.  Nothing here reads /root/reference.
"""
import itertools
import math

import numpy as np
import pytest

from cases import (CASES, SA_CASES, build_case, ref_args, HIST_CASES, build_hist_case, PAIRWISE_CASES, build_pairwise_case,
                   NETWORK_CASES, build_network_case, SAMPLES_CASES, build_samples_case)
from helpers import (NA, check_network_case, compare_golden, compare_search, constrained_problem, load_golden,
                     random_problem, unhex)

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def abi():
    from catnet_b200 import _abi
    _abi.lib()
    return _abi


@pytest.fixture(scope="module")
def port():
    from oracle import port as p
    p.lib()
    return p


def test_device_log_is_glibc_log(abi):
    rng = np.random.default_rng(1)
    xs = [rng.random(200000), rng.random(200000) * 1e-6, 0.9 + 0.2 * rng.random(200000),
          np.exp(rng.uniform(-700, 700, 100000)), np.array([1.0, 0.5, 2.0, 5e-324, 1e-310, 1.7e308])]
    ratios = np.array([k / n for n in range(1, 400) for k in range(1, n + 1)])
    x = np.concatenate(xs + [ratios, np.array([k * (1.0 / n) for n in range(1, 400) for k in range(1, n + 1)])])
    dev = abi.device_log(x)
    host = abi.host_log(x)
    assert np.array_equal(dev.view(np.int64), host.view(np.int64))
    libm = np.array([math.log(v) for v in x[:50000]])
    assert np.array_equal(dev[:50000].view(np.int64), libm.view(np.int64))


@pytest.mark.parametrize("cats,na,pert", [(None, 0.0, False), (None, 0.08, False), (3, 0.05, True), (2, 0.0, False),
                                          (4, 0.1, True), ([2, 3, 4, 2, 3, 4, 2, 3, 4, 2], 0.03, False)])
@pytest.mark.parametrize("m", [1, 31, 32, 33, 257, 1000])
def test_family_counts_and_loglik(abi, port, cats, na, pert, m):
    """empty / ragged sample counts, NA, perturbation masks, 0..3 parents in any order, both kernels"""
    n = 10
    s, c = random_problem(7 + m, n, max(m, 5), cats, na)
    s = s[:m]
    rng = np.random.default_rng(m)
    p = (rng.random(s.shape) < 0.2).astype(np.int32) if pert else None
    s0 = np.where(s < 1, 99, s - 1).astype(np.int32)
    with abi.Context(0) as ctx:
        ctx.set_samples(s, perturbations=p, num_cats=c)
        for d in (0, 1, 2, 3):
            fams = [(ch, pa) for ch in range(n) for pa in itertools.permutations([v for v in range(n) if v != ch], d)]
            fams = [fams[i] for i in rng.choice(len(fams), size=min(60, len(fams)), replace=False)]
            children = [f[0] for f in fams]
            parents = np.array([f[1] for f in fams], dtype=np.int32).reshape(len(fams), d)
            for generic in (False, True):
                counts, ll, nc = ctx.family_scores(children, parents, force_generic=generic)
                for i, (ch, pa) in enumerate(fams):
                    keep = None if p is None else (p[:, ch] == 0)
                    rc, rll, rnc = port.family_score(s0, c, ch, list(pa), keep=keep)
                    assert np.array_equal(counts[i, :len(rc)], rc.astype(np.int64)), (d, generic, ch, pa)
                    assert nc[i] == rnc
                    assert ll[i] == rll, (d, generic, ch, pa, ll[i], rll)   # bit-exact


@pytest.mark.parametrize("bits", [0, 32768, 65536], ids=["warp_atomics", "cta_per_family", "warp_match_any"])
@pytest.mark.parametrize("cats,na,pert,m", [(6, 0.0, False, 1000), (8, 0.05, True, 777), ([5, 12, 3, 7, 2, 9, 6, 4, 10, 5], 0.03, False, 2000),
                                            (16, 0.0, False, 300), (4, 0.1, True, 33), (5, 0.0, False, 1)])
def test_many_categories_histogram_scorers(abi, port, bits, cats, na, pert, m):
    """more than four categories per node: the histogram scorers -- a warp per family with its own shared-memory table
    (atomics, or __match_any_sync aggregation) and the one-CTA-per-family kernel -- against the oracle: counts exact,
    log-likelihoods bit-identical, 0..3 parents, missing values, perturbation masks, whole searches"""
    n = 10
    s, c = random_problem(11 + m, n, max(m, 20), cats, na)
    s = s[:m]
    rng = np.random.default_rng(m)
    p = (rng.random(s.shape) < 0.2).astype(np.int32) if pert else None
    s0 = np.where(s < 1, 99, s - 1).astype(np.int32)
    with abi.Context(0) as ctx:
        ctx.set_samples(s, perturbations=p, num_cats=c)
        ctx.force_generic(bits)
        for d in (0, 1, 2, 3):
            if int(np.max(c)) ** (d + 1) > 16384:
                continue
            fams = [(ch, pa) for ch in range(n) for pa in itertools.permutations([v for v in range(n) if v != ch], d)]
            fams = [fams[i] for i in rng.choice(len(fams), size=min(40, len(fams)), replace=False)]
            children = [f[0] for f in fams]
            parents = np.array([f[1] for f in fams], dtype=np.int32).reshape(len(fams), d)
            counts, ll, nc = ctx.family_scores(children, parents, force_generic=True)
            for i, (ch, pa) in enumerate(fams):
                keep = None if p is None else (p[:, ch] == 0)
                rc, rll, rnc = port.family_score(s0, c, ch, list(pa), keep=keep)
                assert np.array_equal(counts[i, :len(rc)], rc.astype(np.int64)), (d, ch, pa)
                assert nc[i] == rnc and ll[i] == rll, (d, ch, pa, ll[i], rll)
        if m >= 300 and int(np.max(c)) <= 12:
            K = int(np.sum(c - 1)) + 3 * int(np.max(c)) ** 2
            ours = ctx.search_order(max_parent_set=2, max_complexity=K)
            theirs = port.search_order(s, 2, K, num_cats=c, perturbations=p)
            assert compare_search(ours, theirs) == len(theirs)


def test_family_scores_match_golden(abi):
    by_problem = {}
    for g in load_golden("ref_family_scores"):
        by_problem.setdefault((g["seed"], str(g["cats"]), g["na"]), []).append(g)
    for (_, _, _), gs in by_problem.items():
        g0 = gs[0]
        s, c = random_problem(g0["seed"], 8, 150, g0["cats"], g0["na"])
        with abi.Context(0) as ctx:
            ctx.set_samples(s, num_cats=c)
            for g in gs:
                for generic in (False, True):
                    counts, ll, nc = ctx.family_scores([g["child"]], np.array([g["parents"]], dtype=np.int32).reshape(1, -1),
                                                       force_generic=generic)
                    assert [int(v) for v in counts[0, :len(g["counts"])]] == g["counts"]
                    assert ll[0] == float.fromhex(g["loglik"]) and nc[0] == g["ncount"]


@pytest.mark.parametrize("case", CASES, ids=lambda c: "seed%d" % c["seed"])
def test_search_order_matches_reference(abi, port, case):
    s, kw = build_case(case)
    golden = load_golden("ref_search_order")["seed%d" % case["seed"]]
    ours = abi.search_order(s, **kw)
    compare_golden(ours, golden)
    theirs = port.search_order(s, kw["max_parent_set"], kw["max_complexity"], **ref_args(kw))
    assert compare_search(ours, theirs) == len(theirs)        # also checks CPTs and sample sizes
    # same through the generic histogram kernel
    with abi.Context(0) as ctx:
        ctx.set_samples(s, perturbations=kw.get("perturbations"), num_cats=kw.get("num_cats"))
        ctx.force_generic(True)
        kw2 = {k: v for k, v in kw.items() if k not in ("perturbations", "num_cats")}
        compare_golden(ctx.search_order(**kw2), golden)
        # and through the direct bit-plane scorer (no inclusion-exclusion over the pair tables)
        ctx.force_generic(2)
        compare_golden(ctx.search_order(**kw2), golden)


@pytest.mark.parametrize("bits", [32, 64, 2048], ids=["per_node", "grid_barrier", "wave"])
@pytest.mark.parametrize("case", CASES[::3], ids=lambda c: "seed%d" % c["seed"])
def test_search_order_dp_variants(abi, case, bits):
    """the DP normally runs as one cooperative trapezoid-wavefront kernel (several nodes per neighbour exchange); the
    one-node-per-exchange wavefront, the grid-barrier kernel and the per-node launches must give the same networks"""
    s, kw = build_case(case)
    golden = load_golden("ref_search_order")["seed%d" % case["seed"]]
    with abi.Context(0) as ctx:
        ctx.set_samples(s, perturbations=kw.get("perturbations"), num_cats=kw.get("num_cats"))
        ctx.force_generic(bits)
        kw2 = {k: v for k, v in kw.items() if k not in ("perturbations", "num_cats")}
        compare_golden(ctx.search_order(**kw2), golden)


def _tensor_eligible(case, cats, bits):
    """missing values and perturbation masks run as full-table tiles (E2M1 operands only); under pools, fixed parents and
    parent-set sizes the tiles are used while the candidates are at least 1/16 of the unrestricted group (None = depends)"""
    masked_ok = not (bits & 16)
    ok = (int(np.max(cats)) <= 4 and (masked_ok or not any(case.get(k) for k in ("na", "pert"))) and case["n"] >= 3
          and case["P"] >= 2)
    if ok and any(case.get(k) for k in ("pools", "fixed", "psizes")):
        return None
    return ok


TENSOR_KINDS = [(4, 4), (4 | 16, 8)]          # force_generic bits -> cnb_timing.tensor_kind (E2M1 / u8 operands)
TENSOR_KINDS_FULL = TENSOR_KINDS + [(4 | 8192, 4)]   # + full-table tiles forced on complete data


@pytest.mark.parametrize("bits,kind", TENSOR_KINDS, ids=["mxf4", "i8"])
@pytest.mark.parametrize("case", CASES, ids=lambda c: "seed%d" % c["seed"])
def test_search_order_tensor_core_path(abi, case, bits, kind):
    """the tcgen05 one-hot GEMM scorer of the two-parent group, forced on for these small problems"""
    s, kw = build_case(case)
    golden = load_golden("ref_search_order")["seed%d" % case["seed"]]
    with abi.Context(0) as ctx:
        ctx.set_samples(s, perturbations=kw.get("perturbations"), num_cats=kw.get("num_cats"))
        ctx.force_generic(bits)
        kw2 = {k: v for k, v in kw.items() if k not in ("perturbations", "num_cats")}
        nets = ctx.search_order(**kw2)
        used = ctx.timing()["tensor_tiles"] > 0
        want = _tensor_eligible(case, ctx.num_cats(), bits)
        assert want is None or used == want
        assert ctx.timing()["tensor_kind"] == (kind if used else 0)
        compare_golden(nets, golden)


@pytest.mark.parametrize("bits,kind", TENSOR_KINDS_FULL, ids=["mxf4", "i8", "mxf4_full"])
@pytest.mark.parametrize("n,m,cats", [(70, 300, 4), (130, 257, 3), (66, 4100, None), (9, 3, 2), (200, 129, 4), (90, 20001, 4),
                                      (3, 40, 3), (4, 33, 2), (64, 64, 4), (65, 300, 2), (129, 100, 4), (193, 40, 3),
                                      (130, 7000, 4), (200, 6200, 3)])    # the last two: sample-sliced tail wave
def test_tensor_core_path_multi_tile(abi, n, m, cats, bits, kind):
    """several child tiles / pair tiles / K stages, ragged M: tensor-core scores == bit-plane scores"""
    s, c = random_problem(n + m, n, m, cats)
    kw = dict(max_parent_set=2, max_complexity=n * 8, order=(np.random.default_rng(n).permutation(n) + 1))
    with abi.Context(0) as ctx:
        ctx.set_samples(s, num_cats=c)
        ctx.force_generic(8)                      # never the tensor path
        a = ctx.search_order(want_probs=False, **kw)
        ctx.force_generic(bits)                   # always
        b = ctx.search_order(want_probs=False, **kw)
        assert ctx.timing()["tensor_tiles"] > 0 and ctx.timing()["tensor_kind"] == kind
        for world in (1, 3):                      # sharded over tile ranges
            import torch
            seg, nf = ctx.plan(0, world, **kw)
            buf = torch.full((seg * world,), float("nan"), dtype=torch.float64, device="cuda")
            for r in range(world):
                ctx.plan(r, world, **kw)
                ctx.score_shard(buf.data_ptr())
            torch.cuda.synchronize()
            c2 = ctx.finish(buf.data_ptr(), want_probs=False)
            assert compare_search(c2, a, check_probs=False) == len(a)
    assert compare_search(b, a, check_probs=False) == len(a)


@pytest.mark.parametrize("n,m,cats,na,pert", [(70, 300, 4, 0.05, False), (130, 257, 3, 0.0, True), (66, 4100, None, 0.1, True),
                                              (9, 3, 2, 0.2, False), (100, 129, 4, 0.02, True), (60, 9001, 4, 0.05, False),
                                              (3, 40, 3, 0.0, True), (49, 64, 4, 0.3, True), (97, 300, 2, 0.01, False),
                                              (150, 7000, 4, 0.05, True)])
def test_tensor_core_full_tables_masked_data(abi, port, n, m, cats, na, pert):
    """missing values and perturbation masks on the tensor cores: full-table tiles (16 rows per pair, 4 columns per child,
    the child's keep mask in its rows) == the bit-plane scorer, sharded and not; spot-checked tables against the oracle"""
    s, c = random_problem(3 * n + m, n, m, cats, na)
    rng = np.random.default_rng(n + m)
    p = (rng.random(s.shape) < 0.3).astype(np.int32) if pert else None
    kw = dict(max_parent_set=2, max_complexity=n * 8, order=(rng.permutation(n) + 1))
    with abi.Context(0) as ctx:
        ctx.set_samples(s, perturbations=p, num_cats=c)
        ctx.force_generic(8)                      # never the tensor path
        a = ctx.search_order(want_probs=True, **kw)
        assert ctx.timing()["tensor_tiles"] == 0
        ctx.force_generic(4)                      # always
        b = ctx.search_order(want_probs=True, **kw)
        assert ctx.timing()["tensor_tiles"] > 0 and ctx.timing()["tensor_kind"] == 4
        import torch
        for world in (1, 3):
            seg, nf = ctx.plan(0, world, **kw)
            buf = torch.full((seg * world,), float("nan"), dtype=torch.float64, device="cuda")
            for r in range(world):
                ctx.plan(r, world, **kw)
                ctx.score_shard(buf.data_ptr())
            torch.cuda.synchronize()
            c2 = ctx.finish(buf.data_ptr(), want_probs=False)
            assert compare_search(c2, a, check_probs=False) == len(a)
        # the search's own scores of random families against the oracle's family score
        ctx.plan(0, 1, **kw)
        buf = torch.full((ctx.plan(0, 1, **kw)[0],), float("nan"), dtype=torch.float64, device="cuda")
        ctx.score_shard(buf.data_ptr())
        torch.cuda.synchronize()
        if n >= 3:
            order0 = np.asarray(kw["order"]) - 1
            pos = np.array([sorted(rng.choice(n, size=3, replace=False)) for _ in range(40)])
            got = ctx.search_scores(buf.data_ptr(), pos[:, 2], pos[:, :2])
            s0 = np.where(s < 1, 99, s - 1).astype(np.int32)
            for f in range(len(pos)):
                ch, pa = int(order0[pos[f, 2]]), [int(order0[pos[f, 0]]), int(order0[pos[f, 1]])]
                keep = None if p is None else (p[:, ch] == 0)
                assert got[f] == port.family_score(s0, c, ch, pa, keep=keep)[1], (f, ch, pa)
    assert compare_search(b, a) == len(a)


@pytest.mark.parametrize("n,m,sparse", [(70, 300, False), (130, 9000, False), (66, 40, True)])
def test_tensor_core_epilogue_four_categories(abi, n, m, sparse):
    """the epilogue specialised for data whose nodes all have 4 categories == the generic epilogue, also with empty
    parent configurations and empty cells (few samples)"""
    s, c = random_problem(7 * n + m, n, m, 4)
    if sparse:
        s[:, : n // 2] = np.minimum(s[:, : n // 2], 2)  # categories 3 and 4 declared but never seen
    kw = dict(max_parent_set=2, max_complexity=n * 20, want_probs=False)
    with abi.Context(0) as ctx:
        ctx.set_samples(s, num_cats=c)
        ctx.force_generic(4 | 4096)
        a = ctx.search_order(**kw)
        assert ctx.timing()["tensor_tiles"] > 0
        ctx.force_generic(4)
        b = ctx.search_order(**kw)
        assert ctx.timing()["tensor_tiles"] > 0
        ctx.force_generic(8)
        d = ctx.search_order(**kw)
    assert compare_search(b, a, check_probs=False) == len(a)
    assert compare_search(b, d, check_probs=False) == len(d)


def _constrained_on_tiles(abi, port, seed):
    """-> True when the problem is restricted (pools / fixed parents / sizes) AND its two-parent group ran on the tiles"""
    s, P, K, kw = constrained_problem(seed)
    theirs = port.search_order(s, P, K, **kw)
    with abi.Context(0) as ctx:
        ctx.set_samples(s, perturbations=kw.get("perturbations"), num_cats=kw.get("num_cats"))
        ctx.force_generic(4)
        kw2 = {k: v for k, v in kw.items() if k not in ("perturbations", "num_cats")}
        if theirs is None:
            with pytest.raises(abi.CnbError):
                ctx.search_order(max_parent_set=P, max_complexity=K, **kw2)
            return False
        ours = ctx.search_order(max_parent_set=P, max_complexity=K, **kw2)
        restricted = any(k in kw for k in ("parents_pool", "fixed_parents", "parent_sizes"))
        used = ctx.timing()["tensor_tiles"] > 0 and restricted
    assert compare_search(ours, theirs) == len(theirs)
    return used


@pytest.mark.parametrize("seed", range(700, 760))
def test_constrained_search_on_tensor_tiles(abi, port, seed):
    """pools, fixed parents, parent-set sizes, perturbations, priors, missing values with the tensor-core path forced on: the
    candidates are a SUBSET of the unrestricted two-parent group, the tiles cover all of it and the selection looks every
    candidate's slot up from its parents; result == the oracle's"""
    _constrained_on_tiles(abi, port, seed)


def test_constrained_search_reaches_the_tiles(abi, port):
    """the previous test would be vacuous if no restricted problem reached the tiles"""
    assert sum(_constrained_on_tiles(abi, port, seed) for seed in range(760, 800)) >= 5


@pytest.mark.parametrize("seed", range(700, 740))
def test_search_order_constraints_vs_oracle(abi, port, seed):
    """parent pools, fixed parents, per-node parent-set sizes, perturbations, priors, 1..3 parents, tight complexity
    limits -- the problems tests/test_oracle.py pins the restatement on against the compiled reference"""
    s, P, K, kw = constrained_problem(seed)
    ours = abi.search_order(s, max_parent_set=P, max_complexity=K, **kw)
    theirs = port.search_order(s, P, K, **kw)
    assert len(ours) == len(theirs) and compare_search(ours, theirs) == len(theirs)


@pytest.mark.parametrize("seed", range(200, 230))
def test_search_order_random_vs_oracle(abi, port, seed):
    rng = np.random.default_rng(seed)
    n, m = int(rng.integers(2, 12)), int(rng.integers(1, 400))
    cats = [None, 2, 3, 4, None][seed % 5]
    s, c = random_problem(seed, n, m, cats, na_frac=[0.0, 0.1, 0.0][seed % 3])
    kw = dict(num_cats=c, max_parent_set=int(rng.integers(1, 4)))
    if seed % 4 == 0:
        kw["perturbations"] = (rng.random(s.shape) < 0.2).astype(np.int32)
    if seed % 5 == 0:
        kw["edge_prob"] = 0.1 + 0.8 * rng.random((n, n))
    if seed % 7 == 0:
        kw["fixed_parents"] = [[int(v) for v in rng.choice(np.arange(1, n + 1), size=rng.integers(0, 2), replace=False)]
                               for _ in range(n)]
    kw["order"] = (rng.permutation(n) + 1).astype(np.int32)
    kw["max_complexity"] = int(rng.integers(n, n * int(c.max()) ** kw["max_parent_set"] * 3 + 1))
    ours = abi.search_order(s, **kw)
    theirs = port.search_order(s, kw["max_parent_set"], kw["max_complexity"], **ref_args(kw))
    assert compare_search(ours, theirs) == len(theirs)


def test_search_vs_reference_core(abi, ref):
    """the reference's own C++ (if its library travelled here) on fresh problems"""
    for seed in range(300, 306):
        s, c = random_problem(seed, 9, 150, [None, 3][seed % 2], na_frac=0.05)
        theirs = ref.search_order(s, 2, 400, num_cats=c)
        ours = abi.search_order(s, max_parent_set=2, max_complexity=400, num_cats=c)
        assert compare_search(ours, theirs) == len(theirs)


def test_config_c1_alarm(abi):
    """BASELINE configs[0]: cnSearchOrder on cnSamples(alarmnet, 1000), true order, maxParentSet=2 (mixed categories:
    the per-candidate DP path, catnet_search2.h:681-841)"""
    from catnet_b200 import synth
    c = synth.config("C1")
    ours = abi.search_order(c["samples"], max_parent_set=2, max_complexity=c["max_complexity"], order=c["order"] + 1,
                            num_cats=c["cats"], want_probs=False)
    assert compare_golden(ours, load_golden("ref_configs")["C1"]) == 819


def test_config_c2_breast(abi):
    """BASELINE configs[1]: breast, 100 genes x 98 samples, maxParentSet=3, full complexity (4,087,875 families)"""
    from catnet_b200 import synth
    c = synth.config("C2")
    ours = abi.search_order(c["samples"], max_parent_set=3, max_complexity=c["max_complexity"], num_cats=c["cats"],
                            want_probs=False)
    assert compare_golden(ours, load_golden("ref_configs")["C2"]) == 684


def test_config_c3_reduced(abi):
    """BASELINE configs[2] at reduced M (the full-M reference run is ~1.6 h): 100 nodes x 3 cats, maxParentSet=3"""
    from catnet_b200 import synth
    c = synth.config("C3", m=300)
    ours = abi.search_order(c["samples"], max_parent_set=3, max_complexity=c["max_complexity"], order=c["order"] + 1,
                            num_cats=c["cats"], want_probs=False)
    assert compare_golden(ours, load_golden("ref_configs")["C3_m300"]) == 1261


def _check_result_consistency(nets, cats_pos, n):
    """size-independent invariants of a search result"""
    seen = set()
    for net in nets:
        cx = sum((cats_pos[i] - 1) * int(np.prod([cats_pos[p] for p in net["parents"][i]])) for i in range(n))
        assert cx == net["complexity"] and cx not in seen
        seen.add(cx)
        assert all(p < i for i in range(n) for p in net["parents"][i])      # parents precede the child in the order
        acc = 0.0
        for v, pr in zip(net["node_loglik"], net["node_prior"]):
            acc += (v + pr)
        assert acc == net["loglik"]                                         # re-summed in node order
    ll = [net["loglik"] for net in sorted(nets, key=lambda t: t["complexity"])]
    return ll


def test_config_c3_full_size_properties(abi, port):
    """BASELINE configs[2] at full size (100 x 3 cats, M=100,000, P=3: 4,087,875 families): invariants that do not
    need the 1.6 h reference run -- sample-permutation invariance, exact doubling, spot-checked tables"""
    from catnet_b200 import synth
    c = synth.config("C3")
    s, order, cats = c["samples"], c["order"] + 1, c["cats"]
    n = s.shape[1]
    kw = dict(max_parent_set=3, max_complexity=c["max_complexity"], order=order, num_cats=cats)
    a = abi.search_order(s, want_probs=False, **kw)
    cats_pos = [int(cats[v - 1]) for v in order]
    _check_result_consistency(a, cats_pos, n)
    rng = np.random.default_rng(0)
    b = abi.search_order(s[rng.permutation(len(s))], want_probs=False, **kw)     # counts do not depend on sample order
    assert compare_search(b, a, check_probs=False) == len(a)
    d2 = abi.search_order(np.concatenate([s, s]), want_probs=False, **kw)       # n -> 2n scales every term by 2 exactly
    assert compare_search(d2[:len(a)], [dict(t, sample_sizes=[2 * v for v in t["sample_sizes"]]) for t in a],
                          check_probs=False) == len(a)
    # spot-check family tables at full M against the oracle restatement
    s0 = (s - 1).astype(np.int32)
    with abi.Context(0) as ctx:
        ctx.set_samples(s, num_cats=cats)
        fams = [tuple(int(v) for v in rng.choice(n, size=4, replace=False)) for _ in range(40)]
        counts, ll, nc = ctx.family_scores([f[0] for f in fams], np.array([f[1:] for f in fams], dtype=np.int32))
        for i, f in enumerate(fams):
            rc, rll, rnc = port.family_score(s0, cats, f[0], list(f[1:]))
            assert np.array_equal(counts[i, :len(rc)], rc.astype(np.int64)) and ll[i] == rll and nc[i] == rnc


def test_config_c5_shape_properties(abi, port):
    """BASELINE configs[4] shape (500 nodes x 4 cats, P=2, 20,833,250 families) at M=20,000 so that the test stays
    short (the full 1M-sample run is bench.py): result invariants + spot-checked tables + sharded == unsharded"""
    from catnet_b200 import synth
    rng = np.random.Generator(np.random.PCG64(5005))
    net = synth.random_network(500, 4, 2, rng)
    s = synth.sample_network(net, 20000, rng)
    order = np.asarray(synth.order_nodes_descend(net["parents"])) + 1
    K = synth.default_max_complexity(500, 4, 2)
    assert K == 23999
    kw = dict(max_parent_set=2, max_complexity=K, order=order, num_cats=net["cats"])
    with abi.Context(0) as ctx:
        ctx.set_samples(s, num_cats=net["cats"])
        a = ctx.search_order(want_probs=False, **{k: v for k, v in kw.items() if k != "num_cats"})
        t = ctx.timing()
        assert t["num_families"] == 20833250 and t["fast_path"] == 1
        _check_result_consistency(a[::50], [4] * 500, 500)
        s0 = (s - 1).astype(np.int32)
        fams = [tuple(int(v) for v in rng.choice(500, size=3, replace=False)) for _ in range(60)]
        counts, ll, nc = ctx.family_scores([f[0] for f in fams], np.array([f[1:] for f in fams], dtype=np.int32))
        for i, f in enumerate(fams):
            rc, rll, rnc = port.family_score(s0, net["cats"], f[0], list(f[1:]))
            assert np.array_equal(counts[i, :len(rc)], rc.astype(np.int64)) and ll[i] == rll and nc[i] == rnc
    # the true network's families must be recoverable: the best network at the true complexity scores at least
    # as high as the generating structure re-scored on the same data
    true_cx = sum(3 * 4 ** len(p) for p in net["parents"])
    best = {t["complexity"]: t for t in a}
    assert true_cx in best


@pytest.mark.parametrize("case", SA_CASES, ids=lambda c: "seed%d" % c["seed"])
def test_sa_trajectory_matches_reference(abi, case):
    """cnSearchSA: same uniform stream -> same proposals, same accept/reject sequence, same final networks"""
    s, cats = random_problem(case["seed"], case["n"], case["m"], case.get("cats"))
    g = load_golden("ref_search_sa")["seed%d" % case["seed"]]
    sa = dict(case["sa"], start_order=np.arange(1, case["n"] + 1))
    r = abi.search_sa(s, sa, max_parent_set=case["P"], max_complexity=case["K"], num_cats=cats)
    assert [int(v) for v in r["order"]] == g["order"]
    assert (r["niter"], r["naccept"]) == (g["niter"], g["naccept"])
    assert [float(v) for v in r["trace"]] == [float.fromhex(v) for v in g["trace"]]
    assert [int(v) for v in r["trace_accept"]] == g["trace_accept"]
    compare_golden(r["nets"], g["nets"])


@pytest.mark.parametrize("case", HIST_CASES, ids=lambda c: "seed%d" % c["seed"])
def test_search_hist_matches_reference(abi, case):
    """cnSearchHist: same uniform stream -> same random orders, same selected networks, the
    parent histogram of the reference core bit for bit; one-shot call and resident context agree"""
    s, kw = build_hist_case(case)
    g = load_golden("ref_search_hist")["seed%d" % case["seed"]]
    args = dict(max_parent_set=case["P"], max_complexity=case["K"], num_cats=kw["num_cats"],
                perturbations=kw.get("perturbations"))
    h, it = abi.search_hist(s, case["hist"], **args)
    assert it == g["iterations"]
    assert h.tolist() == [[float.fromhex(v) for v in row] for row in g["hist"]]
    with abi.Context(0) as ctx:
        ctx.set_samples(s, perturbations=kw.get("perturbations"), num_cats=kw["num_cats"])
        h2, it2 = ctx.search_hist(case["hist"], max_parent_set=case["P"], max_complexity=case["K"])
        assert it2 == it and np.array_equal(h2, h)


@pytest.mark.parametrize("n,m,cats", [(12, 300, 3), (10, 257, 4), (14, 64, 2), (11, 1000, None), (9, 33, [2, 4, 3, 2, 4, 3, 2, 4, 3])])
def test_three_parent_inclusion_exclusion(abi, port, n, m, cats):
    """three-parent families on complete data: free cells + three-way tables (k_score_planes_ie3) against the direct
    bit-plane / histogram scorers and the oracle"""
    s, c = random_problem(7 * n + m, n, m, cats)
    kw = dict(max_parent_set=3, max_complexity=n * 40, order=(np.random.default_rng(m).permutation(n) + 1))
    with abi.Context(0) as ctx:
        ctx.set_samples(s, num_cats=c)
        a = ctx.search_order(want_probs=True, **kw)
        assert ctx.timing()["fast_path"] == 1
        ctx.force_generic(512)                    # no inclusion-exclusion for three parents
        b = ctx.search_order(want_probs=True, **kw)
    assert compare_search(a, b) == len(b)
    if m <= 300:
        o = port.search_order(s, 3, kw["max_complexity"], order=kw["order"], num_cats=c)
        assert compare_search(a, o) == len(o)


@pytest.mark.parametrize("n,m,cats", [(40, 300, 3), (70, 9000, 2), (100, 8200, 3), (130, 500, 2), (200, 400, 3),
                                      (97, 260, [2, 3] * 48 + [2]), (260, 300, 2), (5, 40, 2), (3, 64, 3)])
def test_reduced_tile_geometry(abi, port, n, m, cats):
    """data with fewer than four categories per node: the tensor-core tiles contract only the free cells there are -- 4 rows
    per pair and 2 columns per child for three categories (64 x 96 family slots per tile), 1 and 1 for binary data
    (256 x 192) -- and must give the tables, log-likelihoods and networks of the 28 x 64 tiles and of the bit-plane scorer"""
    s, c = random_problem(3 * n + m, n, m, cats)
    kw = dict(max_parent_set=2, max_complexity=n * 12, order=(np.random.default_rng(n).permutation(n) + 1))
    with abi.Context(0) as ctx:
        ctx.set_samples(s, num_cats=c)
        ctx.force_generic(4)                       # tensor cores, tiles of the data's own geometry
        a = ctx.search_order(**kw)
        ta = ctx.timing()["tensor_tiles"]
        ctx.force_generic(4 | 131072)              # ... always 28 x 64 tiles
        b = ctx.search_order(**kw)
        tb = ctx.timing()["tensor_tiles"]
        ctx.force_generic(8)                       # bit planes
        d = ctx.search_order(**kw)
        assert ctx.timing()["tensor_tiles"] == 0
    assert 0 < ta <= tb and (n < 64 or ta < tb)
    for other in (b, d):
        assert compare_search(a, other) == len(other)
        for x, y in zip(a, other):
            assert x["loglik"] == y["loglik"] and x["node_loglik"] == y["node_loglik"]
    if m <= 300 and n <= 130:
        o = port.search_order(s, 2, kw["max_complexity"], order=kw["order"], num_cats=c)
        assert compare_search(a, o) == len(o)


@pytest.mark.parametrize("n,m,cats", [(12, 300, 3), (10, 257, 4), (14, 64, 2), (11, 1000, None), (70, 200, 3), (66, 9000, 2),
                                      (30, 8300, 4), (4, 50, 3), (5, 33, 2), (65, 40, 4), (129, 70, 2)])
def test_three_parent_tensor_core_path(abi, n, m, cats):
    """the unrestricted three-parent group as T tiles on the tensor cores (virtual pair nodes x third parent x children)
    against the bit-plane inclusion-exclusion kernel; two column tiles at n > 64; forced on for small M"""
    s, c = random_problem(11 * n + m, n, m, cats)
    kw = dict(max_parent_set=3, max_complexity=n * 30, order=(np.random.default_rng(n + m).permutation(n) + 1))
    with abi.Context(0) as ctx:
        ctx.set_samples(s, num_cats=c)
        ctx.force_generic(1024)                   # bit planes
        a = ctx.search_order(want_probs=False, **kw)
        assert ctx.timing()["tensor3_tiles"] == 0
        ctx.force_generic(4 if m < 8192 else 0)   # tensor cores (forced below the sample threshold)
        b = ctx.search_order(want_probs=False, **kw)
        assert ctx.timing()["tensor3_tiles"] > 0
    assert compare_search(b, a, check_probs=False) == len(a)


@pytest.mark.parametrize("n,m,cats", [(30, 200, 3), (70, 300, 2), (40, 8300, 4), (101, 120, 3), (20, 500, 4)])
def test_three_parent_tiles_batched_and_dealt(abi, monkeypatch, n, m, cats):
    """the T tiles of a column tile contracted in batches that fit a (here tiny) table buffer, and dealt to several ranks by
    triple range -- every rank then scores ALL children of the column tile for its triples, so the rank buffers (started at
    -inf) are combined by an element-wise maximum instead of the all-gather: same networks as one pass on one rank"""
    import torch
    s, c = random_problem(19 * n + m, n, m, cats)
    kw = dict(max_parent_set=3, max_complexity=n * 30, order=(np.random.default_rng(n + m).permutation(n) + 1))
    with abi.Context(0) as ctx:
        ctx.set_samples(s, num_cats=c)
        ctx.force_generic(4 if m < 8192 else 0)
        a = ctx.search_order(want_probs=False, **kw)
        whole = ctx.timing()["tensor3_tiles"]
        assert whole > 0
        monkeypatch.setenv("CNB_T3_BATCH_BYTES", "1")          # one aligned group of tiles per batch
        b = ctx.search_order(want_probs=False, **kw)
        assert ctx.timing()["tensor3_tiles"] == whole
        assert compare_search(b, a, check_probs=False) == len(a)
        monkeypatch.delenv("CNB_T3_BATCH_BYTES")
        ctx.allow_scatter(2)
        for world in (2, 3, 5):
            seg, _ = ctx.plan(0, world, **kw)
            bufs, tiles = [], 0
            for r in range(world):
                ctx.plan(r, world, **kw)
                buf = torch.full((seg * world,), float("nan"), dtype=torch.float64, device="cuda")
                ctx.score_shard(buf.data_ptr())
                assert ctx.scattered()
                tiles += ctx.timing()["tensor3_tiles"]
                bufs.append(buf)
            assert tiles == whole
            full = torch.stack(bufs).amax(dim=0)                # what all_reduce(MAX) leaves on every rank
            nets = ctx.finish(full.data_ptr(), want_probs=False)
            assert compare_search(nets, a, check_probs=False) == len(a)
        ctx.allow_scatter(0)                                   # without the promise: bit planes on several ranks
        ctx.plan(0, 2, **kw)
        buf = torch.empty((seg * 2 + 8,), dtype=torch.float64, device="cuda")
        ctx.score_shard(buf.data_ptr())
        assert not ctx.scattered() and ctx.timing()["tensor3_tiles"] == 0


@pytest.mark.parametrize("seed", [4, 19, 55, 223, 226, 349, 439, 487, 958])
def test_fixed_parent_listing_order_on_tensor_tiles(abi, port, seed):
    """found by tools/fuzz_gpu.py with the tensor path forced: a two-parent candidate (fixed parent f, free parent x) is
    listed [f, x] even when x precedes f in the order, and the reference sums its table in that listing's order; the tile
    epilogue summed every slot parent-ascending -> last-bit different scores -> last-bit different DP totals"""
    s, P, K, kw = constrained_problem(seed)
    theirs = port.search_order(s, P, K, **kw)
    with abi.Context(0) as ctx:
        ctx.set_samples(s, perturbations=kw.get("perturbations"), num_cats=kw.get("num_cats"))
        ctx.force_generic(4)
        kw2 = {k: v for k, v in kw.items() if k not in ("perturbations", "num_cats")}
        ours = ctx.search_order(max_parent_set=P, max_complexity=K, **kw2)
        assert ctx.timing()["tensor_tiles"] > 0
    assert compare_search(ours, theirs) == len(theirs)


@pytest.mark.parametrize("n,m,cats,P", [(30, 400, 3, 2), (12, 200, 4, 3), (70, 9000, 4, 2), (9, 150, 6, 2)])
def test_plan_is_reused_across_orders(abi, port, monkeypatch, n, m, cats, P):
    """without pools, fixed parents, sizes, priors and with one category count the plan depends on the order only through the
    order itself: a context keeps it and a later search replaces just the order (SA chains, histogram batches, benchmark
    loops).  Same networks as a context that plans from scratch, for a sequence of orders; a bad order is still refused;
    a constrained search in between plans afresh"""
    s, c = random_problem(31 * n + m, n, m, cats)
    rng = np.random.default_rng(n)
    orders = [(rng.permutation(n) + 1).astype(np.int32) for _ in range(4)] + [None]
    kw = dict(max_parent_set=P, max_complexity=n * 20)
    with abi.Context(0) as ctx:
        ctx.set_samples(s, num_cats=c)
        got = [ctx.search_order(want_probs=False, order=o, **kw) for o in orders]
        with pytest.raises(abi.CnbError) as e:
            ctx.search_order(order=np.r_[orders[0][:-1], orders[0][0]].astype(np.int32), **kw)
        assert "Invalid nodeOrder" in str(e.value)
        pools = [[j + 1 for j in range(n) if j != i and (i + j) % 3] for i in range(n)]
        mid = ctx.search_order(want_probs=False, order=orders[1], parents_pool=pools, **kw)
        again = ctx.search_order(want_probs=False, order=orders[2], **kw)
        monkeypatch.setenv("CNB_NO_PLAN_REUSE", "1")
        want = [ctx.search_order(want_probs=False, order=o, **kw) for o in orders]
        want_mid = ctx.search_order(want_probs=False, order=orders[1], parents_pool=pools, **kw)
    for a, b in zip(got + [mid, again], want + [want_mid, want[2]]):
        assert compare_search(a, b, check_probs=False) == len(b)
        assert [x["loglik"] for x in a] == [y["loglik"] for y in b]
    if m <= 400:
        o = port.search_order(s, P, kw["max_complexity"], order=orders[3], num_cats=c, want_probs=False)
        assert compare_search(got[3], o, check_probs=False) == len(o)


def test_host_compacted_upload(abi, monkeypatch):
    """large host matrices cross PCIe as bytes (int32 -> uint8 by host threads, 0 = missing): same packed data, hence the
    same search, as the plain int32 upload; NA and category inference included; a value above 255 is still an error"""
    s, c = random_problem(123, 9, 70000, None, na_frac=0.03)
    kw = dict(max_parent_set=2, max_complexity=9 * 16 * 3)
    res = {}
    for mode in ("off", "1"):
        monkeypatch.setenv("CNB_HOST_COMPACT_MIN", mode)
        with abi.Context(0) as ctx:
            ctx.set_samples(s)                          # categories inferred on the device
            res[mode] = (ctx.num_cats().tolist(), ctx.search_order(want_probs=True, **kw), ctx.timing()["h2d_bytes"])
    assert res["off"][0] == res["1"][0] == c.tolist()
    assert compare_search(res["1"][1], res["off"][1]) == len(res["off"][1])
    assert res["1"][2] < res["off"][2] / 3                  # a quarter of the bytes
    bad = s.copy()
    bad[5, 2] = 300
    with abi.Context(0) as ctx, pytest.raises(abi.CnbError) as e:
        ctx.set_samples(bad, num_cats=c)
    assert e.value.code == abi.CNB_ERR_PARAM


def test_byte_matrix_on_device(abi):
    """cnb_ctx_set_samples_dev8 (what the multi-process upload all-gathers) and dist.set_samples_sharded at world 1"""
    import torch
    from catnet_b200 import dist as cdist
    s, c = random_problem(321, 8, 3000, None, na_frac=0.04)
    kw = dict(max_parent_set=2, max_complexity=8 * 16 * 3)
    with abi.Context(0) as ctx:
        ctx.set_samples(s, num_cats=c)
        a = ctx.search_order(want_probs=True, **kw)
        d8 = torch.as_tensor(s, device="cuda").clamp_(0, 255).to(torch.uint8)
        ctx.set_samples_dev8(d8.data_ptr(), 8, 3000, num_cats=c)
        b = ctx.search_order(want_probs=True, **kw)
        host = torch.as_tensor(s).pin_memory()
        full, nbytes = cdist.set_samples_sharded(ctx, host, 0, 1, torch.device("cuda", 0), num_cats=c)
        torch.cuda.synchronize()
        d = ctx.search_order(want_probs=True, **kw)
    assert compare_search(b, a) == len(a) and compare_search(d, a) == len(a) and nbytes == s.size   # one byte per value


def test_one_shot_context_is_parked_and_released(abi):
    """one-shot calls reuse the context the previous one parked (streams, events, device buffers), also when the data set
    changes shape; cnb_release_cache frees it and the next call starts from scratch -- same results every time"""
    L = abi.lib()
    s1, c1 = random_problem(1, 9, 400, None)
    s2, c2 = random_problem(2, 14, 90, 3, na_frac=0.05)
    kw1 = dict(max_parent_set=2, max_complexity=9 * 16 * 3, num_cats=c1)
    kw2 = dict(max_parent_set=3, max_complexity=14 * 27 * 2, num_cats=c2)
    a1 = abi.search_order(s1, **kw1)
    a2 = abi.search_order(s2, **kw2)                 # bigger N, different categories, NA: takes over a1's context
    b1 = abi.search_order(s1, **kw1)
    L.cnb_release_cache()
    d2 = abi.search_order(s2, **kw2)
    e = abi.pairwise("entropy", np.where(s1 < 1, 1, s1))
    L.cnb_release_cache()
    f = abi.pairwise("entropy", np.where(s1 < 1, 1, s1))
    assert compare_search(b1, a1) == len(a1) and compare_search(d2, a2) == len(a2) and np.array_equal(e, f)


def test_more_than_65535_sample_words(abi, port):
    """2.3 M samples = 71,875 words of 32: the pack / permute / operand-row kernels put the sample dimension on grid.x"""
    rng = np.random.default_rng(0)
    m, n = 2_300_000, 6
    s = rng.integers(1, 4, size=(m, n), dtype=np.int32)
    s[:, 1] = np.where(rng.random(m) < 0.5, s[:, 0], s[:, 1])
    c = np.full(n, 3, dtype=np.int32)
    with abi.Context(0) as ctx:
        ctx.set_samples(s, num_cats=c)
        counts, ll, nc = ctx.family_scores([1, 3], np.array([[0], [2]], dtype=np.int32))
        nets = ctx.search_order(want_probs=False, max_parent_set=2, max_complexity=200)
        assert ctx.timing()["tensor_tiles"] > 0
    s0 = (s - 1).astype(np.int32)
    for f, (ch, pa) in enumerate([(1, [0]), (3, [2])]):
        rc, rll, rnc = port.family_score(s0, c, ch, pa)
        assert np.array_equal(counts[f, :9], rc.astype(np.int64)) and ll[f] == rll and nc[f] == rnc
    assert len(nets) > 1 and 1 in [len(p) for p in nets[-1]["parents"]] + [1]


@pytest.mark.parametrize("n,m,cats,P", [(300, 400, 4, 2), (120, 300, 2, 2), (90, 200, 3, 2), (40, 3000, 4, 3), (513, 64, 2, 1)])
def test_dp_trapezoid_equals_wavefront(abi, n, m, cats, P):
    """equal-category searches with many complexity slots (several CTAs of 256 slots, halos of different widths, node
    counts that are not a multiple of the macro-step): k_dp_trap == k_dp_wave == per-node launches, network by network"""
    s, c = random_problem(5 * n + m, n, m, cats)
    kw = dict(max_parent_set=P, max_complexity=n * (cats ** P) * (cats - 1))
    with abi.Context(0) as ctx:
        ctx.set_samples(s, num_cats=c)
        a = ctx.search_order(want_probs=False, **kw)
        ctx.force_generic(2048)
        b = ctx.search_order(want_probs=False, **kw)
        ctx.force_generic(32)
        d = ctx.search_order(want_probs=False, **kw)
        ctx.force_generic(16384)                  # trapezoid without the CTA-wide work list of re-summation chains
        e = ctx.search_order(want_probs=False, **kw)
    assert compare_search(a, b, check_probs=False) == len(b) and compare_search(a, d, check_probs=False) == len(d)
    assert compare_search(a, e, check_probs=False) == len(e)


@pytest.mark.parametrize("n,m,cats,P,grid", [(150, 64, 2, 2, True), (120, 4096, 3, 2, True), (200, 333, 4, 2, False), (90, 1024, 2, 3, True),
                                             (257, 50, 3, 2, False)])
def test_dp_work_list(abi, n, m, cats, P, grid):
    """the trapezoid DP compacts the viable candidates of a CTA into a work list before re-summing them: same networks and
    bit-identical log-likelihoods as the per-slot chains and as the grid-barrier kernel"""
    s, c = random_problem(11 * n + m, n, m, cats, structured=not grid)
    kw = dict(max_parent_set=P, max_complexity=n * (cats ** P) * (cats - 1))
    with abi.Context(0) as ctx:
        ctx.set_samples(s, num_cats=c)
        a = ctx.search_order(want_probs=False, **kw)
        ctx.force_generic(16384)
        b = ctx.search_order(want_probs=False, **kw)
        ctx.force_generic(64)                     # grid-barrier kernel
        d = ctx.search_order(want_probs=False, **kw)
    assert compare_search(a, b, check_probs=False) == len(b) and compare_search(a, d, check_probs=False) == len(d)
    assert [x["loglik"] for x in a] == [x["loglik"] for x in b] == [x["loglik"] for x in d]


def test_pair_tables_on_tensor_cores(abi):
    """the two-way tables of all node pairs come from the tcgen05 kernel when M >= 8192 (P tiles) and from the POPC
    kernel otherwise: both routes must give the same searches (one- and two-parent scores are built on them)"""
    s, c = random_problem(91, 70, 9000, 4)
    kw = dict(max_parent_set=2, max_complexity=70 * 16 * 3)
    with abi.Context(0) as ctx:
        ctx.force_generic(128)                    # POPC pair tables
        ctx.set_samples(s, num_cats=c)
        a = ctx.search_order(want_probs=True, **kw)
    with abi.Context(0) as ctx:
        ctx.set_samples(s, num_cats=c)            # tensor-core pair tables
        b = ctx.search_order(want_probs=True, **kw)
        ctx.force_generic(8)                      # ... consumed by the bit-plane inclusion-exclusion scorer
        d = ctx.search_order(want_probs=True, **kw)
    assert compare_search(b, a) == len(a) and compare_search(d, a) == len(a)


@pytest.mark.parametrize("generic", [False, True], ids=["auto", "histogram"])
@pytest.mark.parametrize("case", PAIRWISE_CASES, ids=lambda c: "seed%d" % c["seed"])
def test_pairwise_matches_reference(abi, case, generic):
    """cnEntropy / cnEdgeDistanceKL / cnEdgeDistancePearson / cnEntropyOrder (catnet_entropy.cpp) against the outputs
    of the reference's own functions, bit for bit; `generic` forces the shared-memory histogram scorer"""
    s, p = build_pairwise_case(case)
    g = load_golden("ref_pairwise")["seed%d" % case["seed"]]
    with abi.Context(0) as ctx:
        if generic:
            ctx.force_generic(1)
        ctx.set_samples(s, perturbations=p)
        for kind in ("entropy", "kl", "pearson"):
            assert ctx.pairwise(kind).tolist() == [[float.fromhex(v) for v in row] for row in g[kind]], kind
        if g["order"] is not None:
            assert ctx.entropy_order().tolist() == g["order"]
    if not generic:
        # the one-shot entry points the R shim calls
        assert abi.pairwise("kl", s, p).tolist() == [[float.fromhex(v) for v in row] for row in g["kl"]]
        if g["order"] is not None:
            assert abi.entropy_order(s, p).tolist() == g["order"]


@pytest.mark.parametrize("seed", range(400, 420))
def test_pairwise_random_vs_oracle(abi, port, seed):
    rng = np.random.default_rng(seed)
    m, n, c = int(rng.integers(1, 700)), int(rng.integers(1, 30)), int(rng.integers(1, 9))
    s = rng.integers(1, c + 1, size=(m, n)).astype(np.int32) + int(rng.integers(0, 3))
    if n > 2:
        s[:, 2] = np.where(rng.random(m) < 0.7, s[:, 1], s[:, 2])
    p = (rng.random((m, n)) < rng.random()).astype(np.int32) if seed % 2 else None
    with abi.Context(0) as ctx:
        ctx.set_samples(s, perturbations=p)
        for kind in ("entropy", "kl", "pearson"):
            assert np.array_equal(ctx.pairwise(kind), port.pairwise(kind, s, p)), kind
        if n > 1:
            assert np.array_equal(ctx.entropy_order(), port.pairwise("order", s, p))


def test_pairwise_large_properties(abi, port):
    """C5-shaped data (reduced M): the matrices agree with the oracle on a sample of entries, H(n1|n2) <= H(n1),
    mutual information is symmetric, and permuting the samples changes nothing"""
    rng = np.random.default_rng(7)
    m, n = 40000, 120
    s = rng.integers(1, 5, size=(m, n)).astype(np.int32)
    s[:, 1::2] = np.where(rng.random((m, n // 2)) < 0.6, s[:, 0::2], s[:, 1::2])
    p = (rng.random((m, n)) < 0.3).astype(np.int32)
    with abi.Context(0) as ctx:
        ctx.set_samples(s)
        h = ctx.pairwise("entropy")
        assert np.array_equal(ctx.pairwise("kl"), np.zeros((n, n)))
    d = np.diag(h)
    assert (h <= d[:, None] + 1e-12).all()
    mi = d[:, None] - h
    assert np.allclose(mi, mi.T, rtol=0, atol=1e-12)
    sub = [0, 1, 2, 57, 119]
    ho = port.pairwise("entropy", s[:, sub])
    assert np.array_equal(h[np.ix_(sub, sub)], ho)
    perm = rng.permutation(m)
    with abi.Context(0) as ctx:
        ctx.set_samples(s, perturbations=p)
        e1, k1, x1 = ctx.pairwise("entropy"), ctx.pairwise("kl"), ctx.pairwise("pearson")
        o1 = ctx.entropy_order()
        ctx.set_samples(s[perm], perturbations=p[perm])
        assert np.array_equal(e1, ctx.pairwise("entropy")) and np.array_equal(k1, ctx.pairwise("kl"))
        assert np.array_equal(x1, ctx.pairwise("pearson")) and np.array_equal(o1, ctx.entropy_order())
    for kind, mat in (("entropy", e1), ("kl", k1), ("pearson", x1)):
        assert np.array_equal(mat[np.ix_(sub, sub)], port.pairwise(kind, s[:, sub], p[:, sub])), kind


def test_pairwise_python_front_end(abi, port):
    from catnet_b200 import search
    s, _ = random_problem(77, 6, 120, 3)
    p = (np.random.default_rng(5).random(s.shape) < 0.3).astype(np.int32)
    data = (s.T * 10).astype(np.float64)           # nodes x samples, arbitrary category values
    assert np.array_equal(search.cnEntropy(data), port.pairwise("entropy", s))
    assert np.array_equal(search.cnEdgeDistanceKL(data, p.T), port.pairwise("kl", s, p))
    assert np.array_equal(search.cnEdgeDistancePearson(data, p.T), port.pairwise("pearson", s, p))
    assert np.array_equal(search.cnEntropyOrder(data, p.T), port.pairwise("order", s, p))
    bad = s.copy()
    bad[4, 2] = NA
    with pytest.raises(abi.CnbError) as e:
        abi.pairwise("entropy", bad)
    assert e.value.code == abi.CNB_ERR_PARAM


class _DeviceNetwork:
    """the product behind the signatures of oracle.port.net_* (zero-based codes in, NA = anything out of range)"""

    def __init__(self, abi, generic=False, one_shot=False):
        self.abi, self.generic, self.one_shot = abi, generic, one_shot

    @staticmethod
    def _lbl(s0, cats):
        s0 = np.asarray(s0)
        return np.where((s0 < 0) | (s0 >= np.asarray(cats)[None, :]), NA, s0 + 1).astype(np.int32)

    def net_set_prob(self, s0, cats, parents, pert=None):
        s = self._lbl(s0, cats)
        if self.one_shot:
            return self.abi.set_prob(s, cats, parents, pert)
        with self.abi.Context(0) as ctx:
            if self.generic:
                ctx.force_generic(1)
            ctx.set_samples(s, perturbations=pert, num_cats=cats)
            return ctx.set_prob(cats, parents)

    def net_loglik(self, mode, s0, cats, parents, probs, pert=None):
        s = self._lbl(s0, cats)
        if mode == "node":        # sampleNodeLoglik of every node: no perturbations, listed one by one
            return self.abi.node_loglik(s, cats, parents, probs, np.arange(1, len(cats) + 1))
        if self.one_shot:
            return self.abi.loglik(s, cats, parents, probs, pert, mode)
        with self.abi.Context(0) as ctx:
            ctx.set_samples(s, perturbations=pert, num_cats=cats)
            return ctx.loglik(cats, parents, probs, mode)


@pytest.mark.parametrize("variant", ["ctx", "histogram", "one_shot"])
@pytest.mark.parametrize("case", NETWORK_CASES, ids=lambda c: "seed%d" % c["seed"])
def test_network_matches_reference(abi, case, variant):
    """cnSetProb / cnLoglik / cnNodeLoglik on the device against the reference's CATNET methods, bit for bit"""
    impl = _DeviceNetwork(abi, generic=variant == "histogram", one_shot=variant == "one_shot")
    check_network_case(impl, case, build_network_case, load_golden("ref_network")["seed%d" % case["seed"]])


@pytest.mark.parametrize("seed", range(600, 616))
def test_network_random_vs_oracle(abi, port, seed):
    rng = np.random.default_rng(seed)
    n, m = int(rng.integers(1, 40)), int(rng.integers(6, 5000))
    s, c = random_problem(seed, n, m, [None, 2, 4, 7][seed % 4], na_frac=[0, 0.05][seed % 2])
    s0 = np.where(s < 1, 99, s - 1).astype(np.int32)
    parents = [sorted(rng.choice(i, size=min(i, int(rng.integers(0, 4))), replace=False).tolist(), reverse=seed % 3 == 0)
               for i in range(n)]
    pert = (rng.random(s.shape) < 0.3).astype(np.int32) if seed % 3 == 1 else None
    impl = _DeviceNetwork(abi)
    pa, la, sa = impl.net_set_prob(s0, c, parents, pert)
    pb, lb, sb = port.net_set_prob(s0, c, parents, pert)
    assert all(np.array_equal(x, y) for x, y in zip(pa, pb)) and np.array_equal(la, lb) and np.array_equal(sa, sb)
    if seed % 4 == 0:
        pa = [np.where(t < 0.1, 0.0, t) for t in pa]
    for mode in ("nodes", "bysample"):
        assert np.array_equal(impl.net_loglik(mode, s0, c, parents, pa, pert), port.net_loglik(mode, s0, c, parents, pa, pert)), mode
    sel = rng.choice(n, size=min(n, 3), replace=False) + 1
    a = abi.node_loglik(impl._lbl(s0, c), c, parents, pa, sel, pert)
    assert np.array_equal(a, port.net_loglik("nodes", s0, c, parents, pa, pert)[sel - 1])


def test_network_errors(abi):
    s, c = random_problem(5, 4, 50, 3)
    parents = [[], [0], [0, 1], [2]]
    probs, _, _ = abi.set_prob(s, c, parents)
    for bad_par in ([[], [1], [0, 1], [2]], [[], [7], [0, 1], [2]]):          # self-parent, out of range
        with pytest.raises(abi.CnbError) as e:
            abi.loglik(s, c, bad_par, probs)
        assert e.value.code == abi.CNB_ERR_PARAM
    s2 = s.copy()
    s2[3, 2] = 4                                                              # more categories than the object
    with pytest.raises(abi.CnbError) as e:
        abi.loglik(s2, c, parents, probs)
    assert e.value.code == abi.CNB_ERR_PARAM
    with pytest.raises(abi.CnbError):
        abi.node_loglik(s, c, parents, probs, [9])


def test_search_then_loglik_round_trip(abi):
    """the networks cnSearchOrder returns carry likelihoods that cnLoglik reproduces from their own tables:
    mean over samples of the by-sample values = sum of the node values = the network's log-lik (to rounding)"""
    s, c = random_problem(21, 9, 400, 3)
    nets = abi.search_order(s, max_parent_set=2, max_complexity=9 * 27 * 2, num_cats=c)
    net = nets[len(nets) // 2]
    parents = [[q for q in p] for p in net["parents"]]           # identity order: positions = labels
    nodes = abi.loglik(s, c, parents, net["probs"], mode="nodes")
    assert np.allclose(nodes, net["node_loglik"], rtol=1e-12, atol=0)
    assert abs(nodes.sum() - net["loglik"]) <= 1e-12 * abs(net["loglik"])
    by = abi.loglik(s, c, parents, net["probs"], mode="bysample")
    assert abs(by.mean() - net["loglik"]) <= 1e-10 * abs(net["loglik"])


def test_network_python_front_end(abi, port):
    """cnSearchOrder -> cnSetProb on other data -> cnLoglik / cnNodeLoglik, through the mirrors of the R front-ends"""
    from catnet_b200 import search
    s, c = random_problem(33, 7, 300, 3)
    data = s.T.astype(np.float64) * 5
    ev = search.cnSearchOrder(data[:, :150], maxParentSet=2)
    net = ev.nets[len(ev.nets) // 2]
    fit = search.cnSetProb(net, data[:, 150:])
    parents = [[q - 1 for q in p] for p in net["parents"]]
    s0 = (s - 1).astype(np.int32)
    probs, ll, sizes = port.net_set_prob(s0[150:], c, parents)
    assert all(np.array_equal(a, b) for a, b in zip(fit["probabilities"], probs)) and fit["likelihood"] == float(ll.sum())
    def inf(v):                                   # the front-ends export -FLT_MAX as -Inf, like rcatnet.cpp:295-300
        return np.where(v <= -3.4028234663852886e38, -np.inf, v)
    assert search.cnLoglik(fit, data) == float(inf(port.net_loglik("nodes", s0, c, parents, probs)).sum())
    assert np.array_equal(search.cnLoglik(fit, data, bysample=True), inf(port.net_loglik("bysample", s0, c, parents, probs)))
    assert np.array_equal(search.cnNodeLoglik(fit, [2, 5], data), inf(port.net_loglik("nodes", s0, c, parents, probs))[[1, 4]])
    assert np.isfinite(search.cnLoglik(fit, data[:, 150:]))        # its own training data: no zero-probability hit
    # cnSamples from the fitted network: indices, category values, perturbations, NA rate
    draw = search.cnSamples(fit, 500, seed=3)
    o, _ = port.net_samples(c, parents, probs, 500, None, 0.0, 3)
    assert draw.shape == (7, 500) and np.array_equal(draw, o.T)
    vals = search.cnSamples(fit, 500, as_index=False, seed=3)
    assert set(np.unique(vals)) <= {5.0, 10.0, 15.0}
    fixed = np.zeros(7, dtype=np.int32)
    fixed[2] = 3
    pert = search.cnSamples(fit, 50, perturbations=fixed, naRate=0.3, seed=4)
    assert ((pert[2] == 3) | (pert[2] == NA)).all() and (pert == NA).sum() == 50 * int(0.3 * 7)


@pytest.mark.parametrize("case", SAMPLES_CASES, ids=lambda c: "seed%d" % c["seed"])
def test_samples_match_reference(abi, case):
    """cnSamples on the device: the reference's own samples under the shared uniform stream, value for value"""
    cats, parents, probs, pert, na = build_samples_case(case)
    g = load_golden("ref_samples")["seed%d" % case["seed"]]
    s = abi.samples(cats, parents, probs, case["m"], pert, na, seed=case["seed"])
    assert s.ravel().tolist() == g["samples"]


@pytest.mark.parametrize("seed", range(800, 812))
def test_samples_random_vs_oracle(abi, port, seed):
    m = 1 + 523 * (seed % 6)
    cats, parents, probs, pert, na = build_samples_case(dict(seed=seed, n=1 + 3 * (seed % 12), m=m, maxcat=6, maxpar=3,
                                                             pert=[0, 0.4][seed % 2], na=[0, 0.2, 0.7][seed % 3]))
    b, _ = port.net_samples(cats, parents, probs, m, pert, na, seed)
    assert np.array_equal(abi.samples(cats, parents, probs, m, pert, na, seed=seed), b)


def test_samples_large_statistics_and_device_output(abi, port):
    """C5-shaped sampling straight into device memory: the first rows equal the oracle's, the empirical conditional
    frequencies approach the tables, and the matrix feeds cnb_ctx_set_samples_dev without a host round trip"""
    import torch
    from catnet_b200 import synth
    rng = np.random.Generator(np.random.PCG64(5))
    net = synth.random_network(60, 4, 2, rng)
    cats, parents, probs = net["cats"], [list(p) for p in net["parents"]], [np.asarray(t).ravel() for t in net["cpts"]]
    m = 400000
    out = torch.empty((m, 60), dtype=torch.int32, device="cuda")
    abi.samples_dev(cats, parents, probs, m, out.data_ptr(), seed=99)
    torch.cuda.synchronize()
    s = out.cpu().numpy()
    assert s.min() >= 1 and (s.max(axis=0) <= cats).all()
    # samples j < 1000 of every node use draws k*M + j: compare with the oracle run on the same M through a slice
    o, _ = port.net_samples(cats, parents, probs, m, None, 0.0, 99)
    assert np.array_equal(s, o)
    i = max(range(60), key=lambda k: len(parents[k]))
    cfg = np.zeros(m, dtype=np.int64)
    for q in parents[i]:
        cfg = cfg * int(cats[q]) + (s[:, q] - 1)
    tab = np.asarray(probs[i]).reshape(-1, cats[i])
    for r in range(tab.shape[0]):
        sel = s[cfg == r, i]
        if len(sel) > 5000:
            freq = np.bincount(sel - 1, minlength=cats[i]) / len(sel)
            assert np.abs(freq - tab[r]).max() < 0.03
    with abi.Context(0) as ctx:
        ctx.set_samples_dev(out.data_ptr(), 60, m, num_cats=cats)
        fit, _, sizes = ctx.set_prob(cats, parents)
        assert all(v == m for v in sizes) and np.abs(fit[i] - np.asarray(probs[i])).max() < 0.03


def test_samples_errors(abi):
    cats = np.array([2, 2], dtype=np.int32)
    with pytest.raises(abi.CnbError) as e:                     # a directed cycle: the reference returns NULL
        abi.samples(cats, [[1], [0]], [np.full(4, 0.5), np.full(4, 0.5)], 10)
    assert e.value.code == abi.CNB_ERR_PARAM


def test_search_hist_python_front_end(abi):
    """catnet_b200.search.cnSearchHist: R's matrix(vhisto, nrow=numnodes) is [child, parent]"""
    from catnet_b200 import search
    s, cats = random_problem(46, 6, 150, 3)
    m = search.cnSearchHist(s.T.astype(float), maxParentSet=2, score="BIC", weight="likelihood", maxIter=6, numThreads=2,
                            seed=3)
    assert m.shape == (6, 6) and np.all(np.diag(m) == 0) and m.sum() > 0


def test_sa_alarm_vs_oracle(abi, port):
    """BASELINE configs[3] shape (ALARM, maxParentSet=2, BIC) on 2,000 samples, a short chain against the oracle"""
    from catnet_b200 import synth
    net = synth.load_alarm()
    s = synth.sample_network(net, 2000, np.random.Generator(np.random.PCG64(4004)))
    sa = dict(select_mode=2, temp_start=1e-2, temp_cool_fact=0.9, temp_check_orders=4, max_iter=12, order_shuffles=4,
              stop_diff=1e-2, num_threads=2, seed=4004)
    start = np.arange(1, 38)
    r = abi.search_sa(s, dict(sa, start_order=start), max_parent_set=2, max_complexity=600, num_cats=net["cats"])
    o = port.search_sa(s, 2, 600, start, num_cats=net["cats"], **sa)
    assert np.array_equal(r["order"], o["order"]) and r["niter"] == o["niter"] and r["naccept"] == o["naccept"]
    assert np.array_equal(r["trace"], o["trace"]) and np.array_equal(r["trace_accept"], o["trace_accept"])
    assert compare_search(r["nets"], o["nets"], check_probs=False) == len(o["nets"])


def test_sharded_search_equals_unsharded(abi):
    """the N>1 path on one GPU: every rank's shard scored into its segment of the rank-major buffer (here by one
    process, segment after segment), then selection + DP on the full buffer -> identical networks"""
    import ctypes
    import torch
    s, c = random_problem(5, 14, 300, None, 0.03)
    kw = dict(max_parent_set=3, max_complexity=2000)
    with abi.Context(0) as ctx:
        ctx.set_samples(s, num_cats=c)
        ref_nets = ctx.search_order(**kw)
        for world in (2, 3, 8):
            seg, nf = ctx.plan(0, world, **kw)
            buf = torch.full((seg * world,), float("nan"), dtype=torch.float64, device="cuda")
            for r in range(world):
                ctx.plan(r, world, **kw)
                ctx.score_shard(buf.data_ptr())
            torch.cuda.synchronize()
            nets = ctx.finish(buf.data_ptr())
            assert compare_search(nets, ref_nets) == len(ref_nets)


@pytest.mark.parametrize("n,m,cats,slices", [(70, 9000, 4, "1,3,4,4,4"), (40, 8192, 3, "1"), (130, 20001, 4, "1,1"), (65, 12345, 2, "1,2,3,4,5,6"),
                                             (30, 100000, 4, "1,3,4,4,4"), (200, 8500, 4, "2,1")])
def test_streamed_one_shot_search(abi, monkeypatch, n, m, cats, slices):
    """cnb_search_order on a large complete matrix uploads the samples in slices and contracts every slice while the next
    one is on its way (pair tables and two-parent tables accumulate over the slices): same networks, tables and
    log-likelihoods as the resident path, whatever the slicing"""
    s, c = random_problem(13 * n + m, n, m, cats)
    kw = dict(max_parent_set=2, max_complexity=n * 16 * 3, num_cats=c, order=(np.random.default_rng(m).permutation(n) + 1))
    monkeypatch.setenv("CNB_STREAM_MIN", "off")
    a = abi.search_order(s, **kw)
    assert abi.last_call_timing()["fast_path"] != 2
    monkeypatch.setenv("CNB_STREAM_MIN", "0")
    monkeypatch.setenv("CNB_STREAM_SLICES", slices)
    b = abi.search_order(s, **kw)
    assert abi.last_call_timing()["fast_path"] == 2 and abi.last_call_timing()["tensor_tiles"] > 0
    assert compare_search(b, a) == len(a)
    for x, y in zip(a, b):
        assert x["loglik"] == y["loglik"] and all(np.array_equal(p, q) for p, q in zip(x["probs"], y["probs"]))
    # three parents: the groups behind the streamed one run on the resident planes
    kw3 = dict(kw, max_parent_set=3, max_complexity=n * 64 * 3)
    if n <= 70:
        monkeypatch.setenv("CNB_STREAM_MIN", "off")
        a3 = abi.search_order(s, want_probs=False, **kw3)
        monkeypatch.setenv("CNB_STREAM_MIN", "0")
        b3 = abi.search_order(s, want_probs=False, **kw3)
        assert abi.last_call_timing()["fast_path"] == 2 and compare_search(b3, a3, check_probs=False) == len(a3)


def test_streamed_search_steps_aside(abi, monkeypatch):
    """a missing value, a value outside its node's categories, perturbations, pools: the streamed form hands over to the
    resident path (which scores the data with the kernels that take them, or reports the error)"""
    monkeypatch.setenv("CNB_STREAM_MIN", "0")
    s, c = random_problem(77, 40, 9000, 4)
    kw = dict(max_parent_set=2, max_complexity=40 * 48, num_cats=c)
    ok = abi.search_order(s, **kw)
    assert abi.last_call_timing()["fast_path"] == 2
    s_na = s.copy()
    s_na[8000, 7] = NA                                      # found while packing the LAST slice
    with abi.Context(0) as ctx:
        ctx.set_samples(s_na, num_cats=c)
        want = ctx.search_order(max_parent_set=2, max_complexity=40 * 48)
    got = abi.search_order(s_na, **kw)
    assert abi.last_call_timing()["fast_path"] != 2 and compare_search(got, want) == len(want)
    bad = s.copy()
    bad[8100, 3] = 9
    with pytest.raises(abi.CnbError) as e:
        abi.search_order(bad, **kw)
    assert "Incorrect categories" in str(e.value)
    pert = np.zeros_like(s)
    pert[::3, 5] = 1
    abi.search_order(s, perturbations=pert, **kw)
    assert abi.last_call_timing()["fast_path"] != 2
    again = abi.search_order(s, **kw)                       # and the parked context streams again afterwards
    assert abi.last_call_timing()["fast_path"] == 2 and compare_search(again, ok) == len(ok)


def test_errors_are_codes_not_crashes(abi):
    s, c = random_problem(1, 5, 50, 3)
    bad = s.copy()
    bad[3, 2] = 9                                      # category outside 1..C with nodeCats given
    with pytest.raises(abi.CnbError) as e:
        abi.search_order(bad, max_parent_set=2, max_complexity=100, num_cats=c)
    assert e.value.code == abi.CNB_ERR_PARAM and "Incorrect categories" in str(e.value)
    with pytest.raises(abi.CnbError) as e:
        abi.search_order(s, max_parent_set=2, max_complexity=100, num_cats=c, order=[1, 2, 3, 4, 9])
    assert "Invalid nodeOrder" in str(e.value)
    with pytest.raises(abi.CnbError):
        abi.search_order(s, max_parent_set=2, max_complexity=100, num_cats=c, device=99)

# ---------------------------------------------------------------- the reference's score cache (use_cache, what R passes)
def _strict_nets(a, b):
    x, y = {n["complexity"]: n for n in a}, {n["complexity"]: n for n in b}
    assert sorted(x) == sorted(y)
    for cx in x:
        u, v = x[cx], y[cx]
        assert u["parents"] == v["parents"], (cx, u["parents"], v["parents"])
        assert u["loglik"] == v["loglik"], (cx, u["loglik"], v["loglik"])
        assert list(u["node_loglik"]) == list(v["node_loglik"]) and list(u["node_prior"]) == list(v["node_prior"]), cx
        assert u["sample_sizes"] == v["sample_sizes"]
        if len(u["probs"]) and len(v["probs"]):
            assert all(np.array_equal(p, q) for p, q in zip(u["probs"], v["probs"])), cx


@pytest.mark.parametrize("seed", __import__("cases").SA_CACHE_SEEDS)
def test_sa_with_score_cache_matches_reference(abi, seed):
    """cnSearchSA / cnSearchHist with use_cache against the COMPILED REFERENCE run with its cache ON (one thread): the
    trajectory, the parents in the sequence the reference lists them in (that of the order which first scored the family),
    the log-likelihoods summed in that sequence, the histogram weighted by them (csrc/cnb_cache.h)"""
    from cases import sa_cache_problem
    s, c, P, K, start, sa, kw, _ = sa_cache_problem(seed)
    g = load_golden("ref_search_sa_cache")["seed%d" % seed]
    r = abi.search_sa(s, dict(sa, num_threads=1, start_order=start, use_cache=True), max_parent_set=P, max_complexity=K,
                      num_cats=c, **kw)
    assert [int(v) for v in r["order"]] == g["order"]
    assert (r["niter"], r["naccept"]) == (g["niter"], g["naccept"])
    assert [float(v) for v in r["trace"]] == [float.fromhex(v) for v in g["trace"]]
    assert [int(v) for v in r["trace_accept"]] == g["trace_accept"]
    compare_golden(r["nets"], g["nets"])
    hk = dict(score=seed % 3, weight=1 + seed % 2, max_iter=10, num_threads=1, seed=sa["seed"], use_cache=True)
    kh = {k: v for k, v in kw.items() if k != "edge_prob"}
    h, it = abi.search_hist(s, hk, max_parent_set=2, max_complexity=K, num_cats=c, **kh)
    assert it == g["hist_iterations"]
    assert h.tolist() == [[float.fromhex(v) for v in row] for row in g["hist"]]


@pytest.mark.parametrize("seed", [s for s in __import__("cases").SA_CACHE_SEEDS if s % 3] + list(range(30, 48)))
def test_sa_with_score_cache_and_threads_vs_oracle(abi, port, seed):
    """several proposals per step pass through the cache one after the other (the reference's threads store as they come:
    its own result depends on timing there) -- against the restatement, which takes them in the same sequence; tables and
    priors included"""
    from cases import sa_cache_problem
    s, c, P, K, start, sa, kw, _ = sa_cache_problem(seed)
    r = abi.search_sa(s, dict(sa, start_order=start, use_cache=True), want_probs=True, max_parent_set=P, max_complexity=K,
                      num_cats=c, **kw)
    o = port.search_sa(s, P, K, start, num_cats=c, use_cache=True, want_probs=True, **sa, **kw)
    assert np.array_equal(r["order"], o["order"]) and r["niter"] == o["niter"] and r["naccept"] == o["naccept"]
    assert np.array_equal(r["trace"], o["trace"]) and np.array_equal(r["trace_accept"], o["trace_accept"])
    _strict_nets(r["nets"], o["nets"])
    hk = dict(score=seed % 3, weight=seed % 3, max_iter=9, num_threads=1 + seed % 4, seed=sa["seed"])
    kh = {k: v for k, v in kw.items() if k != "edge_prob"}
    h, it = abi.search_hist(s, dict(hk, use_cache=True), max_parent_set=2, max_complexity=K, num_cats=c, **kh)
    h2, it2 = port.search_hist(s, 2, K, use_cache=True, num_cats=c, **hk, **kh)
    assert it == it2 and np.array_equal(h, h2)


def test_sa_score_cache_switch(abi, monkeypatch):
    """use_cache = 0 is the cache-off result; on the fixtures that differ, use_cache = 1 is not"""
    from cases import sa_cache_problem
    g = load_golden("ref_search_sa_cache")
    seed = next(s for s in __import__("cases").SA_CACHE_SEEDS if g["seed%d" % s]["differs_from_cache_off"])
    s, c, P, K, start, sa, kw, _ = sa_cache_problem(seed)
    args = dict(max_parent_set=P, max_complexity=K, num_cats=c, **kw)
    on = abi.search_sa(s, dict(sa, num_threads=1, start_order=start, use_cache=True), want_probs=True, **args)
    off = abi.search_sa(s, dict(sa, num_threads=1, start_order=start, use_cache=False), want_probs=True, **args)
    differs = (not np.array_equal(on["trace"], off["trace"]) or [n["parents"] for n in on["nets"]] != [n["parents"] for n in off["nets"]]
               or [n["loglik"] for n in on["nets"]] != [n["loglik"] for n in off["nets"]])
    assert differs


@pytest.mark.parametrize("m,cats,na", [(300, None, 0.05), (8300, None, 0.0), (8300, 4, 0.0), (8300, 3, 0.0), (9000, None, 0.02)])
def test_family_scores_follow_the_listing(abi, port, m, cats, na):
    """an explicit family is counted and summed in the sequence its parents are LISTED in (first listed parent most
    significant), whichever kernel scores it: what the score-cache overrides and cnSetProb rely on -- every permutation
    of 2 and 3 parents, small data (bit planes) and enough complete samples for the pair / three-way tables"""
    import itertools
    s, c = random_problem(77, 8, m, cats, na)
    s0 = np.where(s < 1, 99, s - 1).astype(np.int32)
    with abi.Context(0) as ctx:
        ctx.set_samples(s, num_cats=c)
        for d in (1, 2, 3):
            fams = []
            for child in (7, 3, 5):
                base = [q for q in range(8) if q != child][:d + 1][-d:]
                for perm in itertools.permutations(base):
                    fams.append((child, list(perm)))
            counts, ll, nc = ctx.family_scores([f[0] for f in fams], np.array([f[1] for f in fams], dtype=np.int32))
            for i, (child, pars) in enumerate(fams):
                rc, rll, rnc = port.family_score(s0, c, child, pars)
                assert np.array_equal(counts[i, :len(rc)], rc.astype(np.int64)) and ll[i] == rll and nc[i] == rnc, (d, child, pars)

@pytest.mark.parametrize("seed", [2, 5, 7, 101])
def test_sa_score_cache_pipelined_equals_sequential(abi, monkeypatch, seed):
    """use_cache with several proposals per step: option tables side by side, then the walk through the cache model and the
    DPs one behind the other (the default) -- the same chain as evaluating the proposals strictly one after the other"""
    from cases import sa_cache_problem
    s, c, P, K, start, sa, kw, _ = sa_cache_problem(seed)
    args = dict(max_parent_set=P, max_complexity=K, num_cats=c, **kw)
    sa = dict(sa, num_threads=3, start_order=start, use_cache=True)
    a = abi.search_sa(s, sa, want_probs=True, **args)
    monkeypatch.setenv("CNB_CACHE_NO_PIPELINE", "1")
    b = abi.search_sa(s, sa, want_probs=True, **args)
    assert np.array_equal(a["trace"], b["trace"]) and np.array_equal(a["order"], b["order"]) and a["niter"] == b["niter"]
    _strict_nets(a["nets"], b["nets"])
