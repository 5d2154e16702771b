"""GPU parity tests proper: the CUDA path, through the C ABI, against the reference's own C++ core (oracle/_ref).

Bars (BASELINE.json north_star): family counts bit-exact; family log-lik within 1e-9 relative in fp64 (the device
evaluates glibc's log algorithm, so the tests additionally report/require bit equality); cnSearchOrder selects
identical parent sets for every complexity level.
"""
import itertools
import math

import numpy as np
import pytest

from helpers import NA, compare_search, random_problem

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def abi():
    from catnet_b200 import _abi
    _abi.lib()
    return _abi


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
def test_family_counts_and_loglik(abi, ref, cats, na, pert, m):
    n = 10
    s, c = random_problem(7 + m, n, max(m, 5), cats, na)
    s = s[:m] if m >= 5 else s[:m]
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
                    sub = s0 if p is None else s0[p[:, ch] == 0]
                    if len(sub) == 0:
                        continue          # the reference refuses nsamples < 1 (catnet_class.h:824)
                    rc, rll, rnc = ref.family_score(sub, c, ch, list(pa))
                    assert np.array_equal(counts[i, :len(rc)], rc.astype(np.int64)), (d, generic, ch, pa)
                    assert nc[i] == rnc
                    assert ll[i] == rll, (d, generic, ch, pa, ll[i], rll)   # bit-exact


CASES = [
    dict(seed=1, n=8, m=200, cats=None, P=2),
    dict(seed=2, n=9, m=120, cats=3, P=3),
    dict(seed=3, n=10, m=64, cats=2, P=3),
    dict(seed=4, n=8, m=300, cats=None, P=2, na=0.07),
    dict(seed=5, n=7, m=90, cats=4, P=2, na=0.05, pert=True),
    dict(seed=6, n=12, m=40, cats=2, P=2, K=30),
    dict(seed=7, n=9, m=500, cats=None, P=3, order="random"),
    dict(seed=8, n=9, m=150, cats=3, P=2, pools=True),
    dict(seed=9, n=9, m=150, cats=None, P=3, fixed=True),
    dict(seed=10, n=8, m=150, cats=3, P=2, prior=True),
    dict(seed=11, n=8, m=150, cats=None, P=2, prior=True, fixed=True, pools=True, psizes=True),
    dict(seed=12, n=6, m=10, cats=2, P=2),
    dict(seed=13, n=5, m=2, cats=2, P=2),
    dict(seed=14, n=1, m=20, cats=3, P=2),
    dict(seed=15, n=9, m=100, cats=None, P=2, infer=True, na=0.05),
    dict(seed=16, n=8, m=100, cats=5, P=2),            # > 4 categories: generic histogram kernel
    dict(seed=17, n=7, m=100, cats=[2, 6, 3, 2, 7, 3, 2], P=2),
    dict(seed=18, n=8, m=100, cats=4, P=3),            # d=3 with 4 categories: generic kernel for that group
]


def build_case(c):
    s, cats = random_problem(c["seed"], c["n"], c["m"], c.get("cats"), c.get("na", 0.0))
    n = c["n"]
    rng = np.random.default_rng(1000 + c["seed"])
    kw = dict(max_parent_set=c["P"])
    kw["max_complexity"] = c.get("K", int(n * (int(cats.max()) ** c["P"]) * (int(cats.max()) - 1)))
    kw["order"] = (rng.permutation(n) + 1).astype(np.int32) if c.get("order") == "random" else None
    if not c.get("infer"):
        kw["num_cats"] = cats
    if c.get("pert"):
        kw["perturbations"] = (rng.random(s.shape) < 0.15).astype(np.int32)
    if c.get("pools"):
        kw["parents_pool"] = [list(rng.choice(np.arange(1, n + 1), size=rng.integers(0, n), replace=False)) for _ in range(n)]
    if c.get("fixed"):
        kw["fixed_parents"] = [list(rng.choice(np.arange(1, n + 1), size=rng.integers(0, 3), replace=False)) for _ in range(n)]
    if c.get("prior"):
        e = rng.random((n, n))
        e[rng.random((n, n)) < 0.1] = 0.0
        e[rng.random((n, n)) < 0.05] = 1.0
        kw["edge_prob"] = e
    if c.get("psizes"):
        kw["parent_sizes"] = rng.integers(0, c["P"] + 1, size=n).astype(np.int32)
    return s, kw


@pytest.mark.parametrize("case", CASES, ids=lambda c: "seed%d" % c["seed"])
def test_search_order_matches_reference(abi, ref, case):
    s, kw = build_case(case)
    theirs = ref.search_order(s, kw["max_parent_set"], kw["max_complexity"], order=kw["order"],
                              perturbations=kw.get("perturbations"), parent_sizes=kw.get("parent_sizes"),
                              num_cats=kw.get("num_cats"), parents_pool=kw.get("parents_pool"),
                              fixed_parents=kw.get("fixed_parents"), edge_prob=kw.get("edge_prob"))
    ours = abi.search_order(s, **kw)
    exact = compare_search(ours, theirs)
    assert exact == len(theirs), "log-liks not bit-identical: %d of %d" % (exact, len(theirs))
    # same through the generic histogram kernel
    with abi.Context(0) as ctx:
        ctx.set_samples(s, perturbations=kw.get("perturbations"), num_cats=kw.get("num_cats"))
        ctx.force_generic(True)
        kw2 = {k: v for k, v in kw.items() if k not in ("perturbations", "num_cats")}
        ours2 = ctx.search_order(**kw2)
    assert compare_search(ours2, theirs) == len(theirs)


def test_config_c1_alarm(abi, ref):
    """BASELINE configs[0]: cnSearchOrder on cnSamples(alarmnet, 1000), true order, maxParentSet=2 (mixed categories:
    the per-candidate DP path, catnet_search2.h:681-841)"""
    from catnet_b200 import synth
    c = synth.config("C1")
    order = c["order"] + 1
    theirs = ref.search_order(c["samples"], 2, c["max_complexity"], order=order, num_cats=c["cats"])
    ours = abi.search_order(c["samples"], max_parent_set=2, max_complexity=c["max_complexity"], order=order,
                            num_cats=c["cats"])
    assert len(theirs) > 500
    assert compare_search(ours, theirs) == len(theirs)


def test_config_c2_breast(abi, ref):
    """BASELINE configs[1]: breast, 100 genes x 98 samples, maxParentSet=3, full complexity (4,087,875 families)"""
    from catnet_b200 import synth
    c = synth.config("C2")
    theirs = ref.search_order(c["samples"], 3, c["max_complexity"], num_cats=c["cats"], want_probs=False)
    ours = abi.search_order(c["samples"], max_parent_set=3, max_complexity=c["max_complexity"], num_cats=c["cats"],
                            want_probs=False)
    assert compare_search(ours, theirs, check_probs=False) == len(theirs)


def test_config_c3_reduced(abi, ref):
    """BASELINE configs[2] at reduced M (the full-M reference run is ~1.6 h): 100 nodes x 3 cats, maxParentSet=3"""
    from catnet_b200 import synth
    c = synth.config("C3", m=300)
    order = c["order"] + 1
    theirs = ref.search_order(c["samples"], 3, c["max_complexity"], order=order, num_cats=c["cats"], want_probs=False)
    ours = abi.search_order(c["samples"], max_parent_set=3, max_complexity=c["max_complexity"], order=order,
                            num_cats=c["cats"], want_probs=False)
    assert compare_search(ours, theirs, check_probs=False) == len(theirs)
