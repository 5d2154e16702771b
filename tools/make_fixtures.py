"""Generate the committed fixtures under tests/golden/ from the reference's data files.

Run in the build container (where /root/reference exists):  python tools/make_fixtures.py
  tests/golden/alarmnet.json       ALARM network (37 nodes) from data/alarmnet.rda: categories, parents, CPTs
  tests/golden/breast_100x98.npy   breast data through the demo/breast.r:6-16 pipeline: 100 genes x 98 samples, 2 cats
The GPU box has no /root/reference; bench.py and the tests read only these fixtures.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))
from rda import read_rda  # noqa: E402

from catnet_b200.synth import discretize_uniform  # noqa: E402

REF = os.environ.get("CATNET_REFERENCE", "/root/reference")
OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


def flatten(x):
    """nested probability lists (first listed parent outermost) -> flat row-major list"""
    if isinstance(x, list) and x and isinstance(x[0], list):
        return [v for sub in x for v in flatten(sub)]
    return list(x)


def alarm():
    net = read_rda(os.path.join(REF, "data", "alarmnet.rda"))["alarmnet"].attrs
    cats = [len(c) for c in net["categories"]]
    parents = [[] if p is None else [int(v) - 1 for v in p] for p in net["parents"]]
    cpts = [flatten(p) for p in net["probabilities"]]
    for i, c in enumerate(cpts):
        rows = int(np.prod([cats[p] for p in parents[i]])) if parents[i] else 1
        assert len(c) == rows * cats[i], (i, len(c), rows, cats[i])
    assert sum((cats[i] - 1) * int(np.prod([cats[p] for p in parents[i]])) for i in range(37)) == net["complexity"][0]
    with open(os.path.join(OUT, "alarmnet.json"), "w") as f:
        json.dump({"source": "catnet data/alarmnet.rda", "nodes": net["nodes"], "cats": cats,
                   "parents": parents, "cpts": cpts}, f)
    print("alarmnet.json: 37 nodes, complexity", net["complexity"][0])


def breast():
    df = read_rda(os.path.join(REF, "data", "breast.rda"))["breast"]
    cols = []
    for c in df.value:                       # 100 columns, each a factor of 1214 rows
        lv = c.attrs["levels"]
        cols.append([lv[k - 1] for k in c.value])
    # drop row 1 and cols 1-2; genes are rows of the frame, samples are columns 3..100
    x = np.array([[float(v) for v in col[1:]] for col in cols[2:]], dtype=np.float64)   # [98 samples][1213 genes]
    x = x.T                                                                                # [1213 genes][98]
    d = discretize_uniform(x, 2)
    subset = np.arange(2, 1202, 12) - 1                                                   # seq(2,1201,12), 1-based
    d = d[subset]
    assert d.shape == (100, 98) and d.min() == 1 and d.max() == 2
    np.save(os.path.join(OUT, "breast_100x98.npy"), d.astype(np.int8))
    print("breast_100x98.npy:", d.shape, "ones", int((d == 1).sum()))


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    alarm()
    breast()
