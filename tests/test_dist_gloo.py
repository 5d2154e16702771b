"""CPU suite, part 3: the N>1 path on world_size-2 gloo -- shard layout, all-gather of the score segments, merge of
independent SA chains.  The scores themselves come from the oracle here (a GPU is not available); what is under
test is the exchange logic of catnet_b200/dist.py, which is what runs over NCCL on the GPU box."""
import itertools
import math
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from catnet_b200 import dist as cdist


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import sys
        sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__))))
        from helpers import random_problem
        from oracle import port as orc
        n, P = 7, 2
        s, c = random_problem(3, n, 120, 3)
        s0 = (s - 1).astype(np.int32)
        # candidate families in the planner's order: groups by parent count, child ascending, lexicographic subsets
        groups = [[(ch, pa) for ch in range(n) for pa in itertools.combinations(range(ch), d)] for d in (1, 2)]
        sizes = [len(g) for g in groups]
        chunks, offs, seg = cdist.segment_layout(sizes, world)
        buf = torch.full((seg * world,), float("nan"), dtype=torch.float64)
        for g, fams in enumerate(groups):
            b, e = cdist.owned_range(sizes[g], chunks[g], rank)
            for fid in range(b, e):
                ch, pa = fams[fid]
                buf[cdist.score_index(g, fid, chunks, offs, seg)] = orc.family_score(s0, c, ch, list(pa))[1]
        own = buf[rank * seg:(rank + 1) * seg]
        assert torch.isnan(own).sum() == sum(chunks) - sum(cdist.owned_range(sizes[g], chunks[g], rank)[1]
                                                        - cdist.owned_range(sizes[g], chunks[g], rank)[0] for g in range(2))
        cdist.gather_scores(buf, rank, world, seg)
        # every family's score is now present at its slot, on every rank
        for g, fams in enumerate(groups):
            for fid, (ch, pa) in enumerate(fams):
                v = buf[cdist.score_index(g, fid, chunks, offs, seg)].item()
                assert v == orc.family_score(s0, c, ch, list(pa))[1]
        # SA chains: per-complexity merge keeps the better network
        mine = [{"complexity": 10, "loglik": -5.0 - rank}, {"complexity": 11 + rank, "loglik": -4.0}]
        allr = [None] * world
        dist.all_gather_object(allr, mine)
        merged = cdist.merge_evaluations(allr)
        assert [m["complexity"] for m in merged] == [10, 11, 12] and merged[0]["loglik"] == -5.0
        out.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        out.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def test_world2_gloo_exchange():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    res = [out.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(60)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res


def test_layout_matches_planner():
    """segment_layout is the Python mirror of cnb_build_plan's chunking"""
    from catnet_b200 import _abi
    for n, p, world in [(37, 2, 8), (12, 3, 3), (500, 2, 8)]:
        sizes = [sum(math.comb(i, d) for i in range(n)) for d in range(1, p + 1)]
        chunks, offs, seg = cdist.segment_layout(sizes, world)
        assert seg == sum(chunks) and all(c * world >= s for c, s in zip(chunks, sizes))
        # all slots distinct
        idx = {cdist.score_index(g, f, chunks, offs, seg) for g in range(len(sizes)) for f in range(0, sizes[g], max(1, sizes[g] // 50))}
        assert max(idx) < seg * world
        keep = _abi.Keep()
        prm = _abi.make_params(keep, n, 10, max_parent_set=p, max_complexity=1 << 24)
        assert _abi.lib().cnb_num_families(prm) == sum(sizes)
