"""One process per GPU: shard the candidate parent sets over the ranks, all-gather the per-family scores
(NCCL over NVLink on the GPU box, gloo in the CPU tests), run selection + DP redundantly on every rank.

Score buffer layout (cnb_ctx_plan / dev_score_index in csrc/cnb_device.cuh): rank-major.  Each group g (= all
candidate sets with the same parent count) of nfam_g families is cut into `world` contiguous chunks of
chunk_g = ceil(nfam_g / world); rank r's segment holds its chunk of every group back to back (seg_len = sum chunk_g),
so one all_gather_into_tensor of equal-sized segments completes the buffer.  The two-parent group, when it runs on
the tensor cores, is dealt tile by tile instead (tile t -> rank t % world, csrc/cnb_tiles.h): its chunk is
ceil(ntiles / world) tiles of 28 x 64 score slots.  cnSearchSA runs one independent chain per rank and merges them
per complexity like R's .searchSA (R/catnet.search.R:317-332).

The sample matrix is replicated on every GPU.  When it starts in HOST memory (the R-facing call), every rank copies
only its 1/world of the samples over PCIe and the ranks complete each other's copy with one all-gather over NVLink
(set_samples_sharded): the host-to-device time divides by the number of GPUs instead of multiplying the PCIe load.
"""
import numpy as np


def segment_layout(group_sizes, world):
    """python mirror of the planner's layout: -> (chunks, seg_offsets, seg_len)"""
    chunks = [(int(n) + world - 1) // world for n in group_sizes]
    offs = np.concatenate([[0], np.cumsum(chunks)])[:-1].astype(np.int64)
    return chunks, [int(o) for o in offs], int(sum(chunks))


def score_index(group, fid, chunks, offs, seg_len):
    r = fid // chunks[group]
    return r * seg_len + offs[group] + (fid - r * chunks[group])


def owned_range(group_size, chunk, rank):
    b = rank * chunk
    return b, max(b, min(group_size, b + chunk))


def gather_scores(scores, rank, world, seg_len, group=None, scattered=False):
    """complete the rank-major buffer in place: every rank contributes scores[rank*seg_len:(rank+1)*seg_len].  Over NCCL the
    segment is gathered where it lies (ncclAllGather's in-place form: input = output + rank * count); gloo gets a copy.
    scattered (ctx.scattered() after score_shard, with ctx.allow_scatter(2)): the rank wrote scores into every segment of a
    buffer that started at -inf -- the three-parent T tiles deal triples, not children, to the ranks -- and the buffers are
    combined by an element-wise MAX instead."""
    import torch.distributed as dist
    if world == 1:
        return scores
    if scattered:
        dist.all_reduce(scores, op=dist.ReduceOp.MAX, group=group)
        return scores
    mine = scores[rank * seg_len:(rank + 1) * seg_len]
    if dist.get_backend(group) != "nccl":
        mine = mine.clone()
    dist.all_gather_into_tensor(scores, mine, group=group)
    return scores


def set_samples_sharded(ctx, host_samples, rank, world, device, num_cats=None, group=None):
    """host_samples: int32 [M][N] numpy array or CPU tensor (pageable memory is fine), identical on every rank.  Rank r
    narrows rows [r*ceil(M/world), (r+1)*ceil(M/world)) to bytes with the library's host threads and copies them over its
    own PCIe link (cnb_ctx_upload_rows8; 0 = missing, the form cnb_ctx_set_samples_dev8 takes), one in-place
    all_gather_into_tensor over NVLink completes the BYTE matrix on every GPU (a quarter of the int32 traffic), then the
    matrix is packed.  Returns (device byte tensor, bytes copied H2D by this rank)."""
    import os
    import torch
    import torch.distributed as dist
    ctx.set_stream(torch.cuda.current_stream(device).cuda_stream)   # the copy, the all-gather and the pack stay ordered
    hs = host_samples.numpy() if hasattr(host_samples, "numpy") else host_samples
    m, n = hs.shape
    rows = (m + world - 1) // world
    full = torch.empty((rows * world, n), dtype=torch.uint8, device=device)
    r0, r1 = min(m, rank * rows), min(m, (rank + 1) * rows)
    if r1 > r0:
        ctx.upload_rows8(hs, r0, r1, full.data_ptr(), max(1, (os.cpu_count() or 8) // world))
    if world > 1:
        dist.all_gather_into_tensor(full, full[rank * rows:(rank + 1) * rows], group=group)
    ctx.set_samples_dev8(full.data_ptr(), n, m, num_cats=num_cats)
    return full, (r1 - r0) * n


def search_order_sharded(ctx, rank, world, device, want_probs=True, group=None, **kw):
    """cnSearchOrder on `world` GPUs; returns the same networks on every rank"""
    import torch
    ctx.set_stream(torch.cuda.current_stream(device).cuda_stream)   # kernels and the all-gather on one stream
    ctx.allow_scatter(2)                                            # three-parent T tiles may be dealt to the ranks
    seg, _ = ctx.plan(rank, world, **kw)
    scores = torch.empty(max(seg * world, 1), dtype=torch.float64, device=device)
    ctx.score_shard(scores.data_ptr())
    gather_scores(scores, rank, world, seg, group, scattered=ctx.scattered())
    torch.cuda.synchronize(device)
    return ctx.finish(scores.data_ptr(), want_probs)


_side_streams = {}


def nccl_options():
    """pg_options for init_process_group("nccl", ...): NCCL's kernels on a high-priority stream.  The all-gather of slice k + 1
    has to run while the table kernel's pass over slice k still has thousands of CTAs waiting; at default priority the block
    scheduler starts a later kernel only once the earlier one has issued its last CTA, which would put every all-gather
    (and the packing behind it) between two passes instead of under one."""
    import torch.distributed as dist
    opts = dist.ProcessGroupNCCL.Options()
    opts.is_high_priority_stream = True
    return opts


def search_order_streamed(ctx, host_samples, rank, world, device, num_cats, want_probs=True, group=None, finish_rank=None,
                          raw=False, **kw):
    """cnSearchOrder on `world` GPUs from a HOST matrix (int32 [M][N], pageable is fine, identical on every rank), the upload
    overlapped with the scoring (the multi-process form of the streamed search, include/catnet_b200.h): the samples travel in
    slices; of every slice rank r narrows its 1/world share of the rows to bytes and copies it over its own PCIe link, an
    in-place all-gather over NVLink completes the slice on every GPU (both on a side stream), and the library packs the slice
    and runs one pass of the tensor-core table kernel over this rank's tiles while the host threads already narrow the next
    slice.  Problems the streamed form does not take (missing values, perturbations, pools, more than four categories, small
    matrices) go through set_samples_sharded + search_order_sharded.  finish_rank: only that rank runs selection, DP and
    export (the others return None); default every rank.  Returns (networks | result handle if raw | None, info)."""
    import os
    import torch
    import torch.distributed as dist
    stream = torch.cuda.current_stream(device)
    ctx.set_stream(stream.cuda_stream)
    hs = host_samples.numpy() if hasattr(host_samples, "numpy") else host_samples
    m, n = hs.shape
    seg, _, rows = ctx.stream_begin(rank, world, n, m, num_cats=num_cats, **kw)
    info = {"streamed": bool(rows), "slices": max(len(rows) - 1, 0), "h2d_bytes": 0}

    def finish(scores):
        if finish_rank is not None and rank != finish_rank:
            return None
        return ctx.finish_raw(scores.data_ptr()) if raw else ctx.finish(scores.data_ptr(), want_probs)

    if not rows:
        _, info["h2d_bytes"] = set_samples_sharded(ctx, hs, rank, world, device, num_cats=num_cats, group=group)
        seg, _ = ctx.plan(rank, world, **kw)
        scores = torch.empty(max(seg * world, 1), dtype=torch.float64, device=device)
        ctx.score_shard(scores.data_ptr())
        gather_scores(scores, rank, world, seg, group)
        torch.cuda.synchronize(device)
        return finish(scores), info
    ns = len(rows) - 1
    shares = [(rows[k + 1] - rows[k] + world - 1) // world for k in range(ns)]
    full = torch.empty((max(rows[k] + world * shares[k] for k in range(ns)), n), dtype=torch.uint8, device=device)
    scores = torch.empty(max(seg * world, 1), dtype=torch.float64, device=device)
    side = _side_streams.get(str(device))
    if side is None:
        side = _side_streams[str(device)] = torch.cuda.Stream(device, priority=-1)
    side.wait_stream(stream)                                   # `full` is allocated in the current stream's order
    cap, off = sum(shares) * n, 0
    threads = max(1, (os.cpu_count() or 8) // world)
    events = []
    for k in range(ns):
        r0, r1, sh = rows[k], rows[k + 1], shares[k]
        a, b = min(r1, r0 + rank * sh), min(r1, r0 + (rank + 1) * sh)
        if b > a:
            try:
                ctx.upload_rows8_on(hs, a, b, full.data_ptr(), threads, side.cuda_stream, off, cap)
            except Exception as e:                             # a label above 255: it travelled as 255 = "no category", which the
                if "above 255" not in str(e):                  # packing reports on EVERY rank (stream_end: not ok) -- no rank may
                    raise                                      # leave the collectives early
            off += (b - a) * n
            info["h2d_bytes"] += (b - a) * n
        with torch.cuda.stream(side):
            if world > 1:
                dist.all_gather_into_tensor(full[r0:r0 + world * sh], full[r0 + rank * sh:r0 + (rank + 1) * sh], group=group)
            ev = torch.cuda.Event()
            ev.record(side)
        events.append(ev)
        ctx.stream_slice(k, full.data_ptr(), ev.cuda_event)
    if not ctx.stream_end(scores.data_ptr()):
        info["streamed"] = False                               # missing values: the byte matrix is complete on every rank
        ctx.set_samples_dev8(full.data_ptr(), n, m, num_cats=num_cats)
        seg, _ = ctx.plan(rank, world, **kw)                   # other kernels, another layout of the score buffer
        scores = torch.empty(max(seg * world, 1), dtype=torch.float64, device=device)
        ctx.score_shard(scores.data_ptr())
    full.record_stream(side)
    gather_scores(scores, rank, world, seg, group)
    torch.cuda.synchronize(device)
    return finish(scores), info


def merge_evaluations(evals):
    """per complexity keep the network with the highest likelihood over the chains (first wins ties)"""
    best = {}
    for nets in evals:
        for net in nets or []:
            cur = best.get(net["complexity"])
            if cur is None or net["loglik"] > cur["loglik"]:
                best[net["complexity"]] = net
    return [best[k] for k in sorted(best)]


def search_sa_chains(ctx, rank, world, sa, group=None, **kw):
    """one independent SA chain per rank (seed + rank), merged on every rank"""
    import torch.distributed as dist
    mine = ctx.search_sa(dict(sa, seed=int(sa.get("seed", 1)) + rank), **kw)
    if world == 1:
        return mine, [mine]
    allr = [None] * world
    dist.all_gather_object(allr, {"nets": mine["nets"], "order": mine["order"], "niter": mine["niter"],
                                  "naccept": mine["naccept"]}, group=group)
    return {"nets": merge_evaluations([r["nets"] for r in allr]), "chains": allr}, allr
