"""Multi-GPU host logic (one process per GPU, torch.distributed for the plumbing).

Two ways the path shards:
  * restarts -- independent EM runs (scSplit:473-484); no data-path collective, only the final "first strictly
    greatest LL wins" selection gathers R doubles.  Preferred when one iteration is shorter than an all-reduce.
  * cells    -- every rank holds a contiguous slice of the barcodes (both entry streams restricted to its cells);
    the E-step, posteriors, assignment and refine scores are local; the M-step statistics are packed as
        [ altW (V x KP) | refW (V x KP) | column mass (KP) | LL | pad ]          (KP = K rounded up to even)
    and summed over ranks once per iteration by ncclAllReduce inside libscsplit_b200.so (csrc/comm.cu), issued on the
    library's stream between the M-step SpMM and k_finalize.  This module only plans the shards and carries the
    128-byte NCCL unique id from rank 0 to the others.
"""
import numpy as np


def shard_ranges(n, world, weights=None):
    """Contiguous [lo, hi) ranges over n items for `world` ranks; balanced by count, or by `weights` (e.g. reads per cell)."""
    if weights is None:
        base, extra = divmod(n, world)
        out, lo = [], 0
        for r in range(world):
            hi = lo + base + (1 if r < extra else 0)
            out.append((lo, hi))
            lo = hi
        return out
    w = np.asarray(weights, dtype=np.float64)
    assert w.shape == (n,)
    cum = np.concatenate([[0.0], np.cumsum(w)])
    cuts = [0]
    for r in range(1, world):
        cuts.append(int(np.searchsorted(cum, cum[-1] * r / world, side="left")))
    cuts.append(n)
    cuts = np.maximum.accumulate(np.minimum(cuts, n))
    return [(int(cuts[r]), int(cuts[r + 1])) for r in range(world)]


def shard_restarts(R, world):
    """Restart indices of every rank (round robin keeps early restarts -- the tie-break winners -- spread out)."""
    return [list(range(r, R, world)) for r in range(world)]


def select_best(ll_by_restart):
    """run.core's rule (scSplit:482-484): scan restarts in order, keep the first STRICTLY greater LL; None if none valid."""
    best, best_ll = None, -1e50
    for i, ll in enumerate(ll_by_restart):
        if ll is not None and not np.isnan(ll) and ll > best_ll:
            best, best_ll = i, ll
    return best


def stats_layout(V, K):
    KP = K + (K & 1)
    n = 2 * V * KP
    return dict(KP=KP, alt=(0, V * KP), ref=(V * KP, n), colsum=(n, n + KP), ll=n + KP, size=n + KP + 2)


def pack_stats(altW, refW, colsum, ll):
    """Host mirror of the device buffer that is all-reduced per iteration (for tests and for documentation)."""
    V, K = altW.shape
    lay = stats_layout(V, K)
    KP = lay["KP"]
    buf = np.zeros(lay["size"], dtype=np.float64)
    a = np.zeros((V, KP)); a[:, :K] = altW
    r = np.zeros((V, KP)); r[:, :K] = refW
    buf[lay["alt"][0]:lay["alt"][1]] = a.ravel()
    buf[lay["ref"][0]:lay["ref"][1]] = r.ravel()
    buf[lay["colsum"][0]:lay["colsum"][0] + K] = colsum
    buf[lay["ll"]] = ll
    return buf


def unpack_stats(buf, V, K):
    lay = stats_layout(V, K)
    KP = lay["KP"]
    a = buf[lay["alt"][0]:lay["alt"][1]].reshape(V, KP)[:, :K]
    r = buf[lay["ref"][0]:lay["ref"][1]].reshape(V, KP)[:, :K]
    return a, r, buf[lay["colsum"][0]:lay["colsum"][0] + K], float(buf[lay["ll"]])


def finalize_from_stats(altW, refW, colsum, k_alt):
    """What every rank computes redundantly after the all-reduce (k_finalize and the prior update in its last block): scSplit:196,198."""
    with np.errstate(divide="ignore", invalid="ignore"):
        P = (altW + k_alt[:, None]) / ((altW + refW) + 1)
        lPs = np.log2(colsum) - np.log2(colsum.sum())
    return P, lPs


def broadcast_unique_id(make_id, rank, src=0, group=None, device=None):
    """Rank `src` calls make_id() -> uint8[128] (Engine.comm_unique_id); everyone returns the same bytes."""
    import torch
    import torch.distributed as dist
    t = torch.zeros(128, dtype=torch.uint8, device=device or "cpu")
    if rank == src:
        t = torch.from_numpy(np.ascontiguousarray(make_id(), dtype=np.uint8)).to(t.device)
    dist.broadcast(t, src, group=group)
    return t.cpu().numpy()


def open_cell_shard(ref, alt, rank, world, local_device, weights=None):
    """Engine over this rank's slice of the cells, joined to the NCCL communicator. Returns (engine, (lo, hi))."""
    from .engine import Engine
    lo, hi = shard_ranges(ref.shape[1], world, weights)[rank]
    eng = Engine(ref[:, lo:hi], alt[:, lo:hi], device=local_device)
    if world > 1:
        uid = broadcast_unique_id(Engine.comm_unique_id, rank, device="cuda:%d" % local_device)
        eng.comm_init(uid, rank, world)
    return eng, (lo, hi)
