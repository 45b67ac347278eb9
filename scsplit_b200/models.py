"""Host-side mirror of the reference's `models` class (scSplit:64-336) on top of the CUDA engine.

Same constructor signature, method names, attribute names and attribute types (pandas frames at the seam) as the
reference, so `run.core` / `run` (scSplit:470-600) can drive it unchanged; all sparse arithmetic runs in
libscsplit_b200.so.  `core_batched` is the restart driver (scSplit:470-490) with the restarts batched on the GPU.

Quirks of the reference that are kept on purpose (SURVEY.md sections 3.1, 4, 5.1):
  * E then M inside an iteration: on exit P_s_c / lP_c_s / lP_c_m belong to the last E-step, model_af / lP_s are one
    M-step ahead (scSplit:159-161);
  * assignment is a >= 0.99 threshold, not an argmax (scSplit:207);
  * `reassigned = assigned.copy()` is shallow, so `reassigned[dbl] += found` also grows `assigned[dbl]` (scSplit:234,252);
  * SNV ids whose FIRST CHARACTER is X, Y or M are dropped from the P/A matrix (scSplit:283);
  * `dist_variants` is `list(set(...))` (scSplit:334): order is arbitrary.
"""
import datetime
import os

import numpy as np
import pandas as pd

from . import initialiser
from ._lib import SCS_L, SCS_LPS, SCS_P, SCS_W
from .engine import Engine

_ENGINES = {}


def get_engine(base_calls_mtx, device=0):
    """One device-resident copy of the count matrices per (ref, alt) pair, shared by every restart's `models`.

    The cache entry holds the two matrices themselves, so their ids cannot be recycled by another pool while the
    engine lives, and a hit is re-checked against shape and nnz.
    """
    ref, alt = base_calls_mtx[0], base_calls_mtx[1]
    key = (id(ref), id(alt), device)
    hit = _ENGINES.get(key)
    if hit is not None:
        eng, kref, kalt, sig = hit
        if eng.h is not None and kref is ref and kalt is alt and sig == (ref.shape, ref.nnz, alt.shape, alt.nnz):
            return eng
    for k in [k for k in _ENGINES if k[2] == device]:
        _ENGINES.pop(k)[0].close()
    eng = Engine(ref, alt, device=device)
    _ENGINES[key] = (eng, ref, alt, (ref.shape, ref.nnz, alt.shape, alt.nnz))
    return eng


def release_engines():
    for k in list(_ENGINES):
        _ENGINES.pop(k)[0].close()


def _log(output, text):
    with open(os.path.join(output, r'scSplit.log'), 'a') as logfile:
        logfile.write(text)


class models:
    """Model state of one restart: SNVs, barcodes, P(s|c), log-likelihoods, allele-fraction model, assignments."""

    def __init__(self, base_calls_mtx, num, output, device=0, _defer_seed=False, _trim=None):
        self.ref_bc_mtx, self.alt_bc_mtx = base_calls_mtx[0], base_calls_mtx[1]
        self.all_POS, self.barcodes = base_calls_mtx[2].tolist(), base_calls_mtx[3].tolist()
        self.num = num
        self.engine = get_engine(base_calls_mtx, device)
        self.P_s_c = pd.DataFrame(0.0, index=self.barcodes, columns=range(self.num))
        self.lP_c_s = pd.DataFrame(0.0, index=self.barcodes, columns=range(self.num))
        self.lP_s = [np.log2(1 / self.num)] * self.num
        self.assigned, self.reassigned = [], []
        self.convergence = 0
        for _ in range(self.num):
            self.assigned.append([])
            self.reassigned.append([])
        self.model_af = pd.DataFrame(0.0, index=self.all_POS, columns=range(self.num))
        self.pseudo = 1
        # background alt fraction (scSplit:100-103); (V,1) matrix like the reference's attribute
        self.k_alt = np.asmatrix(self.engine.kalt()).T
        # subset / PCA / KMeans initialisation: same np.random and argsort call sequence as scSplit:106-125
        # (core_batched runs the trimming loops of all restarts first and hands each model its own: _trim)
        irows, icols, nrows, ncols = _trim if _trim is not None else initialiser.trim_subset(
            self.engine.pattern_counts, len(self.all_POS), len(self.barcodes), self.num)
        self._labels = initialiser.initial_labels(self.ref_bc_mtx, self.alt_bc_mtx, irows, icols, nrows, ncols, self.num)
        if self._labels is None:
            _log(output, 'Not enough informative cells to support model initialization. \n')
        else:
            self.initial = []
            for n in range(self.num):
                self.initial.append([self.barcodes[col] for col in icols[self._labels[icols] == n]])
            if not _defer_seed:
                self._set_af(self.engine.seed_af(self._labels, self.num))
        self._em_open = False

    # ---- helpers
    def _set_af(self, P):
        self.model_af = pd.DataFrame(P, index=self.all_POS, columns=range(self.num))

    def _open_em(self):
        """Make this model the engine's single active restart, with this object's model_af / lP_s / P_s_c."""
        eng = self.engine
        eng.em_begin(np.ascontiguousarray(self.model_af.values, dtype=np.float64))
        eng.set(SCS_LPS, np.asarray(self.lP_s, dtype=np.float64))
        self._em_open = True

    def _pull_e(self):
        eng = self.engine
        self.lP_c_s = pd.DataFrame(eng.get(SCS_L), index=self.barcodes, columns=range(self.num))
        self.P_s_c = pd.DataFrame(eng.get(SCS_W), index=self.barcodes, columns=range(self.num))

    def _pull_m(self):
        eng = self.engine
        self._set_af(eng.get(SCS_P))
        self.lP_s = pd.Series(eng.get(SCS_LPS), index=range(self.num))

    # ---- reference API
    def run_EM(self, output):
        """Expectation-Maximisation to the reference's stop rule (scSplit:148-162), entirely on the device."""
        eng = self.engine
        self._open_em()
        eng.em_run()
        it, conv, ll = eng.em_status()
        iterations = int(it[0])
        hist = eng.get(4)[:iterations]
        now = str(datetime.datetime.now())
        _log(output, ''.join('E-M Iteration ' + str(i + 1) + '   ' + now + '\n' for i in range(iterations)))
        self.sum_log_likelihood = [1, 2] + hist.tolist()
        self.lP_c_m = float(ll[0])
        self._pull_e()
        self._pull_m()
        if iterations < 300:
            self.convergence = 1

    def calculate_cell_likelihood(self):
        """One E-step (scSplit:165-188) from this object's model_af and lP_s."""
        self._open_em()
        self.engine.estep()
        self._pull_e()
        self.lP_c_m = self.engine.stats()["ll"]

    def calculate_model_af(self):
        """One M-step (scSplit:191-198) from this object's P_s_c."""
        if not self._em_open:
            self._open_em()
        self.engine.set(SCS_W, np.ascontiguousarray(self.P_s_c.values, dtype=np.float64))
        self.engine.mstep()
        self._pull_m()

    def assign_cells(self):
        """Cells with P(s|c) >= 0.99, sorted by barcode, per state (scSplit:201-207)."""
        W = self.P_s_c.values
        bc = np.asarray(self.barcodes, dtype=object)
        with np.errstate(invalid='ignore'):
            hit = W >= 0.99
        for n in range(self.num):
            self.assigned[n] = sorted(bc[hit[:, n]].tolist())

    def _index_lists(self, lists):
        pos = {b: i for i, b in enumerate(self.barcodes)}
        return [[pos[b] for b in lst] for lst in lists]

    def define_doublet(self):
        """Doublet state = first argmax_i of sum over assigned cells of L_i(c) with the post-M model (scSplit:210-225)."""
        weight = np.zeros(len(self.barcodes), dtype=np.int64)
        for lst in self._index_lists(self.assigned):
            np.add.at(weight, lst, 1)          # cross_state sums every (cluster j, cell) pair once (scSplit:219-223)
        if weight.max(initial=0) > 255:
            raise ValueError('a barcode is listed more than 255 times in `assigned`')
        scores = self.engine.doublet_scores(self.model_af.values, weight.astype(np.uint8))
        result = scores.tolist()
        self.doublet = result.index(max(result))

    def refine_doublets(self, doublets):
        """Top up the doublet cluster to the expected proportion (scSplit:228-252)."""
        found = []
        self.reassigned = self.assigned.copy()
        if doublets == 0:
            self.doublet = -1
        elif doublets > 0:
            W = self.P_s_c.values
            with np.errstate(invalid='ignore'):
                hit = W >= 0.99
            hit[:, self.doublet] = False                       # rpc = rps.drop(doublet, axis=1), scSplit:244
            # (W>=0.99) can hold for one state only, so the row sum of rps has a single non-zero term
            labels = np.where(hit.any(axis=1), hit.argmax(axis=1), -1).astype(np.int32)
            if (hit.sum(axis=1) > 1).any():
                raise ValueError('a cell is assigned to two states')
            score = self.engine.refine_scores(self.model_af.values, labels)
            lack = len(self.barcodes) * doublets - len(self.assigned[self.doublet])
            if lack > 0:
                n = len(self.barcodes)
                order = pd.Series(score, index=self.barcodes).argsort()       # scSplit:247 (quicksort, like the reference)
                found = pd.Index(self.barcodes)[order[int(n - lack):n]]
                fset = set(found.tolist())
                for k in range(self.num):
                    if k != self.doublet:
                        self.reassigned[k] = [x for x in self.assigned[k] if x not in fset]
                    else:
                        self.reassigned[k] += found.tolist()

    def distinguishing_alleles(self, pos=[]):
        """Per-cluster hard allele sums on the GPU, then the reference's variant selection on the host (scSplit:255-336)."""
        self.dist_variants, ncols = [], self.num - 1 + (self.doublet < 0) * 1
        # ---- reductions (device): N_ref / N_alt per cluster and the ternary presence/absence code, scSplit:269-282
        idx_lists = self._index_lists(self.reassigned)
        N = len(self.barcodes)
        labels = np.full(N, -1, dtype=np.int32)
        overlap = False
        for n, lst in enumerate(idx_lists):
            arr = np.asarray(lst, dtype=np.int64)
            if arr.size and (labels[arr] >= 0).any() and (labels[arr][labels[arr] >= 0] != n).any():
                overlap = True
            labels[arr] = n
        if not overlap:
            _, _, code = self.engine.cluster_counts(labels, self.num)
        else:   # a barcode listed under two clusters: one masked pass per cluster keeps the membership semantics
            code = np.zeros((len(self.all_POS), self.num), dtype=np.int8)
            for n, lst in enumerate(idx_lists):
                lab = np.full(N, -1, dtype=np.int32)
                lab[np.asarray(lst, dtype=np.int64)] = 0
                _, _, c1 = self.engine.cluster_counts(lab, 1)
                code[:, n] = c1[:, 0]
        snv = self.all_POS if len(pos) == 0 else [self.all_POS[i] for i in pos]
        if len(pos) != 0:
            code = code[np.asarray(pos, dtype=np.int64)]
        alt_or_ref = pd.DataFrame(code, index=snv, columns=range(self.num))
        if self.doublet != -1:
            alt_or_ref = alt_or_ref.drop(self.doublet, axis=1)
        alt_or_ref = alt_or_ref.astype(np.float64)
        vals = alt_or_ref.values.copy()
        vals[code_is(alt_or_ref.values, 0)] = np.nan          # unknown
        vals[code_is(alt_or_ref.values, -1)] = 0              # REF only -> ALT absent
        alt_or_ref = pd.DataFrame(vals, index=alt_or_ref.index, columns=alt_or_ref.columns)
        alt_or_ref = alt_or_ref.loc[[x for x in alt_or_ref.index if x[0] not in ['X', 'Y', 'M']]]
        # ---- selection (host, tiny): scSplit:284-336
        self.dist_variants = select_distinguishing(alt_or_ref, self.num, ncols)
        self.dist_variants = list(set(self.dist_variants))
        self.dist_matrix = alt_or_ref.reindex(self.dist_variants)
        self.pa_matrix = alt_or_ref


def code_is(values, c):
    return values == c


def _row_var(frame):
    return frame.var(axis=1)


def _representatives_by_text(alt_or_ref, submatrix, informative):
    """scSplit:291-296 as written: rows grouped by the concatenated text of their values."""
    selected = []
    informative_sub = submatrix[informative]
    patt = informative_sub.astype(str).values.sum(axis=1)
    inverse = np.asarray(np.unique(patt, return_inverse=True)[1]).ravel()
    for j in range(int(inverse.max()) + 1):
        members = informative_sub.index[np.flatnonzero(inverse == j)]
        selected.append(alt_or_ref.loc[members].count(axis=1).idxmax())
    return selected


def pattern_representatives(alt_or_ref, submatrix, informative):
    """One SNV per distinct presence/absence pattern among the informative rows of the column window (scSplit:291-296):
    the row observed in most clusters of the ORIGINAL matrix, the first such row on ties.

    The reference groups the rows by the concatenated text of their values and looks every group up again by label; the
    informative rows are fully observed rows of a 0/1 matrix, so the pattern is an integer (one bit per cluster) and the
    representative of every pattern is the head of its group after one sort by (pattern, observed clusters descending,
    row position) -- the only step of the selection whose cost grows with the number of SNVs.  Frames that are not 0/1
    or whose labels repeat (label look-ups would then return several rows) are grouped the reference's way.
    """
    pos = np.flatnonzero(informative)
    vals = submatrix.values[pos]
    if vals.shape[1] > 62 or not alt_or_ref.index.is_unique or not ((vals == 0) | (vals == 1)).all():
        return _representatives_by_text(alt_or_ref, submatrix, informative)
    key = (vals.astype(np.int64) << np.arange(vals.shape[1], dtype=np.int64)).sum(axis=1)
    observed = alt_or_ref.count(axis=1).values[pos]
    order = np.lexsort((pos, -observed, key))
    heads = order[np.concatenate(([True], key[order][1:] != key[order][:-1]))]
    return alt_or_ref.index[pos[heads]].tolist()


def select_distinguishing(alt_or_ref, num, ncols):
    """Variant selection of scSplit:284-333 on the ternary matrix (rows = SNVs, columns = non-doublet clusters).

    Greedy search, window by window of the leading `ncols` columns, for rows that are fully observed and together
    have rank `ncols`; clusters that stay identical on the chosen rows form the next window.  Same numpy / pandas
    primitives in the same order as the reference so that ties resolve identically.
    """
    dist_variants = []
    submatrix = alt_or_ref.copy()
    while len(dist_variants) < num - 1:
        found, todo = [], []
        while len(found) < ncols - 1:
            informative = ((_row_var(submatrix) > 0) & (submatrix.count(axis=1) == ncols)).values
            if informative.any():
                selected = pattern_representatives(alt_or_ref, submatrix, informative)
                subt = submatrix.reindex(np.unique(selected)).iloc[:, 0:ncols]
                subt = subt.fillna(-1)
                d = np.linalg.svd(subt, full_matrices=False)[1]
                if sum(d > 1e-10) >= ncols:
                    rs = subt.sum(axis=1)
                    found = [subt[rs == min(rs)].index[0]]
                    while len(found) < ncols:                      # greedy orthogonal completion
                        basis = subt.reindex(found).transpose()
                        svd = np.linalg.svd(basis, full_matrices=False)
                        U, Vt = svd[0], svd[2]
                        D = np.asmatrix(svd[1]) if len(found) == 1 else np.diag(svd[1])
                        colU = U.shape[1]
                        Vinv = np.linalg.solve(Vt, np.diag([1] * len(Vt)))
                        Dinv = np.linalg.solve(D, np.diag([1] * len(D)))
                        proj = np.asmatrix(np.zeros((colU, subt.shape[0])))
                        rows = np.ascontiguousarray(subt.values, dtype=np.float64)
                        for j in range(subt.shape[0]):
                            for i in range(colU):
                                proj[i, j] = np.dot(U[:, i], rows[j])
                        Wc = np.dot(np.dot(Vinv, Dinv), proj)
                        R = np.dot(basis, Wc)
                        diff = (subt.transpose() - R).transpose()
                        dv = diff.var(axis=1)
                        subt1 = subt.reindex(diff[dv > (0.5 * max(dv))].index)
                        rs1 = subt1.sum(axis=1)
                        found += [subt1[rs1 == min(rs1)].index[0]]
            ncols -= 1
            submatrix = submatrix.iloc[:, 0:ncols]
        dist_variants += found
        if len(found) == 0:
            with open('scSplit_dist_variants.txt', 'w') as logfile:     # CWD, like scSplit:323
                logfile.write('Not all clusters can be distinguished. \n')
            break
        chosen = alt_or_ref.reindex(dist_variants)
        for i in range(1, len(alt_or_ref.columns)):
            for j in range(i):
                if (chosen.iloc[:, [i, j]].var(axis=1)).sum() == 0:
                    todo.extend([i, j])
        submatrix = alt_or_ref.iloc[:, sorted(set(todo))]
        ncols = len(submatrix.columns)
    return dist_variants


def _assign_lists(W, barcodes):
    """assign_cells on a bare posterior matrix (scSplit:201-207)."""
    bc = np.asarray(barcodes, dtype=object)
    with np.errstate(invalid='ignore'):
        hit = W >= 0.99
    return [sorted(bc[hit[:, n]].tolist()) for n in range(W.shape[1])]


def _labels_of_trim(base_calls_mtx, trim, barcodes, num, output):
    """PCA / KMeans on a trimmed block (scSplit:126-142) without the model's frames: (labels or None, initial lists)."""
    irows, icols, nrows, ncols = trim
    labels = initialiser.initial_labels(base_calls_mtx[0], base_calls_mtx[1], irows, icols, nrows, ncols, num)
    if labels is None:
        _log(output, 'Not enough informative cells to support model initialization. \n')
        return None, None
    return labels, [[barcodes[col] for col in icols[labels[icols] == n]] for n in range(num)]


def _initialise(engine, base_calls_mtx, barcodes, num, output):
    """The initialisation part of models.__init__ (scSplit:105-142) without the model's frames: (labels or None, initial lists)."""
    trim = initialiser.trim_subset(engine.pattern_counts, base_calls_mtx[0].shape[0], len(barcodes), num)
    return _labels_of_trim(base_calls_mtx, trim, barcodes, num, output)


def _same_rng_state(a, b):
    return a[0] == b[0] and a[2:] == b[2:] and np.array_equal(a[1], b[1])


def _initialise_all(eng, base_calls_mtx, barcodes, num, output, ems, device):
    """The `ems` initialisations of run.core (scSplit:473-477) on one shared `np.random` stream: (last model, labels, initial).

    The reference alternates, per restart, the trimming loop (the only consumer of `np.random`: scSplit:113-114) and
    StandardScaler / PCA / KMeans(random_state=0) on the surviving block.  Here the trimming loops of all restarts run
    first, back to back on the stream, and the blocks are clustered afterwards: the loops need nothing of scikit-learn,
    so they run while its modules are still being imported (cli.run loads them in a helper thread; after the count CSVs
    that import is the longest single step of a run).  The order only matters if the clustering draws from `np.random`
    as well (PCA's randomised solver on a large surviving block does): every block is therefore clustered from the
    stream state its restart had in the reference's order, and the first clustering that moves the stream sends the
    restarts after it back to the reference's alternating order, from the state it left.
    """
    V, N = base_calls_mtx[0].shape[0], len(barcodes)
    trims, after = [], []
    for i in range(ems):
        trims.append(initialiser.trim_subset(eng.pattern_counts, V, N, num))
        after.append(np.random.get_state())
    model, labels, initial = None, [None] * ems, [None] * ems
    ahead = True                                  # trims[i] / after[i] still are what the reference's order gives
    for i in range(ems):
        _log(output, 'Repeat ' + str(i + 1) + ' of ' + str(ems) + ' in demultiplexing ' + str(num) +
             ' subpopulations (including doublets if expected)' + '\n')
        _log(output, 'Building model... \n')
        if ahead:
            np.random.set_state(after[i])
            trim = trims[i]
        else:
            trim = initialiser.trim_subset(eng.pattern_counts, V, N, num)
        before = np.random.get_state()
        if i == ems - 1:
            model = models(base_calls_mtx, num, output, device=device, _defer_seed=True, _trim=trim)
            labels[i], initial[i] = model._labels, getattr(model, 'initial', None)
        else:
            labels[i], initial[i] = _labels_of_trim(base_calls_mtx, trim, barcodes, num, output)
        if ahead and not _same_rng_state(before, np.random.get_state()):
            ahead = False
    if ahead:
        np.random.set_state(after[-1])
    return model, labels, initial


def _world():
    """(rank, world) of a torchrun launch whose process group is up, else (0, 1)."""
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return dist.get_rank(), dist.get_world_size()
    except Exception:
        pass
    return 0, 1


def _bcast(arr, src, device):
    """Broadcast a numpy array from rank `src` (shape and dtype known on every rank)."""
    import torch
    import torch.distributed as dist
    t = torch.from_numpy(np.ascontiguousarray(arr))
    if dist.get_backend() == 'nccl':
        t = t.cuda(device)
    dist.broadcast(t, src)
    return t.cpu().numpy()


def core_batched(base_calls_mtx, num, output, ems, device=0, chunk=64):
    """Restart driver (run.core, scSplit:470-490): `ems` restarts, first strictly greatest LL wins.

    The R initialisations are generated serially on the host first -- they share one `np.random` stream and never read
    EM results (hard part 6) -- then the EM runs execute in batches on the GPU; a batch holds as many
    restarts as the device has memory for (`scs_em_footprint`), at most `chunk`.  Only the LAST restart is a full `models`
    object (the reference returns that object with four attributes of the best restart grafted on, scSplit:488); the
    others are kept as labels + results.

    Under torchrun (an initialised torch.distributed process group) the restarts are sharded over the ranks, one GPU
    each: rank 0 draws all initialisations (one `np.random` stream, as above) and broadcasts the labels, rank r runs
    restarts r, r + world, ... (dist.shard_restarts), the log-likelihoods are gathered and the reference's selection rule is
    applied to them in restart order on every rank, and the winner's arrays travel to rank 0, the only rank that returns a
    populated model (the others return (None, best)).
    Returns (model, chosen_index); raises ValueError like scSplit:486 when no restart could be initialised.
    """
    from . import dist as D
    rank, world = _world()
    barcodes = base_calls_mtx[3].tolist()
    N, V = len(barcodes), base_calls_mtx[0].shape[0]
    eng = get_engine(base_calls_mtx, device)
    model, labels, initial = None, [None] * ems, [None] * ems
    if rank == 0:
        model, labels, initial = _initialise_all(eng, base_calls_mtx, barcodes, num, output, ems, device)
    if world > 1:
        lab = np.full((ems, N), -2, dtype=np.int32)          # -2 rows: restarts whose initialisation failed
        if rank == 0:
            for i in range(ems):
                if labels[i] is not None:
                    lab[i] = labels[i]
        lab = _bcast(lab, 0, device)
        labels = [None if (lab[i] == -2).all() else lab[i] for i in range(ems)]
    mine = D.shard_restarts(ems, world)[rank]
    live = [i for i in mine if labels[i] is not None]
    per, free = eng.em_footprint(num)
    chunk = int(max(1, min(chunk, (0.7 * free) // max(per, 1))))
    ll_all = np.full(ems, np.nan)
    iters_all = np.zeros(ems, dtype=np.int64)
    best_local, best_ll, keep, last = None, -1e50, None, None
    for c0 in range(0, len(live), chunk):
        part = live[c0:c0 + chunk]
        P0 = eng.seed_af(np.stack([labels[i] for i in part]), num)
        if P0.ndim == 2:
            P0 = P0[None]
        ok = [r for r in range(len(part)) if P0[r].sum() > 0]             # scSplit:478 gate
        if not ok:
            continue
        eng.em_begin(np.ascontiguousarray(P0[ok]))
        eng.em_run()
        it, conv, ll = eng.em_status()
        for r, pr in enumerate(ok):
            i = part[pr]
            ll_all[i], iters_all[i] = float(ll[r]), int(it[r])
            if i == ems - 1:       # the reference returns the LAST model object and later reads ITS lP_c_s (autodetect, scSplit:555 with 488)
                last = dict(L=eng.get(SCS_L, r), P0=P0[pr])
            if float(ll[r]) > best_ll:                                      # first strictly greater within this rank's order
                best_local, best_ll = i, float(ll[r])
                keep = dict(W=eng.get(SCS_W, r), P=eng.get(SCS_P, r), L=eng.get(SCS_L, r), lps=eng.get(SCS_LPS, r),
                            hist=eng.get(4, r)[:int(it[r])], iters=int(it[r]))
    if world > 1:
        import torch
        import torch.distributed as dist
        pack = np.stack([np.nan_to_num(ll_all, nan=-np.inf), iters_all.astype(np.float64)])
        t = torch.from_numpy(pack)
        t = t.cuda(device) if dist.get_backend() == 'nccl' else t
        dist.all_reduce(t, op=dist.ReduceOp.MAX)                            # every restart ran on exactly one rank
        pack = t.cpu().numpy()
        ll_all = np.where(np.isinf(pack[0]), np.nan, pack[0])
        iters_all = pack[1].astype(np.int64)
    best = D.select_best([None if np.isnan(x) else float(x) for x in ll_all])
    if rank == 0:
        now = str(datetime.datetime.now())
        for i in range(ems):
            if not np.isnan(ll_all[i]):
                _log(output, ''.join('E-M Iteration ' + str(t + 1) + '   ' + now + '\n' for t in range(int(iters_all[i]))))
                _log(output, 'Model Log-Likelihood = ' + str(float(ll_all[i])) + '\n')
    if best is None:
        raise ValueError('Model not converged, please allow more repeats!')
    if world > 1:
        owner = best % world                                                # dist.shard_restarts deals round robin
        K = num
        shapes = dict(W=(N, K), P=(V, K), L=(N, K), lps=(K,), hist=(int(iters_all[best]),))
        got = {}
        for name, shp in shapes.items():
            src = keep[name] if (rank == owner and best_local == best) else np.zeros(shp)
            got[name] = _bcast(np.asarray(src, dtype=np.float64).reshape(shp), owner, device)
        keep = dict(got, iters=int(iters_all[best]))
        lastL = np.zeros((N, K))
        last_owner = (ems - 1) % world
        have_last = not np.isnan(ll_all[ems - 1])
        if have_last:
            lastL = _bcast(last["L"] if rank == last_owner else lastL, last_owner, device)
            lastP0 = _bcast(last["P0"] if rank == last_owner else np.zeros((V, K)), last_owner, device)
            last = dict(L=lastL, P0=lastP0)
        if rank != 0:
            return None, best
    # ---- rank 0 (or the only process): the LAST model object with the best restart grafted on (scSplit:488)
    if not np.isnan(ll_all[ems - 1]):
        model._set_af(last["P0"])
        model.lP_c_m = float(ll_all[ems - 1])
        model.convergence = int(iters_all[ems - 1] < 300)
        model.lP_c_s = pd.DataFrame(last["L"], index=model.barcodes, columns=range(num))
    if best == ems - 1:
        model.sum_log_likelihood = [1, 2] + np.asarray(keep["hist"]).tolist()
        model.lP_s = pd.Series(keep["lps"], index=range(num))
    model.P_s_c = pd.DataFrame(keep["W"], index=model.barcodes, columns=range(num))
    model._set_af(keep["P"])
    model.assigned = _assign_lists(keep["W"], barcodes)
    model.initial = initial[best]
    if not hasattr(model, 'lP_c_m'):
        model.lP_c_m = float(ll_all[best])
    print('Repeat ' + str(best + 1) + ' was chosen.')
    return model, best
