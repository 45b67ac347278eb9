"""numpy-in / numpy-out wrapper of one scs_ctx (one GPU, one shard of cells)."""
import ctypes

import numpy as np

from . import _lib
from ._lib import SCS_L, SCS_LLHIST, SCS_LPS, SCS_P, SCS_STATS, SCS_W, check


def _ptr(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def _expect(a, shape, name):
    """The library reads plain pointers of implied length: a wrong shape must stop here, not read past the buffer."""
    if a.shape != tuple(shape):
        raise ValueError("%s has shape %r, expected %r" % (name, a.shape, tuple(shape)))
    return a


def _canonical_csr(m):
    """scipy CSR with sorted, duplicate-free indices and no explicit zeros (what csr_matrix(dense) gives, scSplit:546)."""
    from scipy.sparse import csr_matrix
    m = csr_matrix(m)
    if not m.has_canonical_format:
        m = m.copy()
        m.sum_duplicates()
    if m.nnz and (m.data == 0).any():
        m = m.copy()
        m.eliminate_zeros()
    return m


class Engine:
    """Device-resident count matrices + EM state.  All arrays crossing this boundary are host numpy arrays."""

    def __init__(self, ref, alt, device=0, canonicalize=True):
        """ref, alt: scipy CSR V x N (the reference's base_calls_mtx[0:2], scSplit:546-547).

        canonicalize=False skips the host-side scans (the caller guarantees sorted, duplicate-free indices, no
        explicit zeros and int16/int32/int64 counts); the device build still validates and raises on violation.
        """
        self.lib = _lib.load()
        if not canonicalize:
            self._create(ref, alt, device, ref.data.dtype)
            return
        ref, alt = _canonical_csr(ref), _canonical_csr(alt)
        if ref.shape != alt.shape:
            raise ValueError("ref and alt matrices differ in shape: %r vs %r" % (ref.shape, alt.shape))
        self.V, self.N = int(ref.shape[0]), int(ref.shape[1])
        for m in (ref, alt):
            if m.dtype.kind not in "iu" and not np.array_equal(m.data, np.rint(m.data)):
                raise ValueError("allele counts must be integers")
            if m.nnz and (m.data.min() < 0 or m.data.max() > np.iinfo(np.int32).max):
                raise ValueError("allele counts must lie in [0, 2^31): never wrapped (scSplit:38-39 uses int16)")
        top = max([int(m.data.max()) for m in (ref, alt) if m.nnz] + [0])
        dt = np.dtype(np.int16 if top <= np.iinfo(np.int16).max else np.int32)   # narrowest transfer type that fits
        self._create(ref, alt, device, dt)

    def _create(self, ref, alt, device, dt):
        dt = np.dtype(dt)
        if dt not in (np.dtype(np.int16), np.dtype(np.int32), np.dtype(np.int64)):
            raise ValueError("count dtype must be int16, int32 or int64 (got %s)" % dt)
        if ref.shape != alt.shape:
            raise ValueError("ref and alt matrices differ in shape: %r vs %r" % (ref.shape, alt.shape))
        self.V, self.N = int(ref.shape[0]), int(ref.shape[1])
        arrs = []
        for m in (ref, alt):
            arrs += [np.ascontiguousarray(m.indptr, dtype=np.int64), np.ascontiguousarray(m.indices, dtype=np.int32),
                     np.ascontiguousarray(m.data, dtype=dt)]
        self._keep = arrs
        h = ctypes.c_void_p()
        check(self.lib.scs_create(ctypes.byref(h), int(device), self.V, self.N, _ptr(arrs[0]), _ptr(arrs[1]), _ptr(arrs[2]),
                                  _ptr(arrs[3]), _ptr(arrs[4]), _ptr(arrs[5]), dt.itemsize))
        self.h = h
        self._keep = None
        self.R = self.K = 0
        self.max_iter = 0

    # -- lifetime
    def close(self):
        if getattr(self, "h", None):
            self.lib.scs_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # -- layout
    def layout_info(self):
        out = (ctypes.c_int64 * 6)()
        check(self.lib.scs_layout_info(self.h, out))
        return dict(nnz_ref=out[0], nnz_alt=out[1], nnz_union=out[2], light=out[3], heavy=out[4], device_bytes=out[5])

    def kalt(self):
        out = np.empty(self.V, dtype=np.float64)
        check(self.lib.scs_kalt(self.h, _ptr(out)))
        return out

    def seed_af(self, labels, K):
        labels = np.ascontiguousarray(labels, dtype=np.int32)
        single = labels.ndim == 1
        if labels.ndim > 2 or labels.shape[-1] != self.N:
            raise ValueError("labels must be [N] or [R][N] with N = %d cells (got %r)" % (self.N, labels.shape))
        labels = labels.reshape(-1, self.N)
        R = labels.shape[0]
        out = np.empty((R, self.V, K), dtype=np.float64)
        check(self.lib.scs_seed_af(self.h, R, int(K), _ptr(labels), _ptr(out)))
        return out[0] if single else out

    def pattern_counts(self, keep_rows=None, keep_cols=None):
        kr = None if keep_rows is None else _expect(np.ascontiguousarray(keep_rows, dtype=np.uint8), (self.V,), "keep_rows")
        kc = None if keep_cols is None else _expect(np.ascontiguousarray(keep_cols, dtype=np.uint8), (self.N,), "keep_cols")
        rc = np.empty(self.V, dtype=np.int64)
        cc = np.empty(self.N, dtype=np.int64)
        check(self.lib.scs_pattern_counts(self.h, _ptr(kr), _ptr(kc), _ptr(rc), _ptr(cc)))
        return rc, cc

    # -- EM
    def em_footprint(self, K):
        """(device bytes one batched restart of K states needs, device bytes still available)."""
        out = (ctypes.c_int64 * 2)()
        check(self.lib.scs_em_footprint(self.h, int(K), out))
        return int(out[0]), int(out[1])

    def em_begin(self, P0, max_iter=_lib.SCS_MAX_ITER):
        P0 = np.ascontiguousarray(P0, dtype=np.float64)
        if P0.ndim == 2:
            P0 = P0[None]
        R, V, K = P0.shape
        if V != self.V:
            raise ValueError("P0 has %d SNVs, matrices have %d" % (V, self.V))
        check(self.lib.scs_em_begin(self.h, R, K, _ptr(P0), int(max_iter)))
        self.R, self.K, self.max_iter = R, K, int(max_iter)

    def em_iterate(self, n_iter, check_conv=True):
        check(self.lib.scs_em_iterate(self.h, int(n_iter), 1 if check_conv else 0))

    def em_capture(self, check_conv=True):
        """Capture one iteration's launches as CUDA graphs now (after at least one iteration has run); executes nothing."""
        check(self.lib.scs_em_capture(self.h, 1 if check_conv else 0))

    def last_ms(self):
        """Device milliseconds (CUDA events on the library's stream) of the last em_iterate call."""
        ms = ctypes.c_double(0.0)
        check(self.lib.scs_em_last_ms(self.h, ctypes.byref(ms)))
        return ms.value

    def em_run(self):
        check(self.lib.scs_em_run(self.h))

    def estep(self):
        check(self.lib.scs_estep(self.h))

    def mstep(self):
        check(self.lib.scs_mstep(self.h))

    def em_status(self):
        it = np.empty(self.R, dtype=np.int32)
        cv = np.empty(self.R, dtype=np.int32)
        ll = np.empty(self.R, dtype=np.float64)
        check(self.lib.scs_em_status(self.h, _ptr(it), _ptr(cv), _ptr(ll)))
        return it, cv, ll

    def em_info(self):
        """How the E-step of the current EM state runs (fixed-point row lanes, fraction bits, tiles, restarts on the fp64 stand-in)."""
        out = (ctypes.c_int64 * 6)()
        check(self.lib.scs_em_info(self.h, out))
        return dict(q_lanes=int(out[0]), q_frac_bits=int(out[1]), tiles=int(out[2]), q_unfit_restarts=int(out[3]),
                    q_form={0: "fp64", 1: "row groups", 2: "lock-step", 3: "lock-step, two 64-byte tables"}[int(out[4])], q_tasks=int(out[5]))

    def _shape(self, what):
        return {SCS_P: (self.V, self.K), SCS_W: (self.N, self.K), SCS_L: (self.N, self.K), SCS_LPS: (self.K,),
                SCS_LLHIST: (self.max_iter,)}[what]

    def get(self, what, restart=0):
        out = np.empty(self._shape(what), dtype=np.float64)
        check(self.lib.scs_em_get(self.h, int(restart), int(what), _ptr(out)))
        return out

    def set(self, what, value, restart=0):
        value = np.ascontiguousarray(value, dtype=np.float64)
        if value.shape != self._shape(what):
            raise ValueError("expected shape %r, got %r" % (self._shape(what), value.shape))
        check(self.lib.scs_em_set(self.h, int(restart), int(what), _ptr(value)))

    def stats(self, restart=0):
        """Packed M-step statistics of one restart: S[2V][KP] (alt rows 2v, ref rows 2v+1), column mass, last E-step LL."""
        KP = self.K + (self.K & 1)
        out = np.empty(2 * self.V * KP + KP + 2, dtype=np.float64)
        check(self.lib.scs_em_get(self.h, int(restart), SCS_STATS, _ptr(out)))
        n = 2 * self.V * KP
        return dict(S=out[:n].reshape(2 * self.V, KP), colsum=out[n:n + self.K].copy(), ll=float(out[n + KP]),
                    dense_cells=float(out[n + KP + 1]))   # posterior rows that did not pack (sparse M-step gate)

    def em_end(self):
        check(self.lib.scs_em_end(self.h))
        self.R = self.K = 0

    # -- post-EM
    def doublet_scores(self, P, mask):
        P = np.ascontiguousarray(P, dtype=np.float64)
        if P.ndim != 2:
            raise ValueError("P must be a V x K matrix")
        _expect(P, (self.V, P.shape[1]), "P")
        mask = _expect(np.ascontiguousarray(mask, dtype=np.uint8), (self.N,), "mask")
        out = np.empty(P.shape[1], dtype=np.float64)
        check(self.lib.scs_doublet_scores(self.h, P.shape[1], _ptr(P), _ptr(mask), _ptr(out)))
        return out

    def refine_scores(self, P, labels):
        P = np.ascontiguousarray(P, dtype=np.float64)
        if P.ndim != 2:
            raise ValueError("P must be a V x K matrix")
        _expect(P, (self.V, P.shape[1]), "P")
        labels = _expect(np.ascontiguousarray(labels, dtype=np.int32), (self.N,), "labels")
        out = np.empty(self.N, dtype=np.float64)
        check(self.lib.scs_refine_scores(self.h, P.shape[1], _ptr(P), _ptr(labels), _ptr(out)))
        return out

    def cluster_counts(self, labels, K):
        labels = _expect(np.ascontiguousarray(labels, dtype=np.int32), (self.N,), "labels")
        n_ref = np.empty((self.V, K), dtype=np.int64)
        n_alt = np.empty((self.V, K), dtype=np.int64)
        pa = np.empty((self.V, K), dtype=np.int8)
        check(self.lib.scs_cluster_counts(self.h, int(K), _ptr(labels), _ptr(n_ref), _ptr(n_alt), _ptr(pa)))
        return n_ref, n_alt, pa

    # -- multi-GPU
    @staticmethod
    def comm_unique_id():
        buf = np.zeros(128, dtype=np.uint8)
        check(_lib.load().scs_comm_unique_id(_ptr(buf)))
        return buf

    def comm_init(self, uid, rank, world):
        uid = _expect(np.ascontiguousarray(uid, dtype=np.uint8), (128,), "uid")
        check(self.lib.scs_comm_init(self.h, _ptr(uid), int(rank), int(world)))

    # -- instrumentation
    def launch_count(self):
        n = ctypes.c_int64(0)
        check(self.lib.scs_launch_count(self.h, ctypes.byref(n)))
        return n.value

    def set_timing(self, on=True, every=1):
        """CUDA-event timing of the E-step SpMM and of the M-step sums, on every `every`-th EM iteration."""
        check(self.lib.scs_set_timing(self.h, int(every) if on else 0))

    def kernel_times(self, reset=False):
        out = (ctypes.c_double * 8)()
        check(self.lib.scs_kernel_times(self.h, out, 1 if reset else 0))
        return dict(e_us=out[0], m_us=out[1], e_n=int(out[2]), m_n=int(out[3]), allreduce_us=out[4], allreduce_n=int(out[5]),
                    bayes_us=out[6], bayes_n=int(out[7]))

    def set_variant(self, v):
        check(self.lib.scs_set_variant(self.h, int(v)))

    def selftest_exp2(self, x):
        """2**x through the Bayes kernel's own device routine (diagnostic; tests compare it with numpy's)."""
        x = np.ascontiguousarray(x, dtype=np.float64)
        y = np.empty_like(x)
        check(self.lib.scs_selftest_exp2(self.h, x.ctypes.data, y.ctypes.data, x.size))
        return y
