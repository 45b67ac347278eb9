#!/usr/bin/env python3
"""Benchmark of the scSplit `run` EM hot path on B200 (contract: see DESIGN.md "Measurement").

  python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo's CUDA path
  python bench.py --impl reference [--gpus N] [--steps K] ...    # the reference's CPU path on the host cores

A *step* is one EM iteration (E-step + M-step, scSplit:159-160) over the whole synthetic pool of BASELINE.json's
metric configuration: 30,000 SNVs x 20,000 cells, 8 samples + doublet state (K = 9), Poisson depth lam = 0.05.
`value` = iterations/s with everything resident in HBM, timed with CUDA events on the library's stream.
`e2e`   = the same metric through the public host API (scsplit_b200.engine.Engine -> C-ABI) with pinned HOST
          buffers: H2D of the count matrices + device layout build + `e2e_iters` iterations + D2H of P_s_c,
          model_af and LL, all inside the timed region.
N > 1 (torchrun): independent restarts are sharded over the GPUs (preferred for this
config), every rank runs the same pool from a different initialisation, no data-path collective -> weak scaling.
Every run then adds a second leg under the key `cells_c4`: BASELINE config 4 (60k SNVs x 200k cells, K = 17) with the
CELLS sharded over the N GPUs and the M-step statistics all-reduced by NCCL every iteration (at N = 1: the whole pool on
one GPU, the anchor of that strong-scaling curve).  `--mode cells` runs the main leg itself cell-sharded.
Further keys of the N = 1 line: `regimes` (soft-posterior start of a run, whole run to the stop rule), `small_configs`
(configs 1 and 2 next to their roofline times), `cpu_baseline`.
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "em_iters_per_sec"
UNIT = "iter/s"
KERNEL_TIMING_EVERY = 8  # the per-kernel CUDA events (roofline) are recorded on every 8th iteration of the timed loop: on every
                         # iteration they cost ~11 us of the ~270 us step at config 3 (tools/timing_overhead.py)
E2E_ITERS = 8           # iterations per end-to-end call: the reference's median per restart
L2_BYTES = 126e6


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def algorithmic_bytes(V, N, K, nnz_union):
    """8 B per union-pattern entry per pass + row pointers + dense operands."""
    e = 8 * nnz_union + 8 * (N + 1) + 2 * V * K * 8 + 2 * N * K * 8
    m = 8 * nnz_union + 8 * (V + 1) + N * K * 8 + 3 * V * K * 8 + V * 8
    return e, m


def workload(name, cell_lo=0, cell_hi=None):
    from scsplit_b200 import synth
    cfg = synth.CONFIGS[name]
    pool = synth.make_config(name, cell_lo=cell_lo, cell_hi=cell_hi)
    return cfg, pool


def random_seed_labels(N, K, seed, per_cluster=3):
    """A cheap stand-in for the reference's initial subset (scSplit:126-142): a few random cells per state."""
    rng = np.random.default_rng(seed)
    lab = np.full(N, -1, dtype=np.int32)
    pick = rng.choice(N, K * per_cluster, replace=False)
    lab[pick] = np.arange(K * per_cluster) % K
    return lab


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """Samples SM clock and throttle reasons of one GPU while the timed region runs (NVML, 5 ms period)."""
    REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting"}

    def __init__(self, index):
        self.index, self.samples, self.mask, self.stop, self.thread = index, [], 0, False, None
        self.max_mhz = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[index]) if vis and vis.split(",")[index].isdigit() else index
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nv = None

    def _loop(self):
        nv = self.nv
        while not self.stop:
            try:
                self.samples.append(int(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                try:
                    self.mask |= int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                except Exception:
                    self.mask |= int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
            except Exception:
                pass
            time.sleep(0.005)

    def __enter__(self):
        if self.nv:
            self.thread = threading.Thread(target=self._loop, daemon=True)
            self.thread.start()
        return self

    def __exit__(self, *a):
        self.stop = True
        if self.thread:
            self.thread.join()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unavailable"]}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "samples": len(self.samples),
                "reasons": sorted(v for k, v in self.REASONS.items() if self.mask & k)}


# ------------------------------------------------------------------------------------------------ reference arm
def reference_models(pool, K, P0):
    """An instance of the REFERENCE's `models` class with its state filled in directly (its __init__ costs ~70 s at
    this size: dense toarray + trimming, scSplit:106-125), so that calculate_cell_likelihood / calculate_model_af
    (scSplit:165-198) can be timed.  Returns (step_fn, kind)."""
    import pandas as pd
    from oracle import ref_loader
    ref, alt = pool.ref.astype(np.int64), pool.alt.astype(np.int64)
    if ref_loader.reference_available() or ref_loader.patched_copy_available():
        g = ref_loader.load_reference(prefer_copy=True)
        M = g["models"]
        m = object.__new__(M)
        m.ref_bc_mtx, m.alt_bc_mtx = ref, alt
        m.all_POS, m.barcodes = list(pool.snv_ids), list(pool.barcodes)
        m.num = K
        m.P_s_c = pd.DataFrame(0.0, index=m.barcodes, columns=range(K))
        m.lP_c_s = pd.DataFrame(0.0, index=m.barcodes, columns=range(K))
        m.lP_s = [np.log2(1 / K)] * K
        m.model_af = pd.DataFrame(P0.copy(), index=m.all_POS, columns=range(K))
        m.pseudo = 1
        n_alt = alt.sum(axis=1) + 1
        n_ref = ref.sum(axis=1) + 1
        m.k_alt = n_alt / (n_ref + n_alt)

        def step():
            m.calculate_cell_likelihood()
            m.calculate_model_af()
            return float(m.lP_c_m)
        return step, "reference"
    from oracle import em_oracle as O
    state = {"P": P0.copy(), "lPs": np.array([np.log2(1 / K)] * K)}
    k_alt = O.kalt(ref, alt)

    def step():
        L, W, LL = O.estep(ref, alt, state["P"], state["lPs"], faithful=True)
        state["P"], state["lPs"] = O.mstep(ref, alt, W, k_alt)
        return LL
    return step, "port"


def workload_string(name, V, N, K, nnz_union):
    return "%s: %d SNVs x %d cells, K=%d states incl. doublet state, Poisson lam=0.05, nnz(union)=%d" % (name, V, N, K, nnz_union)


def time_reference(pool, K, n_steps, n_warm, budget_s):
    """Time the reference's E+M iteration (scSplit:165-198).  The whole pool when n_steps + n_warm iterations fit `budget_s`
    seconds (6.6 s per iteration at config 3 on one core), otherwise a stated column sample scaled to the pool."""
    from oracle import em_oracle as O
    V, N = pool.ref.shape
    full_s = 8.7 * (pool.ref.nnz + pool.alt.nnz) / 29.4e6 * K / 9.0   # measured at this shape: BASELINE.md section 2
    frac = min(1.0, budget_s / max(full_s * (n_steps + n_warm), 1e-9))
    n_s = N if frac >= 1.0 else min(N, max(200, int(N * frac)))
    sub = pool if n_s == N else pool._replace(ref=pool.ref[:, :n_s].tocsr(), alt=pool.alt[:, :n_s].tocsr(), barcodes=pool.barcodes[:n_s])
    k_alt = O.kalt(sub.ref, sub.alt)
    P0 = O.seed_af(sub.ref, sub.alt, k_alt, random_seed_labels(n_s, K, 0), K)
    step, kind = reference_models(sub, K, P0)
    for _ in range(n_warm):
        step()
    t0 = time.perf_counter()
    for _ in range(n_steps):
        step()
    dt = (time.perf_counter() - t0) / n_steps
    # cost is linear in the number of stored entries, i.e. in cells at fixed depth: a sample is scaled to the full pool
    full_iter_s = dt * (N / n_s)
    sample = "%d timed + %d warm-up E+M iterations on %s (all %d SNVs, K=%d)" % (
        n_steps, n_warm, "the full pool of %d cells" % N if n_s == N else "the first %d of %d cells, scaled x%.2f to the full pool" % (n_s, N, N / n_s), V, K)
    return dict(value=1.0 / full_iter_s, unit=UNIT, cores=1, kind=kind, sample=sample, same_config=bool(n_s == N)), dt


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from scsplit_b200 import synth
    cfg, pool = workload(args.workload)
    K = cfg["K"]
    # the whole pool for up to 40 iterations (the driver asks for 20 + 5: ~3 minutes at config 3); more than that is sampled
    base, dt = time_reference(pool, K, args.steps, args.warmup, budget_s=8.7 * 40)
    V, N = pool.ref.shape
    line = {
        "impl": "reference", "metric": METRIC, "value": base["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 / base["value"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_string(args.workload, V, N, K, synth.union_nnz(pool.ref, pool.alt)),
                   "host_cores_available": os.cpu_count(), "note": "the reference is single-threaded on this path (scipy sparsetools, pandas)"},
        "cpu_baseline": base,
        "e2e": {"value": base["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "cell_snv_per_sec": base["value"] * V * N,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------ CUDA arm
def pin(a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()


def e2e_call(Engine, csr, P0_pinned, out_W, out_P, iters, device):
    """What a user of the host API does for one restart (scSplit:477-480): upload, run, read back."""
    from scsplit_b200._lib import SCS_P, SCS_W, check
    import ctypes
    with Engine(csr[0], csr[1], device=device, canonicalize=False) as eng:
        eng.em_begin(P0_pinned)
        eng.em_iterate(iters, check_conv=False)
        check(eng.lib.scs_em_get(eng.h, 0, SCS_W, out_W.ctypes.data_as(ctypes.c_void_p)))
        check(eng.lib.scs_em_get(eng.h, 0, SCS_P, out_P.ctypes.data_as(ctypes.c_void_p)))
        it, cv, ll = eng.em_status()
    return float(ll[0])


def join_comm(eng, Engine, rank, world):
    """Carry rank 0's NCCL unique id to every rank (torch.distributed) and join the library's communicator."""
    import torch
    import torch.distributed as dist
    uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        uid = torch.from_numpy(Engine.comm_unique_id()).cuda()
    dist.broadcast(uid, 0)
    eng.comm_init(uid.cpu().numpy(), rank, world)


def timed_iterations(eng, steps, warmup, local, barrier):
    """W untimed + K timed iterations on the resident pool; device time from CUDA events on the library's stream."""
    import torch
    eng.em_iterate(warmup, check_conv=False)
    eng.em_capture(check_conv=False)                     # graph capture + instantiation (~0.4 ms, executes nothing) stays outside the timed steps
    eng.set_timing(True, every=KERNEL_TIMING_EVERY)
    eng.kernel_times(reset=True)
    barrier()
    n0 = eng.launch_count()
    with ClockSampler(local) as clk:
        t0 = time.perf_counter()
        eng.em_iterate(steps, check_conv=False)          # K steps on one stream; the host only polls flags every 8th
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
    out = dict(dev_s=eng.last_ms() / 1000.0, wall=wall, launches=eng.launch_count() - n0, kt=eng.kernel_times(reset=True), clocks=clk.summary())
    eng.set_timing(False)
    barrier()
    return out


def max_over_ranks(values, world):
    import torch
    import torch.distributed as dist
    t = torch.tensor(values, dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(x) for x in t]


def cells_leg(args, rank, world, local, barrier):
    """BASELINE config 4 with the cells sharded over the ranks and the statistics all-reduced by NCCL every iteration
    (world = 1: the whole pool on one GPU).  Returns the dict printed under `cells_c4` (rank 0) or None."""
    import torch
    import torch.distributed as dist
    from scsplit_b200 import synth
    from scsplit_b200.engine import Engine
    name = "c4"
    cfg = synth.CONFIGS[name]
    N_total, V, K = cfg["N"], cfg["V"], cfg["K"]
    per = (N_total + world - 1) // world
    lo, hi = rank * per, min(N_total, (rank + 1) * per)
    steps, warmup = min(args.steps, 40), min(args.warmup, 8)
    # phase 1 (host memory and time: ~75 M entries per 25k cells): every rank reports whether it got its shard up
    eng, err, gen_s = None, None, 0.0
    try:
        import psutil
        need = 45e9 * (hi - lo) / N_total * world        # ~45 GB of host memory in flight for the whole pool on one box
        if psutil.virtual_memory().available < need:
            raise MemoryError("host has %.0f GB available, the config-4 shards of this box need ~%.0f GB" % (psutil.virtual_memory().available / 1e9, need / 1e9))
        t0 = time.perf_counter()
        workers = max(1, (os.cpu_count() or 1) // world - 1)
        pool = synth.make_config(name, cell_lo=lo, cell_hi=hi, workers=workers)
        gen_s = time.perf_counter() - t0
        eng = Engine(pool.ref, pool.alt, device=local)
        del pool
    except Exception as exc:
        err = "%s: %s" % (type(exc).__name__, exc)
    ok = torch.tensor([0 if err else 1], dtype=torch.int64, device="cuda")
    if world > 1:
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    if int(ok[0]) == 0:
        if eng is not None:
            eng.close()
        return {"skipped": err or "another rank could not set up its shard"} if rank == 0 else None
    with eng:
        if world > 1:
            join_comm(eng, Engine, rank, world)
        info = eng.layout_info()
        nnzU = torch.tensor([info["nnz_union"]], dtype=torch.int64, device="cuda")
        if world > 1:
            dist.all_reduce(nnzU)
        lab = random_seed_labels(N_total, K, 0)[lo:hi]
        eng.em_begin(eng.seed_af(lab, K), max_iter=steps + warmup + 1)
        einfo = eng.em_info()
        res = timed_iterations(eng, steps, warmup, local, barrier)
        it, cv, ll = eng.em_status()
    dev_s, wall = max_over_ranks([res["dev_s"], res["wall"]], world)
    if rank != 0:
        return None
    kt = res["kt"]
    eb, mb = algorithmic_bytes(V, N_total, K, int(nnzU[0]))
    peak, _ = load_peaks()
    return {
        "workload": workload_string(name, V, N_total, K, int(nnzU[0])),
        "parallelism": "cells sharded over %d GPUs (%d cells each); per iteration NCCL all-reduces the posterior tail (overlapping the M-step kernels) and the M-step sums S, %d collectives" % (
            world, per, round(kt["allreduce_n"] / max(kt["m_n"], 1))) if world > 1 else "single GPU, whole pool",
        "n_gpus": world, "steps": steps, "warmup": warmup, "value": steps / dev_s, "unit": UNIT, "ms_per_step": 1000.0 * dev_s / steps,
        "scaling": "strong", "e_us": kt["e_us"], "m_us": kt["m_us"], "bayes_us": kt["bayes_us"], "allreduce_us": kt["allreduce_us"],
        "allreduce_us_is": "elapsed on the communication stream per iteration (includes the time a collective waits for SMs behind the kernels it overlaps)",
        "allreduce_pieces_timed": kt["allreduce_n"], "allreduce_doubles": 2 * V * (K + (K & 1)) + (K + (K & 1)) + 2,
        "estep": "fixed point, %d values per table row, F=%d, form: %s" % (2 * einfo["q_lanes"] + 1, einfo["q_frac_bits"], einfo["q_form"]) if einfo["q_lanes"] else "fp64",
        "hbm_frac_of_iteration": (eb + mb) / world / (dev_s / steps) / 1e9 / peak,
        "last_ll": float(ll[0]), "host_generation_s": gen_s, "clocks": res["clocks"],
    }


def regime_times(Engine, pool, P0, local):
    """One restart the way `scSplit run` runs it: from the seed to the reference's stop rule (scSplit:155-161)."""
    out = {}
    with Engine(pool.ref, pool.alt, device=local) as eng:
        eng.em_begin(P0)
        eng.em_iterate(2, check_conv=False)               # the soft-posterior start: tiled M-step, not the sparse kernel
        out["soft_ms_per_step"] = eng.last_ms() / 2
        eng.em_begin(P0)
        t0 = time.perf_counter()
        eng.em_run()
        wall = time.perf_counter() - t0
        it, cv, ll = eng.em_status()
        out["whole_run"] = {"iterations": int(it[0]), "converged": bool(cv[0]), "device_ms": eng.last_ms(), "wall_ms": 1000.0 * wall,
                            "iter_per_s": float(it[0]) / (eng.last_ms() / 1000.0), "ll": float(ll[0])}
    return out


def small_config_times(Engine, local):
    """Configs 1 and 2: iterations that are a few microseconds of HBM time, i.e. launch-bound."""
    from scsplit_b200 import synth
    peak, _ = load_peaks()
    out = {}
    for name in ("c1", "c2"):
        cfg, pool = workload(name)
        K = cfg["K"]
        V, N = pool.ref.shape
        with Engine(pool.ref, pool.alt, device=local) as eng:
            eng.em_begin(eng.seed_af(random_seed_labels(N, K, 0), K), max_iter=400)
            eng.em_iterate(20, check_conv=False)
            n0 = eng.launch_count()
            eng.em_iterate(200, check_conv=False)
            eb, mb = algorithmic_bytes(V, N, K, synth.union_nnz(pool.ref, pool.alt))
            out[name] = {"us_per_iteration": eng.last_ms() * 1000.0 / 200, "hbm_roofline_us": (eb + mb) / peak / 1e3,
                         "launches_per_iteration": (eng.launch_count() - n0) / 200.0, "shape": "%d x %d, K=%d" % (V, N, K)}
    return out


def run_cuda_arm(args):
    import torch
    import torch.distributed as dist
    from scipy.sparse import csr_matrix
    from scsplit_b200 import _lib, synth
    from scsplit_b200.engine import Engine

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if _lib.device_count() < 1:
        raise SystemExit("bench.py needs a CUDA device: scsplit_b200 has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        # NCCL prints its version banner on stdout while the communicator comes up: send fd 1 to stderr meanwhile so that
        # stdout carries exactly one JSON line
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    cells_mode = args.mode == "cells" and world > 1
    if cells_mode:
        cfg = synth.CONFIGS[args.workload]
        per = (cfg["N"] + world - 1) // world
        lo, hi = rank * per, min(cfg["N"], (rank + 1) * per)
        cfg, pool = workload(args.workload, lo, hi)
    else:
        cfg, pool = workload(args.workload)
    K = cfg["K"]
    V, n_local = pool.ref.shape
    N_total = cfg["N"]
    nnzU_local = synth.union_nnz(pool.ref, pool.alt)

    eng = Engine(pool.ref, pool.alt, device=local)
    if cells_mode:
        join_comm(eng, Engine, rank, world)
    info = eng.layout_info()
    assert info["nnz_union"] == nnzU_local
    nnzU_total = nnzU_local
    if cells_mode:
        tt = torch.tensor([nnzU_local], dtype=torch.int64, device="cuda")
        dist.all_reduce(tt)
        nnzU_total = int(tt[0])
    # restart-sharded: every rank starts from its own initialisation; cell-sharded: same seed, local slice of labels
    lab = random_seed_labels(N_total, K, 0 if cells_mode else rank)
    if cells_mode:
        lab = lab[lo:hi]
    P0 = eng.seed_af(lab, K)
    eng.em_begin(P0, max_iter=max(args.steps + args.warmup + 1, 300))
    einfo = eng.em_info()
    res = timed_iterations(eng, args.steps, args.warmup, local, barrier)
    dev_s, wall = max_over_ranks([res["dev_s"], res["wall"]], world)
    launches, kt, clocks = res["launches"], res["kt"], res["clocks"]
    it, cv, ll = eng.em_status()

    # ---- end to end through the host API, pinned host buffers (rank-local; restart-sharded => N independent calls)
    e2e = None
    if not cells_mode:
        ref, alt = pool.ref, pool.alt
        pr = csr_matrix((pin(ref.data.astype(np.int16)), pin(ref.indices.astype(np.int32)), pin(ref.indptr.astype(np.int64))), shape=ref.shape)
        pa = csr_matrix((pin(alt.data.astype(np.int16)), pin(alt.indices.astype(np.int32)), pin(alt.indptr.astype(np.int64))), shape=alt.shape)
        pr.has_canonical_format = True
        pa.has_canonical_format = True
        P0p, outW, outP = pin(P0), pin(np.empty((n_local, K))), pin(np.empty((V, K)))
        eng.em_end()
        n_e2e = 8
        e2e_call(Engine, (pr, pa), P0p, outW, outP, E2E_ITERS, local)        # warm-up (allocator, lazy module load)
        barrier()
        t0 = time.perf_counter()
        for _ in range(n_e2e):
            e2e_call(Engine, (pr, pa), P0p, outW, outP, E2E_ITERS, local)
        torch.cuda.synchronize()
        e2e_s = (time.perf_counter() - t0) / n_e2e
        t0 = time.perf_counter()
        for _ in range(n_e2e):
            e2e_call(Engine, (pr, pa), P0p, outW, outP, 1, local)
        torch.cuda.synchronize()
        strict_s = (time.perf_counter() - t0) / n_e2e
        e2e_s, strict_s = max_over_ranks([e2e_s, strict_s], world)
        h2d = sum(a.nbytes for m in (pr, pa) for a in (m.data, m.indices, m.indptr)) + P0p.nbytes
        d2h = outW.nbytes + outP.nbytes + 16
        e2e = {"value": world * E2E_ITERS / e2e_s, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
               "iters_per_call": E2E_ITERS, "ms_per_call": 1000.0 * e2e_s,
               "what": "Engine(ref,alt) from pinned host CSR (H2D + device layout build) + em_begin(P0) + %d iterations + D2H of P_s_c, model_af, LL; one call = one restart (scSplit:477-480)" % E2E_ITERS,
               "strict_one_iteration_per_call": {"value": world * 1.0 / strict_s, "unit": UNIT, "ms_per_call": 1000.0 * strict_s}}
    eng.close()

    # ---- N = 1 extras: how a real run starts and ends, the launch-bound small configurations
    extras = {}
    if world == 1 and rank == 0 and not args.quick:
        extras["regimes"] = dict(saturated_ms_per_step=1000.0 * dev_s / args.steps, **regime_times(Engine, pool, P0, local))
        extras["small_configs"] = small_config_times(Engine, local)
        try:     # the whole `scSplit run` CLI, CSV files in, result files out (tools/cli_e2e.py), next to the recorded reference runs
            import importlib.util
            import io
            import contextlib
            spec = importlib.util.spec_from_file_location("cli_e2e", os.path.join(ROOT, "tools", "cli_e2e.py"))
            mod = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(mod)
            with contextlib.redirect_stdout(io.StringIO()):          # ("Repeat n was chosen." belongs to the CLI, not to this line)
                r = mod.run(argparse.Namespace(workload=args.workload, ems=30, seed=2003, keep=None))
            extras["cli_e2e"] = {k: r[k] for k in ("workload", "wall_s", "phases_s", "em_iterations", "csv_mb", "assigned_cells", "agreement_with_truth")}
            extras["cli_e2e"]["reference"] = "the reference CLI on the same CSVs (one core of the build container): 1099 s with -e 2 (profiles/r02_ref_cli_c3.json; this CLI 1.3 s, identical partition, profiles/r02_cli_e2e_c3_e2.json) and 2443 s with -e 8 (profiles/r02_ref_cli_c3_e8.json; same optimum, LL -11323037.636696128, same cluster sizes); with -e 30 it was not run (~7400 s extrapolated from those two: 224 s per restart + 650 s)"
            if "reference_wall_s" in r:      # tests/golden/cli_<workload>_e30.npz: the reference CLI's own -e 30 run on the same CSVs and seed
                for k in ("reference_wall_s", "speedup_vs_reference_cli", "reference_assigned_cells", "cells_assigned_by_both", "agreement_with_reference"):
                    extras["cli_e2e"][k] = r[k]
                extras["cli_e2e"]["reference"] = "the reference CLI on the same CSVs and seed, -e 30, one core of the build container (oracle/make_c3_cli_reference.py -e 30, profiles/r02_ref_cli_c3_e30.json): 8254 s = `reference_wall_s` of this object, partition in tests/golden/cli_c3_e30.npz; with -e 2 it took 1099 s, with -e 8 2443 s (profiles/r02_ref_cli_c3.json, _e8.json)"
            if os.path.exists("scSplit_dist_variants.txt"):
                os.remove("scSplit_dist_variants.txt")
        except Exception as exc:
            extras["cli_e2e"] = {"error": "%s: %s" % (type(exc).__name__, exc)}

    # ---- second leg: config 4, cells sharded over the ranks, NCCL all-reduce per iteration (N = 1: the anchor)
    cells = None
    if not cells_mode and not args.no_cells_leg and not args.quick:
        del pool
        cells = cells_leg(args, rank, world, local, barrier)   # (set-up failures, e.g. host memory, come back as {"skipped": why})

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peak, peak_src = load_peaks()
    eb, mb = algorithmic_bytes(V, n_local, K, nnzU_local)
    dom = "E" if kt["e_us"] >= kt["m_us"] else "M"
    dom_bytes, dom_us = (eb, kt["e_us"]) if dom == "E" else (mb, kt["m_us"])
    traffic, traffic_src = None, None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath) and args.workload == "c3" and not cells_mode:   # the capture is of this configuration only
        with open(tpath) as f:
            tj = json.load(f)
        traffic, traffic_src = tj.get(dom), "static: %s (not measured in this run)" % tj.get("source", "profiles/traffic.json")
    iters_total = world * args.steps if not cells_mode else args.steps
    value = iters_total / dev_s
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1000.0 * dev_s / args.steps, "higher_is_better": True,
        "scaling": "strong" if cells_mode else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {
            "workload": workload_string(args.workload, V, N_total, K, nnzU_total),
            "parallelism": ("cells sharded over %d GPUs, NCCL all-reduce of 2VK+K+1 doubles per iteration" % world) if cells_mode
            else ("%d independent restarts, one per GPU, no collective" % world if world > 1 else "single GPU, one restart"),
            "l2": "inputs larger than L2: two entry streams of %.0f MB each are read per step vs %.0f MB of L2; no flush" % (
                info["light"] * 4 / 1e6 + info["heavy"] * 8 / 1e6, L2_BYTES / 1e6),
            "init": "seed P from 3 random cells per state through scs_seed_af (the reference's trim/PCA/KMeans initialiser is host code outside the step)",
            "regime": "saturated posteriors (after %d warm-up iterations): E-step + sparse M-step; the soft start of a run is under `regimes`" % args.warmup,
            "estep": ("log tables in %d-bit fixed point (F=%d), %d-byte rows, exact integer accumulation; fp64 per restart when a table does not fit" % (
                6 + einfo["q_frac_bits"], einfo["q_frac_bits"], 64 if einfo["q_form"].startswith("lock-step,") else 16 * einfo["q_lanes"]) + "; form: " + einfo["q_form"]) if einfo["q_lanes"] else "fp64 tables and accumulation",
        },
        "cell_snv_per_sec": value * V * N_total,
        "wall_ms_per_step": 1000.0 * wall / args.steps,
        "clocks": clocks,
        "gpu_launches": int(launches),
        "last_ll": float(ll[0]),
        "roofline": {"bound": "hbm", "kernel": ("E-step sums (k_estep_lk)" if einfo["q_form"].startswith("lock-step") else "E-step sums (k_estep_q + k_combine_q)") if dom == "E" and einfo["q_lanes"] else "E-step SpMM" if dom == "E" else "M-step sums",
                     "achieved": dom_bytes / (dom_us * 1e-6) / 1e9,
                     "peak": peak, "unit": "GB/s", "frac": dom_bytes / (dom_us * 1e-6) / 1e9 / peak, "traffic": traffic, "traffic_source": traffic_src,
                     "peak_source": peak_src, "algorithmic_bytes_per_launch": int(dom_bytes), "avg_launch_us": dom_us,
                     "e_step": {"us": kt["e_us"], "bytes": int(eb), "GBps": eb / max(kt["e_us"], 1e-9) / 1e3, "launches_timed": kt["e_n"],
                                "what": ("k_estep_lk: a lane per entry, 32 units in lock step, integer totals by RED.ADD.64 (k_bayes converts them to L)" if einfo["q_form"].startswith("lock-step") else
                                         "all launches that produce L: k_estep_q (per-(tile, cell) integer sums) + k_combine_q (fold over the tiles)") if einfo["q_lanes"] else "k_spmm_tiled + k_spmm_combine",
                                "bayes_us": kt["bayes_us"],
                                "with_bayes": {"us": kt["e_us"] + kt["bayes_us"], "frac": eb / max(kt["e_us"] + kt["bayes_us"], 1e-9) / 1e3 / peak,
                                               "what": "the whole E-step (scSplit:165-188): sums + k_bayes (posteriors, LL, column mass)"}},
                     "m_step": {"us": kt["m_us"], "bytes": int(mb), "GBps": mb / max(kt["m_us"], 1e-9) / 1e3, "launches_timed": kt["m_n"]},
                     "iteration_frac": (eb + mb) / (dev_s / args.steps) / 1e9 / peak},
    }
    if e2e:
        line["e2e"] = e2e
    line.update(extras)
    if cells is not None:
        line["cells_c4"] = cells
    if world == 1 and not args.no_cpu_baseline and not args.quick:
        _, pool = workload(args.workload)
        base, _ = time_reference(pool, K, 2, 1, budget_s=30.0)
        line["cpu_baseline"] = base
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=500)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--workload", default="c3", choices=["c1", "c2", "c3", "c4", "c5"])
    ap.add_argument("--mode", default="restarts", choices=["restarts", "cells"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-cells-leg", action="store_true", help="skip the second leg (config 4, cells sharded over the GPUs)")
    ap.add_argument("--quick", action="store_true", help="main leg and e2e only (kernel development)")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_cuda_arm(args)


if __name__ == "__main__":
    main()
