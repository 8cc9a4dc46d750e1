#!/usr/bin/env python
"""bench.py — option PDE solves/sec on BASELINE.json's configs[1]:
a batch of American puts (varied strike / vol / T / r, seeded as SURVEY.md
§8(d)), 1-D Rannacher(CN) + PenaltyMethod, 2049 nodes x 1000 steps, FP64.

    python bench.py --gpus N --steps K --warmup W [--impl reference]
    torchrun ... bench.py --gpus N ...      (one rank per GPU, N > 1)

    python bench.py --config {1,3,4,5,hjb,er,oc}      (1 = BASELINE.json configs[1], the default and the
                                                 driver's contract; 3 / 4 / 5 = configs[2] / [3] / [4],
                                                 hjb = the policy-iteration example: one GPU each)

A "step" is one pass of the hot path over one batch (one qpde_bs1d_plan_run of
all contracts).  `value` = contracts / device time with the batch resident in
HBM; `e2e` = the same through qpde_bs1d_sweep with HOST buffers (H2D + run +
D2H inside the timed region).  Contracts shard across ranks with no data-path
collective (weak scaling: each rank owns --contracts contracts).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_NODES_REFINE = 6          # Axis::special (33 ticks) refined 6x -> 2049 nodes
N_STEPS = 1000
ALG_BYTES_PER_NODE_SOLVE = 56   # 7 FP64 vectors per inner solve


def contract_parameters(n, rank=0, world=1):
    """default_rng(1234); K, sigma, T, r drawn in this
    order for the whole job (n contracts per rank x world ranks); rank k sweeps
    its contiguous slice (quantpde_b200.shard.contract_slice)."""
    from quantpde_b200.shard import contract_slice
    total = n * world
    rng = np.random.default_rng(1234)
    K = rng.uniform(50, 150, total)
    vol = rng.uniform(0.10, 0.60, total)
    T = rng.uniform(0.25, 2.0, total)
    r = rng.uniform(0.01, 0.08, total)
    b, e = contract_slice(total, rank, world)
    return K[b:e], vol[b:e], T[b:e], r[b:e]


def parity_sample_indices(n_total, K, vol, T, r, n_sample=256, wave=148):
    """A stratified sample of the batch for the node-level parity check:
    the first and the last contracts (first / last wave of CTAs), the extremes of every parameter,
    and an evenly spaced remainder.  Deterministic."""
    n_sample = min(n_sample, n_total)
    picks = []
    edge = min(wave // 4, max(1, n_sample // 8))
    picks += list(range(edge)) + list(range(n_total - edge, n_total))
    for arr in (K, vol, T, r):
        order = np.argsort(arr, kind="stable")
        picks += [int(order[0]), int(order[1 % n_total]), int(order[-1]), int(order[-2 % n_total])]
    seen, out = set(), []
    for i in picks:
        if i not in seen and 0 <= i < n_total:
            seen.add(i); out.append(i)
    k = 0
    rest = np.linspace(0, n_total - 1, num=4 * n_sample).astype(np.int64)
    while len(out) < n_sample and k < len(rest):
        i = int(rest[k]); k += 1
        if i not in seen:
            seen.add(i); out.append(i)
    return np.array(sorted(out[:n_sample]), dtype=np.int64)


def _oracle_one(args):
    """One contract of configs[1] on the oracle (the CPU checker): V, per-step inner counts, mask."""
    K, vol, T, r = args
    from oracle import pyoracle as po
    o = po.american_put(float(K), float(r), float(vol), float(T), N_STEPS, N_NODES_REFINE)
    return o["V"], o["inner"], o["mask"], o["S"], o["n_steps"]


def parity_sample(idx, K, vol, T, r, V, inner_total, n_steps, active=None):
    """Node-for-node comparison of the sampled contracts with the oracle (which is bit-identical to the
    reference on the golden cases): worst relative error (|a-b| / max(1,|a|,|b|)), inner-iteration totals,
    step counts, and — when the active sets were fetched — mismatches outside / inside the tie band
    (DESIGN.md §6)."""
    from concurrent.futures import ProcessPoolExecutor
    from oracle import pyoracle as po
    po.lib()
    jobs = [(K[i], vol[i], T[i], r[i]) for i in idx]
    with ProcessPoolExecutor(max_workers=os.cpu_count() or 1) as ex:
        ref = list(ex.map(_oracle_one, jobs, chunksize=4))
    worst, inner_equal, steps_equal, hard, ties, tied_contracts = 0.0, 0, 0, 0, 0, 0
    eps = np.finfo(np.float64).eps
    for k, i in enumerate(idx):
        Vr, inr, mr, Sr, ns = ref[k]
        err = np.abs(V[k] - Vr) / np.maximum(1.0, np.maximum(np.abs(V[k]), np.abs(Vr)))
        worst = max(worst, float(err.max()))
        inner_equal += int(int(inr.sum()) == int(inner_total[k]))
        steps_equal += int(ns == int(n_steps[k]))
        if active is not None:
            obst = np.maximum(K[i] - Sr, 0.0)
            diff = active[k].astype(bool) != mr.astype(bool)
            tie = np.abs(Vr - obst) <= 64 * eps * np.maximum(1.0, np.abs(obst))
            hard += int(np.sum(diff & ~tie)); t = int(np.sum(diff & tie)); ties += t
            tied_contracts += int(t > 0)
    out = {"n": int(len(idx)), "worst_rel_err": worst, "tolerance": 1e-9,
           "inner_totals_equal": inner_equal, "step_counts_equal": steps_equal,
           "against": "oracle/qpde_oracle.c (bit-identical to the reference headers on tests/golden)"}
    if active is not None:
        out.update({"active_set_mismatches": hard, "ties": ties, "contracts_with_ties": tied_contracts})
    return out


SPECIAL = np.array([0., .1, .2, .3, .4, .5, .6, .7, .80, .84, .88, .92, .94, .96, .98, 1.,
                    1.02, 1.04, 1.06, 1.08, 1.10, 1.14, 1.18, 1.23, 1.30, 1.40, 1.50, 1.75,
                    2.25, 3.00, 7.50, 20., 100.])   # Axis::special, Core/Axis.hpp:445-459


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def load_traffic(kernel, contracts):
    """dram__bytes_read + write per launch of the dominant kernel, from the committed ncu
    capture (profiles/traffic.json), scaled to this launch's contract count."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(path):
        d = json.load(open(path)).get(kernel)
        if d:
            return d["bytes_per_contract"] * contracts
    return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                 "--format=csv,noheader,nounits", "-lms", "200"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax = float(f[2])
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": smax,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------
# CPU arm: the reference's own implementation of the path on the host cores
# ---------------------------------------------------------------------------

def _cpu_one(args):
    """One contract on one core.  kind 'reference' = oracle/_ref/qpde_ref (the
    unmodified reference headers on oracle/shim's linear algebra), else the C
    restatement oracle/qpde_oracle.c."""
    kind, K, vol, T, r = args
    from oracle import pyoracle as po
    t0 = time.perf_counter()
    if kind == "reference":
        out = po.run_ref("vanilla", american=1, steps=N_STEPS, refine=N_NODES_REFINE,
                         K=float(K), r=float(r), vol=float(vol), T=float(T))
        inner = int(out["inner"].sum())
        secs = out["seconds"]            # steady_clock around stepper.solve (Results.hpp:103-108)
    else:
        out = po.american_put(float(K), float(r), float(vol), float(T), N_STEPS, N_NODES_REFINE)
        inner = int(out["inner"].sum())
        secs = time.perf_counter() - t0
    return secs, inner


def cpu_sample_indices(n_sample, pool=65536):
    """The contracts the CPU arm times: a stratified sample of the 65,536-contract workload
    (the same strata as the parity sample)."""
    K, vol, T, r = contract_parameters(pool)
    return parity_sample_indices(pool, K, vol, T, r, n_sample), (K, vol, T, r)


def cpu_baseline(n_sample, kind=None, offset=0):
    """Times a bounded, stratified sample of the workload on all host cores (one contract
    per worker process, as BASELINE.md §3 prescribes for the single-threaded
    reference).  Returns solves/s for the whole host.  `offset` rotates through a pool of 256
    stratified contracts so that successive calls time different ones."""
    from concurrent.futures import ProcessPoolExecutor
    from oracle import pyoracle as po
    po.lib()
    if kind is None:
        kind = "reference" if po.have_ref() else "port"
    cores = os.cpu_count() or 1
    idx, (K, vol, T, r) = cpu_sample_indices(max(256, n_sample))
    take = [int(idx[(offset + k) % len(idx)]) for k in range(n_sample)]
    jobs = [(kind, K[i], vol[i], T[i], r[i]) for i in take]
    t0 = time.perf_counter()
    with ProcessPoolExecutor(max_workers=cores) as ex:
        results = list(ex.map(_cpu_one, jobs))
    wall = time.perf_counter() - t0
    per_contract = float(np.mean([s for s, _ in results]))
    return {"value": n_sample / wall, "unit": "solves/s", "cores": cores, "kind": kind,
            "solves_per_s_per_core": 1.0 / per_contract,
            "sample": "%d contracts drawn from the seeded 65,536-contract workload (first / last contracts, extremes "
                      "of K, sigma, T, r, evenly spaced rest; 2049 nodes x 1000 steps), one per worker process on %d "
                      "cores; %.3f s per contract single-threaded; LU is oracle/shim's stand-in, not Eigen's"
                      % (n_sample, cores, per_contract),
            "seconds_per_contract_1core": per_contract, "wall_s": wall}


def reference_seconds(job):
    """(mode, kwargs) -> wall seconds of one run of the reference's CPU path (oracle/_ref).  The
    cpu_baseline leg of the other configs (tools/bench_configs.py) goes through here: bench.py is
    the one place outside tests/ that executes anything under oracle/."""
    from oracle import pyoracle as po
    mode, kw = job
    return po.run_ref(mode, **kw)["seconds"]


# ---------------------------------------------------------------------------

def run_reference_arm(args, rank, world):
    if rank != 0:
        return
    if args.config != "1":
        return run_reference_arm_other(args)
    from oracle import pyoracle as po
    po.lib()
    cores = os.cpu_count() or 1
    # per step: 4 contracts per core of the stratified pool of 256 (3 - 4 s of wall per step), a
    # different slice every step, so that 20 steps cover the pool several times
    n_sample = args.cpu_sample or 4 * cores
    for _ in range(args.warmup):
        cpu_baseline(min(cores, n_sample))
    vals, walls, per_core = [], [], []
    for k in range(args.steps):
        b = cpu_baseline(n_sample, offset=k * n_sample)
        vals.append(b["value"]); walls.append(b["wall_s"]); per_core.append(b["solves_per_s_per_core"])
    value = float(n_sample * len(walls) / sum(walls))
    line = {
        "impl": "reference", "metric": "option PDE solves/sec (N=2049, 1000 CN steps)",
        "value": value, "unit": "solves/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * float(np.mean(walls)),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": "american_put_batch_2049x1000 (BASELINE.json configs[1])",
                   "contracts_per_step": n_sample, "nodes": 2049, "time_steps": N_STEPS,
                   "scheme": "rannacher+penalty"},
        "cpu_baseline": {k: b[k] for k in ("value", "unit", "cores", "kind", "sample", "solves_per_s_per_core")},
        "e2e": {"value": value, "unit": "solves/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    line["cpu_baseline"]["value"] = value
    line["cpu_baseline"]["solves_per_s_per_core"] = float(np.mean(per_core))
    print(json.dumps(line))


# ---------------------------------------------------------------------------
# The other BASELINE.json configs (one GPU): configs[2] Bermudan local-vol, configs[3] Merton jump,
# configs[4] GMWB, and the HJB policy-iteration example.  Same JSON line as configs[1].
# ---------------------------------------------------------------------------

TUTORIAL = np.array([0., 10., 20., 30., 40., 50., 60., 70., 75., 80., 84., 88., 92., 94., 96.,
                     98., 100., 102., 104., 106., 108., 110., 114., 118., 123., 130., 140.,
                     150., 175., 225., 300., 750., 2000., 10000.])          # examples/tutorial.cpp:41-56
GMWB_W = np.array([0., 5., 10., 15., 20., 25., 30., 35., 40., 45., 50., 55., 60., 65., 70., 72.5, 75., 77.5, 80.,
                   82., 84., 86., 88., 90., 91., 92., 93., 94., 95., 96., 97., 98., 99., 100., 101., 102., 103.,
                   104., 105., 106., 107., 108., 109., 110., 112., 114., 116., 118., 120., 123., 126., 130., 135.,
                   140., 145., 150., 160., 175., 200., 225., 250., 300., 500., 750., 1000.])   # examples/hjbqvi/gmwb.cpp:92-105


def _jump_setup_job(args):
    from quantpde_b200 import capi
    S, params = args
    return capi.jump_setup(S, "lognormal", params)


def _oracle_job(job):
    """One contract of configs 3 / 4 / hjb on the oracle: node values (and inner counts)."""
    from oracle import pyoracle as po
    mode, kw = job
    if mode == "bermudan":
        o = po.bermudan_put(kw["K"], kw["r"], kw["alpha"], 1.0, kw["steps"], 10, 6, scheme=kw["scheme"])
    elif mode == "jump":
        o = po.jump_call(kw["K"], 0.05, kw["vol"], 0.25, 0.1, kw["steps"], 6, params=(kw["mu"], kw["sd"]))
    else:
        o = po.hjb_straddle(kw["K"], kw["r_l"], kw["r_b"], kw["vol"], 1.0, kw["steps"], 6, long_position=False, scheme="bdf2")
    return o["V"], o.get("inner")


def _ref_job(job):
    return reference_seconds(job)


def _other_config_problem(config, Cn):
    """(label, specs-by-variant builder inputs, oracle jobs, reference jobs, flops / bytes model) of one config."""
    rng = np.random.default_rng(1234)
    if config == "3":
        Cn = Cn or 16384
        alpha = rng.uniform(10, 30, Cn); r = rng.uniform(.01, .08, Cn); K = rng.uniform(80, 120, Cn)
        return dict(Cn=Cn, alpha=alpha, r=r, K=K, steps=6400, N=2113)
    if config == "4":
        Cn = Cn or 4096
        vol = rng.uniform(.1, .3, Cn); K = rng.uniform(80, 120, Cn)
        pairs = [(-0.1 + 0.02 * (i - 4), 0.45 + 0.01 * (i - 4)) for i in range(8)]
        which = rng.integers(0, 8, Cn)
        return dict(Cn=Cn, vol=vol, K=K, pairs=pairs, which=which, steps=768, N=2113, n_fft=4096)
    Cn = Cn or 8192
    K = rng.uniform(60, 140, Cn); vol = rng.uniform(.15, .5, Cn)
    r_l = rng.uniform(.01, .04, Cn); r_b = r_l + rng.uniform(.005, .04, Cn)
    return dict(Cn=Cn, K=K, vol=vol, r_l=r_l, r_b=r_b, steps=768, N=2049)


def _rel(a, b):
    return float(np.max(np.abs(a - b) / np.maximum(1.0, np.maximum(np.abs(a), np.abs(b)))))


def run_other_config(args):
    import torch
    from concurrent.futures import ProcessPoolExecutor
    from quantpde_b200 import capi
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    sys.stdout.flush()
    result_fd = os.dup(1)
    os.dup2(2, 1)
    torch.cuda.set_device(0)
    handle = capi.Handle(0)
    stream = torch.cuda.ExternalStream(handle.stream, device=torch.device("cuda", 0))
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device="cuda")
    cores = os.cpu_count() or 1
    hbm_peak, hbm_src = load_peaks()
    fp64_peak = handle.fp64_peak_tflops()
    n_par = 8

    def timed(spec):
        """--warmup untimed runs, then --steps runs timed with CUDA events on the library's stream."""
        with capi.Plan(handle, spec) as plan:
            for _ in range(max(1, args.warmup)):
                plan.run(); handle.synchronize()
            l0 = handle.launches
            ms = []
            for _ in range(args.steps):
                flush.fill_(1.0); torch.cuda.synchronize()
                e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
                e0.record(stream); plan.run(); e1.record(stream); e1.synchronize()
                ms.append(e0.elapsed_time(e1))
            launches = handle.launches - l0
            res = plan.fetch(capi.OUT_STATUS | capi.OUT_STEPS | capi.OUT_INNER_TOTAL | capi.OUT_VALUE)
            return float(np.mean(ms)), res, plan.kernel, launches

    def e2e_of(spec, Cn, N, h2d):
        r2 = capi.ResultBuffers(Cn, N, capi.sweep_max_steps(spec), capi.OUT_V | capi.OUT_VALUE | capi.OUT_STATUS, pinned=True)
        ts = []
        for i in range(1 + args.steps):
            t0 = time.perf_counter()
            capi.sweep(handle, spec, outputs=capi.OUT_V | capi.OUT_VALUE | capi.OUT_STATUS, result=r2)
            torch.cuda.synchronize()
            if i > 0:
                ts.append(time.perf_counter() - t0)
        d2h = r2.V.nbytes + r2.value.nbytes + r2.status.nbytes
        V = r2.V[:n_par].copy()
        r2.close()
        return {"value": Cn / float(np.mean(ts)), "unit": "solves/s", "h2d_bytes_per_step": int(h2d),
                "d2h_bytes_per_step": int(d2h), "ms_per_step": 1e3 * float(np.mean(ts)),
                "api": "qpde_bs1d_sweep (host pointers in, V[C][N] + value + status out into page-locked host buffers)"}, V

    def cpu_ref(jobs):
        from oracle import pyoracle as po
        if not po.have_ref():
            return None
        t0 = time.perf_counter()
        with ProcessPoolExecutor(max_workers=cores) as ex:
            secs = list(ex.map(_ref_job, jobs))
        wall = time.perf_counter() - t0
        return {"value": len(jobs) / wall, "unit": "solves/s", "cores": cores, "kind": "reference",
                "solves_per_s_per_core": 1.0 / float(np.mean(secs)),
                "sample": "%d contracts of this workload, one per worker process on %d cores, %.3f s per contract "
                          "single-threaded (oracle/_ref/qpde_ref: the reference headers on oracle/shim's linear algebra)"
                          % (len(jobs), cores, float(np.mean(secs)))}

    sampler = ClockSampler(0)
    sampler.start()
    cfg = args.config
    line = None
    if cfg in ("3", "4", "hjb"):
        P = _other_config_problem(cfg, args.contracts)
        Cn, N, steps = P["Cn"], P["N"], P["steps"]
        variants = {}
        if cfg == "3":
            K, r, alpha = P["K"], P["r"], P["alpha"]
            axis = np.ascontiguousarray(K[:, None] / 100.0 * TUTORIAL[None, :])
            def spec_of(scheme, sl=slice(None), outputs=capi.OUT_VALUE | capi.OUT_STATUS | capi.OUT_STEPS | capi.OUT_INNER_TOTAL):
                return capi.BatchSpec(axis=axis[sl], refine=6, r=r[sl], sigma=alpha[sl], q=0., T=1., dt=1. / steps, strike=K[sl],
                                      scheme=scheme, payoff="put", vol_model="inverse_s", n_exercise=10, exercise="k_minus_s",
                                      kernel=args.kernel, outputs=outputs)
            for scheme in ("bdf2", "rannacher"):
                ms, res, kern, launches = timed(spec_of(scheme))
                variants[scheme] = dict(ms=ms, kernel=kern, launches=launches, steps_taken=int(res.n_steps[0]),
                                        not_ok=int((res.status != 0).sum()), solve_steps=int(res.n_steps.astype(np.int64).sum()))
            head = "bdf2"
            spec = spec_of(head, outputs=capi.OUT_ALL)
            ojobs = [("bermudan", dict(K=float(K[c]), r=float(r[c]), alpha=float(alpha[c]), steps=steps, scheme=head)) for c in range(n_par)]
            rjobs = [("bermudan", dict(steps=steps, refine=6, scheme=head, K=float(K[c]), r=float(r[c]), alpha=float(alpha[c]),
                                       T=1.0, E=10)) for c in range(2 * cores)]
            h2d = axis.nbytes + 6 * 8 * Cn
            workload = "bermudan_put_local_vol_2113x6400x10_events (BASELINE.json configs[2])"
            metric = "option PDE solves/sec (Bermudan put, local vol, N=2113, 6400 BDF2 steps, 10 exercise dates)"
            v = variants[head]
            alg = 40.0 * N * v["solve_steps"]                    # 40 N bytes per linear step
            roofline = {"bound": "hbm", "achieved": alg / (v["ms"] / 1e3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                        "frac": alg / (v["ms"] / 1e3) / 1e9 / hbm_peak, "traffic": None, "peak_source": hbm_src,
                        "algorithmic_bytes_per_launch": alg,
                        "model": "40 B x nodes x steps (SURVEY.md 8d: lo, di, up, V^n read, V^{n+1} written); the RESIDENT sweep "
                                 "keeps them on chip, so frac > 1 means it has left the HBM roofline"}
        elif cfg == "4":
            K, vol, pairs, which = P["K"], P["vol"], P["pairs"], P["which"]
            t0 = time.perf_counter()
            axes, grids, keys = [], [], {}
            for c in range(Cn):
                axis_c = np.union1d(SPECIAL * K[c], [K[c] * 1000.0])
                S = capi.refine_axis(axis_c, 6)
                dx = (np.log(S[-2]) - np.log(S[1])) / (4096 - 1)
                axes.append(axis_c); grids.append(S)
                keys.setdefault((int(which[c]), float(dx)), c)
            reps = sorted(keys.values())
            with ProcessPoolExecutor(max_workers=cores) as ex:
                done = dict(zip(reps, ex.map(_jump_setup_job, [(grids[c], pairs[which[c]]) for c in reps])))
            jk, jx, jd, jw = [], [], [], []
            for c in range(Cn):
                x0 = np.log(grids[c][1]); dx = (np.log(grids[c][-2]) - x0) / (4096 - 1)
                _, _, kappa, w = done[keys[(int(which[c]), float(dx))]]
                jk.append(kappa); jx.append(float(x0)); jd.append(float(dx)); jw.append(w)
            setup_s = time.perf_counter() - t0
            axes = np.stack(axes); jw = np.stack(jw); jk, jx, jd = np.array(jk), np.array(jx), np.array(jd)
            def spec_of(scheme, sl=slice(None), outputs=capi.OUT_VALUE | capi.OUT_STATUS | capi.OUT_STEPS | capi.OUT_INNER_TOTAL):
                return capi.BatchSpec(axis=axes[sl], refine=6, r=0.05, sigma=vol[sl], q=0., T=.25, dt=.25 / steps, strike=K[sl],
                                      scheme="bdf1", payoff="call",
                                      jump=dict(lam=0.1, kappa=jk[sl], x0=jx[sl], dx=jd[sl], weights=jw[sl]), outputs=outputs)
            head = "bdf1"
            ms, res, kern, launches = timed(spec_of(head))
            variants[head] = dict(ms=ms, kernel="resident (jump)", launches=launches, steps_taken=int(res.n_steps[0]),
                                  not_ok=int((res.status != 0).sum()), solve_steps=int(res.n_steps.astype(np.int64).sum()),
                                  host_setup_seconds=setup_s, distinct_weight_vectors=len(reps))
            spec = spec_of(head, outputs=capi.OUT_ALL)
            ojobs = [("jump", dict(K=float(K[c]), vol=float(vol[c]), steps=steps, mu=pairs[which[c]][0], sd=pairs[which[c]][1])) for c in range(n_par)]
            rjobs = [("jump", dict(steps=steps, refine=6, density="lognormal", K=float(K[c]), r=0.05, vol=float(vol[c]), T=0.25,
                                   mu=pairs[which[c]][0], sd=pairs[which[c]][1], **{"lambda": 0.1})) for c in range(2 * cores)]
            h2d = axes.nbytes + jw.nbytes + 10 * 8 * Cn
            workload = "merton_jump_call_2113x768_nfft4096 (BASELINE.json configs[3])"
            metric = "option PDE solves/sec (Merton jump European call, N=2113, n_fft=4096, 768 BDF1 steps)"
            nfft = P["n_fft"]
            per_step = 2 * 5.0 * nfft * np.log2(nfft) + 6.0 * nfft + 3.0 * nfft + 3.0 * N + 20.0 * N
            flops = per_step * variants[head]["solve_steps"]
            tf = flops / (ms / 1e3) / 1e12
            roofline = {"bound": "fp64", "achieved": tf, "peak": fp64_peak, "unit": "TFLOP/s", "frac": tf / fp64_peak,
                        "traffic": None, "peak_source": "measured (qpde_measure_fp64_peak: independent DFMA chains)",
                        "algorithmic_flops_per_launch": flops,
                        "ncu_fp64_pipe_active_pct": 22.8,
                        "model": "per step and contract: two radix-2 FFTs of n_fft (5 n log2 n each), the spectrum product "
                                 "(6 n), two interpolation passes (3 n + 3 N) and the tridiagonal solve with its right-hand "
                                 "side (20 N); ncu_fp64_pipe_active_pct is sm__pipe_fp64_cycles_active of the committed "
                                 "capture (profiles/r1_jump_ncu_raw_subset.csv)"}
        else:
            K, vol, r_l, r_b = P["K"], P["vol"], P["r_l"], P["r_b"]
            axis = np.ascontiguousarray(K[:, None] * SPECIAL[None, :])
            ctl = np.stack([r_l, r_b], axis=1)
            def spec_of(scheme, sl=slice(None), outputs=capi.OUT_VALUE | capi.OUT_STATUS | capi.OUT_STEPS | capi.OUT_INNER_TOTAL):
                return capi.BatchSpec(axis=axis[sl], refine=6, r=0., sigma=vol[sl], q=0., T=1., dt=1. / steps, strike=K[sl],
                                      scheme="bdf2", payoff="straddle", controls=ctl[sl], kernel=args.kernel, outputs=outputs)
            head = "bdf2"
            ms, res, kern, launches = timed(spec_of(head))
            inner = int(res.inner_total.astype(np.int64).sum())
            variants[head] = dict(ms=ms, kernel=kern, launches=launches, steps_taken=int(res.n_steps[0]),
                                  not_ok=int((res.status != 0).sum()), solve_steps=inner,
                                  mean_policy_iterations=inner / float(res.n_steps.astype(np.int64).sum()))
            spec = spec_of(head, outputs=capi.OUT_ALL)
            ojobs = [("hjb", dict(K=float(K[c]), r_l=float(r_l[c]), r_b=float(r_b[c]), vol=float(vol[c]), steps=steps)) for c in range(n_par)]
            rjobs = [("hjb", dict(steps=steps, refine=6, long=0, scheme="bdf2", K=float(K[c]), r_l=float(r_l[c]), r_b=float(r_b[c]),
                                  vol=float(vol[c]), T=1.0)) for c in range(2 * cores)]
            h2d = axis.nbytes + ctl.nbytes + 6 * 8 * Cn
            workload = "hjb_unequal_rates_straddle_2049x768 (examples/unequal_borrowing_lending_rates.cpp)"
            metric = "option PDE solves/sec (HJB unequal borrowing/lending rates, policy iteration, N=2049, 768 BDF2 steps)"
            flops = (2 * 6.0 * N + 20.0 * N) * inner            # per policy iteration: n_controls SpMV rows + one solve
            tf = flops / (ms / 1e3) / 1e12
            roofline = {"bound": "fp64", "achieved": tf, "peak": fp64_peak, "unit": "TFLOP/s", "frac": tf / fp64_peak,
                        "traffic": None, "peak_source": "measured (qpde_measure_fp64_peak: independent DFMA chains)",
                        "algorithmic_flops_per_launch": flops,
                        "model": "per policy iteration and node: 2 controls x 6 flops (the candidate rows A(q) x - b(q)) + 20 for "
                                 "the tridiagonal solve and its right-hand side"}
        # ---- e2e through the one-shot call and the parity sample: the first contracts of the batch, every
        # node against the oracle
        e2e, V8 = e2e_of(spec_of(head, outputs=capi.OUT_V | capi.OUT_VALUE | capi.OUT_STATUS), Cn, N, h2d)
        with ProcessPoolExecutor(max_workers=cores) as ex:
            ref = list(ex.map(_oracle_job, ojobs))
        worst = max(_rel(V8[c], ref[c][0]) for c in range(n_par))
        parity = {"n": n_par, "worst_rel_err": worst, "tolerance": 1e-9,
                  "against": "oracle/qpde_oracle.c (bit-identical to the reference headers on tests/golden)",
                  "drawn_from": "the first %d contracts of the timed batch, V from the e2e call" % n_par}
        assert worst <= 1e-9, "parity sample failed: %.3e" % worst
        v = variants[head]
        line = {"metric": metric, "value": Cn / (v["ms"] / 1e3), "unit": "solves/s", "n_gpus": 1, "steps": args.steps,
                "warmup": max(1, args.warmup), "ms_per_step": v["ms"], "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": workload, "contracts_per_gpu": Cn, "nodes": N, "time_steps": steps, "scheme": head,
                           "kernel": v["kernel"], "not_converged": v["not_ok"], "steps_taken": v["steps_taken"],
                           "l2": "flushed between timed steps (256 MiB write)",
                           "variants": {k: {"solves_per_s": Cn / (vv["ms"] / 1e3), **{a: b for a, b in vv.items() if a != "solve_steps"}}
                                        for k, vv in variants.items()}},
                "roofline": roofline, "gpu_launches": int(v["launches"]), "e2e": e2e, "parity_sample": parity}
        if not args.no_cpu:
            cb = cpu_ref(rjobs)
            if cb:
                line["cpu_baseline"] = cb
    elif cfg == "er":
        # ---- examples/hjbqvi/exchange_rate.cpp: a batch of 1-D impulse-control HJBQVI problems (varied
        # sigma, a, b, lambda, c), one CTA each; qpde_hjbqvi1d_solve takes host arrays in and out, so the
        # timed call IS the end-to-end one
        g = np.load(os.path.join(ROOT, "tests", "golden", "hjbqvi1d_reference.npz"), allow_pickle=False)
        X0, Q0, Z0 = g["er_k0/X0"], g["er_k0/Q0"], g["er_k0/Z0"]
        refine = args.hjbqvi_refine if args.hjbqvi_refine >= 0 else 4
        Pn = args.contracts or 2 * handle.sm_count
        rng = np.random.default_rng(4321)
        par = np.stack([np.full(Pn, .02), rng.uniform(.2, .4, Pn), rng.uniform(.15, .5, Pn), rng.uniform(1., 4., Pn),
                        np.zeros(Pn), rng.uniform(.5, 1.5, Pn), rng.uniform(.05, .15, Pn)], axis=1)
        steps = 16 * 2 ** refine
        # parity first: the reference's own outputs at this model's default parameters, refinements 2 and 3
        worst = 0.
        for name in ("er_k2", "er_k3"):
            k = int(name[-1])
            small = capi.exchange_rate_solve(handle, X0, Q0, Z0, [[.02, .3, .25, 3., 0., 1., .1]], k, 16 * 2 ** k)
            worst = max(worst, _rel(small["V"][0], g[name + "/V"]))
            assert small["mean_inner"][0] == float(g[name + "/mean_inner"]), name
        assert worst <= 1e-9, worst
        capi.exchange_rate_solve(handle, X0, Q0, Z0, par[:2], 1, 32)            # warm-up
        ts = []
        l0 = handle.launches
        nrun = max(1, min(args.steps, 3))
        for _ in range(nrun):
            t0 = time.perf_counter()
            res = capi.exchange_rate_solve(handle, X0, Q0, Z0, par, refine, steps)
            ts.append(time.perf_counter() - t0)
        launches = (handle.launches - l0) // nrun
        secs = float(np.mean(ts))
        N, nq, nz = res["V"].shape[1], res["n_controls"], res["n_impulses"]
        its = float((res["mean_inner"] * res["n_steps"]).sum())              # policy iterations over the batch
        sweeps = float((res["mean_inner"] * res["n_steps"] * res["mean_sweeps"]).sum())
        flops = N * (its * (24.0 * nq + 12.0 * nz) + sweeps * 40.0)
        tf = flops / secs / 1e12
        line = {"metric": "exchange-rate HJBQVI solves/sec (1-D impulse control, %d nodes, %d + %d controls, %d BDF1 steps)" % (N, nq, nz, steps),
                "value": Pn / secs, "unit": "solves/s", "n_gpus": 1, "steps": nrun, "warmup": 1, "ms_per_step": 1e3 * secs,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": "exchange_rate_hjbqvi_1d_batch (examples/hjbqvi/exchange_rate.cpp at refinement %d)" % refine,
                           "problems": Pn, "nodes": N, "stochastic_controls": nq, "impulse_controls": nz, "time_steps": steps,
                           "mean_policy_iterations": float(res["mean_inner"].mean()), "mean_sweeps": float(res["mean_sweeps"].mean()),
                           "not_ok": int((res["status"] != 0).sum())},
                "roofline": {"bound": "fp64", "achieved": tf, "peak": fp64_peak, "unit": "TFLOP/s", "frac": tf / fp64_peak,
                             "traffic": None, "peak_source": "measured (qpde_measure_fp64_peak: independent DFMA chains)",
                             "algorithmic_flops_per_launch": flops,
                             "model": "per node and policy iteration: 24 flops per stochastic candidate (upwinded row, flow, "
                                      "3-term product) + 12 per admissible impulse candidate, + 40 per node and lagged tridiagonal "
                                      "solve; one launch, one CTA per problem (DESIGN.md 4.7)"},
                "gpu_launches": int(launches),
                "e2e": {"value": Pn / secs, "unit": "solves/s", "h2d_bytes_per_step": int(8 * (len(X0) + len(Q0) + len(Z0) + par.size)),
                        "d2h_bytes_per_step": int(Pn * N * 33 + Pn * 24), "ms_per_step": 1e3 * secs,
                        "api": "qpde_hjbqvi1d_solve (host axes and parameters in; V, X, Q, Z, mask out): the timed call itself"},
                "parity_sample": {"n": 2, "worst_rel_err": worst, "tolerance": 1e-9, "policy_iteration_counts_equal": True,
                                  "against": "tests/golden/hjbqvi1d_reference.npz er_k2, er_k3 (the reference's own outputs)"}}
        if not args.no_cpu:
            from oracle import pyoracle as po
            if po.have_ref():
                n_jobs = min(Pn, 2 * cores)
                jobs = [("exchange", dict(refine=refine, sigma=float(par[c, 1]), a=float(par[c, 2]), b=float(par[c, 3]),
                                          c=float(par[c, 6]), **{"lambda": float(par[c, 5])})) for c in range(n_jobs)]
                cb = cpu_ref(jobs)
                if cb:
                    line["cpu_baseline"] = cb
    elif cfg == "oc":
        # ---- examples/hjbqvi/optimal_consumption.cpp: the 2-D HJBQVI with impulses both ways; the timed call is
        # qpde_hjbqvi2d_solve (host in, host out) at refinement 4 (305 x 305 nodes, 241 + 241 controls, 512 steps)
        refine = args.hjbqvi_refine if args.hjbqvi_refine >= 0 else 4
        g = np.load(os.path.join(ROOT, "tests", "golden", "hjbqvi2d_reference.npz"), allow_pickle=False)
        small = capi.optimal_consumption_solve(handle, 2)                       # warm-up + parity: the reference's own output
        worst = _rel(small["V"], g["oc_k2/V"].reshape(small["V"].shape))
        assert worst <= 1e-9 and small["mean_inner"] == float(g["oc_k2/mean_inner"]), worst
        ts = []
        l0 = handle.launches
        nrun = max(1, min(args.steps, 2))
        for _ in range(nrun):
            t0 = time.perf_counter()
            res = capi.optimal_consumption_solve(handle, refine)
            ts.append(time.perf_counter() - t0)
        launches = (handle.launches - l0) // nrun
        secs = float(np.mean(ts))
        nw, nb = len(res["W"]), len(res["B"])
        ncand = 15 * 2 ** refine + 1
        its = float(res["mean_inner"]) * res["n_steps"]
        flops = nw * nb * its * (30.0 * ncand + 90.0 * ncand + float(res["mean_sweeps"]) * 30.0)
        tf = flops / secs / 1e12
        line = {"metric": "optimal-consumption HJBQVI solves/sec (2-D impulse control both ways, %d x %d nodes, %d BDF1 steps)" % (nw, nb, res["n_steps"]),
                "value": 1.0 / secs, "unit": "solves/s", "n_gpus": 1, "steps": nrun, "warmup": 1, "ms_per_step": 1e3 * secs,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": "optimal_consumption_hjbqvi_2d (examples/hjbqvi/optimal_consumption.cpp at refinement %d)" % refine,
                           "nodes_w": nw, "nodes_b": nb, "stochastic_controls": ncand, "impulse_controls": ncand,
                           "time_steps": int(res["n_steps"]), "mean_policy_iterations": float(res["mean_inner"]),
                           "line_sweeps_per_policy_iteration": float(res["mean_sweeps"]), "status": int(res["status"]),
                           "seconds_per_solve": secs, "node_steps_per_s": nw * nb * res["n_steps"] / secs},
                "roofline": {"bound": "fp64", "achieved": tf, "peak": fp64_peak, "unit": "TFLOP/s", "frac": tf / fp64_peak,
                             "traffic": None, "peak_source": "measured (qpde_measure_fp64_peak: independent DFMA chains)",
                             "algorithmic_flops_per_launch": flops / max(1, launches),
                             "model": "per node and policy iteration: 30 flops per stochastic candidate + 90 per impulse candidate "
                                      "(admissible transfer, two brackets, bilinear row) + 30 per node and line sweep; the solve is "
                                      "launch- and latency-bound on small grids: 2 launches per sweep of ~150 sweeps (DESIGN.md 4.8)"},
                "gpu_launches": int(launches),
                "e2e": {"value": 1.0 / secs, "unit": "solves/s", "h2d_bytes_per_step": int(8 * (nw * nb + 2 * 16 + 2 * 20 + 8)),
                        "d2h_bytes_per_step": int(nw * nb * 25), "ms_per_step": 1e3 * secs,
                        "api": "qpde_hjbqvi2d_solve (host axes and parameters in; V, Q, Z, mask out): the timed call itself"},
                "parity_sample": {"n": 1, "worst_rel_err": worst, "tolerance": 1e-9, "policy_iteration_counts_equal": True,
                                  "against": "tests/golden/hjbqvi2d_reference.npz oc_k2 (the reference's own output, 77 x 77 nodes, 128 steps)"}}
        if not args.no_cpu:
            from oracle import pyoracle as po
            if po.have_ref():
                t0 = time.perf_counter()
                ref = po.run_ref("consumption", refine=2)
                wall = time.perf_counter() - t0
                n1 = len(ref["W"]) * len(ref["B"]) * int(ref["steps"])
                line["cpu_baseline"] = {"value": 1.0 / wall, "unit": "solves/s", "cores": 1, "kind": "reference",
                                        "node_steps_per_s": n1 / wall,
                                        "sample": "refinement 2 of the same problem (77 x 77 nodes, 61 + 61 controls, 128 steps): %.1f s on "
                                                  "one core (the reference is single-threaded; refinement 3 takes it 633 s, refinement 4 "
                                                  "hours); compare node_steps_per_s, not solves/s" % wall}
    else:
        # ---- configs[4]: GMWB 2-D impulse control, 4097 x 1025 nodes; qpde_gmwb2d_solve takes host arrays
        # in and out, so the timed call IS the end-to-end one
        refine, a_points, ts0 = args.gmwb_refine, 17, args.gmwb_timesteps0
        A = np.array([0. + (100. / (a_points - 1)) * i for i in range(a_points)])
        Z = np.array([0., .5, 1.])
        capi.gmwb_solve(handle, GMWB_W, A, [0., 10.], Z, 1, 2)                  # warm-up: module load, clocks
        # parity first: refinement 1 of the reference's own default problem against its committed output
        g = np.load(os.path.join(ROOT, "tests", "golden", "gmwb_reference.npz"), allow_pickle=False)
        A51 = np.array([0. + 2. * i for i in range(51)])
        small = capi.gmwb_solve(handle, GMWB_W, A51, [0., 10.], Z, 1, 64)
        worst = _rel(small["V"], g["gmwb_k1/V"].reshape(small["V"].shape))
        assert worst <= 1e-9 and small["mean_inner"] == float(g["gmwb_k1/mean_inner"]), worst
        ts = []
        l0 = handle.launches
        nrun = max(1, min(args.steps, 2))
        for _ in range(nrun):
            t0 = time.perf_counter()
            res = capi.gmwb_solve(handle, GMWB_W, A, [0., 10.], Z, refine, ts0 * 2 ** refine)
            ts.append(time.perf_counter() - t0)
        launches = (handle.launches - l0) // nrun
        secs = float(np.mean(ts))
        nw, na, nz = len(res["W"]), len(res["A"]), 2 * 2 ** refine + 1
        its = float(res["mean_inner"]) * res["n_steps"]
        flops = (60.0 * nz + 2 * 12.0 + 30.0) * nw * na * its
        tf = flops / secs / 1e12
        line = {"metric": "GMWB HJBQVI solves/sec (2-D impulse control, %d x %d nodes, %d BDF1 steps)" % (nw, na, res["n_steps"]),
                "value": 1.0 / secs, "unit": "solves/s", "n_gpus": 1, "steps": nrun, "warmup": 1, "ms_per_step": 1e3 * secs,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": "gmwb_2d_impulse_control_%dx%d (BASELINE.json configs[4])" % (nw, na), "nodes_w": nw,
                           "nodes_a": na, "impulse_candidates": nz, "time_steps": int(res["n_steps"]),
                           "mean_policy_iterations": float(res["mean_inner"]), "status": int(res["status"]),
                           "seconds_per_solve": secs, "node_steps_per_s": nw * na * res["n_steps"] / secs},
                "roofline": {"bound": "fp64", "achieved": tf, "peak": fp64_peak, "unit": "TFLOP/s", "frac": tf / fp64_peak,
                             "traffic": None, "peak_source": "measured (qpde_measure_fp64_peak: independent DFMA chains)",
                             "algorithmic_flops_per_launch": flops / max(1, launches),
                             "ncu_fp64_pipe_active_pct_policy_kernel": 28.5,
                             "model": "per node and policy iteration: 60 flops per impulse candidate + 24 for the two stochastic "
                                      "controls + 30 for the line solve; the whole solve is ~3 launches per policy iteration, the "
                                      "line pass is a serial chain of n_a line solves (latency-bound, DESIGN.md 4.5)"},
                "gpu_launches": int(launches),
                "e2e": {"value": 1.0 / secs, "unit": "solves/s", "h2d_bytes_per_step": int(8 * (len(GMWB_W) + len(A) + len(Z) + 2)),
                        "d2h_bytes_per_step": int(8 * nw * na), "ms_per_step": 1e3 * secs,
                        "api": "qpde_gmwb2d_solve (host axes in, V[n_a][n_w] out): the timed call itself"},
                "parity_sample": {"n": 1, "worst_rel_err": worst, "tolerance": 1e-9,
                                  "against": "tests/golden/gmwb_reference.npz gmwb_k1 (the reference's own output, 129 x 101 nodes, 64 steps)"}}
        if not args.no_cpu:
            from oracle import pyoracle as po
            if po.have_ref():
                t0 = time.perf_counter()
                ref = po.run_ref("gmwb", refine=1)
                wall = time.perf_counter() - t0
                n1 = len(ref["W"]) * len(ref["A"]) * int(ref["steps"])
                line["cpu_baseline"] = {"value": 1.0 / wall, "unit": "solves/s", "cores": 1, "kind": "reference",
                                        "node_steps_per_s": n1 / wall,
                                        "sample": "refinement 1 of the reference's default problem (%d x %d nodes, %d steps, 5 impulse "
                                                  "candidates): %.1f s on one core (the reference is single-threaded; the config-size "
                                                  "solve would take days); compare node_steps_per_s, not solves/s"
                                                  % (len(ref["W"]), len(ref["A"]), int(ref["steps"]), wall)}
    line["clocks"] = sampler.stop()
    handle.close()
    sys.stdout.flush()
    os.dup2(result_fd, 1)
    os.close(result_fd)
    print(json.dumps(line), flush=True)


def run_gmwb_sharded(args, rank, world, local):
    """--config 5 on N > 1 GPUs: configs[4] with the control search sharded by the A-axis over the ranks
    (qpde_comm_init + qpde_gmwb2d_solve_sharded: NCCL inside the library; torch.distributed / gloo only ships
    the 128-byte communicator id and reduces the wall time).  Strong scaling: one problem, N GPUs."""
    import torch
    import torch.distributed as dist
    from quantpde_b200 import capi
    sys.stdout.flush()
    result_fd = os.dup(1)
    os.dup2(2, 1)
    torch.cuda.set_device(local)
    dist.init_process_group("gloo")
    refine, a_points, ts0 = args.gmwb_refine, 17, args.gmwb_timesteps0
    A = np.array([0. + (100. / (a_points - 1)) * i for i in range(a_points)])
    Z = np.array([0., .5, 1.])
    handle = capi.Handle(local)
    ident = torch.zeros(capi.COMM_ID_BYTES, dtype=torch.uint8)
    if rank == 0:
        ident = torch.frombuffer(bytearray(capi.comm_unique_id()), dtype=torch.uint8).clone()
    dist.broadcast(ident, 0)
    handle.comm_init(ident.numpy().tobytes(), rank, world)
    capi.gmwb_solve(handle, GMWB_W, A, [0., 10.], Z, 1, 2, sharded=True)          # warm-up: module load, NCCL channels
    # parity first (every rank takes part): refinement 1 of the reference's default problem against its own output
    g = np.load(os.path.join(ROOT, "tests", "golden", "gmwb_reference.npz"), allow_pickle=False)
    A51 = np.array([0. + 2. * i for i in range(51)])
    small = capi.gmwb_solve(handle, GMWB_W, A51, [0., 10.], Z, 1, 64, sharded=True)
    worst = _rel(small["V"], g["gmwb_k1/V"].reshape(small["V"].shape))
    assert worst <= 1e-9 and small["mean_inner"] == float(g["gmwb_k1/mean_inner"]), worst
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ts = []
    nrun = max(1, min(args.steps, 2))
    l0 = handle.launches
    for _ in range(nrun):
        dist.barrier()
        t0 = time.perf_counter()
        res = capi.gmwb_solve(handle, GMWB_W, A, [0., 10.], Z, refine, ts0 * 2 ** refine, sharded=True)
        secs = torch.tensor([time.perf_counter() - t0], dtype=torch.float64)
        dist.all_reduce(secs, op=dist.ReduceOp.MAX)
        ts.append(float(secs.item()))
    launches = (handle.launches - l0) // nrun
    clocks = sampler.stop() if rank == 0 else None
    # every rank holds the full V: they must agree bit for bit
    digest = torch.tensor([float(np.frombuffer(res["V"].tobytes(), dtype=np.uint64).sum() % (2 ** 52))], dtype=torch.float64)
    lo, hi = digest.clone(), digest.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    secs = float(np.mean(ts))
    nw, na, nz = len(res["W"]), len(res["A"]), 2 * 2 ** refine + 1
    its = float(res["mean_inner"]) * res["n_steps"]
    flops = (60.0 * nz + 2 * 12.0 + 30.0) * nw * na * its
    line = {"metric": "GMWB HJBQVI solves/sec (2-D impulse control, %d x %d nodes, %d BDF1 steps)" % (nw, na, res["n_steps"]),
            "value": 1.0 / secs, "unit": "solves/s", "n_gpus": world, "steps": nrun, "warmup": 1, "ms_per_step": 1e3 * secs,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "gmwb_2d_impulse_control_%dx%d (BASELINE.json configs[4])" % (nw, na), "nodes_w": nw,
                       "nodes_a": na, "impulse_candidates": nz, "time_steps": int(res["n_steps"]),
                       "mean_policy_iterations": float(res["mean_inner"]), "status": int(res["status"]),
                       "seconds_per_solve": secs,
                       "parallelism": "control search sharded by the A-axis over %d ranks, decisions exchanged by NCCL "
                                      "broadcasts inside the library, line chain replicated" % world,
                       "ranks_bit_identical": bool(lo.item() == hi.item())},
            "roofline": {"bound": "fp64", "achieved": flops / secs / 1e12, "peak": None, "unit": "TFLOP/s", "frac": None,
                         "traffic": None, "algorithmic_flops_per_launch": flops / max(1, launches),
                         "model": "per node and policy iteration: 60 flops per impulse candidate + 24 for the two stochastic "
                                  "controls + 30 for the line solve, summed over the ranks"},
            "gpu_launches": int(launches), "clocks": clocks,
            "e2e": {"value": 1.0 / secs, "unit": "solves/s", "h2d_bytes_per_step": int(8 * (len(GMWB_W) + len(A) + len(Z) + 2)),
                    "d2h_bytes_per_step": int(8 * nw * na), "ms_per_step": 1e3 * secs,
                    "api": "qpde_gmwb2d_solve_sharded (host axes in, V[n_a][n_w] out on every rank): the timed call itself"},
            "parity_sample": {"n": 1, "worst_rel_err": worst, "tolerance": 1e-9,
                              "against": "tests/golden/gmwb_reference.npz gmwb_k1 (the reference's own output), solved sharded"}}
    handle.close()
    dist.barrier()
    dist.destroy_process_group()
    sys.stdout.flush()
    os.dup2(result_fd, 1)
    os.close(result_fd)
    if rank == 0:
        print(json.dumps(line), flush=True)


def run_reference_arm_other(args):
    """--impl reference --config {3,4,5,hjb}: the reference's CPU path on a bounded sample of that config."""
    from concurrent.futures import ProcessPoolExecutor
    from oracle import pyoracle as po
    cores = os.cpu_count() or 1
    cfg = args.config
    if cfg == "5":
        walls = []
        for _ in range(max(1, min(args.steps, 2))):
            t0 = time.perf_counter(); ref = po.run_ref("gmwb", refine=1); walls.append(time.perf_counter() - t0)
        value = 1.0 / float(np.mean(walls))
        sample = "refinement 1 of the reference's default GMWB problem (129 x 101 nodes, 64 steps) on one core"
        workload, n_jobs = "gmwb_2d_impulse_control (BASELINE.json configs[4]), bounded sample", 1
    elif cfg == "oc":
        walls = []
        for _ in range(max(1, min(args.steps, 2))):
            t0 = time.perf_counter(); ref = po.run_ref("consumption", refine=2); walls.append(time.perf_counter() - t0)
        value = 1.0 / float(np.mean(walls))
        sample = "refinement 2 of the reference's optimal-consumption problem (77 x 77 nodes, 128 steps) on one core"
        workload, n_jobs = "optimal_consumption_hjbqvi_2d, bounded sample", 1
    elif cfg == "er":
        refine = args.hjbqvi_refine if args.hjbqvi_refine >= 0 else 4
        n_jobs = args.cpu_sample or 2 * cores
        rng = np.random.default_rng(4321)
        Pn = max(n_jobs, 296)
        sig, a_, b_, lam, c_ = rng.uniform(.2, .4, Pn), rng.uniform(.15, .5, Pn), rng.uniform(1., 4., Pn), rng.uniform(.5, 1.5, Pn), rng.uniform(.05, .15, Pn)
        jobs = [("exchange", dict(refine=refine, sigma=float(sig[c]), a=float(a_[c]), b=float(b_[c]), c=float(c_[c]),
                                  **{"lambda": float(lam[c])})) for c in range(n_jobs)]
        walls = []
        for k in range(args.warmup + args.steps):
            t0 = time.perf_counter()
            with ProcessPoolExecutor(max_workers=cores) as ex:
                list(ex.map(_ref_job, jobs))
            if k >= args.warmup:
                walls.append(time.perf_counter() - t0)
        value = n_jobs * len(walls) / float(sum(walls))
        sample = "%d problems of the workload per step, one per worker process on %d cores" % (n_jobs, cores)
        workload = "exchange_rate_hjbqvi_1d_batch (examples/hjbqvi/exchange_rate.cpp at refinement %d)" % refine
    else:
        P = _other_config_problem(cfg, args.contracts)
        n_jobs = args.cpu_sample or 2 * cores
        if cfg == "3":
            jobs = [("bermudan", dict(steps=P["steps"], refine=6, scheme="bdf2", K=float(P["K"][c]), r=float(P["r"][c]),
                                      alpha=float(P["alpha"][c]), T=1.0, E=10)) for c in range(n_jobs)]
            workload = "bermudan_put_local_vol_2113x6400x10_events (BASELINE.json configs[2])"
        elif cfg == "4":
            jobs = [("jump", dict(steps=P["steps"], refine=6, density="lognormal", K=float(P["K"][c]), r=0.05, vol=float(P["vol"][c]),
                                  T=0.25, mu=P["pairs"][P["which"][c]][0], sd=P["pairs"][P["which"][c]][1], **{"lambda": 0.1}))
                    for c in range(n_jobs)]
            workload = "merton_jump_call_2113x768_nfft4096 (BASELINE.json configs[3])"
        else:
            jobs = [("hjb", dict(steps=P["steps"], refine=6, long=0, scheme="bdf2", K=float(P["K"][c]), r_l=float(P["r_l"][c]),
                                 r_b=float(P["r_b"][c]), vol=float(P["vol"][c]), T=1.0)) for c in range(n_jobs)]
            workload = "hjb_unequal_rates_straddle_2049x768"
        walls = []
        for k in range(args.warmup + args.steps):
            t0 = time.perf_counter()
            with ProcessPoolExecutor(max_workers=cores) as ex:
                list(ex.map(_ref_job, jobs))
            if k >= args.warmup:
                walls.append(time.perf_counter() - t0)
        value = n_jobs * len(walls) / float(sum(walls))
        sample = "%d contracts of the workload per step, one per worker process on %d cores" % (n_jobs, cores)
    print(json.dumps({"impl": "reference", "metric": "option PDE solves/sec", "value": value, "unit": "solves/s",
                      "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * float(np.mean(walls)),
                      "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                      "config": {"workload": workload, "contracts_per_step": n_jobs},
                      "cpu_baseline": {"value": value, "unit": "solves/s", "cores": cores if cfg not in ("5", "oc") else 1, "kind": "reference",
                                       "sample": sample},
                      "e2e": {"value": value, "unit": "solves/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                      "gpu_launches": 0}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="1", choices=["1", "3", "4", "5", "hjb", "er", "oc"],
                    help="1: BASELINE.json configs[1] (default); 3 / 4 / 5: configs[2] / [3] / [4]; hjb: policy iteration")
    ap.add_argument("--contracts", type=int, default=0, help="contracts per GPU (default: the config's batch size)")
    ap.add_argument("--kernel", default="auto", choices=["auto", "interleaved", "resident"])
    ap.add_argument("--cpu-sample", type=int, default=0)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--hjbqvi-refine", type=int, default=-1, help="--config er / oc: grid refinement (default 4)")
    ap.add_argument("--gmwb-refine", type=int, default=6, help="--config 5: grid refinement (6 = 4097 x 1025)")
    ap.add_argument("--gmwb-timesteps0", type=int, default=32, help="--config 5: time steps before refinement (32 -> 2048 at refinement 6)")
    ap.add_argument("--no-strong", action="store_true", help="N > 1: skip the strong-scaling leg (65,536 contracts in total)")
    ap.add_argument("--no-parity", action="store_true", help="skip the node-level parity sample against the oracle")
    ap.add_argument("--parity-sample", type=int, default=256)
    ap.add_argument("--no-streaming", action="store_true",
                    help="skip the bounded sample of the HBM-streaming (INTERLEAVED, TMA-staged) family")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return
    if args.config == "5" and world > 1:
        run_gmwb_sharded(args, rank, world, local)         # every rank: the sharded solve is a collective call
        return
    if args.config != "1":
        if rank == 0:
            run_other_config(args)
        return
    if args.contracts <= 0:
        args.contracts = 65536

    import torch
    import torch.distributed as dist
    from quantpde_b200 import capi

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    # stdout carries exactly one line, the result: anything a library prints on the way (NCCL
    # announces its version on fd 1 at the first collective) goes to stderr
    sys.stdout.flush()
    result_fd = os.dup(1)
    os.dup2(2, 1)
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    Cn = args.contracts
    K, vol, T, r = contract_parameters(Cn, rank, world)
    axis = np.ascontiguousarray(K[:, None] * SPECIAL[None, :])     # K * Axis::special (S_0 = K)
    handle = capi.Handle(local)
    outputs = capi.OUT_V | capi.OUT_VALUE | capi.OUT_INNER_TOTAL | capi.OUT_STATUS | capi.OUT_STEPS | capi.OUT_COMPUTED
    spec = capi.BatchSpec(axis=axis, refine=N_NODES_REFINE, r=r, sigma=vol, q=0.0, T=T,
                          dt=T / N_STEPS, strike=K, american=True, payoff="put",
                          kernel=args.kernel, outputs=outputs)
    plan = capi.Plan(handle, spec)
    N = plan.N
    stream = torch.cuda.ExternalStream(handle.stream, device=torch.device("cuda", local))
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device="cuda")  # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def flush_l2():
        flush.fill_(1.0)
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        plan.run()
        handle.synchronize()

    sampler = ClockSampler(local)
    barrier()
    sampler.start()
    launches0 = handle.launches
    times_ms = []
    for _ in range(args.steps):
        flush_l2()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        plan.run()
        e1.record(stream)
        e1.synchronize()
        times_ms.append(e0.elapsed_time(e1))
    launches = handle.launches - launches0
    barrier()
    clocks = sampler.stop()
    total_ms = float(sum(times_ms))

    res = plan.fetch(capi.OUT_INNER_TOTAL | capi.OUT_STATUS | capi.OUT_STEPS | capi.OUT_COMPUTED)
    inner_sum = int(res.inner_total.astype(np.int64).sum())
    computed_sum = int(res.computed.astype(np.int64).sum())
    not_converged = int((res.status != 0).sum())

    # ---- parity of the batch that was just timed (rank 0): a stratified sample of its contracts,
    # node for node against the oracle; active sets and per-step counts from a second plan of the
    # sampled contracts alone, whose V must equal the full batch's bit for bit.
    parity = None
    if rank == 0 and not args.no_parity:
        idx = parity_sample_indices(Cn, K, vol, T, r, args.parity_sample)
        full = plan.fetch(capi.OUT_V)
        Vs = np.ascontiguousarray(full.V[idx]); del full
        sub = capi.BatchSpec(axis=axis[idx], refine=N_NODES_REFINE, r=r[idx], sigma=vol[idx], q=0.0, T=T[idx],
                             dt=T[idx] / N_STEPS, strike=K[idx], american=True, payoff="put", kernel=args.kernel)
        with capi.Plan(handle, sub) as splan:
            splan.run()
            sres = splan.fetch()
        parity = parity_sample(idx, K, vol, T, r, Vs, res.inner_total[idx], res.n_steps[idx], active=sres.active)
        parity["sample_plan_bits_equal_batch"] = bool(np.array_equal(sres.V, Vs))
        parity["drawn_from"] = "the timed %d-contract batch: first/last contracts, extremes of K, sigma, T, r, evenly spaced rest" % Cn
        del Vs, sres

    # ---- e2e: qpde_bs1d_sweep with host buffers, copies inside the timed region
    e2e = None
    if not args.no_e2e:
        # One call of the reference-facing API per step: host pointers in (the
        # contract parameters), host pointers out (V[C][N], value, inner_total,
        # status) into page-locked buffers the caller allocated once.
        e2e_out = capi.OUT_V | capi.OUT_VALUE | capi.OUT_INNER_TOTAL | capi.OUT_STATUS
        r2 = capi.ResultBuffers(Cn, N, capi.sweep_max_steps(spec), e2e_out, pinned=True)
        e2e_times = []
        for i in range(1 + args.steps):            # one untimed call, then --steps timed ones
            barrier()
            t0 = time.perf_counter()
            capi.sweep(handle, spec, outputs=e2e_out, result=r2)
            torch.cuda.synchronize()
            dt_s = time.perf_counter() - t0
            if i > 0:
                e2e_times.append(dt_s)
        h2d = axis.nbytes + 6 * 8 * Cn
        d2h = r2.V.nbytes + r2.value.nbytes + r2.inner_total.nbytes + r2.status.nbytes
        e2e_ms = 1e3 * float(np.mean(e2e_times))
        assert int(r2.inner_total.astype(np.int64).sum()) == inner_sum   # same work as the resident run
        r2.close()
    # ---- strong scaling (N > 1): the BASELINE batch itself, 65,536 contracts in total, split over the ranks
    strong_ms = 0.0
    if world > 1 and not args.no_strong:
        Ks, vs, Ts, rs = contract_parameters(65536 // world, rank, world)
        sspec = capi.BatchSpec(axis=np.ascontiguousarray(Ks[:, None] * SPECIAL[None, :]), refine=N_NODES_REFINE, r=rs,
                               sigma=vs, q=0.0, T=Ts, dt=Ts / N_STEPS, strike=Ks, american=True, payoff="put",
                               kernel=args.kernel, outputs=outputs)
        with capi.Plan(handle, sspec) as sp:
            sp.run(); handle.synchronize()
            barrier()
            for _ in range(args.steps):
                flush_l2()
                e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
                e0.record(stream); sp.run(); e1.record(stream); e1.synchronize()
                strong_ms += e0.elapsed_time(e1)
            barrier()
    if world > 1:
        tt = torch.tensor([total_ms, e2e_ms if not args.no_e2e else 0.0, strong_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ii = torch.tensor([inner_sum, not_converged, computed_sum], device="cuda", dtype=torch.int64)
        dist.all_reduce(ii, op=dist.ReduceOp.SUM)
        total_ms, e2e_ms_max, strong_ms = float(tt[0]), float(tt[1]), float(tt[2])
        inner_sum_all, not_converged, computed_all = int(ii[0]), int(ii[1]), int(ii[2])
    else:
        e2e_ms_max = e2e_ms if not args.no_e2e else 0.0
        inner_sum_all, computed_all = inner_sum, computed_sum
    if not args.no_e2e:
        e2e = {"value": Cn * world / (e2e_ms_max / 1e3), "unit": "solves/s",
               "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
               "ms_per_step": e2e_ms_max,
               "api": "qpde_bs1d_sweep (host pointers in; V[C][N] + value + inner_total + status out into page-locked host buffers)"}

    ms_per_step = total_ms / args.steps
    value = Cn * world / (ms_per_step / 1e3)

    # ---- the HBM-streaming family on the same workload (rank 0, N = 1): same contracts, grid and 1000 time
    # steps, INTERLEAVED kernel (TMA ring): one warm-up sweep, one timed sweep (4.2 s each).
    streaming = None
    if rank == 0 and world == 1 and not args.no_streaming:
        s_steps = N_STEPS
        sspec = capi.BatchSpec(axis=axis, refine=N_NODES_REFINE, r=r, sigma=vol, q=0.0, T=T, dt=T / s_steps,
                               strike=K, american=True, payoff="put", kernel="interleaved",
                               outputs=capi.OUT_VALUE | capi.OUT_INNER_TOTAL | capi.OUT_STATUS)
        with capi.Plan(handle, sspec) as splan:
            splan.run(); handle.synchronize()                      # warm-up
            flush_l2()
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            l0 = handle.launches
            e0.record(stream); splan.run(); e1.record(stream); e1.synchronize()
            s_ms = e0.elapsed_time(e1)
            sres = splan.fetch(capi.OUT_INNER_TOTAL | capi.OUT_STATUS)
            s_inner = int(sres.inner_total.astype(np.int64).sum())
            s_alg = ALG_BYTES_PER_NODE_SOLVE * N * s_inner
            speak, _ = load_peaks()
            streaming = {"kernel": "interleaved (TMA-staged)" if splan.tma_staged else "interleaved (plain loads)",
                         "sample": "%d contracts x %d nodes x %d steps (setup + sweep launch)" % (Cn, N, s_steps),
                         "ms": s_ms, "solves_per_s": Cn / (s_ms / 1e3), "gpu_launches": int(handle.launches - l0),
                         "achieved": s_alg / (s_ms / 1e3) / 1e9, "peak": speak, "unit": "GB/s",
                         "frac": s_alg / (s_ms / 1e3) / 1e9 / speak,
                         "not_converged": int((sres.status != 0).sum()),
                         "note": "achieved = 56 B x nodes x inner solves / time (SURVEY.md 8d); the kernel itself moves up "
                                 "to 104 B per computed inner solve and node, less where a further penalty iteration restarts "
                                 "at the first changed row and on zero-obstacle tiles (DESIGN.md 4.1); ncu capture of the same "
                                 "batch at 20 steps: 3.94 TB of DRAM traffic per launch (round 1: 4.85 TB) at 6.20 TB/s = 95 % "
                                 "of the measured peak (profiles/r2_interleaved_tma_ncu_raw_subset.csv)"}

    peak, peak_src = load_peaks()
    alg_bytes = ALG_BYTES_PER_NODE_SOLVE * N * (inner_sum_all / world)   # per launch (one rank's batch)
    achieved = alg_bytes / (ms_per_step / 1e3) / 1e9
    exe_bytes = ALG_BYTES_PER_NODE_SOLVE * N * (computed_all / world)    # the solves the device actually ran
    achieved_exe = exe_bytes / (ms_per_step / 1e3) / 1e9
    traffic = load_traffic(plan.kernel, Cn)
    line = {
        "metric": "option PDE solves/sec (N=2049, 1000 CN steps)", "value": value, "unit": "solves/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": "american_put_batch_2049x1000 (BASELINE.json configs[1])",
                   "contracts_per_gpu": Cn, "nodes": N, "time_steps": N_STEPS,
                   "scheme": "rannacher+penalty", "kernel": plan.kernel,
                   "mean_inner_iterations_per_step": inner_sum_all / (world * Cn * float(N_STEPS)),
                   "not_converged": not_converged,
                   "l2": "flushed between timed steps (256 MiB write)",
                   "parallelism": "contracts sharded over %d rank(s), no collective" % world},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": alg_bytes,
                     "achieved_executed": achieved_exe, "frac_executed": achieved_exe / peak,
                     "executed_solves_per_counted_solve": computed_all / max(1, inner_sum_all),
                     "model": "56 B x nodes x sum of inner solves (SURVEY.md 8d): 'achieved' bills every inner "
                              "iteration the reference counts, 'achieved_executed' only the linear solves the device "
                              "ran (a confirming iteration with an unchanged active set is counted, not run).  The "
                              "RESIDENT kernel keeps its state on chip (traffic = measured DRAM bytes per launch), so "
                              "frac > 1 means it has left the HBM roofline: its bound is dependent-issue latency "
                              "(profiles/r2_lean_*: issue slots and FP64 pipe)"},
        "clocks": clocks, "gpu_launches": int(launches), "e2e": e2e,
    }
    if parity is not None:
        line["parity_sample"] = parity
    if world > 1 and strong_ms > 0.0:
        line["strong_scaling"] = {"contracts_total": 65536, "contracts_per_gpu": 65536 // world,
                                  "ms_per_step": strong_ms / args.steps,
                                  "value": 65536 // world * world / (strong_ms / args.steps / 1e3), "unit": "solves/s"}
    if streaming is not None:
        line["roofline_streaming_family"] = streaming
    if rank == 0 and world == 1 and not args.no_cpu:    # the CPU baseline is an N = 1 leg
        cores = os.cpu_count() or 1
        line["cpu_baseline"] = {k: v for k, v in cpu_baseline(args.cpu_sample or 256).items()
                                if k in ("value", "unit", "cores", "kind", "sample", "solves_per_s_per_core")}
    plan.close()
    handle.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    sys.stdout.flush()
    os.dup2(result_fd, 1)
    os.close(result_fd)
    if rank == 0:
        print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
