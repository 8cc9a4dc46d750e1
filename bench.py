#!/usr/bin/env python
"""bench.py — option PDE solves/sec on BASELINE.json's configs[1]:
a batch of American puts (varied strike / vol / T / r, seeded as SURVEY.md
§8(d)), 1-D Rannacher(CN) + PenaltyMethod, 2049 nodes x 1000 steps, FP64.

    python bench.py --gpus N --steps K --warmup W [--impl reference]
    torchrun ... bench.py --gpus N ...      (one rank per GPU, N > 1)

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


def cpu_baseline(n_sample, kind=None):
    """Times a bounded sample of the workload on all host cores (one contract
    per worker process, as BASELINE.md §3 prescribes for the single-threaded
    reference).  Returns solves/s for the whole host."""
    from concurrent.futures import ProcessPoolExecutor
    from oracle import pyoracle as po
    po.lib()
    if kind is None:
        kind = "reference" if po.have_ref() else "port"
    cores = os.cpu_count() or 1
    K, vol, T, r = contract_parameters(n_sample)
    jobs = [(kind, K[i], vol[i], T[i], r[i]) for i in range(n_sample)]
    t0 = time.perf_counter()
    with ProcessPoolExecutor(max_workers=cores) as ex:
        results = list(ex.map(_cpu_one, jobs))
    wall = time.perf_counter() - t0
    per_contract = float(np.mean([s for s, _ in results]))
    return {"value": n_sample / wall, "unit": "solves/s", "cores": cores, "kind": kind,
            "sample": "%d contracts of the same seeded workload (2049 nodes x 1000 steps), one per "
                      "worker process on %d cores; %.3f s per contract single-threaded; LU is "
                      "oracle/shim's stand-in, not Eigen's" % (n_sample, cores, per_contract),
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
    from oracle import pyoracle as po
    po.lib()
    cores = os.cpu_count() or 1
    n_sample = args.cpu_sample or max(cores, 8)
    for _ in range(args.warmup):
        cpu_baseline(min(cores, n_sample))
    vals, walls = [], []
    for _ in range(args.steps):
        b = cpu_baseline(n_sample)
        vals.append(b["value"]); walls.append(b["wall_s"])
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
        "cpu_baseline": {k: b[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": value, "unit": "solves/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    line["cpu_baseline"]["value"] = value
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--contracts", type=int, default=65536, help="contracts per GPU")
    ap.add_argument("--kernel", default="auto", choices=["auto", "interleaved", "resident"])
    ap.add_argument("--cpu-sample", type=int, default=0)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true")
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
    outputs = capi.OUT_V | capi.OUT_VALUE | capi.OUT_INNER_TOTAL | capi.OUT_STATUS | capi.OUT_STEPS
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

    res = plan.fetch(capi.OUT_INNER_TOTAL | capi.OUT_STATUS | capi.OUT_STEPS)
    inner_sum = int(res.inner_total.astype(np.int64).sum())
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
    if world > 1:
        tt = torch.tensor([total_ms, e2e_ms if not args.no_e2e else 0.0], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ii = torch.tensor([inner_sum, not_converged], device="cuda", dtype=torch.int64)
        dist.all_reduce(ii, op=dist.ReduceOp.SUM)
        total_ms, e2e_ms_max = float(tt[0]), float(tt[1])
        inner_sum_all, not_converged = int(ii[0]), int(ii[1])
    else:
        e2e_ms_max = e2e_ms if not args.no_e2e else 0.0
        inner_sum_all = inner_sum
    if not args.no_e2e:
        e2e = {"value": Cn * world / (e2e_ms_max / 1e3), "unit": "solves/s",
               "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
               "ms_per_step": e2e_ms_max,
               "api": "qpde_bs1d_sweep (host pointers in; V[C][N] + value + inner_total + status out into page-locked host buffers)"}

    ms_per_step = total_ms / args.steps
    value = Cn * world / (ms_per_step / 1e3)

    # ---- the HBM-streaming family on a bounded sample of the same workload (rank 0, N = 1):
    # same contracts and grid, 200 time steps instead of 1000, INTERLEAVED kernel (TMA ring).
    streaming = None
    if rank == 0 and world == 1 and not args.no_streaming:
        s_steps = 200
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
                         "ms": s_ms, "gpu_launches": int(handle.launches - l0),
                         "achieved": s_alg / (s_ms / 1e3) / 1e9, "peak": speak, "unit": "GB/s",
                         "frac": s_alg / (s_ms / 1e3) / 1e9 / speak,
                         "not_converged": int((sres.status != 0).sum()),
                         "note": "achieved = 56 B x nodes x inner solves / time (SURVEY.md 8d); the kernel itself moves "
                                 "104 B per computed inner solve and node (DESIGN.md 4.1); ncu capture of the same "
                                 "batch at 20 steps: 6.10 TB/s of DRAM traffic = 93 % of the measured peak "
                                 "(profiles/r1_interleaved_tma_ncu_raw_subset.csv)"}

    peak, peak_src = load_peaks()
    alg_bytes = ALG_BYTES_PER_NODE_SOLVE * N * (inner_sum_all / world)   # per launch (one rank's batch)
    achieved = alg_bytes / (ms_per_step / 1e3) / 1e9
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
                     "model": "56 B x nodes x sum of inner solves (SURVEY.md 8d); the RESIDENT kernel "
                              "keeps state on chip, so frac > 1 means it has left the HBM roofline"},
        "clocks": clocks, "gpu_launches": int(launches), "e2e": e2e,
    }
    if parity is not None:
        line["parity_sample"] = parity
    if streaming is not None:
        line["roofline_streaming_family"] = streaming
    if rank == 0 and world == 1 and not args.no_cpu:    # the CPU baseline is an N = 1 leg
        cores = os.cpu_count() or 1
        line["cpu_baseline"] = {k: v for k, v in cpu_baseline(args.cpu_sample or max(cores, 8)).items()
                                if k in ("value", "unit", "cores", "kind", "sample")}
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
