#!/usr/bin/env python
"""Benchmark of the ROMHighContrast hot path on B200.

One "step" = one pass of the batched FEM snapshot solver (GMG-preconditioned fp64 CG, libromhc.so) over the
K parameter vectors of this rank (BASELINE.json configs[2]: (4,4) subdomains, N=64 -> 256^2 cells, D=65025,
K=10000 random-contrast samples up to 1e6).  `value` = snapshot solves/s with y resident in HBM; `e2e` = the same
through the host-buffer C-ABI entry point romhc_generate_solutions_host (y from pinned host memory, U copied
back to pinned host memory every step).  The same JSON line carries the secondary metrics of the path (POD Gram
GEMM on the fp64 tensor cores, 1M online reduced Galerkin solves), the live roofline of the dominant solver
kernel and the CPU baseline (the oracle's SuperLU path on the host cores).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--k-snap 10000] [--quick]
"""
from __future__ import annotations

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

GEO, NPB = (4, 4), 64            # configs[2] of BASELINE.json
CMAX = 1e6
KIND_NAMES = ["k_pcg_p_apply", "k_pcg_update", "k_mg_down(l0)", "k_mg_down(l>=1)", "k_mg_tail", "k_mg_up(l0)",
              "k_mg_up(l>=1)"]
# algorithmic fp64 streams per fine-level DOF per launch (SURVEY 8d stream counting; DESIGN.md "kernels")
KIND_STREAMS = [1.5, 5.0, 1.75, 2.25 / 4, 2.0 / 16, 2.25, 3.25 / 4]   # z_A, z = M r and p travel as fp32 (half a stream each way) on the finest level


def workload_name(K):
    """config.workload -- the same string in both arms (the driver compares them)"""
    return (f"configs[2]: (4,4) subdomains, N=64 (256x256 cells, D=65025), {K} snapshots per GPU, contrast 10^U(0,6); "
            "snapshot solves to rtol 1e-12")


def sample_params(K, seed):
    return 10 ** np.random.default_rng(seed).uniform(0, np.log10(CMAX), size=(K,) + GEO)


# ----------------------------------------------------------------------------------------------------------
# CPU arm: the oracle's sparse direct path (scipy SuperLU), one worker per host core
# ----------------------------------------------------------------------------------------------------------
def _cpu_worker(args):
    geo, N, ys = args
    os.environ["OMP_NUM_THREADS"] = "1"
    from oracle import FEMOracle
    o = FEMOracle(geo, N)
    t0 = time.perf_counter()
    U = o.generate_solutions(ys)
    return time.perf_counter() - t0, float(np.abs(U).sum())


def cpu_snapshot_rate(per_core, cores=None, geo=GEO, N=NPB, seed=123):
    import multiprocessing as mp
    cores = cores or max(1, (os.cpu_count() or 1))
    ys = sample_params(per_core * cores, seed)
    chunks = [(geo, N, ys[i::cores]) for i in range(cores)]
    t0 = time.perf_counter()
    with mp.get_context("spawn").Pool(cores) as pool:
        pool.map(_cpu_worker, chunks)
    wall = time.perf_counter() - t0
    return len(ys) / wall, cores, len(ys), wall


def run_reference(args):
    """--impl reference: the reference algorithm's CPU path (oracle port: sparse assembly + SuperLU) on all host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = max(1, os.cpu_count() or 1)
    per_core = 8
    for _ in range(args.warmup if args.warmup is not None else 1):
        cpu_snapshot_rate(1, cores)
    rates, t_all = [], 0.0
    steps = args.steps or 2
    for s in range(steps):
        r, c, n, wall = cpu_snapshot_rate(per_core, cores, seed=1000 + s)
        rates.append(r); t_all += wall
    val = float(np.mean(rates))
    line = {
        "impl": "reference", "metric": "fem_snapshot_solves_per_s", "value": val, "unit": "solves/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": args.warmup if args.warmup is not None else 1,
        "ms_per_step": 1e3 * t_all / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(args.k_snap),
                   "note": "each step is a bounded sample of that workload (sparse assembly + SuperLU per system)"},
        "cpu_baseline": {"value": val, "unit": "solves/s", "cores": cores, "kind": "port",
                         "sample": f"{per_core * cores} snapshot solves per step (oracle: scipy CSR + SuperLU, "
                                   f"{cores} worker processes)"},
        "e2e": {"value": val, "unit": "solves/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.idx}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
                for nm, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                pass
        busy = [s for s in sm if s > 0]
        return {"sm_mhz": float(np.median(busy)) if busy else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=None)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--k-snap", type=int, default=10000, help="snapshot solves per GPU per step")
    ap.add_argument("--k-online", type=int, default=1000000)
    ap.add_argument("--k-obs", type=int, default=100000, help="configs[3]: observations in the estimation batch (whole job)")
    ap.add_argument("--n-rb", type=int, default=20)
    ap.add_argument("--quick", action="store_true", help="small sizes (debug)")
    ap.add_argument("--workload", default="config2", choices=["config2", "config4"],
                    help="config2 (default): the metric's configuration; config4: only the 512^2 / (8,8) / 100k-snapshot pipeline")
    ap.add_argument("--no-secondary", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-config4", action="store_true", help="skip the secondary configs[4] block (512^2 mesh, (8,8) subdomains)")
    ap.add_argument("--k-config4", type=int, default=12500, help="configs[4] snapshots per GPU (100k / 8)")
    ap.add_argument("--strip-kb", type=float, default=None, help="shared memory per strip CTA (tuning)")
    ap.add_argument("--nu", type=int, default=None, help="Gauss-Seidel sweeps of the V(nu,nu) cycle (tuning)")
    ap.add_argument("--nu-tail", type=int, default=None)
    ap.add_argument("--nu-mid", type=int, default=None)
    ap.add_argument("--threads", type=int, default=None)
    ap.add_argument("--tile", type=int, default=None, help="1: register-tiled multigrid kernels (default), 0: strip kernels")
    ap.add_argument("--tile-ty", type=int, default=None)
    ap.add_argument("--tile-prefetch", type=int, default=None)
    ap.add_argument("--tile-persistent", type=int, default=None)
    ap.add_argument("--fused", type=int, default=None, help="1: PCG update fused into the finest going-down kernel (default)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    steps = args.steps or 3
    warmup = args.warmup if args.warmup is not None else 3
    if args.quick:
        args.k_snap, args.k_online, args.k_config4, args.k_obs = 512, 100000, 256, 4000

    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a GPU (no CPU fallback)"
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from romhighcontrast_b200 import _lib
    from romhighcontrast_b200.engine import Engine

    if args.workload == "config4":
        def barrier4():
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
        c4 = run_config4(args, world, rank, local, barrier4, lambda: torch.cuda.Event(enable_timing=True))
        if rank == 0:
            snap = c4["snapshots"]
            print(json.dumps({"metric": "fem_snapshot_solves_per_s", "value": snap["solves_per_s"], "unit": "solves/s", "n_gpus": world,
                              "steps": 1, "warmup": 1, "ms_per_step": snap["ms"], "higher_is_better": True, "scaling": "weak",
                              "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": {"workload": c4["workload"]},
                              "gpu_launches": int(_lib.launch_count()), "secondary": {"config4": c4}}), flush=True)
        if world > 1:
            dist.destroy_process_group()
        return
    eng = Engine(GEO, NPB)
    if args.strip_kb:
        eng.set_option("strip_kb", args.strip_kb)
    if args.nu:
        eng.set_option("nu", args.nu)
    if args.nu_tail:
        eng.set_option("nu_tail", args.nu_tail)
    if args.nu_mid:
        eng.set_option("nu_mid", args.nu_mid)
    if args.threads:
        eng.set_option("threads", args.threads)
    if args.tile is not None:
        eng.set_option("tile", args.tile)
    if args.tile_ty:
        eng.set_option("tile_ty", args.tile_ty)
    if args.tile_prefetch is not None:
        eng.set_option("tile_prefetch", args.tile_prefetch)
    if args.tile_persistent is not None:
        eng.set_option("tile_persistent", args.tile_persistent)
    if args.fused is not None:
        eng.set_option("fused", args.fused)
    K = args.k_snap
    y_host = sample_params(K, seed=42 + rank)                   # this rank's shard of the training set
    y = eng.params(y_host)
    x = eng.empty(K, eng.Dp)
    ev = lambda: torch.cuda.Event(enable_timing=True)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- resident: warm-up, then `steps` timed passes -----------------------------------------------------
    for _ in range(warmup):
        eng.solve(y, out=x)
    eng.set_option("profile", 1)
    sampler = ClockSampler(local)
    sampler.start()
    launches0 = _lib.launch_count()
    barrier()
    e0, e1 = ev(), ev()
    e0.record()
    for _ in range(steps):
        _, iters, relres = eng.solve(y, out=x)
    e1.record()
    barrier()
    ms_total = e0.elapsed_time(e1)
    launches = _lib.launch_count() - launches0
    eng.set_option("profile", 0) if False else None
    import ctypes as C
    pms, pn = (C.c_double * 8)(), (C.c_int64 * 8)()
    _lib.check(eng.lib.romhc_get_profile(eng.handle, pms, pn))
    eng.set_option("profile", 0)
    tmax = torch.tensor([ms_total], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    ms_step = float(tmax.item()) / steps
    value = world * K / (ms_step * 1e-3)
    it_np = iters.cpu().numpy()
    stats = dict(eng.last_solve_stats)

    # the all-fp64 figure next to the default (fp32 TRANSPORT of z, z_A and p on the finest level; arithmetic, x, r and
    # every reduction are fp64): one timed pass with z32 = 0
    eng.set_option("z32", 0)
    eng.solve(y, out=x)
    barrier()
    e0, e1 = ev(), ev()
    e0.record(); eng.solve(y, out=x); e1.record(); torch.cuda.synchronize()
    t64 = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t64, op=dist.ReduceOp.MAX)
    eng.set_option("z32", 3)
    precision_note = ("fp64 arithmetic, iterate, residual and reductions; fp32 transport of z, z_A, p between kernels on the "
                      "finest level (z32=3); all-fp64 transport (z32=0): %.0f solves/s" % (world * K / (float(t64.item()) * 1e-3)))
    eng.solve(y, out=x)                                       # leave the default-mode solution in x

    # ---- end to end through the host-buffer C ABI (pinned host memory both ways) --------------------------------
    U_pin = torch.empty((K, eng.D), dtype=torch.float64, pin_memory=True)
    y_pin = torch.from_numpy(np.ascontiguousarray(y_host.reshape(K, -1))).pin_memory()
    U_np, y_np = U_pin.numpy(), y_pin.numpy()
    eng.generate_solutions_host(y_np, out=U_np)                 # warm-up (allocations)
    barrier()
    t0 = time.perf_counter()
    e2e_steps = max(1, min(steps, 2))
    for _ in range(e2e_steps):
        eng.generate_solutions_host(y_np, out=U_np)
    torch.cuda.synchronize()
    t_e2e = torch.tensor([(time.perf_counter() - t0) / e2e_steps], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
    e2e_val = world * K / float(t_e2e.item())
    clocks = sampler.stop()
    # the ceiling of e2e on this box: the step's D2H bytes as one plain pinned cudaMemcpyAsync per rank, all ranks at once
    # (the host side of N GPUs shares its PCIe / memory fabric); e2e cannot beat max(resident step, that copy)
    Ud = eng.unpad(x)                                           # the same values the e2e call just delivered: U_pin stays valid
    d2h = []
    for _ in range(3):
        barrier()
        t0 = time.perf_counter()
        U_pin.copy_(Ud, non_blocking=True)
        torch.cuda.synchronize()
        td = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(td, op=dist.ReduceOp.MAX)
        d2h.append(float(td.item()))
    del Ud
    t_d2h = min(d2h[1:])
    ceiling = world * K / max(t_d2h, ms_step * 1e-3)
    e2e_ceiling = {"d2h_s_all_ranks_concurrent": t_d2h, "d2h_GBps_aggregate": world * U_pin.numel() * 8 / t_d2h / 1e9,
                   "ceiling_solves_per_s": ceiling, "e2e_frac_of_ceiling": e2e_val / ceiling}
    # the same through the reference-facing class API (numpy in, fresh pageable numpy out): src.lib mirror -> C ABI
    api_val = None
    try:
        from lib.SolutionsManagers import SolutionsManagerFEM
        sm = SolutionsManagerFEM(GEO, NPB, method="lsqsparse")
        sm.generate_solutions(y_host[:256])
        sm.generate_solutions(y_host)                               # first full call allocates the pinned bounce buffers
        barrier()                                                   # all ranks copy at the same time, as in the e2e leg above
        t0 = time.perf_counter()
        U_api = sm.generate_solutions(y_host)
        t_api = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t_api, op=dist.ReduceOp.MAX)
        api_val = world * K / float(t_api.item())
        del U_api, sm
    except MemoryError:
        api_val = None

    # ---- parity spot check of what was just timed (sub-sample against the CPU oracle) ------------------------------
    parity = None
    if rank == 0:
        from oracle import FEMOracle
        o = FEMOracle(GEO, NPB)
        sel = [0, K // 2, K - 1]
        Uo = o.generate_solutions(y_host[sel])
        parity = float(np.max(np.linalg.norm(U_np[sel] - Uo, axis=1) / np.linalg.norm(Uo, axis=1)))

    # ---- roofline of the dominant solver kernel (live CUDA-event times of the first iterations, all systems active)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak, peak_src = (peaks.get("hbm_gbs"), "measured (MEASURED_PEAKS.json)") if peaks.get("hbm_gbs") else (6650.0, "fallback")
    per_kind = []
    for i, nm in enumerate(KIND_NAMES):
        if pn[i] > 0:
            avg_ms = pms[i] / pn[i]
            streams = KIND_STREAMS[i]
            if nm == "k_mg_down(l0)" and pn[1] == 0:
                streams += 3.5                                    # fused with the PCG update: R p (fp32), x; W x, r on top
                nm = "k_mg_update_down(l0)"
            if nm == "k_mg_tail":
                streams = 2.0 / 4.0 ** eng.tail_level            # reads r, writes z of its first level
            if nm.endswith("(l>=1)") and pn[2] > 0:
                nl = max(1, int(round(pn[i] / pn[2])))          # levels 1..nl share this kind: mean bytes per launch
                streams = KIND_STREAMS[i] * sum(4.0 ** -(l - 1) for l in range(1, nl + 1)) / nl
            gb = streams * 8.0 * eng.D * K / 1e9
            per_kind.append({"kernel": nm, "launches": int(pn[i]), "avg_ms": avg_ms, "share": 0.0,
                             "algorithmic_GB": gb, "GBps": gb / (avg_ms * 1e-3)})
    # one PCG iteration = every kind weighted by its launches per iteration (the level >= 1 kinds run once per level)
    n_it = max(1, int(pn[0]) if pn[0] > 0 else int(pn[2]))
    tot_ms = sum(k["avg_ms"] * k["launches"] for k in per_kind) / n_it or 1.0
    tot_gb = sum(k["algorithmic_GB"] * k["launches"] for k in per_kind) / n_it
    for k in per_kind:
        k["share"] = k["avg_ms"] * k["launches"] / n_it / tot_ms
    dom = max(per_kind, key=lambda k: k["avg_ms"]) if per_kind else None
    roofline = None
    if dom:
        # DRAM bytes per launch of the same kernel from the committed ncu pass (same workload, K = 10000); null otherwise
        traffic = None
        try:
            tr = json.load(open(os.path.join(ROOT, "profiles", "r2_dram_traffic_k10000.json")))
            ncu_name = {"k_mg_update_down(l0)": "k_mgp_update_down<64>",   # (keys of the JSON drop the second template argument) "k_mg_down(l0)": "k_mgp_down<64>",
                        "k_mg_up(l0)": "k_mgp_up<64>"}.get(dom["kernel"], dom["kernel"])
            if tr.get("K") == K and ncu_name in tr["kernels"] and all(v is None for v in (args.strip_kb, args.nu, args.nu_mid, args.nu_tail, args.tile, args.tile_ty, args.fused)):
                traffic = tr["kernels"][ncu_name]["dram_bytes_per_launch"] / 1e9
        except Exception:
            traffic = None
        roofline = {"bound": "hbm", "kernel": dom["kernel"], "achieved": dom["GBps"], "peak": hbm_peak, "unit": "GB/s",
                    "frac": dom["GBps"] / hbm_peak, "traffic": traffic, "traffic_unit": "GB per launch (ncu dram__bytes_read+write)",
                    "algorithmic_GB_per_launch": dom["algorithmic_GB"], "peak_source": peak_src,
                    "whole_iteration_GBps": tot_gb / (tot_ms * 1e-3), "whole_iteration_ms": tot_ms,
                    "whole_iteration_algorithmic_GB": tot_gb, "whole_iteration_frac": tot_gb / (tot_ms * 1e-3) / hbm_peak}

    secondary = {}
    if not args.no_secondary:
        secondary = run_secondary(eng, x, y, K, args, world, rank, barrier, ev, U_np, y_host)

    if not args.no_secondary and rank == 0:
        try:
            secondary["config0"] = run_config0()
        except Exception as exc:
            secondary["config0"] = {"error": repr(exc)[:300]}
    if not args.no_secondary and not args.no_config4:
        del x, y
        eng = None
        import gc
        gc.collect()
        torch.cuda.empty_cache()
        try:
            secondary["config4"] = run_config4(args, world, rank, local, barrier, ev)
        except Exception as exc:
            if world > 1:
                raise
            secondary["config4"] = {"error": repr(exc)[:300]}

    cpu = None
    if rank == 0 and not args.no_cpu:
        r, cores, n, wall = cpu_snapshot_rate(32, None)          # ~10 s of wall time on all host cores
        cpu = {"value": r, "unit": "solves/s", "cores": cores, "kind": "port",
               "sample": f"{n} snapshot solves of the same workload (oracle: scipy CSR + SuperLU, {cores} processes, {wall:.1f} s)"}
        try:      # the UNMODIFIED reference on configs[0], timed where /root/reference exists (oracle/time_reference.py)
            cpu["true_reference_config0"] = json.load(open(os.path.join(ROOT, "tests", "golden", "reference_timing_config0.json")))
        except Exception:
            cpu["true_reference_config0"] = None

    if rank == 0:
        line = {
            "metric": "fem_snapshot_solves_per_s", "value": value, "unit": "solves/s", "n_gpus": world, "steps": steps,
            "warmup": warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(K),
                       "solver": "batched GMG-preconditioned CG, V(2,2)/(3,3)/(4,4), rtol 1e-12 on sqrt(r.z / r0.z0)",
                       "precision": precision_note,
                       "l2": "inputs larger than L2 (%.1f GB working set per step)" % (stats["workspace_bytes"] / 1e9),
                       "tuning": {"strip_kb": args.strip_kb, "nu": args.nu, "nu_tail": args.nu_tail, "threads": args.threads},
                       "pcg_iterations": {"min": int(it_np.min()), "mean": float(it_np.mean()), "max": int(it_np.max())},
                       "parity_rel_l2_vs_oracle": parity},
            "clocks": clocks,
            "e2e": {"value": e2e_val, "unit": "solves/s", "h2d_bytes_per_step": int(y_np.nbytes),
                    "d2h_bytes_per_step": int(U_np.nbytes + K * 12),
                    "path": "romhc_generate_solutions_host (C ABI), pinned host buffers", "ceiling": e2e_ceiling,
                    "class_api_value": api_val, "class_api_frac_of_e2e": (api_val / e2e_val) if api_val else None,
                    "class_api_path": "SolutionsManagerFEM.generate_solutions: numpy in, numpy out (result in a pooled pinned block)"},
            "gpu_launches": int(launches),
            "roofline": roofline, "kernels": per_kind, "cpu_baseline": cpu, "secondary": secondary,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_greedy(U_np, y_host, n):
    """Wall time of the reference-facing greedy builders (ReducedBasis.py:112-139) on this rank's snapshots: numpy in,
    basis out; includes the H2D copy of the (K, D) snapshot matrix (the device-resident figure is in
    secondary.distributed.greedy_sharded)."""
    import torch
    from lib.ReducedBasis import ReducedBasisGreedy, GREEDY_FOR_GALERKIN, GREEDY_FOR_H10
    from lib.SolutionsManagers import SolutionsManagerFEM
    sm = SolutionsManagerFEM(GEO, NPB, method="lsqsparse")
    h1 = sm.H10norm(U_np)
    ReducedBasisGreedy().build(n=2, sm=sm, solutions2train=U_np, a2train=y_host, solutions2train_h1norm=h1)   # one-time allocations
    out = {"K": int(len(U_np)), "n": n}
    for name, crit in (("galerkin", GREEDY_FOR_GALERKIN), ("h10", GREEDY_FOR_H10)):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        rb = ReducedBasisGreedy(greedy_for=crit).build(n=n, sm=sm, solutions2train=U_np, a2train=y_host,
                                                       solutions2train_h1norm=h1)
        torch.cuda.synchronize()
        out[name + "_s"] = time.perf_counter() - t0
        out[name + "_selected_head"] = [int(i) for i in rb.selected_indices[:6]]
        out[name + "_max_rel_error_last"] = float(rb.max_errors[-1])
    return out


def run_distributed(eng, x, y_dev, y_host, K, args, world, rank, barrier, ev):
    """POD (both routes) and the greedy builders on a training set sharded over the ranks (contiguous slices, rank r owns
    [b_r, b_{r+1}) of the union): dist.distributed_pca(method="gram"): all_to_all -> partial centred SYRK -> all_reduce
    of the K x K Gram -> replicated eigensolve -> all_gather of component slices; method="krylov": Gram-free block
    Lanczos with one (32, Dp) all_reduce per step; dist.greedy_build_sharded: all_gather of (error, index) pairs +
    broadcast of the winner.  Every collective is warmed (first repetition discarded) and the best of 3 is reported;
    times are max over ranks.  At world == 1 the same code runs without collectives (the N = 1 point of the scaling)."""
    import torch
    import torch.distributed as dist
    from romhighcontrast_b200 import dist as rd
    from lib.ReducedBasis import ReducedBasisGreedy, GREEDY_FOR_GALERKIN, GREEDY_FOR_H10
    from lib.SolutionsManagers import SolutionsManagerFEM
    n = args.n_rb
    out = {}

    def maxr(v):
        t = torch.tensor([float(v)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- POD of a 10k-snapshot union (configs[2]'s POD, strong scaling over the ranks) ------------------------------
    K_pod = int(min(10000, world * K))
    b = rd.shard_bounds(K_pod, world)
    Kl = b[rank + 1] - b[rank]
    counts = [b[r + 1] - b[r] for r in range(world)]
    stage_best, tot_ms = {}, []
    for rep in range(4):
        Xl = x[:Kl].clone()
        tm = {}
        barrier()
        e0, e1 = ev(), ev()
        e0.record()
        comps_g, sig_g = rd.distributed_pca(eng, Xl, n, counts=counts, timings=tm, method="gram")
        e1.record(); torch.cuda.synchronize()
        if rep:                                                    # repetition 0 warms NCCL channels and allocations
            tot_ms.append(maxr(e0.elapsed_time(e1)))
            for k_, v_ in tm.items():
                v_ = maxr(v_)
                stage_best[k_] = min(stage_best.get(k_, v_), v_)
        del Xl
    gram_bytes = 8.0 * K_pod * K_pod
    out["pod_gram"] = {"K_total": K_pod, "rows_per_rank": Kl, "n": n, "ms": min(tot_ms), "stages_ms": stage_best,
                       "gram_allreduce_bytes": gram_bytes if world > 1 else 0,
                       "gram_allreduce_busbw_GBps": (gram_bytes * 2 * (world - 1) / world / (stage_best["gram_allreduce"] * 1e-3) / 1e9
                                                     if world > 1 and stage_best.get("gram_allreduce") else None),
                       "all_to_all_bytes_per_rank": 8.0 * Kl * eng.Dp * (world - 1) / world,
                       "singular_values_head": [float(v) for v in sig_g[:5].cpu()]}
    kry_ms, kst = [], {}
    for rep in range(3):
        Xl = x[:Kl].clone()
        kst = {}
        barrier()
        e0, e1 = ev(), ev()
        e0.record()
        comps_k, sig_k = rd.distributed_pca(eng, Xl, n, counts=counts, timings=kst, method="krylov")
        e1.record(); torch.cuda.synchronize()
        if rep:
            kry_ms.append(maxr(e0.elapsed_time(e1)))
        del Xl
    sv_diff = float(((sig_k - sig_g).abs() / sig_g).max())
    sub_diff = float((comps_k - comps_g).abs().max())
    out["pod_krylov"] = {"K_total": K_pod, "n": n, "ms": min(kry_ms), **{k_: v_ for k_, v_ in kst.items()},
                         "sv_rel_diff_vs_gram_route": sv_diff, "components_max_abs_diff_vs_gram_route": sub_diff,
                         "agree_1e-9": bool(sv_diff <= 1e-9)}
    del comps_g, comps_k
    # ---- Gram-free POD of the WHOLE union (world x K rows, weak scaling: the configs[4] pattern) --------------------
    if world > 1:
        kw_ms, kst = [], {}
        for rep in range(2):
            Xl = x.clone()
            kst = {}
            barrier()
            e0, e1 = ev(), ev()
            e0.record()
            _, sig_w = rd.distributed_pca(eng, Xl, n, counts=[K] * world, timings=kst, method="krylov")
            e1.record(); torch.cuda.synchronize()
            if rep:
                kw_ms.append(maxr(e0.elapsed_time(e1)))
            del Xl
        out["pod_krylov_whole_union"] = {"K_total": world * K, "n": n, "ms": min(kw_ms), **kst,
                                         "singular_values_head": [float(v) for v in sig_w[:5].cpu()]}
    # ---- sharded greedy on the same 10k union, both criteria; snapshots stay where the solver left them ------------------
    sm = SolutionsManagerFEM(GEO, NPB, method="lsqsparse")
    sm.__dict__["_engine"] = eng
    U_loc, a_loc = x[:Kl].contiguous(), y_dev[:Kl].contiguous()
    h1_loc = eng.h10_norm(U_loc)
    g = {"K_total": K_pod, "n": n, "input": "device-resident padded snapshots (no PCIe)"}
    for name, crit in (("galerkin", GREEDY_FOR_GALERKIN), ("h10", GREEDY_FOR_H10)):
        best = None
        for rep in range(3):
            barrier()
            t0 = time.perf_counter()
            _, _, picked = rd.greedy_build_sharded(sm, n, U_loc, a_loc, h1_loc, K_pod, greedy_for=crit)
            torch.cuda.synchronize()
            dt = maxr(time.perf_counter() - t0)
            if rep:
                best = dt if best is None else min(best, dt)
        tm = {}
        rd.greedy_build_sharded(sm, n, U_loc, a_loc, h1_loc, K_pod, greedy_for=crit, timings=tm)   # stage-synchronised pass
        g[name] = {"s": best, "selected_head": [int(i) for i in picked[:6]], "stages_ms_sum_over_steps_synchronised": tm}
    # index parity: the sharded build must pick exactly what one GPU picks on the gathered set (sub-sample)
    K_chk = int(min(2048, K_pod))
    bc = rd.shard_bounds(K_chk, world)
    kl = bc[rank + 1] - bc[rank]
    Uc, ac = eng.unpad(x[:kl].contiguous()).cpu().numpy(), np.ascontiguousarray(y_host[:kl])
    hc = sm.H10norm(Uc)
    if world > 1:
        parts = [None] * world
        dist.all_gather_object(parts, (Uc, ac, hc))
    else:
        parts = [(Uc, ac, hc)]
    same = True
    for name, crit in (("galerkin", GREEDY_FOR_GALERKIN), ("h10", GREEDY_FOR_H10)):
        _, _, picked = rd.greedy_build_sharded(sm, 10, Uc, ac, hc, K_chk, greedy_for=crit)
        if rank == 0:
            ref = ReducedBasisGreedy(greedy_for=crit).build(n=10, sm=sm, solutions2train=np.vstack([p_[0] for p_ in parts]),
                                                            a2train=np.concatenate([p_[1] for p_ in parts]),
                                                            solutions2train_h1norm=np.concatenate([p_[2] for p_ in parts]))
            same &= [int(i) for i in picked] == [int(i) for i in ref.selected_indices]
    g["indices_equal_single_gpu_on_subsample"] = {"K": K_chk, "n": 10, "equal": bool(same)}
    out["greedy_sharded"] = g
    barrier()
    return out


def run_config4(args, world, rank, local, barrier, ev):
    """BASELINE configs[4]: (8,8) subdomains, N = 64 (512 x 512 cells, D = 261 121), a 100k-snapshot training set sharded
    over the GPUs of the box: snapshots (no collective) -> Gram-free block-Lanczos POD (one (32, Dp) all_reduce per
    step) -> sharded greedy, both criteria (argmax all_gather + winner all_reduce per step) -> 1M online reduced
    Galerkin solves.  K per GPU is bounded at 12 500 = 100 000 / 8: at 8 GPUs this IS the 100k set, on fewer GPUs a
    12 500-per-GPU subset of it (one GPU cannot hold 100k x 261k doubles = 209 GB).  Snapshots never leave the device."""
    import torch
    import torch.distributed as dist
    from romhighcontrast_b200 import dist as rd
    from romhighcontrast_b200.engine import Engine
    from lib.ReducedBasis import GREEDY_FOR_GALERKIN, GREEDY_FOR_H10
    from lib.SolutionsManagers import SolutionsManagerFEM
    geo, N, n = (8, 8), 64, args.n_rb
    K = int(args.k_config4)
    out = {"workload": f"configs[4]: (8,8) subdomains, N=64 (512x512 cells, D=261121), {K} snapshots per GPU "
                       f"({world * K} in total), contrast 10^U(0,6)"}

    def maxr(v):
        t = torch.tensor([float(v)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    eng = Engine(geo, N)
    y_host = 10 ** np.random.default_rng(4000 + rank).uniform(0, np.log10(CMAX), size=(K,) + geo)
    y = eng.params(y_host)
    x = eng.empty(K, eng.Dp)
    eng.set_option("workspace_gb", 72)                             # 180 GB of HBM: 26 GB of snapshots (+ a centred copy for the POD) + 4800-system solver chunks
    kw = int(min(K, 32768, (72 << 30) // eng.solve_bytes_per_system))   # exactly the chunk Context::solve will use
    eng.solve(y[:kw], out=x[:kw])                                  # warm-up at the full chunk size: workspace allocation, kernel attributes
    barrier()
    e0, e1 = ev(), ev()
    e0.record()
    _, iters, relres = eng.solve(y, out=x)
    e1.record(); barrier()
    ms = maxr(e0.elapsed_time(e1))
    it = iters.double()
    out["snapshots"] = {"solves_per_s": world * K / (ms * 1e-3), "ms": ms, "K_per_gpu": K,
                        "dof_solves_per_s": world * K * eng.D / (ms * 1e-3),
                        "pcg_iterations": {"min": int(it.min()), "mean": float(it.mean()), "max": int(it.max())},
                        "max_relres": float(relres.max()), "chunks": eng.last_solve_stats["chunks"]}
    if rank == 0 and not args.no_cpu:                              # parity of what was timed: 3 systems against the CPU oracle
        from oracle import FEMOracle
        sel = [0, K // 2, K - 1]
        t0 = time.perf_counter()
        Uo = FEMOracle(geo, N).generate_solutions(y_host[sel])
        cpu_s = (time.perf_counter() - t0) / len(sel)
        U = eng.unpad(x[sel].contiguous()).cpu().numpy()
        out["snapshots"]["parity_rel_l2_vs_oracle"] = float(np.max(np.linalg.norm(U - Uo, axis=1) / np.linalg.norm(Uo, axis=1)))
        out["snapshots"]["oracle_seconds_per_system_1core"] = cpu_s
    # Gram-free POD of the sharded set
    kst, kms = {}, []
    for rep in range(2):
        Xc = x.clone()
        kst = {}
        barrier()
        e0, e1 = ev(), ev()
        e0.record()
        comps, sig = rd.distributed_pca(eng, Xc, n, counts=[K] * world, timings=kst, method="krylov")
        e1.record(); torch.cuda.synchronize()
        kms.append(maxr(e0.elapsed_time(e1)))
        del Xc
    out["pod_krylov"] = {"K_total": world * K, "n": n, "ms": min(kms), "first_call_ms": kms[0], **kst,
                         "singular_values_head": [float(v) for v in sig[:5].cpu()]}
    # sharded greedy on the resident snapshots
    sm = SolutionsManagerFEM(geo, N, method="lsqsparse")
    sm.__dict__["_engine"] = eng
    h1 = eng.h10_norm(x)
    g = {"K_total": world * K, "n": n}
    for name, crit in (("galerkin", GREEDY_FOR_GALERKIN), ("h10", GREEDY_FOR_H10)):
        best, picked = None, None
        for rep in range(2):
            barrier()
            t0 = time.perf_counter()
            _, _, picked = rd.greedy_build_sharded(sm, n, x, y, h1, world * K, greedy_for=crit)
            torch.cuda.synchronize()
            dt = maxr(time.perf_counter() - t0)
            best = dt if best is None else min(best, dt)
        g[name] = {"s": best, "selected_head": [int(i) for i in picked[:6]]}
    out["greedy_sharded"] = g
    # online stage on the POD basis: 1M reduced Galerkin solves over all ranks
    Ahat, bhat = eng.project_operators(comps.contiguous())
    Ko = args.k_online // world
    yo = eng.params(10 ** np.random.default_rng(4100 + rank).uniform(0, np.log10(CMAX), size=(Ko,) + geo))
    eng.reduced_solve(yo, Ahat, bhat, check=False)
    barrier()
    e0, e1 = ev(), ev()
    e0.record()
    for _ in range(5):
        Cc = eng.reduced_solve(yo, Ahat, bhat, check=False)
    e1.record(); torch.cuda.synchronize()
    oms = maxr(e0.elapsed_time(e1) / 5)
    out["reduced_galerkin"] = {"K": world * Ko, "n": n, "nb": 64, "ms": oms, "solves_per_s": world * Ko / (oms * 1e-3)}
    del eng, x
    torch.cuda.empty_cache()
    return out


def run_config0():
    """BASELINE configs[0] through the reference-shaped class API (numpy in, numpy out), the same calls and inputs that
    oracle/time_reference.py times on the unmodified reference (tests/golden/reference_timing_config0.json)."""
    import torch
    from lib.ReducedBasis import ReducedBasisGreedy, ReducedBasisPCA, GREEDY_FOR_GALERKIN, GREEDY_FOR_H10
    from lib.SolutionsManagers import SolutionsManagerFEM
    geo, N, K, n = (2, 2), 32, 100, 10
    y = 10 ** np.random.default_rng(42).uniform(0, 6, (K,) + geo)
    T = {}

    def tm(name, f, *a, **kw):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        r = f(*a, **kw)
        torch.cuda.synchronize()
        T[name] = time.perf_counter() - t0
        return r

    for rep in range(2):                                           # repetition 0 pays one-time allocations
        sm = tm("assembly_s", SolutionsManagerFEM, geo, N, num_cores=1, method="lsqsparse")
        U = tm("snapshots_s", sm.generate_solutions, y)
        h1 = tm("h10norm_s", sm.H10norm, U)
        rbg = tm("greedy_galerkin_s", ReducedBasisGreedy(greedy_for=GREEDY_FOR_GALERKIN).build, n=n, sm=sm, solutions2train=U,
                 a2train=y, solutions2train_h1norm=h1)
        tm("greedy_h10_s", ReducedBasisGreedy(greedy_for=GREEDY_FOR_H10).build, n=n, sm=sm, solutions2train=U, a2train=y,
           solutions2train_h1norm=h1)
        tm("pca_s", ReducedBasisPCA().build, n=n, sm=sm, solutions2train=U, a2train=y)
        rbg.orthonormalize()
        tm("forward_modeling_s", rbg.forward_modeling, sm, y)
        tm("projection_s", rbg.projection, sm, U)
    out = {"workload": "configs[0]: (2,2) subdomains, N=32 (64x64 cells, D=3969), 100 snapshots, n=10; class API, numpy in/out",
           "seconds": T, "snapshot_solves_per_s": K / T["snapshots_s"], "greedy_galerkin_selected": [int(i) for i in rbg.selected_indices]}
    try:
        ref = json.load(open(os.path.join(ROOT, "tests", "golden", "reference_timing_config0.json")))
        out["true_reference_seconds"] = ref["seconds"]
        out["speedup_vs_true_reference"] = {k_: ref["seconds"][k_] / T[k_] for k_ in T if k_ in ref["seconds"] and T[k_] > 0}
        out["true_reference_host"] = ref["host"]
    except Exception:
        pass
    return out


def run_config3(eng, x, y_dev, y_host, args, world, rank, barrier):
    """BASELINE configs[3]: state + parameter estimation (least squares in the reduced space and the PBDW correction,
    m = 50 point measurements) over a 100 000-observation batch on the (4,4), N = 64 model, the batch sharded over the
    ranks (no collective).  Every observation is solved, measured, estimated and scored on the device."""
    import torch
    import torch.distributed as dist
    from lib.ReducedBasis import ReducedBasisGreedy
    from lib.SolutionsManagers import SolutionsManagerFEM
    from romhighcontrast_b200.inverse import observation_batch_estimation
    n, m = args.n_rb, 50
    Kobs = int(args.k_obs) // world
    sm = SolutionsManagerFEM(GEO, NPB, method="lsqsparse")
    sm.__dict__["_engine"] = eng
    Ktr = min(1000, x.shape[0])
    Utr, ytr = x[:Ktr].contiguous(), y_dev[:Ktr].contiguous()
    rb = ReducedBasisGreedy().build(n=n, sm=sm, solutions2train=Utr, a2train=ytr, solutions2train_h1norm=eng.h10_norm(Utr))
    pts = np.random.default_rng(1).uniform(low=[sm.x_domain[0], sm.y_domain[0]], high=[sm.x_domain[1], sm.y_domain[1]], size=(m, 2))
    yobs = sample_params(Kobs, seed=4400 + rank)
    observation_batch_estimation(sm, rb, pts, yobs[:2000])             # warm-up
    barrier()
    t0 = time.perf_counter()
    res = observation_batch_estimation(sm, rb, pts, yobs)
    torch.cuda.synchronize()
    dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    tm = {}
    observation_batch_estimation(sm, rb, pts, yobs[:20000], timings=tm)   # stage-synchronised pass on a fifth of the batch
    return {"workload": f"configs[3]: {world * Kobs} observations, m={m} point measurements, n={n} greedy basis (trained on {Ktr} snapshots), "
                        "(4,4) subdomains, N=64; solve + measure + LS state estimate + PBDW correction + parameter estimators",
            "observations_per_s": world * Kobs / float(dt.item()), "s": float(dt.item()),
            "stage_seconds_on_20000": tm,
            "median_rel_H10_error_ls": float(np.median(res["err_ls"])), "median_rel_H10_error_pbdw": float(np.median(res["err_pbdw"])),
            "median_abs_rel_param_error_inverse": float(np.median(np.abs(1 - res["a_inverse"] / yobs)))}


def run_secondary(eng, x, y, K, args, world, rank, barrier, ev, U_np=None, y_host=None):
    """POD (centred Gram on the fp64 tensor cores), the greedy builders and 1M online reduced Galerkin solves on the
    snapshots in `x`."""
    import torch
    import torch.distributed as dist
    out = {}
    n = args.n_rb
    # the two communicating stages on the K-sharded union of all ranks' snapshots -- before anything below
    # centres the resident snapshots in place
    try:
        out["distributed"] = run_distributed(eng, x, y, y_host, K, args, world, rank, barrier, ev)
    except Exception as exc:
        import traceback
        out["distributed"] = {"error": repr(exc)[:300], "trace": traceback.format_exc()[-600:]}
        if world > 1:
            raise                                                 # a rank that left a collective early would hang the others
    try:
        out["config3"] = run_config3(eng, x, y, y_host, args, world, rank, barrier)
    except Exception as exc:
        out["config3"] = {"error": repr(exc)[:300]}
        if world > 1:
            raise
    if U_np is not None:
        try:
            out["greedy"] = run_greedy(U_np, y_host, n)
        except MemoryError:
            pass
    # PCG iteration counts against the contrast of the coefficient field (512 systems each, 10^U(0, log10 cmax))
    sweep = {}
    for cmax in (1e0, 1e2, 1e4, 1e6, 1e8, 1e10):
        yc = np.ones((512,) + GEO) if cmax == 1.0 else 10 ** np.random.default_rng(7).uniform(0, np.log10(cmax), (512,) + GEO)
        _, itc, _ = eng.solve(eng.params(yc))
        sweep["%g" % cmax] = {"mean": float(itc.double().mean()), "max": int(itc.max())}
    out["pcg_iterations_vs_contrast"] = sweep
    # measured fp64 GEMM peak of this GPU (cuBLAS DGEMM 8192^3) as the tensor-pipe denominator
    a = torch.randn(8192, 8192, dtype=torch.float64, device="cuda")
    b = torch.randn(8192, 8192, dtype=torch.float64, device="cuda")
    torch.matmul(a, b)
    best = 1e9
    for _ in range(3):
        e0, e1 = ev(), ev()
        e0.record(); torch.matmul(a, b); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    dgemm_peak = 2 * 8192 ** 3 / (best * 1e-3) / 1e12
    del a, b
    # POD: mean, centre in place, Gram (lower tiles), top-n eigenpairs, back-projection
    Kg = min(K, 10000)
    X = x[:Kg]
    mean = eng.column_mean(X)
    eng.center_rows_(X, mean)
    G = eng.gemm_nt(X, X, symmetric=True)        # warm-up
    barrier()
    gram_ms = []
    for _ in range(3):
        e0, e1 = ev(), ev()
        e0.record()
        G = eng.gemm_nt(X, X, symmetric=True)
        e1.record(); torch.cuda.synchronize()
        gram_ms.append(e0.elapsed_time(e1))
    ms = min(gram_ms)
    flop = float(Kg) * (Kg + 1) * eng.D            # triangle only, algorithmic D
    out["gram"] = {"K": Kg, "ms": ms, "TFLOPs": flop / (ms * 1e-3) / 1e12, "peak_TFLOPs_cublas_dgemm_8192": dgemm_peak,
                   "frac": flop / (ms * 1e-3) / 1e12 / dgemm_peak, "flop_counted": "K(K+1)D (lower triangle)", "all_ms": gram_ms}
    from romhighcontrast_b200.pod import top_eigenpairs
    pod_ms = []
    for _ in range(3):       # the first pass pays torch's one-time cuSOLVER initialisation (QR of the Lanczos blocks)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        lam, V = top_eigenpairs(eng, G, n)
        comps = eng.gemm_tn(V, X) / torch.sqrt(lam)[:, None]
        torch.cuda.synchronize()
        pod_ms.append(1e3 * (time.perf_counter() - t0))
    out["pod_eig_backproject_ms"] = min(pod_ms[1:])
    out["pod_eig_backproject_first_call_ms"] = pod_ms[0]
    out["pod_singular_values_head"] = [float(v) for v in torch.sqrt(lam)[:5].cpu()]
    # the Gram-free route (block Lanczos on the rows of X, pod.krylov_pca) on the same centred snapshots, rank-local
    try:
        from romhighcontrast_b200.pod import krylov_pca
        kry_ms, kst = [], {}
        for _ in range(3):
            e0, e1 = ev(), ev()
            e0.record()
            ck, sk, _ = krylov_pca(eng, X, n, center_in_place=True, stats=kst, distributed=False)
            e1.record(); torch.cuda.synchronize()
            kry_ms.append(e0.elapsed_time(e1))
        sg = torch.sqrt(lam)
        W32 = ck.new_zeros(32, ck.shape[1]); W32[:ck.shape[0]] = ck
        e0, e1 = ev(), ev()
        e0.record()
        for _ in range(5):
            Zs = eng.gemm_tn(eng.gemm_nt(X, W32, splitk=True), X)
        e1.record(); torch.cuda.synchronize()
        ap_ms = e0.elapsed_time(e1) / 5
        out["pod_krylov"] = {"K": Kg, "n": n, "ms": min(kry_ms[1:]), "first_call_ms": kry_ms[0],
                             "gram_route_ms": out["gram"]["ms"] + out["pod_eig_backproject_ms"],
                             "steps": kst.get("steps"), "krylov_dim": kst.get("krylov_dim"),
                             "sv_rel_diff_vs_gram_route": float(((sk - sg).abs() / sg).max()),
                             "apply_S_ms": ap_ms, "apply_S_TFLOPs": 4.0 * Kg * X.shape[1] * 32 / (ap_ms * 1e-3) / 1e12,
                             "apply_S_frac_of_cublas_dgemm": 4.0 * Kg * X.shape[1] * 32 / (ap_ms * 1e-3) / 1e12 / dgemm_peak}
    except Exception as exc:                                   # a secondary metric must never take the bench line down
        out["pod_krylov"] = {"error": repr(exc)[:300]}
    # online stage: reduced operators once, then k_online reduced Galerkin solves
    Ahat, bhat = eng.project_operators(comps.contiguous())
    Ko = args.k_online
    yo = eng.params(sample_params(Ko, seed=43 + rank))
    Cc = eng.reduced_solve(yo, Ahat, bhat, check=False)
    barrier()
    e0, e1 = ev(), ev()
    e0.record()
    reps = 5
    for _ in range(reps):
        Cc = eng.reduced_solve(yo, Ahat, bhat, check=False)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    if world > 1:                                              # max over ranks, like every multi-GPU number of the line
        tm = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        ms = float(tm.item())
    nb = eng.nb
    flop_per = 2 * nb * n * (n + 1) / 2 + n ** 3 / 3 + 2 * n * n
    out["reduced_galerkin"] = {"K": Ko, "n": n, "ms": ms, "solves_per_s": world * Ko / (ms * 1e-3),
                               "GFLOPs": Ko * flop_per / (ms * 1e-3) / 1e9}
    yh = yo.cpu().pin_memory().numpy(); Ah = Ahat.cpu().numpy(); bh = bhat.cpu().numpy()
    Ch = torch.empty((Ko, n), dtype=torch.float64, pin_memory=True).numpy()
    eng.reduced_galerkin_host(yh, Ah, bh, out=Ch)              # warm-up (staging buffers)
    barrier()                                                  # all ranks use the host side at the same time
    t0 = time.perf_counter()
    try:
        eng.reduced_galerkin_host(yh, Ah, bh, out=Ch)
        dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        out["reduced_galerkin"]["e2e_solves_per_s"] = world * Ko / float(dt.item())
    except np.linalg.LinAlgError:
        out["reduced_galerkin"]["e2e_solves_per_s"] = None
    return out


if __name__ == "__main__":
    main()
