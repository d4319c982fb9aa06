#!/usr/bin/env python
"""bench.py -- FindNeighbors cells/sec (kNN + nn Graph + SNN) on synthetic PCA-like data.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config C|B] [--impl b200|reference]

A "step" is one pass of the hot path over the whole embedding:
  value : inputs (FP64 embedding, row-major) already resident in HBM, results left in
          HBM (kNN index/distances, nn Graph CSC, SNN CSR); timed with CUDA events.
  e2e   : the same through the host-buffer C ABI (b200nn_find_neighbors_count/_fill)
          with pinned HOST buffers: H2D of the embedding and D2H of idx/dist/nn/snn
          are inside the timed region.
N > 1 (torchrun): query rows sharded over ranks, embedding replicated, one all-gather
of the kNN index, each rank emits its SNN rows.  Total work is fixed -> "strong".

--impl reference times the CPU restatement of the reference (oracle/: ANN-style
kd-tree + Gustavson ComputeSNN, all host threads) on a bounded sample of the same
workload; R, RANN and Eigen are not available so the real reference cannot run.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from seurat_b200.synthetic import CONFIGS, gaussian_mixture  # noqa: E402

METRIC = "FindNeighbors cells/sec (kNN+SNN)"
PRUNE = 1.0 / 15.0


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {"hbm_gbs": p["hbm_gbs"], "bf16_tflops": p["bf16_tflops"],
                "bf16_tflops_sustained": p.get("bf16_tflops_sustained", p["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    """samples nvidia-smi clocks / throttle reasons while the timed region runs"""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.gpu), "-lms", "200"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        # the idle samples at the edges drag the median down; keep the upper half (under load)
        sm_sorted = sorted(sm)
        load = sm_sorted[len(sm_sorted) // 2:] if sm_sorted else []
        return {"sm_mhz": statistics.median(load) if load else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------- CPU arm
def cpu_baseline(X: np.ndarray, k: int, idx1_full: np.ndarray | None, knn_sample: int, snn_rows: int,
                 threads: int) -> dict:
    """ANN-style kd-tree kNN + Gustavson SNN (oracle/) on a bounded sample, extrapolated
    linearly in the number of query rows / SNN rows (tree build and transpose are timed in full)."""
    import oracle
    n = X.shape[0]
    rng = np.random.default_rng(2024)
    qs = np.sort(rng.choice(n, size=min(knn_sample, n), replace=False))
    t = {}
    idx_s, _ = oracle.knn(X, X[qs], k, method="kdtree", nthreads=threads, timings=t)
    knn_total = t["build_s"] + t["search_s"] * (n / len(qs))
    if idx1_full is None:
        # no GPU index available (reference arm): build a plausible full index cheaply is not
        # possible on the CPU, so SNN is timed on the kd-tree result of a contiguous block instead
        blk = min(n, max(snn_rows, len(qs)))
        sub = X[:blk]
        idx_blk, _ = oracle.knn(sub, sub, k, method="kdtree", nthreads=threads)
        t2 = {}
        oracle.compute_snn_rows(idx_blk, PRUNE, 0, blk, nthreads=threads, timings=t2)
        snn_total = (t2["transpose_s"] + t2["rows_s"]) * (n / blk)
        snn_note = f"SNN on a {blk}-cell block's own kNN graph, scaled by n/{blk}"
    else:
        r0 = 0
        r1 = min(n, snn_rows)
        t2 = {}
        oracle.compute_snn_rows(idx1_full, PRUNE, r0, r1, nthreads=threads, timings=t2)
        snn_total = t2["transpose_s"] + t2["rows_s"] * (n / (r1 - r0))
        snn_note = f"SNN rows 0..{r1} of the full index (transpose in full), scaled by n/{r1 - r0}"
    total = knn_total + snn_total
    return {"value": n / total, "unit": "cells/s", "cores": threads, "kind": "port",
            "sample": (f"kd-tree build on all {n} cells ({t['build_s']:.1f}s) + search of {len(qs)} sampled queries "
                       f"({t['search_s']:.1f}s) scaled by n/{len(qs)}; {snn_note}; oracle/ restatement, not R"),
            "knn_s_extrapolated": knn_total, "snn_s_extrapolated": snn_total}


def run_reference_arm(args, cfg):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle
    n, d, k = cfg["n"], cfg["d"], cfg["k"]
    X, _ = gaussian_mixture(n, d, cfg["seed"])
    threads = oracle.num_threads()
    vals = []
    for s in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        r = cpu_baseline(X, k, None, args.cpu_knn_sample, args.cpu_snn_rows, threads)
        if s >= args.warmup:
            vals.append((r, time.perf_counter() - t0))
    r = vals[-1][0]
    value = statistics.median(v[0]["value"] for v in vals)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "cells/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * n / value,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": workload_config(cfg, args.gpus),
            "cpu_baseline": {"value": value, "unit": "cells/s", "cores": threads, "kind": "port", "sample": r["sample"]},
            "e2e": {"value": value, "unit": "cells/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
            "note": "extrapolated from a bounded sample; each step took %.1fs of CPU time" % statistics.median(v[1] for v in vals)}
    print(json.dumps(line))


def traffic_from_profiles(kernel: str):
    """dram bytes per launch of `kernel` from the committed ncu capture (profiles/traffic.json), else None"""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f).get(kernel)
    return None


def workload_config(cfg, gpus):
    return {"workload": f"synthetic Gaussian mixture {cfg['n']} cells x {cfg['d']} PCs, k={cfg['k']}, prune.SNN=1/15 "
                        f"(SURVEY 8d seed {cfg['seed']})",
            "n_cells": cfg["n"], "dims": cfg["d"], "k": cfg["k"], "prune": "1/15",
            "parallelism": f"query-rows sharded x{gpus}, embedding replicated, kNN index all-gathered",
            "l2_flush": "inputs larger than L2 (embedding %.0f MB fp64)" % (cfg["n"] * cfg["d"] * 8 / 1e6)}


# --------------------------------------------------------------------------- GPU arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="C", choices=sorted(CONFIGS))
    ap.add_argument("--n", type=int, default=0, help="override the number of cells (debugging)")
    ap.add_argument("--knn-algo", type=int, default=0, help="0 auto, 1 exact FP64 scan, 2 tensor-core")
    ap.add_argument("--scan-mode", type=int, default=0,
                    help="tensor path: 0 auto (grouped, tile-pruned), 1 exhaustive scan of every tile, 2 grouped")
    ap.add_argument("--n-cand", type=int, default=0, help="candidates kept per query by the tensor path (0 = auto)")
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--cpu-knn-sample", type=int, default=2000)
    ap.add_argument("--cpu-snn-rows", type=int, default=100_000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    cfg = dict(CONFIGS[args.config])
    if args.n:
        cfg["n"] = args.n
    if args.impl == "reference":
        run_reference_arm(args, cfg)
        return

    import torch
    import torch.distributed as dist
    import seurat_b200 as sb
    from seurat_b200 import _lib
    from seurat_b200 import dist as sdist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("--gpus N > 1 must be launched with torchrun (one rank per GPU)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    n, d, k = cfg["n"], cfg["d"], cfg["k"]
    X, _ = gaussian_mixture(n, d, cfg["seed"])           # identical on every rank (seeded)
    be = sdist.CudaBackend(local, knn_algo=args.knn_algo, n_cand=args.n_cand, scan_mode=args.scan_mode)
    ctx = be.ctx
    ref = torch.from_numpy(X).to(dev)
    torch.cuda.synchronize()
    b, e = sdist.shard_bounds(n, world)[rank]
    peaks = load_peaks()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    stage_ms = {"knn_cand": 0.0, "knn_rerank": 0.0, "knn_fallback": 0.0, "nn_graph": 0.0, "snn": 0.0}
    launches = 0
    last = {}

    def step(record: bool):
        nonlocal launches, last
        idx0, dd = be.knn_shard(ref, b, e, k)
        s1 = dict(be.last_stats)
        idx_full = sdist.all_gather_index(idx0, n) if world > 1 else idx0
        if world > 1:
            # the library runs on its own stream: the all-gather (NCCL stream) must have landed
            torch.cuda.current_stream().synchronize()
        g = be.ctx.dev_graphs(idx_full, n, k, PRUNE, True, b, e)
        if record:
            for key in ("knn_cand", "knn_rerank", "knn_fallback"):
                stage_ms[key] += s1.get("ms_" + key, 0.0)
            stage_ms["nn_graph"] += g["ms_nn_graph"]
            stage_ms["snn"] += g["ms_snn"]
            launches += int(s1.get("gpu_launches", 0)) + int(g.get("gpu_launches", 0))
        last = {"knn": s1, "graphs": g, "idx0": idx0}

    for _ in range(args.warmup):
        step(False)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step(True)
    ev1.record()
    barrier()
    wall_ms = (time.perf_counter() - t0) * 1e3
    clocks = sampler.stop() if rank == 0 else None
    ev_ms = ev0.elapsed_time(ev1)
    # the library works on its own stream and synchronises before returning, so the host
    # wall clock between the two barriers is the device-side duration; take the larger of both
    my_ms = max(ev_ms, wall_ms)
    t = torch.tensor([my_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    ms_per_step = total_ms / args.steps
    value = n / (ms_per_step / 1e3)

    # ---- end-to-end through the host-buffer C ABI (pinned host memory) --------------------
    e2e = None
    if world == 1:
        Xp = torch.from_numpy(X).pin_memory()
        Xh = Xp.numpy()
        nnz_nn = last["graphs"]["nnz_nn"]
        nnz_snn = last["graphs"]["nnz_snn"]
        slack = int(nnz_snn * 1.05) + 1024
        pin = lambda count, dt: torch.empty(count, dtype=dt).pin_memory().numpy()  # noqa: E731
        out = {"idx": pin(n * k, torch.int32).reshape(n, k), "dist": pin(n * k, torch.float64).reshape(n, k),
               "nn_p": pin(n + 1, torch.int32), "nn_i": pin(nnz_nn, torch.int32), "nn_x": pin(nnz_nn, torch.float64),
               "snn_p": pin(n + 1, torch.int32), "snn_i": pin(slack, torch.int32), "snn_x": pin(slack, torch.float64)}
        times = []
        for s in range(1 + max(1, args.e2e_steps)):
            t0 = time.perf_counter()
            r = ctx.find_neighbors(Xh, k, PRUNE, True, _lib.ROW_MAJOR, args.knn_algo, args.n_cand, True, out,
                                   scan_mode=args.scan_mode)
            t1 = time.perf_counter()
            if s > 0:
                times.append(t1 - t0)
        h2d = n * d * 8
        d2h = n * k * 12 + 2 * 4 * (n + 1) + 12 * (nnz_nn + r["stats"]["nnz_snn"])
        e2e = {"value": n / statistics.median(times), "unit": "cells/s", "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": d2h, "ms_per_step": 1e3 * statistics.median(times),
               "stage_ms": {kk: vv for kk, vv in r["stats"].items() if kk.startswith("ms_")}}
    else:
        # every rank: H2D of the replicated embedding + its query shard's results back
        Xp = torch.from_numpy(X).pin_memory()
        times = []
        pinned = {}
        refe = torch.empty((n, d), dtype=torch.float64, device=dev)
        for s in range(1 + max(1, args.e2e_steps)):
            barrier()
            t0 = time.perf_counter()
            # rank 0 uploads the embedding once, the other ranks receive it over NVLink
            if rank == 0:
                refe.copy_(Xp, non_blocking=True)
            dist.broadcast(refe, 0)
            torch.cuda.synchronize()
            idx0, dd = be.knn_shard(refe, b, e, k)
            idx_full = sdist.all_gather_index(idx0, n)
            g = be.graph_rows(idx_full, k, PRUNE, True, b, e)
            # results of this rank's shard go to page-locked host buffers (allocated once)
            for name, src in (("idx", idx0), ("dist", dd), ("snn_p", g["snn"][0]), ("snn_i", g["snn"][1]),
                              ("snn_x", g["snn"][2])):
                flat = src.reshape(-1)
                if name not in pinned or pinned[name].numel() < flat.numel():
                    pinned[name] = torch.empty(int(flat.numel() * 1.05) + 1024, dtype=flat.dtype).pin_memory()
                pinned[name][: flat.numel()].copy_(flat, non_blocking=True)
            torch.cuda.synchronize()
            t1 = time.perf_counter()
            tt = torch.tensor([t1 - t0], dtype=torch.float64, device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            if s > 0:
                times.append(float(tt.item()))
        nnz_local = int(g["snn"][1].numel())
        e2e = {"value": n / statistics.median(times), "unit": "cells/s", "h2d_bytes_per_step": n * d * 8,
               "d2h_bytes_per_step": (e - b) * k * 12 + 8 * (e - b + 1) + 12 * nnz_local,
               "ms_per_step": 1e3 * statistics.median(times),
               "note": "H2D once on rank 0 + NCCL broadcast of the embedding; D2H bytes are per rank"}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel + the SNN stage --------------------------------
    st = {kk: vv / args.steps for kk, vv in stage_ms.items()}
    nq_local = e - b
    knn_ms = st["knn_cand"] if st["knn_cand"] > 0 else st["knn_fallback"]
    flops = 2.0 * nq_local * n * d
    tensor_peak = peaks["bf16_tflops_sustained"]
    ks = last["knn"]
    tv, tt = ks.get("knn_tiles_visited", 0), ks.get("knn_tiles_total", 0)
    roofline = {"kernel": "k_knn_tc (tcgen05 candidates, both passes + list building)" if st["knn_cand"] > 0
                else "k_knn_exact (FP64 scan, no tensor cores)",
                "bound": "tensor", "achieved": flops / (knn_ms * 1e-3) / 1e12 if knn_ms > 0 else None,
                "peak": tensor_peak, "unit": "TFLOP/s", "frac": None, "traffic": traffic_from_profiles("k_knn_tc"),
                "peak_source": peaks["source"] + " bf16 sustained", "algorithmic_flops_per_launch": flops,
                "ms_per_launch": knn_ms}
    if roofline["achieved"]:
        roofline["frac"] = roofline["achieved"] / tensor_peak
    if tt:
        # the grouped search skips (item, tile) pairs that provably hold no neighbour, so the
        # ALGORITHMIC rate can exceed the tensor peak; the executed rate cannot
        exec_flops = 2.0 * tv * 256 * 128 * 64
        roofline.update({
            "tiles_visited": tv, "tiles_total": tt, "tiles_visited_frac": tv / tt,
            "executed_mma_tflops": exec_flops / (knn_ms * 1e-3) / 1e12 if knn_ms > 0 else None,
            "candidates_per_s_executed": tv * 256 * 128 / (knn_ms * 1e-3) if knn_ms > 0 else None,
            "note": ("frac uses the algorithmic 2*nq*nr*d FLOPs of SURVEY 8d; the grouped search visits only "
                     "tiles_visited_frac of the key tiles (bound-based skipping, result certified exact). "
                     "The exhaustive scan (--scan-mode 1) is limited by the epilogue (TMEM reads + selection) at 19-20 keys/clk/SM, "
                     "see DESIGN.md section 4")})
    g = last["graphs"]
    rows_local = e - b
    snn_bytes = 4.0 * n * k + (4.0 * (n + 1) + 12.0 * g["nnz_nn"]) + (4.0 * (rows_local + 1) + 12.0 * g["nnz_snn"])
    snn_ms = st["nn_graph"] + st["snn"]
    roofline_snn = {"kernel": "K3 transpose + K4 SNN rows (stage total)", "bound": "hbm",
                    "achieved": snn_bytes / (snn_ms * 1e-3) / 1e9 if snn_ms > 0 else None, "peak": peaks["hbm_gbs"],
                    "unit": "GB/s", "algorithmic_bytes": snn_bytes, "ms": snn_ms,
                    "expanded_bytes": 4.0 * g.get("snn_expanded", 0),
                    "traffic": traffic_from_profiles("k_snn_rows_filter")}
    if roofline_snn["achieved"]:
        roofline_snn["frac"] = roofline_snn["achieved"] / peaks["hbm_gbs"]

    cpu = None
    if not args.no_cpu_baseline and world == 1:
        import oracle
        idx1 = (last["idx0"].cpu().numpy() + 1).astype(np.int32)
        cpu = cpu_baseline(X, k, idx1, args.cpu_knn_sample, args.cpu_snn_rows, oracle.num_threads())

    line = {"metric": METRIC, "value": value, "unit": "cells/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f16 tensor candidates + f64 re-rank" if st["knn_cand"] > 0 else "f64",
            "data": "synthetic", "config": workload_config(cfg, world), "clocks": clocks, "e2e": e2e,
            "gpu_launches": launches, "roofline": roofline, "roofline_snn": roofline_snn, "cpu_baseline": cpu,
            "stage_ms_per_step": st, "n_uncertified": last["knn"].get("n_uncertified"),
            "nnz_snn": g["nnz_snn"], "max_indegree": g.get("max_indegree")}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
