#!/usr/bin/env python
"""bench.py -- AMM-PGO* edge-updates/s on the synthetic 1M-pose SE(3) grid (BASELINE.json
configs[3]) at 1/2/4/8 B200, with the kernel roofline and the CPU baseline beside it.

    python bench.py --gpus N --steps K --warmup W            # this repo (CUDA path)
    python bench.py --impl reference --gpus N ...            # restated CPU reference (oracle)

A "step" is one full AMM-PGO* iteration (iterate + communicate + update,
C++/examples/dist_pgo.cpp:497-521) over the whole graph.  `value` = E * K / seconds with
the graph and the iterate resident in HBM; `e2e` runs the same iteration through the
reference-facing call sequence with HOST matrices (initialize(X_host) -> update -> iterate
-> communicate -> X() back to host) inside the timed region.
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

METRIC = "AMM-PGO* edge-updates/s on 1M-pose SE(3) graph"
UNIT = "edge-updates/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--grid", default="100,100,100", help="nx,ny,nz of the synthetic SE(3) grid")
    ap.add_argument("--workload", default="grid", choices=["grid", "sphere"],
                    help="grid: BASELINE.json configs[3] (the metric's workload); sphere: configs[4]-shaped multi-robot "
                         "sphere, --nodes robots of --poses-per-robot poses (secondary line, not the headline)")
    ap.add_argument("--poses-per-robot", type=int, default=39063)
    ap.add_argument("--nodes", type=int, default=64)
    ap.add_argument("--loss", default="trivial")
    ap.add_argument("--algorithm", default="star", choices=["star", "hash"])
    ap.add_argument("--preconditioner", default="BlockJacobi", choices=["BlockJacobi", "RegularizedCholesky", "Jacobi", "None"],
                    help="tCG preconditioner of both arms (the headline is quoted with BlockJacobi; RegularizedCholesky is the "
                         "reference's default)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=10)
    ap.add_argument("--time-to-cost-cpu", type=int, default=0, metavar="ITERS",
                    help="time-to-cost against the CPU baseline's cost (SURVEY.md section 8d) on the bounded sample the "
                         "baseline can finish: the restated CPU reference runs ITERS iterations of the slab sample, the "
                         "target is (1 + 1e-3) x its final 2F, and both sides are timed from the same initial iterate")
    ap.add_argument("--time-to-cost", type=int, default=0, metavar="ITERS",
                    help="also report BASELINE.json's second metric: seconds until 2F <= (1 + 1e-3) x the cost after "
                         "ITERS iterations (1000 in SURVEY.md section 8d) of this implementation")
    return ap.parse_args()


def workload_name(args, N, E):
    if args.workload == "sphere":
        return ("synthetic multi-robot SE(3) sphere (BASELINE.json configs[4] shape), %d poses / %d edges, %d robot "
                "nodes of %d poses, %s loss, %s" % (N, E, args.nodes, args.poses_per_robot, args.loss,
                                                    "AMM-PGO*" if args.algorithm == "star" else "AMM-PGO#"))
    return ("synthetic %s SE(3) grid, %d poses / %d edges, %d robot nodes, %s loss, %s"
            % (args.grid.replace(",", "x"), N, E, args.nodes, args.loss,
               "AMM-PGO*" if args.algorithm == "star" else "AMM-PGO#"))


class ClockSampler:
    """SM clock and throttle reasons DURING the timed region (B200_PROFILING.md, the nvidia-smi clocks
    line).  Sampled in-process through NVML (the library nvidia-smi itself queries): spawning nvidia-smi
    from a process with torch loaded holds the GIL for tens of milliseconds per fork, which showed up as
    a 30 % slower 0.2 s timed region in one run.  Falls back to the nvidia-smi subprocess without pynvml."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index, uuid=None):
        self.index = index
        self.samples = []          # [sm_mhz, sm_max_mhz, hw_slowdown, hw_thermal, sw_thermal, sw_power_cap]
        self.source = "nvidia-smi"
        self._nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            h = None
            if uuid:
                for cand in (uuid, "GPU-" + uuid):
                    try:
                        h = pynvml.nvmlDeviceGetHandleByUUID(cand.encode() if isinstance(cand, str) else cand)
                        break
                    except Exception:
                        h = None
            if h is None:
                h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self._mx = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            self._bits = [pynvml.nvmlClocksEventReasonHwSlowdown, pynvml.nvmlClocksEventReasonHwThermalSlowdown,
                          pynvml.nvmlClocksEventReasonSwThermalSlowdown, pynvml.nvmlClocksEventReasonSwPowerCap]
            pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)       # probe once outside the timed region
            pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)
            self._nvml, self._h, self.source = pynvml, h, "nvml"
        except Exception:
            self._nvml = None
        self._stop = threading.Event()
        self._t = threading.Thread(target=self._run, daemon=True)

    def _sample_nvml(self):
        n, h = self._nvml, self._h
        sm = float(n.nvmlDeviceGetClockInfo(h, n.NVML_CLOCK_SM))
        r = int(n.nvmlDeviceGetCurrentClocksEventReasons(h))
        self.samples.append([sm, self._mx] + [bool(r & b) for b in self._bits])

    def _sample_smi(self):
        out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                              "--format=csv,noheader,nounits"], capture_output=True, text=True,
                             timeout=5).stdout.strip().splitlines()
        if out:
            f = [x.strip() for x in out[0].split(",")]
            self.samples.append([float(f[0]), float(f[1])] + [v.lower().startswith("active") for v in f[2:6]])

    def _run(self):
        while not self._stop.is_set():
            try:
                if self._nvml is not None:
                    self._sample_nvml()
                else:
                    self._sample_smi()
            except Exception:
                pass
            self._stop.wait(0.02 if self._nvml is not None else 0.2)

    def __enter__(self):
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=6)

    def wait_first(self, timeout=5.0):
        """Returns once the sampling thread has delivered a sample (or after `timeout` seconds)."""
        t0 = time.perf_counter()
        while not self.samples and time.perf_counter() - t0 < timeout:
            time.sleep(0.005)

    def mark(self):
        """Start of the timed region: samples taken before it do not count."""
        self._skip = len(self.samples)

    def summary(self):
        samples = self.samples[getattr(self, "_skip", 0):] or self.samples[-1:]
        if not samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = [s[0] for s in samples]
        mx = [s[1] for s in samples]
        reasons = sorted({n for s in samples for n, v in zip(self.NAMES, s[2:6]) if v})
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": max(mx), "reasons": reasons,
                "samples": len(samples), "source": self.source}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def build_hash():
    import glob
    import hashlib
    h = hashlib.sha256()
    for f in sorted(glob.glob(os.path.join(ROOT, "dpgo_b200", "csrc", "*"))):
        if not f.endswith(".o"):
            h.update(open(f, "rb").read())
    return h.hexdigest()[:16]


def ncu_traffic(kernel_key):
    """DRAM bytes per launch of the dominant kernel from the committed ncu capture -- only if that capture was
    taken on THIS build of the kernels (profiles/ncu_traffic.json records the source hash), else None."""
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(p):
        t = json.load(open(p))
        if t.get("build") == build_hash():
            return t.get(kernel_key)
    return None


# ---------------------------------------------------------------------------
def make_graph(args, dims=None):
    import dpgo_b200 as D
    if args.workload == "sphere":
        return D.sphere_rings(args.nodes, args.poses_per_robot)
    nx, ny, nz = dims or tuple(int(v) for v in args.grid.split(","))
    return D.grid3d(nx, ny, nz)


def cpu_reference_driver(args, g, nodes, workers, threads):
    """The restated CPU reference on graph g: the C++/OpenMP restatement (oracle/cpu_dpgo.cpp, built into
    oracle/_ref/libcpu_dpgo.so together with the reference's own AVX2 projection kernels) when the prebuilt
    library travelled with the repo, else None."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from oracle import cpu_ref
    from oracle import dpgo as odpgo
    from parity import to_measurements
    if not cpu_ref.available() or args.preconditioner not in cpu_ref.PRECON:
        return None                          # (the C++ restatement carries None / Jacobi / BlockJacobi)
    opts = odpgo.Options(loss=args.loss, preconditioner=args.preconditioner)
    return cpu_ref.CpuDPGO(to_measurements(g), g.num_poses, nodes, opts, args.algorithm, workers=workers,
                           threads=threads, mode=0)


def cpu_time_steps(drv, X0, steps, warmup, budget_s=None):
    """initialize + update, `warmup` untimed and up to `steps` timed iterations (iterate, communicate, update:
    what dist_pgo times, C++/examples/dist_pgo.cpp:496-521, plus communicate).  Stops early when the budget
    is spent; returns (seconds, timed steps actually run, per-step seconds)."""
    drv.initialize(X0)
    drv.update()
    per = []
    t_begin = time.perf_counter()
    for k in range(warmup + steps):
        t0 = time.perf_counter()
        drv.iterate()
        drv.communicate()
        drv.update()
        dt = time.perf_counter() - t0
        if k >= warmup:
            per.append(dt)
        if budget_s is not None and k + 1 >= warmup + 1 and time.perf_counter() - t_begin + dt > budget_s:
            break
    return float(sum(per)), len(per), per


def cpu_baseline_sample(args, sample_grid=(100, 125, 5), sample_nodes=4, iters=5):
    """`cpu_baseline` object of the CUDA arm: the restated CPU reference on a BOUNDED sample of the same
    workload (a slab of the same grid generator with the per-node size of the full workload), all host
    threads, nodes looped serially with OpenMP inside the operators like the reference; setup excluded
    as in dist_pgo.  Runs in the process that holds the CUDA context, so nothing is forked."""
    g, _, X0 = make_graph(args, sample_grid)
    threads = os.cpu_count() or 1
    t0 = time.perf_counter()
    drv = cpu_reference_driver(args, g, sample_nodes, workers=1, threads=threads)
    if drv is None:
        return numpy_oracle_sample(args, sample_grid, sample_nodes, 2)
    setup = time.perf_counter() - t0
    secs, n, _ = cpu_time_steps(drv, X0, iters, 1, budget_s=60.0)
    drv.close()
    return {"value": g.num_edges * n / secs, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": "C++/OpenMP restatement (oracle/cpu_dpgo.cpp + the reference's AVX2 projections) on a %dx%dx%d "
                      "SE(3) grid slab: %d poses / %d edges, %d robot nodes of %d poses (the per-node size of the full "
                      "workload), %d timed iterations after 1 warm-up, %.2f s in iterate+communicate+update (setup "
                      "%.1f s excluded), %d OpenMP threads, nodes looped serially" % (
                          sample_grid + (g.num_poses, g.num_edges, sample_nodes, g.num_poses // sample_nodes, n, secs,
                                         setup, threads))}


def time_to_cost_vs_cpu(args, iters, device, sample_grid=(100, 125, 5), sample_nodes=4):
    """BASELINE.json's second metric with the CPU baseline's cost as the target: the restated CPU reference
    (oracle/cpu_dpgo.cpp) runs `iters` iterations of the slab sample (4 robot nodes of the full workload's node
    size); target = (1 + 1e-3) x its final 2F.  Reported: the seconds each side needs from the same initial
    iterate until its own 2F (sum of the nodes' objectives, read every iteration on both sides) is at or below
    that target -- iterate + communicate + update only, setup excluded on both sides."""
    import dpgo_b200 as D
    g, _, X0 = make_graph(args, sample_grid)
    threads = os.cpu_count() or 1
    cdrv = cpu_reference_driver(args, g, sample_nodes, workers=1, threads=threads)
    if cdrv is None:
        return None
    cdrv.initialize(X0)
    cdrv.update()
    ctrace, ctime, t_acc = [2.0 * float(cdrv.node_scalars()[:, 0].sum())], [0.0], 0.0
    for _ in range(iters):
        t0 = time.perf_counter()
        cdrv.iterate(); cdrv.communicate(); cdrv.update()
        t_acc += time.perf_counter() - t0
        ctrace.append(2.0 * float(cdrv.node_scalars()[:, 0].sum()))
        ctime.append(t_acc)
    cdrv.close()
    target = (1.0 + 1e-3) * ctrace[-1]
    k_cpu = next(k for k, f in enumerate(ctrace) if f <= target)
    cls = D.DPGOStar if args.algorithm == "star" else D.DPGOHash
    drv = cls(g, sample_nodes, D.Options(loss=args.loss, device=device, preconditioner=args.preconditioner))
    assert drv.initialize(X0) == 0
    D.lib.check(drv.update())
    for _ in range(3):                                    # warm the kernels, then start again
        D.lib.check(drv.iterate()); D.lib.check(drv.communicate()); D.lib.check(drv.update())
    assert drv.initialize(X0) == 0
    D.lib.check(drv.update())
    drv.synchronize()
    gtrace = [2.0 * drv.objective()[0]]
    t0, k = time.perf_counter(), 0
    while k < 4 * iters and gtrace[-1] > target:
        D.lib.check(drv.iterate()); D.lib.check(drv.communicate()); D.lib.check(drv.update())
        gtrace.append(2.0 * drv.objective()[0])
        k += 1
    drv.synchronize()
    g_secs = time.perf_counter() - t0
    n = min(len(gtrace), len(ctrace))
    dev = max(abs(a - b) / abs(b) for a, b in zip(gtrace[:n], ctrace[:n]))
    return {"target_2F": target, "reference_iterations": iters, "reached": gtrace[-1] <= target,
            "gpu": {"seconds": g_secs, "iterations": k, "final_2F": gtrace[-1]},
            "cpu": {"seconds": ctime[k_cpu], "iterations": k_cpu, "final_2F": ctrace[-1], "cores": threads,
                    "seconds_all_iterations": ctime[-1]},
            "trace_max_rel_dev": dev,
            "sample": "%dx%dx%d SE(3) grid slab, %d poses / %d edges, %d robot nodes of %d poses, %s loss, %s"
                      % (sample_grid + (g.num_poses, g.num_edges, sample_nodes, g.num_poses // sample_nodes, args.loss,
                                        "AMM-PGO*" if args.algorithm == "star" else "AMM-PGO#")),
            "note": "target = (1 + 1e-3) x the restated CPU reference's 2F after reference_iterations iterations; both "
                    "sides start from the same initial iterate and read 2F (sum of the nodes' objectives) after every "
                    "iteration; setup excluded on both sides; trace_max_rel_dev = largest relative difference of the "
                    "two objective traces over the common iterations"}


def numpy_oracle_sample(args, sample_grid, sample_nodes, iters):
    """Fallback when oracle/_ref/libcpu_dpgo.so is absent: the numpy/scipy oracle, one thread."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from oracle import dist_pgo as odist
    from oracle import dpgo as odpgo
    from parity import to_measurements
    g, _, X0 = make_graph(args, sample_grid)
    timing = {}
    odist.run(to_measurements(g), g.num_poses, sample_nodes, odpgo.Options(loss=args.loss, preconditioner=args.preconditioner),
              X0, iters, args.algorithm, log_global=False, timing=timing)
    return {"value": g.num_edges * iters / timing["seconds"], "unit": UNIT, "cores": 1, "kind": "port",
            "sample": "numpy/scipy oracle (libcpu_dpgo.so not built) on a %dx%dx%d slab, %d robot nodes, %d iterations"
                      % (sample_grid + (sample_nodes, iters))}


def run_reference(args):
    """Reference arm: the restated CPU reference on the FULL workload of the CUDA arm (same generator, seed,
    node count, loss, algorithm, initial iterate), all host threads, `--warmup` untimed and `--steps` timed
    iterations (fewer only if a 4-minute budget runs out; the line carries the count actually timed)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    g, _, X0 = make_graph(args)
    N, E = g.num_poses, g.num_edges
    threads = os.cpu_count() or 1
    t0 = time.perf_counter()
    drv = cpu_reference_driver(args, g, args.nodes, workers=min(threads, 16), threads=threads)
    if drv is None:
        cb = numpy_oracle_sample(args, (100, 125, 5), 4, 2)
        line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": 2,
                "warmup": 0, "ms_per_step": 1e3 * 250000 / cb["value"], "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": workload_name(args, N, E), "note": "bounded sample only: " + cb["sample"]},
                "cpu_baseline": cb, "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        emit(line)
        return
    setup = time.perf_counter() - t0
    secs, n, per = cpu_time_steps(drv, X0, args.steps, args.warmup, budget_s=240.0)
    value = E * n / secs
    fobj = float(drv.node_scalars()[:, 0].sum())
    cb = {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
          "sample": "the full workload: %d timed iterations after %d warm-up, %.1f s in iterate+communicate+update "
                    "(setup %.1f s excluded: matrix assembly by oracle/data_matrix.py, sparse Cholesky of G00), %d OpenMP "
                    "threads inside the operators, robot nodes looped serially (C++/examples/dist_pgo.cpp:497-520)"
                    % (n, args.warmup, secs, setup, threads)}
    line = {
        "impl": "reference", "metric": METRIC if args.workload == "grid" else "AMM-PGO%s edge-updates/s on a multi-robot SE(3) sphere" % (
            "*" if args.algorithm == "star" else "#"),
        "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": n, "warmup": args.warmup, "ms_per_step": 1e3 * secs / n, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(args, N, E), "preconditioner": args.preconditioner,
                   "final_2F": 2 * fobj, "step_seconds": per,
                   "note": "upstream dist_pgo cannot be built here (no Eigen / SuiteSparse / glog / Boost); this arm times the "
                           "restated reference algorithm in C++17 + OpenMP (oracle/cpu_dpgo.cpp: the reference's scalar CSR "
                           "operators, an up-looking sparse Cholesky in place of CHOLMOD, the reference's own AVX2 SO(3) "
                           "projection compiled from its sources) on the same graph, partition, initial iterate and options "
                           "as the CUDA arm; `steps` is the number of iterations actually timed"},
        "cpu_baseline": cb,
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ---------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    import dpgo_b200 as D
    from dpgo_b200 import multi

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit("launch with torchrun --nproc-per-node %d for --gpus %d" % (args.gpus, args.gpus))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    g, _, X0 = make_graph(args)
    N, E, d = g.num_poses, g.num_edges, g.d
    opts = D.Options(loss=args.loss, device=local_rank, preconditioner=args.preconditioner)
    t_setup = time.perf_counter()
    drv = multi.make_driver(g, args.nodes, opts, args.algorithm, rank, world)
    setup_s = time.perf_counter() - t_setup     # mmpgo_set_graph: partition, majoriser blocks, G00 factor, upload (+ NCCL init)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step():
        D.lib.check(drv.iterate())
        D.lib.check(drv.communicate())
        D.lib.check(drv.update())

    assert drv.initialize(X0) == 0
    D.lib.check(drv.update())
    stream = torch.cuda.ExternalStream(drv.stream())
    # per-iteration trace from iteration 0: an event behind every step (no synchronisation of its own) and the
    # rank-local objective (host scalars of update(), no device traffic)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.warmup + args.steps + 1)]
    f_trace = [drv.objective()[0]]
    ev[0].record(stream)
    for k in range(args.warmup):
        step()
        ev[k + 1].record(stream)
        f_trace.append(drv.objective()[0])
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    drv.reset_counters()
    try:
        dev_uuid = str(torch.cuda.get_device_properties(local_rank).uuid)
    except Exception:
        dev_uuid = None
    import gc
    gc.collect()
    gc.disable()                     # no collector pause inside the timed region
    # the sampler is created, started and has taken its first sample BEFORE the barrier: NVML start-up takes tens of
    # milliseconds when 8 processes do it at once, and skew between the ranks after the barrier would be billed to
    # the first timed step (24 ms at 8 GPUs in r02m_bench_8gpu.json's first version)
    with ClockSampler(local_rank, dev_uuid) as clk:
        clk.wait_first()
        barrier()
        clk.mark()
        e0.record(stream)
        for k in range(args.steps):
            step()
            ev[args.warmup + k + 1].record(stream)
            f_trace.append(drv.objective()[0])
        e1.record(stream)
        barrier()
    gc.enable()
    ms = e0.elapsed_time(e1)
    # (the first timed interval starts at the barrier, not at the previous step's event)
    iter_ms = [ev[k].elapsed_time(ev[k + 1]) for k in range(args.warmup)] + \
              [e0.elapsed_time(ev[args.warmup + 1])] + \
              [ev[k].elapsed_time(ev[k + 1]) for k in range(args.warmup + 1, args.warmup + args.steps)]
    f_trace = np.array(f_trace)
    if world > 1:
        t = torch.tensor(f_trace, device="cuda", dtype=torch.float64)
        dist.all_reduce(t)
        f_trace = t.cpu().numpy()
    import hashlib
    trace_digest = hashlib.sha256(" ".join("%.11e" % (2 * v) for v in f_trace).encode()).hexdigest()[:16]
    ctr = drv.counters()
    if world > 1:
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    secs = ms / 1e3
    value = E * args.steps / secs
    F, gn = drv.global_objective()

    # ---- per-kernel device times (CUDA events on the library's stream) and the roofline
    sizes = drv.sizes()
    E_intra = sizes["bsr_entries"] // 2
    NO = sizes["own_poses"]
    HE = sizes["inter_half_edges"]
    kinds = [k for k in drv.KERNEL_KINDS if k != "g00_spmv"]
    k_ms = {k: drv.profile_pass(k, 20) for k in kinds if not k.startswith("g00_solve")}
    # the persistent translation solve is timed over the solves of the timed steps themselves
    # (bytes = pose-iterations x bytes per pose-iteration); one extra cold solve gives its duration
    solve_calls = max(int(ctr.solve_calls), 1)
    pose_iters = float(ctr.reserved[0])
    drv.reset_counters()
    k_ms["g00_solve"] = drv.profile_pass("g00_solve", 5)
    c2 = drv.counters()
    solve_pose_iters = float(c2.reserved[0]) / 8.0          # 3 warm-up + 5 timed launches
    per_step = {
        "k2": ctr.intra_passes / args.steps, "k1_inter": ctr.inter_passes / args.steps,
        "k3_prox": ctr.prox_passes / args.steps, "g00_solves": ctr.solve_calls / args.steps,
        "g00_node_iters": ctr.solve_iters / args.steps,
    }
    # K2b algorithmic bytes per pose and CG iteration: ELLPACK entries (12 B
    # each, one per intra-node half-edge) + the diagonal in both phases + phase A: read z, p, Ap,
    # write p, Ap; phase B: read x, p, Ap, z, write x, z  (11 vector streams of 8 d bytes)
    sell_bytes_per_pose = 12.0 * 2 * E_intra / max(NO, 1)
    b_iter = sell_bytes_per_pose + 16 + 8 * d * 11
    sinfo = drv.solver_info()
    direct = sinfo["solver"] == "direct"
    # sparse direct solve (SURVEY.md section 8d solve model): both triangular sweeps read the factor once
    # (8 B per entry of L each; the blocks are dense, no indices) + right-hand side in, solution out
    direct_bytes = 16.0 * sinfo["factor_nnz"] + 2 * 8 * d * NO
    alg_bytes = {
        "k2_eval": 120 * E_intra + 192 * NO, "k2_grad": 120 * E_intra + 192 * NO,
        "k2_hv": 120 * E_intra + 192 * NO,
        # G01 pass (k_g01): per half-edge the column index and row 0 of the block; per pose the rotation rows
        # (read once with ideal reuse), g_t, the diagonal row and the d outputs
        "k2_g01": (4 + 8 * (d + 1)) * 2 * E_intra + (8 * d * d + 16 * d + 8 * (d + 1)) * NO,
        "k1_inter": 120 * HE + 192 * NO, "k3_prox": 680 * NO,
        "edge_objective": 120 * sizes["owned_edges"] + 96 * NO,
        "g00_solve": direct_bytes if direct else b_iter * solve_pose_iters,
    }
    # every translation solve is preceded by one G01 pass; the other K2 passes are full block-CSR passes
    n_g01 = min(per_step["g00_solves"], per_step["k2"])
    k2_full = float(np.mean([k_ms["k2_eval"], k_ms["k2_grad"], k_ms["k2_hv"]]))
    share = {
        "k2 block-CSR pass": (per_step["k2"] - n_g01) * k2_full + n_g01 * k_ms["k2_g01"],
        "k2b translation solve": per_step["g00_solves"] * k_ms["g00_solve"] *
                                 (1.0 if direct else (pose_iters / solve_calls) / max(solve_pose_iters, 1.0)),
        "k1 inter-edge pass": per_step["k1_inter"] * k_ms["k1_inter"],
        "k3 fused proximal": per_step["k3_prox"] * k_ms["k3_prox"],
    }
    peak, peak_src = measured_peak()
    # whole-step figure: algorithmic bytes of everything launched in one step / measured step time
    n_eobj = 2.0 if args.algorithm == "star" else 0.0
    step_bytes = ((per_step["k2"] - n_g01) * alg_bytes["k2_eval"] + n_g01 * alg_bytes["k2_g01"] +
                  (direct_bytes * per_step["g00_solves"] if direct else b_iter * pose_iters / args.steps) +
                  (per_step["k1_inter"] - n_eobj) * alg_bytes["k1_inter"] + n_eobj * alg_bytes["edge_objective"] +
                  per_step["k3_prox"] * alg_bytes["k3_prox"] + ctr.vector_passes / args.steps * 96.0 * 2 * NO)
    lite = int(ctr.reserved[1]) > 0
    names = {"k2b translation solve": ("g00_solve", "k_mf_solve<3> (K2b sparse Cholesky sweeps, one persistent launch per solve)"
                                       if direct else ("k_tsolve_lite<3,7> (K2b persistent Jacobi-PCG translation solve, "
                                                       "small-shard kernel)") if lite else
                                       "k_tsolve<3> + resumed tail in k_tsolve_lite<3,7> (K2b persistent Jacobi-PCG translation solve)"),
             "k2 block-CSR pass": ("k2_eval", "k_gpass<3,G_EVAL> (K2 block-CSR connection-Laplacian pass)"),
             "k1 inter-edge pass": ("k1_inter", "k_inter<3> (K1 inter-node edge pass)"),
             "k3 fused proximal": ("k3_prox", "k_prox<3> (K3 fused proximal + SO(3) projection)")}
    dom, dom_name = names[max(share, key=share.get)]
    achieved = alg_bytes[dom] / (k_ms[dom] * 1e-3) / 1e9
    roofline = {
        "bound": "hbm", "kernel": dom_name,
        "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
        "traffic": None if (lite and dom == "g00_solve") else ncu_traffic(dom), "peak_source": peak_src,
        "algorithmic_bytes_per_launch": alg_bytes[dom], "launch_ms": k_ms[dom],
        "kernel_ms": k_ms, "est_ms_per_step_by_kernel": share, "launches_per_step": per_step,
        "all_kernels_gbs": {k: alg_bytes[k] / (k_ms[k] * 1e-3) / 1e9 for k in k_ms},
        "all_kernels_frac": {k: alg_bytes[k] / (k_ms[k] * 1e-3) / 1e9 / peak for k in k_ms},
        "whole_step": {"algorithmic_bytes": step_bytes, "achieved": step_bytes / (ms / args.steps * 1e-3) / 1e9,
                       "frac": step_bytes / (ms / args.steps * 1e-3) / 1e9 / peak,
                       "note": "all kernels of one AMM-PGO* step incl. host control gaps (this rank)"},
    }

    # ---- e2e: reference-facing call sequence with host matrices inside the timed region
    e2e = multi.e2e_loop(drv, X0, args.e2e_steps, E, d, N)

    # ---- time-to-cost: the reference cost is this implementation's own cost after
    # ITERS iterations; the CPU oracle cannot run 1000 iterations of 4M edges within a bench run
    ttc = None
    if args.time_to_cost > 0:
        assert drv.initialize(X0) == 0
        D.lib.check(drv.update())
        for _ in range(args.time_to_cost):
            step()
        target = (1.0 + 1e-3) * 2.0 * drv.global_objective()[0]
        assert drv.initialize(X0) == 0
        D.lib.check(drv.update())
        barrier()
        t0, k = time.perf_counter(), 0
        while k < args.time_to_cost and 2.0 * drv.global_objective()[0] > target:
            step()
            k += 1
        drv.synchronize()
        barrier()
        ttc = {"seconds": time.perf_counter() - t0, "iterations": k, "target_2F": target,
               "reference_iterations": args.time_to_cost,
               "note": "target = (1 + 1e-3) x this implementation's own 2F after reference_iterations iterations "
                       "from the same initial iterate; objective read from the per-node sums every iteration"}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline and args.workload == "grid":
        cpu = cpu_baseline_sample(args)
    ttc_cpu = None
    if rank == 0 and world == 1 and args.time_to_cost_cpu > 0:
        ttc_cpu = time_to_cost_vs_cpu(args, args.time_to_cost_cpu, local_rank)

    if rank == 0:
        line = {
            "metric": METRIC if args.workload == "grid" else
            "AMM-PGO%s edge-updates/s on a multi-robot SE(3) sphere shard" % ("*" if args.algorithm == "star" else "#"),
            "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(args, N, E),
                       "l2": "working set (graph %.0f MB + iterates) is larger than the 126 MB L2" % (
                           (sizes["bsr_entries"] * 132 + HE * 128) / 1e6),
                       "preconditioner": args.preconditioner, "nodes_per_gpu": args.nodes // world,
                       "translation_solver": sinfo,
                       "setup_s": setup_s,        # rank 0: mmpgo_create + mmpgo_set_graph (partition, blocks, G00 factor, upload)
                       "final_2F": 2 * F, "final_2gradnorm": 2 * gn,
                       "objective_trace": {"digest": trace_digest, "first_2F": 2 * float(f_trace[0]), "last_2F": 2 * float(f_trace[-1]),
                                           "note": "sha256 over the 2F values of every iteration from 0 (warm-up included), 12 "
                                                   "significant digits: equal digests at 1/2/4/8 GPUs = the same trajectory"},
                       "iter_ms": [round(v, 4) for v in iter_ms],
                       "init": "seeded perturbation of the ground truth (sigma_t 0.2, sigma_R 0.1 rad), iteration 0 is the first "
                               "entry of iter_ms; the reference's dist_pgo starts from a chordal initialisation"},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "time_to_cost": ttc, "time_to_cost_vs_cpu": ttc_cpu,
            "gpu_launches": int(ctr.launches), "clocks": clk.summary(),
            "counters_per_step": {"launches": ctr.launches / args.steps, "k2_passes": per_step["k2"],
                                  "g00_solves": ctr.solve_calls / args.steps,
                                  "g00_node_iterations": per_step["g00_node_iters"],
                                  "tcg_iterations": ctr.tcg_iterations / args.steps},
        }
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


_RESULT_OUT = None


def emit(line):
    """The ONE JSON line of the contract, on the process's original stdout."""
    out = _RESULT_OUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    # stdout carries exactly one line (the result); whatever a library prints there on the way (NCCL's version
    # banner when a communicator is created, a warning of a forked tool) is sent to stderr from here on
    global _RESULT_OUT
    args = parse()
    sys.stdout.flush()
    _RESULT_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
