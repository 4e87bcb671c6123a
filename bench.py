#!/usr/bin/env python
"""bench.py -- headline benchmark of the neighbour-list hot path.

Metric (BASELINE.json): neighbour pairs/s for the materialised path build_cell_list +
materialize_pairlist (the timed region of the reference's scripts/benchmark.jl:148-151), random
cubic box, rho = 0.05 A^-3, rc = 5 A, full PBC, Float64 / Int32, 10 M atoms per GPU.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--atoms A]

One JSON line on stdout (rank 0).  `value` = pairs/s with positions already resident in HBM;
`e2e` = the same through the public API with HOST buffers: pinned H2D of the positions, the list, and the
whole PairList (i, j, S, first) complete in pinned host memory inside the timed region (nl_pairs_to_host:
5 B/pair cross the bus, i and S are rebuilt by host threads of the library; d2h_bytes_per_step counts what
actually crosses the bus); `roofline` = the dominant (pair-fill) kernel against the
measured HBM peak; `cpu_baseline` = the C++ restatement of the reference's CPU path (oracle/) on a
bounded sample.  `--impl reference` times that CPU restatement alone (Julia is not installed here, so
the real reference cannot be run).
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
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

DENSITY = 0.05
CUTOFF = 5.0
SEED = 10
FALLBACK_HBM_GBS = 6650.0


def make_positions(n_atoms, seed):
    """SURVEY.md 8d generator: f ~ U[0,1)^3 (PCG64(seed)), x = L f."""
    rng = np.random.Generator(np.random.PCG64(seed))
    L = (n_atoms / DENSITY) ** (1.0 / 3.0)
    X = rng.random((n_atoms, 3))
    X *= L
    return X, np.eye(3) * L, L


def hbm_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback"


def fill_traffic_profiled(n_atoms, world):
    """dram__bytes_read + dram__bytes_write of one fill launch from the COMMITTED ncu capture of this build (profiles/).  It is a
    citation of that capture, not a measurement of this run, so it is reported beside `roofline.traffic` (null), never as it."""
    if n_atoms != 10_000_000 or world != 1:
        return None
    try:
        with open(os.path.join(ROOT, "profiles", "r02_fill_traffic.json")) as f:
            return json.load(f)
    except Exception:
        return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for t, line in self.rows:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 8:
                continue
            try:
                if t0 - 0.05 <= t <= t1 + 0.05:
                    sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if v.lower().startswith("active") and t0 - 0.05 <= t <= t1 + 0.05:
                    reasons.add(name)
        if not sm:  # region shorter than the sampling period: fall back to every sample taken
            for t, line in self.rows:
                try:
                    sm.append(float(line.split(",")[1]))
                except Exception:
                    pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


def host_threads():
    """Threads the CPU arm may use: every core this process is allowed on.  torchrun exports OMP_NUM_THREADS=1 to its
    workers, which is not a statement about the machine, so the count is passed to the oracle explicitly."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def cpu_baseline_run(n_atoms, steps, warmup, cores=None):
    """The oracle (a C++ restatement of the reference's sort-based CPU path: serial binning / sort /
    offsets, count + fill parallel over atoms like the KA CPU backend) with every host thread."""
    from oracle import nl_oracle as O
    X, C, L = make_positions(n_atoms, SEED)
    cores = cores or host_threads()
    times, P = [], 0
    for s in range(warmup + steps):
        t = time.perf_counter()
        r = O.sortbased(X, CUTOFF, C, (True, True, True), nthreads=cores, want_R=False)
        dt = time.perf_counter() - t
        P = r["npairs"]
        del r
        if s >= warmup:
            times.append(dt)
    return P, times, cores


def pick_cpu_atoms(requested, steps, warmup, budget_s, cores):
    """Largest workload of the headline ladder whose (warmup + steps) runs fit `budget_s` seconds of CPU time, estimated
    from one 500 k-atom probe (the path is linear in N at fixed density)."""
    _, t, _ = cpu_baseline_run(500_000, 1, 1, cores)
    per_atom = 1.5 * t[0] / 500_000   # larger workloads run somewhat slower per atom (cache misses of the random gathers)
    for n in (10_000_000, 5_000_000, 2_000_000, 1_000_000, 500_000):
        if n <= requested and per_atom * n * (steps + warmup) <= budget_s:
            return n
    return min(requested, 200_000)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = host_threads()
    n = args.cpu_atoms if args.cpu_atoms > 0 else pick_cpu_atoms(args.atoms, args.steps, args.warmup, args.cpu_budget, cores)
    P, times, cores = cpu_baseline_run(n, args.steps, args.warmup, cores)
    ms = 1e3 * float(np.mean(times))
    val = P / (ms * 1e-3)
    same = n == args.atoms
    sample = f"{n} atoms, rho={DENSITY}, rc={CUTOFF}, pbc TTT, Float64/Int32, seed {SEED} ({P} pairs per step)"
    print(json.dumps({
        "impl": "reference", "metric": "neighbour pairs/s (build_cell_list + materialize_pairlist)", "value": val, "unit": "pairs/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": ("the headline workload: " if same else "headline density / cutoff on the largest CPU sample that fits the time budget: ") + sample,
                   "same_workload_as_gpu_arm": same, "cpu_budget_s": args.cpu_budget,
                   "note": "C++/OpenMP restatement of the reference's CPU path (oracle/); Julia is not installed so julia -t N cannot run"},
        "cpu_baseline": {"value": val, "unit": "pairs/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def run_ours(args):
    import torch
    import torch.distributed as dist

    import neighbourlists_jl_b200 as nl

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    L_ = nl._lib.lib()

    n_atoms = args.atoms
    pbc = (True, True, True)
    it = torch.int32
    sharded = comm = None
    if world > 1:
        sharded = __import__("importlib").import_module("neighbourlists_jl_b200.sharded")
        comm = sharded.make_nccl_comm()

    def make_inputs(n_per_rank, mode):
        """(X, cell, gidx): one rank's share of ONE global box of world * n_per_rank atoms (weak scaling).
        by-index: rank r holds the block [r n, (r+1) n) of the global index range, positions anywhere in the box, so the
                  all-to-all-v of nl_shard_exchange moves (1 - 1/world) of all atoms every list (the worst case);
        slabbed : rank r generates the atoms of its own equal-width z slab (a domain-decomposed MD code: almost nothing moves)."""
        if world == 1:
            X, C, L = make_positions(n_per_rank, SEED)
            return X, C, None
        n_total = n_per_rank * world
        L = (n_total / DENSITY) ** (1.0 / 3.0)
        C = np.eye(3) * L
        rng = np.random.Generator(np.random.PCG64(SEED + rank))
        X = rng.random((n_per_rank, 3))
        if mode == "slabbed":
            X[:, 2] = (X[:, 2] + rank) / world
        X *= L
        return X, C, torch.arange(rank * n_per_rank + 1, (rank + 1) * n_per_rank + 1, dtype=it)

    X, C, gidx = make_inputs(n_atoms, args.input)
    X_host = torch.from_numpy(X).pin_memory()
    X_dev = X_host.to(dev)
    if gidx is not None:
        gidx_host = gidx.pin_memory()
        gidx_dev = gidx_host.to(dev)
    torch.cuda.synchronize()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident throughput ("value") + stage / kernel timings
    def step(timers=None, Xd=None, gd=None, cell=None):
        Xd = X_dev if Xd is None else Xd
        cell = C if cell is None else cell
        if world == 1:
            clist = nl.build_cell_list(Xd, CUTOFF, cell, pbc)
            return nl.materialize_pairlist(clist, with_R=True, timers=timers)
        return sharded.neighbour_list_sharded_native(Xd, gidx_dev if gd is None else gd, CUTOFF, cell, pbc, comm, rank, world,
                                                     with_R=True, timers=timers)

    def npairs_of(pl):
        return int(pl.i.shape[0])

    for _ in range(args.warmup):
        pl = step()
        P = npairs_of(pl)
        del pl
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    timers = {}
    launches0 = L_.nl_launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t_wall0 = time.time()
    ev0.record()
    for _ in range(args.steps):
        pl = step(timers)
        P = npairs_of(pl)
        del pl
    ev1.record()
    barrier()
    t_wall1 = time.time()
    launches = L_.nl_launch_count() - launches0
    total_ms = ev0.elapsed_time(ev1)
    clocks = sampler.stop(t_wall0, t_wall1) if rank == 0 else None
    count_ms = [e[0].elapsed_time(e[1]) for e in timers["events"]]
    fill_ms = [e[2].elapsed_time(e[3]) for e in timers["events"]]

    # ---------------- end to end through the public API with HOST buffers
    # nl.to_host = nl_pairs_to_host: first, j and one byte per pair for S cross the bus, i and S are rebuilt by host threads of the
    # library while the copies run (include/nlcuda.h); the result is the complete (i, j, S, first) in pinned host memory.
    hbuf = nl.HostPairBuffers(int(P * 1.02) + 1024, int(n_atoms * 1.1) + 1024, np.int32, dev)
    # half the logical CPUs: measured on the 16-vCPU GPU box, 8 host threads beat 16 (34.5 vs 37.3 ms for the transfer): the decoders
    # are bound by memory bandwidth, which the DMA writes share
    host_threads = max(1, (os.cpu_count() or 2) // 2 // world)
    d2h_bytes = [0]

    def e2e_step():
        if world == 1:
            # pinned H2D inside; reference layout (no R); first is copied and i rebuilt by host threads while the fill pass runs
            d2h_bytes[0] = 4 * (n_atoms + 1) + 5 * P + 4
            return nl.neighbour_list(X_host, CUTOFF, C, pbc, device=dev, host_out=hbuf, host_threads=host_threads)
        else:
            pl = sharded.neighbour_list_sharded_native(X_host.to(dev, non_blocking=True), gidx_host.to(dev, non_blocking=True), CUTOFF, C, pbc,
                                                       comm, rank, world)
        h = nl.to_host(pl, out=hbuf, nthreads=host_threads)  # i rebuilt through the shard's row -> global index map; returns when complete
        d2h_bytes[0] = nl.to_host_bytes(pl)
        return h

    e2e_warm, e2e_steps = 1, max(1, min(args.steps, 3))
    for _ in range(e2e_warm):
        e2e_step()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(e2e_steps):
        h = e2e_step()
    e1.record()
    barrier()
    e2e_ms = e0.elapsed_time(e1) / e2e_steps
    assert int(h.first[-1]) - 1 == h.i.shape[0] == h.j.shape[0] == h.S.shape[0] == P
    # the host arrays of the last timed step against a plain field-by-field copy of a device list (outside the timed region)
    if world == 1:
        chk = nl.neighbour_list(X_dev, CUTOFF, C, pbc)
        lo_, hi_ = max(0, P // 2 - 2_000_000), min(P, P // 2 + 2_000_000)
        assert np.array_equal(h.first, chk.first.cpu().numpy())
        for name in ("i", "j", "S"):
            assert np.array_equal(getattr(h, name)[lo_:hi_], getattr(chk, name)[lo_:hi_].cpu().numpy()), name
            assert np.array_equal(getattr(h, name)[-100000:], getattr(chk, name)[-100000:].cpu().numpy()), name
        del chk

    # ---------------- multi-GPU extras: the other input distribution, and BASELINE config 4 (100 M atoms) at 8 ranks
    def timed_steps(Xd, gd, cell, nsteps=3, nwarm=2):
        for _ in range(nwarm):
            p_ = step(None, Xd, gd, cell); Pn = npairs_of(p_); del p_
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(nsteps):
            p_ = step(None, Xd, gd, cell); Pn = npairs_of(p_); del p_
        b.record()
        barrier()
        return a.elapsed_time(b) / nsteps, Pn

    extras = {}
    if world > 1:
        del hbuf, h
        torch.cuda.empty_cache()
        other = "slabbed" if args.input == "by-index" else "by-index"
        Xo, Co, go = make_inputs(n_atoms, other)
        ms_o, P_o = timed_steps(torch.from_numpy(Xo).to(dev), go.to(dev), Co)
        extras[other] = [ms_o, float(P_o)]
        if world == 8 and not args.no_c4:
            Xc, Cc, gc = make_inputs(12_500_000, args.input)
            ms_c, P_c = timed_steps(torch.from_numpy(Xc).to(dev), gc.to(dev), Cc)
            extras["c4"] = [ms_c, float(P_c)]
            del Xc, gc
        for k in list(extras):
            t = torch.tensor(extras[k], dtype=torch.float64, device=dev)
            tmax, tsum = t.clone(), t.clone()
            dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
            dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
            extras[k] = {"ms_per_step": float(tmax[0]), "pairs_per_s": float(tsum[1]) / (float(tmax[0]) * 1e-3), "pairs": float(tsum[1])}

    # ---------------- max over ranks
    ms_per_step = total_ms / args.steps
    stats = torch.tensor([ms_per_step, e2e_ms, float(np.mean(fill_ms)), float(np.mean(count_ms))], dtype=torch.float64, device=dev)
    pairs = torch.tensor([float(P)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.MAX)
        dist.all_reduce(pairs, op=dist.ReduceOp.SUM)
    ms_per_step, e2e_ms, fill_mean, count_mean = [float(v) for v in stats.tolist()]
    P_total = float(pairs.item())

    if rank == 0:
        peak, which = hbm_peak()
        # algorithmic bytes of ONE fill launch: pair output i 4 + j 4 + S 12 + R 24 = 44 B/pair,
        # plus one read of the per-atom records (32 B) and of `first` (4 B)
        fill_bytes = 44.0 * P + 36.0 * n_atoms
        achieved = fill_bytes / (fill_mean * 1e-3) / 1e9
        # whole-step figure the north star quotes: B = 24 N + 48 P
        step_bytes = 24.0 * n_atoms + 48.0 * P
        out = {
            "metric": "neighbour pairs/s (build_cell_list + materialize_pairlist)", "value": P_total / (ms_per_step * 1e-3),
            "unit": "pairs/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"{n_atoms} atoms per GPU, random cubic box rho={DENSITY} A^-3, rc={CUTOFF} A, pbc TTT, Float64/Int32, "
                                   f"seed {SEED}+rank; output (i,j,S,R) = 44 B/pair",
                       "pairs_per_gpu": P, "parallelism": "single GPU" if world == 1 else
                       f"{world} spatial slabs along z of one {n_atoms * world}-atom box; nl_shard_prepare / nl_shard_exchange: all-to-all-v + "
                       "cutoff-wide halo exchange with ncclSend / ncclRecv inside libnlcuda.so",
                       "input": None if world == 1 else
                       (args.input + (": every rank holds a block of the global INDEX range, positions anywhere in the box -> the all-to-all-v "
                                      "moves (1 - 1/N) of all atoms inside the timed region" if args.input == "by-index" else
                                      ": every rank holds the atoms of its own equal-width z slab (almost nothing moves)")),
                       "other_input": {k: v for k, v in extras.items() if k != "c4"} or None,
                       "c4": (dict(extras["c4"], workload="BASELINE config 4: 100 M atoms (12.5 M per GPU), same density / cutoff, "
                                   f"{args.input} input, with R") if "c4" in extras else None),
                       "l2": "inputs (240 MB) and outputs (>11 GB) exceed the 126 MB L2; no explicit flush",
                       "stage_ms": {"count_stage": count_mean, "fill_kernel": fill_mean},
                       "step_roofline": {"bytes": step_bytes, "formula": "24 N + 48 P", "gbs": step_bytes / (ms_per_step * 1e-3) / 1e9,
                                         "frac": step_bytes / (ms_per_step * 1e-3) / 1e9 / peak}},
            "roofline": {"bound": "hbm", "kernel": "pair fill (nl_fill_pairs: k_row_starts + k_expand_rows + k_fill3)", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": None, "traffic_profiled": fill_traffic_profiled(n_atoms, world), "peak_source": which,
                         "bytes_per_launch": fill_bytes, "formula": "44 P + 36 N"},
            "e2e": {"value": P_total / (e2e_ms * 1e-3), "unit": "pairs/s",
                    "h2d_bytes_per_step": int(X_host.numel() * 8 + (0 if world == 1 else n_atoms * 8)),
                    "d2h_bytes_per_step": int(d2h_bytes[0]), "ms_per_step": e2e_ms, "host_threads": host_threads,
                    "result_bytes_in_host_memory": int(20 * P + 4 * (n_atoms + 1)),
                    "note": "host positions -> neighbour_list -> nl_pairs_to_host: the whole PairList (i,j,S,first) complete in pinned host "
                            "memory; first, j and one byte per pair of S cross PCIe, i and S are rebuilt by the library's host threads "
                            "while the copies run"},
            "gpu_launches": int(launches),
            "clocks": clocks,
        }
        if world == 1 and not args.no_cpu_baseline:
            n_cpu = args.cpu_atoms if args.cpu_atoms > 0 else 1_000_000
            Pc, times, cores = cpu_baseline_run(n_cpu, 1, 1)
            out["cpu_baseline"] = {"value": Pc / float(np.mean(times)), "unit": "pairs/s", "cores": cores, "kind": "port",
                                   "sample": f"{n_cpu} atoms of the same workload ({Pc} pairs), C++/OpenMP restatement of the "
                                             "reference CPU path (not Julia)"}
        print(json.dumps(out))
    if world > 1:
        sharded.shard_disconnect(comm)
        nl._lib.check(L_.nl_nccl_comm_destroy(comm))
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--atoms", type=int, default=10_000_000)
    ap.add_argument("--cpu-atoms", type=int, default=0,
                    help="atoms of the CPU arm (0: --impl reference takes the largest headline-ladder size that fits --cpu-budget; "
                         "the cpu_baseline leg of the GPU arm uses 1 M)")
    ap.add_argument("--cpu-budget", type=float, default=200.0, help="seconds of CPU time the reference arm may spend in total")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--input", default="by-index", choices=["by-index", "slabbed"],
                    help="multi-GPU only: how the atoms are distributed over the ranks before the list is built")
    ap.add_argument("--no-c4", action="store_true", help="skip the extra 100 M-atom record at 8 GPUs")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
