#!/usr/bin/env python
"""bench.py -- Gibbs draws/s of the collapsed-Gibbs sweep (BASELINE.json metric) on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl repo|reference] [--workload C2|C4]

A *step* is one full Gibbs sweep (LabeledLDA.training_iteration, LabeledLDA.py:101-125) over the synthetic
corpus; a *draw* is one resample of one (document, unique word) pair.  Rank 0 prints ONE JSON line.

  value         draws/s, whole job, corpus + state resident in HBM, CUDA-event time of `gibbs_sweep(K)` on the
                library's stream (sampling kernels + delta all-reduce + merge), max over ranks
  roofline      dominant kernel (the sampling kernel): algorithmic bytes per draw x draws / its
                CUDA-event time, against MEASURED_PEAKS.json hbm_gbs
  e2e           the same sweep driven through the C-ABI with HOST buffers: every step uploads the corpus + z from
                pinned host memory (gibbs_load), sweeps once and reads z / n_wk / n_dk / n_k back (gibbs_get_state);
                wall clock around the calls, max over ranks
  cpu_baseline  oracle/liboracle.so (C restatement of the reference loop, pinned to the unmodified reference by
                tests/golden) on a bounded sample of the same corpus, rank 0, N=1 only
  --impl reference   the CPU path alone (no lda_thesis_b200 import): all host threads, same metric/config

Workloads: N=1 -> C2 (100k docs x ~200 pairs, K=100, V=100k).  N>1 -> weak scaling: every rank holds
one C2-shaped shard (seed + rank) of one model, n_wk deltas all-reduced (NCCL, inside the library) every refresh
block.  `--workload C4` runs the 1M-doc K=500 corpus doc-sharded over the N ranks (strong scaling).
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

ALPHA, BETA = 0.1, 0.01
SEED = 20260201


# ------------------------------------------------------------------------------------------------ helpers
def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic(workload, fetch):
    """dram bytes per launch of the sampling kernel from the committed ncu capture, if any."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return json.load(f).get("%s/%s" % (workload, fetch))
    except Exception:
        return None


class ClockSampler(object):
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md clocks line)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.lines:
            p = [x.strip() for x in line.split(",")]
            if len(p) < 8:
                continue
            try:
                sm.append(float(p[0]))
                mx.append(float(p[1]))
            except ValueError:
                continue
            for n, v in zip(names, p[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def build_workload(name, rank, world):
    """This rank's shard of the synthetic corpus + its global draw/tile offsets."""
    from lda_thesis_b200 import synth
    if name == "C2":
        cfg = dict(synth.CONFIGS["C2"])
        cfg["seed"] += rank                      # weak scaling: one C2-shaped shard per rank
        c = synth.labeled_corpus(**cfg)
        desc = "C2 LabeledLDA synthetic: %d docs x Poisson(200) pairs, K=100, V=100k per GPU" % cfg["D"]
    elif name == "C4":
        cfg = dict(synth.CONFIGS["C4"])
        per = cfg["D"] // world
        cfg["D"] = per
        cfg["seed"] += rank                      # shard r of the 1M-doc corpus (documents are i.i.d.)
        c = synth.labeled_corpus(**cfg)
        desc = "C4 LabeledLDA synthetic: 1M docs x Poisson(250) pairs, K=500, V=100k, doc-sharded over %d GPUs" % world
    else:
        raise SystemExit("unknown workload " + name)
    return c, desc


def pinned_like(a):
    """Copy a NumPy array into page-locked host memory (torch is only the allocator here)."""
    import torch
    t = torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    return t.numpy(), t


# ------------------------------------------------------------------------------------------------ CPU arm
def cpu_port_rate(c, n_docs, sweeps, threads, mode):
    """draws/s of oracle/liboracle.so on the first n_docs documents.  mode: 'exact' (the reference's sequential
    chain, 1 thread) or 'snapshot' (the device schedule, OpenMP over documents)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle_lib
    from lda_thesis_b200 import synth
    s = synth.take_docs(c, 0, min(n_docs, c["D"]))
    o = oracle_lib.LldaOracle(s["doc_ptr"], s["word"], s["freq"], s["lab_ptr"], s["lab_idx"], s["K"], s["V"],
                              ALPHA, BETA, seed=SEED)
    times = []
    for i in range(sweeps + 1):                 # first sweep is the warm-up
        t0 = time.perf_counter()
        if mode == "exact":
            o.exact_sweep(1)
        else:
            o.snapshot_sweep(1, n_refresh=1, n_threads=threads)
        times.append(time.perf_counter() - t0)
    return o.N / float(np.median(times[1:])), o.N, s["D"]


def run_reference(args):
    """--impl reference: the CPU restatement of the reference loop on this box's host cores, nothing else."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle_lib
    oracle_lib.build()
    threads = max(1, oracle_lib.openmp_threads())
    wl = args.workload or "C2"
    c, desc = build_workload(wl, 0, max(1, args.gpus))
    n_docs = 20000 if wl == "C2" else 8000
    from lda_thesis_b200 import synth
    s = synth.take_docs(c, 0, min(n_docs, c["D"]))
    o = oracle_lib.LldaOracle(s["doc_ptr"], s["word"], s["freq"], s["lab_ptr"], s["lab_idx"], s["K"], s["V"],
                              ALPHA, BETA, seed=SEED)
    for _ in range(args.warmup):
        o.snapshot_sweep(1, n_threads=threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        o.snapshot_sweep(1, n_threads=threads)
    dt = time.perf_counter() - t0
    rate = o.N * args.steps / dt
    exact_rate, _, _ = cpu_port_rate(c, 4000, 2, 1, "exact")
    sample = ("first %d docs (%d draws) of the workload per step; C port of LabeledLDA.py:101-125 in the device's "
              "snapshot schedule, OpenMP over documents; the 1-thread sequential (reference-order) port runs at "
              "%.3g draws/s" % (s["D"], o.N, exact_rate))
    out = {"impl": "reference", "metric": "gibbs_draws_per_s", "value": rate, "unit": "draws/s",
           "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "fp32 weights / int32 counts",
           "data": "synthetic", "config": {"workload": desc, "sample_docs": int(s["D"]), "sample_draws": int(o.N)},
           "cpu_baseline": {"value": rate, "unit": "draws/s", "cores": threads, "kind": "port", "sample": sample,
                            "host_cpus": os.cpu_count()},
           "e2e": {"value": rate, "unit": "draws/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    print(json.dumps(out))


# ------------------------------------------------------------------------------------------------ GPU arm
def run_e2e(g, host, keep, n_local, draws_total, args, barrier, allmax):
    """Host buffers in, host state out, every step (see the module docstring)."""
    stt = g.get_state(n_wk=False)
    z_pin, t = pinned_like(stt["z"])
    keep.append(t)
    out_bufs = g.alloc_state_buffers(pinned=True)
    e2e_steps = max(3, min(args.steps, 10))
    h2d = sum(host[k].nbytes for k in host) + z_pin.nbytes
    d2h = sum(b.nbytes for b in out_bufs.values())
    for _ in range(2):
        g.load(host["doc_ptr"], host["word"], host["freq"], z_pin, host["lab_ptr"], host["lab_idx"])
        g.sweep(1)
        g.get_state_into(out_bufs)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        g.load(host["doc_ptr"], host["word"], host["freq"], z_pin, host["lab_ptr"], host["lab_idx"])
        g.sweep(1)
        g.get_state_into(out_bufs)
        z_pin[:] = out_bufs["z"]                    # next step continues the chain from the host-side state
    barrier()
    e2e_s = allmax(time.perf_counter() - t0)
    return {"value": draws_total * e2e_steps / e2e_s, "unit": "draws/s", "h2d_bytes_per_step": int(h2d),
            "d2h_bytes_per_step": int(d2h), "steps": e2e_steps, "ms_per_step": e2e_s / e2e_steps * 1e3,
            "path": "GibbsSampler.load (pinned host CSR + z) -> sweep(1) -> get_state (z, n_wk, n_dk, n_k)"}


def run_repo(args):
    import torch
    import torch.distributed as dist
    from lda_thesis_b200 import _lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != max(1, args.gpus):
        if world == 1 and args.gpus > 1:
            raise SystemExit("bench.py --gpus %d must be launched with torch.distributed.run (one rank per GPU)" % args.gpus)
    if _lib.device_count() < 1:
        raise SystemExit("bench.py: no CUDA device; libgibbs_b200 has no CPU path")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def allmax(x):
        if world == 1:
            return float(x)
        t = torch.tensor([float(x)], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def allsum(x):
        if world == 1:
            return float(x)
        t = torch.tensor([float(x)], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    wl = args.workload or "C2"
    c, desc = build_workload(wl, rank, world)
    n_local = int(c["doc_ptr"][-1])
    # global id of this shard's first document (RNG addressing and refresh-block assignment across shards)
    tile_docs = 256
    if world > 1:
        sizes = torch.zeros(world, dtype=torch.int64, device="cuda")
        sizes[rank] = c["D"]
        dist.all_reduce(sizes)
        doc_base = int(sizes.cpu().numpy()[:rank].sum())
    else:
        doc_base = 0
    n_refresh = args.refresh

    g = _lib.GibbsSampler(c["D"], c["V"], c["K"], ALPHA, BETA, seed=SEED, mode="snapshot", device=local,
                          n_refresh=n_refresh, doc_base=doc_base, tile_docs=tile_docs,
                          row_fetch=args.fetch)
    if world > 1:
        uid = torch.zeros(128, dtype=torch.uint8)
        if rank == 0:
            uid = torch.from_numpy(np.frombuffer(_lib.comm_unique_id(), dtype=np.uint8).copy())
        uid = uid.cuda()
        dist.broadcast(uid, 0)
        g.comm_init(world, rank, uid.cpu().numpy().tobytes())
    # pinned host copies: the e2e leg uploads from these every step
    host, keep = {}, []
    for k in ("doc_ptr", "word", "freq", "lab_ptr", "lab_idx"):
        host[k], t = pinned_like(c[k])
        keep.append(t)
    g.load(host["doc_ptr"], host["word"], host["freq"], None, host["lab_ptr"], host["lab_idx"])

    # ---- value: K sweeps, device-resident
    g.sweep(max(args.warmup, 3))
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    barrier()
    g.sweep(args.steps)
    barrier()
    st = g.stats()
    call_ms = allmax(st["last_call_ms"])
    if rank == 0:
        clk = clocks.stop()
    draws_total = allsum(n_local)
    value = draws_total * args.steps / (call_ms * 1e-3)
    kern_ms = allmax(st["last_sweep_ms"])            # sampling kernels of one sweep
    merge_ms = allmax(st["last_merge_ms"])
    peak, peak_src = measured_peak()
    bpd = st["bytes_per_draw"]
    ach = n_local * bpd / (kern_ms * 1e-3) / 1e9     # this rank's kernel (per GPU)
    fetch = g.row_fetch()
    roof = {"bound": "hbm", "kernel": "llda_%s_kernel (sampling)" % fetch, "achieved": ach, "peak": peak,
            "unit": "GB/s", "frac": ach / peak, "peak_source": peak_src, "bytes_per_draw": bpd,
            "model": "%s: %s" % (fetch, "16 B record r/w + 32 B x |labels| sectors + 16 B RED + per-doc terms"
                                 if fetch == "gather" else "16 B record r/w + 4*ldk row + 16 B RED + per-doc terms"),
            "kernel_ms_per_sweep": kern_ms, "merge_ms_per_sweep": merge_ms, "draws_per_launch_set": n_local,
            "traffic": ncu_traffic(wl, fetch)}

    # ---- e2e: host buffers in, host state out, every step
    e2e = None
    if not args.no_e2e:
        e2e = run_e2e(g, host, keep, n_local, draws_total, args, barrier, allmax)
    launches = int(st["last_launches"]) * args.steps

    # ---- cpu baseline (rank 0, N=1 only)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import oracle_lib
        oracle_lib.build()
        th = max(1, oracle_lib.openmp_threads())
        r1, n1, d1 = cpu_port_rate(c, 20000, 3, 1, "exact")
        rN, nN, dN = cpu_port_rate(c, 20000, 3, th, "snapshot")
        cpu = {"value": r1, "unit": "draws/s", "cores": 1, "kind": "port",
               "sample": "first %d docs (%d draws) of the workload, 1 warm-up + 3 sweeps, C port of the sequential "
                         "reference loop LabeledLDA.py:101-125 (oracle_llda_exact_sweep)" % (d1, n1),
               "all_cores": {"value": rN, "cores": th, "schedule": "snapshot (OpenMP over documents)"},
               "host_cpus": os.cpu_count(),
               "python_reference_note": "unmodified LabeledLDA.training_iteration measured at 6.35e4 draws/s (K=100, "
                                        "1 core, build container; BASELINE.md §2) -- it cannot travel to the GPU box"}
    g.close()
    if rank == 0:
        out = {"metric": "gibbs_draws_per_s", "value": value, "unit": "draws/s", "n_gpus": world, "steps": args.steps,
               "warmup": max(args.warmup, 3), "ms_per_step": call_ms / args.steps, "higher_is_better": True,
               "scaling": "weak" if wl == "C2" else "strong", "vs_baseline": None,
               "dtype": "fp32 weights / int32 counts", "data": "synthetic",
               "config": {"workload": desc, "mode": "snapshot", "n_refresh": n_refresh, "row_fetch": fetch,
                          "draws_per_sweep": int(draws_total), "alpha": ALPHA, "beta": BETA,
                          "l2": "no flush: per-sweep inputs (records %d MB + n_wk/delta tables %d MB) exceed the 126 MB L2"
                                % (n_local * 8 >> 20, 2 * c["V"] * st["ldk"] * 4 >> 20)},
               "roofline": roof, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches, "clocks": clk}
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="repo", choices=["repo", "reference"])
    ap.add_argument("--workload", default=None, choices=[None, "C2", "C4"])
    ap.add_argument("--refresh", type=int, default=1, help="refresh blocks per sweep (snapshot mode)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true", help="skip the e2e leg (profiling runs)")
    ap.add_argument("--fetch", default="auto", choices=["auto", "dense", "gather"], help="row fetch of the sampling kernel")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_repo(args)


if __name__ == "__main__":
    main()
