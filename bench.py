#!/usr/bin/env python
"""bench.py -- MLUPS of the coupled LBM-DEM step on B200, with its roofline and CPU baseline.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload cfg4|cfg2|cfg3|cfg5]

Default workload = BASELINE.json configs[3], the one the metric's roofline target is quoted on: 4096 x 4096
lattice per GPU, scale 2.7, fp32 (-DSINGLE_PRECISION semantics), the reference's own input
bin/a08_a4b4r18_7000.data (6355 grains; committed as tests/golden/a08_a4b4r18_7000.data).  With N GPUs the
lattice is 4096*N x 4096 (x strips, one process per GPU, WEAK scaling) and the sample is repeated N times along x
with the lattice extent of one strip as period.  Every line also carries, under "cfg5", BASELINE.json configs[4]
as it is stated: 8192 x 8192, scale 2.6, fp64, bin/50000.data (49 987 grains), split over the N GPUs (strong
scaling; fits one B200).  With N > 1 the line carries "strip_check": before anything is timed the same global
lattice runs 3 coupled steps on rank 0 ALONE and the strips must reproduce it bit for bit.

One "step" = one LBM step (rasterise grains, fused collide-stream kernel, hydrodynamic forces) plus the npDEM
DEM sub-steps that follow it, i.e. npDEM calls of the reference's renderScene().

Prints ONE JSON line (see README / DESIGN.md "Measurement").
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "2d-lbm-dem_b200"), os.path.join(ROOT, "tools")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLD = os.path.join(ROOT, "tests", "golden")
WORKLOADS = {
    # name: rows per GPU (weak) or total rows (strong), ly, scale, precision, reference input, synthetic stand-in, scaling
    "cfg4": dict(rows=4096, ly=4096, scale=2.7, prec="f32", fixture="a08_a4b4r18_7000.data", preset="a08_7000", scaling="weak",
                 desc="BASELINE configs[3]: 4096x4096 lattice per GPU, scale 2.7, fp32, bin/a08_a4b4r18_7000.data (6355 grains) per 4096 rows"),
    "cfg3": dict(rows=2048, ly=2048, scale=1.0, prec="f64", fixture="a08d83.data", preset="a08d83", scaling="weak",
                 desc="BASELINE configs[2]: 2048x2048 lattice, scale 1, fp64, bin/a08d83.data (726 grains)"),
    "cfg5": dict(rows=8192, ly=8192, scale=2.6, prec="f64", fixture="50000.data", preset="50000", scaling="strong",
                 desc="BASELINE configs[4]: 8192x8192 lattice, scale 2.6, fp64, bin/50000.data (49987 grains), x strips over the GPUs"),
    "cfg2": dict(rows=1024, ly=1024, scale=1.0, prec="f64", fixture=None, preset=None, scaling="weak",
                 desc="BASELINE configs[1]: 1024x1024 lattice, fp64, one grain outside the lattice (pure LBM stencil)"),
}


def make_sample_file(workload, n_gpus, path):
    """The grain file of the run, in the reference's input format (read_sample, src/main.c:609-639).
    Returns (number of grains, what the data is)."""
    import make_sample as ms
    W = WORKLOADS[workload]
    tiles = n_gpus if W["scaling"] == "weak" else 1
    if W["fixture"] is None:
        # SURVEY 8(d) cfg 2: the reference cannot run with zero grains; one grain outside the lattice
        ms.write_sample(path, [1.0], [0.5 * W["rows"] * tiles], [1.0], comment="# far grain")
        return 1, "synthetic (one grain outside the lattice)"
    src = os.path.join(GOLD, W["fixture"])
    if os.path.exists(src):
        with open(src) as fh:
            comment = fh.readline().rstrip("\n")
            n = int(fh.readline())
            rows = [fh.readline().split() for _ in range(n)]
        period = 0.1 * W["rows"] / W["scale"]   # lattice extent of one strip's rows in the file's unit (mm)
        with open(path, "w") as fh:
            fh.write(f"{comment} (x{tiles})\n{n * tiles}\n")
            for k in range(tiles):
                for r_, x_, y_ in rows:
                    fh.write(f"{r_}\t{x_ if k == 0 else repr(float(x_) + k * period)}\t{y_}\n")
        what = f"reference fixture bin/{W['fixture']}" + (f", repeated {tiles} times along x (period {period:.4f} mm)" if tiles > 1 else "")
        return n * tiles, what
    n, r_min, r_max, width = ms.PRESETS[W["preset"]]
    r, x, y = ms.packed_sample(n * tiles, r_min, r_max, width * tiles, seed=12345)
    ms.write_sample(path, r, x, y, comment=f"# synthetic {W['preset']} x{tiles} seed=12345")
    return len(r), f"synthetic ({W['fixture']} is missing from tests/golden)"


class quiet_stdout:
    """The reference printf()s its banner to fd 1; keep this process's stdout to the JSON line."""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)
        return self

    def __exit__(self, *exc):
        try:
            import ctypes
            ctypes.CDLL(None).fflush(None)
        except Exception:  # noqa: BLE001
            pass
        os.dup2(self.saved, 1)
        os.close(self.saved)
        return False


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "20", "-i", str(self.index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
            time.sleep(0.25)   # nvidia-smi needs a moment before its first row
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons = [], None, set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                smax = float(r[2])
            except (ValueError, IndexError):
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples": len(sm)}


def cpu_reference(workload, steps, warmup, sample_path, tmpdir, budget_s=100.0):
    """The reference's own CPU implementation (oracle/_ref, built from /root/reference with its GNU release flags +
    OpenMP) on ALL host cores, on the one-GPU share of the workload; falls back to the plain-C oracle port.
    At most `steps` coupled steps are timed, fewer when they would not fit `budget_s` seconds."""
    W = WORKLOADS[workload]
    rows, ly, scale, prec = W["rows"], W["ly"], W["scale"], W["prec"]
    from oracle import build as obuild
    cwd = os.getcwd()
    os.chdir(tmpdir)  # the reference writes stats.data / VTK into the cwd
    try:
        kind, cores, sim = "reference", os.cpu_count() or 1, None
        scale_tag = ("%g" % scale) if scale != int(scale) else "%d." % int(scale)
        lib = obuild.ref_lib_path(rows, ly, scale_tag, prec, omp=True, release=True)
        if os.path.exists(lib) or obuild.ref_available():
            try:
                from oracle.refwrap import Reference
                # torch.distributed.run exports OMP_NUM_THREADS=1: the reference gets every host core regardless
                os.environ["OMP_NUM_THREADS"] = str(cores)
                sim = Reference(rows, ly, scale_tag, prec, omp=True, release=True)
                sim.set_omp_threads(cores)
                sim.init(sample_path)
                cores = sim.omp_threads()
            except Exception as e:  # noqa: BLE001 - fall back to the port, say why
                print(f"# reference library unusable ({e}); timing the oracle port", file=sys.stderr)
                sim = None
        if sim is None:
            from oracle.oraclewrap import Oracle
            kind, cores = "port", 1
            sim = Oracle(rows, ly, scale, prec)
            sim.init(sample_path)
        npd = sim.scalars()["npDEM"]
        t_w = 0.0
        for _ in range(warmup):
            t_w += sim.time_coupled(npd)[0]
        est = t_w / max(warmup, 1)
        timed = max(1, min(steps, int(budget_s / max(est, 1e-6)))) if warmup else steps
        t, lbm = 0.0, 0
        for _ in range(timed):
            dt, nl = sim.time_coupled(npd)
            t += dt
            lbm += nl
        mlups = rows * ly * lbm / t / 1e6
        return {"value": mlups, "unit": "MLUPS", "cores": cores, "kind": kind,
                "sample": f"{timed} coupled steps ({npd} DEM sub-steps each) of the {rows}x{ly} {prec} lattice "
                          f"(the one-GPU share of the workload), after {warmup} warm-up steps",
                "ms_per_step": 1e3 * t / max(lbm, 1), "steps_timed": timed, "npDEM": npd, "dx": sim.scalars()["dx"]}
    finally:
        os.chdir(cwd)


def emit(line: dict):
    """the ONE line of this process's real stdout (fd 1 is pointed at stderr while the run lasts: NCCL and
    the reference print banners there)"""
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


_REAL_STDOUT = 1


def base_config(workload, n_gpus, n_grains, data):
    W = WORKLOADS[workload]
    strong = W["scaling"] == "strong"
    lx = W["rows"] if strong else W["rows"] * n_gpus
    per_gpu = lx // n_gpus
    return {"workload": W["desc"], "lattice": [lx, W["ly"]], "precision": W["prec"], "scale": W["scale"],
            "decomposition": f"{n_gpus} x-strip(s) of {per_gpu} rows, grains replicated" + (
                "" if n_gpus == 1 else "; ghost rows over NCCL, force sums " + (
                    "by ncclAllReduce" if os.environ.get("LBMDEM_PEER_SUMS", "1") == "0" else "through CUDA IPC peer mappings")),
            "cache": "populations are 2 x %.0f MB per GPU, larger than the 126 MB L2; no explicit flush" %
                     (per_gpu * W["ly"] * 9 * (4 if W["prec"] == "f32" else 8) / 1e6),
            "grains": n_grains, "data": data}


def strip_check(G, D, workload, lx, sample_path, rank, local_rank, s, npd, strict, kernel):
    """Before anything is timed: the same GLOBAL lattice on rank 0 alone, 3 coupled steps, against the strips.
    Grain rows and hydrodynamic forces are compared bit for bit; populations and node indices through the exact
    position-weighted integer fingerprint (lbmdem_state_checksum), whose strip values add up mod 2^64."""
    import numpy as np
    import torch
    import torch.distributed as dist
    W = WORKLOADS[workload]
    steps = 3
    s.step(npd * steps)
    mine = s.state_checksum()
    limbs = torch.tensor([(v >> (16 * k)) & 0xFFFF for v in mine for k in range(4)], dtype=torch.int64, device="cuda")
    dist.all_reduce(limbs)
    dens = torch.tensor([s.total_density()], dtype=torch.float64, device="cuda")
    dist.all_reduce(dens)
    out = None
    if rank == 0:
        tot = [sum(int(limbs[4 * j + k].item()) << (16 * k) for k in range(4)) & ((1 << 64) - 1) for j in range(2)]
        one = G.Solver(lx, W["ly"], W["scale"], W["prec"], device=local_rank, strict_fp=strict, kernel=kernel)
        one.init(sample_path)
        one.step(npd * steps)
        ref = one.state_checksum()
        d1 = one.total_density()
        out = {"coupled_steps": steps, "against": f"the {lx}x{W['ly']} lattice on rank 0 alone",
               "grains_equal": bool(np.array_equal(one.grains(), s.grains())),
               "fhf_sums_equal": bool(np.array_equal(one.fhf(), s.fhf())),
               "f_fingerprint_equal": tot[0] == ref[0], "obst_fingerprint_equal": tot[1] == ref[1],
               "density_equal": bool(abs(dens.item() - d1) <= 1e-12 * abs(d1)), "density": [dens.item(), d1]}
        out["ok"] = all(out[k] for k in ("grains_equal", "fhf_sums_equal", "f_fingerprint_equal", "obst_fingerprint_equal", "density_equal"))
        one.close()
        del one
        torch.cuda.empty_cache()
    D.barrier()
    return out


def measure(workload, a, rank, local_rank, world, tmpdir, with_check, with_e2e=True, steps=None, warmup=None):
    """One workload on the process group's GPUs; returns the fields of the JSON line (rank 0's are complete)."""
    import numpy as np  # noqa: F401
    import torch

    import lbmdem_dist as D
    import lbmdem_gpu as G

    W = WORKLOADS[workload]
    steps = a.steps if steps is None else steps
    warmup = a.warmup if warmup is None else warmup
    n_gpus = world
    ly, scale, prec = W["ly"], W["scale"], W["prec"]
    lx = W["rows"] if W["scaling"] == "strong" else W["rows"] * n_gpus
    sample_path = os.path.join(tmpdir, f"sample_{workload}_r{rank}.data")
    n_grains, data = make_sample_file(workload, a.sample_gpus or n_gpus, sample_path)
    config = base_config(workload, n_gpus, n_grains, data)

    s = D.make_strip_solver(lx, ly, scale, prec, strict_fp=a.strict, kernel=a.kernel)
    s.init(sample_path)
    sc = s.scalars()
    npd = sc["npDEM"]
    config.update(npDEM=npd, dx=sc["dx"], strict_fp=a.strict)
    if a.kernel:
        config.update(kernel=a.kernel)
    stream = torch.cuda.ExternalStream(s.stream(), device=local_rank)
    check = None
    if with_check and world > 1:
        check = strip_check(G, D, workload, lx, sample_path, rank, local_rank, s, npd, a.strict, a.kernel)

    def timed_region(fn, k):
        D.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        fn(k)
        e1.record(stream)
        D.barrier()
        torch.cuda.synchronize()
        return D.max_over_ranks(e0.elapsed_time(e1))

    # ---- device-resident arm: K coupled steps, inputs already in HBM ----
    s.step(npd * warmup)
    s.reset_kernel_timer(True)
    clocks = ClockSampler(local_rank)
    clocks.start()
    ms_total = timed_region(lambda k: s.step(npd * k), steps)
    k1_ms, k1_n, launches = s.kernel_timer()
    clk = clocks.stop()
    lc = s.list_counts()
    config["sparse_work_last_step"] = {"bounce_links": lc["links"], "boundary_nodes": lc["boundary_nodes"],
                                       "tiles_rebuilt": lc["tiles_rebuilt"], "tiles": ((s.nx + (8 if world > 1 else 0) + 31) // 32) * ((ly + 63) // 64)}
    s.reset_kernel_timer(False)
    mlups = lx * ly * steps / (ms_total * 1e-3) / 1e6
    real_b = 4 if prec == "f32" else 8

    e2e = None
    if with_e2e:
        # ---- end-to-end arm: host buffers in, host buffers out, every step ----
        # with N > 1 GPUs every rank moves ITS share of the (replicated) grain rows across the host boundary and the
        # ranks pass them on over NVLink (lbmdem_step_host_share); with one GPU that is lbmdem_step_host
        share = world > 1
        i0, i1 = s.share() if share else (0, n_grains)
        state = s.grains()[i0:i1, :9].copy()
        # the reference prints its density checksum every stepConsole = 400 renderScene() calls (:1715):
        # the end-to-end loop asks for it at that cadence (it costs a stream-only pass over the lattice)
        every = max(1, 400 // npd)
        # grain rows in the precision of the run: float for an fp32 lattice (what a -DSINGLE_PRECISION reference holds)
        grain_rows = "f32" if prec == "f32" else "f64"
        dens = None
        for _ in range(max(3, warmup // 2)):
            state, fh, dens = s.step_host(state, npd, rows=grain_rows, share=share)
        D.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for k in range(steps):
            state, fh, d_ = s.step_host(state, npd, want_density=((k + 1) % every == 0 or k == steps - 1), rows=grain_rows, share=share)
            dens = d_ if d_ is not None else dens
        torch.cuda.synchronize()
        t_e2e = D.max_over_ranks(time.perf_counter() - t0)
        D.barrier()
        # grain rows travel in the precision of the run: 9 values up, 9 + 3 down per grain
        e2e = {"value": lx * ly * steps / t_e2e / 1e6, "unit": "MLUPS",
               "h2d_bytes_per_step": n_grains * 9 * real_b, "d2h_bytes_per_step": n_grains * 12 * real_b + 8 * world,
               "call": ("lbmdem_step_host" + ("_share" if share else "") + ("_f32" if grain_rows == "f32" else "")) +
                       " with page-locked host buffers (lbmdem_host_alloc): grain state up, npDEM renderScene() calls, grain state + fhf down every step, density checksum every 400 calls" +
                       ("; every rank moves its share of the grain rows, the bytes are the sum over the ranks" if share else ""),
               "ms_per_step": 1e3 * t_e2e / steps, "density_checksum": dens}

    # ---- roofline of the dominant kernel (K1), measured live with CUDA events on its stream ----
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (measured copy bandwidth)"
    else:
        peak, peak_src = 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"
    bytes_per_launch = 2 * 9 * real_b * (s.nx * ly)        # 2*9*sizeof(real) per lattice update
    k1_avg_ms = k1_ms / max(k1_n, 1)
    achieved = bytes_per_launch / (k1_avg_ms * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "k1_traffic.json")
    if os.path.exists(tpath) and n_gpus == 1:
        traffic = json.load(open(tpath)).get(f"{workload}_bytes_per_launch")
    roofline = {"bound": "hbm", "kernel": "lbm_rows_kernel<%s>" % ("float" if prec == "f32" else "double"),
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "peak_source": peak_src, "algorithmic_bytes_per_launch": bytes_per_launch,
                "avg_launch_ms": k1_avg_ms, "launches_timed": k1_n, "share_of_step": k1_ms / ms_total}
    s.close()
    del s
    torch.cuda.empty_cache()
    out = {"value": mlups, "ms_per_step": ms_total / steps, "steps": steps, "warmup": warmup, "dtype": prec,
           "scaling": W["scaling"], "data": data, "config": config, "clocks": clk, "gpu_launches": launches,
           "roofline": roofline, "sample_path": sample_path}
    if e2e is not None:
        out["e2e"] = e2e
    if check is not None:
        out["strip_check"] = check
    return out


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cfg4", choices=sorted(WORKLOADS))
    ap.add_argument("--strict", type=int, default=0)
    ap.add_argument("--kernel", type=int, default=0, help="cross-check switches of lbmdem_params.kernel (0 = the product path)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-cfg5", action="store_true", help="skip the extra BASELINE configs[4] measurement")
    ap.add_argument("--no-strip-check", action="store_true")
    ap.add_argument("--cpu-steps", type=int, default=40)
    ap.add_argument("--sample-gpus", type=int, default=0,
                    help="diagnostic: build the grain sample as for this many GPUs (replicated-grain cost on one GPU)")
    a = ap.parse_args()
    W = WORKLOADS[a.workload]
    rank, local_rank, world = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("LOCAL_RANK", 0), ("WORLD_SIZE", 1)))
    n_gpus = a.gpus
    tmpdir = tempfile.mkdtemp(prefix="lbmdem_bench_")

    if a.impl == "reference":
        if rank != 0:
            return 0
        # the reference is one shared-memory process: it runs the ONE-GPU share of the workload on all host cores
        sample_path = os.path.join(tmpdir, "sample_ref.data")
        n_grains, data = make_sample_file(a.workload, 1, sample_path)
        Wn = max(1, min(a.warmup, 5))
        with quiet_stdout():
            cb = cpu_reference(a.workload, a.steps, Wn, sample_path, tmpdir)
        config = base_config(a.workload, 1, n_grains, data)
        config.update(npDEM=cb["npDEM"], dx=cb["dx"], strict_fp=0)
        config["decomposition"] = f"1 process, {cb['cores']} OpenMP threads; the requested --gpus {n_gpus} only names the b200 arm it is compared with"
        config["reference_steps_timed"] = cb["steps_timed"]
        line = {"impl": "reference", "metric": "MLUPS", "value": cb["value"], "unit": "MLUPS", "n_gpus": n_gpus,
                "steps": a.steps, "warmup": Wn, "ms_per_step": cb["ms_per_step"], "higher_is_better": True,
                "scaling": W["scaling"], "vs_baseline": None, "dtype": W["prec"], "data": data, "config": config,
                "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
                "e2e": {"value": cb["value"], "unit": "MLUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        emit(line)
        return 0

    import torch

    import lbmdem_dist as D

    if world != n_gpus:
        if world == 1 and n_gpus > 1:
            raise SystemExit("--gpus N > 1 must be launched with torchrun (one process per GPU)")
        n_gpus = world
    if world > 1:
        D.init_process_group("nccl")
    torch.cuda.set_device(local_rank)

    m = measure(a.workload, a, rank, local_rank, world, tmpdir, with_check=not a.no_strip_check)
    line = {"metric": "MLUPS", "value": m["value"], "unit": "MLUPS", "n_gpus": n_gpus, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": m["ms_per_step"], "higher_is_better": True, "scaling": m["scaling"], "vs_baseline": None,
            "dtype": m["dtype"], "data": m["data"], "config": m["config"], "clocks": m["clocks"], "e2e": m["e2e"],
            "gpu_launches": m["gpu_launches"], "roofline": m["roofline"]}
    if "strip_check" in m:
        line["strip_check"] = m["strip_check"]
    if a.workload == "cfg4" and not a.no_cfg5:
        # BASELINE configs[4] as stated (8192^2 fp64, 49 987 grains), on the same GPUs: a second, shorter measurement
        try:
            c5 = measure("cfg5", a, rank, local_rank, world, tmpdir, with_check=not a.no_strip_check,
                         steps=max(10, min(a.steps, 100)), warmup=max(3, min(a.warmup, 10)))
            line["cfg5"] = {k: c5[k] for k in ("value", "ms_per_step", "steps", "warmup", "dtype", "scaling", "data", "config",
                                               "e2e", "gpu_launches", "roofline") if k in c5}
            line["cfg5"].update(metric="MLUPS", unit="MLUPS", n_gpus=n_gpus)
            if "strip_check" in c5:
                line["cfg5"]["strip_check"] = c5["strip_check"]
        except Exception as e:  # noqa: BLE001 - the headline line must survive
            line["cfg5"] = {"failed": repr(e)}
    if rank == 0 and n_gpus == 1 and not a.no_cpu_baseline:
        one = os.path.join(tmpdir, "sample_cpu.data")
        make_sample_file(a.workload, 1, one)
        try:
            with quiet_stdout():
                cb = cpu_reference(a.workload, a.cpu_steps, 1, one, tmpdir, budget_s=25.0)
            line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
        except Exception as e:  # noqa: BLE001
            line["cpu_baseline"] = {"value": None, "unit": "MLUPS", "cores": 0, "kind": "port", "sample": f"failed: {e}"}
    if rank == 0:
        emit(line)
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
