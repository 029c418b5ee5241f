#!/usr/bin/env python
"""bench.py -- MLUPS of the coupled LBM-DEM step on B200, with its roofline and CPU baseline.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload cfg4|cfg2|cfg3]

Workload (BASELINE.json configs[3], the one the metric's roofline target is quoted on):
4096 x 4096 lattice per GPU, scale 2.7, fp32 (-DSINGLE_PRECISION semantics), a synthetic
6355-grain packing with the radius range / extent of bin/a08_a4b4r18_7000.data per 4096 rows
(tools/make_sample.py, seed 12345).  With N GPUs the lattice is 4096*N x 4096 (x strips, one
process per GPU, weak scaling) and the packing is N times as wide.

One "step" = one LBM step (rasterise grains, fused collide-stream kernel, hydrodynamic forces)
plus the npDEM DEM sub-steps that follow it, i.e. npDEM calls of the reference's renderScene().

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

WORKLOADS = {
    # name: (rows per GPU, ly, scale, precision, sample preset, description)
    "cfg4": (4096, 4096, 2.7, "f32", "a08_7000",
             "BASELINE configs[3]: 4096x4096 lattice per GPU, scale 2.7, fp32, 6355 synthetic grains per 4096 rows"),
    "cfg3": (2048, 2048, 1.0, "f64", "a08d83",
             "BASELINE configs[2]: 2048x2048 lattice, scale 1, fp64, 726 synthetic grains"),
    "cfg5": (1024, 8192, 2.6, "f64", "50000_strip",
             "BASELINE configs[4]: 1024 rows x 8192 columns per GPU (8192x8192 on 8 GPUs), scale 2.6, fp64, 4000 synthetic grains per strip"),
    "cfg2": (1024, 1024, 1.0, "f64", None,
             "BASELINE configs[1]: 1024x1024 lattice, fp64, one grain outside the lattice (pure LBM stencil)"),
}


def make_sample_file(preset, n_gpus, rows, path):
    import make_sample as ms
    if preset is None:
        # SURVEY 8(d) cfg 2: the reference cannot run with zero grains; one grain outside the lattice
        ms.write_sample(path, [1.0], [0.5 * rows * n_gpus], [1.0], comment="# far grain")
        return 1
    n, r_min, r_max, width = ms.PRESETS[preset]
    r, x, y = ms.packed_sample(n * n_gpus, r_min, r_max, width * n_gpus, seed=12345)
    ms.write_sample(path, r, x, y, comment=f"# synthetic {preset} x{n_gpus} seed=12345")
    return len(r)


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
                                          "-lms", "50", "-i", str(self.index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
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


def cpu_reference(workload, steps, warmup, sample_path, tmpdir, as_line):
    """The reference's own CPU implementation (oracle/_ref, built from /root/reference with its
    GNU release flags + OpenMP) on the host cores; falls back to the plain-C oracle port."""
    rows, ly, scale, prec, preset, desc = WORKLOADS[workload]
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
                os.environ.setdefault("OMP_NUM_THREADS", str(cores))
                sim = Reference(rows, ly, scale_tag, prec, omp=True, release=True)
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
        for _ in range(warmup):
            sim.time_coupled(npd)
        t, lbm = 0.0, 0
        for _ in range(steps):
            dt, nl = sim.time_coupled(npd)
            t += dt
            lbm += nl
        mlups = rows * ly * lbm / t / 1e6
        return {"value": mlups, "unit": "MLUPS", "cores": cores, "kind": kind,
                "sample": f"{steps} coupled steps ({npd} DEM sub-steps each) of the {rows}x{ly} {prec} lattice "
                          f"(the per-GPU share of the workload), after {warmup} warm-up steps",
                "ms_per_step": 1e3 * t / max(lbm, 1)}
    finally:
        os.chdir(cwd)


def emit(line: dict):
    """the ONE line of this process's real stdout (fd 1 is pointed at stderr while the run lasts: NCCL and
    the reference print banners there)"""
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


_REAL_STDOUT = 1


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
    ap.add_argument("--cpu-steps", type=int, default=40)
    ap.add_argument("--sample-gpus", type=int, default=0,
                    help="diagnostic: build the grain sample as for this many GPUs (replicated-grain cost on one GPU)")
    a = ap.parse_args()
    rows, ly, scale, prec, preset, desc = WORKLOADS[a.workload]
    rank, local_rank, world = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("LOCAL_RANK", 0), ("WORLD_SIZE", 1)))
    n_gpus = a.gpus
    config = {"workload": desc, "lattice": [rows * n_gpus, ly], "precision": prec, "scale": scale,
              "decomposition": f"{n_gpus} x-strip(s) of {rows} rows, grains replicated",
              "cache": "populations are 2 x %.0f MB per GPU, larger than the 126 MB L2; no explicit flush" %
                       (rows * ly * 9 * (4 if prec == "f32" else 8) / 1e6)}
    tmpdir = tempfile.mkdtemp(prefix="lbmdem_bench_")
    sample_path = os.path.join(tmpdir, f"sample_r{rank}.data")

    if a.impl == "reference":
        if rank != 0:
            return 0
        n_grains = make_sample_file(preset, 1, rows, sample_path)
        W = max(1, min(a.warmup, 2))
        # every coupled step of the full lattice costs the host ~0.25 s: time at most REF_CAP of the K steps
        REF_CAP = 120
        timed = max(1, min(a.steps, REF_CAP))
        with quiet_stdout():
            cb = cpu_reference(a.workload, timed, W, sample_path, tmpdir, True)
        config["reference_steps_timed"] = timed
        config["grains"] = n_grains
        line = {"impl": "reference", "metric": "MLUPS", "value": cb["value"], "unit": "MLUPS", "n_gpus": n_gpus,
                "steps": a.steps, "warmup": W, "ms_per_step": cb["ms_per_step"], "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": prec, "data": "synthetic", "config": config,
                "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
                "e2e": {"value": cb["value"], "unit": "MLUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        emit(line)
        return 0

    import numpy as np
    import torch

    import lbmdem_dist as D
    import lbmdem_gpu as G

    if world != n_gpus:
        if world == 1 and n_gpus > 1:
            raise SystemExit("--gpus N > 1 must be launched with torchrun (one process per GPU)")
        n_gpus = world
    if world > 1:
        D.init_process_group("nccl")
    torch.cuda.set_device(local_rank)
    lx = rows * n_gpus
    n_grains = make_sample_file(preset, a.sample_gpus or n_gpus, rows, sample_path)
    config["grains"] = n_grains

    s = D.make_strip_solver(lx, ly, scale, prec, strict_fp=a.strict, kernel=a.kernel)
    s.init(sample_path)
    sc = s.scalars()
    npd = sc["npDEM"]
    config.update(npDEM=npd, dx=sc["dx"], strict_fp=a.strict)
    if a.kernel:
        config.update(kernel=a.kernel)
    stream = torch.cuda.ExternalStream(s.stream(), device=local_rank)

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
    s.step(npd * a.warmup)
    s.reset_kernel_timer(True)
    clocks = ClockSampler(local_rank)
    clocks.start()
    ms_total = timed_region(lambda k: s.step(npd * k), a.steps)
    k1_ms, k1_n, launches = s.kernel_timer()
    clk = clocks.stop()
    s.reset_kernel_timer(False)
    mlups = lx * ly * a.steps / (ms_total * 1e-3) / 1e6

    # ---- end-to-end arm: host buffers in, host buffers out, every step ----
    state = s.grains()[:, :9].copy()
    e2e_steps = a.steps
    # the reference prints its density checksum every stepConsole = 400 renderScene() calls (:1715):
    # the end-to-end loop asks for it at that cadence (it costs a stream-only pass over the lattice)
    every = max(1, 400 // npd)
    # grain rows in the precision of the run: float for an fp32 lattice (what a -DSINGLE_PRECISION reference holds)
    rows = "f32" if prec == "f32" else "f64"
    for _ in range(max(3, a.warmup // 2)):
        state, fh, dens = s.step_host(state, npd, rows=rows)
    D.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for k in range(e2e_steps):
        state, fh, d_ = s.step_host(state, npd, want_density=((k + 1) % every == 0 or k == e2e_steps - 1), rows=rows)
        dens = d_ if d_ is not None else dens
    torch.cuda.synchronize()
    t_e2e = D.max_over_ranks(time.perf_counter() - t0)
    D.barrier()
    real_b = 4 if prec == "f32" else 8
    # grain rows travel in the precision of the run: 9 values up, 9 + 3 down per grain
    e2e = {"value": lx * ly * e2e_steps / t_e2e / 1e6, "unit": "MLUPS",
           "h2d_bytes_per_step": n_grains * 9 * real_b, "d2h_bytes_per_step": n_grains * 12 * real_b + 8,
           "call": ("lbmdem_step_host_f32" if rows == "f32" else "lbmdem_step_host") +
                   " with page-locked host buffers (lbmdem_host_alloc): grain state up, npDEM renderScene() calls, grain state + fhf down every step, density checksum every 400 calls",
           "ms_per_step": 1e3 * t_e2e / e2e_steps, "density_checksum": dens}

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
    if os.path.exists(tpath):
        traffic = json.load(open(tpath)).get(f"{a.workload}_bytes_per_launch")
    roofline = {"bound": "hbm", "kernel": "lbm_rows_kernel<%s>" % ("float" if prec == "f32" else "double"),
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "peak_source": peak_src, "algorithmic_bytes_per_launch": bytes_per_launch,
                "avg_launch_ms": k1_avg_ms, "launches_timed": k1_n, "share_of_step": k1_ms / ms_total}

    line = {"metric": "MLUPS", "value": mlups, "unit": "MLUPS", "n_gpus": n_gpus, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": ms_total / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": prec, "data": "synthetic", "config": config, "clocks": clk, "e2e": e2e, "gpu_launches": launches,
            "roofline": roofline}
    if rank == 0 and n_gpus == 1 and not a.no_cpu_baseline:
        one = os.path.join(tmpdir, "sample_cpu.data")
        make_sample_file(preset, 1, rows, one)
        try:
            with quiet_stdout():
                cb = cpu_reference(a.workload, a.cpu_steps, 1, one, tmpdir, False)
            line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
        except Exception as e:  # noqa: BLE001
            line["cpu_baseline"] = {"value": None, "unit": "MLUPS", "cores": 0, "kind": "port", "sample": f"failed: {e}"}
    if rank == 0:
        emit(line)
    s.close()
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
