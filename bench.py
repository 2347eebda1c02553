#!/usr/bin/env python
"""Benchmark of the hot path: Wolter-I trace + HPD, rays/s (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this engine (one rank per GPU)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (oracle port)

A STEP is one pass of the hot path over one synthetic bundle:
    transform -> wolterprimary -> reflect -> woltersecondary -> reflect -> flat  (ONE fused kernel)
    -> hpd (centroid reduction + exact radix select; all-reduced over ranks for N>1)
on ``--rays`` rays per GPU (default 1.25e8 = 10 GB of bundle; x8 GPUs = the 1e9-ray bundle of
BASELINE config 5, traced with config 1's chain).  The source bundle is generated on the
device once and stays resident; every step reads it and writes a second bundle, so each step
does the full Newton work on fresh rays.  Inputs (6 GB/GPU) are far larger than L2, so no L2
flush is needed between steps.

Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for every field.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "rays/sec (fp64, Wolter-I trace + HPD)"
R0, Z0, PSI = 220., 8400., 1.
RIN, ROUT = 220., 220.6
SURFACES_PER_RAY = 3                 # primary, secondary, focal plane
# algorithmic HBM bytes per ray of the fused trace kernel: read x,y,z,l,m,n + write x,y,z,l,m,n,ux,uy,uz
FP64_INSTR_PER_RAY = 283        # DADD+DMUL+DFMA+DSETP per ray, profiles/r01f_k_chain.txt (source page)
TRACE_BYTES_PER_RAY = 6 * 8 + 9 * 8
# per-routine API for comparison: 144+96+72+96+72+96
PERCALL_BYTES_PER_RAY = 576


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="engine", choices=["engine", "reference"])
    ap.add_argument("--rays", type=float, default=1.25e8, help="rays per GPU")
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--ref-rays", type=float, default=2.0e7, help="rays per step of the reference arm")
    return ap.parse_args()


# ------------------------------------------------------------------------------- clocks
class ClockSampler:
    """nvidia-smi clocks line of /opt/skills/guides/B200_PROFILING.md, sampled during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.lines = []          # (wall-clock time the line was read, line)
        self.proc = None
        self.t0 = self.t1 = None

    def mark_begin(self):
        self.t0 = time.time()

    def mark_end(self):
        self.t1 = time.time()

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        # nvidia-smi is started BEFORE the warm-up (its start-up holds driver locks for ~100 ms and
        # must not land in the timed region); keep the samples read inside the timed region, or --
        # when the region is shorter than the sampling period -- those taken under load since the
        # warm-up began (the GPU runs the same step back to back from there on).
        inside = [ln for t, ln in self.lines if self.t0 is not None and self.t0 <= t <= (self.t1 or t)]
        window = "timed region"
        if not inside:
            inside = [ln for t, ln in self.lines]
            window = "warm-up + timed region"
        self.window = window
        for ln in inside:
            f = [t.strip() for t in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
                power.append(float(f[3]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(mx), "reasons": sorted(reasons),
                "samples": len(sm), "power_w_max": max(power), "window": self.window}


# ------------------------------------------------------------------------------- CPU arm
def cpu_step_factory(n):
    """One step of the reference's CPU path on n rays: the six Fortran-routine calls (C oracle
    port, OpenMP over rays like the Fortran) under the reference's Python call pattern, then
    analyses.hpd (numpy).  Returns (fresh_inputs(), step(rays))."""
    import numpy as np
    from oracle import chains, pyref

    np.random.seed(0)
    src = pyref.subannulus(RIN, ROUT, 2 * np.pi, n, zhat=-1.)

    def fresh():
        return [r.copy() for r in src]

    def step(rays):
        chains.run_steps_cpu(rays, chains.wolter1_steps(R0, Z0, PSI))
        return pyref.hpd(rays)
    return fresh, step


def omp_threads():
    return int(os.environ.get("OMP_NUM_THREADS", os.cpu_count() or 1))


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path.  Its Fortran cannot
    be compiled in this image (no Fortran compiler), so this is the C oracle port (kind
    'port'), OpenMP on every host core, driven through the reference's call pattern."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import f2py as of
    of.lib()
    n = int(args.ref_rays)
    fresh, step = cpu_step_factory(n)
    for _ in range(max(1, min(args.warmup, 2))):
        step(fresh())
    inputs = [fresh() for _ in range(args.steps)] if n * 80 * args.steps < 24e9 else None
    t = 0.
    hp = None
    for k in range(args.steps):
        rays = inputs[k] if inputs is not None else fresh()
        t0 = time.perf_counter()
        hp = step(rays)
        t += time.perf_counter() - t0
    val = n * args.steps / t
    cores = omp_threads()
    sample = "%d steps x %.3g rays (same chain + hpd), C oracle port with OpenMP" % (args.steps, n)
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": "rays/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "wolter1_trace_hpd", "rays_per_step": n, "surfaces_per_ray": SURFACES_PER_RAY,
                   "r0": R0, "z0": Z0, "psi": PSI, "hpd": hp},
        "cpu_baseline": {"value": val, "unit": "rays/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def cpu_baseline():
    """Bounded sample (~10-20 s) of the same step on the host cores."""
    from oracle import f2py as of
    of.lib()
    n = 4_000_000
    fresh, step = cpu_step_factory(n)
    step(fresh())                                   # warm-up (page faults, thread pool)
    rays = fresh()
    t0 = time.perf_counter()
    step(rays)
    t1 = time.perf_counter() - t0
    reps = max(1, min(20, int(12. / max(t1, 1e-3))))
    ins = [fresh() for _ in range(reps)]
    t0 = time.perf_counter()
    for r in ins:
        step(r)
    t = time.perf_counter() - t0
    return {"value": n * reps / t, "unit": "rays/s", "cores": omp_threads(), "kind": "port",
            "sample": "%d x %.1e rays, six oracle-routine passes + numpy hpd, OpenMP on all host cores "
                      "(C restatement of the f2py Fortran; the Fortran itself cannot be built here)" % (reps, n)}


# ------------------------------------------------------------------------------- engine arm
def run_engine(args):
    import numpy as np
    import torch
    import torch.distributed as td

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (there is no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        td.init_process_group("nccl", device_id=dev)
    import pyxfocus_b200 as pxf
    from pyxfocus_b200 import dist as pdist
    from pyxfocus_b200._call import bundle_alloc

    n = int(args.rays)
    total = n * world
    first = rank * n
    # ---- resident inputs: source bundle generated on the device (counter-based, seed 0,
    # global ray index => identical stream for any sharding)
    src = pxf.sources.subannulus(RIN, ROUT, 2 * np.pi, n, zhat=-1., rng="philox", seed=0, first=first, device=dev)
    out = bundle_alloc(n, dev, zero=True)
    prog = (pxf.Program().transform(0., 0., Z0, 0., 0., 0.).wolterprimary(R0, Z0, PSI).reflect()
            .woltersecondary(R0, Z0, PSI).reflect().flat())

    sums = torch.zeros(16, dtype=torch.float64, device=dev)

    def step():
        prog.run(src, out=out, sums=sums)          # trace kernel also emits the centroid sums
        return pdist.hpd(out, sums=sums, total=total, min_shard=n) if world > 1 else pxf.analyses.hpd(out, sums=sums)

    def barrier():
        if world > 1:
            td.barrier()
        torch.cuda.synchronize()

    clk = ClockSampler(local)
    if rank == 0:
        clk.start()
    # warm-up: at least W steps, and keep stepping (GPU busy, clocks up) until nvidia-smi has had
    # ~0.4 s to finish its start-up, so that none of it lands in the timed region
    t_w = time.time()
    nw = 0
    while nw < max(args.warmup, 3) or (world == 1 and time.time() - t_w < 0.4):
        hp = step()
        nw += 1
    barrier()
    # ---- timed region: K steps, CUDA events on the launching stream, clocks sampled alongside
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    ws = pxf.analyses.hpd_workspace(n, dev)

    def timed_loop(deferred):
        """K steps between two events.  deferred: every step's HPD is left on the device (float64[4] =
        [HPD, lower, upper, valid]) and all K are read after the timed region -- nothing forces a host
        round trip between steps (a stream of bundles analysed back to back).  Otherwise each step
        reads its HPD back before the next trace is launched."""
        res = torch.zeros(args.steps, 4, dtype=torch.float64, device=dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        last = None
        barrier()
        clk.mark_begin()
        e0.record()
        for k in range(args.steps):
            kev[k][0].record()
            prog.run(src, out=out, sums=sums)
            kev[k][1].record()
            if deferred:
                if world > 1:
                    pdist.hpd(out, sums=sums, total=total, min_shard=n, out=res[k])
                else:
                    pxf.analyses.hpd_enqueue(out, res[k], ws, sums=sums)
            else:
                last = pdist.hpd(out, sums=sums, total=total, min_shard=n) if world > 1 else pxf.analyses.hpd(out, sums=sums)
        e1.record()
        barrier()
        clk.mark_end()
        if deferred:
            h = res.cpu().numpy()
            if not (h[:, 3] != 0.).all():
                return None, None          # a bracket missed (~1e-9): the caller re-times with read-backs
            last = float(h[-1, 0])
        return e0.elapsed_time(e1), last

    launches0 = pxf.launch_count()
    readback = "deferred: K results read after the timed region"
    ms, hp_t = timed_loop(True)
    if ms is None:
        launches0 = pxf.launch_count()
        readback = "per step"
        ms, hp_t = timed_loop(False)
    assert hp_t == hp or (hp_t != hp_t and hp != hp), "HPD changed between warm-up and timed steps"
    launches = pxf.launch_count() - launches0
    clocks = clk.stop() if rank == 0 else None
    trace_ms = sum(a.elapsed_time(b) for a, b in kev) / args.steps
    if world > 1:
        t = torch.tensor([ms, trace_ms], dtype=torch.float64, device=dev)
        td.all_reduce(t, op=td.ReduceOp.MAX)
        ms, trace_ms = float(t[0]), float(t[1])
    value = total * args.steps / (ms * 1e-3)

    # ---- end to end through the host-array C ABI entry (pinned host rows -> device -> host)
    e2e = None
    if not args.no_e2e:
        e2e = run_e2e(args, pxf, pdist, src, prog, n, world, dev, barrier)

    if rank != 0:
        if world > 1:
            td.destroy_process_group()
        return
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    achieved = TRACE_BYTES_PER_RAY * n / (trace_ms * 1e-3) / 1e9
    traffic = None
    try:
        # dram__bytes_read+write of the fused kernel from the ncu --set full capture (taken at
        # 2e7 rays/launch; the kernel streams, so bytes scale with the ray count)
        tj = json.load(open(os.path.join(ROOT, "profiles", "trace_kernel_traffic.json")))
        traffic = float(tj["dram_bytes_per_ray"]) * n
    except (OSError, ValueError):
        pass
    line = {
        "metric": METRIC, "value": value, "unit": "rays/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "wolter1_trace_hpd (BASELINE configs[0] chain at configs[4] scale: "
                               "subannulus -> transform -> wolterprimary -> reflect -> woltersecondary -> reflect -> flat -> hpd)",
                   "rays_per_gpu": n, "total_rays": total, "surfaces_per_ray": SURFACES_PER_RAY,
                   "interactions_per_s": value * SURFACES_PER_RAY, "r0": R0, "z0": Z0, "psi": PSI,
                   "parallelism": "rays sharded %d-way, no trace-time communication; HPD = bracketed exact "
                                  "select, 3 small collectives (sums+sample all-gather, counters+bins "
                                  "all-reduce, key lists all-gather)" % world,
                   "l2": "inputs (%.1f GB/GPU) larger than L2, no flush needed" % (48e-9 * n),
                   "hpd": hp, "trace_kernel_ms": trace_ms, "hpd_readback": readback},
        "roofline": {"bound": "hbm", "kernel": "k_program (fused trace)", "achieved": achieved, "peak": peak,
                     "unit": "GB/s", "frac": achieved / peak,
                     "peak_source": "measured (MEASURED_PEAKS.json hbm_gbs)" if peaks else "fallback",
                     "algorithmic_bytes_per_ray": TRACE_BYTES_PER_RAY, "traffic": traffic,
                     "note": "the fused kernel is COMPUTE bound (a compute-only replay of the chain takes the "
                             "same 3.5 ms, profiles/r01_trace_compute.txt): serial fp64 Newton/division chains, "
                             "no FMA contraction for bit parity; see DESIGN.md and profiles/r01_notes.md",
                     # second denominator: fp64 warp instructions per ray from the ncu capture (static),
                     # duration measured live; peak = 148 SMs x 64 lanes x sm_max clock, reached to 99 %
                     # by profiles/micro/fp64_peak.cu
                     "fp64": {"thread_instr_per_ray": FP64_INSTR_PER_RAY,
                              "achieved_Tinstr_s": FP64_INSTR_PER_RAY * n / (trace_ms * 1e-3) / 1e12,
                              "peak_Tinstr_s": 148 * 64 * 1.965e9 / 1e12,
                              "frac": FP64_INSTR_PER_RAY * n / (trace_ms * 1e-3) / (148 * 64 * 1.965e9)}},
        "clocks": clocks,
        "gpu_launches": launches,
    }
    if e2e is not None:
        line["e2e"] = e2e
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline()
    emit(line)
    if world > 1:
        td.destroy_process_group()


def run_e2e(args, pxf, pdist, src, prog, n, world, dev, barrier):
    """Same step through ``pxf_host_trace_program``: the bundle lives in pinned HOST memory,
    every step uploads the rows the chain reads, runs the fused kernel chunk by chunk, downloads
    every row it writes back into the host arrays (the f2py in-place contract) and returns the
    HPD.  For N>1 each rank keeps its final x,y on the device and the HPD is the global one."""
    import torch
    import torch.distributed as td
    rank = int(os.environ.get("RANK", "0"))
    try:
        host = [None] + [torch.empty(n, dtype=torch.float64).pin_memory() for _ in range(9)]
    except RuntimeError as e:
        return {"value": None, "unit": "rays/s", "error": "pinned allocation failed: %s" % str(e)[:80]}
    pristine = [None] + [src[k].cpu() for k in range(1, 7)]
    keep = [torch.empty(n, dtype=torch.float64, device=dev) for _ in range(2)] if world > 1 else None
    rays10 = [None] + [None] * 9

    def reset():
        for k in range(1, 7):
            host[k].copy_(pristine[k])

    def one():
        if world > 1:
            pxf.host.trace(host, prog, write_back=True, keep_xy=keep)
            rays10[1], rays10[2] = keep
            for k in range(3, 10):
                rays10[k] = keep[0]          # placeholders; dist.hpd reads rows 1,2 only
            return pdist.hpd(rays10)
        return pxf.host.trace(host, prog, write_back=True, hpd=True)["hpd"]

    reset()
    one()                                             # warm-up (allocator, page tables)
    t = 0.
    hp = None
    steps = max(1, args.e2e_steps)
    for _ in range(steps):
        reset()
        barrier()
        t0 = time.perf_counter()
        hp = one()
        torch.cuda.synchronize()
        t += time.perf_counter() - t0
    if world > 1:
        tt = torch.tensor([t], dtype=torch.float64, device=dev)
        td.all_reduce(tt, op=td.ReduceOp.MAX)
        t = float(tt[0])
    return {"value": n * world * steps / t, "unit": "rays/s", "steps": steps,
            # x, y uploaded; z, l, m, n of the subannulus source are bitwise constant and are found so by
            # the host-side chunk scan (filled on the device instead of uploaded)
            "h2d_bytes_per_step": 16 * n * world, "d2h_bytes_per_step": 40 * n * world + 8,
            "host_scanned_bytes_per_step": 32 * n * world, "host_filled_bytes_per_step": 24 * n * world,
            "path": "pxf_host_trace_program: pinned host rows -> chunked H2D / fused kernel / D2H on 3 streams "
                    "-> all nine rows mutated in place (constant input chunks are detected by a host scan and "
                    "not uploaded; x,y,l,m,n downloaded; z=0 and the normal (0,0,1) left by flat are filled by "
                    "host threads instead of crossing PCIe) + HPD", "hpd": hp}


def emit(line):
    """The one JSON line, on the real stdout."""
    sys.stdout.flush()
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


if __name__ == "__main__":
    a = parse()
    # rank 0 prints exactly ONE line on stdout: anything a library writes to fd 1 meanwhile (NCCL prints its
    # version banner there when NCCL_DEBUG is set) goes to stderr instead
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    if a.impl == "reference":
        run_reference(a)
    else:
        run_engine(a)
