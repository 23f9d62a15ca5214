"""bench.py -- the headline measurement: 1024^3 fp64 r2c + c2r pair (PHYSICAL_IN_Z) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W]            # our arm (CUDA library through the C ABI)
    python bench.py --impl reference [--gpus N] [--steps K] ...     # CPU arm: the oracle port of the reference's generic backend
    python bench.py --config 512x | 2048f32 ...                     # the other multi-GPU configs of BASELINE.json (not the driver's line)

A "step" is one forward (r2c) + one backward (c2r) 3-D transform of the whole field, i.e. one
iteration of the timed loop of examples/fft_physical_z/fft_r2c_z.f90:96-128 of the reference.
N > 1 runs under torchrun (one rank per GPU) on the pencil grids 1x2 / 2x2 / 2x4 with the
TOTAL field fixed ("strong" scaling), as BASELINE.json asks.

value   = GFLOP/s of the whole job with the reference's own convention (5 N log2 N per 3-D c2c,
          examples/fft_physical_x/fft_c2c_x.f90:159-166; an r2c + c2r pair counts as one c2c),
          device-timed (CUDA events on the library's stream), max over ranks, inputs resident in HBM.
parity  = before anything is timed, the examples' ramp field (i/nx)(j/ny)(k/nz) goes through r2c at the FULL size on
          the grid being measured and every rank compares its spectrum pencil with the closed form
          R_n[0] = (n+1)/2, R_n[k] = 1/(exp(-2 pi i k/n) - 1) (product over the axes); the run refuses to print a line
          when max|delta|/max|ref| exceeds the tolerance (1e-12 fp64, 1e-5 fp32).  Then the round trip of the random
          field is checked the same way.
e2e     = the same pair through the host-array entry points d2d_fft_3d_r2c_host / _c2r_host: pinned host
          buffers, H2D of the input and D2H of the result inside the timed region of every step (wall clock between
          device-synchronised points, max over ranks).  The context is in stream-ordered mode, so the library pipelines
          the PCIe traffic of consecutive calls (upload / transform / download streams).
roofline= the dominant FFT kernel's algorithmic bytes / its average CUDA-event duration inside the
          timed region, against MEASURED_PEAKS.json's hbm_gbs.
cpu_baseline = the oracle (C restatement of the reference's generic-backend CPU path, OpenMP over
          the simulated MPI ranks) on a bounded sample, host cores of this box.  rank 0, N = 1 only.
"""
import argparse
import json
import math
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GRIDS = {1: (1, 1), 2: (1, 2), 4: (2, 2), 8: (2, 4)}
METRIC = "fft3d_r2c_c2r_pair_gflops"
UNIT = "GFLOP/s"
# BASELINE.json configs that bench.py can time: name -> (cube edge, format, precision, BASELINE index)
CONFIGS = {
    "1024z": (1024, "Z", "f64", 3),
    "512x": (512, "X", "f64", 2),
    "2048f32": (2048, "X", "f32", 4),
}
REF_SAMPLE_N = 512  # cube edge of the bounded sample the CPU arm runs, for every N (GFLOP/s is size-normalised)


def pair_flops(nx, ny, nz):
    n = float(nx) * ny * nz
    return 5.0 * n * math.log2(n)


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (NVML, 100 ms period)."""

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        names = {
            "hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
            "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
            "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
            "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4),
            "hw_power_brake": getattr(nv, "nvmlClocksThrottleReasonHwPowerBrakeSlowdown", 0x80),
        }
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            self._stop.wait(0.1)

    def start(self):
        if self.nv is not None:
            self._stop.clear()
            self._thr = threading.Thread(target=self._run, daemon=True)
            self._thr.start()

    def stop(self):
        if self._thr is not None:
            self._stop.set()
            self._thr.join()
            self._thr = None

    def summary(self):
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


def oracle_mod():
    p = os.path.join(ROOT, "oracle")
    if p not in sys.path:
        sys.path.insert(0, p)
    import oracle as orc  # CPU checker / baseline only
    return orc


def cpu_pair_seconds(orc, n, grid, reps, fmt=None):
    """Time `reps` r2c + c2r pairs of an n^3 fp64 field with the oracle (all rank-threads)."""
    import numpy as np
    fmt = orc.PHYSICAL_IN_Z if fmt is None else fmt
    shape = (n, n, n)
    g = np.asfortranarray(np.random.default_rng(20240601).uniform(-1, 1, shape))
    ins = orc.scatter(g, grid, 2 if fmt == orc.PHYSICAL_IN_Z else 0)
    del g
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        spec = orc.fft_3d_r2c_world(shape, grid, fmt, ins)
        orc.fft_3d_c2r_world(shape, grid, fmt, spec)
        ts.append(time.perf_counter() - t0)
    return ts


def cpu_setup():
    """The CPU arm's rank grid and OpenMP threads: one thread per simulated MPI rank, as many ranks as the box has cores
    (power of two).  torchrun exports OMP_NUM_THREADS=1 -- the count is set explicitly and read back."""
    cores = os.cpu_count() or 1
    c = 1
    while c * 2 <= min(cores, 64):
        c *= 2
    os.environ["OMP_NUM_THREADS"] = str(c)  # before the OpenMP runtime starts
    orc = oracle_mod()
    threads = orc.set_threads(c)
    return orc, orc.best_2d_grid(c), c, threads


def cpu_baseline(budget_s=20.0):
    """The oracle on a bounded sample of the workload (about 10-30 s of CPU work)."""
    orc, grid, cores, threads = cpu_setup()
    cpu_pair_seconds(orc, 128, grid, 1)            # warm-up (loads the library, starts the OpenMP team)
    t256 = min(cpu_pair_seconds(orc, 256, grid, 2))
    n = REF_SAMPLE_N if t256 * 9.0 * 2 < budget_s else 256
    reps = 2 if n == REF_SAMPLE_N else max(2, int(budget_s / 2 / max(t256, 1e-3)))
    ts = cpu_pair_seconds(orc, n, grid, reps) if n != 256 or reps > 2 else [t256]
    best = min(ts)
    return {"value": pair_flops(n, n, n) / best / 1e9, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"{n}^3 fp64 r2c+c2r pair (PHYSICAL_IN_Z), oracle port of the reference's generic backend, "
                      f"{grid[0]}x{grid[1]} rank-threads ({threads} OpenMP threads on {os.cpu_count()} cores), best of {len(ts)}; "
                      f"{best * 1e3:.0f} ms/pair",
            "ms_per_pair": best * 1e3, "n": n, "threads_used": threads}


def run_reference(args):
    """--impl reference: the reference's CPU algorithm (oracle port; the Fortran+MPI reference cannot be
    built in this image) on the host cores.  Each step = one pair on a bounded sample of the workload: REF_SAMPLE_N^3 for
    every N, so that the arm's value does not depend on how it was launched."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    n_full, fmtc, prec, _ = CONFIGS[args.config]
    orc, grid, cores, threads = cpu_setup()
    fmt = orc.PHYSICAL_IN_Z if fmtc == "Z" else orc.PHYSICAL_IN_X
    cpu_pair_seconds(orc, 128, grid, 1, fmt)
    n = REF_SAMPLE_N
    import numpy as np
    shape = (n, n, n)
    g = np.asfortranarray(np.random.default_rng(20240601).uniform(-1, 1, shape))
    ins = orc.scatter(g, grid, 2 if fmtc == "Z" else 0)
    del g
    for _ in range(args.warmup):
        spec = orc.fft_3d_r2c_world(shape, grid, fmt, ins)
        orc.fft_3d_c2r_world(shape, grid, fmt, spec)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        spec = orc.fft_3d_r2c_world(shape, grid, fmt, ins)
        orc.fft_3d_c2r_world(shape, grid, fmt, spec)
    dt = (time.perf_counter() - t0) / args.steps
    val = pair_flops(n, n, n) / dt / 1e9
    sample = (f"{n}^3 fp64 r2c+c2r pair per step (the {n_full}^3 workload is sampled at {n}^3: GFLOP/s is size-normalised), "
              f"{grid[0]}x{grid[1]} rank-threads, {threads} OpenMP threads")
    cfg = workload_config(args, GRIDS.get(args.gpus, (1, 1)))
    cfg["cpu_sample"] = {"nx": n, "ny": n, "nz": n, "rank_threads": f"{grid[0]}x{grid[1]}", "threads_used": threads,
                         "host_cores": os.cpu_count()}
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic uniform(-1,1), seed 20240601",
        "config": cfg,
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample, "threads_used": threads},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "reference is Fortran+MPI: no Fortran compiler / MPI in this image, so the CPU arm is the line-traceable C port "
                "of its generic backend (oracle/d2d_oracle.c), one OpenMP thread per simulated MPI rank",
    }
    print(json.dumps(line), flush=True)
    return 0


def workload_config(args, grid):
    n, fmtc, prec, idx = CONFIGS[args.config]
    ex = "fft_physical_z" if fmtc == "Z" else "fft_physical_x"
    return {"workload": f"examples/{ex} {n}^3 {'fp64' if prec == 'f64' else 'fp32'} r2c+c2r pair (BASELINE.json configs[{idx}]"
                        f"{', headline' if idx == 3 else ''})",
            "nx": n, "ny": n, "nz": n, "format": f"PHYSICAL_IN_{fmtc}", "p_row": grid[0], "p_col": grid[1],
            "l2": "working set >> L2 (each sweep streams >= 2 GiB per GPU at the headline size), no flush needed",
            "flops": "5*N*log2(N) per r2c+c2r pair, N = nx*ny*nz (examples/fft_physical_x/fft_c2c_x.f90:159-166)"}


def ramp_dft_1d(n, nk):
    """Closed-form DFT of the examples' ramp i/n, i = 1..n: R[0] = (n+1)/2, R[k] = 1/(exp(-2 pi i k/n) - 1)."""
    import numpy as np
    k = np.arange(nk, dtype=np.float64)
    r = np.empty(nk, dtype=np.complex128)
    r[0] = (n + 1) / 2.0
    w = np.exp(-2j * np.pi * k[1:] / n)
    r[1:] = 1.0 / (w - 1.0)
    return r


def ramp_forward_error(p, torch, eng, d2d, in_r, out_c, fmtc, n):
    """r2c of the ramp field at the full size on this grid; each rank compares ITS spectrum pencil with the closed form.
    Returns max|delta| / max|ref| over this rank's pencil (max|ref| is the global maximum, the DC bin)."""
    dev = in_r.device
    rdt = in_r.dtype
    ph, sp = eng.ph, eng.sp
    ist, isz = (ph.zst, ph.zsz) if fmtc == "Z" else (ph.xst, ph.xsz)      # input pencil (1-based starts)
    ost, osz = (sp.xst, sp.xsz) if fmtc == "Z" else (sp.zst, sp.zsz)      # output pencil
    ax = [(torch.arange(ist[d], ist[d] + isz[d], device=dev, dtype=torch.float64) / n) for d in range(3)]
    # in_r(i, j, k) = (i/nx)(j/ny)(k/nz), filled slab by slab along the slowest axis
    plane = (ax[0][:, None] * ax[1][None, :])
    for k in range(isz[2]):
        in_r[:, :, k] = (plane * ax[2][k]).to(rdt)
    eng.fft_3d(in_r, out_c)
    d2d.sync()
    nk = [n, n, n]
    nk[2 if fmtc == "Z" else 0] = n // 2 + 1
    R = [torch.from_numpy(ramp_dft_1d(n, nk[d])[ost[d] - 1: ost[d] - 1 + osz[d]]).to(dev) for d in range(3)]
    refmax = ((n + 1) / 2.0) ** 3
    rp = R[0][:, None] * R[1][None, :]
    worst = 0.0
    for k in range(osz[2]):
        ref = rp * R[2][k]
        worst = max(worst, float((out_c[:, :, k].to(torch.complex128) - ref).abs().max().item()) if ref.numel() else 0.0)
    return worst / refmax


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="1024z", choices=sorted(CONFIGS), help="BASELINE.json workload (default: the headline)")
    ap.add_argument("--n", type=int, default=0, help="override the cube edge (experiments)")
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    if args.n:
        c = CONFIGS[args.config]
        CONFIGS[args.config] = (args.n, c[1], c[2], c[3])
    if args.impl == "reference":
        return run_reference(args)

    import numpy as np
    import torch
    import torch.distributed as dist
    from __graft_entry__ import package

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: there is no CPU fallback for the product path")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if world != args.gpus and rank == 0:
        print(f"# note: --gpus {args.gpus} but WORLD_SIZE={world}; using {world}", file=sys.stderr)
    grid = GRIDS.get(world)
    p = package()
    if grid is None:
        grid = p.best_2d_grid(world)
    n, fmtc, prec, _ = CONFIGS[args.config]
    rdt, cdt = (torch.float64, torch.complex128) if prec == "f64" else (torch.float32, torch.complex64)
    rbytes_el = 8 if prec == "f64" else 4
    tol = 1e-12 if prec == "f64" else 1e-5
    if world > 1:
        d2d = p.decomp_2d_init_from_torch_distributed(n, n, n, *grid)
    else:
        d2d = p.decomp_2d_init(n, n, n, 1, 1)
    d2d.set_blocking(False)  # stream-ordered: the timed loop has no host synchronisation inside
    eng = p.decomp_2d_fft_init(p.PHYSICAL_IN_Z if fmtc == "Z" else p.PHYSICAL_IN_X, dtype=rdt)
    a_in, a_out = (d2d.alloc_z, d2d.alloc_x) if fmtc == "Z" else (d2d.alloc_x, d2d.alloc_z)
    in_r = a_in(rdt, eng.ph)
    out_c = a_out(cdt, eng.sp)
    back = a_in(rdt, eng.ph)
    lib_stream = torch.cuda.ExternalStream(d2d.stream(), device=in_r.device)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def allmax(x):
        if world == 1:
            return float(x)
        t = torch.tensor([float(x)], dtype=torch.float64, device=in_r.device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- parity gate 1: forward spectrum of the ramp field against its closed form, full size, this grid ----------
    torch.cuda.synchronize()
    fwd_err = allmax(ramp_forward_error(p, torch, eng, d2d, in_r, out_c, fmtc, n))
    if not fwd_err <= tol:
        raise SystemExit(f"forward spectrum of the ramp field differs from the closed form: {fwd_err:.3e} > {tol:.0e} "
                         f"(grid {grid}, world {world}): refusing to time a wrong transform")

    gen = torch.Generator(device=in_r.device)
    gen.manual_seed(20240601 + rank)
    in_r.uniform_(-1, 1, generator=gen)

    def step():
        eng.fft_3d(in_r, out_c)
        eng.fft_3d(out_c, back)

    torch.cuda.synchronize()
    for _ in range(args.warmup):
        step()
    d2d.sync()
    # parity gate 2: the round trip must reproduce the input (size-independent property)
    rt_tol = 1e-12 if prec == "f64" else 2e-5
    rt_err = allmax(float((back / float(n) ** 3 - in_r).abs().max().item()))
    if not rt_err < rt_tol:
        raise SystemExit(f"round trip error {rt_err}: refusing to time a wrong transform")

    sampler = ClockSampler(local)
    d2d.profile_reset()
    d2d.profile(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = d2d.launch_count()
    barrier()
    sampler.start()
    e0.record(lib_stream)
    for _ in range(args.steps):
        step()
    e1.record(lib_stream)
    d2d.sync()
    barrier()
    sampler.stop()
    d2d.profile(False)
    launches = d2d.launch_count() - launches0
    ms = e0.elapsed_time(e1) / args.steps
    prof = d2d.profile_read()
    if world > 1:
        ms = allmax(ms)
        lt = torch.tensor([launches], dtype=torch.int64, device=in_r.device)
        dist.all_reduce(lt, op=dist.ReduceOp.SUM)
        launches = int(lt.item())
    flops = pair_flops(n, n, n)
    value = flops / (ms * 1e-3) / 1e9

    # ---- roofline of the dominant kernel (rank 0's timers; bytes are this rank's share) -------------
    peak, peak_src = load_peaks()
    kern = {}
    for label, (tot, calls, by) in prof.items():
        if label.startswith("fft_") and label not in ("fft_r2c", "fft_c2r", "fft_c2c") and calls:
            kern[label] = {"ms": tot / calls, "GBps": by / calls / (tot / calls) / 1e6, "bytes": by / calls, "calls": calls,
                           "ms_per_step": tot / args.steps}
    comm = {}
    for label, (tot, calls, by) in prof.items():
        if not calls:
            continue
        # a2a_*: NCCL send/recv exchange; p2p_*: a producer kernel whose stores go to the peers (D2D_FUSED=1);
        # ce_*: copy-engine pushes of the pipelined chain -- one span per peer stream, from the first chunk to the last
        if label.startswith("a2a_") or (label.startswith("p2p_") and label[4:] in ("x_y", "y_x", "y_z", "z_y")):
            comm[label] = {"ms": tot / calls, "send_GBps": by / calls / (tot / calls) / 1e6, "send_bytes": by / calls}
        elif label.startswith("ce_"):
            span = tot / calls
            comm[label] = {"ms": span, "send_GBps": (by / args.steps) / span / 1e6, "send_bytes": by / args.steps,
                           "peer_streams": calls / args.steps}
    sync_ms = {k: prof[k][0] / args.steps for k in ("p2p_ready", "p2p_done", "pipe_wait") if k in prof}
    dom = max(kern, key=lambda k: kern[k]["ms"] * kern[k]["calls"]) if kern else None
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            tj = json.load(f)
        if args.config == "1024z":
            traffic = tj.get(f"{dom}@{world}", tj.get(dom) if world == 1 else None)
    except Exception:
        pass
    roofline = None
    if dom:
        launches_per_step = kern[dom]["calls"] / args.steps
        roofline = {"bound": "hbm", "kernel": dom, "achieved": kern[dom]["GBps"], "peak": peak, "unit": "GB/s",
                    "frac": kern[dom]["GBps"] / peak, "traffic": traffic, "peak_source": peak_src,
                    "bytes_per_launch": kern[dom]["bytes"], "ms_per_launch": kern[dom]["ms"],
                    "launches_per_step": launches_per_step,
                    "all_kernels": {k: {"ms": round(v["ms"], 4), "GBps": round(v["GBps"], 1), "frac": round(v["GBps"] / peak, 3),
                                        "ms_per_step": round(v["ms_per_step"], 4)}
                                    for k, v in kern.items()}}
        if sync_ms:
            roofline["flag_wait_ms_per_step"] = {k: round(v, 4) for k, v in sync_ms.items()}
        if comm:
            roofline["exchanges"] = {k: {"ms": round(v["ms"], 4), "send_GBps": round(v["send_GBps"], 1),
                                         "frac_of_900": round(v["send_GBps"] / 900.0, 3)} for k, v in comm.items()}

    # ---- e2e: host arrays through the C ABI (pinned host memory, copies inside the timed region) ----
    e2e = None
    if not args.no_e2e:
        del back
        torch.cuda.empty_cache()
        rbytes, cbytes = in_r.numel() * rbytes_el, out_c.numel() * 2 * rbytes_el
        h_in = torch.empty(in_r.numel(), dtype=rdt).pin_memory()
        h_spec = torch.empty(out_c.numel() * 2, dtype=rdt).pin_memory()
        h_back = torch.empty(in_r.numel(), dtype=rdt).pin_memory()
        h_in.copy_(in_r.permute(2, 1, 0).reshape(-1))
        torch.cuda.synchronize()

        def e2e_step():
            eng.fft_3d_r2c_host(h_in.data_ptr(), h_spec.data_ptr())
            eng.fft_3d_c2r_host(h_spec.data_ptr(), h_back.data_ptr())

        e2e_step()  # warm-up (allocates the device staging buffers)
        d2d.sync()
        d2d.profile_reset()
        d2d.profile(True)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            e2e_step()
        d2d.sync()  # every stream of the context: the last download is on the host
        wall_ms = (time.perf_counter() - t0) * 1e3 / args.e2e_steps
        barrier()
        d2d.profile(False)
        eprof = d2d.profile_read()
        pcie = {k: {"ms_per_call": round(eprof[k][0] / eprof[k][1], 2), "GBps": round(eprof[k][2] / eprof[k][0] / 1e6, 1)}
                for k in ("h2d", "d2h") if k in eprof and eprof[k][1] and eprof[k][0] > 0}
        ems = allmax(wall_ms)
        m = min(h_in.numel(), 1 << 24)  # a sample is enough for a guard (the full check ran on the device above)
        rt2 = float((h_back[:m] / float(n) ** 3 - h_in[:m]).abs().max().item())
        rt3 = float((h_back[-m:] / float(n) ** 3 - h_in[-m:]).abs().max().item())
        if not max(rt2, rt3) < rt_tol:
            raise SystemExit(f"e2e round trip error {max(rt2, rt3)}")
        e2e = {"value": flops / (ems * 1e-3) / 1e9, "unit": UNIT, "ms_per_step": ems,
               "h2d_bytes_per_step": int(rbytes + cbytes) * world, "d2h_bytes_per_step": int(cbytes + rbytes) * world,
               "api": "d2d_fft_3d_r2c_host + d2d_fft_3d_c2r_host (pinned host arrays; context in stream-ordered mode, "
                      "d2d_ctx_sync at the end of the timed loop)",
               "steps": args.e2e_steps, "timing": "wall clock between device-synchronised points, max over ranks",
               "round_trip_max_err": max(rt2, rt3), "pcie_spans": pcie}
        del h_in, h_spec, h_back

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cpu = cpu_baseline()

    p.decomp_2d_finalize()
    if rank == 0:
        rtot = float(n) ** 3 * rbytes_el
        ctot = float(n) * n * (n // 2 + 1) * 2 * rbytes_el
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": prec,
            "data": "synthetic uniform(-1,1), seed 20240601+rank, generated on device",
            "config": workload_config(args, grid),
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches),
            "clocks": sampler.summary(), "round_trip_max_err": rt_err, "forward_max_rel_err": fwd_err,
            "parity": f"ramp field r2c vs closed form at {n}^3 on {grid[0]}x{grid[1]}: {fwd_err:.2e} (tol {tol:.0e}); "
                      f"round trip {rt_err:.2e}",
            "hbm_floor_ms": (2 * rtot + 10 * ctot) / world / peak / 1e6,
            "exchange": None if world == 1 else (
                "copy-engine pushes over peer memory, chunk-pipelined with the stages" if any(k.startswith("ce_") for k in prof) and os.environ.get("D2D_PUSH") != "sm"
                else "exchange kernel (TMA pushes) over peer memory, chunk-pipelined with the stages" if any(k.startswith("ce_") for k in prof)
                else "fused peer stores of the producer kernels over peer memory" if any(k.startswith("p2p_") and k[4:] in ("x_y", "y_x", "y_z", "z_y") for k in prof)
                else "NCCL grouped send/recv"),
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
