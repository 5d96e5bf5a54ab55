#!/usr/bin/env python
"""bench.py -- propagator steps/s of Parament_equiprop on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config C2|C3|C4|C5|C1] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
           bench.py --gpus N --steps K --warmup W

One bench "step" = one pass of the hot path over one batch of synthetic input: the whole pulse of the
configuration (default C2 = BASELINE.json configs[1]: dim 16, 2 controls, 1e6 points, complex64, SIMPSON)
is propagated to one dim x dim propagator.  Unit of `value`: effective time steps per second.

N > 1 (one process per GPU, torch.distributed/NCCL for the plumbing): WEAK scaling along the time axis --
the pulse is N times longer, rank r owns the r-th contiguous slice of 1e6 points, reduces it to a partial
propagator on its GPU, the partials are all-gathered over NCCL and multiplied in order on rank 0
(Parament_combine).  value = (steps of all ranks) / (max-over-ranks device time).

Printed JSON (rank 0, one line): the driver contract plus
  roofline      FP64 tensor pipe (DMMA) is the binding unit of every kernel on this path; `peak` is measured in
                this process with the library's own microbenchmark (MEASURED_PEAKS.json has no FP64 figure),
                `achieved` uses the ALGORITHMIC flops F_step = 8 n^3 M_ref + 8 n^2 A' of SURVEY.md 8(d);
                executed flops (the degree actually evaluated) and the HBM figures are reported beside it.
  cpu_baseline  the scipy.linalg.expm product oracle (oracle/equiprop_oracle.py) timed on this box's cores on a
                bounded sample of the same workload (rank 0, N = 1 only).
  e2e           the same metric through the host-pointer C-ABI call (pinned host buffers, H2D + D2H inside).

--impl reference: the reference's own CUDA build (oracle/_ref/libparament.so, compiled from /root/reference by
oracle/build_ref.sh) driven through its C API on the same workload; if that library is missing the CPU oracle
port is timed instead.  Rank 0 only.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

QUAD = {"none": 0, "midpoint": 0x01000000, "simpson": 0x02000000}


# ----------------------------------------------------------------------------------------------------------
# helpers
# ----------------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                                          "-i", str(self.gpu), "-lms", "100"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [x.strip() for x in line.split(",")]))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [r for t, r in self.rows if t0 - 0.05 <= t <= t1 + 0.15 and len(r) >= 9] or [r for _, r in self.rows if len(r) >= 9]
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm = sorted(float(r[1]) for r in rows)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in rows for n, v in zip(names, r[5:9]) if v.lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(rows[0][2]), "power_w_max": max(float(r[3]) for r in rows),
                "samples": len(rows), "reasons": reasons}


def dist_env():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


def make_slices(name, world, rank):
    """Workload of this rank: slice `rank` of a pulse `world` times longer than the configuration's.  Ensembles
    (C5) shard pulses instead.  Every rank builds its slice from the same seeded generator."""
    from workloads import make_workload, smooth_pulses
    w = make_workload(name)
    if world > 1 and w.batch == 1:
        rng = np.random.default_rng(20260000 + 100 * rank + 7)
        w.carr = smooth_pulses(rng, w.amps, w.pts, dtype=np.float32 if w.precision == "fp32" else np.float64).astype(w.ctype)
    elif world > 1:
        rng = np.random.default_rng(20260000 + 100 * rank + 7)
        w.carr = smooth_pulses(rng, w.amps, w.pts, batch=w.batch, dtype=np.float32).astype(w.ctype)
    return w


def algorithmic_flops_per_step(w, M, real_products=4):
    """8 n^3 real flops per complex product (four real products); `real_products` = 3 where the kernel forms a complex
    product from three real ones (the batched GEMM of dim > 64): 6 n^3 executed."""
    nterms = w.amps + (w.amps + w.amps * (w.amps - 1) // 2 if w.use_magnus else 0)
    return 2.0 * real_products * w.dim ** 3 * M + 8.0 * w.dim ** 2 * nterms


def load_measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f), "measured"
    except OSError:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback"


# ----------------------------------------------------------------------------------------------------------
# our implementation
# ----------------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    rank, local_rank, world = dist_env()
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torch.distributed.run --nproc-per-node N for --gpus N > 1")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    os.environ["PARAMENT_DEVICE"] = str(local_rank)
    import parament_b200 as pb
    from parament_b200 import constants as K
    lib = pb._lib.lib

    w = make_slices(args.config, world, rank)
    n, fp64 = w.dim, w.precision == "fp64"
    tdt = torch.complex128 if fp64 else torch.complex64
    ctx = pb.Parament(w.precision, device=local_rank)
    ctx.set_hamiltonian(w.H0, *w.H1, use_magnus=w.use_magnus, quadrature_mode=w.quadrature)
    steps_rank = w.total_steps
    carr_np = np.ascontiguousarray(w.carr.reshape(w.batch, w.amps, w.pts))

    # Inputs larger than L2 (126 MB): rotate over enough distinct device copies of the amplitude stream.
    nbuf = max(2, int(np.ceil(160e6 / carr_np.nbytes)) + 1) if carr_np.nbytes < 160e6 else 2
    nbuf = min(nbuf, 12)
    rng = np.random.default_rng(1000 + rank)
    dev_in = []
    for i in range(nbuf):
        shift = carr_np if i == 0 else np.roll(carr_np, 17 * i, axis=-1)
        dev_in.append(torch.from_numpy(np.ascontiguousarray(shift)).cuda())
    flush = None
    if nbuf * carr_np.nbytes < 130e6:     # tiny inputs (C1): flush L2 explicitly between iterations
        flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    dev_out = torch.zeros(w.batch, n, n, dtype=tdt, device="cuda")
    gather = [torch.zeros(n, n, dtype=tdt, device="cuda") for _ in range(world)] if world > 1 and w.batch == 1 else None
    combined = torch.zeros(n, n, dtype=tdt, device="cuda")
    stream = torch.cuda.Stream()             # a real (non-default) stream: the library enqueues on it without synchronising
    torch.cuda.set_stream(stream)

    peak_dmma = lib.Parament_measurePeak(K.PEAK_DMMA)     # roofline denominator, measured on this device now
    peak_ffma = lib.Parament_measurePeak(K.PEAK_FFMA)
    launches = [0]

    def device_step(i):
        ctx.equiprop_device(w.dt, dev_in[i % nbuf].data_ptr(), w.pts, w.amps, dev_out.data_ptr(), batch=w.batch,
                            stream=stream.cuda_stream)
        launches[0] += int(ctx.stat(K.STAT_LAUNCHES))
        if gather is not None:
            dist.all_gather(gather, dev_out[0])                    # dim^2 per rank over NCCL / NVLink
            if rank == 0:
                parts = torch.stack(gather)                        # world x n x n on the device, slice order
                ctx.combine_device(parts.data_ptr(), world, combined.data_ptr(), stream=stream.cuda_stream)
                launches[0] += int(ctx.stat(K.STAT_LAUNCHES))
                return combined
        return None

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(args.warmup):
        device_step(i)
    barrier()
    launches[0] = 0
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    # ---- timed region 1: inputs resident in HBM, CUDA events on the launching stream ----
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    t0 = time.time()
    for i in range(args.steps):
        if flush is not None:
            flush.fill_(i & 0xFF)
        evs[i][0].record(stream)
        device_step(args.warmup + i)
        evs[i][1].record(stream)
    barrier()
    t1 = time.time()
    dev_ms = sum(a.elapsed_time(b) for a, b in evs)
    # duration of the dominant kernel launch: at N = 1 the timed region holds nothing but the chain kernel and its
    # small ordered reduction, so the average per-step device time is the launch duration
    kernel_ms = dev_ms / args.steps if world == 1 else ctx.stat(K.STAT_DEVICE_MS)
    M_used, M_ref = int(ctx.stat(K.STAT_DEGREE_USED)), int(ctx.stat(K.STAT_DEGREE_REFERENCE))
    products, horner, family = ctx.stat(K.STAT_PRODUCTS), int(ctx.stat(K.STAT_HORNER)), int(ctx.stat(K.STAT_FAMILY))
    real_products = int(ctx.stat(K.STAT_REAL_PRODUCTS))
    gpu_launches = launches[0]

    # ---- timed region 2: end to end through the host-pointer C-ABI, pinned host buffers ----
    pinned = torch.from_numpy(carr_np).pin_memory()
    host_view = pinned.numpy()
    out_host = np.zeros((w.batch, n, n), dtype=w.ctype)
    fn = getattr(lib, "Parament_equipropBatch" + ("_fp64" if fp64 else ""))

    def host_step():
        ec = fn(ctx._handle, host_view.reshape(-1), float(w.dt), w.pts, w.amps, w.batch, out_host.reshape(-1))
        assert ec == 0, ec
        if gather is not None:
            part = torch.from_numpy(out_host[0]).cuda()
            dist.all_gather(gather, part)
            if rank == 0:
                return ctx.combine(torch.stack(gather).cpu().numpy())
        return out_host

    e2e_steps = max(3, min(args.steps, 20))
    for _ in range(2):
        host_step()
    barrier()
    te0 = time.time()
    for _ in range(e2e_steps):
        host_step()
    barrier()
    e2e_s = time.time() - te0
    clocks = sampler.stop(t0, t1) if rank == 0 else None

    # ---- max over ranks ----
    t = torch.tensor([dev_ms, e2e_s * 1e3], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, e2e_ms = t.tolist()
    total_steps = steps_rank * world
    value = total_steps * args.steps / (dev_ms * 1e-3)
    e2e_value = total_steps * e2e_steps / (e2e_ms * 1e-3)

    line = None
    if rank == 0:
        peaks, how = load_measured_peaks()
        F_alg = algorithmic_flops_per_step(w, M_ref)
        F_exe = algorithmic_flops_per_step(w, products, real_products)     # products actually executed per step
        per_gpu_rate = steps_rank / (kernel_ms * 1e-3)         # last equiprop on rank 0, kernels only
        achieved = F_alg * per_gpu_rate * 1e-12
        traffic = None
        try:
            with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
                traffic = json.load(f).get(args.config)
        except OSError:
            pass
        in_bytes = carr_np.nbytes
        line = {
            "metric": "propagator steps/sec", "value": value, "unit": "steps/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"{args.config}: {w.description}", "dim": n, "controls": w.amps, "points_per_gpu": w.pts,
                       "pulses_per_gpu": w.batch, "effective_steps_per_gpu": steps_rank, "quadrature": w.quadrature,
                       "magnus": w.use_magnus, "io_precision": "complex64" if not fp64 else "complex128",
                       "x_Hnorm_h": w.meta["x"], "degree_reference": M_ref, "degree_used": M_used,
                       "series_evaluation": {0: "Clenshaw recurrence", 1: "Horner in Y^2 (same polynomial)",
                                             2: "Paterson-Stockmeyer blocks of four (same polynomial)",
                                             3: "degree 8 in three matrix products (Sastre 2018)",
                                             4: "degree 12 in four matrix products (Sastre 2018)"}.get(horner, str(horner)),
                       "matrix_products_per_step": products, "real_products_per_complex_product": real_products, "kernel_family": family,
                       "l2": f"inputs rotate over {nbuf} device copies ({nbuf * in_bytes / 1e6:.0f} MB > 126 MB L2)" if flush is None
                             else "L2 flushed by a 256 MiB write between iterations",
                       "parallelism": f"time axis sliced over {world} GPU(s), ordered NCCL all-gather + combine" if w.batch == 1
                                      else f"independent pulses sharded over {world} GPU(s), no communication"},
            "clocks": clocks,
            "gpu_launches": gpu_launches,
            "e2e": {"value": e2e_value, "unit": "steps/s", "h2d_bytes_per_step": in_bytes * world,
                    "d2h_bytes_per_step": out_host.nbytes * world, "ms_per_step": e2e_ms / e2e_steps, "timed_steps": e2e_steps,
                    "api": "Parament_equipropBatch (host pointers, pinned)"},
            # `achieved` = algorithmic flops of the reference's recurrence / time.  When the Horner-in-Y^2
            # evaluation executes fewer products than that recurrence, `frac` is computed from the EXECUTED flops so that
            # it stays a pipe utilisation (<= 1); the algorithmic figure is kept in `algorithmic_frac`.
            "roofline": {"bound": "tensor", "pipe": "FP64 tensor pipe (DMMA.8x8x4)", "achieved": achieved, "peak": peak_dmma, "unit": "TFLOP/s",
                         "frac": min(achieved, F_exe * per_gpu_rate * 1e-12) / peak_dmma if peak_dmma > 0 else None,
                         "algorithmic_frac": achieved / peak_dmma if peak_dmma > 0 else None, "traffic": traffic,
                         "peak_source": "Parament_measurePeak(DMMA mma.sync.m8n8k4.f64) in this process; MEASURED_PEAKS.json has no FP64 figure",
                         "kernel": {1: "k1_chain_kernel", 2: "k4_onchip_kernel (k4_chain_kernel when the shared-memory-resident variant does not fit)", 3: "k4_zgemm_kernel"}[family],
                         "kernel_ms_per_launch": kernel_ms,
                         "flops_per_step_algorithmic": F_alg, "flops_per_step_executed": F_exe,
                         "executed_tflops": F_exe * per_gpu_rate * 1e-12,
                         "executed_frac": F_exe * per_gpu_rate * 1e-12 / peak_dmma if peak_dmma > 0 else None,
                         "fp32_ffma_peak_tflops": peak_ffma,
                         "hbm": {"achieved_gbs": in_bytes / (kernel_ms * 1e-3) * 1e-9, "peak_gbs": peaks.get("hbm_gbs"),
                                 "peak_source": how, "note": "algorithmic HBM input is the amplitude stream only; not the limiter"}},
        }
        if world == 1:
            line["cpu_baseline"] = cpu_baseline(w)
    ctx.destroy()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if line is not None:
        print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------------------
# CPU oracle baseline and the reference arm  (the only places that execute oracle/)
# ----------------------------------------------------------------------------------------------------------
def _oracle_pulse(args):
    from oracle.equiprop_oracle import equiprop_oracle, _single_thread_blas
    H0, H1, carr, dt, quad, mag, prec = args
    with _single_thread_blas():
        return equiprop_oracle(H0, H1, carr, dt, quad, mag, prec)


def cpu_baseline(w, budget_s=12.0):
    from oracle.equiprop_oracle import equiprop_oracle
    cores = os.cpu_count() or 1
    workers = min(cores, 32)
    if w.batch > 1:
        # ensemble: whole pulses, one per worker process at a time
        from concurrent.futures import ProcessPoolExecutor
        t = time.time()
        _oracle_pulse((w.H0, w.H1, w.carr[0], w.dt, w.quadrature, w.use_magnus, w.precision))
        per_pulse = time.time() - t
        npulses = int(max(workers, min(w.batch, budget_s / per_pulse * workers * 0.7)))
        jobs = [(w.H0, w.H1, w.carr[b], w.dt, w.quadrature, w.use_magnus, w.precision) for b in range(npulses)]
        t = time.time()
        with ProcessPoolExecutor(max_workers=workers) as ex:
            list(ex.map(_oracle_pulse, jobs, chunksize=max(1, npulses // (4 * workers))))
        dt = time.time() - t
        return {"value": npulses * w.steps / dt, "unit": "steps/s", "cores": workers, "kind": "port",
                "sample": f"first {npulses} pulses of the ensemble ({npulses * w.steps} effective steps), float64 scipy.linalg.expm + "
                          f"ordered product, {workers} processes x 1 BLAS thread, {dt:.1f} s"}
    carr = w.carr
    per_step_pts = 2 if (w.use_magnus or w.quadrature == "simpson") else 1
    # probe, then size the sample for ~budget_s of CPU work
    probe_steps = max(8, min(w.steps, {2: 4000, 8: 4000, 16: 2000, 64: 200, 256: 16}.get(w.dim, 100)))
    t = time.time()
    equiprop_oracle(w.H0, w.H1, carr[:, :per_step_pts * probe_steps + 1], w.dt, w.quadrature, w.use_magnus, w.precision, workers=1)
    per_step = (time.time() - t) / probe_steps
    sample = int(max(probe_steps, min(w.steps, budget_s / per_step * workers * 0.7)))
    t = time.time()
    equiprop_oracle(w.H0, w.H1, carr[:, :per_step_pts * sample + 1], w.dt, w.quadrature, w.use_magnus, w.precision, workers=workers)
    dt = time.time() - t
    return {"value": sample / dt, "unit": "steps/s", "cores": workers, "kind": "port",
            "sample": f"first {sample} effective steps of the same pulse, float64 scipy.linalg.expm + ordered product, "
                      f"{workers} processes x 1 BLAS thread, {dt:.1f} s"}


def run_reference(args):
    rank, local_rank, world = dist_env()
    if rank != 0:
        return
    from workloads import make_workload
    w = make_workload(args.config)
    n, fp64 = w.dim, w.precision == "fp64"
    sfx = "_fp64" if fp64 else ""
    base = {"impl": "reference", "metric": "propagator steps/sec", "unit": "steps/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "c64" if not fp64 else "c128",
            "data": "synthetic", "config": {"workload": f"{args.config}: {w.description}"}}
    libpath = os.path.join(ROOT, "oracle", "_ref", "libparament.so")
    ref_ok = os.path.exists(libpath) and w.dim ** 2 * w.pts < 2 ** 31 and 3 * w.dim ** 2 * w.pts * (16 if fp64 else 8) < 150e9
    if ref_ok:
        try:
            lib = ctypes.cdll.LoadLibrary(libpath)
            h = ctypes.c_void_p()
            ref_ok = getattr(lib, "Parament_create" + sfx)(ctypes.byref(h)) == 0
        except OSError:
            ref_ok = False
    if not ref_ok:
        # no usable reference CUDA build for this configuration: time the CPU oracle port on a bounded sample
        cb = cpu_baseline(w)
        base.update({"value": cb["value"], "ms_per_step": None, "cpu_baseline": cb,
                     "e2e": {"value": cb["value"], "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                     "note": "reference CUDA build unavailable or cannot hold this configuration (SURVEY.md section 6); CPU oracle port timed instead"})
        print(json.dumps(base), flush=True)
        return
    vp = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    H0 = np.ascontiguousarray(w.H0.ravel())
    H1 = np.ascontiguousarray(w.H1.ravel())
    assert getattr(lib, "Parament_setHamiltonian" + sfx)(h, vp(H0), vp(H1), ctypes.c_uint(n), ctypes.c_uint(w.amps),
                                                         ctypes.c_bool(w.use_magnus), ctypes.c_int(QUAD[w.quadrature])) == 0
    pulses = w.carr.reshape(w.batch, w.amps, w.pts)
    nb = min(w.batch, 500)                  # ensembles: the reference has no batch entry point -> sequential calls, bounded sample
    out = np.zeros(n * n, dtype=w.ctype)
    eq = getattr(lib, "Parament_equiprop" + sfx)

    def step():
        for b in range(nb):
            c = np.ascontiguousarray(pulses[b].ravel())
            assert eq(h, vp(c), ctypes.c_double(w.dt), ctypes.c_uint(w.pts), ctypes.c_uint(w.amps), vp(out)) == 0

    for _ in range(max(1, args.warmup)):
        step()
    t = time.time()
    for _ in range(args.steps):
        step()
    dt = time.time() - t
    getattr(lib, "Parament_destroy" + sfx)(h)
    value = w.steps * nb * args.steps / dt
    base.update({"value": value, "ms_per_step": dt / args.steps * 1e3,
                 "cpu_baseline": {"value": value, "unit": "steps/s", "cores": 1, "kind": "reference",
                                  "sample": f"reference CUDA build (cuBLAS batched, compiled from /root/reference for sm_100) on GPU 0 through "
                                            f"its C API, {nb} pulse(s) of {w.pts} points per step; the reference has no CPU implementation"},
                 "e2e": {"value": value, "unit": "steps/s", "h2d_bytes_per_step": int(pulses[:nb].nbytes), "d2h_bytes_per_step": int(out.nbytes * nb)}})
    print(json.dumps(base), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--config", default="C2", choices=["C1", "C2", "C3", "C4", "C5"])
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
