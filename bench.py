#!/usr/bin/env python
"""bench.py -- propagator steps/s of Parament_equiprop on B200 (BASELINE.json metric: dim 16 / 64 / 256, 1-8 GPUs).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config C2] [--configs C3,C4,...|all|none] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
           bench.py --gpus N --steps K --warmup W

One bench "step" = one pass of the hot path over one batch of synthetic input: the whole pulse (or ensemble) of a
configuration is propagated to its dim x dim propagator(s).  Unit of `value`: effective time steps per second.

ONE JSON line (rank 0).  Top level = `--config` (default C2 = BASELINE.json configs[1]: dim 16, 2 controls, 1e6 points,
complex64, SIMPSON), timed over exactly K steps after W warm-ups.  `configs` holds the same record for the other
BASELINE configurations measured in the same run -- C3 (dim 64), C4 (dim 256), C5 (ensemble), C1, and two variants of C2
the north_star names (Magnus commutators on; complex amplitudes) -- each with its own (smaller, stated) step count so the
whole run stays within minutes.

N > 1 (one process per GPU; torch.distributed / NCCL for the plumbing):
  top level   C2, WEAK scaling along the time axis: the pulse is N times longer, rank r owns the r-th slice of 1e6 points.
  configs     STRONG scaling of the BASELINE multi-GPU configurations: C4 (and C3, C2) -- the SAME pulse cut into N contiguous
              time slices, each rank reduces its slice to a partial propagator, the partials are all-gathered over NCCL and
              multiplied in order on rank 0 (Parament_combineDevice); C5 -- the same 1e4 pulses sharded over the ranks, no
              communication.  value = steps of the whole job / max-over-ranks time.

Every record carries
  value / ms_per_step   inputs resident in HBM, CUDA events on the launching stream, max over ranks
  e2e                   the same work through the host-pointer C-ABI call from PAGEABLE numpy arrays (what the unchanged
                        pyparament wrapper passes, parament.py:263-272), H2D + D2H inside the timed region; `pinned` = the
                        same from page-locked buffers; `wrapper` (N = 1) = through the UNCHANGED reference wrapper class
                        (oracle/_ref/pyparament, PARAMENT_LIB_DIR = this library), Python-side casts and copies included
  roofline              FP64 tensor pipe (DMMA) is the binding unit of every kernel on this path; `peak` is measured in this
                        process (MEASURED_PEAKS.json has no FP64 figure); `achieved` uses the ALGORITHMIC flops
                        F_step = 8 n^3 M_ref + 8 n^2 A' of SURVEY.md 8(d), `frac` the EXECUTED flops (<= 1)
  cpu_baseline          (N = 1) the scipy.linalg.expm product oracle timed on this box's cores on a bounded sample

--impl reference: the reference's own CUDA build (oracle/_ref/libparament.so, compiled from /root/reference by
oracle/build_ref.sh) driven through its C API from the same kind of host buffers, on the configurations it can hold
(C2, C5 as sequential calls, C1, the C2 variants; C3 / C4 need 196 / 315 GB); if that library is missing the CPU oracle
port is timed instead.  Rank 0 only.  This arm never imports parament_b200.
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
SERIES_NAMES = {0: "Clenshaw recurrence", 1: "Horner in Y^2 (same polynomial)",
                2: "Paterson-Stockmeyer blocks of four (same polynomial)",
                3: "degree 8 in three matrix products (Sastre 2018)",
                4: "degree 12 in four matrix products (Sastre 2018)"}
KERNEL_NAMES = {1: "k1_chain_kernel", 2: "k4_onchip_kernel", 3: "k4_zgemm_kernel"}
# timed steps of the sub-records (the top level always uses --steps): bounded so that the default run finishes in minutes
SUB_STEPS = {"C1": 20, "C2": 20, "C2_magnus": 10, "C2_complex": 10, "C3": 5, "C4": 3, "C5": 10}
SUB_E2E = {"C1": 20, "C2": 10, "C2_magnus": 5, "C2_complex": 5, "C3": 3, "C4": 2, "C5": 10}
DEFAULT_SUBS_1 = ["C3", "C4", "C5", "C1", "C2_magnus", "C2_complex"]
DEFAULT_SUBS_N = ["C4", "C3", "C2", "C5"]


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

    def window(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        rows = [r for t, r in self.rows if t0 - 0.05 <= t <= t1 + 0.15 and len(r) >= 9] or [r for _, r in self.rows if len(r) >= 9]
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm = sorted(float(r[1]) for r in rows)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in rows for n, v in zip(names, r[5:9]) if v.lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(rows[0][2]), "power_w_max": max(float(r[3]) for r in rows),
                "samples": len(rows), "reasons": reasons}

    def stop(self):
        if self.proc is not None:
            time.sleep(0.15)
            self.proc.terminate()


def dist_env():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


def nterms_of(w):
    return w.amps + (w.amps + w.amps * (w.amps - 1) // 2 if w.use_magnus else 0)


def flops_per_step(w, products, real_products=4):
    """`products` complex n x n products of `real_products` real ones each (2 n^3 flops per real product) + the assembly."""
    return 2.0 * real_products * w.dim ** 3 * products + 8.0 * w.dim ** 2 * nterms_of(w)


def load_measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f), "measured"
    except OSError:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback"


def workload_config(name, w, world, mode, l2):
    """Workload description shared verbatim by both arms (`config` of the JSON line)."""
    if w.batch > 1:
        par = f"{w.batch} independent pulses sharded over {world} GPU(s), no communication"
    elif mode == "weak":
        par = f"pulse {world}x longer, one contiguous slice of {w.pts} points per GPU, ordered NCCL all-gather + combine"
    else:
        par = f"time axis of the one pulse cut into {world} contiguous slice(s), ordered NCCL all-gather + combine"
    return {"workload": f"{name}: {w.description}", "dim": w.dim, "controls": w.amps, "points": w.pts, "pulses": w.batch,
            "effective_steps_per_pulse": w.steps, "quadrature": w.quadrature, "magnus": w.use_magnus,
            "io_precision": "complex64" if w.precision == "fp32" else "complex128", "x_Hnorm_h": w.meta["x"],
            "l2": l2, "parallelism": par}


def pageable_copy(a):
    """A fresh pageable numpy array (what np.ascontiguousarray(...).astype(...) in the unchanged wrapper produces)."""
    b = np.empty_like(a)
    np.copyto(b, a)
    return b


def reference_wrapper_dir():
    d = os.path.join(ROOT, "oracle", "_ref", "pyparament")
    return d if os.path.isdir(os.path.join(d, "parament")) else None


def time_wrapper(w, lib_dir, calls, warm=1):
    """(seconds per pass, pulses per pass) through the UNCHANGED reference wrapper class (parament.Parament.set_hamiltonian /
    .equiprop), bound to the library in `lib_dir`.  Ensembles go pulse by pulse (the wrapper has no batch call)."""
    d = reference_wrapper_dir()
    if d is None:
        return None
    if not hasattr(np, "float"):
        np.float = float            # the wrapper calls np.float(dt) (parament.py:271), removed in numpy >= 1.24
    os.environ["PARAMENT_LIB_DIR"] = lib_dir
    if d not in sys.path:
        sys.path.insert(0, d)
    import parament                 # noqa: E402  (the reference's Python package, unmodified)
    ctx = parament.Parament(precision=w.precision)
    try:
        ctx.set_hamiltonian(w.H0, *w.H1, use_magnus=w.use_magnus, quadrature_mode=w.quadrature)
        pulses = w.carr.reshape(w.batch, w.amps, w.pts)
        nb = min(w.batch, calls)

        def one_pass():
            for b in range(nb):
                ctx.equiprop(w.dt, *pulses[b])
        for _ in range(warm):
            one_pass()
        t = time.time()
        one_pass()
        return (time.time() - t), nb
    finally:
        ctx.destroy()


# ----------------------------------------------------------------------------------------------------------
# our implementation
# ----------------------------------------------------------------------------------------------------------
class Ours:
    def __init__(self, args):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.args = torch, dist, args
        self.rank, self.local_rank, self.world = dist_env()
        if self.world != args.gpus and self.world == 1 and args.gpus > 1:
            raise SystemExit("launch with torch.distributed.run --nproc-per-node N for --gpus N > 1")
        torch.cuda.set_device(self.local_rank)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local_rank))
        os.environ["PARAMENT_DEVICE"] = str(self.local_rank)
        if self.world > 1 and "PARAMENT_STAGE_THREADS" not in os.environ:
            # the library stages pageable host buffers with up to six copying threads per context; N ranks share the box's cores
            os.environ["PARAMENT_STAGE_THREADS"] = str(max(2, min(6, (os.cpu_count() or 8) // self.world)))
        import parament_b200 as pb
        from parament_b200 import constants as K
        self.pb, self.K, self.lib = pb, K, pb._lib.lib
        self.stream = torch.cuda.Stream()   # a real (non-default) stream: the library enqueues on it without synchronising
        torch.cuda.set_stream(self.stream)
        self.peak_dmma = self.lib.Parament_measurePeak(K.PEAK_DMMA)   # roofline denominator, measured on this device now
        self.peak_ffma = self.lib.Parament_measurePeak(K.PEAK_FFMA)
        self.peak_tf32 = self.lib.Parament_measurePeak(K.PEAK_TF32_MMA)
        self.sampler = ClockSampler(self.local_rank)
        if self.rank == 0:
            self.sampler.start()

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, *vals):
        t = self.torch.tensor(list(vals), dtype=self.torch.float64, device="cuda")
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return t.tolist()

    # ------------------------------------------------------------------------------------------------------
    def measure(self, name, steps, warmup, e2e_steps, mode, cpu_budget, with_wrapper):
        """One record.  mode: 'weak' (each rank a full-size slice of a world-times longer pulse) or 'strong' (the one
        workload cut over the ranks).  All ranks call this with the same arguments."""
        torch, dist, K, lib = self.torch, self.dist, self.K, self.lib
        from workloads import make_workload, smooth_pulses
        from parament_b200.distributed import slice_bounds
        rank, world = self.rank, self.world
        w = make_workload(name)
        n, fp64 = w.dim, w.precision == "fp64"
        tdt = torch.complex128 if fp64 else torch.complex64
        sfx = "_fp64" if fp64 else ""
        rp = 2 if (w.use_magnus or w.quadrature == "simpson") else 1
        ov = 0 if (w.quadrature == "none" and not w.use_magnus) else 1

        # ---- this rank's share ----
        sliced = w.batch == 1 and world > 1
        if w.batch == 1 and mode == "weak" and world > 1:
            rng = np.random.default_rng(20260000 + 100 * rank + 7)
            host_full = smooth_pulses(rng, w.amps, w.pts, dtype=np.float32 if not fp64 else np.float64).astype(w.ctype)
            lo, hi = 0, w.steps                                       # own host arrays, all of their steps
            local = host_full.reshape(1, w.amps, w.pts)
            job_steps = w.steps * world
        elif w.batch == 1:
            b = slice_bounds(w.steps, world)
            lo, hi = b[rank], b[rank + 1]
            host_full = np.ascontiguousarray(w.carr.reshape(w.amps, w.pts))
            local = np.ascontiguousarray(host_full[:, rp * lo: rp * hi + ov]).reshape(1, w.amps, -1)
            job_steps = w.steps
        else:
            if mode == "weak" and world > 1:
                rng = np.random.default_rng(20260000 + 100 * rank + 7)
                local = smooth_pulses(rng, w.amps, w.pts, batch=w.batch, dtype=np.float32).astype(w.ctype)
                job_steps = w.total_steps * world
            else:
                p0, p1 = w.batch * rank // world, w.batch * (rank + 1) // world
                local = np.ascontiguousarray(w.carr.reshape(w.batch, w.amps, w.pts)[p0:p1])
                job_steps = w.total_steps
            host_full, lo, hi = None, 0, w.steps
        lbatch, lpts = local.shape[0], local.shape[2]
        local_steps = (hi - lo) * lbatch

        ctx = self.pb.Parament(w.precision, device=self.local_rank)
        ctx.set_hamiltonian(w.H0, *w.H1, use_magnus=w.use_magnus, quadrature_mode=w.quadrature)

        # Inputs larger than L2 (126 MB): rotate over enough distinct device copies of the amplitude stream.
        nbuf = min(12, max(2, int(np.ceil(160e6 / max(local.nbytes, 1))) + 1)) if local.nbytes < 160e6 else 2
        dev_in = [torch.from_numpy(local if i == 0 else np.ascontiguousarray(np.roll(local, 17 * i, axis=-1))).cuda() for i in range(nbuf)]
        flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda") if nbuf * local.nbytes < 130e6 else None
        l2 = (f"inputs rotate over {nbuf} device copies ({nbuf * local.nbytes / 1e6:.0f} MB > 126 MB L2)" if flush is None
              else "L2 flushed by a 256 MiB write between iterations")
        dev_out = torch.zeros(lbatch, n, n, dtype=tdt, device="cuda")
        gathered = torch.zeros(world, n, n, dtype=tdt, device="cuda") if sliced else None
        combined = torch.zeros(n, n, dtype=tdt, device="cuda")
        stream = self.stream
        launches = [0]

        def device_step(i):
            ctx.equiprop_device(w.dt, dev_in[i % nbuf].data_ptr(), lpts, w.amps, dev_out.data_ptr(), batch=lbatch, stream=stream.cuda_stream)
            launches[0] += int(ctx.stat(K.STAT_LAUNCHES))
            if sliced:
                dist.all_gather_into_tensor(gathered, dev_out[0])      # dim^2 per rank over NCCL / NVLink, slice order
                if rank == 0:
                    ctx.combine_device(gathered.data_ptr(), world, combined.data_ptr(), stream=stream.cuda_stream)
                    launches[0] += int(ctx.stat(K.STAT_LAUNCHES))

        for i in range(warmup):
            device_step(i)
        self.barrier()
        launches[0] = 0
        # ---- timed region 1: inputs resident in HBM, CUDA events on the launching stream ----
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        self.barrier()
        t0 = time.time()
        for i in range(steps):
            if flush is not None:
                flush.fill_(i & 0xFF)
            evs[i][0].record(stream)
            device_step(warmup + i)
            evs[i][1].record(stream)
        self.barrier()
        t1 = time.time()
        dev_ms = sum(a.elapsed_time(b) for a, b in evs)
        kernel_ms = dev_ms / steps if world == 1 else ctx.stat(K.STAT_DEVICE_MS)   # N = 1: nothing but the path's kernels in the region
        stats = {"M_used": int(ctx.stat(K.STAT_DEGREE_USED)), "M_ref": int(ctx.stat(K.STAT_DEGREE_REFERENCE)),
                 "products": ctx.stat(K.STAT_PRODUCTS), "horner": int(ctx.stat(K.STAT_HORNER)), "family": int(ctx.stat(K.STAT_FAMILY)),
                 "real_products": ctx.stat(K.STAT_REAL_PRODUCTS), "hnorm": ctx.stat(K.STAT_HNORM),
                 "series_norm": ctx.stat(K.STAT_SERIES_NORM), "math": int(ctx.stat(K.STAT_MATH))}
        gpu_launches = launches[0]

        # ---- timed region 2: end to end through the host-pointer C-ABI ----
        def e2e_region(pinned):
            if sliced:
                src = torch.from_numpy(host_full).pin_memory().numpy() if pinned else pageable_copy(host_full)
                fn = getattr(lib, "Parament_equipropSliceToDevice" + sfx)
                part = torch.zeros(n, n, dtype=tdt, device="cuda")
                res_host = torch.zeros(n, n, dtype=tdt).pin_memory() if pinned else torch.zeros(n, n, dtype=tdt)
                flat = src.reshape(-1)
                pts_full = src.shape[1]

                def step():
                    ec = fn(ctx._handle, flat, float(w.dt), pts_full, w.amps, lo, hi, ctypes.c_void_p(part.data_ptr()))
                    assert ec == 0, ec
                    dist.all_gather_into_tensor(gathered, part)      # the partial never leaves the GPU
                    if rank == 0:
                        ctx.combine_device(gathered.data_ptr(), world, combined.data_ptr(), stream=stream.cuda_stream)
                        res_host.copy_(combined, non_blocking=True)
                    stream.synchronize()
                in_b, out_b = local.nbytes, n * n * (16 if fp64 else 8)
            else:
                src = torch.from_numpy(local).pin_memory().numpy() if pinned else pageable_copy(local)
                out_host = (torch.zeros(lbatch, n, n, dtype=tdt).pin_memory().numpy() if pinned else np.zeros((lbatch, n, n), dtype=w.ctype))
                fn = getattr(lib, "Parament_equipropBatch" + sfx)
                flat, oflat = src.reshape(-1), out_host.reshape(-1)

                def step():
                    ec = fn(ctx._handle, flat, float(w.dt), lpts, w.amps, lbatch, oflat)
                    assert ec == 0, ec
                in_b, out_b = local.nbytes, out_host.nbytes
            for _ in range(2):
                step()
            self.barrier()
            te = time.time()
            for _ in range(e2e_steps):
                step()
            self.barrier()
            return (time.time() - te) * 1e3, in_b, out_b

        e2e_ms, in_b, out_b = e2e_region(pinned=False)
        pin_ms, _, _ = e2e_region(pinned=True)
        dev_ms, e2e_ms, pin_ms = self.max_over_ranks(dev_ms, e2e_ms, pin_ms)
        tot = self.torch.tensor([float(in_b), float(out_b) if (not sliced or rank == 0) else 0.0], dtype=self.torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(tot)
        in_all, out_all = tot.tolist()

        rec = None
        if rank == 0:
            peaks, how = load_measured_peaks()
            F_alg = flops_per_step(w, stats["M_ref"])
            F_exe = flops_per_step(w, stats["products"], stats["real_products"])
            per_gpu_rate = local_steps / (kernel_ms * 1e-3)
            achieved = F_alg * per_gpu_rate * 1e-12
            executed = F_exe * per_gpu_rate * 1e-12
            traffic = None
            try:
                with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
                    traffic = json.load(f).get(name)
            except OSError:
                pass
            # Denominator: the FP64 tensor pipe (DMMA) -- except where the complex64 kernel of dim <= 8 ran in FP32 arithmetic as
            # 3xTF32 split products: SURVEY 8(d) asks for that fraction against the FP32 ceiling the path replaces (FFMA peak),
            # with the tensor-pipe utilisation (three TF32 MACs per fp32-grade MAC) reported beside it.
            tf32 = stats["math"] == 1
            mixed = stats["math"] == 2
            pk = self.peak_ffma if tf32 else self.peak_dmma
            pipe = ("FP32-grade products as 3xTF32 on the warp-level tensor path (HMMA.1688.F32.TF32); peak = measured FFMA rate, the FP32 ceiling it replaces"
                    if tf32 else "FP64 tensor pipe (DMMA.8x8x4)")
            tf32_exec = 3.0 * executed if tf32 else 0.0
            if mixed:
                # two of the step's products (y02, L'R) run at fp32 grade as 3xTF32 on the other tensor sub-pipe: `frac` is the sum of
                # the two pipes' busy fractions at their measured peaks (what ncu's sm__pipe_tensor_cycles_active counts)
                F_exe = flops_per_step(w, stats["products"] - 2, stats["real_products"])
                executed = F_exe * per_gpu_rate * 1e-12
                tf32_exec = 2 * 3 * 8.0 * w.dim ** 3 * per_gpu_rate * 1e-12
                pipe = ("FP64 tensor pipe (DMMA.8x8x4) for X^2 and the running product + TF32 tensor path (HMMA.1688.F32.TF32, 3xTF32) for the "
                        "two small products of the degree-8 form; frac = sum of the two pipes' busy fractions")
            rec = {
                "value": job_steps * steps / (dev_ms * 1e-3), "unit": "steps/s", "n_gpus": world, "steps": steps, "warmup": warmup,
                "ms_per_step": dev_ms / steps, "scaling": mode, "dtype": "f32 (3xTF32)" if tf32 else ("f64 + f32 (3xTF32)" if mixed else "f64"),
                "config": workload_config(name, w, world, mode, l2),
                "implementation": {"degree_reference": stats["M_ref"], "degree_used": stats["M_used"],
                                   "series_evaluation": SERIES_NAMES.get(stats["horner"], str(stats["horner"])),
                                   "matrix_products_per_step": stats["products"],
                                   "real_products_per_complex_product": stats["real_products"], "kernel_family": stats["family"],
                                   "arithmetic": ("fp32 as 3xTF32 (mma.sync.m16n8k8.tf32)" if tf32 else
                                                  "fp64 (mma.sync.m8n8k4.f64) + 3xTF32 for the two small series products" if stats["math"] == 2
                                                  else "fp64 (mma.sync.m8n8k4.f64)"),
                                   "norm_reference": stats["hnorm"], "norm_series": stats["series_norm"],
                                   "effective_steps_per_gpu": local_steps},
                "clocks": self.sampler.window(t0, t1),
                "gpu_launches": gpu_launches,
                "e2e": {"value": job_steps * e2e_steps / (e2e_ms * 1e-3), "unit": "steps/s", "h2d_bytes_per_step": int(in_all),
                        "d2h_bytes_per_step": int(out_all), "ms_per_step": e2e_ms / e2e_steps, "timed_steps": e2e_steps,
                        "host_buffers": "pageable numpy arrays",
                        "api": ("Parament_equipropSliceToDevice + NCCL all-gather + Parament_combineDevice + D2H" if sliced
                                else "Parament_equipropBatch (host pointers in, host pointers out)"),
                        "pinned": {"value": job_steps * e2e_steps / (pin_ms * 1e-3), "ms_per_step": pin_ms / e2e_steps,
                                   "host_buffers": "page-locked"}},
                # `achieved` = algorithmic flops of the reference's recurrence / time.  The product-saving
                # evaluation executes fewer products than that recurrence, so `frac` is computed from the EXECUTED flops and
                # stays a pipe utilisation (<= 1); the algorithmic figure is kept in `algorithmic_frac`.
                "roofline": {"bound": "tensor", "pipe": pipe, "achieved": achieved, "peak": pk, "unit": "TFLOP/s",
                             "frac": ((executed / pk + tf32_exec / self.peak_tf32) if mixed else min(achieved, executed) / pk) if pk > 0 else None,
                             "algorithmic_frac": achieved / pk if pk > 0 else None, "traffic": traffic,
                             "peak_source": ("Parament_measurePeak(FFMA) in this process" if tf32 else
                                             "Parament_measurePeak(DMMA mma.sync.m8n8k4.f64) in this process") + "; MEASURED_PEAKS.json has no FP32 / FP64 figure",
                             "tensor_pipe": ({"executed_tf32_tflops": tf32_exec, "peak_tflops": self.peak_tf32,
                                              "frac": tf32_exec / self.peak_tf32 if self.peak_tf32 > 0 else None} if (tf32 or mixed) else None),
                             "fp64_dmma_peak_tflops": self.peak_dmma,
                             "kernel": KERNEL_NAMES.get(stats["family"]), "kernel_ms_per_launch": kernel_ms,
                             "flops_per_step_algorithmic": F_alg, "flops_per_step_executed": F_exe,
                             "executed_tflops": executed, "executed_frac": executed / pk if pk > 0 else None,
                             # the same rate counted at 8 n^3 flops per complex product: comparable across kernels that form a complex
                             # product from four or from three real ones (the latter spend part of the pipe on operand sums instead)
                             "complex_product_rate_frac": flops_per_step(w, stats["products"], 4) * per_gpu_rate * 1e-12 / pk if pk > 0 else None,
                             "fp32_ffma_peak_tflops": self.peak_ffma, "tf32_mma_sync_peak_tflops": self.peak_tf32,
                             "hbm": {"achieved_gbs": local.nbytes / (kernel_ms * 1e-3) * 1e-9, "peak_gbs": peaks.get("hbm_gbs"),
                                     "peak_source": how, "note": "algorithmic HBM input is the amplitude stream only; not the limiter"}},
            }
            if world == 1 and with_wrapper:
                from parament_b200._lib import DEFAULT_LIB_DIR
                tw = time_wrapper(w, str(DEFAULT_LIB_DIR), calls=500)
                if tw is not None:
                    sec, nb = tw
                    rec["e2e"]["wrapper"] = {"value": w.steps * nb / sec, "ms_per_pass": sec * 1e3, "pulses_per_pass": nb,
                                             "api": "unchanged pyparament Parament.equiprop (oracle/_ref/pyparament) bound to this library"}
            if world == 1 and cpu_budget > 0:
                rec["cpu_baseline"] = cpu_baseline(w, cpu_budget)
        ctx.destroy()
        del dev_in, flush
        torch.cuda.empty_cache()
        return rec

    def finish(self):
        self.sampler.stop()
        if self.world > 1:
            self.dist.barrier()
            self.dist.destroy_process_group()


def run_ours(args):
    o = Ours(args)
    top = o.measure(args.config, args.steps, args.warmup, max(3, min(args.steps, 20)), "weak", 12.0, True)
    subs = {}
    for name in args.sub_configs:
        if name == args.config and o.world == 1:
            continue
        rec = o.measure(name, min(args.steps, SUB_STEPS[name]), max(3, min(args.warmup, 3)), SUB_E2E[name], "strong",
                        4.0 if name in ("C3", "C4", "C5", "C1") else 0.0, name in ("C1", "C5"))
        if rec is not None:
            subs[name] = rec
    o.finish()
    if top is not None:
        line = {"metric": "propagator steps/sec", "value": top["value"], "unit": "steps/s", "n_gpus": o.world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": top["ms_per_step"], "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": top["dtype"], "data": "synthetic"}
        for k in ("config", "implementation", "clocks", "gpu_launches", "e2e", "roofline", "cpu_baseline"):
            if k in top:
                line[k] = top[k]
        line["configs"] = subs
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


def pin_numpy(a):
    """Page-locked copy of `a` for the reference arm (cudaMallocHost of the CUDA runtime; no torch, no parament_b200 in
    this process).  None when no runtime library can be loaded."""
    import glob
    rt = None
    for cand in ["libcudart.so"] + sorted(glob.glob("/usr/local/cuda/lib64/libcudart.so*"), reverse=True):
        try:
            rt = ctypes.CDLL(cand)
            break
        except OSError:
            continue
    if rt is None:
        return None
    p = ctypes.c_void_p()
    if rt.cudaMallocHost(ctypes.byref(p), ctypes.c_size_t(max(a.nbytes, 16))) != 0 or not p.value:
        return None
    buf = (ctypes.c_char * a.nbytes).from_address(p.value)
    out = np.frombuffer(buf, dtype=a.dtype).reshape(a.shape)
    np.copyto(out, a)
    return out            # never freed: the process ends after the measurement


def reference_record(name, steps, warmup, lib, mode="strong"):
    """One configuration on the reference's CUDA build through its C API; (None, workload) if it cannot hold the configuration."""
    from workloads import make_workload
    w = make_workload(name)
    n, fp64 = w.dim, w.precision == "fp64"
    sfx = "_fp64" if fp64 else ""
    if lib is None or not (w.dim ** 2 * w.pts < 2 ** 31 and 3 * w.dim ** 2 * w.pts * (16 if fp64 else 8) < 150e9):
        return None, w
    h = ctypes.c_void_p()
    if getattr(lib, "Parament_create" + sfx)(ctypes.byref(h)) != 0:
        return None, w
    vp = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    H0 = np.ascontiguousarray(w.H0.ravel())
    H1 = np.ascontiguousarray(w.H1.ravel())
    assert getattr(lib, "Parament_setHamiltonian" + sfx)(h, vp(H0), vp(H1), ctypes.c_uint(n), ctypes.c_uint(w.amps),
                                                         ctypes.c_bool(w.use_magnus), ctypes.c_int(QUAD[w.quadrature])) == 0
    nb = min(w.batch, 500)                  # ensembles: the reference has no batch entry point -> sequential calls, bounded sample
    pulses = np.ascontiguousarray(w.carr.reshape(w.batch, w.amps, w.pts)[:nb])
    eq = getattr(lib, "Parament_equiprop" + sfx)

    def run(src, out):
        flat = [src[b].reshape(-1) for b in range(nb)]

        def step():
            for b in range(nb):
                assert eq(h, vp(flat[b]), ctypes.c_double(w.dt), ctypes.c_uint(w.pts), ctypes.c_uint(w.amps), vp(out)) == 0
        for _ in range(max(1, warmup)):
            step()
        t = time.time()
        for _ in range(steps):
            step()
        return (time.time() - t) / steps

    sec = run(pageable_copy(pulses), np.zeros(n * n, dtype=w.ctype))
    pin_src, pin_out = pin_numpy(pulses), pin_numpy(np.zeros(n * n, dtype=w.ctype))
    sec_pin = run(pin_src, pin_out) if pin_src is not None and pin_out is not None else None
    getattr(lib, "Parament_destroy" + sfx)(h)
    value = w.steps * nb / sec
    rec = {"value": value, "unit": "steps/s", "ms_per_step": sec * 1e3, "steps": steps, "warmup": warmup, "scaling": "strong",
           "dtype": "c64" if not fp64 else "c128",
           "config": workload_config(name, w, 1, mode, "host buffers: every call copies its inputs from host memory"),
           "cpu_baseline": {"value": value, "unit": "steps/s", "cores": 1, "kind": "reference",
                            "sample": f"reference CUDA build (cuBLAS batched, compiled from /root/reference for sm_100) on GPU 0 through "
                                      f"its C API, {nb} pulse(s) of {w.pts} points per step, pageable host arrays; the reference has "
                                      f"no CPU implementation"},
           "e2e": {"value": value, "unit": "steps/s", "h2d_bytes_per_step": int(pulses.nbytes), "d2h_bytes_per_step": int(n * n * (16 if fp64 else 8) * nb),
                   "ms_per_step": sec * 1e3, "host_buffers": "pageable numpy arrays", "api": "Parament_equiprop (reference C API)"}}
    if sec_pin is not None:
        rec["e2e"]["pinned"] = {"value": w.steps * nb / sec_pin, "ms_per_step": sec_pin * 1e3, "host_buffers": "page-locked"}
    return rec, w


def run_reference(args):
    rank, local_rank, world = dist_env()
    if rank != 0:
        return
    base = {"impl": "reference", "metric": "propagator steps/sec", "unit": "steps/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "data": "synthetic"}
    libpath = os.path.join(ROOT, "oracle", "_ref", "libparament.so")
    lib = None
    if os.path.exists(libpath):
        try:
            lib = ctypes.cdll.LoadLibrary(libpath)
            h = ctypes.c_void_p()
            if lib.Parament_create(ctypes.byref(h)) == 0:
                lib.Parament_destroy(h)
            else:
                lib = None
        except OSError:
            lib = None
    top, w = reference_record(args.config, args.steps, args.warmup, lib, mode="weak")   # same `config` text as our top-level record
    if top is None:
        # no usable reference CUDA build for this configuration: time the CPU oracle port on a bounded sample
        cb = cpu_baseline(w)
        base.update({"dtype": "c64" if w.precision == "fp32" else "c128", "config": workload_config(args.config, w, 1, "weak", "n/a (CPU)"),
                     "value": cb["value"], "ms_per_step": None, "cpu_baseline": cb,
                     "e2e": {"value": cb["value"], "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                     "note": "reference CUDA build unavailable or cannot hold this configuration (SURVEY.md section 6); CPU oracle port timed instead"})
        print(json.dumps(base), flush=True)
        return
    base.update({k: top[k] for k in ("dtype", "config", "value", "ms_per_step", "cpu_baseline", "e2e")})
    wrapper_lib_dir = os.path.join(ROOT, "oracle", "_ref")
    tw = time_wrapper(w, wrapper_lib_dir, calls=500)
    if tw is not None:
        base["e2e"]["wrapper"] = {"value": w.steps * tw[1] / tw[0], "ms_per_pass": tw[0] * 1e3, "pulses_per_pass": tw[1],
                                  "api": "unchanged pyparament Parament.equiprop bound to the reference library"}
    subs = {}
    for name in args.sub_configs:
        if name == args.config:
            continue
        rec, ww = reference_record(name, min(args.steps, SUB_STEPS[name]), min(args.warmup, 3), lib)
        if rec is None:
            gb = 3 * ww.dim ** 2 * ww.pts * (16 if ww.precision == "fp64" else 8) / 1e9
            subs[name] = {"unavailable": f"the reference allocates three dim^2 x pts arrays ({gb:.0f} GB) and indexes them with 32-bit "
                                         f"integers (parament.cpp:425-429,590): it cannot hold this configuration"}
            continue
        if name in ("C1", "C5"):
            tw = time_wrapper(ww, wrapper_lib_dir, calls=500)
            if tw is not None:
                rec["e2e"]["wrapper"] = {"value": ww.steps * tw[1] / tw[0], "ms_per_pass": tw[0] * 1e3, "pulses_per_pass": tw[1],
                                         "api": "unchanged pyparament Parament.equiprop bound to the reference library"}
        subs[name] = rec
    base["configs"] = subs
    print(json.dumps(base), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--config", default="C2", choices=list(SUB_STEPS))
    ap.add_argument("--configs", default="all", help="sub-records: 'all' (default set), 'none', or a comma list of configuration names")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    world = dist_env()[2]
    if args.configs == "all":
        args.sub_configs = list(DEFAULT_SUBS_1 if (max(world, args.gpus) == 1 or args.impl == "reference") else DEFAULT_SUBS_N)
    elif args.configs == "none":
        args.sub_configs = []
    else:
        args.sub_configs = [c for c in args.configs.split(",") if c in SUB_STEPS]
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
