#!/usr/bin/env python3
"""bench.py -- the headline measurement of BASELINE.json, on synthetic data.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

metric   : output Mpix/s, resample linear+jacobian 8192^2 (BASELINE.json)
one step : one 8192x8192 float32 frame resampled twice --
             (1) method 'linear'   through Scale(1.3) o Rotation(30 deg) o Offset   (config 2 at 8192^2)
             (2) method 'Gaussian' (Jacobian-adaptive filter) through the polar unwrap
                 Scale o Radial(origin=centre)                                      (config 3)
           = 2 x 67.1 Mpix of output per step.
value    : whole-job output Mpix/s with the frame resident in HBM (CUDA events, max over ranks),
           fp32 fast path.  `parity_mode` carries the same step in the fp64 parity mode.
e2e      : the same step through the public API with HOST buffers: one H2D upload of the pinned frame
           per step, both kernels, both results copied into pinned host outputs -- all inside the timed
           region (transform_b200.AsyncResampler.submit_many; consecutive steps overlap on 4 streams).
N > 1    : `value` -- one process per GPU, each rank resamples its own frame ("scaling": "weak", no
           collective in the data path).  `sharded` -- the strong-scaling job: ONE 8192^2 frame resident
           on rank 0, each rank is sent the inverse-mapped bounding box of its part of the output grid
           (NCCL grouped send/recv), resamples it, and the parts are all-gathered while the next strip is
           being computed (transform_b200.sharding.ShardedResampler); plus BASELINE config 4's 64-frame
           batch split by frames.
roofline : for the kernel that dominates the step, algorithmic bytes (output bytes + input bytes,
           each once: SURVEY.md section 8d) / its CUDA-event duration, against the measured HBM copy
           bandwidth in MEASURED_PEAKS.json.  `linear_kernel` reports the pure-shift map (every source
           pixel is read: DRAM traffic = algorithmic bytes) as THE linear figure, and the named chain
           (42 % of whose output maps inside the source) beside its measured DRAM bytes.
cpu_baseline / --impl reference : the reference's CPU path (oracle/cpu_baseline.py) on the box's host
           cores, on a bounded sample of the same workload, on one persistent pool of forked workers.
"""
import argparse
import json
import os

# keep stdout to exactly one JSON line: NCCL prints its version banner there at VERSION level
if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
    os.environ["NCCL_DEBUG"] = "WARN"
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
N_IMG = 8192
METRIC = "output Mpix/s, resample linear+jacobian 8192^2"
UNIT = "Mpix/s"
ALG_BYTES = 2 * N_IMG * N_IMG * 4            # fp32 out + fp32 in, each once
CPU_KIND = {"linear": "reference (the reference's own Composition.invert and interpND: core.py, basic.py and helpers.pyx "
                      "compiled from the files where they lie into oracle/_ref, which travels to the GPU box)",
            "jacobian": "port (C restatement, bit-identical to the reference's own code built with the repairs of "
                        "oracle/build_ref_repaired.py and ~30x faster per core than it)"}


def config(n_gpus, extra=None):
    c = {"workload": "8192x8192 float32 frame per GPU; step = linear resample through "
                     "Scale(1.3)oRotation(30deg)oOffset + Jacobian-adaptive Gaussian resample through "
                     "ScaleoRadial (polar unwrap); bound='t'",
         "pixels_per_step_per_gpu": 2 * N_IMG * N_IMG,
         "precision": "fp32 fast path (coordinates fp64, interpolation fp32); parity mode in `parity_mode`",
         "cache": "inputs (268 MB) larger than L2 (126 MB); no explicit flush",
         "parallelism": "frames sharded across %d GPU(s), no collective in the data path" % n_gpus}
    if extra:
        c.update(extra)
    return c


# ------------------------------------------------------------------------------------------------
# clocks sampler (nvidia-smi polled while the timed region runs)
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock and throttle reasons of one GPU while the timed region runs: NVML polled from a thread every
    ~2 ms (a 20-step timed region lasts ~25 ms: `nvidia-smi -lms` delivers one or two samples in that time);
    falls back to one streaming `nvidia-smi -lms 20` process when the NVML binding is missing."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index):
        self.index = index
        self.samples = []                       # (sm_mhz, sm_max_mhz, set of reasons, time)
        self.window = [None, None]              # only samples taken inside the timed region count
        self.proc = None
        self.thread = None
        self.stop = threading.Event()
        self.source = None

    def _nvml_loop(self, nv, handle, sm_max):
        bits = [(nv.nvmlClocksEventReasonHwSlowdown, "hw_slowdown"),
                (nv.nvmlClocksEventReasonHwThermalSlowdown, "hw_thermal_slowdown"),
                (nv.nvmlClocksEventReasonSwThermalSlowdown, "sw_thermal_slowdown"),
                (nv.nvmlClocksEventReasonSwPowerCap, "sw_power_cap")]
        while not self.stop.is_set():
            try:
                mhz = nv.nvmlDeviceGetClockInfo(handle, nv.NVML_CLOCK_SM)
                mask = nv.nvmlDeviceGetCurrentClocksEventReasons(handle)
                self.samples.append((float(mhz), float(sm_max), {nm for b, nm in bits if mask & b}, time.perf_counter()))
            except Exception:
                pass
            time.sleep(0.002)

    def _pump(self):
        for line in self.proc.stdout:
            f = [x.strip() for x in line.split(",")]
            if len(f) >= 6:
                try:
                    self.samples.append((float(f[0]), float(f[1]),
                                         {nm for nm, v in zip(self.NAMES, f[2:6]) if v.lower().startswith("active")},
                                         time.perf_counter()))
                except ValueError:
                    continue

    def __enter__(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            # NVML indexes physical devices: honour CUDA_VISIBLE_DEVICES when it lists indices
            vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
            phys = self.index
            if vis and all(v.strip().isdigit() for v in vis.split(",")):
                phys = int(vis.split(",")[self.index])
            handle = nv.nvmlDeviceGetHandleByIndex(phys)
            sm_max = nv.nvmlDeviceGetMaxClockInfo(handle, nv.NVML_CLOCK_SM)
            self.thread = threading.Thread(target=self._nvml_loop, args=(nv, handle, sm_max), daemon=True)
            self.thread.start()
            self.source = "nvml, 2 ms"
            return self
        except Exception:
            self.thread = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
            time.sleep(0.3)                       # let the sampler come up before the timed region
            self.samples.clear()
            self.source = "nvidia-smi -lms 20"
        except Exception:
            self.proc = None
        return self

    def __exit__(self, *a):
        self.stop.set()
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()
        if self.thread is not None:
            self.thread.join(timeout=5)

    def begin(self):
        self.window[0] = time.perf_counter()

    def end(self):
        self.window[1] = time.perf_counter()

    def summary(self):
        t0, t1 = self.window
        if t0 is not None and t1 is not None:
            inside = [s for s in self.samples if t0 <= s[3] <= t1]
            if inside:                          # (the nvidia-smi fallback may deliver nothing inside a short region)
                self.samples = inside
        sm = sorted(s[0] for s in self.samples)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0, "source": self.source}
        reasons = set()
        for s in self.samples:
            reasons |= s[2]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(s[1] for s in self.samples), "reasons": sorted(reasons),
                "samples": len(sm), "source": self.source}


# ------------------------------------------------------------------------------------------------
# CPU side (reference arm and cpu_baseline leg)
# ------------------------------------------------------------------------------------------------
def cpu_arm(nproc):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import cpu_baseline as cb
    return cb.CpuArm(N_IMG, nproc, seed=0)


def cpu_step(arm, rows_linear, rows_jacobian):
    """One bounded sample of the step: rows of the linear resample + rows of the Jacobian resample."""
    t_lin, px_lin = arm.linear(rows_linear)
    t_jac, px_jac = arm.jacobian(rows_jacobian)
    # one step = N^2 pixels of each method: combine the two per-pixel rates accordingly
    full = float(N_IMG * N_IMG)
    t_step = full * (t_lin / px_lin) + full * (t_jac / px_jac)
    return {"value": 2 * full / t_step / 1e6, "t_lin": t_lin, "px_lin": px_lin, "t_jac": t_jac, "px_jac": px_jac,
            "linear_mpix_s": px_lin / t_lin / 1e6, "jacobian_mpix_s": px_jac / t_jac / 1e6}


def cpu_rows(arm, seconds_per_leg):
    """Rows per leg such that each leg of a sample takes about `seconds_per_leg` on this box."""
    nproc = arm.nproc
    cal = cpu_step(arm, nproc * 16, nproc * 4)
    rl = int(seconds_per_leg * (cal["px_lin"] / cal["t_lin"]) / N_IMG)
    rj = int(seconds_per_leg * (cal["px_jac"] / cal["t_jac"]) / N_IMG)
    rl = max(nproc * 8, min(rl, N_IMG)) // nproc * nproc
    rj = max(nproc * 2, min(rj, N_IMG)) // nproc * nproc
    return rl, rj


def cpu_kind(arm):
    k = dict(CPU_KIND)
    if arm.kind_lin == "reference-interp":
        k["linear"] = ("reference interpND (oracle/_ref/helpers*.so) fed by a port of Composition.invert that calls "
                       "np.matmul on [...,2,1] stacks exactly as basic.py:150-152 (oracle/_ref/core*.so is missing)")
    elif arm.kind_lin != "reference":
        k["linear"] = "port (oracle/_ref/helpers*.so is missing: the NumPy restatement of interpND)"
    return k


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    nproc = os.cpu_count() or 1
    arm = cpu_arm(nproc)
    try:
        # each step is a bounded sample sized so that K + W steps end within ~3 minutes
        budget = 170.0 / max(1, args.steps + args.warmup)
        rl, rj = cpu_rows(arm, 0.5 * budget)
        for _ in range(args.warmup):
            cpu_step(arm, rl, rj)
        vals, t0 = [], time.perf_counter()
        for _ in range(args.steps):
            vals.append(cpu_step(arm, rl, rj))
        wall = time.perf_counter() - t0
        kind = cpu_kind(arm)
    finally:
        arm.close()
    value = sum(v["value"] for v in vals) / len(vals)
    lin = sum(v["linear_mpix_s"] for v in vals) / len(vals)
    jac = sum(v["jacobian_mpix_s"] for v in vals) / len(vals)
    sample = ("per step: %d rows of the 8192^2 linear resample + %d rows of the Jacobian-Gaussian resample, spread over "
              "the frame, on one persistent pool of %d forked workers; per-pixel rates combined for a full step"
              % (rl, rj, nproc))
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * wall / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "impl": "reference", "config": config(args.gpus),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": nproc, "kind": kind, "sample": sample,
                             "linear_mpix_s": lin, "jacobian_mpix_s": jac},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def traffic_for(kernel_key):
    """dram bytes per launch from the committed ncu --set full captures (profiles/traffic.json, written
    by tools/ncu_evidence.py from the raw pages of those captures), if one exists for this kernel."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        with open(p) as f:
            v = json.load(f).get(kernel_key)
        return v.get("dram_bytes") if isinstance(v, dict) else v
    except Exception:
        return None


def ncu_for(kernel_key):
    """Pipe utilisations and hit rates of the same committed capture (what bounds a kernel that is not HBM-bound)."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        with open(p) as f:
            v = json.load(f).get(kernel_key)
        return {k: v[k] for k in ("issue_active_pct", "tex_writeback_pct", "l1tex_throughput_pct", "dram_throughput_pct",
                                  "l1_hit_pct", "l2_hit_pct", "source") if k in v}
    except Exception:
        return None


def extra_configs(t, torch, np, timed):
    """BASELINE configs 2, 4 and 5 at full size (untimed for `value`; reported in `breakdown`)."""
    out = {}
    peak, _ = measured_peak()
    # config 2: 4096^2 float32, Scale o Rotation o Offset, nearest and linear
    n = 4096
    g = torch.Generator(device="cuda").manual_seed(7)
    src = torch.rand((n, n), generator=g, device="cuda", dtype=torch.float32)
    dst = torch.empty_like(src)
    tr = t.Composition([t.Scale(1.3, dim=2), t.Rotation(30, unit="deg"), t.Offset([-n / 2.0, -n / 2.0])])
    for m in ("nearest", "linear"):
        ms = timed(lambda: tr.resample(src, method=m, precision="fp32", out=dst), reps=20)
        out["config2_4096_%s_fp32_ms" % m] = ms
        out["config2_4096_%s_fp32_mpix_s" % m] = n * n / (ms * 1e-3) / 1e6
        out["config2_4096_%s_fp32_roofline_frac" % m] = 2 * n * n * 4 / (ms * 1e-3) / 1e9 / peak
    # other affine maps at 8192^2 (the 60 % target should not hinge on one angle)
    n8 = N_IMG
    src8 = torch.rand((n8, n8), generator=g, device="cuda", dtype=torch.float32)
    dst8 = torch.empty_like(src8)
    for name, trx in (("rot45", t.Wrap(t.Rotation(45, unit="deg"), t.Offset([-n8 / 2.0, -n8 / 2.0]))),
                      ("rot5", t.Wrap(t.Rotation(5, unit="deg"), t.Offset([-n8 / 2.0, -n8 / 2.0]))),
                      ("scale0p8", t.Scale(0.8, dim=2)), ("scale2", t.Scale(2.0, dim=2)),
                      ("shift", t.Offset([3.25, -7.5]))):
        ms = timed(lambda: trx.resample(src8, method="linear", precision="fp32", out=dst8), reps=10)
        out["linear_%s_8192_fp32_ms" % name] = ms
        out["linear_%s_8192_fp32_roofline_frac" % name] = ALG_BYTES / (ms * 1e-3) / 1e9 / peak
        if name in ("rot5", "shift"):
            ms = timed(lambda: trx.resample(src8, method="cubic", precision="fp32", out=dst8), reps=10)
            out["cubic_%s_8192_fp32_ms" % name] = ms
            out["cubic_%s_8192_fp32_roofline_frac" % name] = ALG_BYTES / (ms * 1e-3) / 1e9 / peak
    # anti-aliased down-scaling (the constant-Jacobian kernel) and the reference's windowed input-plane filters
    trd = t.Composition([t.Offset([n8 / 2.0 + 0.3, n8 / 2.0 - 0.2]), t.Scale([0.6, 0.7]), t.Rotation(25, unit="deg"),
                         t.Offset([-n8 / 2.0, -n8 / 2.0])])
    out["jacobian_downscale_8192_fp32_ms"] = timed(lambda: trd.resample(src8, method="Gaussian", precision="fp32", out=dst8), reps=5)
    trn = t.Composition([t.Scale(1.3, dim=2), t.Rotation(30, unit="deg"), t.Offset([-n8 / 2.0, -n8 / 2.0])])
    for m in ("hann", "zlanczos"):
        ms = timed(lambda: trn.resample(src8, method=m, precision="fp32", out=dst8), reps=5)
        out["%s_named_8192_fp32_ms" % m] = ms
    del src8, dst8
    # config 4: batch of 64 4096^2 frames, TAN -> CAR, linear (one launch, frames as planes)
    total, nb, n = config4_transform(t)
    frames = torch.rand((nb, n, n), generator=g, device="cuda", dtype=torch.float32)
    outb = torch.empty_like(frames)
    ms = timed(lambda: total.resample(frames, method="linear", precision="fp32", out=outb, shape=(nb, n, n)), reps=2)
    out["config4_64x4096_tan2car_linear_fp32_ms"] = ms
    out["config4_64x4096_tan2car_linear_fp32_mpix_s"] = nb * n * n / (ms * 1e-3) / 1e6
    out["config4_64x4096_tan2car_linear_fp32_roofline_frac"] = 2 * nb * n * n * 4 / (ms * 1e-3) / 1e9 / peak
    del frames, outb
    # config 5: Transform.apply on 1e9 2-vectors through an inverted composed chain (fp64: 32 GB of traffic)
    nv = 10 ** 9
    v = (torch.rand((nv, 2), generator=g, device="cuda", dtype=torch.float64) - 0.5) * 8192
    aff = t.Composition([t.Scale([1.5, 2]), t.Rotation(30, unit="deg"), t.Offset([3, 4])]).inverse()
    rad = t.Composition([t.Scale([1.5, 2]), t.Radial(origin=[0.1, 0.2]), t.Offset([3, 4])]).inverse()
    for name, trx in (("affine", aff), ("radial", rad)):
        ms = timed(lambda: trx.apply(v), reps=2)
        out["config5_apply_1e9_%s_fp64_ms" % name] = ms
        out["config5_apply_1e9_%s_fp64_mvec_s" % name] = nv / (ms * 1e-3) / 1e6
        out["config5_apply_1e9_%s_fp64_roofline_frac" % name] = nv * 2 * 2 * 8 / (ms * 1e-3) / 1e9 / peak
    del v
    torch.cuda.empty_cache()
    return out


def config4_transform(t):
    """BASELINE config 4: 64 FITS-shaped 4096^2 frames, TAN -> CAR (what remap() builds, core.py:920-923)."""
    n = 4096
    hdr = {'NAXIS': 2, 'NAXIS1': n, 'NAXIS2': n, 'CTYPE1': 'RA---TAN', 'CTYPE2': 'DEC--TAN', 'CUNIT1': 'deg',
           'CUNIT2': 'deg', 'CRPIX1': n / 2 + 0.5, 'CRPIX2': n / 2 + 0.5, 'CRVAL1': 150.0, 'CRVAL2': 20.0,
           'CDELT1': -1.5 / n, 'CDELT2': 1.5 / n}
    car = dict(hdr, CTYPE1='RA---CAR', CTYPE2='DEC--CAR')
    return t.Composition([t.WCS(car).inverse(), t.WCS(hdr)]), 64, n


def sharded_job(t, torch, dist, np, src, tr_aff, tr_rad, rank, world, ms_single, reps):
    """The strong-scaling job: ONE 8192^2 frame resident on rank 0, output split across the ranks."""
    from transform_b200 import sharding
    n = N_IMG
    out = {"job": "one 8192^2 float32 frame resident on rank 0; scatter of inverse-mapped source rectangles "
                  "(NCCL grouped send/recv), resample, all_gather of 4 strips per rank overlapped with compute",
           "world": world}
    peer_gbs = 770.0           # measured peer-copy bandwidth per direction per GPU (B200_PROFILING.md)
    for key, tr, method in (("linear", tr_aff, "linear"), ("jacobian", tr_rad, "Gaussian")):
        sr = sharding.ShardedResampler(tr, (n, n), method=method, bound="t", nsub=4, precision="fp32")
        for _ in range(2):
            res = sr(src if rank == 0 else None)
        # three batches of `reps` calls, each timed on the device and reduced with MAX over the ranks; the best
        # batch is reported (one of ~40 small NCCL operations per call occasionally stalls for milliseconds:
        # seen once in four runs at N = 8, 6.2 ms against 2.0-2.1)
        batches = []
        for _ in range(3):
            dist.barrier()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps):
                res = sr(src if rank == 0 else None, sync=False)
            e1.record()
            dist.barrier()
            torch.cuda.synchronize()
            tt = torch.tensor([e0.elapsed_time(e1) / reps], device="cuda", dtype=torch.float64)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            batches.append(float(tt[0]))
        ms = min(batches)
        ok = True
        if rank == 0:
            want = tr.resample(src, method=method, precision="fp32")
            ok = bool(torch.equal(res, want)) if key == "linear" else bool(float((res - want).abs().max()) < 1e-5)
            del want
        link_ms = max(sr.bytes_root_out, sr.bytes_rank_in) / (peer_gbs * 1e9) * 1e3
        out[key] = {"ms": ms, "mpix_s": n * n / (ms * 1e-3) / 1e6, "single_gpu_ms": ms_single[key],
                    "speedup": ms_single[key] / ms, "efficiency": ms_single[key] / ms / world,
                    "split": sr.split, "staged_source_fraction": sr.staged_pixels / float(n * n),
                    "nvlink_bytes_scatter": int(sr.bytes_scatter), "nvlink_bytes_gather": int(sr.bytes_gather),
                    "busiest_link_bytes": int(max(sr.bytes_root_out, sr.bytes_rank_in)),
                    "link_bound_ms": link_ms,
                    "limiter": "NVLink: rank 0 sends %.0f MB of source rectangles and every rank receives %.0f MB of "
                               "gathered output; at the measured 770 GB/s per direction that alone is %.2f ms against "
                               "%.2f ms of kernel time per rank" % (sr.bytes_scatter / 1e6, sr.bytes_rank_in / 1e6,
                                                                     link_ms, ms_single[key] / world),
                    "batches_ms": batches, "equals_unsharded": ok}
        del sr, res
    # BASELINE config 4: the 64-frame batch split by frames (8 per GPU at N = 8), no communication
    total, nb, n4 = config4_transform(t)
    per = nb // world
    g = torch.Generator(device="cuda").manual_seed(4000 + rank)
    frames = torch.rand((per, n4, n4), generator=g, device="cuda", dtype=torch.float32)
    outb = torch.empty_like(frames)
    for _ in range(2):
        total.resample(frames, method="linear", precision="fp32", out=outb, shape=(per, n4, n4))
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        total.resample(frames, method="linear", precision="fp32", out=outb, shape=(per, n4, n4))
    e1.record()
    dist.barrier()
    torch.cuda.synchronize()
    tt = torch.tensor([e0.elapsed_time(e1) / 3], device="cuda", dtype=torch.float64)
    dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    out["config4_batch"] = {"frames_per_gpu": per, "ms": float(tt[0]),
                            "mpix_s": per * world * n4 * n4 / (float(tt[0]) * 1e-3) / 1e6,
                            "job": "64 x 4096^2 TAN->CAR linear, %d frames per GPU, frames resident, no collective" % per}
    del frames, outb
    return out


def gpu_arm(args):
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product path has no CPU fallback)")
    torch.cuda.set_device(local)
    # host buffers of this rank on the NUMA node of its GPU (matters for the end-to-end number at N > 1)
    from transform_b200 import _device as tdev
    all_cpus = os.sched_getaffinity(0)
    numa_cpus = None if os.environ.get("TFM_BENCH_NO_NUMA") else tdev.bind_to_gpu_numa(local)
    # NCCL writes its version banner to stdout; park stdout on stderr until the JSON line is due
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    sys.path.insert(0, ROOT)
    import numpy as np
    import transform_b200 as t
    from transform_b200 import _lib as L

    n = N_IMG
    g = torch.Generator(device="cuda").manual_seed(1000 + rank)
    src = torch.rand((n, n), generator=g, device="cuda", dtype=torch.float32)
    out_lin = torch.empty((n, n), device="cuda", dtype=torch.float32)
    out_jac = torch.empty((n, n), device="cuda", dtype=torch.float32)
    tr_aff = t.Composition([t.Scale(1.3, dim=2), t.Rotation(30, unit="deg"), t.Offset([-n / 2.0, -n / 2.0])])
    c = (n - 1) / 2.0
    tr_rad = t.Composition([t.Scale([n / (2 * np.pi), n / (n / np.sqrt(2.0))]), t.Radial(origin=[c, c])])
    tr_shift = t.Offset([3.25, -7.5])

    def step(prec="fp32", ev=None, o_lin=out_lin, o_jac=out_jac):
        if ev is not None:
            ev[0].record()
        tr_aff.resample(src, method="linear", precision=prec, out=o_lin)
        if ev is not None:
            ev[1].record()
        tr_rad.resample(src, method="Gaussian", precision=prec, out=o_jac)
        if ev is not None:
            ev[2].record()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(3, args.warmup)):
        step()
    kernels = {}
    tr_aff.resample(src, method="linear", precision="fp32", out=out_lin)
    kernels["linear"] = L.last_kernel()
    tr_rad.resample(src, method="Gaussian", precision="fp32", out=out_jac)
    kernels["jacobian"] = L.last_kernel()
    barrier()
    launches0 = L.launch_count()
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(args.steps)]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clk:
        barrier()
        clk.begin()
        e0.record()
        for k in range(args.steps):
            step(ev=evs[k])
        e1.record()
        barrier()
        clk.end()
    ms_total = e0.elapsed_time(e1)
    launches = L.launch_count() - launches0
    ms_lin = sum(ev[0].elapsed_time(ev[1]) for ev in evs) / args.steps
    ms_jac = sum(ev[1].elapsed_time(ev[2]) for ev in evs) / args.steps

    def timed(fn, reps=5):
        fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / reps

    # ---- the same step in the fp64 parity mode (device-timed; the reference's own arithmetic) ----
    o64a = torch.empty((n, n), device="cuda", dtype=torch.float64)
    o64b = torch.empty((n, n), device="cuda", dtype=torch.float64)
    k_par = max(1, min(args.steps, 5))
    ms_par = timed(lambda: step(prec="fp64", o_lin=o64a, o_jac=o64b), reps=k_par)
    ms_par_lin = timed(lambda: tr_aff.resample(src, method="linear", out=o64a), reps=k_par)
    ms_shift = timed(lambda: tr_shift.resample(src, method="linear", precision="fp32", out=out_lin), reps=20)
    kernels["linear_shift"] = L.last_kernel()

    # ---- end to end through the public API with host buffers ----
    h_src = torch.empty((n, n), dtype=torch.float32).pin_memory()
    h_src.copy_(src)
    h_lin = [torch.empty((n, n), dtype=torch.float32).pin_memory() for _ in range(2)]
    h_jac = [torch.empty((n, n), dtype=torch.float32).pin_memory() for _ in range(2)]
    for b in h_lin + h_jac:
        b.zero_()

    def step_e2e_sync():
        tr_aff.resample(h_src, method="linear", precision="fp32", out=h_lin[0])
        tr_rad.resample(h_src, method="Gaussian", precision="fp32", out=h_jac[0])

    step_e2e_sync()
    barrier()
    k_e2e = max(1, min(args.steps, 20))
    t0 = time.perf_counter()
    for _ in range(k_e2e):
        step_e2e_sync()
    barrier()
    s_e2e_sync = (time.perf_counter() - t0) / k_e2e
    step()
    torch.cuda.synchronize()
    ok = bool(torch.equal(h_lin[0], out_lin.cpu()) and torch.equal(h_jac[0], out_jac.cpu()))

    # the public pipelined API: ONE upload of the frame per step, both kernels, both downloads; uploads,
    # kernels and downloads of consecutive steps overlap on 4 streams
    pipe = t.AsyncResampler(depth=4)

    def step_e2e_async(k):
        return pipe.submit_many(h_src, [(tr_aff, h_lin[k & 1], "linear", {"precision": "fp32"}),
                                        (tr_rad, h_jac[k & 1], "Gaussian", {"precision": "fp32"})])

    step_e2e_async(0)
    step_e2e_async(1)
    pipe.drain()
    barrier()
    t0 = time.perf_counter()
    for k in range(k_e2e):
        step_e2e_async(k)
    pipe.drain()
    barrier()
    s_e2e = (time.perf_counter() - t0) / k_e2e
    ok = ok and bool(torch.equal(h_lin[1], h_lin[0]) and torch.equal(h_jac[1], h_jac[0]))
    # the same host<->device traffic with NO kernels, all ranks at once: what this box lets the end-to-end number be
    s_up, s_dn = torch.cuda.Stream(), torch.cuda.Stream()
    scratch = torch.empty_like(src)

    def copies_only():
        with torch.cuda.stream(s_up):
            scratch.copy_(h_src, non_blocking=True)
        with torch.cuda.stream(s_dn):
            h_lin[1].copy_(out_lin, non_blocking=True)
            h_jac[1].copy_(out_jac, non_blocking=True)

    copies_only()
    barrier()
    t0 = time.perf_counter()
    for _ in range(k_e2e):
        copies_only()
    barrier()
    s_copy = (time.perf_counter() - t0) / k_e2e
    del scratch
    # parity mode end to end (float64 results: twice the download)
    h64 = [torch.empty((n, n), dtype=torch.float64).pin_memory() for _ in range(2)]

    def step_e2e_par():
        return pipe.submit_many(h_src, [(tr_aff, h64[0], "linear", {}), (tr_rad, h64[1], "Gaussian", {})])

    step_e2e_par()
    pipe.drain()
    barrier()
    k_pe = max(1, min(args.steps, 3))
    t0 = time.perf_counter()
    for _ in range(k_pe):
        step_e2e_par()
    pipe.drain()
    barrier()
    s_e2e_par = (time.perf_counter() - t0) / k_pe
    del h64, o64b

    # ---- breakdown (untimed for `value`) ----
    brk = {}
    if rank == 0:
        brk["linear_affine_fp32_ms"] = ms_lin
        brk["jacobian_radial_fp32_ms"] = ms_jac
        brk["linear_radial_fp32_ms"] = timed(lambda: tr_rad.resample(src, method="linear", precision="fp32", out=out_lin))
        for tn, nm in ((128, "texture_gather_first_form"), (16, "ldg_gather"), (256, "tma_one_box_per_cta")):
            brk["linear_affine_fp32_%s_ms" % nm] = timed(
                lambda: tr_aff.resample(src, method="linear", precision="fp32", out=out_lin, tuning=tn))
        for tn, nm in ((32768, "ld64"), (4096, "ld32"), (512, "finite_difference_kernel")):
            brk["jacobian_radial_fp32_%s_ms" % nm] = timed(
                lambda: tr_rad.resample(src, method="Gaussian", precision="fp32", out=out_jac, tuning=tn), reps=3)
        brk["nearest_affine_fp32_ms"] = timed(lambda: tr_aff.resample(src, method="nearest", precision="fp32", out=out_lin))
        brk["cubic_affine_fp32_ms"] = timed(lambda: tr_aff.resample(src, method="cubic", precision="fp32", out=out_lin))
        brk["linear_affine_fp64_parity_ms"] = ms_par_lin
        brk["jacobian_radial_fp64_parity_ms"] = ms_par - ms_par_lin
        for k in list(brk):
            brk[k.replace("_ms", "_mpix_s")] = n * n / (brk[k] * 1e-3) / 1e6
        if not args.no_extra:
            brk.update(extra_configs(t, torch, np, timed))
    del o64a

    sharded = None
    if world > 1:
        # single-GPU times of the two calls on THIS run's rank 0 are the strong-scaling reference
        ms1 = torch.tensor([ms_lin, ms_jac], device="cuda", dtype=torch.float64)
        dist.broadcast(ms1, 0)
        sharded = sharded_job(t, torch, dist, np, src, tr_aff, tr_rad, rank, world,
                              {"linear": float(ms1[0]), "jacobian": float(ms1[1])}, reps=max(3, min(args.steps, 10)))

    # max over ranks
    tt = torch.tensor([ms_total, s_e2e, s_e2e_sync, ms_par, s_e2e_par, s_copy], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    ms_total, s_e2e, s_e2e_sync, ms_par, s_e2e_par, s_copy = (float(v) for v in tt)

    if rank == 0:
        px_step = 2.0 * n * n * world
        peak, peak_src = measured_peak()
        dom_is_jac = ms_jac >= ms_lin
        dom_ms = ms_jac if dom_is_jac else ms_lin
        achieved = ALG_BYTES / (dom_ms * 1e-3) / 1e9
        lin_gbs = ALG_BYTES / (ms_lin * 1e-3) / 1e9
        shift_gbs = ALG_BYTES / (ms_shift * 1e-3) / 1e9
        line = {"metric": METRIC, "value": px_step / (ms_total / args.steps * 1e-3) / 1e6, "unit": UNIT,
                "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
                "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config(world),
                "roofline": {"bound": "hbm", "kernel": kernels["jacobian" if dom_is_jac else "linear"],
                             "achieved": achieved, "peak": peak,
                             "unit": "GB/s", "frac": achieved / peak, "peak_source": peak_src,
                             "traffic": traffic_for("jacobian" if dom_is_jac else "linear"),
                             "ncu": ncu_for("jacobian" if dom_is_jac else "linear"),
                             "algorithmic_bytes_per_launch": ALG_BYTES,
                             "share_of_step": dom_ms / (ms_lin + ms_jac),
                             "note": "the Jacobian kernel is bound by instruction issue and the texture pipe, not by "
                                     "HBM (DESIGN.md section 4): its fraction of the HBM roofline is reported, not hidden",
                             "linear_kernel": {
                                 "map": "pure shift (every source pixel is read once: DRAM traffic = algorithmic bytes)",
                                 "kernel": kernels["linear_shift"], "ms": ms_shift, "achieved": shift_gbs,
                                 "frac": shift_gbs / peak, "traffic": traffic_for("linear_shift"),
                                 "ncu": ncu_for("linear_shift"),
                                 "named_chain": {"kernel": kernels["linear"], "ms": ms_lin, "achieved": lin_gbs,
                                                 "frac_of_algorithmic_bytes": lin_gbs / peak,
                                                 "traffic": traffic_for("linear"),
                                                 "note": "58 % of this chain's output tiles map outside the source and "
                                                         "are rejected before any load: measured DRAM traffic is about "
                                                         "half of the algorithmic bytes, so this fraction flatters"}}},
                "parity_mode": {"precision": "fp64 (the reference's own arithmetic; default of the API)",
                                "value": px_step / (ms_par * 1e-3) / 1e6, "ms_per_step": ms_par, "unit": UNIT,
                                "e2e_value": px_step / s_e2e_par / 1e6,
                                "e2e_d2h_bytes_per_step": 2 * n * n * 8},
                "e2e": {"value": px_step / s_e2e / 1e6, "unit": UNIT,
                        "h2d_bytes_per_step": n * n * 4, "d2h_bytes_per_step": 2 * n * n * 4,
                        "steps": k_e2e, "matches_resident_result": ok,
                        "mode": "transform_b200.AsyncResampler.submit_many, 4 streams: one upload of the frame per "
                                "step, H2D / kernels / D2H of consecutive steps overlap; pinned host buffers",
                        "copies_only_ceiling_value": px_step / s_copy / 1e6,
                        "copies_only_note": "the same pinned H2D / D2H traffic per step with no kernels, all ranks at once, "
                                            "two streams per rank: the most this box lets `e2e.value` be",
                        "synchronous_calls_value": px_step / s_e2e_sync / 1e6,
                        "synchronous_calls_note": "plain Transform.resample(host, out=host) twice per step: two uploads",
                        "host_affinity": "NVML ideal CPU set of the rank's GPU (%d cores)" % len(numa_cpus)
                                         if numa_cpus else "unchanged"},
                "gpu_launches": int(launches), "kernels": kernels, "clocks": clk.summary(), "breakdown": brk}
        if sharded is not None:
            line["sharded"] = sharded
        if world == 1 and not args.no_cpu:
            os.sched_setaffinity(0, all_cpus)          # the CPU baseline uses every host core again
            nproc = os.cpu_count() or 1
            arm = cpu_arm(nproc)
            try:
                rl, rj = cpu_rows(arm, 8.0)            # ~8 s per leg on this box
                cb = cpu_step(arm, rl, rj)
                kind = cpu_kind(arm)
            finally:
                arm.close()
            line["cpu_baseline"] = {
                "value": cb["value"], "unit": UNIT, "cores": nproc, "kind": kind,
                "sample": "%d rows (linear) + %d rows (Jacobian-Gaussian) of the 8192^2 frame on one persistent pool of "
                          "%d forked workers; per-pixel rates combined for a full step" % (rl, rj, nproc),
                "linear_mpix_s": cb["linear_mpix_s"], "jacobian_mpix_s": cb["jacobian_mpix_s"]}
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        print(json.dumps(line))
        sys.stdout.flush()
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg (profiling runs)")
    ap.add_argument("--no-extra", action="store_true", help="skip the configs 2/4/5 breakdown (profiling runs)")
    args = ap.parse_args()
    if args.impl == "reference":
        return reference_arm(args)
    return gpu_arm(args)


if __name__ == "__main__":
    sys.exit(main())
