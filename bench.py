#!/usr/bin/env python
"""bench.py -- headline benchmark of the libiqo resize hot path on B200.

Workload (BASELINE.json configs[3], the configuration the metric is quoted on):
    a batch of independent 1920x1080 U8 frames, Lanczos3 -> 960x540, frames sharded over the
    GPUs (one process per GPU, no collective: nothing reduces across devices).
One "step" = one pass of the hot path over the whole per-GPU batch (4096 frames, 8.5 GB in +
2.1 GB out per GPU: far larger than the 126 MB L2, so no L2 flush is needed between steps).

    python bench.py [--gpus N] [--steps K] [--warmup W]            # our CUDA arm
    python bench.py --impl reference [--gpus N] [--steps K] ...    # the reference's CPU path

One JSON line on stdout (rank 0).  `value` is device-resident throughput (CUDA events, max
over ranks); `e2e` is the same metric through the host-pointer C-ABI call
(iqo_cuda_resize_batch_host) with the H2D / D2H copies inside the timed region.
"""
import argparse
import ctypes
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

WORKLOADS = {
    # name: (kind, degree, pxScale, srcW, srcH, dstW, dstH, default frames per GPU)
    "cfg4_lanczos3_1080p_to_540p": (0, 3, 1, 1920, 1080, 960, 540, 4096),
    "cfg1_lanczos3_1080p_to_720p": (0, 3, 1, 1920, 1080, 1280, 720, 2048),
    "cfg3y_lanczos2_2160p_to_1080p": (0, 2, 1, 3840, 2160, 1920, 1080, 768),
    "cfg3uv_lanczos2_px2_1080p_to_540p": (0, 2, 2, 1920, 1080, 960, 540, 4096),
    "cfg2a_area_2160p_to_1080p": (1, 0, 1, 3840, 2160, 1920, 1080, 768),
    "cfg2b_linear_720p_to_2160p": (2, 0, 1, 1280, 720, 3840, 2160, 768),
    # cfg5's ratio (1024:375, Lanczos4, 22 taps, 375 phases) on a size that fits a batch
    "cfg5s_lanczos4_8192_to_3000": (0, 4, 1, 8192, 8192, 3000, 3000, 16),
}
DEFAULT_WORKLOAD = "cfg4_lanczos3_1080p_to_540p"
METRIC = "dst_mpix_per_s_lanczos3_u8_resize"
UNIT = "Mpix/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS))
    ap.add_argument("--frames", type=int, default=0, help="frames per GPU (default: workload's)")
    ap.add_argument("--e2e-frames", type=int, default=1024, help="frames per end-to-end step")
    ap.add_argument("--e2e-steps", type=int, default=0, help="default: min(steps, 5)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=10.0, help="budget of the cpu_baseline sample")
    ap.add_argument("--path", default="auto", choices=["auto", "generic", "no_tma", "no_stream", "stream"])
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: every GPU gets the workload's frame count; strong: that count is sharded over the GPUs")
    return ap.parse_args()


# ----------------------------------------------------------------------------- parity of the timed batch

def golden_positions(frames):
    return sorted(set([0, frames // 2, frames - 1]))


def plant_golden_frames(src, work):
    """Overwrite three frames of the (frames, sh, sw) batch with the SURVEY 8c LCG image (seed 1), whose
    reference output hash is committed in tests/golden/cases.json.  Returns the hash (hex) or None when the
    workload has no golden vector.  Uses libiqo_b200.vectors (numpy), nothing under oracle/."""
    import torch
    from libiqo_b200 import vectors
    kind, deg, px, sw, sh, dw, dh, _ = work
    want = vectors.golden_hash(kind, deg, px, sw, sh, dw, dh, seed=1)
    if want is None or src.shape[0] == 0:
        return None
    frame = torch.from_numpy(vectors.lcg_image(sh, sw, seed=1)).to(src.device)
    for f in golden_positions(src.shape[0]):
        src[f].copy_(frame)
    return want


def check_golden_frames(dst, want):
    """The planted frames of the timed batch against the reference's Generic output: the first one by its
    FNV-1a-64 hash, the others by equality with the first (same source, same result)."""
    import torch
    from libiqo_b200 import vectors
    pos = golden_positions(dst.shape[0])
    got = "%016x" % vectors.fnv1a64(dst[pos[0]].cpu().numpy())
    same = all(bool(torch.equal(dst[f], dst[pos[0]])) for f in pos[1:])
    ok = (got == want) and same
    return {"frames_checked": len(pos), "mismatches": 0 if ok else None, "max_abs_diff": 0 if ok else None,
            "fnv1a64": got, "golden": want, "bit_exact": ok,
            "against": "reference Generic output hash (tests/golden/cases.json, LCG source seed 1 planted in the timed batch)"}


# ----------------------------------------------------------------------------- CPU arms

def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def time_cpu(work, seconds, steps=None, warmup=1, max_frames=256):
    """Time the reference's CPU implementation of the path on a bounded sample of `work`.
    Returns dict(value Mpix/s, cores, kind, sample, ms_per_step).  Uses oracle/_ref's full
    reference library (SIMD dispatch + OpenMP, all host threads) when it is present, otherwise
    the scalar oracle port.  This is the checker being *timed as a baseline*, never the product."""
    import numpy as np
    import oracle_lib as O
    kind, deg, px, sw, sh, dw, dh, _ = work
    full = O.ref_full()
    cores = host_cores()
    if full is not None:
        full.iqo_ref_set_threads(cores)   # torchrun exports OMP_NUM_THREADS=1
        h = full.iqo_ref_public_new(kind, deg, sw, sh, dw, dh, px)

        def run(n, src, dst):
            full.iqo_ref_public_resize_batch(h, n, sw, sw * sh, src.ctypes.data, dw, dw * dh, dst.ctypes.data)
        label, used = "reference", full.iqo_ref_threads()
    else:
        def run(n, src, dst):
            rc = O.oracle().iqo_oracle_resize_batch(kind, deg, sw, sh, dw, dh, px, n, sw, sw * sh,
                                                    src.ctypes.data, dw, dw * dh, dst.ctypes.data)
            assert rc == 0
        label, used = "port", 1
    # calibrate the sample so that one step is ~seconds/(steps+warmup)
    probe = 4
    src = np.stack([O.lcg_image(sh, sw, seed=1 + f) for f in range(probe)])
    dst = np.zeros((probe, dh, dw), dtype=np.uint8)
    run(probe, src, dst)
    t = time.perf_counter()
    run(probe, src, dst)
    per_frame = (time.perf_counter() - t) / probe
    nsteps = steps if steps else 3
    frames = int(max(1, min(max_frames, seconds / max(per_frame, 1e-9) / (nsteps + warmup))))
    reps = -(-frames // probe)
    src = np.concatenate([src] * reps)[:frames].copy()
    dst = np.zeros((frames, dh, dw), dtype=np.uint8)
    for _ in range(warmup):
        run(frames, src, dst)
    times = []
    for _ in range(nsteps):
        t = time.perf_counter()
        run(frames, src, dst)
        times.append(time.perf_counter() - t)
    if full is not None:
        full.iqo_ref_public_delete(h)
    mean = sum(times) / len(times)
    return dict(value=round(frames * dw * dh / mean / 1e6, 1), unit=UNIT, cores=used, kind=label,
                sample="%d frames/step x %d steps (+%d warm-up) of %dx%d->%dx%d, host memory, %s"
                       % (frames, nsteps, warmup, sw, sh, dw, dh,
                          "reference public API, CPUID dispatch + OpenMP (oracle/_ref)" if label == "reference"
                          else "scalar oracle port (oracle/iqo_oracle.c)"),
                ms_per_step=mean * 1e3, frames=frames)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    os.environ["OMP_NUM_THREADS"] = str(host_cores())
    work = WORKLOADS[args.workload]
    kind, deg, px, sw, sh, dw, dh, _ = work
    res = time_cpu(work, seconds=60.0, steps=args.steps, warmup=args.warmup, max_frames=256)
    line = {
        "impl": "reference", "metric": METRIC, "value": res["value"], "unit": UNIT,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": round(res["ms_per_step"], 3), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "int32", "data": "synthetic",
        "config": {"workload": args.workload, "frames_per_step": res["frames"],
                   "src": [sw, sh], "dst": [dw, dh], "device": "host CPU"},
        "cpu_baseline": {k: res[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": res["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# ----------------------------------------------------------------------------- clocks

class ClockSampler(object):
    """Samples SM clock and throttle reasons with NVML while the timed regions run."""

    def __init__(self, index, period=0.02):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._active = threading.Event()
        self._thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self._nv = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM)
            self._period = period
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()
        except Exception as exc:  # pragma: no cover
            self.error = repr(exc)

    _NAMES = (("hw_slowdown", 0x8), ("sw_thermal_slowdown", 0x20), ("hw_thermal_slowdown", 0x40),
              ("hw_power_brake_slowdown", 0x80), ("sw_power_cap", 0x4), ("sync_boost", 0x10))

    def _run(self):
        nv = self._nv
        while not self._stop.is_set():
            if self._active.is_set():
                try:
                    self.samples.append(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM))
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self._h)
                    for name, bit in self._NAMES:
                        if mask & bit:
                            self.reasons.add(name)
                except Exception:
                    pass
            time.sleep(self._period)

    def start(self):
        self._active.set()

    def pause(self):
        self._active.clear()

    def summary(self):
        self._stop.set()
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(s)}


def visible_physical_index(local_rank):
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis:
        try:
            return int(vis.split(",")[local_rank])
        except Exception:
            return local_rank
    return local_rank


# ----------------------------------------------------------------------------- CUDA arm

def run_cuda(args):
    # libraries (NCCL's version banner, ...) may print to stdout: keep fd 1 for the JSON line only
    json_fd = os.dup(1)
    os.dup2(2, 1)
    try:
        return _run_cuda(args, json_fd)
    finally:
        os.close(json_fd)


def _run_cuda(args, json_fd):
    import numpy as np
    import torch
    import libiqo_b200 as iqo

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU path")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    iqo.lib()  # fail loudly if the CUDA library is missing

    work = WORKLOADS[args.workload]
    kind, deg, px, sw, sh, dw, dh, default_frames = work
    frames = args.frames or default_frames
    if args.scaling == "strong":
        from libiqo_b200.sharding import frame_shard
        frames = frame_shard(frames, world, rank)[1]
    dev = torch.device("cuda", local)
    gen = torch.Generator(device=dev)
    gen.manual_seed(1234 + rank)
    src = torch.randint(0, 256, (frames, sh, sw), dtype=torch.uint8, device=dev, generator=gen)
    try:
        golden = plant_golden_frames(src, work)   # three frames with a known reference result (parity below)
    except Exception as e:  # the parity report must never cost the measurement
        golden = None
        print("bench: golden frames not planted: %r" % (e,), file=sys.stderr)
    dst = torch.zeros((frames, dh, dw), dtype=torch.uint8, device=dev)
    r = iqo.make_resizer(kind, deg, sw, sh, dw, dh, px, device=local)
    r.set_path({"generic": iqo.PATH_GENERIC, "no_tma": iqo.PATH_NO_TMA, "no_stream": iqo.PATH_NO_STREAM,
                "stream": iqo.PATH_STREAM}.get(args.path, iqo.PATH_AUTO))
    stream = torch.cuda.current_stream().cuda_stream

    def step():
        r.resize_batch(frames, sw, sw * sh, src, dw, dw * dh, dst, stream)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(visible_physical_index(local)) if rank == 0 else None

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    launches0 = iqo.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if sampler:
        sampler.start()
    barrier()
    ev0.record()
    for _ in range(args.steps):
        step()
    ev1.record()
    barrier()
    if sampler:
        sampler.pause()
    launches = iqo.launch_count() - launches0
    ms_total = ev0.elapsed_time(ev1)
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    ms_step = ms_total / args.steps
    nframes_all = torch.tensor([float(frames)], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(nframes_all)
    total_px = float(nframes_all.item()) * dw * dh
    value = total_px / (ms_step * 1e-3) / 1e6
    kernel_name = r.last_kernel()

    # ---- parity of the timed configuration (outside the timed region): the planted frames of the batch the
    # timed launches wrote, against the committed hash of the reference's output (no oracle on this arm)
    parity = None
    if rank == 0:
        try:
            parity = check_golden_frames(dst, golden) if golden else {"frames_checked": 0, "against": "no golden vector for this workload"}
        except Exception as e:
            parity = {"frames_checked": 0, "error": repr(e)}

    # ---- end to end: host (pinned) buffers through the public host-pointer call
    e2e = None
    if not args.no_e2e:
        ef = min(args.e2e_frames, frames)
        esteps = args.e2e_steps or min(args.steps, 5)
        hs = torch.empty((ef, sh, sw), dtype=torch.uint8, pin_memory=True)
        hd = torch.empty((ef, dh, dw), dtype=torch.uint8, pin_memory=True)
        hs.copy_(src[:ef])
        torch.cuda.synchronize()
        for _ in range(2):
            r.resize_batch_host(ef, sw, sw * sh, hs, dw, dw * dh, hd)
        barrier()
        if sampler:
            sampler.start()
        t0 = time.perf_counter()
        for _ in range(esteps):
            r.resize_batch_host(ef, sw, sw * sh, hs, dw, dw * dh, hd)   # returns after the D2H completed
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if sampler:
            sampler.pause()
        t = torch.tensor([dt], dtype=torch.float64, device=dev)
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
        e2e_ok = bool(torch.equal(hd.to(dev), dst[:ef]))
        e2e = {"value": round(float(ef) * dw * dh * world * esteps / dt / 1e6, 1), "unit": UNIT,
               "h2d_bytes_per_step": ef * sw * sh, "d2h_bytes_per_step": ef * dw * dh,
               "frames_per_step": ef, "steps": esteps, "ms_per_step": round(dt / esteps * 1e3, 3),
               "api": "iqo_cuda_resize_batch_host (pinned host buffers, double-buffered H2D/kernel/D2H)",
               "matches_device_run": e2e_ok}

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return 0

    # ---- roofline of the dominant (only) kernel
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    launches_per_step = max(1, launches // args.steps)
    alg_bytes = float(frames) * (sw * sh + dw * dh) / launches_per_step
    launch_ms = ms_step / launches_per_step
    achieved = alg_bytes / (launch_ms * 1e-3) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        try:
            traffic = json.load(open(tp)).get(args.workload, {}).get(kernel_name)
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
                "frac": round(achieved / peak, 4), "traffic": traffic, "peak_source": peak_src,
                "kernel": kernel_name, "launches_per_step": launches_per_step,
                "algorithmic_bytes_per_launch": alg_bytes, "avg_launch_ms": round(launch_ms, 4),
                "macs_per_dst_px": None}
    try:
        # compute side: all taps counted, zeros included; the scalar-IMAD peak is the measured
        # 64 lanes/clk/SM (profiles/r1_microbench_pipe_rates.txt) at the sampled SM clock.  dp4a/dp2a do 4/2 MACs
        # per lane-instruction and mirrored taps are pre-added, so the achieved MAC rate may exceed that peak.
        ny = iqo.plan_query(kind, deg, sw, sh, dw, dh, px, 1)["numCoefs"] if sh != dh else 1
        nx = iqo.plan_query(kind, deg, sw, sh, dw, dh, px, 0)["numCoefs"] if sw != dw else 1
        macs = (ny * sw * dh + nx * dw * dh) / float(dw * dh)
        roofline["macs_per_dst_px"] = round(macs, 2)
        roofline["achieved_tmac_s"] = round(value * 1e6 * macs / 1e12, 2)
        roofline["imad_peak_tmac_s"] = round(148 * 64 * 1.965e9 / 1e12, 2)
    except Exception:
        pass

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        res = time_cpu(work, seconds=args.cpu_seconds)
        cpu = {k: res[k] for k in ("value", "unit", "cores", "kind", "sample")}

    line = {
        "metric": METRIC, "value": round(value, 1), "unit": UNIT, "n_gpus": world,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": round(ms_step, 4),
        "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
        "dtype": "int32", "data": "synthetic",
        "config": {"workload": args.workload, "frames_per_gpu": frames, "src": [sw, sh], "dst": [dw, dh],
                   "kernel": kernel_name, "sharding": "independent frames per GPU, no collective",
                   "cache": "inputs %.1f GB per GPU exceed the 126 MB L2 (no flush needed)"
                            % (frames * sw * sh / 1e9),
                   "timing": "CUDA events on the launch stream, max over ranks"},
        "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches),
        "clocks": sampler.summary() if sampler else None, "parity": parity,
    }
    os.write(json_fd, (json.dumps(line) + "\n").encode())
    if dist is not None:
        dist.destroy_process_group()
    return 0


def main():
    args = parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_cuda(args)


if __name__ == "__main__":
    sys.exit(main())
