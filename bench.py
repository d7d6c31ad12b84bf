#!/usr/bin/env python3
"""bench.py -- BASELINE.json's metric: Mrays/s and ms/frame, Teapot BVH scene at 1080p.

  python bench.py --gpus N --steps K --warmup W            (our arm; N>1 under torchrun)
  python bench.py --impl reference --gpus N --steps K --warmup W   (the reference's CPU code)

A step = one Whitted frame (Blinn + shadow + reflection/refraction, 5 bounces, the reference's
Halton(4,5) sample pattern) of scenes/Teapot/scene2.xml at 1920x1080 with --spp samples per pixel
per GPU.  With N GPUs the frame has N*spp samples per pixel, rank r renders samples
[r*spp,(r+1)*spp) of every pixel (spp-sliced) and the FP32 accumulators are summed onto rank 0
with one NCCL reduce per frame: per-GPU work is fixed, so scaling is "weak".

value : rays/s of the whole job with the scene resident in HBM (device work only, CUDA events).
e2e   : same frame through the C ABI with HOST buffers: rtu_scene_upload (pack + H2D) then
        rtu_render -> host RGB8 + Z8 (D2H), every step.
A ray = one root-level Trace or ShadowTrace; the ray set is exactly the
reference recursion's (no culling flags).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "raytracer-utah_b200", "python"))

SCENE = "Teapot/scene2.xml"   # default workload; --scene / --mode select the GI workload of BASELINE config 4
MODE = "whitted"
HARNESS = os.path.join(ROOT, "oracle", "_ref", "ref_harness")
FALLBACK_HBM_GBS = 6650.0  # /opt/skills/guides/B200_PROFILING.md fallback
# dram bytes per launch of the dominant kernel from profiles/ (ncu --set full); None until captured
NCU_TRAFFIC_BYTES_PER_LAUNCH = 1144421000  # dram read+write of the wave-0 k_shadow_wave launch of a 64-spp frame (profiles/r01_final_ncu_summary.txt)


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f), "measured"
    except Exception:
        return {"hbm_gbs": FALLBACK_HBM_GBS}, "fallback"


class ClockSampler:
    """nvidia-smi clocks/throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.monotonic(), [c.strip() for c in line.split(",")]))

    def wait_alive(self, timeout=5.0):
        """nvidia-smi can take most of a second to print its first row: block until it has, so that a short timed
        region is not over before the sampler runs."""
        t0 = time.monotonic()
        while self.proc and not self.rows and time.monotonic() - t0 < timeout:
            time.sleep(0.01)

    def mark(self):
        return time.monotonic()

    def stop(self, t_begin=None, t_end=None):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        return self.summarise(self.rows, t_begin, t_end)

    @staticmethod
    def summarise(stamped_rows, t_begin=None, t_end=None):
        """Median SM clock, maximum clock and throttle reasons of the rows that arrived inside [t_begin, t_end]."""
        rows = [r for t, r in stamped_rows if (t_begin is None or t >= t_begin) and (t_end is None or t <= t_end + 0.05)]
        window = "timed region"
        if not rows:  # nothing fell inside the window: report what was sampled under the warm-up load instead
            rows, window = [r for _, r in stamped_rows], "warm-up + timed region"
        sm, mx, reasons = [], [], set()
        for r in rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "window": window}


def _mode_id(R):
    return R.MODE_PATH if MODE == "path" else R.MODE_WHITTED


def run_harness(width, height, spp, s0, s1, threads=0):
    cmd = [HARNESS, os.path.join(ROOT, "scenes", SCENE), "--root", os.path.join(ROOT, "scenes"), "--mode", "whitted",
           "--width", str(width), "--height", str(height), "--spp", str(spp), "--pattern", "ref",
           "--samples", str(s0), str(s1), "--threads", str(threads), "--out", "-"]
    r = subprocess.run(cmd, stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, text=True, check=True)
    return json.loads([l for l in r.stderr.splitlines() if l.startswith("{")][-1])


def run_port(width, height, spp, s0, s1):
    """The C restatement (oracle/liboracle.so) as CPU baseline when the reference binary is absent."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle_py
    import rtu_b200 as R
    hs = R.HostScene(os.path.join(R.SCENES, SCENE))
    p = R.default_params(width=width, height=height, spp=spp, pattern=R.PATTERN_REFERENCE, mode=_mode_id(R),
                         sample_begin=s0, sample_end=s1)
    o = oracle_py.render(hs.desc, params=p, want=("rgb",))
    st = o["stats"]
    rays = st["trace_rays"] + st["shadow_rays"]
    return {"rays": rays, "seconds": st["seconds"], "threads": st["threads"], "mrays_per_s": rays / st["seconds"] * 1e-6}


def cpu_sample(width, height, spp, target_seconds=12.0):
    """Times the reference's CPU implementation on a bounded sample of the workload."""
    # the reference binary's headless modes cover Whitted frames; its GI estimator is only reachable through
    # Render() at a fixed 1024 spp, so the GI workload is timed with the C port
    kind = "reference" if (os.path.exists(HARNESS) and MODE == "whitted") else "port"
    fn = run_harness if kind == "reference" else run_port
    probe = fn(width, height, spp, 0, 1)
    k = int(max(1, min(spp, target_seconds / max(probe["seconds"], 1e-3))))
    res = fn(width, height, spp, 0, k) if k > 1 else probe
    return kind, k, res


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    W, H, spp = args.width, args.height, args.spp
    kind, k, first = cpu_sample(W, H, spp, target_seconds=max(2.0, 60.0 / max(1, args.steps + args.warmup)))
    fn = run_harness if kind == "reference" else run_port
    for _ in range(max(0, args.warmup - 1)):
        fn(W, H, spp, 0, k)
    rays = secs = 0.0
    threads = first["threads"]
    for _ in range(args.steps):
        r = fn(W, H, spp, 0, k)
        rays += r["rays"]
        secs += r["seconds"]
    value = rays / secs * 1e-6
    sample = "samples [0,%d) of the %d-spp pattern, all %dx%d pixels, per step" % (k, spp, W, H)
    line = {"impl": "reference", "metric": "Mrays/s", "value": value, "unit": "Mrays/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": secs / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, 1, note="CPU arm: %s" % sample),
            "cpu_baseline": {"value": value, "unit": "Mrays/s", "cores": threads, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))
    return 0


def workload_config(args, n, note=None):
    what = "Whitted (Blinn + shadow + reflection/refraction, 5 bounces)" if MODE == "whitted" else \
        "HEAD estimator (4-bounce MonteCarlo GI + Whitted, RenderFunctions.cpp:129-135)"
    c = {"workload": "%s %dx%d %s, %d spp/GPU, reference Halton(4,5) pattern" % (SCENE, args.width, args.height, what, args.spp),
         "scene": SCENE, "width": args.width, "height": args.height, "spp_per_gpu": args.spp, "spp_total": args.spp * n,
         "parallelism": "spp-sliced x%d%s" % (n, " + ncclReduce(FP32 accum) to rank 0" if n > 1 else ""),
         "l2": "flushed (256 MiB device write) between timed steps",
         "ray_set": "identical to the reference recursion (no culling flags)"}
    if note:
        c["note"] = note
    return c


def ours(args):
    import numpy as np
    import torch
    import rtu_b200 as R

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; this renderer has no CPU path")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if os.environ.get("NCCL_DEBUG", "").upper() in ("VERSION", "WARN"):
            os.environ.pop("NCCL_DEBUG")  # these levels print NCCL's version banner on stdout; rank 0 prints ONE JSON line
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local))
    W, H, spp = args.width, args.height, args.spp
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    hs = R.HostScene(os.path.join(R.SCENES, SCENE))
    ctx = R.Context(local, stream.cuda_stream)
    sc = R.Scene(ctx, hs.desc)
    accum = torch.zeros(W * H * 4, dtype=torch.float32, device="cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    p = R.default_params(width=W, height=H, spp=spp * world, sample_begin=rank * spp, sample_end=(rank + 1) * spp,
                         pattern=R.PATTERN_REFERENCE, mode=_mode_id(R), shade_bounces=5, gi_bounces=4)

    def step():
        sc.render_device(p, accum.data_ptr(), clear=True)
        if world > 1:
            dist.reduce(accum, dst=0)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()  # started before the warm-up so that it is certainly printing when the timed region begins
    for _ in range(max(3, args.warmup)):
        step()
    barrier()
    if rank == 0:
        sampler.wait_alive()
    st = sc.stats()
    rays_rank = st["trace_rays"] + st["shadow_rays"]
    launches_step = st["kernel_launches"]

    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    t_begin = sampler.mark()
    for a, b in ev:
        flush.zero_()          # L2 flush, outside the timed bracket
        a.record()
        step()
        b.record()
    barrier()
    t_end = sampler.mark()
    clocks = sampler.stop(t_begin, t_end) if rank == 0 else None
    ms_total = sum(a.elapsed_time(b) for a, b in ev)
    t = torch.tensor([ms_total, float(rays_rank)], dtype=torch.float64, device="cuda")
    if world > 1:
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone()
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        ms_total, rays_all = float(tmax[0]), float(tsum[1])
    else:
        rays_all = float(rays_rank)
    value = rays_all * args.steps / (ms_total * 1e-3) * 1e-6

    # ---- e2e: pack + H2D + render + (reduce) + resolve + D2H to host buffers, every step
    brk = {"upload": 0.0, "render_and_readback": 0.0, "destroy": 0.0}

    # page-locked host buffers for the frame's outputs (Result.png and ZBuffer.png pixels)
    pinned = {"rgb8": torch.empty((H, W, 3), dtype=torch.uint8, pin_memory=True).numpy(),
              "z8": torch.empty((H, W), dtype=torch.uint8, pin_memory=True).numpy()}

    def e2e_step():
        t_a = time.perf_counter()
        s2 = R.Scene(ctx, hs.desc)
        t_b = time.perf_counter()
        if world > 1:
            s2.render_device(p, accum.data_ptr(), clear=True)
            dist.reduce(accum, dst=0)
            out = s2.resolve(p, accum.data_ptr(), want=("rgb8", "z8")) if rank == 0 else None
        else:
            out = s2.render(p, want=("rgb8", "z8"), out=pinned)
        t_c = time.perf_counter()
        nbytes = s2.stats()["scene_device_bytes"]
        s2.close()
        t_d = time.perf_counter()
        brk["upload"] += t_b - t_a
        brk["render_and_readback"] += t_c - t_b
        brk["destroy"] += t_d - t_c
        return out, nbytes

    e2e_step()
    barrier()
    for k in brk:
        brk[k] = 0.0
    t0 = time.perf_counter()
    h2d = 0
    for _ in range(args.steps):
        _, h2d = e2e_step()
    barrier()
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = rays_all * args.steps / float(te[0]) * 1e-6

    # ---- roofline of the dominant kernel: one extra frame with per-launch CUDA events
    pk = R.default_params(width=W, height=H, spp=spp * world, sample_begin=rank * spp, sample_end=(rank + 1) * spp,
                          pattern=R.PATTERN_REFERENCE, mode=_mode_id(R), shade_bounces=5, gi_bounces=4, flags=R.FLAG_TIME_KERNELS)
    sc.render_device(pk, accum.data_ptr(), clear=True)
    ks = sc.stats()
    line = None
    if rank == 0:
        peaks, peak_src = measured_peaks()
        classes = {"k_extend<primary>": ks["primary_wave"], "k_extend<queue>": ks["secondary_waves"], "k_shadow_wave": ks["shadow_waves"],
                   "k_shade": ks["shade_kernels"]}
        name, dom = max(((k, v) for k, v in classes.items() if k != "k_shade"), key=lambda kv: kv[1]["ms"])
        nl = max(1, dom["launches"])
        # algorithmic bytes: 28 B per child-box test, 52 B per triangle test, 48 B per node transform
        alg_bytes = 28 * dom["box_tests"] + 52 * dom["tri_tests"] + 48 * dom["node_visits"]
        # useful launches only: empty waves are launched but do nothing
        ach = alg_bytes / (dom["ms"] * 1e-3) / 1e9 if dom["ms"] > 0 else 0.0
        sm_mhz = (clocks or {}).get("sm_mhz") or 1500.0
        flops = 21 * dom["box_tests"] + 70 * dom["tri_tests"] + 36 * dom["node_visits"]
        fp32_peak = 148 * 128 * sm_mhz * 1e6  # FADD/FMUL issue rate; -fmad=false so no FMA doubling
        roof = {"bound": "hbm", "kernel": name, "achieved": ach, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                "frac": ach / peaks["hbm_gbs"], "peak_source": peak_src,
                "traffic": NCU_TRAFFIC_BYTES_PER_LAUNCH,
                "alg_bytes_per_launch": alg_bytes / nl, "launches_per_step": dom["launches"], "kernel_ms_per_step": dom["ms"],
                "kernel_share_of_step": dom["ms"] / max(1e-9, sum(c["ms"] for c in classes.values())),
                "rays_per_step": dom["rays"], "kernel_mrays_per_s": dom["rays"] / (dom["ms"] * 1e-3) * 1e-6 if dom["ms"] > 0 else 0.0,
                "fp32": {"achieved_tflops": flops / (dom["ms"] * 1e-3) / 1e12 if dom["ms"] > 0 else 0.0,
                         "peak_tflops_no_fma": fp32_peak / 1e12,
                         "frac": (flops / (dom["ms"] * 1e-3)) / fp32_peak if dom["ms"] > 0 else 0.0, "note": "the Teapot BVH (0.6 MB) is L1/L2 resident: latency/issue bound, not HBM bound"},
                "all_kernels_ms": {k: v["ms"] for k, v in classes.items()}}
        line = {"metric": "Mrays/s", "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
                "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic", "config": workload_config(args, world),
                "clocks": clocks,
                "e2e": {"value": e2e_value, "unit": "Mrays/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(W * H * 4),
                        "ms_per_step": float(te[0]) / args.steps * 1e3,
                        "breakdown_ms_per_step": {k: v / args.steps * 1e3 for k, v in brk.items()}, "timer": "host perf_counter around a synchronized region (includes scene packing on the host)"},
                "gpu_launches": int(launches_step * args.steps),
                "rays_per_step": rays_all,
                "roofline": roof}
        if world == 1 and not args.no_cpu:
            kind, k, res = cpu_sample(W, H, spp)
            line["cpu_baseline"] = {"value": res["mrays_per_s"], "unit": "Mrays/s", "cores": res["threads"], "kind": kind,
                                    "sample": "samples [0,%d) of the %d-spp pattern, all %dx%d pixels (%.1f s)" % (k, spp, W, H, res["seconds"])}
        print(json.dumps(line))
    sc.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    global SCENE, MODE
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--spp", type=int, default=64)
    ap.add_argument("--width", type=int, default=1920)
    ap.add_argument("--height", type=int, default=1080)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--scene", default=SCENE, help="scene file under scenes/ (default: the Teapot headline scene)")
    ap.add_argument("--mode", default="whitted", choices=["whitted", "path"], help="path = HEAD estimator with Monte-Carlo GI")
    args = ap.parse_args()
    SCENE, MODE = args.scene, args.mode
    if args.impl == "reference":
        return reference_arm(args)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.gpus != world and world == 1 and args.gpus > 1:
        # convenience: re-launch under torchrun
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(args.gpus), "--master-addr", "127.0.0.1",
               "--master-port", "29511", os.path.abspath(__file__)] + sys.argv[1:]
        return subprocess.call(cmd)
    return ours(args)


if __name__ == "__main__":
    sys.exit(main())
