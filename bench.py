#!/usr/bin/env python3
"""bench.py -- BASELINE.json's metric: Mrays/s and ms/frame, Teapot BVH scene at 1080p, 1/2/4/8 B200 vs host-CPU ref.

  python bench.py --gpus N --steps K --warmup W            (our arm; N>1 under torchrun)
  python bench.py --impl reference --gpus N --steps K --warmup W   (the reference's CPU code)

A step = one Whitted frame (Blinn + shadow + reflection/refraction, 5 bounces, the reference's Halton(4,5) sample pattern)
of scenes/Teapot/scene2.xml at 1920x1080 with --spp-total samples per pixel (default 1024 = the reference's own
maxSampleSize, RenderFunctions.cpp:27).  The sample budget is FIXED: with N GPUs rank r renders samples
[r*S/N,(r+1)*S/N) of every pixel (partitioning B) through rtu_render_device and the partial images are brought
together by rtu_reduce_resolve (RGB planes, one ncclReduce onto rank 0 on the render stream, resolve on rank 0): scaling is
"strong".  --scaling weak keeps --spp samples per GPU instead (round-1 behaviour).

value : rays/s of the whole job with the scene resident in HBM (device work only, CUDA events on the render stream, max
        over ranks); for N>1 the timed step includes the reduce.
e2e   : the same frame through the C ABI with HOST buffers, every step: rtu_scene_upload (pack + H2D), render,
        (reduce,) resolve, D2H of RGB8 + Z8 into page-locked buffers.
gi    : BASELINE config 4 in the same run: Project11/scene.xml 800x600, HEAD estimator (4-bounce Monte-Carlo GI + Whitted,
        RenderFunctions.cpp:129-135), same sample budget, spp-sliced the same way; with N>1 rank 0 also renders the whole
        budget alone, so the line carries the N-GPU efficiency of the GI scene.
A ray = one root-level Trace or ShadowTrace; the ray set is exactly the reference recursion's.
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

SCENE = "Teapot/scene2.xml"   # default workload; --scene / --mode select another one
MODE = "whitted"
GI_SCENE, GI_SIZE = "Project11/scene.xml", (800, 600)
REF_DIR = os.path.join(ROOT, "oracle", "_ref")
HARNESS = os.path.join(REF_DIR, "ref_harness")
FALLBACK_HBM_GBS = 6650.0  # /opt/skills/guides/B200_PROFILING.md fallback
NCU_TABLE = os.path.join(ROOT, "profiles", "ncu_table.json")  # written by tools/ncu_summary.py --json from ncu --set full captures


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


def _mode_id(R, mode=None):
    return R.MODE_PATH if (mode or MODE) == "path" else R.MODE_WHITTED


# ------------------------------------------------------------------------------------------------ CPU legs
def run_harness(width, height, spp, s0, s1, threads=0, binary=HARNESS, scene=None):
    cmd = [binary, os.path.join(ROOT, "scenes", scene or SCENE), "--root", os.path.join(ROOT, "scenes"), "--mode", "whitted",
           "--width", str(width), "--height", str(height), "--spp", str(spp), "--pattern", "ref",
           "--samples", str(s0), str(s1), "--threads", str(threads), "--out", "-"]
    r = subprocess.run(cmd, stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, text=True, check=True)
    return json.loads([l for l in r.stderr.splitlines() if l.startswith("{")][-1])


def run_port(width, height, spp, s0, s1, threads=0, scene=None, mode=None):
    """The C restatement (oracle/liboracle.so) as CPU baseline when the reference binary cannot run the workload."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle_py
    import rtu_b200 as R
    hs = R.HostScene(os.path.join(R.SCENES, scene or SCENE))
    p = R.default_params(width=width, height=height, spp=spp, pattern=R.PATTERN_REFERENCE, mode=_mode_id(R, mode),
                         sample_begin=s0, sample_end=s1)
    o = oracle_py.render(hs.desc, params=p, want=("rgb",), threads=threads or None)
    st = o["stats"]
    rays = st["trace_rays"] + st["shadow_rays"]
    return {"rays": rays, "seconds": st["seconds"], "threads": st["threads"], "mrays_per_s": rays / st["seconds"] * 1e-6}


def cpu_sample(width, height, spp, target_seconds=12.0, threads=0, binary=HARNESS):
    """Times the reference's CPU implementation on a bounded sample of the workload: samples [0,k) of the pattern."""
    # the reference binary's headless modes cover Whitted frames; its GI estimator is only reachable through
    # Render() at a fixed 1024 spp, so a GI workload is timed with the C port
    kind = "reference" if (os.path.exists(binary) and MODE == "whitted") else "port"
    fn = (lambda *a, **k: run_harness(*a, binary=binary, **k)) if kind == "reference" else run_port
    probe = fn(width, height, spp, 0, 1, threads=threads)
    k = int(max(1, min(spp, target_seconds / max(probe["seconds"], 1e-3))))
    res = fn(width, height, spp, 0, k, threads=threads) if k > 1 else probe
    return kind, k, res


def cpu_variants(width, height, spp, seconds_each=5.0):
    """The CPU figures SURVEY 8(d) / BASELINE.md section 3 ask for, each on a bounded sample: 1 thread and N threads of the
    build as it is (glibc's rand() lock makes the two close on scenes whose Shade calls SampleSphere), N threads with a
    thread-local rand() interposed at link time, and that again built -O3 -march=x86-64-v3."""
    out = {}
    for name, binary, threads in (("1_thread", HARNESS, 1), ("n_threads", HARNESS, 0),
                                  ("n_threads_tlrand", os.path.join(REF_DIR, "ref_harness_tlrand"), 0),
                                  ("n_threads_tlrand_O3_avx2", os.path.join(REF_DIR, "ref_harness_o3"), 0)):
        if not os.path.exists(binary) or MODE != "whitted":
            continue
        try:
            _, k, res = cpu_sample(width, height, spp, target_seconds=seconds_each, threads=threads, binary=binary)
            out[name] = {"mrays_per_s": res["mrays_per_s"], "threads": res["threads"], "samples": k, "seconds": res["seconds"]}
        except Exception as e:  # a variant binary that cannot run on this host (instruction set) is reported, not fatal
            out[name] = {"error": str(e)[:120]}
    return out


def sample_budget(args, world):
    """(samples per pixel of the frame, samples per rank)"""
    if args.scaling == "strong":
        total = args.spp_total
        if total % world:
            raise SystemExit("bench.py: --spp-total must be a multiple of the number of GPUs")
        return total, total // world
    return args.spp * world, args.spp


def workload_config(args, n):
    total, per = sample_budget(args, n)
    what = "Whitted (Blinn + shadow + reflection/refraction, 5 bounces)" if MODE == "whitted" else \
        "HEAD estimator (4-bounce MonteCarlo GI + Whitted, RenderFunctions.cpp:129-135)"
    return {"workload": "%s %dx%d %s, %d spp per frame, reference Halton(4,5) pattern" % (SCENE, args.width, args.height, what, total),
            "scene": SCENE, "width": args.width, "height": args.height, "spp_total": total, "spp_per_gpu": per,
            "parallelism": "spp-sliced x%d%s" % (n, " + rtu_reduce_resolve (ncclReduce of FP32 RGB planes to rank 0)" if n > 1 else ""),
            "l2": "flushed (256 MiB device write) between timed steps",
            "ray_set": "identical to the reference recursion (no culling flags)"}


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    world = int(os.environ.get("WORLD_SIZE", "1"))
    W, H = args.width, args.height
    spp, _ = sample_budget(args, max(1, world))
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
            "warmup": args.warmup, "ms_per_step": secs / args.steps * 1e3, "higher_is_better": True, "scaling": args.scaling,
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, max(1, world)),
            "cpu_baseline": {"value": value, "unit": "Mrays/s", "cores": threads, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------------------ roofline
def ncu_entry(scene, mode, kernel):
    """The ncu --set full capture of `kernel` on (scene, mode), if profiles/ holds one; None otherwise."""
    try:
        with open(NCU_TABLE) as f:
            table = json.load(f)
    except Exception:
        return None
    return table.get("%s:%s" % (scene, mode), {}).get(kernel)


def roofline(ks, clocks, scene, mode, scene_bytes, ref=None):
    peaks, peak_src = measured_peaks()
    names = {"k_extend<primary>": "primary_wave", "k_extend<queue>": "secondary_waves", "k_shadow_wave": "shadow_waves", "k_shade": "shade_kernels"}
    classes = {k: ks[v] for k, v in names.items()}
    name, dom = max(((k, v) for k, v in classes.items() if k != "k_shade"), key=lambda kv: kv[1]["ms"])
    nl = max(1, dom["launches"])
    # bytes of the tests the kernel executed (its own counters): 28 B per child-box test, 52 B per triangle test, 48 B per node transform
    exe_bytes = 28 * dom["box_tests"] + 52 * dom["tri_tests"] + 48 * dom["node_visits"]
    exe = exe_bytes / (dom["ms"] * 1e-3) / 1e9 if dom["ms"] > 0 else 0.0
    # algorithmic bytes: the same formula on the reference's own walk, per ray of this scene, x the rays of the step
    rc = ref[names[name]] if ref else None
    if rc and rc["rays"] > 0:
        per_ray = (28 * rc["box_tests"] + 52 * rc["tri_tests"] + 48 * rc["node_visits"]) / rc["rays"]
        alg_bytes = per_ray * dom["rays"]
        alg_src = "reference walk (RTU_FLAG_REFERENCE_WALK counters of %d rays) x rays of the step" % rc["rays"]
    else:
        per_ray = exe_bytes / max(1, dom["rays"])
        alg_bytes = exe_bytes
        alg_src = "the kernel's own counters (no reference walk for this scene)"
    ach = alg_bytes / (dom["ms"] * 1e-3) / 1e9 if dom["ms"] > 0 else 0.0
    sm_mhz = (clocks or {}).get("sm_mhz") or 1500.0
    flops = 21 * dom["box_tests"] + 70 * dom["tri_tests"] + 36 * dom["node_visits"]
    fp32_peak = 148 * 128 * sm_mhz * 1e6  # FADD/FMUL issue rate; -fmad=false so no FMA doubling
    cap = ncu_entry(scene, mode, name)
    traffic = cap.get("dram_bytes_per_launch") if cap else None
    # Which ceiling binds is read off the capture: DRAM traffic far below the algorithmic bytes means the BVH is served by
    # the caches; with most sectors hitting L1 the kernel is bound by instruction issue / dependent-load latency, otherwise by
    # L2.  Without a capture of this (scene, kernel) the scene's footprint against the 126 MB L2 decides, and says so.
    if cap and traffic is not None:
        if (cap.get("dram_gbs") or 0.0) >= 0.5 * ach:   # DRAM rate of the captured launch vs the algorithmic rate
            bound = "hbm"
        elif cap.get("l1_hit_pct", 0) >= 70.0:
            bound = "issue"
        else:
            bound = "l2"
        bound_source = "ncu capture %s" % cap.get("source", "profiles/")
    else:
        bound = "issue" if scene_bytes < 16 << 20 else ("l2" if scene_bytes < 126 << 20 else "hbm")
        bound_source = "no ncu capture for this (scene, kernel): scene footprint %.1f MB vs L1/L2 capacity" % (scene_bytes / 1e6)
    roof = {"bound": bound, "bound_source": bound_source, "kernel": name, "achieved": ach, "peak": peaks["hbm_gbs"], "unit": "GB/s",
            "frac": ach / peaks["hbm_gbs"], "peak_source": peak_src, "alg_gbs": ach,
            "traffic": traffic,
            "alg_bytes_per_launch": alg_bytes / nl, "alg_bytes_per_ray": per_ray, "alg_source": alg_src,
            "executed_gbs": exe, "executed_bytes_per_ray": exe_bytes / max(1, dom["rays"]),
            "note": "achieved = the reference algorithm's bytes per ray x rays / kernel time; the kernel reaches the same answers with "
                    "fewer tests (executed_*), from caches (traffic = DRAM bytes per launch under ncu): frac can pass 1, the binding limit is `bound`",
            "launches_per_step": dom["launches"], "kernel_ms_per_step": dom["ms"],
            "kernel_share_of_step": dom["ms"] / max(1e-9, sum(c["ms"] for c in classes.values())),
            "rays_per_step": dom["rays"], "kernel_mrays_per_s": dom["rays"] / (dom["ms"] * 1e-3) * 1e-6 if dom["ms"] > 0 else 0.0,
            "issue": ({"issue_slot_frac": cap.get("issue_pct", 0) / 100.0, "lanes_per_inst": cap.get("threads_per_inst"),
                       "ipc_per_sm": cap.get("ipc_per_sm"), "l1_hit_pct": cap.get("l1_hit_pct"), "l2_hit_pct": cap.get("l2_hit_pct"),
                       "dram_gbs": cap.get("dram_gbs")} if cap else None),
            "fp32": {"achieved_tflops": flops / (dom["ms"] * 1e-3) / 1e12 if dom["ms"] > 0 else 0.0,
                     "peak_tflops_no_fma": fp32_peak / 1e12,
                     "frac": (flops / (dom["ms"] * 1e-3)) / fp32_peak if dom["ms"] > 0 else 0.0},
            "all_kernels_ms": {k: v["ms"] for k, v in classes.items()}}
    return roof


# ------------------------------------------------------------------------------------------------ our arm
class Job:
    """One workload (scene, size, mode, sample budget) on this rank's GPU."""

    def __init__(self, R, ctx, comm, rank, world, scene, width, height, mode, spp_total, spp_rank):
        self.R, self.ctx, self.comm, self.rank, self.world = R, ctx, comm, rank, world
        self.W, self.H = width, height
        self.hs = R.HostScene(os.path.join(R.SCENES, scene))
        self.sc = R.Scene(ctx, self.hs.desc)
        self.spp_total = spp_total
        self.p = R.default_params(width=width, height=height, spp=spp_total, sample_begin=rank * spp_rank, sample_end=(rank + 1) * spp_rank,
                                  pattern=R.PATTERN_REFERENCE, mode=mode, shade_bounces=5, gi_bounces=4)
        self.p_all = R.default_params(width=width, height=height, spp=spp_total, pattern=R.PATTERN_REFERENCE, mode=mode, shade_bounces=5, gi_bounces=4)

    def step(self, scene=None):
        sc = scene or self.sc
        sc.render_device(self.p, 0, clear=True)
        if self.world > 1:
            sc.reduce_resolve(self.comm, self.p, 0, root=0, want=())   # reduce only: no host buffers, nothing to wait for

    def close(self):
        self.sc.close()
        self.hs.close()


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
        # torch.distributed is only plumbing here: barriers, max-over-ranks of the timings and the hand-over of the 128-byte
        # communicator id; the data path's collective is the library's own (rtu_reduce_resolve)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local))
    W, H = args.width, args.height
    spp_total, spp_rank = sample_budget(args, world)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    ctx = R.Context(local, stream.cuda_stream)
    comm = None
    if world > 1:
        idt = torch.zeros(R.COMM_ID_BYTES, dtype=torch.uint8, device="cuda")
        if rank == 0:
            idt.copy_(torch.frombuffer(bytearray(R.comm_unique_id()), dtype=torch.uint8))
        dist.broadcast(idt, 0)
        comm = R.Comm(ctx, bytes(idt.cpu().numpy().tobytes()), rank, world)
    job = Job(R, ctx, comm, rank, world, SCENE, W, H, _mode_id(R), spp_total, spp_rank)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0])

    def sum_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t[0])

    def timed(step, steps, warmup):
        """`steps` steps bracketed by CUDA events on the render stream, L2 flushed in between; ms summed, max over ranks."""
        for _ in range(warmup):
            step()
        barrier()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        barrier()
        for a, b in ev:
            flush.zero_()          # L2 flush, outside the timed bracket
            a.record()
            step()
            b.record()
        barrier()
        return max_over_ranks(sum(a.elapsed_time(b) for a, b in ev))

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()  # started before the warm-up so that it is certainly printing when the timed region begins
    warm = max(3, args.warmup)
    for _ in range(warm):
        job.step()
    barrier()
    if rank == 0:
        sampler.wait_alive()
    st = job.sc.stats()
    rays_rank = st["trace_rays"] + st["shadow_rays"]
    launches_step = st["kernel_launches"] + (2 if world > 1 else 0)   # + k_pack_rgb and NCCL's reduce kernel
    rays_all = sum_over_ranks(float(rays_rank))
    t_begin = sampler.mark()
    ms_total = timed(job.step, args.steps, 0)
    t_end = sampler.mark()
    clocks = sampler.stop(t_begin, t_end) if rank == 0 else None
    value = rays_all * args.steps / (ms_total * 1e-3) * 1e-6

    # ---- the N-GPU image equals the 1-GPU image ("up to FP32 summation order")
    parity = None
    if world > 1:
        job.sc.render_device(job.p, 0, clear=True)
        multi = job.sc.reduce_resolve(comm, job.p, 0, root=0, want=("rgb",))
        if rank == 0:
            single = job.sc.render(job.p_all, want=("rgb",))["rgb"].astype(np.float64)
            m = multi["rgb"].astype(np.float64)
            rel = np.abs(m - single) / np.maximum(np.maximum(np.abs(m), np.abs(single)), 1e-2)
            parity = {"max_rel": float(rel.max()), "tolerance": 1e-4, "ok": bool(rel.max() <= 1e-4),
                      "what": "rtu_reduce_resolve of %d slices vs one rtu_render of all %d samples, linear RGB, |d| / max(|a|,|b|,0.01)" % (world, spp_total)}
        barrier()

    # ---- e2e: pack + H2D + render + (reduce) + resolve + D2H to page-locked host buffers, every step
    brk = {"upload": 0.0, "render_and_readback": 0.0, "destroy": 0.0}
    pinned = {"rgb8": torch.empty((H, W, 3), dtype=torch.uint8, pin_memory=True).numpy(),
              "z8": torch.empty((H, W), dtype=torch.uint8, pin_memory=True).numpy()}

    def e2e_step():
        t_a = time.perf_counter()
        s2 = R.Scene(ctx, job.hs.desc)
        t_b = time.perf_counter()
        if world > 1:
            s2.render_device(job.p, 0, clear=True)
            s2.reduce_resolve(comm, job.p, 0, root=0, want=("rgb8", "z8"), out=pinned)
        else:
            s2.render(job.p, want=("rgb8", "z8"), out=pinned)
        t_c = time.perf_counter()
        nbytes = s2.stats()["scene_device_bytes"]
        s2.close()
        t_d = time.perf_counter()
        brk["upload"] += t_b - t_a
        brk["render_and_readback"] += t_c - t_b
        brk["destroy"] += t_d - t_c
        return nbytes

    e2e_step()
    barrier()
    for k in brk:
        brk[k] = 0.0
    t0 = time.perf_counter()
    h2d = 0
    for _ in range(args.steps):
        h2d = e2e_step()
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    e2e_value = rays_all * args.steps / e2e_s * 1e-6

    # ---- roofline of the dominant kernel: one extra frame with per-launch CUDA events
    pk = R.default_params(width=W, height=H, spp=spp_total, sample_begin=rank * spp_rank, sample_end=(rank + 1) * spp_rank,
                          pattern=R.PATTERN_REFERENCE, mode=_mode_id(R), shade_bounces=5, gi_bounces=4, flags=R.FLAG_TIME_KERNELS)
    job.sc.render_device(pk, 0, clear=True)
    ks = job.sc.stats()
    # ... and a few samples with RTU_FLAG_REFERENCE_WALK: that mode books the reference's own work (every node visit, child-box test
    # and triangle test of Trace / ShadowTrace), i.e. SURVEY 8d's algorithmic bytes per ray of this scene.  The frame kernels
    # find the same answers with less work (pruning, light masks and lists), so their own counters say what they executed, not
    # what the path costs per unit.
    rs = None
    try:
        pr = R.default_params(width=W, height=H, spp=spp_total, sample_begin=0, sample_end=min(4, spp_total), pattern=R.PATTERN_REFERENCE,
                              mode=_mode_id(R), shade_bounces=5, gi_bounces=4, flags=R.FLAG_REFERENCE_WALK)
        job.sc.render_device(pr, 0, clear=True)
        rs = job.sc.stats()
    except R.RtuError:
        rs = None  # (a mesh without a cyBVH: there is no reference walk to count)

    # ---- BASELINE config 4: the GI scene with the same sample budget, same slicing
    gi = None
    if not args.no_gi and not (SCENE == GI_SCENE and MODE == "path"):
        gsteps = max(2, args.steps // 5)
        gjob = Job(R, ctx, comm, rank, world, GI_SCENE, GI_SIZE[0], GI_SIZE[1], R.MODE_PATH, spp_total, spp_rank)
        gjob.step()
        gst = gjob.sc.stats()
        grays = sum_over_ranks(float(gst["trace_rays"] + gst["shadow_rays"]))
        gms = timed(gjob.step, gsteps, 1)
        gi = {"workload": "%s %dx%d HEAD estimator (4-bounce MonteCarlo GI + Whitted), %d spp per frame, spp-sliced x%d" % (GI_SCENE, GI_SIZE[0], GI_SIZE[1], spp_total, world),
              "value": grays * gsteps / (gms * 1e-3) * 1e-6, "unit": "Mrays/s", "ms_per_step": gms / gsteps, "steps": gsteps, "rays_per_step": grays,
              "queue_retries": gst["queue_retries"]}
        if world > 1:
            # the same budget on rank 0 alone, in the same run: the denominator of the scaling efficiency
            t1 = None
            if rank == 0:
                def solo():
                    gjob.sc.render_device(gjob.p_all, 0, clear=True)
                solo()
                torch.cuda.synchronize()
                evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(2)]
                for a, b in evs:
                    flush.zero_()
                    a.record()
                    solo()
                    b.record()
                torch.cuda.synchronize()
                t1 = sum(a.elapsed_time(b) for a, b in evs) / len(evs)
                gi["ms_per_step_1gpu_same_run"] = t1
                gi["speedup_vs_n1"] = t1 / (gms / gsteps)
                gi["efficiency_vs_n1"] = t1 / (gms / gsteps) / world
            barrier()
        gjob.close()

    if rank == 0:
        roof = roofline(ks, clocks, SCENE, MODE, ks["scene_device_bytes"], rs)
        line = {"metric": "Mrays/s", "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": args.steps, "warmup": warm,
                "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
                "dtype": "f32", "data": "synthetic", "config": workload_config(args, world),
                "clocks": clocks,
                "e2e": {"value": e2e_value, "unit": "Mrays/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(W * H * 4),
                        "ms_per_step": e2e_s / args.steps * 1e3,
                        "breakdown_ms_per_step": {k: v / args.steps * 1e3 for k, v in brk.items()},
                        "timer": "host perf_counter around a synchronized region (includes scene packing on the host), max over ranks"},
                "gpu_launches": int(launches_step * args.steps),
                "rays_per_step": rays_all,
                "queue_retries": st["queue_retries"],
                "roofline": roof}
        if parity is not None:
            line["multi_gpu_parity"] = parity
        if gi is not None:
            line["gi"] = gi
        if world == 1 and not args.no_cpu:
            kind, k, res = cpu_sample(W, H, spp_total)
            line["cpu_baseline"] = {"value": res["mrays_per_s"], "unit": "Mrays/s", "cores": res["threads"], "kind": kind,
                                    "sample": "samples [0,%d) of the %d-spp pattern, all %dx%d pixels (%.1f s)" % (k, spp_total, W, H, res["seconds"]),
                                    "build": "-O2 -ffp-contract=off, all host threads, glibc rand() as it is",
                                    "variants": cpu_variants(W, H, spp_total)}
        print(json.dumps(line))
    job.close()
    if comm is not None:
        comm.close()
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
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"],
                    help="strong: --spp-total samples per frame shared by the GPUs; weak: --spp samples per GPU")
    ap.add_argument("--spp-total", type=int, default=1024, help="samples per pixel of the frame (strong scaling); 1024 = the reference's maxSampleSize")
    ap.add_argument("--spp", type=int, default=64, help="samples per pixel per GPU (weak scaling)")
    ap.add_argument("--width", type=int, default=1920)
    ap.add_argument("--height", type=int, default=1080)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-gi", action="store_true", help="skip the GI workload (the \"gi\" key)")
    ap.add_argument("--scene", default=SCENE, help="scene file under scenes/ (default: the Teapot headline scene)")
    ap.add_argument("--mode", default="whitted", choices=["whitted", "path"], help="path = HEAD estimator with Monte-Carlo GI")
    ap.add_argument("--workload", default=None, choices=["teapot", "gi"], help="shorthand: gi = --scene Project11/scene.xml --mode path --width 800 --height 600")
    args = ap.parse_args()
    if args.workload == "gi":
        args.scene, args.mode, args.width, args.height = GI_SCENE, "path", GI_SIZE[0], GI_SIZE[1]
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
