#!/usr/bin/env python
"""bench.py — the headline benchmark of BASELINE.json: Mrays/s and ms/frame for horse_and_mug at 7680x3840
with 16x16 SSAA (config 5; 7.55 G sub-samples, 25.99 G rays per frame) on N B200s.

    python bench.py --gpus 1 --steps 3 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...     # the reference's own CPU renderer on the box's host cores

A step = one frame.  One process per GPU; the frame is cut into bands of pixel rows dealt round-robin to the ranks
(one row per band at 16x16 SSAA, exactly the reference's row deal; scene replicated), every rank renders its bands
with ONE persistent kernel.
  value  device time with the scene resident in HBM and the frame ending up on GPU 0: CUDA events on the launching
         stream, max over ranks; at N > 1 the packed bands are gathered to rank 0 over NCCL and scattered into the
         row-major frame by one small kernel.
  e2e    the same frame through the C-ABI call a user makes, RGB8 landing in page-locked HOST memory inside the
         timed region: rt_render at N = 1; at N > 1 rt_render_part_to_host on every rank — each GPU copies its own
         bands over its own PCIe link into a frame in shared memory, no gather, then one barrier.
At N = 1 the line also carries BASELINE.json's configurations 1-4 (`configs`: kernel ms, end-to-end ms, the
reference's render-only time on the host cores, and the process wall time of `raytracer scene.xml` against the
reference's binary), the scene-creation times (cold and warm) and one dynamic-scene number (build + frame).
"""
import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tools"))

FLOP_PER_RAY = {"horse_and_mug": 614, "dragon_lowres": 712, "bunny": 432, "mirror_spheres": 178, "simple": 78}
CACHE_B_PER_RAY = {"horse_and_mug": 731, "dragon_lowres": 845, "bunny": 520, "mirror_spheres": 185, "simple": 92}
SMALL_CONFIGS = [("1", "simple", "simple.aa1"), ("2", "bunny", "bunny.aa1"), ("3", "horse_and_mug", "horse_and_mug.aa1"),
                 ("4a", "dragon_lowres", "dragon_lowres.aa1"), ("4b", "mirror_spheres", "mirror_spheres.aa1")]


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--scene", default="horse_and_mug")
    ap.add_argument("--width", type=int, default=7680)
    ap.add_argument("--height", type=int, default=3840)
    ap.add_argument("--aa", type=int, default=16)
    ap.add_argument("--builder", default="default", choices=["default", "sah", "lbvh", "ploc", "sah_gpu"])
    ap.add_argument("--cpu-rows", type=int, default=0, help="sub-sample rows in the CPU sample (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip BASELINE.json's configurations 1-4 (N = 1 only)")
    ap.add_argument("--full-frame", action="store_true", help="--impl reference: one pass over EVERY sub-sample row (~4 min on 16 cores)")
    ap.add_argument("--gather", default="nccl", choices=["nccl", "p2p"],
                    help="N>1, device-resident frame: nccl = packed bands + one NCCL gather + scatter kernel; p2p = every rank's kernel "
                         "stores its pixels straight into rank 0's frame over NVLink (CUDA IPC peer mapping), then one barrier")
    return ap.parse_args()


# ----------------------------------------------------------------------------------------------- clocks


class ClockSampler:
    """nvidia-smi sampled every 200 ms during the timed region (B200_PROFILING.md's clocks line)."""
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(prefix="clocks_", suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if not self.proc:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        try:
            for line in open(self.path):
                f = [x.strip() for x in line.split(",")]
                if len(f) < 9:
                    continue
                try:
                    sm.append(float(f[1]))
                    mx.append(float(f[2]))
                except ValueError:
                    continue
                for name, val in zip(names, f[5:9]):
                    if val.lower() == "active":
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            busy = [s for s in sm if s >= 0.5 * max(sm)] or sm
            out.update(sm_mhz=statistics.median(busy), sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))
        return out


# ----------------------------------------------------------------------------------------------- reference arm


class quiet_stdout:
    """The reference prints "Rendering ... with N cores..." from C (raytracer.cpp:368) on every render call: keep this
    process's stdout to the one JSON line by pointing fd 1 at /dev/null for the duration."""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        devnull = os.open(os.devnull, os.O_WRONLY)
        os.dup2(devnull, 1)
        os.close(devnull)

    def __exit__(self, *exc):
        try:
            import ctypes as C
            C.CDLL(None).fflush(None)
        except Exception:
            pass
        os.dup2(self.saved, 1)
        os.close(self.saved)


def workload_name(args):
    return f"{args.scene}.xml {args.width}x{args.height} output, {args.aa}x{args.aa} SSAA ({args.width * args.aa}x{args.height * args.aa} sub-samples)"


def workload_config(args):
    """The `config` object both arms print (identical, so that the driver's same_config check holds)."""
    return {"workload": workload_name(args), "scene": args.scene, "width": args.width, "height": args.height, "aa": args.aa}


def cpu_sample_rows(args, threads, passes=1):
    """Bounded sample of the workload for the CPU: every k-th sub-sample row of the (width*aa) x (height*aa) grid,
    sized for roughly 15 s of CPU work per pass (~2.5 Mrays/s per core, 3.44 rays per sub-sample on horse_and_mug)
    and at most ~2 minutes over all passes of a run."""
    total_rows = args.height * args.aa
    if args.cpu_rows > 0:
        n = min(args.cpu_rows, total_rows)
    else:
        seconds = max(2.0, min(15.0, 120.0 / max(1, passes)))
        target_rays = seconds * 2.5e6 * max(1, threads)
        n = int(target_rays / (3.44 * args.width * args.aa))
        n = max(4, min(n, total_rows))
    stride = max(1, total_rows // n)
    n = min(n, (total_rows + stride - 1) // stride)
    return stride // 2 if stride > 1 else 0, stride, n


def reference_sample(args, steps, warmup, full_frame=False):
    """Times the reference's own generate/rayTrace/toPixel (oracle/_ref/libref.so when it was built from
    /root/reference, else the C port) on the sample; ray count of the sample from the C port (same decisions).
    The scene comes from the reference's own loader (RefScene.to_scene) when the reference is available: no product
    library is loaded on this path."""
    import harness as H
    threads = os.cpu_count() or 8
    if full_frame:
        row0, stride, n_rows = 0, 1, args.height * args.aa
    else:
        row0, stride, n_rows = cpu_sample_rows(args, threads, passes=steps + warmup + 1)
    kind = "reference" if H.ref_available() else "port"
    ref = None
    if kind == "reference":
        ref = H.RefScene(H.golden_scene_path(args.scene))
        sc = ref.to_scene()
    else:
        sc = H.golden_scene(args.scene)
    cam = sc.camera(0, args.width, args.height)
    orc = H.OracleScene(sc)
    t0 = time.perf_counter()
    _, st = orc.render_rows(cam, args.aa, row0, stride, n_rows, threads=threads, keep=False)
    port_secs = time.perf_counter() - t0
    rays = st.total_rays
    times = []
    if kind == "reference":
        for i in range(warmup + steps):
            secs, _ = ref.time_rows(0, args.aa, args.width, args.height, row0, stride, n_rows, threads)
            if i >= warmup:
                times.append(secs)
        ref.close()
    else:
        times.append(port_secs)
        for i in range(max(0, warmup + steps - 1)):
            t0 = time.perf_counter()
            orc.render_rows(cam, args.aa, row0, stride, n_rows, threads=threads, keep=False)
            times.append(time.perf_counter() - t0)
        times = times[-steps:]
    orc.close()
    secs = sum(times) / len(times)
    sample = (f"{n_rows} of {args.height * args.aa} sub-sample rows (every {stride}th from row {row0}) of the "
              f"{args.width * args.aa}x{args.height * args.aa} grid = {n_rows * args.width * args.aa} sub-samples, {rays} rays")
    return {"value": rays / secs / 1e6, "unit": "Mrays/s", "cores": threads, "kind": kind, "sample": sample,
            "seconds_per_sample": secs, "rays": rays, "fraction_of_frame": n_rows / (args.height * args.aa)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    # a step of the reference arm is one pass over the bounded sample; same step / warm-up counts as the B200 arm
    steps, warmup = max(1, args.steps), max(args.warmup, 3)
    if args.full_frame:
        steps, warmup = 1, 0
    base = reference_sample(args, steps, warmup, full_frame=args.full_frame)
    line = {"impl": "reference", "metric": "Mrays/s (primary + shadow + reflection)", "value": base["value"], "unit": "Mrays/s",
            "n_gpus": args.gpus, "steps": steps, "warmup": warmup, "ms_per_step": base["seconds_per_sample"] * 1e3,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
            "data": "reference's shipped scene (tests/golden/scenes, deterministic; the metric is defined on it, not on synthetic data)",
            "config": workload_config(args),
            "step": "one pass over the bounded CPU sample" if not args.full_frame else "one pass over the whole frame",
            "frame_seconds_extrapolated": base["seconds_per_sample"] / base["fraction_of_frame"],
            "cpu_baseline": {k: base[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": base["value"], "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))
    return 0


# ----------------------------------------------------------------------------------------------- B200 arm


def small_configs(H, torch):
    """BASELINE.json configurations 1-4 on this GPU (outside the timed region of the headline number): render-kernel ms
    (CUDA events, best of 20), end-to-end ms of rt_render with the frame landing in page-locked / pageable host memory
    (median of 20), scene creation (warm), byte identity with the golden frame rendered by the unmodified reference,
    and the reference's render-only time on this box's host cores (its own thread fan-out)."""
    import numpy as np
    out = []
    for cfg, scene, key in SMALL_CONFIGS:
        gold, m = H.golden_image(key)
        sc = H.golden_scene(scene)
        cam = sc.camera(m["camera"], m["width"], m["height"])
        t0 = time.perf_counter()
        rt = H.RayTracer(sc)
        create_ms = (time.perf_counter() - t0) * 1e3
        inf = rt.info()
        pinned = torch.empty(gold.size, dtype=torch.uint8, pin_memory=True)
        pageable = np.empty_like(gold)
        for _ in range(3):
            rt.render(cam, m["aa"], out=pinned)
        k_ms, e_pin, e_page = [], [], []
        for _ in range(20):
            t0 = time.perf_counter()
            rt.render(cam, m["aa"], out=pinned)
            e_pin.append((time.perf_counter() - t0) * 1e3)
            k_ms.append(rt.last_stats.ms_render)
            t0 = time.perf_counter()
            rt.render(cam, m["aa"], out=pageable)
            e_page.append((time.perf_counter() - t0) * 1e3)
        st = rt.last_stats
        rec = {"config": cfg, "scene": scene, "width": m["width"], "height": m["height"], "aa": m["aa"], "rays": st.total_rays,
               "kernel_ms": min(k_ms), "e2e_ms": statistics.median(e_pin), "e2e_pageable_ms": statistics.median(e_page),
               "kernel_mrays_s": st.total_rays / min(k_ms) / 1e3, "e2e_mrays_s": st.total_rays / statistics.median(e_pin) / 1e3,
               "scene_create_ms": create_ms, "build_device_ms": inf.ms_build_device,
               "identical_to_reference_frame": bool(np.array_equal(pinned.numpy().reshape(gold.shape), gold) and np.array_equal(pageable, gold))}
        rt.close()
        if H.ref_available():
            ref = H.RefScene(H.golden_scene_path(scene))
            with quiet_stdout():
                secs = min(ref.render(m["camera"], m["aa"])[1] for _ in range(3))
            rec.update(reference_render_s=secs, reference_build_s=ref.L.ref_build_seconds(ref.h), reference_mrays_s=st.total_rays / secs / 1e6,
                       e2e_speedup_vs_reference_render=secs * 1e3 / rec["e2e_ms"])
            ref.close()
        out.append(rec)
    return out


def cli_walls():
    """Process wall time of `raytracer scene.xml --aa 1` against the reference's binary for configurations 1-4
    (tools/cli_wall.py; README.md:1,8 quotes 0.452 s for horse_and_mug)."""
    try:
        import cli_wall
        rows = []
        for _, scene, _k in SMALL_CONFIGS:
            r = cli_wall.measure(scene, 1, 3)
            rows.append({"scene": scene, "aa": 1, "ours_wall_s": r["ours"]["wall_s"], "ours_first_run_wall_s": r["ours"]["first_run_wall_s"],
                         "ours_planted_s": r["ours"]["planted_s"], "ours_rendered_s": r["ours"]["rendered_s"],
                         "reference_wall_s": r["reference"]["wall_s"] if r.get("reference") else None,
                         "reference_total_s": r["reference"]["total_s"] if r.get("reference") else None,
                         "ppm_identical": r.get("identical")})
        return {"cuda_startup_probe_s": cli_wall.probe(3), "rows": rows,
                "note": "ours_wall_s includes the CUDA start-up of a fresh process (cuda_startup_probe_s: `raytracer --probe`)"}
    except Exception as e:  # the CLI binaries are optional for the benchmark
        return {"unavailable": str(e)[:200]}


def run_b200(args):
    import torch
    import harness as H

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the B200 path has no CPU fallback")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    sc = H.golden_scene(args.scene)
    cam = sc.camera(0, args.width, args.height)
    builder = {"default": 0, "lbvh": 1, "sah": 2, "ploc": 3, "sah_gpu": 5}[args.builder]
    torch.zeros(1, device="cuda")  # CUDA context up before the scene build is timed
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    rt = H.RayTracer(sc, builder=builder)  # first creation in this process: includes loading the library's kernels
    build_cold_s = time.perf_counter() - t0
    rt.close()
    warm = []
    for _ in range(3):
        t0 = time.perf_counter()
        rt = H.RayTracer(sc, builder=builder)
        warm.append(time.perf_counter() - t0)
        info = rt.info()
        if len(warm) < 3:
            rt.close()
    build_warm_s = statistics.median(warm)

    aa = args.aa
    frame_bytes = cam.image_width * cam.image_height * 3
    stride = rt.part_bytes(cam, aa, 0, world)
    stream = torch.cuda.current_stream()
    my_rows = torch.zeros(max(stride, 1), dtype=torch.uint8, device="cuda")
    frame = torch.zeros(frame_bytes, dtype=torch.uint8, device="cuda") if rank == 0 else None
    all_parts = torch.zeros((world, stride), dtype=torch.uint8, device="cuda") if (rank == 0 and world > 1) else None
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2
    if world == 1:
        host_frame = torch.empty(frame_bytes, dtype=torch.uint8, pin_memory=True)
        shared = None
    else:
        shared = H.rt_b200.SharedHostFrame(dist, rank, frame_bytes)  # page-locked in every rank
        host_frame = None

    peer = None
    if world > 1 and args.gather == "p2p":
        peer = H.rt_b200.PeerFrame(dist, rank, frame_bytes)
        if rank == 0:
            frame = peer.as_tensor()

    def device_step(e1=None):
        """render my bands -> (gather -> assemble on rank 0); everything on torch's current stream"""
        if world == 1:
            rt.render_part_into_frame(cam, aa, 0, 1, frame.data_ptr(), stream.cuda_stream, want_stats=False)
            if e1:
                e1.record(stream)
        elif peer is not None:
            rt.render_part_into_frame(cam, aa, rank, world, peer.ptr.value, stream.cuda_stream, want_stats=False)
            if e1:
                e1.record(stream)
            dist.barrier()  # stream-ordered: rank 0 may read the frame once every rank's kernel has retired
        else:
            rt.render_part(cam, aa, rank, world, my_rows.data_ptr(), stream.cuda_stream, want_stats=False)
            if e1:
                e1.record(stream)
            H.rt_b200.gather_parts(dist, my_rows, all_parts, rank)
            if rank == 0:
                rt.assemble(cam, aa, world, all_parts.data_ptr(), stride, frame.data_ptr(), stream.cuda_stream)

    def barrier():
        torch.cuda.synchronize()
        if dist:
            dist.barrier()
            torch.cuda.synchronize()

    # exact ray counts of the frame (deterministic): one untimed pass with the device counters read back
    st = rt.render_part(cam, aa, rank, world, my_rows.data_ptr(), stream.cuda_stream, want_stats=True)
    counts = torch.tensor([st.primary_rays, st.reflection_rays, st.shadow_rays, st.shadow_occluded], dtype=torch.int64, device="cuda")
    if dist:
        dist.all_reduce(counts)
    primary, reflection, shadow, occluded = [int(x) for x in counts.tolist()]
    rays = primary + reflection + shadow

    # timing rule: at least 3 untimed warm-up frames (the first seconds under load run slower on these boards)
    warmup = max(args.warmup, 3)
    for _ in range(warmup):
        flush.zero_()
        device_step()
    barrier()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()

    # ---- value: device time, scene resident, CUDA events on the launching stream, L2 flushed between steps
    barrier()
    wall0 = time.perf_counter()
    step_ms, render_ms = [], []
    for _ in range(args.steps):
        flush.zero_()
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        e0.record(stream)
        device_step(e1)
        e2.record(stream)
        e2.synchronize()
        step_ms.append(e0.elapsed_time(e2))
        render_ms.append(e0.elapsed_time(e1))
    barrier()
    wall_s = time.perf_counter() - wall0
    t = torch.tensor([sum(step_ms), sum(render_ms)], dtype=torch.float64, device="cuda")
    if dist:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, total_render_ms = [float(x) for x in t.tolist()]
    ms_per_step = total_ms / args.steps
    kernel_ms = total_render_ms / args.steps  # the render kernel alone (max over ranks)
    per_rank_kernel_ms = [sum(render_ms) / args.steps]
    if dist:
        g = [torch.zeros(1, dtype=torch.float64, device="cuda") for _ in range(world)]
        dist.all_gather(g, torch.tensor([per_rank_kernel_ms[0]], dtype=torch.float64, device="cuda"))
        per_rank_kernel_ms = [float(x.item()) for x in g]

    # ---- e2e: the user-facing C-ABI call, frame into page-locked HOST memory inside the timed region
    e2e_ms = []
    for i in range(1 + args.steps):
        flush.zero_()
        barrier()
        t0 = time.perf_counter()
        if world == 1:
            rt.render(cam, aa, out=host_frame)  # rt_render: kernel + one D2H, synchronous
        else:
            rt.render_part_to_host(cam, aa, rank, world, shared.ptr.value)  # kernel + this GPU's own strided D2H, synchronous
            dist.barrier()  # the frame is complete once every rank has returned
        if i > 0:
            e2e_ms.append((time.perf_counter() - t0) * 1e3)
    t = torch.tensor([sum(e2e_ms)], dtype=torch.float64, device="cuda")
    if dist:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_ms_per_step = float(t.item()) / args.steps

    clocks = sampler.stop() if rank == 0 else None

    if rank == 0:
        import hashlib
        final = host_frame.numpy() if world == 1 else shared.as_numpy()
        frame_sha = hashlib.sha256(final.tobytes()).hexdigest()
        device_sha = hashlib.sha256(frame.cpu().numpy().tobytes()).hexdigest()
        value = rays / (ms_per_step * 1e3)  # Mrays/s, whole job
        f_clk = (clocks.get("sm_mhz") or 1965.0) * 1e6
        n_sm = torch.cuda.get_device_properties(local).multi_processor_count
        flop_ray = FLOP_PER_RAY.get(args.scene)
        kernel_rays_per_s = rays / world / (kernel_ms * 1e-3)  # per GPU
        peak_tflops = n_sm * 128 * f_clk / 1e12  # FP32 non-FMA issue peak at the clock observed under load
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        wl_key = f"{args.scene}:{args.width}x{args.height}:{args.aa}:{world}"
        traffic, counters = None, {}
        try:
            tr = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
            traffic = tr.get(wl_key, {}).get("bytes")
            counters = tr.get(wl_key, {}).get("ncu", {})
        except Exception:
            pass
        roofline = None
        if flop_ray:
            achieved = kernel_rays_per_s * flop_ray / 1e12
            roofline = {"bound": "fp32_issue", "bound_note": "neither hbm nor tensor: the scene is L2-resident and the path is not a contraction (SURVEY.md 8d)",
                        "kernel": "rtb::render_kernel_v2", "achieved": achieved, "peak": peak_tflops, "unit": "TFLOP/s",
                        "frac": achieved / peak_tflops,
                        "frac_note": "SURVEY.md 8d's definition: the REFERENCE tree's work per ray; the utilisation of what this kernel "
                                     "actually issues is issue_slot_util x lanes_active / 32 below (ncu, same build, profiles/)",
                        "issue_slot_util": counters.get("issue_slot_util"), "lanes_active": counters.get("lanes_active"),
                        "warp_inst_per_ray": counters.get("warp_inst_per_ray"), "l1_wavefront_util": counters.get("l1_wavefront_util"),
                        "ncu_source": counters.get("source"),
                        "peak_source": f"{n_sm} SMs x 128 FP32 lanes x {f_clk / 1e6:.0f} MHz observed under load (non-FMA; no measured FP32 peak in MEASURED_PEAKS.json)",
                        "algorithmic_flop_per_ray": flop_ray, "rays_per_launch": rays // world, "kernel_ms": kernel_ms,
                        "traffic": traffic, "traffic_unit": "bytes per launch (ncu dram read+write)",
                        "hbm": {"algorithmic_bytes_per_launch": frame_bytes // world, "achieved_gbs": frame_bytes / world / (kernel_ms * 1e-3) / 1e9,
                                "peak_gbs": hbm_peak, "peak_source": "measured" if peaks else "fallback",
                                "frac": frame_bytes / world / (kernel_ms * 1e-3) / 1e9 / hbm_peak},
                        "l1l2_cache": {"algorithmic_bytes_per_ray": CACHE_B_PER_RAY.get(args.scene),
                                       "achieved_gbs": kernel_rays_per_s * CACHE_B_PER_RAY.get(args.scene, 0) / 1e9}}
        details = {"rays_per_frame": rays, "primary": primary, "reflection": reflection,
                   "shadow": shadow, "shadow_occluded": occluded, "parallelism": f"row_bands_h{H.rt_b200.band_height(cam, aa, world)}_interleaved_x{world}",
                   "gather": "none" if world == 1 else ("fused: kernels store into rank 0's frame over NVLink P2P (CUDA IPC), one barrier" if peer is not None
                                                       else "nccl gather of packed bands + scatter kernel"),
                   "l2": "flushed between timed iterations (256 MiB write)", "bvh": {0: "default", 1: "lbvh_gpu", 2: "sah_host", 3: "ploc_gpu", 4: "auto", 5: "sah_gpu"}[info.builder],
                   "bvh_nodes": info.bvh_nodes, "frame_sha256": frame_sha, "device_frame_sha256": device_sha}
        line = {"metric": "Mrays/s (primary + shadow + reflection)", "value": value, "unit": "Mrays/s", "n_gpus": world,
                "steps": args.steps, "warmup": warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "f32",
                "data": "reference's shipped scene (tests/golden/scenes, deterministic; the metric is defined on it, not on synthetic data)",
                "config": workload_config(args),  # the same object the reference arm prints
                "frame": details,
                "ms_per_frame": ms_per_step, "render_kernel_ms": kernel_ms, "render_kernel_ms_per_rank": [round(x, 3) for x in per_rank_kernel_ms],
                "wall_s_timed_region": wall_s,
                "step_ms_rank0": [round(x, 3) for x in step_ms],
                "clocks": {k: clocks[k] for k in ("sm_mhz", "sm_max_mhz", "reasons")},
                "e2e": {"value": rays / (e2e_ms_per_step * 1e3), "unit": "Mrays/s", "ms_per_frame": e2e_ms_per_step,
                        "h2d_bytes_per_step": ctypes.sizeof(H.RtCamera) * world, "d2h_bytes_per_step": frame_bytes,
                        "api": "rt_render (C-ABI) via rt_b200.RayTracer.render into pinned host memory; for frames >= 32 MB the kernel stores finished "
                               "pixels straight into that memory over PCIe (no separate D2H copy), timed with the host clock around the call" if world == 1
                        else "rt_render_part_to_host (C-ABI) on every rank: own bands over own PCIe link into a shared page-locked frame, one barrier"},
                "gpu_launches": args.steps * (world + (1 if (world > 1 and peer is None) else 0)),
                "scene_build": {"cold_s": build_cold_s, "warm_s": build_warm_s, "device_ms": info.ms_build_device, "host_enqueue_ms": info.ms_build_host,
                                "sah_cost_ploc": info.sah_cost_ploc, "sah_cost_sah": info.sah_cost_sah, "sah_cost_kept": info.bvh_sah_cost,
                                "reinsertion": {"moves": info.reinsert_moves, "rounds": info.reinsert_rounds, "sah_cost_before": info.reinsert_cost_before,
                                                "sah_cost_after": info.reinsert_cost_after, "kept": bool(info.reinsert_accepted)},
                                "note": "cold = first rt_scene_create of the process (loads the library's kernels); warm = median of 3 more",
                                "dynamic_scene_ms_per_frame": build_warm_s * 1e3 + e2e_ms_per_step},
                "roofline": roofline}
        if world == 1 and not args.no_cpu_baseline:
            base = reference_sample(args, 1, 0)
            line["cpu_baseline"] = {k: base[k] for k in ("value", "unit", "cores", "kind", "sample")}
        if world == 1 and not args.no_configs:
            rt.close()
            line["configs"] = small_configs(H, torch)
            line["cli_wall"] = cli_walls()
        print(json.dumps(line))
    rt.close()
    if shared is not None:
        dist.barrier()
        shared.close()
    if peer is not None:
        dist.barrier()
        peer.close()
    if dist:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    a = parse_args()
    sys.exit(run_reference(a) if a.impl == "reference" else run_b200(a))
