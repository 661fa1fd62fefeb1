#!/usr/bin/env python
"""bench.py — the headline benchmark of BASELINE.json: Mrays/s and ms/frame for horse_and_mug at 7680x3840
with 16x16 SSAA (config 5; 7.55 G sub-samples, 25.99 G rays per frame) on N B200s.

    python bench.py --gpus 1 --steps 3 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...     # the reference's own CPU renderer on the box's host cores

A step = one frame.  One process per GPU; the frame is cut into 32x32-pixel tiles dealt round-robin to the ranks
(scene replicated), every rank renders its tiles with ONE persistent kernel, the packed tiles are gathered to rank 0
over NCCL and scattered into the row-major frame by one small kernel.  `value` is device time with the scene
resident in HBM (CUDA events on the launching stream, max over ranks); `e2e` is the same frame through the public
API with the RGB8 frame landing in pinned host memory (rt_render at N=1).
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

FLOP_PER_RAY = {"horse_and_mug": 614, "dragon_lowres": 712, "bunny": 432, "mirror_spheres": 178, "simple": 78}
CACHE_B_PER_RAY = {"horse_and_mug": 731, "dragon_lowres": 845, "bunny": 520, "mirror_spheres": 185, "simple": 92}


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
    ap.add_argument("--gather", default="nccl", choices=["nccl", "p2p"],
                    help="N>1: nccl = packed tiles + one NCCL gather + scatter kernel; p2p = every rank's kernel stores its "
                         "pixels straight into rank 0's frame over NVLink (CUDA IPC peer mapping), then one barrier")
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


def reference_sample(args, steps, warmup):
    """Times the reference's own generate/rayTrace/toPixel (oracle/_ref/libref.so when it was built from
    /root/reference, else the C port) on the sample; ray count of the sample from the C port (same decisions)."""
    import harness as H
    threads = os.cpu_count() or 8
    row0, stride, n_rows = cpu_sample_rows(args, threads, passes=steps + warmup + 1)
    sc = H.golden_scene(args.scene)
    cam = sc.camera(0, args.width, args.height)
    orc = H.OracleScene(sc)
    _, st = orc.render_rows(cam, args.aa, row0, stride, n_rows, threads=threads, keep=False)
    rays = st.total_rays
    kind = "reference" if H.ref_available() else "port"
    times = []
    if kind == "reference":
        ref = H.RefScene(H.golden_scene_path(args.scene))
        for i in range(warmup + steps):
            secs, _ = ref.time_rows(0, args.aa, args.width, args.height, row0, stride, n_rows, threads)
            if i >= warmup:
                times.append(secs)
        ref.close()
    else:
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            orc.render_rows(cam, args.aa, row0, stride, n_rows, threads=threads, keep=False)
            if i >= warmup:
                times.append(time.perf_counter() - t0)
    orc.close()
    secs = sum(times) / len(times)
    sample = (f"{n_rows} of {args.height * args.aa} sub-sample rows (every {stride}th from row {row0}) of the "
              f"{args.width * args.aa}x{args.height * args.aa} grid = {n_rows * args.width * args.aa} sub-samples, {rays} rays")
    return {"value": rays / secs / 1e6, "unit": "Mrays/s", "cores": threads, "kind": kind, "sample": sample,
            "seconds_per_sample": secs, "rays": rays}


def workload_name(args):
    return f"{args.scene}.xml {args.width}x{args.height} output, {args.aa}x{args.aa} SSAA ({args.width * args.aa}x{args.height * args.aa} sub-samples)"


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    # a step of the reference arm is one pass over the bounded sample
    steps, warmup = max(1, args.steps), max(0, min(args.warmup, 1))
    base = reference_sample(args, steps, warmup)
    line = {"impl": "reference", "metric": "Mrays/s (primary + shadow + reflection)", "value": base["value"], "unit": "Mrays/s",
            "n_gpus": args.gpus, "steps": steps, "warmup": warmup, "ms_per_step": base["seconds_per_sample"] * 1e3,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
            "data": "reference's shipped scene (tests/golden/scenes, deterministic)",
            "config": {"workload": workload_name(args), "step": "one pass over the bounded CPU sample"},
            "cpu_baseline": {k: base[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": base["value"], "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))
    return 0


# ----------------------------------------------------------------------------------------------- B200 arm


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
    rt = H.RayTracer(sc, builder=builder)
    build_s = time.perf_counter() - t0
    info = rt.info()

    aa = args.aa
    frame_bytes = cam.image_width * cam.image_height * 3
    stride = rt.part_bytes(cam, 0, world)
    stream = torch.cuda.current_stream()
    my_tiles = torch.zeros(max(stride, 1), dtype=torch.uint8, device="cuda")
    frame = torch.zeros(frame_bytes, dtype=torch.uint8, device="cuda") if rank == 0 else None
    all_parts = torch.zeros((world, stride), dtype=torch.uint8, device="cuda") if (rank == 0 and world > 1) else None
    host_frame = torch.empty(frame_bytes, dtype=torch.uint8, pin_memory=True) if rank == 0 else None
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2

    peer = None
    if world > 1 and args.gather == "p2p":
        peer = H.rt_b200.PeerFrame(dist, rank, frame_bytes)
        if rank == 0:
            frame = peer.as_tensor()

    def device_step():
        """render my tiles -> (gather -> assemble on rank 0); everything on torch's current stream"""
        if world == 1:
            rt.render_part_into_frame(cam, aa, 0, 1, frame.data_ptr(), stream.cuda_stream, want_stats=False)
        elif peer is not None:
            rt.render_part_into_frame(cam, aa, rank, world, peer.ptr.value, stream.cuda_stream, want_stats=False)
            dist.barrier()  # stream-ordered: rank 0 may read the frame once every rank's kernel has retired
        else:
            rt.render_part(cam, aa, rank, world, my_tiles.data_ptr(), stream.cuda_stream, want_stats=False)
            H.rt_b200.gather_parts(dist, my_tiles, all_parts, rank)
            if rank == 0:
                rt.assemble(cam, world, all_parts.data_ptr(), stride, frame.data_ptr(), stream.cuda_stream)

    def barrier():
        torch.cuda.synchronize()
        if dist:
            dist.barrier()
            torch.cuda.synchronize()

    # exact ray counts of the frame (deterministic): one untimed pass with the device counters read back
    st = rt.render_part(cam, aa, rank, world, my_tiles.data_ptr(), stream.cuda_stream, want_stats=True)
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
        if world == 1:
            rt.render_part_into_frame(cam, aa, 0, 1, frame.data_ptr(), stream.cuda_stream, want_stats=False)
            e1.record(stream)
        elif peer is not None:
            rt.render_part_into_frame(cam, aa, rank, world, peer.ptr.value, stream.cuda_stream, want_stats=False)
            e1.record(stream)
            dist.barrier()
        else:
            rt.render_part(cam, aa, rank, world, my_tiles.data_ptr(), stream.cuda_stream, want_stats=False)
            e1.record(stream)
            H.rt_b200.gather_parts(dist, my_tiles, all_parts, rank)
            if rank == 0:
                rt.assemble(cam, world, all_parts.data_ptr(), stride, frame.data_ptr(), stream.cuda_stream)
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

    # ---- e2e: the user-facing call, frame into pinned HOST memory inside the timed region
    e2e_ms = []
    for i in range(1 + args.steps):
        flush.zero_()
        barrier()
        t0 = time.perf_counter()
        if world == 1:
            rt.render(cam, aa, out=host_frame)  # rt_render: kernel + one D2H, synchronous
        else:
            device_step()
            if rank == 0:
                host_frame.copy_(frame, non_blocking=True)
            torch.cuda.synchronize()
        barrier()
        if i > 0:
            e2e_ms.append((time.perf_counter() - t0) * 1e3)
    t = torch.tensor([sum(e2e_ms)], dtype=torch.float64, device="cuda")
    if dist:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_ms_per_step = float(t.item()) / args.steps

    clocks = sampler.stop() if rank == 0 else None

    if rank == 0:
        import hashlib
        frame_sha = hashlib.sha256(host_frame.numpy().tobytes()).hexdigest()
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
        traffic = None
        try:
            tr = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
            traffic = tr.get(f"{args.scene}:{args.width}x{args.height}:{args.aa}:{world}", {}).get("bytes")
        except Exception:
            pass
        roofline = None
        if flop_ray:
            achieved = kernel_rays_per_s * flop_ray / 1e12
            roofline = {"bound": "fp32_issue", "bound_note": "neither hbm nor tensor: the scene is L2-resident and the path is not a contraction (SURVEY.md 8d)",
                        "kernel": "rtb::render_kernel_v2", "achieved": achieved, "peak": peak_tflops, "unit": "TFLOP/s",
                        "frac": achieved / peak_tflops,
                        "peak_source": f"{n_sm} SMs x 128 FP32 lanes x {f_clk / 1e6:.0f} MHz observed under load (non-FMA; no measured FP32 peak in MEASURED_PEAKS.json)",
                        "algorithmic_flop_per_ray": flop_ray, "rays_per_launch": rays // world, "kernel_ms": kernel_ms,
                        "traffic": traffic, "traffic_unit": "bytes per launch (ncu dram read+write)",
                        "hbm": {"algorithmic_bytes_per_launch": frame_bytes // world, "achieved_gbs": frame_bytes / world / (kernel_ms * 1e-3) / 1e9,
                                "peak_gbs": hbm_peak, "peak_source": "measured" if peaks else "fallback",
                                "frac": frame_bytes / world / (kernel_ms * 1e-3) / 1e9 / hbm_peak},
                        "l1l2_cache": {"algorithmic_bytes_per_ray": CACHE_B_PER_RAY.get(args.scene),
                                       "achieved_gbs": kernel_rays_per_s * CACHE_B_PER_RAY.get(args.scene, 0) / 1e9}}
        line = {"metric": "Mrays/s (primary + shadow + reflection)", "value": value, "unit": "Mrays/s", "n_gpus": world,
                "steps": args.steps, "warmup": warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "f32",
                "data": "reference's shipped scene (tests/golden/scenes, deterministic; the metric is defined on it, not on synthetic data)",
                "config": {"workload": workload_name(args), "rays_per_frame": rays, "primary": primary, "reflection": reflection,
                           "shadow": shadow, "shadow_occluded": occluded, "parallelism": f"tiles32x32_interleaved_x{world}",
                           "gather": "none" if world == 1 else ("fused: kernels store into rank 0's frame over NVLink P2P (CUDA IPC), one barrier" if peer is not None
                                                               else "nccl gather of packed tiles + scatter kernel"),
                           "l2": "flushed between timed iterations (256 MiB write)", "bvh": {0: "default", 1: "lbvh_gpu", 2: "sah_host", 3: "ploc_gpu", 4: "auto", 5: "sah_gpu"}[info.builder],
                           "bvh_nodes": info.bvh_nodes, "scene_build_s": build_s, "frame_sha256": frame_sha},
                "ms_per_frame": ms_per_step, "render_kernel_ms": kernel_ms, "wall_s_timed_region": wall_s,
                "step_ms_rank0": [round(x, 3) for x in step_ms],
                "clocks": {k: clocks[k] for k in ("sm_mhz", "sm_max_mhz", "reasons")},
                "e2e": {"value": rays / (e2e_ms_per_step * 1e3), "unit": "Mrays/s", "ms_per_frame": e2e_ms_per_step,
                        "h2d_bytes_per_step": ctypes.sizeof(H.RtCamera) * world, "d2h_bytes_per_step": frame_bytes,
                        "api": "rt_render (C-ABI) via rt_b200.RayTracer.render into pinned host memory" if world == 1
                        else "rt_render_part + NCCL gather + rt_assemble_tiles + D2H into pinned host memory"},
                "gpu_launches": args.steps * (world + (1 if (world > 1 and peer is None) else 0)),
                "roofline": roofline}
        if world == 1 and not args.no_cpu_baseline:
            base = reference_sample(args, 1, 0)
            line["cpu_baseline"] = {k: base[k] for k in ("value", "unit", "cores", "kind", "sample")}
        print(json.dumps(line))
    rt.close()
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
