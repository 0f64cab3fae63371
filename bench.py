#!/usr/bin/env python
"""Headline benchmark: BASELINE.json's metric -- Mrays/s (closest-hit primary + any-hit AO rays) and BVH build Mtris/s --
on config C3 as BASELINE.json words it: the 10M-triangle displaced terrain, PreferBalancedBuild, 1920x1080 primary rays +
16-spp ambient-occlusion rays, one process per GPU.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N ...             # the reference's CPU arm (Embree 4.3 bundled with the
                                                              # reference, all host threads, same frame)

A "step" is ONE frame of ONE job: every primary ray traced closest-hit, `spp` AO rays per hit pixel traced any-hit, the
occlusion resolved to one float per pixel in a framebuffer on GPU 0. With N GPUs the frame's pixels are sharded in 32x32
tiles round-robin over the ranks (BVH replicated), every rank's resolve kernel stores its pixels straight into rank 0's
framebuffer over NVLink and rank 0 waits for the completion flags INSIDE the timed region: "scaling": "strong".
  value  rays of the frame / frame time, rays resident in HBM, device-timed (CUDA events), max over ranks
  e2e    the same rays pushed from / pulled back to pinned HOST buffers through the C-ABI's host entry point
  build  hiprtBuildGeometry of the mesh, timed separately over the same number of steps (one GPU's rate: a single large
         build does not shard -- "replicas only")
  hiprt_reference_gpu   the unmodified HIPRT v3.1.0 kernels on the same GPU and the same rays (oracle/_ref, N = 1 only),
         next to this repo's kernels under that harness's protocol (warm L2, best of 5, 48-byte hits)

Other workloads (--workload): c2_sphere_1M / c3_terrain_10M (primary + one shadow ray per hit pixel, PreferFastBuild).
Scene (C4), batch (C5) and path-tracing (C5) numbers come from tools/bench_scene.py, bench_batch.py, bench_pathtrace.py.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WIDTH, HEIGHT = 1920, 1080
WORKLOADS = {
    # C3 as BASELINE.json words it: PreferBalancedBuild (H-PLOC) + 16-spp ambient-occlusion rays (incoherent any-hit rays, maxT = 0.05)
    "c3_terrain_10M_ao": dict(mesh=("terrain", 2236), eye=(0.5, 0.6, -0.6), look=(0.5, 0.0, 0.5), light=(2.0, 4.0, 3.0), flags="balanced", secondary="ao"),
    "c2_sphere_1M": dict(mesh=("sphere", 1000, 500), eye=(0.0, 0.0, 3.0), look=(0.0, 0.0, 0.0), light=(2.0, 4.0, 3.0), flags="fast"),
    "c3_terrain_10M": dict(mesh=("terrain", 2236), eye=(0.5, 0.6, -0.6), look=(0.5, 0.0, 0.5), light=(2.0, 4.0, 3.0), flags="fast"),
    "small_sphere_40k": dict(mesh=("sphere", 200, 100), eye=(0.0, 0.0, 3.0), look=(0.0, 0.0, 0.0), light=(2.0, 4.0, 3.0), flags="fast"),
    "small_terrain_ao": dict(mesh=("terrain", 200), eye=(0.5, 0.6, -0.6), look=(0.5, 0.0, 0.5), light=(2.0, 4.0, 3.0), flags="balanced", secondary="ao"),
}
DEFAULT_WORKLOAD = "c3_terrain_10M_ao"
BYTES_PER_RAY = 80  # algorithmic HBM traffic of a traced ray: 32 B hiprtRay in + 48 B hiprtHit out
# compulsory HBM bytes per primitive (pair) of each build phase. `boxes`: 24 B of indices in + the 32-byte
# {box, indices} record out + every VERTEX once (12 B x vertices / pairs, filled in from the mesh: shared vertices are served by
# L2 after their first touch -- round 1 counted 4 vertices per pair and read 1.02 of the HBM peak); H-PLOC replaces emit: 200 B model;
# qnodes: 128 B box node in + 64 B record out per BOX NODE (~ 1/3 per primitive pair)
BUILD_BYTES_PER_PRIM = {"pair": 32, "boxes": 56, "morton+histogram": 32, "sort": 68, "emit+refit": 128, "hploc": 200, "collapse+pack": 123 + 152, "qnodes": 64}


def metric_name(wl, spp):
    sec = f"{spp}-spp AO any-hit" if wl.get("secondary") == "ao" else "shadow any-hit"
    return f"Mrays/s (primary closest-hit + {sec}, {WIDTH}x{HEIGHT}); build Mtris/s under 'build'"


def workload_config(name, wl, n_tris, spp):
    """The keys that define the workload -- printed identically by both arms."""
    return {"workload": name, "triangles": int(n_tris), "resolution": [WIDTH, HEIGHT], "primary_rays": WIDTH * HEIGHT,
            "secondary": f"{spp}-spp ambient occlusion (any-hit, maxT 0.05), rays only for hit pixels" if wl.get("secondary") == "ao"
            else "one shadow ray per hit pixel (any-hit)",
            "build": "PreferBalancedBuild" if wl.get("flags") == "balanced" else "PreferFastBuild",
            "ray_order": "32x32 pixel tiles, 8x4 blocks inside a tile (the reference's tests launch 8x8 / 2-D blocks); same order for every arm",
            "frame": "one frame of one job per step; N GPUs share the frame (tile i -> GPU i mod N)"}


def make_mesh(spec):
    from hiprt_b200 import scenes

    if spec[0] == "sphere":
        return scenes.sphere_mesh(spec[1], spec[2])
    if spec[0] == "terrain":
        return scenes.terrain_mesh(spec[1])
    raise ValueError(spec)


def load_json(name):
    try:
        return json.load(open(os.path.join(ROOT, "profiles", name)))
    except Exception:
        return None


def read_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock and throttle reasons sampled through NVML (in-process, every ~1 ms) while the timed
    regions run; falls back to one nvidia-smi query when NVML is not importable."""
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, device_index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self.stop_flag = threading.Event()
        self.thread = None
        try:
            import pynvml

            pynvml.nvmlInit()
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(visible.split(",")[device_index]) if visible and visible.split(",")[device_index].isdigit() else device_index
            self.nv, self.h = pynvml, pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.thread = threading.Thread(target=self._run, daemon=True)
            self.thread.start()
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        while not self.stop_flag.is_set():
            try:
                mhz = float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                bits = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") \
                    else int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                self.samples.append(mhz)
                for bit, name in self.REASONS.items():
                    if bits & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.001)

    def stop(self):
        if self.thread is None:
            try:
                out = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,clocks.max.sm", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=10).stdout.strip().splitlines()[0].split(",")
                return {"sm_mhz": float(out[0]), "sm_max_mhz": float(out[1]), "reasons": [], "samples": 1, "source": "nvidia-smi (single query)"}
            except Exception:
                return {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        self.stop_flag.set()
        self.thread.join(timeout=2)
        sm = sorted(self.samples)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(sm), "source": "NVML, ~1 ms period over the timed regions"}


def frame_rays_host(wl, spp, hits_of):
    """The frame's rays in the single-GPU trace order: (pixels, primary rays, primary hits, secondary rays).
    hits_of(primary rays) -> hits, supplied by the arm (its own engine)."""
    from hiprt_b200 import frame, scenes

    pixels = frame.shard_pixels(WIDTH, HEIGHT, 0, 1)
    primary = frame.frame_rays(WIDTH, HEIGHT, wl["eye"], wl["look"], pixels)
    hits = hits_of(primary)
    if wl.get("secondary") == "ao":
        secondary = scenes.ao_rays(primary, hits, spp=spp, pixel_ids=pixels)
    else:
        secondary = scenes.shadow_rays(primary, hits, wl["light"])[0]
    return pixels, primary, hits, secondary


def embree_arm(verts, idx, wl, steps, warmup, threads, spp):
    """Reference CPU arm: Embree 4.3 build (LOW = Morton, the counterpart of PreferFastBuild; MEDIUM = SAH for the balanced
    workloads) and rtcIntersect1 / rtcOccluded1 over all host threads, on the same frame."""
    from oracle import embree_binding as emb

    if not emb.available():
        return None
    quality = 1 if wl.get("flags") == "balanced" else 0
    build_ms = []
    for _ in range(max(2, min(steps, 3))):
        s = emb.EmbreeScene(verts, idx, quality=quality, threads=threads)
        build_ms.append(s.build_ms)
        s.release()
    scene = emb.EmbreeScene(verts, idx, quality=quality, threads=threads)
    _, primary, hits, secondary = frame_rays_host(wl, spp, lambda rays: scene.trace(rays, kind=2)[0])
    times = []
    for it in range(warmup + steps):
        _, ms_p = scene.trace(primary, kind=2, want_hits=False)
        _, ms_s = scene.trace(secondary, kind=1, want_hits=False)
        if it >= warmup:
            times.append(ms_p + ms_s)
    n_rays = len(primary) + len(secondary)
    ms = float(np.mean(times))
    return {"mrays": n_rays / ms / 1e3, "ms": ms, "rays": n_rays, "build_ms": float(min(build_ms[1:])), "quality": ["LOW", "MEDIUM"][quality],
            "build_mtris": len(idx) / float(min(build_ms[1:])) / 1e3, "hits": hits}


def oracle_port_arm(verts, idx, wl, sample=65536):
    from hiprt_b200 import frame
    from oracle import binding as orc

    t0 = time.perf_counter()
    g = orc.Geometry(verts, idx)
    build_s = time.perf_counter() - t0
    pixels = frame.shard_pixels(WIDTH, HEIGHT, 0, 1)
    rays = frame.frame_rays(WIDTH, HEIGHT, wl["eye"], wl["look"], pixels[:: max(1, len(pixels) // sample)])
    t0 = time.perf_counter()
    g.trace(rays, tie_mode=1)
    trace_s = time.perf_counter() - t0
    return {"mrays": len(rays) / trace_s / 1e6, "rays": len(rays), "build_mtris": len(idx) / build_s / 1e6}


def reference_arm(args, wl, verts, idx, host_threads):
    r = embree_arm(verts, idx, wl, args.steps, args.warmup, host_threads, args.ao_spp)
    kind = "reference"
    if r is None:
        r = oracle_port_arm(verts, idx, wl)
        r["ms"] = r["rays"] / r["mrays"] / 1e3
        kind, cores, sample, engine = "port", 1, f"{r['rays']} primary rays (every k-th pixel), single thread", "oracle port (oracle/hiprt_oracle.cpp)"
    else:
        cores = host_threads
        sample = f"full frame: {r['rays']} rays/step (primary + secondary), {args.steps} steps; build = best of 2 x rtcCommitScene({r['quality']})"
        engine = f"Embree 4.3.0 (contrib/embree of the reference), build quality {r['quality']}, rtcIntersect1 / rtcOccluded1"
    return {
        "impl": "reference", "metric": metric_name(wl, args.ao_spp), "value": round(r["mrays"], 3), "unit": "Mrays/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(r["ms"], 4), "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(args.workload, wl, len(idx), args.ao_spp),
        "rays_per_step": r["rays"], "engine": engine, "build": {"mtris_per_s": round(r["build_mtris"], 3)},
        "cpu_baseline": {"value": round(r["mrays"], 3), "unit": "Mrays/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": round(r["mrays"], 3), "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }


def hiprt_reference_gpu(ctx, api, geom, geom_h, inp, temp, flags, verts, idx, primary, secondary, d_primary, d_secondary, n_tris, ev):
    """The unmodified HIPRT v3.1.0 (oracle/_ref) on this GPU and these rays, and this repo's kernels under the same protocol."""
    import torch

    from oracle import ref_binding as ref

    n_rays = len(primary) + len(secondary)
    rg = ref.RefGeometry(verts, idx, build_flags=0, reps=6)
    _, ms_p = rg.trace(primary, kind=2, reps=6)
    _, ms_s = rg.trace(secondary, kind=1, reps=6)
    d_hits_p = torch.empty((len(primary), 12), dtype=torch.int32, device=ctx.device)
    d_hits_s = torch.empty((len(secondary), 12), dtype=torch.int32, device=ctx.device)

    def best_of(fn, reps=6):  # the reference harness's protocol: no L2 flush, best of 5 after one warm-up launch
        best = 1e30
        for r in range(reps):
            e0, e1 = ev(), ev()
            e0.record(); fn(); e1.record(); e1.synchronize()
            if r > 0:
                best = min(best, e0.elapsed_time(e1))
        return best

    own_p = best_of(lambda: ctx.trace(geom, d_primary, d_hits_p, api.CLOSEST_HIT))
    own_s = best_of(lambda: ctx.trace(geom, d_secondary, d_hits_s, api.ANY_HIT))
    own_b = best_of(lambda: ctx.build_geometry(inp, geom_h, temp, flags))
    # the reference's traversal kernel through THIS repo's structure (the formats are byte-compatible): same BVH, their kernel
    try:
        _, x_p = ref.trace_geometry(geom_h, primary, kind=2, reps=6)
        _, x_s = ref.trace_geometry(geom_h, secondary, kind=1, reps=6)
        cross = {"mrays_per_s": round(n_rays / (x_p + x_s) / 1e3, 2), "primary_ms": round(x_p, 4), "secondary_ms": round(x_s, 4)}
    except Exception as e:
        cross = {"unavailable": str(e)[:120]}
    theirs, ours = n_rays / (ms_p + ms_s) / 1e3, n_rays / (own_p + own_s) / 1e3
    best_theirs = max(theirs, cross.get("mrays_per_s", 0.0))
    return {"mrays_per_s": round(theirs, 2), "build_mtris_per_s": round(n_tris / rg.build_ms / 1e3, 2), "build_ms": round(rg.build_ms, 4),
            "primary_ms": round(ms_p, 4), "secondary_ms": round(ms_s, 4),
            "protocol": "same rays, 48-byte hiprtHit out, warm L2, best of 5 launches after a warm-up (oracle/ref/ref_harness.cpp)",
            "this_repo_same_protocol": {"mrays_per_s": round(ours, 2), "primary_ms": round(own_p, 4), "secondary_ms": round(own_s, 4),
                                        "build_mtris_per_s": round(n_tris / own_b / 1e3, 2)},
            "hiprt_kernel_on_this_repos_bvh": cross,
            "trace_ratio": round(ours / best_theirs, 4), "build_ratio": round((n_tris / own_b) / (n_tris / rg.build_ms), 2),
            "note": "unmodified HIPRT v3.1.0 kernels via Orochi/CUDA on this GPU (oracle/_ref); its own build is always PreferFastBuild "
                    "(its H-PLOC does not run on sm_100); trace_ratio = this repo / the faster of (HIPRT on its own BVH, HIPRT's kernel on this repo's BVH)"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--ao-spp", type=int, default=16, help="AO rays per hit pixel of the *_ao workloads")
    ap.add_argument("--no-hiprt-reference", action="store_true", help="skip the reference's own GPU kernels (oracle/_ref) at N = 1")
    ap.add_argument("--with-hiprt-reference", action="store_true", help="(default at N = 1 when oracle/_ref is present; kept for old command lines)")
    ap.add_argument("--exchange", default="p2p", choices=["p2p", "nccl"], help="framebuffer exchange: peer stores from the resolve kernel, or dist.gather")
    ap.add_argument("--tile", type=int, default=32, help="edge of the image tiles that are dealt round-robin to the ranks")
    ap.add_argument("--no-graph", action="store_true", help="plain launches instead of one CUDA graph launch per frame")
    ap.add_argument("--replicate", default="broadcast", choices=["broadcast", "rebuild"], help="N > 1: how the BVH reaches every GPU")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    wl = WORKLOADS[args.workload]
    host_threads = os.cpu_count() or 1

    if args.impl == "reference" and rank != 0:
        return 0
    verts, idx = make_mesh(wl["mesh"])
    n_tris = len(idx)
    if args.impl == "reference":
        print(json.dumps(reference_arm(args, wl, verts, idx, host_threads)))
        return 0

    # ------------------------------------------------------------------------------------------
    import torch

    from hiprt_b200 import api, frame

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    ctx = api.Context(local_rank)
    dev = ctx.device
    flags = api.BALANCED_BUILD if wl.get("flags") == "balanced" else api.FAST_BUILD
    ev = lambda: torch.cuda.Event(enable_timing=True)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- the BVH on every GPU: built on rank 0 and broadcast (NCCL) + rebased, or rebuilt everywhere
    d_verts = torch.from_numpy(verts).to(dev)
    d_idx = torch.from_numpy(idx.view(np.int32)).to(dev)
    inp = ctx.mesh_input(d_verts, d_idx)
    temp = ctx.alloc_temp(ctx.geometry_temp_size(inp, flags))
    geom_h = ctx.create_geometry(inp, flags)
    ctx.build_geometry(inp, geom_h, temp, flags)
    geom = api.Geometry(ctx, geom_h, inp, temp, flags)
    torch.cuda.synchronize()
    replicate = {"how": "single GPU"}
    trace_geom = geom
    if world > 1 and args.replicate == "broadcast":
        handle, seconds = frame.replicate_geometry(ctx, geom if rank == 0 else None, dist, src=0)
        trace_geom = api.Geometry(ctx, handle, inp, temp, flags)
        nbytes = ctx.geometry_size(trace_geom)
        replicate = {"how": "built on GPU 0, ncclBroadcast of the blob + hiprtB200RelocateGeometry on the others", "bytes": nbytes,
                     "ms": round(seconds * 1e3, 3), "gbs": round(nbytes / seconds / 1e9, 1)}
    elif world > 1:
        replicate = {"how": "rebuilt on every GPU from the same input (deterministic)"}

    # ---- the frame, sharded
    fr = frame.AoFrame(ctx, trace_geom, WIDTH, HEIGHT, wl["eye"], wl["look"], spp=args.ao_spp, rank=rank, world=world, tile=args.tile, dist=dist,
                       exchange=args.exchange, light=None if wl.get("secondary") == "ao" else wl["light"])
    n_primary, n_secondary = fr.n_primary, fr.n_ao
    n_rays = n_primary + n_secondary
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # 256 MiB > 126 MB L2

    for _ in range(args.warmup):
        flush.fill_(1)
        fr.step()
    barrier()
    # the frame as ONE CUDA graph launch (4 kernels on two streams + 2 counter resets); every rank decides the same way
    graphed = False
    if not args.no_graph:
        try:
            graphed = fr.capture()
        except Exception as e:  # noqa: BLE001 -- capture is an optimisation, the plain launches are the fallback
            sys.stderr.write(f"[bench] rank {rank}: CUDA graph capture failed, plain launches: {e}\n")
            fr.graph = None
        if world > 1:
            ok = torch.tensor([1 if graphed else 0], dtype=torch.int32, device=dev)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            graphed = bool(ok.item())
            if not graphed:
                fr.graph = None
        for _ in range(2 if graphed else 0):
            flush.fill_(1)
            fr.step()
        barrier()
    sampler = ClockSampler(local_rank)
    t_wall0 = time.perf_counter()
    step_ms = []
    for _ in range(args.steps):
        flush.fill_(1)  # cold L2 for every step
        if world > 1:
            dist.barrier()  # aligns the start of the frame on all GPUs (outside the timed events)
        e0, e3 = ev(), ev()
        e0.record()
        fr.step()
        e3.record()
        e3.synchronize()
        step_ms.append(e0.elapsed_time(e3))
    barrier()
    wall_s = time.perf_counter() - t_wall0
    total_ms = float(np.sum(step_ms))
    image = fr.image() if rank == 0 else None
    # (diagnostic, untimed) the same frame with plain launches and an event after the two trace launches
    trace_ms = []
    for _ in range(3):
        flush.fill_(1)
        if world > 1:
            dist.barrier()
        e0, e1, e2 = ev(), ev(), ev()
        e0.record()
        fr.step(marks=(e1, e2))
        e2.synchronize()
        trace_ms.append(e0.elapsed_time(e2))
    barrier()

    # ---- the two trace launches one at a time (cold L2 each): what the roofline lines and the per-kernel figures are quoted on
    prim_ms, sec_ms = [], []
    for it in range(3 + args.steps):
        e0, e1, e2, e3 = ev(), ev(), ev(), ev()
        flush.fill_(1)
        e0.record()
        ctx.trace_format(trace_geom, fr.d_primary, fr.primary_format, fr.d_primary_rec, api.CLOSEST_HIT)
        e1.record()
        flush.fill_(1)
        e2.record()
        if n_secondary:
            ctx.trace_format(trace_geom, fr.d_ao, fr.ao_format, fr.d_ao_rec, api.ANY_HIT)
        e3.record()
        e3.synchronize()
        if it >= 3:
            prim_ms.append(e0.elapsed_time(e1))
            sec_ms.append(e2.elapsed_time(e3))

    # ---- build: K full builds of the same mesh (every internal launch, no allocation); one GPU's rate
    ctx.enable_build_profile(True)
    build_ms, phases = [], {}
    for it in range(args.warmup + args.steps):
        flush.fill_(1)
        e0, e1 = ev(), ev()
        e0.record()
        ctx.build_geometry(inp, geom_h, temp, flags)
        e1.record()
        e1.synchronize()
        if it >= args.warmup:
            build_ms.append(e0.elapsed_time(e1))
            for name, ms in ctx.build_profile():
                phases.setdefault(name, []).append(ms)
    ctx.enable_build_profile(False)
    clocks = sampler.stop()
    build_total_ms = float(np.sum(build_ms))

    # ---- end to end through the C-ABI with host buffers (pinned), H2D + D2H inside the timed region: this rank's rays
    pin = lambda a: torch.from_numpy(a.view(np.uint8).reshape(-1)).pin_memory().numpy().view(a.dtype)
    h_primary, h_secondary = pin(fr.primary), pin(fr.ao)
    h_rec_p = pin(np.zeros(n_primary, api.COMPACT_DTYPE))
    h_rec_s = pin(np.zeros(max(n_secondary, 1), np.dtype("<u4")))
    h_hits_p = pin(np.zeros(n_primary, api.HIT_DTYPE))
    h_hits_s = pin(np.zeros(max(n_secondary, 1), api.HIT_DTYPE))

    def e2e_compact():
        ctx.trace_format_host(trace_geom, h_primary, api.FORMAT_COMPACT, h_rec_p, api.CLOSEST_HIT)
        if n_secondary:
            ctx.trace_format_host(trace_geom, h_secondary, api.FORMAT_PRIM_ID, h_rec_s, api.ANY_HIT)

    def e2e_full():
        ctx.trace_host(trace_geom, h_primary, h_hits_p, api.CLOSEST_HIT)
        if n_secondary:
            ctx.trace_host(trace_geom, h_secondary, h_hits_s, api.ANY_HIT)

    def timed_host(fn, reps):
        total = 0.0
        for it in range(3 + reps):
            if world > 1:
                dist.barrier()
            t0 = time.perf_counter()
            fn()
            if it >= 3:
                total += (time.perf_counter() - t0) * 1e3
        return total

    e2e_total_ms = timed_host(e2e_compact, args.steps)
    e2e_full_ms = timed_host(e2e_full, min(args.steps, 5)) * args.steps / min(args.steps, 5)
    assert np.array_equal(h_rec_p["primID"], fr.hits0["primID"]), "host-buffer path disagrees with the device path"
    assert np.array_equal(h_hits_p["primID"], fr.hits0["primID"]), "host-buffer path (full hits) disagrees with the device path"

    # ---- max over ranks, sum of rays
    rays_all, prim_all, sec_all = float(n_rays), float(n_primary), float(n_secondary)
    if world > 1:
        t = torch.tensor([total_ms, e2e_total_ms, e2e_full_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms, e2e_total_ms, e2e_full_ms = [float(x) for x in t.tolist()]
        cnt = torch.tensor([n_rays, n_primary, n_secondary], dtype=torch.float64, device=dev)
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
        rays_all, prim_all, sec_all = [float(x) for x in cnt.tolist()]
        per_rank = torch.zeros(world, dtype=torch.float64, device=dev)
        per_rank[rank] = float(np.mean(trace_ms))
        dist.all_reduce(per_rank, op=dist.ReduceOp.SUM)
        per_rank_trace_ms = [round(float(x), 4) for x in per_rank.tolist()]
    else:
        per_rank_trace_ms = [round(float(np.mean(trace_ms)), 4)]

    if rank == 0:
        peak, peak_src = read_peaks()
        ms_per_step = total_ms / args.steps
        value = rays_all * args.steps / total_ms / 1e3  # Mrays/s, whole job
        prim_ms_avg, sec_ms_avg = float(np.mean(prim_ms)), float(np.mean(sec_ms))
        # the dominant kernel of the step: the longer of the two trace launches (timed alone, cold L2)
        sec_bytes_per_ray = 32 + 4 * api.FORMAT_WORDS[fr.ao_format]
        kernels = {"traceTrianglesKernel<closest-hit> (primary rays of GPU 0)": (n_primary, BYTES_PER_RAY, prim_ms_avg, "closest"),
                   "traceTrianglesKernel<any-hit, quantised box records> (secondary rays of GPU 0)": (n_secondary, sec_bytes_per_ray, sec_ms_avg, "any")}
        dominant = max(kernels, key=lambda k: kernels[k][2])
        dn, dbytes, dms, dtag = kernels[dominant]
        achieved = dn * dbytes / (dms * 1e-3) / 1e9
        traffic_all = (load_json("r02_traffic.json") or {}).get(args.workload, {})
        build_ms_avg = build_total_ms / args.steps
        n_prims = int(geom.header()["primNodeCount"])
        # build passes against the HBM roofline: algorithmic bytes per primitive pair, ONLY the phases that ran
        build_roofline, build_bytes = {}, 0
        per_prim = dict(BUILD_BYTES_PER_PRIM)
        per_prim["boxes"] = 56 + 12.0 * len(verts) / n_prims
        per_prim["qnodes"] = 192.0 * int(geom.header()["boxNodeCount"]) / n_prims
        for name, v in phases.items():
            ms = float(np.mean(v))
            nbytes = per_prim.get(name, 0) * n_prims
            build_bytes += nbytes
            build_roofline[name] = {"ms": round(ms, 4), "bytes_per_prim": round(per_prim.get(name, 0), 1),
                                    "achieved_gbs": round(nbytes / (ms * 1e-3) / 1e9, 1) if ms > 0 else 0.0,
                                    "frac": round(nbytes / (ms * 1e-3) / 1e9 / peak, 4) if ms > 0 else 0.0}
        sort_emit = [k for k in ("sort", "emit+refit", "hploc") if k in phases]
        se_bytes = sum(per_prim[k] for k in sort_emit) * n_prims
        se_ms = sum(float(np.mean(phases[k])) for k in sort_emit)
        ao = image[image >= 0]
        line = {
            "metric": metric_name(wl, args.ao_spp), "value": round(value, 2), "unit": "Mrays/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": round(ms_per_step, 4), "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": workload_config(args.workload, wl, n_tris, args.ao_spp),
            "rays_per_step": int(rays_all), "primary_rays": int(prim_all), "secondary_rays": int(sec_all),
            "engine": "libhiprt_b200.so: hiprtB200TraceGeometryFormat x2 + hiprtB200ResolveAoFrame per GPU"
                      + (" + hiprtB200WaitFlags on GPU 0" if fr.exchange == "p2p" else " + NCCL gather"),
            "detail": {"primitives_after_pairing": n_prims, "l2": "256 MiB flush write before every timed step",
                       "records": "primary: 48-byte hiprtHit; secondary: 4-byte primID (occlusion word); framebuffer: float32 per pixel on GPU 0",
                       "exchange": {"p2p": "resolve kernel stores into GPU 0's framebuffer over NVLink (CUDA IPC mapping) + release flag per GPU; "
                                           "GPU 0 waits on the flags inside the timed region"}.get(fr.exchange, fr.exchange),
                       "tile": args.tile, "bvh_replication": replicate,
                       "launch_overlap": "the closest-hit and the any-hit launch of a frame run on two streams (the first ends in a tail of a few "
                                         "long warps); per_gpu_trace_ms = both launches together inside the frame",
                       "frame_launch": "one CUDA graph launch per frame and GPU" if graphed else "plain launches",
                       "per_gpu_trace_ms_plain_launches": per_rank_trace_ms,
                       "gpu0_primary_ms_alone": round(prim_ms_avg, 4), "gpu0_secondary_ms_alone": round(sec_ms_avg, 4),
                       "gpu0_resolve_exchange_ms": round(max(float(np.mean(step_ms)) - float(np.mean(trace_ms)), 0.0), 4),
                       "image_check": {"hit_pixels": int(ao.size), "mean_ao": round(float(ao.mean()), 5) if ao.size else None}},
            "build": {"mtris_per_s": round(n_tris * args.steps / build_total_ms / 1e3, 2), "ms_per_build": round(build_ms_avg, 4),
                      "scope": "one GPU (a single large build is replicas-only)",
                      "bytes_per_triangle_model": round(build_bytes / n_tris, 1), "achieved_gbs": round(build_bytes / (build_ms_avg * 1e-3) / 1e9, 1),
                      "frac_of_hbm_peak": round(build_bytes / (build_ms_avg * 1e-3) / 1e9 / peak, 4),
                      "sort_plus_hierarchy": {"phases": sort_emit, "ms": round(se_ms, 4), "frac_of_hbm_peak": round(se_bytes / (se_ms * 1e-3) / 1e9 / peak, 4) if se_ms else None,
                                              "frac_of_8TBs_spec": round(se_bytes / (se_ms * 1e-3) / 1e9 / 8000.0, 4) if se_ms else None},
                      "phases": build_roofline},
            "roofline": {"kernel": dominant, "bound": "hbm", "achieved": round(achieved, 2), "peak": peak,
                         "unit": "GB/s", "frac": round(achieved / peak, 5), "traffic": traffic_all.get(f"trace_{dtag}_dram_bytes"),
                         "peak_source": peak_src, "launch_ms": round(dms, 4), "rays_per_launch": int(dn), "algorithmic_bytes_per_ray": dbytes,
                         "other_kernels": {k: {"launch_ms": round(v[2], 4), "achieved": round(v[0] * v[1] / (v[2] * 1e-3) / 1e9, 2) if v[2] > 0 else None,
                                               "algorithmic_bytes_per_ray": v[1]} for k, v in kernels.items() if k != dominant},
                         "note": "compulsory I/O per ray (32 B hiprtRay in + the output record: 48 B hiprtHit / 4 B occlusion word) x rays / launch "
                                 "time; traversal is bound by L1 data-pipe wavefronts and issue slots, not HBM: cache_level below is the roofline "
                                 "that explains the kernels"},
            "e2e": {"value": round(rays_all * args.steps / e2e_total_ms / 1e3, 2), "unit": "Mrays/s", "h2d_bytes_per_step": int(32 * rays_all),
                    "d2h_bytes_per_step": int(16 * prim_all + 4 * sec_all),
                    "api": "hiprtB200TraceGeometryFormatHost (pinned host buffers; 16-byte {primID,t,u,v} records for primary rays, 4-byte occlusion words for secondary rays)",
                    "full_hits": {"value": round(rays_all * args.steps / e2e_full_ms / 1e3, 2), "d2h_bytes_per_step": int(48 * rays_all),
                                  "api": "hiprtB200TraceGeometryHost (48-byte hiprtHit for every ray)"}},
            # timed frames: 2 trace kernels + resolve (+ the flag wait on GPU 0) per step and GPU; timed single launches: 2 per step on
            # GPU 0; timed builds: init, pair (indexed meshes), boxes, morton, sort passes, emit or hploc, collapse, qnodes (memsets not counted)
            "gpu_launches": args.steps * fr.launches_per_step * world + 2 * args.steps + args.steps * (len(phases) + 3),
            "clocks": clocks,
            "wall_s_timed_region": round(wall_s, 3),
        }
        if world == 1:
            # cache-level traffic: node visits counted by the instrumented kernel, against the measured gather peaks
            try:
                bp, lp = ctx.trace_stats(trace_geom, fr.d_primary, api.CLOSEST_HIT)
                bs, ls = ctx.trace_stats(trace_geom, fr.d_ao, api.ANY_HIT) if n_secondary else (0, 0)
                peaks = load_json("r02_cache_peaks.json") or {}
                cl = {"box_visits_per_primary_ray": round(bp / n_primary, 3), "leaf_visits_per_primary_ray": round(lp / n_primary, 3),
                      "box_visits_per_secondary_ray": round(bs / max(n_secondary, 1), 3), "leaf_visits_per_secondary_ray": round(ls / max(n_secondary, 1), 3),
                      "bytes_model": "box_node_bytes x box visits + 64 B x leaf visits (SURVEY.md 8d), counted by hiprtB200TraceGeometryStats"}
                # any-hit launches walk the 64-byte quantised records (csrc/qnode.h), closest-hit launches the public 128-byte nodes
                q_any = os.environ.get("HIPRT_B200_TRACE_VARIANT") not in ("1", "2", "3")
                for tag, b, l, ms, node_bytes in (("primary", bp, lp, prim_ms_avg, 128), ("secondary", bs, ls, sec_ms_avg, 64 if q_any else 128)):
                    gbs = (node_bytes * b + 64 * l) / (ms * 1e-3) / 1e9 if ms > 0 else 0.0
                    cl[tag] = {"achieved_gbs": round(gbs, 1), "box_node_bytes": node_bytes}
                    for level in ("l1_resident_gather_gbs", "l2_resident_gather_gbs"):
                        if peaks.get(level):
                            cl[tag]["frac_of_" + level] = round(gbs / peaks[level], 4)
                cl["peaks"] = {k: peaks.get(k) for k in ("l1_resident_gather_gbs", "l2_resident_gather_gbs", "hbm_gather_gbs")} if peaks else \
                    "profiles/r02_cache_peaks.json missing: run hiprt_b200/csrc/build/cache_peaks"
                line["roofline"]["cache_level"] = cl
            except Exception as e:
                line["roofline"]["cache_level"] = {"unavailable": str(e)[:200]}
        if world == 1 and not args.no_cpu_baseline:
            r = embree_arm(verts, idx, wl, min(args.steps, 10), 2, host_threads, args.ao_spp)
            if r is not None:
                line["cpu_baseline"] = {"value": round(r["mrays"], 3), "unit": "Mrays/s", "cores": host_threads, "kind": "reference",
                                        "sample": f"full frame ({r['rays']} rays/step) x {min(args.steps, 10)} steps on Embree 4.3 bundled with the reference "
                                                  f"(build quality {r['quality']})",
                                        "build_mtris_per_s": round(r["build_mtris"], 3)}
                # cross-check against Embree: same primitive away from ties
                line["cpu_baseline"]["primID_agreement_with_gpu"] = round(float((r["hits"]["primID"] == fr.hits0["primID"]).mean()), 6)
            else:
                r = oracle_port_arm(verts, idx, wl)
                line["cpu_baseline"] = {"value": round(r["mrays"], 4), "unit": "Mrays/s", "cores": 1, "kind": "port",
                                        "sample": f"{r['rays']} primary rays (every k-th pixel) through the oracle", "build_mtris_per_s": round(r["build_mtris"], 4)}
        if world == 1 and not args.no_hiprt_reference:
            try:
                from oracle import ref_binding as ref

                if ref.available():
                    line["hiprt_reference_gpu"] = hiprt_reference_gpu(ctx, api, geom, geom_h, inp, temp, flags, verts, idx, fr.primary, fr.ao,
                                                                      fr.d_primary, fr.d_ao, n_tris, ev)
                else:
                    line["hiprt_reference_gpu"] = {"unavailable": "oracle/_ref is not built (make -C oracle/ref needs /root/reference)"}
            except Exception as e:  # the reference is optional infrastructure
                line["hiprt_reference_gpu"] = {"unavailable": str(e)[:200]}
        if world == 1:
            # (untimed side figure, N = 1) the third builder of the API on the same mesh: PreferHighQualityBuild = SBVH, and the
            # reference's own SBVH when oracle/_ref is there (tools/bench_sbvh.py compares the trees as well)
            try:
                hq = ctx.build(d_verts, d_idx, flags=api.HIGH_QUALITY_BUILD)
                best = 1e30
                for _ in range(3):
                    e0, e1 = ev(), ev()
                    e0.record(); ctx.build_geometry(hq.inp, hq.h, hq.temp, api.HIGH_QUALITY_BUILD); e1.record(); e1.synchronize()
                    best = min(best, e0.elapsed_time(e1))
                line["build"]["high_quality"] = {"ms_per_build": round(best, 3), "mtris_per_s": round(n_tris / best / 1e3, 1),
                                                 "leaves": int(hq.header()["primNodeCount"]), "builder": "SBVH: binned SAH + spatial splits, one cooperative launch"}
                hq.destroy()
                del hq
                ref_line = line.get("hiprt_reference_gpu", {})
                if "unavailable" not in ref_line and not args.no_hiprt_reference:
                    from oracle import ref_binding as ref

                    rg = ref.RefGeometry(verts, idx, build_flags=api.HIGH_QUALITY_BUILD, reps=2)
                    ref_line["high_quality_build_ms"] = round(rg.build_ms, 3)
                    ref_line["high_quality_build_ratio"] = round(rg.build_ms / best, 2)
                    rg.destroy()
            except Exception as e:
                line["build"]["high_quality"] = {"unavailable": str(e)[:200]}
        print(json.dumps(line))
    fr.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
