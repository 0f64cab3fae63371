"""Config C5, render half: diffuse path tracing (cosine sampling, fixed maximum depth, no Russian roulette) as a WAVEFRONT
on top of the library's rays-in / hits-out entry points -- every bounce is one `hiprtB200TraceGeometry/Scene` launch over the
rays that are still alive, the bounce itself (hit point, face-forward normal, cosine-hemisphere direction, throughput) is a
handful of torch ops on the device. Pixels are given as a list, so a rank renders exactly the pixels of its image tiles
(hiprt_b200/multi_gpu.py) and nothing is exchanged until the framebuffer gather.

The random numbers are a counter-based integer hash of (pixel, sample, bounce, dimension): a ray depends only on its pixel and
sample index, never on which rank or in which order it is traced, so an image rendered on N GPUs is bit-identical to the one
rendered on one.
"""
import numpy as np

HIT_WORDS = 12  # hiprtHit as int32 words: instanceIDs[4], primID, pad, uv[2], normal[3], t


def _hash_u01(torch, a, b, c, d):
    """Counter-based hash of four int64 tensors / scalars -> float32 uniform in [0, 1)."""
    m = 0xFFFFFFFF
    x = (a * 0x9E3779B1 + b * 0x85EBCA77 + c * 0xC2B2AE3D + d * 0x27D4EB2F + 0x165667B1) & m
    x = ((x ^ (x >> 15)) * 0x2C1B3C6D) & m
    x = ((x ^ (x >> 12)) * 0x297A2D39) & m
    x = x ^ (x >> 15)
    return (x >> 8).to(torch.float32) * (1.0 / 16777216.0)


def camera_rays(torch, pixels, width, height, eye, look_at, sample, up=(0.0, 1.0, 0.0), fov_deg=45.0):
    """Jittered pinhole rays for row-major pixel indices `pixels` (int64 device tensor). Returns float32 [n, 8] hiprtRay rows."""
    dev = pixels.device
    f64 = np.float64
    eye_n, look = np.asarray(eye, f64), np.asarray(look_at, f64)
    w = look - eye_n
    w /= np.linalg.norm(w)
    u = np.cross(w, np.asarray(up, f64))
    u /= np.linalg.norm(u)
    v = np.cross(u, w)
    half = np.tan(np.radians(fov_deg) / 2)
    px, py = (pixels % width), (pixels // width)
    jx, jy = _hash_u01(torch, pixels, sample, 0, 1), _hash_u01(torch, pixels, sample, 0, 2)
    sx = ((px.to(torch.float32) + jx) / width * 2 - 1) * (half * width / height)
    sy = (1 - (py.to(torch.float32) + jy) / height * 2) * half
    t32 = lambda a: torch.tensor(a, dtype=torch.float32, device=dev)
    d = sx[:, None] * t32(u) + sy[:, None] * t32(v) + t32(w)
    d = d / d.norm(dim=1, keepdim=True)
    rays = torch.empty((len(pixels), 8), dtype=torch.float32, device=dev)
    rays[:, 0:3] = t32(eye_n)
    rays[:, 3] = 0.0
    rays[:, 4:7] = d
    rays[:, 7] = 3.0e38
    return rays


def render(ctx, target, pixels, width, height, eye, look_at, spp=4, max_depth=5, albedo=0.7, record=None, stats=None):
    """Radiance (float32 [n, 3], device) of the pixel list `pixels` (int64 device tensor of row-major indices).
    target: api.Geometry or api.Scene. record: optional list that receives (rays, hits) device tensors of every launch
    (parity tests trace the same rays through the oracle). stats: optional dict, 'rays' is incremented by the rays traced."""
    torch = ctx.torch
    dev = pixels.device
    n = len(pixels)
    radiance = torch.zeros((n, 3), dtype=torch.float32, device=dev)
    sky_lo, sky_hi = torch.tensor([1.0, 1.0, 1.0], device=dev), torch.tensor([0.5, 0.7, 1.0], device=dev)
    for s in range(spp):
        rays = camera_rays(torch, pixels, width, height, eye, look_at, s)
        slot = torch.arange(n, device=dev)  # pixel slot of every live ray
        throughput = torch.ones(n, dtype=torch.float32, device=dev)
        for depth in range(max_depth):
            if len(rays) == 0:
                break
            hits = ctx.trace(target, rays.view(torch.uint8).reshape(-1, 32))
            if stats is not None:
                stats["rays"] = stats.get("rays", 0) + len(rays)
            if record is not None:
                record.append((rays.clone(), hits.clone()))
            hf = hits.view(torch.float32).reshape(-1, HIT_WORDS)
            hit = hits.reshape(-1, HIT_WORDS)[:, 4] != -1  # primID != hiprtInvalidValue
            d = rays[:, 4:7]
            # rays that leave the scene pick up the sky
            miss = ~hit
            if miss.any():
                tsky = (0.5 * (d[miss, 1] + 1.0)).clamp(0, 1)[:, None]
                radiance.index_add_(0, slot[miss], throughput[miss, None] * ((1 - tsky) * sky_lo + tsky * sky_hi))
            if not hit.any() or depth + 1 == max_depth:
                break
            rays, hf, slot, throughput, d = rays[hit], hf[hit], slot[hit], throughput[hit] * albedo, d[hit]
            nrm = hf[:, 8:11]
            nrm = nrm / nrm.norm(dim=1, keepdim=True).clamp_min(1e-30)
            nrm = torch.where(((nrm * d).sum(1) > 0)[:, None], -nrm, nrm)
            p = rays[:, 0:3] + d * hf[:, 11:12] + nrm * 1e-4
            pix = pixels[slot]
            u0, u1 = _hash_u01(torch, pix, s, depth + 1, 1), _hash_u01(torch, pix, s, depth + 1, 2)
            phi, sin_t = 6.2831853 * u0, u1.sqrt()
            axis = torch.where((nrm[:, 0].abs() > 0.5)[:, None], torch.tensor([0.0, 1.0, 0.0], device=dev), torch.tensor([1.0, 0.0, 0.0], device=dev))
            tng = torch.linalg.cross(axis, nrm)
            tng = tng / tng.norm(dim=1, keepdim=True)
            btg = torch.linalg.cross(nrm, tng)
            nd = tng * (phi.cos() * sin_t)[:, None] + btg * (phi.sin() * sin_t)[:, None] + nrm * (1 - u1).clamp_min(0).sqrt()[:, None]
            rays = torch.empty((len(p), 8), dtype=torch.float32, device=dev)
            rays[:, 0:3], rays[:, 3], rays[:, 4:7], rays[:, 7] = p, 0.0, nd / nd.norm(dim=1, keepdim=True), 3.0e38
    return radiance / spp


# ---------------------------------------------------------------------------------------------------------------------------
# The same renderer with the bounce fused into one CUDA kernel (hiprt_b200/pathtrace_kernels.cu, compiled through
# hiprtBuildTraceKernels like an application's kernels): per bounce one trace launch + one bounce launch, no torch glue.
import ctypes as C
import os

_BOUNCE_SRC = os.path.join(os.path.dirname(os.path.abspath(__file__)), "pathtrace_kernels.cu")
_cuda = None


def _launch(fn, n_threads, block, args, stream=None):
    """cuLaunchKernel in the current CUDA context on `stream` (a CUstream as c_void_p; None: the default stream). args: ctypes values."""
    global _cuda
    if _cuda is None:
        _cuda = C.CDLL("libcuda.so.1")
        _cuda.cuLaunchKernel.argtypes = [C.c_void_p] + [C.c_uint] * 6 + [C.c_uint, C.c_void_p, C.POINTER(C.c_void_p), C.c_void_p]
    ptrs = (C.c_void_p * len(args))(*[C.cast(C.pointer(a), C.c_void_p) for a in args])
    rc = _cuda.cuLaunchKernel(fn, (n_threads + block - 1) // block, 1, 1, block, 1, 1, 0, stream, ptrs, None)
    if rc != 0:
        raise RuntimeError(f"cuLaunchKernel failed: {rc}")


def _kernels(ctx):
    """(bounce kernel, camera kernel) of pathtrace_kernels.cu, compiled once per context through hiprtBuildTraceKernels."""
    fns = getattr(ctx, "_pathtrace_kernels", None)
    if fns is None:
        from . import api

        names = (C.c_char_p * 2)(b"diffuseBounceKernel", b"cameraRaysKernel")
        out = (C.c_void_p * 2)()
        module = C.c_void_p()
        ctx._check(api.lib().hiprtBuildTraceKernels(ctx.h, 2, names, open(_BOUNCE_SRC).read().encode(), b"pathtrace_kernels", 0, None, None, 0, None,
                                                    0, 1, None, out, C.byref(module), False), "hiprtBuildTraceKernels")
        fns = ctx._pathtrace_kernels = (C.c_void_p(out[0]), C.c_void_p(out[1]))
    return fns


class _Float3(C.Structure):
    _fields_ = [("x", C.c_float), ("y", C.c_float), ("z", C.c_float)]


def _camera_basis(eye, look_at, up=(0.0, 1.0, 0.0), fov_deg=45.0):
    f64 = np.float64
    eye_n, look = np.asarray(eye, f64), np.asarray(look_at, f64)
    w = look - eye_n
    w /= np.linalg.norm(w)
    u = np.cross(w, np.asarray(up, f64))
    u /= np.linalg.norm(u)
    v = np.cross(u, w)
    f3 = lambda a: _Float3(*[float(np.float32(x)) for x in a])
    return f3(eye_n), f3(u), f3(v), f3(w), float(np.tan(np.radians(fov_deg) / 2))


def render_fused(ctx, target, pixels, width, height, eye, look_at, spp=4, max_depth=5, albedo=0.7, stats=None, batch_paths=1 << 25):
    """As render(), as a queue of launches the host never waits for. Samples are traced in BATCHES: a wavefront holds the paths of
    `g` samples of every pixel (g * len(pixels) <= batch_paths), because the deep bounces of one sample of a small pixel list are a
    few thousand long rays and a launch there is bound by its slowest warp, not by its ray count: eight ranks that
    render an eighth of the image each would pay that tail 64 x 2 times per frame, batched they pay it 4 x 2 times. Per batch one
    camera launch, per bounce one INDIRECT trace launch (hiprtB200TraceGeometryIndirect: the ray count is the device word the
    previous bounce's append counter left) and one bounce launch; every (sample-in-batch, pixel) has its own radiance slot, summed
    at the end, so the image does not depend on the order in which paths finish. Nothing is read back until the caller looks at
    the image; stats['rays'] (one read-back at the end) is the sum of the per-launch counts. The order of the rays inside a
    wavefront depends on the order in which warps append, so images agree with render() up to float rounding of the bounce math,
    not bit for bit."""
    from . import api

    torch = ctx.torch
    dev = pixels.device
    n_pix = len(pixels)
    bounce, camera = _kernels(ctx)
    g = max(1, min(spp, batch_paths // max(n_pix, 1)))  # samples per batch
    batches = [(s0, min(g, spp - s0)) for s0 in range(0, spp, g)]
    cap = g * n_pix
    radiance = torch.zeros((g, n_pix, 3), dtype=torch.float32, device=dev)
    rays = [torch.empty((cap, 8), dtype=torch.float32, device=dev) for _ in range(2)]
    slots = [torch.empty(cap, dtype=torch.int32, device=dev) for _ in range(2)]
    thr = [torch.empty(cap, dtype=torch.float32, device=dev) for _ in range(2)]
    hits = torch.empty((cap, HIT_WORDS), dtype=torch.int32, device=dev)
    counts = torch.zeros(len(batches) * (max_depth + 1), dtype=torch.int32, device=dev)  # live paths before every launch, one word each
    pixels = pixels.contiguous()
    p = lambda t, off=0: C.c_void_p(t.data_ptr() + off)
    eye3, u3, v3, w3, half = _camera_basis(eye, look_at)
    handle = target.h if hasattr(target, "h") else target
    stream = api._stream_handle(None)
    timed = stats is not None and stats.get("time_trace")
    import time as _time

    host0 = _time.perf_counter()
    for bi, (s0, gs) in enumerate(batches):
        k = bi * (max_depth + 1)
        n = gs * n_pix
        _launch(camera, n, 128, [p(pixels), C.c_uint32(n_pix), C.c_uint32(n), C.c_uint32(width), C.c_uint32(height), eye3, u3, v3, w3, C.c_float(half),
                                 C.c_uint32(s0), p(rays[0]), p(slots[0]), p(thr[0]), p(counts, 4 * k)], stream)
        cur = 0
        for depth in range(max_depth):
            if timed:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
            ctx._check(api.lib().hiprtB200TraceGeometryIndirect(ctx.h, handle, p(rays[cur]), n, p(counts, 4 * (k + depth)), p(hits), api.FORMAT_FULL,
                                                                api.CLOSEST_HIT, stream), "hiprtB200TraceGeometryIndirect")
            if timed:
                e1.record()
                e1.synchronize()
                stats["trace_ms"] = stats.get("trace_ms", 0.0) + e0.elapsed_time(e1)
                stats.setdefault("per_depth_ms", [0.0] * max_depth)[depth] += e0.elapsed_time(e1)
            last = int(depth + 1 == max_depth)
            _launch(bounce, n, 128, [p(rays[cur]), p(hits), p(slots[cur]), p(thr[cur]), p(counts, 4 * (k + depth)), p(pixels), C.c_uint32(n_pix), C.c_uint32(s0),
                                     C.c_uint32(depth), C.c_float(albedo), C.c_int(last), p(rays[1 - cur]), p(slots[1 - cur]), p(thr[1 - cur]),
                                     p(counts, 4 * (k + depth + 1)), p(radiance)], stream)
            cur = 1 - cur
    if stats is not None:
        stats["host_enqueue_ms"] = stats.get("host_enqueue_ms", 0.0) + (_time.perf_counter() - host0) * 1e3
        stats["samples_per_batch"] = g
        per = counts.view(len(batches), max_depth + 1)[:, :max_depth].to(torch.int64).sum(0).cpu()
        stats["rays"] = stats.get("rays", 0) + int(per.sum())
        stats["per_depth_rays"] = [int(x) for x in per]
    return radiance.sum(0) / spp
