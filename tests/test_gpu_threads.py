"""Concurrent use of one context from several host threads, each with its own stream (reference: the MultiThreading tests
of test/hiprtTest.cpp:1226-1345 build from up to 16 threads on one context): builds, traces through the library kernels and
destroys must not interfere -- every thread's structure equals the one a serial build produces."""
import threading

import numpy as np
import pytest

from hiprt_b200 import api, scenes
from oracle import binding as orc
from tests.test_gpu_parity import compare_hits, to_dev

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = api.Context(0)
    yield c
    c.close()


def test_builds_and_traces_from_eight_threads(ctx):
    t = ctx.torch
    n_threads, rounds = 8, 6
    meshes = [scenes.sphere_mesh(40 + 8 * k, 20 + 4 * k) for k in range(n_threads)]
    rays = scenes.pinhole_rays(96, 96, (0.0, 0.0, 3.0), (0, 0, 0))
    expected = [orc.Geometry(v, i).trace(rays, tie_mode=1) for v, i in meshes]
    errors, results = [], [None] * n_threads

    def worker(k):
        try:
            stream = t.cuda.Stream(device=ctx.device)
            v, i = meshes[k]
            with t.cuda.stream(stream):
                dv, di = to_dev(ctx, v), to_dev(ctx, i)
                d_rays = ctx.rays_to_device(rays)
                for r in range(rounds):
                    inp = ctx.mesh_input(dv, di)
                    flags = api.FAST_BUILD if (r + k) % 2 == 0 else api.BALANCED_BUILD
                    temp = ctx.alloc_temp(ctx.geometry_temp_size(inp, flags))
                    g = ctx.create_geometry(inp, flags)
                    ctx.build_geometry(inp, g, temp, flags, stream=stream)
                    geom = api.Geometry(ctx, g, inp, temp, flags)
                    hits = ctx.trace(geom, d_rays, stream=stream)
                    stream.synchronize()
                    results[k] = ctx.hits_to_host(hits)
                    geom.destroy()
        except Exception as e:  # surfaced in the main thread
            errors.append((k, repr(e)))

    threads = [threading.Thread(target=worker, args=(k,)) for k in range(n_threads)]
    for th in threads:
        th.start()
    for th in threads:
        th.join()
    assert not errors, errors
    for k in range(n_threads):
        compare_hits(results[k], expected[k])
