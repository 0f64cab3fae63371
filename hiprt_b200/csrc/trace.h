// Launchers of the library's trace kernels (trace_kernels.cu).
#pragma once
#include <cuda_runtime.h>
#include <hiprt/hiprt_b200.h>
#include <hiprt/hiprt_types.h>

namespace hiprt
{
// workCounter: one device uint32 the persistent variant uses as its ray-batch counter (reset on `stream` before the launch)
// triangles: the geometry holds triangle pairs (fast kernel) rather than custom primitives (generic kernel)
// format: hiprtB200HitFormat of the output records (`hits` then points to 48-, 16- or 4-byte records); formats other than the
// full hiprtHit exist for triangle geometries only
// qRegion: the geometry's traversal-only Q region (qnode.h) when it has one -- box nodes are then read from its 64-byte records
// numRaysDevice: optional device word holding the actual ray count (<= numRays, which then is the capacity): triangle geometries only
void traceGeometry(
	hiprtGeometry geom, bool triangles, const hiprtRay* rays, uint32_t numRays, void* hits, hiprtTraversalType type, uint32_t* workCounter, uint32_t smCount,
	cudaStream_t stream, uint32_t format = 0, const void* qRegion = nullptr, const uint32_t* numRaysDevice = nullptr );
// instrumented variant of the triangle kernel: adds {box-node visits, leaf visits} of the launch to stats[0..1]
void traceGeometryStats(
	hiprtGeometry geom, const hiprtRay* rays, uint32_t numRays, hiprtHit* hits, hiprtTraversalType type, uint32_t* workCounter, uint32_t smCount, cudaStream_t stream,
	unsigned long long* stats, const void* qRegion = nullptr );
// times: optional device array of one time per ray (motion blur); nullptr = every ray at `time`
void traceScene(
	hiprtScene scene, const hiprtRay* rays, uint32_t numRays, hiprtHit* hits, hiprtTraversalType type, hiprtRayMask mask, uint32_t* workCounter, uint32_t smCount,
	cudaStream_t stream, const float* times = nullptr, float time = 0.0f );
} // namespace hiprt
