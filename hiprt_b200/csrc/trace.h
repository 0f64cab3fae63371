// Launchers of the library's trace kernels (trace_kernels.cu).
#pragma once
#include <cuda_runtime.h>
#include <hiprt/hiprt_types.h>

namespace hiprt
{
void traceGeometry( hiprtGeometry geom, const hiprtRay* rays, uint32_t numRays, hiprtHit* hits, hiprtTraversalType type, cudaStream_t stream );
void traceScene(
	hiprtScene scene, const hiprtRay* rays, uint32_t numRays, hiprtHit* hits, hiprtTraversalType type, hiprtRayMask mask, cudaStream_t stream );
} // namespace hiprt
