// Launchers of the multi-GPU frame kernels (frame_kernels.cu).
#pragma once
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>

namespace hiprt
{
void resolveAoFrame(
	const void* primaryRecords, uint32_t primaryStride, uint32_t primIdOffset, const uint32_t* aoPrimIDs, const uint32_t* firstAoRay, const uint32_t* pixelIndex,
	uint32_t numPixels, uint32_t spp, float* framebuffer, uint32_t* doneFlag, uint32_t doneValue, uint32_t* blockCounter, uint32_t smCount, cudaStream_t stream );
void waitFlags( const uint32_t* flags, uint32_t count, uint32_t strideWords, uint32_t value, uint32_t* timeoutOut, cudaStream_t stream );
void relocateGeometry( void* geometry, cudaStream_t stream );
} // namespace hiprt
