// Multi-GPU frame plumbing (include/hiprt/hiprt_b200.h "multi-GPU plumbing"): the AO resolve fused with the framebuffer
// exchange over peer memory, the completion wait, and the header rebase of a replicated geometry blob.
// The reference has no counterpart (single GPU: SURVEY.md section 5, "Distributed communication backend").
#include "frame.h"

#include <hiprt/hiprt_b200_layout.h>

namespace hiprt
{
// One thread per local pixel. The `spp` occlusion words of a pixel are consecutive (spp = 16: one 64-byte line, read as 4 x 16 B).
// framebuffer / doneFlag may be peer mappings: plain stores to them are NVLink writes, posted and pipelined; the completion
// flag is written by the LAST block to finish (device-scope counter), after a system-scope fence, so that the owner that
// observes the flag also observes every pixel this GPU stored.
__global__ void __launch_bounds__( 256 ) resolveAoFrameKernel(
	const uint8_t* __restrict__ primaryRecords,
	uint32_t primaryStride,
	uint32_t primIdOffset,
	const uint32_t* __restrict__ aoPrimIDs,
	const uint32_t* __restrict__ firstAoRay,
	const uint32_t* __restrict__ pixelIndex,
	uint32_t  numPixels,
	uint32_t  spp,
	float*	  framebuffer,
	uint32_t* doneFlag,
	uint32_t  doneValue,
	uint32_t* blockCounter )
{
	for ( uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < numPixels; i += gridDim.x * blockDim.x )
	{
		const uint32_t prim = __ldcs( reinterpret_cast<const uint32_t*>( primaryRecords + size_t( i ) * primaryStride + primIdOffset ) );
		float		   ao	= -1.0f;
		if ( prim != InvalidValue )
		{
			const uint32_t first = firstAoRay[i];
			uint32_t	   open	 = 0;
			if ( ( spp & 3u ) == 0 && ( first & 3u ) == 0 )
			{
				const uint4* q = reinterpret_cast<const uint4*>( aoPrimIDs + first );
				for ( uint32_t k = 0; k < spp / 4; ++k )
				{
					const uint4 v = __ldcs( q + k );
					open += uint32_t( v.x == InvalidValue ) + uint32_t( v.y == InvalidValue ) + uint32_t( v.z == InvalidValue ) + uint32_t( v.w == InvalidValue );
				}
			}
			else
				for ( uint32_t k = 0; k < spp; ++k )
					open += uint32_t( __ldcs( aoPrimIDs + first + k ) == InvalidValue );
			ao = float( open ) / float( spp );
		}
		framebuffer[pixelIndex[i]] = ao;
	}
	if ( doneFlag == nullptr ) return;
	__threadfence_system(); // this thread's pixel stores are visible system-wide before the block counts itself in
	__syncthreads();
	if ( threadIdx.x == 0 )
	{
		if ( atomicAdd( blockCounter, 1u ) + 1u == gridDim.x )
		{
			*blockCounter = 0; // ready for the next frame
			// doneValue == 0: the frame number lives on the device (blockCounter[1]) and counts the launches -- a launch recorded
			// in a CUDA graph carries the same arguments every replay
			if ( doneValue == 0u ) doneValue = ++blockCounter[1];
			__threadfence_system();
			asm volatile( "st.release.sys.global.u32 [%0], %1;" ::"l"( doneFlag ), "r"( doneValue ) : "memory" );
		}
	}
}

void resolveAoFrame(
	const void* primaryRecords, uint32_t primaryStride, uint32_t primIdOffset, const uint32_t* aoPrimIDs, const uint32_t* firstAoRay, const uint32_t* pixelIndex,
	uint32_t numPixels, uint32_t spp, float* framebuffer, uint32_t* doneFlag, uint32_t doneValue, uint32_t* blockCounter, uint32_t smCount, cudaStream_t stream )
{
	const uint32_t blocks = std::max( 1u, std::min( ( numPixels + 255u ) / 256u, smCount * 8u ) );
	resolveAoFrameKernel<<<blocks, 256, 0, stream>>>(
		static_cast<const uint8_t*>( primaryRecords ), primaryStride, primIdOffset, aoPrimIDs, firstAoRay, pixelIndex, numPixels, spp, framebuffer, doneFlag,
		doneValue, blockCounter );
}

// Spins (bounded) until every flag has reached `value`; acquire loads at system scope pair with the release stores above.
__global__ void waitFlagsKernel( const uint32_t* flags, uint32_t count, uint32_t strideWords, uint32_t value, uint32_t* timeoutOut )
{
	const uint32_t i = threadIdx.x;
	bool		   ok = true;
	// value == 0: the expected frame number lives on the device (timeoutOut[1]) and counts the launches (CUDA graph replays)
	const bool counting = value == 0u;
	if ( counting ) value = timeoutOut[1] + 1u;
	__syncthreads();
	if ( counting && i == 0 ) timeoutOut[1] = value;
	if ( i < count )
	{
		ok = false;
		for ( uint32_t spins = 0; spins < ( 1u << 24 ); ++spins ) // ~ seconds: never hang the device
		{
			uint32_t v;
			asm volatile( "ld.acquire.sys.global.u32 %0, [%1];" : "=r"( v ) : "l"( flags + size_t( i ) * strideWords ) : "memory" );
			if ( int32_t( v - value ) >= 0 )
			{
				ok = true;
				break;
			}
			__nanosleep( 100 );
		}
	}
	if ( !ok && timeoutOut ) *timeoutOut = 1u;
}

void waitFlags( const uint32_t* flags, uint32_t count, uint32_t strideWords, uint32_t value, uint32_t* timeoutOut, cudaStream_t stream )
{
	waitFlagsKernel<<<1, std::max( 32u, ( count + 31u ) & ~31u ), 0, stream>>>( flags, count, strideWords, value, timeoutOut );
}

// A geometry blob copied to a new address: its header holds absolute pointers to its own box-node and leaf arrays.
__global__ void relocateGeometryKernel( GeomHeader* header )
{
	uint8_t* const	base	= reinterpret_cast<uint8_t*>( header );
	const uintptr_t oldBox	= reinterpret_cast<uintptr_t>( header->boxNodes );
	const uintptr_t oldPrim = reinterpret_cast<uintptr_t>( header->primNodes );
	// the box nodes directly follow the 64-byte header slot in every layout this library (and the reference) writes
	constexpr uintptr_t BoxOffset = 64;
	header->boxNodes  = reinterpret_cast<Box4Node*>( base + BoxOffset );
	header->primNodes = base + BoxOffset + ( oldPrim - oldBox );
}

void relocateGeometry( void* geometry, cudaStream_t stream ) { relocateGeometryKernel<<<1, 1, 0, stream>>>( static_cast<GeomHeader*>( geometry ) ); }
} // namespace hiprt
