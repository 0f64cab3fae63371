// Library-provided trace kernels (rays in -> hits out), built from the public device API in
// include/hiprt/hiprt_device.h exactly as a user kernel would be (reference counterpart:
// test/kernels/HiprtTestKernel.h:102-116 TraceKernel).
//
// Stack: the top of each thread's stack lives in shared memory (lane-interleaved, so a warp's
// accesses hit 32 distinct banks), deeper entries overflow to a per-thread local array that
// is rarely touched. One thread per ray; blocks of 128 threads.
#define HIPRT_B200_DEFAULT_FUNC_TABLE
#include <hiprt/hiprt_device.h>

#include "builders.h"
#include "trace.h"

#include <algorithm>
#include <cstdlib>

namespace hiprt
{
constexpr uint32_t TraceBlock		= 128; // scene kernels
constexpr uint32_t SharedStackDepth = 12;

// Top `SharedDepth` entries in shared memory (entry e of a thread at base[e * 32], one bank per lane), the rest of the
// 64 entries the reference's private stack offers in a per-thread local array.
template <uint32_t SharedDepth>
class HybridStackT
{
  public:
	__device__ HybridStackT( uint32_t* sharedBase ) : m_shared( sharedBase + ( threadIdx.x & 31u ) + ( threadIdx.x >> 5 ) * 32u * SharedDepth ) { reset(); }
	__device__ void push( uint32_t v )
	{
		if ( m_top < SharedDepth )
			m_shared[m_top * 32u] = v;
		else
			m_local[m_top - SharedDepth] = v;
		++m_top;
	}
	__device__ uint32_t pop()
	{
		--m_top;
		return m_top < SharedDepth ? m_shared[m_top * 32u] : m_local[m_top - SharedDepth];
	}
	__device__ bool		empty() const { return m_top == 0; }
	__device__ uint32_t vacancy() const { return SharedDepth + LocalDepth - m_top; }
	__device__ void		reset() { m_top = 0; }

  private:
	static constexpr uint32_t LocalDepth = 64 - SharedDepth;
	uint32_t*				  m_shared;
	uint32_t				  m_local[LocalDepth];
	uint32_t				  m_top;
};
using HybridStack = HybridStackT<SharedStackDepth>;

__device__ __forceinline__ hiprtRay loadRay( const hiprtRay* rays, uint32_t index )
{
	const float4* rp = reinterpret_cast<const float4*>( rays + index );
	const float4  r0 = __ldg( rp ), r1 = __ldg( rp + 1 );
	hiprtRay	  ray;
	ray.origin = make_float3( r0.x, r0.y, r0.z ), ray.minT = r0.w;
	ray.direction = make_float3( r1.x, r1.y, r1.z ), ray.maxT = r1.w;
	return ray;
}
__device__ __forceinline__ void storeHit( hiprtHit* hits, uint32_t index, const hiprtHit& hit )
{
	float4* hp = reinterpret_cast<float4*>( hits + index );
	hp[0] = make_float4( __uint_as_float( hit.instanceIDs[0] ), __uint_as_float( hit.instanceIDs[1] ), __uint_as_float( hit.instanceIDs[2] ), __uint_as_float( hit.instanceIDs[3] ) );
	hp[1] = make_float4( __uint_as_float( hit.primID ), 0.0f, hit.uv.x, hit.uv.y );
	hp[2] = make_float4( hit.normal.x, hit.normal.y, hit.normal.z, hit.t );
}

// One thread per ray. Persistent = the grid only fills the machine and every warp keeps fetching batches of 32
// consecutive rays from a global counter until the buffer is exhausted (no block-granular tail: a block slot is
// never held by one long ray while its other warps sit idle).
template <hiprtTraversalType Type, uint32_t Block, uint32_t SharedDepth, bool Persistent>
__global__ void __launch_bounds__( Block ) traceGeometryKernel(
	hiprtGeometry geom, const hiprtRay* __restrict__ rays, uint32_t numRays, hiprtHit* __restrict__ hits, uint32_t* workCounter )
{
	__shared__ uint32_t sStack[Block * SharedDepth];
	using Stack		= HybridStackT<SharedDepth>;
	const bool mesh = reinterpret_cast<const GeomHeader*>( geom )->geomType & 1u;
	if constexpr ( Persistent )
	{
		const uint32_t lane = threadIdx.x & 31u;
		while ( true )
		{
			uint32_t base = 0;
			if ( lane == 0 ) base = atomicAdd( workCounter, 32u );
			base = __shfl_sync( 0xFFFFFFFFu, base, 0 );
			if ( base >= numRays ) return;
			const uint32_t index = base + lane;
			if ( index < numRays )
			{
				Stack		   stack( sStack );
				const hiprtRay ray = loadRay( rays, index );
				hiprtHit	   hit;
				if ( mesh )
				{
					hiprtGeomTraversalT<Stack, hiprtTriangleNode, Type> tr( geom, ray, stack );
					hit = tr.getNextHit();
				}
				else
				{
					hiprtGeomTraversalT<Stack, hiprtCustomNode, Type> tr( geom, ray, stack );
					hit = tr.getNextHit();
				}
				storeHit( hits, index, hit );
			}
			__syncwarp();
		}
	}
	else
	{
		const uint32_t index = blockIdx.x * Block + threadIdx.x;
		if ( index >= numRays ) return;
		Stack		   stack( sStack );
		const hiprtRay ray = loadRay( rays, index );
		hiprtHit	   hit;
		if ( mesh )
		{
			hiprtGeomTraversalT<Stack, hiprtTriangleNode, Type> tr( geom, ray, stack );
			hit = tr.getNextHit();
		}
		else
		{
			hiprtGeomTraversalT<Stack, hiprtCustomNode, Type> tr( geom, ray, stack );
			hit = tr.getNextHit();
		}
		storeHit( hits, index, hit );
	}
}

template <hiprtTraversalType Type>
__global__ void __launch_bounds__( TraceBlock ) traceSceneKernel(
	hiprtScene scene, const hiprtRay* __restrict__ rays, uint32_t numRays, hiprtHit* __restrict__ hits, hiprtRayMask mask )
{
	__shared__ uint32_t sStack[TraceBlock * SharedStackDepth];
	const uint32_t		index = blockIdx.x * TraceBlock + threadIdx.x;
	if ( index >= numRays ) return;
	HybridStack				stack( sStack );
	hiprtEmptyInstanceStack noInstanceStack;
	const float4*			rp = reinterpret_cast<const float4*>( rays + index );
	const float4			r0 = __ldg( rp ), r1 = __ldg( rp + 1 );
	hiprtRay				ray;
	ray.origin = make_float3( r0.x, r0.y, r0.z ), ray.minT = r0.w;
	ray.direction = make_float3( r1.x, r1.y, r1.z ), ray.maxT = r1.w;
	hiprtSceneTraversalT<HybridStack, hiprtEmptyInstanceStack, Type> tr( scene, ray, stack, noInstanceStack, mask );
	const hiprtHit hit = tr.getNextHit();
	float4*		   hp  = reinterpret_cast<float4*>( hits + index );
	hp[0] = make_float4( __uint_as_float( hit.instanceIDs[0] ), __uint_as_float( hit.instanceIDs[1] ), __uint_as_float( hit.instanceIDs[2] ), __uint_as_float( hit.instanceIDs[3] ) );
	hp[1] = make_float4( __uint_as_float( hit.primID ), 0.0f, hit.uv.x, hit.uv.y );
	hp[2] = make_float4( hit.normal.x, hit.normal.y, hit.normal.z, hit.t );
}

static int traceVariant()
{
	static const int v = [] {
		const char* e = getenv( "HIPRT_B200_TRACE_VARIANT" );
		return e ? atoi( e ) : 1;
	}();
	return v;
}

template <hiprtTraversalType Type, uint32_t Block, uint32_t SharedDepth, bool Persistent>
static void launchGeometry( hiprtGeometry geom, const hiprtRay* rays, uint32_t numRays, hiprtHit* hits, uint32_t* workCounter, uint32_t smCount, cudaStream_t stream )
{
	auto kernel = traceGeometryKernel<Type, Block, SharedDepth, Persistent>;
	if constexpr ( Persistent )
	{
		static int blocksPerSm = 0; // same value for every caller: a benign race
		if ( blocksPerSm == 0 ) cudaOccupancyMaxActiveBlocksPerMultiprocessor( &blocksPerSm, kernel, int( Block ), 0 );
		cudaMemsetAsync( workCounter, 0, sizeof( uint32_t ), stream );
		const uint32_t blocks = std::min( DivideRoundUp( numRays, Block ), smCount * uint32_t( std::max( blocksPerSm, 1 ) ) );
		kernel<<<blocks, Block, 0, stream>>>( geom, rays, numRays, hits, workCounter );
	}
	else
		kernel<<<DivideRoundUp( numRays, Block ), Block, 0, stream>>>( geom, rays, numRays, hits, workCounter );
}

// Tuning variants (HIPRT_B200_TRACE_VARIANT): block size / shared stack depth / persistent warps.
void traceGeometry(
	hiprtGeometry geom, const hiprtRay* rays, uint32_t numRays, hiprtHit* hits, hiprtTraversalType type, uint32_t* workCounter, uint32_t smCount, cudaStream_t stream )
{
	if ( numRays == 0 ) return;
	const bool any = type == hiprtTraversalTerminateAtAnyHit;
#define HIPRT_B200_LAUNCH( B, D, P )                                                                                                       \
	if ( any )                                                                                                                             \
		launchGeometry<hiprtTraversalTerminateAtAnyHit, B, D, P>( geom, rays, numRays, hits, workCounter, smCount, stream );               \
	else                                                                                                                                   \
		launchGeometry<hiprtTraversalTerminateAtClosestHit, B, D, P>( geom, rays, numRays, hits, workCounter, smCount, stream );
	switch ( traceVariant() )
	{
	case 0: HIPRT_B200_LAUNCH( 128, 12, false ) break;
	case 2: HIPRT_B200_LAUNCH( 64, 12, false ) break;
	case 3: HIPRT_B200_LAUNCH( 64, 32, false ) break;
	case 4: HIPRT_B200_LAUNCH( 128, 32, true ) break;
	case 5: HIPRT_B200_LAUNCH( 64, 32, true ) break;
	case 6: HIPRT_B200_LAUNCH( 128, 24, false ) break;
	default: HIPRT_B200_LAUNCH( 128, 32, false ) break;
	}
#undef HIPRT_B200_LAUNCH
}

void traceScene(
	hiprtScene scene, const hiprtRay* rays, uint32_t numRays, hiprtHit* hits, hiprtTraversalType type, hiprtRayMask mask, cudaStream_t stream )
{
	if ( numRays == 0 ) return;
	const uint32_t blocks = DivideRoundUp( numRays, TraceBlock );
	if ( type == hiprtTraversalTerminateAtAnyHit )
		traceSceneKernel<hiprtTraversalTerminateAtAnyHit><<<blocks, TraceBlock, 0, stream>>>( scene, rays, numRays, hits, mask );
	else
		traceSceneKernel<hiprtTraversalTerminateAtClosestHit><<<blocks, TraceBlock, 0, stream>>>( scene, rays, numRays, hits, mask );
}
} // namespace hiprt
