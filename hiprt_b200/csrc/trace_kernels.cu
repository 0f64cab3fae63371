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

namespace hiprt
{
constexpr uint32_t TraceBlock		= 128;
constexpr uint32_t SharedStackDepth = 12;

class HybridStack
{
  public:
	__device__ HybridStack( uint32_t* sharedBase ) : m_shared( sharedBase + ( threadIdx.x & 31u ) + ( threadIdx.x >> 5 ) * 32u * SharedStackDepth ) { reset(); }
	__device__ void push( uint32_t v )
	{
		if ( m_top < SharedStackDepth )
			m_shared[m_top * 32u] = v;
		else
			m_local[m_top - SharedStackDepth] = v;
		++m_top;
	}
	__device__ uint32_t pop()
	{
		--m_top;
		return m_top < SharedStackDepth ? m_shared[m_top * 32u] : m_local[m_top - SharedStackDepth];
	}
	__device__ bool		empty() const { return m_top == 0; }
	__device__ uint32_t vacancy() const { return SharedStackDepth + LocalDepth - m_top; }
	__device__ void		reset() { m_top = 0; }

  private:
	static constexpr uint32_t LocalDepth = 52;
	uint32_t*				  m_shared;
	uint32_t				  m_local[LocalDepth];
	uint32_t				  m_top;
};

template <hiprtTraversalType Type>
__global__ void __launch_bounds__( TraceBlock )
	traceGeometryKernel( hiprtGeometry geom, const hiprtRay* __restrict__ rays, uint32_t numRays, hiprtHit* __restrict__ hits )
{
	__shared__ uint32_t sStack[TraceBlock * SharedStackDepth];
	const uint32_t		index = blockIdx.x * TraceBlock + threadIdx.x;
	if ( index >= numRays ) return;
	HybridStack stack( sStack );
	const float4* rp = reinterpret_cast<const float4*>( rays + index );
	const float4  r0 = __ldg( rp ), r1 = __ldg( rp + 1 );
	hiprtRay	  ray;
	ray.origin = make_float3( r0.x, r0.y, r0.z ), ray.minT = r0.w;
	ray.direction = make_float3( r1.x, r1.y, r1.z ), ray.maxT = r1.w;
	hiprtGeomTraversalT<HybridStack, hiprtTriangleNode, Type> tr( geom, ray, stack );
	const hiprtHit hit = tr.getNextHit();
	float4*		   hp  = reinterpret_cast<float4*>( hits + index );
	hp[0] = make_float4( __uint_as_float( hit.instanceIDs[0] ), __uint_as_float( hit.instanceIDs[1] ), __uint_as_float( hit.instanceIDs[2] ), __uint_as_float( hit.instanceIDs[3] ) );
	hp[1] = make_float4( __uint_as_float( hit.primID ), 0.0f, hit.uv.x, hit.uv.y );
	hp[2] = make_float4( hit.normal.x, hit.normal.y, hit.normal.z, hit.t );
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

void traceGeometry( hiprtGeometry geom, const hiprtRay* rays, uint32_t numRays, hiprtHit* hits, hiprtTraversalType type, cudaStream_t stream )
{
	if ( numRays == 0 ) return;
	const uint32_t blocks = DivideRoundUp( numRays, TraceBlock );
	if ( type == hiprtTraversalTerminateAtAnyHit )
		traceGeometryKernel<hiprtTraversalTerminateAtAnyHit><<<blocks, TraceBlock, 0, stream>>>( geom, rays, numRays, hits );
	else
		traceGeometryKernel<hiprtTraversalTerminateAtClosestHit><<<blocks, TraceBlock, 0, stream>>>( geom, rays, numRays, hits );
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
