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

#include <cstdlib>

namespace hiprt
{
constexpr uint32_t TraceBlock		= 128;
constexpr uint32_t SharedStackDepth = 12;
constexpr uint32_t FastStackDepth	= 32; // shared-memory stack entries per ray on the fast path (16 KiB per block)

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

// General path (custom primitives, or a ray whose stack outgrew the fast path's shared memory): the public
// traversal object with a per-thread local stack.
template <hiprtTraversalType Type>
__device__ __noinline__ hiprtHit traceGeometryGeneral( hiprtGeometry geom, const hiprtRay& ray, bool triangles )
{
	hiprtPrivateStack stack;
	if ( triangles )
	{
		hiprtGeomTraversalT<hiprtPrivateStack, hiprtTriangleNode, Type> tr( geom, ray, stack );
		return tr.getNextHit();
	}
	hiprtGeomTraversalT<hiprtPrivateStack, hiprtCustomNode, Type> tr( geom, ray, stack );
	return tr.getNextHit();
}

// Variant: 0 = general traversal object with the hybrid (shared + local) stack, 1 = fast path, 2 = fast path without
// leaf postponement, 3 / 4 = 1 / 2 compiled for 10 resident blocks per SM (<= 51 registers). Selected at run time by HIPRT_B200_TRACE_VARIANT (tuning knob; default 1).
template <hiprtTraversalType Type, int Variant>
__global__ void __launch_bounds__( TraceBlock, Variant >= 3 ? 10 : 1 )
	traceGeometryKernel( hiprtGeometry geom, const hiprtRay* __restrict__ rays, uint32_t numRays, hiprtHit* __restrict__ hits )
{
	__shared__ uint32_t sStack[TraceBlock * ( Variant == 0 ? SharedStackDepth : FastStackDepth )];
	const uint32_t		index = blockIdx.x * TraceBlock + threadIdx.x;
	if ( index >= numRays ) return;
	const float4* rp = reinterpret_cast<const float4*>( rays + index );
	const float4  r0 = __ldg( rp ), r1 = __ldg( rp + 1 );
	hiprtRay	  ray;
	ray.origin = make_float3( r0.x, r0.y, r0.z ), ray.minT = r0.w;
	ray.direction = make_float3( r1.x, r1.y, r1.z ), ray.maxT = r1.w;
	hiprtHit hit;
	if constexpr ( Variant == 0 )
	{
		HybridStack												  stack( sStack );
		hiprtGeomTraversalT<HybridStack, hiprtTriangleNode, Type> tr( geom, ray, stack );
		hit = tr.getNextHit();
	}
	else
	{
		const GeomHeader* header	= reinterpret_cast<const GeomHeader*>( geom );
		const bool		  triangles = header->geomType & 1u;
		if ( !triangles || !traceTrianglesFast<Type == hiprtTraversalTerminateAtAnyHit, TraceBlock, FastStackDepth, Variant == 1 || Variant == 3>(
							   header, ray, sStack + threadIdx.x, hit ) )
			hit = traceGeometryGeneral<Type>( geom, ray, triangles );
	}
	float4* hp = reinterpret_cast<float4*>( hits + index );
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

static int traceVariant()
{
	static const int v = [] {
		const char* e = getenv( "HIPRT_B200_TRACE_VARIANT" );
		return e ? atoi( e ) : 1;
	}();
	return v;
}

void traceGeometry( hiprtGeometry geom, const hiprtRay* rays, uint32_t numRays, hiprtHit* hits, hiprtTraversalType type, cudaStream_t stream )
{
	if ( numRays == 0 ) return;
	const uint32_t blocks = DivideRoundUp( numRays, TraceBlock );
	const bool	   any	  = type == hiprtTraversalTerminateAtAnyHit;
#define HIPRT_B200_LAUNCH( V )                                                                                                          \
	if ( any )                                                                                                                          \
		traceGeometryKernel<hiprtTraversalTerminateAtAnyHit, V><<<blocks, TraceBlock, 0, stream>>>( geom, rays, numRays, hits );        \
	else                                                                                                                                \
		traceGeometryKernel<hiprtTraversalTerminateAtClosestHit, V><<<blocks, TraceBlock, 0, stream>>>( geom, rays, numRays, hits );
	switch ( traceVariant() )
	{
	case 0: HIPRT_B200_LAUNCH( 0 ) break;
	case 2: HIPRT_B200_LAUNCH( 2 ) break;
	case 3: HIPRT_B200_LAUNCH( 3 ) break;
	case 4: HIPRT_B200_LAUNCH( 4 ) break;
	default: HIPRT_B200_LAUNCH( 1 ) break;
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
