// TEST INFRASTRUCTURE: the PRODUCT's device-side traversal (include/hiprt/hiprt_device.h -- the traversal objects a user kernel and
// the library's generic trace kernels instantiate) built for the HOST, so that tests/test_product_traversal_host.py can run it on the
// CPU over node arrays in host memory (the oracle's, whose layouts are the product's) without a GPU. Nothing of the header is
// changed: device qualifiers are defined away, the CUDA vector types come from the toolkit's host-usable headers, the few intrinsics
// become their plain C++ meaning (this file is compiled with -ffp-contract=off, so only the header's explicit fmaf calls fuse).
#include <vector_functions.h>
#include <vector_types.h>

#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#define __CUDACC__ 1
#define __device__
#define __host__
#define __forceinline__ inline
static inline float	   __fmul_rn( float a, float b ) { return a * b; }
static inline float	   __fadd_rn( float a, float b ) { return a + b; }
static inline float	   __fsub_rn( float a, float b ) { return a - b; }
static inline float	   __fdiv_rn( float a, float b ) { return a / b; }
static inline uint32_t __float_as_uint( float f )
{
	uint32_t u;
	memcpy( &u, &f, 4 );
	return u;
}
static inline float __uint_as_float( uint32_t u )
{
	float f;
	memcpy( &f, &u, 4 );
	return f;
}
template <typename T>
static inline T __ldg( const T* p )
{
	return *p;
}
// one host thread plays one lane of a one-lane warp (only the dynamic global stack uses these)
static inline unsigned __activemask() { return 1u; }
static inline int	   __ffs( int v ) { return __builtin_ffs( v ); }
static inline void	   __threadfence() {}
template <typename T>
static inline T __shfl_sync( unsigned, T v, int )
{
	return v;
}
template <typename T>
static inline T atomicCAS( T* p, T compare, T value )
{
	const T old = *p;
	if ( old == compare ) *p = value;
	return old;
}
template <typename T>
static inline T atomicExch( T* p, T value )
{
	const T old = *p;
	*p			= value;
	return old;
}
struct HostDim3
{
	unsigned x = 1, y = 1, z = 1;
};
static HostDim3 threadIdx{ 0, 0, 0 }, blockIdx{ 0, 0, 0 }, blockDim, gridDim;
#define HIPRT_B200_DEFAULT_FUNC_TABLE
#include <hiprt/hiprt_device.h>

using namespace hiprt;

static GeomHeader geometryHeader( const void* boxNodes, const void* leaves, uint32_t boxCount, uint32_t leafCount )
{
	GeomHeader h{};
	h.boxNodes		= static_cast<Box4Node*>( const_cast<void*>( boxNodes ) );
	h.primNodes		= const_cast<void*>( leaves );
	h.boxNodeCount	= boxCount;
	h.primNodeCount = leafCount;
	h.geomType		= ( InvalidValue << 1 ) | 1u; // a triangle mesh without a user geometry type
	return h;
}

extern "C" {
// kind 2: closest hit, 1: any hit (hiprtTraversalType)
void prod_trace_geometry( const void* boxNodes, const void* leaves, uint32_t boxCount, uint32_t leafCount, const hiprtRay* rays, uint32_t n, hiprtHit* hits, int kind )
{
	GeomHeader h = geometryHeader( boxNodes, leaves, boxCount, leafCount );
	for ( uint32_t i = 0; i < n; ++i )
	{
		if ( kind == hiprtTraversalTerminateAtClosestHit )
		{
			hiprtGeomTraversalClosest tr( reinterpret_cast<hiprtGeometry>( &h ), rays[i] );
			hits[i] = tr.getNextHit();
		}
		else
		{
			hiprtGeomTraversalAnyHit tr( reinterpret_cast<hiprtGeometry>( &h ), rays[i] );
			hits[i] = tr.getNextHit();
		}
	}
}

// every hit along the ray through the re-entrant any-hit iterator; returns the total number of hits
uint64_t prod_trace_geometry_all(
	const void* boxNodes, const void* leaves, uint32_t boxCount, uint32_t leafCount, const hiprtRay* rays, uint32_t n, uint32_t maxPerRay, uint32_t* counts,
	uint32_t* prims )
{
	GeomHeader h	 = geometryHeader( boxNodes, leaves, boxCount, leafCount );
	uint64_t   total = 0;
	for ( uint32_t i = 0; i < n; ++i )
	{
		hiprtGeomTraversalAnyHit tr( reinterpret_cast<hiprtGeometry>( &h ), rays[i] );
		uint32_t				 c = 0;
		while ( true )
		{
			const hiprtHit hit = tr.getNextHit();
			if ( tr.getCurrentState() != hiprtTraversalStateHit ) break;
			if ( c < maxPerRay ) prims[size_t( i ) * maxPerRay + c] = hit.primID;
			++c;
		}
		counts[i] = c;
		total += c;
	}
	return total;
}

// A single-level scene in host memory: sceneBoxNodes + instanceNodes (64-byte instance leaves whose geometry pointer is patched here:
// leaf l refers to geometry geometryOfLeaf[l] of the `geomCount` (boxNodes, leaves) pairs)
void prod_trace_scene(
	const void* sceneBoxNodes, const void* instanceNodes, uint32_t leafCount, const uint32_t* geometryOfLeaf, const void* const* geomBoxNodes,
	const void* const* geomLeaves, uint32_t geomCount, const hiprtRay* rays, uint32_t n, hiprtHit* hits, int kind, uint32_t rayMask )
{
	std::vector<GeomHeader> geoms( geomCount );
	for ( uint32_t g = 0; g < geomCount; ++g )
		geoms[g] = geometryHeader( geomBoxNodes[g], geomLeaves[g], 0, 0 );
	std::vector<InstanceNode> leaves( leafCount );
	memcpy( leaves.data(), instanceNodes, sizeof( InstanceNode ) * leafCount );
	for ( uint32_t l = 0; l < leafCount; ++l )
		leaves[l].geometry = &geoms[geometryOfLeaf[l]];
	SceneHeader s{};
	s.boxNodes		= static_cast<Box4Node*>( const_cast<void*>( sceneBoxNodes ) );
	s.primNodes		= leaves.data();
	s.primCount		= leafCount;
	s.primNodeCount = leafCount;
	for ( uint32_t i = 0; i < n; ++i )
	{
		if ( kind == hiprtTraversalTerminateAtClosestHit )
		{
			hiprtSceneTraversalClosest tr( reinterpret_cast<hiprtScene>( &s ), rays[i], rayMask );
			hits[i] = tr.getNextHit();
		}
		else
		{
			hiprtSceneTraversalAnyHit tr( reinterpret_cast<hiprtScene>( &s ), rays[i], rayMask );
			hits[i] = tr.getNextHit();
		}
	}
}
}
