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
#include <algorithm>
#include <vector>

#define __CUDACC__ 1
#define __device__
#define __host__
#define __forceinline__ inline
static inline float	   __fmul_rn( float a, float b ) { return a * b; }
static inline float	   __fadd_rn( float a, float b ) { return a + b; }
static inline float	   __fsub_rn( float a, float b ) { return a - b; }
static inline float	   __fdiv_rn( float a, float b ) { return a / b; }
static inline int __float_as_int( float f )
{
	int i;
	memcpy( &i, &f, 4 );
	return i;
}
static inline float __int_as_float( int i )
{
	float f;
	memcpy( &f, &i, 4 );
	return f;
}
// (single host thread: the float atomics of hiprt_math.h only need to compile)
static inline int atomicMin( int* p, int v )
{
	const int old = *p;
	*p			  = v < old ? v : old;
	return old;
}
static inline int atomicMax( int* p, int v )
{
	const int old = *p;
	*p			  = v > old ? v : old;
	return old;
}
static inline unsigned int atomicMin( unsigned int* p, unsigned int v )
{
	const unsigned int old = *p;
	*p					   = v < old ? v : old;
	return old;
}
static inline unsigned int atomicMax( unsigned int* p, unsigned int v )
{
	const unsigned int old = *p;
	*p					   = v > old ? v : old;
	return old;
}
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
#include <hiprt/hiprt_device.h>

using namespace hiprt;

// The two dispatchers every translation unit that includes the device header provides (hiprtBuildTraceKernels generates them from the
// function-name table): here they route to the host copies of the test kernels' custom intersection functions.
typedef bool ( *HostIntersector )( const hiprtRay& ray, const void* data, void* payload, hiprtHit& hit );

// tests/kernels/user_kernels.cu intersectCircle: a disc of radius 0.1 around (centres[primID], 0.5) in the plane z = 1
static bool hostIntersectCircle( const hiprtRay& ray, const void* data, void*, hiprtHit& hit )
{
	const float dx = static_cast<const float*>( data )[hit.primID] - ray.origin.x;
	const float dy = 0.5f - ray.origin.y;
	if ( sqrtf( dx * dx + dy * dy ) >= 0.1f ) return false;
	hit.t = ( 1.0f - ray.origin.z ) / ray.direction.z;
	return true;
}

// tests/kernels/user_kernels.cu intersectSphere: data = float4 (centre, radius) per primitive
static bool hostIntersectSphere( const hiprtRay& ray, const void* data, void*, hiprtHit& hit )
{
	const float4 s	= static_cast<const float4*>( data )[hit.primID];
	const float3 oc = make_float3( ray.origin.x - s.x, ray.origin.y - s.y, ray.origin.z - s.z );
	const float3 d	= ray.direction;
	const float	 a	= d.x * d.x + d.y * d.y + d.z * d.z;
	const float	 b	= 2.0f * ( oc.x * d.x + oc.y * d.y + oc.z * d.z );
	const float	 c	= oc.x * oc.x + oc.y * oc.y + oc.z * oc.z - s.w * s.w;
	const float	 disc = b * b - 4.0f * a * c;
	if ( disc < 0.0f ) return false;
	const float t = ( -b - sqrtf( disc ) ) / ( 2.0f * a );
	if ( t < ray.minT || t > ray.maxT ) return false;
	hit.t	   = t;
	hit.normal = make_float3( oc.x + t * d.x, oc.y + t * d.y, oc.z + t * d.z );
	hit.uv	   = make_float2( 0.0f, 0.0f );
	return true;
}

static HostIntersector g_intersectors[8] = {};

bool intersectFunc( uint32_t geomType, uint32_t rayType, const hiprtFuncTableHeader& tableHeader, const hiprtRay& ray, void* payload, hiprtHit& hit )
{
	const uint32_t index = tableHeader.numGeomTypes * rayType + geomType;
	if ( index >= 8 || !g_intersectors[index] ) return false;
	return g_intersectors[index]( ray, tableHeader.funcDataSets ? tableHeader.funcDataSets[index].intersectFuncData : nullptr, payload, hit );
}

// the reference's Cutout test filter (cutoutFilter, HiprtTestKernel.h:88-100) for the unit quad of two triangles: texture
// coordinates interpolated from hit.uv per primitive, a 16 x 16 checkerboard of holes; true REJECTS the hit
static bool hostCutoutFilter( const hiprtRay&, const void*, void*, const hiprtHit& hit )
{
	const float u = hit.uv.x, v = hit.uv.y;
	const float tx = hit.primID == 0 ? v : u + v;
	const float ty = hit.primID == 0 ? u + v : u;
	return ( ( uint32_t( 16.0f * tx ) + uint32_t( 16.0f * ty ) ) & 1u ) != 0u;
}

typedef bool ( *HostFilter )( const hiprtRay& ray, const void* data, void* payload, const hiprtHit& hit );
static HostFilter g_filters[8] = {};

bool filterFunc( uint32_t geomType, uint32_t rayType, const hiprtFuncTableHeader& tableHeader, const hiprtRay& ray, void* payload, const hiprtHit& hit )
{
	const uint32_t index = tableHeader.numGeomTypes * rayType + geomType;
	if ( index >= 8 || !g_filters[index] ) return false;
	return g_filters[index]( ray, tableHeader.funcDataSets ? tableHeader.funcDataSets[index].filterFuncData : nullptr, payload, hit );
}

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
// hit filter of a function-table slot: 0 none, 1 the Cutout checkerboard
void prod_set_filter( uint32_t index, uint32_t kind )
{
	if ( index < 8 ) g_filters[index] = kind == 1 ? hostCutoutFilter : nullptr;
}

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

// the same closest-hit query with the traversal stack in "shared" + "global" memory (hiprtGlobalStack / hiprtDynamicStack): a ring of
// sharedSize entries that spills its oldest entry to the global region. One host thread is lane 0 of warp 0 of block 0.
void prod_trace_geometry_global_stack(
	const void* boxNodes, const void* leaves, uint32_t boxCount, uint32_t leafCount, const hiprtRay* rays, uint32_t n, hiprtHit* hits, uint32_t sharedSize,
	uint32_t globalSize, int dynamic, uint32_t* maxStateOut )
{
	GeomHeader			  h = geometryHeader( boxNodes, leaves, boxCount, leafCount );
	std::vector<uint32_t> shared( 32u * ( sharedSize + 1 ) ), global( 64u + 32u * 32u * ( globalSize + 1 ) );
	hiprtSharedStackBuffer sharedBuffer{ sharedSize, shared.data() };
	hiprtGlobalStackBuffer globalBuffer{ globalSize, 32u * 32u, global.data() }; // 32 warp slots: the dynamic variant locks one of them
	uint32_t			   worst = 0;
	for ( uint32_t i = 0; i < n; ++i )
	{
		if ( dynamic )
		{
			hiprtDynamicStack									  stack( globalBuffer, sharedBuffer );
			hiprtGeomTraversalClosestCustomStack<hiprtDynamicStack> tr( reinterpret_cast<hiprtGeometry>( &h ), rays[i], stack );
			hits[i] = tr.getNextHit();
			worst	= std::max<uint32_t>( worst, tr.getCurrentState() );
		}
		else
		{
			hiprtGlobalStack									 stack( globalBuffer, sharedBuffer );
			hiprtGeomTraversalClosestCustomStack<hiprtGlobalStack> tr( reinterpret_cast<hiprtGeometry>( &h ), rays[i], stack );
			hits[i] = tr.getNextHit();
			worst	= std::max<uint32_t>( worst, tr.getCurrentState() );
		}
	}
	if ( maxStateOut ) *maxStateOut = worst;
	for ( uint32_t k = 0; k < 32; ++k ) // every lock of the dynamic variant must have been released
		if ( dynamic && global[k] != 0 && maxStateOut ) *maxStateOut = 0xFFFFFFFFu;
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

// ---- general form: custom primitives through the function table, scenes instanced in scenes, animated instances ----------------
// objects: children before parents. kind 0: triangle geometry, 1: custom-primitive geometry (leaves = CustomNode), 2: scene (leaves =
// 64-byte instance nodes; childOfLeaf[l] = index of the object leaf l instances; frames = the scene's 64-byte key frames or null).
struct HostObject
{
	uint32_t	kind, geomType;
	const void* boxNodes;
	const void* leaves;
	uint32_t	leafCount, pad;
	const uint32_t* childOfLeaf;
	const void* frames;
};

// intersectorOfType[g]: 0 none, 1 circle, 2 sphere; dataOfType[g]: its data. maxPerRay == 0: closest hit (one record per ray);
// otherwise every hit through the any-hit scene iterator (counts[i] hits, the first maxPerRay stored).
void prod_trace_general(
	const HostObject* objects, uint32_t objectCount, uint32_t root, uint32_t numGeomTypes, const uint32_t* intersectorOfType, const void* const* dataOfType,
	const hiprtRay* rays, uint32_t n, float time, uint32_t maxPerRay, uint32_t* counts, hiprtHit* hits )
{
	std::vector<GeomHeader>				   geoms( objectCount );
	std::vector<SceneHeader>			   scenes( objectCount );
	std::vector<std::vector<InstanceNode>> nodes( objectCount );
	for ( uint32_t k = 0; k < objectCount; ++k )
	{
		const HostObject& o = objects[k];
		if ( o.kind != 2 )
		{
			geoms[k]		  = geometryHeader( o.boxNodes, o.leaves, 0, o.leafCount );
			geoms[k].geomType = ( o.geomType << 1 ) | ( o.kind == 0 ? 1u : 0u );
			continue;
		}
		nodes[k].resize( o.leafCount );
		memcpy( nodes[k].data(), o.leaves, sizeof( InstanceNode ) * o.leafCount );
		for ( uint32_t l = 0; l < o.leafCount; ++l )
		{
			const uint32_t child = o.childOfLeaf[l];
			if ( objects[child].kind == 2 )
				nodes[k][l].scene = &scenes[child];
			else
				nodes[k][l].geometry = &geoms[child];
		}
		scenes[k]				= SceneHeader{};
		scenes[k].boxNodes		= static_cast<Box4Node*>( const_cast<void*>( o.boxNodes ) );
		scenes[k].primNodes		= nodes[k].data();
		scenes[k].frames		= static_cast<SRTFrame*>( const_cast<void*>( o.frames ) );
		scenes[k].primCount		= o.leafCount;
		scenes[k].primNodeCount = o.leafCount;
	}
	std::vector<hiprtFuncDataSet> sets( numGeomTypes );
	for ( uint32_t g = 0; g < numGeomTypes; ++g )
	{
		sets[g].intersectFuncData = dataOfType[g];
		g_intersectors[g]		  = intersectorOfType[g] == 1 ? hostIntersectCircle : intersectorOfType[g] == 2 ? hostIntersectSphere : nullptr;
	}
	hiprtFuncTableHeader table{ numGeomTypes, 1, sets.data() };
	const hiprtFuncTable tablePtr = reinterpret_cast<hiprtFuncTable>( &table );
	const bool			 isScene  = objects[root].kind == 2;
	for ( uint32_t i = 0; i < n; ++i )
	{
		if ( maxPerRay == 0 )
		{
			if ( isScene )
			{
				hiprtSceneTraversalClosest tr( reinterpret_cast<hiprtScene>( &scenes[root] ), rays[i], hiprtFullRayMask, hiprtTraversalHintDefault, nullptr, tablePtr, 0, time );
				hits[i] = tr.getNextHit();
			}
			else if ( objects[root].kind == 1 )
			{
				hiprtGeomCustomTraversalClosest tr( reinterpret_cast<hiprtGeometry>( &geoms[root] ), rays[i], hiprtTraversalHintDefault, nullptr, tablePtr );
				hits[i] = tr.getNextHit();
			}
			else
			{
				hiprtGeomTraversalClosest tr( reinterpret_cast<hiprtGeometry>( &geoms[root] ), rays[i], hiprtTraversalHintDefault, nullptr, tablePtr );
				hits[i] = tr.getNextHit();
			}
			continue;
		}
		hiprtSceneTraversalAnyHit tr( reinterpret_cast<hiprtScene>( &scenes[root] ), rays[i], hiprtFullRayMask, hiprtTraversalHintDefault, nullptr, tablePtr, 0, time );
		uint32_t				  c = 0;
		while ( true )
		{
			const hiprtHit hit = tr.getNextHit();
			if ( hit.hasHit() )
			{
				if ( c < maxPerRay ) hits[size_t( i ) * maxPerRay + c] = hit;
				++c;
			}
			if ( tr.getCurrentState() != hiprtTraversalStateHit ) break;
		}
		counts[i] = c;
	}
}

// The transform query functions of the device API (hiprtPoint/VectorObjectToWorld/WorldToObject) for instance 0 of a scene whose only
// instance has the given 64-byte key frames: per input vector the four results, in the order object->world point, world->object point,
// object->world vector, world->object vector. twoLevels != 0: the instance sits, under an identity frame, inside an outer scene and the
// multi-level overloads are used.
void prod_transform_queries( const void* keyFrames, uint32_t frameCount, float time, uint32_t n, const float* in, float* out, int twoLevels )
{
	InstanceRecord record{};
	record.frameIndex = 0;
	record.frameCount = frameCount;
	record.type		  = hiprtInstanceTypeGeometry;
	SceneHeader inner{};
	inner.instances	 = &record;
	inner.frames	 = static_cast<SRTFrame*>( const_cast<void*>( keyFrames ) );
	inner.frameCount = frameCount;
	inner.primCount	 = 1;
	SRTFrame identity{};
	identity.rotation = { 0.0f, 0.0f, 0.0f, 1.0f };
	identity.scale	  = { 1.0f, 1.0f, 1.0f };
	InstanceRecord outerRecord{};
	outerRecord.frameIndex = 0;
	outerRecord.frameCount = 1;
	outerRecord.type	   = hiprtInstanceTypeScene;
	outerRecord.scene	   = &inner;
	SceneHeader outer{};
	outer.instances	 = &outerRecord;
	outer.frames	 = &identity;
	outer.frameCount = 1;
	outer.primCount	 = 1;
	const uint32_t ids[hiprtMaxInstanceLevels] = { 0, 0, hiprtInvalidValue, hiprtInvalidValue };
	for ( uint32_t i = 0; i < n; ++i )
	{
		const float3 p = make_float3( in[3 * i], in[3 * i + 1], in[3 * i + 2] );
		float3		 r[4];
		if ( twoLevels )
		{
			const hiprtScene s = reinterpret_cast<hiprtScene>( &outer );
			r[0] = hiprtPointObjectToWorld( p, s, ids, time ), r[1] = hiprtPointWorldToObject( p, s, ids, time );
			r[2] = hiprtVectorObjectToWorld( p, s, ids, time ), r[3] = hiprtVectorWorldToObject( p, s, ids, time );
		}
		else
		{
			const hiprtScene s = reinterpret_cast<hiprtScene>( &inner );
			r[0] = hiprtPointObjectToWorld( p, s, 0u, time ), r[1] = hiprtPointWorldToObject( p, s, 0u, time );
			r[2] = hiprtVectorObjectToWorld( p, s, 0u, time ), r[3] = hiprtVectorWorldToObject( p, s, 0u, time );
		}
		memcpy( out + 12 * i, r, sizeof( r ) );
	}
}
}
