// User trace kernels for tests/test_gpu_user_kernels.py, written against the public device API exactly as an
// application of the reference would write them (cf. the reference's test/kernels/HiprtTestKernel.h:
// CustomIntersectionKernel :177-214, CutoutKernel :216-260, TraceKernel :102-116). Compiled at test time
// through hiprtBuildTraceKernels (NVRTC), which prepends <hiprt/hiprt_device.h> and the func-table dispatchers.

// custom primitive: sphere (centre.xyz, radius) per AABB; data = float4 array
__device__ bool intersectSphere( const hiprtRay& ray, const void* data, void* payload, hiprtHit& hit )
{
	const float4 s	= reinterpret_cast<const float4*>( data )[hit.primID];
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
	if ( payload ) atomicAdd( reinterpret_cast<unsigned int*>( payload ), 1u );
	return true;
}

// custom primitive of the reference's scene tests (intersectCircle, HiprtTestKernel.h:62-79): a disc of radius 0.1 around
// (centres[primID], 0.5) in the plane z = 1, for rays that travel along +z; data = one float (the x of the centre) per primitive
__device__ bool intersectCircle( const hiprtRay& ray, const void* data, void*, hiprtHit& hit )
{
	const float dx = reinterpret_cast<const float*>( data )[hit.primID] - ray.origin.x;
	const float dy = 0.5f - ray.origin.y;
	if ( sqrtf( dx * dx + dy * dy ) >= 0.1f ) return false;
	hit.t = ( 1.0f - ray.origin.z ) / ray.direction.z;
	return true;
}

// hit filter ("cutout"): returns true to REJECT the hit (reference convention); data = one float threshold
__device__ bool cutoutFilter( const hiprtRay&, const void* data, void*, const hiprtHit& hit )
{
	const float threshold = *reinterpret_cast<const float*>( data );
	return hit.uv.x + hit.uv.y > threshold;
}

extern "C" __global__ void customKernel(
	hiprtGeometry geom, hiprtFuncTable table, const hiprtRay* rays, hiprtHit* hits, uint32_t numRays, unsigned int* counters )
{
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if ( i >= numRays ) return;
	hiprtGeomCustomTraversalClosest tr( geom, rays[i], hiprtTraversalHintDefault, counters, table );
	hits[i] = tr.getNextHit();
}

extern "C" __global__ void cutoutKernel( hiprtGeometry geom, hiprtFuncTable table, const hiprtRay* rays, hiprtHit* hits, uint32_t numRays )
{
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if ( i >= numRays ) return;
	hiprtGeomTraversalClosest tr( geom, rays[i], hiprtTraversalHintDefault, nullptr, table );
	hits[i] = tr.getNextHit();
}

// every hit along the ray, in traversal order, through the re-entrant any-hit iterator
extern "C" __global__ void allHitsKernel(
	hiprtGeometry geom, const hiprtRay* rays, uint32_t numRays, uint32_t maxPerRay, uint32_t* counts, uint32_t* prims, float* ts )
{
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if ( i >= numRays ) return;
	hiprtGeomTraversalAnyHit tr( geom, rays[i] );
	uint32_t				 n = 0;
	while ( true )
	{
		const hiprtHit h = tr.getNextHit();
		if ( tr.getCurrentState() != hiprtTraversalStateHit ) break;
		if ( n < maxPerRay ) prims[i * maxPerRay + n] = h.primID, ts[i * maxPerRay + n] = h.t;
		++n;
	}
	counts[i] = n;
}

// every hit of a (possibly multi-level) scene through the re-entrant any-hit iterator, custom primitives through the function
// table: per hit the primitive and the instance ids of the first two levels (cf. SceneIntersectionKernel, HiprtTestKernel.h:286-326)
extern "C" __global__ void sceneAllHitsKernel(
	hiprtScene scene, hiprtFuncTable table, const hiprtRay* rays, uint32_t numRays, uint32_t maxPerRay, uint32_t* counts, uint32_t* prims,
	uint32_t* inst0, uint32_t* inst1 )
{
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if ( i >= numRays ) return;
	hiprtSceneTraversalAnyHit tr( scene, rays[i], hiprtFullRayMask, hiprtTraversalHintDefault, nullptr, table );
	uint32_t				  n = 0;
	while ( true )
	{
		const hiprtHit h = tr.getNextHit();
		if ( h.hasHit() )
		{
			if ( n < maxPerRay )
			{
				prims[i * maxPerRay + n] = h.primID;
				inst0[i * maxPerRay + n] = h.instanceIDs[0];
				inst1[i * maxPerRay + n] = h.instanceIDs[1];
			}
			++n;
		}
		if ( tr.getCurrentState() != hiprtTraversalStateHit ) break;
	}
	counts[i] = n;
}

// the reference's TraceKernel layout: global stack with a shared-memory cache (test/hiprtTest.cpp:1510-1530)
#define SHARED_STACK_SIZE 16
#define BLOCK_SIZE 64
extern "C" __global__ void __launch_bounds__( BLOCK_SIZE ) globalStackKernel(
	hiprtGeometry geom, const hiprtRay* rays, hiprtHit* hits, uint32_t numRays, hiprtGlobalStackBuffer globalStackBuffer, int dynamic )
{
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	__shared__ uint32_t	   cache[SHARED_STACK_SIZE * BLOCK_SIZE];
	hiprtSharedStackBuffer sharedStackBuffer{ SHARED_STACK_SIZE, cache };
	if ( i >= numRays ) return;
	if ( dynamic )
	{
		hiprtDynamicStack									 stack( globalStackBuffer, sharedStackBuffer );
		hiprtGeomTraversalClosestCustomStack<hiprtDynamicStack> tr( geom, rays[i], stack );
		hits[i] = tr.getNextHit();
	}
	else
	{
		hiprtGlobalStack									stack( globalStackBuffer, sharedStackBuffer );
		hiprtGeomTraversalClosestCustomStack<hiprtGlobalStack> tr( geom, rays[i], stack );
		hits[i] = tr.getNextHit();
	}
}

extern "C" __global__ void sceneKernel( hiprtScene scene, const hiprtRay* rays, hiprtHit* hits, uint32_t numRays, uint32_t mask )
{
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if ( i >= numRays ) return;
	hiprtSceneTraversalClosest tr( scene, rays[i], mask );
	hits[i] = tr.getNextHit();
}

// closest hit of a scene that holds custom primitives: the function table goes to the traversal object
extern "C" __global__ void sceneTableKernel( hiprtScene scene, hiprtFuncTable table, const hiprtRay* rays, hiprtHit* hits, uint32_t numRays )
{
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if ( i >= numRays ) return;
	hiprtSceneTraversalClosest tr( scene, rays[i], hiprtFullRayMask, hiprtTraversalHintDefault, nullptr, table );
	hits[i] = tr.getNextHit();
}

// motion blur: every ray is traced at the same `time` (an application would carry a time per ray in its own ray payload)
extern "C" __global__ void motionSceneKernel( hiprtScene scene, const hiprtRay* rays, hiprtHit* hits, uint32_t numRays, float time )
{
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if ( i >= numRays ) return;
	hiprtSceneTraversalClosest tr( scene, rays[i], hiprtFullRayMask, hiprtTraversalHintDefault, nullptr, nullptr, 0, time );
	hits[i] = tr.getNextHit();
}

// transform queries of hiprt_device.h for one instance: out[4 * i + {0,1,2,3}] = point object->world, that point back to
// object space, vector object->world, that vector back to object space
extern "C" __global__ void transformQueryKernel( hiprtScene scene, uint32_t instanceID, const float3* in, float3* out, uint32_t count, float time )
{
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if ( i >= count ) return;
	const float3 pw = hiprtPointObjectToWorld( in[i], scene, instanceID, time );
	const float3 vw = hiprtVectorObjectToWorld( in[i], scene, instanceID, time );
	out[4 * i + 0]	= pw;
	out[4 * i + 1]	= hiprtPointWorldToObject( pw, scene, instanceID, time );
	out[4 * i + 2]	= vw;
	out[4 * i + 3]	= hiprtVectorWorldToObject( vw, scene, instanceID, time );
	// the multi-level overloads with a single level must agree
	const uint32_t ids[hiprtMaxInstanceLevels] = { instanceID, hiprtInvalidValue, hiprtInvalidValue, hiprtInvalidValue };
	const float3   pw2 = hiprtPointObjectToWorld( in[i], scene, ids, time );
	if ( pw2.x != pw.x || pw2.y != pw.y || pw2.z != pw.z ) out[4 * i + 0] = make_float3( 1e30f, 1e30f, 1e30f );
}

// ---- Monte-Carlo kernels for the golden-image tests (tests/test_gpu_golden.py). The random numbers are this file's own
// (a PCG output function over an LCG state): the pictures are compared statistically, not sample by sample.
__device__ float uniform01( uint32_t& state )
{
	state	   = state * 747796405u + 2891336453u;
	uint32_t w = ( ( state >> ( ( state >> 28u ) + 4u ) ) ^ state ) * 277803737u;
	w		   = ( w >> 22u ) ^ w;
	return float( w >> 8 ) * ( 1.0f / 16777216.0f );
}

__device__ float3 unit( float3 v )
{
	const float s = rsqrtf( v.x * v.x + v.y * v.y + v.z * v.z );
	return make_float3( v.x * s, v.y * s, v.z * s );
}

// Ambient occlusion as the reference's AoRayKernel renders it (test/kernels/AoRayKernel.h:40-131): a pinhole camera at `eye`
// looking down -z (sensor 0.024 high, y down), `spp` primary closest-hit rays per pixel with one random offset for x and y,
// from every hit that is not on an emissive instance `aoSamples` cosine-weighted any-hit rays of length `aoRadius` starting at
// 99 % of the hit distance; unoccluded[pixel] counts the rays that escape. Instance frames are taken to be the identity
// (hit.normal is used as the world-space normal).
extern "C" __global__ void aoKernel(
	hiprtScene scene, uint32_t width, uint32_t height, float eyeX, float eyeY, float eyeZ, float focal, uint32_t spp, uint32_t aoSamples,
	float aoRadius, uint32_t emissiveInstances, uint32_t* unoccluded )
{
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if ( i >= width * height ) return;
	const uint32_t x = i % width, y = i / width;
	uint32_t	   rng = i * 9781u + 0x68bc21ebu;
	uniform01( rng );
	uint32_t	   count = 0;
	const float	   sensorH = 0.024f, sensorW = 0.024f * ( float( width ) / float( height ) );
	for ( uint32_t p = 0; p < spp; ++p )
	{
		const float offset = uniform01( rng );
		hiprtRay	ray;
		ray.origin	  = make_float3( eyeX, eyeY, eyeZ );
		ray.direction = unit( make_float3(
			( ( float( x ) + offset ) / float( width ) - 0.5f ) * sensorW, -( ( float( y ) + offset ) / float( height ) - 0.5f ) * sensorH, -focal ) );
		hiprtSceneTraversalClosest tr( scene, ray );
		const hiprtHit			   hit = tr.getNextHit();
		if ( !hit.hasHit() || ( ( emissiveInstances >> hit.instanceID ) & 1u ) ) continue;
		float3 n = hit.normal;
		if ( ray.direction.x * n.x + ray.direction.y * n.y + ray.direction.z * n.z > 0.0f ) n = make_float3( -n.x, -n.y, -n.z );
		n = unit( n );
		// orthonormal basis around n
		const float3 a = fabsf( n.x ) > 0.5f ? make_float3( 0.0f, 1.0f, 0.0f ) : make_float3( 1.0f, 0.0f, 0.0f );
		const float3 t = unit( make_float3( a.y * n.z - a.z * n.y, a.z * n.x - a.x * n.z, a.x * n.y - a.y * n.x ) );
		const float3 b = make_float3( n.y * t.z - n.z * t.y, n.z * t.x - n.x * t.z, n.x * t.y - n.y * t.x );
		const float	 s = hit.t * 0.99f;
		hiprtRay	 ao;
		ao.origin = make_float3( ray.origin.x + s * ray.direction.x, ray.origin.y + s * ray.direction.y, ray.origin.z + s * ray.direction.z );
		ao.maxT	  = aoRadius;
		for ( uint32_t k = 0; k < aoSamples; ++k )
		{
			const float phi = 6.28318530718f * uniform01( rng ), r2 = uniform01( rng );
			const float r = sqrtf( r2 ), h = sqrtf( 1.0f - r2 ), cu = r * cosf( phi ), cv = r * sinf( phi );
			ao.direction = unit( make_float3( cu * t.x + cv * b.x + h * n.x, cu * t.y + cv * b.y + h * n.y, cu * t.z + cv * b.z + h * n.z ) );
			hiprtSceneTraversalAnyHit shadow( scene, ao );
			if ( !shadow.getNextHit().hasHit() ) ++count;
		}
	}
	unoccluded[i] = count;
}

// The reference's MotionBlurSlerpKernel (HiprtTestKernel.h:361-405): one jittered ray per pixel from (0, 0, 5) through a
// 2 x 2 * 1.576494 window at z = 0, traced at `samples` random times; hits[pixel] counts the times at which it hits.
extern "C" __global__ void motionSamplesKernel( hiprtScene scene, hiprtFuncTable table, uint32_t width, uint32_t height, uint32_t samples, uint32_t* hits )
{
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if ( i >= width * height ) return;
	uint32_t rng = i * 2654435761u + 12345u;
	uniform01( rng );
	const float dx = 2.0f * ( float( i % width ) + uniform01( rng ) ) / float( width ) - 1.0f;
	const float dy = 2.0f * ( float( i / width ) + uniform01( rng ) ) / float( height ) - 1.0f;
	hiprtRay	ray;
	ray.origin	  = make_float3( 0.0f, 0.0f, 5.0f );
	ray.direction = unit( make_float3( dx * 1.576494f, dy * 1.576494f, -5.0f ) );
	uint32_t count = 0;
	for ( uint32_t s = 0; s < samples; ++s )
	{
		const float				   time = uniform01( rng );
		hiprtSceneTraversalClosest tr( scene, ray, hiprtFullRayMask, hiprtTraversalHintDefault, nullptr, table, 0, time );
		if ( tr.getNextHit().hasHit() ) ++count;
	}
	hits[i] = count;
}
