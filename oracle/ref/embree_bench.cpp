// ORACLE -- TEST / BENCHMARK INFRASTRUCTURE ONLY.
// CPU baseline on the Embree 4.3 binary that ships with the reference (contrib/embree/linux,
// headers contrib/embree/include/embree4), as BASELINE.md section 2b prescribes: triangle scene
// with build quality LOW (Morton) or MEDIUM (SAH) timed around rtcCommitScene; rtcIntersect1 /
// rtcOccluded1 over all host threads with FTZ/DAZ enabled in every worker.
// C interface for ctypes (oracle/embree_binding.py).
#include <embree4/rtcore.h>

#include <pmmintrin.h>
#include <xmmintrin.h>

#include <atomic>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <limits>
#include <thread>
#include <vector>

namespace
{
RTCDevice g_device = nullptr;

struct Ray32
{
	float ox, oy, oz, tmin, dx, dy, dz, tmax;
};
struct Hit48
{
	uint32_t inst[4];
	uint32_t prim, pad;
	float	 u, v, nx, ny, nz, t;
};

template <typename Fn>
void parallelFor( uint32_t n, uint32_t threads, Fn fn )
{
	std::atomic<uint32_t>	 next{ 0 };
	const uint32_t			 chunk = 4096;
	std::vector<std::thread> pool;
	for ( uint32_t t = 0; t < threads; ++t )
		pool.emplace_back( [&] {
			_MM_SET_FLUSH_ZERO_MODE( _MM_FLUSH_ZERO_ON );
			_MM_SET_DENORMALS_ZERO_MODE( _MM_DENORMALS_ZERO_ON );
			while ( true )
			{
				const uint32_t b = next.fetch_add( chunk );
				if ( b >= n ) break;
				const uint32_t e = b + chunk < n ? b + chunk : n;
				for ( uint32_t i = b; i < e; ++i )
					fn( i );
			}
		} );
	for ( auto& th : pool )
		th.join();
}
} // namespace

extern "C" {

int emb_init( int threads )
{
	if ( g_device ) return 0;
	_MM_SET_FLUSH_ZERO_MODE( _MM_FLUSH_ZERO_ON );
	_MM_SET_DENORMALS_ZERO_MODE( _MM_DENORMALS_ZERO_ON );
	char cfg[64];
	snprintf( cfg, sizeof( cfg ), "threads=%d", threads );
	g_device = rtcNewDevice( threads > 0 ? cfg : nullptr );
	return g_device ? 0 : -1;
}

// quality: 0 low (Morton), 1 medium (SAH), 2 high. Returns the scene; *msOut = rtcCommitScene time.
void* emb_build( const float* verts, uint32_t vertCount, const uint32_t* indices, uint32_t triCount, int quality, double* msOut )
{
	RTCScene	scene = rtcNewScene( g_device );
	RTCGeometry geom  = rtcNewGeometry( g_device, RTC_GEOMETRY_TYPE_TRIANGLE );
	float*		v	  = static_cast<float*>( rtcSetNewGeometryBuffer( geom, RTC_BUFFER_TYPE_VERTEX, 0, RTC_FORMAT_FLOAT3, 12, vertCount ) );
	memcpy( v, verts, size_t( vertCount ) * 12 );
	uint32_t* idx = static_cast<uint32_t*>( rtcSetNewGeometryBuffer( geom, RTC_BUFFER_TYPE_INDEX, 0, RTC_FORMAT_UINT3, 12, triCount ) );
	if ( indices )
		memcpy( idx, indices, size_t( triCount ) * 12 );
	else
		for ( uint32_t i = 0; i < triCount * 3; ++i )
			idx[i] = i;
	const RTCBuildQuality q = quality == 0 ? RTC_BUILD_QUALITY_LOW : ( quality == 1 ? RTC_BUILD_QUALITY_MEDIUM : RTC_BUILD_QUALITY_HIGH );
	rtcSetGeometryBuildQuality( geom, q );
	rtcCommitGeometry( geom );
	rtcAttachGeometry( scene, geom );
	rtcReleaseGeometry( geom );
	rtcSetSceneBuildQuality( scene, q );
	const auto t0 = std::chrono::high_resolution_clock::now();
	rtcCommitScene( scene );
	if ( msOut ) *msOut = std::chrono::duration<double, std::milli>( std::chrono::high_resolution_clock::now() - t0 ).count();
	return scene;
}

void emb_release( void* scene ) { rtcReleaseScene( static_cast<RTCScene>( scene ) ); }

// kind 2: closest hit (rtcIntersect1), kind 1: any hit (rtcOccluded1; hit.prim = 0 when occluded).
// rays: hiprtRay[n] (32 B), hits: hiprtHit[n] (48 B). Returns wall milliseconds.
double emb_trace( void* scenePtr, const void* raysIn, uint32_t n, void* hitsOut, int kind, int threads )
{
	RTCScene	 scene = static_cast<RTCScene>( scenePtr );
	const Ray32* rays  = static_cast<const Ray32*>( raysIn );
	Hit48*		 hits  = static_cast<Hit48*>( hitsOut );
	const auto	 t0	   = std::chrono::high_resolution_clock::now();
	parallelFor( n, threads > 0 ? threads : std::thread::hardware_concurrency(), [&]( uint32_t i ) {
		const Ray32& r = rays[i];
		Hit48		 h;
		h.inst[0] = h.inst[1] = h.inst[2] = h.inst[3] = 0xFFFFFFFFu;
		h.prim = 0xFFFFFFFFu, h.pad = 0, h.u = h.v = h.nx = h.ny = h.nz = 0.0f, h.t = -1.0f;
		if ( kind == 2 )
		{
			RTCRayHit rh;
			rh.ray.org_x = r.ox, rh.ray.org_y = r.oy, rh.ray.org_z = r.oz, rh.ray.tnear = r.tmin;
			rh.ray.dir_x = r.dx, rh.ray.dir_y = r.dy, rh.ray.dir_z = r.dz, rh.ray.tfar = r.tmax;
			rh.ray.time = 0.0f, rh.ray.mask = 0xFFFFFFFFu, rh.ray.id = 0, rh.ray.flags = 0;
			rh.hit.geomID = RTC_INVALID_GEOMETRY_ID, rh.hit.primID = RTC_INVALID_GEOMETRY_ID;
			rh.hit.instID[0] = RTC_INVALID_GEOMETRY_ID;
			rtcIntersect1( scene, &rh, nullptr );
			if ( rh.hit.geomID != RTC_INVALID_GEOMETRY_ID )
			{
				h.prim = rh.hit.primID, h.u = rh.hit.u, h.v = rh.hit.v;
				h.nx = rh.hit.Ng_x, h.ny = rh.hit.Ng_y, h.nz = rh.hit.Ng_z, h.t = rh.ray.tfar;
			}
		}
		else
		{
			RTCRay ray;
			ray.org_x = r.ox, ray.org_y = r.oy, ray.org_z = r.oz, ray.tnear = r.tmin;
			ray.dir_x = r.dx, ray.dir_y = r.dy, ray.dir_z = r.dz, ray.tfar = r.tmax;
			ray.time = 0.0f, ray.mask = 0xFFFFFFFFu, ray.id = 0, ray.flags = 0;
			rtcOccluded1( scene, &ray, nullptr );
			if ( ray.tfar < 0.0f ) h.prim = 0, h.t = 0.0f;
		}
		if ( hits ) hits[i] = h;
	} );
	return std::chrono::duration<double, std::milli>( std::chrono::high_resolution_clock::now() - t0 ).count();
}

} // extern "C"
