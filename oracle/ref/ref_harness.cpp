// ORACLE -- TEST / BENCHMARK INFRASTRUCTURE ONLY.
//
// Driver for the UNMODIFIED reference (GPUOpen HIPRT v3.1.0): host code calls the reference's
// own C API (libhiprt_ref.so built by oracle/ref/Makefile from /root/reference), the kernels
// in ref_trace.cu (loaded from ref_trace.fatbin next to this library) instantiate the reference's
// own traversal objects. Used for (a) pinning parity: hits and node arrays produced by
// the reference on the same GPU and inputs, (b) the "HIPRT-on-Orochi" timing baseline.
// Exposed as a small C interface for ctypes (oracle/ref_binding.py).
#include <cuda_runtime.h>

#include <hiprt/hiprt.h>

#include <chrono>
#include <cstdio>
#include <cstring>
#include <dlfcn.h>
#include <string>
#include <vector>

namespace
{
hiprtContext  g_ctx = nullptr;
char		  g_err[512];
cudaLibrary_t g_lib = nullptr;
cudaKernel_t  g_kernels[4]; // geomClosest, geomAnyHit, sceneClosest, sceneAnyHit
constexpr int RefBlockSize = 64;

#define REF_CHECK( expr )                                                               \
	do                                                                                  \
	{                                                                                   \
		hiprtError e_ = ( expr );                                                       \
		if ( e_ != hiprtSuccess )                                                       \
		{                                                                               \
			snprintf( g_err, sizeof( g_err ), "%s -> hiprtError %d", #expr, int( e_ ) ); \
			return -1;                                                                  \
		}                                                                               \
	} while ( 0 )
#define CUDA_CHECK( expr )                                                              \
	do                                                                                  \
	{                                                                                   \
		cudaError_t e_ = ( expr );                                                      \
		if ( e_ != cudaSuccess )                                                        \
		{                                                                               \
			snprintf( g_err, sizeof( g_err ), "%s -> %s", #expr, cudaGetErrorString( e_ ) ); \
			return -2;                                                                  \
		}                                                                               \
	} while ( 0 )

struct RefGeometry
{
	hiprtGeometry geom = nullptr;
	void *		  dVerts = nullptr, *dIdx = nullptr;
};

struct RefScene
{
	hiprtScene scene = nullptr;
	void *	   dInstances = nullptr, *dFrames = nullptr;
};
} // namespace

extern "C" {

const char* ref_last_error() { return g_err; }

int ref_init( int device )
{
	if ( g_ctx ) return 0;
	CUDA_CHECK( cudaSetDevice( device ) );
	CUDA_CHECK( cudaFree( nullptr ) );
	// the reference wants the CUcontext: fetch the primary context the runtime just made current
	void* cu = dlopen( "libcuda.so.1", RTLD_NOW );
	if ( !cu )
	{
		snprintf( g_err, sizeof( g_err ), "libcuda.so.1 not found" );
		return -3;
	}
	auto  getCurrent = reinterpret_cast<int ( * )( void** )>( dlsym( cu, "cuCtxGetCurrent" ) );
	void* cuCtx		 = nullptr;
	getCurrent( &cuCtx );
	hiprtContextCreationInput input;
	input.ctxt		 = cuCtx;
	input.device	 = device;
	input.deviceType = hiprtDeviceNVIDIA;
	REF_CHECK( hiprtCreateContext( HIPRT_API_VERSION, input, g_ctx ) );
	hiprtSetLogLevel( g_ctx, hiprtLogLevelError | hiprtLogLevelWarn );
	// trace kernels: ref_trace.fatbin sits next to this shared library
	Dl_info info;
	dladdr( reinterpret_cast<const void*>( &ref_init ), &info );
	std::string path( info.dli_fname );
	path = path.substr( 0, path.find_last_of( '/' ) + 1 ) + "ref_trace.fatbin";
	CUDA_CHECK( cudaLibraryLoadFromFile( &g_lib, path.c_str(), nullptr, nullptr, 0, nullptr, nullptr, 0 ) );
	const char* names[4] = { "refTraceGeomClosest", "refTraceGeomAnyHit", "refTraceSceneClosest", "refTraceSceneAnyHit" };
	for ( int i = 0; i < 4; ++i )
		CUDA_CHECK( cudaLibraryGetKernel( &g_kernels[i], g_lib, names[i] ) );
	return 0;
}

// Build a triangle geometry `reps` times (first build = warm-up: module loading); msOut = best time.
int ref_build_geometry(
	const float* verts, uint32_t vertCount, const uint32_t* indices, uint32_t triCount, uint32_t buildFlags, int reps, void** handleOut, double* msOut )
{
	RefGeometry* g = new RefGeometry;
	CUDA_CHECK( cudaMalloc( &g->dVerts, size_t( vertCount ) * 12 ) );
	CUDA_CHECK( cudaMemcpy( g->dVerts, verts, size_t( vertCount ) * 12, cudaMemcpyHostToDevice ) );
	hiprtTriangleMeshPrimitive mesh{};
	mesh.vertices	  = g->dVerts;
	mesh.vertexCount  = vertCount;
	mesh.vertexStride = 12;
	// non-indexed meshes: the reference sizes its buffers from triangleCount (BvhCommon.h:89-99) although its
	// TriangleMesh falls back to vertexCount / 3 (TriangleMesh.h:66) -- so the count must always be given
	mesh.triangleCount = vertCount / 3;
	if ( indices )
	{
		CUDA_CHECK( cudaMalloc( &g->dIdx, size_t( triCount ) * 12 ) );
		CUDA_CHECK( cudaMemcpy( g->dIdx, indices, size_t( triCount ) * 12, cudaMemcpyHostToDevice ) );
		mesh.triangleIndices = g->dIdx;
		mesh.triangleCount	 = triCount;
		mesh.triangleStride	 = 12;
	}
	hiprtGeometryBuildInput input{};
	input.type					 = hiprtPrimitiveTypeTriangleMesh;
	input.primitive.triangleMesh = mesh;
	input.geomType				 = hiprtInvalidValue;
	hiprtBuildOptions options{};
	options.buildFlags = buildFlags;
	size_t tempSize	   = 0;
	REF_CHECK( hiprtGetGeometryBuildTemporaryBufferSize( g_ctx, input, options, tempSize ) );
	void* temp = nullptr;
	CUDA_CHECK( cudaMalloc( &temp, tempSize ) );
	REF_CHECK( hiprtCreateGeometry( g_ctx, input, options, g->geom ) );
	double best = 1e30;
	for ( int r = 0; r < reps; ++r )
	{
		CUDA_CHECK( cudaDeviceSynchronize() );
		const auto t0 = std::chrono::high_resolution_clock::now();
		REF_CHECK( hiprtBuildGeometry( g_ctx, hiprtBuildOperationBuild, input, options, temp, 0, g->geom ) );
		CUDA_CHECK( cudaDeviceSynchronize() );
		const double ms = std::chrono::duration<double, std::milli>( std::chrono::high_resolution_clock::now() - t0 ).count();
		if ( r > 0 || reps == 1 ) best = ms < best ? ms : best;
	}
	CUDA_CHECK( cudaFree( temp ) );
	*handleOut = g;
	if ( msOut ) *msOut = best;
	return 0;
}

// header: 48 bytes (hiprt/impl/Header.h GeomHeader)
int ref_geometry_header( void* handle, void* headerOut )
{
	RefGeometry* g = static_cast<RefGeometry*>( handle );
	CUDA_CHECK( cudaMemcpy( headerOut, g->geom, 48, cudaMemcpyDeviceToHost ) );
	return 0;
}

int ref_geometry_nodes( void* handle, void* boxNodesOut, uint32_t boxCount, void* primNodesOut, uint32_t primCount )
{
	RefGeometry* g = static_cast<RefGeometry*>( handle );
	void*		 h[2]; // GeomHeader starts with {boxNodes, primNodes} (hiprt/impl/Header.h:33-40)
	CUDA_CHECK( cudaMemcpy( h, g->geom, sizeof( h ), cudaMemcpyDeviceToHost ) );
	CUDA_CHECK( cudaMemcpy( boxNodesOut, h[0], size_t( boxCount ) * 128, cudaMemcpyDeviceToHost ) );
	CUDA_CHECK( cudaMemcpy( primNodesOut, h[1], size_t( primCount ) * 64, cudaMemcpyDeviceToHost ) );
	return 0;
}

void* ref_geometry_device_handle( void* handle ) { return static_cast<RefGeometry*>( handle )->geom; }

int ref_destroy_geometry( void* handle )
{
	RefGeometry* g = static_cast<RefGeometry*>( handle );
	hiprtDestroyGeometry( g_ctx, g->geom );
	cudaFree( g->dVerts );
	cudaFree( g->dIdx );
	delete g;
	return 0;
}

// Trace host rays through a geometry/scene handle that lives on the device (the reference's own or,
// since the formats are byte-compatible, one built by libhiprt_b200). kind: 1 any hit, 2 closest.
static int traceImpl( void* deviceHandle, bool isScene, const hiprtRay* rays, uint32_t numRays, hiprtHit* hits, int kind, int reps, double* msOut )
{
	hiprtRay* dRays = nullptr;
	hiprtHit* dHits = nullptr;
	CUDA_CHECK( cudaMalloc( reinterpret_cast<void**>( &dRays ), size_t( numRays ) * sizeof( hiprtRay ) ) );
	CUDA_CHECK( cudaMalloc( reinterpret_cast<void**>( &dHits ), size_t( numRays ) * sizeof( hiprtHit ) ) );
	CUDA_CHECK( cudaMemcpy( dRays, rays, size_t( numRays ) * sizeof( hiprtRay ), cudaMemcpyHostToDevice ) );
	const uint32_t				blocks = ( numRays + RefBlockSize - 1 ) / RefBlockSize;
	hiprtGlobalStackBufferInput stackInput{};
	stackInput.type		   = hiprtStackTypeGlobal;
	stackInput.entryType   = hiprtStackEntryTypeInteger;
	stackInput.stackSize   = 64;
	stackInput.threadCount = blocks * RefBlockSize;
	hiprtGlobalStackBuffer stackBuffer{};
	REF_CHECK( hiprtCreateGlobalStackBuffer( g_ctx, stackInput, stackBuffer ) );
	cudaEvent_t e0, e1;
	cudaEventCreate( &e0 );
	cudaEventCreate( &e1 );
	float best = 1e30f;
	for ( int r = 0; r < reps; ++r )
	{
		cudaEventRecord( e0 );
		uint32_t n		= numRays;
		void*	 args[] = { &deviceHandle, &n, &stackBuffer, &dRays, &dHits };
		CUDA_CHECK( cudaLaunchKernel( reinterpret_cast<const void*>( g_kernels[( isScene ? 2 : 0 ) + ( kind == 1 ? 1 : 0 )] ), dim3( blocks ), dim3( RefBlockSize ), args, 0, nullptr ) );
		cudaEventRecord( e1 );
		CUDA_CHECK( cudaEventSynchronize( e1 ) );
		float ms = 0;
		cudaEventElapsedTime( &ms, e0, e1 );
		if ( r > 0 || reps == 1 ) best = ms < best ? ms : best;
	}
	CUDA_CHECK( cudaGetLastError() );
	CUDA_CHECK( cudaMemcpy( hits, dHits, size_t( numRays ) * sizeof( hiprtHit ), cudaMemcpyDeviceToHost ) );
	hiprtDestroyGlobalStackBuffer( g_ctx, stackBuffer );
	cudaFree( dRays );
	cudaFree( dHits );
	cudaEventDestroy( e0 );
	cudaEventDestroy( e1 );
	if ( msOut ) *msOut = best;
	return 0;
}

int ref_trace_geometry( void* deviceGeometry, const void* rays, uint32_t numRays, void* hits, int kind, int reps, double* msOut )
{
	return traceImpl( deviceGeometry, false, static_cast<const hiprtRay*>( rays ), numRays, static_cast<hiprtHit*>( hits ), kind, reps, msOut );
}

int ref_trace_scene( void* deviceScene, const void* rays, uint32_t numRays, void* hits, int kind, int reps, double* msOut )
{
	return traceImpl( deviceScene, true, static_cast<const hiprtRay*>( rays ), numRays, static_cast<hiprtHit*>( hits ), kind, reps, msOut );
}

// Scene of `instanceCount` geometry instances with matrix frames (row-major 3x4 + time = 16 floats,
// hiprtFrameMatrix is 64 bytes). geometryHandles: device hiprtGeometry per instance.
int ref_build_scene(
	void* const* geometryHandles, const float* frames, const uint32_t* masks, uint32_t instanceCount, uint32_t buildFlags, int reps, void** handleOut, double* msOut )
{
	RefScene*				   s = new RefScene;
	std::vector<hiprtInstance> inst( instanceCount );
	for ( uint32_t i = 0; i < instanceCount; ++i )
	{
		inst[i].type	 = hiprtInstanceTypeGeometry;
		inst[i].geometry = static_cast<hiprtGeometry>( geometryHandles[i] );
	}
	CUDA_CHECK( cudaMalloc( &s->dInstances, sizeof( hiprtInstance ) * instanceCount ) );
	CUDA_CHECK( cudaMemcpy( s->dInstances, inst.data(), sizeof( hiprtInstance ) * instanceCount, cudaMemcpyHostToDevice ) );
	CUDA_CHECK( cudaMalloc( &s->dFrames, sizeof( hiprtFrameMatrix ) * instanceCount ) );
	CUDA_CHECK( cudaMemcpy( s->dFrames, frames, sizeof( hiprtFrameMatrix ) * instanceCount, cudaMemcpyHostToDevice ) );
	void* dMasks = nullptr;
	if ( masks )
	{
		CUDA_CHECK( cudaMalloc( &dMasks, 4 * size_t( instanceCount ) ) );
		CUDA_CHECK( cudaMemcpy( dMasks, masks, 4 * size_t( instanceCount ), cudaMemcpyHostToDevice ) );
	}
	hiprtSceneBuildInput input{};
	input.instances				   = s->dInstances;
	input.instanceTransformHeaders = nullptr;
	input.instanceFrames		   = s->dFrames;
	input.instanceMasks			   = dMasks;
	input.instanceCount			   = instanceCount;
	input.frameCount			   = instanceCount;
	input.frameType				   = hiprtFrameTypeMatrix;
	hiprtBuildOptions options{};
	options.buildFlags = buildFlags;
	size_t tempSize	   = 0;
	REF_CHECK( hiprtGetSceneBuildTemporaryBufferSize( g_ctx, input, options, tempSize ) );
	void* temp = nullptr;
	CUDA_CHECK( cudaMalloc( &temp, tempSize ) );
	REF_CHECK( hiprtCreateScene( g_ctx, input, options, s->scene ) );
	double best = 1e30;
	for ( int r = 0; r < reps; ++r )
	{
		CUDA_CHECK( cudaDeviceSynchronize() );
		const auto t0 = std::chrono::high_resolution_clock::now();
		REF_CHECK( hiprtBuildScene( g_ctx, hiprtBuildOperationBuild, input, options, temp, 0, s->scene ) );
		CUDA_CHECK( cudaDeviceSynchronize() );
		const double ms = std::chrono::duration<double, std::milli>( std::chrono::high_resolution_clock::now() - t0 ).count();
		if ( r > 0 || reps == 1 ) best = ms < best ? ms : best;
	}
	CUDA_CHECK( cudaFree( temp ) );
	if ( dMasks ) cudaFree( dMasks );
	*handleOut = s;
	if ( msOut ) *msOut = best;
	return 0;
}

void* ref_scene_device_handle( void* handle ) { return static_cast<RefScene*>( handle )->scene; }

// header 64 bytes; instance nodes 64 bytes each
int ref_scene_dump( void* handle, void* headerOut, void* boxNodesOut, uint32_t boxCount, void* instNodesOut, uint32_t instCount )
{
	RefScene* s = static_cast<RefScene*>( handle );
	void*	  h[8]; // SceneHeader (64 bytes) starts with {boxNodes, primNodes} (hiprt/impl/Header.h:50-60)
	CUDA_CHECK( cudaMemcpy( h, s->scene, 64, cudaMemcpyDeviceToHost ) );
	if ( headerOut ) memcpy( headerOut, h, 64 );
	if ( boxNodesOut ) CUDA_CHECK( cudaMemcpy( boxNodesOut, h[0], size_t( boxCount ) * 128, cudaMemcpyDeviceToHost ) );
	if ( instNodesOut ) CUDA_CHECK( cudaMemcpy( instNodesOut, h[1], size_t( instCount ) * 64, cudaMemcpyDeviceToHost ) );
	return 0;
}

int ref_destroy_scene( void* handle )
{
	RefScene* s = static_cast<RefScene*>( handle );
	hiprtDestroyScene( g_ctx, s->scene );
	cudaFree( s->dInstances );
	cudaFree( s->dFrames );
	delete s;
	return 0;
}

} // extern "C"
