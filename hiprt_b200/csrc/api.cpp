// C-ABI shim: the 37 hiprt* entry points + the hiprtB200* additions.
// Mirrors the behaviour of the reference's shim/context (hiprt/impl/hiprt.cpp:34-776,
// hiprt/impl/Context.cpp): argument validation -> hiprtErrorInvalidParameter, API version gate,
// internal failures are caught, logged through the context's log mask and returned as
// hiprtErrorInternal; geometries/scenes created together share one device allocation that is
// released when its last member is destroyed (Context.cpp:99-142).
// Runtime: plain CUDA runtime API on the caller's device (no Orochi, no backend switch).
#include <hiprt/hiprt_b200.h>

#include "builders.h"
#include "compiler.h"
#include "frame.h"
#include "qnode.h"
#include "trace.h"

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>
#include <unordered_map>
#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>

using namespace hiprt;

namespace
{
void checkCuda( cudaError_t e, const char* what )
{
	if ( e != cudaSuccess ) throw std::runtime_error( std::string( what ) + ": " + cudaGetErrorString( e ) );
}

struct OutOfDeviceMemory : std::runtime_error
{
	using std::runtime_error::runtime_error;
};

// End-to-end path with HOST buffers: rays are pushed, traced and the hits pulled back in chunks on
// three streams so that the H2D copy of chunk c+1, the traversal of chunk c and the D2H copy of chunk
// c-1 overlap (PCIe is full duplex; per ray 32 B go up and 48 B come down, so the call is bound by the
// D2H direction). Device staging buffers and streams live in the context and are reused across calls.
// Pageable host memory still works (the copies then serialise inside the driver).
class HostTracePipeline
{
  public:
	static constexpr uint32_t MaxChunk = 1u << 20;
	// rays per chunk (default 256 Ki: 8 MiB up, 12 MiB down); HIPRT_B200_HOST_CHUNK overrides it for tuning
	const uint32_t Chunk = [] {
		const char* e = getenv( "HIPRT_B200_HOST_CHUNK" );
		const long	v = e ? atol( e ) : 0;
		return v >= 1024 && v <= long( MaxChunk ) ? uint32_t( v ) : 1u << 18;
	}();
	static constexpr int	  Slots = 3;

	~HostTracePipeline()
	{
		if ( !m_ready ) return;
		for ( int i = 0; i < Slots; ++i )
		{
			cudaFree( m_rays[i] );
			cudaFree( m_hits[i] );
			cudaEventDestroy( m_up[i] );
			cudaEventDestroy( m_traced[i] );
			cudaEventDestroy( m_down[i] );
		}
		cudaEventDestroy( m_begin );
		cudaStreamDestroy( m_h2d );
		cudaStreamDestroy( m_compute );
		cudaStreamDestroy( m_d2h );
	}

	void run(
		hiprtGeometry geometry, bool triangles, hiprtScene scene, const hiprtRay* hostRays, uint32_t numRays, void* hostRecords, hiprtTraversalType type, hiprtRayMask mask, cudaStream_t user,
		uint32_t smCount, uint32_t format = hiprtB200HitFormatFull, const void* qRegion = nullptr )
	{
		const size_t recordSize = hiprtB200HitFormatSize( hiprtB200HitFormat( format ) );
		uint8_t*	 hostHits	= static_cast<uint8_t*>( hostRecords );
		std::lock_guard<std::mutex> lock( m_mutex );
		init();
		// order after whatever the caller has queued on its stream (e.g. the build)
		checkCuda( cudaEventRecord( m_begin, user ), "cudaEventRecord" );
		checkCuda( cudaStreamWaitEvent( m_h2d, m_begin, 0 ), "cudaStreamWaitEvent" );
		checkCuda( cudaStreamWaitEvent( m_compute, m_begin, 0 ), "cudaStreamWaitEvent" );
		const uint32_t chunks = ( numRays + Chunk - 1 ) / Chunk;
		for ( uint32_t c = 0; c < chunks; ++c )
		{
			const int	   slot	 = int( c % Slots );
			const uint32_t first = c * Chunk, count = std::min( Chunk, numRays - first );
			if ( c >= uint32_t( Slots ) ) checkCuda( cudaStreamWaitEvent( m_h2d, m_traced[slot], 0 ), "cudaStreamWaitEvent" );
			checkCuda( cudaMemcpyAsync( m_rays[slot], hostRays + first, size_t( count ) * sizeof( hiprtRay ), cudaMemcpyHostToDevice, m_h2d ), "H2D rays" );
			checkCuda( cudaEventRecord( m_up[slot], m_h2d ), "cudaEventRecord" );

			checkCuda( cudaStreamWaitEvent( m_compute, m_up[slot], 0 ), "cudaStreamWaitEvent" );
			if ( c >= uint32_t( Slots ) ) checkCuda( cudaStreamWaitEvent( m_compute, m_down[slot], 0 ), "cudaStreamWaitEvent" );
			if ( geometry )
				traceGeometry( geometry, triangles, m_rays[slot], count, m_hits[slot], type, m_counters + slot * 16, smCount, m_compute, format, qRegion );
			else
				traceScene( scene, m_rays[slot], count, m_hits[slot], type, mask, m_counters + slot * 16, smCount, m_compute );
			checkCuda( cudaGetLastError(), "trace launch" );
			checkCuda( cudaEventRecord( m_traced[slot], m_compute ), "cudaEventRecord" );

			checkCuda( cudaStreamWaitEvent( m_d2h, m_traced[slot], 0 ), "cudaStreamWaitEvent" );
			checkCuda( cudaMemcpyAsync( hostHits + size_t( first ) * recordSize, m_hits[slot], size_t( count ) * recordSize, cudaMemcpyDeviceToHost, m_d2h ), "D2H hits" );
			checkCuda( cudaEventRecord( m_down[slot], m_d2h ), "cudaEventRecord" );
		}
		checkCuda( cudaStreamSynchronize( m_d2h ), "cudaStreamSynchronize" );
		launches += chunks;
	}

	uint64_t launches = 0;

  private:
	void init()
	{
		if ( m_ready ) return;
		checkCuda( cudaStreamCreateWithFlags( &m_h2d, cudaStreamNonBlocking ), "cudaStreamCreate" );
		checkCuda( cudaStreamCreateWithFlags( &m_compute, cudaStreamNonBlocking ), "cudaStreamCreate" );
		checkCuda( cudaStreamCreateWithFlags( &m_d2h, cudaStreamNonBlocking ), "cudaStreamCreate" );
		checkCuda( cudaEventCreateWithFlags( &m_begin, cudaEventDisableTiming ), "cudaEventCreate" );
		checkCuda( cudaMalloc( reinterpret_cast<void**>( &m_counters ), Slots * 64 ), "cudaMalloc" );
		for ( int i = 0; i < Slots; ++i )
		{
			checkCuda( cudaMalloc( reinterpret_cast<void**>( &m_rays[i] ), size_t( Chunk ) * sizeof( hiprtRay ) ), "cudaMalloc" );
			checkCuda( cudaMalloc( reinterpret_cast<void**>( &m_hits[i] ), size_t( Chunk ) * sizeof( hiprtHit ) ), "cudaMalloc" );
			checkCuda( cudaEventCreateWithFlags( &m_up[i], cudaEventDisableTiming ), "cudaEventCreate" );
			checkCuda( cudaEventCreateWithFlags( &m_traced[i], cudaEventDisableTiming ), "cudaEventCreate" );
			checkCuda( cudaEventCreateWithFlags( &m_down[i], cudaEventDisableTiming ), "cudaEventCreate" );
		}
		m_ready = true;
	}

	bool		 m_ready = false;
	std::mutex	 m_mutex;
	cudaStream_t m_h2d = nullptr, m_compute = nullptr, m_d2h = nullptr;
	cudaEvent_t	 m_begin = nullptr, m_up[Slots]{}, m_traced[Slots]{}, m_down[Slots]{};
	hiprtRay*	 m_rays[Slots]{};
	uint32_t*	 m_counters = nullptr;
	hiprtHit*	 m_hits[Slots]{};
};

class Context
{
  public:
	explicit Context( const hiprtContextCreationInput& input )
	{
		if ( input.deviceType != hiprtDeviceNVIDIA ) throw std::runtime_error( "hiprt_b200 supports NVIDIA devices only" );
		m_info.device = input.device;
		m_userCtx	  = input.ctxt;
		bind();
		cudaDeviceProp prop{};
		checkCuda( cudaGetDeviceProperties( &prop, m_info.device ), "cudaGetDeviceProperties" );
		m_info.smCount	 = uint32_t( prop.multiProcessorCount );
		m_maxThreadsPerSm = uint32_t( prop.maxThreadsPerMultiProcessor > 0 ? prop.maxThreadsPerMultiProcessor : 2048 );
		m_ccMajor = prop.major, m_ccMinor = prop.minor;
		m_info.profile = &m_profile;
	}

	~Context()
	{
		// `modules` unloads its CUmodules when the members are destroyed: make this context's device current for that
		try
		{
			bind();
		}
		catch ( ... )
		{
		}
		if ( m_profileEventsCreated )
			for ( auto& e : m_profile.events )
				cudaEventDestroy( e );
	}

	void enableProfile( bool on )
	{
		if ( on && !m_profileEventsCreated )
		{
			for ( auto& e : m_profile.events )
				checkCuda( cudaEventCreate( &e ), "cudaEventCreate" );
			m_profileEventsCreated = true;
		}
		m_profile.enabled = on;
		m_profile.count	  = 0;
	}
	BuildProfile& profile() { return m_profile; }

	// Make the context's device current on the calling thread (builds may come from many host
	// threads, each with its own stream: reference test/hiprtTest.cpp:1226-1352).
	void bind() const
	{
		checkCuda( cudaSetDevice( m_info.device ), "cudaSetDevice" );
		if ( m_userCtx ) compiler::setCurrentContext( m_userCtx );
	}

	const DeviceInfo& info() const { return m_info; }
	uint32_t		  maxThreadsPerSm() const { return m_maxThreadsPerSm; }
	int				  ccMajor() const { return m_ccMajor; }
	int				  ccMinor() const { return m_ccMinor; }

	void log( hiprtLogLevel level, const char* fmt, ... )
	{
		va_list args;
		va_start( args, fmt );
		char buf[1024];
		vsnprintf( buf, sizeof( buf ), fmt, args );
		va_end( args );
		if ( level == hiprtLogLevelError )
		{
			std::lock_guard<std::mutex> lock( m_mutex );
			m_lastError = buf;
		}
		if ( m_logMask & level ) fprintf( level == hiprtLogLevelInfo ? stdout : stderr, "[hiprt_b200] %s\n", buf );
	}
	void		setLogLevel( hiprtLogLevel mask ) { m_logMask = mask; }
	const char* lastError() const { return m_lastError.c_str(); }

	// ---- pooled device allocations ------------------------------------------------------------
	std::vector<void*> allocatePool( const std::vector<size_t>& sizes )
	{
		size_t total = 0;
		for ( size_t s : sizes )
			total += s;
		void*		base = nullptr;
		cudaError_t e	 = cudaMalloc( &base, std::max<size_t>( total, 64 ) );
		if ( e == cudaErrorMemoryAllocation ) throw OutOfDeviceMemory( "cudaMalloc: out of device memory" );
		checkCuda( e, "cudaMalloc" );
		std::vector<void*> out( sizes.size() );
		uint8_t*		   p = static_cast<uint8_t*>( base );
		for ( size_t i = 0; i < sizes.size(); ++i )
		{
			out[i] = p;
			p += sizes[i];
		}
		std::lock_guard<std::mutex> lock( m_mutex );
		m_pools[base] = Pool{ total, uint32_t( sizes.size() ) };
		return out;
	}

	void releaseFromPool( void* member )
	{
		std::lock_guard<std::mutex> lock( m_mutex );
		auto						it = m_pools.upper_bound( member );
		if ( it == m_pools.begin() )
		{
			log( hiprtLogLevelWarn, "destroying an object that was not allocated by this context" );
			return;
		}
		--it;
		uint8_t* base = static_cast<uint8_t*>( it->first );
		if ( static_cast<uint8_t*>( member ) >= base + std::max<size_t>( it->second.size, 64 ) )
		{
			log( hiprtLogLevelWarn, "destroying an object that was not allocated by this context" );
			return;
		}
		if ( --it->second.members == 0 )
		{
			checkCuda( cudaFree( it->first ), "cudaFree" );
			m_pools.erase( it );
		}
	}

	std::string			  cacheDir;
	HostTracePipeline	  hostPipeline;
	compiler::ModuleCache modules; // user trace kernels loaded into THIS context's CUDA context

	// ---- what the library knows about a geometry handle without a device read: its primitive kind (chooses the trace kernel)
	// and its traversal-only Q region (qnode.h), nullptr when it has none ----
	struct GeomInfo
	{
		bool   triangles = false;
		void*  qRegion	 = nullptr;
		size_t blobSize	 = 0; // hiprtB200AllocateGeometryBlob: bytes reserved for the public structure (the Q region follows)
	};
	void noteGeometryKind( void* geometry, bool triangles, void* qRegion = nullptr, size_t blobSize = 0 )
	{
		std::lock_guard<std::mutex> lock( m_mutex );
		m_geomKind[geometry] = GeomInfo{ triangles, qRegion, blobSize };
	}
	// (batch builds register 100 000 geometries per call: one lock, one rehash)
	void noteGeometryKinds( hiprtGeometry* geometries, const hiprtGeometryBuildInput* inputs, uint32_t count, const std::vector<void*>& qRegions )
	{
		std::lock_guard<std::mutex> lock( m_mutex );
		m_geomKind.reserve( m_geomKind.size() + count );
		for ( uint32_t i = 0; i < count; ++i )
			m_geomKind[geometries[i]] = GeomInfo{ inputs[i].type == hiprtPrimitiveTypeTriangleMesh, qRegions.empty() ? nullptr : qRegions[i], 0 };
	}
	void forgetGeometry( void* geometry )
	{
		std::lock_guard<std::mutex> lock( m_mutex );
		m_geomKind.erase( geometry );
	}
	void forgetGeometries( hiprtGeometry* geometries, uint32_t count )
	{
		std::lock_guard<std::mutex> lock( m_mutex );
		for ( uint32_t i = 0; i < count; ++i )
			m_geomKind.erase( geometries[i] );
	}
	GeomInfo geometryInfo( void* geometry )
	{
		std::lock_guard<std::mutex> lock( m_mutex );
		auto						it = m_geomKind.find( geometry );
		return it != m_geomKind.end() ? it->second : GeomInfo{};
	}
	// Geometries this context did not build (e.g. written by another library with the same format): the header
	// word is read back once per call, after the work queued on `stream`.
	bool geometryHasTriangles( void* geometry, cudaStream_t stream, const void** qRegionOut = nullptr )
	{
		{
			std::lock_guard<std::mutex> lock( m_mutex );
			auto						it = m_geomKind.find( geometry );
			if ( it != m_geomKind.end() )
			{
				if ( qRegionOut ) *qRegionOut = it->second.qRegion;
				return it->second.triangles;
			}
		}
		if ( qRegionOut ) *qRegionOut = nullptr;
		uint32_t geomType = 0;
		checkCuda(
			cudaMemcpyAsync( &geomType, static_cast<uint8_t*>( geometry ) + offsetof( GeomHeader, geomType ), sizeof( uint32_t ), cudaMemcpyDeviceToHost, stream ),
			"cudaMemcpyAsync" );
		checkCuda( cudaStreamSynchronize( stream ), "cudaStreamSynchronize" );
		// remembered: a handle this context did not build itself must not cost a synchronisation per launch;
		// hiprtDestroyGeometry forgets it again
		noteGeometryKind( geometry, geomType & 1u );
		return geomType & 1u;
	}

	// Ray-batch counter of the persistent trace kernels: ONE 64-byte slot per stream. Launches on one stream run one after the other
	// (the counter is reset on the stream right before each launch), launches on different streams never share a slot however many
	// of them are queued -- a wavefront renderer queues hundreds per stream (hiprt_b200/pathtrace.py). (The reference has no
	// counterpart: its kernels are the caller's, one thread per ray.)
	uint32_t* workCounterFor( cudaStream_t stream )
	{
		std::lock_guard<std::mutex> lock( m_mutex );
		constexpr uint32_t			Slots = 1024;
		if ( !m_workCounters ) checkCuda( cudaMalloc( reinterpret_cast<void**>( &m_workCounters ), size_t( Slots ) * 64 ), "cudaMalloc" );
		auto it = m_workCounterOfStream.find( stream );
		if ( it == m_workCounterOfStream.end() )
		{
			// (more streams than slots: the oldest assignments are forgotten; a slot is then shared by two streams only if both are
			// still in use 1024 streams apart)
			it = m_workCounterOfStream.emplace( stream, m_workCounterNext++ % Slots ).first;
		}
		return m_workCounters + size_t( it->second ) * 16;
	}

  private:
	struct Pool
	{
		size_t	 size;
		uint32_t members;
	};
	DeviceInfo			  m_info;
	BuildProfile		  m_profile;
	bool				  m_profileEventsCreated = false;
	void*				  m_userCtx			= nullptr;
	uint32_t			  m_maxThreadsPerSm = 2048;
	int					  m_ccMajor = 10, m_ccMinor = 0;
	hiprtLogLevel		  m_logMask = hiprtLogLevelNone;
	std::string			  m_lastError;
	std::map<void*, Pool> m_pools;
	std::unordered_map<void*, GeomInfo> m_geomKind;
	std::mutex			  m_mutex;
	uint32_t*			  m_workCounters	= nullptr;
	uint32_t			  m_workCounterNext = 0;
	std::unordered_map<cudaStream_t, uint32_t> m_workCounterOfStream;
};

inline Context* ctx( hiprtContext c ) { return reinterpret_cast<Context*>( c ); }

// Runs `fn` with the reference's error convention.
template <typename Fn>
hiprtError guarded( hiprtContext context, Fn&& fn )
{
	try
	{
		ctx( context )->bind();
		fn();
	}
	catch ( const OutOfDeviceMemory& e )
	{
		ctx( context )->log( hiprtLogLevelError, "%s", e.what() );
		return hiprtErrorOutOfDeviceMemory;
	}
	catch ( const std::bad_alloc& )
	{
		return hiprtErrorOutOfHostMemory;
	}
	catch ( const compiler::NotAvailable& e )
	{
		ctx( context )->log( hiprtLogLevelError, "%s", e.what() );
		return hiprtErrorNotImplemented;
	}
	catch ( const std::exception& e )
	{
		ctx( context )->log( hiprtLogLevelError, "%s", e.what() );
		return hiprtErrorInternal;
	}
	return hiprtSuccess;
}

bool usesBatch( size_t primCount, hiprtBuildOptions o )
{
	// reference: hiprt/impl/BvhCommon.h:155-166
	return primCount <= o.batchBuildMaxPrimCount && ( o.buildFlags & 7 ) != hiprtBuildFlagBitCustomBvhImport;
}

void checkBuildFlags( Context* c, hiprtBuildOptions o )
{
	if ( ( o.buildFlags & 3 ) == hiprtBuildFlagBitPreferHighQualityBuild ) c->log( hiprtLogLevelInfo, "HighQualityBuild (SBVH)" );
}
} // namespace

extern "C" {

// ------------------------------------------------------------------------------------------------
hiprtError hiprtCreateContext( uint32_t hiprtApiVersion, const hiprtContextCreationInput& input, hiprtContext& contextOut )
{
	if ( hiprtApiVersion != HIPRT_API_VERSION ) return hiprtErrorInvalidApiVersion;
	try
	{
		contextOut = reinterpret_cast<hiprtContext>( new Context( input ) );
	}
	catch ( const std::exception& e )
	{
		fprintf( stderr, "[hiprt_b200] hiprtCreateContext failed: %s\n", e.what() );
		contextOut = nullptr;
		return hiprtErrorInternal;
	}
	return hiprtSuccess;
}

hiprtError hiprtDestroyContext( hiprtContext context )
{
	if ( !context ) return hiprtErrorInvalidParameter;
	delete ctx( context );
	return hiprtSuccess;
}

// ------------------------------------------------------------------------------------------------
hiprtError hiprtCreateGeometry(
	hiprtContext context, const hiprtGeometryBuildInput& buildInput, const hiprtBuildOptions buildOptions, hiprtGeometry& geometryOut )
{
	hiprtGeometry* addr = &geometryOut;
	return hiprtCreateGeometries( context, 1, &buildInput, buildOptions, &addr );
}

hiprtError hiprtCreateGeometries(
	hiprtContext				   context,
	uint32_t					   numGeometries,
	const hiprtGeometryBuildInput* buildInputs,
	const hiprtBuildOptions		   buildOptions,
	hiprtGeometry**				   geometriesOut )
{
	if ( !context || numGeometries == 0 || buildInputs == nullptr || geometriesOut == nullptr ) return hiprtErrorInvalidParameter;
	for ( uint32_t i = 0; i < numGeometries; ++i )
		if ( geometriesOut[i] == nullptr ) return hiprtErrorInvalidParameter;
	return guarded( context, [&] {
		checkBuildFlags( ctx( context ), buildOptions );
		std::vector<size_t> sizes( numGeometries );
		for ( uint32_t i = 0; i < numGeometries; ++i )
			sizes[i] = geometryAllocationSize( buildInputs[i], buildOptions );
		const std::vector<void*> ptrs = ctx( context )->allocatePool( sizes );
		for ( uint32_t i = 0; i < numGeometries; ++i )
			*geometriesOut[i] = reinterpret_cast<hiprtGeometry>( ptrs[i] );
	} );
}

hiprtError hiprtDestroyGeometry( hiprtContext context, hiprtGeometry geometry ) { return hiprtDestroyGeometries( context, 1, &geometry ); }

hiprtError hiprtDestroyGeometries( hiprtContext context, uint32_t numGeometries, hiprtGeometry* geometries )
{
	if ( !context || numGeometries == 0 || geometries == nullptr ) return hiprtErrorInvalidParameter;
	for ( uint32_t i = 0; i < numGeometries; ++i )
		if ( !geometries[i] ) return hiprtErrorInvalidParameter;
	return guarded( context, [&] {
		ctx( context )->forgetGeometries( geometries, numGeometries );
		for ( uint32_t i = 0; i < numGeometries; ++i )
			ctx( context )->releaseFromPool( geometries[i] );
	} );
}

hiprtError hiprtBuildGeometry(
	hiprtContext				   context,
	hiprtBuildOperation			   buildOperation,
	const hiprtGeometryBuildInput& buildInput,
	const hiprtBuildOptions		   buildOptions,
	hiprtDevicePtr				   temporaryBuffer,
	hiprtApiStream				   stream,
	hiprtGeometry				   geometryOut )
{
	return hiprtBuildGeometries( context, buildOperation, 1, &buildInput, buildOptions, temporaryBuffer, stream, &geometryOut );
}

hiprtError hiprtBuildGeometries(
	hiprtContext				   context,
	hiprtBuildOperation			   buildOperation,
	uint32_t					   numGeometries,
	const hiprtGeometryBuildInput* buildInputs,
	const hiprtBuildOptions		   buildOptions,
	hiprtDevicePtr				   temporaryBuffer,
	hiprtApiStream				   stream,
	hiprtGeometry*				   geometriesOut )
{
	if ( !context || numGeometries == 0 || buildInputs == nullptr || geometriesOut == nullptr ) return hiprtErrorInvalidParameter;
	for ( uint32_t i = 0; i < numGeometries; ++i )
		if ( geometriesOut[i] == nullptr ) return hiprtErrorInvalidParameter;
	if ( buildOperation != hiprtBuildOperationBuild && buildOperation != hiprtBuildOperationUpdate ) return hiprtErrorInvalidParameter;
	return guarded( context, [&] {
		Context* c = ctx( context );
		checkBuildFlags( c, buildOptions );
		cudaStream_t s = static_cast<cudaStream_t>( stream );
		if ( buildOperation == hiprtBuildOperationUpdate )
		{
			// the structure is refitted where it lies (it may have been compacted, loaded or relocated since the build): its Q
			// region, if it has one, is the one on record for the handle
			for ( uint32_t i = 0; i < numGeometries; ++i )
				updateGeometry( c->info(), buildInputs[i], buildOptions, temporaryBuffer, s, geometriesOut[i], c->geometryInfo( geometriesOut[i] ).qRegion );
			return;
		}
		// Q regions: single (non-batch) builds of triangle meshes only
		std::vector<void*> qRegions;
		{
			bool any = false;
			for ( uint32_t i = 0; i < numGeometries && !any; ++i )
				any = !usesBatch( geometryPrimCount( buildInputs[i] ), buildOptions ) && geometryQRegion( buildInputs[i], buildOptions, geometriesOut[i] );
			if ( any )
			{
				qRegions.resize( numGeometries, nullptr );
				for ( uint32_t i = 0; i < numGeometries; ++i )
					if ( !usesBatch( geometryPrimCount( buildInputs[i] ), buildOptions ) ) qRegions[i] = geometryQRegion( buildInputs[i], buildOptions, geometriesOut[i] );
			}
		}
		c->noteGeometryKinds( geometriesOut, buildInputs, numGeometries, qRegions );
		// inputs small enough for the batch builder are built together first (Context.cpp:153-168)
		uint32_t batched = 0;
		for ( uint32_t i = 0; i < numGeometries; ++i )
			batched += usesBatch( geometryPrimCount( buildInputs[i] ), buildOptions ) ? 1u : 0u;
		if ( batched == numGeometries )
		{
			// the common case of a pure batch: the caller's arrays go to the device as they are
			buildGeometriesBatch( c->info(), buildInputs, numGeometries, buildOptions, temporaryBuffer, s, reinterpret_cast<void* const*>( geometriesOut ) );
			return;
		}
		std::vector<hiprtGeometryBuildInput> batchInputs;
		std::vector<void*>					 batchOut;
		batchInputs.reserve( batched ), batchOut.reserve( batched );
		for ( uint32_t i = 0; i < numGeometries; ++i )
			if ( usesBatch( geometryPrimCount( buildInputs[i] ), buildOptions ) )
			{
				batchInputs.push_back( buildInputs[i] );
				batchOut.push_back( geometriesOut[i] );
			}
		if ( !batchInputs.empty() )
			buildGeometriesBatch( c->info(), batchInputs.data(), uint32_t( batchInputs.size() ), buildOptions, temporaryBuffer, s, batchOut.data() );
		for ( uint32_t i = 0; i < numGeometries; ++i )
			if ( !usesBatch( geometryPrimCount( buildInputs[i] ), buildOptions ) )
				buildGeometry( c->info(), buildInputs[i], buildOptions, temporaryBuffer, s, geometriesOut[i], qRegions.empty() ? nullptr : qRegions[i] );
	} );
}

hiprtError hiprtGetGeometryBuildTemporaryBufferSize(
	hiprtContext context, const hiprtGeometryBuildInput& buildInput, const hiprtBuildOptions buildOptions, size_t& sizeOut )
{
	return hiprtGetGeometriesBuildTemporaryBufferSize( context, 1, &buildInput, buildOptions, sizeOut );
}

hiprtError hiprtGetGeometriesBuildTemporaryBufferSize(
	hiprtContext				   context,
	uint32_t					   numGeometries,
	const hiprtGeometryBuildInput* buildInputs,
	const hiprtBuildOptions		   buildOptions,
	size_t&						   sizeOut )
{
	if ( !context || numGeometries == 0 || buildInputs == nullptr ) return hiprtErrorInvalidParameter;
	return guarded( context, [&] {
		// non-batch builds run one after another on the stream: max; the batch needs its own slice
		size_t	 size	   = 0;
		uint32_t batchSize = 0;
		for ( uint32_t i = 0; i < numGeometries; ++i )
		{
			if ( usesBatch( geometryPrimCount( buildInputs[i] ), buildOptions ) )
				++batchSize;
			else
				size = std::max( size, geometryTempSize( buildInputs[i], buildOptions ) );
		}
		if ( batchSize ) size = std::max( size, batchTempSize( batchSize ) );
		sizeOut = size;
	} );
}

hiprtError hiprtCompactGeometry( hiprtContext context, hiprtApiStream stream, hiprtGeometry geometryIn, hiprtGeometry& geometryOut )
{
	hiprtGeometry* addr = &geometryOut;
	return hiprtCompactGeometries( context, 1, stream, &geometryIn, &addr );
}

hiprtError hiprtCompactGeometries(
	hiprtContext context, uint32_t numGeometries, hiprtApiStream stream, hiprtGeometry* geometriesIn, hiprtGeometry** geometriesOut )
{
	if ( !context || numGeometries == 0 || geometriesIn == nullptr || geometriesOut == nullptr ) return hiprtErrorInvalidParameter;
	for ( uint32_t i = 0; i < numGeometries; ++i )
		if ( geometriesIn[i] == nullptr || geometriesOut[i] == nullptr ) return hiprtErrorInvalidParameter;
	return guarded( context, [&] {
		// reference: Context::compactGeometries (Context.cpp:292-355): read the headers back, allocate
		// exactly what is used, copy box and leaf arrays, rebase the header pointers, free the inputs.
		Context*				c = ctx( context );
		cudaStream_t			s = static_cast<cudaStream_t>( stream );
		std::vector<GeomHeader> headers( numGeometries );
		for ( uint32_t i = 0; i < numGeometries; ++i )
			checkCuda( cudaMemcpyAsync( &headers[i], geometriesIn[i], sizeof( GeomHeader ), cudaMemcpyDeviceToHost, s ), "cudaMemcpyAsync" );
		checkCuda( cudaStreamSynchronize( s ), "cudaStreamSynchronize" );
		std::vector<size_t> sizes( numGeometries ), boxBytes( numGeometries ), leafBytes( numGeometries );
		for ( uint32_t i = 0; i < numGeometries; ++i )
		{
			const size_t leafSize = ( headers[i].geomType & 1u ) ? sizeof( TrianglePairNode ) : sizeof( CustomNode );
			boxBytes[i]			  = RoundUp( size_t( headers[i].boxNodeCount ) * sizeof( Box4Node ), size_t( 64 ) );
			leafBytes[i]		  = RoundUp( size_t( headers[i].primNodeCount ) * leafSize, size_t( 64 ) );
			sizes[i]			  = 64 + boxBytes[i] + leafBytes[i];
			// a geometry that has a Q region keeps one: regenerated behind the compacted structure
			if ( c->geometryInfo( geometriesIn[i] ).qRegion ) sizes[i] += qRegionSize( headers[i].boxNodeCount );
		}
		const std::vector<void*> out = c->allocatePool( sizes );
		for ( uint32_t i = 0; i < numGeometries; ++i )
		{
			uint8_t*   dst = static_cast<uint8_t*>( out[i] );
			GeomHeader h   = headers[i];
			checkCuda( cudaMemcpyAsync( dst + 64, headers[i].boxNodes, boxBytes[i], cudaMemcpyDeviceToDevice, s ), "copy box nodes" );
			checkCuda( cudaMemcpyAsync( dst + 64 + boxBytes[i], headers[i].primNodes, leafBytes[i], cudaMemcpyDeviceToDevice, s ), "copy leaves" );
			h.boxNodes	= reinterpret_cast<Box4Node*>( dst + 64 );
			h.primNodes = dst + 64 + boxBytes[i];
			h.size		= 64 + boxBytes[i] + leafBytes[i];
			checkCuda( cudaMemcpyAsync( dst, &h, sizeof( GeomHeader ), cudaMemcpyHostToDevice, s ), "write header" );
			*geometriesOut[i] = reinterpret_cast<hiprtGeometry>( dst );
			void* q = sizes[i] > h.size ? dst + h.size : nullptr;
			compressBoxNodes( reinterpret_cast<const GeomHeader*>( dst ), q, headers[i].boxNodeCount, s );
			c->noteGeometryKind( dst, headers[i].geomType & 1u, q );
		}
		checkCuda( cudaStreamSynchronize( s ), "cudaStreamSynchronize" );
		for ( uint32_t i = 0; i < numGeometries; ++i )
		{
			c->forgetGeometry( geometriesIn[i] );
			c->releaseFromPool( geometriesIn[i] );
		}
	} );
}

// ------------------------------------------------------------------------------------------------
hiprtError
hiprtCreateScene( hiprtContext context, const hiprtSceneBuildInput& buildInput, const hiprtBuildOptions buildOptions, hiprtScene& sceneOut )
{
	hiprtScene* addr = &sceneOut;
	return hiprtCreateScenes( context, 1, &buildInput, buildOptions, &addr );
}

hiprtError hiprtCreateScenes(
	hiprtContext context, uint32_t numScenes, const hiprtSceneBuildInput* buildInputs, const hiprtBuildOptions buildOptions, hiprtScene** scenesOut )
{
	if ( !context || numScenes == 0 || buildInputs == nullptr || scenesOut == nullptr ) return hiprtErrorInvalidParameter;
	for ( uint32_t i = 0; i < numScenes; ++i )
		if ( scenesOut[i] == nullptr ) return hiprtErrorInvalidParameter;
	return guarded( context, [&] {
		checkBuildFlags( ctx( context ), buildOptions );
		std::vector<size_t> sizes( numScenes );
		for ( uint32_t i = 0; i < numScenes; ++i )
		{
			// instance ids are 24 bits wide (reference: Context.cpp:366-373)
			if ( buildInputs[i].instanceCount >= ( 1u << InstanceIDBits ) ) throw std::runtime_error( "too many instances (24-bit instance ids)" );
			sizes[i] = sceneStorageSize( buildInputs[i], buildOptions );
		}
		const std::vector<void*> ptrs = ctx( context )->allocatePool( sizes );
		for ( uint32_t i = 0; i < numScenes; ++i )
			*scenesOut[i] = reinterpret_cast<hiprtScene>( ptrs[i] );
	} );
}

hiprtError hiprtDestroyScene( hiprtContext context, hiprtScene scene ) { return hiprtDestroyScenes( context, 1, &scene ); }

hiprtError hiprtDestroyScenes( hiprtContext context, uint32_t numScenes, hiprtScene* scenes )
{
	if ( !context || numScenes == 0 || scenes == nullptr ) return hiprtErrorInvalidParameter;
	for ( uint32_t i = 0; i < numScenes; ++i )
		if ( !scenes[i] ) return hiprtErrorInvalidParameter;
	return guarded( context, [&] {
		for ( uint32_t i = 0; i < numScenes; ++i )
			ctx( context )->releaseFromPool( scenes[i] );
	} );
}

hiprtError hiprtBuildScene(
	hiprtContext				context,
	hiprtBuildOperation			buildOperation,
	const hiprtSceneBuildInput& buildInput,
	const hiprtBuildOptions		buildOptions,
	hiprtDevicePtr				temporaryBuffer,
	hiprtApiStream				stream,
	hiprtScene					sceneOut )
{
	return hiprtBuildScenes( context, buildOperation, 1, &buildInput, buildOptions, temporaryBuffer, stream, &sceneOut );
}

hiprtError hiprtBuildScenes(
	hiprtContext				context,
	hiprtBuildOperation			buildOperation,
	uint32_t					numScenes,
	const hiprtSceneBuildInput* buildInputs,
	const hiprtBuildOptions		buildOptions,
	hiprtDevicePtr				temporaryBuffer,
	hiprtApiStream				stream,
	hiprtScene*					scenesOut )
{
	if ( !context || numScenes == 0 || buildInputs == nullptr || scenesOut == nullptr ) return hiprtErrorInvalidParameter;
	for ( uint32_t i = 0; i < numScenes; ++i )
		if ( scenesOut[i] == nullptr ) return hiprtErrorInvalidParameter;
	if ( buildOperation != hiprtBuildOperationBuild && buildOperation != hiprtBuildOperationUpdate ) return hiprtErrorInvalidParameter;
	return guarded( context, [&] {
		Context* c = ctx( context );
		checkBuildFlags( c, buildOptions );
		cudaStream_t s	  = static_cast<cudaStream_t>( stream );
		uint8_t*	 temp = static_cast<uint8_t*>( temporaryBuffer );
		for ( uint32_t i = 0; i < numScenes; ++i )
		{
			// several scenes get disjoint slices of the temporary buffer (sizes add up: Context.cpp:598-629)
			if ( buildOperation == hiprtBuildOperationUpdate )
				updateScene( c->info(), buildInputs[i], buildOptions, temp, s, scenesOut[i] );
			else
				buildScene( c->info(), buildInputs[i], buildOptions, temp, s, scenesOut[i] );
			temp += sceneTempSize( buildInputs[i], buildOptions );
		}
	} );
}

hiprtError hiprtGetSceneBuildTemporaryBufferSize(
	hiprtContext context, const hiprtSceneBuildInput& buildInput, const hiprtBuildOptions buildOptions, size_t& sizeOut )
{
	return hiprtGetScenesBuildTemporaryBufferSize( context, 1, &buildInput, buildOptions, sizeOut );
}

hiprtError hiprtGetScenesBuildTemporaryBufferSize(
	hiprtContext context, uint32_t numScenes, const hiprtSceneBuildInput* buildInputs, const hiprtBuildOptions buildOptions, size_t& sizeOut )
{
	if ( !context || numScenes == 0 || buildInputs == nullptr ) return hiprtErrorInvalidParameter;
	return guarded( context, [&] {
		size_t size = 0;
		for ( uint32_t i = 0; i < numScenes; ++i )
			size += sceneTempSize( buildInputs[i], buildOptions );
		sizeOut = size;
	} );
}

hiprtError hiprtCompactScene( hiprtContext context, hiprtApiStream stream, hiprtScene sceneIn, hiprtScene& sceneOut )
{
	hiprtScene* addr = &sceneOut;
	return hiprtCompactScenes( context, 1, stream, &sceneIn, &addr );
}

hiprtError hiprtCompactScenes( hiprtContext context, uint32_t numScenes, hiprtApiStream stream, hiprtScene* scenesIn, hiprtScene** scenesOut )
{
	if ( !context || numScenes == 0 || scenesIn == nullptr || scenesOut == nullptr ) return hiprtErrorInvalidParameter;
	for ( uint32_t i = 0; i < numScenes; ++i )
		if ( scenesIn[i] == nullptr || scenesOut[i] == nullptr ) return hiprtErrorInvalidParameter;
	return guarded( context, [&] {
		Context*				 c = ctx( context );
		cudaStream_t			 s = static_cast<cudaStream_t>( stream );
		std::vector<SceneHeader> headers( numScenes );
		for ( uint32_t i = 0; i < numScenes; ++i )
			checkCuda( cudaMemcpyAsync( &headers[i], scenesIn[i], sizeof( SceneHeader ), cudaMemcpyDeviceToHost, s ), "cudaMemcpyAsync" );
		checkCuda( cudaStreamSynchronize( s ), "cudaStreamSynchronize" );
		std::vector<size_t> sizes( numScenes );
		auto				up = []( size_t v ) { return RoundUp( v, size_t( 64 ) ); };
		for ( uint32_t i = 0; i < numScenes; ++i )
			sizes[i] = 64 + up( size_t( headers[i].boxNodeCount ) * sizeof( Box4Node ) ) + up( size_t( headers[i].primNodeCount ) * sizeof( InstanceNode ) ) +
					   up( size_t( headers[i].primCount ) * sizeof( InstanceRecord ) ) + up( size_t( headers[i].frameCount ) * sizeof( SRTFrame ) );
		const std::vector<void*> out = c->allocatePool( sizes );
		for ( uint32_t i = 0; i < numScenes; ++i )
		{
			uint8_t*	dst = static_cast<uint8_t*>( out[i] );
			SceneHeader h	= headers[i];
			size_t		off = 64;
			auto		copy = [&]( const void* src, size_t bytes ) {
				   uint8_t* p = dst + off;
				   if ( bytes ) checkCuda( cudaMemcpyAsync( p, src, bytes, cudaMemcpyDeviceToDevice, s ), "compact copy" );
				   off += up( bytes );
				   return p;
			};
			h.boxNodes	= reinterpret_cast<Box4Node*>( copy( headers[i].boxNodes, size_t( h.boxNodeCount ) * sizeof( Box4Node ) ) );
			h.primNodes = reinterpret_cast<InstanceNode*>( copy( headers[i].primNodes, size_t( h.primNodeCount ) * sizeof( InstanceNode ) ) );
			h.instances = reinterpret_cast<InstanceRecord*>( copy( headers[i].instances, size_t( h.primCount ) * sizeof( InstanceRecord ) ) );
			h.frames	= headers[i].frames ? reinterpret_cast<SRTFrame*>( copy( headers[i].frames, size_t( h.frameCount ) * sizeof( SRTFrame ) ) ) : nullptr;
			h.size		= sizes[i];
			checkCuda( cudaMemcpyAsync( dst, &h, sizeof( SceneHeader ), cudaMemcpyHostToDevice, s ), "write header" );
			*scenesOut[i] = reinterpret_cast<hiprtScene>( dst );
		}
		checkCuda( cudaStreamSynchronize( s ), "cudaStreamSynchronize" );
		for ( uint32_t i = 0; i < numScenes; ++i )
			c->releaseFromPool( scenesIn[i] );
	} );
}

// ------------------------------------------------------------------------------------------------
hiprtError hiprtCreateFuncTable( hiprtContext context, uint32_t numGeomTypes, uint32_t numRayTypes, hiprtFuncTable& funcTableOut )
{
	if ( !context ) return hiprtErrorInvalidParameter;
	return guarded( context, [&] {
		// device layout: header followed by numGeomTypes*numRayTypes data sets (Context.cpp:714-733)
		const size_t bytes = sizeof( hiprtFuncTableHeader ) + size_t( numGeomTypes ) * numRayTypes * sizeof( hiprtFuncDataSet );
		uint8_t*	 p	   = nullptr;
		checkCuda( cudaMalloc( reinterpret_cast<void**>( &p ), bytes ), "cudaMalloc" );
		checkCuda( cudaMemset( p, 0, bytes ), "cudaMemset" );
		hiprtFuncTableHeader h{ numGeomTypes, numRayTypes, reinterpret_cast<hiprtFuncDataSet*>( p + sizeof( hiprtFuncTableHeader ) ) };
		checkCuda( cudaMemcpy( p, &h, sizeof( h ), cudaMemcpyHostToDevice ), "cudaMemcpy" );
		funcTableOut = reinterpret_cast<hiprtFuncTable>( p );
	} );
}

hiprtError hiprtSetFuncTable( hiprtContext context, hiprtFuncTable funcTable, uint32_t geomType, uint32_t rayType, hiprtFuncDataSet set )
{
	if ( !context || !funcTable ) return hiprtErrorInvalidParameter;
	return guarded( context, [&] {
		hiprtFuncTableHeader h;
		checkCuda( cudaMemcpy( &h, funcTable, sizeof( h ), cudaMemcpyDeviceToHost ), "cudaMemcpy" );
		if ( geomType >= h.numGeomTypes || rayType >= h.numRayTypes ) throw std::runtime_error( "hiprtSetFuncTable: index out of range" );
		checkCuda( cudaMemcpy( &h.funcDataSets[h.numGeomTypes * rayType + geomType], &set, sizeof( set ), cudaMemcpyHostToDevice ), "cudaMemcpy" );
	} );
}

hiprtError hiprtDestroyFuncTable( hiprtContext context, hiprtFuncTable funcTable )
{
	if ( !context || !funcTable ) return hiprtErrorInvalidParameter;
	return guarded( context, [&] { checkCuda( cudaFree( funcTable ), "cudaFree" ); } );
}

// ------------------------------------------------------------------------------------------------
hiprtError hiprtCreateGlobalStackBuffer( hiprtContext context, const hiprtGlobalStackBufferInput& input, hiprtGlobalStackBuffer& stackBufferOut )
{
	if ( !context ) return hiprtErrorInvalidParameter;
	return guarded( context, [&] {
		// sizes as in the reference (Context.cpp:752-779); layout documented in hiprt_device.h
		Context*	 c		   = ctx( context );
		const size_t entrySize = input.entryType == hiprtStackEntryTypeInstance ? sizeof( hiprtInstanceStackEntry ) : sizeof( uint32_t );
		hiprtGlobalStackBuffer b{ input.stackSize, 0, nullptr };
		size_t				   bytes = 0, zeroBytes = 0;
		if ( input.type == hiprtStackTypeDynamic )
		{
			b.stackCount = c->info().smCount * c->maxThreadsPerSm();
			zeroBytes	 = size_t( b.stackCount / 32 ) * sizeof( uint32_t ); // lock table
			bytes		 = zeroBytes + size_t( b.stackCount ) * input.stackSize * entrySize;
		}
		else
		{
			b.stackCount = input.threadCount;
			bytes		 = size_t( input.stackSize ) * input.threadCount * entrySize;
		}
		cudaError_t e = cudaMalloc( &b.stackData, std::max<size_t>( bytes, 64 ) );
		if ( e == cudaErrorMemoryAllocation ) throw OutOfDeviceMemory( "cudaMalloc: out of device memory" );
		checkCuda( e, "cudaMalloc" );
		if ( zeroBytes ) checkCuda( cudaMemset( b.stackData, 0, zeroBytes ), "cudaMemset" );
		stackBufferOut = b;
	} );
}

hiprtError hiprtDestroyGlobalStackBuffer( hiprtContext context, hiprtGlobalStackBuffer stackBuffer )
{
	if ( !context ) return hiprtErrorInvalidParameter;
	return guarded( context, [&] { checkCuda( cudaFree( stackBuffer.stackData ), "cudaFree" ); } );
}

// ------------------------------------------------------------------------------------------------
// Serialisation: the storage buffer with the header's pointers rebased to offsets
// (reference: Context.cpp:787-853).
hiprtError hiprtSaveGeometry( hiprtContext context, hiprtGeometry geometry, const char* filename )
{
	if ( !context || !geometry || !filename ) return hiprtErrorInvalidParameter;
	return guarded( context, [&] {
		GeomHeader h;
		checkCuda( cudaMemcpy( &h, geometry, sizeof( h ), cudaMemcpyDeviceToHost ), "cudaMemcpy" );
		std::vector<uint8_t> buf( h.size );
		checkCuda( cudaMemcpy( buf.data(), geometry, h.size, cudaMemcpyDeviceToHost ), "cudaMemcpy" );
		const uintptr_t base = reinterpret_cast<uintptr_t>( geometry );
		h.boxNodes			 = reinterpret_cast<Box4Node*>( reinterpret_cast<uintptr_t>( h.boxNodes ) - base );
		h.primNodes			 = reinterpret_cast<void*>( reinterpret_cast<uintptr_t>( h.primNodes ) - base );
		memcpy( buf.data(), &h, sizeof( h ) );
		std::ofstream f( filename, std::ios::binary );
		if ( !f ) throw std::runtime_error( std::string( "cannot open " ) + filename );
		f.write( reinterpret_cast<const char*>( buf.data() ), std::streamsize( buf.size() ) );
	} );
}

hiprtError hiprtLoadGeometry( hiprtContext context, hiprtGeometry& geometryOut, const char* filename )
{
	if ( !context || !filename ) return hiprtErrorInvalidParameter;
	return guarded( context, [&] {
		std::ifstream f( filename, std::ios::binary );
		if ( !f ) throw std::runtime_error( std::string( "cannot open " ) + filename );
		GeomHeader h;
		f.read( reinterpret_cast<char*>( &h ), sizeof( h ) );
		if ( !f || h.rtip != 0 ) throw std::runtime_error( "not a geometry saved by an RTIP-0 (software traversal) build" );
		std::vector<uint8_t> buf( h.size );
		f.seekg( 0 );
		f.read( reinterpret_cast<char*>( buf.data() ), std::streamsize( buf.size() ) );
		if ( !f ) throw std::runtime_error( "truncated geometry file" );
		// (no Q region: a file may hold an imported hierarchy whose leaf boxes are not the boxes of the vertices -- qnode.h)
		const std::vector<void*> out = ctx( context )->allocatePool( { size_t( h.size ) } );
		const uintptr_t			 base = reinterpret_cast<uintptr_t>( out[0] );
		h.boxNodes					  = reinterpret_cast<Box4Node*>( reinterpret_cast<uintptr_t>( h.boxNodes ) + base );
		h.primNodes					  = reinterpret_cast<void*>( reinterpret_cast<uintptr_t>( h.primNodes ) + base );
		memcpy( buf.data(), &h, sizeof( h ) );
		checkCuda( cudaMemcpy( out[0], buf.data(), buf.size(), cudaMemcpyHostToDevice ), "cudaMemcpy" );
		geometryOut = reinterpret_cast<hiprtGeometry>( out[0] );
		ctx( context )->noteGeometryKind( out[0], h.geomType & 1u );
	} );
}

hiprtError hiprtSaveScene( hiprtContext context, hiprtScene, const char* ) { return context ? hiprtErrorNotImplemented : hiprtErrorInvalidParameter; }
hiprtError hiprtLoadScene( hiprtContext context, hiprtScene&, const char* ) { return context ? hiprtErrorNotImplemented : hiprtErrorInvalidParameter; }

// ------------------------------------------------------------------------------------------------
static void exportRootBox( const Box4Node* deviceRoot, hiprtFloat3& lo, hiprtFloat3& hi )
{
	Box4Node root;
	checkCuda( cudaMemcpy( &root, deviceRoot, sizeof( root ), cudaMemcpyDeviceToHost ), "cudaMemcpy" );
	lo = { FltMax, FltMax, FltMax };
	hi = { -FltMax, -FltMax, -FltMax };
	for ( int s = 0; s < 4; ++s )
	{
		if ( root.child[s] == InvalidValue ) continue;
		lo = { std::min( lo.x, root.box[s].lo.x ), std::min( lo.y, root.box[s].lo.y ), std::min( lo.z, root.box[s].lo.z ) };
		hi = { std::max( hi.x, root.box[s].hi.x ), std::max( hi.y, root.box[s].hi.y ), std::max( hi.z, root.box[s].hi.z ) };
	}
}

hiprtError hiprtExportGeometryAabb( hiprtContext context, hiprtGeometry geometry, hiprtFloat3& aabbMinOut, hiprtFloat3& aabbMaxOut )
{
	if ( !context || !geometry ) return hiprtErrorInvalidParameter;
	return guarded( context, [&] {
		GeomHeader h;
		checkCuda( cudaMemcpy( &h, geometry, sizeof( h ), cudaMemcpyDeviceToHost ), "cudaMemcpy" );
		exportRootBox( h.boxNodes, aabbMinOut, aabbMaxOut );
	} );
}

hiprtError hiprtExportSceneAabb( hiprtContext context, hiprtScene scene, hiprtFloat3& aabbMinOut, hiprtFloat3& aabbMaxOut )
{
	if ( !context || !scene ) return hiprtErrorInvalidParameter;
	return guarded( context, [&] {
		SceneHeader h;
		checkCuda( cudaMemcpy( &h, scene, sizeof( h ), cudaMemcpyDeviceToHost ), "cudaMemcpy" );
		exportRootBox( h.boxNodes, aabbMinOut, aabbMaxOut );
	} );
}

// ------------------------------------------------------------------------------------------------
hiprtError hiprtBuildTraceKernels(
	hiprtContext	  context,
	uint32_t		  numFunctions,
	const char**	  funcNames,
	const char*		  src,
	const char*		  moduleName,
	uint32_t		  numHeaders,
	const char**	  headers,
	const char**	  includeNames,
	uint32_t		  numOptions,
	const char**	  options,
	uint32_t		  numGeomTypes,
	uint32_t		  numRayTypes,
	hiprtFuncNameSet* funcNameSets,
	hiprtApiFunction* functionsOut,
	hiprtApiModule*	  moduleOut,
	bool			  cache )
{
	if ( !context || numFunctions == 0 || funcNames == nullptr || src == nullptr || functionsOut == nullptr ) return hiprtErrorInvalidParameter;
	(void)cache; // compiled modules are cached in memory by source hash
	return guarded( context, [&] {
		Context* c = ctx( context );
		compiler::buildTraceKernels(
			c->modules, c->ccMajor(), c->ccMinor(), numFunctions, funcNames, src, moduleName ? moduleName : "hiprt_user_module", numHeaders, headers,
			includeNames, numOptions, options, numGeomTypes, numRayTypes, funcNameSets, functionsOut, moduleOut );
	} );
}

hiprtError hiprtBuildTraceKernelsFromBitcode(
	hiprtContext	  context,
	uint32_t		  numFunctions,
	const char**	  funcNames,
	const char*		  moduleName,
	const char*		  bitcodeBinary,
	size_t			  bitcodeBinarySize,
	uint32_t		  numGeomTypes,
	uint32_t		  numRayTypes,
	hiprtFuncNameSet* funcNameSets,
	hiprtApiFunction* functionsOut,
	bool			  cache )
{
	if ( !context || numFunctions == 0 || funcNames == nullptr || bitcodeBinary == nullptr || bitcodeBinarySize == 0 || functionsOut == nullptr )
		return hiprtErrorInvalidParameter;
	(void)cache; // linked modules are cached in memory
	return guarded( context, [&] {
		Context* c = ctx( context );
		compiler::buildTraceKernelsFromBitcode(
			c->modules, c->ccMajor(), c->ccMinor(), numFunctions, funcNames, moduleName ? moduleName : "hiprt_user_module", bitcodeBinary, bitcodeBinarySize,
			numGeomTypes, numRayTypes, funcNameSets, functionsOut );
	} );
}

hiprtError hiprtSetCacheDirPath( hiprtContext context, const char* path )
{
	if ( !context || !path ) return hiprtErrorInvalidParameter;
	ctx( context )->cacheDir = path;
	return hiprtSuccess;
}

hiprtError hiprtSetLogLevel( hiprtContext context, hiprtLogLevel level )
{
	if ( !context ) return hiprtErrorInvalidParameter;
	ctx( context )->setLogLevel( level );
	return hiprtSuccess;
}

// ================================================================================================
// hiprt_b200 additions
hiprtError hiprtB200TraceGeometry(
	hiprtContext context, hiprtGeometry geometry, const hiprtRay* rays, uint32_t numRays, hiprtHit* hits, hiprtTraversalType type, hiprtApiStream stream )
{
	if ( !context || !geometry || ( numRays && ( !rays || !hits ) ) ) return hiprtErrorInvalidParameter;
	return guarded( context, [&] {
		const void* q		 = nullptr;
		const bool	triangles = ctx( context )->geometryHasTriangles( geometry, static_cast<cudaStream_t>( stream ), &q );
		traceGeometry(
			geometry, triangles, rays, numRays, hits, type, ctx( context )->workCounterFor( static_cast<cudaStream_t>( stream ) ), ctx( context )->info().smCount, static_cast<cudaStream_t>( stream ),
			hiprtB200HitFormatFull, q );
		checkCuda( cudaGetLastError(), "trace launch" );
	} );
}

hiprtError hiprtB200TraceScene(
	hiprtContext	   context,
	hiprtScene		   scene,
	const hiprtRay*	   rays,
	uint32_t		   numRays,
	hiprtHit*		   hits,
	hiprtTraversalType type,
	hiprtRayMask	   mask,
	hiprtApiStream	   stream )
{
	if ( !context || !scene || ( numRays && ( !rays || !hits ) ) ) return hiprtErrorInvalidParameter;
	return guarded( context, [&] {
		traceScene( scene, rays, numRays, hits, type, mask, ctx( context )->workCounterFor( static_cast<cudaStream_t>( stream ) ), ctx( context )->info().smCount, static_cast<cudaStream_t>( stream ) );
		checkCuda( cudaGetLastError(), "trace launch" );
	} );
}

hiprtError hiprtB200TraceSceneAtTime(
	hiprtContext	   context,
	hiprtScene		   scene,
	const hiprtRay*	   rays,
	const float*	   times,
	float			   time,
	uint32_t		   numRays,
	hiprtHit*		   hits,
	hiprtTraversalType type,
	hiprtRayMask	   mask,
	hiprtApiStream	   stream )
{
	if ( !context || !scene || ( numRays && ( !rays || !hits ) ) ) return hiprtErrorInvalidParameter;
	return guarded( context, [&] {
		traceScene(
			scene, rays, numRays, hits, type, mask, ctx( context )->workCounterFor( static_cast<cudaStream_t>( stream ) ), ctx( context )->info().smCount, static_cast<cudaStream_t>( stream ), times,
			time );
		checkCuda( cudaGetLastError(), "trace launch" );
	} );
}

hiprtError hiprtB200TraceGeometryHost(
	hiprtContext context, hiprtGeometry geometry, const hiprtRay* hostRays, uint32_t numRays, hiprtHit* hostHits, hiprtTraversalType type, hiprtApiStream stream )
{
	if ( !context || !geometry || numRays == 0 || !hostRays || !hostHits ) return hiprtErrorInvalidParameter;
	return guarded( context, [&] {
		const void* q		 = nullptr;
		const bool	triangles = ctx( context )->geometryHasTriangles( geometry, static_cast<cudaStream_t>( stream ), &q );
		ctx( context )->hostPipeline.run(
			geometry, triangles, nullptr, hostRays, numRays, hostHits, type, hiprtFullRayMask, static_cast<cudaStream_t>( stream ), ctx( context )->info().smCount,
			hiprtB200HitFormatFull, q );
	} );
}

hiprtError hiprtB200TraceSceneHost(
	hiprtContext	   context,
	hiprtScene		   scene,
	const hiprtRay*	   hostRays,
	uint32_t		   numRays,
	hiprtHit*		   hostHits,
	hiprtTraversalType type,
	hiprtRayMask	   mask,
	hiprtApiStream	   stream )
{
	if ( !context || !scene || numRays == 0 || !hostRays || !hostHits ) return hiprtErrorInvalidParameter;
	return guarded( context, [&] { ctx( context )->hostPipeline.run( nullptr, false, scene, hostRays, numRays, hostHits, type, mask, static_cast<cudaStream_t>( stream ), ctx( context )->info().smCount ); } );
}

// ---- compact output records, multi-GPU plumbing (include/hiprt/hiprt_b200.h) ------------------------------------------------
size_t hiprtB200HitFormatSize( hiprtB200HitFormat format )
{
	return format == hiprtB200HitFormatFull ? sizeof( hiprtHit ) : ( format == hiprtB200HitFormatCompact ? 16 : 4 );
}

hiprtError hiprtB200TraceGeometryFormat(
	hiprtContext context, hiprtGeometry geometry, const hiprtRay* rays, uint32_t numRays, void* records, hiprtB200HitFormat format, hiprtTraversalType type,
	hiprtApiStream stream )
{
	if ( !context || !geometry || ( numRays && ( !rays || !records ) ) || uint32_t( format ) > hiprtB200HitFormatPrimID ) return hiprtErrorInvalidParameter;
	const void* q		 = nullptr;
	const bool	triangles = ctx( context )->geometryHasTriangles( geometry, static_cast<cudaStream_t>( stream ), &q );
	if ( !triangles && format != hiprtB200HitFormatFull ) return hiprtErrorInvalidParameter;
	return guarded( context, [&] {
		traceGeometry(
			geometry, triangles, rays, numRays, records, type, ctx( context )->workCounterFor( static_cast<cudaStream_t>( stream ) ), ctx( context )->info().smCount, static_cast<cudaStream_t>( stream ),
			uint32_t( format ), q );
		checkCuda( cudaGetLastError(), "trace launch" );
	} );
}

hiprtError hiprtB200TraceGeometryIndirect(
	hiprtContext context, hiprtGeometry geometry, const hiprtRay* rays, uint32_t maxRays, const uint32_t* numRaysDevice, void* records, hiprtB200HitFormat format,
	hiprtTraversalType type, hiprtApiStream stream )
{
	if ( !context || !geometry || !numRaysDevice || ( maxRays && ( !rays || !records ) ) || uint32_t( format ) > hiprtB200HitFormatPrimID ) return hiprtErrorInvalidParameter;
	const void* q		 = nullptr;
	const bool	triangles = ctx( context )->geometryHasTriangles( geometry, static_cast<cudaStream_t>( stream ), &q );
	if ( !triangles ) return hiprtErrorInvalidParameter;
	return guarded( context, [&] {
		traceGeometry(
			geometry, true, rays, maxRays, records, type, ctx( context )->workCounterFor( static_cast<cudaStream_t>( stream ) ), ctx( context )->info().smCount, static_cast<cudaStream_t>( stream ),
			uint32_t( format ), q, numRaysDevice );
		checkCuda( cudaGetLastError(), "trace launch" );
	} );
}

hiprtError hiprtB200TraceGeometryFormatHost(
	hiprtContext context, hiprtGeometry geometry, const hiprtRay* hostRays, uint32_t numRays, void* hostRecords, hiprtB200HitFormat format, hiprtTraversalType type,
	hiprtApiStream stream )
{
	if ( !context || !geometry || numRays == 0 || !hostRays || !hostRecords || uint32_t( format ) > hiprtB200HitFormatPrimID ) return hiprtErrorInvalidParameter;
	const void* q		 = nullptr;
	const bool	triangles = ctx( context )->geometryHasTriangles( geometry, static_cast<cudaStream_t>( stream ), &q );
	if ( !triangles && format != hiprtB200HitFormatFull ) return hiprtErrorInvalidParameter;
	return guarded( context, [&] {
		ctx( context )->hostPipeline.run(
			geometry, triangles, nullptr, hostRays, numRays, hostRecords, type, hiprtFullRayMask, static_cast<cudaStream_t>( stream ), ctx( context )->info().smCount,
			uint32_t( format ), q );
	} );
}

hiprtError hiprtB200TraceGeometryStats(
	hiprtContext context, hiprtGeometry geometry, const hiprtRay* rays, uint32_t numRays, hiprtHit* hits, hiprtTraversalType type, uint64_t* visitCounts,
	hiprtApiStream stream )
{
	if ( !context || !geometry || !visitCounts || ( numRays && ( !rays || !hits ) ) ) return hiprtErrorInvalidParameter;
	const void* q = nullptr;
	if ( !ctx( context )->geometryHasTriangles( geometry, static_cast<cudaStream_t>( stream ), &q ) ) return hiprtErrorInvalidParameter;
	return guarded( context, [&] {
		traceGeometryStats(
			geometry, rays, numRays, hits, type, ctx( context )->workCounterFor( static_cast<cudaStream_t>( stream ) ), ctx( context )->info().smCount, static_cast<cudaStream_t>( stream ),
			reinterpret_cast<unsigned long long*>( visitCounts ), q );
		checkCuda( cudaGetLastError(), "trace launch" );
	} );
}

hiprtError hiprtB200GetGeometrySize( hiprtContext context, hiprtGeometry geometry, size_t* sizeOut )
{
	if ( !context || !geometry || !sizeOut ) return hiprtErrorInvalidParameter;
	return guarded( context, [&] {
		GeomHeader h;
		checkCuda( cudaMemcpy( &h, geometry, sizeof( h ), cudaMemcpyDeviceToHost ), "cudaMemcpy" );
		*sizeOut = h.size;
	} );
}

hiprtError hiprtB200AllocateGeometryBlob( hiprtContext context, size_t size, hiprtGeometry* geometryOut )
{
	if ( !context || size < 64 || !geometryOut ) return hiprtErrorInvalidParameter;
	// room for the Q region of the structure that will arrive (box nodes are at most the whole blob: 64 B per 128 B of them)
	return guarded( context, [&] {
		const size_t publicSize = ( size + 63 ) & ~size_t( 63 );
		void*		 p			= ctx( context )->allocatePool( { publicSize + qRegionSize( size / sizeof( Box4Node ) ) } )[0];
		ctx( context )->noteGeometryKind( p, false, nullptr, publicSize );
		*geometryOut = reinterpret_cast<hiprtGeometry>( p );
	} );
}

hiprtError hiprtB200RelocateGeometry( hiprtContext context, hiprtGeometry geometry, hiprtApiStream stream )
{
	if ( !context || !geometry ) return hiprtErrorInvalidParameter;
	return guarded( context, [&] {
		relocateGeometry( geometry, static_cast<cudaStream_t>( stream ) );
		checkCuda( cudaGetLastError(), "relocate launch" );
		// a blob from hiprtB200AllocateGeometryBlob has room for a Q region behind it: the sender's was not part of the copy,
		// regenerate it here (a structure that is not a triangle mesh marks the region invalid on the device)
		const auto info = ctx( context )->geometryInfo( geometry );
		if ( info.blobSize )
		{
			uint32_t head[12];
			checkCuda( cudaMemcpyAsync( head, geometry, sizeof( head ), cudaMemcpyDeviceToHost, static_cast<cudaStream_t>( stream ) ), "cudaMemcpyAsync" );
			checkCuda( cudaStreamSynchronize( static_cast<cudaStream_t>( stream ) ), "cudaStreamSynchronize" );
			const GeomHeader* h = reinterpret_cast<const GeomHeader*>( head );
			const bool		  triangles = h->geomType & 1u;
			void*			  q = ( triangles && h->boxNodeCount <= info.blobSize / sizeof( Box4Node ) ) ? reinterpret_cast<uint8_t*>( geometry ) + info.blobSize : nullptr;
			compressBoxNodes( reinterpret_cast<const GeomHeader*>( geometry ), q, h->boxNodeCount, static_cast<cudaStream_t>( stream ) );
			ctx( context )->noteGeometryKind( geometry, triangles, q, info.blobSize );
		}
	} );
}

hiprtError hiprtB200SharedAlloc( hiprtContext context, size_t bytes, void** devicePtrOut, void* handleOut )
{
	if ( !context || bytes == 0 || !devicePtrOut || !handleOut ) return hiprtErrorInvalidParameter;
	static_assert( sizeof( cudaIpcMemHandle_t ) == 64, "the IPC handle is documented as 64 opaque bytes" );
	return guarded( context, [&] {
		void*		p = nullptr;
		cudaError_t e = cudaMalloc( &p, bytes );
		if ( e == cudaErrorMemoryAllocation ) throw OutOfDeviceMemory( "cudaMalloc: out of device memory" );
		checkCuda( e, "cudaMalloc" );
		checkCuda( cudaMemset( p, 0, bytes ), "cudaMemset" );
		cudaIpcMemHandle_t h;
		e = cudaIpcGetMemHandle( &h, p );
		if ( e != cudaSuccess )
		{
			cudaFree( p );
			checkCuda( e, "cudaIpcGetMemHandle" );
		}
		memcpy( handleOut, &h, sizeof( h ) );
		*devicePtrOut = p;
	} );
}

hiprtError hiprtB200SharedFree( hiprtContext context, void* devicePtr )
{
	if ( !context || !devicePtr ) return hiprtErrorInvalidParameter;
	return guarded( context, [&] { checkCuda( cudaFree( devicePtr ), "cudaFree" ); } );
}

hiprtError hiprtB200SharedOpen( hiprtContext context, const void* handle, void** devicePtrOut )
{
	if ( !context || !handle || !devicePtrOut ) return hiprtErrorInvalidParameter;
	return guarded( context, [&] {
		cudaIpcMemHandle_t h;
		memcpy( &h, handle, sizeof( h ) );
		void* p = nullptr;
		checkCuda( cudaIpcOpenMemHandle( &p, h, cudaIpcMemLazyEnablePeerAccess ), "cudaIpcOpenMemHandle" );
		*devicePtrOut = p;
	} );
}

hiprtError hiprtB200SharedClose( hiprtContext context, void* devicePtr )
{
	if ( !context || !devicePtr ) return hiprtErrorInvalidParameter;
	return guarded( context, [&] { checkCuda( cudaIpcCloseMemHandle( devicePtr ), "cudaIpcCloseMemHandle" ); } );
}

hiprtError hiprtB200ResolveAoFrame(
	hiprtContext context, const void* primaryRecords, hiprtB200HitFormat primaryFormat, const uint32_t* aoPrimIDs, const uint32_t* firstAoRay,
	const uint32_t* pixelIndex, uint32_t numPixels, uint32_t spp, float* framebuffer, uint32_t* doneFlag, uint32_t doneValue, uint32_t* blockCounter,
	hiprtApiStream stream )
{
	if ( !context || uint32_t( primaryFormat ) > hiprtB200HitFormatPrimID || spp == 0 || !framebuffer || ( doneFlag && !blockCounter ) ) return hiprtErrorInvalidParameter;
	if ( numPixels && ( !primaryRecords || !aoPrimIDs || !firstAoRay || !pixelIndex ) ) return hiprtErrorInvalidParameter;
	return guarded( context, [&] {
		const uint32_t stride = uint32_t( hiprtB200HitFormatSize( primaryFormat ) );
		const uint32_t offset = primaryFormat == hiprtB200HitFormatFull ? 16u /* offsetof( hiprtHit, primID ) */ : 0u;
		resolveAoFrame(
			primaryRecords, stride, offset, aoPrimIDs, firstAoRay, pixelIndex, numPixels, spp, framebuffer, doneFlag, doneValue, blockCounter,
			ctx( context )->info().smCount, static_cast<cudaStream_t>( stream ) );
		checkCuda( cudaGetLastError(), "resolve launch" );
	} );
}

hiprtError hiprtB200WaitFlags(
	hiprtContext context, const uint32_t* flags, uint32_t count, uint32_t strideWords, uint32_t value, uint32_t* timeoutOut, hiprtApiStream stream )
{
	if ( !context || !flags || count == 0 || count > 1024 || strideWords == 0 || ( value == 0 && !timeoutOut ) ) return hiprtErrorInvalidParameter;
	return guarded( context, [&] {
		waitFlags( flags, count, strideWords, value, timeoutOut, static_cast<cudaStream_t>( stream ) );
		checkCuda( cudaGetLastError(), "wait launch" );
	} );
}

hiprtError hiprtB200GetTempLayout( size_t primCount, size_t pairedTriangleCount, size_t* offsetsOut, uint32_t numOffsets )
{
	if ( !offsetsOut || numOffsets < 14 ) return hiprtErrorInvalidParameter;
	const TempLayout L = tempLayout( primCount, pairedTriangleCount );
	const size_t	 v[14] = { L.control, L.pairs,	  L.pairScan,	L.primRefs,	  L.keys[0], L.vals[0], L.keys[1],
							   L.vals[1], L.sortScratch, L.sortedRefs, L.scratch, L.flags,	 L.taskQueue, L.total };
	memcpy( offsetsOut, v, sizeof( v ) );
	return hiprtSuccess;
}

hiprtError hiprtB200GetHighQualityTempLayout( size_t primCount, size_t pairedTriangleCount, int spatialSplits, size_t* offsetsOut, uint32_t numOffsets )
{
	if ( !offsetsOut || numOffsets < 8 ) return hiprtErrorInvalidParameter;
	highQualityTempLayout( primCount, pairedTriangleCount, spatialSplits != 0, offsetsOut );
	return hiprtSuccess;
}

uint32_t hiprtB200GetDeviceHeaders( const char*** namesOut, const char*** sourcesOut ) { return compiler::deviceHeaders( namesOut, sourcesOut ); }

hiprtError hiprtB200LinkTraceKernels(
	int				  ccMajor,
	int				  ccMinor,
	const char*		  moduleName,
	const char*		  bitcodeBinary,
	size_t			  bitcodeBinarySize,
	uint32_t		  numGeomTypes,
	uint32_t		  numRayTypes,
	hiprtFuncNameSet* funcNameSets,
	void*			  cubinOut,
	size_t*			  cubinSizeInOut,
	char*			  logOut,
	size_t			  logCapacity )
{
	if ( !bitcodeBinary || bitcodeBinarySize == 0 || !cubinSizeInOut ) return hiprtErrorInvalidParameter;
	try
	{
		const std::vector<char> cubin = compiler::linkTraceKernels(
			ccMajor, ccMinor, bitcodeBinary, bitcodeBinarySize, moduleName ? moduleName : "hiprt_user_module", numGeomTypes, numRayTypes, funcNameSets );
		const size_t capacity = *cubinSizeInOut;
		*cubinSizeInOut		  = cubin.size();
		if ( cubinOut == nullptr ) return hiprtSuccess; // size query
		if ( capacity < cubin.size() ) return hiprtErrorInvalidParameter;
		memcpy( cubinOut, cubin.data(), cubin.size() );
		return hiprtSuccess;
	}
	catch ( const std::exception& e )
	{
		if ( logOut && logCapacity ) snprintf( logOut, logCapacity, "%s", e.what() );
		return hiprtErrorInternal;
	}
}

const char* hiprtB200GetLastErrorString( hiprtContext context ) { return context ? ctx( context )->lastError() : ""; }

hiprtError hiprtB200EnableBuildProfile( hiprtContext context, int enable )
{
	if ( !context ) return hiprtErrorInvalidParameter;
	return guarded( context, [&] { ctx( context )->enableProfile( enable != 0 ); } );
}

hiprtError hiprtB200GetBuildProfile( hiprtContext context, uint32_t* phaseCountOut, const char** phaseNamesOut, float* phaseMsOut, uint32_t maxPhases )
{
	if ( !context || !phaseCountOut ) return hiprtErrorInvalidParameter;
	return guarded( context, [&] {
		BuildProfile& p = ctx( context )->profile();
		uint32_t	  n = 0;
		if ( p.count > 1 )
		{
			checkCuda( cudaEventSynchronize( p.events[p.count - 1] ), "cudaEventSynchronize" );
			for ( int i = 1; i < p.count && n < maxPhases; ++i, ++n )
			{
				float ms = 0.0f;
				checkCuda( cudaEventElapsedTime( &ms, p.events[i - 1], p.events[i] ), "cudaEventElapsedTime" );
				if ( phaseNamesOut ) phaseNamesOut[n] = p.names[i];
				if ( phaseMsOut ) phaseMsOut[n] = ms;
			}
		}
		*phaseCountOut = n;
	} );
}

hiprtError hiprtB200CopyToHost( hiprtContext context, void* hostDst, const void* deviceSrc, size_t bytes )
{
	if ( !context || !hostDst || !deviceSrc ) return hiprtErrorInvalidParameter;
	return guarded( context, [&] {
		checkCuda( cudaDeviceSynchronize(), "cudaDeviceSynchronize" );
		checkCuda( cudaMemcpy( hostDst, deviceSrc, bytes, cudaMemcpyDeviceToHost ), "cudaMemcpy" );
	} );
}

} // extern "C"
