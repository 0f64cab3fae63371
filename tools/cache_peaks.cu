// Measured cache-level peaks for the traversal roofline ("reported against a measured L2 bandwidth
// microbenchmark"): the access pattern of a box-node visit -- every lane of a warp fetches its own 128-byte node with four
// 256-bit read-only loads (ld.global.nc.v8.f32, what visitBox4 issues) at an unpredictable index -- against working sets that
// live in L1 (16 KB per block), in L2 (48 MB) and in HBM (4 GB). Independent loads, 8 blocks of 128 threads per SM as the
// trace kernel runs: an upper bound for node bytes per second that no traversal order can exceed.
//   build: make -C hiprt_b200/csrc build/cache_peaks      run: hiprt_b200/csrc/build/cache_peaks > profiles/r02_cache_peaks.json
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <vector>

struct F8
{
	float v[8];
};
__device__ __forceinline__ F8 ldg8( const void* p )
{
	F8 r;
	asm volatile( "ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
				  : "=f"( r.v[0] ), "=f"( r.v[1] ), "=f"( r.v[2] ), "=f"( r.v[3] ), "=f"( r.v[4] ), "=f"( r.v[5] ), "=f"( r.v[6] ), "=f"( r.v[7] )
				  : "l"( p ) );
	return r;
}
__device__ __forceinline__ uint32_t hash( uint32_t x )
{
	x ^= x >> 16, x *= 0x7feb352du, x ^= x >> 15, x *= 0x846ca68bu, x ^= x >> 16;
	return x;
}

// nodesPerWindow nodes of 128 B per window; window = whole buffer (shared == 1) or a private slice per block.
// sameNodePerWarp: all lanes of a warp fetch one node (the coherent-ray case: one wavefront serves the warp).
__global__ void __launch_bounds__( 128 ) gatherKernel( const uint8_t* nodes, uint32_t nodesPerWindow, int sharedWindow, int sameNodePerWarp, uint32_t iters, float* sink )
{
	const uint8_t* base = nodes + ( sharedWindow ? 0 : size_t( blockIdx.x ) * nodesPerWindow * 128 );
	const uint32_t id	= sameNodePerWarp ? ( blockIdx.x * blockDim.x + threadIdx.x ) >> 5 : blockIdx.x * blockDim.x + threadIdx.x;
	float		   acc	= 0.0f;
	uint32_t	   h	= hash( id * 2654435761u + 1u );
	for ( uint32_t it = 0; it < iters; ++it )
	{
		// two independent node fetches per iteration (a warp of the trace kernel has one in flight per lane; 8 blocks per SM hide the rest)
		const uint32_t a = h % nodesPerWindow;
		h				 = hash( h + it );
		const uint32_t b = h % nodesPerWindow;
		h				 = hash( h + 0x9e3779b9u );
		const uint8_t* pa = base + size_t( a ) * 128;
		const uint8_t* pb = base + size_t( b ) * 128;
		const F8	   a0 = ldg8( pa ), a1 = ldg8( pa + 32 ), a2 = ldg8( pa + 64 ), a3 = ldg8( pa + 96 );
		const F8	   b0 = ldg8( pb ), b1 = ldg8( pb + 32 ), b2 = ldg8( pb + 64 ), b3 = ldg8( pb + 96 );
		acc += a0.v[0] + a1.v[1] + a2.v[2] + a3.v[3] + b0.v[4] + b1.v[5] + b2.v[6] + b3.v[7];
	}
	if ( acc == 12345.678f ) *sink = acc;
}

static double run( const uint8_t* nodes, uint32_t nodesPerWindow, int sharedWindow, int sameNode, uint32_t blocks, uint32_t iters, float* sink )
{
	cudaEvent_t e0, e1;
	cudaEventCreate( &e0 ), cudaEventCreate( &e1 );
	gatherKernel<<<blocks, 128>>>( nodes, nodesPerWindow, sharedWindow, sameNode, iters / 4, sink ); // warm-up (fills the caches)
	double best = 0.0;
	for ( int rep = 0; rep < 5; ++rep )
	{
		cudaEventRecord( e0 );
		gatherKernel<<<blocks, 128>>>( nodes, nodesPerWindow, sharedWindow, sameNode, iters, sink );
		cudaEventRecord( e1 );
		cudaEventSynchronize( e1 );
		float ms = 0;
		cudaEventElapsedTime( &ms, e0, e1 );
		const double bytes = double( blocks ) * 128.0 * iters * 2.0 * 128.0; // bytes delivered to lanes
		best			   = std::max( best, bytes / ( ms * 1e-3 ) / 1e9 );
	}
	cudaEventDestroy( e0 ), cudaEventDestroy( e1 );
	return best;
}

int main()
{
	cudaDeviceProp prop{};
	cudaGetDeviceProperties( &prop, 0 );
	const uint32_t sms	  = uint32_t( prop.multiProcessorCount );
	const uint32_t blocks = sms * 8;
	const size_t   big	  = size_t( 4 ) << 30;
	uint8_t*	   nodes  = nullptr;
	float*		   sink	  = nullptr;
	if ( cudaMalloc( &nodes, big ) != cudaSuccess || cudaMalloc( &sink, 4 ) != cudaSuccess )
	{
		fprintf( stderr, "cudaMalloc failed\n" );
		return 1;
	}
	cudaMemset( nodes, 0, big );
	int clockKHz = 0;
	cudaDeviceGetAttribute( &clockKHz, cudaDevAttrClockRate, 0 );
	const double l1	   = run( nodes, 128, 0, 0, blocks, 4096, sink );							  // 16 KB per block, 128 KB per SM: L1 hits
	const double l1Brd = run( nodes, 128, 0, 1, blocks, 4096, sink );							  // same, one node per warp (broadcast)
	const double l2	   = run( nodes, uint32_t( ( size_t( 48 ) << 20 ) / 128 ), 1, 0, blocks, 1024, sink ); // 48 MB: L2 hits, L1 misses
	const double l2Brd = run( nodes, uint32_t( ( size_t( 48 ) << 20 ) / 128 ), 1, 1, blocks, 1024, sink );
	const double hbm   = run( nodes, uint32_t( big / 128 ), 1, 0, blocks, 256, sink );			  // 4 GB: DRAM
	printf(
		"{\n \"gpu\": \"%s\", \"sms\": %u, \"sm_clock_mhz_nominal\": %d,\n"
		" \"pattern\": \"per-lane 128-byte node gather, 4 x ld.global.nc.v8.f32 per node, 2 independent nodes per iteration, %u blocks x 128 threads\",\n"
		" \"l1_resident_gather_gbs\": %.1f, \"l1_resident_gather_bytes_per_clk_per_sm\": %.1f,\n"
		" \"l1_resident_same_node_per_warp_gbs\": %.1f,\n"
		" \"l2_resident_gather_gbs\": %.1f, \"l2_resident_same_node_per_warp_gbs\": %.1f,\n"
		" \"hbm_gather_gbs\": %.1f,\n"
		" \"working_sets\": {\"l1\": \"16 KB per block (8 blocks per SM)\", \"l2\": \"48 MB shared\", \"hbm\": \"4 GB shared\"},\n"
		" \"note\": \"bytes delivered to lanes per second, best of 5 launches after a warm-up launch; same_node_per_warp = all 32 lanes fetch one node\"\n}\n",
		prop.name, sms, clockKHz / 1000, blocks, l1, l1 * 1e9 / ( double( clockKHz ) * 1e3 ) / sms, l1Brd, l2, l2Brd, hbm );
	cudaFree( nodes ), cudaFree( sink );
	return 0;
}
