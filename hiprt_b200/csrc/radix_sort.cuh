// Onesweep least-significant-digit radix sort of (key u32, value u32) pairs.
//
// What it replaces: hiprt/impl/RadixSort.cpp:38-70 -> Orochi's ParallelPrimitives onesweep
// (contrib/Orochi/ParallelPrimitives/RadixSortKernels.h:347-707, host RadixSort.cpp:237-310).
// Contract: stable, ascending, all 32 key bits (the LBVH relies on stability: equal Morton
// codes keep ascending primitive index).
//
// B200 design: the digit histograms of ALL passes come from the kernel that produces the keys
// (one read of the keys saved), each pass is ONE kernel: a tile of TILE keys is ranked with
// warp-level match_any multi-split, the tile's per-digit counts are published with a
// decoupled look-back chain, keys and values are staged through shared memory so that global
// writes of one digit run are contiguous. Compulsory traffic: PASSES x 16 B per pair.
#pragma once
#include "device_util.cuh"

namespace hiprt
{
namespace rsort
{
constexpr int	   Bits		 = 8;
constexpr int	   Radix	 = 1 << Bits;
constexpr int	   Passes	 = 4; // 32-bit keys
constexpr int	   Block	 = 512;
constexpr int	   Warps	 = Block / 32;
constexpr int	   Items	 = 12; // keys per thread
constexpr int	   Tile		 = Block * Items;
constexpr uint32_t FlagAgg	 = 1u << 30; // tile aggregate available
constexpr uint32_t FlagIncl	 = 2u << 30; // inclusive prefix available
constexpr uint32_t ValueMask = ( 1u << 30 ) - 1u;

__host__ __device__ inline uint32_t tileCount( uint32_t n ) { return ( n + Tile - 1 ) / Tile; }

// Scratch layout (uint32 words), zeroed before the first pass:
//   hist[Passes][Radix] | tileCounter[Passes] (+pad to 16) | lookback[Passes][tiles][Radix]
__host__ __device__ inline size_t scratchWords( uint32_t n )
{
	return size_t( Passes ) * Radix + 16 + size_t( Passes ) * tileCount( n ) * Radix;
}

// Accumulate the digit histograms of every pass for one key into shared counters
// (called by the Morton kernel; smemHist has Passes*Radix entries).
__device__ __forceinline__ void histogramKey( uint32_t* smemHist, uint32_t key )
{
#pragma unroll
	for ( int p = 0; p < Passes; ++p )
		atomicAdd( &smemHist[p * Radix + ( ( key >> ( p * Bits ) ) & ( Radix - 1 ) )], 1u );
}

struct __align__( 16 ) Smem
{
	union
	{
		uint32_t warpHist[Warps][Radix]; // ranking phase
		uint2	 stage[Tile];			 // scatter phase: (key, value) by local rank
	};
	uint32_t digitStartLocal[Radix];  // first local rank of each digit inside the tile
	uint32_t digitBaseGlobal[Radix];  // global position of the tile's first element of each digit minus its local start
	uint32_t tile;
};

// One pass: keysIn/valsIn -> keysOut/valsOut ordered by digit `pass`.
__global__ void __launch_bounds__( Block ) onesweepPass(
	const uint32_t* __restrict__ keysIn,
	const uint32_t* __restrict__ valsIn,
	uint32_t* __restrict__ keysOut,
	uint32_t* __restrict__ valsOut,
	const uint32_t* nPtr, // element count in device memory (known only after triangle pairing)
	uint32_t		nMax, // host-side upper bound: fixes the scratch layout and the grid
	uint32_t*		scratch,
	int				pass )
{
	extern __shared__ __align__( 16 ) uint8_t smemRaw[];
	Smem&						  s = *reinterpret_cast<Smem*>( smemRaw );
	const uint32_t				  tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
	const uint32_t				  n			  = *nPtr;
	const uint32_t				  tiles		  = tileCount( nMax );
	uint32_t*					  hist		  = scratch + pass * Radix;
	uint32_t*					  tileCounter = scratch + Passes * Radix + pass;
	uint32_t*					  lookback	  = scratch + Passes * Radix + 16 + size_t( pass ) * tiles * Radix;
	const int					  shift		  = pass * Bits;

	if ( tid == 0 ) s.tile = atomicAdd( tileCounter, 1u );
	for ( int i = tid; i < Warps * Radix; i += Block )
		( &s.warpHist[0][0] )[i] = 0;
	__syncthreads();
	const uint32_t tile		= s.tile;
	const uint32_t tileBase = tile * Tile;
	if ( tileBase >= n ) return; // tickets are handed out in order: every lower tile is live
	const uint32_t warpBase = tileBase + warp * ( 32 * Items );

	// ---- load (warp-striped: element order inside a warp is item-major, lane-minor)
	uint32_t key[Items], rank[Items];
#pragma unroll
	for ( int i = 0; i < Items; ++i )
	{
		const uint32_t idx = warpBase + i * 32 + lane;
		key[i]			   = idx < n ? keysIn[idx] : 0xFFFFFFFFu; // padding sorts last and is never written
	}

	// ---- rank inside the warp with match_any: the lowest lane of each digit group adds the group's size to the warp's counter of
	// that digit with ONE shared-memory atomic and hands the previous value to its group by shuffle. The atomics of one warp retire
	// in program order (item i before item i + 1: stability), and nothing else orders the items, so the twelve matches, atomics and
	// shuffles pipeline (ncu on the read-modify-write loop this replaces -- LDS, __syncwarp, STS, __syncwarp per item -- showed the
	// short-scoreboard stall on top).
	uint32_t peers[Items];
#pragma unroll
	for ( int i = 0; i < Items; ++i )
		peers[i] = __match_any_sync( FullMask, ( key[i] >> shift ) & ( Radix - 1 ) );
#pragma unroll
	for ( int i = 0; i < Items; ++i )
	{
		const uint32_t d	  = ( key[i] >> shift ) & ( Radix - 1 );
		const uint32_t leader = __ffs( peers[i] ) - 1u;
		uint32_t	   prev	  = 0;
		if ( lane == leader ) prev = atomicAdd( &s.warpHist[warp][d], uint32_t( __popc( peers[i] ) ) );
		rank[i] = __shfl_sync( FullMask, prev, leader ) + __popc( peers[i] & laneMaskLt() );
	}
	__syncthreads();

	// ---- per digit: exclusive scan over warps, tile total, look-back, local digit start
	uint32_t tileTotal = 0; // valid for tid < Radix
	if ( tid < Radix )
	{
		uint32_t sum = 0;
#pragma unroll
		for ( int w = 0; w < Warps; ++w )
		{
			const uint32_t c	= s.warpHist[w][tid];
			s.warpHist[w][tid] = sum;
			sum += c;
		}
		tileTotal = sum;
		// publish the aggregate (tile 0 publishes its inclusive prefix right away)
		volatile uint32_t* lb = lookback + size_t( tile ) * Radix + tid;
		*lb					  = ( tile == 0 ? FlagIncl : FlagAgg ) | tileTotal;
	}
	// exclusive scan of tile totals over digits -> digitStartLocal (block-wide, Radix <= Block)
	{
		uint32_t v = tid < Radix ? tileTotal : 0;
		// warp inclusive scan
		uint32_t x = v;
#pragma unroll
		for ( int o = 1; o < 32; o <<= 1 )
		{
			const uint32_t y = __shfl_up_sync( FullMask, x, o );
			if ( lane >= o ) x += y;
		}
		__shared__ uint32_t warpSums[Warps];
		if ( lane == 31 ) warpSums[warp] = x;
		__syncthreads();
		uint32_t base = 0;
		for ( uint32_t w = 0; w < warp; ++w )
			base += warpSums[w];
		if ( tid < Radix ) s.digitStartLocal[tid] = base + x - v;
	}
	if ( tid < Radix )
	{
		// global start of this digit = sum of the global histogram over smaller digits
		// (recomputed per tile: Radix loads from L2, cheaper than another launch)
		uint32_t before = 0;
		{
			uint32_t x = hist[tid];
			uint32_t incl = x;
#pragma unroll
			for ( int o = 1; o < 32; o <<= 1 )
			{
				const uint32_t y = __shfl_up_sync( FullMask, incl, o );
				if ( lane >= o ) incl += y;
			}
			__shared__ uint32_t histWarpSums[Radix / 32];
			if ( lane == 31 ) histWarpSums[warp] = incl;
			// only the first Radix/32 warps take part; they are whole warps so a named barrier is not needed
			asm volatile( "bar.sync 1, %0;" ::"n"( Radix ) );
			for ( uint32_t w = 0; w < warp; ++w )
				before += histWarpSums[w];
			before += incl - x;
		}
		// decoupled look-back over the preceding tiles
		uint32_t prefix = 0;
		if ( tile > 0 )
		{
			int t = int( tile ) - 1;
			for ( uint32_t spins = 0; spins < ( 1u << 26 ); ++spins ) // bounded: never hang the device
			{
				const uint32_t v = ldVolatile( lookback + size_t( t ) * Radix + tid );
				if ( ( v >> 30 ) == 0 ) continue; // not published yet
				prefix += v & ValueMask;
				if ( v & FlagIncl ) break;
				--t;
			}
			volatile uint32_t* lb = lookback + size_t( tile ) * Radix + tid;
			*lb					  = FlagIncl | ( prefix + tileTotal );
		}
		s.digitBaseGlobal[tid] = before + prefix - s.digitStartLocal[tid]; // destination = this + local rank
	}
	__syncthreads();

	// ---- local rank of every element inside the tile (digit-major)
#pragma unroll
	for ( int i = 0; i < Items; ++i )
	{
		const uint32_t d = ( key[i] >> shift ) & ( Radix - 1 );
		rank[i] += s.digitStartLocal[d] + s.warpHist[warp][d];
	}
	__syncthreads(); // warpHist is dead from here on: the union becomes the staging buffer

	const uint32_t valid = min( uint32_t( Tile ), n - tileBase );
	// keys and values travel together: staged by local rank as 8-byte pairs, written as runs of one digit
#pragma unroll
	for ( int i = 0; i < Items; ++i )
	{
		const uint32_t idx = warpBase + i * 32 + lane;
		s.stage[rank[i]]   = make_uint2( key[i], idx < n ? valsIn[idx] : 0u );
	}
	__syncthreads();
#pragma unroll
	for ( int i = 0; i < Items; ++i )
	{
		const uint32_t r = i * Block + tid;
		if ( r < valid )
		{
			const uint2	   kv  = s.stage[r];
			const uint32_t dst = s.digitBaseGlobal[( kv.x >> shift ) & ( Radix - 1 )] + r;
			keysOut[dst]	   = kv.x;
			valsOut[dst]	   = kv.y;
		}
	}
}
} // namespace rsort
} // namespace hiprt
