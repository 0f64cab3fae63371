// PreferHighQualityBuild: top-down binned-SAH builder with spatial splits (SBVH), hand-written for sm_100a.
//
// Reference path this replaces: hiprt/impl/SbvhBuilder.h:174-560 with the kernels
// SetupLeavesAndReferences / ResetBins / BinReferencesObject / FindObjectSplit / BinReferencesSpatial / SplitReferences /
// DistributeReferences (SbvhBuilderKernels.h:65-1077), Bin / Split / Task (SbvhCommon.h), Triangle::split (Triangle.h:132-162),
// constants BvhConfig.h:46-51. The reference runs one level of the tree per host iteration: seven launches and four blocking
// read-backs of counters per level.
//
// B200 design: ONE persistent cooperative launch builds the whole binary tree. The grid (a few CTAs per SM, all resident) walks
// the levels on the device; the phases of a level are separated by a grid barrier (one atomic per CTA) instead of a launch and a
// host round trip, and the loop ends exactly when a level creates no task -- the host never learns the depth.
//   * References MOVE: a level streams the 32-byte reference records of its tasks from one buffer and writes the children's
//     records to the other, segment by segment (the reference keeps the records in place and permutes two index arrays, which
//     turns every later pass into a gather).
//   * Binning is privatised per CTA while a CTA's span of references stays inside one task (the top of the tree): shared-memory
//     bins, flushed once per task change. Below that, tasks are many and the global atomics spread out.
//   * Split search: one sub-warp of 8 / 16 / 32 lanes (the bin count rounded up) per task, prefix / suffix unions by shuffles (the reference: one thread per task
//     looping over bins held in local memory).
//   * Child segments, nodes, tasks and leaf records are allocated in the split phase, so a node is written once, complete.
// The result is the same pair of record arrays every other builder hands to collapsePackKernel (build_kernels.cuh): binary
// nodes {box, child0, child1} and leaf records {clipped box, primitive words}; several leaves may name one primitive.
//
// The arithmetic (bin of a centroid, bin planes, SAH cost, clipping) is spelled with explicit single roundings so that the CPU
// restatement in oracle/hiprt_oracle.cpp (orc_build_sbvh) produces the same tree: tests compare the two leaf for leaf.
#pragma once
#include "builders.h"
#include "prim_sources.cuh"

namespace hiprt
{
namespace sbvh
{
constexpr uint32_t Block   = 256;
constexpr uint32_t MinBins = 8;	 // BvhConfig.h:46
constexpr uint32_t MaxBins = 32; // BvhConfig.h:47
constexpr float	   Beta	   = 1.0e-4f; // overlap of the object split's children, relative to the scene area, that makes a task try spatial splits
constexpr float	   Gamma   = 1.0e-3f; // axes shorter than this fraction of the scene extent are not split spatially
constexpr float	   Epsilon = 1.0e-2f; // centroid bins are scaled by (1 - Epsilon)
constexpr uint32_t SmallTask = 4; // on thread-per-task levels, tasks of at most this many references are split straight from their references
constexpr uint32_t TimelineSlots  = 8;	// per level: start, reset, object bins, object split, spatial bins, demand, split, distribute
constexpr uint32_t TimelineLevels = 128;

// A task = one binary node whose references are still to be partitioned. 128 bytes, two 64-byte halves.
struct alignas( 128 ) Task
{
	float	 box[6];	// union of the node's reference boxes (what the parent's bins measured)
	uint32_t nextTask1; // index of the right child's task on the next level
	uint32_t count;		// references of this node
	float	 box0[6];	// left child
	uint32_t count0;
	uint32_t child0; // node index of the left child: internal node or leaf record
	float	 box1[6]; // right child
	uint32_t count1;
	uint32_t child1;
	uint32_t node;	// this task's binary node
	uint32_t split; // axis (3 = by arrival) | spatial << 2 | tries spatial << 3 | (bin index, or left count of a split by arrival) << 4
	union
	{
		float	 cost;	  // of the object split, until the split is final
		uint32_t arrival; // then: arrival counter of a split by arrival
	};
	uint32_t out0, out1;	   // first position of the child's references in the next buffer, or the child's leaf record
							   // (until the split phase, out0 of a two-reference task collects the smaller first word of its references)
	uint32_t cursor0, cursor1; // references placed so far
	uint32_t nextTask0;
};
static_assert( sizeof( Task ) == 128, "Task is one 128-byte line" );

struct LevelCounts
{
	uint32_t taskCount;
	uint32_t activeRefs;   // references of this level's tasks
	uint32_t spatialTasks; // tasks of this level that try spatial splits
	uint32_t demand;	   // references those splits would duplicate
};

struct State
{
	uint32_t	barrier; // CTAs arrived, all barriers so far
	uint32_t	error;	 // 1: a barrier timed out, 2: an allocation ran past its array
	uint32_t	totalRefs;
	uint32_t	leafCount;
	uint32_t	nodeCount;
	uint32_t	levels;
	uint32_t	pad[2];
	LevelCounts level[2];
};
static_assert( sizeof( State ) == 64, "State is one 64-byte line" );

struct Params
{
	State*		   state;
	BuildControl*  ctl;
	const void*	   primRefs; // {box, words} per primitive, in input order (primBoxesKernel)
	void*		   refs[2];	 // 32 B x maxRefs, ping-pong
	uint32_t*	   refTask[2];
	Task*		   tasks[2]; // maxRefs / 2 each
	float*		   objectBins;
	float*		   spatialBins; // null: object splits only
	void*		   nodes;		// binary nodes out (scratch records)
	void*		   leaves;		// leaf records out
	uint64_t*	   taskQueue;	// collapse queue: slot 0 receives the root task
	uint64_t*	   timeline; // optional: %globaltimer after every grid barrier, TimelineSlots values per level (profiling)
	uint32_t	   serialMin; // levels with at least this many tasks run their per-task phases with one thread per task
	uint32_t	   maxRefs; // capacity of the reference / node / leaf arrays (sized on the host from the primitive count BEFORE pairing)
};

// ---- arithmetic shared with the oracle (single roundings, no contraction) ---------------------------------------------------
__device__ __forceinline__ float axisOf( const DBox& b, bool hi, uint32_t a )
{
	const float3 v = hi ? b.hi : b.lo;
	return a == 0 ? v.x : ( a == 1 ? v.y : v.z );
}
__device__ __forceinline__ void setAxis( float3& v, uint32_t a, float x )
{
	if ( a == 0 )
		v.x = x;
	else if ( a == 1 )
		v.y = x;
	else
		v.z = x;
}
__device__ __forceinline__ bool validBox( const DBox& b ) { return b.lo.x <= b.hi.x && b.lo.y <= b.hi.y && b.lo.z <= b.hi.z; }
__device__ __forceinline__ float areaOrZero( const DBox& b ) { return validBox( b ) ? boxArea( b ) : 0.0f; }
__device__ __forceinline__ DBox intersectBox( const DBox& a, const DBox& b )
{
	return DBox{
		make_float3( fmaxf( a.lo.x, b.lo.x ), fmaxf( a.lo.y, b.lo.y ), fmaxf( a.lo.z, b.lo.z ) ),
		make_float3( fminf( a.hi.x, b.hi.x ), fminf( a.hi.y, b.hi.y ), fminf( a.hi.z, b.hi.z ) ) };
}
// bin of a centroid along one axis: floor( (1 - Epsilon) * bins / extent * (centre - lo) ), clamped; a flat axis bins to 0
__device__ __forceinline__ uint32_t centroidBin( float lo, float hi, float taskLo, float taskHi, uint32_t bins )
{
	const float k = __fmul_rn( 1.0f - Epsilon, __fdiv_rn( float( bins ), __fsub_rn( taskHi, taskLo ) ) );
	const float c = __fmul_rn( __fadd_rn( hi, lo ), 0.5f );
	const float v = __fmul_rn( k, __fsub_rn( c, taskLo ) );
	if ( !( v > 0.0f ) ) return 0u; // negative, zero, NaN (0 * inf on a flat axis)
	if ( v >= float( bins ) ) return bins - 1u;
	return uint32_t( v );
}
__device__ __forceinline__ float binPlane( float taskLo, float binSize, uint32_t i ) { return __fmaf_rn( binSize, float( i + 1u ), taskLo ); }
// First and last spatial bin a box touches along one axis: first = the first bin whose upper plane lies above `lo`, last = the first
// whose upper plane reaches `hi` (the last bin if none does). The planes are non-decreasing in i, so the reference's linear scan over
// the planes (SbvhBuilderKernels.h:273-281) is a lower-bound search: 5 probes instead of 31, the same result.
__device__ __forceinline__ void spatialBinRange( float lo, float hi, float taskLo, float binSize, uint32_t bins, uint32_t& first, uint32_t& last )
{
	uint32_t f = 0, l = 0; // number of planes i in [0, bins - 1) with !( lo < plane_i ) / !( hi <= plane_i )
	for ( uint32_t step = 16u; step > 0u; step >>= 1 )
	{
		if ( f + step < bins && !( lo < binPlane( taskLo, binSize, f + step - 1u ) ) ) f += step;
		if ( l + step < bins && !( hi <= binPlane( taskLo, binSize, l + step - 1u ) ) ) l += step;
	}
	first = min( f, l ), last = l;
}
__device__ __forceinline__ float sahCost( const DBox& l, uint32_t nl, const DBox& r, uint32_t nr )
{
	return __fadd_rn( __fmul_rn( boxArea( l ), float( nl ) ), __fmul_rn( boxArea( r ), float( nr ) ) );
}

// ---- what a primitive looks like on either side of a plane ------------------------------------------------------------------
__device__ __forceinline__ void clipTriangle( const float3 v[3], uint32_t axis, float position, DBox& left, DBox& right )
{
	for ( uint32_t i = 0; i < 3; ++i )
	{
		const float3 a = v[( i + 2u ) % 3u], b = v[i];
		const float	 ap = axis == 0 ? a.x : ( axis == 1 ? a.y : a.z ), bp = axis == 0 ? b.x : ( axis == 1 ? b.y : b.z );
		if ( ap <= position ) growBox( left, a );
		if ( ap >= position ) growBox( right, a );
		if ( ( ap < position && bp > position ) || ( ap > position && bp < position ) )
		{
			const float	 t = fminf( fmaxf( __fdiv_rn( __fsub_rn( position, ap ), __fsub_rn( bp, ap ) ), 0.0f ), 1.0f );
			const float3 p = make_float3(
				__fmaf_rn( t, __fsub_rn( b.x, a.x ), a.x ), __fmaf_rn( t, __fsub_rn( b.y, a.y ), a.y ), __fmaf_rn( t, __fsub_rn( b.z, a.z ), a.z ) );
			growBox( left, p );
			growBox( right, p );
		}
	}
}

// the vertices a reference's primitive is clipped by (fetched once per reference)
template <typename Source>
struct Clipper;

template <>
struct Clipper<TriangleMeshSource>
{
	float3 v[6];
	bool   pair;
	__device__ __forceinline__ Clipper( const TriangleMeshSource& src, uint32_t w0, uint32_t w1 )
	{
		const uint3 i0 = src.indices( w0 );
		v[0] = src.vertex( i0.x ), v[1] = src.vertex( i0.y ), v[2] = src.vertex( i0.z );
		pair = w0 != w1;
		if ( pair )
		{
			const uint3 i1 = src.indices( w1 );
			v[3] = src.vertex( i1.x ), v[4] = src.vertex( i1.y ), v[5] = src.vertex( i1.z );
		}
	}
	__device__ __forceinline__ void split( uint32_t axis, float position, const DBox& box, DBox& left, DBox& right ) const
	{
		left = emptyBox(), right = emptyBox();
		clipTriangle( v, axis, position, left, right );
		if ( pair ) clipTriangle( v + 3, axis, position, left, right );
		setAxis( left.hi, axis, position );
		setAxis( right.lo, axis, position );
		left = intersectBox( left, box ), right = intersectBox( right, box );
	}
};
template <>
struct Clipper<AabbListSource>
{
	__device__ __forceinline__ Clipper( const AabbListSource&, uint32_t, uint32_t ) {}
	__device__ __forceinline__ void split( uint32_t axis, float position, const DBox& box, DBox& left, DBox& right ) const
	{
		left = box, right = box;
		setAxis( left.hi, axis, position );
		setAxis( right.lo, axis, position );
	}
};
template <>
struct Clipper<InstanceSource> // (an instance is cut like a box: the part of its world box on either side of the plane)
{
	__device__ __forceinline__ Clipper( const InstanceSource&, uint32_t, uint32_t ) {}
	__device__ __forceinline__ void split( uint32_t axis, float position, const DBox& box, DBox& left, DBox& right ) const
	{
		left = box, right = box;
		setAxis( left.hi, axis, position );
		setAxis( right.lo, axis, position );
	}
};

// a side that came out empty keeps the slab of the box it was cut from (degenerate primitives)
__device__ __forceinline__ void repairSplit( uint32_t axis, float position, const DBox& box, DBox& left, DBox& right )
{
	if ( !validBox( left ) )
	{
		left = box;
		setAxis( left.hi, axis, position );
		growBox( left, make_float3( __fmul_rn( __fadd_rn( left.hi.x, left.lo.x ), 0.5f ), __fmul_rn( __fadd_rn( left.hi.y, left.lo.y ), 0.5f ), __fmul_rn( __fadd_rn( left.hi.z, left.lo.z ), 0.5f ) ) );
	}
	if ( !validBox( right ) )
	{
		right = box;
		setAxis( right.lo, axis, position );
		growBox( right, make_float3( __fmul_rn( __fadd_rn( right.hi.x, right.lo.x ), 0.5f ), __fmul_rn( __fadd_rn( right.hi.y, right.lo.y ), 0.5f ), __fmul_rn( __fadd_rn( right.hi.z, right.lo.z ), 0.5f ) ) );
	}
}

// ---- plumbing ------------------------------------------------------------------------------------------------------------------
// Bins: 8 words {lo.xyz, hi.xyz, enter (object bins: count), exit}; a task's bins are contiguous, [axis][bin].
__device__ __forceinline__ void resetBin( float* bin )
{
	reinterpret_cast<float4*>( bin )[0] = make_float4( FltMax, FltMax, FltMax, -FltMax );
	reinterpret_cast<float4*>( bin )[1] = make_float4( -FltMax, -FltMax, 0.0f, 0.0f ); // (the two counters: zero bits)
}
__device__ __forceinline__ void growBinAtomic( float* bin, const DBox& b )
{
	atomicMinF( bin + 0, b.lo.x ), atomicMinF( bin + 1, b.lo.y ), atomicMinF( bin + 2, b.lo.z );
	atomicMaxF( bin + 3, b.hi.x ), atomicMaxF( bin + 4, b.hi.y ), atomicMaxF( bin + 5, b.hi.z );
}
__device__ __forceinline__ DBox loadBox6( const float* p ) // L2: written by other CTAs of this launch
{
	const float2 a = __ldcg( reinterpret_cast<const float2*>( p ) ), b = __ldcg( reinterpret_cast<const float2*>( p ) + 1 ),
				 c = __ldcg( reinterpret_cast<const float2*>( p ) + 2 );
	return DBox{ make_float3( a.x, a.y, b.x ), make_float3( b.y, c.x, c.y ) };
}
__device__ __forceinline__ void storeBox6( float* p, const DBox& b )
{
	reinterpret_cast<float2*>( p )[0] = make_float2( b.lo.x, b.lo.y );
	reinterpret_cast<float2*>( p )[1] = make_float2( b.lo.z, b.hi.x );
	reinterpret_cast<float2*>( p )[2] = make_float2( b.hi.y, b.hi.z );
}
__device__ __forceinline__ uint32_t binsOfLevel( uint32_t taskCount, uint32_t maxRefs )
{
	// the bin arrays hold 3 x MinBins bins for maxRefs / 2 tasks: fewer tasks, more bins each (SbvhBuilder.h:352)
	return min( MinBins * ( maxRefs / 2u ) / max( taskCount, 1u ), MaxBins );
}

// Grid barrier: every CTA of the (cooperative, fully resident) grid arrives once per call. The counter only grows.
// Returns false when the barrier gave up (some CTA never arrived): the caller leaves the kernel.
__device__ __forceinline__ bool gridBarrier( State* state, uint32_t& epoch )
{
	__shared__ uint32_t sOk;
	__syncthreads();
	if ( threadIdx.x == 0 )
	{
		++epoch;
		const uint32_t target = epoch * gridDim.x;
		__threadfence();
		atomicAdd( &state->barrier, 1u );
		uint32_t polls = 0;
		bool	 ok	   = true;
		while ( int32_t( ldVolatile( &state->barrier ) - target ) < 0 )
		{
			if ( ldVolatile( &state->error ) == 1u || ++polls > ( 1u << 24 ) )
			{
				atomicMax( &state->error, 1u );
				ok = false;
				break;
			}
			__nanosleep( 64 );
		}
		__threadfence();
		sOk = ok ? 1u : 0u;
	}
	else
		++epoch;
	__syncthreads();
	return sOk != 0u;
}

__device__ __forceinline__ Rec32 loadRecL2( const void* base, size_t index ) { return loadRecCg( base, index ); }

// ---- phase: reset the object bins of the level ----------------------------------------------------------------------------------
__device__ __forceinline__ void resetBinsPhase( const Params& P, uint32_t cur, uint32_t taskCount, uint32_t bins, bool serial )
{
	const uint32_t total = taskCount * 3u * bins;
	for ( uint32_t i = blockIdx.x * Block + threadIdx.x; i < total; i += gridDim.x * Block )
	{
		// (small tasks of a thread-per-task level never touch their bins: findObjectSplitPhase)
		if ( serial && __ldcg( &P.tasks[cur][i / ( 3u * bins )].count ) <= SmallTask ) continue;
		resetBin( P.objectBins + size_t( i ) * 8 );
	}
}

// ---- phases over references: a CTA walks a contiguous span, 256 references at a time; while the span stays inside one task the
// bins live in shared memory --------------------------------------------------------------------------------------------------
struct SpanWalk
{
	uint32_t chunk, chunkEnd;
	__device__ __forceinline__ SpanWalk( uint32_t refCount )
	{
		const uint32_t chunks	= ( refCount + Block - 1u ) / Block;
		const uint32_t perBlock = ( chunks + gridDim.x - 1u ) / gridDim.x;
		chunk					= min( blockIdx.x * perBlock, chunks );
		chunkEnd				= min( chunk + perBlock, chunks );
	}
};

template <typename Source, bool Spatial>
__device__ __forceinline__ void binReference(
	const Params& P, const Source& src, const Rec32& ref, const DBox& taskBox, float* bins, uint32_t binCount, float edgeThreshold )
{
	if constexpr ( !Spatial )
	{
		const uint32_t bx = centroidBin( ref.box.lo.x, ref.box.hi.x, taskBox.lo.x, taskBox.hi.x, binCount );
		const uint32_t by = centroidBin( ref.box.lo.y, ref.box.hi.y, taskBox.lo.y, taskBox.hi.y, binCount );
		const uint32_t bz = centroidBin( ref.box.lo.z, ref.box.hi.z, taskBox.lo.z, taskBox.hi.z, binCount );
		float* b = bins + size_t( bx ) * 8;
		growBinAtomic( b, ref.box ), atomicAdd( reinterpret_cast<uint32_t*>( b ) + 6, 1u );
		b = bins + size_t( binCount + by ) * 8;
		growBinAtomic( b, ref.box ), atomicAdd( reinterpret_cast<uint32_t*>( b ) + 6, 1u );
		b = bins + size_t( 2u * binCount + bz ) * 8;
		growBinAtomic( b, ref.box ), atomicAdd( reinterpret_cast<uint32_t*>( b ) + 6, 1u );
	}
	else
	{
		const Clipper<Source> clipper( src, ref.a, ref.b );
		for ( uint32_t axis = 0; axis < 3; ++axis )
		{
			const float taskLo = axisOf( taskBox, false, axis ), taskHi = axisOf( taskBox, true, axis );
			if ( __fsub_rn( taskHi, taskLo ) < edgeThreshold ) continue;
			const float binSize = __fdiv_rn( __fsub_rn( taskHi, taskLo ), float( binCount ) );
			uint32_t	first, last;
			spatialBinRange( axisOf( ref.box, false, axis ), axisOf( ref.box, true, axis ), taskLo, binSize, binCount, first, last );
			DBox   cur		= ref.box;
			float* axisBins = bins + size_t( axis * binCount ) * 8;
			for ( uint32_t i = first; i < last; ++i )
			{
				const float position = binPlane( taskLo, binSize, i );
				DBox		left, right;
				clipper.split( axis, position, cur, left, right );
				repairSplit( axis, position, cur, left, right );
				growBinAtomic( axisBins + size_t( i ) * 8, left );
				cur = right;
			}
			atomicAdd( reinterpret_cast<uint32_t*>( axisBins + size_t( first ) * 8 ) + 6, 1u ); // enter
			atomicAdd( reinterpret_cast<uint32_t*>( axisBins + size_t( last ) * 8 ) + 7, 1u );	// exit
			growBinAtomic( axisBins + size_t( last ) * 8, cur );
		}
	}
}

// adds a CTA's shared-memory bins to a task's bins
__device__ __forceinline__ void flushBins( const float* sBins, float* taskBins, uint32_t binCount )
{
	for ( uint32_t i = threadIdx.x; i < 3u * binCount; i += Block )
	{
		const float*   s = sBins + i * 8;
		const uint32_t enter = reinterpret_cast<const uint32_t*>( s )[6], exit = reinterpret_cast<const uint32_t*>( s )[7];
		float*		   g = taskBins + size_t( i ) * 8;
		if ( s[0] <= s[3] ) growBinAtomic( g, DBox{ make_float3( s[0], s[1], s[2] ), make_float3( s[3], s[4], s[5] ) } );
		if ( enter ) atomicAdd( reinterpret_cast<uint32_t*>( g ) + 6, enter );
		if ( exit ) atomicAdd( reinterpret_cast<uint32_t*>( g ) + 7, exit );
	}
}

template <typename Source, bool Spatial>
__device__ __forceinline__ void binPhase(
	const Params& P, const Source& src, uint32_t cur, uint32_t refCount, uint32_t binCount, float edgeThreshold, float* sBins, uint32_t* sWord, bool serial )
{
	const void*		refs	 = P.refs[cur];
	const uint32_t* refTask	 = P.refTask[cur];
	Task*			tasks	 = P.tasks[cur];
	float*			binArray = Spatial ? P.spatialBins : P.objectBins;
	uint32_t		held	 = InvalidValue; // the task whose bins are in shared memory
	for ( SpanWalk w( refCount ); w.chunk < w.chunkEnd; ++w.chunk )
	{
		const uint32_t index = w.chunk * Block + threadIdx.x;
		const bool	   valid = index < refCount;
		const uint32_t task	 = valid ? __ldcg( refTask + index ) : InvalidValue;
		if ( threadIdx.x == 0 ) *sWord = task;
		__syncthreads();
		const uint32_t first   = *sWord;
		const bool	   uniform = __syncthreads_and( !valid || task == first ) != 0;
		if ( uniform && first != held )
		{
			if ( held != InvalidValue ) flushBins( sBins, binArray + size_t( held ) * 3u * binCount * 8, binCount );
			__syncthreads();
			for ( uint32_t i = threadIdx.x; i < 3u * binCount; i += Block )
				resetBin( sBins + i * 8 );
			held = first;
			__syncthreads();
		}
		if ( valid )
		{
			bool wanted = true;
			if constexpr ( Spatial )
				wanted = ( __ldcg( &tasks[task].split ) & 8u ) != 0u;
			else
				wanted = !( serial && __ldcg( &tasks[task].count ) <= SmallTask ); // (split straight from the references)
			if ( wanted )
			{
				const Rec32 ref		= loadRecL2( refs, index );
				const DBox	taskBox = loadBox6( tasks[task].box );
				// Two references whose centroids no bin plane separates are told apart by their primitive (a task holds a
				// primitive at most once): the one with the smaller first word goes left, whatever the arrival order.
				if constexpr ( !Spatial )
					if ( __ldcg( &tasks[task].count ) == 2u ) atomicMin( &tasks[task].out0, ref.a );
				float*		bins	= ( uniform && task == held ) ? sBins : binArray + size_t( task ) * 3u * binCount * 8;
				binReference<Source, Spatial>( P, src, ref, taskBox, bins, binCount, edgeThreshold );
			}
		}
	}
	__syncthreads();
	if ( held != InvalidValue ) flushBins( sBins, binArray + size_t( held ) * 3u * binCount * 8, binCount );
	__syncthreads();
}

// ---- sub-warp scans over a task's bins: lane i of a group of `width` lanes holds bin i --------------------------------------------
struct BinAcc
{
	DBox	 box;
	uint32_t enter, exit;
};
__device__ __forceinline__ BinAcc loadBin( const float* bin )
{
	const float4 a = __ldcg( reinterpret_cast<const float4*>( bin ) ), b = __ldcg( reinterpret_cast<const float4*>( bin ) + 1 );
	return BinAcc{ DBox{ make_float3( a.x, a.y, a.z ), make_float3( a.w, b.x, b.y ) }, __float_as_uint( b.z ), __float_as_uint( b.w ) };
}
__device__ __forceinline__ BinAcc shflBin( const BinAcc& v, int delta, bool up, uint32_t width )
{
	BinAcc o;
#define HIPRT_SHFL( x ) ( up ? __shfl_up_sync( FullMask, x, delta, width ) : __shfl_down_sync( FullMask, x, delta, width ) )
	o.box.lo = make_float3( HIPRT_SHFL( v.box.lo.x ), HIPRT_SHFL( v.box.lo.y ), HIPRT_SHFL( v.box.lo.z ) );
	o.box.hi = make_float3( HIPRT_SHFL( v.box.hi.x ), HIPRT_SHFL( v.box.hi.y ), HIPRT_SHFL( v.box.hi.z ) );
	o.enter = HIPRT_SHFL( v.enter ), o.exit = HIPRT_SHFL( v.exit );
#undef HIPRT_SHFL
	return o;
}
// inclusive union over lanes <= i (up) or >= i (down) of the group
__device__ __forceinline__ BinAcc scanBins( BinAcc v, bool up, uint32_t width, uint32_t laneInGroup )
{
	for ( uint32_t d = 1; d < width; d <<= 1 )
	{
		const BinAcc o	= shflBin( v, int( d ), up, width );
		const bool	 in = up ? laneInGroup >= d : laneInGroup + d < width;
		if ( in ) v.box = uniteBox( v.box, o.box ), v.enter += o.enter, v.exit += o.exit;
	}
	return v;
}

struct SplitChoice
{
	float	 cost;
	uint32_t key; // axis * 32 + bin; InvalidValue: none
	DBox	 box0, box1;
	uint32_t count0, count1;
};
// lexicographic minimum of (cost, key) over the group, then the winner's boxes and counts for everybody
__device__ __forceinline__ void reduceChoice( SplitChoice& c, uint32_t width, uint32_t groupBase )
{
	float	 cost = c.cost;
	uint32_t key  = c.key;
	for ( uint32_t d = width >> 1; d > 0; d >>= 1 )
	{
		const float	   oc = __shfl_xor_sync( FullMask, cost, d );
		const uint32_t ok = __shfl_xor_sync( FullMask, key, d );
		if ( ok != InvalidValue && ( key == InvalidValue || oc < cost || ( oc == cost && ok < key ) ) ) cost = oc, key = ok;
	}
	// the winner is the lane that proposed (cost, key): lane (key & 31) of the group, with its own best equal to the group's
	const uint32_t srcLane = groupBase + ( key == InvalidValue ? 0u : ( key & 31u ) );
#define HIPRT_PULL( x ) x = __shfl_sync( FullMask, x, srcLane )
	HIPRT_PULL( c.box0.lo.x ), HIPRT_PULL( c.box0.lo.y ), HIPRT_PULL( c.box0.lo.z );
	HIPRT_PULL( c.box0.hi.x ), HIPRT_PULL( c.box0.hi.y ), HIPRT_PULL( c.box0.hi.z );
	HIPRT_PULL( c.box1.lo.x ), HIPRT_PULL( c.box1.lo.y ), HIPRT_PULL( c.box1.lo.z );
	HIPRT_PULL( c.box1.hi.x ), HIPRT_PULL( c.box1.hi.y ), HIPRT_PULL( c.box1.hi.z );
	HIPRT_PULL( c.count0 ), HIPRT_PULL( c.count1 );
#undef HIPRT_PULL
	c.cost = cost, c.key = key;
}

// Best split of one task over its bins. Object bins: counts in `enter`; spatial bins: enter on the left, exit on the right.
// A lane keeps the best candidate of ITS bin index over the three axes (ties: the lower axis), the group then reduces.
template <bool Spatial>
__device__ __forceinline__ SplitChoice searchBins(
	const float* taskBins, bool active, const DBox& taskBox, float startCost, float edgeThreshold, uint32_t binCount, uint32_t width, uint32_t laneInGroup,
	uint32_t groupBase )
{
	SplitChoice best;
	best.cost = startCost, best.key = InvalidValue, best.box0 = emptyBox(), best.box1 = emptyBox(), best.count0 = 0, best.count1 = 0;
	// the bins of the three axes are fetched together: the phase is bound by the latency of these loads (the build runs at two
	// CTAs per SM), three in flight per lane instead of one
	BinAcc bins3[3];
	bool   on3[3];
#pragma unroll
	for ( uint32_t axis = 0; axis < 3; ++axis )
	{
		on3[axis] = active;
		if constexpr ( Spatial ) on3[axis] = active && !( __fsub_rn( axisOf( taskBox, true, axis ), axisOf( taskBox, false, axis ) ) < edgeThreshold );
		bins3[axis] = BinAcc{ emptyBox(), 0u, 0u }; // (lanes past the last bin hold empty bins)
		if ( on3[axis] && laneInGroup < binCount ) bins3[axis] = loadBin( taskBins + size_t( axis * binCount + laneInGroup ) * 8 );
	}
#pragma unroll
	for ( uint32_t axis = 0; axis < 3; ++axis )
	{
		const bool	 axisOn = on3[axis];
		const BinAcc mine	= bins3[axis];
		const BinAcc left  = scanBins( mine, true, width, laneInGroup );
		const BinAcc right = shflBin( scanBins( mine, false, width, laneInGroup ), 1, false, width ); // bins above mine
		const uint32_t nl = left.enter, nr = Spatial ? right.exit : right.enter;
		if ( axisOn && laneInGroup + 1u < binCount && nl > 0u && nr > 0u )
		{
			const float cost = sahCost( left.box, nl, right.box, nr );
			if ( best.cost > cost ) best.cost = cost, best.key = axis * 32u + laneInGroup, best.box0 = left.box, best.box1 = right.box, best.count0 = nl, best.count1 = nr;
		}
	}
	reduceChoice( best, width, groupBase );
	return best;
}

// The same search by ONE thread (deep levels: millions of tasks with few bins each -- a sub-warp per task spends about 400 warp
// instructions per task on shuffles there, a thread per task about 20). Suffix unions go to a local array, the bins are read twice (the
// second time out of L2), four bins in flight per batch. Same candidates in the same order, same arithmetic: the same choice.
template <bool Spatial>
__device__ __forceinline__ SplitChoice searchBinsSerial( const float* taskBins, bool active, const DBox& taskBox, float startCost, float edgeThreshold, uint32_t binCount )
{
	SplitChoice best;
	best.cost = startCost, best.key = InvalidValue, best.box0 = emptyBox(), best.box1 = emptyBox(), best.count0 = 0, best.count1 = 0;
	if ( !active ) return best;
	float	 suffix[MaxBins][6];
	uint32_t suffixCount[MaxBins];
	for ( uint32_t axis = 0; axis < 3; ++axis )
	{
		if constexpr ( Spatial )
			if ( __fsub_rn( axisOf( taskBox, true, axis ), axisOf( taskBox, false, axis ) ) < edgeThreshold ) continue;
		const float* bins = taskBins + size_t( axis * binCount ) * 8;
		BinAcc		 acc{ emptyBox(), 0u, 0u };
		for ( int hi = int( binCount ) - 1; hi >= 1; hi -= 4 ) // bins hi, hi-1, hi-2, hi-3 (down to 1): loads first, unions after
		{
			BinAcc b[4];
#pragma unroll
			for ( int k = 0; k < 4; ++k )
				b[k] = hi - k >= 1 ? loadBin( bins + size_t( hi - k ) * 8 ) : BinAcc{ emptyBox(), 0u, 0u };
#pragma unroll
			for ( int k = 0; k < 4; ++k )
			{
				if ( hi - k < 1 ) break;
				acc.box = uniteBox( acc.box, b[k].box ), acc.enter += b[k].enter, acc.exit += b[k].exit;
				float* d = suffix[hi - k];
				d[0] = acc.box.lo.x, d[1] = acc.box.lo.y, d[2] = acc.box.lo.z, d[3] = acc.box.hi.x, d[4] = acc.box.hi.y, d[5] = acc.box.hi.z;
				suffixCount[hi - k] = Spatial ? acc.exit : acc.enter;
			}
		}
		BinAcc left{ emptyBox(), 0u, 0u };
		for ( uint32_t lo = 0; lo + 1u < binCount; lo += 4 )
		{
			BinAcc b[4];
#pragma unroll
			for ( uint32_t k = 0; k < 4; ++k )
				b[k] = lo + k + 1u < binCount ? loadBin( bins + size_t( lo + k ) * 8 ) : BinAcc{ emptyBox(), 0u, 0u };
#pragma unroll
			for ( uint32_t k = 0; k < 4; ++k )
			{
				const uint32_t i = lo + k;
				if ( i + 1u >= binCount ) break;
				left.box = uniteBox( left.box, b[k].box ), left.enter += b[k].enter;
				const uint32_t nl = left.enter, nr = suffixCount[i + 1u];
				if ( nl > 0u && nr > 0u )
				{
					const float* r = suffix[i + 1u];
					const DBox	 right{ make_float3( r[0], r[1], r[2] ), make_float3( r[3], r[4], r[5] ) };
					const float	 cost = sahCost( left.box, nl, right, nr );
					if ( best.cost > cost ) best.cost = cost, best.key = axis * 32u + i, best.box0 = left.box, best.box1 = right, best.count0 = nl, best.count1 = nr;
				}
			}
		}
	}
	return best;
}

// A task of at most SmallTask references, split straight from its references (a contiguous segment of the reference buffer): the
// bins would hold at most four boxes in a kilobyte. Same candidates as the binned search -- the planes right above a reference's
// centroid bin, except the highest -- same costs; among equal costs the lowest (axis, bin) wins, as in the ordered scan.
// Also returns the smallest first word of the references (what the binning phase collects for two-reference tasks).
__device__ __forceinline__ SplitChoice searchSmallTask( const void* refs, uint32_t begin, uint32_t count, const DBox& taskBox, uint32_t binCount, uint32_t& minWord )
{
	SplitChoice best;
	best.cost = FltMax, best.key = InvalidValue, best.box0 = emptyBox(), best.box1 = emptyBox(), best.count0 = 0, best.count1 = 0;
	DBox	 box[SmallTask];
	uint32_t bin[SmallTask][3];
	minWord = InvalidValue;
#pragma unroll
	for ( uint32_t r = 0; r < SmallTask; ++r )
	{
		if ( r >= count ) break;
		const Rec32 ref = loadRecCg( refs, begin + r );
		box[r]			= ref.box;
		minWord			= min( minWord, ref.a );
		bin[r][0]		= centroidBin( ref.box.lo.x, ref.box.hi.x, taskBox.lo.x, taskBox.hi.x, binCount );
		bin[r][1]		= centroidBin( ref.box.lo.y, ref.box.hi.y, taskBox.lo.y, taskBox.hi.y, binCount );
		bin[r][2]		= centroidBin( ref.box.lo.z, ref.box.hi.z, taskBox.lo.z, taskBox.hi.z, binCount );
	}
#pragma unroll
	for ( uint32_t axis = 0; axis < 3; ++axis )
#pragma unroll
		for ( uint32_t c = 0; c < SmallTask; ++c )
		{
			if ( c >= count ) break;
			const uint32_t plane = bin[c][axis];
			DBox		   l = emptyBox(), r = emptyBox();
			uint32_t	   nl = 0, nr = 0;
#pragma unroll
			for ( uint32_t k = 0; k < SmallTask; ++k )
			{
				if ( k >= count ) break;
				if ( bin[k][axis] <= plane )
					l = uniteBox( l, box[k] ), ++nl;
				else
					r = uniteBox( r, box[k] ), ++nr;
			}
			if ( nr == 0u ) continue; // the highest bin: no plane above it separates anything
			const float	   cost = sahCost( l, nl, r, nr );
			const uint32_t key	= axis * 32u + plane;
			if ( cost < best.cost || ( cost == best.cost && key < best.key ) ) best.cost = cost, best.key = key, best.box0 = l, best.box1 = r, best.count0 = nl, best.count1 = nr;
		}
	return best;
}

// ---- phase: object split of every task ----------------------------------------------------------------------------------------
__device__ __forceinline__ void findObjectSplitPhase(
	const Params& P, uint32_t cur, uint32_t level, uint32_t forceLevel, uint32_t taskCount, uint32_t binCount, float overlapThreshold )
{
	Task*		   tasks		 = P.tasks[cur];
	const uint32_t lane			 = laneId();
	const bool	   serial		 = taskCount >= P.serialMin;										   // one thread per task
	const uint32_t width		 = serial ? 1u : ( binCount <= 8u ? 8u : ( binCount <= 16u ? 16u : 32u ) ); // lanes per task
	const uint32_t laneInGroup	 = lane % width, group = lane / width, groupsPerWarp = 32u / width;
	const uint32_t groupBase	 = group * width;
	const uint32_t warpsInGrid	 = gridDim.x * ( Block / 32u );
	const uint32_t warp			 = blockIdx.x * ( Block / 32u ) + ( threadIdx.x >> 5 );
	uint32_t	   spatialTasks	 = 0;
	for ( uint32_t base = warp * groupsPerWarp; base < taskCount; base += warpsInGrid * groupsPerWarp )
	{
		const uint32_t t	  = base + group;
		const bool	   active = t < taskCount;
		Task*		   task	  = tasks + ( active ? t : 0u );
		const DBox	   box	  = loadBox6( task->box );
		const uint32_t count  = __ldcg( &task->count );
		const bool	   forced = level >= forceLevel;
		const float*   taskBins = P.objectBins + size_t( active ? t : 0u ) * 3u * binCount * 8;
		const bool	   small	= serial && active && count <= SmallTask;
		SplitChoice	   c;
		if ( small )
		{
			uint32_t minWord;
			c = searchSmallTask( P.refs[cur], __ldcg( &task->cursor0 ), count, box, binCount, minWord );
			if ( forced ) c.key = InvalidValue;
			if ( count == 2u ) task->out0 = minWord;
		}
		else
			c = serial ? searchBinsSerial<false>( taskBins, active && !forced, box, FltMax, 0.0f, binCount )
					   : searchBins<false>( taskBins, active && !forced, box, FltMax, 0.0f, binCount, width, laneInGroup, groupBase );
		uint32_t axis = c.key >> 5, index = c.key & 31u;
		if ( c.key == InvalidValue )
		{
			// no plane separates the centroids (or the depth bound is reached): halve the references by arrival
			c.cost	 = __fmul_rn( boxArea( box ), float( count ) );
			c.count0 = count >> 1, c.count1 = count - ( count >> 1 );
			c.box0 = box, c.box1 = box;
			axis = 3u, index = count >> 1;
		}
		bool trySpatial = false;
		if ( P.spatialBins != nullptr && !forced && active ) trySpatial = areaOrZero( intersectBox( c.box0, c.box1 ) ) >= overlapThreshold;
		if ( active )
		{
			if ( laneInGroup == 0 )
			{
				storeBox6( task->box0, c.box0 ), storeBox6( task->box1, c.box1 );
				task->count0 = c.count0, task->count1 = c.count1;
				task->cost	 = c.cost;
				task->split = axis | ( trySpatial ? 8u : 0u ) | ( index << 4 );
				if ( trySpatial ) ++spatialTasks;
			}
			if ( trySpatial && serial )
				for ( uint32_t i = 0; i < 3u * binCount; ++i )
					resetBin( P.spatialBins + ( size_t( t ) * 3u * binCount + i ) * 8 );
			else if ( trySpatial && laneInGroup < binCount )
				for ( uint32_t a = 0; a < 3; ++a )
					resetBin( P.spatialBins + ( size_t( t ) * 3u * binCount + a * binCount + laneInGroup ) * 8 );
		}
	}
	// one atomic per warp
	for ( int o = 16; o > 0; o >>= 1 )
		spatialTasks += __shfl_xor_sync( FullMask, spatialTasks, o );
	if ( lane == 0 && spatialTasks ) atomicAdd( &P.state->level[level & 1u].spatialTasks, spatialTasks );
}

// ---- phase: final split of every task, allocation of what it creates ----------------------------------------------------------------
// DemandOnly: the spatial search alone, to add up the references the level's spatial splits would duplicate (LevelCounts::demand);
// the split phase proper repeats the search (cheap next to the binning) knowing whether the level's demand fits the budget.
template <typename Source, bool DemandOnly>
__device__ __forceinline__ void splitPhase(
	const Params& P, uint32_t cur, uint32_t level, uint32_t remaining, uint32_t demand, uint32_t refCount, uint32_t taskCount, uint32_t binCount,
	float edgeThreshold )
{
	Task*		   tasks	   = P.tasks[cur];
	Task*		   nextTasks   = P.tasks[cur ^ 1u];
	State*		   state	   = P.state;
	LevelCounts*   next		   = &state->level[( level + 1u ) & 1u];
	const uint32_t lane		   = laneId();
	const bool	   serial	   = taskCount >= P.serialMin;											 // one thread per task
	const uint32_t width	   = serial ? 1u : ( binCount <= 8u ? 8u : ( binCount <= 16u ? 16u : 32u ) ); // lanes per task
	const uint32_t laneInGroup = lane % width, group = lane / width, groupsPerWarp = 32u / width;
	const uint32_t groupBase   = group * width;
	const uint32_t warpsInGrid = gridDim.x * ( Block / 32u );
	const uint32_t warp		   = blockIdx.x * ( Block / 32u ) + ( threadIdx.x >> 5 );
	for ( uint32_t base = warp * groupsPerWarp; base < taskCount; base += warpsInGrid * groupsPerWarp )
	{
		const uint32_t t	  = base + group;
		const bool	   active = t < taskCount;
		Task*		   task	  = tasks + ( active ? t : 0u );
		const DBox	   box	  = loadBox6( task->box );
		uint32_t	   split  = __ldcg( &task->split );
		const bool	   trySpatial = active && ( split & 8u ) != 0u;
		// (whole warps run the search: its shuffles name every lane)
		const bool anySpatial = __any_sync( FullMask, trySpatial );
		SplitChoice c;
		c.key = InvalidValue;
		if ( anySpatial )
		{
			const float* taskBins  = P.spatialBins + size_t( active ? t : 0u ) * 3u * binCount * 8;
			const float	 startCost = trySpatial ? __ldcg( &task->cost ) : 0.0f;
			c = serial ? searchBinsSerial<true>( taskBins, trySpatial, box, startCost, edgeThreshold, binCount )
					   : searchBins<true>( taskBins, trySpatial, box, startCost, edgeThreshold, binCount, width, laneInGroup, groupBase );
		}
		const bool leader = active && laneInGroup == 0;
		DBox	   box0 = emptyBox(), box1 = emptyBox();
		uint32_t   count0 = 0, count1 = 0;
		if ( leader )
		{
			box0 = loadBox6( task->box0 ), box1 = loadBox6( task->box1 );
			count0 = __ldcg( &task->count0 ), count1 = __ldcg( &task->count1 );
		}
		if constexpr ( DemandOnly )
		{
			if ( leader && trySpatial && c.key != InvalidValue ) atomicAdd( &state->level[level & 1u].demand, c.count0 + c.count1 - count0 - count1 );
			continue;
		}
		if ( leader && trySpatial && c.key != InvalidValue )
		{
			// Every reference the split duplicates must fit the reference budget (SbvhBuilderKernels.h:571-587, where the tasks of
			// a level draw on the budget in atomic arrival order until it is gone). Here the outcome does not depend on any order:
			// a level whose whole demand fits what is left gets all its spatial splits; in the level that would overdraw, a task
			// may use the share of the rest that its reference count has among the level's references (the shares cannot
			// overdraw either), and what a level leaves goes to the next.
			const uint32_t duplicates = c.count0 + c.count1 - count0 - count1;
			const uint32_t share	  = demand <= remaining ? remaining : uint32_t( uint64_t( remaining ) * __ldcg( &task->count ) / refCount );
			if ( duplicates <= share )
			{
				atomicAdd( &state->totalRefs, duplicates );
				box0 = c.box0, box1 = c.box1, count0 = c.count0, count1 = c.count1;
				split = ( c.key >> 5 ) | 4u | ( ( c.key & 31u ) << 4 );
				storeBox6( task->box0, box0 ), storeBox6( task->box1, box1 );
				task->count0 = count0, task->count1 = count1;
			}
		}
		split &= ~8u;
		// What the task creates -- a binary node + a task + a segment of references for a child with several references, a leaf
		// record for a child with one -- is allocated once per warp: the four counters are single hot lines of the whole grid
		// (ncu: a sixth of all stall samples of the build sat on per-task atomics here).
		const uint32_t inner = leader ? uint32_t( count0 > 1u ) + uint32_t( count1 > 1u ) : 0u;
		const uint32_t leafs = leader ? 2u - inner : 0u;
		const uint32_t refs	 = leader ? ( count0 > 1u ? count0 : 0u ) + ( count1 > 1u ? count1 : 0u ) : 0u;
		uint32_t	   innerScan = inner, leafScan = leafs, refScan = refs; // inclusive scans over the warp's leaders
		for ( uint32_t d = width; d < 32u; d <<= 1 )
		{
			const uint32_t a = __shfl_up_sync( FullMask, innerScan, d ), b = __shfl_up_sync( FullMask, leafScan, d ), e = __shfl_up_sync( FullMask, refScan, d );
			if ( lane >= d ) innerScan += a, leafScan += b, refScan += e;
		}
		uint32_t nodeBase = 0, taskBase = 0, refBase = 0, leafBase = 0;
		if ( lane == 31u - ( 31u % width ) ) // the last leader lane holds the totals
		{
			if ( innerScan )
			{
				nodeBase = atomicAdd( &state->nodeCount, innerScan );
				taskBase = atomicAdd( &next->taskCount, innerScan );
				refBase	 = atomicAdd( &next->activeRefs, refScan );
			}
			if ( leafScan ) leafBase = atomicAdd( &state->leafCount, leafScan );
		}
		const uint32_t last = 31u - ( 31u % width );
		nodeBase = __shfl_sync( FullMask, nodeBase, last ) + innerScan - inner;
		taskBase = __shfl_sync( FullMask, taskBase, last ) + innerScan - inner;
		refBase	 = __shfl_sync( FullMask, refBase, last ) + refScan - refs;
		leafBase = __shfl_sync( FullMask, leafBase, last ) + leafScan - leafs;
		if ( !leader ) continue;
		if ( nodeBase + inner > P.maxRefs || taskBase + inner > P.maxRefs / 2u || refBase + refs > P.maxRefs || leafBase + leafs > P.maxRefs )
		{
			atomicMax( &state->error, 2u );
			continue;
		}

		const uint32_t minWord = __ldcg( &task->out0 );
		uint32_t	   child[2], out[2], nextTask[2];
		for ( uint32_t side = 0; side < 2; ++side )
		{
			const uint32_t n = side ? count1 : count0;
			nextTask[side]	 = InvalidValue;
			if ( n > 1u )
			{
				const uint32_t node = nodeBase++, nt = taskBase++;
				out[side]	   = refBase, refBase += n;
				child[side]	   = makeNodeIndex( node, BoxType );
				nextTask[side] = nt;
				Task* c2	   = nextTasks + nt;
				storeBox6( c2->box, side ? box1 : box0 );
				c2->count = n, c2->node = node, c2->split = 0u, c2->out0 = InvalidValue;
				c2->cursor0 = out[side]; // (until the split phase: first reference of the task in the next buffer)
			}
			else
			{
				out[side]	= leafBase++;
				child[side] = makeNodeIndex( out[side], uint32_t( Source::LeafType ) );
			}
		}
		task->child0 = child[0], task->child1 = child[1];
		task->out0 = out[0], task->out1 = out[1];
		task->cursor0 = 0u, task->cursor1 = 0u, task->arrival = 0u;
		task->nextTask0 = nextTask[0], task->nextTask1 = nextTask[1];
		if ( ( split & 7u ) == 3u && count0 + count1 == 2u ) task->nextTask0 = minWord; // (both children are leaves: the field is free)
		task->split = split;
		storeRec( P.nodes, __ldcg( &task->node ), box, child[0], child[1] );
	}
}

// ---- phase: references move to the children ---------------------------------------------------------------------------------------
// place `n` (0, 1) items of a (task, side) stream: lanes of a warp that feed the same cursor share one atomic
__device__ __forceinline__ uint32_t claimSlot( uint32_t* cursor, bool wants )
{
	uint32_t	   slot	 = 0;
	const uint32_t peers = __match_any_sync( FullMask, wants ? reinterpret_cast<uint64_t>( cursor ) : 0ull );
	if ( wants )
	{
		const uint32_t leader = __ffs( peers ) - 1u;
		uint32_t	   base	  = 0;
		if ( laneId() == leader ) base = atomicAdd( cursor, __popc( peers ) );
		base = __shfl_sync( peers, base, leader );
		slot = base + __popc( peers & laneMaskLt() );
	}
	return slot;
}

template <typename Source>
__device__ __forceinline__ void distributePhase( const Params& P, const Source& src, uint32_t cur, uint32_t refCount, uint32_t binCount )
{
	const void*		refs	 = P.refs[cur];
	const uint32_t* refTask	 = P.refTask[cur];
	Task*			tasks	 = P.tasks[cur];
	void*			outRefs	 = P.refs[cur ^ 1u];
	uint32_t*		outTask	 = P.refTask[cur ^ 1u];
	for ( SpanWalk w( refCount ); w.chunk < w.chunkEnd; ++w.chunk )
	{
		const uint32_t index = w.chunk * Block + threadIdx.x;
		const bool	   valid = index < refCount;
		bool		   goes[2] = { false, false };
		DBox		   boxes[2];
		Rec32		   ref{};
		Task*		   task = tasks;
		if ( valid )
		{
			const uint32_t t = __ldcg( refTask + index );
			task			 = tasks + t;
			ref				 = loadRecL2( refs, index );
			boxes[0] = ref.box, boxes[1] = ref.box;
			const DBox	   taskBox = loadBox6( task->box );
			const uint32_t split   = __ldcg( &task->split );
			const uint32_t axis = split & 3u, splitIndex = split >> 4;
			if ( !( split & 4u ) )
			{
				bool onLeft;
				if ( axis < 3u )
					onLeft = centroidBin( axisOf( ref.box, false, axis ), axisOf( ref.box, true, axis ), axisOf( taskBox, false, axis ), axisOf( taskBox, true, axis ), binCount ) <= splitIndex;
				else if ( __ldcg( &task->count ) == 2u )
					onLeft = ref.a == __ldcg( &task->nextTask0 );
				else
					onLeft = atomicAdd( &task->arrival, 1u ) < splitIndex;
				goes[onLeft ? 0 : 1] = true;
			}
			else
			{
				const float taskLo	= axisOf( taskBox, false, axis );
				const float binSize = __fdiv_rn( __fsub_rn( axisOf( taskBox, true, axis ), taskLo ), float( binCount ) );
				uint32_t	first, last;
				spatialBinRange( axisOf( ref.box, false, axis ), axisOf( ref.box, true, axis ), taskLo, binSize, binCount, first, last );
				const float position = binPlane( taskLo, binSize, splitIndex );
				if ( first <= splitIndex && last > splitIndex )
				{
					const Clipper<Source> clipper( src, ref.a, ref.b );
					clipper.split( axis, position, ref.box, boxes[0], boxes[1] );
					repairSplit( axis, position, ref.box, boxes[0], boxes[1] );
					goes[0] = goes[1] = true;
				}
				else
					goes[last <= splitIndex ? 0 : 1] = true;
			}
		}
		for ( uint32_t side = 0; side < 2; ++side )
		{
			const bool	   wants	= valid && goes[side];
			const uint32_t count	= wants ? __ldcg( side ? &task->count1 : &task->count0 ) : 0u;
			const uint32_t out		= wants ? __ldcg( side ? &task->out1 : &task->out0 ) : 0u;
			const bool	   toLeaf	= count == 1u;
			const uint32_t slot		= claimSlot( side ? &task->cursor1 : &task->cursor0, wants && !toLeaf );
			if ( !wants ) continue;
			if ( toLeaf )
				storeRec( P.leaves, out, boxes[side], ref.a, ref.b );
			else if ( slot < count ) // (always: the counts are what the bins measured; a corrupted input must not write elsewhere)
			{
				storeRec( outRefs, out + slot, boxes[side], ref.a, ref.b );
				outTask[out + slot] = __ldcg( side ? &task->nextTask1 : &task->nextTask0 );
			}
			else
				atomicMax( &P.state->error, 2u );
		}
	}
}

// ---- the build --------------------------------------------------------------------------------------------------------------------
template <typename Source>
__global__ void __launch_bounds__( Block, 2 ) buildKernel( Source src, Params P )
{
	__shared__ __align__( 16 ) float sBins[3 * MaxBins * 8];
	__shared__ uint32_t				 sWord;
	State*		   state = P.state;
	uint32_t	   epoch = 0;
	const uint32_t n	 = P.ctl->primCount;
	const DBox	   sceneBox{ make_float3( P.ctl->sceneBox.lo.x, P.ctl->sceneBox.lo.y, P.ctl->sceneBox.lo.z ), make_float3( P.ctl->sceneBox.hi.x, P.ctl->sceneBox.hi.y, P.ctl->sceneBox.hi.z ) };
	// The reference budget follows the primitive count AFTER pairing, known on the device only: half as many again with spatial
	// splits (SbvhAlpha, BvhConfig.h:48). From level `forceLevel` on tasks are halved by arrival, which bounds the depth whatever the input.
	const uint32_t budget = min( P.maxRefs, P.spatialBins != nullptr ? n + n / 2u : n );
	uint32_t	   log2n  = 0;
	while ( ( 1ull << log2n ) < n )
		++log2n;
	const uint32_t forceLevel		= max( 32u, 2u * log2n + 8u );
	const float	   overlapThreshold = __fmul_rn( boxArea( sceneBox ), Beta );
	const float	   edgeThreshold =
		__fmul_rn( fmaxf( fmaxf( __fsub_rn( sceneBox.hi.x, sceneBox.lo.x ), __fsub_rn( sceneBox.hi.y, sceneBox.lo.y ) ), __fsub_rn( sceneBox.hi.z, sceneBox.lo.z ) ), Gamma );

	// level 0: every primitive is a reference of the root task
	for ( uint32_t i = blockIdx.x * Block + threadIdx.x; i < n; i += gridDim.x * Block )
	{
		const Rec32 r = loadRec( P.primRefs, i );
		storeRec( P.refs[0], i, r.box, r.a, r.b );
		P.refTask[0][i] = 0u;
	}
	if ( blockIdx.x == 0 && threadIdx.x == 0 )
	{
		Task* root = P.tasks[0];
		storeBox6( root->box, sceneBox );
		root->count = n, root->node = 0u, root->split = 0u, root->out0 = InvalidValue, root->cursor0 = 0u;
		state->level[0] = LevelCounts{ 1u, n, 0u, 0u };
		state->level[1] = LevelCounts{ 0u, 0u, 0u, 0u };
		state->totalRefs = n, state->leafCount = 0u, state->nodeCount = 1u;
	}
	// a barrier that gives up (a CTA of the grid is missing: cannot happen under a cooperative launch) ends the build: the collapse
	// that follows finds nothing to do and the error flag is raised
#define HIPRT_SBVH_BARRIER()                                                   \
	if ( !gridBarrier( state, epoch ) )                                        \
	{                                                                          \
		if ( threadIdx.x == 0 ) P.ctl->primCount = 0u, P.ctl->pad[0] = 1u; \
		return;                                                                \
	}
	HIPRT_SBVH_BARRIER();

	uint32_t level = 0;
	auto	 stamp = [&]( uint32_t slot ) {
		if ( P.timeline != nullptr && blockIdx.x == 0 && threadIdx.x == 0 && level < TimelineLevels )
		{
			uint64_t now;
			asm volatile( "mov.u64 %0, %%globaltimer;" : "=l"( now ) );
			P.timeline[level * TimelineSlots + slot] = now;
		}
	};
	while ( true )
	{
		const uint32_t cur		 = level & 1u;
		const uint32_t taskCount = ldVolatile( &state->level[cur].taskCount );
		const uint32_t refCount	 = ldVolatile( &state->level[cur].activeRefs );
		if ( taskCount == 0u ) break;
		const uint32_t binCount	 = binsOfLevel( taskCount, budget );
		const uint32_t remaining = budget - min( ldVolatile( &state->totalRefs ), budget ); // (changes in the split phase only)
		stamp( 0 );

		const bool serial = taskCount >= P.serialMin; // one thread per task in the per-task phases
		resetBinsPhase( P, cur, taskCount, binCount, serial );
		if ( blockIdx.x == 0 && threadIdx.x == 0 ) state->level[cur ^ 1u] = LevelCounts{ 0u, 0u, 0u, 0u };
		HIPRT_SBVH_BARRIER();
		stamp( 1 );

		binPhase<Source, false>( P, src, cur, refCount, binCount, edgeThreshold, sBins, &sWord, serial );
		HIPRT_SBVH_BARRIER();
		stamp( 2 );

		findObjectSplitPhase( P, cur, level, forceLevel, taskCount, binCount, overlapThreshold );
		HIPRT_SBVH_BARRIER();
		stamp( 3 );

		const uint32_t spatialTasks = ldVolatile( &state->level[cur].spatialTasks );
		if ( spatialTasks != 0u )
		{
			binPhase<Source, true>( P, src, cur, refCount, binCount, edgeThreshold, sBins, &sWord, serial );
			HIPRT_SBVH_BARRIER();
		}
		stamp( 4 );

		uint32_t demand = 0;
		if ( spatialTasks != 0u )
		{
			splitPhase<Source, true>( P, cur, level, remaining, 0u, refCount, taskCount, binCount, edgeThreshold );
			HIPRT_SBVH_BARRIER();
			demand = ldVolatile( &state->level[cur].demand );
		}
		stamp( 5 );
		splitPhase<Source, false>( P, cur, level, remaining, demand, refCount, taskCount, binCount, edgeThreshold );
		HIPRT_SBVH_BARRIER();
		stamp( 6 );

		distributePhase<Source>( P, src, cur, refCount, binCount );
		HIPRT_SBVH_BARRIER();
		stamp( 7 );
		++level;
	}

#undef HIPRT_SBVH_BARRIER

	// hand over to the collapse: leaf count, root, first task
	if ( blockIdx.x == 0 && threadIdx.x == 0 )
	{
		state->levels	 = level;
		P.ctl->primCount = ldVolatile( &state->leafCount );
		P.ctl->rootNode	 = 0u;
		P.taskQueue[0]	 = ( uint64_t( InvalidValue ) << 32 ) | makeNodeIndex( 0u, BoxType );
		if ( ldVolatile( &state->error ) != 0u ) P.ctl->pad[0] = 1u; // build error flag
	}
}
} // namespace sbvh
} // namespace hiprt
