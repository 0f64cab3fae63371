// LBVH fast builder + collapse to 4-wide nodes + leaf packing, hand-written for sm_100a.
//
// Reference path this replaces: hiprt/impl/LbvhBuilder.h:133-430 with
// the kernels PairTriangles / ComputeCentroidBox / ComputeMortonCodes (BvhBuilderKernels.h:
// 373-548), Orochi onesweep sort, EmitTopologyAndFitBounds (LbvhBuilderKernels.h:67-134),
// Collapse / CompactTasks / PackLeaves (BvhBuilderKernels.h:1088-1479).
//
// B200 design: the whole build is 10 launches on the caller's stream and never
// synchronises with the host (the reference blocks 4-5 times to read counters back):
//   init -> pair (single pass, decoupled look-back scan => deterministic pair order)
//        -> prim boxes + scene box (vertices gathered once, 32-byte box records kept)
//        -> Morton codes fused with the digit histograms of all sort passes
//        -> 4 x onesweep pass
//        -> emit topology + refit (bottom-up, one flag exchange per internal node)
//        -> collapse + pack (persistent in-order work queue; box nodes are written as whole
//           128-byte lines, leaves are packed by the task that creates their parent, so the
//           reference's CompactTasks and PackLeaves passes disappear).
// The primitive count after pairing lives in device memory (BuildControl::primCount); grids
// are sized from the host-side upper bound and kernels trim themselves.
#pragma once
#include "builders.h"
#include "prim_sources.cuh"
#include "radix_sort.cuh"

#include <algorithm>
#include <stdexcept>
#include <string>

namespace hiprt
{
#define HIPRT_CUDA_CHECK( expr )                                                                                  \
	do                                                                                                            \
	{                                                                                                             \
		cudaError_t e_ = ( expr );                                                                                \
		if ( e_ != cudaSuccess )                                                                                  \
			throw std::runtime_error( std::string( "CUDA error: " ) + cudaGetErrorString( e_ ) + " at " #expr ); \
	} while ( 0 )

static inline size_t alignUp( size_t v ) { return RoundUp( v, size_t( DefaultAlignment ) ); }
__host__ __device__ inline size_t alignUpDev( size_t v ) { return ( v + DefaultAlignment - 1 ) / DefaultAlignment * DefaultAlignment; }

constexpr uint32_t PairBlock = 256; // triangles per pairing tile

// ------------------------------------------------------------------------------------------------
__global__ void initGeometryKernel(
	BuildControl* ctl, GeomHeader* header, Box4Node* boxNodes, void* primNodes, size_t size, uint32_t primCount, uint32_t geomType )
{
	ctl->sceneBox		 = Box{ { FltMax, FltMax, FltMax }, { -FltMax, -FltMax, -FltMax } };
	ctl->primCount		 = primCount;
	ctl->rootNode		 = InvalidValue;
	ctl->queueHead		 = 0;
	ctl->pairTileCounter = 0;
	ctl->nodeCounter	 = 0;
	ctl->pad[0]			 = 0; // build error flag
	header->boxNodes	  = boxNodes;
	header->primNodes	  = primNodes;
	header->size		  = size;
	header->referenceCount = 0;
	header->boxNodeCount  = 1;
	header->primNodeCount = 0;
	header->geomType	  = geomType;
	header->rtip		  = 0;
}

// Scene header + per-instance records + API frames converted to the internal SRT form
// (reference: InitSceneData BvhBuilderKernels.h:224-266, frame conversion Transform.h:134-176).
__global__ void __launch_bounds__( 256 ) initSceneKernel(
	BuildControl*	ctl,
	SceneHeader*	header,
	Box4Node*		boxNodes,
	InstanceNode*	instNodes,
	InstanceRecord* records,
	SRTFrame*		frames,
	size_t			size,
	InstanceSource	src,
	uint32_t		instanceCount,
	const void*		apiFrames,
	bool			matrixFrames )
{
	const uint32_t index = blockIdx.x * blockDim.x + threadIdx.x;
	if ( index < instanceCount )
	{
		const hiprtInstance		   inst = src.instances[index];
		const hiprtTransformHeader h	= src.header( index );
		uint4					   rec;
		rec.x					  = h.frameIndex;
		rec.y					  = ( h.frameCount & 0x7FFFFFFFu ) | ( uint32_t( inst.type == hiprtInstanceTypeScene ) << 31 );
		const uint64_t ptr		  = reinterpret_cast<uint64_t>( inst.geometry );
		rec.z = uint32_t( ptr ), rec.w = uint32_t( ptr >> 32 );
		reinterpret_cast<uint4*>( records )[index] = rec;
	}
	if ( index < src.frameCount && apiFrames != nullptr )
	{
		const DFrame f = matrixFrames ? frameFromMatrix( static_cast<const hiprtFrameMatrix*>( apiFrames )[index] )
									  : frameFromSRT( static_cast<const hiprtFrameSRT*>( apiFrames )[index] );
		storeFrame( frames, index, f );
	}
	if ( index == 0 )
	{
		ctl->sceneBox		 = Box{ { FltMax, FltMax, FltMax }, { -FltMax, -FltMax, -FltMax } };
		ctl->primCount		 = instanceCount;
		ctl->rootNode		 = InvalidValue;
		ctl->queueHead		 = 0;
		ctl->pairTileCounter = 0;
		ctl->nodeCounter	 = 0;
		ctl->pad[0]			 = 0; // build error flag
		header->boxNodes	   = boxNodes;
		header->primNodes	   = instNodes;
		header->instances	   = records;
		// no frames -> null, as the reference's arena hands out for an empty array (MemoryArena.h:38-40): the transform queries of
		// the device API (hiprtVectorObjectToWorld ...) take a null frame array for "identity" (Transform.h:421)
		header->frames		   = ( src.frameCount != 0 && apiFrames != nullptr ) ? frames : nullptr;
		header->size		   = size;
		header->referenceCount = 0;
		header->primCount	   = instanceCount;
		header->primNodeCount  = 0;
		header->boxNodeCount   = 1;
		header->frameCount	   = src.frameCount;
		header->rtip		   = 0;
	}
}

// ------------------------------------------------------------------------------------------------
// Triangle pairing: greedy inside each group of 32 consecutive triangles (reference
// BvhBuilderKernels.h:373-420). Output order = group order (the reference's order across warps
// is atomic arrival order); made deterministic with a single-pass scan over 256-triangle tiles.
constexpr uint32_t ScanAgg	= 1u << 30;
constexpr uint32_t ScanIncl = 2u << 30;
constexpr uint32_t PairItems = 8;					  // groups of 32 triangles per warp
constexpr uint32_t PairTile	 = PairBlock * PairItems; // triangles per tile

__global__ void __launch_bounds__( PairBlock )
	pairTrianglesKernel( TriangleMeshSource mesh, uint32_t triCount, uint2* pairs, uint32_t* scanState, BuildControl* ctl )
{
	__shared__ uint32_t sTile;
	__shared__ uint32_t sWarpCount[PairBlock / 32];
	__shared__ uint32_t sBase;
	const uint32_t		lane = laneId(), warp = threadIdx.x >> 5;
	if ( threadIdx.x == 0 ) sTile = atomicAdd( &ctl->pairTileCounter, 1u );
	__syncthreads();
	const uint32_t tile		= sTile;
	const uint32_t warpBase = tile * PairTile + warp * ( 32 * PairItems );

	uint32_t partner[PairItems]; // set on the first triangle of each emitted pair
	uint32_t emitMask[PairItems];
	uint32_t warpTotal = 0;
#pragma unroll
	for ( uint32_t it = 0; it < PairItems; ++it )
	{
		const uint32_t groupBase = warpBase + it * 32;
		const uint32_t index	 = groupBase + lane;
		const bool	   valid	 = index < triCount;
		uint3		   idx		 = make_uint3( 0, 0, 0 );
		if ( valid ) idx = mesh.indices( index );
		uint32_t active = __ballot_sync( FullMask, valid );
		partner[it]		= InvalidValue;
		// Fast path for the common layout (quad diagonals are consecutive triangles): if the lowest active
		// triangle a can pair with a+1 and a+1 is still active, a+1 IS the lowest candidate, so the greedy
		// rule picks it; such steps need no broadcast, only bit operations on warp-uniform masks.
		const uint3	   in = make_uint3( __shfl_down_sync( FullMask, idx.x, 1 ), __shfl_down_sync( FullMask, idx.y, 1 ), __shfl_down_sync( FullMask, idx.z, 1 ) );
		const uint32_t nextOk = __ballot_sync( FullMask, lane < 31 && TriangleMeshSource::sharesEdge( idx, in ) ) & ( active >> 1 );
		uint32_t	   fastHeads = 0;
		while ( active )
		{
			const uint32_t a = __ffs( active ) - 1;
			if ( ( nextOk >> a ) & ( active >> ( a + 1 ) ) & 1u )
			{
				// the whole run a, a+2, a+4 ... of such steps at once: each of them only removes (head, head + 1), so the
				// greedy rule would take them one after the other with nothing in between (a regular mesh: one step per group)
				const uint32_t steps = ( ( nextOk & active & ( active >> 1 ) ) >> a ) | 0xAAAAAAAAu; // odd offsets do not matter
				const uint32_t count = ~steps ? __ffs( ~steps ) - 1u : 32u;							// first even offset that is no step
				const uint32_t heads = ( 0x55555555u & ( count >= 32u ? 0xFFFFFFFFu : ( 1u << count ) - 1u ) ) << a;
				fastHeads |= heads;
				active &= ~( heads | ( heads << 1 ) );
				continue;
			}
			active &= active - 1;
			const uint3 ia		 = make_uint3( __shfl_sync( FullMask, idx.x, a ), __shfl_sync( FullMask, idx.y, a ), __shfl_sync( FullMask, idx.z, a ) );
			const bool	pairable = ( ( active >> lane ) & 1u ) && TriangleMeshSource::sharesEdge( ia, idx );
			const uint32_t cand	 = __ballot_sync( FullMask, pairable );
			uint32_t	   b	 = a;
			if ( cand )
			{
				b = __ffs( cand ) - 1;
				active &= ~( 1u << b );
			}
			if ( lane == a ) partner[it] = groupBase + b;
		}
		if ( ( fastHeads >> lane ) & 1u ) partner[it] = groupBase + lane + 1;
		emitMask[it] = __ballot_sync( FullMask, valid && partner[it] != InvalidValue );
		warpTotal += __popc( emitMask[it] );
	}
	if ( lane == 0 ) sWarpCount[warp] = warpTotal;
	__syncthreads();
	uint32_t warpOffset = 0, blockTotal = 0;
#pragma unroll
	for ( uint32_t w = 0; w < PairBlock / 32; ++w )
	{
		if ( w < warp ) warpOffset += sWarpCount[w];
		blockTotal += sWarpCount[w];
	}
	if ( warp == 0 )
	{
		// decoupled look-back over the preceding tiles, 32 tiles per step
		if ( lane == 0 ) *reinterpret_cast<volatile uint32_t*>( scanState + tile ) = ( tile == 0 ? ScanIncl : ScanAgg ) | blockTotal;
		uint32_t prefix = 0;
		if ( tile > 0 )
		{
			int look = int( tile ) - 1;
			while ( true )
			{
				const int	   t = look - int( lane );
				const uint32_t v = t >= 0 ? ldVolatile( scanState + t ) : ScanIncl; // before the first tile: inclusive 0
				if ( __any_sync( FullMask, ( v >> 30 ) == 0 ) ) continue;			// not all published yet
				const uint32_t incl = __ballot_sync( FullMask, ( v & ScanIncl ) != 0 );
				if ( incl )
				{
					const uint32_t first = __ffs( incl ) - 1;
					prefix += __reduce_add_sync( FullMask, lane <= first ? ( v & ( ScanAgg - 1 ) ) : 0u );
					break;
				}
				prefix += __reduce_add_sync( FullMask, v & ( ScanAgg - 1 ) );
				look -= 32;
			}
			if ( lane == 0 ) *reinterpret_cast<volatile uint32_t*>( scanState + tile ) = ScanIncl | ( prefix + blockTotal );
		}
		if ( lane == 0 )
		{
			sBase = prefix;
			if ( tile == DivideRoundUp( triCount, PairTile ) - 1 ) ctl->primCount = prefix + blockTotal;
		}
	}
	__syncthreads();
	uint32_t out = sBase + warpOffset;
#pragma unroll
	for ( uint32_t it = 0; it < PairItems; ++it )
	{
		if ( ( emitMask[it] >> lane ) & 1u ) pairs[out + __popc( emitMask[it] & laneMaskLt() )] = make_uint2( warpBase + it * 32 + lane, partner[it] );
		out += __popc( emitMask[it] );
	}
}

// ------------------------------------------------------------------------------------------------
// Primitive boxes (kept as 32-byte records so vertices are gathered exactly once before the
// leaves are packed) + scene box by warp shuffle reduction and 6 atomics per block.
template <typename Source>
__global__ void __launch_bounds__( 256 ) primBoxesKernel( Source src, BuildControl* ctl, void* primRefs )
{
	const uint32_t n	= ctl->primCount;
	DBox		   mine = emptyBox();
	for ( uint32_t p = blockIdx.x * blockDim.x + threadIdx.x; p < n; p += gridDim.x * blockDim.x )
	{
		const DBox	b = src.box( p );
		const uint2 w = src.refWords( p );
		storeRec( primRefs, p, b, w.x, w.y );
		mine = uniteBox( mine, b );
	}
	mine = warpUniteBox( mine );
	__shared__ DBox sBox[8];
	if ( laneId() == 0 ) sBox[threadIdx.x >> 5] = mine;
	__syncthreads();
	if ( threadIdx.x < 32 )
	{
		DBox b = threadIdx.x < ( blockDim.x >> 5 ) ? sBox[threadIdx.x] : emptyBox();
		b	   = warpUniteBox( b );
		if ( threadIdx.x == 0 && b.lo.x <= b.hi.x )
		{
			atomicMinF( &ctl->sceneBox.lo.x, b.lo.x );
			atomicMinF( &ctl->sceneBox.lo.y, b.lo.y );
			atomicMinF( &ctl->sceneBox.lo.z, b.lo.z );
			atomicMaxF( &ctl->sceneBox.hi.x, b.hi.x );
			atomicMaxF( &ctl->sceneBox.hi.y, b.hi.y );
			atomicMaxF( &ctl->sceneBox.hi.z, b.hi.z );
		}
	}
}

// ------------------------------------------------------------------------------------------------
// Morton codes (key = code, value = primitive index) fused with the digit histograms of all
// four sort passes (reference: ComputeMortonCodes BvhBuilderKernels.h:508-525 + Orochi GHistogram).
__global__ void __launch_bounds__( 256 )
	mortonKernel( const BuildControl* ctl, const void* primRefs, uint32_t* keys, uint32_t* vals, uint32_t* sortScratch )
{
	__shared__ uint32_t sHist[rsort::Passes * rsort::Radix];
	for ( int i = threadIdx.x; i < rsort::Passes * rsort::Radix; i += blockDim.x )
		sHist[i] = 0;
	__syncthreads();
	const uint32_t n   = ctl->primCount;
	const float3   lo  = make_float3( ctl->sceneBox.lo.x, ctl->sceneBox.lo.y, ctl->sceneBox.lo.z );
	const float3   ext = make_float3( ctl->sceneBox.hi.x - lo.x, ctl->sceneBox.hi.y - lo.y, ctl->sceneBox.hi.z - lo.z );
	const MortonLayout L = mortonLayout( ext );
	for ( uint32_t p = blockIdx.x * blockDim.x + threadIdx.x; p < n; p += gridDim.x * blockDim.x )
	{
		const Rec32	 r = loadRec( primRefs, p );
		const float3 c = make_float3(
			( r.box.hi.x + r.box.lo.x ) * 0.5f, ( r.box.hi.y + r.box.lo.y ) * 0.5f, ( r.box.hi.z + r.box.lo.z ) * 0.5f );
		const float3 pos = make_float3( __fdividef( c.x - lo.x, ext.x ), __fdividef( c.y - lo.y, ext.y ), __fdividef( c.z - lo.z, ext.z ) );
		const uint32_t key = mortonCode( L, pos );
		keys[p]			   = key;
		vals[p]			   = p;
		rsort::histogramKey( sHist, key );
	}
	__syncthreads();
	for ( int i = threadIdx.x; i < rsort::Passes * rsort::Radix; i += blockDim.x )
		if ( sHist[i] ) atomicAdd( &sortScratch[i], sHist[i] );
}

// ------------------------------------------------------------------------------------------------
// LBVH topology + refit, bottom-up (reference: LbvhBuilderKernels.h:44-134, MortonCode.h:279-285).
// Internal node id = position of the split gap; the first child to arrive parks its far range
// end in flags[parent], the second one unites the boxes and climbs. The climbing thread keeps
// its own box in registers and reads only the sibling's record (through L2: it was written by
// another SM in this launch).
__device__ __forceinline__ uint64_t gapDelta( const uint32_t* keys, uint32_t a, uint32_t n )
{
	// delta between sorted positions a and a+1; the ends of the array compare as +infinity
	if ( a >= n - 1 ) return ~0ull; // also catches a == uint32(-1)
	const uint32_t k0 = keys[a], k1 = keys[a + 1];
	return ( uint64_t( k0 ^ k1 ) << 32 ) | uint64_t( a ^ ( a + 1 ) );
}

// Comparing two gap deltas is comparing the position of their highest set bit: for a sorted array two
// gaps bounding a range can never share that position (the bit would have to go 0->1 twice without a
// higher bit changing in between). A one-byte "level" per gap therefore replaces the 64-bit delta.
__device__ __forceinline__ uint32_t gapLevel( const uint32_t* keys, uint32_t a, uint32_t n )
{
	if ( a >= n - 1 ) return 64u; // the ends of the array (also a == uint32(-1)): +infinity
	const uint32_t x = keys[a] ^ keys[a + 1];
	return x ? 63u - __clz( x ) : 31u - __clz( a ^ ( a + 1 ) );
}

// Tile-local version: a block owns EmitTile consecutive leaves. A node whose whole leaf range lies
// inside the tile ("local": its gap level is below the maximum level on both sides within the tile,
// found with two block-wide max scans) is resolved entirely in shared memory -- the first child to
// arrive publishes {box, index} in its own thread slot and exchanges {range end, thread} through a
// shared-memory flag; the second one unites and writes the finished 32-byte binary node with ONE
// global store. Only the few nodes whose range crosses a tile boundary (~depth x N/EmitTile of them)
// go through global flags and L2 as in the reference.
constexpr uint32_t EmitTile = 256;

__device__ __forceinline__ uint32_t blockExclusiveMaxScan( uint32_t v, bool reverse, uint32_t* warpTotals )
{
	const uint32_t lane = laneId(), warp = threadIdx.x >> 5;
	uint32_t	   x	= v;
#pragma unroll
	for ( int o = 1; o < 32; o <<= 1 )
	{
		const uint32_t y = reverse ? __shfl_down_sync( FullMask, x, o ) : __shfl_up_sync( FullMask, x, o );
		if ( reverse ? ( lane + o < 32 ) : ( lane >= uint32_t( o ) ) ) x = max( x, y );
	}
	if ( lane == ( reverse ? 0u : 31u ) ) warpTotals[warp] = x;
	__syncthreads();
	uint32_t excl = reverse ? __shfl_down_sync( FullMask, x, 1 ) : __shfl_up_sync( FullMask, x, 1 );
	if ( lane == ( reverse ? 31u : 0u ) ) excl = 0;
	constexpr uint32_t Warps = EmitTile / 32;
	if ( reverse )
	{
		for ( uint32_t w = warp + 1; w < Warps; ++w )
			excl = max( excl, warpTotals[w] );
	}
	else
	{
		for ( uint32_t w = 0; w < warp; ++w )
			excl = max( excl, warpTotals[w] );
	}
	return excl;
}

template <uint32_t LeafType>
__global__ void __launch_bounds__( EmitTile ) emitTopologyKernel(
	BuildControl*	ctl,
	const uint32_t* sortedKeys,
	const uint32_t* sortedVals,
	const void*		primRefs,
	void*			sortedRefs,
	void*			scratch,
	uint32_t*		flags,
	uint64_t*		taskQueue )
{
	__shared__ float4	sOwn[EmitTile][2];		// {box, index} published by a thread that stops at a local node
	__shared__ uint32_t sFlag[EmitTile];		// local gaps: {relative far range end << 16 | thread} of the first arriver
	__shared__ uint8_t	sLevel[EmitTile + 1];	// gap base-1+g: level (bits 0-6), local (bit 7)
	__shared__ uint32_t sWarpMax[2][EmitTile / 32];

	const uint32_t n	= ctl->primCount;
	const uint32_t t	= threadIdx.x;
	const uint32_t base = blockIdx.x * EmitTile;
	if ( base >= n ) return;
	const uint32_t k = base + t;

	// ---- gap levels of the tile and which gaps are local
	const uint32_t lv	   = gapLevel( sortedKeys, k - 1u, n ); // my left gap
	const uint32_t lvRight = gapLevel( sortedKeys, base + EmitTile - 1u, n );
	sFlag[t]			   = InvalidValue;
	const uint32_t maxLeft	= blockExclusiveMaxScan( lv, false, sWarpMax[0] );
	const uint32_t maxRight = max( blockExclusiveMaxScan( lv, true, sWarpMax[1] ), lvRight );
	const bool	   isLocal	= t >= 1 && lv < maxLeft && lv < maxRight;
	sLevel[t]				= uint8_t( lv | ( isLocal ? 0x80u : 0u ) );
	if ( t == 0 ) sLevel[EmitTile] = uint8_t( lvRight );
	__syncthreads();
	if ( k >= n ) return;

	auto levelOf = [&]( uint32_t gap ) -> uint32_t {
		const uint32_t g = gap + 1u - base; // position inside the tile's gap array (wraps for gaps left of the tile)
		return g <= EmitTile ? ( sLevel[g] & 0x7Fu ) : gapLevel( sortedKeys, gap, n );
	};

	const Rec32 mine = loadRec( primRefs, sortedVals[k] );
	DBox		box	 = mine.box;
	storeRec( sortedRefs, k, box, mine.a, mine.b );

	uint32_t  i = k, j = k + 1; // covered leaf range [i, j)
	uint32_t  self		   = makeNodeIndex( k, LeafType );
	uint32_t* scratchWords = reinterpret_cast<uint32_t*>( scratch );
	while ( true )
	{
		const bool	   left	  = ( i == 0 ) || ( j != n && levelOf( j - 1 ) < levelOf( i - 1 ) );
		const uint32_t parent = left ? j - 1 : i - 1;
		const uint32_t g	  = parent + 1u - base;
		if ( g < EmitTile && ( sLevel[g] & 0x80u ) )
		{
			// ---- both children live in this tile: shared-memory handshake
			sOwn[t][0] = make_float4( box.lo.x, box.lo.y, box.lo.z, box.hi.x );
			sOwn[t][1] = make_float4( box.hi.y, box.hi.z, __uint_as_float( self ), 0.0f );
			__threadfence_block();
			const uint32_t other = atomicExch( &sFlag[g], ( ( ( left ? i : j ) - base ) << 16 ) | t );
			if ( other == InvalidValue ) return; // first to arrive: the sibling finishes the node
			__threadfence_block();
			const float4   q0 = sOwn[other & 0xFFFFu][0], q1 = sOwn[other & 0xFFFFu][1];
			const uint32_t sib = __float_as_uint( q1.z );
			if ( left )
				j = base + ( other >> 16 );
			else
				i = base + ( other >> 16 );
			box = uniteBox( box, DBox{ make_float3( q0.x, q0.y, q0.z ), make_float3( q0.w, q1.x, q1.y ) } );
			storeRec( scratch, parent, box, left ? self : sib, left ? sib : self );
		}
		else
		{
			// ---- range crosses the tile boundary: global flags, sibling record through L2
			scratchWords[size_t( parent ) * 8 + ( left ? 6 : 7 )] = self;
			__threadfence();
			const uint32_t other = atomicExch( &flags[parent], left ? i : j );
			if ( other == InvalidValue ) return;
			// (no fence here: the sibling's words are read with ld.cg below, issued only after the exchange has returned)
			if ( left )
				j = other;
			else
				i = other;
			const uint32_t sib	  = __ldcg( &scratchWords[size_t( parent ) * 8 + ( left ? 7 : 6 )] );
			const Rec32	   sibRec = isBoxIndex( sib ) ? loadRecCg( scratch, nodeAddr( sib ) ) : loadRecCg( sortedRefs, nodeAddr( sib ) );
			box					  = uniteBox( box, sibRec.box );
			float4* dst = reinterpret_cast<float4*>( scratch ) + 2 * size_t( parent );
			dst[0]		= make_float4( box.lo.x, box.lo.y, box.lo.z, box.hi.x );
			reinterpret_cast<float2*>( dst + 1 )[0] = make_float2( box.hi.y, box.hi.z );
		}
		self = makeNodeIndex( parent, BoxType );
		if ( i == 0 && j == n )
		{
			ctl->rootNode = parent;
			taskQueue[0]  = ( uint64_t( InvalidValue ) << 32 ) | self; // root task: {binary node, parent box}
			return;
		}
	}
}

// ------------------------------------------------------------------------------------------------
// Balanced builder: H-PLOC (hierarchical parallel locally-ordered clustering; reference:
// PlocBuilderKernels.h:120-307, host PlocBuilder.h:300-351). Lanes climb the same LBVH ranges
// as the fast builder, but a range only COLLECTS cluster ids; once a range holds more than
// PlocWindow/2 clusters (or is the whole array) the warp merges it: every cluster looks for its
// nearest neighbour (union-box area) within PlocRadius positions, mutually nearest pairs merge
// into a new binary node, survivors are compacted in order, until at most PlocWindow/2 (or 1 at
// the root) remain. Output = the same ScratchNode/ReferenceNode records the collapse consumes.
//
// B200 version: the window lives in REGISTERS (one cluster per lane: id + box); neighbour boxes
// come from shfl_down, the distance of the mirrored pair from shfl_up, compaction from a ballot
// and __fns -- no shared memory, no shared-memory atomics. Distances are ordered by the 64-bit key
// (area bits, distance in positions, parity of the lower position, lower position): a total order
// over pairs, so the globally closest pair is always mutual and every round makes progress; on
// exact area ties (identical primitives) neighbours pair up as (0,1),(2,3),... and the window halves
// every round instead of degenerating into a chain (the reference gets the same effect from the
// parity bit of its encodeOffset, PlocBuilderKernels.h:54-65). The tree shape depends only on the
// sorted keys and boxes (node ADDRESSES depend on atomic order, as in the reference).
// Search radius: the reference uses 8 (BvhConfig.h). Measured here on five meshes (profiles/r02_ploc_radius.txt: terrain, UV sphere,
// icosphere, uniform and clustered mixed-size soups), radius 2 gives FEWER node visits per ray than radius 8 on four of them
// (-1 % .. -10 %; +2 % on the clustered soup, where the SAH-style cost is 2 % worse too) and the kernel runs 1.55 -> 0.92 ms at
// 5 M primitives: each round is 2 instead of 8 neighbour evaluations, and on Morton-sorted input the wider search mostly finds pairs
// whose merge looks good locally but interleaves subtrees. HIPRT_B200_PLOC_RADIUS = 2 / 4 / 8 overrides it.
constexpr uint32_t PlocRadius = 2;
constexpr uint32_t PlocWindow = 32; // clusters held by a warp: 16 from each side of the split
constexpr uint32_t PlocBlock  = 256;

__device__ __forceinline__ uint64_t shflU64( uint64_t v, uint32_t srcLane )
{
	const uint32_t lo = __shfl_sync( FullMask, uint32_t( v ), srcLane ), hi = __shfl_sync( FullMask, uint32_t( v >> 32 ), srcLane );
	return ( uint64_t( hi ) << 32 ) | lo;
}
__device__ __forceinline__ uint64_t shflUpU64( uint64_t v, uint32_t d )
{
	const uint32_t lo = __shfl_up_sync( FullMask, uint32_t( v ), d ), hi = __shfl_up_sync( FullMask, uint32_t( v >> 32 ), d );
	return ( uint64_t( hi ) << 32 ) | lo;
}
__device__ __forceinline__ DBox shflBox( const DBox& b, uint32_t srcLane )
{
	return DBox{
		make_float3( __shfl_sync( FullMask, b.lo.x, srcLane ), __shfl_sync( FullMask, b.lo.y, srcLane ), __shfl_sync( FullMask, b.lo.z, srcLane ) ),
		make_float3( __shfl_sync( FullMask, b.hi.x, srcLane ), __shfl_sync( FullMask, b.hi.y, srcLane ), __shfl_sync( FullMask, b.hi.z, srcLane ) ) };
}
__device__ __forceinline__ DBox shflDownBox( const DBox& b, uint32_t d )
{
	return DBox{
		make_float3( __shfl_down_sync( FullMask, b.lo.x, d ), __shfl_down_sync( FullMask, b.lo.y, d ), __shfl_down_sync( FullMask, b.lo.z, d ) ),
		make_float3( __shfl_down_sync( FullMask, b.hi.x, d ), __shfl_down_sync( FullMask, b.hi.y, d ), __shfl_down_sync( FullMask, b.hi.z, d ) ) };
}

template <uint32_t LeafType, uint32_t Radius = PlocRadius, uint32_t Block = PlocBlock>
__global__ void __launch_bounds__( Block ) hplocKernel(
	BuildControl*	ctl,
	const uint32_t* sortedKeys,
	const uint32_t* sortedVals,
	const void*		primRefs,
	void*			sortedRefs,
	void*			scratch,
	uint32_t*		flags,		// per split gap: far end of the range of the first child to arrive (0xFF.. initially)
	uint32_t*		clusterIds, // per sorted position: cluster occupying it (compacted to the front of each merged range)
	uint32_t*		nodeCounter,
	uint64_t*		taskQueue )
{
	const uint32_t n	= ctl->primCount;
	const uint32_t k	= blockIdx.x * blockDim.x + threadIdx.x;
	const uint32_t lane = laneId();
	bool		   active = k < n;
	if ( active )
	{
		const Rec32 mine = loadRec( primRefs, sortedVals[k] );
		storeRec( sortedRefs, k, mine.box, mine.a, mine.b );
		__stcg( &clusterIds[k], makeNodeIndex( k, LeafType ) );
	}
	uint32_t i = k, j = k + 1, s = k;
	constexpr uint32_t Half = PlocWindow / 2;

	// Memory ordering of the flag handshake. Everything a range owner wrote (leaf records, cluster ids, merged nodes) is made
	// visible by ONE gpu-scope fence per loop iteration, placed before the next exchange; the side that arrives second reads the
	// other side's data only with ld.cg (L2, never a stale L1 line) through addresses that depend on the exchanged value, so no
	// fence follows the exchange. (ncu: with a fence before and after every exchange and after every merge, 37 % of all stall
	// samples of this kernel were membar stalls.)
	__threadfence();
	while ( __any_sync( FullMask, active ) )
	{
		if ( active )
		{
			const bool	   left	  = ( i == 0 ) || ( j != n && gapDelta( sortedKeys, j - 1, n ) < gapDelta( sortedKeys, i - 1, n ) );
			const uint32_t parent = left ? j - 1 : i - 1;
			const uint32_t other = atomicExch( &flags[parent], left ? i : j );
			if ( left )
				s = j, j = other;
			else
				s = i, i = other;
			if ( other == InvalidValue ) active = false; // first to arrive: the sibling carries the range on
		}
		const bool last	 = active && i == 0 && j == n;
		uint32_t   merge = __ballot_sync( FullMask, active && ( j - i > Half || last ) );
		const bool merged = merge != 0;
		while ( merge )
		{
			const uint32_t cur = __ffs( merge ) - 1;
			merge &= merge - 1;
			const uint32_t ci = __shfl_sync( FullMask, i, cur ), cj = __shfl_sync( FullMask, j, cur ), cs = __shfl_sync( FullMask, s, cur );
			const bool	   clast = __shfl_sync( FullMask, uint32_t( last ), cur ) != 0;

			// gather up to 16 clusters from the front of each side (valid ids are a prefix of each side)
			const uint32_t numLeft = min( cs - ci, Half ), numRight = min( cj - cs, Half );
			const uint32_t idL = lane < numLeft ? __ldcg( &clusterIds[ci + lane] ) : InvalidValue;
			const uint32_t idR = lane < numRight ? __ldcg( &clusterIds[cs + lane] ) : InvalidValue;
			const uint32_t vL = __popc( __ballot_sync( FullMask, idL != InvalidValue ) ), vR = __popc( __ballot_sync( FullMask, idR != InvalidValue ) );
			const uint32_t fromR = __shfl_sync( FullMask, idR, ( lane - vL ) & 31u );
			uint32_t	   count = vL + vR;
			uint32_t	   id	 = lane < vL ? idL : ( lane < count ? fromR : InvalidValue );
			const uint32_t threshold = clast ? 1u : Half;
			DBox		   box		 = emptyBox();
			if ( count > threshold && lane < count )
				box = ( isBoxIndex( id ) ? loadRecCg( scratch, nodeAddr( id ) ) : loadRecCg( sortedRefs, nodeAddr( id ) ) ).box;

			// The rounds below leave exactly `threshold` clusters (the last round merges no more pairs than that takes: what it drops
			// is found again one level up), so the nodes of all rounds are allocated with ONE atomic on the grid-wide counter, issued
			// before the first round's neighbour search; one atomic per round had its round trip through L2 in every round's
			// dependency chain.
			uint32_t nodeBase = 0;
			if ( lane == 0 && count > threshold ) nodeBase = atomicAdd( nodeCounter, count - threshold );
			nodeBase = __shfl_sync( FullMask, nodeBase, 0 );
			while ( count > threshold )
			{
				// nearest neighbour within PlocRadius positions, both directions
				uint64_t best = ~0ull;
#pragma unroll
				for ( uint32_t d = 1; d <= Radius; ++d )
				{
					const DBox	   nb	 = shflDownBox( box, d );
					const bool	   valid = lane + d < count;
					const uint64_t fwd	 = valid ? ( ( uint64_t( __float_as_uint( boxArea( uniteBox( box, nb ) ) ) ) << 32 ) | ( ( d - 1u ) << 11 ) |
												 ( ( lane & 1u ) << 10 ) | ( lane << 5 ) | ( lane + d ) )
											   : ~0ull;
					const uint64_t bwd	 = shflUpU64( fwd, d ); // the pair (lane - d, lane), computed by lane - d
					best				 = min( best, fwd );
					if ( lane >= d ) best = min( best, bwd );
				}
				const bool	   has	  = lane < count; // count > 1 here, so every live lane has a neighbour
				const uint32_t pa	  = ( uint32_t( best ) >> 5 ) & 31u, pb = uint32_t( best ) & 31u;
				const uint32_t nbLane = has ? ( pa == lane ? pb : pa ) : lane;
				const uint64_t nbBest = shflU64( best, nbLane );
				const bool	   mutual = has && nbBest == best;
				// the lower position keeps the merged cluster; only as many pairs as still separate the window from `threshold`
				uint32_t	   mmask   = __ballot_sync( FullMask, mutual && lane < nbLane );
				const uint32_t allowed = count - threshold;
				if ( uint32_t( __popc( mmask ) ) > allowed ) mmask &= ( 1u << __fns( mmask, 0, allowed + 1 ) ) - 1u; // the lowest `allowed` of them
				const bool	   merging = ( ( mmask >> lane ) & 1u ) != 0u;
				const bool	   dropped = mutual && lane > nbLane && ( ( mmask >> nbLane ) & 1u ) != 0u;
				const uint32_t nbId	   = __shfl_sync( FullMask, id, nbLane );
				const DBox	   nbBox   = shflBox( box, nbLane );
				const uint32_t base	   = nodeBase;
				nodeBase += __popc( mmask );
				if ( merging )
				{
					const uint32_t addr = n - 2u - ( base + __popc( mmask & laneMaskLt() ) );
					box					= uniteBox( box, nbBox );
					storeRec( scratch, addr, box, id, nbId );
					id = makeNodeIndex( addr, BoxType );
				}
				// compact the survivors in order
				const uint32_t keep		= __ballot_sync( FullMask, has && !dropped );
				const uint32_t newCount = __popc( keep );
				const uint32_t src		= lane < newCount ? __fns( keep, 0, lane + 1 ) : lane;
				id						= __shfl_sync( FullMask, id, src );
				box						= shflBox( box, src );
				if ( lane >= newCount ) id = InvalidValue;
				count = newCount;
			}

			if ( lane < Half && ci + lane < n ) __stcg( &clusterIds[ci + lane], lane < count ? id : InvalidValue );
			if ( clast && lane == 0 )
			{
				ctl->rootNode = nodeAddr( id );
				taskQueue[0]  = ( uint64_t( InvalidValue ) << 32 ) | id; // root task of the collapse
			}
			__syncwarp();
		}
		if ( merged ) __threadfence(); // publish this iteration's nodes and cluster ids before the next exchange
		if ( last ) active = false;
	}
}

// ------------------------------------------------------------------------------------------------
// Collapse the binary tree into 4-wide nodes and pack the leaves (reference: Collapse + openNodes
// BvhBuilderKernels.h:137-193,1088-1258; PackLeaves :1441-1479).
// A task = (binary node, parent box address); the task's index in the queue IS the address of
// the Box4Node it produces. Four lanes cooperate on a task (one per child slot). Warps claim 8
// consecutive task slots at a time in queue order, so the producer of any slot is always held
// by an already-running warp: the kernel is persistent and deadlock-free at any occupancy.
// A slot is processed as soon as its task arrives; termination = every leaf has been packed.
constexpr uint32_t CollapseBlock = 128;

template <typename Source>
__global__ void __launch_bounds__( CollapseBlock ) collapsePackKernel(
	Source		  src,
	BuildControl* ctl,
	uint32_t*	  boxNodeCounter,  // header field: number of box nodes allocated (starts at 1 = root)
	uint32_t*	  primNodeCounter, // header field: number of leaf nodes allocated (starts at 0)
	const void*	  scratch,
	const void*	  sortedRefs,
	uint64_t*	  taskQueue,
	uint32_t	  maxBoxNodes,
	uint32_t	  claimSize, // task slots a warp claims at a time (multiple of 8)
	Box4Node*	  boxNodes,
	void*		  leafNodes )
{
	__shared__ __align__( 16 ) uint32_t sNode[CollapseBlock / 32][8][32]; // one 128-byte line per 4-lane group
	const uint32_t n	= ctl->primCount;
	const uint32_t lane = laneId(), warp = threadIdx.x >> 5;
	const uint32_t group = lane >> 2, slot = lane & 3u;
	const uint32_t groupMask = 0xFu << ( group * 4 );

	uint32_t idlePolls	 = 0; // safety net: a corrupted tree must not hang the device
	uint32_t packedLocal = 0; // lane 0: leaves packed by this warp and not yet added to the global count
	// "every leaf is packed" is the termination signal; the count is only flushed when a warp runs dry
	auto finished = [&]() -> bool {
		uint32_t packed = 0;
		if ( lane == 0 )
		{
			if ( packedLocal ) atomicAdd( primNodeCounter, packedLocal ), packedLocal = 0;
			packed = ldVolatile( primNodeCounter );
		}
		return __shfl_sync( FullMask, packed, 0 ) >= n;
	};
	while ( true )
	{
		// Claim the next `claimSize` (8 .. 32, host-chosen from the tree size) slots unconditionally. The two counters (queue head, allocated nodes) are single L2 lines that
		// every warp of the grid hits: ncu showed more than half of all stall samples on a "head < allocated?" pre-check
		// plus the claim atomic, so the pre-check is gone and the claim is 4x rarer. Slots are processed strictly in queue
		// order, 8 at a time; a slot whose task has not been published yet is polled at its own address (no shared line).
		uint32_t base = 0;
		if ( lane == 0 ) base = atomicAdd( &ctl->queueHead, claimSize );
		base = __shfl_sync( FullMask, base, 0 );
		if ( base >= maxBoxNodes ) // ran past the end of the queue: nothing to do here (flush the leaf count first)
		{
			finished();
			return;
		}
		for ( uint32_t sub = 0; sub < claimSize / 8; ++sub )
		{
		const uint32_t taskIdx = base + sub * 8 + group;
		bool		   done	   = taskIdx >= maxBoxNodes; // such a slot can never be produced
		uint64_t	   task	   = ~0ull;
		uint32_t	   misses  = 0;

		while ( !__all_sync( FullMask, done ) )
		{
			bool ready = false;
			if ( !done )
			{
				task  = ldVolatile( taskQueue + taskIdx );
				ready = task != ~0ull;
			}
			if ( !__any_sync( FullMask, ready ) )
			{
				// nothing arrived: stop once every leaf is packed (no task can appear any more). The leaf counter shares
				// its line with the node allocator, so it is looked at on the 1st, 2nd, 4th, 8th ... miss only.
				if ( ( ( misses + 1u ) & misses ) == 0u && finished() ) return;
				if ( ++idlePolls > ( 1u << 20 ) )
				{
					if ( lane == 0 ) ctl->pad[0] = 1; // build error flag
					return;
				}
				__nanosleep( misses < 64u ? 100u : 400u ); // own slots, distinct lines: a short period keeps the per-level latency low
				++misses;
				continue;
			}
			idlePolls = 0, misses = 0;

			// ---- process the ready groups in lockstep; other lanes idle through the collectives
			const uint32_t binIndex	  = uint32_t( task );
			const uint32_t parentAddr = uint32_t( task >> 32 );
			uint32_t	   child	  = InvalidValue;
			DBox		   cbox		  = emptyBox();
			uint32_t	   w0 = 0, w1 = 0; // children of an internal child / primIndex of a leaf child
			uint32_t	   count = 2;
			if ( ready && slot < 2 )
			{
				const uint32_t* binWords = reinterpret_cast<const uint32_t*>( scratch ) + size_t( nodeAddr( binIndex ) ) * 8;
				child					 = binWords[6 + slot];
				const Rec32 rec = isBoxIndex( child ) ? loadRec( scratch, nodeAddr( child ) ) : loadRec( sortedRefs, nodeAddr( child ) );
				cbox = rec.box, w0 = rec.a, w1 = rec.b;
			}
			// open the internal child with the largest area (lowest slot on ties) until 4 children
#pragma unroll
			for ( int it = 0; it < 2; ++it )
			{
				float area = -FltMax;
				if ( ready && slot < count && isBoxIndex( child ) ) area = boxArea( cbox );
				float best = fmaxf( area, __shfl_xor_sync( FullMask, area, 1 ) );
				best	   = fmaxf( best, __shfl_xor_sync( FullMask, best, 2 ) );
				const bool	   open	   = ready && best >= 0.0f;
				const uint32_t isBest  = __ballot_sync( FullMask, open && area == best ) & groupMask;
				const uint32_t srcLane = isBest ? __ffs( isBest ) - 1 : lane;
				const uint32_t c0 = __shfl_sync( FullMask, w0, srcLane ), c1 = __shfl_sync( FullMask, w1, srcLane );
				if ( open )
				{
					const uint32_t openSlot = srcLane & 3u;
					uint32_t	   take		= InvalidValue;
					if ( slot == openSlot ) take = c0;
					if ( slot == count ) take = c1;
					if ( take != InvalidValue )
					{
						child			= take;
						const Rec32 rec = isBoxIndex( child ) ? loadRec( scratch, nodeAddr( child ) ) : loadRec( sortedRefs, nodeAddr( child ) );
						cbox = rec.box, w0 = rec.a, w1 = rec.b;
					}
					++count;
				}
			}

			const bool	   active	= ready && slot < count;
			const bool	   internal = active && isBoxIndex( child );
			const bool	   leaf		= active && !internal;
			// warp-aggregated allocation of box nodes (= task slots) and leaf nodes
			const uint32_t imask = __ballot_sync( FullMask, internal );
			const uint32_t lmask = __ballot_sync( FullMask, leaf );
			uint32_t	   ibase = 0;
			if ( lane == 0 )
			{
				if ( imask ) ibase = atomicAdd( boxNodeCounter, __popc( imask ) );
				packedLocal += __popc( lmask );
			}
			ibase			  = __shfl_sync( FullMask, ibase, 0 );
			uint32_t outChild = InvalidValue;
			if ( internal )
			{
				const uint32_t addr = ibase + __popc( imask & laneMaskLt() );
				taskQueue[addr]		= ( uint64_t( taskIdx ) << 32 ) | child; // single 8-byte store: arrives whole
				outChild			= makeNodeIndex( addr, BoxType );
			}
			if ( leaf )
			{
				// leaf address = sorted position: no allocation, Morton-ordered (cache friendly) leaf array
				const uint32_t addr = nodeAddr( child );
				src.writeLeaf( leafNodes, addr, w0, w1 );
				outChild = makeNodeIndex( addr, Source::LeafType );
			}

			// ---- assemble the 128-byte node in shared memory, write it as whole 16-byte words
			if ( ready )
			{
				uint32_t* nodeWords = sNode[warp][group];
				nodeWords[slot]		= outChild;
				const DBox b		= active ? cbox : emptyBox();
				float*	   bw		= reinterpret_cast<float*>( nodeWords ) + 4 + 6 * slot;
				bw[0] = b.lo.x, bw[1] = b.lo.y, bw[2] = b.lo.z, bw[3] = b.hi.x, bw[4] = b.hi.y, bw[5] = b.hi.z;
				if ( slot == 0 )
				{
					nodeWords[28] = taskIdx == 0 ? InvalidValue : parentAddr;
					nodeWords[29] = 0;
					nodeWords[30] = count;
					nodeWords[31] = 0;
				}
			}
			__syncwarp();
			if ( ready )
			{
				const uint4* srcWords = reinterpret_cast<const uint4*>( sNode[warp][group] );
				uint4*		 dst	  = reinterpret_cast<uint4*>( boxNodes + taskIdx );
				dst[slot]			  = srcWords[slot];
				dst[slot + 4]		  = srcWords[slot + 4];
				done				  = true;
			}
			__syncwarp();
		}
		} // sub-batches of the claim
	}
}

// ------------------------------------------------------------------------------------------------
// Update (refit) of a built structure: hiprtBuildOperationUpdate (reference: ResetCountersAndUpdateLeaves
// + FitBounds, BvhBuilderKernels.h:550-814, host LbvhBuilder.h:432-484). Topology and node addresses are
// kept; leaves are rewritten from the (moved) primitives and every child box is recomputed bottom-up.
// One thread per box node, ONE launch: the thread refreshes its leaf children, then counts itself in on
// its own node; each finished internal child counts in on its parent. The arrival that completes a node
// (#internal children + 1) carries the node's union box one level up. Counters are left at zero again.
template <typename Source>
__device__ __forceinline__ DBox refreshLeaf( const Source& src, void* leafNodes, uint32_t leafAddr );

template <>
__device__ __forceinline__ DBox refreshLeaf<TriangleMeshSource>( const TriangleMeshSource& src, void* leafNodes, uint32_t leafAddr )
{
	float4*		   leaf	 = reinterpret_cast<float4*>( static_cast<TrianglePairNode*>( leafNodes ) + leafAddr );
	const float4   meta	 = leaf[3]; // {pad, primIndex0, primIndex1, flags}: the pairing of the original build is kept
	const uint32_t tri0 = __float_as_uint( meta.y ), tri1 = __float_as_uint( meta.z );
	src.writeLeaf( leafNodes, leafAddr, tri0, tri1 );
	const float4 q0 = leaf[0], q1 = leaf[1], q2 = leaf[2];
	DBox		 b	= emptyBox();
	growBox( b, make_float3( q0.x, q0.y, q0.z ) );
	growBox( b, make_float3( q0.w, q1.x, q1.y ) );
	growBox( b, make_float3( q1.z, q1.w, q2.x ) );
	growBox( b, make_float3( q2.y, q2.z, q2.w ) );
	return b;
}
template <>
__device__ __forceinline__ DBox refreshLeaf<AabbListSource>( const AabbListSource& src, void* leafNodes, uint32_t leafAddr )
{
	return src.box( static_cast<const CustomNode*>( leafNodes )[leafAddr].primIndex );
}
template <>
__device__ __forceinline__ DBox refreshLeaf<InstanceSource>( const InstanceSource& src, void* leafNodes, uint32_t leafAddr )
{
	const uint32_t p = reinterpret_cast<const uint32_t*>( static_cast<InstanceNode*>( leafNodes ) + leafAddr )[15] & 0xFFFFFFu;
	src.writeLeaf( leafNodes, leafAddr, p, 0u );
	return src.box( p );
}

// The node arrays are taken from the structure's own header (both header kinds start with {boxNodes, primNodes}; a scene's
// converted frames follow at `framesPtr`): after hiprtCompactGeometry / hiprtCompactScene, hiprtLoadGeometry or a custom-BVH
// import they do not sit where a fresh build of the same input would put them (reference: LbvhBuilder.h:445-452 reads the header back).
template <typename Source>
__device__ __forceinline__ void rebaseSource( Source&, const void* ) {}
template <>
__device__ __forceinline__ void rebaseSource<InstanceSource>( InstanceSource& src, const void* header )
{
	if ( src.frames != nullptr ) src.frames = static_cast<const SceneHeader*>( header )->frames;
}

template <typename Source>
__global__ void __launch_bounds__( 128 ) refitKernel( Source src, const void* header, const uint32_t* boxNodeCount )
{
	uint32_t node = blockIdx.x * blockDim.x + threadIdx.x;
	if ( node >= *boxNodeCount ) return;
	Box4Node* const boxNodes  = static_cast<Box4Node* const*>( header )[0];
	void* const		leafNodes = static_cast<void* const*>( header )[1];
	rebaseSource<Source>( src, header );
	uint32_t* words = reinterpret_cast<uint32_t*>( boxNodes + node );
	const uint4 children = *reinterpret_cast<const uint4*>( words );
	const uint32_t c[4]	 = { children.x, children.y, children.z, children.w };
	uint32_t	   internal = 0;
#pragma unroll
	for ( uint32_t s = 0; s < 4; ++s )
	{
		if ( c[s] == InvalidValue ) continue;
		if ( isBoxIndex( c[s] ) )
		{
			++internal;
			continue;
		}
		const DBox b  = refreshLeaf<Source>( src, leafNodes, nodeAddr( c[s] ) );
		float*	   bw = reinterpret_cast<float*>( words ) + 4 + 6 * s;
		bw[0] = b.lo.x, bw[1] = b.lo.y, bw[2] = b.lo.z, bw[3] = b.hi.x, bw[4] = b.hi.y, bw[5] = b.hi.z;
	}
	// climb while this thread is the arrival that completes the node
	while ( true )
	{
		__threadfence();
		const uint32_t arrived = atomicAdd( &words[29], 1u ) + 1u; // updateCounter
		if ( arrived < internal + 1u ) return;
		words[29] = 0; // ready for the next update
		__threadfence();
		const uint32_t parent = __ldcg( &words[28] );
		if ( parent == InvalidValue ) return; // the root is complete
		// union of this node's child boxes (written by other threads of this launch: read through L2)
		DBox		   u  = emptyBox();
		const float4*  q  = reinterpret_cast<const float4*>( words );
		const float4   q0 = __ldcg( q ), q1 = __ldcg( q + 1 ), q2 = __ldcg( q + 2 ), q3 = __ldcg( q + 3 ), q4 = __ldcg( q + 4 ), q5 = __ldcg( q + 5 ), q6 = __ldcg( q + 6 );
		const float	   b[24] = { q1.x, q1.y, q1.z, q1.w, q2.x, q2.y, q2.z, q2.w, q3.x, q3.y, q3.z, q3.w,
								 q4.x, q4.y, q4.z, q4.w, q5.x, q5.y, q5.z, q5.w, q6.x, q6.y, q6.z, q6.w };
		const uint32_t cc[4] = { __float_as_uint( q0.x ), __float_as_uint( q0.y ), __float_as_uint( q0.z ), __float_as_uint( q0.w ) };
#pragma unroll
		for ( uint32_t s = 0; s < 4; ++s )
			if ( cc[s] != InvalidValue )
				u = uniteBox( u, DBox{ make_float3( b[6 * s], b[6 * s + 1], b[6 * s + 2] ), make_float3( b[6 * s + 3], b[6 * s + 4], b[6 * s + 5] ) } );
		// my slot in the parent, and how many internal children the parent waits for
		uint32_t*	   pw	= reinterpret_cast<uint32_t*>( boxNodes + parent );
		const uint4	   pc	= *reinterpret_cast<const uint4*>( pw );
		const uint32_t me	= makeNodeIndex( node, BoxType );
		const uint32_t slot = pc.x == me ? 0u : ( pc.y == me ? 1u : ( pc.z == me ? 2u : 3u ) );
		float*		   bw	= reinterpret_cast<float*>( pw ) + 4 + 6 * slot;
		bw[0] = u.lo.x, bw[1] = u.lo.y, bw[2] = u.lo.z, bw[3] = u.hi.x, bw[4] = u.hi.y, bw[5] = u.hi.z;
		internal = uint32_t( pc.x != InvalidValue && isBoxIndex( pc.x ) ) + uint32_t( pc.y != InvalidValue && isBoxIndex( pc.y ) ) +
				   uint32_t( pc.z != InvalidValue && isBoxIndex( pc.z ) ) + uint32_t( pc.w != InvalidValue && isBoxIndex( pc.w ) );
		node  = parent;
		words = pw;
	}
}

// Scene update: refresh the per-instance records and the converted frames (the part of initSceneKernel that
// depends on the inputs), leaving header counters and topology alone.
__global__ void __launch_bounds__( 256 ) updateSceneInputsKernel(
	const SceneHeader* header, InstanceSource src, uint32_t instanceCount, const void* apiFrames, bool matrixFrames )
{
	const uint32_t	index	= blockIdx.x * blockDim.x + threadIdx.x;
	InstanceRecord* records = header->instances;
	SRTFrame*		frames	= header->frames;
	if ( index < instanceCount )
	{
		const hiprtInstance		   inst = src.instances[index];
		const hiprtTransformHeader h	= src.header( index );
		uint4					   rec;
		rec.x			   = h.frameIndex;
		rec.y			   = ( h.frameCount & 0x7FFFFFFFu ) | ( uint32_t( inst.type == hiprtInstanceTypeScene ) << 31 );
		const uint64_t ptr = reinterpret_cast<uint64_t>( inst.geometry );
		rec.z = uint32_t( ptr ), rec.w = uint32_t( ptr >> 32 );
		reinterpret_cast<uint4*>( records )[index] = rec;
	}
	if ( index < src.frameCount && apiFrames != nullptr && frames != nullptr )
	{
		const DFrame f = matrixFrames ? frameFromMatrix( static_cast<const hiprtFrameMatrix*>( apiFrames )[index] )
									  : frameFromSRT( static_cast<const hiprtFrameSRT*>( apiFrames )[index] );
		storeFrame( frames, index, f );
	}
}

// A single primitive: root with one child (reference: SingletonConstruction BvhBuilderKernels.h:296-331).
template <typename Source>
__global__ void singletonKernel( Source src, uint32_t* primNodeCounter, Box4Node* boxNodes, void* leafNodes )
{
	const DBox	b = src.box( 0 );
	const uint2 w = src.refWords( 0 );
	src.writeLeaf( leafNodes, 0, w.x, w.y );
	Box4Node node;
	for ( int s = 0; s < 4; ++s )
	{
		node.child[s] = InvalidValue;
		node.box[s]	  = Box{ { FltMax, FltMax, FltMax }, { -FltMax, -FltMax, -FltMax } };
	}
	node.child[0]	   = makeNodeIndex( 0, Source::LeafType );
	node.box[0]		   = Box{ { b.lo.x, b.lo.y, b.lo.z }, { b.hi.x, b.hi.y, b.hi.z } };
	node.parentAddr	   = InvalidValue;
	node.updateCounter = 0;
	node.childCount	   = 1;
	node.pad		   = 0;
	boxNodes[0]		   = node;
	*primNodeCounter   = 1;
}

} // namespace hiprt
