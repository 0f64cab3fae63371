// Batch builder: many small geometries (<= 512 primitives each) in ONE launch, one CTA per geometry,
// the whole LBVH pipeline in shared memory.
//
// Reference path this replaces: hiprt/impl/BatchBuilder.h:81-132 and
// BatchBuilderKernels.h:75-260 (block union, Morton, 32 one-bit split passes, Emit, Collapse loop,
// PackLeaves; no triangle pairing on this path). Results are the same structure the big builder
// produces for unpaired input: same Morton codes, stable order, radix-tree topology, collapse rule.
//
// B200 version: fixed 128-thread CTAs that loop over their primitives (a 2-triangle geometry does
// not pay for a 512-thread block; ~6 CTAs are resident per SM), the sort is a bitonic network on the
// 64-bit composite (code << 32 | index) -- unique keys, so it is stable by construction and replaces the
// reference's 32 scan passes -- topology + refit use shared-memory flags only, the collapse is a
// shared-memory work queue processed level by level by 4-lane groups, box nodes leave the CTA as whole
// 128-byte lines.
#pragma once
#include "build_kernels.cuh"

namespace hiprt
{
constexpr uint32_t BatchBlock	 = 128;
constexpr uint32_t BatchMaxPrims = 512;
constexpr uint32_t BatchMaxBox	 = 344; // maxBoxNodeCount( 512 ) = 256 + 1 + 86 = 343

struct BatchSmem
{
	float	 box[BatchMaxPrims][6];		   // primitive boxes: input order, then sorted order
	float	 nodeBox[BatchMaxPrims][6];	   // binary nodes
	uint32_t nodeChild[BatchMaxPrims][2];
	uint64_t key[BatchMaxPrims];		   // (code << 32) | primitive index
	uint32_t flag[BatchMaxPrims];
	uint64_t task[BatchMaxBox];			   // (parent box address << 32) | binary node index
	uint32_t stage[BatchBlock / 4][32];	   // one 128-byte node per 4-lane group
	float	 red[BatchBlock / 32][6];
	uint32_t tail, leafCount, root;
};

__device__ __forceinline__ DBox smemBox( const float ( *arr )[6], uint32_t i )
{
	return DBox{ make_float3( arr[i][0], arr[i][1], arr[i][2] ), make_float3( arr[i][3], arr[i][4], arr[i][5] ) };
}
__device__ __forceinline__ void setSmemBox( float ( *arr )[6], uint32_t i, const DBox& b )
{
	arr[i][0] = b.lo.x, arr[i][1] = b.lo.y, arr[i][2] = b.lo.z, arr[i][3] = b.hi.x, arr[i][4] = b.hi.y, arr[i][5] = b.hi.z;
}

template <typename Source>
__device__ void batchBuildOne( BatchSmem& s, Source src, uint32_t n, uint32_t geomType, size_t storageSize, uint8_t* storage )
{
	const uint32_t t = threadIdx.x, lane = t & 31u, warp = t >> 5;
	GeomHeader*	   header	 = reinterpret_cast<GeomHeader*>( storage );
	Box4Node*	   boxNodes	 = reinterpret_cast<Box4Node*>( storage + alignUpDev( sizeof( GeomHeader ) ) );
	uint8_t*	   leafNodes = storage + alignUpDev( sizeof( GeomHeader ) ) + alignUpDev( maxBoxNodeCount( n ) * sizeof( Box4Node ) );

	if ( n <= 1 )
	{
		if ( t == 0 )
		{
			Box4Node node;
			for ( int k = 0; k < 4; ++k )
			{
				node.child[k] = InvalidValue;
				node.box[k]	  = Box{ { FltMax, FltMax, FltMax }, { -FltMax, -FltMax, -FltMax } };
			}
			if ( n == 1 )
			{
				const DBox b = src.box( 0 );
				const uint2 w = src.refWords( 0 );
				src.writeLeaf( leafNodes, 0, w.x, w.y );
				node.child[0] = makeNodeIndex( 0, Source::LeafType );
				node.box[0]	  = Box{ { b.lo.x, b.lo.y, b.lo.z }, { b.hi.x, b.hi.y, b.hi.z } };
			}
			node.parentAddr = InvalidValue, node.updateCounter = 0, node.childCount = n, node.pad = 0;
			boxNodes[0] = node;
			*header		= GeomHeader{ boxNodes, leafNodes, storageSize, 0u, 1u, n, geomType, 0u };
		}
		return;
	}

	// ---- primitive boxes + scene box
	DBox mine = emptyBox();
	for ( uint32_t p = t; p < n; p += BatchBlock )
	{
		const DBox b = src.box( p );
		setSmemBox( s.box, p, b );
		mine = uniteBox( mine, b );
	}
	mine = warpUniteBox( mine );
	if ( lane == 0 ) setSmemBox( s.red, warp, mine );
	if ( t == 0 ) s.tail = 1, s.leafCount = 0, s.root = InvalidValue;
	__syncthreads();
	DBox scene = smemBox( s.red, 0 );
#pragma unroll
	for ( uint32_t w = 1; w < BatchBlock / 32; ++w )
		scene = uniteBox( scene, smemBox( s.red, w ) );

	// ---- Morton codes; padded to a power of two for the bitonic network
	uint32_t padded = 2;
	while ( padded < n )
		padded <<= 1;
	{
		const float3	   ext = make_float3( scene.hi.x - scene.lo.x, scene.hi.y - scene.lo.y, scene.hi.z - scene.lo.z );
		const MortonLayout L   = mortonLayout( ext );
		for ( uint32_t p = t; p < padded; p += BatchBlock )
		{
			uint64_t key = ~0ull;
			if ( p < n )
			{
				const DBox	 b = smemBox( s.box, p );
				const float3 c = make_float3( ( b.hi.x + b.lo.x ) * 0.5f, ( b.hi.y + b.lo.y ) * 0.5f, ( b.hi.z + b.lo.z ) * 0.5f );
				const float3 pos =
					make_float3( __fdividef( c.x - scene.lo.x, ext.x ), __fdividef( c.y - scene.lo.y, ext.y ), __fdividef( c.z - scene.lo.z, ext.z ) );
				key = ( uint64_t( mortonCode( L, pos ) ) << 32 ) | p;
			}
			s.key[p] = key;
		}
	}
	__syncthreads();

	// ---- bitonic sort of the composite keys (ascending)
	for ( uint32_t k = 2; k <= padded; k <<= 1 )
		for ( uint32_t j = k >> 1; j > 0; j >>= 1 )
		{
			for ( uint32_t i = t; i < padded; i += BatchBlock )
			{
				const uint32_t partner = i ^ j;
				if ( partner > i )
				{
					const uint64_t a = s.key[i], b = s.key[partner];
					const bool	   up = ( i & k ) == 0;
					if ( ( a > b ) == up ) s.key[i] = b, s.key[partner] = a;
				}
			}
			__syncthreads();
		}

	// ---- boxes into sorted order (through registers), flags reset
	{
		DBox regs[BatchMaxPrims / BatchBlock];
#pragma unroll
		for ( uint32_t q = 0; q < BatchMaxPrims / BatchBlock; ++q )
		{
			const uint32_t k = t + q * BatchBlock;
			if ( k < n ) regs[q] = smemBox( s.box, uint32_t( s.key[k] ) );
		}
		__syncthreads();
#pragma unroll
		for ( uint32_t q = 0; q < BatchMaxPrims / BatchBlock; ++q )
		{
			const uint32_t k = t + q * BatchBlock;
			if ( k < n ) setSmemBox( s.box, k, regs[q] ), s.flag[k] = InvalidValue;
		}
	}
	__syncthreads();

	// ---- radix-tree topology + refit, everything in shared memory
	auto levelOf = [&]( uint32_t a ) -> uint32_t {
		if ( a >= n - 1 ) return 64u;
		const uint32_t x = uint32_t( s.key[a] >> 32 ) ^ uint32_t( s.key[a + 1] >> 32 );
		return x ? 63u - __clz( x ) : 31u - __clz( a ^ ( a + 1 ) );
	};
	for ( uint32_t k = t; k < n; k += BatchBlock )
	{
		uint32_t i = k, j = k + 1, self = makeNodeIndex( k, Source::LeafType );
		DBox	 box = smemBox( s.box, k );
		while ( true )
		{
			const bool	   left	  = ( i == 0 ) || ( j != n && levelOf( j - 1 ) < levelOf( i - 1 ) );
			const uint32_t parent = left ? j - 1 : i - 1;
			s.nodeChild[parent][left ? 0 : 1] = self;
			__threadfence_block();
			const uint32_t other = atomicExch( &s.flag[parent], left ? i : j );
			if ( other == InvalidValue ) break;
			__threadfence_block();
			if ( left )
				j = other;
			else
				i = other;
			const uint32_t sib = reinterpret_cast<volatile uint32_t*>( &s.nodeChild[parent][left ? 1 : 0] )[0];
			const volatile float* sb = isBoxIndex( sib ) ? s.nodeBox[nodeAddr( sib )] : s.box[nodeAddr( sib )];
			box = uniteBox( box, DBox{ make_float3( sb[0], sb[1], sb[2] ), make_float3( sb[3], sb[4], sb[5] ) } );
			setSmemBox( s.nodeBox, parent, box );
			self = makeNodeIndex( parent, BoxType );
			if ( i == 0 && j == n )
			{
				s.root	  = parent;
				s.task[0] = ( uint64_t( InvalidValue ) << 32 ) | self;
				break;
			}
		}
	}
	__syncthreads();

	// ---- collapse to 4-wide nodes + leaf packing: queue processed level by level, 4 lanes per task
	auto loadRecS = [&]( uint32_t index ) -> Rec32 {
		Rec32		   r;
		const uint32_t a = nodeAddr( index );
		if ( isBoxIndex( index ) )
			r.box = smemBox( s.nodeBox, a ), r.a = s.nodeChild[a][0], r.b = s.nodeChild[a][1];
		else
			r.box = smemBox( s.box, a ), r.a = uint32_t( s.key[a] ), r.b = 0;
		return r;
	};
	const uint32_t group = t >> 2, slot = t & 3u, groupInWarp = lane >> 2;
	const uint32_t groupMask = 0xFu << ( groupInWarp * 4 );
	uint32_t	   head = 0;
	while ( true )
	{
		const uint32_t tail = s.tail; // uniform: read after a barrier
		if ( head >= tail ) break;
		__syncthreads(); // everybody has read the tail before anyone appends
		for ( uint32_t base = head; base < tail; base += BatchBlock / 4 )
		{
			const uint32_t taskIdx = base + group;
			const bool	   ready   = taskIdx < tail;
			const uint64_t task	   = ready ? s.task[taskIdx] : 0ull;
			const uint32_t binIndex = uint32_t( task ), parentAddr = uint32_t( task >> 32 );
			uint32_t	   child = InvalidValue, w0 = 0, w1 = 0, count = 2;
			DBox		   cbox	 = emptyBox();
			if ( ready && slot < 2 )
			{
				child			= s.nodeChild[nodeAddr( binIndex )][slot];
				const Rec32 rec = loadRecS( child );
				cbox = rec.box, w0 = rec.a, w1 = rec.b;
			}
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
						const Rec32 rec = loadRecS( child );
						cbox = rec.box, w0 = rec.a, w1 = rec.b;
					}
					++count;
				}
			}
			const bool	   active	= ready && slot < count;
			const bool	   internal = active && isBoxIndex( child );
			const bool	   leaf		= active && !internal;
			const uint32_t imask = __ballot_sync( FullMask, internal ), lmask = __ballot_sync( FullMask, leaf );
			uint32_t	   ibase = 0, lbase = 0;
			if ( lane == 0 )
			{
				if ( imask ) ibase = atomicAdd( &s.tail, __popc( imask ) );
				if ( lmask ) lbase = atomicAdd( &s.leafCount, __popc( lmask ) );
			}
			ibase			  = __shfl_sync( FullMask, ibase, 0 );
			lbase			  = __shfl_sync( FullMask, lbase, 0 );
			uint32_t outChild = InvalidValue;
			if ( internal )
			{
				const uint32_t addr = ibase + __popc( imask & laneMaskLt() );
				s.task[addr]		= ( uint64_t( taskIdx ) << 32 ) | child;
				outChild			= makeNodeIndex( addr, BoxType );
			}
			if ( leaf )
			{
				const uint32_t addr = lbase + __popc( lmask & laneMaskLt() );
				const uint2 w = src.refWords( w0 );
				src.writeLeaf( leafNodes, addr, w.x, w.y );
				outChild = makeNodeIndex( addr, Source::LeafType );
			}
			if ( ready )
			{
				uint32_t* nodeWords = s.stage[group];
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
				const uint4* srcWords = reinterpret_cast<const uint4*>( s.stage[group] );
				uint4*		 dst	  = reinterpret_cast<uint4*>( boxNodes + taskIdx );
				dst[slot]			  = srcWords[slot];
				dst[slot + 4]		  = srcWords[slot + 4];
			}
			__syncwarp();
		}
		head = tail;
		__syncthreads();
	}
	if ( t == 0 ) *header = GeomHeader{ boxNodes, leafNodes, storageSize, 0u, s.tail, n, geomType, 0u };
}

// inputs / storages: device copies of the caller's host arrays (hiprtGeometryBuildInput is 128 bytes).
__global__ void __launch_bounds__( BatchBlock )
	batchBuildKernel( const hiprtGeometryBuildInput* inputs, void* const* storages, uint32_t count )
{
	extern __shared__ __align__( 16 ) uint8_t batchSmemRaw[];
	BatchSmem&								  s = *reinterpret_cast<BatchSmem*>( batchSmemRaw );
	const uint32_t							  g = blockIdx.x + blockIdx.y * gridDim.x;
	if ( g >= count ) return;
	const hiprtGeometryBuildInput& in	   = inputs[g];
	uint8_t*					   storage = static_cast<uint8_t*>( storages[g] );
	if ( in.type == hiprtPrimitiveTypeTriangleMesh )
	{
		const auto&		   m = in.primitive.triangleMesh;
		TriangleMeshSource src{};
		src.vertices	   = static_cast<const uint8_t*>( m.vertices );
		src.vertexStride   = m.vertexStride;
		src.triangles	   = ( m.triangleCount == 0 ) ? nullptr : static_cast<const uint8_t*>( m.triangleIndices );
		src.triangleStride = m.triangleStride;
		src.pairs		   = m.trianglePairCount > 0 ? static_cast<const uint2*>( m.trianglePairIndices ) : nullptr;
		const uint32_t n   = m.trianglePairCount > 0 ? m.trianglePairCount
							 : ( m.triangleCount == 0 || m.triangleIndices == nullptr ) ? m.vertexCount / 3
																						: m.triangleCount;
		const size_t   size = alignUpDev( sizeof( GeomHeader ) ) + alignUpDev( maxBoxNodeCount( n ) * sizeof( Box4Node ) ) + alignUpDev( size_t( n ) * sizeof( TrianglePairNode ) );
		batchBuildOne( s, src, n, ( in.geomType << 1 ) | 1u, size, storage );
	}
	else
	{
		const auto&	   a = in.primitive.aabbList;
		AabbListSource src{ static_cast<const uint8_t*>( a.aabbs ), a.aabbStride };
		const uint32_t n	= a.aabbCount;
		const size_t   size = alignUpDev( sizeof( GeomHeader ) ) + alignUpDev( maxBoxNodeCount( n ) * sizeof( Box4Node ) ) + alignUpDev( size_t( n ) * sizeof( CustomNode ) );
		batchBuildOne( s, src, n, in.geomType << 1, size, storage );
	}
}
} // namespace hiprt
