// Primitive containers of the builders: what a "primitive" is (triangle pair, custom AABB,
// instance), how its box is fetched and how its leaf node is written.
// Reference counterparts: TriangleMesh.h:56-287, AabbList.h:31-60, InstanceList.h:31-235.
#pragma once
#include "device_util.cuh"
#include "transform_util.cuh"
#include <hiprt/hiprt_b200_frame.h>

namespace hiprt
{
// ------------------------------------------------------------------------------------------------
struct TriangleMeshSource
{
	static constexpr uint32_t LeafType = TriangleType;
	static constexpr uint32_t LeafSize = sizeof( TrianglePairNode );

	const uint8_t* vertices;
	const uint8_t* triangles; // null: triangle t uses vertices 3t..3t+2
	const uint2*   pairs;	  // null: primitive p is the lone triangle p
	uint32_t	   vertexStride;
	uint32_t	   triangleStride;

	__device__ __forceinline__ float3 vertex( uint32_t i ) const
	{
		const float* p = reinterpret_cast<const float*>( vertices + size_t( i ) * vertexStride );
		return make_float3( __ldg( p ), __ldg( p + 1 ), __ldg( p + 2 ) );
	}
	__device__ __forceinline__ uint3 indices( uint32_t t ) const
	{
		if ( triangles == nullptr ) return make_uint3( 3 * t, 3 * t + 1, 3 * t + 2 );
		const uint32_t* p = reinterpret_cast<const uint32_t*>( triangles + size_t( t ) * triangleStride );
		return make_uint3( __ldg( p ), __ldg( p + 1 ), __ldg( p + 2 ) );
	}
	__device__ __forceinline__ uint2 pairOf( uint32_t p ) const { return pairs ? pairs[p] : make_uint2( p, p ); }

	__device__ __forceinline__ DBox box( uint32_t p ) const
	{
		const uint2 pr = pairOf( p );
		const uint3 i0 = indices( pr.x );
		DBox		b  = emptyBox();
		growBox( b, vertex( i0.x ) );
		growBox( b, vertex( i0.y ) );
		growBox( b, vertex( i0.z ) );
		if ( pr.x != pr.y )
		{
			const uint3 i1 = indices( pr.y );
			growBox( b, vertex( i1.x ) );
			growBox( b, vertex( i1.y ) );
			growBox( b, vertex( i1.z ) );
		}
		return b;
	}

	// For each index of b its position in a (3 = absent); pairable iff at most one is absent
	// (TriangleMesh.h:32-50).
	static __device__ __forceinline__ bool mapIndices( const uint3 a, const uint3 b, uint3& m )
	{
		m.x = b.x == a.z ? 2u : ( b.x == a.y ? 1u : ( b.x == a.x ? 0u : 3u ) );
		m.y = b.y == a.z ? 2u : ( b.y == a.y ? 1u : ( b.y == a.x ? 0u : 3u ) );
		m.z = b.z == a.z ? 2u : ( b.z == a.y ? 1u : ( b.z == a.x ? 0u : 3u ) );
		return ( m.x == 3u ) + ( m.y == 3u ) + ( m.z == 3u ) <= 1;
	}

	// pairable without the mapping: at least two of b's indices occur in a
	static __device__ __forceinline__ bool sharesEdge( const uint3 a, const uint3 b )
	{
		const uint32_t hits = uint32_t( b.x == a.x || b.x == a.y || b.x == a.z ) + uint32_t( b.y == a.x || b.y == a.y || b.y == a.z ) +
							  uint32_t( b.z == a.x || b.z == a.y || b.z == a.z );
		return hits >= 2;
	}

	// 64-byte pair leaf: triangle 0 rotated so that its unshared vertex comes first, the shared
	// edge is (v1,v2) and v3 is the unshared vertex of triangle 1 (TriangleMesh.h:89-157).
	// the two words a primitive's 32-byte record carries through the build: its two triangle indices, so that
	// the leaf can be written without going back to the pair list
	__device__ __forceinline__ uint2 refWords( uint32_t p ) const { return pairOf( p ); }

	__device__ __forceinline__ void writeLeaf( void* leafArray, uint32_t leafAddr, uint32_t tri0, uint32_t tri1 ) const
	{
		const uint2 pr	  = make_uint2( tri0, tri1 );
		uint3		i0	  = indices( pr.x );
		uint32_t	flags = DefaultTriangleFlags;
		uint32_t	i3	  = 0;
		if ( pr.x != pr.y )
		{
			uint3 i1 = indices( pr.y );
			uint3 m;
			mapIndices( i1, i0, m );
			uint3 f0 = make_uint3( 1, 2, 0 );
			if ( m.y == 3u )
			{
				m  = make_uint3( m.y, m.z, m.x );
				i0 = make_uint3( i0.y, i0.z, i0.x );
				f0 = make_uint3( f0.z, f0.x, f0.y );
			}
			else if ( m.z == 3u )
			{
				m  = make_uint3( m.z, m.x, m.y );
				i0 = make_uint3( i0.z, i0.x, i0.y );
				f0 = make_uint3( f0.y, f0.z, f0.x );
			}
			const bool keep = ( m.y == 0u && m.z == 2u ) || ( m.y == 2u && m.z == 1u ) || ( m.y == 1u && m.z == 0u );
			const bool flip = !keep;
			uint3	   f1	= flip ? make_uint3( 2, 1, 0 ) : make_uint3( 0, 1, 2 );
			if ( ( m.y == 2u && m.z == 1u ) || ( m.y == 1u && m.z == 2u ) )
			{
				i1 = make_uint3( i1.y, i1.z, i1.x );
				f1 = make_uint3( f1.z, f1.x, f1.y );
			}
			else if ( ( m.y == 2u && m.z == 0u ) || ( m.y == 0u && m.z == 2u ) )
			{
				i1 = make_uint3( i1.z, i1.x, i1.y );
				f1 = make_uint3( f1.y, f1.z, f1.x );
			}
			flags = ( uint32_t( flip ) << 13 ) | ( f1.y << 10 ) | ( f1.x << 8 ) | ( f0.y << 2 ) | f0.x;
			i3	  = i1.z;
		}
		const float3 v0 = vertex( i0.x ), v1 = vertex( i0.y ), v2 = vertex( i0.z );
		const float3 v3 = ( pr.x != pr.y ) ? vertex( i3 ) : v2;
		float4*		 dst = reinterpret_cast<float4*>( static_cast<TrianglePairNode*>( leafArray ) + leafAddr );
		dst[0]			 = make_float4( v0.x, v0.y, v0.z, v1.x );
		dst[1]			 = make_float4( v1.y, v1.z, v2.x, v2.y );
		dst[2]			 = make_float4( v2.z, v3.x, v3.y, v3.z );
		dst[3]			 = make_float4( 0.0f, __uint_as_float( pr.x ), __uint_as_float( pr.y ), __uint_as_float( flags ) );
	}
};

// ------------------------------------------------------------------------------------------------
struct AabbListSource
{
	static constexpr uint32_t LeafType = CustomType;
	static constexpr uint32_t LeafSize = sizeof( CustomNode );

	const uint8_t* aabbs;
	uint32_t	   stride; // 24 (2 x float3) or 32 (2 x float4); max starts at stride/2

	__device__ __forceinline__ DBox box( uint32_t p ) const
	{
		const float* lo = reinterpret_cast<const float*>( aabbs + size_t( p ) * stride );
		const float* hi = reinterpret_cast<const float*>( aabbs + size_t( p ) * stride + ( stride >> 1 ) );
		return DBox{ make_float3( __ldg( lo ), __ldg( lo + 1 ), __ldg( lo + 2 ) ), make_float3( __ldg( hi ), __ldg( hi + 1 ), __ldg( hi + 2 ) ) };
	}
	__device__ __forceinline__ uint2 refWords( uint32_t p ) const { return make_uint2( p, 0u ); }
	__device__ __forceinline__ void	 writeLeaf( void* leafArray, uint32_t leafAddr, uint32_t p, uint32_t ) const
	{
		static_cast<CustomNode*>( leafArray )[leafAddr].primIndex = p;
	}
};

// ------------------------------------------------------------------------------------------------
// Instances of a scene. Box = union over the (up to 4) child boxes of the referenced structure's
// root node of the box of their 8 transformed corners (reference: InstanceList.h:104-121,
// Transform.h:496-517); leaf = 64-byte InstanceNode holding the world->object matrix
// (BvhNode.h:1142-1182). Frames have been converted to SRT form by initSceneKernel.
__device__ __forceinline__ motion::Frame toMotionFrame( const DFrame& f ) { return motion::Frame{ f.rotation, f.scale, f.shear, f.translation, f.time }; }

struct InstanceSource
{
	static constexpr uint32_t LeafType = InstanceType;
	static constexpr uint32_t LeafSize = sizeof( InstanceNode );

	const hiprtInstance*		instances;
	const hiprtTransformHeader* headers; // null: frame k belongs to instance k
	const SRTFrame*				frames;	 // converted frames inside the scene storage
	const uint32_t*				masks;	 // null: all ones
	uint32_t					frameCount;

	__device__ __forceinline__ hiprtTransformHeader header( uint32_t p ) const
	{
		return headers ? headers[p] : hiprtTransformHeader{ p, 1u };
	}
	__device__ __forceinline__ DFrame frame( uint32_t index ) const
	{
		return ( frameCount == 0 || frames == nullptr ) ? identityFrame() : loadFrame( frames, index );
	}

	__device__ __forceinline__ DBox box( uint32_t p ) const
	{
		const hiprtTransformHeader h = header( p );
		// both header kinds start with the box-node pointer (Header.h:33-60)
		const Box4Node* root = *reinterpret_cast<Box4Node* const*>( instances[p].geometry );
		const float4*	q	 = reinterpret_cast<const float4*>( root );
		const float4	q0 = q[0], q1 = q[1], q2 = q[2], q3 = q[3], q4 = q[4], q5 = q[5], q6 = q[6];
		const float		b[24]	   = { q1.x, q1.y, q1.z, q1.w, q2.x, q2.y, q2.z, q2.w, q3.x, q3.y, q3.z, q3.w,
									   q4.x, q4.y, q4.z, q4.w, q5.x, q5.y, q5.z, q5.w, q6.x, q6.y, q6.z, q6.w };
		const uint32_t	child[4]   = { __float_as_uint( q0.x ), __float_as_uint( q0.y ), __float_as_uint( q0.z ), __float_as_uint( q0.w ) };
		const uint32_t	childCount = reinterpret_cast<const uint32_t*>( root )[30];
		DBox			out		   = emptyBox();
		const uint32_t	nFrames	   = max( h.frameCount, 1u );
		// Animated instances are bounded over their motion as the reference does (Transform::boundPointMotion,
		// Transform.h:461-494): every key frame plus three interpolated poses (t = 1/4, 1/2, 3/4) between consecutive keys.
		constexpr uint32_t Steps = 3;
		for ( uint32_t f = 0; f < nFrames; ++f )
		{
			const DFrame key = frame( h.frameIndex + f );
			for ( uint32_t step = 0; step <= ( f + 1 < nFrames ? Steps : 0u ); ++step )
			{
				DFrame fr = key;
				if ( step > 0 )
				{
					const DFrame		next = frame( h.frameIndex + f + 1 );
					const motion::Frame m	 = motion::interpolate( toMotionFrame( key ), toMotionFrame( next ), float( step ) * ( 1.0f / float( Steps + 1 ) ) );
					fr						 = DFrame{ m.rotation, m.scale, m.shear, m.translation, 0.0f };
				}
				const bool ident = frameIsIdentity( fr );
#pragma unroll
				for ( uint32_t c = 0; c < 4; ++c )
				{
					if ( c >= childCount || child[c] == InvalidValue ) continue;
					const float3 lo = make_float3( b[6 * c], b[6 * c + 1], b[6 * c + 2] ), hi = make_float3( b[6 * c + 3], b[6 * c + 4], b[6 * c + 5] );
#pragma unroll
					for ( uint32_t k = 0; k < 8; ++k )
					{
						const float3 corner = make_float3( ( k & 4 ) ? hi.x : lo.x, ( k & 2 ) ? hi.y : lo.y, ( k & 1 ) ? hi.z : lo.z );
						growBox( out, ident ? corner : frameTransformPoint( fr, corner ) );
					}
				}
			}
		}
		return out;
	}

	__device__ __forceinline__ uint2 refWords( uint32_t p ) const { return make_uint2( p, 0u ); }
	__device__ __forceinline__ void	 writeLeaf( void* leafArray, uint32_t leafAddr, uint32_t p, uint32_t ) const
	{
		const hiprtTransformHeader h	= header( p );
		const hiprtInstance		   inst = instances[p];
		const DFrame			   fr	= frame( h.frameIndex );
		// static = exactly one key frame (UserInstanceNode::init, BvhNode.h:1152-1168); an animated instance keeps its key-frame
		// range {frameIndex, frameCount} in the first eight bytes instead of a matrix and is never flagged identity
		const bool isStatic = h.frameCount <= 1;
		const bool ident	= isStatic && frameIsIdentity( fr );
		float	   m[3][4]	= { { 1, 0, 0, 0 }, { 0, 1, 0, 0 }, { 0, 0, 1, 0 } };
		if ( isStatic && !ident ) frameInverseMatrix( fr, m );
		if ( !isStatic )
		{
			m[0][0] = __uint_as_float( h.frameIndex ), m[0][1] = __uint_as_float( h.frameCount );
			m[1][1] = 0.0f, m[2][2] = 0.0f;
		}
		const uint32_t mask = masks ? masks[p] : FullRayMask;
		// bits: primIndex:24 | type:1 | static:1 | identity:1
		const uint32_t bits =
			( p & 0xFFFFFFu ) | ( uint32_t( inst.type == hiprtInstanceTypeScene ) << 24 ) | ( uint32_t( isStatic ) << 25 ) | ( uint32_t( ident ) << 26 );
		const uint64_t ptr	= reinterpret_cast<uint64_t>( inst.geometry );
		float4*		   dst	= reinterpret_cast<float4*>( static_cast<InstanceNode*>( leafArray ) + leafAddr );
		dst[0]				= make_float4( m[0][0], m[0][1], m[0][2], m[0][3] );
		dst[1]				= make_float4( m[1][0], m[1][1], m[1][2], m[1][3] );
		dst[2]				= make_float4( m[2][0], m[2][1], m[2][2], m[2][3] );
		dst[3] = make_float4( __uint_as_float( uint32_t( ptr ) ), __uint_as_float( uint32_t( ptr >> 32 ) ), __uint_as_float( mask ), __uint_as_float( bits ) );
	}
};
} // namespace hiprt
