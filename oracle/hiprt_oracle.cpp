// ORACLE -- TEST INFRASTRUCTURE ONLY. Not part of the product.
//
// Single-threaded CPU restatement of the reference's (GPUOpen HIPRT v3.1.0) hot path:
// triangle pairing -> primitive boxes -> extended Morton codes -> stable sort -> LBVH
// topology + refit -> collapse to 4-wide nodes -> leaf packing -> stack traversal with
// Moller-Trumbore triangles. Each function cites the reference file:line it follows.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load this
// library; the product (hiprt_b200/) never does.
//
// Parity pinning: checked against the reference's golden images (tests/golden/, generated
// by tests/golden/make_golden.py from test/references/*.png) and against hit records
// dumped from the reference's own kernels on a B200 (oracle/_ref, see oracle/Makefile).
//
// Build: g++ -O2 -ffp-contract=off (no implicit FMA: every fused op below is explicit, as
// in the reference's hiprt_math.h:1571-1581).

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <deque>
#include <limits>
#include <numeric>
#include <thread>
#include <vector>

#include <hiprt/hiprt_b200_layout.h>

using namespace hiprt;

namespace
{
using V3 = hiprtFloat3;

inline V3	 mk( float x, float y, float z ) { return V3{ x, y, z }; }
inline V3	 operator-( V3 a, V3 b ) { return mk( a.x - b.x, a.y - b.y, a.z - b.z ); }
inline V3	 operator+( V3 a, V3 b ) { return mk( a.x + b.x, a.y + b.y, a.z + b.z ); }
inline V3	 operator*( V3 a, V3 b ) { return mk( a.x * b.x, a.y * b.y, a.z * b.z ); }
inline V3	 operator*( V3 a, float s ) { return mk( a.x * s, a.y * s, a.z * s ); }
inline V3	 vmin( V3 a, V3 b ) { return mk( fminf( a.x, b.x ), fminf( a.y, b.y ), fminf( a.z, b.z ) ); }
inline V3	 vmax( V3 a, V3 b ) { return mk( fmaxf( a.x, b.x ), fmaxf( a.y, b.y ), fmaxf( a.z, b.z ) ); }
inline float comp( const V3& v, int a ) { return ( &v.x )[a]; }

// hiprt_math.h:1571-1581 (Kahan-style compensated products)
inline float sumProd( float a, float b, float c, float d )
{
	const float cd = c * d;
	return fmaf( a, b, cd ) + fmaf( c, d, -cd );
}
inline float diffProd( float a, float b, float c, float d )
{
	const float cd = c * d;
	return fmaf( a, b, -cd ) - fmaf( c, d, -cd );
}
// hiprt_math.h:1592-1603
inline V3 cross( V3 a, V3 b )
{
	return mk( diffProd( a.y, b.z, a.z, b.y ), diffProd( a.z, b.x, a.x, b.z ), diffProd( a.x, b.y, a.y, b.x ) );
}
inline float dot( V3 a, V3 b ) { return fmaf( a.x, b.x, sumProd( a.y, b.y, a.z, b.z ) ); }

inline Box emptyBox() { return Box{ mk( FltMax, FltMax, FltMax ), mk( -FltMax, -FltMax, -FltMax ) }; }
inline Box grow( Box b, V3 p ) { return Box{ vmin( b.lo, p ), vmax( b.hi, p ) }; }
inline Box unite( Box a, Box b ) { return Box{ vmin( a.lo, b.lo ), vmax( a.hi, b.hi ) }; }
// Aabb.h:71-75
inline float area( const Box& b )
{
	const V3 e = b.hi - b.lo;
	return 2 * ( e.x * e.y + e.x * e.z + e.y * e.z );
}

// ---------------------------------------------------------------------------------------------
// Mesh access (TriangleMesh.h:56-87)
struct Mesh
{
	const uint8_t* verts;
	uint32_t	   vertCount, vertStride;
	const uint8_t* tris; // may be null
	uint32_t	   triCount, triStride;

	V3 vertex( uint32_t i ) const
	{
		const float* p = reinterpret_cast<const float*>( verts + size_t( i ) * vertStride );
		return mk( p[0], p[1], p[2] );
	}
	void indices( uint32_t t, uint32_t out[3] ) const
	{
		if ( !tris )
		{
			out[0] = 3 * t, out[1] = 3 * t + 1, out[2] = 3 * t + 2;
			return;
		}
		const uint32_t* p = reinterpret_cast<const uint32_t*>( tris + size_t( t ) * triStride );
		out[0] = p[0], out[1] = p[1], out[2] = p[2];
	}
};

// TriangleMesh.h:32-50. For each index of b: position (0..2) of the equal index in a, 3 if
// absent (later positions win). Valid pairing iff at most one index of b is absent from a.
bool vertexMapping( const uint32_t a[3], const uint32_t b[3], uint32_t m[3] )
{
	for ( int k = 0; k < 3; ++k )
	{
		m[k] = 3;
		for ( uint32_t j = 0; j < 3; ++j )
			if ( b[k] == a[j] ) m[k] = j;
	}
	return ( m[0] == 3 ) + ( m[1] == 3 ) + ( m[2] == 3 ) <= 1;
}

inline void rotL( uint32_t v[3] )
{
	const uint32_t t = v[0];
	v[0] = v[1], v[1] = v[2], v[2] = t;
}
inline void rotR( uint32_t v[3] )
{
	const uint32_t t = v[2];
	v[2] = v[1], v[1] = v[0], v[0] = t;
}

// TriangleMesh.h:89-157: build the 64-byte leaf for the pair (t0,t1); t0 == t1 is a lone triangle.
TrianglePairNode makePairNode( const Mesh& mesh, uint32_t t0, uint32_t t1 )
{
	uint32_t i0[3], i1[3] = { 0, 0, 0 };
	mesh.indices( t0, i0 );
	uint32_t flags = DefaultTriangleFlags;
	if ( t0 != t1 )
	{
		mesh.indices( t1, i1 );
		uint32_t m[3];
		vertexMapping( i1, i0, m ); // where each vertex of triangle 0 sits in triangle 1
		uint32_t f0[3] = { 1, 2, 0 };
		if ( m[1] == 3 )
		{
			rotL( m ), rotL( i0 ), rotR( f0 );
		}
		else if ( m[2] == 3 )
		{
			rotR( m ), rotR( i0 ), rotL( f0 );
		}
		// now the vertex of triangle 0 that triangle 1 lacks is first; shared edge is (1,2)
		const bool sameWinding = ( m[1] == 0 && m[2] == 2 ) || ( m[1] == 2 && m[2] == 1 ) || ( m[1] == 1 && m[2] == 0 );
		const bool flip		   = !sameWinding;
		uint32_t   f1[3]	   = { 0, 1, 2 };
		if ( flip ) f1[0] = 2, f1[2] = 0;
		if ( ( m[1] == 2 && m[2] == 1 ) || ( m[1] == 1 && m[2] == 2 ) )
		{
			rotL( i1 ), rotR( f1 );
		}
		else if ( ( m[1] == 2 && m[2] == 0 ) || ( m[1] == 0 && m[2] == 2 ) )
		{
			rotR( i1 ), rotL( f1 );
		}
		flags = ( uint32_t( flip ) << 13 ) | ( f1[1] << 10 ) | ( f1[0] << 8 ) | ( f0[1] << 2 ) | f0[0];
	}
	TrianglePairNode n;
	memset( &n, 0, sizeof( n ) );
	n.flags		 = flags;
	n.primIndex0 = t0;
	n.primIndex1 = t1;
	n.v0		 = mesh.vertex( i0[0] );
	n.v1		 = mesh.vertex( i0[1] );
	n.v2		 = mesh.vertex( i0[2] );
	n.v3		 = ( t0 != t1 ) ? mesh.vertex( i1[2] ) : n.v2;
	return n;
}

inline Box pairBox( const TrianglePairNode& n ) { return grow( grow( grow( grow( emptyBox(), n.v0 ), n.v1 ), n.v2 ), n.v3 ); }

// BvhBuilderKernels.h:373-420 (also restated on the CPU by the reference's own test harness,
// test/hiprtTest.cpp:480-509): greedy pairing inside each group of 32 consecutive triangles.
// Groups are emitted in ascending order here (the reference's cross-warp order is atomic).
uint32_t pairTriangles( const Mesh& mesh, uint32_t* pairsOut )
{
	uint32_t count = 0;
	for ( uint32_t base = 0; base < mesh.triCount; base += 32 )
	{
		const uint32_t n = std::min( 32u, mesh.triCount - base );
		uint32_t	   idx[32][3];
		bool		   active[32];
		for ( uint32_t l = 0; l < n; ++l )
		{
			mesh.indices( base + l, idx[l] );
			active[l] = true;
		}
		for ( uint32_t a = 0; a < n; ++a )
		{
			if ( !active[a] ) continue;
			active[a]		 = false;
			uint32_t partner = a;
			for ( uint32_t b = 0; b < n; ++b )
			{
				uint32_t m[3];
				if ( b != a && active[b] && vertexMapping( idx[a], idx[b], m ) )
				{
					partner	  = b;
					active[b] = false;
					break;
				}
			}
			pairsOut[2 * count + 0] = base + a;
			pairsOut[2 * count + 1] = base + partner;
			++count;
		}
	}
	return count;
}

// ---------------------------------------------------------------------------------------------
// Extended Morton code (MortonCode.h:60-277). 32-bit wrap-around arithmetic is made explicit.
inline uint32_t spread2( uint32_t v )
{
	v &= 0x0000ffffu;
	v = ( v ^ ( v << 8 ) ) & 0x00ff00ffu;
	v = ( v ^ ( v << 4 ) ) & 0x0f0f0f0fu;
	v = ( v ^ ( v << 2 ) ) & 0x33333333u;
	v = ( v ^ ( v << 1 ) ) & 0x55555555u;
	return v;
}
inline uint32_t spread3( uint32_t v )
{
	v = ( v * 0x00010001u ) & 0xFF0000FFu;
	v = ( v * 0x00000101u ) & 0x0F00F00Fu;
	v = ( v * 0x00000011u ) & 0xC30C30C3u;
	v = ( v * 0x00000005u ) & 0x49249249u;
	return v;
}
// float -> int32 as the GPU converts (round toward zero, saturating, NaN -> 0)
inline int32_t f2i( float f )
{
	if ( f != f ) return 0;
	if ( f >= 2147483648.0f ) return std::numeric_limits<int32_t>::max();
	if ( f <= -2147483648.0f ) return std::numeric_limits<int32_t>::min();
	return int32_t( f );
}
inline uint32_t f2u( float f )
{
	if ( f != f || f <= 0.0f ) return 0;
	if ( f >= 4294967296.0f ) return 0xFFFFFFFFu;
	return uint32_t( f );
}
inline int32_t wadd( int32_t a, int32_t b ) { return int32_t( uint32_t( a ) + uint32_t( b ) ); }
inline int32_t wsub( int32_t a, int32_t b ) { return int32_t( uint32_t( a ) - uint32_t( b ) ); }
inline uint32_t shl( uint32_t v, int32_t s ) { return ( s < 0 || s > 31 ) ? 0u : v << s; }
inline uint32_t shr( uint32_t v, int32_t s ) { return ( s < 0 || s > 31 ) ? 0u : v >> s; }

uint32_t extendedMorton( V3 pos, V3 ext )
{
	// axes by decreasing extent, stable for equal extents (the if-ladder at :69-148)
	int ax[3] = { 0, 1, 2 };
	if ( comp( ext, ax[0] ) < comp( ext, ax[1] ) ) std::swap( ax[0], ax[1] );
	if ( comp( ext, ax[1] ) < comp( ext, ax[2] ) ) std::swap( ax[1], ax[2] );
	if ( comp( ext, ax[0] ) < comp( ext, ax[1] ) ) std::swap( ax[0], ax[1] );
	const float eb = comp( ext, ax[0] ), em = comp( ext, ax[1] ), es = comp( ext, ax[2] );

	int32_t px	 = f2i( log2f( eb / em ) );
	int32_t py	 = f2i( log2f( em / es ) );
	int32_t pz	 = f2i( log2f( eb / es ) );
	int32_t swap = wsub( pz, wadd( px, py ) );

	px = std::min( px, 30 );
	py = std::min( int32_t( uint32_t( py ) * 2u ), 30 - px ) / 2;

	int32_t preSum = px + py * 2;
	if ( preSum != 30 )
		preSum = wadd( preSum, swap );
	else
		swap = 0;

	int32_t bx, by, bz;
	bz = ( es != 0.0f ) ? std::max( 0, wsub( 30, preSum ) / 3 ) : 0;
	if ( swap > 0 )
	{
		bx = std::max( 0, wadd( wadd( wadd( wsub( wsub( 30, bz ), preSum ) / 2, py ), px ), 1 ) );
		by = 30 - bx - bz;
	}
	else
	{
		by = std::max( 0, wadd( wsub( wsub( 30, bz ), preSum ) / 2, py ) );
		bx = 30 - by - bz;
	}

	auto quant = []( float p, int32_t bits ) {
		const uint32_t scale = shl( 1u, bits );
		return std::min( f2u( fmaxf( p * float( scale ), 0.0f ) ), scale - 1u );
	};
	uint32_t cx = quant( comp( pos, ax[0] ), bx );
	uint32_t cy = quant( comp( pos, ax[1] ), by );
	uint32_t cz = quant( comp( pos, ax[2] ), bz );

	uint32_t code = 0;
	int32_t	 d0 = 0, d1 = 0;
	if ( preSum > 0 )
	{
		bx -= px;
		code = shr( cx & shl( shl( 1u, px ) - 1u, bx ), bx );
		code = shl( code, py * 2 );
		bx -= py;
		by -= py;
		const uint32_t t0 = spread2( shr( cx & shl( shl( 1u, py ) - 1u, bx ), bx ) );
		const uint32_t t1 = spread2( shr( cy & shl( shl( 1u, py ) - 1u, by ), by ) );
		code |= t0 * 2 + t1;
		if ( swap > 0 )
		{
			code <<= 1;
			bx -= 1;
			code |= shr( cx & shl( 1u, bx ), bx );
		}
		code = shl( code, bx + by + bz );
		cx &= shl( 1u, bx ) - 1u;
		cy &= shl( 1u, by ) - 1u;
		if ( swap > 0 )
		{
			d0 = by - bx;
			cx = shl( cx, d0 );
			d1 = by - bz;
			cz = shl( cz, d1 );
		}
		else
		{
			d0 = bx - by;
			cy = shl( cy, d0 );
			d1 = bx - bz;
			cz = shl( cz, d1 );
		}
	}
	if ( bz == 0 )
	{
		code |= spread2( cx ) * 2 + spread2( cy );
	}
	else
	{
		cx = cx ? spread3( cx ) : 0;
		cy = cy ? spread3( cy ) : 0;
		cz = cz ? spread3( cz ) : 0;
		if ( swap > 0 )
			code |= shr( cy * 4 + cx * 2 + cz, d0 + d1 );
		else
			code |= shr( cx * 4 + cy * 2 + cz, d0 + d1 );
	}
	return code;
}

// ---------------------------------------------------------------------------------------------
// LBVH topology + refit (LbvhBuilderKernels.h:44-134, MortonCode.h:279-285), sequential.
inline uint64_t delta( const uint32_t* keys, int64_t a, int64_t n )
{
	if ( a < 0 || a + 1 >= n ) return ~0ull;
	const uint64_t x = ( uint64_t( keys[a] ) << 32 ) | uint64_t( a );
	const uint64_t y = ( uint64_t( keys[a + 1] ) << 32 ) | uint64_t( a + 1 );
	return x ^ y;
}

// Recursive-free construction: process ranges bottom-up exactly as the parent rule dictates.
// Returns the root internal node id.
uint32_t buildTopology(
	const uint32_t* keys, uint32_t n, const Box* leafBoxes, uint32_t leafType, std::vector<ScratchNode>& nodes )
{
	nodes.assign( n > 1 ? n - 1 : 0, ScratchNode{ emptyBox(), InvalidValue, InvalidValue } );
	// other end of the range of the first child that reached each internal node
	std::vector<uint32_t> pending( n, InvalidValue );
	uint32_t			  root = InvalidValue;
	for ( uint32_t leaf = 0; leaf < n; ++leaf )
	{
		uint32_t i = leaf, j = leaf + 1;
		uint32_t self = makeNodeIndex( leaf, leafType );
		while ( true )
		{
			if ( i == 0 && j == n )
			{
				root = nodeAddr( self );
				break;
			}
			uint32_t parent;
			bool	 left;
			if ( i == 0 || ( j != n && delta( keys, int64_t( j ) - 1, n ) < delta( keys, int64_t( i ) - 1, n ) ) )
			{
				parent = j - 1, left = true;
				nodes[parent].child0 = self;
			}
			else
			{
				parent = i - 1, left = false;
				nodes[parent].child1 = self;
			}
			if ( pending[parent] == InvalidValue )
			{
				pending[parent] = left ? i : j; // first arrival: remember our far end and stop
				break;
			}
			if ( left )
				j = pending[parent];
			else
				i = pending[parent];
			ScratchNode& p = nodes[parent];
			auto		 childBox = [&]( uint32_t c ) { return isBoxIndex( c ) ? nodes[nodeAddr( c )].box : leafBoxes[nodeAddr( c )]; };
			p.box		 = unite( childBox( p.child0 ), childBox( p.child1 ) );
			self		 = makeNodeIndex( parent, BoxType );
		}
	}
	return root;
}

// ---------------------------------------------------------------------------------------------
// Collapse to 4-wide nodes + leaf packing (BvhBuilderKernels.h:137-193, 1088-1258, 1441-1479).
// Canonical addressing: box nodes in FIFO (breadth-first, slot order) allocation order; leaves
// in (box node address, slot) order.
struct Wide
{
	std::vector<Box4Node> boxes;
	std::vector<uint32_t> leafRefs; // reference index (sorted position / scratch leaf address) per packed leaf
};

void collapse(
	const std::vector<ScratchNode>& bin, const Box* leafBoxes, uint32_t n, uint32_t rootAddr, uint32_t leafType, Wide& out )
{
	struct Task
	{
		uint32_t node, addr, parent;
	};
	std::deque<Task> queue;
	out.boxes.clear();
	out.leafRefs.clear();
	out.boxes.push_back( Box4Node{} );
	queue.push_back( Task{ makeNodeIndex( rootAddr, BoxType ), 0, InvalidValue } );
	auto boxOf = [&]( uint32_t c ) { return isBoxIndex( c ) ? bin[nodeAddr( c )].box : leafBoxes[nodeAddr( c )]; };
	std::vector<std::pair<uint32_t, uint32_t>> leafSlots; // (box addr, slot) in allocation order
	while ( !queue.empty() )
	{
		const Task t = queue.front();
		queue.pop_front();
		const ScratchNode& b		= bin[nodeAddr( t.node )];
		uint32_t		   child[4] = { b.child0, b.child1, InvalidValue, InvalidValue };
		Box				   cbox[4]	= { boxOf( b.child0 ), boxOf( b.child1 ), emptyBox(), emptyBox() };
		uint32_t		   count	= 2;
		while ( count < 4 )
		{
			float best	   = -FltMax;
			int	  bestSlot = -1;
			for ( uint32_t s = 0; s < count; ++s )
			{
				if ( !isBoxIndex( child[s] ) ) continue;
				const float a = area( cbox[s] );
				if ( a > best ) best = a, bestSlot = int( s );
			}
			if ( bestSlot < 0 || best < 0.0f ) break;
			const ScratchNode& open = bin[nodeAddr( child[bestSlot] )];
			child[bestSlot]			= open.child0;
			cbox[bestSlot]			= boxOf( open.child0 );
			child[count]			= open.child1;
			cbox[count]				= boxOf( open.child1 );
			++count;
		}
		Box4Node node;
		memset( &node, 0, sizeof( node ) );
		node.parentAddr	   = t.parent;
		node.updateCounter = 0;
		node.childCount	   = count;
		for ( uint32_t s = 0; s < 4; ++s )
		{
			if ( s >= count )
			{
				node.child[s] = InvalidValue;
				node.box[s]	  = emptyBox();
				continue;
			}
			node.box[s] = cbox[s];
			if ( isBoxIndex( child[s] ) )
			{
				const uint32_t addr = uint32_t( out.boxes.size() );
				out.boxes.push_back( Box4Node{} );
				queue.push_back( Task{ child[s], addr, t.addr } );
				node.child[s] = makeNodeIndex( addr, BoxType );
			}
			else
			{
				node.child[s] = child[s]; // patched below
				leafSlots.push_back( { t.addr, s } );
			}
		}
		out.boxes[t.addr] = node;
	}
	// leaves in (box address, slot) order: FIFO processing order == ascending box address
	std::stable_sort( leafSlots.begin(), leafSlots.end() );
	for ( auto& ls : leafSlots )
	{
		uint32_t& c = out.boxes[ls.first].child[ls.second];
		out.leafRefs.push_back( nodeAddr( c ) );
		c = makeNodeIndex( uint32_t( out.leafRefs.size() - 1 ), leafType );
	}
	(void)n;
}

// ---------------------------------------------------------------------------------------------
// The geometry object the tests poke at.
struct Geometry
{
	uint32_t					  primCount = 0;
	uint32_t					  leafType	= TriangleType;
	uint32_t					  geomType	= InvalidValue; // user geometry type of custom primitives (hiprtGeometryBuildInput::geomType)
	std::vector<uint32_t>		  pairs;	 // 2 per primitive (triangle meshes)
	std::vector<TrianglePairNode> primNodes; // per primitive, input order
	std::vector<Box>			  primBoxes; // per primitive, input order
	Box							  sceneBox;
	std::vector<uint32_t>		  keys, sortedKeys, sortedVals;
	std::vector<Box>			  sortedBoxes;
	std::vector<ScratchNode>	  scratch;
	uint32_t					  root = InvalidValue;
	Wide						  wide;
	std::vector<TrianglePairNode> leaves;	   // packed triangle leaves
	std::vector<CustomNode>		  customLeaves; // packed custom leaves
};

void finishBuild( Geometry& g )
{
	const uint32_t n = g.primCount;
	g.sceneBox		 = emptyBox();
	for ( uint32_t i = 0; i < n; ++i )
		g.sceneBox = unite( g.sceneBox, g.primBoxes[i] );
	// BvhBuilderKernels.h:508-525
	const V3 ext = g.sceneBox.hi - g.sceneBox.lo;
	g.keys.resize( n );
	for ( uint32_t i = 0; i < n; ++i )
	{
		const V3 c	 = ( g.primBoxes[i].hi + g.primBoxes[i].lo ) * 0.5f;
		const V3 d	 = c - g.sceneBox.lo;
		const V3 pos = mk( d.x / ext.x, d.y / ext.y, d.z / ext.z );
		g.keys[i]	 = extendedMorton( pos, ext );
	}
}

void sortAndBuild( Geometry& g )
{
	const uint32_t n = g.primCount;
	g.sortedVals.resize( n );
	std::iota( g.sortedVals.begin(), g.sortedVals.end(), 0u );
	std::stable_sort(
		g.sortedVals.begin(), g.sortedVals.end(), [&]( uint32_t a, uint32_t b ) { return g.keys[a] < g.keys[b]; } );
	g.sortedKeys.resize( n );
	g.sortedBoxes.resize( n );
	for ( uint32_t k = 0; k < n; ++k )
	{
		g.sortedKeys[k]	 = g.keys[g.sortedVals[k]];
		g.sortedBoxes[k] = g.primBoxes[g.sortedVals[k]];
	}
}

void wideAndPack( Geometry& g )
{
	const uint32_t n = g.primCount;
	if ( n == 1 )
	{
		// BvhBuilderKernels.h:296-331
		Box4Node node;
		memset( &node, 0, sizeof( node ) );
		node.parentAddr = InvalidValue;
		node.childCount = 1;
		for ( int s = 0; s < 4; ++s )
			node.child[s] = InvalidValue, node.box[s] = emptyBox();
		node.child[0] = makeNodeIndex( 0, g.leafType );
		node.box[0]	  = g.primBoxes[0];
		g.wide.boxes.assign( 1, node );
		g.wide.leafRefs.assign( 1, 0 );
	}
	else
	{
		collapse( g.scratch, g.sortedBoxes.data(), n, g.root, g.leafType, g.wide );
	}
	g.leaves.clear();
	g.customLeaves.clear();
	for ( uint32_t ref : g.wide.leafRefs )
	{
		const uint32_t prim = ( n == 1 ) ? 0 : g.sortedVals[ref];
		if ( g.leafType == TriangleType )
			g.leaves.push_back( g.primNodes[prim] );
		else
			g.customLeaves.push_back( CustomNode{ prim } );
	}
}

// ---------------------------------------------------------------------------------------------
// Traversal (hiprt_device_impl.h:338-411, 413-634, 779-850; Aabb.h:92-101; Triangle.h:91-130)
struct Slab
{
	float t0, t1;
};
inline Slab slab( const Box& b, V3 o, V3 invD, float maxT )
{
	const V3	f  = ( b.hi - o ) * invD;
	const V3	n  = ( b.lo - o ) * invD;
	const V3	hi = vmax( f, n );
	const V3	lo = vmin( f, n );
	const float t1 = fminf( fminf( fminf( hi.x, hi.y ), hi.z ), maxT );
	const float t0 = fmaxf( fmaxf( fmaxf( lo.x, lo.y ), lo.z ), 0.0f );
	return Slab{ t0, t1 };
}

// Triangle.h:96-130
bool hitTriangle( V3 v0, V3 v1, V3 v2, const hiprtRay& ray, uint32_t flags, float& t, hiprtFloat2& uv )
{
	const V3	e1	  = v1 - v0;
	const V3	e2	  = v2 - v0;
	const V3	s1	  = cross( ray.direction, e2 );
	const float denom = dot( s1, e1 );
	if ( denom == 0.0f ) return false;
	const float inv = 1.0f / denom;
	const V3	d	= ray.origin - v0;
	float		b[3];
	b[1]		  = dot( d, s1 ) * inv;
	const V3 s2	  = cross( d, e1 );
	b[2]		  = dot( ray.direction, s2 ) * inv;
	const float t0 = dot( e2, s2 ) * inv;
	if ( b[1] < 0.0f || b[1] > 1.0f || b[2] < 0.0f || b[1] + b[2] > 1.0f || t0 < ray.minT || t0 > ray.maxT ) return false;
	b[0] = 1.0f - b[1] - b[2];
	uv.x = b[flags & 3];
	uv.y = b[( flags >> 2 ) & 3];
	t	 = t0;
	return true;
}

bool hitPairTriangle( const TrianglePairNode& n, uint32_t which, const hiprtRay& ray, hiprtHit& hit )
{
	if ( which == 1 && n.primIndex0 == n.primIndex1 ) return false;
	const V3	   a = which ? n.v1 : n.v0, b = which ? n.v3 : n.v1, c = n.v2;
	const uint32_t flags = n.flags >> ( which * 8 );
	if ( !hitTriangle( a, b, c, ray, flags, hit.t, hit.uv ) ) return false;
	hit.primID = which ? n.primIndex1 : n.primIndex0;
	hit.normal = ( ( flags >> 5 ) & 1 ) ? cross( c - a, b - a ) : cross( b - a, c - a );
	return true;
}

struct BvhView
{
	const Box4Node*			boxes;
	const TrianglePairNode* tris;
};

// Order the (up to) four children of a box node: hit children by entry distance through the
// 5-comparator network, misses last (hiprt_device_impl.h:344-371).
void orderChildren( const Box4Node& node, V3 o, V3 invD, float maxT, uint32_t res[4] )
{
	float d[4];
	for ( int s = 0; s < 4; ++s )
	{
		const Slab sl = slab( node.box[s], o, invD, maxT );
		d[s]		  = sl.t0;
		res[s]		  = ( sl.t0 <= sl.t1 ) ? node.child[s] : InvalidValue;
	}
	auto cswap = [&]( int a, int b ) {
		if ( ( res[b] != InvalidValue && d[b] < d[a] ) || res[a] == InvalidValue )
		{
			std::swap( res[a], res[b] );
			std::swap( d[a], d[b] );
		}
	};
	cswap( 0, 2 ), cswap( 1, 3 ), cswap( 0, 1 ), cswap( 2, 3 ), cswap( 1, 2 );
}

// tieMode 0: the reference's scalar rule (a later-visited hit with t <= maxT replaces; inside a
// pair triangle 1 replaces triangle 0 only when strictly nearer). tieMode 1: order-independent
// rule min over (t, primID), which is what the B200 kernels implement.
// box visits of traceClosest in which no child was hit (statistic for tools/simt_model.py: an upper bound of what culling popped
// nodes by their stored entry distance could save)
static uint64_t g_fruitlessBoxVisits = 0;

hiprtHit traceClosest( const BvhView& bvh, const hiprtRay& inRay, int tieMode, uint32_t* boxVisits, uint32_t* leafVisits )
{
	hiprtRay			  ray  = inRay;
	const V3			  invD = mk( 1.0f / ray.direction.x, 1.0f / ray.direction.y, 1.0f / ray.direction.z );
	hiprtHit			  best;
	std::vector<uint32_t> stack;
	stack.push_back( InvalidValue );
	uint32_t node = RootIndex;
	while ( node != InvalidValue )
	{
		if ( isBoxIndex( node ) )
		{
			if ( boxVisits ) ++*boxVisits;
			uint32_t res[4];
			orderChildren( bvh.boxes[nodeAddr( node )], ray.origin, invD, ray.maxT, res );
			for ( int s = 3; s >= 1; --s )
				if ( res[s] != InvalidValue ) stack.push_back( res[s] );
			if ( res[0] != InvalidValue )
				node = res[0];
			else
			{
				if ( boxVisits ) ++g_fruitlessBoxVisits; // statistic runs are single-threaded (they pass counters)
				node = stack.back();
				stack.pop_back();
			}
			continue;
		}
		if ( leafVisits ) ++*leafVisits;
		const TrianglePairNode& leaf = bvh.tris[nodeAddr( node )];
		// both triangles are tested against the ray as it is on entry (hiprt_device_impl.h:574-587)
		hiprtHit cand[2];
		bool	 ok[2];
		for ( uint32_t which = 0; which < 2; ++which )
			ok[which] = hitPairTriangle( leaf, which, ray, cand[which] );
		if ( tieMode == 0 )
		{
			int pick = ok[0] ? 0 : -1;
			if ( ok[1] && ( !ok[0] || cand[0].t > cand[1].t ) ) pick = 1;
			if ( pick >= 0 ) best = cand[pick];
		}
		else
		{
			for ( uint32_t which = 0; which < 2; ++which )
			{
				const hiprtHit& c = cand[which];
				if ( ok[which] && ( !best.hasHit() || c.t < best.t || ( c.t == best.t && c.primID < best.primID ) ) ) best = c;
			}
		}
		if ( best.hasHit() ) ray.maxT = best.t;
		node = stack.back();
		stack.pop_back();
	}
	return best;
}

// All hits along the ray in traversal order (the any-hit iterator drained to completion).
void traceAll( const BvhView& bvh, const hiprtRay& ray, std::vector<hiprtHit>& hits )
{
	const V3			  invD = mk( 1.0f / ray.direction.x, 1.0f / ray.direction.y, 1.0f / ray.direction.z );
	std::vector<uint32_t> stack;
	stack.push_back( InvalidValue );
	uint32_t node = RootIndex;
	while ( node != InvalidValue )
	{
		if ( isBoxIndex( node ) )
		{
			uint32_t res[4];
			orderChildren( bvh.boxes[nodeAddr( node )], ray.origin, invD, ray.maxT, res );
			for ( int s = 3; s >= 1; --s )
				if ( res[s] != InvalidValue ) stack.push_back( res[s] );
			if ( res[0] != InvalidValue )
				node = res[0];
			else
			{
				node = stack.back();
				stack.pop_back();
			}
			continue;
		}
		const TrianglePairNode& leaf = bvh.tris[nodeAddr( node )];
		for ( uint32_t which = 0; which < 2; ++which )
		{
			hiprtHit h;
			if ( hitPairTriangle( leaf, which, ray, h ) ) hits.push_back( h );
		}
		node = stack.back();
		stack.pop_back();
	}
}
// ---------------------------------------------------------------------------------------------
// Scene (top level): instance transforms (Transform.h:41-240, Quaternion.h:30-140,
// QrDecomposition.h:30-135), instance boxes (InstanceList.h:104-121, Transform.h:460-517),
// instance leaves (BvhNode.h:1109-1182), two-level traversal (hiprt_device_impl.h:959-1204).
struct Quat
{
	float x, y, z, w;
};
struct Frame
{
	Quat  rot{ 0, 0, 0, 1 };
	V3	  scale{ 1, 1, 1 }, shear{ 0, 0, 0 }, trans{ 0, 0, 0 };
	float time = 0;
	bool  identity() const
	{
		return scale.x == 1 && scale.y == 1 && scale.z == 1 && shear.x == 0 && shear.y == 0 && shear.z == 0 && trans.x == 0 &&
			   trans.y == 0 && trans.z == 0 && rot.w == 1;
	}
};

inline float hypot3( V3 a )
{
	const float x = fabsf( a.x ), y = fabsf( a.y ), z = fabsf( a.z );
	const float d = x < y ? ( y < z ? z : y ) : ( x < z ? z : x );
	return d != 0.0f ? d * sqrtf( ( x / d ) * ( x / d ) + ( y / d ) * ( y / d ) + ( z / d ) * ( z / d ) ) : 0.0f;
}
inline Quat qmul( Quat a, Quat b )
{
	const V3 av = mk( a.x, a.y, a.z ), bv = mk( b.x, b.y, b.z );
	const V3 c	= cross( av, bv ) + bv * a.w + av * b.w;
	return Quat{ c.x, c.y, c.z, a.w * b.w - dot( av, bv ) };
}
inline V3 qrotate( Quat q, V3 p )
{
	const Quat r = qmul( qmul( q, Quat{ p.x, p.y, p.z, 0 } ), Quat{ -q.x, -q.y, -q.z, q.w } );
	return mk( r.x, r.y, r.z );
}
inline Quat qFromAxisAngle( const float aa[4] )
{
	if ( aa[3] == 0.0f ) return Quat{ 0, 0, 0, 1 };
	const V3	axis = mk( aa[0], aa[1], aa[2] );
	const float len	 = hypot3( axis );
	const float s	 = sinf( aa[3] / 2.0f );
	return Quat{ axis.x / len * s, axis.y / len * s, axis.z / len * s, cosf( aa[3] / 2.0f ) };
}
inline Quat qFromMatrix( const float R[3][3] )
{
	const float tr = R[0][0] + R[1][1] + R[2][2];
	if ( tr > 0.0f )
	{
		const float S = sqrtf( tr + 1.0f ) * 2.0f;
		return Quat{ ( R[2][1] - R[1][2] ) / S, ( R[0][2] - R[2][0] ) / S, ( R[1][0] - R[0][1] ) / S, 0.25f * S };
	}
	if ( R[0][0] > R[1][1] && R[0][0] > R[2][2] )
	{
		const float S = sqrtf( 1.0f + R[0][0] - R[1][1] - R[2][2] ) * 2.0f;
		return Quat{ 0.25f * S, ( R[0][1] + R[1][0] ) / S, ( R[0][2] + R[2][0] ) / S, ( R[2][1] - R[1][2] ) / S };
	}
	if ( R[1][1] > R[2][2] )
	{
		const float S = sqrtf( 1.0f + R[1][1] - R[0][0] - R[2][2] ) * 2.0f;
		return Quat{ ( R[0][1] + R[1][0] ) / S, 0.25f * S, ( R[1][2] + R[2][1] ) / S, ( R[0][2] - R[2][0] ) / S };
	}
	const float S = sqrtf( 1.0f + R[2][2] - R[0][0] - R[1][1] ) * 2.0f;
	return Quat{ ( R[0][2] + R[2][0] ) / S, ( R[1][2] + R[2][1] ) / S, 0.25f * S, ( R[1][0] - R[0][1] ) / S };
}
inline void qToMatrix( Quat q, float R[3][3] )
{
	const float x2 = q.x * q.x, y2 = q.y * q.y, z2 = q.z * q.z;
	R[0][0] = 1.0f - 2.0f * y2 - 2.0f * z2;
	R[0][1] = 2.0f * q.x * q.y - 2.0f * q.w * q.z;
	R[0][2] = 2.0f * q.x * q.z + 2.0f * q.w * q.y;
	R[1][0] = 2.0f * q.x * q.y + 2.0f * q.w * q.z;
	R[1][1] = 1.0f - 2.0f * x2 - 2.0f * z2;
	R[1][2] = 2.0f * q.y * q.z - 2.0f * q.w * q.x;
	R[2][0] = 2.0f * q.x * q.z - 2.0f * q.w * q.y;
	R[2][1] = 2.0f * q.y * q.z + 2.0f * q.w * q.x;
	R[2][2] = 1.0f - 2.0f * x2 - 2.0f * y2;
}
// modified Gram-Schmidt on columns, det(Q) forced positive
void qrDecompose( const float A[3][3], float Q[3][3], float R[3][3] )
{
	for ( int i = 0; i < 3; ++i )
		for ( int j = 0; j < 3; ++j )
			Q[i][j] = A[i][j], R[i][j] = 0.0f;
	for ( int c = 0; c < 3; ++c )
	{
		for ( int p = 0; p < c; ++p )
		{
			const float w = Q[0][p] * Q[0][c] + Q[1][p] * Q[1][c] + Q[2][p] * Q[2][c];
			R[p][c]		  = w;
			for ( int r = 0; r < 3; ++r )
				Q[r][c] -= w * Q[r][p];
		}
		float w = hypot3( mk( Q[0][c], Q[1][c], Q[2][c] ) );
		R[c][c] = w;
		if ( w != 0.0f )
		{
			w = 1.0f / w;
			for ( int r = 0; r < 3; ++r )
				Q[r][c] *= w;
		}
		else
			Q[0][c] = 1.0f;
	}
	const float d = Q[0][0] * Q[1][1] * Q[2][2] + Q[0][1] * Q[1][2] * Q[2][0] + Q[1][0] * Q[2][1] * Q[0][2] - Q[0][2] * Q[1][1] * Q[2][0] -
					Q[0][1] * Q[1][0] * Q[2][2] - Q[1][2] * Q[2][1] * Q[0][0];
	if ( d < 0.0f )
		for ( int k = 0; k < 3; ++k )
			Q[k][0] = -Q[k][0], R[0][k] = -R[0][k];
}

Frame frameFromMatrix( const float* m16 ) // hiprtFrameMatrix: 3x4 row-major + time
{
	const float( *m )[4] = reinterpret_cast<const float( * )[4]>( m16 );
	Frame f;
	const bool ident = m[0][0] == 1 && m[1][1] == 1 && m[2][2] == 1 && m[0][1] == 0 && m[0][2] == 0 && m[0][3] == 0 && m[1][0] == 0 &&
					   m[1][2] == 0 && m[1][3] == 0 && m[2][0] == 0 && m[2][1] == 0 && m[2][3] == 0;
	if ( !ident )
	{
		float A[3][3], Q[3][3], R[3][3];
		for ( int i = 0; i < 3; ++i )
			for ( int j = 0; j < 3; ++j )
				A[i][j] = m[i][j];
		qrDecompose( A, Q, R );
		f.trans = mk( m[0][3], m[1][3], m[2][3] );
		f.rot	= qFromMatrix( Q );
		f.scale = mk( R[0][0], R[1][1], R[2][2] );
		f.shear = mk( R[0][1], R[0][2], R[1][2] );
	}
	f.time = m16[12];
	return f;
}

Frame frameFromSRT( const float* s12 ) // hiprtFrameSRT: rotation(4) scale(3) translation(3) time + pad
{
	Frame f;
	f.rot	= qFromAxisAngle( s12 );
	f.scale = mk( s12[4], s12[5], s12[6] );
	f.trans = mk( s12[7], s12[8], s12[9] );
	f.time	= s12[10];
	return f;
}

V3 frameTransform( const Frame& f, V3 p )
{
	if ( f.identity() ) return p;
	V3 r = p * f.scale;
	r	 = r + mk( p.y * f.shear.x + p.z * f.shear.y, p.z * f.shear.z, 0.0f );
	r	 = qrotate( f.rot, r );
	return r + f.trans;
}

// ---- key-frame (motion blur) transforms: Quaternion.h:110-190 (qtDot, qtNormalize, qtMix), Transform.h:178-240 (SRTFrame::interpolate,
// invTransform), :417-449 (Transform::interpolateFrames, transformRay)
inline float qdot( Quat a, Quat b ) { return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w; }
inline Quat	 qnormalize( Quat q )
{
	const float len = sqrtf( qdot( q, q ) );
	return Quat{ q.x / len, q.y / len, q.z / len, q.w / len };
}
Quat qmix( Quat v0, Quat v1, float t )
{
	v0		= qnormalize( v0 );
	v1		= qnormalize( v1 );
	float d = qdot( v0, v1 );
	if ( d < 0.0f ) v1 = Quat{ -v1.x, -v1.y, -v1.z, -v1.w }, d = -d;
	if ( d > 0.9995f ) return qnormalize( Quat{ v0.x + ( v1.x - v0.x ) * t, v0.y + ( v1.y - v0.y ) * t, v0.z + ( v1.z - v0.z ) * t, v0.w + ( v1.w - v0.w ) * t } );
	const float theta0 = acosf( d ), theta = theta0 * t;
	const float s0 = cosf( theta ) - d * sinf( theta ) / sinf( theta0 ), s1 = sinf( theta ) / sinf( theta0 );
	return Quat{ v0.x * s0 + v1.x * s1, v0.y * s0 + v1.y * s1, v0.z * s0 + v1.z * s1, v0.w * s0 + v1.w * s1 };
}

Frame frameInterpolate( const Frame& f0, const Frame& f1, float t )
{
	Frame f;
	f.scale = f0.scale * ( 1.0f - t ) + f1.scale * t;
	f.shear = f0.shear * ( 1.0f - t ) + f1.shear * t;
	f.trans = f0.trans * ( 1.0f - t ) + f1.trans * t;
	f.rot	= qmix( f0.rot, f1.rot, t );
	return f;
}

// pose of an instance with `count` key frames at `time` (Transform::interpolateFrames)
Frame frameAt( const Frame* frames, uint32_t count, float time )
{
	if ( count == 0 || frames == nullptr ) return Frame();
	Frame f0 = frames[0];
	if ( count == 1 || time == 0.0f || time <= f0.time ) return f0;
	Frame f1 = frames[count - 1];
	if ( time >= f1.time ) return f1;
	for ( uint32_t i = 1; i < count; ++i )
	{
		f1 = frames[i];
		if ( time >= f0.time && time <= f1.time ) break;
		f0 = f1;
	}
	return frameInterpolate( f0, f1, ( time - f0.time ) / ( f1.time - f0.time ) );
}

// world -> object of a point (SRTFrame::invTransform; the shear terms read the ORIGINAL point's z, as the reference does)
V3 frameInvTransform( const Frame& f, V3 p )
{
	if ( f.identity() ) return p;
	V3 r = p - f.trans;
	r	 = qrotate( Quat{ -f.rot.x, -f.rot.y, -f.rot.z, f.rot.w }, r );
	r	 = mk( r.x / f.scale.x, r.y / f.scale.y, r.z / f.scale.z );
	r.y -= p.z * f.shear.z / f.scale.y;
	r.x -= ( f.shear.x * r.y + f.shear.y * r.z ) / f.scale.x;
	return r;
}

void frameInverseMatrix( const Frame& f, float out[3][4] )
{
	float Q[3][3];
	qToMatrix( f.rot, Q );
	float Ri[3][3] = {};
	Ri[0][0]	   = 1.0f / f.scale.x;
	Ri[1][1]	   = 1.0f / f.scale.y;
	Ri[2][2]	   = 1.0f / f.scale.z;
	Ri[0][1]	   = -f.shear.x / ( f.scale.x * f.scale.y );
	Ri[0][2]	   = ( f.shear.x * f.shear.z - f.shear.y * f.scale.y ) / ( f.scale.x * f.scale.y * f.scale.z );
	Ri[1][2]	   = -f.shear.z / ( f.scale.y * f.scale.z );
	for ( int i = 0; i < 3; ++i )
	{
		for ( int j = 0; j < 3; ++j )
			out[i][j] = dot( mk( Ri[i][0], Ri[i][1], Ri[i][2] ), mk( Q[j][0], Q[j][1], Q[j][2] ) );
		out[i][3] = -dot( mk( out[i][0], out[i][1], out[i][2] ), f.trans );
	}
}

struct SceneObj
{
	uint32_t				  count = 0;
	std::vector<Geometry*>	  geoms; // per instance (null when the instance is a scene)
	std::vector<SceneObj*>	  scenes; // per instance: the nested scene of a hiprtInstanceTypeScene instance, else null
	std::vector<Frame>		  frames;
	std::vector<uint32_t>	  masks;
	std::vector<Box>		  primBoxes, sortedBoxes;
	std::vector<uint32_t>	  keys, sortedKeys, sortedVals;
	std::vector<ScratchNode>  scratch;
	uint32_t				  root = InvalidValue;
	Wide					  wide;
	std::vector<InstanceNode> leaves; // packed order
	std::vector<uint32_t>	  leafGeom; // geometry slot per packed leaf
	Box						  sceneBox;
};

// (a geometry and a scene are bounded alike: through the child boxes of their root node, InstanceList.h:104-121)
Box instanceBox( const Wide& wide, const Frame& f )
{
	const Box4Node& root = wide.boxes[0];
	Box				out	 = emptyBox();
	for ( uint32_t c = 0; c < root.childCount; ++c )
	{
		const Box& b = root.box[c];
		for ( int k = 0; k < 8; ++k )
			out = grow( out, frameTransform( f, mk( ( k & 4 ) ? b.hi.x : b.lo.x, ( k & 2 ) ? b.hi.y : b.lo.y, ( k & 1 ) ? b.hi.z : b.lo.z ) ) );
	}
	return out;
}

hiprtRay toObjectSpace( const InstanceNode& n, const hiprtRay& ray )
{
	hiprtRay out;
	const V3 o = ray.origin, d = ray.direction;
	out.origin = mk( dot( mk( n.invMatrix[0][0], n.invMatrix[0][1], n.invMatrix[0][2] ), o ),
					 dot( mk( n.invMatrix[1][0], n.invMatrix[1][1], n.invMatrix[1][2] ), o ),
					 dot( mk( n.invMatrix[2][0], n.invMatrix[2][1], n.invMatrix[2][2] ), o ) ) +
				 mk( n.invMatrix[0][3], n.invMatrix[1][3], n.invMatrix[2][3] );
	out.direction = mk( dot( mk( n.invMatrix[0][0], n.invMatrix[0][1], n.invMatrix[0][2] ), d ),
						dot( mk( n.invMatrix[1][0], n.invMatrix[1][1], n.invMatrix[1][2] ), d ),
						dot( mk( n.invMatrix[2][0], n.invMatrix[2][1], n.invMatrix[2][2] ), d ) );
	out.minT	  = ray.minT;
	out.maxT	  = ray.maxT;
	return out;
}

// Closest hit through a two-level structure with the order-independent rule min (t, instanceID, primID)
// (tieMode 1) or the reference's visiting-order rule (tieMode 0).
hiprtHit traceSceneClosest( const SceneObj& s, const hiprtRay& worldRay, uint32_t rayMask, int tieMode )
{
	hiprtRay			  ray  = worldRay;
	V3					  invD = mk( 1.0f / ray.direction.x, 1.0f / ray.direction.y, 1.0f / ray.direction.z );
	hiprtHit			  best;
	std::vector<uint32_t> stack;
	stack.push_back( InvalidValue );
	uint32_t		node	 = RootIndex;
	const Geometry* geom	 = nullptr;
	uint32_t		instance = InvalidValue;
	auto			pop		 = [&]() {
		   uint32_t n = stack.back();
		   stack.pop_back();
		   while ( n == InvalidValue && !stack.empty() )
		   {
			   // leave the instance: world ray again, distance kept
			   const float maxT = ray.maxT;
			   ray				= worldRay;
			   ray.maxT			= maxT;
			   invD				= mk( 1.0f / ray.direction.x, 1.0f / ray.direction.y, 1.0f / ray.direction.z );
			   geom				= nullptr;
			   instance			= InvalidValue;
			   n				= stack.back();
			   stack.pop_back();
		   }
		   return n;
	};
	while ( node != InvalidValue )
	{
		if ( isBoxIndex( node ) )
		{
			const Box4Node* boxes = geom ? geom->wide.boxes.data() : s.wide.boxes.data();
			uint32_t		res[4];
			orderChildren( boxes[nodeAddr( node )], ray.origin, invD, ray.maxT, res );
			for ( int k = 3; k >= 1; --k )
				if ( res[k] != InvalidValue ) stack.push_back( res[k] );
			node = res[0] != InvalidValue ? res[0] : pop();
			continue;
		}
		if ( !geom )
		{
			const uint32_t		addr = nodeAddr( node );
			const InstanceNode& in	 = s.leaves[addr];
			if ( in.mask & rayMask )
			{
				const float maxT = ray.maxT;
				if ( !in.identity ) ray = toObjectSpace( in, ray );
				ray.maxT = maxT;
				invD	 = mk( 1.0f / ray.direction.x, 1.0f / ray.direction.y, 1.0f / ray.direction.z );
				geom	 = s.geoms[s.leafGeom[addr]];
				instance = in.primIndex;
				stack.push_back( InvalidValue );
				node = RootIndex;
				continue;
			}
			node = pop();
			continue;
		}
		const TrianglePairNode& leaf = geom->leaves[nodeAddr( node )];
		hiprtHit				cand[2];
		bool					ok[2];
		for ( uint32_t which = 0; which < 2; ++which )
		{
			ok[which]				= hitPairTriangle( leaf, which, ray, cand[which] );
			cand[which].instanceID = instance;
		}
		if ( tieMode == 0 )
		{
			int pick = ok[0] ? 0 : -1;
			if ( ok[1] && ( !ok[0] || cand[0].t > cand[1].t ) ) pick = 1;
			if ( pick >= 0 ) best = cand[pick];
		}
		else
		{
			for ( uint32_t which = 0; which < 2; ++which )
			{
				const hiprtHit& c = cand[which];
				if ( !ok[which] ) continue;
				bool better = !best.hasHit() || c.t < best.t;
				if ( !better && c.t == best.t )
					better = c.instanceID != best.instanceID ? c.instanceID < best.instanceID : c.primID < best.primID;
				if ( better ) best = c;
			}
		}
		if ( best.hasHit() ) ray.maxT = best.t;
		node = pop();
	}
	return best;
}

// ---------------------------------------------------------------------------------------------
// General traversal: custom primitives through a caller-supplied intersection function (the oracle's stand-in for a
// hiprtFuncTable entry, hiprt_device_impl.h:767-776), scenes instanced in scenes (instance stack, hiprt_device_impl.h:1105-1193),
// and the any-hit iterator drained to completion (every hit, in visiting order) next to the closest-hit query.
// The function gets the ray in the primitive's object space and hit.primID set; it fills hit.t (and uv / normal) and returns
// non-zero for a hit.
typedef int ( *CustomIntersectFn )( const hiprtRay* ray, uint32_t geomType, void* user, hiprtHit* hit );

struct Walk
{
	bool				   closest;	 // false: collect every hit
	int					   tieMode;
	uint32_t			   rayMask;
	CustomIntersectFn	   fn;
	void*				   user;
	hiprtHit			   best;
	std::vector<hiprtHit>* all;
	uint32_t			   ids[hiprtMaxInstanceLevels];

	void report( hiprtHit h, uint32_t level, hiprtRay& ray, bool custom )
	{
		for ( uint32_t l = 0; l < hiprtMaxInstanceLevels; ++l )
			h.instanceIDs[l] = l < level ? ids[l] : InvalidValue;
		if ( !closest )
		{
			all->push_back( h );
			return;
		}
		bool better = custom || tieMode == 0 || !best.hasHit() || h.t < best.t; // (a custom hit always replaces: :767-776, 838-842)
		if ( !better && h.t == best.t )
		{
			better = h.primID < best.primID;
			for ( uint32_t l = 0; l < hiprtMaxInstanceLevels; ++l )
				if ( h.instanceIDs[l] != best.instanceIDs[l] )
				{
					better = h.instanceIDs[l] < best.instanceIDs[l];
					break;
				}
		}
		if ( better ) best = h, ray.maxT = h.t;
	}
};

void walkGeometry( const Geometry& g, hiprtRay& ray, uint32_t level, Walk& w )
{
	const V3			  invD = mk( 1.0f / ray.direction.x, 1.0f / ray.direction.y, 1.0f / ray.direction.z );
	std::vector<uint32_t> stack{ InvalidValue };
	uint32_t			  node = RootIndex;
	while ( node != InvalidValue )
	{
		if ( isBoxIndex( node ) )
		{
			uint32_t res[4];
			orderChildren( g.wide.boxes[nodeAddr( node )], ray.origin, invD, ray.maxT, res );
			for ( int k = 3; k >= 1; --k )
				if ( res[k] != InvalidValue ) stack.push_back( res[k] );
			if ( res[0] != InvalidValue )
			{
				node = res[0];
				continue;
			}
		}
		else if ( g.leafType == TriangleType )
		{
			const TrianglePairNode& leaf = g.leaves[nodeAddr( node )];
			hiprtHit				cand[2];
			bool					ok[2];
			for ( uint32_t which = 0; which < 2; ++which )
				ok[which] = hitPairTriangle( leaf, which, ray, cand[which] );
			if ( w.closest && w.tieMode == 0 )
			{
				// the reference's pair rule: the second triangle replaces the first only when strictly nearer
				int pick = ok[0] ? 0 : -1;
				if ( ok[1] && ( !ok[0] || cand[0].t > cand[1].t ) ) pick = 1;
				if ( pick >= 0 ) w.report( cand[pick], level, ray, false );
			}
			else
				for ( uint32_t which = 0; which < 2; ++which )
					if ( ok[which] ) w.report( cand[which], level, ray, false );
		}
		else
		{
			hiprtHit h;
			h.primID = g.customLeaves[nodeAddr( node )].primIndex;
			if ( w.fn && w.fn( &ray, g.geomType, w.user, &h ) ) w.report( h, level, ray, true );
		}
		node = stack.back();
		stack.pop_back();
	}
}

void walkScene( const SceneObj& s, hiprtRay& ray, uint32_t level, Walk& w )
{
	std::vector<uint32_t> stack{ InvalidValue };
	uint32_t			  node = RootIndex;
	while ( node != InvalidValue )
	{
		if ( isBoxIndex( node ) )
		{
			const V3 invD = mk( 1.0f / ray.direction.x, 1.0f / ray.direction.y, 1.0f / ray.direction.z );
			uint32_t res[4];
			orderChildren( s.wide.boxes[nodeAddr( node )], ray.origin, invD, ray.maxT, res );
			for ( int k = 3; k >= 1; --k )
				if ( res[k] != InvalidValue ) stack.push_back( res[k] );
			if ( res[0] != InvalidValue )
			{
				node = res[0];
				continue;
			}
		}
		else
		{
			const uint32_t		addr = nodeAddr( node );
			const InstanceNode& in	 = s.leaves[addr];
			if ( ( in.mask & w.rayMask ) && level < hiprtMaxInstanceLevels )
			{
				hiprtRay inner = in.identity ? ray : toObjectSpace( in, ray );
				inner.maxT	   = ray.maxT;
				w.ids[level]   = in.primIndex;
				const uint32_t k = s.leafGeom[addr];
				if ( s.scenes[k] )
					walkScene( *s.scenes[k], inner, level + 1, w );
				else
					walkGeometry( *s.geoms[k], inner, level + 1, w );
				ray.maxT = inner.maxT; // the distance found inside carries over (t is invariant under the instance transform)
			}
		}
		node = stack.back();
		stack.pop_back();
	}
}
} // namespace

// =================================================================================================
// C interface (ctypes)
// hiprtHit leaves the deeper instance ids, and uv / normal of a miss, uninitialised (as the reference's does): give the words a
// fixed value so that two runs of the oracle compare equal byte for byte.
static hiprtHit canonical( hiprtHit h, bool scene )
{
	if ( !scene ) h.instanceIDs[0] = InvalidValue;
	for ( uint32_t l = 1; l < hiprtMaxInstanceLevels; ++l )
		h.instanceIDs[l] = InvalidValue;
	if ( !h.hasHit() ) h.uv = { 0.0f, 0.0f }, h.normal = { 0.0f, 0.0f, 0.0f };
	return h;
}

// Rays are independent: the plain trace entry points split them over the host's cores (test infrastructure only; the per-ray
// arithmetic is untouched, so results do not depend on the thread count).
template <typename F>
static void forEachRay( uint32_t n, F&& body )
{
	const uint32_t threads = std::max( 1u, std::min( std::thread::hardware_concurrency(), n / 512u ) );
	if ( threads <= 1 )
	{
		for ( uint32_t i = 0; i < n; ++i )
			body( i );
		return;
	}
	std::vector<std::thread> pool;
	const uint32_t			 chunk = ( n + threads - 1 ) / threads;
	for ( uint32_t t = 0; t < threads; ++t )
		pool.emplace_back( [&, t] {
			for ( uint32_t i = t * chunk, e = std::min( n, ( t + 1 ) * chunk ); i < e; ++i )
				body( i );
		} );
	for ( std::thread& th : pool )
		th.join();
}

extern "C" {

// Pair triangles of an indexed mesh; pairsOut holds 2*triCount uint32. Returns the pair count.
uint32_t orc_pair_triangles( const uint32_t* indices, uint32_t triCount, uint32_t triStrideBytes, uint32_t* pairsOut )
{
	Mesh m{ nullptr, 0, 0, reinterpret_cast<const uint8_t*>( indices ), triCount, triStrideBytes };
	return pairTriangles( m, pairsOut );
}

uint32_t orc_morton( const float pos[3], const float ext[3] )
{
	return extendedMorton( mk( pos[0], pos[1], pos[2] ), mk( ext[0], ext[1], ext[2] ) );
}

// Build a triangle geometry with the LBVH pipeline. pairs == null: pair automatically when
// `allowPairing` and the mesh is indexed with > 2 triangles (TriangleMesh.h:246), else one
// primitive per triangle.
void* orc_geometry_create(
	const float*	verts,
	uint32_t		vertCount,
	uint32_t		vertStride,
	const uint32_t* indices,
	uint32_t		triCount,
	uint32_t		triStride,
	const uint32_t* pairs,
	uint32_t		pairCount,
	int				allowPairing )
{
	Mesh mesh{ reinterpret_cast<const uint8_t*>( verts ), vertCount, vertStride, reinterpret_cast<const uint8_t*>( indices ),
			   triCount, triStride };
	if ( !indices || triCount == 0 ) mesh.triCount = vertCount / 3;
	Geometry* g = new Geometry;
	if ( pairs && pairCount )
		g->pairs.assign( pairs, pairs + 2 * size_t( pairCount ) );
	else if ( allowPairing && indices && mesh.triCount > 2 )
	{
		g->pairs.resize( 2 * size_t( mesh.triCount ) );
		g->pairs.resize( 2 * size_t( pairTriangles( mesh, g->pairs.data() ) ) );
	}
	else
	{
		g->pairs.resize( 2 * size_t( mesh.triCount ) );
		for ( uint32_t t = 0; t < mesh.triCount; ++t )
			g->pairs[2 * t] = g->pairs[2 * t + 1] = t;
	}
	g->primCount = uint32_t( g->pairs.size() / 2 );
	g->leafType	 = TriangleType;
	g->primNodes.resize( g->primCount );
	g->primBoxes.resize( g->primCount );
	for ( uint32_t p = 0; p < g->primCount; ++p )
	{
		g->primNodes[p] = makePairNode( mesh, g->pairs[2 * p], g->pairs[2 * p + 1] );
		g->primBoxes[p] = pairBox( g->primNodes[p] );
	}
	finishBuild( *g );
	sortAndBuild( *g );
	if ( g->primCount > 1 )
		g->root = buildTopology( g->sortedKeys.data(), g->primCount, g->sortedBoxes.data(), g->leafType, g->scratch );
	wideAndPack( *g );
	return g;
}

void orc_geometry_destroy( void* h ) { delete static_cast<Geometry*>( h ); }

// Replace keys/order with externally supplied sorted keys + values (e.g. downloaded from the
// GPU build) and rebuild topology/collapse/pack: "topology bit-exact for identical Morton codes".
void orc_geometry_rebuild_from_sorted( void* h, const uint32_t* sortedKeys, const uint32_t* sortedVals )
{
	Geometry& g = *static_cast<Geometry*>( h );
	const uint32_t n = g.primCount;
	g.sortedKeys.assign( sortedKeys, sortedKeys + n );
	g.sortedVals.assign( sortedVals, sortedVals + n );
	for ( uint32_t k = 0; k < n; ++k )
		g.sortedBoxes[k] = g.primBoxes[g.sortedVals[k]];
	if ( n > 1 ) g.root = buildTopology( g.sortedKeys.data(), n, g.sortedBoxes.data(), g.leafType, g.scratch );
	wideAndPack( g );
}

uint32_t orc_geometry_count( void* h, int what )
{
	Geometry& g = *static_cast<Geometry*>( h );
	switch ( what )
	{
	case 0: return g.primCount;
	case 1: return uint32_t( g.wide.boxes.size() );
	case 2: return uint32_t( g.wide.leafRefs.size() );
	case 3: return g.root;
	}
	return 0;
}

// what: 0 pairs(u32 x2) 1 primBoxes(24B) 2 keys 3 sortedKeys 4 sortedVals 5 scratch(32B)
//       6 boxNodes(128B) 7 triangle leaves(64B) 8 sceneBox(24B) 9 primNodes input order (64B)
void orc_geometry_read( void* h, int what, void* out )
{
	Geometry& g = *static_cast<Geometry*>( h );
	auto	  cp = [&]( const void* p, size_t bytes ) {
		 if ( bytes ) memcpy( out, p, bytes );
	};
	switch ( what )
	{
	case 0: cp( g.pairs.data(), g.pairs.size() * 4 ); break;
	case 1: cp( g.primBoxes.data(), g.primBoxes.size() * sizeof( Box ) ); break;
	case 2: cp( g.keys.data(), g.keys.size() * 4 ); break;
	case 3: cp( g.sortedKeys.data(), g.sortedKeys.size() * 4 ); break;
	case 4: cp( g.sortedVals.data(), g.sortedVals.size() * 4 ); break;
	case 5: cp( g.scratch.data(), g.scratch.size() * sizeof( ScratchNode ) ); break;
	case 6: cp( g.wide.boxes.data(), g.wide.boxes.size() * sizeof( Box4Node ) ); break;
	case 7: cp( g.leaves.data(), g.leaves.size() * sizeof( TrianglePairNode ) ); break;
	case 8: cp( &g.sceneBox, sizeof( Box ) ); break;
	case 9: cp( g.primNodes.data(), g.primNodes.size() * sizeof( TrianglePairNode ) ); break;
	}
}

// Binary LBVH from sorted keys and sorted leaf boxes alone. scratchOut: (n-1) x 32 B. Returns root id.
uint32_t orc_lbvh_topology( const uint32_t* sortedKeys, const float* sortedBoxes, uint32_t n, uint32_t leafType, void* scratchOut )
{
	std::vector<ScratchNode> nodes;
	const uint32_t root = buildTopology( sortedKeys, n, reinterpret_cast<const Box*>( sortedBoxes ), leafType, nodes );
	if ( !nodes.empty() ) memcpy( scratchOut, nodes.data(), nodes.size() * sizeof( ScratchNode ) );
	return root;
}

// Trace through the oracle's own structure. mode 0 = closest (tieMode 0), 1 = closest (tieMode 1).
void orc_geometry_trace( void* h, const hiprtRay* rays, uint32_t n, hiprtHit* hits, int tieMode, uint32_t* visits )
{
	Geometry& g = *static_cast<Geometry*>( h );
	BvhView	  v{ g.wide.boxes.data(), g.leaves.data() };
	if ( !visits )
	{
		forEachRay( n, [&]( uint32_t i ) { hits[i] = canonical( traceClosest( v, rays[i], tieMode, nullptr, nullptr ), false ); } );
		return;
	}
	uint32_t bv = 0, lv = 0;
	for ( uint32_t i = 0; i < n; ++i )
		hits[i] = canonical( traceClosest( v, rays[i], tieMode, &bv, &lv ), false );
	visits[0] = bv, visits[1] = lv;
}

// Key-frame interpolation for the motion-blur tests. frames: hiprtFrameSRT rows (12 floats: axis-angle rotation, scale, translation,
// time, pad). matrixOut: 3x4 object -> world matrix of the interpolated pose (hiprt_device_impl.h:1980-2005);
// rayOut: the world ray carried into the instance (Transform::transformRay: origin and origin + direction through the inverse).
void orc_motion_pose_typed(
	const float* keyFrames, uint32_t count, int frameType, float time, float* matrixOut, const hiprtRay* rayIn, hiprtRay* rayOut );

void orc_motion_pose( const float* srtFrames, uint32_t count, float time, float* matrixOut, const hiprtRay* rayIn, hiprtRay* rayOut )
{
	orc_motion_pose_typed( srtFrames, count, 0, time, matrixOut, rayIn, rayOut );
}

// frameType 0: hiprtFrameSRT rows (12 floats), 1: hiprtFrameMatrix rows (16 floats: the matrix is decomposed into scale / shear /
// rotation first, Transform.h:61-127, and the key frames are interpolated in that form)
void orc_motion_pose_typed(
	const float* keyFrames, uint32_t count, int frameType, float time, float* matrixOut, const hiprtRay* rayIn, hiprtRay* rayOut )
{
	std::vector<Frame> frames( count );
	for ( uint32_t i = 0; i < count; ++i )
		frames[i] = frameType == 1 ? frameFromMatrix( keyFrames + 16 * size_t( i ) ) : frameFromSRT( keyFrames + 12 * size_t( i ) );
	const Frame f = frameAt( frames.data(), count, time );
	if ( matrixOut )
	{
		float Q[3][3];
		qToMatrix( f.rot, Q );
		for ( int i = 0; i < 3; ++i )
		{
			matrixOut[4 * i + 0] = Q[i][0] * f.scale.x;
			matrixOut[4 * i + 1] = Q[i][1] * f.scale.y + Q[i][0] * f.shear.x;
			matrixOut[4 * i + 2] = Q[i][2] * f.scale.z + Q[i][1] * f.shear.z + Q[i][0] * f.shear.y;
		}
		matrixOut[3] = f.trans.x, matrixOut[7] = f.trans.y, matrixOut[11] = f.trans.z;
	}
	if ( rayIn && rayOut )
	{
		*rayOut = *rayIn;
		if ( !f.identity() )
		{
			const V3 o = rayIn->origin, d = rayIn->direction;
			const V3 oo = frameInvTransform( f, o ), e = frameInvTransform( f, o + d );
			rayOut->origin = oo, rayOut->direction = e - oo;
		}
	}
}

uint64_t orc_fruitless_box_visits()
{
	const uint64_t v	 = g_fruitlessBoxVisits;
	g_fruitlessBoxVisits = 0;
	return v;
}

// Per-ray visit counts (box nodes, leaves) of the closest-hit traversal: input of the SIMT-efficiency model
// in tools/simt_model.py.
void orc_geometry_trace_visits( void* h, const hiprtRay* rays, uint32_t n, uint32_t* boxVisits, uint32_t* leafVisits )
{
	Geometry& g = *static_cast<Geometry*>( h );
	BvhView	  v{ g.wide.boxes.data(), g.leaves.data() };
	for ( uint32_t i = 0; i < n; ++i )
	{
		uint32_t bv = 0, lv = 0;
		traceClosest( v, rays[i], 1, &bv, &lv );
		boxVisits[i] = bv, leafVisits[i] = lv;
	}
}

// Trace through arbitrary node arrays in the library's HBM format (e.g. a GPU-built structure
// copied to the host): checks traversal independently of the builder.
void orc_trace_nodes( const void* boxNodes, const void* triNodes, const hiprtRay* rays, uint32_t n, hiprtHit* hits, int tieMode )
{
	BvhView v{ static_cast<const Box4Node*>( boxNodes ), static_cast<const TrianglePairNode*>( triNodes ) };
	for ( uint32_t i = 0; i < n; ++i )
		hits[i] = traceClosest( v, rays[i], tieMode, nullptr, nullptr );
}

// Any-hit iterator drained: for each ray the number of hits and, if primOut != null, up to
// maxPerRay prim ids in traversal order. Returns the total hit count.
uint64_t orc_geometry_trace_all(
	void* h, const hiprtRay* rays, uint32_t n, uint32_t* countOut, uint32_t* primOut, float* tOut, uint32_t maxPerRay )
{
	Geometry&			  g = *static_cast<Geometry*>( h );
	BvhView				  v{ g.wide.boxes.data(), g.leaves.data() };
	std::vector<hiprtHit> hits;
	uint64_t			  total = 0;
	for ( uint32_t i = 0; i < n; ++i )
	{
		hits.clear();
		traceAll( v, rays[i], hits );
		countOut[i] = uint32_t( hits.size() );
		total += hits.size();
		for ( uint32_t k = 0; k < hits.size() && k < maxPerRay; ++k )
		{
			if ( primOut ) primOut[size_t( i ) * maxPerRay + k] = hits[k].primID;
			if ( tOut ) tOut[size_t( i ) * maxPerRay + k] = hits[k].t;
		}
	}
	return total;
}

// Renumber an arbitrary 4-wide structure into the canonical order used above (breadth-first
// box nodes, leaves by (box, slot)) so that structures whose addresses depend on atomic
// arrival order can be compared bit-for-bit. Returns the number of box nodes reached, and
// the number of leaves through *leafCountOut. leafStride is the size of one leaf node.
uint32_t orc_canonicalize(
	const void* boxNodesIn, const void* leavesIn, uint32_t leafStride, void* boxNodesOut, void* leavesOut, uint32_t* leafCountOut )
{
	const Box4Node*		  in = static_cast<const Box4Node*>( boxNodesIn );
	Box4Node*			  out = static_cast<Box4Node*>( boxNodesOut );
	std::vector<uint32_t> order{ 0 }; // old address per new address
	std::vector<uint32_t> parentNew{ InvalidValue };
	uint32_t			  leafCount = 0;
	for ( size_t k = 0; k < order.size(); ++k )
	{
		Box4Node node	= in[order[k]];
		node.parentAddr = parentNew[k];
		node.updateCounter = 0;
		node.pad		= 0;
		for ( int s = 0; s < 4; ++s )
		{
			const uint32_t c = node.child[s];
			if ( c == InvalidValue ) continue;
			if ( isBoxIndex( c ) )
			{
				order.push_back( nodeAddr( c ) );
				parentNew.push_back( uint32_t( k ) );
				node.child[s] = makeNodeIndex( uint32_t( order.size() - 1 ), BoxType );
			}
			else
			{
				memcpy(
					static_cast<uint8_t*>( leavesOut ) + size_t( leafCount ) * leafStride,
					static_cast<const uint8_t*>( leavesIn ) + size_t( nodeAddr( c ) ) * leafStride,
					leafStride );
				node.child[s] = makeNodeIndex( leafCount, nodeType( c ) );
				++leafCount;
			}
		}
		out[k] = node;
	}
	*leafCountOut = leafCount;
	return uint32_t( order.size() );
}

// ---- scenes -------------------------------------------------------------------------------------
// One frame per instance (frame k belongs to instance k). frameType 0: hiprtFrameSRT (12 floats),
// 1: hiprtFrameMatrix (16 floats). geometryOfInstance[k] indexes `geometries`.
// objectIsScene (optional, one flag per handle in `geometries`): the handle is an oracle scene, instanced as hiprtInstanceTypeScene
void* orc_scene_create_multilevel(
	void* const*	geometries,
	uint32_t		geometryCount,
	const uint32_t* geometryOfInstance,
	const float*	frames,
	int				frameType,
	const uint32_t* masks,
	uint32_t		instanceCount,
	const uint32_t* objectIsScene );

void* orc_scene_create(
	void* const*	geometries,
	uint32_t		geometryCount,
	const uint32_t* geometryOfInstance,
	const float*	frames,
	int				frameType,
	const uint32_t* masks,
	uint32_t		instanceCount )
{
	return orc_scene_create_multilevel( geometries, geometryCount, geometryOfInstance, frames, frameType, masks, instanceCount, nullptr );
}

void* orc_scene_create_multilevel(
	void* const*	geometries,
	uint32_t		geometryCount,
	const uint32_t* geometryOfInstance,
	const float*	frames,
	int				frameType,
	const uint32_t* masks,
	uint32_t		instanceCount,
	const uint32_t* objectIsScene )
{
	(void)geometryCount;
	SceneObj* s = new SceneObj;
	s->count	= instanceCount;
	s->geoms.resize( instanceCount );
	s->scenes.assign( instanceCount, nullptr );
	s->frames.resize( instanceCount );
	s->masks.resize( instanceCount );
	s->primBoxes.resize( instanceCount );
	for ( uint32_t k = 0; k < instanceCount; ++k )
	{
		const bool nested = objectIsScene && objectIsScene[geometryOfInstance[k]];
		s->geoms[k]		  = nested ? nullptr : static_cast<Geometry*>( geometries[geometryOfInstance[k]] );
		s->scenes[k]	  = nested ? static_cast<SceneObj*>( geometries[geometryOfInstance[k]] ) : nullptr;
		s->frames[k] = frames ? ( frameType == 1 ? frameFromMatrix( frames + 16 * size_t( k ) ) : frameFromSRT( frames + 12 * size_t( k ) ) ) : Frame();
		s->masks[k]	 = masks ? masks[k] : FullRayMask;
		s->primBoxes[k] = instanceBox( s->geoms[k] ? s->geoms[k]->wide : s->scenes[k]->wide, s->frames[k] );
	}
	// same LBVH pipeline as geometries, leaves are instances
	s->sceneBox = emptyBox();
	for ( uint32_t k = 0; k < instanceCount; ++k )
		s->sceneBox = unite( s->sceneBox, s->primBoxes[k] );
	const V3 ext = s->sceneBox.hi - s->sceneBox.lo;
	s->keys.resize( instanceCount );
	for ( uint32_t k = 0; k < instanceCount; ++k )
	{
		const V3 c = ( s->primBoxes[k].hi + s->primBoxes[k].lo ) * 0.5f;
		const V3 d = c - s->sceneBox.lo;
		s->keys[k] = extendedMorton( mk( d.x / ext.x, d.y / ext.y, d.z / ext.z ), ext );
	}
	s->sortedVals.resize( instanceCount );
	std::iota( s->sortedVals.begin(), s->sortedVals.end(), 0u );
	std::stable_sort( s->sortedVals.begin(), s->sortedVals.end(), [&]( uint32_t a, uint32_t b ) { return s->keys[a] < s->keys[b]; } );
	s->sortedKeys.resize( instanceCount );
	s->sortedBoxes.resize( instanceCount );
	for ( uint32_t k = 0; k < instanceCount; ++k )
	{
		s->sortedKeys[k]  = s->keys[s->sortedVals[k]];
		s->sortedBoxes[k] = s->primBoxes[s->sortedVals[k]];
	}
	if ( instanceCount == 1 )
	{
		Box4Node node;
		memset( &node, 0, sizeof( node ) );
		node.parentAddr = InvalidValue;
		node.childCount = 1;
		for ( int k = 0; k < 4; ++k )
			node.child[k] = InvalidValue, node.box[k] = emptyBox();
		node.child[0] = makeNodeIndex( 0, InstanceType );
		node.box[0]	  = s->primBoxes[0];
		s->wide.boxes.assign( 1, node );
		s->wide.leafRefs.assign( 1, 0 );
	}
	else
	{
		s->root = buildTopology( s->sortedKeys.data(), instanceCount, s->sortedBoxes.data(), InstanceType, s->scratch );
		collapse( s->scratch, s->sortedBoxes.data(), instanceCount, s->root, InstanceType, s->wide );
	}
	for ( uint32_t ref : s->wide.leafRefs )
	{
		const uint32_t k = instanceCount == 1 ? 0 : s->sortedVals[ref];
		InstanceNode   n;
		memset( &n, 0, sizeof( n ) );
		const bool ident = s->frames[k].identity();
		float	   m[3][4] = { { 1, 0, 0, 0 }, { 0, 1, 0, 0 }, { 0, 0, 1, 0 } };
		if ( !ident ) frameInverseMatrix( s->frames[k], m );
		memcpy( n.invMatrix, m, sizeof( m ) );
		n.geometry	= nullptr;
		n.mask		= s->masks[k];
		n.primIndex = k;
		n.type		= s->scenes[k] ? 1 : 0;
		n.isStatic	= 1;
		n.identity	= ident ? 1 : 0;
		s->leaves.push_back( n );
		s->leafGeom.push_back( k );
	}
	return s;
}

void orc_scene_destroy( void* h ) { delete static_cast<SceneObj*>( h ); }

uint32_t orc_scene_count( void* h, int what )
{
	SceneObj& s = *static_cast<SceneObj*>( h );
	return what == 0 ? s.count : what == 1 ? uint32_t( s.wide.boxes.size() ) : uint32_t( s.leaves.size() );
}

// what: 0 primBoxes 1 keys 2 boxNodes 3 instance leaves (64 B, geometry pointer zeroed) 4 sortedKeys 5 sortedVals
void orc_scene_read( void* h, int what, void* out )
{
	SceneObj& s	 = *static_cast<SceneObj*>( h );
	auto	  cp = [&]( const void* p, size_t bytes ) {
		 if ( bytes ) memcpy( out, p, bytes );
	};
	switch ( what )
	{
	case 0: cp( s.primBoxes.data(), s.primBoxes.size() * sizeof( Box ) ); break;
	case 1: cp( s.keys.data(), s.keys.size() * 4 ); break;
	case 2: cp( s.wide.boxes.data(), s.wide.boxes.size() * sizeof( Box4Node ) ); break;
	case 3: cp( s.leaves.data(), s.leaves.size() * sizeof( InstanceNode ) ); break;
	case 4: cp( s.sortedKeys.data(), s.sortedKeys.size() * 4 ); break;
	case 5: cp( s.sortedVals.data(), s.sortedVals.size() * 4 ); break;
	}
}

// Replace the world->object matrices of the instance leaves by externally built ones (12 floats per instance,
// indexed by instance id). Lets a test separate "the transform was built the same way up to rounding"
// (checked on the matrices themselves) from "given identical instance nodes, traversal agrees to the strict bar".
void orc_scene_set_inv_matrices( void* h, const float* invMatrices )
{
	SceneObj& s = *static_cast<SceneObj*>( h );
	for ( InstanceNode& n : s.leaves )
		if ( !n.identity ) memcpy( n.invMatrix, invMatrices + 12 * size_t( n.primIndex ), 12 * sizeof( float ) );
}

void orc_scene_trace( void* h, const hiprtRay* rays, uint32_t n, hiprtHit* hits, uint32_t rayMask, int tieMode )
{
	SceneObj& s = *static_cast<SceneObj*>( h );
	forEachRay( n, [&]( uint32_t i ) { hits[i] = canonical( traceSceneClosest( s, rays[i], rayMask, tieMode ), true ); } );
}

// Custom primitives: an AABB list (hiprtAABBListPrimitive: min xyz at byte 0, max xyz at byte strideBytes / 2 of every record),
// built by the same LBVH pipeline with custom leaves.
void* orc_geometry_create_aabbs( const float* aabbs, uint32_t count, uint32_t strideBytes, uint32_t geomType )
{
	Geometry* g	 = new Geometry;
	g->primCount = count;
	g->leafType	 = CustomType;
	g->geomType	 = geomType;
	g->primBoxes.resize( count );
	for ( uint32_t p = 0; p < count; ++p )
	{
		const float* lo = reinterpret_cast<const float*>( reinterpret_cast<const uint8_t*>( aabbs ) + size_t( p ) * strideBytes );
		const float* hi = reinterpret_cast<const float*>( reinterpret_cast<const uint8_t*>( lo ) + strideBytes / 2 );
		g->primBoxes[p] = Box{ mk( lo[0], lo[1], lo[2] ), mk( hi[0], hi[1], hi[2] ) };
	}
	finishBuild( *g );
	sortAndBuild( *g );
	if ( g->primCount > 1 )
		g->root = buildTopology( g->sortedKeys.data(), g->primCount, g->sortedBoxes.data(), g->leafType, g->scratch );
	wideAndPack( *g );
	return g;
}

// General trace (single-threaded: `fn` may be a Python callback). isScene: `h` is a scene, else a geometry.
// maxPerRay == 0: closest hit, one hiprtHit per ray in hitsOut. maxPerRay > 0: every hit along the ray in visiting order,
// countOut[i] of them (all are counted, the first maxPerRay are stored at hitsOut[i * maxPerRay ...]).
void orc_trace_general(
	void* h, int isScene, const hiprtRay* rays, uint32_t n, uint32_t rayMask, int tieMode, CustomIntersectFn fn, void* user, uint32_t maxPerRay,
	uint32_t* countOut, hiprtHit* hitsOut )
{
	std::vector<hiprtHit> all;
	for ( uint32_t i = 0; i < n; ++i )
	{
		Walk w;
		w.closest = maxPerRay == 0, w.tieMode = tieMode, w.rayMask = rayMask, w.fn = fn, w.user = user, w.all = &all;
		for ( uint32_t& id : w.ids )
			id = InvalidValue;
		all.clear();
		hiprtRay ray = rays[i];
		if ( isScene )
			walkScene( *static_cast<SceneObj*>( h ), ray, 0, w );
		else
			walkGeometry( *static_cast<Geometry*>( h ), ray, 0, w );
		if ( maxPerRay == 0 )
		{
			hiprtHit b = w.best;
			if ( !b.hasHit() ) b.uv = { 0.0f, 0.0f }, b.normal = { 0.0f, 0.0f, 0.0f };
			for ( uint32_t l = 0; l < hiprtMaxInstanceLevels; ++l )
				if ( !b.hasHit() || !isScene ) b.instanceIDs[l] = InvalidValue;
			hitsOut[i] = b;
			continue;
		}
		countOut[i] = uint32_t( all.size() );
		for ( uint32_t k = 0; k < all.size() && k < maxPerRay; ++k )
			hitsOut[size_t( i ) * maxPerRay + k] = all[k];
	}
}

} // extern "C"
