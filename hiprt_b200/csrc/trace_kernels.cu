// Library-provided trace kernels (rays in -> hits out), built from the public device API in
// include/hiprt/hiprt_device.h exactly as a user kernel would be (reference counterpart:
// test/kernels/HiprtTestKernel.h:102-116 TraceKernel).
//
// Stack: the top of each thread's stack lives in shared memory (lane-interleaved, so a warp's
// accesses hit 32 distinct banks), deeper entries overflow to a per-thread local array that
// is rarely touched. One thread per ray; blocks of 128 threads.
#define HIPRT_B200_DEFAULT_FUNC_TABLE
#include <hiprt/hiprt_device.h>

#include "builders.h"
#include "qnode.h"
#include "trace.h"

#include <algorithm>
#include <cstdlib>

namespace hiprt
{
constexpr uint32_t TraceBlock		= 128; // scene kernels
constexpr uint32_t SharedStackDepth = 12;
constexpr uint32_t FullMask			= 0xFFFFFFFFu;

// Top `SharedDepth` entries in shared memory (entry e of a thread at base[e * 32], one bank per lane), the rest of the
// 64 entries the reference's private stack offers in a per-thread local array.
template <uint32_t SharedDepth>
class HybridStackT
{
  public:
	__device__ HybridStackT( uint32_t* sharedBase ) : m_shared( sharedBase + ( threadIdx.x & 31u ) + ( threadIdx.x >> 5 ) * 32u * SharedDepth ) { reset(); }
	__device__ void push( uint32_t v )
	{
		if ( m_top < SharedDepth )
			m_shared[m_top * 32u] = v;
		else
			m_local[m_top - SharedDepth] = v;
		++m_top;
	}
	__device__ uint32_t pop()
	{
		--m_top;
		return m_top < SharedDepth ? m_shared[m_top * 32u] : m_local[m_top - SharedDepth];
	}
	__device__ bool		empty() const { return m_top == 0; }
	__device__ uint32_t vacancy() const { return SharedDepth + LocalDepth - m_top; }
	__device__ void		reset() { m_top = 0; }

  private:
	static constexpr uint32_t LocalDepth = 64 - SharedDepth;
	uint32_t*				  m_shared;
	uint32_t				  m_local[LocalDepth];
	uint32_t				  m_top;
};
using HybridStack = HybridStackT<SharedStackDepth>;

__device__ __forceinline__ hiprtRay loadRay( const hiprtRay* rays, uint32_t index )
{
	const float4* rp = reinterpret_cast<const float4*>( rays + index );
	const float4  r0 = __ldg( rp ), r1 = __ldg( rp + 1 );
	hiprtRay	  ray;
	ray.origin = make_float3( r0.x, r0.y, r0.z ), ray.minT = r0.w;
	ray.direction = make_float3( r1.x, r1.y, r1.z ), ray.maxT = r1.w;
	return ray;
}
// A default-constructed hiprtHit initialises instanceIDs[0] only (the union's first member, as in the reference's hiprt_types.h:237-258);
// the library kernels write all four words of the record, so the deeper levels are set to "none" explicitly.
__device__ __forceinline__ hiprtHit emptyHit()
{
	hiprtHit hit;
	hit.instanceIDs[1] = hit.instanceIDs[2] = hit.instanceIDs[3] = InvalidValue;
	return hit;
}
// (hits of the single-level traversal objects: only instanceIDs[0] is theirs)
__device__ __forceinline__ void storeHit( hiprtHit* hits, uint32_t index, const hiprtHit& hit );
__device__ __forceinline__ void storeSingleLevelHit( hiprtHit* hits, uint32_t index, hiprtHit hit )
{
	hit.instanceIDs[1] = hit.instanceIDs[2] = hit.instanceIDs[3] = InvalidValue;
	storeHit( hits, index, hit );
}
__device__ __forceinline__ void storeHit( hiprtHit* hits, uint32_t index, const hiprtHit& hit )
{
	float4* hp = reinterpret_cast<float4*>( hits + index );
	hp[0] = make_float4( __uint_as_float( hit.instanceIDs[0] ), __uint_as_float( hit.instanceIDs[1] ), __uint_as_float( hit.instanceIDs[2] ), __uint_as_float( hit.instanceIDs[3] ) );
	hp[1] = make_float4( __uint_as_float( hit.primID ), 0.0f, hit.uv.x, hit.uv.y );
	hp[2] = make_float4( hit.normal.x, hit.normal.y, hit.normal.z, hit.t );
}
// Output record formats of the library kernels (hiprtB200HitFormat, include/hiprt/hiprt_b200.h): the full 48-byte hiprtHit, a
// 16-byte {primID, t, u, v} record, or the primitive id alone (occlusion queries: hiprtInvalidValue = not occluded). The
// traversal is the same; the smaller records cut the write traffic of a launch (48 -> 16 / 4 bytes per ray).
template <uint32_t Format>
__device__ __forceinline__ void storeHitAs( void* out, uint32_t index, const hiprtHit& hit )
{
	if constexpr ( Format == hiprtB200HitFormatFull )
		storeHit( static_cast<hiprtHit*>( out ), index, hit );
	else if constexpr ( Format == hiprtB200HitFormatCompact )
		static_cast<float4*>( out )[index] = make_float4( __uint_as_float( hit.primID ), hit.t, hit.uv.x, hit.uv.y );
	else
		static_cast<uint32_t*>( out )[index] = hit.primID;
}

// One thread per ray. Persistent = the grid only fills the machine and every warp keeps fetching batches of 32
// consecutive rays from a global counter until the buffer is exhausted (no block-granular tail: a block slot is
// never held by one long ray while its other warps sit idle).
template <hiprtTraversalType Type, uint32_t Block, uint32_t SharedDepth, bool Persistent>
__global__ void __launch_bounds__( Block ) traceGeometryKernel(
	hiprtGeometry geom, const hiprtRay* __restrict__ rays, uint32_t numRays, hiprtHit* __restrict__ hits, uint32_t* workCounter )
{
	__shared__ uint32_t sStack[Block * SharedDepth];
	using Stack		= HybridStackT<SharedDepth>;
	const bool mesh = reinterpret_cast<const GeomHeader*>( geom )->geomType & 1u;
	if constexpr ( Persistent )
	{
		const uint32_t lane = threadIdx.x & 31u;
		while ( true )
		{
			uint32_t base = 0;
			if ( lane == 0 ) base = atomicAdd( workCounter, 32u );
			base = __shfl_sync( 0xFFFFFFFFu, base, 0 );
			if ( base >= numRays ) return;
			const uint32_t index = base + lane;
			if ( index < numRays )
			{
				Stack		   stack( sStack );
				const hiprtRay ray = loadRay( rays, index );
				hiprtHit	   hit;
				if ( mesh )
				{
					hiprtGeomTraversalT<Stack, hiprtTriangleNode, Type> tr( geom, ray, stack );
					hit = tr.getNextHit();
				}
				else
				{
					hiprtGeomTraversalT<Stack, hiprtCustomNode, Type> tr( geom, ray, stack );
					hit = tr.getNextHit();
				}
				storeSingleLevelHit( hits, index, hit );
			}
			__syncwarp();
		}
	}
	else
	{
		const uint32_t index = blockIdx.x * Block + threadIdx.x;
		if ( index >= numRays ) return;
		Stack		   stack( sStack );
		const hiprtRay ray = loadRay( rays, index );
		hiprtHit	   hit;
		if ( mesh )
		{
			hiprtGeomTraversalT<Stack, hiprtTriangleNode, Type> tr( geom, ray, stack );
			hit = tr.getNextHit();
		}
		else
		{
			hiprtGeomTraversalT<Stack, hiprtCustomNode, Type> tr( geom, ray, stack );
			hit = tr.getNextHit();
		}
		storeSingleLevelHit( hits, index, hit );
	}
}

// ------------------------------------------------------------------------------------------------
// Fast path for triangle geometries without custom functions (the library's own default table rejects
// nothing): same slab and Moller-Trumbore arithmetic as the generic traversal objects (the helpers
// of hiprt_device.h), restructured around what ncu shows to be the limiter of one-thread-per-ray
// traversal on sm_100a: the half-rate ALU pipe (min/max, select, compare), not the FMA pipe and not memory
// (profiles/r01_ncu_trace_fast_*.csv).
//  * while-while loop with warp votes (box phase / leaf phase), so lanes of a warp execute the same phase;
//    the box phase ends when no lane has a box node, or fewer than `BoxQuorum` have one while some lane waits
//    with a leaf;
//  * when every ray of a warp has the same direction octant and finite non-zero 1/d, the near and far plane of
//    each slab are known at compile time of the octant's code copy: (near - o) * invD and (far - o) * invD are
//    the very products the reference takes min/max of (Aabb.h:92-101; rounding is monotonic, lo <= hi), so the
//    six per-child min/max disappear with bit-identical results. Mixed warps and rays with a zero direction
//    component (where the reference's min/max see NaN/inf) run the general code;
//  * stack top in a register, entries in lane-interleaved shared memory (local array only beyond SharedDepth);
//  * children ordered by a 5-exchange network on (entry distance, index) with misses at -inf, so the hit
//    children are the LAST n of the four: all four are stored unconditionally far-to-near above the stack top,
//    the top moves by n and the nearest child is simply popped -- no per-child predicates;
//  * closest hit keeps only (t, leaf, triangle) in the loop; uv mapping and the normal are produced once
//    per ray at the end by re-running the reference-order test on the winning triangle.
// The winner is the minimum of (t, primID), as in hiprt::acceptCloser.
constexpr uint32_t GeneralOctant = 8;
#ifndef HIPRT_B200_SORT_ANY_HIT
#define HIPRT_B200_SORT_ANY_HIT 0 // 1 = any-hit rays also visit children nearest first (the round-1 kernel)
#endif

// 32-byte read-only load (LDG.E.256, sm_100+): a 128-byte box node is 4 of these instead of 7 16-byte loads, a
// 64-byte triangle pair 2 instead of 4 -- the traversal is bound by L1TEX wavefronts (one per distinct line and
// instruction when the lanes of a warp are in different nodes), so fewer, wider loads are the lever.
struct F8
{
	float v[8];
};
// Node loads carry an L2 evict-last hint, ray loads and hit stores an evict-first one (and bypass L1): a launch streams
// 32 B of ray in and up to 48 B of hit out per ray through the same L2 that has to keep the node working set (C3: 0.36 GB of
// rays + 0.54 GB of hits per frame against 126 MB of L2), and the top stall of the kernel is the scoreboard wait on node loads.
template <bool Hints>
__device__ __forceinline__ F8 ldg8( const void* p, uint32_t i )
{
	F8 r;
	if constexpr ( Hints )
		asm( "ld.global.nc.L2::evict_last.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
			 : "=f"( r.v[0] ), "=f"( r.v[1] ), "=f"( r.v[2] ), "=f"( r.v[3] ), "=f"( r.v[4] ), "=f"( r.v[5] ), "=f"( r.v[6] ), "=f"( r.v[7] )
			 : "l"( reinterpret_cast<const char*>( p ) + 32u * i ) );
	else
		asm( "ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
			 : "=f"( r.v[0] ), "=f"( r.v[1] ), "=f"( r.v[2] ), "=f"( r.v[3] ), "=f"( r.v[4] ), "=f"( r.v[5] ), "=f"( r.v[6] ), "=f"( r.v[7] )
			 : "l"( reinterpret_cast<const char*>( p ) + 32u * i ) );
	return r;
}
// A hiprtRay is 32 bytes and 32-byte aligned inside the ray buffer: ONE streaming 256-bit load.
template <bool Hints>
__device__ __forceinline__ hiprtRay loadRayT( const hiprtRay* rays, uint32_t index )
{
	if constexpr ( !Hints ) return loadRay( rays, index );
	F8 r;
	asm( "ld.global.nc.L1::no_allocate.L2::evict_first.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
		 : "=f"( r.v[0] ), "=f"( r.v[1] ), "=f"( r.v[2] ), "=f"( r.v[3] ), "=f"( r.v[4] ), "=f"( r.v[5] ), "=f"( r.v[6] ), "=f"( r.v[7] )
		 : "l"( rays + index ) );
	hiprtRay ray;
	ray.origin = make_float3( r.v[0], r.v[1], r.v[2] ), ray.minT = r.v[3];
	ray.direction = make_float3( r.v[4], r.v[5], r.v[6] ), ray.maxT = r.v[7];
	return ray;
}
__device__ __forceinline__ void stStream16( void* p, float4 v )
{
	uint64_t policy;
	asm( "createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"( policy ) );
	asm volatile( "st.global.L1::no_allocate.L2::cache_hint.v4.f32 [%0], {%1,%2,%3,%4}, %5;" ::"l"( p ), "f"( v.x ), "f"( v.y ), "f"( v.z ), "f"( v.w ), "l"( policy ) : "memory" );
}
template <uint32_t Format, bool Hints>
__device__ __forceinline__ void storeHitT( void* out, uint32_t index, const hiprtHit& hit )
{
	if constexpr ( !Hints )
		storeHitAs<Format>( out, index, hit );
	else if constexpr ( Format == hiprtB200HitFormatFull )
	{
		float4* hp = reinterpret_cast<float4*>( static_cast<hiprtHit*>( out ) + index );
		stStream16( hp, make_float4( __uint_as_float( hit.instanceIDs[0] ), __uint_as_float( hit.instanceIDs[1] ), __uint_as_float( hit.instanceIDs[2] ), __uint_as_float( hit.instanceIDs[3] ) ) );
		stStream16( hp + 1, make_float4( __uint_as_float( hit.primID ), 0.0f, hit.uv.x, hit.uv.y ) );
		stStream16( hp + 2, make_float4( hit.normal.x, hit.normal.y, hit.normal.z, hit.t ) );
	}
	else if constexpr ( Format == hiprtB200HitFormatCompact )
		stStream16( static_cast<float4*>( out ) + index, make_float4( __uint_as_float( hit.primID ), hit.t, hit.uv.x, hit.uv.y ) );
	else
		__stcs( static_cast<uint32_t*>( out ) + index, hit.primID );
}

// Slab test of child `s` of a box node. Octant bit k set = direction component k negative (near plane is hi).
template <uint32_t Octant>
__device__ __forceinline__ bool slabTest( const float* b, const float3 o, const float3 invD, float maxT, float& t0Out )
{
	if constexpr ( Octant == GeneralOctant )
	{
		const float fx = __fmul_rn( __fsub_rn( b[3], o.x ), invD.x ), nx = __fmul_rn( __fsub_rn( b[0], o.x ), invD.x );
		const float fy = __fmul_rn( __fsub_rn( b[4], o.y ), invD.y ), ny = __fmul_rn( __fsub_rn( b[1], o.y ), invD.y );
		const float fz = __fmul_rn( __fsub_rn( b[5], o.z ), invD.z ), nz = __fmul_rn( __fsub_rn( b[2], o.z ), invD.z );
		const float t1 = fminf( fminf( fminf( fmaxf( fx, nx ), fmaxf( fy, ny ) ), fmaxf( fz, nz ) ), maxT );
		const float t0 = fmaxf( fmaxf( fmaxf( fminf( fx, nx ), fminf( fy, ny ) ), fminf( fz, nz ) ), 0.0f );
		t0Out		   = t0;
		return t0 <= t1;
	}
	else
	{
		constexpr int nxI = ( Octant & 1u ) ? 3 : 0, nyI = ( Octant & 2u ) ? 4 : 1, nzI = ( Octant & 4u ) ? 5 : 2;
		constexpr int fxI = 3 - nxI, fyI = 5 - nyI, fzI = 7 - nzI;
		const float	  nx = __fmul_rn( __fsub_rn( b[nxI], o.x ), invD.x ), fx = __fmul_rn( __fsub_rn( b[fxI], o.x ), invD.x );
		const float	  ny = __fmul_rn( __fsub_rn( b[nyI], o.y ), invD.y ), fy = __fmul_rn( __fsub_rn( b[fyI], o.y ), invD.y );
		const float	  nz = __fmul_rn( __fsub_rn( b[nzI], o.z ), invD.z ), fz = __fmul_rn( __fsub_rn( b[fzI], o.z ), invD.z );
		const float	  t1 = fminf( fminf( fminf( fx, fy ), fz ), maxT );
		const float	  t0 = fmaxf( fmaxf( fmaxf( nx, ny ), nz ), 0.0f );
		t0Out			 = t0;
		return t0 <= t1;
	}
}

// Exchange network shared by the exact and the quantised visit: leaves the hit children sorted, nearest last.
__device__ __forceinline__ void sortChildren( float d[4], uint32_t c[4] )
{
#define HIPRT_B200_CSWAP( A, B )                                     \
	{                                                                \
		const bool	   sw = d[B] < d[A];                             \
		const float	   lo = sw ? d[B] : d[A], hi = sw ? d[A] : d[B]; \
		const uint32_t cl = sw ? c[B] : c[A], ch = sw ? c[A] : c[B]; \
		d[A] = lo, d[B] = hi, c[A] = cl, c[B] = ch;                  \
	}
	HIPRT_B200_CSWAP( 0, 2 )
	HIPRT_B200_CSWAP( 1, 3 )
	HIPRT_B200_CSWAP( 0, 1 )
	HIPRT_B200_CSWAP( 2, 3 )
	HIPRT_B200_CSWAP( 1, 2 )
#undef HIPRT_B200_CSWAP
}

// One box-node visit: tests the four children, leaves the hit ones sorted (nearest last) in c[4-n..3].
// Sort == false (any-hit queries: the answer is occluded / not occluded, whichever order finds it): no ordering network and no
// entry distances -- returns the 4-bit mask of hit children and leaves c[] as loaded.
template <uint32_t Octant, bool Hints = false, bool Sort = true>
__device__ __forceinline__ uint32_t visitBox4( const Box4Node* __restrict__ boxNodes, uint32_t node, const float3 o, const float3 invD, float maxT, uint32_t c[4] )
{
	const void* p  = boxNodes + ( node >> 4 );
	const F8	w0 = ldg8<Hints>( p, 0 ), w1 = ldg8<Hints>( p, 1 ), w2 = ldg8<Hints>( p, 2 ), w3 = ldg8<Hints>( p, 3 );
	const float b[24] = { w0.v[4], w0.v[5], w0.v[6], w0.v[7], w1.v[0], w1.v[1], w1.v[2], w1.v[3], w1.v[4], w1.v[5], w1.v[6], w1.v[7],
						  w2.v[0], w2.v[1], w2.v[2], w2.v[3], w2.v[4], w2.v[5], w2.v[6], w2.v[7], w3.v[0], w3.v[1], w3.v[2], w3.v[3] };
	c[0] = __float_as_uint( w0.v[0] ), c[1] = __float_as_uint( w0.v[1] ), c[2] = __float_as_uint( w0.v[2] ), c[3] = __float_as_uint( w0.v[3] );
	if constexpr ( !Sort )
	{
		uint32_t mask = 0;
#pragma unroll
		for ( int s = 0; s < 4; ++s )
		{
			float t0;
			if ( slabTest<Octant>( b + 6 * s, o, invD, maxT, t0 ) && c[s] != InvalidValue ) mask |= 1u << s;
		}
		return mask;
	}
	float	 d[4];
	uint32_t n = 0;
#pragma unroll
	for ( int s = 0; s < 4; ++s )
	{
		float	   t0;
		const bool hit = slabTest<Octant>( b + 6 * s, o, invD, maxT, t0 ) && c[s] != InvalidValue; // (unused slots need not hold the empty box)
		d[s]		   = hit ? t0 : __int_as_float( 0xff800000 );
		c[s]		   = hit ? c[s] : InvalidValue;
		n += hit ? 1u : 0u;
	}
	sortChildren( d, c );
	return n;
}

// One box-node visit through the 64-byte quantised record (qnode.h): two 32-byte loads instead of four. The plane distance of
// plane index q on axis a is ((origin + q * cell) - o) * invD = val * A + B with val = 1 + q * 2^-15 (a float assembled from the
// byte by ONE byte-permute), A = cell * 2^15 * invD (exact: cell is a power of two) and B = (origin - o) * invD - A, one FFMA
// per plane, rounded down for entry and up for exit planes. Entry / exit distances are then widened by 2^-20 relative: together
// with the (at least) half grid cell of margin the builder leaves around every child box this makes the test conservative with respect to the
// exact float32 slab test (see qnode.h for the argument). Same contract as visitBox4.
template <bool Hints, bool Sort>
__device__ __forceinline__ uint32_t visitBox4Q( const QBox4Node* __restrict__ qNodes, uint32_t node, const float3 o, const float3 invD, float maxT, uint32_t c[4] )
{
	const void* p  = qNodes + ( node >> 4 );
	const F8	w0 = ldg8<Hints>( p, 0 ), w1 = ldg8<Hints>( p, 1 );
	const float ax = __fmul_rn( w0.v[3], invD.x ), ay = __fmul_rn( w0.v[4], invD.y ), az = __fmul_rn( w0.v[5], invD.z );
	const float bx = __fsub_rn( __fmul_rn( __fsub_rn( w0.v[0], o.x ), invD.x ), ax );
	const float by = __fsub_rn( __fmul_rn( __fsub_rn( w0.v[1], o.y ), invD.y ), ay );
	const float bz = __fsub_rn( __fmul_rn( __fsub_rn( w0.v[2], o.z ), invD.z ), az );
	c[0] = __float_as_uint( w0.v[6] ), c[1] = __float_as_uint( w0.v[7] ), c[2] = __float_as_uint( w1.v[0] ), c[3] = __float_as_uint( w1.v[1] );
	const uint32_t lox = __float_as_uint( w1.v[2] ), loy = __float_as_uint( w1.v[3] ), loz = __float_as_uint( w1.v[4] );
	const uint32_t hix = __float_as_uint( w1.v[5] ), hiy = __float_as_uint( w1.v[6] ), hiz = __float_as_uint( w1.v[7] );
	// entry planes: lo where the ray travels in +axis, hi otherwise
	const bool	   nx = invD.x < 0.0f, ny = invD.y < 0.0f, nz = invD.z < 0.0f;
	const uint32_t enx = nx ? hix : lox, eny = ny ? hiy : loy, enz = nz ? hiz : loz;
	const uint32_t exx = nx ? lox : hix, exy = ny ? loy : hiy, exz = nz ? loz : hiz;
	constexpr float Shrink = 0.99999904632568359375f, Grow = 1.00000095367431640625f, Tiny = 1.0e-30f; // 1 -+ 2^-20
	float			d[4];
	uint32_t		n = 0;
#pragma unroll
	for ( int s = 0; s < 4; ++s )
	{
		const uint32_t sel = 0x7604u | ( uint32_t( s ) << 4 ); // result bytes {0x00, byte s, 0x80, 0x3F}: 1 + q * 2^-15
		const float	   t0x = __fmaf_rd( __uint_as_float( __byte_perm( enx, 0x3F800000u, sel ) ), ax, bx );
		const float	   t0y = __fmaf_rd( __uint_as_float( __byte_perm( eny, 0x3F800000u, sel ) ), ay, by );
		const float	   t0z = __fmaf_rd( __uint_as_float( __byte_perm( enz, 0x3F800000u, sel ) ), az, bz );
		const float	   t1x = __fmaf_ru( __uint_as_float( __byte_perm( exx, 0x3F800000u, sel ) ), ax, bx );
		const float	   t1y = __fmaf_ru( __uint_as_float( __byte_perm( exy, 0x3F800000u, sel ) ), ay, by );
		const float	   t1z = __fmaf_ru( __uint_as_float( __byte_perm( exz, 0x3F800000u, sel ) ), az, bz );
		const float	   t0  = __fmaf_rd( fmaxf( fmaxf( fmaxf( t0x, t0y ), t0z ), 0.0f ), Shrink, -Tiny );
		const float	   t1  = __fmaf_ru( fminf( fminf( fminf( t1x, t1y ), t1z ), maxT ), Grow, Tiny );
		const bool	   hit = t0 <= t1; // unused slots hold an inverted box (entry plane 255 cells behind the exit plane): never hit
		if constexpr ( Sort )
		{
			d[s] = hit ? t0 : __int_as_float( 0xff800000 );
			c[s] = hit ? c[s] : InvalidValue;
			n += hit ? 1u : 0u;
		}
		else
			n |= hit ? ( 1u << s ) : 0u;
	}
	if constexpr ( Sort ) sortChildren( d, c );
	return n;
}

// The quantised walk is valid for rays whose exact slab arithmetic is free of NaN / infinity (qnode.h): every direction
// component finite, non-zero and of sane magnitude, origin finite and within the coordinate range the Q grid covers.
__device__ __forceinline__ bool regularForQ( const hiprtRay& ray )
{
	const float dx = fabsf( ray.direction.x ), dy = fabsf( ray.direction.y ), dz = fabsf( ray.direction.z );
	constexpr float Lo = 8.6736173798840355e-19f, Hi = 1152921504606846976.0f; // 2^-60, 2^60
	return dx >= Lo && dx <= Hi && dy >= Lo && dy <= Hi && dz >= Lo && dz <= Hi && fabsf( ray.origin.x ) <= QMaxCoordinate &&
		   fabsf( ray.origin.y ) <= QMaxCoordinate && fabsf( ray.origin.z ) <= QMaxCoordinate; // (comparisons are false for NaN)
}

// One triangle-pair leaf against the ray: shortens ray.maxT and records the winner as the traversal loops below expect.
// CheckBox (quantised walk only): the leaf's own box -- the box of its vertices, exactly what the builder stored in the parent
// node -- first goes through the EXACT slab test; a leaf the exact traversal would not have reached is skipped.
// The box test runs against the ray's ORIGINAL extent `boxMaxT`, not the current one: the exact walk tested this box when it
// visited the parent, i.e. against an extent that was at least the current one -- an exact tie (a duplicated triangle, or a
// flat box whose slab distance lies an ulp beyond the triangle's t) must not be culled because its twin was intersected first.
template <hiprtTraversalType Type, bool Hints, bool CheckBox>
__device__ __forceinline__ bool visitPairLeaf(
	const TrianglePairNode* __restrict__ leaves, uint32_t addr, hiprtRay& ray, const float3 invD, uint32_t& bestRef, uint32_t& bestPrim, float boxMaxT = 0.0f )
{
	const void*	   p  = leaves + addr;
	const F8	   l0 = ldg8<Hints>( p, 0 ), l1 = ldg8<Hints>( p, 1 );
	const uint32_t prim0 = __float_as_uint( l1.v[5] ), prim1 = __float_as_uint( l1.v[6] );
	const float3   pv0 = make_float3( l0.v[0], l0.v[1], l0.v[2] ), pv1 = make_float3( l0.v[3], l0.v[4], l0.v[5] );
	const float3   pv2 = make_float3( l0.v[6], l0.v[7], l1.v[0] ), pv3 = make_float3( l1.v[1], l1.v[2], l1.v[3] );
	if constexpr ( CheckBox )
	{
		// (a lone triangle repeats v2 in v3: the box of four vertices is the box of three)
		const float b[6] = { fminf( fminf( pv0.x, pv1.x ), fminf( pv2.x, pv3.x ) ), fminf( fminf( pv0.y, pv1.y ), fminf( pv2.y, pv3.y ) ),
							 fminf( fminf( pv0.z, pv1.z ), fminf( pv2.z, pv3.z ) ), fmaxf( fmaxf( pv0.x, pv1.x ), fmaxf( pv2.x, pv3.x ) ),
							 fmaxf( fmaxf( pv0.y, pv1.y ), fmaxf( pv2.y, pv3.y ) ), fmaxf( fmaxf( pv0.z, pv1.z ), fmaxf( pv2.z, pv3.z ) ) };
		float		t0;
		if ( !slabTest<GeneralOctant>( b, ray.origin, invD, boxMaxT, t0 ) ) return false;
	}
	bool		found = false;
	HitTriangle h;
	if ( intersectTriangle( pv0, pv1, pv2, ray, h ) && ( h.t < ray.maxT || prim0 < bestPrim ) )
	{
		ray.maxT = h.t, bestRef = addr << 1, bestPrim = prim0;
		found = true;
	}
	if ( prim0 != prim1 && !( Type == hiprtTraversalTerminateAtAnyHit && found ) )
	{
		if ( intersectTriangle( pv1, pv3, pv2, ray, h ) && ( h.t < ray.maxT || prim1 < bestPrim ) )
		{
			ray.maxT = h.t, bestRef = ( addr << 1 ) | 1u, bestPrim = prim1;
			found = true;
		}
	}
	return found;
}

// Exact traversal of ONE lane's ray, callable from divergent code (no warp votes): what a lane of the quantised kernels runs
// for a ray regularForQ() rejects. Stack = the lane's shared-memory column, then its local array. Returns false on overflow.
template <hiprtTraversalType Type, uint32_t SharedDepth, bool Hints>
__device__ __forceinline__ bool traceLaneExact(
	const Box4Node* __restrict__ boxNodes, const TrianglePairNode* __restrict__ leaves, hiprtRay& ray, uint32_t* const stackBase, uint32_t* localStack,
	uint32_t& bestRef, uint32_t& bestPrim )
{
	constexpr uint32_t TotalDepth = 64;
	const float3	   invD		  = rcp3( ray.direction );
	uint32_t		   top = 0, node = RootIndex;
	while ( node != InvalidValue )
	{
		if ( isBoxIndex( node ) )
		{
			uint32_t c[4];
			visitBox4<GeneralOctant, Hints>( boxNodes, node, ray.origin, invD, ray.maxT, c );
			if ( top + 3 > TotalDepth ) return false;
#pragma unroll
			for ( int k = 3; k >= 0; --k )
				if ( c[k] != InvalidValue )
				{
					if ( top < SharedDepth )
						stackBase[top * 32u] = c[k];
					else
						localStack[top - SharedDepth] = c[k];
					++top;
				}
		}
		else if ( visitPairLeaf<Type, Hints, false>( leaves, node >> 3, ray, invD, bestRef, bestPrim ) && Type == hiprtTraversalTerminateAtAnyHit )
			return true;
		if ( top == 0 ) break;
		--top;
		node = top < SharedDepth ? stackBase[top * 32u] : localStack[top - SharedDepth];
	}
	return true;
}

// Traversal of one ray per lane; every lane of the warp must call it (warp votes inside), `live` = has a ray.
// Refill == 0: traces ray `index` of the lane and returns when the whole warp is done.
// Refill  > 0: lane-granular persistent threads -- whenever at least `Refill` lanes are idle (their ray is finished) those
// lanes write their hits and fetch the next rays from `workCounter`, the others keep their state; returns when the ray
// buffer is exhausted and every lane is done. Pays off for incoherent rays only (see the kernel below).
// Q: box nodes are visited through the quantised 64-byte records `qNodes` (qnode.h); a lane whose ray regularForQ() rejects
// runs the exact traversal of that ray on its own (divergent, rare) and rejoins the warp as an idle lane.
template <hiprtTraversalType Type, uint32_t SharedDepth, uint32_t BoxQuorum, uint32_t Refill = 0, uint32_t Format = hiprtB200HitFormatFull, bool Hints = false, bool Stats = false, bool Q = false>
__device__ __forceinline__ void traceTrianglesWarp(
	const Box4Node* __restrict__ boxNodes,
	const TrianglePairNode* __restrict__ leaves,
	const hiprtRay* __restrict__ rays,
	void* __restrict__ hits,
	uint32_t		index,
	bool			live,
	uint32_t* const stackBase,
	uint32_t*		workCounter = nullptr,
	uint32_t		numRays		= 0,
	unsigned long long* stats	= nullptr, // Stats: {box-node visits, leaf visits} of the launch are added here
	const QBox4Node* __restrict__ qNodes = nullptr,
	bool				qInvalid		 = false ) // the Q region is marked unusable: every ray takes the exact path
{
	constexpr uint32_t TotalDepth = 64; // the reference's private stack size
	constexpr bool	   Sort		  = Type != hiprtTraversalTerminateAtAnyHit || HIPRT_B200_SORT_ANY_HIT;
	uint32_t		   localStack[TotalDepth - SharedDepth];
	uint32_t		   top = 0;
	uint32_t		   boxVisits = 0, leafVisits = 0;

	hiprtRay ray	  = loadRayT<Hints>( rays, live ? index : 0u );
	float3	 invD	  = rcp3( ray.direction );
	float	 maxT0	  = ray.maxT;
	uint32_t bestRef  = InvalidValue; // (leaf address << 1) | triangle of the pair
	uint32_t bestPrim = InvalidValue;
	bool	 overflow = false;
	uint32_t node	  = live ? RootIndex : InvalidValue;

	// direction octant of the warp, or GeneralOctant when its rays disagree / have a zero, denormal-reciprocal or
	// non-finite direction component
	uint32_t octant = GeneralOctant;
	auto	 chooseOctant = [&]() {
		const uint32_t mine = ( ray.direction.x < 0.0f ? 1u : 0u ) | ( ray.direction.y < 0.0f ? 2u : 0u ) | ( ray.direction.z < 0.0f ? 4u : 0u );
		const bool regular = isfinite( invD.x ) && isfinite( invD.y ) && isfinite( invD.z ) && invD.x != 0.0f && invD.y != 0.0f && invD.z != 0.0f &&
							 isfinite( ray.origin.x ) && isfinite( ray.origin.y ) && isfinite( ray.origin.z );
		const uint32_t liveLanes = __ballot_sync( FullMask, live );
		const uint32_t first	 = __shfl_sync( FullMask, mine, liveLanes ? __ffs( liveLanes ) - 1 : 0 );
		octant					 = __all_sync( FullMask, !live || ( regular && mine == first ) ) ? first : GeneralOctant;
	};
	if ( Refill == 0 && !Q ) chooseOctant();
	// (Q) a ray the quantised walk must not see: traced exactly by its lane right here; the lane then waits as a finished one
#define HIPRT_B200_EXACT_IF_IRREGULAR()                                                                                                    \
	if constexpr ( Q )                                                                                                                     \
	{                                                                                                                                      \
		if ( live && node != InvalidValue && ( qInvalid || !regularForQ( ray ) ) )                                                         \
		{                                                                                                                                  \
			overflow = !traceLaneExact<Type, SharedDepth, Hints>( boxNodes, leaves, ray, stackBase, localStack, bestRef, bestPrim );       \
			node	 = InvalidValue;                                                                                                       \
		}                                                                                                                                  \
	}
	HIPRT_B200_EXACT_IF_IRREGULAR()

	// result of the lane's finished ray: uv mapping and normal by re-running the reference-order test on the winning triangle
	auto finish = [&]() {
		hiprtHit hit = emptyHit();
		if ( bestRef != InvalidValue && !overflow )
		{
			if constexpr ( Format == hiprtB200HitFormatPrimID )
				hit.primID = bestPrim;
			else
			{
				ray.maxT = maxT0;
				intersectPairTriangle( leaves, bestRef >> 1, bestRef & 1u, ray, hit );
			}
		}
		storeHitT<Format, Hints>( hits, index, hit );
	};

	auto pop = [&]() -> uint32_t {
		if ( top == 0 ) return InvalidValue;
		--top;
		return top < SharedDepth ? stackBase[top * 32u] : localStack[top - SharedDepth];
	};

	bool exhausted = false; // (Refill) the ray buffer has been handed out completely
	while ( true )
	{
		if constexpr ( Refill != 0 )
		{
			const bool	   idle		= node == InvalidValue;
			const uint32_t idleMask = __ballot_sync( FullMask, idle );
			if ( idleMask == FullMask || ( !exhausted && uint32_t( __popc( idleMask ) ) >= Refill ) )
			{
				if ( idle && live )
				{
					finish();
					live = false;
				}
				if ( !exhausted )
				{
					const uint32_t count = __popc( idleMask );
					uint32_t	   base	 = 0;
					if ( ( threadIdx.x & 31u ) == 0 ) base = atomicAdd( workCounter, count );
					base	  = __shfl_sync( FullMask, base, 0 );
					exhausted = base + count >= numRays;
					const uint32_t mine = base + __popc( idleMask & ( ( 1u << ( threadIdx.x & 31u ) ) - 1u ) );
					if ( idle && mine < numRays )
					{
						index	 = mine;
						ray		 = loadRayT<Hints>( rays, index );
						invD	 = rcp3( ray.direction );
						maxT0	 = ray.maxT;
						bestRef	 = InvalidValue;
						bestPrim = InvalidValue;
						overflow = false;
						top		 = 0;
						node	 = RootIndex;
						live	 = true;
					}
					HIPRT_B200_EXACT_IF_IRREGULAR()
				}
				if ( !__any_sync( FullMask, node != InvalidValue ) && exhausted ) break;
			}
		}
		else
		{
			if ( !__any_sync( FullMask, node != InvalidValue ) ) break;
		}

		// ---- box phase
		while ( true )
		{
			const bool	   isBox	= isBoxIndex( node );
			const uint32_t boxLanes = __ballot_sync( FullMask, isBox );
			if ( boxLanes == 0 ) break;
			if ( BoxQuorum > 1 && __popc( boxLanes ) < BoxQuorum && __any_sync( FullMask, !isBox && node != InvalidValue ) ) break;
			if ( isBox )
			{
				uint32_t c[4];
				uint32_t n;
				if constexpr ( Stats ) ++boxVisits;
				if constexpr ( Q )
					n = visitBox4Q<Hints, Sort>( qNodes, node, ray.origin, invD, ray.maxT, c );
				else
				{
				switch ( octant ) // warp-uniform
				{
				case 0: n = visitBox4<0, Hints, Sort>( boxNodes, node, ray.origin, invD, ray.maxT, c ); break;
				case 1: n = visitBox4<1, Hints, Sort>( boxNodes, node, ray.origin, invD, ray.maxT, c ); break;
				case 2: n = visitBox4<2, Hints, Sort>( boxNodes, node, ray.origin, invD, ray.maxT, c ); break;
				case 3: n = visitBox4<3, Hints, Sort>( boxNodes, node, ray.origin, invD, ray.maxT, c ); break;
				case 4: n = visitBox4<4, Hints, Sort>( boxNodes, node, ray.origin, invD, ray.maxT, c ); break;
				case 5: n = visitBox4<5, Hints, Sort>( boxNodes, node, ray.origin, invD, ray.maxT, c ); break;
				case 6: n = visitBox4<6, Hints, Sort>( boxNodes, node, ray.origin, invD, ray.maxT, c ); break;
				case 7: n = visitBox4<7, Hints, Sort>( boxNodes, node, ray.origin, invD, ray.maxT, c ); break;
				default: n = visitBox4<GeneralOctant, Hints, Sort>( boxNodes, node, ray.origin, invD, ray.maxT, c ); break;
				}
				}
				if constexpr ( !Sort )
				{
					// n = mask of hit children: pushed in slot order (predicated stores), the last one pushed is popped again
					if ( top + 4 <= SharedDepth )
					{
#pragma unroll
						for ( int k = 0; k < 4; ++k )
							if ( ( n >> k ) & 1u )
							{
								stackBase[top * 32u] = c[k];
								++top;
							}
						node = InvalidValue;
						if ( top != 0 )
						{
							--top;
							node = stackBase[top * 32u];
						}
					}
					else if ( top + 3 > TotalDepth )
					{
						overflow = true;
						node	 = InvalidValue;
						top		 = 0;
					}
					else
					{
#pragma unroll
						for ( int k = 0; k < 4; ++k )
							if ( ( n >> k ) & 1u )
							{
								if ( top < SharedDepth )
									stackBase[top * 32u] = c[k];
								else
									localStack[top - SharedDepth] = c[k];
								++top;
							}
						node = pop();
					}
				}
				// hit children: c[4-n..3], nearest first. Far-to-near onto the stack, then the nearest is popped.
				else if ( top + 4 <= SharedDepth )
				{
					uint32_t* slot = stackBase + top * 32u;
					slot[0] = c[3], slot[32] = c[2], slot[64] = c[1], slot[96] = c[0];
					top += n;
					node = InvalidValue;
					if ( top != 0 )
					{
						--top;
						node = stackBase[top * 32u];
					}
				}
				else if ( top + 3 > TotalDepth )
				{
					overflow = true;
					node	 = InvalidValue;
					top		 = 0;
				}
				else
				{
#pragma unroll
					for ( int k = 3; k >= 0; --k )
						if ( c[k] != InvalidValue )
						{
							if ( top < SharedDepth )
								stackBase[top * 32u] = c[k];
							else
								localStack[top - SharedDepth] = c[k];
							++top;
						}
					node = pop();
				}
			}
		}

		// ---- leaf phase
		if ( node != InvalidValue && !isBoxIndex( node ) )
		{
			if constexpr ( Stats ) ++leafVisits;
			const bool found = visitPairLeaf<Type, Hints, Q>( leaves, node >> 3, ray, invD, bestRef, bestPrim, maxT0 );
			if ( Type == hiprtTraversalTerminateAtAnyHit && found )
			{
				node = InvalidValue;
				top	 = 0;
			}
			else
				node = pop();
		}
	}

#undef HIPRT_B200_EXACT_IF_IRREGULAR
	if ( Refill == 0 && live ) finish();
	if constexpr ( Stats )
	{
		boxVisits  = __reduce_add_sync( FullMask, boxVisits );
		leafVisits = __reduce_add_sync( FullMask, leafVisits );
		if ( ( threadIdx.x & 31u ) == 0 )
		{
			atomicAdd( stats, static_cast<unsigned long long>( boxVisits ) );
			atomicAdd( stats + 1, static_cast<unsigned long long>( leafVisits ) );
		}
	}
}

// Persistent = the grid only fills the machine and every warp keeps fetching batches of 32 consecutive rays from a
// global counter: ray costs are heavy-tailed (silhouette rays visit 10x the mean number of nodes), and a block-per-128-rays
// grid holds a block slot until its slowest warp is done while the other three warp slots sit empty.
// Refill > 0 (persistent only): a warp whose fetched batch is INCOHERENT (mixed direction octants: AO / diffuse bounces)
// keeps that batch and from then on refills idle lanes one by one; ncu shows 15 of 32 lanes active on 8-spp AO rays without
// it. Coherent batches (primary, shadow rays) stay warp-granular: refilling them lane by lane was measured 12-18 % slower.
// Q: `qRegion` = the geometry's Q region (QHeader + QBox4Node per box node, qnode.h); box nodes are read from there.
template <hiprtTraversalType Type, uint32_t Block, uint32_t SharedDepth, uint32_t BoxQuorum, bool Persistent, uint32_t Refill = 0, uint32_t Format = hiprtB200HitFormatFull, bool Hints = false, bool Stats = false, bool Q = false>
__global__ void __launch_bounds__( Block, Refill != 0 ? 1024u / Block : 1u ) traceTrianglesKernel(
	hiprtGeometry geom, const hiprtRay* __restrict__ rays, uint32_t numRays, void* __restrict__ hits, uint32_t* workCounter, unsigned long long* stats,
	const void* __restrict__ qRegion, const uint32_t* __restrict__ numRaysDevice )
{
	__shared__ uint32_t sStack[Block * SharedDepth];
	// indirect launch (hiprtB200TraceGeometryIndirect): the ray count is a device word written by an earlier launch of the stream
	// (a wavefront's append counter); `numRays` is then the capacity the grid was sized for
	if ( numRaysDevice != nullptr ) numRays = min( numRays, __ldg( numRaysDevice ) );
	const GeomHeader*	header					 = reinterpret_cast<const GeomHeader*>( geom );
	const Box4Node* __restrict__ boxNodes		 = header->boxNodes;
	const TrianglePairNode* __restrict__ leaves = reinterpret_cast<const TrianglePairNode*>( header->primNodes );
	uint32_t* const stackBase = sStack + ( threadIdx.x & 31u ) + ( threadIdx.x >> 5 ) * 32u * SharedDepth;
	const QBox4Node* __restrict__ qNodes = nullptr;
	bool qInvalid						 = false;
	if constexpr ( Q )
	{
		qNodes	 = reinterpret_cast<const QBox4Node*>( static_cast<const uint8_t*>( qRegion ) + sizeof( QHeader ) );
		qInvalid = __ldg( &static_cast<const QHeader*>( qRegion )->invalid ) != 0u;
	}
	if constexpr ( Persistent )
	{
		const uint32_t lane = threadIdx.x & 31u;
		while ( true )
		{
			uint32_t base = 0;
			if ( lane == 0 ) base = atomicAdd( workCounter, 32u );
			base = __shfl_sync( FullMask, base, 0 );
			if ( base >= numRays ) return;
			const uint32_t index = base + lane;
			const bool	   live	 = index < numRays;
			if constexpr ( Refill != 0 )
			{
				// direction octant of the lane's ray (the two 16-byte loads are repeated inside the traversal: L1 hits)
				const float4   r1	  = __ldg( reinterpret_cast<const float4*>( rays + ( live ? index : base ) ) + 1 );
				const uint32_t mine	  = ( r1.x < 0.0f ? 1u : 0u ) | ( r1.y < 0.0f ? 2u : 0u ) | ( r1.z < 0.0f ? 4u : 0u );
				const uint32_t first  = __shfl_sync( FullMask, mine, 0 );
				const bool	   coherent = __all_sync( FullMask, mine == first );
				if ( !coherent )
				{
					traceTrianglesWarp<Type, SharedDepth, BoxQuorum, Refill, Format, Hints, Stats, Q>( boxNodes, leaves, rays, hits, index, live, stackBase, workCounter, numRays, stats, qNodes, qInvalid );
					return; // the refilling traversal drains the ray buffer
				}
			}
			traceTrianglesWarp<Type, SharedDepth, BoxQuorum, 0, Format, Hints, Stats, Q>( boxNodes, leaves, rays, hits, index, live, stackBase, nullptr, 0, stats, qNodes, qInvalid );
		}
	}
	else
	{
		const uint32_t index = blockIdx.x * Block + threadIdx.x;
		traceTrianglesWarp<Type, SharedDepth, BoxQuorum, 0, Format, Hints, Stats, Q>( boxNodes, leaves, rays, hits, index, index < numRays, stackBase, nullptr, 0, stats, qNodes, qInvalid );
	}
}

template <hiprtTraversalType Type>
__global__ void __launch_bounds__( TraceBlock ) traceSceneKernel(
	hiprtScene scene, const hiprtRay* __restrict__ rays, uint32_t numRays, hiprtHit* __restrict__ hits, hiprtRayMask mask, const float* __restrict__ times, float time )
{
	__shared__ uint32_t sStack[TraceBlock * SharedStackDepth];
	const uint32_t		index = blockIdx.x * TraceBlock + threadIdx.x;
	if ( index >= numRays ) return;
	HybridStack				stack( sStack );
	hiprtEmptyInstanceStack noInstanceStack;
	const float4*			rp = reinterpret_cast<const float4*>( rays + index );
	const float4			r0 = __ldg( rp ), r1 = __ldg( rp + 1 );
	hiprtRay				ray;
	ray.origin = make_float3( r0.x, r0.y, r0.z ), ray.minT = r0.w;
	ray.direction = make_float3( r1.x, r1.y, r1.z ), ray.maxT = r1.w;
	if ( times ) time = __ldg( times + index );
	hiprtSceneTraversalT<HybridStack, hiprtEmptyInstanceStack, Type> tr(
		scene, ray, stack, noInstanceStack, mask, hiprtTraversalHintDefault, nullptr, nullptr, 0, time );
	storeSingleLevelHit( hits, index, tr.getNextHit() );
}

static int traceVariant()
{
	static const int v = [] {
		const char* e = getenv( "HIPRT_B200_TRACE_VARIANT" );
		return e ? atoi( e ) : -1;
	}();
	return v;
}

template <hiprtTraversalType Type, uint32_t Block, uint32_t SharedDepth, bool Persistent>
static void launchGeometry( hiprtGeometry geom, const hiprtRay* rays, uint32_t numRays, hiprtHit* hits, uint32_t* workCounter, uint32_t smCount, cudaStream_t stream )
{
	auto kernel = traceGeometryKernel<Type, Block, SharedDepth, Persistent>;
	if constexpr ( Persistent )
	{
		static int blocksPerSm = 0; // same value for every caller: a benign race
		if ( blocksPerSm == 0 ) cudaOccupancyMaxActiveBlocksPerMultiprocessor( &blocksPerSm, kernel, int( Block ), 0 );
		cudaMemsetAsync( workCounter, 0, sizeof( uint32_t ), stream );
		const uint32_t blocks = std::min( DivideRoundUp( numRays, Block ), smCount * uint32_t( std::max( blocksPerSm, 1 ) ) );
		kernel<<<blocks, Block, 0, stream>>>( geom, rays, numRays, hits, workCounter );
	}
	else
		kernel<<<DivideRoundUp( numRays, Block ), Block, 0, stream>>>( geom, rays, numRays, hits, workCounter );
}

// The configuration the round-1 sweeps settled on: blocks of 128 threads, 32 stack entries per lane in
// shared memory, box-phase quorum 16, persistent warps, lane refill (threshold 8) for incoherent batches.
// HIPRT_B200_TRACE_VARIANT: 1 = generic traversal objects (what a user kernel compiles), 2 = fast kernel without cache hints,
// 3 = fast kernel on the public 128-byte box nodes even when the geometry has a Q region (the round-1 kernel).
template <hiprtTraversalType Type, uint32_t Format, bool Hints, bool Stats = false, bool Q = false>
static void launchTrianglesT(
	hiprtGeometry geom, const hiprtRay* rays, uint32_t numRays, void* hits, uint32_t* workCounter, uint32_t smCount, cudaStream_t stream, unsigned long long* stats = nullptr,
	const void* qRegion = nullptr, const uint32_t* numRaysDevice = nullptr )
{
	constexpr uint32_t Block = 128;
	auto			   kernel = traceTrianglesKernel<Type, Block, 32, 16, true, 8, Format, Hints, Stats, Q>;
	static int		   blocksPerSm = 0; // same value for every caller: a benign race
	if ( blocksPerSm == 0 ) cudaOccupancyMaxActiveBlocksPerMultiprocessor( &blocksPerSm, kernel, int( Block ), 0 );
	cudaMemsetAsync( workCounter, 0, sizeof( uint32_t ), stream );
	const uint32_t blocks = std::min( DivideRoundUp( numRays, Block ), smCount * uint32_t( std::max( blocksPerSm, 1 ) ) );
	kernel<<<blocks, Block, 0, stream>>>( geom, rays, numRays, hits, workCounter, stats, qRegion, numRaysDevice );
}

// Which queries walk the Q records (qnode.h) when the geometry has them. ncu on the C3 AO rays (profiles/r02_ncu_trace_c3_ao_q.csv):
// the records halve the L1 data-pipe load (81 % -> 46 %) but their visit costs more ALU-pipe instructions (byte -> float, entry /
// exit selection), and with the child ordering network of closest-hit queries on top the ALU pipe becomes the limiter (75 %):
// closest-hit came out 13 % SLOWER on incoherent rays and 2 % slower on primary rays, any-hit (no ordering) 12 % faster. So
// any-hit queries take the Q walk, closest-hit queries the exact one. HIPRT_B200_TRACE_VARIANT=4 forces the Q walk for both.
static bool useQ( bool any, const void* qRegion ) { return qRegion != nullptr && ( any || traceVariant() == 4 ); }

template <uint32_t Format, bool Hints, bool Q = false>
static void launchTriangles(
	hiprtGeometry geom, const hiprtRay* rays, uint32_t numRays, void* hits, bool any, uint32_t* workCounter, uint32_t smCount, cudaStream_t stream, const void* qRegion = nullptr,
	const uint32_t* numRaysDevice = nullptr )
{
	if ( any )
		launchTrianglesT<hiprtTraversalTerminateAtAnyHit, Format, Hints, false, Q>( geom, rays, numRays, hits, workCounter, smCount, stream, nullptr, qRegion, numRaysDevice );
	else
		launchTrianglesT<hiprtTraversalTerminateAtClosestHit, Format, Hints, false, Q>( geom, rays, numRays, hits, workCounter, smCount, stream, nullptr, qRegion, numRaysDevice );
}

// Instrumented launch (SURVEY.md 8d "counted by an instrumented build"): the same kernel with per-lane visit counters, summed
// into stats[0] (box-node visits) and stats[1] (leaf visits; on the quantised walk this counts the leaves REACHED, of which the
// exact box test then admits the same ones as the exact walk). Used by bench.py outside the timed region for the cache-level
// traffic of a ray batch (64 or 128 B per box visit + 64 B per leaf visit).
void traceGeometryStats(
	hiprtGeometry geom, const hiprtRay* rays, uint32_t numRays, hiprtHit* hits, hiprtTraversalType type, uint32_t* workCounter, uint32_t smCount, cudaStream_t stream,
	unsigned long long* stats, const void* qRegion )
{
	if ( numRays == 0 ) return;
	const bool any = type == hiprtTraversalTerminateAtAnyHit;
	if ( useQ( any, qRegion ) && traceVariant() != 3 )
	{
		if ( any )
			launchTrianglesT<hiprtTraversalTerminateAtAnyHit, hiprtB200HitFormatFull, true, true, true>( geom, rays, numRays, hits, workCounter, smCount, stream, stats, qRegion );
		else
			launchTrianglesT<hiprtTraversalTerminateAtClosestHit, hiprtB200HitFormatFull, true, true, true>( geom, rays, numRays, hits, workCounter, smCount, stream, stats, qRegion );
	}
	else if ( any )
		launchTrianglesT<hiprtTraversalTerminateAtAnyHit, hiprtB200HitFormatFull, true, true>( geom, rays, numRays, hits, workCounter, smCount, stream, stats );
	else
		launchTrianglesT<hiprtTraversalTerminateAtClosestHit, hiprtB200HitFormatFull, true, true>( geom, rays, numRays, hits, workCounter, smCount, stream, stats );
}

// Generic kernel output in another record format: traced into full hits is not possible without a staging buffer, so the
// generic (custom-primitive) path supports the full format only; api.cpp rejects the others for such geometries.
// qRegion: the geometry's Q region when it has one (api.cpp's handle map), else nullptr.
void traceGeometry(
	hiprtGeometry geom, bool triangles, const hiprtRay* rays, uint32_t numRays, void* hits, hiprtTraversalType type, uint32_t* workCounter, uint32_t smCount,
	cudaStream_t stream, uint32_t format, const void* qRegion, const uint32_t* numRaysDevice )
{
	if ( numRays == 0 ) return;
	if ( numRaysDevice != nullptr && ( !triangles || traceVariant() == 1 ) ) return; // (api.cpp rejects indirect launches on custom-primitive geometries)
	const bool any = type == hiprtTraversalTerminateAtAnyHit;
	const int  variant = traceVariant();
	if ( triangles && ( variant != 1 || format != hiprtB200HitFormatFull ) )
	{
		const bool hints = variant != 2;
		if ( useQ( any, qRegion ) && hints && variant != 3 )
		{
			switch ( format )
			{
			case hiprtB200HitFormatCompact: launchTriangles<hiprtB200HitFormatCompact, true, true>( geom, rays, numRays, hits, any, workCounter, smCount, stream, qRegion, numRaysDevice ); return;
			case hiprtB200HitFormatPrimID: launchTriangles<hiprtB200HitFormatPrimID, true, true>( geom, rays, numRays, hits, any, workCounter, smCount, stream, qRegion, numRaysDevice ); return;
			default: launchTriangles<hiprtB200HitFormatFull, true, true>( geom, rays, numRays, hits, any, workCounter, smCount, stream, qRegion, numRaysDevice ); return;
			}
		}
		switch ( format )
		{
		case hiprtB200HitFormatCompact:
			if ( hints )
				launchTriangles<hiprtB200HitFormatCompact, true>( geom, rays, numRays, hits, any, workCounter, smCount, stream, nullptr, numRaysDevice );
			else
				launchTriangles<hiprtB200HitFormatCompact, false>( geom, rays, numRays, hits, any, workCounter, smCount, stream, nullptr, numRaysDevice );
			return;
		case hiprtB200HitFormatPrimID:
			if ( hints )
				launchTriangles<hiprtB200HitFormatPrimID, true>( geom, rays, numRays, hits, any, workCounter, smCount, stream, nullptr, numRaysDevice );
			else
				launchTriangles<hiprtB200HitFormatPrimID, false>( geom, rays, numRays, hits, any, workCounter, smCount, stream, nullptr, numRaysDevice );
			return;
		default:
			if ( hints )
				launchTriangles<hiprtB200HitFormatFull, true>( geom, rays, numRays, hits, any, workCounter, smCount, stream, nullptr, numRaysDevice );
			else
				launchTriangles<hiprtB200HitFormatFull, false>( geom, rays, numRays, hits, any, workCounter, smCount, stream, nullptr, numRaysDevice );
			return;
		}
	}
	hiprtHit* fullHits = static_cast<hiprtHit*>( hits );
	if ( any )
		launchGeometry<hiprtTraversalTerminateAtAnyHit, 128, 32, false>( geom, rays, numRays, fullHits, workCounter, smCount, stream );
	else
		launchGeometry<hiprtTraversalTerminateAtClosestHit, 128, 32, false>( geom, rays, numRays, fullHits, workCounter, smCount, stream );
}

// ------------------------------------------------------------------------------------------------
// Fast path for single-level scenes (what the library's own scene kernel covers: instances of geometries, static
// transforms, no custom intersection functions -- the library's default table rejects every custom primitive, so
// instances of custom-primitive geometries are skipped). Same structure as traceTrianglesWarp: while-while with warp votes,
// stack top in a register, four children stored unconditionally, closest hit kept as (t, instance leaf, pair leaf,
// triangle) and completed once at the end. An instance is entered by pushing a marker; popping it restores the world ray.
// Arithmetic (transform, slab, triangle) is that of the generic hiprtSceneTraversalT, winner = min (t, instanceID, primID).
constexpr uint32_t InstanceExitMarker = 0xFFFFFFFEu; // instance-typed index no scene can hold (2^24 instance limit)
constexpr uint32_t DeferredMarker	  = 0xFFFFFFFEu; // primID written for a ray that is finished by traceSceneDeferredKernel

// world -> object ray of an instance leaf: inverse matrix for static instances; animated ones (key-frame range in the node)
// are posed at the ray's time (hiprtB200TraceSceneAtTime; the plain entry points trace at time 0)
__device__ __forceinline__ hiprtRay instanceRayToObject( const SceneHeader* scene, const InstanceNode* in, uint32_t bits, const hiprtRay& ray, float time )
{
	if ( ( bits >> 25 ) & 1u ) return transformRayToObject( in, ray );
	const uint2 range = __ldg( reinterpret_cast<const uint2*>( in ) );
	return motion::transformRay( motion::frameAt( scene->frames, range.x, range.y, time ), ray );
}

template <hiprtTraversalType Type, uint32_t SharedDepth, uint32_t BoxQuorum>
__device__ __forceinline__ void traceSceneWarp(
	const SceneHeader* __restrict__ scene,
	const hiprtRay* __restrict__ rays,
	hiprtHit* __restrict__ hits,
	uint32_t		index,
	bool			live,
	hiprtRayMask	rayMask,
	uint32_t* const stackBase,
	float			time,
	uint32_t*		deferredCount ) // rays that met an instance of a SCENE are counted here and left to traceSceneDeferredKernel
{
	constexpr uint32_t TotalDepth = 64;
	uint32_t		   localStack[TotalDepth - SharedDepth];
	uint32_t		   top = 0;

	const Box4Node* __restrict__ sceneBoxes	   = scene->boxNodes;
	const InstanceNode* __restrict__ instNodes = scene->primNodes;
	const Box4Node*			boxNodes		   = sceneBoxes;
	const TrianglePairNode* leaves			   = nullptr;

	hiprtRay	ray		 = loadRay( rays, live ? index : 0u );
	float3		invD	 = rcp3( ray.direction );
	const float maxT0	 = ray.maxT;
	uint32_t	bestRef	 = InvalidValue; // (pair leaf address << 1) | triangle
	uint32_t	bestPrim = InvalidValue, bestId = InvalidValue, bestInst = InvalidValue;
	uint32_t	curInst	 = InvalidValue; // instance leaf (node index) the ray is inside of
	uint32_t	curId	 = InvalidValue;
	bool		overflow = false;
	bool		deferred = false;
	uint32_t	node	 = live ? RootIndex : InvalidValue;

	auto push = [&]( uint32_t v ) {
		if ( top < SharedDepth )
			stackBase[top * 32u] = v;
		else
			localStack[top - SharedDepth] = v;
		++top;
	};
	auto pop = [&]() -> uint32_t {
		if ( top == 0 ) return InvalidValue;
		--top;
		return top < SharedDepth ? stackBase[top * 32u] : localStack[top - SharedDepth];
	};

	while ( __any_sync( FullMask, node != InvalidValue ) )
	{
		// ---- box phase (scene level and geometry level alike)
		while ( true )
		{
			const bool	   isBox	= isBoxIndex( node );
			const uint32_t boxLanes = __ballot_sync( FullMask, isBox );
			if ( boxLanes == 0 ) break;
			if ( BoxQuorum > 1 && __popc( boxLanes ) < BoxQuorum && __any_sync( FullMask, !isBox && node != InvalidValue ) ) break;
			if ( isBox )
			{
				uint32_t	   c[4];
				const uint32_t n = visitBox4<GeneralOctant>( boxNodes, node, ray.origin, invD, ray.maxT, c );
				if ( top + 4 <= SharedDepth )
				{
					uint32_t* slot = stackBase + top * 32u;
					slot[0] = c[3], slot[32] = c[2], slot[64] = c[1], slot[96] = c[0];
					top += n;
					node = InvalidValue;
					if ( top != 0 )
					{
						--top;
						node = stackBase[top * 32u];
					}
				}
				else if ( top + 4 > TotalDepth )
				{
					overflow = true;
					node	 = InvalidValue;
					top		 = 0;
				}
				else
				{
#pragma unroll
					for ( int k = 3; k >= 0; --k )
						if ( c[k] != InvalidValue ) push( c[k] );
					node = pop();
				}
			}
		}

		// ---- leaf phase: leave an instance / enter an instance / test a triangle pair
		if ( node != InvalidValue && !isBoxIndex( node ) )
		{
			if ( node == InstanceExitMarker )
			{
				const float	   maxT	 = ray.maxT;
				ray					 = loadRay( rays, index );
				ray.maxT			 = maxT;
				invD				 = rcp3( ray.direction );
				boxNodes			 = sceneBoxes;
				curInst				 = InvalidValue;
				node				 = pop();
			}
			else if ( curInst == InvalidValue )
			{
				const InstanceNode* in	 = instNodes + ( node >> 4 );
				const uint4			meta = __ldg( reinterpret_cast<const uint4*>( in ) + 3 ); // {geometry.lo, geometry.hi, mask, bits}
				bool				enter = ( meta.z & rayMask ) != 0u;
				if ( enter && ( ( meta.w >> 24 ) & 1u ) )
				{
					// an instance of a scene (multi-level instancing): this kernel walks one instance level; the ray is handed to
					// the generic multi-level traversal that runs after it (traceSceneDeferredKernel)
					deferred = true;
					enter	 = false;
					top		 = 0;
				}
				if ( enter )
				{
					const GeomHeader* g = reinterpret_cast<const GeomHeader*>( ( uint64_t( meta.y ) << 32 ) | meta.x );
					const uint4 ptrs	= __ldg( reinterpret_cast<const uint4*>( g ) ); // {boxNodes, primNodes}
					const uint32_t geomType = __ldg( &g->geomType );
					enter					= geomType & 1u; // custom primitives cannot hit through the default function table
					if ( enter && top + 1 >= TotalDepth )
					{
						overflow = true;
						enter	 = false;
						top		 = 0;
					}
					if ( enter )
					{
						push( InstanceExitMarker );
						if ( !( ( meta.w >> 26 ) & 1u ) )
						{
							const float maxT = ray.maxT;
							ray				 = instanceRayToObject( scene, in, meta.w, ray, time );
							ray.maxT		 = maxT;
							invD			 = rcp3( ray.direction );
						}
						boxNodes = reinterpret_cast<const Box4Node*>( ( uint64_t( ptrs.y ) << 32 ) | ptrs.x );
						leaves	 = reinterpret_cast<const TrianglePairNode*>( ( uint64_t( ptrs.w ) << 32 ) | ptrs.z );
						curInst	 = node;
						curId	 = meta.w & 0xFFFFFFu;
						node	 = RootIndex;
					}
				}
				if ( !enter ) node = pop();
			}
			else
			{
				const uint32_t addr = node >> 3;
				const void*	   p	= leaves + addr;
				const F8	   l0 = ldg8<false>( p, 0 ), l1 = ldg8<false>( p, 1 );
				const uint32_t prim0 = __float_as_uint( l1.v[5] ), prim1 = __float_as_uint( l1.v[6] );
				const float3   pv0 = make_float3( l0.v[0], l0.v[1], l0.v[2] ), pv1 = make_float3( l0.v[3], l0.v[4], l0.v[5] );
				const float3   pv2 = make_float3( l0.v[6], l0.v[7], l1.v[0] ), pv3 = make_float3( l1.v[1], l1.v[2], l1.v[3] );
				bool		   found = false;
				HitTriangle	   h;
				if ( intersectTriangle( pv0, pv1, pv2, ray, h ) && ( h.t < ray.maxT || curId < bestId || ( curId == bestId && prim0 < bestPrim ) ) )
				{
					ray.maxT = h.t, bestRef = addr << 1, bestPrim = prim0, bestId = curId, bestInst = curInst;
					found = true;
				}
				if ( prim0 != prim1 && !( Type == hiprtTraversalTerminateAtAnyHit && found ) )
				{
					if ( intersectTriangle( pv1, pv3, pv2, ray, h ) && ( h.t < ray.maxT || curId < bestId || ( curId == bestId && prim1 < bestPrim ) ) )
					{
						ray.maxT = h.t, bestRef = ( addr << 1 ) | 1u, bestPrim = prim1, bestId = curId, bestInst = curInst;
						found = true;
					}
				}
				if ( Type == hiprtTraversalTerminateAtAnyHit && found )
				{
					node = InvalidValue;
					top	 = 0;
				}
				else
					node = pop();
			}
		}
	}

	if ( !live ) return;
	hiprtHit hit = emptyHit();
	if ( deferred )
	{
		hit.primID = DeferredMarker;
		atomicAdd( deferredCount, 1u );
	}
	else if ( bestRef != InvalidValue && !overflow )
	{
		const InstanceNode* in	 = instNodes + ( bestInst >> 4 );
		const uint4			meta = __ldg( reinterpret_cast<const uint4*>( in ) + 3 );
		const GeomHeader*	g	 = reinterpret_cast<const GeomHeader*>( ( uint64_t( meta.y ) << 32 ) | meta.x );
		hiprtRay			objectRay = loadRay( rays, index );
		if ( !( ( meta.w >> 26 ) & 1u ) ) objectRay = instanceRayToObject( scene, in, meta.w, objectRay, time );
		objectRay.maxT	= maxT0;
		hit.instanceID	= meta.w & 0xFFFFFFu;
		intersectPairTriangle( reinterpret_cast<const TrianglePairNode*>( g->primNodes ), bestRef >> 1, bestRef & 1u, objectRay, hit );
	}
	storeHit( hits, index, hit );
}

template <hiprtTraversalType Type, uint32_t Block, uint32_t SharedDepth, uint32_t BoxQuorum>
__global__ void __launch_bounds__( Block, 768 / Block ) traceSceneFastKernel( // 6 blocks of 128 per SM: caps the kernel at 80 registers (84 cost a block)
	hiprtScene scene, const hiprtRay* __restrict__ rays, uint32_t numRays, hiprtHit* __restrict__ hits, hiprtRayMask mask, uint32_t* workCounter,
	const float* __restrict__ times, float time )
{
	__shared__ uint32_t sStack[Block * SharedDepth];
	uint32_t* const		stackBase = sStack + ( threadIdx.x & 31u ) + ( threadIdx.x >> 5 ) * 32u * SharedDepth;
	const uint32_t		lane	  = threadIdx.x & 31u;
	while ( true )
	{
		uint32_t base = 0;
		if ( lane == 0 ) base = atomicAdd( workCounter, 32u );
		base = __shfl_sync( FullMask, base, 0 );
		if ( base >= numRays ) return;
		const float rayTime = ( times && base + lane < numRays ) ? __ldg( times + base + lane ) : time;
		traceSceneWarp<Type, SharedDepth, BoxQuorum>(
			reinterpret_cast<const SceneHeader*>( scene ), rays, hits, base + lane, base + lane < numRays, mask, stackBase, rayTime, workCounter + 1 );
	}
}

// Second launch of every library scene trace: finishes the rays the fast kernel deferred (multi-level instancing: instances of
// scenes, hiprt_device_impl.h:1105-1193 in the reference) with the generic traversal object and an instance stack, as a user kernel
// would. When the fast kernel deferred nothing -- every single-level scene -- each block reads one word and leaves.
template <hiprtTraversalType Type>
__global__ void __launch_bounds__( TraceBlock ) traceSceneDeferredKernel(
	hiprtScene scene, const hiprtRay* __restrict__ rays, uint32_t numRays, hiprtHit* __restrict__ hits, hiprtRayMask mask, const uint32_t* deferredCount,
	const float* __restrict__ times, float time )
{
	if ( *reinterpret_cast<const volatile uint32_t*>( deferredCount ) == 0u ) return;
	__shared__ uint32_t sStack[TraceBlock * SharedStackDepth];
	for ( uint32_t index = blockIdx.x * TraceBlock + threadIdx.x; index < numRays; index += gridDim.x * TraceBlock )
	{
		if ( hits[index].primID != DeferredMarker ) continue;
		HybridStack				  stack( sStack );
		hiprtPrivateInstanceStack instanceStack;
		const hiprtRay			  ray = loadRay( rays, index );
		hiprtSceneTraversalT<HybridStack, hiprtPrivateInstanceStack, Type> tr(
			scene, ray, stack, instanceStack, mask, hiprtTraversalHintDefault, nullptr, nullptr, 0, times ? __ldg( times + index ) : time );
		storeHit( hits, index, tr.getNextHit() );
	}
}

template <hiprtTraversalType Type>
static void launchSceneFast(
	hiprtScene scene, const hiprtRay* rays, uint32_t numRays, hiprtHit* hits, hiprtRayMask mask, uint32_t* workCounter, uint32_t smCount, cudaStream_t stream,
	const float* times, float time )
{
	constexpr uint32_t Block = 128;
	auto			   kernel = traceSceneFastKernel<Type, Block, 32, 16>;
	static int		   blocksPerSm = 0; // same value for every caller: a benign race
	if ( blocksPerSm == 0 ) cudaOccupancyMaxActiveBlocksPerMultiprocessor( &blocksPerSm, kernel, int( Block ), 0 );
	cudaMemsetAsync( workCounter, 0, 2 * sizeof( uint32_t ), stream ); // [0] ray batches handed out, [1] rays deferred to the multi-level pass
	const uint32_t blocks = std::min( DivideRoundUp( numRays, Block ), smCount * uint32_t( std::max( blocksPerSm, 1 ) ) );
	kernel<<<blocks, Block, 0, stream>>>( scene, rays, numRays, hits, mask, workCounter, times, time );
	traceSceneDeferredKernel<Type><<<std::min( DivideRoundUp( numRays, TraceBlock ), smCount * 4u ), TraceBlock, 0, stream>>>(
		scene, rays, numRays, hits, mask, workCounter + 1, times, time );
}

void traceScene(
	hiprtScene scene, const hiprtRay* rays, uint32_t numRays, hiprtHit* hits, hiprtTraversalType type, hiprtRayMask mask, uint32_t* workCounter, uint32_t smCount,
	cudaStream_t stream, const float* times, float time )
{
	if ( numRays == 0 ) return;
	if ( traceVariant() != 1 )
	{
		if ( type == hiprtTraversalTerminateAtAnyHit )
			launchSceneFast<hiprtTraversalTerminateAtAnyHit>( scene, rays, numRays, hits, mask, workCounter, smCount, stream, times, time );
		else
			launchSceneFast<hiprtTraversalTerminateAtClosestHit>( scene, rays, numRays, hits, mask, workCounter, smCount, stream, times, time );
		return;
	}

	const uint32_t blocks = DivideRoundUp( numRays, TraceBlock );
	if ( type == hiprtTraversalTerminateAtAnyHit )
		traceSceneKernel<hiprtTraversalTerminateAtAnyHit><<<blocks, TraceBlock, 0, stream>>>( scene, rays, numRays, hits, mask, times, time );
	else
		traceSceneKernel<hiprtTraversalTerminateAtClosestHit><<<blocks, TraceBlock, 0, stream>>>( scene, rays, numRays, hits, mask, times, time );
}
} // namespace hiprt
