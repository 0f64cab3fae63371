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
				storeHit( hits, index, hit );
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
		storeHit( hits, index, hit );
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

// One box-node visit: tests the four children, leaves the hit ones sorted (nearest last) in c[4-n..3].
template <uint32_t Octant, bool Hints = false>
__device__ __forceinline__ uint32_t visitBox4( const Box4Node* __restrict__ boxNodes, uint32_t node, const float3 o, const float3 invD, float maxT, uint32_t c[4] )
{
	const void* p  = boxNodes + ( node >> 4 );
	const F8	w0 = ldg8<Hints>( p, 0 ), w1 = ldg8<Hints>( p, 1 ), w2 = ldg8<Hints>( p, 2 ), w3 = ldg8<Hints>( p, 3 );
	const float b[24] = { w0.v[4], w0.v[5], w0.v[6], w0.v[7], w1.v[0], w1.v[1], w1.v[2], w1.v[3], w1.v[4], w1.v[5], w1.v[6], w1.v[7],
						  w2.v[0], w2.v[1], w2.v[2], w2.v[3], w2.v[4], w2.v[5], w2.v[6], w2.v[7], w3.v[0], w3.v[1], w3.v[2], w3.v[3] };
	c[0] = __float_as_uint( w0.v[0] ), c[1] = __float_as_uint( w0.v[1] ), c[2] = __float_as_uint( w0.v[2] ), c[3] = __float_as_uint( w0.v[3] );
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
	return n;
}

// Traversal of one ray per lane; every lane of the warp must call it (warp votes inside), `live` = has a ray.
// Refill == 0: traces ray `index` of the lane and returns when the whole warp is done.
// Refill  > 0: lane-granular persistent threads -- whenever at least `Refill` lanes are idle (their ray is finished) those
// lanes write their hits and fetch the next rays from `workCounter`, the others keep their state; returns when the ray
// buffer is exhausted and every lane is done. Pays off for incoherent rays only (see the kernel below).
template <hiprtTraversalType Type, uint32_t SharedDepth, uint32_t BoxQuorum, uint32_t Refill = 0, uint32_t Format = hiprtB200HitFormatFull, bool Hints = false, bool Stats = false>
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
	unsigned long long* stats	= nullptr ) // Stats: {box-node visits, leaf visits} of the launch are added here
{
	constexpr uint32_t TotalDepth = 64; // the reference's private stack size
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
	if ( Refill == 0 ) chooseOctant();

	// result of the lane's finished ray: uv mapping and normal by re-running the reference-order test on the winning triangle
	auto finish = [&]() {
		hiprtHit hit;
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
				}
				if ( !__any_sync( FullMask, node != InvalidValue ) ) break;
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
				switch ( octant ) // warp-uniform
				{
				case 0: n = visitBox4<0, Hints>( boxNodes, node, ray.origin, invD, ray.maxT, c ); break;
				case 1: n = visitBox4<1, Hints>( boxNodes, node, ray.origin, invD, ray.maxT, c ); break;
				case 2: n = visitBox4<2, Hints>( boxNodes, node, ray.origin, invD, ray.maxT, c ); break;
				case 3: n = visitBox4<3, Hints>( boxNodes, node, ray.origin, invD, ray.maxT, c ); break;
				case 4: n = visitBox4<4, Hints>( boxNodes, node, ray.origin, invD, ray.maxT, c ); break;
				case 5: n = visitBox4<5, Hints>( boxNodes, node, ray.origin, invD, ray.maxT, c ); break;
				case 6: n = visitBox4<6, Hints>( boxNodes, node, ray.origin, invD, ray.maxT, c ); break;
				case 7: n = visitBox4<7, Hints>( boxNodes, node, ray.origin, invD, ray.maxT, c ); break;
				default: n = visitBox4<GeneralOctant, Hints>( boxNodes, node, ray.origin, invD, ray.maxT, c ); break;
				}
				// hit children: c[4-n..3], nearest first. Far-to-near onto the stack, then the nearest is popped.
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
			const uint32_t addr = node >> 3;
			const void*	   p	= leaves + addr;
			if constexpr ( Stats ) ++leafVisits;
			const F8	   l0 = ldg8<Hints>( p, 0 ), l1 = ldg8<Hints>( p, 1 );
			const uint32_t prim0 = __float_as_uint( l1.v[5] ), prim1 = __float_as_uint( l1.v[6] );
			const float3   pv0 = make_float3( l0.v[0], l0.v[1], l0.v[2] ), pv1 = make_float3( l0.v[3], l0.v[4], l0.v[5] );
			const float3   pv2 = make_float3( l0.v[6], l0.v[7], l1.v[0] ), pv3 = make_float3( l1.v[1], l1.v[2], l1.v[3] );
			bool		   found = false;
			HitTriangle	   h;
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
			if ( Type == hiprtTraversalTerminateAtAnyHit && found )
			{
				node = InvalidValue;
				top	 = 0;
			}
			else
				node = pop();
		}
	}

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
template <hiprtTraversalType Type, uint32_t Block, uint32_t SharedDepth, uint32_t BoxQuorum, bool Persistent, uint32_t Refill = 0, uint32_t Format = hiprtB200HitFormatFull, bool Hints = false, bool Stats = false>
__global__ void __launch_bounds__( Block, Refill != 0 ? 1024u / Block : 1u ) traceTrianglesKernel(
	hiprtGeometry geom, const hiprtRay* __restrict__ rays, uint32_t numRays, void* __restrict__ hits, uint32_t* workCounter, unsigned long long* stats )
{
	__shared__ uint32_t sStack[Block * SharedDepth];
	const GeomHeader*	header					 = reinterpret_cast<const GeomHeader*>( geom );
	const Box4Node* __restrict__ boxNodes		 = header->boxNodes;
	const TrianglePairNode* __restrict__ leaves = reinterpret_cast<const TrianglePairNode*>( header->primNodes );
	uint32_t* const stackBase = sStack + ( threadIdx.x & 31u ) + ( threadIdx.x >> 5 ) * 32u * SharedDepth;
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
					traceTrianglesWarp<Type, SharedDepth, BoxQuorum, Refill, Format, Hints, Stats>( boxNodes, leaves, rays, hits, index, live, stackBase, workCounter, numRays, stats );
					return; // the refilling traversal drains the ray buffer
				}
			}
			traceTrianglesWarp<Type, SharedDepth, BoxQuorum, 0, Format, Hints, Stats>( boxNodes, leaves, rays, hits, index, live, stackBase, nullptr, 0, stats );
		}
	}
	else
	{
		const uint32_t index = blockIdx.x * Block + threadIdx.x;
		traceTrianglesWarp<Type, SharedDepth, BoxQuorum, 0, Format, Hints, Stats>( boxNodes, leaves, rays, hits, index, index < numRays, stackBase, nullptr, 0, stats );
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
	const hiprtHit hit = tr.getNextHit();
	float4*		   hp  = reinterpret_cast<float4*>( hits + index );
	hp[0] = make_float4( __uint_as_float( hit.instanceIDs[0] ), __uint_as_float( hit.instanceIDs[1] ), __uint_as_float( hit.instanceIDs[2] ), __uint_as_float( hit.instanceIDs[3] ) );
	hp[1] = make_float4( __uint_as_float( hit.primID ), 0.0f, hit.uv.x, hit.uv.y );
	hp[2] = make_float4( hit.normal.x, hit.normal.y, hit.normal.z, hit.t );
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
// HIPRT_B200_TRACE_VARIANT: 1 = generic traversal objects (what a user kernel compiles), 2 = fast kernel without cache hints.
template <hiprtTraversalType Type, uint32_t Format, bool Hints, bool Stats = false>
static void launchTrianglesT(
	hiprtGeometry geom, const hiprtRay* rays, uint32_t numRays, void* hits, uint32_t* workCounter, uint32_t smCount, cudaStream_t stream, unsigned long long* stats = nullptr )
{
	constexpr uint32_t Block = 128;
	auto			   kernel = traceTrianglesKernel<Type, Block, 32, 16, true, 8, Format, Hints, Stats>;
	static int		   blocksPerSm = 0; // same value for every caller: a benign race
	if ( blocksPerSm == 0 ) cudaOccupancyMaxActiveBlocksPerMultiprocessor( &blocksPerSm, kernel, int( Block ), 0 );
	cudaMemsetAsync( workCounter, 0, sizeof( uint32_t ), stream );
	const uint32_t blocks = std::min( DivideRoundUp( numRays, Block ), smCount * uint32_t( std::max( blocksPerSm, 1 ) ) );
	kernel<<<blocks, Block, 0, stream>>>( geom, rays, numRays, hits, workCounter, stats );
}

template <uint32_t Format, bool Hints>
static void launchTriangles( hiprtGeometry geom, const hiprtRay* rays, uint32_t numRays, void* hits, bool any, uint32_t* workCounter, uint32_t smCount, cudaStream_t stream )
{
	if ( any )
		launchTrianglesT<hiprtTraversalTerminateAtAnyHit, Format, Hints>( geom, rays, numRays, hits, workCounter, smCount, stream );
	else
		launchTrianglesT<hiprtTraversalTerminateAtClosestHit, Format, Hints>( geom, rays, numRays, hits, workCounter, smCount, stream );
}

// Instrumented launch (SURVEY.md 8d "counted by an instrumented build"): the same kernel with per-lane visit counters, summed
// into stats[0] (box-node visits) and stats[1] (leaf visits). Used by bench.py outside the timed region for the cache-level
// traffic of a ray batch (128 B per box visit + 64 B per leaf visit).
void traceGeometryStats(
	hiprtGeometry geom, const hiprtRay* rays, uint32_t numRays, hiprtHit* hits, hiprtTraversalType type, uint32_t* workCounter, uint32_t smCount, cudaStream_t stream,
	unsigned long long* stats )
{
	if ( numRays == 0 ) return;
	if ( type == hiprtTraversalTerminateAtAnyHit )
		launchTrianglesT<hiprtTraversalTerminateAtAnyHit, hiprtB200HitFormatFull, true, true>( geom, rays, numRays, hits, workCounter, smCount, stream, stats );
	else
		launchTrianglesT<hiprtTraversalTerminateAtClosestHit, hiprtB200HitFormatFull, true, true>( geom, rays, numRays, hits, workCounter, smCount, stream, stats );
}

// Generic kernel output in another record format: traced into full hits is not possible without a staging buffer, so the
// generic (custom-primitive) path supports the full format only; api.cpp rejects the others for such geometries.
void traceGeometry(
	hiprtGeometry geom, bool triangles, const hiprtRay* rays, uint32_t numRays, void* hits, hiprtTraversalType type, uint32_t* workCounter, uint32_t smCount,
	cudaStream_t stream, uint32_t format )
{
	if ( numRays == 0 ) return;
	const bool any = type == hiprtTraversalTerminateAtAnyHit;
	const int  variant = traceVariant();
	if ( triangles && ( variant != 1 || format != hiprtB200HitFormatFull ) )
	{
		const bool hints = variant != 2;
		switch ( format )
		{
		case hiprtB200HitFormatCompact:
			if ( hints )
				launchTriangles<hiprtB200HitFormatCompact, true>( geom, rays, numRays, hits, any, workCounter, smCount, stream );
			else
				launchTriangles<hiprtB200HitFormatCompact, false>( geom, rays, numRays, hits, any, workCounter, smCount, stream );
			return;
		case hiprtB200HitFormatPrimID:
			if ( hints )
				launchTriangles<hiprtB200HitFormatPrimID, true>( geom, rays, numRays, hits, any, workCounter, smCount, stream );
			else
				launchTriangles<hiprtB200HitFormatPrimID, false>( geom, rays, numRays, hits, any, workCounter, smCount, stream );
			return;
		default:
			if ( hints )
				launchTriangles<hiprtB200HitFormatFull, true>( geom, rays, numRays, hits, any, workCounter, smCount, stream );
			else
				launchTriangles<hiprtB200HitFormatFull, false>( geom, rays, numRays, hits, any, workCounter, smCount, stream );
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
	float			time )
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
	hiprtHit hit;
	if ( bestRef != InvalidValue && !overflow )
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
__global__ void __launch_bounds__( Block ) traceSceneFastKernel(
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
			reinterpret_cast<const SceneHeader*>( scene ), rays, hits, base + lane, base + lane < numRays, mask, stackBase, rayTime );
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
	cudaMemsetAsync( workCounter, 0, sizeof( uint32_t ), stream );
	const uint32_t blocks = std::min( DivideRoundUp( numRays, Block ), smCount * uint32_t( std::max( blocksPerSm, 1 ) ) );
	kernel<<<blocks, Block, 0, stream>>>( scene, rays, numRays, hits, mask, workCounter, times, time );
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
