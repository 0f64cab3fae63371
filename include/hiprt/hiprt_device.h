// hiprt_b200 -- device-side API: traversal stacks and traversal objects that a user's own
// CUDA kernel instantiates per thread. Same class names, constructor signatures and
// getNextHit()/getCurrentState() protocol as the reference's hiprt/hiprt_device.h:38-583;
// the implementation (reference: hiprt/impl/hiprt_device_impl.h) is written for sm_100a:
// header-only (no pimpl indirection, so ptxas sees straight-line code), 128-bit read-only
// node fetches (ld.global.nc.v4), IEEE reciprocal/division exactly as the reference's
// NVIDIA path, and an order-independent closest-hit rule (see acceptCloser below).
//
// A translation unit that includes this header must provide the two dispatchers
// `intersectFunc` / `filterFunc` declared below. hiprtBuildTraceKernels generates them from
// the hiprtFuncNameSet table (reference: hiprt/impl/Compiler.cpp:573-625); code compiled
// offline can `#define HIPRT_B200_DEFAULT_FUNC_TABLE` before including to get empty ones.
#pragma once

#if !defined( __CUDACC__ )
#error "hiprt_device.h is device code: compile with nvcc or NVRTC for sm_100a"
#endif

#include <hiprt/hiprt_common.h>
#include <hiprt/hiprt_types.h>
#include <hiprt/hiprt_vec.h>
#include <hiprt/hiprt_math.h> // the reference injects its vector arithmetic into every trace-kernel compilation: user kernels call hiprt::dot etc.
#include <hiprt/hiprt_b200_layout.h>
#include <hiprt/hiprt_b200_frame.h>

#ifndef HIPRT_SHARED_STACK_STRIDE
#define HIPRT_SHARED_STACK_STRIDE 32
#endif

HIPRT_DEVICE bool intersectFunc(
	uint32_t					geomType,
	uint32_t					rayType,
	const hiprtFuncTableHeader& tableHeader,
	const hiprtRay&				ray,
	void*						payload,
	hiprtHit&					hit );
HIPRT_DEVICE bool filterFunc(
	uint32_t					geomType,
	uint32_t					rayType,
	const hiprtFuncTableHeader& tableHeader,
	const hiprtRay&				ray,
	void*						payload,
	const hiprtHit&				hit );

#if defined( HIPRT_B200_DEFAULT_FUNC_TABLE )
HIPRT_DEVICE bool
intersectFunc( uint32_t, uint32_t, const hiprtFuncTableHeader&, const hiprtRay&, void*, hiprtHit& )
{
	return false;
}
HIPRT_DEVICE bool
filterFunc( uint32_t, uint32_t, const hiprtFuncTableHeader&, const hiprtRay&, void*, const hiprtHit& )
{
	return false;
}
#endif

namespace hiprt
{
// ------------------------------------------------------------------------------------------------
// float3 arithmetic with the exact operation order of the reference (hiprt_math.h:1571-1603):
// compensated cross/dot built from explicit fmaf, so results do not depend on -fmad.
HIPRT_DEVICE HIPRT_INLINE float3 sub3( const float3 a, const float3 b ) { return make_float3( a.x - b.x, a.y - b.y, a.z - b.z ); }
HIPRT_DEVICE HIPRT_INLINE float3 mul3( const float3 a, const float3 b ) { return make_float3( a.x * b.x, a.y * b.y, a.z * b.z ); }
// (the traversal's own copies, with explicit roundings: hiprt::sumOfProducts / differenceOfProducts of hiprt_math.h are the
// user-facing ones and follow the translation unit's contraction setting)
HIPRT_DEVICE HIPRT_INLINE float	 sumOfProductsRn( float a, float b, float c, float d )
{
	const float cd = __fmul_rn( c, d );
	return __fadd_rn( fmaf( a, b, cd ), fmaf( c, d, -cd ) );
}
HIPRT_DEVICE HIPRT_INLINE float differenceOfProductsRn( float a, float b, float c, float d )
{
	const float cd = __fmul_rn( c, d );
	return __fsub_rn( fmaf( a, b, -cd ), fmaf( c, d, -cd ) );
}
HIPRT_DEVICE HIPRT_INLINE float3 cross3( const float3 a, const float3 b )
{
	return make_float3(
		differenceOfProductsRn( a.y, b.z, a.z, b.y ),
		differenceOfProductsRn( a.z, b.x, a.x, b.z ),
		differenceOfProductsRn( a.x, b.y, a.y, b.x ) );
}
HIPRT_DEVICE HIPRT_INLINE float	 dot3( const float3 a, const float3 b ) { return fmaf( a.x, b.x, sumOfProductsRn( a.y, b.y, a.z, b.z ) ); }
HIPRT_DEVICE HIPRT_INLINE float3 rcp3( const float3 a )
{
	return make_float3( __fdiv_rn( 1.0f, a.x ), __fdiv_rn( 1.0f, a.y ), __fdiv_rn( 1.0f, a.z ) );
}

// 16-byte read-only loads: every node is 64-byte aligned, so a Box4Node is 7 and a triangle
// pair 4 of these.
HIPRT_DEVICE HIPRT_INLINE float4 ldg4( const void* p, uint32_t i ) { return __ldg( reinterpret_cast<const float4*>( p ) + i ); }

struct HitTriangle
{
	float t, u, v;
};

// Moller-Trumbore with the reference's arithmetic and rejection tests (Triangle.h:96-130).
// Returns barycentrics (u,v) of (v1,v2); the caller maps them through the pair flags.
HIPRT_DEVICE HIPRT_INLINE bool
intersectTriangle( const float3 v0, const float3 v1, const float3 v2, const hiprtRay& ray, HitTriangle& out )
{
	const float3 e1	   = sub3( v1, v0 );
	const float3 e2	   = sub3( v2, v0 );
	const float3 s1	   = cross3( ray.direction, e2 );
	const float	 denom = dot3( s1, e1 );
	if ( denom == 0.0f ) return false;
	const float	 inv = __fdiv_rn( 1.0f, denom );
	const float3 d	 = sub3( ray.origin, v0 );
	const float	 u	 = __fmul_rn( dot3( d, s1 ), inv );
	const float3 s2	 = cross3( d, e1 );
	const float	 v	 = __fmul_rn( dot3( ray.direction, s2 ), inv );
	const float	 t	 = __fmul_rn( dot3( e2, s2 ), inv );
	if ( u < 0.0f || u > 1.0f || v < 0.0f || __fadd_rn( u, v ) > 1.0f || t < ray.minT || t > ray.maxT ) return false;
	out.t = t, out.u = u, out.v = v;
	return true;
}

// One triangle of a 64-byte pair leaf. `which` selects (v0,v1,v2) or (v1,v3,v2).
HIPRT_DEVICE HIPRT_INLINE bool
intersectPairTriangle( const TrianglePairNode* leaves, uint32_t leafAddr, uint32_t which, const hiprtRay& ray, hiprtHit& hit )
{
	const void*	 p	= leaves + leafAddr;
	const float4 q0 = ldg4( p, 0 ), q1 = ldg4( p, 1 ), q2 = ldg4( p, 2 ), q3 = ldg4( p, 3 );
	const uint32_t prim0 = __float_as_uint( q3.y ), prim1 = __float_as_uint( q3.z ), flagsAll = __float_as_uint( q3.w );
	if ( which && prim0 == prim1 ) return false;
	const float3 pv0 = make_float3( q0.x, q0.y, q0.z ), pv1 = make_float3( q0.w, q1.x, q1.y );
	const float3 pv2 = make_float3( q1.z, q1.w, q2.x ), pv3 = make_float3( q2.y, q2.z, q2.w );
	const float3 a = which ? pv1 : pv0, b = which ? pv3 : pv1, c = pv2;
	const uint32_t flags = flagsAll >> ( which * 8 );
	HitTriangle	   h;
	if ( !intersectTriangle( a, b, c, ray, h ) ) return false;
	const float w  = __fsub_rn( __fsub_rn( 1.0f, h.u ), h.v );
	const uint32_t sx = flags & 3u, sy = ( flags >> 2 ) & 3u;
	hit.uv.x   = sx == 0 ? w : ( sx == 1 ? h.u : h.v );
	hit.uv.y   = sy == 0 ? w : ( sy == 1 ? h.u : h.v );
	hit.t	   = h.t;
	hit.primID = which ? prim1 : prim0;
	hit.normal = ( ( flags >> 5 ) & 1u ) ? cross3( sub3( c, a ), sub3( b, a ) ) : cross3( sub3( b, a ), sub3( c, a ) );
	return true;
}

// Closest-hit acceptance. The reference accepts any later-visited hit with t <= maxT, which
// makes exact ties depend on traversal order (and, through its warp vote, on neighbouring
// rays). Here the winner is the minimum of (t, instanceID, primID): independent of visiting order.
HIPRT_DEVICE HIPRT_INLINE bool acceptCloser( const hiprtHit& cand, const hiprtHit& best )
{
	if ( !best.hasHit() || cand.t < best.t ) return true;
	if ( cand.t > best.t ) return false;
	if ( cand.instanceID != best.instanceID ) return cand.instanceID < best.instanceID;
	return cand.primID < best.primID;
}

// Slab test of the four children of one box node + ordering (hiprt_device_impl.h:338-411,
// Aabb.h:92-101). out[0..3]: hit children nearest first, InvalidValue for misses.
HIPRT_DEVICE HIPRT_INLINE void
intersectBox4( const Box4Node* nodes, uint32_t addr, const float3 o, const float3 invD, float maxT, uint32_t out[4] )
{
	const void*	 p	= nodes + addr;
	const float4 q0 = ldg4( p, 0 ), q1 = ldg4( p, 1 ), q2 = ldg4( p, 2 ), q3 = ldg4( p, 3 );
	const float4 q4 = ldg4( p, 4 ), q5 = ldg4( p, 5 ), q6 = ldg4( p, 6 );
	const float	 b[24] = { q1.x, q1.y, q1.z, q1.w, q2.x, q2.y, q2.z, q2.w, q3.x, q3.y, q3.z, q3.w,
						   q4.x, q4.y, q4.z, q4.w, q5.x, q5.y, q5.z, q5.w, q6.x, q6.y, q6.z, q6.w };
	const uint32_t c[4] = { __float_as_uint( q0.x ), __float_as_uint( q0.y ), __float_as_uint( q0.z ), __float_as_uint( q0.w ) };
	float		   d[4];
#pragma unroll
	for ( int s = 0; s < 4; ++s )
	{
		const float fx = __fmul_rn( __fsub_rn( b[6 * s + 3], o.x ), invD.x ), nx = __fmul_rn( __fsub_rn( b[6 * s + 0], o.x ), invD.x );
		const float fy = __fmul_rn( __fsub_rn( b[6 * s + 4], o.y ), invD.y ), ny = __fmul_rn( __fsub_rn( b[6 * s + 1], o.y ), invD.y );
		const float fz = __fmul_rn( __fsub_rn( b[6 * s + 5], o.z ), invD.z ), nz = __fmul_rn( __fsub_rn( b[6 * s + 2], o.z ), invD.z );
		const float t1 = fminf( fminf( fminf( fmaxf( fx, nx ), fmaxf( fy, ny ) ), fmaxf( fz, nz ) ), maxT );
		const float t0 = fmaxf( fmaxf( fmaxf( fminf( fx, nx ), fminf( fy, ny ) ), fminf( fz, nz ) ), 0.0f );
		d[s]		   = t0;
		out[s]		   = ( t0 <= t1 ) ? c[s] : InvalidValue;
	}
#define HIPRT_B200_CSWAP( A, B )                                                  \
	if ( ( out[B] != InvalidValue && d[B] < d[A] ) || out[A] == InvalidValue )    \
	{                                                                             \
		const float	   td = d[A];                                                 \
		const uint32_t ti = out[A];                                               \
		out[A] = out[B], d[A] = d[B], out[B] = ti, d[B] = td;                     \
	}
	HIPRT_B200_CSWAP( 0, 2 )
	HIPRT_B200_CSWAP( 1, 3 )
	HIPRT_B200_CSWAP( 0, 1 )
	HIPRT_B200_CSWAP( 2, 3 )
	HIPRT_B200_CSWAP( 1, 2 )
#undef HIPRT_B200_CSWAP
}

// world -> object for a static instance (BvhNode.h:1111-1126): direction is not re-normalised,
// so t is preserved across the instance boundary.
HIPRT_DEVICE HIPRT_INLINE hiprtRay transformRayToObject( const InstanceNode* node, const hiprtRay& ray )
{
	const float4 r0 = ldg4( node, 0 ), r1 = ldg4( node, 1 ), r2 = ldg4( node, 2 );
	hiprtRay	 out;
	const float3 o = ray.origin, d = ray.direction;
	out.origin.x	= __fadd_rn( dot3( make_float3( r0.x, r0.y, r0.z ), o ), r0.w );
	out.origin.y	= __fadd_rn( dot3( make_float3( r1.x, r1.y, r1.z ), o ), r1.w );
	out.origin.z	= __fadd_rn( dot3( make_float3( r2.x, r2.y, r2.z ), o ), r2.w );
	out.direction.x = dot3( make_float3( r0.x, r0.y, r0.z ), d );
	out.direction.y = dot3( make_float3( r1.x, r1.y, r1.z ), d );
	out.direction.z = dot3( make_float3( r2.x, r2.y, r2.z ), d );
	out.minT		= ray.minT;
	out.maxT		= ray.maxT;
	return out;
}

} // namespace hiprt

// ================================================================================================
// Stacks (reference: hiprt_device.h:48-173, hiprt_device_impl.h:86-247)

class hiprtEmptyInstanceStack
{
};

namespace hiprt
{
// (NVRTC has neither <type_traits> nor the __is_same builtin)
template <typename A, typename B>
struct IsSame
{
	static constexpr bool value = false;
};
template <typename A>
struct IsSame<A, A>
{
	static constexpr bool value = true;
};
} // namespace hiprt

// Per-thread stack in local memory.
template <typename Entry, uint32_t Size>
class hiprtPrivateStackT
{
  public:
	static constexpr uint32_t StackSize = Size;
	HIPRT_DEVICE			  hiprtPrivateStackT() : m_top( 0u ) {}
	HIPRT_DEVICE Entry		  pop() { return m_data[--m_top]; }
	HIPRT_DEVICE void		  push( Entry e ) { m_data[m_top++] = e; }
	HIPRT_DEVICE bool		  empty() const { return m_top == 0u; }
	HIPRT_DEVICE uint32_t	  vacancy() const { return Size - m_top; }
	HIPRT_DEVICE void		  reset() { m_top = 0u; }

  private:
	Entry	 m_data[Size];
	uint32_t m_top;
};
using hiprtPrivateStack			= hiprtPrivateStackT<uint32_t, 64u>;
using hiprtPrivateInstanceStack = hiprtPrivateStackT<hiprtInstanceStackEntry, hiprtMaxInstanceLevels - 1>;

// Shared-memory ring (lane-interleaved, stride 32 entries) that spills its oldest entry to a
// per-thread region of a global buffer. Dynamic = warps claim a global region by CAS on a lock
// table sized for the resident warps (hiprtCreateGlobalStackBuffer, hiprtStackTypeDynamic).
template <typename Entry, bool Dynamic>
class hiprtGlobalStackT
{
  public:
	HIPRT_DEVICE hiprtGlobalStackT( hiprtGlobalStackBuffer globalStackBuffer, hiprtSharedStackBuffer sharedStackBuffer )
	{
		const uint32_t tid	= threadIdx.x + threadIdx.y * blockDim.x;
		const uint32_t warp = tid >> 5, lane = tid & 31u;
		m_shared	 = reinterpret_cast<Entry*>( sharedStackBuffer.stackData ) + lane + warp * 32u * sharedStackBuffer.stackSize;
		m_sharedSize = sharedStackBuffer.stackSize;
		m_globalSize = globalStackBuffer.stackSize;
		m_lock		 = nullptr;
		if constexpr ( Dynamic )
		{
			const uint32_t warpsPerBlock = ( blockDim.x * blockDim.y + 31u ) >> 5;
			const uint32_t slots		 = globalStackBuffer.stackCount >> 5;
			uint32_t*	   locks		 = reinterpret_cast<uint32_t*>( globalStackBuffer.stackData );
			const uint32_t active		 = __activemask();
			const uint32_t leader		 = __ffs( active ) - 1;
			uint32_t	   slot			 = InvalidSlot;
			uint32_t	   cand			 = ( warp + ( blockIdx.x + blockIdx.y * gridDim.x ) * warpsPerBlock ) % slots;
			while ( slot == InvalidSlot )
			{
				if ( lane == leader && atomicCAS( &locks[cand], 0u, 1u ) == 0u ) slot = cand;
				cand = ( cand + 1 ) % slots;
				slot = __shfl_sync( active, slot, leader );
			}
			__threadfence();
			m_lock	 = &locks[slot];
			m_global = reinterpret_cast<Entry*>( locks + slots ) + lane + size_t( slot ) * 32u * m_globalSize;
		}
		else
		{
			const uint32_t block = blockIdx.x + blockIdx.y * gridDim.x;
			m_global = reinterpret_cast<Entry*>( globalStackBuffer.stackData ) + lane +
					   ( size_t( warp ) * 32u + size_t( block ) * ( blockDim.x * blockDim.y ) ) * m_globalSize;
		}
		reset();
	}

	HIPRT_DEVICE ~hiprtGlobalStackT()
	{
		if constexpr ( Dynamic )
		{
			__threadfence();
			const uint32_t lane = ( threadIdx.x + threadIdx.y * blockDim.x ) & 31u;
			if ( lane == __ffs( __activemask() ) - 1 ) atomicExch( m_lock, 0u );
		}
	}

	HIPRT_DEVICE Entry pop()
	{
		if ( m_sharedCount > 0 )
		{
			--m_sharedCount;
			m_sharedHead = m_sharedHead == 0 ? m_sharedSize - 1 : m_sharedHead - 1;
			return m_shared[m_sharedHead * 32u];
		}
		return m_global[size_t( --m_globalCount ) * 32u];
	}

	HIPRT_DEVICE void push( Entry e )
	{
		if ( m_sharedSize == 0 )
		{
			m_global[size_t( m_globalCount++ ) * 32u] = e;
			return;
		}
		if ( m_sharedCount == m_sharedSize )
			m_global[size_t( m_globalCount++ ) * 32u] = m_shared[m_sharedHead * 32u]; // evict the oldest
		else
			++m_sharedCount;
		m_shared[m_sharedHead * 32u] = e;
		m_sharedHead				 = m_sharedHead + 1 == m_sharedSize ? 0 : m_sharedHead + 1;
	}

	HIPRT_DEVICE uint32_t vacancy() const { return m_globalSize - m_globalCount + m_sharedSize - m_sharedCount; }
	HIPRT_DEVICE bool	  empty() const { return m_sharedCount == 0 && m_globalCount == 0; }
	HIPRT_DEVICE void	  reset() { m_sharedHead = m_sharedCount = m_globalCount = 0; }

  private:
	static constexpr uint32_t InvalidSlot = 0xFFFFFFFFu;
	uint32_t*				  m_lock;
	Entry*					  m_global;
	Entry*					  m_shared;
	uint32_t				  m_globalSize, m_sharedSize;
	uint32_t				  m_sharedHead, m_sharedCount, m_globalCount;
};
using hiprtGlobalStack			= hiprtGlobalStackT<uint32_t, false>;
using hiprtGlobalInstanceStack	= hiprtGlobalStackT<hiprtInstanceStackEntry, false>;
using hiprtDynamicStack			= hiprtGlobalStackT<uint32_t, true>;
using hiprtDynamicInstanceStack = hiprtGlobalStackT<hiprtInstanceStackEntry, true>;

// ================================================================================================
// Geometry traversal (reference: hiprt_device_impl.h:700-850)

template <typename Stack, hiprtPrimitiveNodeType PrimType, hiprtTraversalType Type>
class hiprtGeomTraversalT
{
  public:
	HIPRT_DEVICE hiprtGeomTraversalT(
		hiprtGeometry	   geom,
		const hiprtRay&	   ray,
		Stack&			   stack,
		hiprtTraversalHint hint		 = hiprtTraversalHintDefault,
		void*			   payload	 = nullptr,
		hiprtFuncTable	   funcTable = nullptr,
		uint32_t		   rayType	 = 0 )
		: m_ray( ray ), m_stack( stack ), m_payload( payload ), m_rayType( rayType )
	{
		(void)hint;
		const hiprt::GeomHeader* h = reinterpret_cast<const hiprt::GeomHeader*>( geom );
		m_boxNodes				   = h->boxNodes;
		m_primNodes				   = h->primNodes;
		m_geomType				   = h->geomType;
		m_table					   = hiprtFuncTableHeader{ 0, 1, nullptr };
		if ( funcTable != nullptr ) m_table = *reinterpret_cast<const hiprtFuncTableHeader*>( funcTable );
		m_stack.reset();
	}

	HIPRT_DEVICE hiprtTraversalState getCurrentState() { return m_state; }

	HIPRT_DEVICE hiprtHit getNextHit()
	{
		using namespace hiprt;
		hiprtRay	 ray  = m_ray;
		const float3 invD = rcp3( ray.direction );
		hiprtHit	 best;
		if ( m_stack.empty() ) m_stack.push( InvalidValue );
		const bool useFilter = m_geomType != InvalidValue && m_table.funcDataSets != nullptr;

		while ( m_node != InvalidValue )
		{
			if ( isBoxIndex( m_node ) )
			{
				uint32_t next[4];
				intersectBox4( m_boxNodes, nodeAddr( m_node ), ray.origin, invD, ray.maxT, next );
				if ( m_stack.vacancy() < 3 )
				{
					m_state = hiprtTraversalStateStackOverflow;
					return hiprtHit();
				}
				if ( next[3] != InvalidValue ) m_stack.push( next[3] );
				if ( next[2] != InvalidValue ) m_stack.push( next[2] );
				if ( next[1] != InvalidValue ) m_stack.push( next[1] );
				m_node = next[0] != InvalidValue ? next[0] : m_stack.pop();
				continue;
			}

			// leaf
			if constexpr ( PrimType == hiprtTriangleNode )
			{
				const TrianglePairNode* leaves = reinterpret_cast<const TrianglePairNode*>( m_primNodes );
				const uint32_t			addr   = nodeAddr( m_node );
				if constexpr ( Type == hiprtTraversalTerminateAtAnyHit )
				{
					// re-entrant: m_phase remembers which triangle of the pair was already reported
					while ( m_phase < 2 )
					{
						hiprtHit   hit;
						const bool found = intersectPairTriangle( leaves, addr, m_phase, ray, hit ) &&
										   !( useFilter && filterFunc( m_geomType >> 1, m_rayType, m_table, ray, m_payload, hit ) );
						++m_phase;
						if ( found )
						{
							if ( m_phase == 2 )
							{
								m_phase = 0;
								m_node	= m_stack.pop();
							}
							m_state = hiprtTraversalStateHit;
							return hit;
						}
					}
					m_phase = 0;
				}
				else
				{
#pragma unroll
					for ( uint32_t which = 0; which < 2; ++which )
					{
						hiprtHit hit;
						if ( intersectPairTriangle( leaves, addr, which, ray, hit ) &&
							 !( useFilter && filterFunc( m_geomType >> 1, m_rayType, m_table, ray, m_payload, hit ) ) &&
							 acceptCloser( hit, best ) )
						{
							best	 = hit;
							ray.maxT = hit.t;
						}
					}
				}
			}
			else
			{
				const CustomNode* leaves = reinterpret_cast<const CustomNode*>( m_primNodes );
				hiprtHit		  hit;
				hit.primID = leaves[nodeAddr( m_node )].primIndex;
				bool found = intersectFunc( m_geomType >> 1, m_rayType, m_table, ray, m_payload, hit );
				if ( found && useFilter && filterFunc( m_geomType >> 1, m_rayType, m_table, ray, m_payload, hit ) ) found = false;
				if ( found )
				{
					if constexpr ( Type == hiprtTraversalTerminateAtAnyHit )
					{
						m_node	= m_stack.pop();
						m_state = hiprtTraversalStateHit;
						return hit;
					}
					else
					{
						best	 = hit;
						ray.maxT = hit.t;
					}
				}
			}
			m_node = m_stack.pop();
		}
		m_state = hiprtTraversalStateFinished;
		return best;
	}

  private:
	hiprtRay				m_ray;
	Stack&					m_stack;
	const hiprt::Box4Node*	m_boxNodes;
	const void*				m_primNodes;
	void*					m_payload;
	hiprtFuncTableHeader	m_table;
	uint32_t				m_geomType;
	uint32_t				m_rayType;
	uint32_t				m_node	= hiprt::RootIndex;
	uint32_t				m_phase = 0;
	hiprtTraversalState		m_state = hiprtTraversalStateInit;
};

// ================================================================================================
// Scene traversal, single- and multi-level (reference: hiprt_device_impl.h:852-1204)

template <typename Stack, typename InstanceStack, hiprtTraversalType Type>
class hiprtSceneTraversalT
{
	static constexpr bool MultiLevel = !hiprt::IsSame<InstanceStack, hiprtEmptyInstanceStack>::value;

  public:
	HIPRT_DEVICE hiprtSceneTraversalT(
		hiprtScene		   scene,
		const hiprtRay&	   ray,
		Stack&			   stack,
		InstanceStack&	   instanceStack,
		hiprtRayMask	   mask		 = hiprtFullRayMask,
		hiprtTraversalHint hint		 = hiprtTraversalHintDefault,
		void*			   payload	 = nullptr,
		hiprtFuncTable	   funcTable = nullptr,
		uint32_t		   rayType	 = 0,
		float			   time		 = 0.0f )
		: m_ray( ray ), m_stack( stack ), m_instanceStack( instanceStack ), m_payload( payload ), m_mask( mask ),
		  m_rayType( rayType ), m_time( time )
	{
		(void)hint;
		m_scene = reinterpret_cast<const hiprt::SceneHeader*>( scene );
		m_table = hiprtFuncTableHeader{ 0, 1, nullptr };
		if ( funcTable != nullptr ) m_table = *reinterpret_cast<const hiprtFuncTableHeader*>( funcTable );
		m_stack.reset();
		if constexpr ( MultiLevel ) m_instanceStack.reset();
#pragma unroll
		for ( uint32_t i = 0; i < hiprtMaxInstanceLevels; ++i )
			m_ids[i] = hiprt::InvalidValue;
	}

	HIPRT_DEVICE hiprtTraversalState getCurrentState() { return m_state; }

	HIPRT_DEVICE hiprtHit getNextHit()
	{
		using namespace hiprt;
		// state that lives in registers for the duration of the call
		const Box4Node*		boxNodes  = m_scene->boxNodes;
		const InstanceNode* instNodes = m_scene->primNodes;
		const void*			primNodes = nullptr;
		uint32_t			geomType  = InvalidValue;
		hiprtRay			ray		  = m_ray;
		if ( m_instance != InvalidValue )
		{
			// re-entry of the any-hit iterator while inside an instance
			const InstanceNode* in = instNodes + nodeAddr( m_instance );
			enterGeometry( in, ray, boxNodes, primNodes, geomType );
		}
		float3	 invD = rcp3( ray.direction );
		hiprtHit best;
		if ( m_stack.empty() ) m_stack.push( InvalidValue );

		while ( m_node != InvalidValue )
		{
			if ( isBoxIndex( m_node ) )
			{
				uint32_t next[4];
				intersectBox4( boxNodes, nodeAddr( m_node ), ray.origin, invD, ray.maxT, next );
				if ( m_stack.vacancy() < 3 )
				{
					m_state = hiprtTraversalStateStackOverflow;
					return hiprtHit();
				}
				if ( next[3] != InvalidValue ) m_stack.push( next[3] );
				if ( next[2] != InvalidValue ) m_stack.push( next[2] );
				if ( next[1] != InvalidValue ) m_stack.push( next[1] );
				m_node = next[0] != InvalidValue ? next[0] : popNode( ray, invD, boxNodes, instNodes );
				continue;
			}

			if ( m_instance == InvalidValue )
			{
				// instance leaf of the current scene level
				const InstanceNode* in = instNodes + nodeAddr( m_node );
				const uint4			meta = __ldg( reinterpret_cast<const uint4*>( in ) + 3 ); // {ptr.lo, ptr.hi, mask, bits}
				if ( ( meta.z & m_mask ) != 0u )
				{
					if ( m_stack.vacancy() < 1 )
					{
						m_state = hiprtTraversalStateStackOverflow;
						return hiprtHit();
					}
					const bool isScene = ( meta.w >> 24 ) & 1u;
					if constexpr ( MultiLevel )
					{
						if ( isScene )
						{
							hiprtRay inner = ray;
							if ( !( ( meta.w >> 26 ) & 1u ) ) inner = transformInto( in, ray );
							m_ids[m_level] = meta.w & 0xFFFFFFu;
							m_instanceStack.push( hiprtInstanceStackEntry{ m_ray, (hiprtScene)m_scene } );
							m_stack.push( InvalidValue );
							m_ray	= inner;
							m_ray.maxT = ray.maxT;
							m_scene = in->scene;
							++m_level;
							ray		  = m_ray;
							invD	  = rcp3( ray.direction );
							boxNodes  = m_scene->boxNodes;
							instNodes = m_scene->primNodes;
							m_node	  = RootIndex;
							continue;
						}
					}
					(void)isScene;
					m_instance	   = m_node;
					m_ids[m_level] = meta.w & 0xFFFFFFu;
					m_stack.push( InvalidValue );
					const float maxT = ray.maxT;
					enterGeometry( in, ray, boxNodes, primNodes, geomType );
					ray.maxT = maxT;
					invD	 = rcp3( ray.direction );
					m_node	 = RootIndex;
					continue;
				}
				m_node = popNode( ray, invD, boxNodes, instNodes );
				continue;
			}

			// primitive leaf inside an instance
			const bool useFilter = geomType != InvalidValue && m_table.funcDataSets != nullptr;
			if ( geomType & 1u )
			{
				const TrianglePairNode* leaves = reinterpret_cast<const TrianglePairNode*>( primNodes );
				const uint32_t			addr   = nodeAddr( m_node );
				if constexpr ( Type == hiprtTraversalTerminateAtAnyHit )
				{
					while ( m_phase < 2 )
					{
						hiprtHit hit;
						setIds( hit );
						const bool found = intersectPairTriangle( leaves, addr, m_phase, ray, hit ) &&
										   !( useFilter && filterFunc( geomType >> 1, m_rayType, m_table, ray, m_payload, hit ) );
						++m_phase;
						if ( found )
						{
							if ( m_phase == 2 )
							{
								m_phase = 0;
								m_node	= popNode( ray, invD, boxNodes, instNodes );
							}
							m_state = hiprtTraversalStateHit;
							return hit;
						}
					}
					m_phase = 0;
				}
				else
				{
#pragma unroll
					for ( uint32_t which = 0; which < 2; ++which )
					{
						hiprtHit hit;
						setIds( hit );
						if ( intersectPairTriangle( leaves, addr, which, ray, hit ) &&
							 !( useFilter && filterFunc( geomType >> 1, m_rayType, m_table, ray, m_payload, hit ) ) &&
							 acceptCloser( hit, best ) )
						{
							best	 = hit;
							ray.maxT = hit.t;
						}
					}
				}
			}
			else
			{
				const CustomNode* leaves = reinterpret_cast<const CustomNode*>( primNodes );
				hiprtHit		  hit;
				setIds( hit );
				hit.primID = leaves[nodeAddr( m_node )].primIndex;
				bool found = intersectFunc( geomType >> 1, m_rayType, m_table, ray, m_payload, hit );
				if ( found && useFilter && filterFunc( geomType >> 1, m_rayType, m_table, ray, m_payload, hit ) ) found = false;
				if ( found )
				{
					if constexpr ( Type == hiprtTraversalTerminateAtAnyHit )
					{
						m_node	= popNode( ray, invD, boxNodes, instNodes );
						m_state = hiprtTraversalStateHit;
						return hit;
					}
					else
					{
						best	 = hit;
						ray.maxT = hit.t;
					}
				}
			}
			m_node = popNode( ray, invD, boxNodes, instNodes );
		}
		m_state = hiprtTraversalStateFinished;
		return best;
	}

  private:
	HIPRT_DEVICE void setIds( hiprtHit& hit ) const
	{
		if constexpr ( MultiLevel )
		{
#pragma unroll
			for ( uint32_t i = 0; i < hiprtMaxInstanceLevels; ++i )
				hit.instanceIDs[i] = i <= m_level ? m_ids[i] : hiprt::InvalidValue;
		}
		else
			hit.instanceID = m_ids[0];
	}

	// world/parent-space ray -> instance space (static matrix or motion-blur frames)
	HIPRT_DEVICE hiprtRay transformInto( const hiprt::InstanceNode* in, const hiprtRay& ray ) const
	{
		const uint32_t bits = __ldg( reinterpret_cast<const uint32_t*>( in ) + 15 );
		if ( ( bits >> 25 ) & 1u ) return hiprt::transformRayToObject( in, ray );
		// animated instance (hiprt_device_impl.h:989-993): the node holds the key-frame range instead of a matrix; the pose
		// at the ray's time is interpolated and origin / origin + direction are carried through its inverse
		const uint2				   range = __ldg( reinterpret_cast<const uint2*>( in ) ); // hiprtTransformHeader {frameIndex, frameCount}
		const hiprt::motion::Frame f	 = hiprt::motion::frameAt( m_scene->frames, range.x, range.y, m_time );
		return hiprt::motion::transformRay( f, ray );
	}

	HIPRT_DEVICE void enterGeometry(
		const hiprt::InstanceNode* in, hiprtRay& ray, const hiprt::Box4Node*& boxNodes, const void*& primNodes, uint32_t& geomType ) const
	{
		const uint32_t bits = __ldg( reinterpret_cast<const uint32_t*>( in ) + 15 );
		if ( !( ( bits >> 26 ) & 1u ) ) ray = transformInto( in, ray );
		const hiprt::GeomHeader* g = in->geometry;
		boxNodes				   = g->boxNodes;
		primNodes				   = g->primNodes;
		geomType				   = g->geomType;
	}

	// Pop the next node; a sentinel means "leave the current instance (or nested scene)".
	HIPRT_DEVICE uint32_t popNode( hiprtRay& ray, float3& invD, const hiprt::Box4Node*& boxNodes, const hiprt::InstanceNode*& instNodes )
	{
		uint32_t node = m_stack.pop();
		while ( node == hiprt::InvalidValue && !m_stack.empty() )
		{
			const float maxT = ray.maxT;
			if constexpr ( MultiLevel )
			{
				if ( m_instance == hiprt::InvalidValue )
				{
					const hiprtInstanceStackEntry e = m_instanceStack.pop();
					m_ray							= e.ray;
					m_scene							= reinterpret_cast<const hiprt::SceneHeader*>( e.scene );
					--m_level;
				}
			}
			m_instance	   = hiprt::InvalidValue;
			m_ids[m_level] = hiprt::InvalidValue;
			ray			   = m_ray;
			ray.maxT	   = maxT;
			invD		   = hiprt::rcp3( ray.direction );
			boxNodes	   = m_scene->boxNodes;
			instNodes	   = m_scene->primNodes;
			node		   = m_stack.pop();
		}
		return node;
	}

	hiprtRay					m_ray;
	Stack&						m_stack;
	InstanceStack&				m_instanceStack;
	const hiprt::SceneHeader*	m_scene;
	void*						m_payload;
	hiprtFuncTableHeader		m_table;
	hiprtRayMask				m_mask;
	uint32_t					m_rayType;
	float						m_time;
	uint32_t					m_ids[hiprtMaxInstanceLevels];
	uint32_t					m_level	   = 0;
	uint32_t					m_instance = hiprt::InvalidValue; // instance node index we are inside of
	uint32_t					m_node	   = hiprt::RootIndex;
	uint32_t					m_phase	   = 0;
	hiprtTraversalState			m_state	   = hiprtTraversalStateInit;
};

// ================================================================================================
// Public traversal objects with the reference's names and signatures.

#define HIPRT_B200_GEOM_TRAVERSAL( Name, Prim, Type )                                                                      \
	class Name                                                                                                             \
	{                                                                                                                      \
	  public:                                                                                                              \
		HIPRT_DEVICE Name(                                                                                                 \
			hiprtGeometry	   geom,                                                                                       \
			const hiprtRay&	   ray,                                                                                        \
			hiprtTraversalHint hint		 = hiprtTraversalHintDefault,                                                      \
			void*			   payload	 = nullptr,                                                                        \
			hiprtFuncTable	   funcTable = nullptr,                                                                        \
			uint32_t		   rayType	 = 0 )                                                                             \
			: m_impl( geom, ray, m_stack, hint, payload, funcTable, rayType )                                              \
		{                                                                                                                  \
		}                                                                                                                  \
		HIPRT_DEVICE hiprtHit			 getNextHit() { return m_impl.getNextHit(); }                                      \
		HIPRT_DEVICE hiprtTraversalState getCurrentState() { return m_impl.getCurrentState(); }                            \
                                                                                                                           \
	  private:                                                                                                             \
		hiprtPrivateStack								   m_stack;                                                        \
		hiprtGeomTraversalT<hiprtPrivateStack, Prim, Type> m_impl;                                                         \
	};

HIPRT_B200_GEOM_TRAVERSAL( hiprtGeomTraversalClosest, hiprtTriangleNode, hiprtTraversalTerminateAtClosestHit )
HIPRT_B200_GEOM_TRAVERSAL( hiprtGeomTraversalAnyHit, hiprtTriangleNode, hiprtTraversalTerminateAtAnyHit )
HIPRT_B200_GEOM_TRAVERSAL( hiprtGeomCustomTraversalClosest, hiprtCustomNode, hiprtTraversalTerminateAtClosestHit )
HIPRT_B200_GEOM_TRAVERSAL( hiprtGeomCustomTraversalAnyHit, hiprtCustomNode, hiprtTraversalTerminateAtAnyHit )
#undef HIPRT_B200_GEOM_TRAVERSAL

#define HIPRT_B200_GEOM_TRAVERSAL_CS( Name, Prim, Type )                                                                   \
	template <typename hiprtStack>                                                                                         \
	class Name : public hiprtGeomTraversalT<hiprtStack, Prim, Type>                                                        \
	{                                                                                                                      \
	  public:                                                                                                              \
		HIPRT_DEVICE Name(                                                                                                 \
			hiprtGeometry	   geom,                                                                                       \
			const hiprtRay&	   ray,                                                                                        \
			hiprtStack&		   stack,                                                                                      \
			hiprtTraversalHint hint		 = hiprtTraversalHintDefault,                                                      \
			void*			   payload	 = nullptr,                                                                        \
			hiprtFuncTable	   funcTable = nullptr,                                                                        \
			uint32_t		   rayType	 = 0 )                                                                             \
			: hiprtGeomTraversalT<hiprtStack, Prim, Type>( geom, ray, stack, hint, payload, funcTable, rayType )           \
		{                                                                                                                  \
		}                                                                                                                  \
	};

HIPRT_B200_GEOM_TRAVERSAL_CS( hiprtGeomTraversalClosestCustomStack, hiprtTriangleNode, hiprtTraversalTerminateAtClosestHit )
HIPRT_B200_GEOM_TRAVERSAL_CS( hiprtGeomTraversalAnyHitCustomStack, hiprtTriangleNode, hiprtTraversalTerminateAtAnyHit )
HIPRT_B200_GEOM_TRAVERSAL_CS( hiprtGeomCustomTraversalClosestCustomStack, hiprtCustomNode, hiprtTraversalTerminateAtClosestHit )
HIPRT_B200_GEOM_TRAVERSAL_CS( hiprtGeomCustomTraversalAnyHitCustomStack, hiprtCustomNode, hiprtTraversalTerminateAtAnyHit )
#undef HIPRT_B200_GEOM_TRAVERSAL_CS

#define HIPRT_B200_SCENE_TRAVERSAL( Name, Type )                                                                           \
	class Name                                                                                                             \
	{                                                                                                                      \
	  public:                                                                                                              \
		HIPRT_DEVICE Name(                                                                                                 \
			hiprtScene		   scene,                                                                                      \
			const hiprtRay&	   ray,                                                                                        \
			hiprtRayMask	   mask		 = hiprtFullRayMask,                                                               \
			hiprtTraversalHint hint		 = hiprtTraversalHintDefault,                                                      \
			void*			   payload	 = nullptr,                                                                        \
			hiprtFuncTable	   funcTable = nullptr,                                                                        \
			uint32_t		   rayType	 = 0,                                                                              \
			float			   time		 = 0.0f )                                                                          \
			: m_impl( scene, ray, m_stack, m_instanceStack, mask, hint, payload, funcTable, rayType, time )                \
		{                                                                                                                  \
		}                                                                                                                  \
		HIPRT_DEVICE hiprtHit			 getNextHit() { return m_impl.getNextHit(); }                                      \
		HIPRT_DEVICE hiprtTraversalState getCurrentState() { return m_impl.getCurrentState(); }                            \
                                                                                                                           \
	  private:                                                                                                             \
		hiprtPrivateStack													   m_stack;                                    \
		hiprtPrivateInstanceStack											   m_instanceStack;                            \
		hiprtSceneTraversalT<hiprtPrivateStack, hiprtPrivateInstanceStack, Type> m_impl;                                   \
	};

HIPRT_B200_SCENE_TRAVERSAL( hiprtSceneTraversalClosest, hiprtTraversalTerminateAtClosestHit )
HIPRT_B200_SCENE_TRAVERSAL( hiprtSceneTraversalAnyHit, hiprtTraversalTerminateAtAnyHit )
#undef HIPRT_B200_SCENE_TRAVERSAL

template <typename hiprtStack, typename hiprtInstanceStack>
class hiprtSceneTraversalClosestCustomStack
	: public hiprtSceneTraversalT<hiprtStack, hiprtInstanceStack, hiprtTraversalTerminateAtClosestHit>
{
  public:
	using hiprtSceneTraversalT<hiprtStack, hiprtInstanceStack, hiprtTraversalTerminateAtClosestHit>::hiprtSceneTraversalT;
};

template <typename hiprtStack, typename hiprtInstanceStack>
class hiprtSceneTraversalAnyHitCustomStack
	: public hiprtSceneTraversalT<hiprtStack, hiprtInstanceStack, hiprtTraversalTerminateAtAnyHit>
{
  public:
	using hiprtSceneTraversalT<hiprtStack, hiprtInstanceStack, hiprtTraversalTerminateAtAnyHit>::hiprtSceneTraversalT;
};

// ================================================================================================
// Transform queries (reference: hiprt_device.h:508-583, hiprt_device_impl.h:1775-1927). Points use the frame of the
// instance at `time`; vectors are normal-like (inverse transpose of the scale / shear part).
namespace hiprt
{
HIPRT_DEVICE HIPRT_INLINE motion::Frame instanceFrameAt( const SceneHeader* scene, uint32_t instanceID, float time )
{
	const InstanceRecord rec = scene->instances[instanceID];
	return motion::frameAt( scene->frames, rec.frameIndex, rec.frameCount, time );
}
} // namespace hiprt

HIPRT_DEVICE HIPRT_INLINE float3 hiprtPointObjectToWorld( const float3& point, hiprtScene scene, uint32_t instanceID, float time = 0.0f )
{
	return hiprt::motion::transformPoint( hiprt::instanceFrameAt( reinterpret_cast<const hiprt::SceneHeader*>( scene ), instanceID, time ), point );
}
HIPRT_DEVICE HIPRT_INLINE float3 hiprtPointWorldToObject( const float3& point, hiprtScene scene, uint32_t instanceID, float time = 0.0f )
{
	return hiprt::motion::invTransformPoint( hiprt::instanceFrameAt( reinterpret_cast<const hiprt::SceneHeader*>( scene ), instanceID, time ), point );
}
HIPRT_DEVICE HIPRT_INLINE float3 hiprtVectorObjectToWorld( const float3& vector, hiprtScene scene, uint32_t instanceID, float time = 0.0f )
{
	return hiprt::motion::transformVector( hiprt::instanceFrameAt( reinterpret_cast<const hiprt::SceneHeader*>( scene ), instanceID, time ), vector );
}
HIPRT_DEVICE HIPRT_INLINE float3 hiprtVectorWorldToObject( const float3& vector, hiprtScene scene, uint32_t instanceID, float time = 0.0f )
{
	return hiprt::motion::invTransformVector( hiprt::instanceFrameAt( reinterpret_cast<const hiprt::SceneHeader*>( scene ), instanceID, time ), vector );
}

// Multi-level instancing: instanceIDs[k] selects the instance at level k; the chain ends at the first instance of a geometry.
// Object -> world applies the innermost level first, world -> object the outermost (hiprt_device_impl.h:1819-1927).
namespace hiprt
{
template <bool Point>
HIPRT_DEVICE HIPRT_INLINE float3 objectToWorldLevels( float3 x, hiprtScene scene, const uint32_t ( &instanceIDs )[hiprtMaxInstanceLevels], float time )
{
	const SceneHeader* levels[hiprtMaxInstanceLevels];
	const SceneHeader* header = reinterpret_cast<const SceneHeader*>( scene );
	int32_t			   depth  = 0;
	for ( uint32_t i = 0; i < hiprtMaxInstanceLevels; ++i )
	{
		levels[i]				 = header;
		const InstanceRecord rec = header->instances[instanceIDs[i]];
		++depth;
		if ( rec.type != hiprtInstanceTypeScene ) break;
		header = rec.scene;
	}
	for ( int32_t j = depth - 1; j >= 0; --j )
	{
		const InstanceRecord rec = levels[j]->instances[instanceIDs[j]];
		const motion::Frame	 f	 = motion::frameAt( levels[j]->frames, rec.frameIndex, rec.frameCount, time );
		x						 = Point ? motion::transformPoint( f, x ) : motion::transformVector( f, x );
	}
	return x;
}

template <bool Point>
HIPRT_DEVICE HIPRT_INLINE float3 worldToObjectLevels( float3 x, hiprtScene scene, const uint32_t ( &instanceIDs )[hiprtMaxInstanceLevels], float time )
{
	const SceneHeader* header = reinterpret_cast<const SceneHeader*>( scene );
	for ( uint32_t i = 0; i < hiprtMaxInstanceLevels; ++i )
	{
		const InstanceRecord rec = header->instances[instanceIDs[i]];
		const motion::Frame	 f	 = motion::frameAt( header->frames, rec.frameIndex, rec.frameCount, time );
		x						 = Point ? motion::invTransformPoint( f, x ) : motion::invTransformVector( f, x );
		if ( rec.type != hiprtInstanceTypeScene ) break;
		header = rec.scene;
	}
	return x;
}
} // namespace hiprt

HIPRT_DEVICE HIPRT_INLINE float3
hiprtPointObjectToWorld( const float3& point, hiprtScene scene, const uint32_t ( &instanceIDs )[hiprtMaxInstanceLevels], float time = 0.0f )
{
	return hiprt::objectToWorldLevels<true>( point, scene, instanceIDs, time );
}
HIPRT_DEVICE HIPRT_INLINE float3
hiprtPointWorldToObject( const float3& point, hiprtScene scene, const uint32_t ( &instanceIDs )[hiprtMaxInstanceLevels], float time = 0.0f )
{
	return hiprt::worldToObjectLevels<true>( point, scene, instanceIDs, time );
}
HIPRT_DEVICE HIPRT_INLINE float3
hiprtVectorObjectToWorld( const float3& vector, hiprtScene scene, const uint32_t ( &instanceIDs )[hiprtMaxInstanceLevels], float time = 0.0f )
{
	return hiprt::objectToWorldLevels<false>( vector, scene, instanceIDs, time );
}
HIPRT_DEVICE HIPRT_INLINE float3
hiprtVectorWorldToObject( const float3& vector, hiprtScene scene, const uint32_t ( &instanceIDs )[hiprtMaxInstanceLevels], float time = 0.0f )
{
	return hiprt::worldToObjectLevels<false>( vector, scene, instanceIDs, time );
}
