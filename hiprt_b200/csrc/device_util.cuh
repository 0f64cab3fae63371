// Device-side helpers shared by the build kernels: boxes, warp reductions, Morton codes.
#pragma once
#include <cuda_runtime.h>
#include <hiprt/hiprt_b200_layout.h>

namespace hiprt
{
constexpr uint32_t FullMask = 0xFFFFFFFFu;

__device__ __forceinline__ uint32_t laneId() { return threadIdx.x & 31u; }
__device__ __forceinline__ uint32_t laneMaskLt()
{
	uint32_t m;
	asm( "mov.u32 %0, %%lanemask_lt;" : "=r"( m ) );
	return m;
}

// volatile (L2-coherent) loads for data produced by other CTAs of the same launch
__device__ __forceinline__ uint32_t ldVolatile( const uint32_t* p ) { return *reinterpret_cast<const volatile uint32_t*>( p ); }
__device__ __forceinline__ uint64_t ldVolatile( const uint64_t* p ) { return *reinterpret_cast<const volatile uint64_t*>( p ); }

struct DBox
{
	float3 lo, hi;
};

__device__ __forceinline__ DBox emptyBox()
{
	return DBox{ make_float3( FltMax, FltMax, FltMax ), make_float3( -FltMax, -FltMax, -FltMax ) };
}
__device__ __forceinline__ void growBox( DBox& b, const float3 p )
{
	b.lo = make_float3( fminf( b.lo.x, p.x ), fminf( b.lo.y, p.y ), fminf( b.lo.z, p.z ) );
	b.hi = make_float3( fmaxf( b.hi.x, p.x ), fmaxf( b.hi.y, p.y ), fmaxf( b.hi.z, p.z ) );
}
__device__ __forceinline__ DBox uniteBox( const DBox& a, const DBox& b )
{
	return DBox{
		make_float3( fminf( a.lo.x, b.lo.x ), fminf( a.lo.y, b.lo.y ), fminf( a.lo.z, b.lo.z ) ),
		make_float3( fmaxf( a.hi.x, b.hi.x ), fmaxf( a.hi.y, b.hi.y ), fmaxf( a.hi.z, b.hi.z ) ) };
}
// surface area with the reference's expression 2*(xy + xz + yz) (Aabb.h:71-75), no contraction
__device__ __forceinline__ float boxArea( const DBox& b )
{
	const float ex = __fsub_rn( b.hi.x, b.lo.x ), ey = __fsub_rn( b.hi.y, b.lo.y ), ez = __fsub_rn( b.hi.z, b.lo.z );
	return __fmul_rn( 2.0f, __fadd_rn( __fadd_rn( __fmul_rn( ex, ey ), __fmul_rn( ex, ez ) ), __fmul_rn( ey, ez ) ) );
}

// 32-byte records {lo.xyz, hi.xyz, a, b} (ReferenceNode / ScratchNode) as two 16-byte accesses
struct Rec32
{
	DBox	 box;
	uint32_t a, b;
};
__device__ __forceinline__ Rec32 unpackRec( const float4 q0, const float4 q1 )
{
	Rec32 r;
	r.box.lo = make_float3( q0.x, q0.y, q0.z );
	r.box.hi = make_float3( q0.w, q1.x, q1.y );
	r.a		 = __float_as_uint( q1.z );
	r.b		 = __float_as_uint( q1.w );
	return r;
}
__device__ __forceinline__ Rec32 loadRec( const void* base, size_t index )
{
	const float4* p = reinterpret_cast<const float4*>( base ) + 2 * index;
	return unpackRec( p[0], p[1] );
}
// L2-coherent variant for records written earlier in the same launch by another SM
__device__ __forceinline__ Rec32 loadRecCg( const void* base, size_t index )
{
	const float4* p = reinterpret_cast<const float4*>( base ) + 2 * index;
	return unpackRec( __ldcg( p ), __ldcg( p + 1 ) );
}
__device__ __forceinline__ void storeRec( void* base, size_t index, const DBox& box, uint32_t a, uint32_t b )
{
	float4* p = reinterpret_cast<float4*>( base ) + 2 * index;
	p[0]	  = make_float4( box.lo.x, box.lo.y, box.lo.z, box.hi.x );
	p[1]	  = make_float4( box.hi.y, box.hi.z, __uint_as_float( a ), __uint_as_float( b ) );
}

// float atomics on min/max through the integer ordering trick (reference: hiprt_math.h:1632-1649)
__device__ __forceinline__ void atomicMinF( float* addr, float v )
{
	if ( __float_as_int( v ) >= 0 )
		atomicMin( reinterpret_cast<int*>( addr ), __float_as_int( v ) );
	else
		atomicMax( reinterpret_cast<unsigned int*>( addr ), __float_as_uint( v ) );
}
__device__ __forceinline__ void atomicMaxF( float* addr, float v )
{
	if ( __float_as_int( v ) >= 0 )
		atomicMax( reinterpret_cast<int*>( addr ), __float_as_int( v ) );
	else
		atomicMin( reinterpret_cast<unsigned int*>( addr ), __float_as_uint( v ) );
}

__device__ __forceinline__ DBox warpUniteBox( DBox b )
{
#pragma unroll
	for ( int o = 16; o > 0; o >>= 1 )
	{
		b.lo.x = fminf( b.lo.x, __shfl_xor_sync( FullMask, b.lo.x, o ) );
		b.lo.y = fminf( b.lo.y, __shfl_xor_sync( FullMask, b.lo.y, o ) );
		b.lo.z = fminf( b.lo.z, __shfl_xor_sync( FullMask, b.lo.z, o ) );
		b.hi.x = fmaxf( b.hi.x, __shfl_xor_sync( FullMask, b.hi.x, o ) );
		b.hi.y = fmaxf( b.hi.y, __shfl_xor_sync( FullMask, b.hi.y, o ) );
		b.hi.z = fmaxf( b.hi.z, __shfl_xor_sync( FullMask, b.hi.z, o ) );
	}
	return b;
}

// ---------------------------------------------------------------------------------------------
// 30-bit "extended" Morton code: the axis with the larger extent receives more bits
// (reference: hiprt/impl/MortonCode.h:60-277). Division and log2 use the fast hardware
// approximations because the reference compiles its build kernels with --use_fast_math
// (hiprt/impl/Compiler.cpp:516-522); 32-bit wrap-around is spelled out with unsigned arithmetic.
__device__ __forceinline__ uint32_t spreadBits2( uint32_t v )
{
	v &= 0x0000ffffu;
	v = ( v ^ ( v << 8 ) ) & 0x00ff00ffu;
	v = ( v ^ ( v << 4 ) ) & 0x0f0f0f0fu;
	v = ( v ^ ( v << 2 ) ) & 0x33333333u;
	v = ( v ^ ( v << 1 ) ) & 0x55555555u;
	return v;
}
__device__ __forceinline__ uint32_t spreadBits3( uint32_t v )
{
	v = ( v * 0x00010001u ) & 0xFF0000FFu;
	v = ( v * 0x00000101u ) & 0x0F00F00Fu;
	v = ( v * 0x00000011u ) & 0xC30C30C3u;
	v = ( v * 0x00000005u ) & 0x49249249u;
	return v;
}
__device__ __forceinline__ uint32_t shlClamp( uint32_t v, int s ) { return ( s < 0 || s > 31 ) ? 0u : v << s; }
__device__ __forceinline__ uint32_t shrClamp( uint32_t v, int s ) { return ( s < 0 || s > 31 ) ? 0u : v >> s; }
__device__ __forceinline__ int wrapAdd( int a, int b ) { return int( uint32_t( a ) + uint32_t( b ) ); }
__device__ __forceinline__ int wrapSub( int a, int b ) { return int( uint32_t( a ) - uint32_t( b ) ); }

// Per-build constants derived from the scene extent only (hoisted out of the per-primitive path).
struct MortonLayout
{
	int ax0, ax1, ax2; // axes by decreasing extent
	int px, py, swap, preSum;
	int bx, by, bz;
};

__device__ __forceinline__ float axisOf( const float3 v, int a ) { return a == 0 ? v.x : ( a == 1 ? v.y : v.z ); }

__device__ __forceinline__ MortonLayout mortonLayout( const float3 ext )
{
	MortonLayout L;
	int			 a0 = 0, a1 = 1, a2 = 2;
	int			 t;
	if ( axisOf( ext, a0 ) < axisOf( ext, a1 ) ) t = a0, a0 = a1, a1 = t;
	if ( axisOf( ext, a1 ) < axisOf( ext, a2 ) ) t = a1, a1 = a2, a2 = t;
	if ( axisOf( ext, a0 ) < axisOf( ext, a1 ) ) t = a0, a0 = a1, a1 = t;
	const float eb = axisOf( ext, a0 ), em = axisOf( ext, a1 ), es = axisOf( ext, a2 );
	int			px = __float2int_rz( __log2f( __fdividef( eb, em ) ) );
	int			py = __float2int_rz( __log2f( __fdividef( em, es ) ) );
	const int	pz = __float2int_rz( __log2f( __fdividef( eb, es ) ) );
	int			swap = wrapSub( pz, wrapAdd( px, py ) );
	px			   = min( px, 30 );
	py			   = min( int( uint32_t( py ) * 2u ), 30 - px ) / 2;
	int preSum	   = px + py * 2;
	if ( preSum != 30 )
		preSum = wrapAdd( preSum, swap );
	else
		swap = 0;
	int bx, by;
	const int bz = ( es != 0.0f ) ? max( 0, wrapSub( 30, preSum ) / 3 ) : 0;
	if ( swap > 0 )
	{
		bx = max( 0, wrapAdd( wrapAdd( wrapAdd( wrapSub( wrapSub( 30, bz ), preSum ) / 2, py ), px ), 1 ) );
		by = 30 - bx - bz;
	}
	else
	{
		by = max( 0, wrapAdd( wrapSub( wrapSub( 30, bz ), preSum ) / 2, py ) );
		bx = 30 - by - bz;
	}
	L.ax0 = a0, L.ax1 = a1, L.ax2 = a2;
	L.px = px, L.py = py, L.swap = swap, L.preSum = preSum;
	L.bx = bx, L.by = by, L.bz = bz;
	return L;
}

__device__ __forceinline__ uint32_t quantizeAxis( float p, int bits )
{
	const uint32_t scale = shlClamp( 1u, bits );
	return min( __float2uint_rz( fmaxf( p * float( scale ), 0.0f ) ), scale - 1u );
}

__device__ __forceinline__ uint32_t mortonCode( const MortonLayout& L, const float3 pos )
{
	int		 bx = L.bx, by = L.by;
	const int bz = L.bz;
	uint32_t cx = quantizeAxis( axisOf( pos, L.ax0 ), bx );
	uint32_t cy = quantizeAxis( axisOf( pos, L.ax1 ), by );
	uint32_t cz = quantizeAxis( axisOf( pos, L.ax2 ), bz );
	uint32_t code = 0;
	int		 d0 = 0, d1 = 0;
	if ( L.preSum > 0 )
	{
		bx -= L.px;
		code = shrClamp( cx & shlClamp( shlClamp( 1u, L.px ) - 1u, bx ), bx );
		code = shlClamp( code, L.py * 2 );
		bx -= L.py;
		by -= L.py;
		const uint32_t t0 = spreadBits2( shrClamp( cx & shlClamp( shlClamp( 1u, L.py ) - 1u, bx ), bx ) );
		const uint32_t t1 = spreadBits2( shrClamp( cy & shlClamp( shlClamp( 1u, L.py ) - 1u, by ), by ) );
		code |= t0 * 2 + t1;
		if ( L.swap > 0 )
		{
			code <<= 1;
			bx -= 1;
			code |= shrClamp( cx & shlClamp( 1u, bx ), bx );
		}
		code = shlClamp( code, bx + by + bz );
		cx &= shlClamp( 1u, bx ) - 1u;
		cy &= shlClamp( 1u, by ) - 1u;
		if ( L.swap > 0 )
		{
			d0 = by - bx;
			cx = shlClamp( cx, d0 );
			d1 = by - bz;
			cz = shlClamp( cz, d1 );
		}
		else
		{
			d0 = bx - by;
			cy = shlClamp( cy, d0 );
			d1 = bx - bz;
			cz = shlClamp( cz, d1 );
		}
	}
	if ( bz == 0 )
	{
		code |= spreadBits2( cx ) * 2 + spreadBits2( cy );
	}
	else
	{
		cx = cx ? spreadBits3( cx ) : 0;
		cy = cy ? spreadBits3( cy ) : 0;
		cz = cz ? spreadBits3( cz ) : 0;
		if ( L.swap > 0 )
			code |= shrClamp( cy * 4 + cx * 2 + cz, d0 + d1 );
		else
			code |= shrClamp( cx * 4 + cy * 2 + cz, d0 + d1 );
	}
	return code;
}
} // namespace hiprt
