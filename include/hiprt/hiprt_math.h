// hiprt_b200 -- vector arithmetic of the public API: the names a HIPRT user's kernels and host code call
// (reference surface: hiprt/hiprt_math.h -- make_<type><n>, component-wise operators, min / max, fma, mix, clamp, sign,
// dot, cross, hypot, normalize, sumOfProducts / differenceOfProducts, atomicMinFloat / atomicMaxFloat).
//
// Written once over the nine vector types (int / unsigned / float x 2 / 3 / 4) with generator macros instead of per-type
// copies. In device translation units the types are CUDA's built-in ones, which live in the global namespace: their
// operators are declared there too (argument-dependent lookup would not find them in `hiprt`); on the host the types are
// hiprt::Vector<T, N> and the operators sit beside them. Named functions are always in namespace hiprt.
// The float3 products use the compensated forms (an fma recovers the rounding error of one product), as the reference
// does: user code such as its test kernels normalises through hypot() and expects those digits.
#pragma once
#include <hiprt/hiprt_common.h>
#include <hiprt/hiprt_vec.h>

#if !defined( __CUDACC_RTC__ )
#include <cmath>
#endif

// ---- construction: from a scalar, from / with smaller vectors (the all-components forms follow below) ------------------------
#define HIPRT_B200_MAKE( S, N2, N3, N4, V2, V3, V4 )                                                                  \
	HIPRT_HOST_DEVICE HIPRT_INLINE V2 N2( const S c ) { return V2{ c, c }; }                                          \
	HIPRT_HOST_DEVICE HIPRT_INLINE V3 N3( const S c ) { return V3{ c, c, c }; }                                       \
	HIPRT_HOST_DEVICE HIPRT_INLINE V4 N4( const S c ) { return V4{ c, c, c, c }; }                                    \
	HIPRT_HOST_DEVICE HIPRT_INLINE V2 N2( const V3& a ) { return V2{ a.x, a.y }; }                                    \
	HIPRT_HOST_DEVICE HIPRT_INLINE V2 N2( const V4& a ) { return V2{ a.x, a.y }; }                                    \
	HIPRT_HOST_DEVICE HIPRT_INLINE V3 N3( const V4& a ) { return V3{ a.x, a.y, a.z }; }                               \
	HIPRT_HOST_DEVICE HIPRT_INLINE V3 N3( const V2& a, const S c ) { return V3{ a.x, a.y, c }; }                      \
	HIPRT_HOST_DEVICE HIPRT_INLINE V3 N3( const S c, const V2& a ) { return V3{ c, a.x, a.y }; }                      \
	HIPRT_HOST_DEVICE HIPRT_INLINE V4 N4( const V2& a, const S c0, const S c1 ) { return V4{ a.x, a.y, c0, c1 }; }    \
	HIPRT_HOST_DEVICE HIPRT_INLINE V4 N4( const S c0, const V2& a, const S c1 ) { return V4{ c0, a.x, a.y, c1 }; }    \
	HIPRT_HOST_DEVICE HIPRT_INLINE V4 N4( const S c0, const S c1, const V2& a ) { return V4{ c0, c1, a.x, a.y }; }    \
	HIPRT_HOST_DEVICE HIPRT_INLINE V4 N4( const V2& a, const V2& b ) { return V4{ a.x, a.y, b.x, b.y }; }             \
	HIPRT_HOST_DEVICE HIPRT_INLINE V4 N4( const V3& a, const S c ) { return V4{ a.x, a.y, a.z, c }; }                 \
	HIPRT_HOST_DEVICE HIPRT_INLINE V4 N4( const S c, const V3& a ) { return V4{ c, a.x, a.y, a.z }; }

// conversions between element types: N(src vector of another element type)
#define HIPRT_B200_CONVERT( S, N2, N3, N4, V2, V3, V4, W2, W3, W4 )                                                                        \
	HIPRT_HOST_DEVICE HIPRT_INLINE V2 N2( const W2& a ) { return V2{ static_cast<S>( a.x ), static_cast<S>( a.y ) }; }                      \
	HIPRT_HOST_DEVICE HIPRT_INLINE V3 N3( const W3& a ) { return V3{ static_cast<S>( a.x ), static_cast<S>( a.y ), static_cast<S>( a.z ) }; } \
	HIPRT_HOST_DEVICE HIPRT_INLINE V4 N4( const W4& a )                                                                                      \
	{                                                                                                                                        \
		return V4{ static_cast<S>( a.x ), static_cast<S>( a.y ), static_cast<S>( a.z ), static_cast<S>( a.w ) };                             \
	}

// ---- component-wise operators -------------------------------------------------------------------------------------
#define HIPRT_B200_BINARY( S, V2, V3, V4, OP )                                                                                              \
	HIPRT_HOST_DEVICE HIPRT_INLINE V2 operator OP( const V2& a, const V2& b ) { return V2{ a.x OP b.x, a.y OP b.y }; }                      \
	HIPRT_HOST_DEVICE HIPRT_INLINE V3 operator OP( const V3& a, const V3& b ) { return V3{ a.x OP b.x, a.y OP b.y, a.z OP b.z }; }          \
	HIPRT_HOST_DEVICE HIPRT_INLINE V4 operator OP( const V4& a, const V4& b ) { return V4{ a.x OP b.x, a.y OP b.y, a.z OP b.z, a.w OP b.w }; } \
	HIPRT_HOST_DEVICE HIPRT_INLINE V2 operator OP( const V2& a, const S c ) { return V2{ a.x OP c, a.y OP c }; }                            \
	HIPRT_HOST_DEVICE HIPRT_INLINE V3 operator OP( const V3& a, const S c ) { return V3{ a.x OP c, a.y OP c, a.z OP c }; }                  \
	HIPRT_HOST_DEVICE HIPRT_INLINE V4 operator OP( const V4& a, const S c ) { return V4{ a.x OP c, a.y OP c, a.z OP c, a.w OP c }; }        \
	HIPRT_HOST_DEVICE HIPRT_INLINE V2 operator OP( const S c, const V2& a ) { return V2{ c OP a.x, c OP a.y }; }                            \
	HIPRT_HOST_DEVICE HIPRT_INLINE V3 operator OP( const S c, const V3& a ) { return V3{ c OP a.x, c OP a.y, c OP a.z }; }                  \
	HIPRT_HOST_DEVICE HIPRT_INLINE V4 operator OP( const S c, const V4& a ) { return V4{ c OP a.x, c OP a.y, c OP a.z, c OP a.w }; }        \
	HIPRT_HOST_DEVICE HIPRT_INLINE V2& operator OP##=( V2& a, const V2& b ) { return a = a OP b; }                                          \
	HIPRT_HOST_DEVICE HIPRT_INLINE V3& operator OP##=( V3& a, const V3& b ) { return a = a OP b; }                                          \
	HIPRT_HOST_DEVICE HIPRT_INLINE V4& operator OP##=( V4& a, const V4& b ) { return a = a OP b; }                                          \
	HIPRT_HOST_DEVICE HIPRT_INLINE V2& operator OP##=( V2& a, const S c ) { return a = a OP c; }                                            \
	HIPRT_HOST_DEVICE HIPRT_INLINE V3& operator OP##=( V3& a, const S c ) { return a = a OP c; }                                            \
	HIPRT_HOST_DEVICE HIPRT_INLINE V4& operator OP##=( V4& a, const S c ) { return a = a OP c; }

#define HIPRT_B200_NEGATE( V2, V3, V4 )                                                           \
	HIPRT_HOST_DEVICE HIPRT_INLINE V2 operator-( const V2& a ) { return V2{ -a.x, -a.y }; }       \
	HIPRT_HOST_DEVICE HIPRT_INLINE V3 operator-( const V3& a ) { return V3{ -a.x, -a.y, -a.z }; } \
	HIPRT_HOST_DEVICE HIPRT_INLINE V4 operator-( const V4& a ) { return V4{ -a.x, -a.y, -a.z, -a.w }; }

#define HIPRT_B200_ARITHMETIC( S, V2, V3, V4 ) \
	HIPRT_B200_BINARY( S, V2, V3, V4, + )      \
	HIPRT_B200_BINARY( S, V2, V3, V4, - )      \
	HIPRT_B200_BINARY( S, V2, V3, V4, * )      \
	HIPRT_B200_BINARY( S, V2, V3, V4, / )

// ---- component-wise min / max (F = the scalar function) --------------------------------------------------------------
#define HIPRT_B200_MINMAX( S, V2, V3, V4, NAME, F )                                                                                          \
	HIPRT_HOST_DEVICE HIPRT_INLINE V2 NAME( const V2& a, const V2& b ) { return V2{ F( a.x, b.x ), F( a.y, b.y ) }; }                        \
	HIPRT_HOST_DEVICE HIPRT_INLINE V3 NAME( const V3& a, const V3& b ) { return V3{ F( a.x, b.x ), F( a.y, b.y ), F( a.z, b.z ) }; }         \
	HIPRT_HOST_DEVICE HIPRT_INLINE V4 NAME( const V4& a, const V4& b ) { return V4{ F( a.x, b.x ), F( a.y, b.y ), F( a.z, b.z ), F( a.w, b.w ) }; } \
	HIPRT_HOST_DEVICE HIPRT_INLINE V2 NAME( const V2& a, const S c ) { return V2{ F( a.x, c ), F( a.y, c ) }; }                              \
	HIPRT_HOST_DEVICE HIPRT_INLINE V3 NAME( const V3& a, const S c ) { return V3{ F( a.x, c ), F( a.y, c ), F( a.z, c ) }; }                 \
	HIPRT_HOST_DEVICE HIPRT_INLINE V4 NAME( const V4& a, const S c ) { return V4{ F( a.x, c ), F( a.y, c ), F( a.z, c ), F( a.w, c ) }; }    \
	HIPRT_HOST_DEVICE HIPRT_INLINE V2 NAME( const S c, const V2& a ) { return V2{ F( c, a.x ), F( c, a.y ) }; }                              \
	HIPRT_HOST_DEVICE HIPRT_INLINE V3 NAME( const S c, const V3& a ) { return V3{ F( c, a.x ), F( c, a.y ), F( c, a.z ) }; }                 \
	HIPRT_HOST_DEVICE HIPRT_INLINE V4 NAME( const S c, const V4& a ) { return V4{ F( c, a.x ), F( c, a.y ), F( c, a.z ), F( c, a.w ) }; }

// Operators: next to the types (global namespace for CUDA's built-in vector types, hiprt for the host structs).
#if !defined( __CUDACC__ )
namespace hiprt
{
#endif
HIPRT_B200_ARITHMETIC( int, hiprtInt2, hiprtInt3, hiprtInt4 )
HIPRT_B200_ARITHMETIC( unsigned int, hiprtUint2, hiprtUint3, hiprtUint4 )
HIPRT_B200_ARITHMETIC( float, hiprtFloat2, hiprtFloat3, hiprtFloat4 )
HIPRT_B200_NEGATE( hiprtInt2, hiprtInt3, hiprtInt4 )
HIPRT_B200_NEGATE( hiprtFloat2, hiprtFloat3, hiprtFloat4 )
#if !defined( __CUDACC__ )
} // namespace hiprt
#endif

namespace hiprt
{
// all-components forms: CUDA's own where they exist (brought into this namespace so that the overloads below do not hide them and
// unqualified calls under `using namespace hiprt` stay unambiguous), ours for plain host translation units
#if defined( __CUDACC__ ) || defined( __VECTOR_FUNCTIONS_H__ )
using ::make_int2;
using ::make_int3;
using ::make_int4;
using ::make_uint2;
using ::make_uint3;
using ::make_uint4;
using ::make_float2;
using ::make_float3;
using ::make_float4;
#else
#define HIPRT_B200_MAKE_ALL( S, N2, N3, N4, V2, V3, V4 )                                                             \
	HIPRT_HOST_DEVICE HIPRT_INLINE V2 N2( const S x, const S y ) { return V2{ x, y }; }                              \
	HIPRT_HOST_DEVICE HIPRT_INLINE V3 N3( const S x, const S y, const S z ) { return V3{ x, y, z }; }                \
	HIPRT_HOST_DEVICE HIPRT_INLINE V4 N4( const S x, const S y, const S z, const S w ) { return V4{ x, y, z, w }; }
HIPRT_B200_MAKE_ALL( int, make_int2, make_int3, make_int4, hiprtInt2, hiprtInt3, hiprtInt4 )
HIPRT_B200_MAKE_ALL( unsigned int, make_uint2, make_uint3, make_uint4, hiprtUint2, hiprtUint3, hiprtUint4 )
HIPRT_B200_MAKE_ALL( float, make_float2, make_float3, make_float4, hiprtFloat2, hiprtFloat3, hiprtFloat4 )
#undef HIPRT_B200_MAKE_ALL
#endif
HIPRT_B200_MAKE( int, make_int2, make_int3, make_int4, hiprtInt2, hiprtInt3, hiprtInt4 )
HIPRT_B200_MAKE( unsigned int, make_uint2, make_uint3, make_uint4, hiprtUint2, hiprtUint3, hiprtUint4 )
HIPRT_B200_MAKE( float, make_float2, make_float3, make_float4, hiprtFloat2, hiprtFloat3, hiprtFloat4 )
HIPRT_B200_CONVERT( int, make_int2, make_int3, make_int4, hiprtInt2, hiprtInt3, hiprtInt4, hiprtUint2, hiprtUint3, hiprtUint4 )
HIPRT_B200_CONVERT( int, make_int2, make_int3, make_int4, hiprtInt2, hiprtInt3, hiprtInt4, hiprtFloat2, hiprtFloat3, hiprtFloat4 )
HIPRT_B200_CONVERT( unsigned int, make_uint2, make_uint3, make_uint4, hiprtUint2, hiprtUint3, hiprtUint4, hiprtInt2, hiprtInt3, hiprtInt4 )
HIPRT_B200_CONVERT( unsigned int, make_uint2, make_uint3, make_uint4, hiprtUint2, hiprtUint3, hiprtUint4, hiprtFloat2, hiprtFloat3, hiprtFloat4 )
HIPRT_B200_CONVERT( float, make_float2, make_float3, make_float4, hiprtFloat2, hiprtFloat3, hiprtFloat4, hiprtInt2, hiprtInt3, hiprtInt4 )
HIPRT_B200_CONVERT( float, make_float2, make_float3, make_float4, hiprtFloat2, hiprtFloat3, hiprtFloat4, hiprtUint2, hiprtUint3, hiprtUint4 )

HIPRT_HOST_DEVICE HIPRT_INLINE int			minScalar( int a, int b ) { return a < b ? a : b; }
HIPRT_HOST_DEVICE HIPRT_INLINE int			maxScalar( int a, int b ) { return a > b ? a : b; }
HIPRT_HOST_DEVICE HIPRT_INLINE unsigned int minScalar( unsigned int a, unsigned int b ) { return a < b ? a : b; }
HIPRT_HOST_DEVICE HIPRT_INLINE unsigned int maxScalar( unsigned int a, unsigned int b ) { return a > b ? a : b; }
HIPRT_HOST_DEVICE HIPRT_INLINE float		minScalar( float a, float b ) { return fminf( a, b ); }
HIPRT_HOST_DEVICE HIPRT_INLINE float		maxScalar( float a, float b ) { return fmaxf( a, b ); }
HIPRT_B200_MINMAX( int, hiprtInt2, hiprtInt3, hiprtInt4, min, minScalar )
HIPRT_B200_MINMAX( int, hiprtInt2, hiprtInt3, hiprtInt4, max, maxScalar )
HIPRT_B200_MINMAX( unsigned int, hiprtUint2, hiprtUint3, hiprtUint4, min, minScalar )
HIPRT_B200_MINMAX( unsigned int, hiprtUint2, hiprtUint3, hiprtUint4, max, maxScalar )
HIPRT_B200_MINMAX( float, hiprtFloat2, hiprtFloat3, hiprtFloat4, min, minScalar )
HIPRT_B200_MINMAX( float, hiprtFloat2, hiprtFloat3, hiprtFloat4, max, maxScalar )
// scalars, mixed types included (clamp( uint32_t, 0, 255 )): inside namespace hiprt the vector overloads above hide ::min / ::max.
// The result has the type of a + b; float operands keep fminf / fmaxf semantics (a NaN operand yields the other one).
HIPRT_HOST_DEVICE HIPRT_INLINE float min( const float a, const float b ) { return fminf( a, b ); }
HIPRT_HOST_DEVICE HIPRT_INLINE float max( const float a, const float b ) { return fmaxf( a, b ); }
template <typename A, typename B>
HIPRT_HOST_DEVICE HIPRT_INLINE auto min( const A a, const B b ) -> decltype( a + b )
{
	using R = decltype( a + b );
	return static_cast<R>( a ) < static_cast<R>( b ) ? static_cast<R>( a ) : static_cast<R>( b );
}
template <typename A, typename B>
HIPRT_HOST_DEVICE HIPRT_INLINE auto max( const A a, const B b ) -> decltype( a + b )
{
	using R = decltype( a + b );
	return static_cast<R>( a ) > static_cast<R>( b ) ? static_cast<R>( a ) : static_cast<R>( b );
}

HIPRT_HOST_DEVICE HIPRT_INLINE hiprtFloat3 fma( const hiprtFloat3& a, const hiprtFloat3& b, const hiprtFloat3& c )
{
	return hiprtFloat3{ fmaf( a.x, b.x, c.x ), fmaf( a.y, b.y, c.y ), fmaf( a.z, b.z, c.z ) };
}

// a*b + c*d and a*b - c*d with the rounding error of c*d recovered by an fma (Kahan)
HIPRT_HOST_DEVICE HIPRT_INLINE float sumOfProducts( const float a, const float b, const float c, const float d )
{
	const float cd = c * d;
	return fmaf( a, b, cd ) + fmaf( c, d, -cd );
}
HIPRT_HOST_DEVICE HIPRT_INLINE float differenceOfProducts( const float a, const float b, const float c, const float d )
{
	const float cd = c * d;
	return fmaf( a, b, -cd ) - fmaf( c, d, -cd );
}

// length without overflow / underflow of the squares: the largest magnitude is factored out first
HIPRT_HOST_DEVICE HIPRT_INLINE float hypot( const hiprtFloat3& a )
{
	const float x = fabsf( a.x ), y = fabsf( a.y ), z = fabsf( a.z );
	const float m = x < y ? ( y < z ? z : y ) : ( x < z ? z : x );
	if ( m == 0.0f ) return 0.0f;
	const float sx = x / m, sy = y / m, sz = z / m;
	return m * sqrtf( sx * sx + sy * sy + sz * sz );
}

HIPRT_HOST_DEVICE HIPRT_INLINE hiprtFloat3 cross( const hiprtFloat3& a, const hiprtFloat3& b )
{
	return hiprtFloat3{ differenceOfProducts( a.y, b.z, a.z, b.y ), differenceOfProducts( a.z, b.x, a.x, b.z ), differenceOfProducts( a.x, b.y, a.y, b.x ) };
}

HIPRT_HOST_DEVICE HIPRT_INLINE float dot( const hiprtFloat3& a, const hiprtFloat3& b ) { return fmaf( a.x, b.x, sumOfProducts( a.y, b.y, a.z, b.z ) ); }

HIPRT_HOST_DEVICE HIPRT_INLINE hiprtFloat3 normalize( const hiprtFloat3& a ) { return a / hypot( a ); }

template <typename V>
HIPRT_HOST_DEVICE HIPRT_INLINE auto* ptr( V& a )
{
	return &a.x;
}

template <typename T, typename V>
HIPRT_HOST_DEVICE HIPRT_INLINE V mix( const V& lo, const V& hi, const T& t )
{
	return lo * ( static_cast<T>( 1 ) - t ) + hi * t;
}

template <typename T, typename V>
HIPRT_HOST_DEVICE HIPRT_INLINE V clamp( const V& v, const T& lo, const T& hi )
{
	return max( min( v, hi ), lo );
}

template <typename T>
HIPRT_HOST_DEVICE HIPRT_INLINE T sign( T val )
{
	return val < T( 0 ) ? T( -1 ) : ( val == T( 0 ) ? T( 0 ) : T( 1 ) );
}

#if defined( __CUDACC__ )
// float min / max through the integer atomics: non-negative floats order like ints, negative ones like reversed unsigneds
HIPRT_DEVICE HIPRT_INLINE float atomicMinFloat( float* addr, float value )
{
	return __float_as_int( value ) >= 0 ? __int_as_float( atomicMin( reinterpret_cast<int*>( addr ), __float_as_int( value ) ) )
										: __uint_as_float( atomicMax( reinterpret_cast<unsigned int*>( addr ), __float_as_uint( value ) ) );
}
HIPRT_DEVICE HIPRT_INLINE float atomicMaxFloat( float* addr, float value )
{
	return __float_as_int( value ) >= 0 ? __int_as_float( atomicMax( reinterpret_cast<int*>( addr ), __float_as_int( value ) ) )
										: __uint_as_float( atomicMin( reinterpret_cast<unsigned int*>( addr ), __float_as_uint( value ) ) );
}
#endif
} // namespace hiprt

#undef HIPRT_B200_MAKE
#undef HIPRT_B200_CONVERT
#undef HIPRT_B200_BINARY
#undef HIPRT_B200_NEGATE
#undef HIPRT_B200_ARITHMETIC
#undef HIPRT_B200_MINMAX
