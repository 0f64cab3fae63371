// hiprt_b200 -- vector types of the public API (reference: hiprt/hiprt_vec.h:29-93).
// Host translation units get plain structs with CUDA's size/alignment; device
// translation units alias CUDA's built-in vector types.
#pragma once
#include <hiprt/hiprt_common.h>

#if defined( __CUDACC__ )
using hiprtInt2	  = int2;
using hiprtInt3	  = int3;
using hiprtInt4	  = int4;
using hiprtUint2  = uint2;
using hiprtUint3  = uint3;
using hiprtUint4  = uint4;
using hiprtFloat2 = float2;
using hiprtFloat3 = float3;
using hiprtFloat4 = float4;
#else
namespace hiprt
{
template <typename T>
struct alignas( 2 * sizeof( T ) ) Vec2
{
	T x, y;
};
template <typename T>
struct Vec3
{
	T x, y, z;
};
template <typename T>
struct alignas( 4 * sizeof( T ) ) Vec4
{
	T x, y, z, w;
};
} // namespace hiprt
using hiprtInt2	  = hiprt::Vec2<int>;
using hiprtInt3	  = hiprt::Vec3<int>;
using hiprtInt4	  = hiprt::Vec4<int>;
using hiprtUint2  = hiprt::Vec2<unsigned int>;
using hiprtUint3  = hiprt::Vec3<unsigned int>;
using hiprtUint4  = hiprt::Vec4<unsigned int>;
using hiprtFloat2 = hiprt::Vec2<float>;
using hiprtFloat3 = hiprt::Vec3<float>;
using hiprtFloat4 = hiprt::Vec4<float>;
#endif

HIPRT_STATIC_ASSERT( sizeof( hiprtFloat2 ) == 8 && alignof( hiprtFloat2 ) == 8 );
HIPRT_STATIC_ASSERT( sizeof( hiprtFloat3 ) == 12 && alignof( hiprtFloat3 ) == 4 );
HIPRT_STATIC_ASSERT( sizeof( hiprtFloat4 ) == 16 && alignof( hiprtFloat4 ) == 16 );
HIPRT_STATIC_ASSERT( sizeof( hiprtUint2 ) == 8 && sizeof( hiprtUint3 ) == 12 && sizeof( hiprtUint4 ) == 16 );
HIPRT_STATIC_ASSERT( sizeof( hiprtInt2 ) == 8 && sizeof( hiprtInt3 ) == 12 && sizeof( hiprtInt4 ) == 16 );
