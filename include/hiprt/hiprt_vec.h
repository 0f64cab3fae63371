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
// (the reference names the host-side template hiprt::Vector<T, N>: hiprt/hiprt_vec.h:32-52; user host code spells it that way)
template <typename T, uint32_t N>
struct Vector;
template <typename T>
struct alignas( 2 * sizeof( T ) ) Vector<T, 2>
{
	T x, y;
};
template <typename T>
struct Vector<T, 3>
{
	T x, y, z;
};
template <typename T>
struct alignas( 4 * sizeof( T ) ) Vector<T, 4>
{
	T x, y, z, w;
};
template <typename T>
using Vec2 = Vector<T, 2>;
template <typename T>
using Vec3 = Vector<T, 3>;
template <typename T>
using Vec4 = Vector<T, 4>;
} // namespace hiprt
using hiprtInt2	  = hiprt::Vector<int, 2>;
using hiprtInt3	  = hiprt::Vector<int, 3>;
using hiprtInt4	  = hiprt::Vector<int, 4>;
using hiprtUint2  = hiprt::Vector<unsigned int, 2>;
using hiprtUint3  = hiprt::Vector<unsigned int, 3>;
using hiprtUint4  = hiprt::Vector<unsigned int, 4>;
using hiprtFloat2 = hiprt::Vector<float, 2>;
using hiprtFloat3 = hiprt::Vector<float, 3>;
using hiprtFloat4 = hiprt::Vector<float, 4>;
#endif

HIPRT_STATIC_ASSERT( sizeof( hiprtFloat2 ) == 8 && alignof( hiprtFloat2 ) == 8 );
HIPRT_STATIC_ASSERT( sizeof( hiprtFloat3 ) == 12 && alignof( hiprtFloat3 ) == 4 );
HIPRT_STATIC_ASSERT( sizeof( hiprtFloat4 ) == 16 && alignof( hiprtFloat4 ) == 16 );
HIPRT_STATIC_ASSERT( sizeof( hiprtUint2 ) == 8 && sizeof( hiprtUint3 ) == 12 && sizeof( hiprtUint4 ) == 16 );
HIPRT_STATIC_ASSERT( sizeof( hiprtInt2 ) == 8 && sizeof( hiprtInt3 ) == 12 && sizeof( hiprtInt4 ) == 16 );
