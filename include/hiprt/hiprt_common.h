// hiprt_b200 -- common macros and constants shared by host and device code.
//
// Drop-in counterpart of the reference's hiprt/hiprt_common.h (constants at :170-179,
// RoundUp/DivideRoundUp at :324-340). Only the pieces the public API needs are kept;
// there is no multi-backend switch here: device code is CUDA for sm_100a.
#pragma once

#if defined( __CUDACC__ )
#define HIPRT_B200_DEVICE_CODE 1
// what HIPRT user code tests to tell kernel translation units from host ones (reference: hiprt/hiprt_common.h:27-29)
#if !defined( __KERNELCC__ )
#define __KERNELCC__
#endif
#endif

#if !defined( __CUDACC_RTC__ )
#include <cstddef>
#include <cstdint>
#else
using int8_t   = signed char;
using uint8_t  = unsigned char;
using int16_t  = short;
using uint16_t = unsigned short;
using int32_t  = int;
using uint32_t = unsigned int;
using int64_t  = long long;
using uint64_t = unsigned long long;
using size_t   = unsigned long;
#endif

#if defined( __CUDACC__ )
#define HIPRT_HOST __host__
#define HIPRT_DEVICE __device__
#define HIPRT_HOST_DEVICE __host__ __device__
#define HIPRT_INLINE __forceinline__
#else
#define HIPRT_HOST
#define HIPRT_DEVICE
#define HIPRT_HOST_DEVICE
#define HIPRT_INLINE inline
#endif

#define HIPRT_STATIC_ASSERT( cond ) static_assert( ( cond ), #cond )

namespace hiprt
{
constexpr float Pi	   = 3.14159265358979323846f;
constexpr float TwoPi  = 2.0f * Pi;
constexpr float FltMin = 1.175494351e-38f;
constexpr float FltMax = 3.402823466e+38f;

constexpr uint32_t InvalidValue				 = 0xFFFFFFFFu;
constexpr uint32_t FullRayMask				 = 0xFFFFFFFFu;
constexpr uint32_t MaxBatchBuildMaxPrimCount = 512u;
constexpr uint32_t MaxInstanceLevels		 = 4u;
constexpr uint32_t DefaultAlignment			 = 64u;
constexpr uint32_t InstanceIDBits			 = 24u;
constexpr uint32_t WarpSize					 = 32u;
constexpr uint32_t BranchingFactor			 = 4u;

// type selection usable in NVRTC translation units, where <type_traits> is not available (reference: hiprt/hiprt_common.h:343-378)
template <bool B, class T, class F>
struct conditional
{
	using type = T;
};
template <class T, class F>
struct conditional<false, T, F>
{
	using type = F;
};
template <class T, class U>
struct is_same
{
	static constexpr bool value = false;
};
template <class T>
struct is_same<T, T>
{
	static constexpr bool value = true;
};

template <typename T, typename U>
constexpr HIPRT_HOST_DEVICE T RoundUp( T value, U factor )
{
	return ( value + factor - 1 ) / factor * factor;
}

template <typename T, typename U>
constexpr HIPRT_HOST_DEVICE T DivideRoundUp( T value, U factor )
{
	return ( value + factor - 1 ) / factor;
}
} // namespace hiprt
