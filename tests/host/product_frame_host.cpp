// TEST INFRASTRUCTURE: the PRODUCT's device arithmetic for animated instances (include/hiprt/hiprt_b200_frame.h, the very header
// the CUDA traversal and the scene builder compile) built for the host, so that tests/test_product_frame_math.py can hold it to the
// reference's own code bit for bit without a GPU. Device qualifiers are defined away and the CUDA vector types come from the
// toolkit's host-usable headers; the non-contracting intrinsics become plain operations (this file is compiled with
// -ffp-contract=off, so nothing fuses that the header does not fuse explicitly).
#include <vector_functions.h>
#include <vector_types.h>

#include <cmath>
#include <cstdint>
#include <cstring>

#define __CUDACC__ 1
#define __device__
#define __host__
#define __forceinline__ inline
static inline float __fmul_rn( float a, float b ) { return a * b; }
static inline float __fadd_rn( float a, float b ) { return a + b; }
static inline float __fsub_rn( float a, float b ) { return a - b; }
template <typename T>
static inline T __ldg( const T* p )
{
	return *p;
}
#include <hiprt/hiprt_b200_frame.h>

using namespace hiprt;

extern "C" {
// frames: the scene storage's 64-byte key frames (SRTFrame: rotation, scale, shear, translation, time) of one instance
void prod_motion_ray( uint32_t frameCount, const void* frames, uint32_t n, const float* times, const hiprtRay* raysIn, hiprtRay* raysOut )
{
	for ( uint32_t i = 0; i < n; ++i )
	{
		const motion::Frame f = motion::frameAt( static_cast<const SRTFrame*>( frames ), 0, frameCount, times[i] );
		raysOut[i]			  = motion::transformRay( f, raysIn[i] );
	}
}

// object -> world of points and (normal-like) vectors, and back, under the pose at `time`
void prod_motion_points( uint32_t frameCount, const void* frames, float time, uint32_t n, const float* in, float* out )
{
	const motion::Frame f = motion::frameAt( static_cast<const SRTFrame*>( frames ), 0, frameCount, time );
	for ( uint32_t i = 0; i < n; ++i )
	{
		const float3 p = make_float3( in[3 * i], in[3 * i + 1], in[3 * i + 2] );
		const float3 r[4] = { motion::transformPoint( f, p ), motion::invTransformPoint( f, p ), motion::transformVector( f, p ), motion::invTransformVector( f, p ) };
		memcpy( out + 12 * i, r, sizeof( r ) );
	}
}
}
