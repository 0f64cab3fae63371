// Bounce kernel of the wavefront diffuse path tracer (hiprt_b200/pathtrace.py), compiled at run time through
// hiprtBuildTraceKernels exactly like an application's own kernels (NVRTC, <hiprt/hiprt_device.h> prepended).
// One thread per path slot of the wavefront's capacity (the live count is a device word: the host never reads it); per live path: a miss adds the sky to the path's pixel slot (one writer per slot and launch, no atomics on
// the image), a hit scatters a cosine-distributed ray from the hit point and appends it to the next wavefront.
#if !defined( __CUDACC_RTC__ ) // offline compile (nvcc -I include): NVRTC gets the header and dispatchers from hiprtBuildTraceKernels
#define HIPRT_B200_DEFAULT_FUNC_TABLE
#include <hiprt/hiprt_device.h>
#endif

__device__ __forceinline__ float hashU01( unsigned long long a, unsigned long long b, unsigned long long c, unsigned long long d )
{
	// the same counter-based hash as pathtrace._hash_u01: a ray depends on (pixel, sample, bounce, dimension) only
	unsigned long long x = ( a * 0x9E3779B1ull + b * 0x85EBCA77ull + c * 0xC2B2AE3Dull + d * 0x27D4EB2Full + 0x165667B1ull ) & 0xFFFFFFFFull;
	x = ( ( x ^ ( x >> 15 ) ) * 0x2C1B3C6Dull ) & 0xFFFFFFFFull;
	x = ( ( x ^ ( x >> 12 ) ) * 0x297A2D39ull ) & 0xFFFFFFFFull;
	x = x ^ ( x >> 15 );
	return float( x >> 8 ) * ( 1.0f / 16777216.0f );
}

extern "C" __global__ void diffuseBounceKernel(
	const hiprtRay* __restrict__ rays,
	const hiprtHit* __restrict__ hits,
	const unsigned int* __restrict__ slotIn,
	const float* __restrict__ throughputIn,
	const unsigned int* __restrict__ countIn, // live paths of this wavefront (device word: the previous bounce's append counter)
	const long long* __restrict__ pixels, // pixel index of every pixel slot (numPixels of them)
	unsigned int	 numPixels,			  // a path slot = sample-in-batch * numPixels + pixel slot
	unsigned int	 sampleBase,		  // first sample of the batch
	unsigned int	 depth,
	float			 albedo,
	int				 lastDepth,
	hiprtRay* __restrict__ raysOut,
	unsigned int* __restrict__ slotOut,
	float* __restrict__ throughputOut,
	unsigned int* __restrict__ countOut,
	float* __restrict__ radiance ) // [slots][3]
{
	const unsigned int i	= blockIdx.x * blockDim.x + threadIdx.x;
	const bool		   live = i < *countIn;
	bool			   scatter = false;
	hiprtRay		   next;
	unsigned int	   slot = 0;
	float			   thr	= 0.0f;
	if ( live )
	{
		const hiprtRay ray = rays[i];
		const hiprtHit hit = hits[i];
		slot			   = slotIn[i];
		thr				   = throughputIn[i];
		if ( hit.primID == hiprtInvalidValue )
		{
			const float t = fminf( fmaxf( 0.5f * ( ray.direction.y + 1.0f ), 0.0f ), 1.0f );
			radiance[3 * slot + 0] += thr * ( ( 1.0f - t ) + t * 0.5f );
			radiance[3 * slot + 1] += thr * ( ( 1.0f - t ) + t * 0.7f );
			radiance[3 * slot + 2] += thr * 1.0f;
		}
		else if ( !lastDepth )
		{
			float3		n	= hit.normal;
			const float len = sqrtf( n.x * n.x + n.y * n.y + n.z * n.z );
			n				= make_float3( n.x / len, n.y / len, n.z / len );
			if ( n.x * ray.direction.x + n.y * ray.direction.y + n.z * ray.direction.z > 0.0f ) n = make_float3( -n.x, -n.y, -n.z );
			const float3 p = make_float3(
				ray.origin.x + ray.direction.x * hit.t + n.x * 1e-4f, ray.origin.y + ray.direction.y * hit.t + n.y * 1e-4f,
				ray.origin.z + ray.direction.z * hit.t + n.z * 1e-4f );
			const unsigned long long pix	= (unsigned long long)pixels[slot % numPixels];
			const unsigned int		 sample = sampleBase + slot / numPixels;
			const float u0 = hashU01( pix, sample, depth + 1, 1 ), u1 = hashU01( pix, sample, depth + 1, 2 );
			const float phi = 6.2831853f * u0, sinT = sqrtf( u1 ), cosT = sqrtf( fmaxf( 1.0f - u1, 0.0f ) );
			const float3 axis = fabsf( n.x ) > 0.5f ? make_float3( 0.0f, 1.0f, 0.0f ) : make_float3( 1.0f, 0.0f, 0.0f );
			float3		 t	  = make_float3( axis.y * n.z - axis.z * n.y, axis.z * n.x - axis.x * n.z, axis.x * n.y - axis.y * n.x );
			const float	 tl	  = sqrtf( t.x * t.x + t.y * t.y + t.z * t.z );
			t				  = make_float3( t.x / tl, t.y / tl, t.z / tl );
			const float3 b	  = make_float3( n.y * t.z - n.z * t.y, n.z * t.x - n.x * t.z, n.x * t.y - n.y * t.x );
			const float	 cx = cosf( phi ) * sinT, cy = sinf( phi ) * sinT;
			float3		 d	= make_float3( t.x * cx + b.x * cy + n.x * cosT, t.y * cx + b.y * cy + n.y * cosT, t.z * cx + b.z * cy + n.z * cosT );
			const float	 dl = sqrtf( d.x * d.x + d.y * d.y + d.z * d.z );
			next.origin		= p;
			next.minT		= 0.0f;
			next.direction	= make_float3( d.x / dl, d.y / dl, d.z / dl );
			next.maxT		= 3.0e38f;
			scatter			= true;
		}
	}
	// warp-aggregated append to the next wavefront
	const unsigned int mask = __ballot_sync( 0xFFFFFFFFu, scatter );
	if ( mask == 0 ) return;
	const unsigned int lane = threadIdx.x & 31u, leader = __ffs( mask ) - 1;
	unsigned int	   base = 0;
	if ( lane == leader ) base = atomicAdd( countOut, __popc( mask ) );
	base = __shfl_sync( 0xFFFFFFFFu, base, leader );
	if ( scatter )
	{
		const unsigned int j = base + __popc( mask & ( ( 1u << lane ) - 1u ) );
		raysOut[j]			 = next;
		slotOut[j]			 = slot;
		throughputOut[j]	 = thr * albedo;
	}
}

// Jittered pinhole camera rays of a BATCH of samples (sampleBase, sampleBase + 1, ...) for the pixel list: path slot i belongs to
// pixel slot i % numPixels and sample sampleBase + i / numPixels (the torch expression of pathtrace.camera_rays in one launch);
// also resets the wavefront: slot = identity, throughput = 1, live count = n.
extern "C" __global__ void cameraRaysKernel(
	const long long* __restrict__ pixels, unsigned int numPixels, unsigned int n, unsigned int width, unsigned int height, float3 eye, float3 u, float3 v, float3 w,
	float halfTan, unsigned int sampleBase, hiprtRay* __restrict__ raysOut, unsigned int* __restrict__ slotOut, float* __restrict__ throughputOut,
	unsigned int* __restrict__ countOut )
{
	const unsigned int i = blockIdx.x * blockDim.x + threadIdx.x;
	if ( i == 0 ) *countOut = n;
	if ( i >= n ) return;
	const unsigned long long pix	= (unsigned long long)pixels[i % numPixels];
	const unsigned int		 sample = sampleBase + i / numPixels;
	const float px = float( pix % width ), py = float( pix / width );
	const float jx = hashU01( pix, sample, 0, 1 ), jy = hashU01( pix, sample, 0, 2 );
	const float sx = ( ( px + jx ) / float( width ) * 2.0f - 1.0f ) * ( halfTan * float( width ) / float( height ) );
	const float sy = ( 1.0f - ( py + jy ) / float( height ) * 2.0f ) * halfTan;
	float3		d  = make_float3( sx * u.x + sy * v.x + w.x, sx * u.y + sy * v.y + w.y, sx * u.z + sy * v.z + w.z );
	const float dl = sqrtf( d.x * d.x + d.y * d.y + d.z * d.z );
	hiprtRay	ray;
	ray.origin	  = eye;
	ray.minT	  = 0.0f;
	ray.direction = make_float3( d.x / dl, d.y / dl, d.z / dl );
	ray.maxT	  = 3.0e38f;
	raysOut[i]		 = ray;
	slotOut[i]		 = i;
	throughputOut[i] = 1.0f;
}
