// ORACLE -- TEST / BENCHMARK INFRASTRUCTURE ONLY.
// Device half of the reference driver (see ref_harness.cpp): trace kernels that instantiate the
// UNMODIFIED reference's own traversal objects, included from /root/reference at build time
// (hiprt/hiprt_device.h + hiprt/impl/hiprt_device_impl.h). Written like the reference's unit-test
// kernel TraceKernel (test/kernels/HiprtTestKernel.h:102-116): one thread per ray, hiprtGlobalStack
// with a 16-entry shared part, blocks of 64 threads (test/hiprtTest.cpp:1510-1530).
// Compiled device-only (nvcc -fatbin), as the reference compiles its own kernels.
#include <cuda_runtime.h>
#include <cmath>

#include <hiprt/hiprt_device.h>
#include <hiprt/impl/hiprt_device_impl.h>

// no custom functions are registered by this harness
__device__ bool intersectFunc( uint32_t, uint32_t, const hiprtFuncTableHeader&, const hiprtRay&, void*, hiprtHit& ) { return false; }
__device__ bool filterFunc( uint32_t, uint32_t, const hiprtFuncTableHeader&, const hiprtRay&, void*, const hiprtHit& ) { return false; }

#define REF_SHARED_STACK_SIZE 16
#define REF_BLOCK_SIZE 64

template <bool AnyHit>
__device__ void refTraceGeomImpl( hiprtGeometry geom, uint32_t numRays, hiprtGlobalStackBuffer globalStackBuffer, const hiprtRay* rays, hiprtHit* hits )
{
	const uint32_t index = blockIdx.x * blockDim.x + threadIdx.x;
	if ( index >= numRays ) return;
	__shared__ uint32_t	   sharedStackCache[REF_SHARED_STACK_SIZE * REF_BLOCK_SIZE];
	hiprtSharedStackBuffer sharedStackBuffer{ REF_SHARED_STACK_SIZE, sharedStackCache };
	hiprtGlobalStack	   stack( globalStackBuffer, sharedStackBuffer );
	if constexpr ( AnyHit )
	{
		hiprtGeomTraversalAnyHitCustomStack<hiprtGlobalStack> tr( geom, rays[index], stack );
		hits[index] = tr.getNextHit();
	}
	else
	{
		hiprtGeomTraversalClosestCustomStack<hiprtGlobalStack> tr( geom, rays[index], stack );
		hits[index] = tr.getNextHit();
	}
}

template <bool AnyHit>
__device__ void refTraceSceneImpl( hiprtScene scene, uint32_t numRays, hiprtGlobalStackBuffer globalStackBuffer, const hiprtRay* rays, hiprtHit* hits )
{
	const uint32_t index = blockIdx.x * blockDim.x + threadIdx.x;
	if ( index >= numRays ) return;
	__shared__ uint32_t		sharedStackCache[REF_SHARED_STACK_SIZE * REF_BLOCK_SIZE];
	hiprtSharedStackBuffer	sharedStackBuffer{ REF_SHARED_STACK_SIZE, sharedStackCache };
	hiprtGlobalStack		stack( globalStackBuffer, sharedStackBuffer );
	hiprtEmptyInstanceStack instanceStack;
	if constexpr ( AnyHit )
	{
		hiprtSceneTraversalAnyHitCustomStack<hiprtGlobalStack, hiprtEmptyInstanceStack> tr( scene, rays[index], stack, instanceStack );
		hits[index] = tr.getNextHit();
	}
	else
	{
		hiprtSceneTraversalClosestCustomStack<hiprtGlobalStack, hiprtEmptyInstanceStack> tr( scene, rays[index], stack, instanceStack );
		hits[index] = tr.getNextHit();
	}
}


#define REF_KERNEL( name, impl, handleType )                                                                                         \
	extern "C" __global__ void __launch_bounds__( REF_BLOCK_SIZE )                                                                   \
		name( handleType h, uint32_t numRays, hiprtGlobalStackBuffer globalStackBuffer, const hiprtRay* rays, hiprtHit* hits )       \
	{                                                                                                                                \
		impl( h, numRays, globalStackBuffer, rays, hits );                                                                           \
	}
REF_KERNEL( refTraceGeomClosest, refTraceGeomImpl<false>, hiprtGeometry )
REF_KERNEL( refTraceGeomAnyHit, refTraceGeomImpl<true>, hiprtGeometry )
REF_KERNEL( refTraceSceneClosest, refTraceSceneImpl<false>, hiprtScene )
REF_KERNEL( refTraceSceneAnyHit, refTraceSceneImpl<true>, hiprtScene )
