// hiprt_b200 -- C-ABI entry points this build adds next to the reference API (hiprt/hiprt.h).
//
// The reference has no host-side "trace" call: users launch their own kernels that include
// hiprt_device.h (reference: test/kernels/HiprtTestKernel.h:102-116 `TraceKernel` is the
// canonical rays-in/hits-out kernel). For the benchmark, the parity tests and bindings from
// other languages we export that kernel as a library call; it is built from the same
// hiprt_device.h traversal objects a user kernel would use.
#pragma once
#include <hiprt/hiprt.h>

#ifdef __cplusplus
extern "C" {
#endif

// rays/hits are DEVICE pointers (hiprtRay[numRays], hiprtHit[numRays]); asynchronous on `stream`.
// traversalType: hiprtTraversalTerminateAtClosestHit or hiprtTraversalTerminateAtAnyHit (first hit found).
HIPRT_API hiprtError hiprtB200TraceGeometry(
	hiprtContext	   context,
	hiprtGeometry	   geometry,
	const hiprtRay*	   rays,
	uint32_t		   numRays,
	hiprtHit*		   hits,
	hiprtTraversalType traversalType,
	hiprtApiStream	   stream );

HIPRT_API hiprtError hiprtB200TraceScene(
	hiprtContext	   context,
	hiprtScene		   scene,
	const hiprtRay*	   rays,
	uint32_t		   numRays,
	hiprtHit*		   hits,
	hiprtTraversalType traversalType,
	hiprtRayMask	   mask,
	hiprtApiStream	   stream );

// Motion blur through the library kernel: `times` is a DEVICE array of one float per ray (the reference passes the time to the
// traversal object's constructor, hiprt_device.h:415-440; a hiprtRay itself carries none). times == nullptr traces every ray at
// `time`. Instances with several key frames are posed at the ray's time, static instances are unaffected.
HIPRT_API hiprtError hiprtB200TraceSceneAtTime(
	hiprtContext	   context,
	hiprtScene		   scene,
	const hiprtRay*	   rays,
	const float*	   times,
	float			   time,
	uint32_t		   numRays,
	hiprtHit*		   hits,
	hiprtTraversalType traversalType,
	hiprtRayMask	   mask,
	hiprtApiStream	   stream );

// Same with HOST buffers: copies the rays to the device, traces, copies the hits back and
// synchronises the stream (the end-to-end path measured by bench.py's "e2e").
HIPRT_API hiprtError hiprtB200TraceGeometryHost(
	hiprtContext	   context,
	hiprtGeometry	   geometry,
	const hiprtRay*	   hostRays,
	uint32_t		   numRays,
	hiprtHit*		   hostHits,
	hiprtTraversalType traversalType,
	hiprtApiStream	   stream );

HIPRT_API hiprtError hiprtB200TraceSceneHost(
	hiprtContext	   context,
	hiprtScene		   scene,
	const hiprtRay*	   hostRays,
	uint32_t		   numRays,
	hiprtHit*		   hostHits,
	hiprtTraversalType traversalType,
	hiprtRayMask	   mask,
	hiprtApiStream	   stream );

// Byte offsets of the builder's intermediate arrays inside the caller's temporary buffer
// (parity tests read sorted Morton keys, binary nodes etc. from there). 16 size_t values,
// see hiprt_b200/csrc/builders.h TempLayout.
HIPRT_API hiprtError hiprtB200GetTempLayout( size_t primCount, size_t pairedTriangleCount, size_t* offsetsOut, uint32_t numOffsets );

// The device-side API header tree (hiprt_device.h and what it includes) as embedded strings,
// for callers that NVRTC-compile their own kernels. Returns the number of headers.
HIPRT_API uint32_t hiprtB200GetDeviceHeaders( const char*** namesOut, const char*** sourcesOut );

// The device-independent half of hiprtBuildTraceKernelsFromBitcode (reference: Compiler::buildKernelsFromBitcode,
// hiprt/impl/Compiler.cpp:338-470), usable without a GPU or a context: generates the intersectFunc / filterFunc dispatchers for
// the function-name table, compiles them as relocatable code for sm_<ccMajor><ccMinor>[a] and links them with the caller's
// relocatable device code (nvcc -dc against <hiprt/hiprt_device.h>: cubin, PTX, fatbin or LTO-IR) through nvJitLink. Writes the
// linked cubin (cuModuleLoadData-ready) to cubinOut; *cubinSizeInOut is the capacity on entry and the cubin size on return
// (cubinOut == NULL: size query). On failure the compiler / linker log is copied to logOut.
HIPRT_API hiprtError hiprtB200LinkTraceKernels(
	int				  ccMajor,
	int				  ccMinor,
	const char*		  moduleName,
	const char*		  bitcodeBinary,
	size_t			  bitcodeBinarySize,
	uint32_t		  numGeomTypes,
	uint32_t		  numRayTypes,
	hiprtFuncNameSet* funcNameSets,
	void*			  cubinOut,
	size_t*			  cubinSizeInOut,
	char*			  logOut,
	size_t			  logCapacity );

HIPRT_API const char* hiprtB200GetLastErrorString( hiprtContext context );

// Per-phase timing of the most recent build on this context (CUDA events on the build stream;
// the reference's counterpart is its compile-time Timer, hiprt/impl/Timer.h:41).
HIPRT_API hiprtError hiprtB200EnableBuildProfile( hiprtContext context, int enable );
HIPRT_API hiprtError hiprtB200GetBuildProfile(
	hiprtContext context, uint32_t* phaseCountOut, const char** phaseNamesOut, float* phaseMsOut, uint32_t maxPhases );

// Synchronous device-to-host copy of raw bytes (lets bindings without a CUDA runtime of their
// own inspect built structures through the handle).
HIPRT_API hiprtError hiprtB200CopyToHost( hiprtContext context, void* hostDst, const void* deviceSrc, size_t bytes );

#ifdef __cplusplus
}
#endif
