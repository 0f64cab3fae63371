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

// ---- compact output records ------------------------------------------------------------------------------------------
// The traversal is the same as above; what changes is the record written per ray. Occlusion / ambient-occlusion queries need
// one word, most shading needs {primID, t, u, v}: the launch writes 4 or 16 bytes per ray instead of 48 (C3 AO frame: 0.54 GB
// of hit records per frame become 45 MB), and the host-buffer variant pulls that much less over PCIe.
// (Reference: the test kernels keep hits in registers and write one colour per pixel, test/kernels/AoRayKernel.h:60-100.)
typedef enum
{
	hiprtB200HitFormatFull	  = 0, // hiprtHit, 48 bytes
	hiprtB200HitFormatCompact = 1, // {uint32 primID, float t, float u, float v}, 16 bytes (primID hiprtInvalidValue = miss)
	hiprtB200HitFormatPrimID  = 2, // uint32 primID, 4 bytes
} hiprtB200HitFormat;

// bytes per record: 48 / 16 / 4
HIPRT_API size_t hiprtB200HitFormatSize( hiprtB200HitFormat format );

// Triangle geometries only (hiprtErrorInvalidParameter for custom-primitive geometries with a format other than Full).
HIPRT_API hiprtError hiprtB200TraceGeometryFormat(
	hiprtContext	   context,
	hiprtGeometry	   geometry,
	const hiprtRay*	   rays,
	uint32_t		   numRays,
	void*			   records,
	hiprtB200HitFormat format,
	hiprtTraversalType traversalType,
	hiprtApiStream	   stream );

// Indirect launch for wavefront renderers: the number of rays is read from the DEVICE word *numRaysDevice when the kernel starts
// (e.g. the append counter a shading kernel earlier on the stream has filled), `maxRays` is the capacity of the buffers. The host
// never learns the count, so a path tracer's bounces queue up without a read-back per bounce (hiprt_b200/pathtrace.py). Triangle
// geometries only. (The reference's test renderers launch one thread per pixel and loop over bounces inside the kernel,
// test/kernels/AoRayKernel.h; a library that takes rays and returns hits needs the count on the device instead.)
HIPRT_API hiprtError hiprtB200TraceGeometryIndirect(
	hiprtContext	   context,
	hiprtGeometry	   geometry,
	const hiprtRay*	   rays,
	uint32_t		   maxRays,
	const uint32_t*	   numRaysDevice,
	void*			   records,
	hiprtB200HitFormat format,
	hiprtTraversalType traversalType,
	hiprtApiStream	   stream );

HIPRT_API hiprtError hiprtB200TraceGeometryFormatHost(
	hiprtContext	   context,
	hiprtGeometry	   geometry,
	const hiprtRay*	   hostRays,
	uint32_t		   numRays,
	void*			   hostRecords,
	hiprtB200HitFormat format,
	hiprtTraversalType traversalType,
	hiprtApiStream	   stream );

// Instrumented trace of a triangle geometry: as hiprtB200TraceGeometry, and adds the number of box-node visits and leaf visits
// of all rays to visitCounts[0] / visitCounts[1] (DEVICE uint64[2], zeroed by the caller). Cache-level traffic of the batch =
// 128 B x box visits + 64 B x leaf visits; not meant for timed runs.
HIPRT_API hiprtError hiprtB200TraceGeometryStats(
	hiprtContext	   context,
	hiprtGeometry	   geometry,
	const hiprtRay*	   rays,
	uint32_t		   numRays,
	hiprtHit*		   hits,
	hiprtTraversalType traversalType,
	uint64_t*		   visitCounts,
	hiprtApiStream	   stream );

// ---- multi-GPU plumbing (one process per GPU) ----------------------------------------------------------------------------
// The reference is single-GPU. What a tile-sharded renderer needs from the library, besides tracing: (1) the structure built
// once and replicated -- a built geometry is one relocatable blob: copy `size` bytes (hiprtB200GetGeometrySize) to another
// device / process (ncclBroadcast, cudaMemcpyPeer) and call hiprtB200RelocateGeometry there to rebase the two pointers of its
// header (the reference's save / load does the same rebase through a file, Context.cpp:801-805); (2) a framebuffer on ONE GPU
// that every GPU's resolve kernel writes into directly over NVLink (peer stores from inside the kernel: no staging copy, no
// collective call), with a device-side completion flag.

// Bytes of a built geometry blob (header + box nodes + leaves, as laid out in its storage).
HIPRT_API hiprtError hiprtB200GetGeometrySize( hiprtContext context, hiprtGeometry geometry, size_t* sizeOut );
// Allocates storage for a geometry blob of `size` bytes owned by this context (free with hiprtDestroyGeometry).
HIPRT_API hiprtError hiprtB200AllocateGeometryBlob( hiprtContext context, size_t size, hiprtGeometry* geometryOut );
// After the blob bytes have been copied into `geometry` on `stream`: rebases header pointers to this address (asynchronous).
HIPRT_API hiprtError hiprtB200RelocateGeometry( hiprtContext context, hiprtGeometry geometry, hiprtApiStream stream );

// Device memory that other processes on the same box can map (CUDA IPC). handleOut receives 64 opaque bytes to be sent to the
// peers by any means (torch.distributed, MPI, a pipe); a peer maps it with hiprtB200SharedOpen and unmaps with hiprtB200SharedClose.
HIPRT_API hiprtError hiprtB200SharedAlloc( hiprtContext context, size_t bytes, void** devicePtrOut, void* handleOut /* 64 bytes */ );
HIPRT_API hiprtError hiprtB200SharedFree( hiprtContext context, void* devicePtr );
HIPRT_API hiprtError hiprtB200SharedOpen( hiprtContext context, const void* handle /* 64 bytes */, void** devicePtrOut );
HIPRT_API hiprtError hiprtB200SharedClose( hiprtContext context, void* devicePtr );

// Ambient-occlusion resolve of this GPU's pixels, fused with the framebuffer exchange: for local pixel i (global pixel
// pixelIndex[i]) whose primary record (primaryRecords[i], any hiprtB200HitFormat) is a hit, counts the unoccluded rays among its
// `spp` occlusion records aoPrimIDs[firstAoRay[i] .. +spp) (hiprtB200HitFormatPrimID records: hiprtInvalidValue = unoccluded) and
// stores ao = unoccluded / spp (misses: -1) to framebuffer[pixelIndex[i]] -- `framebuffer` may live on ANOTHER GPU (peer-mapped
// through hiprtB200SharedOpen): the stores then travel over NVLink while the rest of the grid is still resolving. When the
// last block has issued its stores the kernel releases `doneFlag` (also peer-mappable; may be null) by storing `doneValue` at
// system scope. `blockCounter` is one zero-initialised device uint32 (the kernel leaves it at zero again).
// doneValue == 0 selects the counting mode for launches recorded in a CUDA graph (same arguments every replay): `blockCounter`
// then points to TWO zero-initialised words, the second one is incremented by every launch and its new value is what is released.
HIPRT_API hiprtError hiprtB200ResolveAoFrame(
	hiprtContext	   context,
	const void*		   primaryRecords,
	hiprtB200HitFormat primaryFormat,
	const uint32_t*	   aoPrimIDs,
	const uint32_t*	   firstAoRay,
	const uint32_t*	   pixelIndex,
	uint32_t		   numPixels,
	uint32_t		   spp,
	float*			   framebuffer,
	uint32_t*		   doneFlag,
	uint32_t		   doneValue,
	uint32_t*		   blockCounter,
	hiprtApiStream	   stream );

// Blocks `stream` (a spinning one-block kernel, bounded) until the `count` flags flags[k * strideWords] have all reached `value`
// (flag >= value: frame numbers only grow): the owner of the framebuffer waits for every GPU's hiprtB200ResolveAoFrame.
// If the bound is hit (a rank died), *timeoutOut (device uint32, may be null) is set to 1 and the stream continues.
// value == 0 selects the counting mode (CUDA graph replays): `timeoutOut` then points to TWO zero-initialised words, the second one
// is incremented by every launch and its new value is the frame number waited for.
HIPRT_API hiprtError hiprtB200WaitFlags(
	hiprtContext context, const uint32_t* flags, uint32_t count, uint32_t strideWords, uint32_t value, uint32_t* timeoutOut, hiprtApiStream stream );

// Byte offsets of the builder's intermediate arrays inside the caller's temporary buffer
// (parity tests read sorted Morton keys, binary nodes etc. from there). 16 size_t values,
// see hiprt_b200/csrc/builders.h TempLayout.
HIPRT_API hiprtError hiprtB200GetTempLayout( size_t primCount, size_t pairedTriangleCount, size_t* offsetsOut, uint32_t numOffsets );
// The same for a PreferHighQualityBuild of `primCount` primitives: 8 values -- control, state, primRefs, binary nodes, leaf records,
// the reference budget (a count), the collapse queue, total (hiprt_b200/csrc/builders.h highQualityTempLayout).
HIPRT_API hiprtError
hiprtB200GetHighQualityTempLayout( size_t primCount, size_t pairedTriangleCount, int spatialSplits, size_t* offsetsOut, uint32_t numOffsets );

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
