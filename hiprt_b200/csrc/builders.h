// Host-side entry points of the CUDA builders (implemented in build_lbvh.cu & friends).
// The API shim (api.cpp) validates arguments and calls these; nothing here synchronises
// with the host: every build is a fixed sequence of launches on the caller's stream.
#pragma once
#include <cuda_runtime.h>
#include <hiprt/hiprt_b200_layout.h>
#include <hiprt/hiprt_types.h>

namespace hiprt
{
// Optional per-phase timing of a build (the reference has a compile-time Timer around every
// phase, hiprt/impl/Timer.h:41, LbvhBuilder.h:411-429). Events are recorded on the build stream,
// nothing synchronises until the profile is read.
struct BuildProfile
{
	static constexpr int MaxMarks = 20;
	cudaEvent_t			 events[MaxMarks];
	const char*			 names[MaxMarks]; // names[i] = phase that ENDS at events[i]; events[0] is the start
	int					 count	 = 0;
	bool				 enabled = false;
};

struct DeviceInfo
{
	int			  device  = 0;
	uint32_t	  smCount = 148;
	BuildProfile* profile = nullptr;

	void mark( const char* phaseJustFinished, cudaStream_t stream ) const
	{
		if ( !profile || !profile->enabled || profile->count >= BuildProfile::MaxMarks ) return;
		profile->names[profile->count] = phaseJustFinished;
		cudaEventRecord( profile->events[profile->count], stream );
		++profile->count;
	}
	void beginProfile( cudaStream_t stream ) const
	{
		if ( !profile || !profile->enabled ) return;
		profile->count = 0;
		mark( "start", stream );
	}
};

// ---- sizes -------------------------------------------------------------------------------------
size_t geometryPrimCount( const hiprtGeometryBuildInput& input ); // triangles (or pairs if given) / AABBs
// (the options matter: a high-quality build with spatial splits may create up to 1.5 leaves per primitive)
size_t geometryStorageSize( const hiprtGeometryBuildInput& input, hiprtBuildOptions options );	  // the public structure: header.size
size_t geometryAllocationSize( const hiprtGeometryBuildInput& input, hiprtBuildOptions options ); // + the traversal-only Q region of triangle meshes (qnode.h)
void*  geometryQRegion( const hiprtGeometryBuildInput& input, hiprtBuildOptions options, void* storage ); // nullptr: this build has none
size_t geometryTempSize( const hiprtGeometryBuildInput& input, hiprtBuildOptions options );
size_t sceneStorageSize( const hiprtSceneBuildInput& input, hiprtBuildOptions options );
size_t sceneTempSize( const hiprtSceneBuildInput& input, hiprtBuildOptions options );

// ---- builds (asynchronous on `stream`) -----------------------------------------------------------
// qRegion: geometryQRegion() of the same input (the Q records are generated as the last step), or nullptr
void buildGeometry(
	const DeviceInfo& dev, const hiprtGeometryBuildInput& input, hiprtBuildOptions options, void* temp, cudaStream_t stream, void* storage,
	void* qRegion = nullptr );
void buildScene(
	const DeviceInfo& dev, const hiprtSceneBuildInput& input, hiprtBuildOptions options, void* temp, cudaStream_t stream, void* storage );
// refit in place (hiprtBuildOperationUpdate)
void updateGeometry(
	const DeviceInfo& dev, const hiprtGeometryBuildInput& input, hiprtBuildOptions options, void* temp, cudaStream_t stream, void* storage,
	void* qRegion = nullptr );
void updateScene(
	const DeviceInfo& dev, const hiprtSceneBuildInput& input, hiprtBuildOptions options, void* temp, cudaStream_t stream, void* storage );
// many small geometries in one launch (one CTA per geometry)
size_t batchTempSize( uint32_t numGeometries );
void   buildGeometriesBatch(
	  const DeviceInfo&				 dev,
	  const hiprtGeometryBuildInput* inputs,
	  uint32_t						 numGeometries,
	  hiprtBuildOptions				 options,
	  void*							 temp,
	  cudaStream_t					 stream,
	  void* const*					 storages );

// ---- introspection for the parity tests: where the intermediate arrays of the last build of
// `primCount` primitives live inside the caller's temporary buffer (byte offsets).
struct TempLayout
{
	size_t control;	   // BuildControl
	size_t pairs;	   // uint2[triangles]          (triangle meshes that are paired)
	size_t pairScan;   // look-back words of the pairing scan
	size_t primRefs;   // 32 B x N  {box, primIndex} in input order
	size_t keys[2];	   // u32 x N ping-pong
	size_t vals[2];	   // u32 x N ping-pong
	size_t sortScratch;
	size_t sortedRefs; // 32 B x N  {box, primIndex} in sorted order
	size_t scratch;	   // 32 B x N  binary nodes
	size_t flags;	   // u32 x N
	size_t clusterIds; // u32 x N  (H-PLOC)
	size_t taskQueue;  // u64 x maxBoxNodes (aliases keys/vals)
	size_t total;
};
TempLayout tempLayout( size_t primCount, size_t triangleCountForPairing );
// the same for a high-quality build (sbvh_kernels.cuh): control, state (sbvh::State), primRefs, nodes (binary nodes, 32 B),
// leaves (leaf records, 32 B), maxRefs (a count, not an offset), taskQueue, total
void highQualityTempLayout( size_t primCount, size_t triangleCountForPairing, bool spatialSplits, size_t out[8] );

struct BuildControl
{
	Box		 sceneBox;		 // union of primitive boxes (atomics)
	uint32_t primCount;		 // N after pairing
	uint32_t rootNode;		 // binary root id
	uint32_t queueHead;		 // collapse: next task to claim
	uint32_t pairTileCounter;
	uint32_t nodeCounter;	 // H-PLOC: binary nodes allocated so far
	uint32_t pad[5];
};
static_assert( sizeof( BuildControl ) == 64, "BuildControl is one 64-byte line" );
} // namespace hiprt
