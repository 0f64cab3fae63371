// hiprt_b200 -- public API types, ABI-compatible with the reference's hiprt/hiprt_types.h
// (handles :29-47, enums :52-216, hiprtRay :221, hiprtHit :237, func table :263-307,
// context/build inputs :312-516, stack inputs :521-542, frames :549-586).
// Byte layouts are pinned by the static_asserts below.
#pragma once
#include <hiprt/hiprt_vec.h>

// ---- opaque handles: device pointers to the acceleration-structure header -------------------
struct _hiprtGeometry;
struct _hiprtScene;
struct _hiprtContext;
struct _hiprtFuncTable;

typedef void*			 hiprtDevicePtr;
typedef _hiprtGeometry*	 hiprtGeometry;
typedef _hiprtScene*	 hiprtScene;
typedef _hiprtContext*	 hiprtContext;
typedef _hiprtFuncTable* hiprtFuncTable;
typedef uint32_t		 hiprtLogLevel;
typedef uint32_t		 hiprtBuildFlags;
typedef uint32_t		 hiprtRayMask;
typedef int				 hiprtApiDevice;   // CUdevice ordinal
typedef void*			 hiprtApiCtx;	   // CUcontext (may be null: primary context of `device`)
typedef void*			 hiprtApiStream;   // CUstream / cudaStream_t
typedef void*			 hiprtApiFunction; // CUfunction
typedef void*			 hiprtApiModule;   // CUmodule

// ---- enums ------------------------------------------------------------------------------------
enum : uint32_t
{
	hiprtInvalidValue			   = hiprt::InvalidValue,
	hiprtFullRayMask			   = hiprt::FullRayMask,
	hiprtMaxBatchBuildMaxPrimCount = hiprt::MaxBatchBuildMaxPrimCount,
	hiprtMaxInstanceLevels		   = hiprt::MaxInstanceLevels,
};

enum hiprtError
{
	hiprtSuccess				= 0,
	hiprtErrorNotImplemented	= 1,
	hiprtErrorInternal			= 2,
	hiprtErrorOutOfHostMemory	= 3,
	hiprtErrorOutOfDeviceMemory = 4,
	hiprtErrorInvalidApiVersion = 5,
	hiprtErrorInvalidParameter	= 6
};

enum hiprtTraversalType
{
	hiprtTraversalTerminateAtAnyHit		= 1,
	hiprtTraversalTerminateAtClosestHit = 2
};

enum hiprtTraversalState
{
	hiprtTraversalStateInit,
	hiprtTraversalStateFinished,
	hiprtTraversalStateHit,
	hiprtTraversalStateStackOverflow
};

// Accepted for source compatibility; B200 has no ray-tracing unit to hint.
enum hiprtTraversalHint
{
	hiprtTraversalHintDefault		 = 0,
	hiprtTraversalHintShadowRays	 = 1,
	hiprtTraversalHintReflectionRays = 2
};

enum hiprtLogLevelBits
{
	hiprtLogLevelNone  = 0,
	hiprtLogLevelInfo  = 1 << 0,
	hiprtLogLevelWarn  = 1 << 1,
	hiprtLogLevelError = 1 << 2
};

enum hiprtBuildOperation
{
	hiprtBuildOperationBuild  = 1,
	hiprtBuildOperationUpdate = 2
};

// Low two bits select the builder, the rest are modifier bits.
enum hiprtBuildFlagBits
{
	hiprtBuildFlagBitPreferFastBuild			  = 0, // LBVH
	hiprtBuildFlagBitPreferBalancedBuild		  = 1, // H-PLOC
	hiprtBuildFlagBitPreferHighQualityBuild		  = 2, // routed to H-PLOC in this build
	hiprtBuildFlagBitCustomBvhImport			  = 3,
	hiprtBuildFlagBitDisableSpatialSplits		  = 1 << 2,
	hiprtBuildFlagBitDisableTrianglePairing		  = 1 << 3,
	hiprtBuildFlagBitDisableOrientedBoundingBoxes = 1 << 4
};

enum hiprtPrimitiveType
{
	hiprtPrimitiveTypeTriangleMesh,
	hiprtPrimitiveTypeAABBList
};

enum hiprtInstanceType
{
	hiprtInstanceTypeGeometry,
	hiprtInstanceTypeScene
};

enum hiprtPrimitiveNodeType
{
	hiprtTriangleNode = 0,
	hiprtCustomNode	  = 1
};

enum hiprtFrameType
{
	hiprtFrameTypeSRT,
	hiprtFrameTypeMatrix
};

enum hiprtStackType
{
	hiprtStackTypeGlobal,
	hiprtStackTypeDynamic
};

enum hiprtStackEntryType
{
	hiprtStackEntryTypeInteger,
	hiprtStackEntryTypeInstance
};

enum hiprtBvhNodeType
{
	hiprtBvhNodeTypeInternal = 0,
	hiprtBvhNodeTypeLeaf	 = 1
};

enum hiprtDeviceType
{
	hiprtDeviceAMD,
	hiprtDeviceNVIDIA
};

// ---- rays and hits -----------------------------------------------------------------------------
struct alignas( 16 ) hiprtRay
{
	hiprtFloat3 origin;
	float		minT = 0.0f;
	hiprtFloat3 direction;
	float		maxT = hiprt::FltMax;
};
HIPRT_STATIC_ASSERT( sizeof( hiprtRay ) == 32 );

struct alignas( 16 ) hiprtHit
{
	union
	{
		uint32_t instanceID = hiprtInvalidValue;	 // single-level instancing
		uint32_t instanceIDs[hiprtMaxInstanceLevels]; // multi-level instancing
	};
	uint32_t	primID = hiprtInvalidValue;
	hiprtFloat2 uv;		// barycentrics in the input vertex order
	hiprtFloat3 normal; // geometric normal, object space, not normalised
	float		t = -1.0f;

	HIPRT_HOST_DEVICE bool hasHit() const { return primID != hiprtInvalidValue; }
};
HIPRT_STATIC_ASSERT( sizeof( hiprtHit ) == 48 );

// ---- custom intersection / filter function table ----------------------------------------------
struct hiprtFuncDataSet
{
	const void* intersectFuncData = nullptr;
	const void* filterFuncData	  = nullptr;
};

struct hiprtFuncTableHeader
{
	uint32_t		  numGeomTypes;
	uint32_t		  numRayTypes;
	hiprtFuncDataSet* funcDataSets;
};

struct hiprtFuncNameSet
{
	const char* intersectFuncName = nullptr;
	const char* filterFuncName	  = nullptr;
};

// ---- traversal stacks -----------------------------------------------------------------------------
struct hiprtGlobalStackBuffer
{
	uint32_t stackSize;
	uint32_t stackCount;
	void*	 stackData;
};

struct hiprtSharedStackBuffer
{
	uint32_t stackSize;
	void*	 stackData;
};

struct hiprtGlobalStackBufferInput
{
	hiprtStackType		type	  = hiprtStackTypeGlobal;
	hiprtStackEntryType entryType = hiprtStackEntryTypeInteger;
	uint32_t			stackSize;
	uint32_t			threadCount; // hiprtStackTypeGlobal only
};

struct hiprtInstanceStackEntry
{
	hiprtRay   ray;
	hiprtScene scene;
};

// ---- context and build inputs -------------------------------------------------------------------
struct hiprtContextCreationInput
{
	hiprtApiCtx		ctxt;
	hiprtApiDevice	device;
	hiprtDeviceType deviceType;
};

struct hiprtBuildOptions
{
	hiprtBuildFlags buildFlags;
	uint32_t		batchBuildMaxPrimCount = 0u; // 0 disables the batch builder
};
HIPRT_STATIC_ASSERT( sizeof( hiprtBuildOptions ) == 8 );

// Indexed triangle mesh in device memory; without indices triangle t uses vertices 3t..3t+2.
struct hiprtTriangleMeshPrimitive
{
	hiprtDevicePtr vertices;
	uint32_t	   vertexCount;
	uint32_t	   vertexStride;
	hiprtDevicePtr triangleIndices	   = nullptr;
	uint32_t	   triangleCount	   = 0u;
	uint32_t	   triangleStride	   = 0u;
	hiprtDevicePtr trianglePairIndices = nullptr; // optional uint2 per pair
	uint32_t	   trianglePairCount   = 0u;
};
HIPRT_STATIC_ASSERT( sizeof( hiprtTriangleMeshPrimitive ) == 48 );

// Bounding boxes of custom primitives: (min,max) as 2 x float3 or 2 x float4.
struct hiprtAABBListPrimitive
{
	hiprtDevicePtr aabbs;
	uint32_t	   aabbCount;
	uint32_t	   aabbStride;
};

struct alignas( 64 ) hiprtInternalNode
{
	hiprtFloat3		 aabbMin;
	hiprtFloat3		 aabbMax;
	uint32_t		 childIndices[2];
	hiprtBvhNodeType childNodeTypes[2];
};
HIPRT_STATIC_ASSERT( sizeof( hiprtInternalNode ) == 64 );

struct alignas( 32 ) hiprtLeafNode
{
	hiprtFloat3 aabbMin;
	hiprtFloat3 aabbMax;
	uint32_t	primID;
};
HIPRT_STATIC_ASSERT( sizeof( hiprtLeafNode ) == 32 );

struct hiprtBvhNodeList
{
	hiprtDevicePtr internalNodes;
	hiprtDevicePtr leafNodes;
	uint32_t	   nodeCount;
};

struct alignas( 64 ) hiprtGeometryBuildInput
{
	hiprtPrimitiveType type;
	uint32_t		   geomType = hiprtInvalidValue; // row of the function table
	union
	{
		hiprtTriangleMeshPrimitive triangleMesh;
		hiprtAABBListPrimitive	   aabbList;
	} primitive{};
	hiprtBvhNodeList nodeList;
};
HIPRT_STATIC_ASSERT( sizeof( hiprtGeometryBuildInput ) == 128 );

struct alignas( 16 ) hiprtInstance
{
	hiprtInstanceType type;
	union
	{
		hiprtGeometry geometry;
		hiprtScene	  scene;
	};
};
HIPRT_STATIC_ASSERT( sizeof( hiprtInstance ) == 16 );

struct alignas( 16 ) hiprtSceneBuildInput
{
	hiprtDevicePtr	 instances;				   // hiprtInstance[instanceCount]
	hiprtDevicePtr	 instanceTransformHeaders; // optional hiprtTransformHeader[instanceCount]
	hiprtDevicePtr	 instanceFrames;		   // hiprtFrameSRT / hiprtFrameMatrix [frameCount]
	hiprtDevicePtr	 instanceMasks;			   // optional uint32_t[instanceCount]
	hiprtBvhNodeList nodeList;
	uint32_t		 instanceCount;
	uint32_t		 frameCount;
	hiprtFrameType	 frameType = hiprtFrameTypeSRT;
};
HIPRT_STATIC_ASSERT( sizeof( hiprtSceneBuildInput ) == 80 );

// ---- transformation frames: object-to-world = T * R * S -------------------------------------------
struct alignas( 16 ) hiprtFrameSRT
{
	hiprtFloat4 rotation; // axis.xyz, angle
	hiprtFloat3 scale;
	hiprtFloat3 translation;
	float		time;
};
HIPRT_STATIC_ASSERT( sizeof( hiprtFrameSRT ) == 48 );

struct alignas( 64 ) hiprtFrameMatrix
{
	float matrix[3][4];
	float time;
};
HIPRT_STATIC_ASSERT( sizeof( hiprtFrameMatrix ) == 64 );

struct alignas( 8 ) hiprtTransformHeader
{
	uint32_t frameIndex;
	uint32_t frameCount;
};
HIPRT_STATIC_ASSERT( sizeof( hiprtTransformHeader ) == 8 );
