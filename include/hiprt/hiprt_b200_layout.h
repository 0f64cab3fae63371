// hiprt_b200 -- in-memory format of a built geometry / scene (what a hiprtGeometry or
// hiprtScene handle points to). The format is byte-compatible with the reference's
// NVIDIA path (RTIP 0): GeomHeader/SceneHeader hiprt/impl/Header.h:33-70, Box4Node
// BvhNode.h:145-276, TrianglePairNode :682-703, CustomNode :1103-1107, UserInstanceNode
// :1109-1182, Instance hiprt/impl/Instance.h:32-43, SRT frame Transform.h:129-240,
// ScratchNode/ReferenceNode BvhNode.h:1254-1316, node index encoding :32-103.
// Keeping it compatible lets a structure built here be traversed by the reference's
// kernels and vice versa, which is how the parity tests A/B the two builders.
//
// HBM layout of one geometry (64-byte aligned bump allocation, LbvhBuilder.h:152-154):
//   [GeomHeader 64 B][Box4Node x maxBoxNodes, 128 B each][leaf nodes x primCount]
// and of one scene:
//   [SceneHeader 64 B][Box4Node ...][InstanceNode x I, 64 B][InstanceRecord x I, 16 B][Frame x F, 64 B]
#pragma once
#include <hiprt/hiprt_types.h>

namespace hiprt
{
enum NodeType : uint32_t
{
	TriangleType = 0, // leaf: TrianglePairNode (type 1 = second triangle of the pair)
	BoxType		 = 5, // internal: Box4Node
	InstanceType = 6, // leaf of a scene: InstanceNode
	CustomType	 = 7  // leaf: CustomNode
};

constexpr uint32_t RootIndex			= BoxType; // box node 0
constexpr uint32_t DefaultTriangleFlags = ( 2u << 2 ) | ( 1u << 0 );

// Node index = (address << shift) | type; box and instance nodes use shift 4, triangle and
// custom nodes shift 3 (the reference addresses nodes in 64-byte units on AMD hardware).
HIPRT_HOST_DEVICE HIPRT_INLINE uint32_t nodeType( uint32_t index ) { return index & 7u; }
HIPRT_HOST_DEVICE HIPRT_INLINE uint32_t nodeAddr( uint32_t index )
{
	const uint32_t t = index & 7u;
	return index >> ( ( t == BoxType || t == InstanceType ) ? 4 : 3 );
}
HIPRT_HOST_DEVICE HIPRT_INLINE uint32_t makeNodeIndex( uint32_t addr, uint32_t type )
{
	return ( addr << ( ( type == BoxType || type == InstanceType ) ? 4 : 3 ) ) | type;
}
HIPRT_HOST_DEVICE HIPRT_INLINE bool isBoxIndex( uint32_t index ) { return ( index & 7u ) == BoxType; }
HIPRT_HOST_DEVICE HIPRT_INLINE bool isLeafIndex( uint32_t index )
{
	return ( index & 7u ) != BoxType && index != InvalidValue;
}

struct Box
{
	hiprtFloat3 lo;
	hiprtFloat3 hi;
};
HIPRT_STATIC_ASSERT( sizeof( Box ) == 24 );

struct alignas( 64 ) Box4Node
{
	uint32_t child[4]; // node indices, InvalidValue for unused slots
	Box		 box[4];   // child bounds; unused slots hold the empty box (+FltMax, -FltMax)
	uint32_t parentAddr;
	uint32_t updateCounter; // scratch for refit
	uint32_t childCount;
	uint32_t pad;
};
HIPRT_STATIC_ASSERT( sizeof( Box4Node ) == 128 );

// Two triangles sharing the edge (v1,v2): triangle 0 = (v0,v1,v2), triangle 1 = (v1,v3,v2).
// flags byte k (k = triangle): bits 0-1 / 2-3 select which of (1-u-v, u, v) is reported as
// uv.x / uv.y so that barycentrics follow the caller's vertex order; bit 5 flips the normal.
struct alignas( 64 ) TrianglePairNode
{
	hiprtFloat3 v0, v1, v2, v3;
	uint32_t	pad;
	uint32_t	primIndex0;
	uint32_t	primIndex1; // == primIndex0 for a lone triangle
	uint32_t	flags;
};
HIPRT_STATIC_ASSERT( sizeof( TrianglePairNode ) == 64 );

struct CustomNode
{
	uint32_t primIndex;
};

struct GeomHeader;
struct SceneHeader;

struct alignas( 64 ) InstanceNode
{
	union
	{
		float				 invMatrix[3][4]; // world -> object, static instances
		hiprtTransformHeader transform;		  // frame range, animated instances
	};
	union
	{
		GeomHeader*	 geometry;
		SceneHeader* scene;
	};
	uint32_t mask;
	uint32_t primIndex : 24;
	uint32_t type : 1;	   // hiprtInstanceType
	uint32_t isStatic : 1; // exactly one frame
	uint32_t identity : 1; // transform is the identity: ray is not transformed
	uint32_t : 5;
};
HIPRT_STATIC_ASSERT( sizeof( InstanceNode ) == 64 );

struct alignas( 16 ) InstanceRecord
{
	uint32_t frameIndex;
	uint32_t frameCount : 31;
	uint32_t type : 1;
	union
	{
		GeomHeader*	 geometry;
		SceneHeader* scene;
	};
};
HIPRT_STATIC_ASSERT( sizeof( InstanceRecord ) == 16 );

struct alignas( 64 ) SRTFrame
{
	hiprtFloat4 rotation; // unit quaternion (x,y,z,w)
	hiprtFloat3 scale;
	hiprtFloat3 shear;
	hiprtFloat3 translation;
	float		time;
};
HIPRT_STATIC_ASSERT( sizeof( SRTFrame ) == 64 );

struct GeomHeader
{
	Box4Node* boxNodes;
	void*	  primNodes;
	size_t	  size;
	uint32_t  referenceCount;
	uint32_t  boxNodeCount;
	uint32_t  primNodeCount;
	uint32_t  geomType; // (user geomType << 1) | isTriangleMesh
	uint32_t  rtip;		// always 0 here
};
HIPRT_STATIC_ASSERT( sizeof( GeomHeader ) == 48 );

struct SceneHeader
{
	Box4Node*		boxNodes;
	InstanceNode*	primNodes;
	InstanceRecord* instances;
	SRTFrame*		frames;
	size_t			size;
	uint32_t		referenceCount;
	uint32_t		primCount;
	uint32_t		primNodeCount;
	uint32_t		boxNodeCount;
	uint32_t		frameCount;
	uint32_t		rtip;
};
HIPRT_STATIC_ASSERT( sizeof( SceneHeader ) == 64 );

// Build-time scratch: binary LBVH / H-PLOC node and sorted primitive reference.
struct alignas( 32 ) ScratchNode
{
	Box		 box;
	uint32_t child0;
	uint32_t child1;
};
HIPRT_STATIC_ASSERT( sizeof( ScratchNode ) == 32 );

struct alignas( 32 ) ReferenceNode
{
	Box		 box;
	uint32_t primIndex;
	uint32_t pad;
};
HIPRT_STATIC_ASSERT( sizeof( ReferenceNode ) == 32 );

// Worst-case node counts (reference: hiprt/impl/BvhCommon.h:38-44,63-75).
HIPRT_HOST_DEVICE HIPRT_INLINE size_t maxBoxNodeCount( size_t primCount )
{
	if ( primCount <= BranchingFactor ) return 1;
	const size_t leaves = primCount / 2 + ( primCount & 1 );
	return leaves + 1 + ( leaves + BranchingFactor - 2 ) / ( BranchingFactor - 1 );
}
} // namespace hiprt
