// ORACLE INFRASTRUCTURE -- the reference's own arithmetic, run on the CPU.
//
// The parity-critical functions of the reference (Morton codes, the slab test, the Moller-Trumbore test, frame decomposition /
// inversion / key-frame interpolation) are HIPRT_HOST_DEVICE code in its headers. This file includes those headers WHERE THEY LIE
// under /root/reference (nothing is copied) and exposes them through a C interface, so that the oracle's restatement can be held
// to the reference's code itself, bit for bit, without a GPU: tests/test_oracle_vs_reference_math.py, and the vectors of
// tests/golden/ref_math.npz (tests/golden/make_ref_math.py) for machines where /root/reference does not exist.
// Built by oracle/ref/Makefile into oracle/_ref/libref_host_math.so.
#include <hiprt/hiprt_types.h>
#include <hiprt/hiprt_math.h>
#include <hiprt/impl/Aabb.h>
#include <hiprt/impl/MortonCode.h>
#include <hiprt/impl/Transform.h>
#include <hiprt/impl/Triangle.h>
#include <hiprt/impl/TriangleMesh.h>

#include <cstdint>
#include <cstring>
#include <vector>

using namespace hiprt;

static float3 f3( const float* p ) { return float3{ p[0], p[1], p[2] }; }

static hiprtFrameSRT srtOf( const float* s )
{
	hiprtFrameSRT f;
	f.rotation	  = float4{ s[0], s[1], s[2], s[3] };
	f.scale		  = f3( s + 4 );
	f.translation = f3( s + 7 );
	f.time		  = s[10];
	return f;
}

static hiprtFrameMatrix matrixOf( const float* m )
{
	hiprtFrameMatrix f;
	for ( int i = 0; i < 3; ++i )
		for ( int j = 0; j < 4; ++j )
			f.matrix[i][j] = m[4 * i + j];
	f.time = m[12];
	return f;
}

static SRTFrame frameOf( const float* data, int frameType ) { return frameType == 1 ? SRTFrame( matrixOf( data ) ) : SRTFrame( srtOf( data ) ); }

extern "C" {

// MortonCode.h:60-277
void ref_morton( uint32_t n, const float* normalizedPos, const float* sceneExtent, uint32_t* out )
{
	for ( uint32_t i = 0; i < n; ++i )
		out[i] = computeExtendedMortonCode( f3( normalizedPos + 3 * i ), f3( sceneExtent ) );
}

// Triangle.h:96-130. rays: hiprtRay records (32 B). hit[i] = 0 / 1.
void ref_triangle( uint32_t n, const float* vertices, const hiprtRay* rays, const uint32_t* flags, uint8_t* hit, float* t, float* uv )
{
	for ( uint32_t i = 0; i < n; ++i )
	{
		const Triangle tri( f3( vertices + 9 * i ), f3( vertices + 9 * i + 3 ), f3( vertices + 9 * i + 6 ) );
		float2		   b{ 0.0f, 0.0f };
		float		   tt = 0.0f;
		hit[i]			  = tri.intersect( rays[i], b, tt, flags[i] ) ? 1 : 0;
		t[i] = tt, uv[2 * i] = b.x, uv[2 * i + 1] = b.y;
	}
}

// Aabb.h:92-101
void ref_slab( uint32_t n, const float* boxes, const float* origins, const float* invDirections, const float* maxT, float* t01 )
{
	for ( uint32_t i = 0; i < n; ++i )
	{
		Aabb		 box( f3( boxes + 6 * i ), f3( boxes + 6 * i + 3 ) );
		const float2 r = box.intersect( f3( origins + 3 * i ), f3( invDirections + 3 * i ), maxT[i] );
		t01[2 * i] = r.x, t01[2 * i + 1] = r.y;
	}
}

// Transform.h:61-127,129-176,369-395: API frame -> scale / shear / rotation / translation, and the world -> object matrix the
// scene builder stores in an instance node. frameType 0: hiprtFrameSRT (12 floats), 1: hiprtFrameMatrix (16 floats).
// srtOut: rotation quaternion (4), scale (3), shear (3), translation (3) = 13 floats per frame; invOut: 12 floats per frame.
void ref_frames( uint32_t n, const float* frames, int frameType, float* srtOut, float* invOut )
{
	const uint32_t stride = frameType == 1 ? 16 : 12;
	for ( uint32_t i = 0; i < n; ++i )
	{
		const SRTFrame f = frameOf( frames + stride * i, frameType );
		float*		   o = srtOut + 13 * i;
		o[0] = f.m_rotation.x, o[1] = f.m_rotation.y, o[2] = f.m_rotation.z, o[3] = f.m_rotation.w;
		o[4] = f.m_scale.x, o[5] = f.m_scale.y, o[6] = f.m_scale.z;
		o[7] = f.m_shear.x, o[8] = f.m_shear.y, o[9] = f.m_shear.z;
		o[10] = f.m_translation.x, o[11] = f.m_translation.y, o[12] = f.m_translation.z;
		float inv[3][4];
		computeInvTransformMatrix( f, inv );
		for ( int r = 0; r < 3; ++r )
			for ( int c = 0; c < 4; ++c )
				invOut[12 * i + 4 * r + c] = inv[r][c];
	}
}

// TriangleMesh.h:32-50,89-157 and BvhNode.h:682-703: the 64-byte leaf of a triangle pair (the four vertices after the rotation that
// puts the shared edge at [1,2], both primitive indices, the flags that undo the rotation in the reported uv and normal) and its box.
// vertices: float3 stride 12; triangles: uint3 stride 12; pairs: uint2 per node (x == y: single triangle).
void ref_pair_nodes( const float* vertices, uint32_t vertexCount, const uint32_t* triangles, uint32_t triangleCount, const uint32_t* pairs, uint32_t n, void* nodesOut, float* boxesOut )
{
	hiprtTriangleMeshPrimitive prim{};
	prim.vertices		 = const_cast<float*>( vertices );
	prim.vertexCount	 = vertexCount;
	prim.vertexStride	 = 12;
	prim.triangleIndices = const_cast<uint32_t*>( triangles );
	prim.triangleCount	 = triangleCount;
	prim.triangleStride	 = 12;
	const TriangleMesh mesh( prim );
	for ( uint32_t i = 0; i < n; ++i )
	{
		const TrianglePairNode node = mesh.fetchTrianglePairNode( uint2{ pairs[2 * i], pairs[2 * i + 1] } );
		memcpy( static_cast<uint8_t*>( nodesOut ) + 64 * size_t( i ), &node, 64 );
		const Aabb box = node.aabb();
		boxesOut[6 * i + 0] = box.m_min.x, boxesOut[6 * i + 1] = box.m_min.y, boxesOut[6 * i + 2] = box.m_min.z;
		boxesOut[6 * i + 3] = box.m_max.x, boxesOut[6 * i + 4] = box.m_max.y, boxesOut[6 * i + 5] = box.m_max.z;
	}
}

// Transform.h:460-517: world-space bounds of a box under the key frames of one instance (every key frame and three interpolated
// poses between consecutive keys): what the scene builder bounds an instance with (InstanceList.h:104-121).
void ref_motion_bounds( uint32_t frameCount, const float* frames, int frameType, uint32_t n, const float* boxesIn, float* boxesOut )
{
	const uint32_t		  stride = frameType == 1 ? 16 : 12;
	std::vector<SRTFrame> keys( frameCount );
	for ( uint32_t k = 0; k < frameCount; ++k )
		keys[k] = frameOf( frames + stride * k, frameType );
	const Transform transform( reinterpret_cast<const Frame*>( keys.data() ), 0, frameCount );
	for ( uint32_t i = 0; i < n; ++i )
	{
		const Aabb out = transform.motionBounds( Aabb( f3( boxesIn + 6 * i ), f3( boxesIn + 6 * i + 3 ) ) );
		boxesOut[6 * i + 0] = out.m_min.x, boxesOut[6 * i + 1] = out.m_min.y, boxesOut[6 * i + 2] = out.m_min.z;
		boxesOut[6 * i + 3] = out.m_max.x, boxesOut[6 * i + 4] = out.m_max.y, boxesOut[6 * i + 5] = out.m_max.z;
	}
}

// Transform.h:188-230 under the pose at `time`: per input vector p: transform( p ), invTransform( p ), transformVector( p ),
// invTransformVector( p ) (what hiprtPoint/VectorObjectToWorld/WorldToObject return)
void ref_motion_points( uint32_t frameCount, const float* frames, int frameType, float time, uint32_t n, const float* in, float* out )
{
	const uint32_t		  stride = frameType == 1 ? 16 : 12;
	std::vector<SRTFrame> keys( frameCount );
	for ( uint32_t k = 0; k < frameCount; ++k )
		keys[k] = frameOf( frames + stride * k, frameType );
	const Transform transform( reinterpret_cast<const Frame*>( keys.data() ), 0, frameCount );
	const SRTFrame	f = transform.interpolateFrames( time );
	for ( uint32_t i = 0; i < n; ++i )
	{
		const float3 p	  = f3( in + 3 * i );
		const float3 r[4] = { f.transform( p ), f.invTransform( p ), f.transformVector( p ), f.invTransformVector( p ) };
		for ( int k = 0; k < 4; ++k )
			out[12 * i + 3 * k] = r[k].x, out[12 * i + 3 * k + 1] = r[k].y, out[12 * i + 3 * k + 2] = r[k].z;
	}
}

// Transform.h:410-452: the key frames of one instance, the ray carried into the instance at `time` (motion blur).
void ref_motion_ray( uint32_t frameCount, const float* frames, int frameType, uint32_t n, const float* times, const hiprtRay* raysIn, hiprtRay* raysOut )
{
	const uint32_t		  stride = frameType == 1 ? 16 : 12;
	std::vector<SRTFrame> keys( frameCount );
	for ( uint32_t k = 0; k < frameCount; ++k )
		keys[k] = frameOf( frames + stride * k, frameType );
	const Transform transform( reinterpret_cast<const Frame*>( keys.data() ), 0, frameCount );
	for ( uint32_t i = 0; i < n; ++i )
		raysOut[i] = transform.transformRay( raysIn[i], times[i] );
}
}
