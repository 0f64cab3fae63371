// hiprt_b200 -- key-frame (motion blur) transforms on the device: the internal SRT frame of an instance
// (scale / shear / rotation quaternion / translation / time, 64 bytes in the scene storage), its forward and
// inverse application to points and vectors, and the interpolation between key frames.
// Restates the reference's hiprt/impl/Transform.h:129-240 (SRTFrame), :410-517 (Transform::interpolateFrames,
// transformRay, boundPointMotion) and hiprt/impl/Quaternion.h:110-190 (qtMul, qtRotate, qtMix), operation by operation.
// Used by the traversal objects (animated instances, hiprt_device.h), by the transform query functions
// hiprtPoint/VectorObjectToWorld/WorldToObject, and by the scene builder for the motion bounds of an instance.
#pragma once
#include <hiprt/hiprt_b200_layout.h>

#if defined( __CUDACC__ )
namespace hiprt
{
namespace motion
{
struct Frame
{
	float4 rotation; // unit quaternion (x, y, z, w)
	float3 scale, shear, translation;
	float  time;
};

HIPRT_DEVICE HIPRT_INLINE Frame identityFrame()
{
	return Frame{ make_float4( 0.0f, 0.0f, 0.0f, 1.0f ), make_float3( 1.0f, 1.0f, 1.0f ), make_float3( 0.0f, 0.0f, 0.0f ), make_float3( 0.0f, 0.0f, 0.0f ), 0.0f };
}

HIPRT_DEVICE HIPRT_INLINE Frame loadKeyFrame( const SRTFrame* frames, uint32_t i )
{
	const float4* p = reinterpret_cast<const float4*>( frames + i );
	const float4  a = p[0], b = p[1], c = p[2], d = p[3];
	Frame		  f;
	f.rotation	  = a;
	f.scale		  = make_float3( b.x, b.y, b.z );
	f.shear		  = make_float3( b.w, c.x, c.y );
	f.translation = make_float3( c.z, c.w, d.x );
	f.time		  = d.y;
	return f;
}

// identitySRT (Transform.h:41-50)
HIPRT_DEVICE HIPRT_INLINE bool isIdentity( const Frame& f )
{
	return f.scale.x == 1.0f && f.scale.y == 1.0f && f.scale.z == 1.0f && f.shear.x == 0.0f && f.shear.y == 0.0f && f.shear.z == 0.0f &&
		   f.translation.x == 0.0f && f.translation.y == 0.0f && f.translation.z == 0.0f && f.rotation.w == 1.0f;
}

// (the compensated products of hiprt_math.h:1571-1603, as in the traversal arithmetic)
HIPRT_DEVICE HIPRT_INLINE float sumProd( float a, float b, float c, float d )
{
	const float cd = __fmul_rn( c, d );
	return __fadd_rn( fmaf( a, b, cd ), fmaf( c, d, -cd ) );
}
HIPRT_DEVICE HIPRT_INLINE float diffProd( float a, float b, float c, float d )
{
	const float cd = __fmul_rn( c, d );
	return __fsub_rn( fmaf( a, b, -cd ), fmaf( c, d, -cd ) );
}
HIPRT_DEVICE HIPRT_INLINE float3 cross( const float3 a, const float3 b )
{
	return make_float3( diffProd( a.y, b.z, a.z, b.y ), diffProd( a.z, b.x, a.x, b.z ), diffProd( a.x, b.y, a.y, b.x ) );
}
HIPRT_DEVICE HIPRT_INLINE float dot( const float3 a, const float3 b ) { return fmaf( a.x, b.x, sumProd( a.y, b.y, a.z, b.z ) ); }

HIPRT_DEVICE HIPRT_INLINE float4 qtMul( const float4 a, const float4 b )
{
	const float3 av = make_float3( a.x, a.y, a.z ), bv = make_float3( b.x, b.y, b.z );
	const float3 c	= cross( av, bv );
	return make_float4(
		__fadd_rn( __fadd_rn( c.x, __fmul_rn( a.w, b.x ) ), __fmul_rn( b.w, a.x ) ),
		__fadd_rn( __fadd_rn( c.y, __fmul_rn( a.w, b.y ) ), __fmul_rn( b.w, a.y ) ),
		__fadd_rn( __fadd_rn( c.z, __fmul_rn( a.w, b.z ) ), __fmul_rn( b.w, a.z ) ),
		__fsub_rn( __fmul_rn( a.w, b.w ), dot( av, bv ) ) );
}
HIPRT_DEVICE HIPRT_INLINE float4 qtInvert( const float4 q ) { return make_float4( -q.x, -q.y, -q.z, q.w ); }
HIPRT_DEVICE HIPRT_INLINE float3 qtRotate( const float4 q, const float3 p )
{
	const float4 r = qtMul( qtMul( q, make_float4( p.x, p.y, p.z, 0.0f ) ), qtInvert( q ) );
	return make_float3( r.x, r.y, r.z );
}
HIPRT_DEVICE HIPRT_INLINE float3 qtInvRotate( const float4 q, const float3 p ) { return qtRotate( qtInvert( q ), p ); }
HIPRT_DEVICE HIPRT_INLINE float	 qtDot( const float4 a, const float4 b ) { return fmaf( a.x, b.x, fmaf( a.y, b.y, sumProd( a.z, b.z, a.w, b.w ) ) ); }
HIPRT_DEVICE HIPRT_INLINE float4 qtNormalize( const float4 q )
{
	const float len = sqrtf( qtDot( q, q ) );
	return make_float4( q.x / len, q.y / len, q.z / len, q.w / len );
}

// shortest-path slerp, linear + renormalise when the inputs nearly coincide (Quaternion.h:136-188)
HIPRT_DEVICE HIPRT_INLINE float4 qtMix( float4 v0, float4 v1, float t )
{
	v0		= qtNormalize( v0 );
	v1		= qtNormalize( v1 );
	float d = qtDot( v0, v1 );
	if ( d < 0.0f )
	{
		v1 = make_float4( -v1.x, -v1.y, -v1.z, -v1.w );
		d  = -d;
	}
	if ( d > 0.9995f )
		return qtNormalize( make_float4( v0.x + ( v1.x - v0.x ) * t, v0.y + ( v1.y - v0.y ) * t, v0.z + ( v1.z - v0.z ) * t, v0.w + ( v1.w - v0.w ) * t ) );
	const float theta0 = acosf( d );
	const float theta  = theta0 * t;
	const float sinT = sinf( theta ), sinT0 = sinf( theta0 );
	const float s0 = cosf( theta ) - d * sinT / sinT0;
	const float s1 = sinT / sinT0;
	return make_float4( v0.x * s0 + v1.x * s1, v0.y * s0 + v1.y * s1, v0.z * s0 + v1.z * s1, v0.w * s0 + v1.w * s1 );
}

HIPRT_DEVICE HIPRT_INLINE float3 mix3( const float3 lo, const float3 hi, float t )
{
	const float u = 1.0f - t;
	return make_float3( lo.x * u + hi.x * t, lo.y * u + hi.y * t, lo.z * u + hi.z * t );
}

// SRTFrame::interpolate (Transform.h:178-186): the time of the result is left at zero
HIPRT_DEVICE HIPRT_INLINE Frame interpolate( const Frame& f0, const Frame& f1, float t )
{
	Frame f;
	f.scale		  = mix3( f0.scale, f1.scale, t );
	f.shear		  = mix3( f0.shear, f1.shear, t );
	f.translation = mix3( f0.translation, f1.translation, t );
	f.rotation	  = qtMix( f0.rotation, f1.rotation, t );
	f.time		  = 0.0f;
	return f;
}

// object -> world of a point (Transform.h:188-197)
HIPRT_DEVICE HIPRT_INLINE float3 transformPoint( const Frame& f, const float3 p )
{
	if ( isIdentity( f ) ) return p;
	float3 r = make_float3( p.x * f.scale.x, p.y * f.scale.y, p.z * f.scale.z );
	r.x += p.y * f.shear.x + p.z * f.shear.y;
	r.y += p.z * f.shear.z;
	r = qtRotate( f.rotation, r );
	return make_float3( r.x + f.translation.x, r.y + f.translation.y, r.z + f.translation.z );
}

// object -> world of a normal-like vector: inverse transpose of the scale/shear part (Transform.h:199-208)
HIPRT_DEVICE HIPRT_INLINE float3 transformVector( const Frame& f, const float3 v )
{
	if ( isIdentity( f ) ) return v;
	float3 r = make_float3( v.x / f.scale.x, v.y / f.scale.y, v.z / f.scale.z );
	r.y -= v.x * f.shear.x / f.scale.y;
	r.z -= ( f.shear.y * r.x + f.shear.z * r.y ) / f.scale.z;
	return qtRotate( f.rotation, r );
}

// world -> object of a point (Transform.h:210-220)
HIPRT_DEVICE HIPRT_INLINE float3 invTransformPoint( const Frame& f, const float3 p )
{
	if ( isIdentity( f ) ) return p;
	float3 r = make_float3( p.x - f.translation.x, p.y - f.translation.y, p.z - f.translation.z );
	r		 = qtInvRotate( f.rotation, r );
	r		 = make_float3( r.x / f.scale.x, r.y / f.scale.y, r.z / f.scale.z );
	r.y -= p.z * f.shear.z / f.scale.y;
	r.x -= ( f.shear.x * r.y + f.shear.y * r.z ) / f.scale.x;
	return r;
}

// world -> object of a normal-like vector (Transform.h:222-230)
HIPRT_DEVICE HIPRT_INLINE float3 invTransformVector( const Frame& f, const float3 v )
{
	if ( isIdentity( f ) ) return v;
	float3 r = qtInvRotate( f.rotation, v );
	r		 = make_float3( r.x * f.scale.x, r.y * f.scale.y, r.z * f.scale.z );
	r.y += v.x * f.shear.x;
	r.z += v.x * f.shear.y + v.y * f.shear.z;
	return r;
}

// Transform::interpolateFrames (Transform.h:417-436): pose of an instance with key frames [frameIndex, frameIndex + frameCount) at `time`
HIPRT_DEVICE HIPRT_INLINE Frame frameAt( const SRTFrame* frames, uint32_t frameIndex, uint32_t frameCount, float time )
{
	if ( frameCount == 0 || frames == nullptr ) return identityFrame();
	frames += frameIndex;
	Frame f0 = loadKeyFrame( frames, 0 );
	if ( frameCount == 1 || time == 0.0f || time <= f0.time ) return f0;
	Frame f1 = loadKeyFrame( frames, frameCount - 1 );
	if ( time >= f1.time ) return f1;
	for ( uint32_t i = 1; i < frameCount; ++i )
	{
		f1 = loadKeyFrame( frames, i );
		if ( time >= f0.time && time <= f1.time ) break;
		f0 = f1;
	}
	const float t = ( time - f0.time ) / ( f1.time - f0.time );
	return interpolate( f0, f1, t );
}

// Transform::transformRay (Transform.h:438-449): the direction is the difference of two transformed points, t is preserved
HIPRT_DEVICE HIPRT_INLINE hiprtRay transformRay( const Frame& f, const hiprtRay& ray )
{
	if ( isIdentity( f ) ) return ray;
	hiprtRay	 out;
	const float3 o = ray.origin, d = ray.direction;
	out.origin	   = invTransformPoint( f, o );
	const float3 e = invTransformPoint( f, make_float3( o.x + d.x, o.y + d.y, o.z + d.z ) );
	out.direction  = make_float3( e.x - out.origin.x, e.y - out.origin.y, e.z - out.origin.z );
	out.minT	   = ray.minT;
	out.maxT	   = ray.maxT;
	return out;
}
} // namespace motion
} // namespace hiprt
#endif
