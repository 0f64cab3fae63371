// Instance transforms for the scene (top-level) builder: API frames -> internal SRT frames,
// forward transform of box corners, inverse 3x4 matrix stored in the instance leaves.
// Reference counterparts: hiprt/impl/Transform.h:41-240 (SRTFrame, SRTToInvMatrix),
// Quaternion.h:30-140, QrDecomposition.h:30-135 (modified Gram-Schmidt).
#pragma once
#include "device_util.cuh"

namespace hiprt
{
__device__ __forceinline__ float sumProd( float a, float b, float c, float d )
{
	const float cd = __fmul_rn( c, d );
	return __fadd_rn( fmaf( a, b, cd ), fmaf( c, d, -cd ) );
}
__device__ __forceinline__ float diffProd( float a, float b, float c, float d )
{
	const float cd = __fmul_rn( c, d );
	return __fsub_rn( fmaf( a, b, -cd ), fmaf( c, d, -cd ) );
}
__device__ __forceinline__ float3 crossC( const float3 a, const float3 b )
{
	return make_float3( diffProd( a.y, b.z, a.z, b.y ), diffProd( a.z, b.x, a.x, b.z ), diffProd( a.x, b.y, a.y, b.x ) );
}
__device__ __forceinline__ float dotC( const float3 a, const float3 b ) { return fmaf( a.x, b.x, sumProd( a.y, b.y, a.z, b.z ) ); }
// overflow-safe length (hiprt_math.h:1583-1590)
__device__ __forceinline__ float hypot3( const float3 a )
{
	const float x = fabsf( a.x ), y = fabsf( a.y ), z = fabsf( a.z );
	const float d = fmaxf( x, fmaxf( y, z ) );
	if ( d == 0.0f ) return 0.0f;
	const float rx = __fdiv_rn( x, d ), ry = __fdiv_rn( y, d ), rz = __fdiv_rn( z, d );
	return __fmul_rn( d, __fsqrt_rn( __fadd_rn( __fadd_rn( __fmul_rn( rx, rx ), __fmul_rn( ry, ry ) ), __fmul_rn( rz, rz ) ) ) );
}

__device__ __forceinline__ float4 quatMul( const float4 a, const float4 b )
{
	const float3 av = make_float3( a.x, a.y, a.z ), bv = make_float3( b.x, b.y, b.z );
	const float3 c	= crossC( av, bv );
	return make_float4(
		__fadd_rn( __fadd_rn( c.x, __fmul_rn( a.w, b.x ) ), __fmul_rn( b.w, a.x ) ),
		__fadd_rn( __fadd_rn( c.y, __fmul_rn( a.w, b.y ) ), __fmul_rn( b.w, a.y ) ),
		__fadd_rn( __fadd_rn( c.z, __fmul_rn( a.w, b.z ) ), __fmul_rn( b.w, a.z ) ),
		__fsub_rn( __fmul_rn( a.w, b.w ), dotC( av, bv ) ) );
}
__device__ __forceinline__ float3 quatRotate( const float4 q, const float3 p )
{
	const float4 r = quatMul( quatMul( q, make_float4( p.x, p.y, p.z, 0.0f ) ), make_float4( -q.x, -q.y, -q.z, q.w ) );
	return make_float3( r.x, r.y, r.z );
}

__device__ __forceinline__ float4 quatFromAxisAngle( const float4 aa )
{
	if ( aa.w == 0.0f ) return make_float4( 0.0f, 0.0f, 0.0f, 1.0f );
	const float len = hypot3( make_float3( aa.x, aa.y, aa.z ) );
	const float s	= sinf( aa.w * 0.5f );
	return make_float4( __fdiv_rn( aa.x, len ) * s, __fdiv_rn( aa.y, len ) * s, __fdiv_rn( aa.z, len ) * s, cosf( aa.w * 0.5f ) );
}

__device__ __forceinline__ float4 quatFromRotationMatrix( const float R[3][3] )
{
	const float tr = R[0][0] + R[1][1] + R[2][2];
	float4		q;
	if ( tr > 0.0f )
	{
		const float S = sqrtf( tr + 1.0f ) * 2.0f;
		q			  = make_float4( ( R[2][1] - R[1][2] ) / S, ( R[0][2] - R[2][0] ) / S, ( R[1][0] - R[0][1] ) / S, 0.25f * S );
	}
	else if ( R[0][0] > R[1][1] && R[0][0] > R[2][2] )
	{
		const float S = sqrtf( 1.0f + R[0][0] - R[1][1] - R[2][2] ) * 2.0f;
		q			  = make_float4( 0.25f * S, ( R[0][1] + R[1][0] ) / S, ( R[0][2] + R[2][0] ) / S, ( R[2][1] - R[1][2] ) / S );
	}
	else if ( R[1][1] > R[2][2] )
	{
		const float S = sqrtf( 1.0f + R[1][1] - R[0][0] - R[2][2] ) * 2.0f;
		q			  = make_float4( ( R[0][1] + R[1][0] ) / S, 0.25f * S, ( R[1][2] + R[2][1] ) / S, ( R[0][2] - R[2][0] ) / S );
	}
	else
	{
		const float S = sqrtf( 1.0f + R[2][2] - R[0][0] - R[1][1] ) * 2.0f;
		q			  = make_float4( ( R[0][2] + R[2][0] ) / S, ( R[1][2] + R[2][1] ) / S, 0.25f * S, ( R[1][0] - R[0][1] ) / S );
	}
	return q;
}

__device__ __forceinline__ void quatToRotationMatrix( const float4 q, float R[3][3] )
{
	const float x2 = q.x * q.x, y2 = q.y * q.y, z2 = q.z * q.z;
	R[0][0] = 1.0f - 2.0f * y2 - 2.0f * z2;
	R[0][1] = 2.0f * q.x * q.y - 2.0f * q.w * q.z;
	R[0][2] = 2.0f * q.x * q.z + 2.0f * q.w * q.y;
	R[1][0] = 2.0f * q.x * q.y + 2.0f * q.w * q.z;
	R[1][1] = 1.0f - 2.0f * x2 - 2.0f * z2;
	R[1][2] = 2.0f * q.y * q.z - 2.0f * q.w * q.x;
	R[2][0] = 2.0f * q.x * q.z - 2.0f * q.w * q.y;
	R[2][1] = 2.0f * q.y * q.z + 2.0f * q.w * q.x;
	R[2][2] = 1.0f - 2.0f * x2 - 2.0f * y2;
}

// A = Q R by modified Gram-Schmidt on the columns of A (row-major 3x3); det(Q) forced to +1
__device__ __forceinline__ void qrDecompose( const float A[3][3], float Q[3][3], float R[3][3] )
{
#pragma unroll
	for ( int i = 0; i < 3; ++i )
#pragma unroll
		for ( int j = 0; j < 3; ++j )
			Q[i][j] = A[i][j], R[i][j] = 0.0f;
#pragma unroll
	for ( int c = 0; c < 3; ++c )
	{
#pragma unroll
		for ( int p = 0; p < c; ++p )
		{
			const float w = Q[0][p] * Q[0][c] + Q[1][p] * Q[1][c] + Q[2][p] * Q[2][c];
			R[p][c]		  = w;
			Q[0][c] -= w * Q[0][p], Q[1][c] -= w * Q[1][p], Q[2][c] -= w * Q[2][p];
		}
		float w = hypot3( make_float3( Q[0][c], Q[1][c], Q[2][c] ) );
		R[c][c] = w;
		if ( w != 0.0f )
		{
			w = 1.0f / w;
			Q[0][c] *= w, Q[1][c] *= w, Q[2][c] *= w;
		}
		else
			Q[0][c] = 1.0f;
	}
	const float d = Q[0][0] * Q[1][1] * Q[2][2] + Q[0][1] * Q[1][2] * Q[2][0] + Q[1][0] * Q[2][1] * Q[0][2] - Q[0][2] * Q[1][1] * Q[2][0] -
					Q[0][1] * Q[1][0] * Q[2][2] - Q[1][2] * Q[2][1] * Q[0][0];
	if ( d < 0.0f )
	{
		Q[0][0] = -Q[0][0], Q[1][0] = -Q[1][0], Q[2][0] = -Q[2][0];
		R[0][0] = -R[0][0], R[0][1] = -R[0][1], R[0][2] = -R[0][2];
	}
}

struct DFrame
{
	float4 rotation;
	float3 scale, shear, translation;
	float  time;
};

__device__ __forceinline__ DFrame identityFrame()
{
	return DFrame{ make_float4( 0, 0, 0, 1 ), make_float3( 1, 1, 1 ), make_float3( 0, 0, 0 ), make_float3( 0, 0, 0 ), 0.0f };
}

__device__ __forceinline__ DFrame frameFromSRT( const hiprtFrameSRT& f )
{
	DFrame d;
	d.rotation	  = quatFromAxisAngle( f.rotation );
	d.scale		  = f.scale;
	d.shear		  = make_float3( 0, 0, 0 );
	d.translation = f.translation;
	d.time		  = f.time;
	return d;
}

__device__ __forceinline__ DFrame frameFromMatrix( const hiprtFrameMatrix& f )
{
	const float( &m )[3][4] = f.matrix;
	const bool identity		= m[0][0] == 1.0f && m[1][1] == 1.0f && m[2][2] == 1.0f && m[0][1] == 0.0f && m[0][2] == 0.0f && m[0][3] == 0.0f &&
						  m[1][0] == 0.0f && m[1][2] == 0.0f && m[1][3] == 0.0f && m[2][0] == 0.0f && m[2][1] == 0.0f && m[2][3] == 0.0f;
	DFrame d = identityFrame();
	if ( !identity )
	{
		float A[3][3], Q[3][3], R[3][3];
#pragma unroll
		for ( int i = 0; i < 3; ++i )
#pragma unroll
			for ( int j = 0; j < 3; ++j )
				A[i][j] = m[i][j];
		qrDecompose( A, Q, R );
		d.translation = make_float3( m[0][3], m[1][3], m[2][3] );
		d.rotation	  = quatFromRotationMatrix( Q );
		d.scale		  = make_float3( R[0][0], R[1][1], R[2][2] );
		d.shear		  = make_float3( R[0][1], R[0][2], R[1][2] );
	}
	d.time = f.time;
	return d;
}

__device__ __forceinline__ bool frameIsIdentity( const DFrame& f )
{
	return f.scale.x == 1.0f && f.scale.y == 1.0f && f.scale.z == 1.0f && f.shear.x == 0.0f && f.shear.y == 0.0f && f.shear.z == 0.0f &&
		   f.translation.x == 0.0f && f.translation.y == 0.0f && f.translation.z == 0.0f && f.rotation.w == 1.0f;
}

__device__ __forceinline__ DFrame loadFrame( const SRTFrame* frames, uint32_t i )
{
	const float4* p = reinterpret_cast<const float4*>( frames + i );
	const float4  a = p[0], b = p[1], c = p[2], d = p[3];
	DFrame		  f;
	f.rotation	  = a;
	f.scale		  = make_float3( b.x, b.y, b.z );
	f.shear		  = make_float3( b.w, c.x, c.y );
	f.translation = make_float3( c.z, c.w, d.x );
	f.time		  = d.y;
	return f;
}

__device__ __forceinline__ void storeFrame( SRTFrame* frames, uint32_t i, const DFrame& f )
{
	float4* p = reinterpret_cast<float4*>( frames + i );
	p[0]	  = f.rotation;
	p[1]	  = make_float4( f.scale.x, f.scale.y, f.scale.z, f.shear.x );
	p[2]	  = make_float4( f.shear.y, f.shear.z, f.translation.x, f.translation.y );
	p[3]	  = make_float4( f.translation.z, f.time, 0.0f, 0.0f );
}

// object -> world: scale, shear, rotate, translate (Transform.h:188-197)
__device__ __forceinline__ float3 frameTransformPoint( const DFrame& f, const float3 p )
{
	float3 r = make_float3( p.x * f.scale.x, p.y * f.scale.y, p.z * f.scale.z );
	r.x += p.y * f.shear.x + p.z * f.shear.y;
	r.y += p.z * f.shear.z;
	r = quatRotate( f.rotation, r );
	return make_float3( r.x + f.translation.x, r.y + f.translation.y, r.z + f.translation.z );
}

// world -> object 3x4 matrix (Transform.h:61-97): inverse of the upper-triangular scale/shear times
// the transposed rotation, then the translation column.
__device__ __forceinline__ void frameInverseMatrix( const DFrame& f, float out[3][4] )
{
	float Q[3][3];
	quatToRotationMatrix( f.rotation, Q );
	float Ri[3][3];
	Ri[0][0] = 1.0f / f.scale.x;
	Ri[1][1] = 1.0f / f.scale.y;
	Ri[2][2] = 1.0f / f.scale.z;
	Ri[0][1] = -f.shear.x / ( f.scale.x * f.scale.y );
	Ri[0][2] = ( f.shear.x * f.shear.z - f.shear.y * f.scale.y ) / ( f.scale.x * f.scale.y * f.scale.z );
	Ri[1][2] = -f.shear.z / ( f.scale.y * f.scale.z );
	Ri[1][0] = Ri[2][0] = Ri[2][1] = 0.0f;
#pragma unroll
	for ( int i = 0; i < 3; ++i )
	{
#pragma unroll
		for ( int j = 0; j < 3; ++j )
			out[i][j] = dotC( make_float3( Ri[i][0], Ri[i][1], Ri[i][2] ), make_float3( Q[j][0], Q[j][1], Q[j][2] ) );
		out[i][3] = -dotC( make_float3( out[i][0], out[i][1], out[i][2] ), f.translation );
	}
}
} // namespace hiprt
