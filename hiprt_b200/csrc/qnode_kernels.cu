// Generation of a geometry's Q region (qnode.h) from its public Box4Node array: one thread per box node, 128 bytes in,
// 64 bytes out, streaming. Runs at the end of every build / update / compaction / load / relocation of a triangle geometry.
#include "qnode.h"

#include <hiprt/hiprt_b200_layout.h>

#include <cfloat>

namespace hiprt
{
namespace
{
// smallest power of two >= x and its reciprocal (x > 0, finite, normal), by exponent arithmetic: no division in the kernel
__device__ __forceinline__ void pow2Ceil( double x, double& p, double& invP )
{
	const long long bits = __double_as_longlong( x );
	const long long e	 = ( ( bits >> 52 ) & 0x7FF ) + ( ( bits & 0xFFFFFFFFFFFFFll ) != 0 ? 1 : 0 );
	p					 = __longlong_as_double( e << 52 );
	invP				 = __longlong_as_double( ( 2046ll - e ) << 52 );
}

__global__ void __launch_bounds__( 128 ) compressBoxNodesKernel( const GeomHeader* __restrict__ header, QHeader* __restrict__ qHeader, QBox4Node* __restrict__ qNodes )
{
	const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
	if ( k >= header->boxNodeCount ) return;
	if ( !( header->geomType & 1u ) )
	{
		if ( k == 0 ) qHeader->invalid = 1u;
		return;
	}
	const float4* src = reinterpret_cast<const float4*>( header->boxNodes + k );
	float		  w[32];
#pragma unroll
	for ( int i = 0; i < 8; ++i )
	{
		const float4 v = __ldcs( src + i );
		w[4 * i] = v.x, w[4 * i + 1] = v.y, w[4 * i + 2] = v.z, w[4 * i + 3] = v.w;
	}
	uint32_t child[4];
	bool	 valid[4];
	double	 lo[3] = { DBL_MAX, DBL_MAX, DBL_MAX }, hi[3] = { -DBL_MAX, -DBL_MAX, -DBL_MAX };
	bool	 bad   = false;
#pragma unroll
	for ( int s = 0; s < 4; ++s )
	{
		child[s] = __float_as_uint( w[s] );
		valid[s] = child[s] != InvalidValue;
		if ( !valid[s] ) continue;
#pragma unroll
		for ( int a = 0; a < 3; ++a )
		{
			const float l = w[4 + 6 * s + a], h = w[4 + 6 * s + 3 + a];
			// (an EMPTY child box -- lo > hi -- can only come from an import; such a geometry has no Q region, see api.cpp)
			if ( !( fabsf( l ) <= QMaxCoordinate ) || !( fabsf( h ) <= QMaxCoordinate ) || l > h ) bad = true;
			lo[a] = fmin( lo[a], double( l ) ), hi[a] = fmax( hi[a], double( h ) );
		}
	}
	QBox4Node out;
	uint32_t  q[6] = { 0, 0, 0, 0, 0, 0 };
	if ( bad || !( valid[0] || valid[1] || valid[2] || valid[3] ) )
	{
		if ( bad ) atomicOr( &qHeader->invalid, 1u );
		for ( int a = 0; a < 3; ++a )
			out.origin[a] = 0.0f, out.scale[a] = 32768.0f;
		for ( int a = 0; a < 3; ++a )
			q[a] = 0xFFFFFFFFu; // every slot inverted
	}
	else
	{
#pragma unroll
		for ( int a = 0; a < 3; ++a )
		{
			// at most 250 cells span the node, two more below it leave room for the margin; the cell never drops below 2^-20 of the coordinate
			// magnitude (8 ulp), where quantisation could not separate planes anyway
			const double ext  = hi[a] - lo[a];
			double		 cell, invCell;
			pow2Ceil( fmax( fmax( ext * 0.004, fmax( fabs( lo[a] ), fabs( hi[a] ) ) * 9.5367431640625e-07 ), 8.6736173798840355e-19 ), cell, invCell );
			const float	 org  = __double2float_rd( lo[a] - 2.0 * cell );
			out.origin[a]	  = org;
			out.scale[a]	  = float( cell * 32768.0 ); // exact: a power of two
#pragma unroll
			for ( int s = 0; s < 4; ++s )
			{
				uint32_t ql = 255u, qh = 0u; // unused slot: inverted
				if ( valid[s] )
				{
					const double l = double( w[4 + 6 * s + a] ), h = double( w[4 + 6 * s + 3 + a] );
					// plane ql lies at least one whole cell below l, plane qh at least one whole cell above h
					ql = uint32_t( fmin( fmax( floor( ( l - double( org ) ) * invCell - 0.5 ), 0.0 ), 255.0 ) );
					qh = uint32_t( fmin( fmax( ceil( ( h - double( org ) ) * invCell + 0.5 ), 0.0 ), 255.0 ) );
					// The margin the traversal's error analysis needs is a small fraction of a cell (its own rounding noise is below
					// 2^-8 of one, qnode.h): the planes are placed between half a cell and one and a half cells outside the box, a
					// quarter is asserted. (Cannot trigger for finite nested boxes: origin sits 2 cells below the node, 254 cells reach past it.)
					if ( double( org ) + ql * cell > l - 0.25 * cell || double( org ) + qh * cell < h + 0.25 * cell ) atomicOr( &qHeader->invalid, 1u );
				}
				q[a] |= ql << ( 8 * s );
				q[3 + a] |= qh << ( 8 * s );
			}
		}
	}
#pragma unroll
	for ( int s = 0; s < 4; ++s )
		out.child[s] = child[s];
#pragma unroll
	for ( int i = 0; i < 6; ++i )
		out.q[i] = q[i];
	float4*		  dst = reinterpret_cast<float4*>( qNodes + k );
	const float4* o4  = reinterpret_cast<const float4*>( &out );
#pragma unroll
	for ( int i = 0; i < 4; ++i )
		__stcs( dst + i, o4[i] );
}
} // namespace

void compressBoxNodes( const GeomHeader* header, void* qRegion, uint32_t maxBoxNodes, cudaStream_t stream )
{
	if ( !qRegion || maxBoxNodes == 0 ) return;
	cudaMemsetAsync( qRegion, 0, sizeof( QHeader ), stream );
	QBox4Node* nodes = reinterpret_cast<QBox4Node*>( static_cast<uint8_t*>( qRegion ) + sizeof( QHeader ) );
	compressBoxNodesKernel<<<( maxBoxNodes + 127u ) / 128u, 128, 0, stream>>>( header, static_cast<QHeader*>( qRegion ), nodes );
}
} // namespace hiprt
