// Traversal-only companion of the public Box4Node array ("Q nodes"): one 64-byte record per box node with the four child boxes
// quantised to 8 bits per plane on a power-of-two grid local to the node. The library's triangle trace kernels are bound by L1
// data-pipe wavefronts -- a lane that fetches its own 128-byte Box4Node costs four of them (profiles/r02_cache_peaks.json:
// 31.8 B/clk/SM for divergent 32-byte loads) -- so the kernels walk these 64-byte records (two wavefronts) instead and touch
// the public format only for leaves. The public, reference-compatible layout (include/hiprt/hiprt_b200_layout.h) is untouched:
// the Q region is appended to the library's own allocation behind the `size` bytes the header reports, is never saved, and is
// regenerated after every build / update / compaction / load / relocation (api.cpp keeps the handle -> region map).
//
// Results stay bit-identical to the exact traversal:
//   * the quantised box of a child CONTAINS its exact box with at least half a grid cell of margin per side, the plane distances
//     are evaluated with directed rounding and widened by a relative 2^-20, so a child whose exact slab test passes always
//     passes the quantised one: the quantised walk visits a superset of the exact walk's nodes;
//   * a leaf that is reached is first put through the EXACT slab test of its own box (recomputed from the leaf's vertices,
//     which is how the builder produced it), against the ray's original extent. The exact float32 slab arithmetic is monotonic
//     in the plane positions and the boxes of a BVH are nested, so "the leaf's exact box passes" is equivalent to "every exact
//     ancestor box passes": the triangles tested are those of the exact walk, plus at most triangles whose box the exact walk
//     culled by DISTANCE (entry beyond the closest hit so far) -- which the triangle test then rejects by that same distance,
//     an exact tie excepted, where the smaller primitive id wins as in the exact walk. The (t, primID) minimum is the same;
//   * rays for which that monotonicity argument does not hold (a zero, denormal, infinite or NaN direction component, a
//     non-finite or astronomically large origin) never enter the quantised walk: their lane runs the exact traversal.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

namespace hiprt
{
struct alignas( 64 ) QBox4Node
{
	float	 origin[3]; // plane q of axis a lies at origin[a] + q * cell[a]
	float	 scale[3];	// cell[a] * 2^15 (cell is a power of two): t = (1 + q * 2^-15) * (scale * invD) + ((origin - o) * invD - scale * invD)
	uint32_t child[4];	// as Box4Node::child
	uint32_t q[6];		// lo.x, lo.y, lo.z, hi.x, hi.y, hi.z: byte s = plane index of child s; unused slots hold lo = 255, hi = 0
};
static_assert( sizeof( QBox4Node ) == 64, "QBox4Node is half a Box4Node" );

// First 64 bytes of a Q region.
struct alignas( 64 ) QHeader
{
	uint32_t invalid; // != 0: the records must not be used (a node with non-finite or > 2^40 coordinates, not a triangle mesh)
	uint32_t pad[15];
};

constexpr float	 QMaxCoordinate = 1099511627776.0f; // 2^40: coordinates / ray origins beyond it take the exact path
constexpr size_t qRegionSize( size_t maxBoxNodes ) { return sizeof( QHeader ) + maxBoxNodes * sizeof( QBox4Node ); }

struct GeomHeader;
// (Re)generates the Q region of a built triangle geometry from its Box4Node array; header is a DEVICE pointer, everything is
// read on the device (box-node count, node array address), nothing synchronises. maxBoxNodes bounds the launch.
void compressBoxNodes( const GeomHeader* header, void* qRegion, uint32_t maxBoxNodes, cudaStream_t stream );
} // namespace hiprt
