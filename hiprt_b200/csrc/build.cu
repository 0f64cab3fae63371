// Host orchestration of the builders: temporary-buffer layout, launch sequences.
// (kernels: build_kernels.cuh; reference host side: hiprt/impl/LbvhBuilder.{h,cpp},
// PlocBuilder.{h,cpp}, BatchBuilder.{h,cpp}, Context.cpp:144-290.)
#include "batch_kernels.cuh"
#include "sbvh_kernels.cuh"
#include <cstdlib>
#include "qnode.h"

namespace hiprt
{
TempLayout tempLayout( size_t n, size_t pairTriangles )
{
	TempLayout L{};
	size_t	   off = 0;
	auto	   take = [&]( size_t bytes ) {
		  const size_t o = off;
		  off += alignUp( bytes );
		  return o;
	};
	L.control	  = take( sizeof( BuildControl ) );
	L.pairs		  = take( pairTriangles * sizeof( uint2 ) );
	L.pairScan	  = take( DivideRoundUp( pairTriangles, size_t( PairTile ) ) * sizeof( uint32_t ) );
	L.primRefs	  = take( n * 32 );
	L.keys[0]	  = take( n * 4 );
	L.vals[0]	  = take( n * 4 );
	L.keys[1]	  = take( n * 4 );
	L.vals[1]	  = take( n * 4 );
	L.sortScratch = take( rsort::scratchWords( uint32_t( n ) ) * 4 );
	L.sortedRefs  = take( n * 32 );
	L.scratch	  = take( n * 32 );
	L.flags		  = take( n * 4 );
	L.clusterIds  = take( n * 4 ); // balanced builder only
	L.taskQueue	  = L.keys[1]; // 8 B x maxBox(n) <= 8 B x n: reuses the second key/value buffer
	L.total		  = off;
	return L;
}

// ------------------------------------------------------------------------------------------------
// High-quality build (SBVH): temporary buffer and launch. `maxRefs` = leafCapacity(): the reference budget.
struct SbvhLayout
{
	size_t control, state, timeline, pairs, pairScan, primRefs, refs[2], refTask[2], tasks[2], objectBins, spatialBins, nodes, leaves, taskQueue, total;
};
static SbvhLayout sbvhLayout( size_t n, size_t pairTriangles, size_t maxRefs, bool spatial )
{
	SbvhLayout L{};
	size_t	   off	= 0;
	auto	   take = [&]( size_t bytes ) {
		  const size_t o = off;
		  off += alignUp( bytes );
		  return o;
	};
	const size_t maxTasks = std::max<size_t>( maxRefs / 2, 1 );
	L.control	  = take( sizeof( BuildControl ) );
	L.state		  = take( sizeof( sbvh::State ) );
	L.timeline	  = take( sbvh::TimelineLevels * sbvh::TimelineSlots * sizeof( uint64_t ) );
	L.pairs		  = take( pairTriangles * sizeof( uint2 ) );
	L.pairScan	  = take( DivideRoundUp( pairTriangles, size_t( PairTile ) ) * sizeof( uint32_t ) );
	L.primRefs	  = take( n * 32 );
	L.refs[0]	  = take( maxRefs * 32 );
	L.refs[1]	  = take( maxRefs * 32 );
	L.refTask[0]  = take( maxRefs * 4 );
	L.refTask[1]  = take( maxRefs * 4 );
	L.tasks[0]	  = take( maxTasks * sizeof( sbvh::Task ) );
	L.tasks[1]	  = take( maxTasks * sizeof( sbvh::Task ) );
	L.objectBins  = take( maxTasks * 3 * sbvh::MinBins * 32 );
	L.spatialBins = take( spatial ? maxTasks * 3 * sbvh::MinBins * 32 : 0 );
	L.nodes		  = take( maxRefs * 32 );
	L.leaves	  = take( maxRefs * 32 );
	L.taskQueue	  = take( maxBoxNodeCount( maxRefs ) * 8 );
	L.total		  = off;
	return L;
}

void highQualityTempLayout( size_t n, size_t pairTriangles, bool spatial, size_t out[8] )
{
	const size_t	 maxRefs = spatial ? n + n / 2 : n;
	const SbvhLayout L		 = sbvhLayout( n, pairTriangles, maxRefs, spatial );
	out[0] = L.control, out[1] = L.state, out[2] = L.primRefs, out[3] = L.nodes, out[4] = L.leaves, out[5] = maxRefs, out[6] = L.taskQueue, out[7] = L.total;
}

// ------------------------------------------------------------------------------------------------
// Custom BVH import (hiprtBuildFlagBitCustomBvhImport): the caller's binary tree decides the leaf count, which exceeds the
// primitive count when several leaves reference one primitive (spatial splits). Storage is sized for whichever is larger.
static bool isImport( hiprtBuildOptions options ) { return ( options.buildFlags & 3u ) == hiprtBuildFlagBitCustomBvhImport; }
static size_t importLeafCount( const hiprtGeometryBuildInput& input )
{
	return ( input.nodeList.internalNodes != nullptr || input.nodeList.leafNodes != nullptr ) ? input.nodeList.nodeCount : 0;
}

size_t geometryPrimCount( const hiprtGeometryBuildInput& input )
{
	if ( input.type == hiprtPrimitiveTypeTriangleMesh )
	{
		const auto& m = input.primitive.triangleMesh;
		if ( m.trianglePairCount > 0 ) return m.trianglePairCount;
		if ( m.triangleCount == 0 || m.triangleIndices == nullptr ) return m.vertexCount / 3;
		return m.triangleCount;
	}
	if ( input.type == hiprtPrimitiveTypeAABBList ) return input.primitive.aabbList.aabbCount;
	throw std::runtime_error( "unsupported primitive type" );
}

// PreferBalancedBuild -> H-PLOC; PreferHighQualityBuild -> SBVH (sbvh_kernels.cuh); anything else -> LBVH (Context.cpp:174-198).
static bool usesBalancedBuilder( hiprtBuildOptions options ) { return ( options.buildFlags & 3u ) == hiprtBuildFlagBitPreferBalancedBuild; }
static bool usesSbvhBuilder( hiprtBuildOptions options ) { return ( options.buildFlags & 3u ) == hiprtBuildFlagBitPreferHighQualityBuild; }

// Leaves a build of `input` may create: one per primitive, except that a high-quality build with spatial splits may reference a
// primitive from several leaves -- half as many again at most (SbvhAlpha = 1.5, BvhConfig.h:48; SbvhBuilder.h:158-160) -- and
// an imported hierarchy brings its own leaf count.
static size_t leafCapacity( const hiprtGeometryBuildInput& input, hiprtBuildOptions options )
{
	const size_t n = geometryPrimCount( input );
	if ( isImport( options ) ) return std::max( n, importLeafCount( input ) );
	if ( usesSbvhBuilder( options ) && !( options.buildFlags & hiprtBuildFlagBitDisableSpatialSplits ) ) return n + n / 2;
	return n;
}

static bool wantsPairing( const hiprtGeometryBuildInput& input, hiprtBuildOptions options )
{
	if ( input.type != hiprtPrimitiveTypeTriangleMesh ) return false;
	const auto& m = input.primitive.triangleMesh;
	// reference: TriangleMesh::pairable() (TriangleMesh.h:246) and LbvhBuilder.h:186-189
	return m.triangleIndices != nullptr && m.triangleCount > 2 && m.trianglePairCount == 0 &&
		   !( options.buildFlags & hiprtBuildFlagBitDisableTrianglePairing ) && !isImport( options );
}

size_t geometryStorageSize( const hiprtGeometryBuildInput& input, hiprtBuildOptions options )
{
	const size_t n		  = leafCapacity( input, options );
	const size_t leafSize = input.type == hiprtPrimitiveTypeTriangleMesh ? sizeof( TrianglePairNode ) : sizeof( CustomNode );
	return alignUp( sizeof( GeomHeader ) ) + alignUp( maxBoxNodeCount( n ) * sizeof( Box4Node ) ) + alignUp( n * leafSize );
}

// What hiprtCreateGeometry allocates: the public structure (`size` of the header, what hiprtSaveGeometry writes) followed, for
// triangle meshes, by the traversal-only Q region of the library's own trace kernels (qnode.h).
size_t geometryAllocationSize( const hiprtGeometryBuildInput& input, hiprtBuildOptions options )
{
	const size_t n = leafCapacity( input, options );
	return geometryStorageSize( input, options ) + ( input.type == hiprtPrimitiveTypeTriangleMesh ? qRegionSize( maxBoxNodeCount( n ) ) : 0 );
}

// Where the Q region of a geometry built from `input` lies inside its allocation, or nullptr when such a build has none
// (custom primitives, imported hierarchies whose leaf boxes are the caller's, a single primitive).
void* geometryQRegion( const hiprtGeometryBuildInput& input, hiprtBuildOptions options, void* storage )
{
	if ( input.type != hiprtPrimitiveTypeTriangleMesh || isImport( options ) || geometryPrimCount( input ) < 2 ) return nullptr;
	return static_cast<uint8_t*>( storage ) + geometryStorageSize( input, options );
}

size_t geometryTempSize( const hiprtGeometryBuildInput& input, hiprtBuildOptions options )
{
	if ( isImport( options ) ) return tempLayout( std::max<size_t>( importLeafCount( input ), 1 ), 0 ).total;
	const size_t n = geometryPrimCount( input );
	if ( usesSbvhBuilder( options ) ) return sbvhLayout( n, wantsPairing( input, options ) ? n : 0, leafCapacity( input, options ), leafCapacity( input, options ) > n ).total;
	return tempLayout( n, wantsPairing( input, options ) ? n : 0 ).total;
}

// Binary tree by sbvh::buildKernel (one cooperative launch, all levels), then the shared collapse + pack.
template <typename Source>
static void runSbvh(
	const DeviceInfo& dev, Source src, uint32_t nMax, uint32_t maxRefs, bool spatial, uint8_t* temp, const SbvhLayout& L, cudaStream_t stream,
	uint32_t* boxNodeCounter, uint32_t* primNodeCounter, Box4Node* boxNodes, void* leafNodes )
{
	BuildControl*  ctl		 = reinterpret_cast<BuildControl*>( temp + L.control );
	uint64_t*	   taskQueue = reinterpret_cast<uint64_t*>( temp + L.taskQueue );
	const uint32_t maxBox	 = uint32_t( maxBoxNodeCount( maxRefs ) );
	HIPRT_CUDA_CHECK( cudaMemsetAsync( temp + L.state, 0, sizeof( sbvh::State ), stream ) );
	HIPRT_CUDA_CHECK( cudaMemsetAsync( taskQueue, 0xFF, size_t( maxBox ) * 8, stream ) );

	sbvh::Params P{};
	P.state		  = reinterpret_cast<sbvh::State*>( temp + L.state );
	P.ctl		  = ctl;
	P.primRefs	  = temp + L.primRefs;
	P.refs[0] = temp + L.refs[0], P.refs[1] = temp + L.refs[1];
	P.refTask[0] = reinterpret_cast<uint32_t*>( temp + L.refTask[0] ), P.refTask[1] = reinterpret_cast<uint32_t*>( temp + L.refTask[1] );
	P.tasks[0] = reinterpret_cast<sbvh::Task*>( temp + L.tasks[0] ), P.tasks[1] = reinterpret_cast<sbvh::Task*>( temp + L.tasks[1] );
	P.objectBins  = reinterpret_cast<float*>( temp + L.objectBins );
	P.spatialBins = spatial ? reinterpret_cast<float*>( temp + L.spatialBins ) : nullptr;
	P.nodes		  = temp + L.nodes;
	P.leaves	  = temp + L.leaves;
	P.taskQueue	  = taskQueue;
	P.maxRefs	  = maxRefs;
	// (read per build, so that a test can force either form of the per-task phases on small inputs)
	const char* serialMin = getenv( "HIPRT_B200_SBVH_SERIAL_MIN" );
	P.serialMin			  = serialMin ? uint32_t( atoll( serialMin ) ) : 32768u;
	static const bool wantTimeline = getenv( "HIPRT_B200_SBVH_TIMELINE" ) != nullptr; // tools/profile_sbvh.py
	P.timeline	  = wantTimeline ? reinterpret_cast<uint64_t*>( temp + L.timeline ) : nullptr;
	if ( wantTimeline ) HIPRT_CUDA_CHECK( cudaMemsetAsync( temp + L.timeline, 0, sbvh::TimelineLevels * sbvh::TimelineSlots * sizeof( uint64_t ), stream ) );

	// every CTA of the grid must be resident (the levels are separated by grid barriers): sized from the occupancy of this kernel,
	// launched cooperatively so that the runtime refuses rather than deadlocks if the device cannot hold it
	static int perSm = 0; // per Source instantiation; same value every time: a benign race
	if ( perSm == 0 )
	{
		int v = 0;
		HIPRT_CUDA_CHECK( cudaOccupancyMaxActiveBlocksPerMultiprocessor( &v, sbvh::buildKernel<Source>, int( sbvh::Block ), 0 ) );
		if ( v < 1 ) throw std::runtime_error( "high-quality build: kernel does not fit an SM" );
		perSm = std::min( v, 4 );
	}
	const uint32_t wanted = DivideRoundUp( std::max( nMax, 1u ), sbvh::Block );
	const uint32_t grid	  = std::max( 1u, std::min( wanted, dev.smCount * uint32_t( perSm ) ) );
	void*		   args[] = { &src, &P };
	HIPRT_CUDA_CHECK( cudaLaunchCooperativeKernel( reinterpret_cast<const void*>( sbvh::buildKernel<Source> ), dim3( grid ), dim3( sbvh::Block ), args, 0, stream ) );
	dev.mark( "sbvh", stream );

	const uint32_t warpsNeeded	 = DivideRoundUp( maxBox, 8u );
	const uint32_t blocks		 = std::max( 1u, std::min( DivideRoundUp( warpsNeeded, CollapseBlock / 32 ), dev.smCount * 16 ) );
	const uint32_t residentWarps = dev.smCount * 40;
	const uint32_t claimSize	 = std::max( 8u, std::min( 32u, ( maxBox / ( residentWarps * 16 ) ) & ~7u ) );
	collapsePackKernel<Source><<<blocks, CollapseBlock, 0, stream>>>(
		src, ctl, boxNodeCounter, primNodeCounter, temp + L.nodes, temp + L.leaves, taskQueue, maxBox, claimSize, boxNodes, leafNodes );
	dev.mark( "collapse+pack", stream );
	HIPRT_CUDA_CHECK( cudaGetLastError() );
}

// Sort + emit + collapse for primitives whose 32-byte box records are already in temp.
template <typename Source>
static void runLbvh(
	const DeviceInfo& dev,
	Source			  src,
	bool			  balanced,
	uint32_t		  nMax,
	uint8_t*		  temp,
	const TempLayout& L,
	cudaStream_t	  stream,
	uint32_t*		  boxNodeCounter,
	uint32_t*		  primNodeCounter,
	Box4Node*		  boxNodes,
	void*			  leafNodes )
{
	BuildControl* ctl		 = reinterpret_cast<BuildControl*>( temp + L.control );
	void*		  primRefs	 = temp + L.primRefs;
	uint32_t*	  keys[2]	 = { reinterpret_cast<uint32_t*>( temp + L.keys[0] ), reinterpret_cast<uint32_t*>( temp + L.keys[1] ) };
	uint32_t*	  vals[2]	 = { reinterpret_cast<uint32_t*>( temp + L.vals[0] ), reinterpret_cast<uint32_t*>( temp + L.vals[1] ) };
	uint32_t*	  sortScr	 = reinterpret_cast<uint32_t*>( temp + L.sortScratch );
	void*		  sortedRefs = temp + L.sortedRefs;
	void*		  scratch	 = temp + L.scratch;
	uint32_t*	  flags		 = reinterpret_cast<uint32_t*>( temp + L.flags );
	uint64_t*	  taskQueue	 = reinterpret_cast<uint64_t*>( temp + L.taskQueue );
	const uint32_t maxBox	 = uint32_t( maxBoxNodeCount( nMax ) );

	const uint32_t streamGrid = std::min<uint32_t>( DivideRoundUp( nMax, 256u ), dev.smCount * 8 );
	HIPRT_CUDA_CHECK( cudaMemsetAsync( sortScr, 0, rsort::scratchWords( nMax ) * 4, stream ) );
	mortonKernel<<<streamGrid, 256, 0, stream>>>( ctl, primRefs, keys[0], vals[0], sortScr );
	dev.mark( "morton+histogram", stream );

	const uint32_t tiles = rsort::tileCount( nMax );
	// The staging buffer holds key + value of a whole tile: 48 KB, above the default dynamic shared-memory limit. Function attributes
	// belong to the CUDA context the caller made current (an application may create one per hiprtContext and destroy it again, as the
	// reference's unit tests do), so the attribute is set before every build rather than remembered: a host-side call of about a microsecond.
	HIPRT_CUDA_CHECK( cudaFuncSetAttribute( rsort::onesweepPass, cudaFuncAttributeMaxDynamicSharedMemorySize, int( sizeof( rsort::Smem ) ) ) );
	for ( int pass = 0; pass < rsort::Passes; ++pass )
	{
		const int in = pass & 1, out = in ^ 1;
		rsort::onesweepPass<<<tiles, rsort::Block, sizeof( rsort::Smem ), stream>>>(
			keys[in], vals[in], keys[out], vals[out], &ctl->primCount, nMax, sortScr, pass );
	}
	static_assert( rsort::Passes % 2 == 0, "sorted pairs must end up in buffer 0" );
	dev.mark( "sort", stream );

	// keys[1]/vals[1] are dead now: they become the collapse task queue
	HIPRT_CUDA_CHECK( cudaMemsetAsync( taskQueue, 0xFF, size_t( maxBox ) * 8, stream ) );
	HIPRT_CUDA_CHECK( cudaMemsetAsync( flags, 0xFF, size_t( nMax ) * 4, stream ) );
	if ( balanced )
	{
		uint32_t* clusterIds = reinterpret_cast<uint32_t*>( temp + L.clusterIds );
		static const int radius = [] {
			const char* e = getenv( "HIPRT_B200_PLOC_RADIUS" ); // 2 (default) / 4 / 8: see PlocRadius
			return e ? atoi( e ) : int( PlocRadius );
		}();
		auto kernel = radius == 4 ? hplocKernel<Source::LeafType, 4> : ( radius == 8 ? hplocKernel<Source::LeafType, 8> : hplocKernel<Source::LeafType, PlocRadius> );
		kernel<<<DivideRoundUp( nMax, PlocBlock ), PlocBlock, 0, stream>>>(
			ctl, keys[0], vals[0], primRefs, sortedRefs, scratch, flags, clusterIds, &ctl->nodeCounter, taskQueue );
		dev.mark( "hploc", stream );
	}
	else
	{
		emitTopologyKernel<Source::LeafType><<<DivideRoundUp( nMax, EmitTile ), EmitTile, 0, stream>>>(
			ctl, keys[0], vals[0], primRefs, sortedRefs, scratch, flags, taskQueue );
		dev.mark( "emit+refit", stream );
	}

	const uint32_t warpsNeeded = DivideRoundUp( maxBox, 8u );
	const uint32_t blocks	   = std::max( 1u, std::min( DivideRoundUp( warpsNeeded, CollapseBlock / 32 ), dev.smCount * 16 ) );
	// slots per claim: large trees amortise the claim atomic over 32 tasks, small ones keep 8 so that the tasks of a level
	// spread over many warps instead of queueing inside one
	const uint32_t residentWarps = dev.smCount * 40;
	const uint32_t claimSize	 = std::max( 8u, std::min( 32u, ( maxBox / ( residentWarps * 16 ) ) & ~7u ) );
	collapsePackKernel<Source><<<blocks, CollapseBlock, 0, stream>>>(
		src, ctl, boxNodeCounter, primNodeCounter, scratch, sortedRefs, taskQueue, maxBox, claimSize, boxNodes, leafNodes );
	dev.mark( "collapse+pack", stream );
	HIPRT_CUDA_CHECK( cudaGetLastError() );
}

// Custom BVH import (reference: BvhImporter.h:150-330): the caller's binary nodes already ARE what the collapse consumes --
// an internal node carries its own box and two typed child indices, a leaf its box and a primitive id -- so the import is
// one conversion launch into the scratch / reference records followed by the same collapse + pack kernel as every other build
// (the reference patches the child types in place, then runs its Collapse / CompactTasks / PackLeaves with two host syncs).
template <typename Source>
__global__ void __launch_bounds__( 256 ) importNodesKernel(
	Source src, const hiprtInternalNode* internals, const hiprtLeafNode* leafs, uint32_t leafCount, BuildControl* ctl, void* scratch, void* sortedRefs,
	uint64_t* taskQueue )
{
	const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
	if ( k < leafCount )
	{
		const hiprtLeafNode l = leafs[k];
		const uint2			w = src.refWords( l.primID );
		storeRec( sortedRefs, k, DBox{ make_float3( l.aabbMin.x, l.aabbMin.y, l.aabbMin.z ), make_float3( l.aabbMax.x, l.aabbMax.y, l.aabbMax.z ) }, w.x, w.y );
	}
	if ( k + 1 < leafCount )
	{
		const hiprtInternalNode n = internals[k];
		const uint32_t c0 = makeNodeIndex( n.childIndices[0], n.childNodeTypes[0] == hiprtBvhNodeTypeLeaf ? uint32_t( Source::LeafType ) : uint32_t( BoxType ) );
		const uint32_t c1 = makeNodeIndex( n.childIndices[1], n.childNodeTypes[1] == hiprtBvhNodeTypeLeaf ? uint32_t( Source::LeafType ) : uint32_t( BoxType ) );
		storeRec( scratch, k, DBox{ make_float3( n.aabbMin.x, n.aabbMin.y, n.aabbMin.z ), make_float3( n.aabbMax.x, n.aabbMax.y, n.aabbMax.z ) }, c0, c1 );
	}
	if ( k == 0 )
	{
		ctl->primCount = leafCount;
		ctl->rootNode  = 0; // "the root node is assumed to be at the first position" (hiprt_types.h:421-424)
		taskQueue[0]   = ( uint64_t( InvalidValue ) << 32 ) | makeNodeIndex( 0u, BoxType );
	}
}

template <typename Source>
static void runImport(
	const DeviceInfo& dev, Source src, const hiprtBvhNodeList& nodes, uint8_t* temp, const TempLayout& L, cudaStream_t stream, uint32_t* boxNodeCounter,
	uint32_t* primNodeCounter, Box4Node* boxNodes, void* leafNodes )
{
	const uint32_t n = nodes.nodeCount;
	if ( n == 1 )
	{
		singletonKernel<Source><<<1, 1, 0, stream>>>( src, primNodeCounter, boxNodes, leafNodes );
		HIPRT_CUDA_CHECK( cudaGetLastError() );
		return;
	}
	BuildControl*  ctl		 = reinterpret_cast<BuildControl*>( temp + L.control );
	uint64_t*	   taskQueue = reinterpret_cast<uint64_t*>( temp + L.taskQueue );
	const uint32_t maxBox	 = uint32_t( maxBoxNodeCount( n ) );
	HIPRT_CUDA_CHECK( cudaMemsetAsync( taskQueue, 0xFF, size_t( maxBox ) * 8, stream ) );
	importNodesKernel<Source><<<DivideRoundUp( n, 256u ), 256, 0, stream>>>(
		src, static_cast<const hiprtInternalNode*>( nodes.internalNodes ), static_cast<const hiprtLeafNode*>( nodes.leafNodes ), n, ctl, temp + L.scratch,
		temp + L.sortedRefs, taskQueue );
	dev.mark( "import", stream );
	const uint32_t warpsNeeded	 = DivideRoundUp( maxBox, 8u );
	const uint32_t blocks		 = std::max( 1u, std::min( DivideRoundUp( warpsNeeded, CollapseBlock / 32 ), dev.smCount * 16 ) );
	const uint32_t residentWarps = dev.smCount * 40;
	const uint32_t claimSize	 = std::max( 8u, std::min( 32u, ( maxBox / ( residentWarps * 16 ) ) & ~7u ) );
	collapsePackKernel<Source><<<blocks, CollapseBlock, 0, stream>>>(
		src, ctl, boxNodeCounter, primNodeCounter, temp + L.scratch, temp + L.sortedRefs, taskQueue, maxBox, claimSize, boxNodes, leafNodes );
	dev.mark( "collapse+pack", stream );
	HIPRT_CUDA_CHECK( cudaGetLastError() );
}

void buildGeometry(
	const DeviceInfo& dev, const hiprtGeometryBuildInput& input, hiprtBuildOptions options, void* tempBuffer, cudaStream_t stream, void* storage, void* qRegion )
{
	const bool import = isImport( options );
	if ( import && ( input.nodeList.nodeCount == 0 || input.nodeList.leafNodes == nullptr || ( input.nodeList.nodeCount > 1 && input.nodeList.internalNodes == nullptr ) ) )
		throw std::runtime_error( "custom BVH import needs a node list" );
	const uint32_t nMax = import ? input.nodeList.nodeCount : uint32_t( geometryPrimCount( input ) );
	if ( nMax == 0 ) throw std::runtime_error( "empty geometry" );
	const bool		 pairing  = wantsPairing( input, options );
	const bool		 balanced = usesBalancedBuilder( options );
	const bool		 sbvhTree = usesSbvhBuilder( options ) && !import && nMax > 1;
	const uint32_t	 maxLeaves = uint32_t( leafCapacity( input, options ) ); // = nMax unless a high-quality build may duplicate references
	const TempLayout L		  = tempLayout( nMax, pairing ? nMax : 0 );
	const SbvhLayout S		  = sbvhTree ? sbvhLayout( nMax, pairing ? nMax : 0, maxLeaves, maxLeaves > nMax ) : SbvhLayout{};
	uint8_t*		 temp	 = static_cast<uint8_t*>( tempBuffer );
	uint8_t*		 out	 = static_cast<uint8_t*>( storage );

	GeomHeader* header	  = reinterpret_cast<GeomHeader*>( out );
	Box4Node*	boxNodes  = reinterpret_cast<Box4Node*>( out + alignUp( sizeof( GeomHeader ) ) );
	void*		leafNodes = out + alignUp( sizeof( GeomHeader ) ) + alignUp( maxBoxNodeCount( maxLeaves ) * sizeof( Box4Node ) );
	BuildControl* ctl	  = reinterpret_cast<BuildControl*>( temp + L.control ); // (first in both layouts)

	const bool	   isMesh	= input.type == hiprtPrimitiveTypeTriangleMesh;
	const uint32_t geomType = ( input.geomType << 1 ) | ( isMesh ? 1u : 0u );
	dev.beginProfile( stream );
	initGeometryKernel<<<1, 1, 0, stream>>>( ctl, header, boxNodes, leafNodes, geometryStorageSize( input, options ), nMax, geomType );

	const uint32_t streamGrid = std::min<uint32_t>( DivideRoundUp( nMax, 256u ), dev.smCount * 8 );
	if ( isMesh )
	{
		const auto&		   m = input.primitive.triangleMesh;
		TriangleMeshSource src{};
		src.vertices	   = static_cast<const uint8_t*>( m.vertices );
		src.vertexStride   = m.vertexStride;
		src.triangles	   = ( m.triangleCount == 0 ) ? nullptr : static_cast<const uint8_t*>( m.triangleIndices );
		src.triangleStride = m.triangleStride;
		src.pairs		   = m.trianglePairCount > 0 ? static_cast<const uint2*>( m.trianglePairIndices ) : nullptr;
		if ( import )
		{
			runImport( dev, src, input.nodeList, temp, L, stream, &header->boxNodeCount, &header->primNodeCount, boxNodes, leafNodes );
			return;
		}
		if ( nMax == 1 )
		{
			singletonKernel<TriangleMeshSource><<<1, 1, 0, stream>>>( src, &header->primNodeCount, boxNodes, leafNodes );
			HIPRT_CUDA_CHECK( cudaGetLastError() );
			return;
		}
		if ( pairing )
		{
			uint2*	  pairs = reinterpret_cast<uint2*>( temp + ( sbvhTree ? S.pairs : L.pairs ) );
			uint32_t* scan	= reinterpret_cast<uint32_t*>( temp + ( sbvhTree ? S.pairScan : L.pairScan ) );
			const uint32_t tiles = DivideRoundUp( nMax, PairTile );
			HIPRT_CUDA_CHECK( cudaMemsetAsync( scan, 0, size_t( tiles ) * 4, stream ) );
			pairTrianglesKernel<<<tiles, PairBlock, 0, stream>>>( src, nMax, pairs, scan, ctl );
			src.pairs = pairs;
			dev.mark( "pair", stream );
		}
		primBoxesKernel<TriangleMeshSource><<<streamGrid, 256, 0, stream>>>( src, ctl, temp + ( sbvhTree ? S.primRefs : L.primRefs ) );
		dev.mark( "boxes", stream );
		if ( sbvhTree )
			runSbvh( dev, src, nMax, maxLeaves, maxLeaves > nMax, temp, S, stream, &header->boxNodeCount, &header->primNodeCount, boxNodes, leafNodes );
		else
			runLbvh( dev, src, balanced, nMax, temp, L, stream, &header->boxNodeCount, &header->primNodeCount, boxNodes, leafNodes );
		if ( qRegion )
		{
			compressBoxNodes( header, qRegion, uint32_t( maxBoxNodeCount( maxLeaves ) ), stream );
			dev.mark( "qnodes", stream );
			HIPRT_CUDA_CHECK( cudaGetLastError() );
		}
	}
	else
	{
		const auto&	   a = input.primitive.aabbList;
		AabbListSource src{ static_cast<const uint8_t*>( a.aabbs ), a.aabbStride };
		if ( import )
		{
			runImport( dev, src, input.nodeList, temp, L, stream, &header->boxNodeCount, &header->primNodeCount, boxNodes, leafNodes );
			return;
		}
		if ( nMax == 1 )
		{
			singletonKernel<AabbListSource><<<1, 1, 0, stream>>>( src, &header->primNodeCount, boxNodes, leafNodes );
			HIPRT_CUDA_CHECK( cudaGetLastError() );
			return;
		}
		primBoxesKernel<AabbListSource><<<streamGrid, 256, 0, stream>>>( src, ctl, temp + ( sbvhTree ? S.primRefs : L.primRefs ) );
		if ( sbvhTree )
			runSbvh( dev, src, nMax, maxLeaves, maxLeaves > nMax, temp, S, stream, &header->boxNodeCount, &header->primNodeCount, boxNodes, leafNodes );
		else
			runLbvh( dev, src, balanced, nMax, temp, L, stream, &header->boxNodeCount, &header->primNodeCount, boxNodes, leafNodes );
	}
}

// ------------------------------------------------------------------------------------------------
// not yet implemented in this revision
// ------------------------------------------------------------------------------------------------
// Scene (top-level) build: primitives are instances (reference: LbvhBuilder.cpp:126-164 with
// InstanceList<Frame>, InitSceneData BvhBuilderKernels.h:224-294). Storage layout:
// [SceneHeader][Box4Node x maxBox][InstanceNode x I][InstanceRecord x I][SRTFrame x F].
struct SceneLayout
{
	size_t boxNodes, primNodes, instances, frames, total;
};

// Instance leaves a scene build may create: one per instance, except that a high-quality build with spatial splits may reference an
// instance from several leaves (each with the part of the instance's box it is responsible for), up to half as many again.
static size_t sceneLeafCapacity( const hiprtSceneBuildInput& input, hiprtBuildOptions options )
{
	const size_t n = input.instanceCount;
	if ( usesSbvhBuilder( options ) && !isImport( options ) && !( options.buildFlags & hiprtBuildFlagBitDisableSpatialSplits ) ) return n + n / 2;
	return n;
}

static SceneLayout sceneLayout( const hiprtSceneBuildInput& input, hiprtBuildOptions options )
{
	SceneLayout	 L{};
	const size_t n = input.instanceCount, leaves = sceneLeafCapacity( input, options );
	L.boxNodes	   = alignUp( sizeof( SceneHeader ) );
	L.primNodes	   = L.boxNodes + alignUp( maxBoxNodeCount( leaves ) * sizeof( Box4Node ) );
	L.instances	   = L.primNodes + alignUp( leaves * sizeof( InstanceNode ) );
	L.frames	   = L.instances + alignUp( n * sizeof( InstanceRecord ) );
	L.total		   = L.frames + alignUp( size_t( input.frameCount ) * sizeof( SRTFrame ) );
	return L;
}

size_t sceneStorageSize( const hiprtSceneBuildInput& input, hiprtBuildOptions options ) { return sceneLayout( input, options ).total; }
size_t sceneTempSize( const hiprtSceneBuildInput& input, hiprtBuildOptions options )
{
	if ( usesSbvhBuilder( options ) && !isImport( options ) )
	{
		const size_t leaves = sceneLeafCapacity( input, options );
		return sbvhLayout( input.instanceCount, 0, leaves, leaves > input.instanceCount ).total;
	}
	return tempLayout( input.instanceCount, 0 ).total;
}

void buildScene(
	const DeviceInfo& dev, const hiprtSceneBuildInput& input, hiprtBuildOptions options, void* tempBuffer, cudaStream_t stream, void* storage )
{
	const bool	   balanced = usesBalancedBuilder( options );
	const uint32_t n		= input.instanceCount;
	if ( n == 0 ) throw std::runtime_error( "empty scene" );
	if ( input.instances == nullptr ) throw std::runtime_error( "scene without instances" );
	const SceneLayout S	   = sceneLayout( input, options );
	const TempLayout  L	   = tempLayout( n, 0 );
	uint8_t*		  temp = static_cast<uint8_t*>( tempBuffer );
	uint8_t*		  out  = static_cast<uint8_t*>( storage );
	SceneHeader*	  header	= reinterpret_cast<SceneHeader*>( out );
	Box4Node*		  boxNodes	= reinterpret_cast<Box4Node*>( out + S.boxNodes );
	InstanceNode*	  instNodes = reinterpret_cast<InstanceNode*>( out + S.primNodes );
	InstanceRecord*	  records	= reinterpret_cast<InstanceRecord*>( out + S.instances );
	SRTFrame*		  frames	= reinterpret_cast<SRTFrame*>( out + S.frames );
	BuildControl*	  ctl		= reinterpret_cast<BuildControl*>( temp + L.control );

	InstanceSource src{};
	src.instances  = static_cast<const hiprtInstance*>( input.instances );
	src.headers	   = static_cast<const hiprtTransformHeader*>( input.instanceTransformHeaders );
	src.frames	   = ( input.frameCount > 0 && input.instanceFrames != nullptr ) ? frames : nullptr;
	src.masks	   = static_cast<const uint32_t*>( input.instanceMasks );
	src.frameCount = input.frameCount;

	const uint32_t work = std::max( n, input.frameCount );
	initSceneKernel<<<DivideRoundUp( work, 256u ), 256, 0, stream>>>(
		ctl, header, boxNodes, instNodes, records, frames, S.total, src, n, input.instanceFrames, input.frameType == hiprtFrameTypeMatrix );
	if ( isImport( options ) )
	{
		// custom top-level BVH: one leaf per instance (the instance-node array is sized by the instance count)
		if ( input.nodeList.nodeCount != n || input.nodeList.leafNodes == nullptr || ( n > 1 && input.nodeList.internalNodes == nullptr ) )
			throw std::runtime_error( "custom scene BVH import needs a node list with exactly one leaf per instance" );
		runImport( dev, src, input.nodeList, temp, L, stream, &header->boxNodeCount, &header->primNodeCount, boxNodes, instNodes );
		return;
	}
	if ( n == 1 )
	{
		singletonKernel<InstanceSource><<<1, 1, 0, stream>>>( src, &header->primNodeCount, boxNodes, instNodes );
		HIPRT_CUDA_CHECK( cudaGetLastError() );
		return;
	}
	const uint32_t streamGrid = std::min<uint32_t>( DivideRoundUp( n, 256u ), dev.smCount * 8 );
	if ( usesSbvhBuilder( options ) )
	{
		// instances are cut like custom primitives: the half of the instance's world box on either side of the plane (the reference
		// cuts along the boxes of the instance's root children, InstanceList.h:125-166: tighter, same guarantee)
		const uint32_t	 leaves = uint32_t( sceneLeafCapacity( input, options ) );
		const SbvhLayout Q		= sbvhLayout( n, 0, leaves, leaves > n );
		primBoxesKernel<InstanceSource><<<streamGrid, 256, 0, stream>>>( src, ctl, temp + Q.primRefs );
		runSbvh( dev, src, n, leaves, leaves > n, temp, Q, stream, &header->boxNodeCount, &header->primNodeCount, boxNodes, instNodes );
		return;
	}
	primBoxesKernel<InstanceSource><<<streamGrid, 256, 0, stream>>>( src, ctl, temp + L.primRefs );
	runLbvh( dev, src, balanced, n, temp, L, stream, &header->boxNodeCount, &header->primNodeCount, boxNodes, instNodes );
}
// Refit in place (hiprtBuildOperationUpdate): same primitive count and order as the build, moved positions. The kernels take
// the node arrays from the device header, so a compacted, loaded or imported structure is updated where it actually lies.
void updateGeometry(
	const DeviceInfo& dev, const hiprtGeometryBuildInput& input, hiprtBuildOptions options, void*, cudaStream_t stream, void* storage, void* qRegion )
{
	(void)dev;
	// (an upper bound of the box-node count, for the grids below; whatever the options say now, the structure may have been built
	// with duplicated references or imported)
	const uint32_t nMax = uint32_t( std::max( { leafCapacity( input, options ), importLeafCount( input ), geometryPrimCount( input ) + geometryPrimCount( input ) / 2 } ) );
	if ( nMax == 0 ) throw std::runtime_error( "empty geometry" );
	const GeomHeader* header = static_cast<const GeomHeader*>( storage );
	const uint32_t	  grid	 = DivideRoundUp( uint32_t( maxBoxNodeCount( nMax ) ), 128u ); // upper bound; trimmed by header->boxNodeCount
	if ( input.type == hiprtPrimitiveTypeTriangleMesh )
	{
		const auto&		   m = input.primitive.triangleMesh;
		TriangleMeshSource src{};
		src.vertices	   = static_cast<const uint8_t*>( m.vertices );
		src.vertexStride   = m.vertexStride;
		src.triangles	   = ( m.triangleCount == 0 ) ? nullptr : static_cast<const uint8_t*>( m.triangleIndices );
		src.triangleStride = m.triangleStride;
		src.pairs		   = nullptr; // the pairing is read back from the leaves
		refitKernel<TriangleMeshSource><<<grid, 128, 0, stream>>>( src, header, &header->boxNodeCount );
		compressBoxNodes( header, qRegion, uint32_t( maxBoxNodeCount( nMax ) ), stream ); // (no-op without a Q region)
	}
	else
	{
		const auto&	   a = input.primitive.aabbList;
		AabbListSource src{ static_cast<const uint8_t*>( a.aabbs ), a.aabbStride };
		refitKernel<AabbListSource><<<grid, 128, 0, stream>>>( src, header, &header->boxNodeCount );
	}
	HIPRT_CUDA_CHECK( cudaGetLastError() );
}

void updateScene(
	const DeviceInfo& dev, const hiprtSceneBuildInput& input, hiprtBuildOptions, void*, cudaStream_t stream, void* storage )
{
	(void)dev;
	const uint32_t n = input.instanceCount;
	if ( n == 0 ) throw std::runtime_error( "empty scene" );
	const SceneHeader* header = static_cast<const SceneHeader*>( storage );
	InstanceSource	   src{};
	src.instances  = static_cast<const hiprtInstance*>( input.instances );
	src.headers	   = static_cast<const hiprtTransformHeader*>( input.instanceTransformHeaders );
	// non-null = "this scene has converted frames"; the kernels replace it by header->frames
	src.frames	   = ( input.frameCount > 0 && input.instanceFrames != nullptr ) ? reinterpret_cast<const SRTFrame*>( storage ) : nullptr;
	src.masks	   = static_cast<const uint32_t*>( input.instanceMasks );
	src.frameCount = input.frameCount;
	const uint32_t work = std::max( n, input.frameCount );
	updateSceneInputsKernel<<<DivideRoundUp( work, 256u ), 256, 0, stream>>>(
		header, src, n, input.instanceFrames, input.frameType == hiprtFrameTypeMatrix );
	// (grid: an upper bound of the box-node count whatever builder made the scene; trimmed by header->boxNodeCount)
	refitKernel<InstanceSource><<<DivideRoundUp( uint32_t( maxBoxNodeCount( n + n / 2 ) ), 128u ), 128, 0, stream>>>( src, header, &header->boxNodeCount );
	HIPRT_CUDA_CHECK( cudaGetLastError() );
}

// Batch of small geometries: the caller's host arrays (inputs, storage pointers) are copied into the
// temporary buffer, then ONE launch builds everything (reference: BatchBuilder.h:58-132).
size_t batchTempSize( uint32_t count )
{
	return alignUp( size_t( count ) * sizeof( hiprtGeometryBuildInput ) ) + alignUp( size_t( count ) * sizeof( void* ) );
}

void buildGeometriesBatch(
	const DeviceInfo&			   dev,
	const hiprtGeometryBuildInput* inputs,
	uint32_t					   count,
	hiprtBuildOptions			   options,
	void*						   tempBuffer,
	cudaStream_t				   stream,
	void* const*				   storages )
{
	(void)options;
	if ( count == 0 ) return;
	for ( uint32_t i = 0; i < count; ++i )
		if ( geometryPrimCount( inputs[i] ) > BatchMaxPrims ) throw std::runtime_error( "batch build: geometry exceeds hiprtMaxBatchBuildMaxPrimCount" );
	uint8_t* temp	  = static_cast<uint8_t*>( tempBuffer );
	auto*	 dInputs  = reinterpret_cast<hiprtGeometryBuildInput*>( temp );
	auto*	 dStorage = reinterpret_cast<void**>( temp + alignUp( size_t( count ) * sizeof( hiprtGeometryBuildInput ) ) );
	// pageable sources: the runtime stages them before returning, so the caller's arrays may go away
	HIPRT_CUDA_CHECK( cudaMemcpyAsync( dInputs, inputs, size_t( count ) * sizeof( hiprtGeometryBuildInput ), cudaMemcpyHostToDevice, stream ) );
	HIPRT_CUDA_CHECK( cudaMemcpyAsync( dStorage, storages, size_t( count ) * sizeof( void* ), cudaMemcpyHostToDevice, stream ) );
	// (per CUDA context, see runLbvh: set every time)
	HIPRT_CUDA_CHECK( cudaFuncSetAttribute( batchBuildKernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int( sizeof( BatchSmem ) ) ) );
	dev.beginProfile( stream );
	const uint32_t gx = std::min( count, 65535u * 32u );
	const dim3	   grid( gx, DivideRoundUp( count, gx ) );
	batchBuildKernel<<<grid, BatchBlock, sizeof( BatchSmem ), stream>>>( dInputs, dStorage, count );
	dev.mark( "batch", stream );
	HIPRT_CUDA_CHECK( cudaGetLastError() );
}
} // namespace hiprt
