// rb_internal.h — handle layouts and helpers shared by the translation units of librendu_b200.so.
#pragma once
#include "../../include/rendu_b200.h"
#include "rb_types.h"
#include "rb_wide.h"

#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <memory>
#include <string>
#include <thread>
#include <utility>
#include <vector>

// ---------------------------------------------------------------- errors
void rb_set_error(const char * fmt, ...);
rb_status rb_fail(rb_status code, const char * fmt, ...);
bool rb_have_device();

#define RB_CUDA(call)                                                                                              \
	do {                                                                                                           \
		cudaError_t _e = (call);                                                                                   \
		if(_e != cudaSuccess) {                                                                                    \
			return rb_fail(RB_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(_e), __FILE__, __LINE__); \
		}                                                                                                          \
	} while(0)

#define RB_REQUIRE_DEVICE()                                                                                     \
	do {                                                                                                        \
		if(!rb_have_device()) return rb_fail(RB_ERR_NO_DEVICE, "no CUDA device available (no CPU fallback)");     \
	} while(0)

// ---------------------------------------------------------------- host helpers
// Splits [0, n) into contiguous ranges over a few host threads (large meshes: packing and per-triangle set-up are
// memory-bound loops over hundreds of MB). fn(begin, end) must only touch its own range.
template<typename F>
void rb_parallel_for(size_t n, size_t grain, F && fn) {
	const size_t hw = std::thread::hardware_concurrency();
	size_t nt		= hw > 8 ? 8 : (hw ? hw : 1);
	if(grain && n / grain < nt) nt = n / grain;
	if(nt <= 1) {
		fn(size_t(0), n);
		return;
	}
	std::vector<std::thread> pool;
	const size_t span = (n + nt - 1) / nt;
	for(size_t t = 0; t < nt; ++t) {
		const size_t b = t * span, e = (b + span < n) ? (b + span) : n;
		if(b >= e) break;
		pool.emplace_back([&fn, b, e] { fn(b, e); });
	}
	for(auto & th : pool) th.join();
}

// std::vector whose resize() leaves new elements uninitialised: a 10M-triangle mesh is 640 MB of records that the
// next (multi-threaded) loop overwrites anyway; value-initialising them first costs a single-threaded pass and all
// the page faults on one core.
template<typename T>
struct RbNoInitAlloc : std::allocator<T> {
	template<typename U>
	struct rebind {
		using other = RbNoInitAlloc<U>;
	};
	template<typename U, typename... Args>
	void construct(U * p, Args &&... args) {
		if(sizeof...(Args) == 0) ::new(static_cast<void *>(p)) U;
		else ::new(static_cast<void *>(p)) U(std::forward<Args>(args)...);
	}
};
using RbTriVec = std::vector<rb_tri, RbNoInitAlloc<rb_tri>>;

// ---------------------------------------------------------------- raycaster handle
struct rb_raycaster {
	// Host mirror in the reference's own format (kept so RaycasterVisualisation-style consumers can read it back).
	std::vector<float> verts; // merged world-space vertices, xyz
	RbTriVec tris;
	std::vector<rb_node> nodes;
	uint32_t meshCount = 0;
	std::vector<uint64_t> meshFirst, meshTris; // per mesh: first triangle and triangle count (stable across rebuilds)
	std::vector<uint64_t> meshVertFirst, meshVerts; // per mesh: first vertex and vertex count in `verts` (addMesh order)
	int maxDepth	   = 0; // max number of interior nodes on a root-to-leaf path
	bool built		   = false;
	// rb_raycaster_tree_cost: surface-area cost of the tree when it was built and after the last refit
	double areaAtBuild = 0.0, areaNow = 0.0;
	uint64_t areaSerial = ~uint64_t(0);
	bool areaStale = false;
	uint64_t buildSerial = 0; // counts successful builds / uploads: the triangle order changed (rb_scene keys its per-triangle records to it)
	bool hostMirrorStale = false; // device tree newer than the host mirror (after LBVH build / refit)
	bool triBoxesStale	 = false; // `verts` changed since the boxes in `tris` were computed (refit): redone before a rebuild

	// Device tree.
	int device		  = -1; // CUDA device holding the tree (-1: nothing uploaded yet)
	float4 * d_nodes  = nullptr;
	float4 * d_tris	  = nullptr;
	float4 * d_roots  = nullptr;
	float4 * d_rootBlocks = nullptr; // union boxes of aligned root blocks (many-mesh scenes), see rb_next_root
	bool rootBlocksWanted = false;	  // decided at upload (rb_root_blocks_worthwhile), kept across refits
	size_t rootBlocksCapacity = 0;	  // entries d_rootBlocks was allocated for
	float4 * d_rootTree	  = nullptr; // tree over the root boxes (9..64 meshes): nroots - 1 nodes of 4 float4, re-fitted with the roots
	int rootTreeNodes	  = 0;
	int rootTreeDepth	  = 0;		 // interior levels of that tree (stack entries its walk may need)
	size_t nInterior  = 0;
	size_t nTris	  = 0;
	RbDeviceBvh dev	  = {};
	// LBVH bookkeeping (device): parent links and leaf->node map for refits, world-space vertices, vertex ids.
	int * d_parent	   = nullptr;	 // per interior node: (parent << 1) | side, or ~rootSlot
	int * d_leafParent = nullptr;	 // per leaf: same encoding
	int * d_leafCode   = nullptr;	 // per leaf: (first triangle << 3) | (count - 1)
	unsigned int * d_flags = nullptr; // per interior node: arrival counter of the bottom-up fit
	float * d_verts	   = nullptr;
	uint32_t * d_triVerts = nullptr; // 3 vertex ids per triangle in hierarchy order
	uint32_t * d_triIds	  = nullptr; // (localId, meshId) per triangle in hierarchy order
	size_t nVerts	   = 0;
	size_t nLeaves	   = 0;
	size_t lbvhLeaves  = 0;			 // LBVH trees: leaves of the finished tree (ranges of up to RB_LBVH_LEAF triangles)
	int buildMode	   = RB_BUILD_REFERENCE;

	// Eight-wide quantised copy of the host-built tree (rb_wide.h; RB_TRACE_WIDE): host arrays and device copy.
	std::unique_ptr<RbWideBvh> wide;
	uint4 * d_wideNodes	 = nullptr;
	uint4 * d_wideNodesOct = nullptr; // the same nodes, slots by direction (RbWideBvh::nodesByOctant)
	float4 * d_wideTris	 = nullptr;
	int * d_wideRoots	 = nullptr;
	RbDeviceWide devWide = {};
	bool wideOnDevice	 = false;

	// Work-queue cursors of the persistent kernels: a ring of counters, one taken per launch (no allocation on the
	// launch path; a slot is reused only after RB_CURSOR_RING further launches).
	unsigned int * d_cursors = nullptr;
	mutable std::atomic<uint32_t> cursorNext{0};

	double buildDeviceMs = 0.0, refitDeviceMs = 0.0;

	int traceMode = RB_TRACE_EXACT;
	bool persistent	   = true; // persistent threads + warp work queue (false: one ray per thread)
	int leafVotes	   = 6;	   // > 0: phase-voting scheduler, leaf phase runs once this many lanes wait for it (0: off)
	int nodeSteps	   = 6;	   // phase-voting scheduler: node steps per vote
	bool dual		   = false; // two rays per lane (rb_k_trace_dual)
	bool ring		   = false; // prepared-ray ring per warp (rb_k_trace_ring, experiment)
	int wideOctant	   = 0;	   // RB_TRACE_WIDE: 1 = closest-hit queries visit slots by direction (front to back), 2 = any-hit queries too (experiment)
	bool widePhased	   = false; // RB_TRACE_WIDE under the phase-voting scheduler (rb_k_trace_wide_phased, experiment)
	int dualRefill	   = 16;   // dual: refill once this many of a warp's 64 slots are free
	int fetchThreshold = 12;   // refill a warp from the queue when fewer lanes than this are still traversing
	mutable std::atomic<uint64_t> launches{0};
};

// Host builder mirroring Raycaster::updateHierarchy (rb_bvh_host.cpp).
void rb_host_build_reference(rb_raycaster * rc);
// Depth of the host-format tree (max interior nodes on a path) or -1 if malformed.
int rb_host_tree_depth(const rb_raycaster * rc, std::string * why);

// Flatten the host-format tree into the device layout (host arrays of float4 packed as floats).
struct RbFlatBvh {
	std::vector<float> nodes; // 16 floats per interior node
	std::vector<float> tris;  // 12 floats per triangle
	std::vector<float> roots; // 8 floats per root
	std::vector<uint32_t> triVerts;
	std::vector<int> parent;	 // per interior node: (parent << 1) | side, or ~rootSlot for a root
	std::vector<int> leafParent; // per leaf, same encoding
	std::vector<int> leafCode;	 // per leaf: (first triangle << 3) | (count - 1)
	std::vector<uint32_t> triIds; // (localId, meshId) per triangle
	size_t dummyTris = 0;		  // degenerate records appended to `tris` (the leaf an empty mesh's root points at); not counted as triangles
};
bool rb_host_flatten(const rb_raycaster * rc, RbFlatBvh & out, std::string * why);

// Upload a flattened tree into the handle (rb_raycaster.cu).
rb_status rb_raycaster_upload_flat(rb_raycaster * rc, const RbFlatBvh & flat);
void rb_raycaster_free_device(rb_raycaster * rc);

// GPU LBVH build / refit (rb_lbvh.cu).
rb_status rb_lbvh_build(rb_raycaster * rc);
rb_status rb_lbvh_refit(rb_raycaster * rc, const float * worldPositions, size_t nverts);
rb_status rb_lbvh_download(const rb_raycaster * rc, rb_node * nodes, rb_tri * tris, float * verts);
// (Re)builds the root-block boxes from the root table on the device; called after an upload and after every fit.
rb_status rb_build_root_levels(rb_raycaster * rc, cudaStream_t s);
bool rb_root_blocks_worthwhile(const float * roots8, size_t nroots);
// Entries of a ray's stack column: the deepest hierarchy, or the walk of the root tree (one pending node per level) if that is
// deeper (many tiny meshes), behind the mask words of a big root tree.
inline int rb_stack_entries(const rb_raycaster * rc) {
	return rc->dev.rootWords + (rc->maxDepth + 1 > rc->rootTreeDepth + 1 ? rc->maxDepth + 1 : rc->rootTreeDepth + 1);
}
rb_status rb_root_tree_topology(rb_raycaster * rc, const float * roots8, size_t nroots); // at upload; boxes by rb_build_root_levels

#define RB_TRACE_BLOCK_THREADS 128
#define RB_CURSOR_RING 1024
// Launch helpers used by the path tracer (rb_raycaster.cu).
size_t rb_trace_smem_bytes(const rb_raycaster * rc, int blockThreads);
