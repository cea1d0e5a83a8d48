// rb_lbvh.cu — GPU hierarchy build and refit (north_star "GPU LBVH build: Morton codes, radix sort, Karras hierarchy,
// SAH-free refit"). It sits beside RB_BUILD_REFERENCE (rb_bvh_host.cpp), which reproduces the reference's own tree
// (Raycaster::updateHierarchy, src/engine/raycaster/Raycaster.cpp:54-142); this file has no reference counterpart.
//
//   build : triangle boxes + centroid bounds -> 63-bit Morton codes -> LSD radix sort (8 passes of 8 bits, hand-written:
//           per-block histograms, one scan over the digit-major table, stable scatter ranked with __match_any_sync)
//           -> Karras 2012 hierarchy, one thread per internal node -> bottom-up AABB fit with one atomic flag per
//           node -> triangles re-gathered in sorted order. One tree over all meshes, one triangle per leaf.
//   refit : same bottom-up pass over the existing topology (either build mode) after the vertices moved.
// Boxes are unions of exact fp32 triangle boxes (min/max only), so a box is bit-identical to what the reference's
// BoundingBox::merge chain (Bounds.cpp:12-20) would give for the same triangle set.
#include "rb_internal.h"
#include "rb_math.cuh"

#include <algorithm>
#include <cfloat>
#include <climits>
#include <memory>
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <vector>

#define RB_SORT_BLOCK 256
#define RB_SORT_ITEMS 4
#define RB_SORT_TILE (RB_SORT_BLOCK * RB_SORT_ITEMS)

// ------------------------------------------------------------------------------------------------ radix sort
__global__ void __launch_bounds__(RB_SORT_BLOCK) rb_k_sort_hist(const uint64_t * __restrict__ keys, uint32_t n, int shift, uint32_t * __restrict__ table, uint32_t nblocks) {
	__shared__ uint32_t hist[256];
	hist[threadIdx.x] = 0;
	__syncthreads();
	const uint32_t base = blockIdx.x * RB_SORT_TILE;
	for(int j = 0; j < RB_SORT_ITEMS; ++j) {
		const uint32_t i = base + j * RB_SORT_BLOCK + threadIdx.x;
		if(i < n) atomicAdd(&hist[(keys[i] >> shift) & 255u], 1u);
	}
	__syncthreads();
	table[threadIdx.x * nblocks + blockIdx.x] = hist[threadIdx.x]; // digit-major: a scan gives global offsets
}

// Exclusive scan of a u32 array in three steps (tile sums, scan of the sums by one block, add back).
#define RB_SCAN_TILE 2048
__global__ void __launch_bounds__(256) rb_k_scan_tiles(uint32_t * __restrict__ data, uint32_t n, uint32_t * __restrict__ sums) {
	__shared__ uint32_t s[256];
	const uint32_t base = blockIdx.x * RB_SCAN_TILE + threadIdx.x * 8;
	uint32_t v[8], total = 0;
	for(int k = 0; k < 8; ++k) {
		v[k] = (base + k < n) ? data[base + k] : 0u;
		total += v[k];
	}
	s[threadIdx.x] = total;
	__syncthreads();
	for(int off = 1; off < 256; off <<= 1) {
		const uint32_t t = (threadIdx.x >= off) ? s[threadIdx.x - off] : 0u;
		__syncthreads();
		s[threadIdx.x] += t;
		__syncthreads();
	}
	uint32_t run = s[threadIdx.x] - total;
	for(int k = 0; k < 8; ++k) {
		if(base + k < n) data[base + k] = run;
		run += v[k];
	}
	if(threadIdx.x == 255) sums[blockIdx.x] = s[255];
}
__global__ void __launch_bounds__(256) rb_k_scan_sums(uint32_t * __restrict__ sums, uint32_t n) {
	__shared__ uint32_t s[256];
	__shared__ uint32_t carry;
	if(threadIdx.x == 0) carry = 0;
	__syncthreads();
	for(uint32_t base = 0; base < n; base += 256) {
		const uint32_t i = base + threadIdx.x;
		const uint32_t v = (i < n) ? sums[i] : 0u;
		s[threadIdx.x]	 = v;
		__syncthreads();
		for(int off = 1; off < 256; off <<= 1) {
			const uint32_t t = (threadIdx.x >= off) ? s[threadIdx.x - off] : 0u;
			__syncthreads();
			s[threadIdx.x] += t;
			__syncthreads();
		}
		if(i < n) sums[i] = carry + s[threadIdx.x] - v;
		__syncthreads();
		if(threadIdx.x == 255) carry += s[255];
		__syncthreads();
	}
}
__global__ void __launch_bounds__(256) rb_k_scan_add(uint32_t * __restrict__ data, uint32_t n, const uint32_t * __restrict__ sums) {
	const uint32_t base = blockIdx.x * RB_SCAN_TILE + threadIdx.x * 8;
	const uint32_t add	= sums[blockIdx.x];
	for(int k = 0; k < 8; ++k) {
		if(base + k < n) data[base + k] += add;
	}
}

// Stable scatter of one digit. Keys of a tile are visited warp by warp, and inside a warp in 32-wide stripes, so the
// running per-warp digit counts give each key its rank among equal digits in input order.
__global__ void __launch_bounds__(RB_SORT_BLOCK) rb_k_sort_scatter(const uint64_t * __restrict__ keysIn, const uint32_t * __restrict__ valsIn, uint64_t * __restrict__ keysOut, uint32_t * __restrict__ valsOut, uint32_t n, int shift, const uint32_t * __restrict__ table, uint32_t nblocks) {
	__shared__ uint32_t warpHist[RB_SORT_BLOCK / 32][256];
	const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
	for(int w = 0; w < RB_SORT_BLOCK / 32; ++w) warpHist[w][threadIdx.x] = 0;
	__syncthreads();
	const uint32_t warpBase = blockIdx.x * RB_SORT_TILE + warp * (32 * RB_SORT_ITEMS);
	uint64_t key[RB_SORT_ITEMS];
	uint32_t val[RB_SORT_ITEMS], rank[RB_SORT_ITEMS], digit[RB_SORT_ITEMS];
	for(int j = 0; j < RB_SORT_ITEMS; ++j) {
		const uint32_t i = warpBase + j * 32 + lane;
		const bool valid = i < n;
		key[j]	 = valid ? keysIn[i] : ~0ull;
		val[j]	 = valid ? valsIn[i] : 0u;
		digit[j] = valid ? uint32_t((key[j] >> shift) & 255u) : 256u; // 256: padding lanes match only each other
		const uint32_t peers = __match_any_sync(0xffffffffu, digit[j]);
		uint32_t before = 0;
		if(valid) before = warpHist[warp][digit[j]];
		__syncwarp();
		rank[j] = before + __popc(peers & ((1u << lane) - 1u));
		if(valid && lane == uint32_t(__ffs(peers) - 1)) warpHist[warp][digit[j]] = before + __popc(peers);
		__syncwarp();
	}
	__syncthreads();
	// Exclusive prefix over the warps for digit = threadIdx.x, plus the block's global offset for that digit.
	{
		uint32_t run = table[threadIdx.x * nblocks + blockIdx.x];
		for(int w = 0; w < RB_SORT_BLOCK / 32; ++w) {
			const uint32_t c	   = warpHist[w][threadIdx.x];
			warpHist[w][threadIdx.x] = run;
			run += c;
		}
	}
	__syncthreads();
	for(int j = 0; j < RB_SORT_ITEMS; ++j) {
		if(digit[j] < 256u) {
			const uint32_t dst = warpHist[warp][digit[j]] + rank[j];
			keysOut[dst] = key[j];
			valsOut[dst] = val[j];
		}
	}
}

struct RbSortScratch {
	uint64_t * keysAlt = nullptr;
	uint32_t * valsAlt = nullptr;
	uint32_t * table   = nullptr;
	uint32_t * sums	   = nullptr;
};

// Sorts (keys, vals) ascending by the low `bits` bits of the key; the result ends in the input buffers.
static rb_status rb_radix_sort(uint64_t * keys, uint32_t * vals, uint32_t n, int bits, cudaStream_t s, uint64_t * launches) {
	if(n < 2) return RB_OK;
	const uint32_t nblocks = (n + RB_SORT_TILE - 1) / RB_SORT_TILE;
	const uint32_t tableN  = 256u * nblocks;
	const uint32_t ntiles  = (tableN + RB_SCAN_TILE - 1) / RB_SCAN_TILE;
	RbSortScratch sc;
	// Stream-ordered scratch: no device-wide synchronisation. The pool keeps up to 256 MB cached across
	// synchronisation points (default: everything is returned to the driver), so the next build or refit reuses it.
	{
		int dev = 0;
		cudaMemPool_t pool = nullptr;
		if(cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
			uint64_t keep = 0;
			if(cudaMemPoolGetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep) == cudaSuccess && keep < (uint64_t(256) << 20)) {
				keep = uint64_t(256) << 20;
				cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
			}
		}
	}
	RB_CUDA(cudaMallocAsync(&sc.keysAlt, size_t(n) * sizeof(uint64_t), s));
	RB_CUDA(cudaMallocAsync(&sc.valsAlt, size_t(n) * sizeof(uint32_t), s));
	RB_CUDA(cudaMallocAsync(&sc.table, size_t(tableN) * sizeof(uint32_t), s));
	RB_CUDA(cudaMallocAsync(&sc.sums, size_t(ntiles) * sizeof(uint32_t), s));
	uint64_t *kin = keys, *kout = sc.keysAlt;
	uint32_t *vin = vals, *vout = sc.valsAlt;
	int passes = (bits + 7) / 8;
	if(passes & 1) ++passes; // an even number of passes leaves the result in the caller's buffers
	for(int p = 0; p < passes; ++p) {
		const int shift = 8 * p;
		rb_k_sort_hist<<<nblocks, RB_SORT_BLOCK, 0, s>>>(kin, n, shift, sc.table, nblocks);
		rb_k_scan_tiles<<<ntiles, 256, 0, s>>>(sc.table, tableN, sc.sums);
		rb_k_scan_sums<<<1, 256, 0, s>>>(sc.sums, ntiles);
		rb_k_scan_add<<<ntiles, 256, 0, s>>>(sc.table, tableN, sc.sums);
		rb_k_sort_scatter<<<nblocks, RB_SORT_BLOCK, 0, s>>>(kin, vin, kout, vout, n, shift, sc.table, nblocks);
		std::swap(kin, kout);
		std::swap(vin, vout);
		*launches += 5;
	}
	cudaError_t e = cudaGetLastError();
	cudaFreeAsync(sc.keysAlt, s);
	cudaFreeAsync(sc.valsAlt, s);
	cudaFreeAsync(sc.table, s);
	cudaFreeAsync(sc.sums, s);
	if(e != cudaSuccess) return rb_fail(RB_ERR_CUDA, "radix sort failed: %s", cudaGetErrorString(e));
	return RB_OK;
}

// ------------------------------------------------------------------------------------------------ build kernels
RB_D unsigned int rb_float_ordered(float f) { // monotone float -> uint map for atomicMin/Max
	const unsigned int u = __float_as_uint(f);
	return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
RB_D float rb_float_unordered(unsigned int u) {
	return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

RB_D void rb_tri_box(const float * __restrict__ verts, uint32_t a, uint32_t b, uint32_t c, V3 & lo, V3 & hi) {
	const V3 va = V3{verts[3 * a], verts[3 * a + 1], verts[3 * a + 2]};
	const V3 vb = V3{verts[3 * b], verts[3 * b + 1], verts[3 * b + 2]};
	const V3 vc = V3{verts[3 * c], verts[3 * c + 1], verts[3 * c + 2]};
	lo = glm_min3(glm_min3(va, vb), vc); // BoundingBox(v0,v1,v2), Bounds.cpp:12-15
	hi = glm_max3(glm_max3(va, vb), vc);
}

__global__ void rb_k_lbvh_bounds(const float * __restrict__ verts, const uint32_t * __restrict__ triVerts, uint32_t ntris, unsigned int * __restrict__ bounds6) {
	// Grid-stride partial min/max per thread, shuffle reduction per warp, one atomic per warp and component.
	unsigned int lo[3] = {0xffffffffu, 0xffffffffu, 0xffffffffu}, hi[3] = {0u, 0u, 0u};
	for(uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < ntris; i += gridDim.x * blockDim.x) {
		V3 a, b;
		rb_tri_box(verts, triVerts[3 * i], triVerts[3 * i + 1], triVerts[3 * i + 2], a, b);
		const V3 c = 0.5f * (a + b);
		const unsigned int k[3] = {rb_float_ordered(c.x), rb_float_ordered(c.y), rb_float_ordered(c.z)};
		for(int d = 0; d < 3; ++d) {
			lo[d] = min(lo[d], k[d]);
			hi[d] = max(hi[d], k[d]);
		}
	}
	for(int off = 16; off > 0; off >>= 1) {
		for(int d = 0; d < 3; ++d) {
			lo[d] = min(lo[d], __shfl_down_sync(0xffffffffu, lo[d], off));
			hi[d] = max(hi[d], __shfl_down_sync(0xffffffffu, hi[d], off));
		}
	}
	if((threadIdx.x & 31u) == 0u) {
		for(int d = 0; d < 3; ++d) {
			atomicMin(bounds6 + d, lo[d]);
			atomicMax(bounds6 + 3 + d, hi[d]);
		}
	}
}

RB_D uint64_t rb_expand21(uint64_t v) { // spread 21 bits so that two zero bits separate consecutive bits
	v &= 0x1fffffull;
	v = (v | (v << 32)) & 0x1f00000000ffffull;
	v = (v | (v << 16)) & 0x1f0000ff0000ffull;
	v = (v | (v << 8)) & 0x100f00f00f00f00full;
	v = (v | (v << 4)) & 0x10c30c30c30c30c3ull;
	v = (v | (v << 2)) & 0x1249249249249249ull;
	return v;
}

// `cubic`: all three axes are quantised with the cell size of the LONGEST one, so the grid cells are cubes. With per-axis
// cell sizes a flat scene (a terrain: C5's displaced grid is 2 x 0.1 x 2) spends every third level of the tree splitting its
// thin axis — two children that cover the same footprint, both of which every ray below them has to enter. In the cubic
// grid the high bits of the thin axis are equal for all triangles; Karras' construction splits on the highest DIFFERING bit
// of a range, so those levels simply do not exist.
__global__ void rb_k_lbvh_morton(const float * __restrict__ verts, const uint32_t * __restrict__ triVerts, uint32_t ntris, const unsigned int * __restrict__ bounds6, uint64_t * __restrict__ keys, uint32_t * __restrict__ vals, int cubic) {
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if(i >= ntris) return;
	V3 lo, hi;
	rb_tri_box(verts, triVerts[3 * i], triVerts[3 * i + 1], triVerts[3 * i + 2], lo, hi);
	const V3 c	= 0.5f * (lo + hi);
	const V3 bl = V3{rb_float_unordered(bounds6[0]), rb_float_unordered(bounds6[1]), rb_float_unordered(bounds6[2])};
	const V3 bh = V3{rb_float_unordered(bounds6[3]), rb_float_unordered(bounds6[4]), rb_float_unordered(bounds6[5])};
	float ex = fmaxf(bh.x - bl.x, 1e-30f), ey = fmaxf(bh.y - bl.y, 1e-30f), ez = fmaxf(bh.z - bl.z, 1e-30f);
	if(cubic) ex = ey = ez = fmaxf(ex, fmaxf(ey, ez));
	const double scale = 2097151.0; // 2^21 - 1
	const uint64_t qx = uint64_t(fmin(fmax(double(c.x - bl.x) / double(ex) * scale, 0.0), scale));
	const uint64_t qy = uint64_t(fmin(fmax(double(c.y - bl.y) / double(ey) * scale, 0.0), scale));
	const uint64_t qz = uint64_t(fmin(fmax(double(c.z - bl.z) / double(ez) * scale, 0.0), scale));
	keys[i] = (rb_expand21(qx) << 2) | (rb_expand21(qy) << 1) | rb_expand21(qz);
	vals[i] = i;
}

// Karras 2012, "Maximizing parallelism in the construction of BVHs": common-prefix length with the index as tie-break.
RB_D int rb_delta(const uint64_t * __restrict__ keys, int n, int i, int j) {
	if(j < 0 || j >= n) return -1;
	const uint64_t a = keys[i], b = keys[j];
	if(a == b) return 64 + __clz(i ^ j);
	return __clzll((long long)(a ^ b));
}

#define RB_LBVH_DEAD INT_MIN
// Leaves of the finished tree = entries of the leaf table that carry a code.
__global__ void rb_k_lbvh_count_leaves(const int * __restrict__ leafCode, int n, unsigned int * __restrict__ out) {
	unsigned int c = 0;
	for(int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) c += leafCode[i] >= 0 ? 1u : 0u;
	for(int off = 16; off > 0; off >>= 1) c += __shfl_down_sync(0xffffffffu, c, off);
	if((threadIdx.x & 31u) == 0 && c) atomicAdd(out, c);
}

// One thread per internal node: children and parent links. Child references use the traversal encoding (rb_types.h).
// `leafMax` (1..8) collapses the bottom of the tree: a child whose range holds at most that many triangles becomes ONE leaf
// (the triangles of a range are contiguous in the sorted order, which is all a leaf code needs), so a ray pays one node
// visit for them instead of a little binary tree of visits, and the part of the 64-byte node array that is ever touched
// shrinks by the same factor — at 10 M triangles the tree is several times the L2. Internal nodes inside a collapsed range
// are dead: they keep valid references (a download stays well formed) but link to nobody and are never fitted.
__global__ void rb_k_lbvh_hierarchy(const uint64_t * __restrict__ keys, int n, int leafMax, int * __restrict__ childRefs /* 2 per internal */, int * __restrict__ interiorParent, int * __restrict__ leafParent, int * __restrict__ leafCode) {
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if(i >= n - 1) return;
	const int d		  = (rb_delta(keys, n, i, i + 1) - rb_delta(keys, n, i, i - 1)) >= 0 ? 1 : -1;
	const int deltaMin = rb_delta(keys, n, i, i - d);
	int lmax = 2;
	while(rb_delta(keys, n, i, i + lmax * d) > deltaMin) lmax <<= 1;
	int l = 0;
	for(int t = lmax >> 1; t >= 1; t >>= 1) {
		if(rb_delta(keys, n, i, i + (l + t) * d) > deltaMin) l += t;
	}
	const int j			= i + l * d;
	const int deltaNode = rb_delta(keys, n, i, j);
	int s = 0;
	int t = l;
	do {
		t = (t + 1) >> 1;
		if(rb_delta(keys, n, i, i + (s + t) * d) > deltaNode) s += t;
	} while(t > 1);
	const int split = i + s * d + min(d, 0);
	const int lo = min(i, j), hi = max(i, j);
	const int countL = split - lo + 1, countR = hi - split;
	const bool leftLeaf = countL <= leafMax, rightLeaf = countR <= leafMax;
	childRefs[2 * i + 0] = leftLeaf ? ~((lo << RB_LEAF_BITS) | (countL - 1)) : split;
	childRefs[2 * i + 1] = rightLeaf ? ~(((split + 1) << RB_LEAF_BITS) | (countR - 1)) : (split + 1);
	if(i != 0 && hi - lo + 1 <= leafMax) {
		// inside a collapsed range: some ancestor holds these triangles as one leaf. Nobody else writes this node's parent
		// link (its parent is dead too, or holds it as a leaf), so the mark below is what tells a download to skip it.
		interiorParent[i] = RB_LBVH_DEAD;
		return;
	}
	// Leaf table (what the fit kernel walks): entry `first triangle` carries the leaf, the other triangles of it are blank.
	if(leftLeaf) {
		leafParent[lo] = (i << 1) | 0;
		leafCode[lo]   = (lo << RB_LEAF_BITS) | (countL - 1);
		for(int k = 1; k < countL; ++k) leafCode[lo + k] = -1;
	} else {
		interiorParent[split] = (i << 1) | 0;
	}
	if(rightLeaf) {
		leafParent[split + 1] = (i << 1) | 1;
		leafCode[split + 1]	  = ((split + 1) << RB_LEAF_BITS) | (countR - 1);
		for(int k = 1; k < countR; ++k) leafCode[split + 1 + k] = -1;
	} else {
		interiorParent[split + 1] = (i << 1) | 1;
	}
	if(i == 0) interiorParent[0] = ~0; // root slot 0
}

// Re-gather the triangles in hierarchy order: records for the traversal (v0, v1-v0, v2-v0 + ids) and vertex ids for refits.
__global__ void rb_k_lbvh_gather(const uint32_t * __restrict__ order, const uint32_t * __restrict__ triVertsIn, const uint32_t * __restrict__ triIdsIn, uint32_t ntris, uint32_t * __restrict__ triVertsOut, uint32_t * __restrict__ triIdsOut, int * __restrict__ leafCode) {
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if(i >= ntris) return;
	if(leafCode) leafCode[i] = int(i << RB_LEAF_BITS); // leaf j holds triangle j
	const uint32_t src = order ? order[i] : i;
	for(int k = 0; k < 3; ++k) triVertsOut[3 * i + k] = triVertsIn[3 * src + k];
	triIdsOut[2 * i + 0] = triIdsIn[2 * src + 0];
	triIdsOut[2 * i + 1] = triIdsIn[2 * src + 1];
}

RB_D void rb_write_child_box(float4 * nodes, int parent, int side, V3 lo, V3 hi) {
	float * f = reinterpret_cast<float *>(nodes + 4 * size_t(parent)) + (side ? 6 : 0);
	f[0] = lo.x; f[1] = lo.y; f[2] = lo.z; f[3] = hi.x; f[4] = hi.y; f[5] = hi.z;
}
RB_D void rb_read_child_box(const float4 * nodes, int node, int side, V3 & lo, V3 & hi) {
	const volatile float * f = reinterpret_cast<const volatile float *>(nodes + 4 * size_t(node)) + (side ? 6 : 0);
	lo = V3{f[0], f[1], f[2]};
	hi = V3{f[3], f[4], f[5]};
}

// Bottom-up fit / refit. One thread per leaf: rebuild its triangle records from the current vertices, take the union of
// their boxes, store it in the parent's child slot and climb; at every internal node the first thread to arrive stops,
// the second one (which therefore sees both child boxes) continues. Roots write into the root table.
__global__ void rb_k_fit(const float * __restrict__ verts, const uint32_t * __restrict__ triVerts, const uint32_t * __restrict__ triIds, const int * __restrict__ leafCode, const int * __restrict__ leafParent, const int * __restrict__ interiorParent, int nleaves, float4 * __restrict__ nodes, float4 * __restrict__ tris, float4 * __restrict__ roots, unsigned int * __restrict__ flags, unsigned int * __restrict__ maxDepth) {
	const int leaf = blockIdx.x * blockDim.x + threadIdx.x;
	if(leaf >= nleaves) return;
	const int code	= leafCode[leaf];
	if(code < 0) return; // blank entry of the leaf table (a triangle inside a collapsed LBVH leaf that is not its first)
	const int first = code >> RB_LEAF_BITS;
	const int count = (code & (RB_LEAF_MAX_TRIS - 1)) + 1;
	V3 lo = v3(0.0f), hi = v3(0.0f);
	for(int k = 0; k < count; ++k) {
		const int t = first + k;
		const uint32_t a = triVerts[3 * t], b = triVerts[3 * t + 1], c = triVerts[3 * t + 2];
		const V3 va = V3{verts[3 * a], verts[3 * a + 1], verts[3 * a + 2]};
		const V3 vb = V3{verts[3 * b], verts[3 * b + 1], verts[3 * b + 2]};
		const V3 vc = V3{verts[3 * c], verts[3 * c + 1], verts[3 * c + 2]};
		const V3 e1 = vb - va, e2 = vc - va;
		tris[3 * size_t(t) + 0] = make_float4(va.x, va.y, va.z, __uint_as_float(triIds[2 * t]));
		tris[3 * size_t(t) + 1] = make_float4(e1.x, e1.y, e1.z, __uint_as_float(triIds[2 * t + 1]));
		tris[3 * size_t(t) + 2] = make_float4(e2.x, e2.y, e2.z, 0.0f);
		const V3 tlo = glm_min3(glm_min3(va, vb), vc), thi = glm_max3(glm_max3(va, vb), vc);
		if(k == 0) { lo = tlo; hi = thi; } else { lo = glm_min3(lo, tlo); hi = glm_max3(hi, thi); } // BoundingBox::merge
	}
	int link = leafParent[leaf];
	unsigned int depth = 0;
	for(;;) {
		if(link < 0) {
			// A root: its box lives in the root table.
			const int slot = ~link;
			float * r = reinterpret_cast<float *>(roots + 2 * size_t(slot));
			r[0] = lo.x; r[1] = lo.y; r[2] = lo.z;
			r[4] = hi.x; r[5] = hi.y; r[6] = hi.z;
			break;
		}
		const int parent = link >> 1, side = link & 1;
		rb_write_child_box(nodes, parent, side, lo, hi);
		__threadfence();
		++depth;
		if(atomicAdd(flags + parent, 1u) == 0u) break; // first to arrive: the sibling will finish this node
		V3 slo, shi;
		rb_read_child_box(nodes, parent, side ^ 1, slo, shi);
		lo	 = glm_min3(lo, slo);
		hi	 = glm_max3(hi, shi);
		link = interiorParent[parent];
	}
	// Depth of the deepest leaf (interior nodes on its path): every leaf climbs at least partially, so a separate walk
	// up the parent links gives the exact value.
	unsigned int d = 0;
	for(int l2 = leafParent[leaf]; l2 >= 0; l2 = interiorParent[l2 >> 1]) ++d;
	atomicMax(maxDepth, d);
	(void)depth;
}

__global__ void rb_k_lbvh_write_refs(const int * __restrict__ childRefs, int ninterior, float4 * __restrict__ nodes) {
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if(i >= ninterior) return;
	nodes[4 * size_t(i) + 3] = make_float4(__int_as_float(childRefs[2 * i]), __int_as_float(childRefs[2 * i + 1]), 0.0f, 0.0f);
}

// ------------------------------------------------------------------------------------------------ host
static unsigned rb_blocks(size_t n, unsigned block) { return unsigned((n + block - 1) / block); }

// ------------------------------------------------------------------------------------------------ root blocks
// dst[b] = union of src[2b] and src[2b+1] (entries laid out like roots: (min, ·) (max, ·)); empty entries — inverted
// boxes of meshes without triangles — are ignored, a block of empty entries stays empty.
__global__ void rb_k_root_level(const float4 * __restrict__ src, int nsrc, float4 * __restrict__ dst, int ndst) {
	const int b = blockIdx.x * blockDim.x + threadIdx.x;
	if(b >= ndst) return;
	float lo[3] = {1.0f, 1.0f, 1.0f}, hi[3] = {-1.0f, -1.0f, -1.0f};
	bool any = false;
	for(int j = 2 * b; j < 2 * b + 2 && j < nsrc; ++j) {
		const float4 a = src[2 * j], c = src[2 * j + 1];
		if(a.x > c.x || a.y > c.y || a.z > c.z) continue;
		const float l[3] = {a.x, a.y, a.z}, h[3] = {c.x, c.y, c.z};
		for(int k = 0; k < 3; ++k) {
			lo[k] = any ? fminf(lo[k], l[k]) : l[k];
			hi[k] = any ? fmaxf(hi[k], h[k]) : h[k];
		}
		any = true;
	}
	dst[2 * b]	   = make_float4(lo[0], lo[1], lo[2], 0.0f);
	dst[2 * b + 1] = make_float4(hi[0], hi[1], hi[2], 0.0f);
}

// Blocks are made of CONSECUTIVE roots (the visiting order is the mesh order and must stay so), which only helps when
// consecutive meshes are also neighbours in space — objects listed row by row, instanced grids. The host decides at
// upload time: pairs of consecutive root boxes must, on average, stay small against the whole scene.
bool rb_root_blocks_worthwhile(const float * roots8, size_t nroots) {
	if(nroots < 16 || std::getenv("RB_NO_ROOT_BLOCKS")) return false; // a handful of roots is cheaper to test one by one
	auto area = [](const float * lo, const float * hi) {
		const float x = hi[0] - lo[0], y = hi[1] - lo[1], z = hi[2] - lo[2];
		return (x < 0.0f || y < 0.0f || z < 0.0f) ? 0.0 : 2.0 * (double(x) * y + double(y) * z + double(z) * x);
	};
	float slo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, shi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
	double pairs = 0.0;
	size_t npairs = 0;
	for(size_t r = 0; r < nroots; ++r) {
		const float * lo = roots8 + 8 * r;
		const float * hi = lo + 4;
		if(lo[0] > hi[0]) continue; // empty mesh
		for(int k = 0; k < 3; ++k) {
			slo[k] = std::min(slo[k], lo[k]);
			shi[k] = std::max(shi[k], hi[k]);
		}
		if((r & 1) && roots8[8 * (r - 1)] <= roots8[8 * (r - 1) + 4]) {
			const float * plo = roots8 + 8 * (r - 1);
			float ulo[3], uhi[3];
			for(int k = 0; k < 3; ++k) {
				ulo[k] = std::min(lo[k], plo[k]);
				uhi[k] = std::max(hi[k], plo[4 + k]);
			}
			pairs += area(ulo, uhi);
			++npairs;
		}
	}
	const double scene = area(slo, shi);
	return npairs > 0 && scene > 0.0 && (pairs / double(npairs)) < 0.02 * scene;
}

// ------------------------------------------------------------------------------------------------ root tree
// The reference box-tests EVERY root before it pops the first (Raycaster.cpp:149-155): 52 tests per ray in its largest
// bundled scene, more than the nodes the ray then visits inside the meshes. Which roots a ray enters is a set, whatever the
// order the tests are made in, and the slab test is monotone in the box (rb_next_root): a union box that the ray misses
// hides no root it would enter. So scenes of 9..64 meshes get a small binary tree over their root boxes — topology from
// the boxes at upload (surface-area sweep over the box centres), boxes re-computed on the device after every fit — and the path
// tracer's wavefronts walk it once per ray (rb_roots_tree_pretest) for the bit mask of entered roots, then visit those
// last mesh first as before: the same hits, bit for bit (irregular rays test every root with the NaN-faithful form).
// Up to 64 meshes the mask is two registers; up to 512 it lives in the first entries of the ray's stack column.
static int rb_root_tree_split(std::vector<int> & ids, int begin, int end, const float * roots8, std::vector<int> & refs, int level, int & depth) {
	// returns the reference of the subtree over ids[begin, end): ~root for one root, else the index of a new node
	if(end - begin == 1) return ~ids[begin];
	depth = std::max(depth, level + 1);
	auto centre = [&](int r, int k) { return 0.5f * (roots8[8 * r + k] + roots8[8 * r + 4 + k]); };
	auto area = [](const float * lo, const float * hi) {
		const float dx = std::max(hi[0] - lo[0], 0.0f), dy = std::max(hi[1] - lo[1], 0.0f), dz = std::max(hi[2] - lo[2], 0.0f);
		return dx * dy + dy * dz + dz * dx;
	};
	// Sweep of the surface-area cost along each axis over the roots sorted by box centre (at most 512 of them, once per
	// upload): the split that leaves the two smallest boxes for the most roots. RB_ROOT_TREE_SPLIT=median: the middle one.
	static const bool median = [] { const char * e = std::getenv("RB_ROOT_TREE_SPLIT"); return e && std::string(e) == "median"; }();
	const int n = end - begin;
	int bestAxis = 0, bestAt = n / 2;
	float bestCost = FLT_MAX;
	std::vector<int> order(n), bestOrder;
	std::vector<float> rightArea(n);
	for(int axis = 0; axis < 3; ++axis) {
		for(int i = 0; i < n; ++i) order[i] = ids[begin + i];
		std::sort(order.begin(), order.end(), [&](int a, int b) {
			const float ca = centre(a, axis), cb = centre(b, axis);
			return ca < cb || (ca == cb && a < b);
		});
		float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
		for(int i = n - 1; i >= 1; --i) {
			for(int k = 0; k < 3; ++k) {
				lo[k] = std::min(lo[k], roots8[8 * order[i] + k]);
				hi[k] = std::max(hi[k], roots8[8 * order[i] + 4 + k]);
			}
			rightArea[i] = area(lo, hi);
		}
		float llo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, lhi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
		for(int i = 1; i < n; ++i) { // left = order[0, i), right = order[i, n)
			for(int k = 0; k < 3; ++k) {
				llo[k] = std::min(llo[k], roots8[8 * order[i - 1] + k]);
				lhi[k] = std::max(lhi[k], roots8[8 * order[i - 1] + 4 + k]);
			}
			// (no side below a quarter of the roots: the depth of the tree — stack entries of every ray — stays logarithmic)
			if(i < std::max(1, n / 4) || n - i < std::max(1, n / 4)) continue;
			const float cost = median ? (i == n / 2 ? float(axis == 0 ? 0 : 1) + 1.0f : FLT_MAX) : area(llo, lhi) * float(i) + rightArea[i] * float(n - i);
			if(cost < bestCost) {
				bestCost = cost; bestAxis = axis; bestAt = i;
				bestOrder = order;
			}
		}
	}
	(void)bestAxis;
	// (boxes with NaNs or infinities — a mesh with broken vertices — make every cost unordered: the roots stay as they are
	// and are cut in the middle; any topology gives the right hits)
	if(bestOrder.empty()) bestAt = n / 2;
	else for(int i = 0; i < n; ++i) ids[begin + i] = bestOrder[i];
	const int mid = begin + bestAt;
	const int node = int(refs.size() / 2);
	refs.push_back(0);
	refs.push_back(0); // preorder: a parent's index is smaller than its children's
	const int l = rb_root_tree_split(ids, begin, mid, roots8, refs, level + 1, depth);
	const int r = rb_root_tree_split(ids, mid, end, roots8, refs, level + 1, depth);
	refs[2 * node]	   = l;
	refs[2 * node + 1] = r;
	return node;
}

rb_status rb_root_tree_topology(rb_raycaster * rc, const float * roots8, size_t nroots) {
	cudaFree(rc->d_rootTree);
	rc->d_rootTree	  = nullptr;
	rc->rootTreeNodes = 0;
	rc->rootTreeDepth = 0;
	rc->dev.rootWords = 0;
	const char * e	  = std::getenv("RB_ROOT_TREE_MIN"); // (A/B aid: smallest mesh count that gets a tree; below it the mask pretest tests every root)
	const size_t least = e ? size_t(std::max(2, std::atoi(e))) : size_t(RB_ROOT_TREE_MIN);
	if(nroots < least || nroots > RB_ROOT_TREE_BIG_MAX || std::getenv("RB_NO_ROOT_TREE")) return RB_OK;
	std::vector<int> ids(nroots), refs;
	for(size_t r = 0; r < nroots; ++r) ids[r] = int(r);
	int depth = 0;
	rb_root_tree_split(ids, 0, int(nroots), roots8, refs, 0, depth);
	rc->rootTreeDepth = depth; // interior levels: the walk keeps at most one pending node per level on the ray's stack
	rc->dev.rootWords = nroots > RB_ROOT_TREE_MAX ? int((nroots + 31) / 32) : 0;
	const int nnodes = int(refs.size() / 2);
	std::vector<float> nodes(size_t(nnodes) * 16, 0.0f);
	for(int i = 0; i < nnodes; ++i) {
		std::memcpy(&nodes[size_t(i) * 16 + 12], &refs[2 * i], 4);
		std::memcpy(&nodes[size_t(i) * 16 + 13], &refs[2 * i + 1], 4);
	}
	RB_CUDA(cudaMalloc(&rc->d_rootTree, nodes.size() * sizeof(float) + 64));
	RB_CUDA(cudaMemcpy(rc->d_rootTree, nodes.data(), nodes.size() * sizeof(float), cudaMemcpyHostToDevice));
	rc->rootTreeNodes = nnodes;
	return RB_OK;
}

// Child boxes of the root tree from the current root boxes: 63 nodes at most, children behind their parent — one thread
// walks them backwards (a few microseconds, after a fit that took a millisecond).
__global__ void rb_k_root_tree_fit(const float4 * __restrict__ roots, float4 * __restrict__ tree, int nnodes) {
	if(blockIdx.x != 0 || threadIdx.x != 0) return;
	for(int i = nnodes - 1; i >= 0; --i) {
		float * f = reinterpret_cast<float *>(tree + 4 * size_t(i));
		for(int side = 0; side < 2; ++side) {
			const int ref = __float_as_int(f[12 + side]);
			float lo[3], hi[3];
			if(ref < 0) {
				const float4 a = roots[2 * (~ref)], b = roots[2 * (~ref) + 1];
				lo[0] = a.x; lo[1] = a.y; lo[2] = a.z; hi[0] = b.x; hi[1] = b.y; hi[2] = b.z;
			} else {
				const float * c = reinterpret_cast<const float *>(tree + 4 * size_t(ref));
				for(int k = 0; k < 3; ++k) {
					lo[k] = fminf(c[k], c[6 + k]);
					hi[k] = fmaxf(c[3 + k], c[9 + k]);
				}
			}
			for(int k = 0; k < 3; ++k) {
				f[6 * side + k]		= lo[k];
				f[6 * side + 3 + k] = hi[k];
			}
		}
	}
}

rb_status rb_build_root_levels(rb_raycaster * rc, cudaStream_t s) {
	const int n = rc->dev.nroots;
	rc->dev.rootTree = nullptr;
	if(rc->d_rootTree && rc->rootTreeNodes == n - 1) {
		rb_k_root_tree_fit<<<1, 32, 0, s>>>(rc->d_roots, rc->d_rootTree, rc->rootTreeNodes);
		RB_CUDA(cudaGetLastError());
		rc->dev.rootTree = rc->d_rootTree;
	}
	rc->dev.rootLevels = 0;
	rc->dev.rootBlocks = nullptr;
	if(!rc->rootBlocksWanted || n < 16) return RB_OK;
	int levels = 0;
	while((2 << levels) <= n && levels < RB_MAX_ROOT_LEVELS) ++levels; // 2^levels <= n
	int offset[RB_MAX_ROOT_LEVELS + 2] = {0};
	int total = 0;
	for(int k = 1; k <= levels; ++k) {
		offset[k] = total;
		total += (n + (1 << k) - 1) >> k;
	}
	if(!rc->d_rootBlocks || rc->rootBlocksCapacity < size_t(total)) { // a refit of the same roots reuses the table
		if(rc->d_rootBlocks) RB_CUDA(cudaFree(rc->d_rootBlocks));
		rc->d_rootBlocks	   = nullptr;
		rc->rootBlocksCapacity = 0;
		RB_CUDA(cudaMalloc(&rc->d_rootBlocks, size_t(total) * 2 * sizeof(float4) + 64));
		rc->rootBlocksCapacity = size_t(total);
	}
	const float4 * src = rc->d_roots;
	int nsrc = n;
	for(int k = 1; k <= levels; ++k) {
		const int ndst = (n + (1 << k) - 1) >> k;
		float4 * dst   = rc->d_rootBlocks + 2 * size_t(offset[k]);
		rb_k_root_level<<<rb_blocks(uint32_t(ndst), 128), 128, 0, s>>>(src, nsrc, dst, ndst);
		src	 = dst;
		nsrc = ndst;
	}
	RB_CUDA(cudaGetLastError());
	for(int k = 0; k <= RB_MAX_ROOT_LEVELS; ++k) rc->dev.rootLevelOffset[k] = (k >= 1 && k <= levels) ? offset[k] : 0;
	rc->dev.rootBlocks = rc->d_rootBlocks;
	rc->dev.rootLevels = levels;
	return RB_OK;
}

static rb_status rb_run_fit(rb_raycaster * rc, cudaStream_t s) {
	const int nleaves = int(rc->nLeaves);
	// The spare entry behind the arrival counters receives the depth: a per-frame refit allocates nothing (a 4-byte
	// cudaMallocAsync per call made the driver map and unmap pool memory at every synchronisation: 8-14 ms per refit).
	unsigned int * d_depth = rc->d_flags + rc->nInterior;
	RB_CUDA(cudaMemsetAsync(rc->d_flags, 0, (rc->nInterior + 1) * sizeof(unsigned int), s));
	if(nleaves) {
		rb_k_fit<<<rb_blocks(nleaves, 128), 128, 0, s>>>(rc->d_verts, rc->d_triVerts, rc->d_triIds, rc->d_leafCode, rc->d_leafParent, rc->d_parent, nleaves, rc->d_nodes, rc->d_tris, rc->d_roots, rc->d_flags, d_depth);
		rc->launches++;
	}
	unsigned int depth = 0;
	cudaError_t e = cudaMemcpyAsync(&depth, d_depth, sizeof(depth), cudaMemcpyDeviceToHost, s);
	if(e == cudaSuccess) e = cudaStreamSynchronize(s);
	if(e != cudaSuccess) return rb_fail(RB_ERR_CUDA, "hierarchy fit failed: %s", cudaGetErrorString(e));
	rc->maxDepth	   = int(depth);
	rc->dev.stackDepth = rb_stack_entries(rc);
	return rb_build_root_levels(rc, s); // root boxes may have moved
}

rb_status rb_lbvh_build(rb_raycaster * rc) {
	RB_REQUIRE_DEVICE();
	const size_t ntris = rc->tris.size();
	if(ntris == 0) return rb_fail(RB_ERR_INVALID, "rb_raycaster_build: no triangle to build a hierarchy on");
	if(ntris >= (size_t(1) << (31 - RB_LEAF_BITS))) return rb_fail(RB_ERR_UNSUPPORTED, "too many triangles for 32-bit leaf references");
	const bool timing = std::getenv("RB_BUILD_TIMING") != nullptr;
	auto now = [] { return std::chrono::steady_clock::now(); };
	auto lap = [&](const char * what, std::chrono::steady_clock::time_point & t0) {
		if(timing) {
			cudaDeviceSynchronize();
			const auto t1 = now();
			std::fprintf(stderr, "[rb_lbvh_build] %-22s %8.2f ms\n", what, std::chrono::duration<double, std::milli>(t1 - t0).count());
			t0 = now();
		}
	};
	auto t0 = now();
	rb_raycaster_free_device(rc);
	lap("free previous", t0);
	RB_CUDA(cudaGetDevice(&rc->device));
	cudaStream_t s = cudaStreamPerThread;
	const uint32_t n = uint32_t(ntris);
	// Input triangles in addMesh order: vertex ids and (localId, meshId).
	// Uninitialised staging (a value-initialised vector would first zero 200 MB on one thread).
	std::unique_ptr<uint32_t[]> triVerts(new uint32_t[ntris * 3]), triIds(new uint32_t[ntris * 2]);
	rb_parallel_for(ntris, size_t(1) << 16, [&](size_t begin, size_t end) {
		for(size_t i = begin; i < end; ++i) {
			const rb_tri & t = rc->tris[i];
			triVerts[3 * i] = uint32_t(t.v0); triVerts[3 * i + 1] = uint32_t(t.v1); triVerts[3 * i + 2] = uint32_t(t.v2);
			triIds[2 * i] = uint32_t(t.localId); triIds[2 * i + 1] = t.meshId;
		}
	});
	lap("pack triangles", t0);
	uint32_t *d_tvIn = nullptr, *d_idIn = nullptr, *d_vals = nullptr;
	uint64_t * d_keys = nullptr;
	unsigned int * d_bounds = nullptr;
	int * d_childRefs = nullptr;
	rc->nInterior = ntris - 1;
	rc->nTris	  = ntris;
	rc->nLeaves	  = ntris;
	rc->nVerts	  = rc->verts.size() / 3;
	RB_CUDA(cudaMalloc(&rc->d_verts, rc->verts.size() * sizeof(float) + 64));
	RB_CUDA(cudaMemcpyAsync(rc->d_verts, rc->verts.data(), rc->verts.size() * sizeof(float), cudaMemcpyHostToDevice, s));
	RB_CUDA(cudaMalloc(&d_tvIn, ntris * 3 * sizeof(uint32_t)));
	RB_CUDA(cudaMalloc(&d_idIn, ntris * 2 * sizeof(uint32_t)));
	RB_CUDA(cudaMemcpyAsync(d_tvIn, triVerts.get(), ntris * 3 * sizeof(uint32_t), cudaMemcpyHostToDevice, s));
	RB_CUDA(cudaMemcpyAsync(d_idIn, triIds.get(), ntris * 2 * sizeof(uint32_t), cudaMemcpyHostToDevice, s));
	RB_CUDA(cudaMalloc(&d_keys, ntris * sizeof(uint64_t)));
	RB_CUDA(cudaMalloc(&d_vals, ntris * sizeof(uint32_t)));
	RB_CUDA(cudaMalloc(&d_bounds, 6 * sizeof(unsigned int)));
	RB_CUDA(cudaMalloc(&d_childRefs, (2 * rc->nInterior + 2) * sizeof(int)));
	RB_CUDA(cudaMalloc(&rc->d_nodes, (rc->nInterior + 1) * 64));
	RB_CUDA(cudaMalloc(&rc->d_tris, ntris * 48 + 64));
	RB_CUDA(cudaMalloc(&rc->d_roots, 64));
	RB_CUDA(cudaMalloc(&rc->d_parent, (rc->nInterior + 1) * sizeof(int)));
	RB_CUDA(cudaMalloc(&rc->d_leafParent, ntris * sizeof(int)));
	RB_CUDA(cudaMalloc(&rc->d_leafCode, ntris * sizeof(int)));
	RB_CUDA(cudaMalloc(&rc->d_triVerts, ntris * 3 * sizeof(uint32_t)));
	RB_CUDA(cudaMalloc(&rc->d_triIds, ntris * 2 * sizeof(uint32_t)));
	RB_CUDA(cudaMalloc(&rc->d_flags, (rc->nInterior + 1) * sizeof(unsigned int)));
	RB_CUDA(cudaMalloc(&rc->d_cursors, RB_CURSOR_RING * sizeof(unsigned int)));
	RB_CUDA(cudaMemsetAsync(rc->d_nodes, 0, (rc->nInterior + 1) * 64, s));
	const unsigned int initBounds[6] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0u, 0u, 0u};
	RB_CUDA(cudaMemcpyAsync(d_bounds, initBounds, sizeof(initBounds), cudaMemcpyHostToDevice, s));
	lap("allocate + upload", t0);
	uint64_t launches = 0;
	cudaEvent_t ev0, ev1;
	cudaEventCreate(&ev0);
	cudaEventCreate(&ev1);
	cudaEventRecord(ev0, s);
	rb_k_lbvh_bounds<<<std::min<unsigned>(rb_blocks(n, 256), 148u * 8u), 256, 0, s>>>(rc->d_verts, d_tvIn, n, d_bounds);
	static const int cubic = [] { const char * e = getenv("RB_LBVH_QUANT"); return (e && std::string(e) == "axis") ? 0 : 1; }();
	rb_k_lbvh_morton<<<rb_blocks(n, 256), 256, 0, s>>>(rc->d_verts, d_tvIn, n, d_bounds, d_keys, d_vals, cubic);
	launches += 2;
	rb_status st = rb_radix_sort(d_keys, d_vals, n, 63, s, &launches);
	unsigned int liveLeaves = (unsigned int)n; // one triangle per leaf unless the hierarchy kernel collapses ranges
	if(st == RB_OK) {
		rb_k_lbvh_gather<<<rb_blocks(n, 256), 256, 0, s>>>(d_vals, d_tvIn, d_idIn, n, rc->d_triVerts, rc->d_triIds, rc->d_leafCode);
		++launches;
		if(n > 1) {
			// Triangles per leaf: RB_LBVH_LEAF=1..8 overrides the default of 2 (measured on C5, profiles/r02_lbvh_quality.json:
			// closest-hit 1.64 / 1.79 / 1.76 / 1.79 / 1.68 / 1.59 Grays/s and 480 / 495 / 490 / 471 / 443 / 418 spp/s at 1 / 2 / 3 / 4 / 6 / 8).
			static const int leafMax = [] { const char * e = getenv("RB_LBVH_LEAF"); const int v = e ? atoi(e) : 0; return (v >= 1 && v <= RB_LEAF_MAX_TRIS) ? v : 2; }();
			rb_k_lbvh_hierarchy<<<rb_blocks(n - 1, 128), 128, 0, s>>>(d_keys, int(n), leafMax, d_childRefs, rc->d_parent, rc->d_leafParent, rc->d_leafCode);
			rb_k_lbvh_write_refs<<<rb_blocks(n - 1, 256), 256, 0, s>>>(d_childRefs, int(n - 1), rc->d_nodes);
			launches += 2;
			if(leafMax > 1) {
				// (the bounds words are free again once the codes exist)
				RB_CUDA(cudaMemsetAsync(d_bounds, 0, sizeof(unsigned int), s));
				rb_k_lbvh_count_leaves<<<std::min<unsigned>(rb_blocks(n, 256), 148u * 8u), 256, 0, s>>>(rc->d_leafCode, int(n), d_bounds);
				RB_CUDA(cudaMemcpyAsync(&liveLeaves, d_bounds, sizeof(unsigned int), cudaMemcpyDeviceToHost, s));
				++launches;
			}
		}
		// Leaf table (leaf j holds triangle j) was written by the gather kernel. A single triangle is a root that is
		// itself a leaf.
		float rootRec[8] = {0, 0, 0, 0, 0, 0, 0, 0};
		int rootRef = (n > 1) ? 0 : ~0;
		std::memcpy(&rootRec[3], &rootRef, 4);
		RB_CUDA(cudaMemcpyAsync(rc->d_roots, rootRec, sizeof(rootRec), cudaMemcpyHostToDevice, s));
		if(n == 1) {
			const int rootLink = ~0;
			RB_CUDA(cudaMemcpyAsync(rc->d_leafParent, &rootLink, sizeof(int), cudaMemcpyHostToDevice, s));
		}
		rc->dev.nodes  = rc->d_nodes;
		rc->dev.tris   = rc->d_tris;
		rc->dev.roots  = rc->d_roots;
		rc->dev.nroots = 1;
		st = rb_run_fit(rc, s);
		++launches;
	}
	cudaEventRecord(ev1, s);
	cudaEventSynchronize(ev1);
	float buildMs = 0.0f;
	cudaEventElapsedTime(&buildMs, ev0, ev1);
	rc->buildDeviceMs = buildMs;
	rc->lbvhLeaves	  = liveLeaves; // (read back before the event above completed)
	cudaEventDestroy(ev0);
	cudaEventDestroy(ev1);
	lap("build kernels", t0);
	cudaFree(d_tvIn);
	cudaFree(d_idIn);
	cudaFree(d_keys);
	cudaFree(d_vals);
	cudaFree(d_bounds);
	cudaFree(d_childRefs);
	lap("free scratch", t0);
	if(st) return st;
	if(size_t(rc->dev.stackDepth) * RB_TRACE_BLOCK_THREADS * sizeof(int) > 160 * 1024) {
		return rb_fail(RB_ERR_UNSUPPORTED, "LBVH depth %d needs a traversal stack larger than shared memory", rc->maxDepth);
	}
	rc->launches += launches;
	rc->built			= true;
	rc->buildSerial++;
	rc->hostMirrorStale = true;
	return RB_OK;
}

rb_status rb_lbvh_refit(rb_raycaster * rc, const float * worldPositions, size_t nverts) {
	RB_REQUIRE_DEVICE();
	// worldPositions == NULL: fit the vertices the handle already holds (after rb_raycaster_update_mesh).
	if(worldPositions && nverts != rc->nVerts) return rb_fail(RB_ERR_INVALID, "rb_raycaster_refit: %zu vertices given, hierarchy has %zu", nverts, rc->nVerts);
	RB_CUDA(cudaSetDevice(rc->device));
	cudaStream_t s = cudaStreamPerThread;
	if(worldPositions) {
		RB_CUDA(cudaMemcpyAsync(rc->d_verts, worldPositions, nverts * 3 * sizeof(float), cudaMemcpyHostToDevice, s));
		rc->verts.assign(worldPositions, worldPositions + nverts * 3);
		rc->triBoxesStale = true;
	}
	cudaEvent_t ev0, ev1;
	cudaEventCreate(&ev0);
	cudaEventCreate(&ev1);
	cudaEventRecord(ev0, s);
	rb_status st = rb_run_fit(rc, s);
	cudaEventRecord(ev1, s);
	cudaEventSynchronize(ev1);
	float ms = 0.0f;
	cudaEventElapsedTime(&ms, ev0, ev1);
	rc->refitDeviceMs = ms;
	cudaEventDestroy(ev0);
	cudaEventDestroy(ev1);
	if(st) return st;
	rc->hostMirrorStale = true;
	return RB_OK;
}

// Device tree -> the reference's Node / TriangleInfos arrays (what RaycasterVisualisation reads). Roots come first
// (slots 0..nroots-1), then the children of every device interior node; a leaf keeps (first triangle, count).
rb_status rb_lbvh_download(const rb_raycaster * rc, rb_node * nodesOut, rb_tri * trisOut, float * vertsOut) {
	RB_REQUIRE_DEVICE();
	RB_CUDA(cudaSetDevice(rc->device));
	RB_CUDA(cudaDeviceSynchronize());
	const size_t ni = rc->nInterior, nt = rc->nTris, nr = size_t(rc->dev.nroots);
	// Uninitialised staging: value-initialised vectors would zero more than a gigabyte at 10 M triangles first.
	std::unique_ptr<float[]> nodes(new float[ni * 16 + 16]), roots(new float[nr * 8 + 8]), verts(new float[rc->nVerts * 3 + 3]);
	std::unique_ptr<uint32_t[]> triVerts(new uint32_t[nt * 3 + 3]), triIds(new uint32_t[nt * 2 + 2]);
	if(ni) RB_CUDA(cudaMemcpy(nodes.get(), rc->d_nodes, ni * 64, cudaMemcpyDeviceToHost));
	if(nr) RB_CUDA(cudaMemcpy(roots.get(), rc->d_roots, nr * 32, cudaMemcpyDeviceToHost));
	if(nt) RB_CUDA(cudaMemcpy(triVerts.get(), rc->d_triVerts, nt * 3 * sizeof(uint32_t), cudaMemcpyDeviceToHost));
	if(nt) RB_CUDA(cudaMemcpy(triIds.get(), rc->d_triIds, nt * 2 * sizeof(uint32_t), cudaMemcpyDeviceToHost));
	if(rc->nVerts) RB_CUDA(cudaMemcpy(verts.get(), rc->d_verts, rc->nVerts * 3 * sizeof(float), cudaMemcpyDeviceToHost));
	if(vertsOut && rc->nVerts) std::memcpy(vertsOut, verts.get(), rc->nVerts * 3 * sizeof(float));
	if(trisOut) {
		rb_parallel_for(nt, size_t(1) << 16, [&](size_t begin, size_t end) {
			for(size_t i = begin; i < end; ++i) {
				rb_tri & t = trisOut[i];
				std::memset(&t, 0, sizeof(t));
				t.v0 = triVerts[3 * i]; t.v1 = triVerts[3 * i + 1]; t.v2 = triVerts[3 * i + 2];
				t.localId = triIds[2 * i];
				t.meshId  = triIds[2 * i + 1];
				const float *a = &verts[3 * t.v0], *b = &verts[3 * t.v1], *c = &verts[3 * t.v2];
				for(int k = 0; k < 3; ++k) {
					const float mn = (b[k] < a[k]) ? b[k] : a[k], mx = (a[k] < b[k]) ? b[k] : a[k];
					t.bmin[k] = (c[k] < mn) ? c[k] : mn;
					t.bmax[k] = (mx < c[k]) ? c[k] : mx;
				}
			}
		});
	}
	if(nodesOut) {
		// Numbering: roots in slots [0, nr); the two children of the k-th LIVE device interior node in slots nr + 2k and
		// nr + 2k + 1 (nodes inside a collapsed leaf range are dead and get no record: the output is a proper binary tree).
		// Every record then depends on one device node only, so the conversion runs on all host threads (the reference's
		// consumers follow left/right indices and never assume an order).
		std::unique_ptr<int[]> parent(new int[ni + 1]);
		std::unique_ptr<uint32_t[]> rank(new uint32_t[ni + 1]);
		if(ni) RB_CUDA(cudaMemcpy(parent.get(), rc->d_parent, ni * sizeof(int), cudaMemcpyDeviceToHost));
		uint32_t live = 0;
		for(size_t i = 0; i < ni; ++i) {
			rank[i] = live;
			if(parent[i] != RB_LBVH_DEAD) ++live;
		}
		auto fill = [&](rb_node & nd, int ref, const float * lo, const float * hi) {
			std::memset(&nd, 0, sizeof(nd));
			std::memcpy(nd.bmin, lo, 12);
			std::memcpy(nd.bmax, hi, 12);
			if(ref < 0 || ref == RB_REF_NONE) {
				nd.leaf = 1;
				if(ref != RB_REF_NONE) {
					const int code = ~ref;
					nd.left	 = uint64_t(code >> RB_LEAF_BITS);
					nd.right = uint64_t((code & (RB_LEAF_MAX_TRIS - 1)) + 1);
				}
			} else {
				nd.leaf	 = 0;
				nd.left	 = nr + 2 * size_t(rank[ref]);
				nd.right = nr + 2 * size_t(rank[ref]) + 1;
			}
		};
		for(size_t r = 0; r < nr; ++r) {
			int ref;
			std::memcpy(&ref, &roots[8 * r + 3], 4);
			fill(nodesOut[r], ref, &roots[8 * r], &roots[8 * r + 4]);
		}
		rb_parallel_for(ni, size_t(1) << 15, [&](size_t begin, size_t end) {
			for(size_t i = begin; i < end; ++i) {
				if(parent[i] == RB_LBVH_DEAD) continue;
				const float * f = &nodes[i * 16];
				int refL, refR;
				std::memcpy(&refL, &f[12], 4);
				std::memcpy(&refR, &f[13], 4);
				fill(nodesOut[nr + 2 * size_t(rank[i])], refL, f, f + 3);
				fill(nodesOut[nr + 2 * size_t(rank[i]) + 1], refR, f + 6, f + 9);
			}
		});
	}
	return RB_OK;
}
