// rb_bvh_host.cpp — host side of RB_BUILD_REFERENCE: a load-time restatement of Raycaster::updateHierarchy
// (reference: src/engine/raycaster/Raycaster.cpp:54-142) that reproduces the reference's node numbering and
// triangle order, plus the conversion of a reference-format hierarchy into the device layout of rb_trace.cuh.
//
// The reference's tree is defined by libstdc++: it reorders `_triangles` with std::partition (mean of box centroids
// along the longest axis, Raycaster.cpp:100-115) and std::nth_element (median fallback, :120-125). Both algorithms
// move elements purely as a function of the predicate / comparator results, so running the SAME std algorithms on a
// permutation of triangle indices, with predicates that read precomputed centroids, yields the same permutation
// without moving 64-byte records around. Sub-trees are processed in the reference's LIFO order (roots pushed 0..M-1,
// left child pushed before right) so that emplace_back numbering matches node for node.
#include "rb_internal.h"

#include <algorithm>
#include <cstring>

namespace {

struct Pending {
	size_t id, begin, count;
	int depth;
};

inline float minf(float x, float y) { return (y < x) ? y : x; } // glm::min
inline float maxf(float x, float y) { return (x < y) ? y : x; } // glm::max

} // namespace

void rb_host_build_reference(rb_raycaster * rc) {
	const size_t ntris = rc->tris.size();
	// One root per mesh (Raycaster.cpp:43-46); a rebuild starts again from the per-mesh triangle ranges, which stay
	// valid because the build only permutes triangles inside their mesh's range.
	rc->nodes.assign(rc->meshCount, rb_node{});

	// BoundingBox::getCentroid (Bounds.cpp:61-63): 0.5f * (minis + maxis), per axis.
	std::vector<float> centroid[3];
	for(int a = 0; a < 3; ++a) {
		centroid[a].resize(ntris);
		for(size_t i = 0; i < ntris; ++i) {
			centroid[a][i] = 0.5f * (rc->tris[i].bmin[a] + rc->tris[i].bmax[a]);
		}
	}
	std::vector<uint32_t> order(ntris);
	for(size_t i = 0; i < ntris; ++i) order[i] = uint32_t(i);

	std::vector<Pending> todo;
	todo.reserve(64);
	for(size_t m = 0; m < rc->meshCount; ++m) {
		todo.push_back({m, size_t(rc->meshFirst[m]), size_t(rc->meshTris[m]), 0});
	}
	int maxDepth = 0;
	const std::vector<rb_tri> & tris = rc->tris;

	while(!todo.empty()) {
		const Pending cur = todo.back();
		todo.pop_back();
		uint32_t * first = order.data() + cur.begin;
		uint32_t * last	 = first + cur.count;

		// Union of the triangle boxes (BoundingBox::merge, Bounds.cpp:17-20).
		float lo[3], hi[3];
		if(cur.count > 0) {
			const rb_tri & t0 = tris[first[0]];
			for(int a = 0; a < 3; ++a) {
				lo[a] = t0.bmin[a];
				hi[a] = t0.bmax[a];
			}
			for(size_t k = 1; k < cur.count; ++k) {
				const rb_tri & t = tris[first[k]];
				for(int a = 0; a < 3; ++a) {
					lo[a] = minf(lo[a], t.bmin[a]);
					hi[a] = maxf(hi[a], t.bmax[a]);
				}
			}
		} else {
			// Empty mesh: the reference would read past its triangle array here; give the root an inverted box
			// that no ray can enter.
			for(int a = 0; a < 3; ++a) {
				lo[a] = 1.0f;
				hi[a] = -1.0f;
			}
		}
		{
			rb_node & node = rc->nodes[cur.id];
			std::memcpy(node.bmin, lo, sizeof(lo));
			std::memcpy(node.bmax, hi, sizeof(hi));
		}
		if(cur.count < 3) {
			rb_node & node = rc->nodes[cur.id];
			node.leaf	   = 1;
			node.left	   = cur.begin;
			node.right	   = cur.count;
			continue;
		}
		maxDepth = std::max(maxDepth, cur.depth + 1);

		// Longest axis of the box (ties resolved as in Raycaster.cpp:97-98).
		const float sx = hi[0] - lo[0], sy = hi[1] - lo[1], sz = hi[2] - lo[2];
		const int axis = (sx >= sy && sx >= sz) ? 0 : (sy >= sz ? 1 : 2);
		const float * c = centroid[axis].data();

		size_t split = 0;
		if(cur.count >= 5) {
			float mean = 0.0f;
			for(size_t k = 0; k < cur.count; ++k) mean += c[first[k]];
			mean /= float(cur.count);
			uint32_t * mid = std::partition(first, last, [c, mean](uint32_t t) { return c[t] < mean; });
			split		   = size_t(mid - first);
		}
		if(split == 0 || split == cur.count || cur.count < 5) {
			split = cur.count / 2;
			std::nth_element(first, first + split, last, [c](uint32_t a, uint32_t b) { return c[a] < c[b]; });
		}

		const size_t leftId = rc->nodes.size();
		rc->nodes.emplace_back();
		const size_t rightId = rc->nodes.size();
		rc->nodes.emplace_back();
		rb_node & node = rc->nodes[cur.id];
		node.leaf	   = 0;
		node.left	   = leftId;
		node.right	   = rightId;
		todo.push_back({leftId, cur.begin, split, cur.depth + 1});
		todo.push_back({rightId, cur.begin + split, cur.count - split, cur.depth + 1});
	}

	// Apply the permutation to the triangle records.
	std::vector<rb_tri> sorted(ntris);
	for(size_t i = 0; i < ntris; ++i) sorted[i] = tris[order[i]];
	rc->tris.swap(sorted);
	rc->maxDepth = maxDepth;
}

int rb_host_tree_depth(const rb_raycaster * rc, std::string * why) {
	const size_t nn = rc->nodes.size();
	if(rc->meshCount > nn) {
		if(why) *why = "mesh_count exceeds node count";
		return -1;
	}
	std::vector<std::pair<size_t, int>> todo;
	std::vector<uint8_t> seen(nn, 0);
	int depth = 0;
	for(size_t m = 0; m < rc->meshCount; ++m) todo.push_back({m, 0});
	while(!todo.empty()) {
		const auto cur = todo.back();
		todo.pop_back();
		if(cur.first >= nn || seen[cur.first]) {
			if(why) *why = "node index out of range or visited twice (not a forest)";
			return -1;
		}
		seen[cur.first]	  = 1;
		const rb_node & n = rc->nodes[cur.first];
		if(n.leaf) {
			if(n.left + n.right > rc->tris.size()) {
				if(why) *why = "leaf triangle range out of bounds";
				return -1;
			}
			if(n.right > RB_LEAF_MAX_TRIS) {
				if(why) *why = "leaf holds more than 8 triangles";
				return -1;
			}
			continue;
		}
		depth = std::max(depth, cur.second + 1);
		todo.push_back({size_t(n.left), cur.second + 1});
		todo.push_back({size_t(n.right), cur.second + 1});
	}
	return depth;
}

bool rb_host_flatten(const rb_raycaster * rc, RbFlatBvh & out, std::string * why) {
	const size_t nn = rc->nodes.size();
	const size_t nt = rc->tris.size();
	if(nt >= (size_t(1) << (31 - RB_LEAF_BITS))) {
		if(why) *why = "too many triangles for 32-bit leaf references";
		return false;
	}
	// Compact index of every interior node, in node order.
	std::vector<int> interiorIndex(nn, -1);
	size_t nInterior = 0;
	for(size_t i = 0; i < nn; ++i) {
		if(!rc->nodes[i].leaf) interiorIndex[i] = int(nInterior++);
	}
	auto childRef = [&](size_t id) -> int {
		const rb_node & n = rc->nodes[id];
		if(!n.leaf) return interiorIndex[id];
		if(n.right == 0) return RB_REF_NONE; // empty leaf: never traversed
		return ~int((uint32_t(n.left) << RB_LEAF_BITS) | uint32_t(n.right - 1));
	};
	auto asFloat = [](int v) {
		float f;
		std::memcpy(&f, &v, 4);
		return f;
	};
	out.nodes.assign(nInterior * 16, 0.0f);
	out.parent.assign(nInterior, -1);
	out.leafParent.clear();
	out.leafCode.clear();
	auto addLeaf = [&](size_t id, int link) {
		const rb_node & n = rc->nodes[id];
		if(n.right == 0) return;
		out.leafParent.push_back(link);
		out.leafCode.push_back(int((uint32_t(n.left) << RB_LEAF_BITS) | uint32_t(n.right - 1)));
	};
	for(size_t i = 0; i < nn; ++i) {
		const rb_node & n = rc->nodes[i];
		if(n.leaf) continue;
		const rb_node & L = rc->nodes[n.left];
		const rb_node & R = rc->nodes[n.right];
		float * d		  = out.nodes.data() + size_t(interiorIndex[i]) * 16;
		d[0] = L.bmin[0]; d[1] = L.bmin[1]; d[2] = L.bmin[2]; d[3] = L.bmax[0];
		d[4] = L.bmax[1]; d[5] = L.bmax[2]; d[6] = R.bmin[0]; d[7] = R.bmin[1];
		d[8] = R.bmin[2]; d[9] = R.bmax[0]; d[10] = R.bmax[1]; d[11] = R.bmax[2];
		d[12] = asFloat(childRef(n.left));
		d[13] = asFloat(childRef(n.right));
		if(!L.leaf) out.parent[interiorIndex[n.left]] = (interiorIndex[i] << 1) | 0; else addLeaf(n.left, (interiorIndex[i] << 1) | 0);
		if(!R.leaf) out.parent[interiorIndex[n.right]] = (interiorIndex[i] << 1) | 1; else addLeaf(n.right, (interiorIndex[i] << 1) | 1);
	}
	out.roots.assign(size_t(rc->meshCount) * 8, 0.0f);
	for(size_t m = 0; m < rc->meshCount; ++m) {
		const rb_node & n = rc->nodes[m];
		float * d		  = out.roots.data() + m * 8;
		d[0] = n.bmin[0]; d[1] = n.bmin[1]; d[2] = n.bmin[2];
		d[3] = asFloat(childRef(m));
		d[4] = n.bmax[0]; d[5] = n.bmax[1]; d[6] = n.bmax[2];
		d[7] = 0.0f;
		if(!n.leaf) out.parent[interiorIndex[m]] = ~int(m); // negative: a root, ~value = root slot
		else addLeaf(m, ~int(m));
		if(n.right == 0 && n.leaf) {
			// Empty mesh: inverted box, never accepted.
			d[0] = d[1] = d[2] = 1.0f;
			d[4] = d[5] = d[6] = -1.0f;
		}
	}
	out.tris.assign(nt * 12, 0.0f);
	out.triVerts.resize(nt * 3);
	out.triIds.resize(nt * 2);
	const float * V = rc->verts.data();
	const size_t nv = rc->verts.size() / 3;
	for(size_t i = 0; i < nt; ++i) {
		const rb_tri & t = rc->tris[i];
		if(t.v0 >= nv || t.v1 >= nv || t.v2 >= nv) {
			if(why) *why = "triangle vertex index out of range";
			return false;
		}
		const float * a = V + 3 * t.v0;
		const float * b = V + 3 * t.v1;
		const float * c = V + 3 * t.v2;
		float * d		= out.tris.data() + i * 12;
		d[0] = a[0]; d[1] = a[1]; d[2] = a[2];
		d[3] = asFloat(int(uint32_t(t.localId)));
		d[4] = b[0] - a[0]; d[5] = b[1] - a[1]; d[6] = b[2] - a[2];
		d[7] = asFloat(int(t.meshId));
		d[8] = c[0] - a[0]; d[9] = c[1] - a[1]; d[10] = c[2] - a[2];
		d[11] = 0.0f;
		out.triVerts[3 * i + 0] = uint32_t(t.v0);
		out.triVerts[3 * i + 1] = uint32_t(t.v1);
		out.triVerts[3 * i + 2] = uint32_t(t.v2);
		out.triIds[2 * i + 0]	= uint32_t(t.localId);
		out.triIds[2 * i + 1]	= t.meshId;
	}
	return true;
}
