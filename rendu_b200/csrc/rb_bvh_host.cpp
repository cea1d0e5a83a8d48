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
#include <cmath>
#include <cstring>

namespace {

struct Pending {
	size_t id, begin, count;
	int depth;
};

inline float minf(float x, float y) { return (y < x) ? y : x; } // glm::min
inline float maxf(float x, float y) { return (x < y) ? y : x; } // glm::max

// What the builder reads and permutes.
struct BuildInput {
	const RbTriVec * tris;
	const float * centroid[3];
	uint32_t * order;
	bool sah = false; // RB_BUILD_SAH: binned surface-area splits instead of the reference's mean splits
};

// Processes ONE set as Raycaster.cpp:73-139 does: box, leaf test, axis, mean partition / median fallback. Fills `node`
// (children left unset) and returns the split position, or SIZE_MAX for a leaf.
size_t split_set(const BuildInput & in, size_t begin, size_t count, rb_node & node) {
	const RbTriVec & tris = *in.tris;
	uint32_t * first = in.order + begin;
	uint32_t * last	 = first + count;
	// Union of the triangle boxes (BoundingBox::merge, Bounds.cpp:17-20).
	float lo[3], hi[3];
	if(count > 0) {
		const rb_tri & t0 = tris[first[0]];
		for(int a = 0; a < 3; ++a) {
			lo[a] = t0.bmin[a];
			hi[a] = t0.bmax[a];
		}
		for(size_t k = 1; k < count; ++k) {
			const rb_tri & t = tris[first[k]];
			for(int a = 0; a < 3; ++a) {
				lo[a] = minf(lo[a], t.bmin[a]);
				hi[a] = maxf(hi[a], t.bmax[a]);
			}
		}
	} else {
		// Empty mesh: the reference would read past its triangle array here; give the root an inverted box that no
		// ray can enter.
		for(int a = 0; a < 3; ++a) {
			lo[a] = 1.0f;
			hi[a] = -1.0f;
		}
	}
	std::memcpy(node.bmin, lo, sizeof(lo));
	std::memcpy(node.bmax, hi, sizeof(hi));
	if(count < 3) {
		node.leaf  = 1;
		node.left  = begin;
		node.right = count;
		return SIZE_MAX;
	}
	node.leaf = 0;
	// Longest axis of the box (ties resolved as in Raycaster.cpp:97-98).
	const float sx = hi[0] - lo[0], sy = hi[1] - lo[1], sz = hi[2] - lo[2];
	const int axis = (sx >= sy && sx >= sz) ? 0 : (sy >= sz ? 1 : 2);
	const float * c = in.centroid[axis];
	size_t split = 0;
	if(count >= 5) {
		float mean = 0.0f;
		for(size_t k = 0; k < count; ++k) mean += c[first[k]];
		mean /= float(count);
		uint32_t * mid = std::partition(first, last, [c, mean](uint32_t t) { return c[t] < mean; });
		split		   = size_t(mid - first);
	}
	if(split == 0 || split == count || count < 5) {
		split = count / 2;
		std::nth_element(first, first + split, last, [c](uint32_t a, uint32_t b) { return c[a] < c[b]; });
	}
	return split;
}

// RB_BUILD_SAH: the same step with the split chosen by the binned surface-area heuristic (16 bins per axis over the
// centroid bounds, cost = area(left) * count(left) + area(right) * count(right)) instead of the reference's mean of the
// longest axis. Same leaf rule (fewer than three triangles), same node format, same framework around it — only the tree
// is a different one, so hits agree with the reference except on exact-distance ties and box-edge grazing, as for
// RB_BUILD_LBVH. Falls back to the median of the longest axis where no plane separates the centroids.
size_t split_set_sah(const BuildInput & in, size_t begin, size_t count, rb_node & node) {
	const RbTriVec & tris = *in.tris;
	uint32_t * first = in.order + begin;
	uint32_t * last	 = first + count;
	float lo[3] = {1.0f, 1.0f, 1.0f}, hi[3] = {-1.0f, -1.0f, -1.0f};
	float clo[3] = {0, 0, 0}, chi[3] = {0, 0, 0};
	for(size_t k = 0; k < count; ++k) {
		const uint32_t t  = first[k];
		const rb_tri & tr = tris[t];
		for(int a = 0; a < 3; ++a) {
			const float c = in.centroid[a][t];
			if(k == 0) {
				lo[a]  = tr.bmin[a];
				hi[a]  = tr.bmax[a];
				clo[a] = chi[a] = c;
			} else {
				lo[a]  = minf(lo[a], tr.bmin[a]);
				hi[a]  = maxf(hi[a], tr.bmax[a]);
				clo[a] = minf(clo[a], c);
				chi[a] = maxf(chi[a], c);
			}
		}
	}
	std::memcpy(node.bmin, lo, sizeof(lo));
	std::memcpy(node.bmax, hi, sizeof(hi));
	if(count < 3) {
		node.leaf  = 1;
		node.left  = begin;
		node.right = count;
		return SIZE_MAX;
	}
	node.leaf = 0;
	const int B = 16;
	struct Bin {
		float lo[3], hi[3];
		size_t n;
	};
	auto halfArea = [](const float * l, const float * h) {
		const float x = h[0] - l[0], y = h[1] - l[1], z = h[2] - l[2];
		return x * y + y * z + z * x;
	};
	double bestCost = 0.0;
	int bestAxis = -1, bestBin = 0;
	float bestScale = 0.0f;
	for(int axis = 0; axis < 3; ++axis) {
		const float extent = chi[axis] - clo[axis];
		if(!(extent > 0.0f)) continue;
		const float scale = float(B) / extent;
		if(!std::isfinite(scale)) continue;
		Bin bins[B];
		for(int b = 0; b < B; ++b) bins[b].n = 0;
		const float * c = in.centroid[axis];
		for(size_t k = 0; k < count; ++k) {
			const uint32_t t = first[k];
			int b			 = int((c[t] - clo[axis]) * scale);
			b				 = b < 0 ? 0 : (b >= B ? B - 1 : b);
			Bin & bin		 = bins[b];
			const rb_tri & tr = tris[t];
			if(bin.n++ == 0) {
				std::memcpy(bin.lo, tr.bmin, 12);
				std::memcpy(bin.hi, tr.bmax, 12);
			} else {
				for(int a = 0; a < 3; ++a) {
					bin.lo[a] = minf(bin.lo[a], tr.bmin[a]);
					bin.hi[a] = maxf(bin.hi[a], tr.bmax[a]);
				}
			}
		}
		// Suffix boxes, then a prefix sweep over the 15 planes.
		float rlo[B][3], rhi[B][3];
		size_t rn[B];
		size_t acc = 0;
		float alo[3] = {0, 0, 0}, ahi[3] = {0, 0, 0};
		for(int b = B - 1; b >= 0; --b) {
			if(bins[b].n) {
				for(int a = 0; a < 3; ++a) {
					alo[a] = acc ? minf(alo[a], bins[b].lo[a]) : bins[b].lo[a];
					ahi[a] = acc ? maxf(ahi[a], bins[b].hi[a]) : bins[b].hi[a];
				}
				acc += bins[b].n;
			}
			std::memcpy(rlo[b], alo, 12);
			std::memcpy(rhi[b], ahi, 12);
			rn[b] = acc;
		}
		acc = 0;
		for(int b = 0; b + 1 < B; ++b) {
			if(bins[b].n) {
				for(int a = 0; a < 3; ++a) {
					alo[a] = acc ? minf(alo[a], bins[b].lo[a]) : bins[b].lo[a];
					ahi[a] = acc ? maxf(ahi[a], bins[b].hi[a]) : bins[b].hi[a];
				}
				acc += bins[b].n;
			}
			if(acc == 0 || rn[b + 1] == 0) continue;
			const double cost = double(halfArea(alo, ahi)) * double(acc) + double(halfArea(rlo[b + 1], rhi[b + 1])) * double(rn[b + 1]);
			if(bestAxis < 0 || cost < bestCost) {
				bestCost  = cost;
				bestAxis  = axis;
				bestBin	  = b;
				bestScale = scale;
			}
		}
	}
	size_t split = 0;
	if(bestAxis >= 0) {
		const float * c	 = in.centroid[bestAxis];
		const float base = clo[bestAxis];
		uint32_t * mid	 = std::partition(first, last, [=](uint32_t t) {
			  int b = int((c[t] - base) * bestScale);
			  b		= b < 0 ? 0 : (b >= B ? B - 1 : b);
			  return b <= bestBin;
		  });
		split = size_t(mid - first);
	}
	if(split == 0 || split == count) {
		const float sx = hi[0] - lo[0], sy = hi[1] - lo[1], sz = hi[2] - lo[2];
		const int axis = (sx >= sy && sx >= sz) ? 0 : (sy >= sz ? 1 : 2);
		const float * c = in.centroid[axis];
		split = count / 2;
		std::nth_element(first, first + split, last, [c](uint32_t a, uint32_t b) { return c[a] < c[b]; });
	}
	return split;
}

// The same step for the few very large sets at the top of the tree, on several threads, with the same result bit for
// bit: min/max unions are associative; the centroid mean stays a sequential fp32 sum (its rounding depends on the order);
// and libstdc++'s std::partition for bidirectional iterators (bits/stl_algo.h, __partition) is a two-pointer sweep whose
// outcome has a closed form — with m elements satisfying the predicate, the k-th element of [0, m) that fails it (from
// the left) is swapped with the k-th element of [m, n) that satisfies it (from the right), everything else stays — so
// the two index lists are built with per-chunk counts and the swaps are applied in parallel.
size_t split_set_parallel(const BuildInput & in, size_t begin, size_t count, rb_node & node) {
	const RbTriVec & tris = *in.tris;
	uint32_t * first = in.order + begin;
	const size_t chunks = 64;
	const size_t span	= (count + chunks - 1) / chunks;
	// Union of the triangle boxes.
	std::vector<float> plo(chunks * 3, 0.0f), phi(chunks * 3, 0.0f);
	std::vector<uint8_t> used(chunks, 0);
	rb_parallel_for(chunks, 1, [&](size_t cb, size_t ce) {
		for(size_t c = cb; c < ce; ++c) {
			const size_t b = c * span, e = std::min(count, b + span);
			if(b >= e) continue;
			float lo[3], hi[3];
			const rb_tri & t0 = tris[first[b]];
			for(int a = 0; a < 3; ++a) {
				lo[a] = t0.bmin[a];
				hi[a] = t0.bmax[a];
			}
			for(size_t k = b + 1; k < e; ++k) {
				const rb_tri & t = tris[first[k]];
				for(int a = 0; a < 3; ++a) {
					lo[a] = minf(lo[a], t.bmin[a]);
					hi[a] = maxf(hi[a], t.bmax[a]);
				}
			}
			for(int a = 0; a < 3; ++a) {
				plo[3 * c + a] = lo[a];
				phi[3 * c + a] = hi[a];
			}
			used[c] = 1;
		}
	});
	float lo[3] = {0, 0, 0}, hi[3] = {0, 0, 0};
	bool any = false;
	for(size_t c = 0; c < chunks; ++c) { // chunk order = element order: minf/maxf pick the same values as the plain loop
		if(!used[c]) continue;
		for(int a = 0; a < 3; ++a) {
			lo[a] = any ? minf(lo[a], plo[3 * c + a]) : plo[3 * c + a];
			hi[a] = any ? maxf(hi[a], phi[3 * c + a]) : phi[3 * c + a];
		}
		any = true;
	}
	std::memcpy(node.bmin, lo, sizeof(lo));
	std::memcpy(node.bmax, hi, sizeof(hi));
	node.leaf = 0;
	const float sx = hi[0] - lo[0], sy = hi[1] - lo[1], sz = hi[2] - lo[2];
	const int axis = (sx >= sy && sx >= sz) ? 0 : (sy >= sz ? 1 : 2);
	const float * c = in.centroid[axis];
	float mean = 0.0f;
	for(size_t k = 0; k < count; ++k) mean += c[first[k]];
	mean /= float(count);
	// Predicate counts per chunk -> m.
	std::vector<size_t> trues(chunks + 1, 0);
	rb_parallel_for(chunks, 1, [&](size_t cb, size_t ce) {
		for(size_t ch = cb; ch < ce; ++ch) {
			const size_t b = ch * span, e = std::min(count, b + span);
			size_t n = 0;
			for(size_t k = b; k < e; ++k) n += (c[first[k]] < mean) ? 1 : 0;
			trues[ch + 1] = n;
		}
	});
	for(size_t ch = 0; ch < chunks; ++ch) trues[ch + 1] += trues[ch];
	const size_t m = trues[chunks];
	size_t split  = m;
	if(m != 0 && m != count) {
		// leftFalse: positions < m failing the predicate, ascending; rightTrue: positions >= m satisfying it, ascending
		// here and consumed from the back. Both have the same length.
		std::vector<size_t> nFalse(chunks + 1, 0), nTrue(chunks + 1, 0);
		rb_parallel_for(chunks, 1, [&](size_t cb, size_t ce) {
			for(size_t ch = cb; ch < ce; ++ch) {
				const size_t b = ch * span, e = std::min(count, b + span);
				size_t f = 0, t = 0;
				for(size_t k = b; k < e; ++k) {
					const bool p = c[first[k]] < mean;
					if(k < m) f += p ? 0 : 1;
					else t += p ? 1 : 0;
				}
				nFalse[ch + 1] = f;
				nTrue[ch + 1]  = t;
			}
		});
		for(size_t ch = 0; ch < chunks; ++ch) {
			nFalse[ch + 1] += nFalse[ch];
			nTrue[ch + 1] += nTrue[ch];
		}
		const size_t swaps = nFalse[chunks];
		std::vector<uint32_t> leftFalse(swaps), rightTrue(swaps);
		rb_parallel_for(chunks, 1, [&](size_t cb, size_t ce) {
			for(size_t ch = cb; ch < ce; ++ch) {
				const size_t b = ch * span, e = std::min(count, b + span);
				size_t f = nFalse[ch], t = nTrue[ch];
				for(size_t k = b; k < e; ++k) {
					const bool p = c[first[k]] < mean;
					if(k < m) {
						if(!p) leftFalse[f++] = uint32_t(k);
					} else if(p) {
						rightTrue[t++] = uint32_t(k);
					}
				}
			}
		});
		rb_parallel_for(swaps, size_t(1) << 14, [&](size_t b, size_t e) {
			for(size_t k = b; k < e; ++k) std::swap(first[leftFalse[k]], first[rightTrue[swaps - 1 - k]]);
		});
	} else {
		// Everything on one side: the reference falls back to the median (Raycaster.cpp:116-123).
		split = count / 2;
		std::nth_element(first, first + split, first + count, [c](uint32_t a, uint32_t b) { return c[a] < c[b]; });
	}
	return split;
}

// The reference's LIFO loop on one sub-tree, with local numbering: nodes[0] is the set itself, its descendants follow
// in the order the reference would emplace them (left child, right child, then the whole right sub-tree before the
// left one). Because the loop is LIFO, that is also the order in which the global build creates them — contiguously.
struct SubTree {
	std::vector<rb_node> nodes;
	int maxDepth = 0;
};

void build_subtree(const BuildInput & in, size_t begin, size_t count, int depth, SubTree & out) {
	out.nodes.assign(1, rb_node{});
	std::vector<Pending> todo;
	todo.reserve(64);
	todo.push_back({0, begin, count, depth});
	while(!todo.empty()) {
		const Pending cur = todo.back();
		todo.pop_back();
		rb_node node{};
		const size_t split = in.sah ? split_set_sah(in, cur.begin, cur.count, node) : split_set(in, cur.begin, cur.count, node);
		if(split != SIZE_MAX) {
			out.maxDepth = std::max(out.maxDepth, cur.depth + 1);
			node.left	 = out.nodes.size();
			node.right	 = out.nodes.size() + 1;
			out.nodes.emplace_back();
			out.nodes.emplace_back();
			todo.push_back({size_t(node.left), cur.begin, split, cur.depth + 1});
			todo.push_back({size_t(node.right), cur.begin + split, cur.count - split, cur.depth + 1});
		}
		out.nodes[cur.id] = node;
	}
}

} // namespace

void rb_host_build_reference(rb_raycaster * rc) {
	const size_t ntris = rc->tris.size();
	// BoundingBox::getCentroid (Bounds.cpp:61-63): 0.5f * (minis + maxis), per axis.
	std::vector<float> centroid[3];
	for(int a = 0; a < 3; ++a) centroid[a].resize(ntris);
	std::vector<uint32_t> order(ntris);
	rb_parallel_for(ntris, size_t(1) << 16, [&](size_t b, size_t e) {
		for(size_t i = b; i < e; ++i) {
			for(int a = 0; a < 3; ++a) centroid[a][i] = 0.5f * (rc->tris[i].bmin[a] + rc->tris[i].bmax[a]);
			order[i] = uint32_t(i);
		}
	});
	BuildInput in{&rc->tris, {centroid[0].data(), centroid[1].data(), centroid[2].data()}, order.data(), rc->buildMode == RB_BUILD_SAH};

	// The upper part of the tree is built here, in the reference's order; a set of at most `taskSize` triangles is not
	// expanded but handed to a worker thread (disjoint triangle ranges, nothing shared). Since the reference's loop is
	// LIFO, all descendants of such a set are created contiguously, right when the set is popped: `sequence` records
	// the creation order of the upper nodes and of the deferred sub-trees, and the final numbering — one root per mesh
	// first (Raycaster.cpp:43-46), then emplace_back order — is assigned once the sub-tree sizes are known.
	const size_t taskSize = ntris >= (size_t(1) << 16) ? std::max<size_t>(ntris / 64, 4096) : 0;
	struct Upper {
		rb_node node;
		long left = -1, right = -1; // upper-node indices of the children
		long task = -1;				// deferred sub-tree rooted here
	};
	struct Task {
		size_t upper, begin, count;
		int depth;
		SubTree tree;
		size_t base = 0; // final id of local node 1
	};
	std::vector<Upper> upper(rc->meshCount);
	std::vector<Task> tasks;
	std::vector<long> sequence; // >= 0: upper node created; < 0: ~task index (its descendants are created here)
	std::vector<Pending> todo;
	todo.reserve(64);
	// A rebuild starts again from the per-mesh triangle ranges, which stay valid because the build only permutes
	// triangles inside their mesh's range.
	for(size_t m = 0; m < rc->meshCount; ++m) todo.push_back({m, size_t(rc->meshFirst[m]), size_t(rc->meshTris[m]), 0});
	int maxDepth = 0;
	while(!todo.empty()) {
		const Pending cur = todo.back();
		todo.pop_back();
		if(taskSize && cur.count <= taskSize && cur.count >= 3) {
			upper[cur.id].task = long(tasks.size());
			sequence.push_back(~long(tasks.size()));
			tasks.push_back(Task{cur.id, cur.begin, cur.count, cur.depth, SubTree{}, 0});
			continue;
		}
		rb_node node{};
		const size_t split = in.sah ? split_set_sah(in, cur.begin, cur.count, node) : (cur.count >= (size_t(1) << 17)) ? split_set_parallel(in, cur.begin, cur.count, node) : split_set(in, cur.begin, cur.count, node);
		if(split != SIZE_MAX) {
			maxDepth = std::max(maxDepth, cur.depth + 1);
			const long l = long(upper.size()), r = l + 1;
			upper.emplace_back();
			upper.emplace_back();
			sequence.push_back(l);
			sequence.push_back(r);
			upper[cur.id].left	= l;
			upper[cur.id].right = r;
			todo.push_back({size_t(l), cur.begin, split, cur.depth + 1});
			todo.push_back({size_t(r), cur.begin + split, cur.count - split, cur.depth + 1});
		}
		upper[cur.id].node = node;
	}
	rb_parallel_for(tasks.size(), 1, [&](size_t b, size_t e) {
		for(size_t k = b; k < e; ++k) build_subtree(in, tasks[k].begin, tasks[k].count, tasks[k].depth, tasks[k].tree);
	});

	// Final numbering.
	std::vector<size_t> finalId(upper.size());
	for(size_t m = 0; m < rc->meshCount; ++m) finalId[m] = m;
	size_t next = rc->meshCount;
	for(const long ev : sequence) {
		if(ev >= 0) {
			finalId[size_t(ev)] = next++;
		} else {
			Task & t = tasks[size_t(~ev)];
			t.base	 = next;
			next += t.tree.nodes.size() - 1;
		}
	}
	rc->nodes.assign(next, rb_node{});
	for(size_t u = 0; u < upper.size(); ++u) {
		if(upper[u].task >= 0) continue; // written with its sub-tree below
		rb_node node = upper[u].node;
		if(!node.leaf) {
			node.left  = finalId[size_t(upper[u].left)];
			node.right = finalId[size_t(upper[u].right)];
		}
		rc->nodes[finalId[u]] = node;
	}
	rb_parallel_for(tasks.size(), 1, [&](size_t b, size_t e) {
		for(size_t k = b; k < e; ++k) {
			const Task & t = tasks[k];
			for(size_t j = 0; j < t.tree.nodes.size(); ++j) {
				rb_node node = t.tree.nodes[j];
				if(!node.leaf) {
					node.left  = t.base + size_t(node.left) - 1;
					node.right = t.base + size_t(node.right) - 1;
				}
				rc->nodes[j == 0 ? finalId[t.upper] : t.base + j - 1] = node;
			}
		}
	});
	for(const Task & t : tasks) maxDepth = std::max(maxDepth, t.tree.maxDepth);

	// Apply the permutation to the triangle records.
	RbTriVec sorted(ntris);
	rb_parallel_for(ntris, size_t(1) << 16, [&](size_t b, size_t e) {
		for(size_t i = b; i < e; ++i) sorted[i] = rc->tris[order[i]];
	});
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
	out.dummyTris = 0;
	if(nt + 1 >= (size_t(1) << (31 - RB_LEAF_BITS))) {
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
	// Every interior record is dereferenced below, reachable from a root or not (an uploaded hierarchy is caller data).
	for(size_t i = 0; i < nn; ++i) {
		const rb_node & n = rc->nodes[i];
		if(n.leaf) {
			if(n.right > RB_LEAF_MAX_TRIS || n.left + n.right > nt) {
				if(why) *why = "leaf " + std::to_string(i) + " holds a triangle range outside the triangle array (or more than " + std::to_string(RB_LEAF_MAX_TRIS) + " triangles)";
				return false;
			}
			continue;
		}
		if(n.left >= nn || n.right >= nn) {
			if(why) *why = "interior node " + std::to_string(i) + " has a child index outside the node array";
			return false;
		}
		for(const uint64_t c : {n.left, n.right}) {
			// An empty leaf below an interior node would never report to its parent's arrival counter: a refit would
			// silently stop fitting that subtree's ancestors. The reference never creates one (Raycaster.cpp:88-91).
			if(rc->nodes[c].leaf && rc->nodes[c].right == 0) {
				if(why) *why = "interior node " + std::to_string(i) + " has an empty leaf as a child";
				return false;
			}
		}
	}
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
			// Empty mesh. Its root must not carry RB_REF_NONE: the traversal reads that value as "no roots left" and would
			// never test the meshes added before it (roots are consumed last mesh first), and no box keeps every ray out —
			// the slab test is symmetric in min / max. The root gets a leaf of ONE degenerate triangle instead (all zeros,
			// appended after the real ones): det = 0 fails `abs(det) < FLT_EPSILON`, a NaN ray fails the final range test.
			// The root's box is the point (0, 0, 0): a valid (if empty) box for the refit and wide-tree code.
			d[0] = d[1] = d[2] = 0.0f;
			d[4] = d[5] = d[6] = 0.0f;
			d[3] = asFloat(~int(uint32_t(nt) << RB_LEAF_BITS));
			out.dummyTris = 1;
		}
	}
	out.tris.assign((nt + out.dummyTris) * 12, 0.0f);
	out.triVerts.resize(nt * 3);
	out.triIds.resize(nt * 2);
	const float * V = rc->verts.data();
	const size_t nv = rc->verts.size() / 3;
	std::atomic<bool> bad{false};
	rb_parallel_for(nt, size_t(1) << 16, [&](size_t begin, size_t end) {
		for(size_t i = begin; i < end; ++i) {
			const rb_tri & t = rc->tris[i];
			if(t.v0 >= nv || t.v1 >= nv || t.v2 >= nv) {
				bad = true;
				return;
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
	});
	if(bad) {
		if(why) *why = "triangle vertex index out of range";
		return false;
	}
	return true;
}
