// rb_wide.cpp — host collapse of the reference-order binary tree into the eight-wide quantised tree of rb_wide.h.
//
// Which descendants a wide node adopts is decided by the surface-area dynamic programme of Ylitie, Karras and Laine,
// "Efficient Incoherent Ray Traversal on GPUs Through Compressed Wide BVHs" (HPG 2017), section 3: for every binary
// node n and every i in 1..7, C(n, i) is the cheapest way to cover the subtree of n with at most i children of a wide
// node, where one child is either a leaf of up to three triangles or a wide node of its own. Whatever is chosen, the
// children are listed — and later visited — in the order the reference's stack loop reaches them (right before left,
// Raycaster.cpp:160-173), and a leaf child lists its triangles in that order too.
#include "rb_wide.h"
#include "rb_internal.h"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>

namespace {

// Entering a wide node against testing one triangle, in the kernel's own currency (about 200 against 80 instructions,
// five against three L1 wavefronts).
const double kCostNode = 1.0, kCostTri = 0.5;

struct WChild {
	int node;  // >= 0: binary interior node that becomes a wide node; -1: leaf child
	int ntris; // leaf child: its triangles (hierarchy indices) in visiting order
	int tri[RB_WIDE_MAX_LEAF_TRIS];
	float lo[3], hi[3];
};

inline int refOf(float f) {
	int v;
	std::memcpy(&v, &f, 4);
	return v;
}
inline bool isNone(int ref) { return ref == RB_REF_NONE; }
inline bool isLeaf(int ref) { return ref < 0; }
inline int leafFirst(int ref) { return (~ref) >> RB_LEAF_BITS; }
inline int leafCount(int ref) { return ((~ref) & (RB_LEAF_MAX_TRIS - 1)) + 1; }

inline double halfArea(const float * lo, const float * hi) {
	const double x = double(hi[0]) - lo[0], y = double(hi[1]) - lo[1], z = double(hi[2]) - lo[2];
	return x * y + y * z + z * x;
}

inline float roundDown(double x) {
	float f = float(x);
	if(double(f) > x) f = std::nextafterf(f, -INFINITY);
	return f;
}

struct Plan {
	const float * nodes = nullptr; // flattened binary nodes, 16 floats each
	std::vector<uint32_t> prims;   // triangles below each interior node
	std::vector<double> cost;	   // 8 per node: [i] = C(n, i), i = 1..7
	std::vector<uint8_t> split;	   // 8 per node: [i] = slots given to the LEFT child when covering with i (0: cover with i-1);
								   // [0] = the same for the eight slots of the node's own wide node
	std::vector<uint8_t> asLeaf;   // covering with one child: 1 = a leaf of all its triangles, 0 = a wide node

	const float * loL(int b) const { return nodes + size_t(b) * 16; }
	const float * hiL(int b) const { return nodes + size_t(b) * 16 + 3; }
	const float * loR(int b) const { return nodes + size_t(b) * 16 + 6; }
	const float * hiR(int b) const { return nodes + size_t(b) * 16 + 9; }
	int refL(int b) const { return refOf(nodes[size_t(b) * 16 + 12]); }
	int refR(int b) const { return refOf(nodes[size_t(b) * 16 + 13]); }

	uint32_t primsOf(int ref) const { return isNone(ref) ? 0u : (isLeaf(ref) ? uint32_t(leafCount(ref)) : prims[ref]); }
	void table(int ref, const float * lo, const float * hi, double out[8]) const {
		for(int i = 1; i <= 7; ++i) {
			if(isNone(ref)) out[i] = 0.0;
			else if(isLeaf(ref)) out[i] = halfArea(lo, hi) * leafCount(ref) * kCostTri;
			else out[i] = cost[size_t(ref) * 8 + i];
		}
	}
};

// Fills the plan bottom-up. `box` of an interior node is the one its parent (or the root table) stores for it.
bool makePlan(const RbFlatBvh & flat, size_t nroots, Plan & P, std::string * why) {
	const size_t nI = flat.nodes.size() / 16;
	P.nodes = flat.nodes.data();
	P.prims.assign(nI, 0u);
	P.cost.assign(nI * 8, 0.0);
	P.split.assign(nI * 8, 0);
	P.asLeaf.assign(nI, 0);
	std::vector<float> box(nI * 6, 0.0f);
	std::vector<uint8_t> seen(nI, 0);
	std::string malformed;
	// Children's boxes live in their parent's record: expanding a node hands them down.
	auto expand = [&](int b, std::vector<int> & into) {
		const int refs[2]	 = {P.refL(b), P.refR(b)};
		const float * los[2] = {P.loL(b), P.loR(b)};
		const float * his[2] = {P.hiL(b), P.hiR(b)};
		for(int c = 0; c < 2; ++c) {
			if(isNone(refs[c]) || isLeaf(refs[c])) continue;
			if(size_t(refs[c]) >= nI || seen[refs[c]]) return false; // seen[] entries of different sub-trees are different bytes
			seen[refs[c]] = 1;
			std::memcpy(&box[size_t(refs[c]) * 6], los[c], 12);
			std::memcpy(&box[size_t(refs[c]) * 6 + 3], his[c], 12);
			into.push_back(refs[c]);
		}
		return true;
	};
	const double inf = std::numeric_limits<double>::infinity();
	// One node of the dynamic programme; its children are done.
	auto solve = [&](int b) {
		const int rl = P.refL(b), rr = P.refR(b);
		P.prims[b] = P.primsOf(rl) + P.primsOf(rr);
		double cl[8], cr[8];
		P.table(rl, P.loL(b), P.hiL(b), cl);
		P.table(rr, P.loR(b), P.hiR(b), cr);
		double dist[9];
		uint8_t dk[9];
		for(int j = 2; j <= 8; ++j) {
			dist[j] = inf;
			dk[j]	= 1;
			for(int k = 1; k < j && k <= 7; ++k) {
				if(j - k > 7) continue;
				const double c = cl[k] + cr[j - k];
				if(c < dist[j]) {
					dist[j] = c;
					dk[j]	= uint8_t(k);
				}
			}
		}
		const double area	   = halfArea(&box[size_t(b) * 6], &box[size_t(b) * 6 + 3]);
		const double cInternal = area * kCostNode + dist[8];
		const double cLeaf	   = P.prims[b] <= RB_WIDE_MAX_LEAF_TRIS ? area * P.prims[b] * kCostTri : inf;
		double * C			   = &P.cost[size_t(b) * 8];
		uint8_t * S			   = &P.split[size_t(b) * 8];
		S[0]				   = dk[8];
		P.asLeaf[b]			   = !(cInternal < cLeaf); // NaN areas (rejected later) fall on the leaf side harmlessly
		C[1]				   = P.asLeaf[b] ? cLeaf : cInternal;
		if(!(C[1] < inf)) {	   // more than three triangles and a NaN area: keep it a node
			P.asLeaf[b] = 0;
			C[1]		= cInternal;
		}
		for(int i = 2; i <= 7; ++i) {
			if(dist[i] < C[i - 1]) {
				C[i] = dist[i];
				S[i] = dk[i];
			} else {
				C[i] = C[i - 1];
				S[i] = 0;
			}
		}
	};
	// The top of the forest breadth first (parents before children) until there are enough sub-trees to share out;
	// every sub-tree is then walked and solved bottom-up by one thread; the top is solved last, in reverse.
	std::vector<int> top, frontier;
	for(size_t r = 0; r < nroots; ++r) {
		const float * d = flat.roots.data() + r * 8;
		const int ref	= refOf(d[3]);
		if(isNone(ref) || isLeaf(ref)) continue;
		if(size_t(ref) >= nI || seen[ref]) {
			if(why) *why = "malformed hierarchy (root reference)";
			return false;
		}
		seen[ref] = 1;
		for(int a = 0; a < 3; ++a) { box[size_t(ref) * 6 + a] = d[a]; box[size_t(ref) * 6 + 3 + a] = d[4 + a]; }
		frontier.push_back(ref);
	}
	const size_t wanted = nI >= (size_t(1) << 15) ? 256 : 0;
	for(size_t head = 0; head < frontier.size() && frontier.size() - head < wanted;) {
		const int b = frontier[head++];
		top.push_back(b);
		if(!expand(b, frontier)) {
			if(why) *why = "malformed hierarchy (a node is referenced twice)";
			return false;
		}
	}
	frontier.erase(frontier.begin(), frontier.begin() + long(top.size()));
	std::vector<uint8_t> bad(frontier.size(), 0);
	rb_parallel_for(frontier.size(), 1, [&](size_t fb, size_t fe) {
		std::vector<int> order, stack;
		for(size_t f = fb; f < fe; ++f) {
			order.clear();
			stack.assign(1, frontier[f]);
			while(!stack.empty()) {
				const int b = stack.back();
				stack.pop_back();
				order.push_back(b);
				if(!expand(b, stack)) {
					bad[f] = 1;
					break;
				}
			}
			if(bad[f]) continue;
			for(size_t q = order.size(); q-- > 0;) solve(order[q]);
		}
	});
	for(uint8_t b : bad) {
		if(b) {
			if(why) *why = "malformed hierarchy (a node is referenced twice)";
			return false;
		}
	}
	for(size_t q = top.size(); q-- > 0;) solve(top[q]);
	return true;
}

void gatherTris(const Plan & P, int ref, WChild & c) {
	if(isNone(ref)) return;
	if(isLeaf(ref)) {
		for(int k = 0; k < leafCount(ref) && c.ntris < RB_WIDE_MAX_LEAF_TRIS; ++k) c.tri[c.ntris++] = leafFirst(ref) + k;
		return;
	}
	gatherTris(P, P.refR(ref), c); // right first: the reference's visiting order
	gatherTris(P, P.refL(ref), c);
}

// Appends to `list` the children that cover `ref` with at most i slots, in visiting order.
void collect(const Plan & P, int ref, const float * lo, const float * hi, int i, std::vector<WChild> & list) {
	if(isNone(ref)) return;
	WChild c;
	c.node	= -1;
	c.ntris = 0;
	std::memcpy(c.lo, lo, 12);
	std::memcpy(c.hi, hi, 12);
	if(isLeaf(ref)) {
		gatherTris(P, ref, c);
		list.push_back(c);
		return;
	}
	while(i > 1 && P.split[size_t(ref) * 8 + i] == 0) --i;
	if(i == 1) {
		if(P.asLeaf[ref]) gatherTris(P, ref, c);
		else c.node = ref;
		list.push_back(c);
		return;
	}
	const int k = P.split[size_t(ref) * 8 + i];
	collect(P, P.refR(ref), P.loR(ref), P.hiR(ref), i - k, list);
	collect(P, P.refL(ref), P.loL(ref), P.hiL(ref), k, list);
}

struct WorkItem {
	int wide;  // index of the wide node to fill
	int ref;   // binary reference it stands for (an interior node, or a leaf for a mesh that is a single leaf)
	float lo[3], hi[3];
};

} // namespace

bool rb_wide_collapse(const RbFlatBvh & flat, size_t nroots, RbWideBvh & out, std::string * why) {
	const size_t nTris = flat.tris.size() / 12;
	for(int code : flat.leafCode) {
		if((code & (RB_LEAF_MAX_TRIS - 1)) + 1 > RB_WIDE_MAX_LEAF_TRIS) {
			if(why) *why = "leaves of more than three triangles are not supported by the wide layout";
			return false;
		}
	}
	Plan P;
	const auto t0 = std::chrono::steady_clock::now();
	if(!makePlan(flat, nroots, P, why)) return false;
	if(std::getenv("RB_BUILD_TIMING")) fprintf(stderr, "[rb_wide_collapse] plan %.1f ms\n", std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
	out.nodes.clear();
	out.nodesByOctant.clear();
	out.tris.clear();
	out.tris.reserve(flat.tris.size());
	out.rootNode.assign(nroots, -1);
	out.depth = 0;
	std::vector<WorkItem> work; // the nodes of the current level
	for(size_t r = 0; r < nroots; ++r) {
		const float * d = flat.roots.data() + r * 8;
		const int ref	= refOf(d[3]);
		if(isNone(ref)) continue;
		WorkItem w;
		w.wide	= int(out.nodes.size() / RB_WIDE_NODE_WORDS);
		w.ref	= ref;
		for(int a = 0; a < 3; ++a) { w.lo[a] = d[a]; w.hi[a] = d[4 + a]; }
		out.rootNode[r] = w.wide;
		out.nodes.resize(out.nodes.size() + RB_WIDE_NODE_WORDS, 0u);
		work.push_back(w);
	}
	// Level by level: the nodes of a level are prepared on all threads (children, quantised planes, counts), a prefix sum
	// hands out the node and triangle ranges of their children, and a second parallel pass writes them.
	struct Prepared {
		uint32_t words[RB_WIDE_NODE_WORDS];
		int nInterior = 0, nTris = 0;
		int childRef[8];
		float childLo[8][3], childHi[8][3];
		int tri[8 * RB_WIDE_MAX_LEAF_TRIS];
		std::string error;
	};
	auto prepare = [&](const WorkItem & item, Prepared & pr) {
		std::vector<WChild> list;
		list.reserve(8);
		if(isLeaf(item.ref)) {
			collect(P, item.ref, item.lo, item.hi, 1, list);
		} else {
			const int b = item.ref, k = P.split[size_t(b) * 8];
			collect(P, P.refR(b), P.loR(b), P.hiR(b), 8 - k, list);
			collect(P, P.refL(b), P.loL(b), P.hiL(b), k, list);
		}
		if(list.size() > 8) {
			pr.error = "internal error: a wide node with more than eight children";
			return;
		}
		uint32_t * node = pr.words;
		std::memset(node, 0, sizeof(pr.words));
		uint8_t q[6][8];
		for(int s = 0; s < 8; ++s) {
			q[0][s] = q[1][s] = q[2][s] = 255;
			q[3][s] = q[4][s] = q[5][s] = 0;
		}
		uint32_t expo = 0;
		for(int a = 0; a < 3; ++a) {
			double lo = INFINITY, hi = -INFINITY;
			for(const WChild & c : list) {
				if(!std::isfinite(c.lo[a]) || !std::isfinite(c.hi[a]) || c.lo[a] > c.hi[a]) {
					pr.error = "a box with non-finite or inverted bounds cannot be quantised";
					return;
				}
				lo = std::fmin(lo, double(c.lo[a]));
				hi = std::fmax(hi, double(c.hi[a]));
			}
			if(list.empty()) lo = hi = 0.0;
			int e = -100;
			if(hi > lo) {
				const int guess = int(std::floor(std::log2((hi - lo) / 255.0))) - 1;
				if(guess > e) e = guess;
			}
			float p	 = 0.0f;
			double s = 0.0;
			for(;; ++e) {
				if(e > 40) {
					pr.error = "box extent out of the range of the quantisation exponents";
					return;
				}
				s = std::ldexp(1.0, e);
				p = roundDown(lo - 2.0 * RB_WIDE_MARGIN * s);
				// The margin below the lowest plane must survive the rounding of p to fp32 (a flat or tiny node far from
				// the origin: the quantum has to grow until it does).
				if(std::isfinite(p) && (lo - double(p)) / s >= 2.0 * RB_WIDE_MARGIN && (hi - double(p)) / s + RB_WIDE_MARGIN <= 255.0) break;
			}
			std::memcpy(&node[a], &p, 4);
			expo |= uint32_t(e + 127) << (8 * a);
			for(size_t i = 0; i < list.size(); ++i) {
				const double l = std::floor((double(list[i].lo[a]) - double(p)) / s - RB_WIDE_MARGIN);
				const double h = std::ceil((double(list[i].hi[a]) - double(p)) / s + RB_WIDE_MARGIN);
				if(!(l >= 0.0 && h <= 255.0 && l <= h)) {
					char buf[256];
					snprintf(buf, sizeof(buf), "internal error: quantised plane out of range (axis %d: child [%g, %g], origin %g, scale 2^%d, q [%g, %g])", a, double(list[i].lo[a]), double(list[i].hi[a]), double(p), e, l, h);
					pr.error = buf;
					return;
				}
				q[a][i]		= uint8_t(l);
				q[3 + a][i] = uint8_t(h);
			}
		}
		node[3] = expo | (uint32_t(list.size()) << 24);
		// Children: interior ones get consecutive nodes, leaves consecutive triangles, both in slot order.
		uint8_t meta[8] = {0, 0, 0, 0, 0, 0, 0, 0};
		pr.nInterior = pr.nTris = 0;
		for(size_t i = 0; i < list.size(); ++i) {
			if(list[i].node >= 0) {
				meta[i] = uint8_t(0x20 | (24 + pr.nInterior));
				pr.childRef[pr.nInterior] = list[i].node;
				std::memcpy(pr.childLo[pr.nInterior], list[i].lo, 12);
				std::memcpy(pr.childHi[pr.nInterior], list[i].hi, 12);
				++pr.nInterior;
				continue;
			}
			const int count = list[i].ntris;
			if(count < 1 || count > RB_WIDE_MAX_LEAF_TRIS) {
				pr.error = "internal error: leaf child without triangles";
				return;
			}
			meta[i] = uint8_t((((1u << count) - 1u) << 5) | uint32_t(pr.nTris));
			for(int k = 0; k < count; ++k) {
				if(size_t(list[i].tri[k]) >= nTris) {
					pr.error = "leaf triangle range out of bounds";
					return;
				}
				pr.tri[pr.nTris++] = list[i].tri[k];
			}
		}
		std::memcpy(&node[6], meta, 8);
		for(int plane = 0; plane < 6; ++plane) std::memcpy(&node[8 + 2 * plane], q[plane], 8);
	};
	std::vector<WorkItem> next;
	std::vector<Prepared> prepared;
	std::vector<size_t> nodeAt, triAt;
	while(!work.empty()) {
		++out.depth;
		const size_t n = work.size();
		prepared.assign(n, Prepared());
		rb_parallel_for(n, 256, [&](size_t b, size_t e) {
			for(size_t i = b; i < e; ++i) prepare(work[i], prepared[i]);
		});
		nodeAt.assign(n + 1, out.nodes.size() / RB_WIDE_NODE_WORDS);
		triAt.assign(n + 1, out.tris.size() / 12);
		for(size_t i = 0; i < n; ++i) {
			if(!prepared[i].error.empty()) {
				if(why) *why = prepared[i].error;
				return false;
			}
			nodeAt[i + 1] = nodeAt[i] + size_t(prepared[i].nInterior);
			triAt[i + 1]  = triAt[i] + size_t(prepared[i].nTris);
		}
		if(triAt[n] >= (size_t(1) << 31) || nodeAt[n] >= (size_t(1) << 31)) {
			if(why) *why = "too many nodes or triangles for 32-bit references";
			return false;
		}
		const size_t firstChild = nodeAt[0];
		out.nodes.resize(nodeAt[n] * RB_WIDE_NODE_WORDS, 0u);
		out.tris.resize(triAt[n] * 12);
		next.assign(nodeAt[n] - firstChild, WorkItem());
		rb_parallel_for(n, 256, [&](size_t b, size_t e) {
			for(size_t i = b; i < e; ++i) {
				Prepared & pr	= prepared[i];
				pr.words[4]		= uint32_t(nodeAt[i]);
				pr.words[5]		= uint32_t(triAt[i]);
				std::memcpy(out.nodes.data() + size_t(work[i].wide) * RB_WIDE_NODE_WORDS, pr.words, sizeof(pr.words));
				for(int k = 0; k < pr.nTris; ++k) {
					float * dst = out.tris.data() + (triAt[i] + size_t(k)) * 12;
					std::memcpy(dst, flat.tris.data() + size_t(pr.tri[k]) * 12, 48);
					const int32_t hier = int32_t(pr.tri[k]);
					std::memcpy(dst + 11, &hier, 4);
				}
				for(int k = 0; k < pr.nInterior; ++k) {
					WorkItem & w = next[nodeAt[i] - firstChild + size_t(k)];
					w.wide		 = int(nodeAt[i] + size_t(k));
					w.ref		 = pr.childRef[k];
					std::memcpy(w.lo, pr.childLo[k], 12);
					std::memcpy(w.hi, pr.childHi[k], 12);
				}
			}
		});
		work.swap(next);
	}
	// The copy with slots by direction, node by node on all threads: greedy assignment of (child, slot) pairs by decreasing
	// dot(child centre - node centre, corner direction of the slot), centres taken from the stored (quantised) boxes.
	out.nodesByOctant.assign(out.nodes.size(), 0u);
	rb_parallel_for(out.nodes.size() / RB_WIDE_NODE_WORDS, 4096, [&](size_t nb, size_t ne) {
		for(size_t n = nb; n < ne; ++n) {
			const uint32_t * src = out.nodes.data() + n * RB_WIDE_NODE_WORDS;
			uint32_t * dst		 = out.nodesByOctant.data() + n * RB_WIDE_NODE_WORDS;
			std::memcpy(dst, src, RB_WIDE_NODE_WORDS * sizeof(uint32_t));
			uint8_t meta[8], q[6][8];
			std::memcpy(meta, &src[6], 8);
			for(int plane = 0; plane < 6; ++plane) std::memcpy(q[plane], &src[8 + 2 * plane], 8);
			double scale[3], lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
			for(int a = 0; a < 3; ++a) scale[a] = std::ldexp(1.0, int((src[3] >> (8 * a)) & 0xffu) - 127);
			int count = 0;
			for(int i = 0; i < 8; ++i) {
				if(meta[i] == 0) continue;
				++count;
				for(int a = 0; a < 3; ++a) {
					lo[a] = std::fmin(lo[a], double(q[a][i]) * scale[a]);
					hi[a] = std::fmax(hi[a], double(q[3 + a][i]) * scale[a]);
				}
			}
			double towards[8][8]; // [child][slot]
			for(int i = 0; i < 8; ++i) {
				if(meta[i] == 0) continue;
				double off[3];
				for(int a = 0; a < 3; ++a) off[a] = 0.5 * (double(q[a][i]) + double(q[3 + a][i])) * scale[a] - 0.5 * (lo[a] + hi[a]);
				for(int sl = 0; sl < 8; ++sl) towards[i][sl] = ((sl & 1) ? off[0] : -off[0]) + ((sl & 2) ? off[1] : -off[1]) + ((sl & 4) ? off[2] : -off[2]);
			}
			int slotOf[8], childAt[8];
			for(int i = 0; i < 8; ++i) slotOf[i] = childAt[i] = -1;
			for(int round = 0; round < count; ++round) {
				double best = -INFINITY;
				int bi = -1, bs = -1;
				for(int i = 0; i < 8; ++i) {
					if(meta[i] == 0 || slotOf[i] >= 0) continue;
					for(int sl = 0; sl < 8; ++sl) {
						if(childAt[sl] >= 0) continue;
						if(towards[i][sl] > best) {
							best = towards[i][sl];
							bi	 = i;
							bs	 = sl;
						}
					}
				}
				slotOf[bi]	= bs;
				childAt[bs] = bi;
			}
			uint8_t metaO[8], qO[6][8];
			for(int sl = 0; sl < 8; ++sl) {
				const int i = childAt[sl];
				metaO[sl]	= i >= 0 ? meta[i] : 0;
				for(int plane = 0; plane < 6; ++plane) qO[plane][sl] = i >= 0 ? q[plane][i] : (plane < 3 ? 255 : 0);
			}
			std::memcpy(&dst[6], metaO, 8);
			for(int plane = 0; plane < 6; ++plane) std::memcpy(&dst[8 + 2 * plane], qO[plane], 8);
		}
	});
	return true;
}
