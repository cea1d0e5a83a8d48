// rb_wide.cpp — host collapse of the reference-order binary tree into the eight-wide quantised tree of rb_wide.h.
#include "rb_wide.h"
#include "rb_internal.h"

#include <cmath>
#include <cstdio>
#include <cstring>

namespace {

struct WChild {
	int ref; // binary child reference: >= 0 interior node, < 0 leaf code
	float lo[3], hi[3];
};

inline int refOf(float f) {
	int v;
	std::memcpy(&v, &f, 4);
	return v;
}

// The two children of binary interior node b in VISITING order (right first, Raycaster.cpp:160-173); empty ones dropped.
inline int childrenOf(const float * nodes, int b, WChild out[2]) {
	const float * d = nodes + size_t(b) * 16;
	int n = 0;
	const int refR = refOf(d[13]), refL = refOf(d[12]);
	if(refR != RB_REF_NONE) {
		out[n].ref = refR;
		for(int a = 0; a < 3; ++a) { out[n].lo[a] = d[6 + a]; out[n].hi[a] = d[9 + a]; }
		++n;
	}
	if(refL != RB_REF_NONE) {
		out[n].ref = refL;
		for(int a = 0; a < 3; ++a) { out[n].lo[a] = d[a]; out[n].hi[a] = d[3 + a]; }
		++n;
	}
	return n;
}

inline double halfArea(const WChild & c) {
	const double x = double(c.hi[0]) - c.lo[0], y = double(c.hi[1]) - c.lo[1], z = double(c.hi[2]) - c.lo[2];
	return x * y + y * z + z * x;
}

inline float roundDown(double x) {
	float f = float(x);
	if(double(f) > x) f = std::nextafterf(f, -INFINITY);
	return f;
}

struct WorkItem {
	int wide;	// index of the wide node to fill
	int ref;	// binary reference it stands for (interior node, or a leaf code for a mesh that is a single leaf)
	int depth;
	float lo[3], hi[3]; // box of that reference (only used for the single-leaf case)
};

} // namespace

bool rb_wide_collapse(const RbFlatBvh & flat, size_t nroots, RbWideBvh & out, std::string * why) {
	const float * bn = flat.nodes.data();
	const size_t nTris = flat.tris.size() / 12;
	out.nodes.clear();
	out.tris.clear();
	out.tris.reserve(flat.tris.size());
	out.rootNode.assign(nroots, -1);
	out.depth = 0;
	std::vector<WorkItem> work;
	for(size_t r = 0; r < nroots; ++r) {
		const float * d = flat.roots.data() + r * 8;
		const int ref	= refOf(d[3]);
		if(ref == RB_REF_NONE) continue;
		WorkItem w;
		w.wide	= int(out.nodes.size() / RB_WIDE_NODE_WORDS);
		w.ref	= ref;
		w.depth = 1;
		for(int a = 0; a < 3; ++a) { w.lo[a] = d[a]; w.hi[a] = d[4 + a]; }
		out.rootNode[r] = w.wide;
		out.nodes.resize(out.nodes.size() + RB_WIDE_NODE_WORDS, 0u);
		work.push_back(w);
	}
	std::vector<WChild> list;
	while(!work.empty()) {
		const WorkItem item = work.back();
		work.pop_back();
		if(item.depth > out.depth) out.depth = item.depth;
		// The up-to-eight descendants this node adopts, in visiting order.
		list.clear();
		if(item.ref < 0) {
			WChild c;
			c.ref = item.ref;
			std::memcpy(c.lo, item.lo, sizeof(c.lo));
			std::memcpy(c.hi, item.hi, sizeof(c.hi));
			list.push_back(c);
		} else {
			WChild two[2];
			const int n = childrenOf(bn, item.ref, two);
			list.assign(two, two + n);
			for(;;) {
				if(list.size() >= 8) break;
				int pick = -1;
				double best = -1.0;
				for(size_t i = 0; i < list.size(); ++i) {
					if(list[i].ref < 0) continue;
					const double area = halfArea(list[i]);
					if(!(area <= best)) { // NaN areas are picked too: they are rejected below
						best = area;
						pick = int(i);
					}
				}
				if(pick < 0) break;
				WChild two2[2];
				const int m = childrenOf(bn, list[pick].ref, two2);
				list.erase(list.begin() + pick);
				list.insert(list.begin() + pick, two2, two2 + m);
			}
		}
		uint32_t * node = out.nodes.data() + size_t(item.wide) * RB_WIDE_NODE_WORDS;
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
					if(why) *why = "a box with non-finite or inverted bounds cannot be quantised";
					return false;
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
					if(why) *why = "box extent out of the range of the quantisation exponents";
					return false;
				}
				s = std::ldexp(1.0, e);
				p = roundDown(lo - 2.0 * RB_WIDE_MARGIN * s);
				// The margin below the lowest plane must survive the rounding of p to fp32 (a flat or tiny node far from
				// the origin: the quantum has to grow until it does).
				if(std::isfinite(p) && (lo - double(p)) / s >= 2.0 * RB_WIDE_MARGIN && (hi - double(p)) / s + RB_WIDE_MARGIN <= 255.0) break;
			}
			float pf = p;
			std::memcpy(&node[a], &pf, 4);
			expo |= uint32_t(e + 127) << (8 * a);
			for(size_t i = 0; i < list.size(); ++i) {
				const double l = std::floor((double(list[i].lo[a]) - double(p)) / s - RB_WIDE_MARGIN);
				const double h = std::ceil((double(list[i].hi[a]) - double(p)) / s + RB_WIDE_MARGIN);
				if(!(l >= 0.0 && h <= 255.0 && l <= h)) {
					if(why) {
						char buf[256];
						snprintf(buf, sizeof(buf), "internal error: quantised plane out of range (axis %d: child [%g, %g], origin %g, scale 2^%d, q [%g, %g])", a, double(list[i].lo[a]), double(list[i].hi[a]), double(p), e, l, h);
						*why = buf;
					}
					return false;
				}
				q[a][i]		= uint8_t(l);
				q[3 + a][i] = uint8_t(h);
			}
		}
		node[3] = expo | (uint32_t(list.size()) << 24);
		// Children: interior ones get consecutive nodes, leaves consecutive triangles, both in slot order.
		const int childBase = int(out.nodes.size() / RB_WIDE_NODE_WORDS);
		const size_t triBase = out.tris.size() / 12;
		uint8_t meta[8] = {0, 0, 0, 0, 0, 0, 0, 0};
		int nInterior = 0, triOffset = 0;
		for(size_t i = 0; i < list.size(); ++i) {
			if(list[i].ref >= 0) {
				meta[i] = uint8_t(0x20 | (24 + nInterior));
				++nInterior;
			} else {
				const int code	= ~list[i].ref;
				const int first = code >> RB_LEAF_BITS;
				const int count = (code & (RB_LEAF_MAX_TRIS - 1)) + 1;
				if(count > RB_WIDE_MAX_LEAF_TRIS) {
					if(why) *why = "leaves of more than three triangles are not supported by the wide layout";
					return false;
				}
				if(size_t(first) + size_t(count) > nTris) {
					if(why) *why = "leaf triangle range out of bounds";
					return false;
				}
				meta[i] = uint8_t((((1u << count) - 1u) << 5) | uint32_t(triOffset));
				for(int k = 0; k < count; ++k) {
					const float * src = flat.tris.data() + size_t(first + k) * 12;
					const size_t at	  = out.tris.size();
					out.tris.insert(out.tris.end(), src, src + 12);
					const int32_t hier = int32_t(first + k);
					std::memcpy(&out.tris[at + 11], &hier, 4);
				}
				triOffset += count;
			}
		}
		if(triBase + size_t(triOffset) >= (size_t(1) << 31)) {
			if(why) *why = "too many triangles for 32-bit references";
			return false;
		}
		out.nodes.resize(out.nodes.size() + size_t(nInterior) * RB_WIDE_NODE_WORDS, 0u);
		node = out.nodes.data() + size_t(item.wide) * RB_WIDE_NODE_WORDS; // resize may have moved the array
		node[4] = uint32_t(childBase);
		node[5] = uint32_t(triBase);
		std::memcpy(&node[6], meta, 8);
		for(int plane = 0; plane < 6; ++plane) std::memcpy(&node[8 + 2 * plane], q[plane], 8);
		int k = 0;
		for(size_t i = 0; i < list.size(); ++i) {
			if(list[i].ref < 0) continue;
			WorkItem w;
			w.wide	= childBase + k++;
			w.ref	= list[i].ref;
			w.depth = item.depth + 1;
			std::memcpy(w.lo, list[i].lo, sizeof(w.lo));
			std::memcpy(w.hi, list[i].hi, sizeof(w.hi));
			work.push_back(w);
		}
	}
	return true;
}
