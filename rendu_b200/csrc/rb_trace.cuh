// rb_trace.cuh — device-side BVH traversal: the B200 restatement of Raycaster::intersects / intersectsAny
// (reference: src/engine/raycaster/Raycaster.cpp:144-262, src/engine/raycaster/Intersection.cpp:3-5,22-32).
//
// Device layout (all 16-byte aligned, fetched with 128-bit loads through the read-only path):
//   nodes : 4 x float4 per INTERIOR node (64 B): both child boxes and both child references live in the parent, so
//           one visit = one 64-byte fetch instead of the reference's two dependent 48-byte Node loads.
//             n0 = (L.min.x, L.min.y, L.min.z, L.max.x)   n1 = (L.max.y, L.max.z, R.min.x, R.min.y)
//             n2 = (R.min.z, R.max.x, R.max.y, R.max.z)   n3 = (bits(leftRef), bits(rightRef), 0, 0)
//   tris  : 3 x float4 per triangle (48 B) in hierarchy order, vertices pre-gathered:
//             t0 = (v0.xyz, bits(localId))  t1 = (v1-v0, bits(meshId))  t2 = (v2-v0, 0)
//           v1-v0 / v2-v0 are the reference's `v01`/`v02` (Raycaster.cpp:235-236) computed once at upload; fp32
//           subtraction is deterministic so the stored values equal what the reference recomputes per test.
//   roots : 2 x float4 per mesh root: (min.xyz, bits(ref)) (max.xyz, 0); tested linearly as the reference does
//           (Raycaster.cpp:149-155).
//   A child reference is >= 0 for an interior node index, < 0 for a leaf: ~ref = (firstTriangle << 3) | (count-1).
#pragma once
#include "rb_math.cuh"
#include "rb_types.h"
#include <stdint.h>

struct RbRay {
	V3 pos, dir, inv;
	// Packed copies for the two-wide fp32 instructions of sm_100 (FADD2 / FMUL2): pairs (x,y), (z,x), (y,z) line up with
	// the pairs a node's float4 loads deliver (see rb_node_test).
	unsigned long long pxy, pzx, pyz, ixy, izx, iyz;
	bool regular; // every component of pos and inv is finite: min/max can never see a NaN
};

#ifndef RB_HOST_CHECK
RB_D unsigned long long rb_pack2(float a, float b) {
	unsigned long long r;
	asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
	return r;
}
RB_D void rb_unpack2(unsigned long long v, float & a, float & b) {
	asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
}
// a * b + c on two fp32 lanes at once, each lane one fused multiply-add (round to nearest even) like the scalar FFMA.
RB_D unsigned long long rb_fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
	unsigned long long d;
	asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
	return d;
}
// (a - b) * c on two fp32 lanes at once; each lane rounds exactly like the scalar sub.rn / mul.rn pair (no fusion).
RB_D unsigned long long rb_submul2(unsigned long long a, unsigned long long b, unsigned long long c) {
	unsigned long long d, r;
	asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
	asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(d), "l"(c));
	return r;
}
// Shared-memory word at a 32-bit shared-window address (the traversal stacks).
RB_D void rb_sts(unsigned int address, unsigned int value) {
	asm volatile("st.shared.b32 [%0], %1;" ::"r"(address), "r"(value) : "memory");
}
RB_D unsigned int rb_lds(unsigned int address) {
	unsigned int value;
	asm volatile("ld.shared.b32 %0, [%1];" : "=r"(value) : "r"(address) : "memory");
	return value;
}
#endif // RB_HOST_CHECK: tests/host_check/ compiles this header with g++ and supplies these six primitives itself

RB_D RbRay rb_make_ray(V3 origin, V3 direction) {
	RbRay r;
	r.pos = origin;
	r.dir = normalize3(direction);
	r.inv = V3{1.0f / r.dir.x, 1.0f / r.dir.y, 1.0f / r.dir.z};
	r.regular = isfinite(r.inv.x) && isfinite(r.inv.y) && isfinite(r.inv.z) && isfinite(origin.x) && isfinite(origin.y) && isfinite(origin.z);
	r.pxy = rb_pack2(r.pos.x, r.pos.y); r.pzx = rb_pack2(r.pos.z, r.pos.x); r.pyz = rb_pack2(r.pos.y, r.pos.z);
	r.ixy = rb_pack2(r.inv.x, r.inv.y); r.izx = rb_pack2(r.inv.z, r.inv.x); r.iyz = rb_pack2(r.inv.y, r.inv.z);
	return r;
}

// Intersection::box (Intersection.cpp:22-32). The literal form keeps the reference's NaN behaviour (0*inf when a ray
// is parallel to a slab and starts on it); when no NaN can occur (regular ray, finite box) fminf/fmaxf select the
// same operands, so the cheaper FMNMX form returns the same boolean.
template<bool REGULAR>
RB_D bool rb_box_test_t(const RbRay & ray, float bminx, float bminy, float bminz, float bmaxx, float bmaxy, float bmaxz, float mini, float maxi) {
	const float minRx = (bminx - ray.pos.x) * ray.inv.x;
	const float minRy = (bminy - ray.pos.y) * ray.inv.y;
	const float minRz = (bminz - ray.pos.z) * ray.inv.z;
	const float maxRx = (bmaxx - ray.pos.x) * ray.inv.x;
	const float maxRy = (bmaxy - ray.pos.y) * ray.inv.y;
	const float maxRz = (bmaxz - ray.pos.z) * ray.inv.z;
	if(REGULAR) {
		const float closest	 = fmaxf(fmaxf(fminf(minRx, maxRx), fminf(minRy, maxRy)), fmaxf(fminf(minRz, maxRz), mini));
		const float furthest = fminf(fminf(fmaxf(minRx, maxRx), fmaxf(minRy, maxRy)), fminf(fmaxf(minRz, maxRz), maxi));
		return closest <= furthest;
	}
	const float minFx = glm_min(minRx, maxRx), minFy = glm_min(minRy, maxRy), minFz = glm_min(minRz, maxRz);
	const float maxFx = glm_max(minRx, maxRx), maxFy = glm_max(minRy, maxRy), maxFz = glm_max(minRz, maxRz);
	const float closest	 = std_max(minFx, std_max(minFy, minFz));
	const float furthest = std_min(maxFx, std_min(maxFy, maxFz));
	return std_max(closest, mini) <= std_min(furthest, maxi);
}

RB_D bool rb_box_test(const RbRay & ray, float bminx, float bminy, float bminz, float bmaxx, float bmaxy, float bmaxz, float mini, float maxi) {
	return ray.regular ? rb_box_test_t<true>(ray, bminx, bminy, bminz, bmaxx, bmaxy, bmaxz, mini, maxi)
					   : rb_box_test_t<false>(ray, bminx, bminy, bminz, bmaxx, bmaxy, bmaxz, mini, maxi);
}

// Raycaster::intersects(ray, tri, mini, maxi) (Raycaster.cpp:232-262), Möller-Trumbore with the reference's
// comparisons written literally (their outcome under NaN matters).
RB_D bool rb_tri_test(const RbRay & ray, V3 v0, V3 v01, V3 v02, float mini, float maxi, float & t, float & u, float & v) {
	const V3 p		= cross3(ray.dir, v02);
	const float det = dot3(v01, p);
	if(fabsf(det) < FLT_EPSILON) {
		return false;
	}
	const float invDet = 1.0f / det;
	const V3 q		   = ray.pos - v0;
	u				   = invDet * dot3(q, p);
	if(u < 0.0f || u > 1.0f) {
		return false;
	}
	const V3 r = cross3(q, v01);
	v		   = invDet * dot3(ray.dir, r);
	if(v < 0.0f || (u + v) > 1.0f) {
		return false;
	}
	t = invDet * dot3(v02, r);
	return (t > mini && t < maxi);
}

struct RbHitRec {
	float t, u, v;
	int tri; // index in hierarchy order, -1 = miss
};

struct RbWork {
	unsigned int interior, tris, roots, maxStack;
};

template<bool REGULAR>
RB_D void rb_node_test(const RbRay & ray, const float4 & n0, const float4 & n1, const float4 & n2, float mini, float maxi, bool & hitL, bool & hitR, float & dL, float & dR);

RB_D float4 rb_ldg4(const float4 * p) {
	return __ldg(p);
}

// A 64-byte interior node as two 256-bit loads (LDG.E.256, sm_100 / PTX 8.8) instead of four 128-bit ones: a divergent
// load costs one L1 data-pipe wavefront per lane and request whatever its width (up to the 32-byte sector), and that pipe
// is what bounds the incoherent launches. Same bytes, same read-only path: results are unchanged.
#ifndef RB_HOST_CHECK
RB_D void rb_ldg_node(const float4 * n, float4 & n0, float4 & n1, float4 & n2, float4 & n3) {
#if defined(RB_NODE_LDG128)
	n0 = __ldg(n); n1 = __ldg(n + 1); n2 = __ldg(n + 2); n3 = __ldg(n + 3);
#else
	asm("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
		: "=f"(n0.x), "=f"(n0.y), "=f"(n0.z), "=f"(n0.w), "=f"(n1.x), "=f"(n1.y), "=f"(n1.z), "=f"(n1.w) : "l"(n));
	asm("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
		: "=f"(n2.x), "=f"(n2.y), "=f"(n2.z), "=f"(n2.w), "=f"(n3.x), "=f"(n3.y), "=f"(n3.z), "=f"(n3.w) : "l"(n + 2));
#endif
}
#endif

// Closest hit / any hit in the REFERENCE ORDER: roots pre-tested against the caller's (mini, maxi) and consumed last
// mesh first, depth-first with the right child popped before the left, `maxi` shrinking on every accepted hit
// (strict <, so the first triangle found wins exact ties). Equivalent, decision for decision, to the std::stack loop
// of Raycaster.cpp:144-185 (closest) and :187-224 (any): the reference pushes left then right and pops right; here
// the right child is followed directly and only the left one is stored. Roots are box-tested lazily but with the
// ORIGINAL maxi, which is the value the reference used when it pre-pushed them.
//
// The traversal is a resumable state machine (RbTraversal + rb_step_*) so that the one-ray-per-thread kernels and
// the persistent dynamic-fetch kernels share one body. One step = take the next pending node (stack, else next
// accepted root), descend while on interior nodes, then test the triangles of the leaf reached.
// `stack` points at this thread's column of a shared-memory array, entries `stride` ints apart.
struct RbTraversal {
	RbHitRec best;
	float mini, maxi, maxi0;
	int cur, root;
	unsigned int top; // shared-memory address of the next free stack entry (== base when the stack is empty)
	int rootHi;		  // RB_ROOTS_TREE only: roots 32..63 of the mask of entered roots (`root` holds 0..31)
};

// The stack is addressed in the shared window directly (st.shared / ld.shared on a 32-bit address kept in a register):
// going through a generic int* costs a generic-to-shared conversion (S2UR/ULEA/LEA) at every push and pop.
struct RbStack {
	unsigned int base; // address of this thread's entry 0
	int strideBytes;   // distance between consecutive entries of one thread
};
RB_D RbStack rb_make_stack(int * sharedColumn, int strideEntries) {
	RbStack s;
	s.base		  = (unsigned int)__cvta_generic_to_shared(sharedColumn);
	s.strideBytes = strideEntries * (int)sizeof(int);
	return s;
}
RB_D void rb_push(RbTraversal & T, const RbStack & st, int v) {
	rb_sts(T.top, (unsigned int)v);
	T.top += st.strideBytes;
}
RB_D int rb_pop(RbTraversal & T, const RbStack & st) {
	T.top -= st.strideBytes;
	return (int)rb_lds(T.top);
}
RB_D bool rb_stack_empty(const RbTraversal & T, const RbStack & st) { return T.top == st.base; }
RB_D unsigned int rb_stack_size(const RbTraversal & T, const RbStack & st) { return (T.top - st.base) / (unsigned int)st.strideBytes; }

RB_D void rb_traversal_init(RbTraversal & T, const RbDeviceBvh & bvh, float mini, float maxi, const RbStack & st) {
	T.best.t   = FLT_MAX;
	T.best.u   = 0.0f;
	T.best.v   = 0.0f;
	T.best.tri = -1;
	T.mini	   = mini;
	T.maxi	   = maxi;
	T.maxi0	   = maxi;
	T.cur	   = RB_REF_NONE;
	T.top	   = st.base;
	T.root	   = bvh.nroots;
	T.rootHi   = 0;
}

// Next root whose box the ray enters, last mesh first: the reference tests every root against the caller's (mini, maxi)
// and its stack hands them back in reverse (Raycaster.cpp:149-155). A scene with hundreds of objects has hundreds of
// roots; for them the handle carries union boxes of aligned blocks of 2^k consecutive roots. Arriving at the last root
// of such a block, the block's box is tested first and the whole block is skipped on a miss. The visiting order of the
// roots that are entered is unchanged, and no root the reference would enter is skipped: the slab test is monotone in
// the box (every fp32 operation in it is), so a ray that misses the union misses each member — for regular rays; the
// NaN-carrying irregular rays take the plain loop. BLOCKS is a template switch: with the block code inlined the node
// step of every kernel runs 15 % slower (register allocation of the hot loop; out of line it is worse still, the call
// forces the traversal state into local memory), so only handles that carry blocks launch the BLOCKS = true kernels.
#define RB_ROOTS_LAZY 0	  // every root box-tested when the ray gets to it
#define RB_ROOTS_BLOCKS 1 // ... with whole blocks of consecutive roots skipped on a miss (many-mesh scenes)
#define RB_ROOTS_MASK 2	  // at most 32 roots, all box-tested when the ray starts (rb_roots_pretest): T.root is the bit mask of those entered
#define RB_ROOTS_TREE 3	  // 9..64 roots: the mask (T.root, T.rootHi) comes from one walk of the tree over the root boxes (rb_roots_tree_pretest)
#define RB_ROOTS_TREE_BIG 4 // 65..512 roots: the same walk, the mask in the first bvh.rootWords entries of the ray's stack column (T.root = word index)

// The reference box-tests EVERY root against the caller's (mini, maxi) before it pops the first one (Raycaster.cpp:149-155).
// For a handful of roots the same is done here, up front: the tests are independent (their loads overlap), where the lazy
// loop pays one dependent load-test-branch round trip per root in the middle of the traversal — in a scene of seven meshes
// that was 40 % of the closest-hit wavefront's instructions and most of its serial latency. Same tests, same (mini, maxi),
// same order of the roots entered: same hits. Callers use it for nroots <= 32 in the reference's visiting order.
template<bool REGULAR, bool COUNT>
RB_D unsigned int rb_roots_pretest_t(const RbDeviceBvh & bvh, const RbRay & ray, float mini, float maxi, RbWork * work) {
	unsigned int mask = 0u;
	for(int r = 0; r < bvh.nroots; ++r) {
		const float4 r0 = rb_ldg4(bvh.roots + 2 * r);
		const float4 r1 = rb_ldg4(bvh.roots + 2 * r + 1);
		if(rb_box_test_t<REGULAR>(ray, r0.x, r0.y, r0.z, r1.x, r1.y, r1.z, mini, maxi)) mask |= 1u << r;
	}
	if(COUNT) work->roots += (unsigned int)bvh.nroots;
	return mask;
}
template<bool COUNT>
RB_D unsigned int rb_roots_pretest(const RbDeviceBvh & bvh, const RbRay & ray, float mini, float maxi, RbWork * work) {
	return ray.regular ? rb_roots_pretest_t<true, COUNT>(bvh, ray, mini, maxi, work) : rb_roots_pretest_t<false, COUNT>(bvh, ray, mini, maxi, work);
}

// The same set of entered roots from the tree over the root boxes (RbDeviceBvh::rootTree; rb_lbvh.cu "root tree"): a node
// carries the boxes of its two children, so one visit is two slab tests with the packed arithmetic of rb_node_test, and a
// child that is a root is tested against exactly its own box with the caller's (mini, maxi) — what the reference does for
// it. A union box the ray misses hides no root it would enter (the test is monotone in the box for regular rays). The
// ray's own stack column is empty at this point and serves the walk. Irregular rays (NaNs in the slab arithmetic) test every
// root with the literal form, like rb_roots_pretest.
template<bool COUNT>
RB_D void rb_roots_tree_pretest(RbTraversal & T, const RbDeviceBvh & bvh, const RbRay & ray, const RbStack & st, float mini, float maxi, RbWork * work) {
	unsigned int lo = 0u, hi = 0u;
	if(!ray.regular) {
		for(int r = 0; r < bvh.nroots; ++r) {
			const float4 r0 = rb_ldg4(bvh.roots + 2 * r);
			const float4 r1 = rb_ldg4(bvh.roots + 2 * r + 1);
			if(rb_box_test_t<false>(ray, r0.x, r0.y, r0.z, r1.x, r1.y, r1.z, mini, maxi)) {
				if(r < 32) lo |= 1u << r; else hi |= 1u << (r - 32);
			}
		}
		if(COUNT) work->roots += (unsigned int)bvh.nroots;
	} else {
		int cur = 0;
		for(;;) {
			float4 n0, n1, n2, n3;
			rb_ldg_node(bvh.rootTree + 4 * (size_t)cur, n0, n1, n2, n3);
			if(COUNT) work->roots += 2u;
			bool hitL, hitR;
			float dL, dR;
			rb_node_test<true>(ray, n0, n1, n2, mini, maxi, hitL, hitR, dL, dR);
			const int refL = __float_as_int(n3.x), refR = __float_as_int(n3.y);
			int next = RB_REF_NONE;
			if(hitL) {
				if(refL < 0) { const int r = ~refL; if(r < 32) lo |= 1u << r; else hi |= 1u << (r - 32); }
				else next = refL;
			}
			if(hitR) {
				if(refR < 0) { const int r = ~refR; if(r < 32) lo |= 1u << r; else hi |= 1u << (r - 32); }
				else {
					if(next != RB_REF_NONE) rb_push(T, st, next);
					next = refR;
				}
			}
			if(next == RB_REF_NONE) {
				if(rb_stack_empty(T, st)) break;
				next = rb_pop(T, st);
			}
			cur = next;
		}
	}
	T.root	 = int(lo);
	T.rootHi = int(hi);
}

// RB_ROOTS_TREE_BIG: the mask of entered roots in shared memory. `st` is the traversal stack proper; the mask words sit in
// the bvh.rootWords entries in front of it (the kernels shift the stack base by that many entries), private to the ray.
RB_D unsigned int rb_roots_word_address(const RbDeviceBvh & bvh, const RbStack & st, int word) {
	return st.base - (unsigned int)((bvh.rootWords - word) * st.strideBytes);
}
RB_D void rb_roots_set_bit(const RbDeviceBvh & bvh, const RbStack & st, int r) {
	const unsigned int a = rb_roots_word_address(bvh, st, r >> 5);
	rb_sts(a, rb_lds(a) | (1u << (r & 31)));
}
template<bool COUNT>
RB_D void rb_roots_tree_pretest_big(RbTraversal & T, const RbDeviceBvh & bvh, const RbRay & ray, const RbStack & st, float mini, float maxi, RbWork * work) {
	for(int w = 0; w < bvh.rootWords; ++w) rb_sts(rb_roots_word_address(bvh, st, w), 0u);
	if(!ray.regular) {
		for(int r = 0; r < bvh.nroots; ++r) {
			const float4 r0 = rb_ldg4(bvh.roots + 2 * r);
			const float4 r1 = rb_ldg4(bvh.roots + 2 * r + 1);
			if(rb_box_test_t<false>(ray, r0.x, r0.y, r0.z, r1.x, r1.y, r1.z, mini, maxi)) rb_roots_set_bit(bvh, st, r);
		}
		if(COUNT) work->roots += (unsigned int)bvh.nroots;
	} else {
		int cur = 0;
		for(;;) {
			float4 n0, n1, n2, n3;
			rb_ldg_node(bvh.rootTree + 4 * (size_t)cur, n0, n1, n2, n3);
			if(COUNT) work->roots += 2u;
			bool hitL, hitR;
			float dL, dR;
			rb_node_test<true>(ray, n0, n1, n2, mini, maxi, hitL, hitR, dL, dR);
			const int refL = __float_as_int(n3.x), refR = __float_as_int(n3.y);
			int next = RB_REF_NONE;
			if(hitL) {
				if(refL < 0) rb_roots_set_bit(bvh, st, ~refL);
				else next = refL;
			}
			if(hitR) {
				if(refR < 0) rb_roots_set_bit(bvh, st, ~refR);
				else {
					if(next != RB_REF_NONE) rb_push(T, st, next);
					next = refR;
				}
			}
			if(next == RB_REF_NONE) {
				if(rb_stack_empty(T, st)) break;
				next = rb_pop(T, st);
			}
			cur = next;
		}
	}
	T.root = bvh.rootWords - 1; // highest word that may still hold a root
}
// Next entered root, last mesh first: the highest bit of the highest non-empty word.
RB_D int rb_next_root_big(RbTraversal & T, const RbDeviceBvh & bvh, const RbStack & st) {
	for(int w = T.root; w >= 0; --w) {
		const unsigned int a	= rb_roots_word_address(bvh, st, w);
		const unsigned int word = rb_lds(a);
		if(word != 0u) {
			const int b = 31 - __clz(word);
			rb_sts(a, word & ~(1u << b));
			T.root = w;
			return __float_as_int(rb_ldg4(bvh.roots + 2 * (32 * w + b)).w);
		}
	}
	T.root = -1;
	return RB_REF_NONE;
}

template<bool REGULAR, bool COUNT, int ROOTS>
RB_D int rb_next_root(RbTraversal & T, const RbDeviceBvh & bvh, const RbRay & ray, float maxi, RbWork * work) {
	if(ROOTS == RB_ROOTS_MASK) {
		const unsigned int mask = (unsigned int)T.root;
		if(mask == 0u) return RB_REF_NONE;
		const int r = 31 - __clz(mask); // last mesh first
		T.root		= int(mask & ~(1u << r));
		return __float_as_int(rb_ldg4(bvh.roots + 2 * r).w);
	}
	if(ROOTS == RB_ROOTS_TREE) {
		const unsigned int hi = (unsigned int)T.rootHi, lo = (unsigned int)T.root;
		int r;
		if(hi != 0u) {
			r		 = 31 - __clz(hi);
			T.rootHi = int(hi & ~(1u << r));
			r += 32;
		} else if(lo != 0u) {
			r	   = 31 - __clz(lo);
			T.root = int(lo & ~(1u << r));
		} else {
			return RB_REF_NONE;
		}
		return __float_as_int(rb_ldg4(bvh.roots + 2 * r).w);
	}
	const bool BLOCKS = ROOTS == RB_ROOTS_BLOCKS;
	while(T.root > 0) {
		const int r = T.root - 1;
		if(BLOCKS && REGULAR && bvh.rootLevels > 0) {
			int k		 = min(__ffs(r + 1) - 1, bvh.rootLevels); // r + 1 is a multiple of 2^k
			bool skipped = false;
			for(; k >= 1; --k) {
				const float4 * b = bvh.rootBlocks + 2 * size_t(bvh.rootLevelOffset[k] + (r >> k));
				const float4 b0 = rb_ldg4(b), b1 = rb_ldg4(b + 1);
				if(COUNT) work->roots++;
				if(!rb_box_test_t<true>(ray, b0.x, b0.y, b0.z, b1.x, b1.y, b1.z, T.mini, maxi)) {
					T.root	= (r >> k) << k;
					skipped = true;
					break;
				}
			}
			if(skipped) continue;
		}
		T.root = r;
		const float4 r0 = rb_ldg4(bvh.roots + 2 * r);
		const float4 r1 = rb_ldg4(bvh.roots + 2 * r + 1);
		if(COUNT) work->roots++;
		// (A mesh without triangles is given a leaf of one degenerate triangle by rb_host_flatten, so its root never reads as
		// "no roots left". Testing the reference for RB_REF_NONE here instead costs every kernel 21 % — measured, same box:
		// the register allocation of the hot loop is that sensitive to this loop.)
		if(rb_box_test_t<REGULAR>(ray, r0.x, r0.y, r0.z, r1.x, r1.y, r1.z, T.mini, maxi)) return __float_as_int(r0.w);
	}
	return RB_REF_NONE;
}

// Returns true when the traversal is finished (for ANY: also as soon as a triangle is accepted).
template<bool ANY, bool COUNT, bool REGULAR>
RB_D bool rb_step_exact(RbTraversal & T, const RbDeviceBvh & bvh, const RbRay & ray, const RbStack & st, RbWork * work) {
	int cur = T.cur;
	if(cur == RB_REF_NONE) {
		if(!rb_stack_empty(T, st)) {
			cur = rb_pop(T, st);
		} else {
			cur = rb_next_root<REGULAR, COUNT, false>(T, bvh, ray, T.maxi0, work);
			if(cur == RB_REF_NONE) {
				return true;
			}
		}
	}
	// Descend while on interior nodes.
	while(cur >= 0 && cur != RB_REF_NONE) {
		const float4 * n = bvh.nodes + 4 * (size_t)cur;
		float4 n0, n1, n2, n3;
		rb_ldg_node(n, n0, n1, n2, n3);
		if(COUNT) work->interior++;
		bool hitL, hitR;
		float dL, dR;
		rb_node_test<REGULAR>(ray, n0, n1, n2, T.mini, T.maxi, hitL, hitR, dL, dR);
		const int refL = __float_as_int(n3.x), refR = __float_as_int(n3.y);
		if(hitR) {
			if(hitL) {
				rb_push(T, st, refL);
				if(COUNT) work->maxStack = max(work->maxStack, rb_stack_size(T, st));
			}
			cur = refR;
		} else if(hitL) {
			cur = refL;
		} else if(!rb_stack_empty(T, st)) {
			cur = rb_pop(T, st);
		} else {
			cur = RB_REF_NONE;
		}
	}
	if(cur < 0) {
		const int code	= ~cur;
		const int first = code >> RB_LEAF_BITS;
		const int count = (code & (RB_LEAF_MAX_TRIS - 1)) + 1;
		for(int k = 0; k < count; ++k) {
			const float4 * tp = bvh.tris + 3 * (size_t)(first + k);
			const float4 t0 = rb_ldg4(tp), t1 = rb_ldg4(tp + 1), t2 = rb_ldg4(tp + 2);
			if(COUNT) work->tris++;
			float t, u, v;
			if(rb_tri_test(ray, V3{t0.x, t0.y, t0.z}, V3{t1.x, t1.y, t1.z}, V3{t2.x, t2.y, t2.z}, T.mini, T.maxi, t, u, v)) {
				if(ANY) {
					T.best.t = t; T.best.u = u; T.best.v = v; T.best.tri = first + k;
					T.cur = RB_REF_NONE;
					return true;
				}
				if(t < T.best.t) {
					T.best.t = t; T.best.u = u; T.best.v = v; T.best.tri = first + k;
					T.maxi = t;
				}
			}
		}
		cur = RB_REF_NONE;
	}
	T.cur = cur;
	return false;
}


// Near-child-first traversal of the same tree (RB_TRACE_ORDERED): both children are box-tested exactly as above, the
// one whose slab entry distance is smaller is visited first, so `maxi` shrinks sooner and far subtrees are culled.
// All roots are still considered (in reverse order). Same triangle test; the closest hit is the same triangle except
// where two triangles are hit at exactly the same distance or where the reference's non-watertight box test makes
// the outcome order-dependent ("ordered mode").
template<bool REGULAR>
RB_D bool rb_box_test_dist(const RbRay & ray, float bminx, float bminy, float bminz, float bmaxx, float bmaxy, float bmaxz, float mini, float maxi, float & entry) {
	const float minRx = (bminx - ray.pos.x) * ray.inv.x;
	const float minRy = (bminy - ray.pos.y) * ray.inv.y;
	const float minRz = (bminz - ray.pos.z) * ray.inv.z;
	const float maxRx = (bmaxx - ray.pos.x) * ray.inv.x;
	const float maxRy = (bmaxy - ray.pos.y) * ray.inv.y;
	const float maxRz = (bmaxz - ray.pos.z) * ray.inv.z;
	if(REGULAR) {
		const float closest	 = fmaxf(fmaxf(fminf(minRx, maxRx), fminf(minRy, maxRy)), fmaxf(fminf(minRz, maxRz), mini));
		const float furthest = fminf(fminf(fmaxf(minRx, maxRx), fmaxf(minRy, maxRy)), fminf(fmaxf(minRz, maxRz), maxi));
		entry = closest;
		return closest <= furthest;
	}
	const float minFx = glm_min(minRx, maxRx), minFy = glm_min(minRy, maxRy), minFz = glm_min(minRz, maxRz);
	const float maxFx = glm_max(minRx, maxRx), maxFy = glm_max(minRy, maxRy), maxFz = glm_max(minRz, maxRz);
	const float closest	 = std_max(minFx, std_max(minFy, minFz));
	const float furthest = std_min(maxFx, std_min(maxFy, maxFz));
	entry = std_max(closest, mini);
	return entry <= std_min(furthest, maxi);
}

template<bool ANY, bool COUNT, bool REGULAR>
RB_D bool rb_step_ordered(RbTraversal & T, const RbDeviceBvh & bvh, const RbRay & ray, const RbStack & st, RbWork * work) {
	int cur = T.cur;
	if(cur == RB_REF_NONE) {
		if(!rb_stack_empty(T, st)) {
			cur = rb_pop(T, st);
		} else {
			cur = rb_next_root<REGULAR, COUNT, false>(T, bvh, ray, T.maxi, work);
			if(cur == RB_REF_NONE) {
				return true;
			}
		}
	}
	while(cur >= 0 && cur != RB_REF_NONE) {
		const float4 * n = bvh.nodes + 4 * (size_t)cur;
		float4 n0, n1, n2, n3;
		rb_ldg_node(n, n0, n1, n2, n3);
		if(COUNT) work->interior++;
		float dL, dR;
		bool hitL, hitR;
		rb_node_test<REGULAR>(ray, n0, n1, n2, T.mini, T.maxi, hitL, hitR, dL, dR);
		const int refL = __float_as_int(n3.x), refR = __float_as_int(n3.y);
		if(hitL && hitR) {
			// Visit the nearer child first; on equal entry distance keep the reference's order (right first).
			const bool leftFirst = dL < dR;
			rb_push(T, st, leftFirst ? refR : refL);
			if(COUNT) work->maxStack = max(work->maxStack, rb_stack_size(T, st));
			cur = leftFirst ? refL : refR;
		} else if(hitR) {
			cur = refR;
		} else if(hitL) {
			cur = refL;
		} else if(!rb_stack_empty(T, st)) {
			cur = rb_pop(T, st);
		} else {
			cur = RB_REF_NONE;
		}
	}
	if(cur < 0) {
		const int code	= ~cur;
		const int first = code >> RB_LEAF_BITS;
		const int count = (code & (RB_LEAF_MAX_TRIS - 1)) + 1;
		for(int k = 0; k < count; ++k) {
			const float4 * tp = bvh.tris + 3 * (size_t)(first + k);
			const float4 t0 = rb_ldg4(tp), t1 = rb_ldg4(tp + 1), t2 = rb_ldg4(tp + 2);
			if(COUNT) work->tris++;
			float t, u, v;
			if(rb_tri_test(ray, V3{t0.x, t0.y, t0.z}, V3{t1.x, t1.y, t1.z}, V3{t2.x, t2.y, t2.z}, T.mini, T.maxi, t, u, v)) {
				if(ANY) {
					T.best.t = t; T.best.u = u; T.best.v = v; T.best.tri = first + k;
					T.cur = RB_REF_NONE;
					return true;
				}
				if(t < T.best.t) {
					T.best.t = t; T.best.u = u; T.best.v = v; T.best.tri = first + k;
					T.maxi = t;
				}
			}
		}
		cur = RB_REF_NONE;
	}
	T.cur = cur;
	return false;
}


// Both children of an interior node at once (two Intersection::box tests). For regular rays the twelve subtractions and
// twelve multiplications run as six FADD2 + six FMUL2: the node layout delivers exactly the pairs
// (L.min.x,L.min.y) (L.min.z,L.max.x) (L.max.y,L.max.z) (R.min.x,R.min.y) (R.min.z,R.max.x) (R.max.y,R.max.z).
// dL/dR receive max(closest, mini), the slab entry distance used by the ordered traversal.
template<bool REGULAR>
RB_D void rb_node_test(const RbRay & ray, const float4 & n0, const float4 & n1, const float4 & n2, float mini, float maxi, bool & hitL, bool & hitR, float & dL, float & dR) {
	if(REGULAR) {
		float lminx, lminy, lminz, lmaxx, lmaxy, lmaxz, rminx, rminy, rminz, rmaxx, rmaxy, rmaxz;
		rb_unpack2(rb_submul2(rb_pack2(n0.x, n0.y), ray.pxy, ray.ixy), lminx, lminy);
		rb_unpack2(rb_submul2(rb_pack2(n0.z, n0.w), ray.pzx, ray.izx), lminz, lmaxx);
		rb_unpack2(rb_submul2(rb_pack2(n1.x, n1.y), ray.pyz, ray.iyz), lmaxy, lmaxz);
		rb_unpack2(rb_submul2(rb_pack2(n1.z, n1.w), ray.pxy, ray.ixy), rminx, rminy);
		rb_unpack2(rb_submul2(rb_pack2(n2.x, n2.y), ray.pzx, ray.izx), rminz, rmaxx);
		rb_unpack2(rb_submul2(rb_pack2(n2.z, n2.w), ray.pyz, ray.iyz), rmaxy, rmaxz);
		dL = fmaxf(fmaxf(fminf(lminx, lmaxx), fminf(lminy, lmaxy)), fmaxf(fminf(lminz, lmaxz), mini));
		const float fL = fminf(fminf(fmaxf(lminx, lmaxx), fmaxf(lminy, lmaxy)), fminf(fmaxf(lminz, lmaxz), maxi));
		dR = fmaxf(fmaxf(fminf(rminx, rmaxx), fminf(rminy, rmaxy)), fmaxf(fminf(rminz, rmaxz), mini));
		const float fR = fminf(fminf(fmaxf(rminx, rmaxx), fmaxf(rminy, rmaxy)), fminf(fmaxf(rminz, rmaxz), maxi));
		hitL = dL <= fL;
		hitR = dR <= fR;
	} else {
		hitL = rb_box_test_dist<false>(ray, n0.x, n0.y, n0.z, n0.w, n1.x, n1.y, mini, maxi, dL);
		hitR = rb_box_test_dist<false>(ray, n1.z, n1.w, n2.x, n2.y, n2.z, n2.w, mini, maxi, dR);
	}
}

// ---------------------------------------------------------------------------------------------- phase-split steps
// The same traversal cut into its two kinds of work so that a warp can decide, by vote, which kind to execute next:
//   node phase : take the pending node if none is current (stack, else next accepted root), then — if it is an interior
//                node — box-test its two children and choose the next node;
//   leaf phase : test the triangles of the current leaf.
// Order of operations per ray is unchanged (a ray never starts a node before finishing its pending leaf), so results
// are identical to rb_step_*; only the interleaving BETWEEN rays of a warp differs.
RB_D bool rb_wants_leaf(const RbTraversal & T) { return T.cur < 0; }

// Returns true when the traversal has finished (no node left).
template<bool COUNT, bool ORDERED, bool REGULAR, int ROOTS>
RB_D bool rb_phase_node_t(RbTraversal & T, const RbDeviceBvh & bvh, const RbRay & ray, const RbStack & st, RbWork * work) {
	int cur = T.cur;
	if(cur == RB_REF_NONE) {
		if(!rb_stack_empty(T, st)) {
			cur = rb_pop(T, st);
		} else {
			if(ROOTS == RB_ROOTS_TREE_BIG) cur = rb_next_root_big(T, bvh, st);
			else cur = rb_next_root<REGULAR, COUNT, ROOTS>(T, bvh, ray, ORDERED ? T.maxi : T.maxi0, work);
			if(cur == RB_REF_NONE) return true;
		}
	}
	if(cur >= 0) {
		const float4 * n = bvh.nodes + 4 * (size_t)cur;
		float4 n0, n1, n2, n3;
		rb_ldg_node(n, n0, n1, n2, n3);
		if(COUNT) work->interior++;
		const int refL = __float_as_int(n3.x), refR = __float_as_int(n3.y);
		if(ORDERED) {
			float dL, dR;
			bool hitL, hitR;
			rb_node_test<REGULAR>(ray, n0, n1, n2, T.mini, T.maxi, hitL, hitR, dL, dR);
			if(hitL && hitR) {
				const bool leftFirst = dL < dR;
				rb_push(T, st, leftFirst ? refR : refL);
				cur = leftFirst ? refL : refR;
			} else {
				cur = hitR ? refR : (hitL ? refL : RB_REF_NONE);
			}
		} else {
			bool hitL, hitR;
			float dL, dR;
			rb_node_test<REGULAR>(ray, n0, n1, n2, T.mini, T.maxi, hitL, hitR, dL, dR);
			if(hitR) {
				if(hitL) rb_push(T, st, refL);
				cur = refR;
			} else {
				cur = hitL ? refL : RB_REF_NONE;
			}
		}
		if(COUNT) work->maxStack = max(work->maxStack, rb_stack_size(T, st));
	}
	T.cur = cur;
	return false;
}

// One triangle of the current leaf per call: a leaf with several triangles stays "current" (with its first triangle
// consumed) and is continued by the next leaf phase, so lanes with one triangle never wait for lanes with two, and the
// second triangles are batched with other lanes' first ones. Returns true when an ANY query has found its hit.
template<bool ANY, bool COUNT>
RB_D bool rb_phase_leaf(RbTraversal & T, const RbDeviceBvh & bvh, const RbRay & ray, RbWork * work) {
	const int code	= ~T.cur;
	const int first = code >> RB_LEAF_BITS;
	const int rest	= code & (RB_LEAF_MAX_TRIS - 1); // triangles left after this one
	T.cur = rest > 0 ? ~(((first + 1) << RB_LEAF_BITS) | (rest - 1)) : RB_REF_NONE;
	const float4 * tp = bvh.tris + 3 * (size_t)first;
	const float4 t0 = rb_ldg4(tp), t1 = rb_ldg4(tp + 1), t2 = rb_ldg4(tp + 2);
	if(COUNT) work->tris++;
	float t, u, v;
	if(rb_tri_test(ray, V3{t0.x, t0.y, t0.z}, V3{t1.x, t1.y, t1.z}, V3{t2.x, t2.y, t2.z}, T.mini, T.maxi, t, u, v)) {
		if(ANY) {
			T.best.t = t; T.best.u = u; T.best.v = v; T.best.tri = first;
			return true;
		}
		if(t < T.best.t) {
			T.best.t = t; T.best.u = u; T.best.v = v; T.best.tri = first;
			T.maxi = t;
		}
	}
	return false;
}

template<bool COUNT, bool ORDERED, int ROOTS = RB_ROOTS_LAZY>
RB_D bool rb_phase_node(RbTraversal & T, const RbDeviceBvh & bvh, const RbRay & ray, const RbStack & st, RbWork * work) {
	if(ray.regular) return rb_phase_node_t<COUNT, ORDERED, true, ROOTS>(T, bvh, ray, st, work);
	return rb_phase_node_t<COUNT, ORDERED, false, ROOTS>(T, bvh, ray, st, work);
}

// One traversal step. The NaN-faithful box test is only compiled into the copy taken by irregular rays (a direction
// component that is zero or denormal, or a non-finite origin); every other ray runs the FMNMX form.
template<bool ANY, bool COUNT, bool ORDERED>
RB_D bool rb_step(RbTraversal & T, const RbDeviceBvh & bvh, const RbRay & ray, const RbStack & st, RbWork * work) {
	if(ray.regular) {
		if(ORDERED) return rb_step_ordered<ANY, COUNT, true>(T, bvh, ray, st, work);
		return rb_step_exact<ANY, COUNT, true>(T, bvh, ray, st, work);
	}
	if(ORDERED) return rb_step_ordered<ANY, COUNT, false>(T, bvh, ray, st, work);
	return rb_step_exact<ANY, COUNT, false>(T, bvh, ray, st, work);
}

template<bool ANY, bool COUNT>
RB_D bool rb_trace_exact(const RbDeviceBvh & bvh, const RbRay & ray, float mini, float maxi, const RbStack & st, RbHitRec & best, RbWork * work) {
	RbTraversal T;
	rb_traversal_init(T, bvh, mini, maxi, st);
	while(!rb_step<ANY, COUNT, false>(T, bvh, ray, st, work)) {
	}
	best = T.best;
	return best.tri >= 0;
}

template<bool ANY, bool COUNT>
RB_D bool rb_trace_ordered(const RbDeviceBvh & bvh, const RbRay & ray, float mini, float maxi, const RbStack & st, RbHitRec & best, RbWork * work) {
	RbTraversal T;
	rb_traversal_init(T, bvh, mini, maxi, st);
	while(!rb_step<ANY, COUNT, true>(T, bvh, ray, st, work)) {
	}
	best = T.best;
	return best.tri >= 0;
}
