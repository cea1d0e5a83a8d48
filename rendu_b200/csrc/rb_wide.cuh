// rb_wide.cuh — traversal of the eight-wide quantised tree (layout and rationale in rb_wide.h).
//
// State of a ray: the node it is in (first interior child, first triangle, the eight meta bytes) and a 24-bit list of
// pending items, three bits per child slot: bit 3s of an interior child that was hit, bits 3s..3s+count-1 of a leaf
// child that was hit. Items are consumed lowest bit first, which is the reference's visiting order. Entering an interior
// child parks what is left of the current node on the stack (five words per entry: re-reading the node's header on the
// way back would cost a divergent global load — one L1 wavefront per lane — where a shared-memory stack access costs
// one per warp-instruction).
//
// Box test of child i on axis a, for a regular ray: the byte q is spliced into the mantissa of 1.0f (PRMT), giving
// f = 1 + q * 2^-15, and t = fma(f, A, B) with A = 2^15 * scale_a / dir_a and B = (p_a - o_a) / dir_a - A, so that
// t = (p_a + q * scale_a - o_a) / dir_a up to 2^-7 of a quantum — inside the RB_WIDE_MARGIN the builder leaves. The sign
// of dir_a selects once per node which of the lo / hi words is the near plane; no per-child min/max per axis is needed.
#pragma once
#include "rb_trace.cuh"
#include "rb_wide.h"

struct RbWideTraversal {
	RbHitRec best; // best.tri: index in hierarchy order (word 11 of the wide triangle record)
	float mini, maxi, maxi0;
	unsigned int childBase, triBase, metaLo, metaHi;
	unsigned int pending; // 24 bits, three per slot (OCT: per visiting position, see rb_wide_permute)
	unsigned int flip;	  // OCT: visiting position of a slot = slot ^ flip (bit a set where the ray runs towards -a)
	int root;
	unsigned int top; // shared-memory address of the next free stack entry
};

RB_D void rb_wide_init(RbWideTraversal & T, const RbDeviceBvh & bvh, float mini, float maxi, const RbStack & st) {
	T.best.t   = FLT_MAX;
	T.best.u   = 0.0f;
	T.best.v   = 0.0f;
	T.best.tri = -1;
	T.mini	   = mini;
	T.maxi	   = maxi;
	T.maxi0	   = maxi;
	T.pending  = 0u;
	T.flip	   = 0u;
	T.top	   = st.base;
	T.root	   = bvh.nroots;
}

// Stack entries are five words; word w of entry e of a thread lives at base + (5 e + w) * strideBytes.
RB_D void rb_wide_push(RbWideTraversal & T, const RbStack & st) {
	unsigned int a = T.top;
	rb_sts(a, T.childBase);
	a += st.strideBytes;
	rb_sts(a, T.triBase);
	a += st.strideBytes;
	rb_sts(a, T.metaLo);
	a += st.strideBytes;
	rb_sts(a, T.metaHi);
	a += st.strideBytes;
	rb_sts(a, T.pending);
	T.top = a + st.strideBytes;
}
RB_D void rb_wide_pop(RbWideTraversal & T, const RbStack & st) {
	unsigned int a = T.top - st.strideBytes;
	T.pending = rb_lds(a);
	a -= st.strideBytes;
	T.metaHi = rb_lds(a);
	a -= st.strideBytes;
	T.metaLo = rb_lds(a);
	a -= st.strideBytes;
	T.triBase = rb_lds(a);
	a -= st.strideBytes;
	T.childBase = rb_lds(a);
	T.top = a;
}

// 1 + q * 2^-15 for byte `k` of word `w` (k is a compile-time constant: one PRMT). `one` is 0x3f800000 read from the
// kernel arguments: PRMT takes one immediate, and if the compiler knows the constant it makes THAT the immediate and
// spends a MOV on every selector (48 per node).
#define RB_WIDE_F(w, k) __uint_as_float(__byte_perm((w), one, 0x7604u | ((k) << 4)))

// Four 3-bit fields sitting in the low bits of the four bytes of c, packed into 12 contiguous bits.
RB_D unsigned int rb_wide_pack3(unsigned int c) {
	return (c & 0x7u) | ((c >> 5) & 0x38u) | ((c >> 10) & 0x1c0u) | ((c >> 15) & 0xe00u);
}

// Slots by direction (RbWideBvh::nodesByOctant): a ray visits slot s at position s ^ flip. The pending list is kept per
// POSITION, so that its lowest bit is still the next item: the three-bit groups are permuted by the XOR, i.e. swapped in
// neighbours, pairs and halves as the bits of `flip` say.
RB_D unsigned int rb_wide_permute(unsigned int x, unsigned int flip) {
	if(flip & 1u) x = ((x & 0x1c71c7u) << 3) | ((x >> 3) & 0x1c71c7u);
	if(flip & 2u) x = ((x & 0x03f03fu) << 6) | ((x >> 6) & 0x03f03fu);
	if(flip & 4u) x = ((x & 0x000fffu) << 12) | (x >> 12);
	return x;
}
RB_D unsigned int rb_wide_flip(const RbRay & ray) {
	return (ray.inv.x < 0.0f ? 1u : 0u) | (ray.inv.y < 0.0f ? 2u : 0u) | (ray.inv.z < 0.0f ? 4u : 0u);
}

// Loads node `n`, box-tests its eight children against (mini, maxi) and makes it the current node of T.
template<bool OCT>
RB_D void rb_wide_enter(RbWideTraversal & T, const RbDeviceWide & wide, const RbRay & ray, int n) {
	const uint4 * p = (OCT ? wide.nodesByOctant : wide.nodes) + 5 * (size_t)n;
	const uint4 n0 = __ldg(p), n1 = __ldg(p + 1), n2 = __ldg(p + 2), n3 = __ldg(p + 3), n4 = __ldg(p + 4);
	T.childBase = n1.x;
	T.triBase	= n1.y;
	T.metaLo	= n1.z;
	T.metaHi	= n1.w;
	// Per axis: A = 2^15 * 2^e / dir, B = (p - o) / dir - A. 2^(e+15) is assembled from the biased exponent.
	const float sx = __uint_as_float(((n0.w & 0xffu) + 15u) << 23);
	const float sy = __uint_as_float((((n0.w >> 8) & 0xffu) + 15u) << 23);
	const float sz = __uint_as_float((((n0.w >> 16) & 0xffu) + 15u) << 23);
	const float Ax = sx * ray.inv.x, Ay = sy * ray.inv.y, Az = sz * ray.inv.z;
	const float Bx = (__uint_as_float(n0.x) - ray.pos.x) * ray.inv.x - Ax;
	const float By = (__uint_as_float(n0.y) - ray.pos.y) * ray.inv.y - Ay;
	const float Bz = (__uint_as_float(n0.z) - ray.pos.z) * ray.inv.z - Az;
	// Near / far plane words per axis (slots 0-3 and 4-7), chosen by the sign of the direction.
	const bool nx = ray.inv.x < 0.0f, ny = ray.inv.y < 0.0f, nz = ray.inv.z < 0.0f;
	const unsigned int nearX0 = nx ? n3.z : n2.x, nearX1 = nx ? n3.w : n2.y, farX0 = nx ? n2.x : n3.z, farX1 = nx ? n2.y : n3.w;
	const unsigned int nearY0 = ny ? n4.x : n2.z, nearY1 = ny ? n4.y : n2.w, farY0 = ny ? n2.z : n4.x, farY1 = ny ? n2.w : n4.y;
	const unsigned int nearZ0 = nz ? n4.z : n3.x, nearZ1 = nz ? n4.w : n3.y, farZ0 = nz ? n3.x : n4.z, farZ1 = nz ? n3.y : n4.w;
	unsigned int hit = 0u;
	const unsigned int one = wide.oneBits;
	// Two children at a time: the six plane distances of each are three two-wide FMAs (fma.rn.f32x2, sm_100) per pair
	// of planes — each half rounds exactly like the scalar FFMA.
	const unsigned long long Ax2 = rb_pack2(Ax, Ax), Ay2 = rb_pack2(Ay, Ay), Az2 = rb_pack2(Az, Az);
	const unsigned long long Bx2 = rb_pack2(Bx, Bx), By2 = rb_pack2(By, By), Bz2 = rb_pack2(Bz, Bz);
#define RB_WIDE_PAIR(W, k, s)                                                                                               \
	{                                                                                                                       \
		float tnx0, tnx1, tny0, tny1, tnz0, tnz1, tfx0, tfx1, tfy0, tfy1, tfz0, tfz1;                                       \
		rb_unpack2(rb_fma2(rb_pack2(RB_WIDE_F(nearX##W, k), RB_WIDE_F(nearX##W, k + 1)), Ax2, Bx2), tnx0, tnx1);            \
		rb_unpack2(rb_fma2(rb_pack2(RB_WIDE_F(nearY##W, k), RB_WIDE_F(nearY##W, k + 1)), Ay2, By2), tny0, tny1);            \
		rb_unpack2(rb_fma2(rb_pack2(RB_WIDE_F(nearZ##W, k), RB_WIDE_F(nearZ##W, k + 1)), Az2, Bz2), tnz0, tnz1);            \
		rb_unpack2(rb_fma2(rb_pack2(RB_WIDE_F(farX##W, k), RB_WIDE_F(farX##W, k + 1)), Ax2, Bx2), tfx0, tfx1);              \
		rb_unpack2(rb_fma2(rb_pack2(RB_WIDE_F(farY##W, k), RB_WIDE_F(farY##W, k + 1)), Ay2, By2), tfy0, tfy1);              \
		rb_unpack2(rb_fma2(rb_pack2(RB_WIDE_F(farZ##W, k), RB_WIDE_F(farZ##W, k + 1)), Az2, Bz2), tfz0, tfz1);              \
		const float tn0 = fmaxf(fmaxf(tnx0, tny0), fmaxf(tnz0, T.mini)), tf0 = fminf(fminf(tfx0, tfy0), fminf(tfz0, T.maxi)); \
		const float tn1 = fmaxf(fmaxf(tnx1, tny1), fmaxf(tnz1, T.mini)), tf1 = fminf(fminf(tfx1, tfy1), fminf(tfz1, T.maxi)); \
		if(tn0 <= tf0) hit |= 7u << (3 * (s));                                                                              \
		if(tn1 <= tf1) hit |= 7u << (3 * (s) + 3);                                                                          \
	}
	RB_WIDE_PAIR(0, 0, 0)
	RB_WIDE_PAIR(0, 2, 2)
	RB_WIDE_PAIR(1, 0, 4)
	RB_WIDE_PAIR(1, 2, 6)
#undef RB_WIDE_PAIR
	// The items each slot really has: the unary counts (top three bits of every meta byte) packed three bits per slot.
	const unsigned int cLo = (n1.z >> 5) & 0x07070707u, cHi = (n1.w >> 5) & 0x07070707u;
	T.pending = hit & (rb_wide_pack3(cLo) | (rb_wide_pack3(cHi) << 12));
	if(OCT) T.pending = rb_wide_permute(T.pending, T.flip);
}

// One step: consume the next pending item — enter an interior child, or test one triangle of a leaf child. Returns
// true when the ray has finished (ANY: also as soon as a triangle is accepted). Regular rays only (the caller sends
// rays that can produce NaNs in the slab test through the binary tree, whose box test reproduces them).
template<bool ANY, bool OCT = false>
RB_D bool rb_wide_step(RbWideTraversal & T, const RbDeviceBvh & bvh, const RbDeviceWide & wide, const RbRay & ray, const RbStack & st) {
	int enter = -1; // node to enter in this step (one call site of rb_wide_enter: it is 180 instructions)
	if(T.pending == 0u) {
		if(T.top != st.base) {
			rb_wide_pop(T, st);
		} else {
			// Next mesh whose root box the ray enters, last mesh first, tested with the caller's maxi (Raycaster.cpp:149-155).
			RbTraversal R;
			R.mini = T.mini;
			R.root = T.root;
			const int ref = rb_next_root<true, false, false>(R, bvh, ray, T.maxi0, nullptr);
			T.root		  = R.root;
			if(ref == RB_REF_NONE) return true;
			enter = __ldg(wide.rootNode + T.root);
		}
	}
	if(enter < 0) {
		const unsigned int bit = __ffs(T.pending) - 1;
		T.pending &= T.pending - 1u;
		const unsigned int spot = (bit * 11u) >> 5; // bit / 3 for bit < 24: the visiting position
		const unsigned int slot = OCT ? (spot ^ T.flip) : spot;
		const unsigned int low	= __byte_perm(T.metaLo, T.metaHi, slot) & 31u;
		if(low >= 24u) {
			enter = int(T.childBase + low - 24u);
			if(T.pending != 0u) rb_wide_push(T, st);
		} else {
			const int tri	  = int(T.triBase + low + (bit - 3u * spot));
			const float4 * tp = wide.tris + 3 * (size_t)tri;
			const float4 t0 = rb_ldg4(tp), t1 = rb_ldg4(tp + 1), t2 = rb_ldg4(tp + 2);
			float t, u, v;
			if(rb_tri_test(ray, V3{t0.x, t0.y, t0.z}, V3{t1.x, t1.y, t1.z}, V3{t2.x, t2.y, t2.z}, T.mini, T.maxi, t, u, v)) {
				if(ANY) {
					T.best.t = t; T.best.u = u; T.best.v = v; T.best.tri = __float_as_int(t2.w);
					return true;
				}
				if(t < T.best.t) {
					T.best.t = t; T.best.u = u; T.best.v = v; T.best.tri = __float_as_int(t2.w);
					T.maxi = t;
				}
			}
			return false;
		}
	}
	rb_wide_enter<OCT>(T, wide, ray, enter);
	return false;
}

// ---------------------------------------------------------------------------------------------- phase-split steps
// The same traversal cut into its two kinds of work (as rb_phase_node / rb_phase_leaf do for the binary tree), so that a
// warp can vote which kind to run: entering a node is ≈180 instructions, a triangle test ≈80, and a warp that mixes them
// pays for both in every iteration. Between steps a ray's pending list is kept non-empty whenever its stack is not
// (the pop is taken eagerly), so "what does this lane want next" is a look at the lowest pending bit.
RB_D void rb_wide_settle(RbWideTraversal & T, const RbStack & st) {
	if(T.pending == 0u && T.top != st.base) rb_wide_pop(T, st);
}

// True when the next item of the ray is a triangle (false: an interior child, or the next root).
template<bool OCT = false>
RB_D bool rb_wide_wants_triangle(const RbWideTraversal & T) {
	if(T.pending == 0u) return false;
	const unsigned int bit	= __ffs(T.pending) - 1;
	const unsigned int spot = (bit * 11u) >> 5;
	return (__byte_perm(T.metaLo, T.metaHi, OCT ? (spot ^ T.flip) : spot) & 31u) < 24u;
}

// Node phase: enter the next interior child, or — with nothing pending — the next root. Returns true when the ray has
// finished (no root left).
template<bool OCT = false>
RB_D bool rb_wide_phase_node(RbWideTraversal & T, const RbDeviceBvh & bvh, const RbDeviceWide & wide, const RbRay & ray, const RbStack & st) {
	int enter;
	if(T.pending == 0u) {
		RbTraversal R;
		R.mini = T.mini;
		R.root = T.root;
		const int ref = rb_next_root<true, false, false>(R, bvh, ray, T.maxi0, nullptr);
		T.root		  = R.root;
		if(ref == RB_REF_NONE) return true;
		enter = __ldg(wide.rootNode + T.root);
	} else {
		const unsigned int bit = __ffs(T.pending) - 1;
		T.pending &= T.pending - 1u;
		const unsigned int spot = (bit * 11u) >> 5;
		enter = int(T.childBase + (__byte_perm(T.metaLo, T.metaHi, OCT ? (spot ^ T.flip) : spot) & 31u) - 24u);
		if(T.pending != 0u) rb_wide_push(T, st);
	}
	rb_wide_enter<OCT>(T, wide, ray, enter);
	rb_wide_settle(T, st);
	return false;
}

// Leaf phase: one triangle. Returns true when an ANY query has found its hit.
template<bool ANY, bool OCT = false>
RB_D bool rb_wide_phase_triangle(RbWideTraversal & T, const RbDeviceWide & wide, const RbRay & ray, const RbStack & st) {
	const unsigned int bit = __ffs(T.pending) - 1;
	T.pending &= T.pending - 1u;
	const unsigned int spot = (bit * 11u) >> 5;
	const unsigned int low	= __byte_perm(T.metaLo, T.metaHi, OCT ? (spot ^ T.flip) : spot) & 31u;
	const int tri			= int(T.triBase + low + (bit - 3u * spot));
	const float4 * tp		= wide.tris + 3 * (size_t)tri;
	const float4 t0 = rb_ldg4(tp), t1 = rb_ldg4(tp + 1), t2 = rb_ldg4(tp + 2);
	float t, u, v;
	if(rb_tri_test(ray, V3{t0.x, t0.y, t0.z}, V3{t1.x, t1.y, t1.z}, V3{t2.x, t2.y, t2.z}, T.mini, T.maxi, t, u, v)) {
		if(ANY) {
			T.best.t = t; T.best.u = u; T.best.v = v; T.best.tri = __float_as_int(t2.w);
			return true;
		}
		if(t < T.best.t) {
			T.best.t = t; T.best.u = u; T.best.v = v; T.best.tri = __float_as_int(t2.w);
			T.maxi = t;
		}
	}
	rb_wide_settle(T, st);
	return false;
}
