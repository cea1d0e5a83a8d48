// rb_wide.h — eight-wide hierarchy with quantised child boxes (RB_TRACE_WIDE; experiment: parity-checked, not yet timed).
//
// Why: the traversal of the binary tree is bound by the L1 data pipe — every divergent 16-byte load of a lane is one
// wavefront, four per node visit, ≈130 per incoherent ray (profiles/r01_l1_data_pipe.txt). An eight-wide node with
// 8-bit child boxes is 80 bytes (five loads) and replaces three to four binary visits.
//
// The wide tree is a COLLAPSE of the host-built binary tree (the reference-order one, Raycaster.cpp:58-142 as mirrored by
// rb_bvh_host.cpp, or its RB_BUILD_SAH sibling),
// made on the host: a wide node adopts up to eight descendants of one binary node, listed in the order the reference's
// stack loop would reach them (right child before left, Raycaster.cpp:160-173), and the traversal consumes the listed
// children strictly in that order, leaves and interior children interleaved. Child boxes are stored relative to the
// node (origin + per-axis power-of-two scale) and rounded OUTWARDS by at least RB_WIDE_MARGIN of a quantum, so a wide
// traversal enters a superset of the leaves the reference enters, in the same relative order: the closest hit is the
// same record (same triangle on exact ties) unless a triangle is accepted whose own binary boxes the reference's
// rounded slab test rejects — the same class of exception RB_TRACE_ORDERED documents.
//
// Node, 20 words (five 16-byte loads):
//   w0..w2  origin p (fp32)            w3   ex | ey << 8 | ez << 16 (biased exponents of the per-axis scale) | n << 24
//   w4      index of the first interior child (interior children are consecutive, in slot order)
//   w5      index of the first triangle of the leaf children (wide triangle order, consecutive in slot order)
//   w6, w7  meta[8]: 0 = empty; interior: 0x20 | (24 + rank among interior children); leaf: unary(count) << 5 | offset
//   w8..w13 qlo x[8], y[8], z[8]       w14..w19 qhi x[8], y[8], z[8]      (empty slots: lo = 255, hi = 0)
// Triangles are re-listed in wide order (12 floats as in the binary layout; word 11 = index in hierarchy order).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <vector>

#define RB_WIDE_NODE_WORDS 20
#define RB_WIDE_MAX_LEAF_TRIS 3
#define RB_WIDE_MARGIN 0.0625 // quanta every stored plane is pushed outwards beyond exact rounding

struct RbFlatBvh;

struct RbWideBvh {
	std::vector<uint32_t> nodes; // RB_WIDE_NODE_WORDS per node
	// The same nodes with their children moved to slots chosen by direction: the child that lies furthest towards the
	// corner (sx, sy, sz) of the node, sign bit a of s set = positive, sits in (or near) slot s. A ray whose direction has
	// the sign bits o then meets the children roughly front to back in the order slot ^ (~o & 7) = 0, 1, ... 7 (Ylitie,
	// Karras, Laine 2017, section 3.2). Indices, triangle offsets and boxes are those of `nodes`; only the slots differ.
	// For closest-hit queries that do not promise the reference's order (RB_WIDE_OCTANT): same hits except exact-distance ties.
	std::vector<uint32_t> nodesByOctant;
	std::vector<float> tris;	 // 12 floats per triangle, wide order
	std::vector<int32_t> rootNode; // per root (mesh): wide node index, -1 = empty mesh
	int depth = 0;				 // wide nodes on the longest root-to-leaf path
};

// Collapses the flattened binary tree (which descendants a node adopts: surface-area dynamic programme, rb_wide.cpp; a
// leaf child may merge binary leaves of up to RB_WIDE_MAX_LEAF_TRIS triangles in total, listed in visiting order).
// Fails (with a reason) on binary leaves of more than RB_WIDE_MAX_LEAF_TRIS triangles and on non-finite boxes.
bool rb_wide_collapse(const RbFlatBvh & flat, size_t nroots, RbWideBvh & out, std::string * why);

struct RbDeviceWide {
	const uint4 * nodes; // 5 per node (slots in the reference's visiting order)
	const uint4 * nodesByOctant; // the same nodes, slots by direction
	const float4 * tris; // 3 per triangle
	const int * rootNode;
	int stackDepth; // stack entries a traversal needs
	unsigned int oneBits; // 0x3f800000, handed in as DATA so that it sits in a register (see RB_WIDE_F in rb_wide.cuh)
};
