// wide_host_check.cpp — TEST INFRASTRUCTURE: rb_wide.cuh (the eight-wide traversal the GPU kernels inline) compiled by g++
// and stepped one ray at a time on the CPU. tests/test_wide.py compares what it returns with the reference's golden hits.
// Build: g++ -O1 -std=c++17 -ffp-contract=off -shared -fPIC -I/usr/local/cuda/include wide_host_check.cpp
#include "cuda_shim.h"

#include "../../rendu_b200/csrc/rb_wide.cuh"

// roots: 8 floats per root as the library lays them out (min xyz, reference, max xyz, 0). mode 0: rb_wide_step (what
// rb_k_trace_wide runs); mode 1: rb_wide_phase_node / rb_wide_phase_triangle in the order one lane of
// rb_k_trace_wide_phased would run them; byOctant != 0: wideNodes holds the slots-by-direction copy and the steps run
// with OCT = true. Outputs per ray: triangle in hierarchy order (-1: miss), t, u, v.
// Returns the number of rays that could not take the wide tree (irregular rays: the kernels walk the binary tree for them).
extern "C" int rb_hc_wide_trace(const float * roots, int nroots, const uint32_t * wideNodes, const float * wideTris, const int * wideRoots, int wideDepth, const float * org, const float * dir, int nrays, float mini, float maxi, int anyHit, int mode, int byOctant, int * outTri, float * outTuv, int * outEntered) {
	RbDeviceBvh bvh = {};
	bvh.roots		= reinterpret_cast<const float4 *>(roots);
	bvh.nroots		= nroots;
	RbDeviceWide wide = {};
	wide.nodes		  = reinterpret_cast<const uint4 *>(wideNodes);
	wide.nodesByOctant = wide.nodes;
	wide.tris		  = reinterpret_cast<const float4 *>(wideTris);
	wide.rootNode	  = wideRoots;
	wide.stackDepth	  = wideDepth + 1;
	wide.oneBits	  = 0x3f800000u;
	if(size_t(5 * wide.stackDepth) * 4 > sizeof(rb_hc_shared)) return -1;
	const RbStack st = rb_make_stack(reinterpret_cast<int *>(rb_hc_shared), 1);
	int skipped = 0;
	for(int i = 0; i < nrays; ++i) {
		const RbRay ray = rb_make_ray(V3{org[3 * i], org[3 * i + 1], org[3 * i + 2]}, V3{dir[3 * i], dir[3 * i + 1], dir[3 * i + 2]});
		outTri[i] = -2;
		outTuv[3 * i] = outTuv[3 * i + 1] = outTuv[3 * i + 2] = 0.0f;
		const float big = 1152921504606846976.0f; // rb_wide_ray_ok of rb_raycaster.cu
		if(!(ray.regular && std::fabs(ray.inv.x) < big && std::fabs(ray.inv.y) < big && std::fabs(ray.inv.z) < big)) {
			++skipped;
			continue;
		}
		RbWideTraversal T;
		rb_wide_init(T, bvh, mini, maxi, st);
		if(byOctant) T.flip = rb_wide_flip(ray);
		bool done = false;
		long steps = 0;
		for(; !done && steps < 100000000; ++steps) {
			if(byOctant) {
				if(mode == 0) done = anyHit ? rb_wide_step<true, true>(T, bvh, wide, ray, st) : rb_wide_step<false, true>(T, bvh, wide, ray, st);
				else if(rb_wide_wants_triangle<true>(T)) done = anyHit ? rb_wide_phase_triangle<true, true>(T, wide, ray, st) : rb_wide_phase_triangle<false, true>(T, wide, ray, st);
				else done = rb_wide_phase_node<true>(T, bvh, wide, ray, st);
			} else if(mode == 0) {
				done = anyHit ? rb_wide_step<true>(T, bvh, wide, ray, st) : rb_wide_step<false>(T, bvh, wide, ray, st);
			} else if(rb_wide_wants_triangle(T)) {
				done = anyHit ? rb_wide_phase_triangle<true>(T, wide, ray, st) : rb_wide_phase_triangle<false>(T, wide, ray, st);
			} else {
				done = rb_wide_phase_node(T, bvh, wide, ray, st);
			}
		}
		if(outEntered) outEntered[i] = int(steps); // node entries + triangle tests (+ pops in the plain stepping)
		outTri[i]		  = T.best.tri;
		outTuv[3 * i]	  = T.best.t;
		outTuv[3 * i + 1] = T.best.u;
		outTuv[3 * i + 2] = T.best.v;
	}
	return skipped;
}

// The binary traversal the default kernels inline (rb_trace.cuh), one ray at a time: mode 0 = rb_step (what
// rb_k_closest / rb_k_any / rb_k_trace_persistent run), mode 1 = rb_phase_node / rb_phase_leaf in the order one lane of
// rb_k_trace_phased runs them (`nodeSteps` node steps per vote). ordered != 0: near child first (RB_TRACE_ORDERED).
// nodes: 16 floats per interior node, tris: 12 floats per triangle, roots: 8 floats per root — the device layout
// (rb_bvh_host.cpp, rb_host_flatten). Every ray is traced, irregular ones (NaN / inf / axis-parallel) included.
extern "C" int rb_hc_binary_trace(const float * nodes, const float * tris, const float * roots, int nroots, int stackDepth, const float * org, const float * dir, int nrays, float mini, float maxi, int anyHit, int ordered, int mode, int nodeSteps, int * outTri, float * outTuv) {
	RbDeviceBvh bvh = {};
	bvh.nodes		= reinterpret_cast<const float4 *>(nodes);
	bvh.tris		= reinterpret_cast<const float4 *>(tris);
	bvh.roots		= reinterpret_cast<const float4 *>(roots);
	bvh.nroots		= nroots;
	bvh.stackDepth	= stackDepth;
	if(size_t(stackDepth + 1) * 4 > sizeof(rb_hc_shared)) return -1;
	const RbStack st = rb_make_stack(reinterpret_cast<int *>(rb_hc_shared), 1);
	for(int i = 0; i < nrays; ++i) {
		const RbRay ray = rb_make_ray(V3{org[3 * i], org[3 * i + 1], org[3 * i + 2]}, V3{dir[3 * i], dir[3 * i + 1], dir[3 * i + 2]});
		RbTraversal T;
		rb_traversal_init(T, bvh, mini, maxi, st);
		// mode 2: the phase-split steps with the roots box-tested up front (RB_ROOTS_MASK, what the path tracer's closest-hit
		// wavefront runs on scenes of at most 32 meshes); reference order only.
		if(mode == 2) T.root = int(rb_roots_pretest<false>(bvh, ray, mini, maxi, nullptr));
		bool done = false;
		for(long guard = 0; !done && guard < 100000000; ++guard) {
			if(mode == 2) {
				if(rb_wants_leaf(T)) {
					done = anyHit ? rb_phase_leaf<true, false>(T, bvh, ray, nullptr) : rb_phase_leaf<false, false>(T, bvh, ray, nullptr);
				} else {
					for(int k = 0; k < nodeSteps && !done; ++k) {
						done = rb_phase_node<false, false, RB_ROOTS_MASK>(T, bvh, ray, st, nullptr);
						if(rb_wants_leaf(T)) break;
					}
				}
				continue;
			}
			if(mode == 0) {
				if(anyHit) done = ordered ? rb_step<true, false, true>(T, bvh, ray, st, nullptr) : rb_step<true, false, false>(T, bvh, ray, st, nullptr);
				else done = ordered ? rb_step<false, false, true>(T, bvh, ray, st, nullptr) : rb_step<false, false, false>(T, bvh, ray, st, nullptr);
			} else if(rb_wants_leaf(T)) {
				done = anyHit ? rb_phase_leaf<true, false>(T, bvh, ray, nullptr) : rb_phase_leaf<false, false>(T, bvh, ray, nullptr);
			} else {
				for(int k = 0; k < nodeSteps && !done; ++k) {
					done = ordered ? rb_phase_node<false, true>(T, bvh, ray, st, nullptr) : rb_phase_node<false, false>(T, bvh, ray, st, nullptr);
					if(rb_wants_leaf(T)) break;
				}
			}
		}
		outTri[i]		  = T.best.tri;
		outTuv[3 * i]	  = T.best.t;
		outTuv[3 * i + 1] = T.best.u;
		outTuv[3 * i + 2] = T.best.v;
	}
	return 0;
}
