// rb_raycaster.cu — the Raycaster half of the C ABI (include/rendu_b200.h) and its sm_100a kernels.
// Reference behaviour: src/engine/raycaster/Raycaster.cpp (addMesh :14-52, updateHierarchy :54-142,
// intersects :144-185, intersectsAny :187-224, visible :226-230).
#include "rb_internal.h"
#include "rb_trace.cuh"
#include "rb_wide.cuh"

#include <chrono>
#include <cstdlib>
#include <cstring>
#include <mutex>

// ---------------------------------------------------------------- error plumbing
static thread_local std::string g_lastError;

void rb_set_error(const char * fmt, ...) {
	char buf[1024];
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(buf, sizeof(buf), fmt, ap);
	va_end(ap);
	g_lastError = buf;
}

rb_status rb_fail(rb_status code, const char * fmt, ...) {
	char buf[1024];
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(buf, sizeof(buf), fmt, ap);
	va_end(ap);
	g_lastError = buf;
	return code;
}

bool rb_have_device() {
	static int count = -1;
	static std::once_flag once;
	std::call_once(once, [] {
		int n = 0;
		if(cudaGetDeviceCount(&n) != cudaSuccess) {
			cudaGetLastError();
			n = 0;
		}
		count = n;
	});
	return count > 0;
}

extern "C" const char * rb_last_error(void) {
	return g_lastError.c_str();
}

extern "C" const char * rb_version(void) {
	return "rendu_b200 0.1 (sm_100a)";
}

extern "C" int rb_device_count(void) {
	if(!rb_have_device()) return 0;
	int n = 0;
	cudaGetDeviceCount(&n);
	return n;
}

extern "C" rb_status rb_set_device(int device) {
	RB_REQUIRE_DEVICE();
	RB_CUDA(cudaSetDevice(device));
	return RB_OK;
}

// ---------------------------------------------------------------- kernels

#define RB_TRACE_BLOCK RB_TRACE_BLOCK_THREADS

// Reads ray i from tightly packed float3 arrays.
// Rays and hit records stream through once: they are loaded / stored with the evict-first hint (ld.global.cs /
// st.global.cs) so that they do not push the tree, which every ray re-reads, out of the 126 MB L2.
RB_D void rb_load_ray(const float * org, const float * dir, size_t i, V3 & o, V3 & d) {
	o = V3{__ldcs(org + 3 * i), __ldcs(org + 3 * i + 1), __ldcs(org + 3 * i + 2)};
	d = V3{__ldcs(dir + 3 * i), __ldcs(dir + 3 * i + 1), __ldcs(dir + 3 * i + 2)};
}

// Raycaster::Hit constructors (Raycaster.cpp:6-12) written as three 16-byte stores.
RB_D void rb_store_hit(rb_hit * out, size_t i, const RbDeviceBvh & bvh, const RbHitRec & h) {
	uint4 a, b, c;
	if(h.tri >= 0) {
		const float4 t0 = rb_ldg4(bvh.tris + 3 * (size_t)h.tri);
		const float4 t1 = rb_ldg4(bvh.tris + 3 * (size_t)h.tri + 1);
		const float w	= 1.0f - h.u - h.v;
		a = make_uint4(1u, __float_as_uint(h.t), __float_as_uint(h.u), __float_as_uint(h.v));
		b = make_uint4(__float_as_uint(w), 0u, __float_as_uint(t0.w), 0u);
		c = make_uint4(__float_as_uint(t1.w), 0u, (unsigned int)h.tri, 0u);
	} else {
		a = make_uint4(0u, __float_as_uint(FLT_MAX), 0u, 0u);
		b = make_uint4(0u, 0u, 0u, 0u);
		c = make_uint4(0u, 0u, 0u, 0u);
	}
	uint4 * o = reinterpret_cast<uint4 *>(out + i);
	__stcs(o, a);
	__stcs(o + 1, b);
	__stcs(o + 2, c);
}

// One ray per thread. The traversal stack lives in shared memory, one column per thread (bank-conflict free).
template<bool ORDERED>
__global__ void __launch_bounds__(RB_TRACE_BLOCK) rb_k_closest(RbDeviceBvh bvh, const float * __restrict__ org, const float * __restrict__ dir, float mini, float maxi, size_t n, rb_hit * __restrict__ out) {
	extern __shared__ int s_stack[];
	const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	if(i >= n) return;
	V3 o, d;
	rb_load_ray(org, dir, i, o, d);
	const RbRay ray = rb_make_ray(o, d);
	RbHitRec best;
	if(ORDERED) rb_trace_ordered<false, false>(bvh, ray, mini, maxi, rb_make_stack(s_stack + threadIdx.x, blockDim.x), best, nullptr);
	else rb_trace_exact<false, false>(bvh, ray, mini, maxi, rb_make_stack(s_stack + threadIdx.x, blockDim.x), best, nullptr);
	rb_store_hit(out, i, bvh, best);
}

template<bool ORDERED>
__global__ void __launch_bounds__(RB_TRACE_BLOCK) rb_k_any(RbDeviceBvh bvh, const float * __restrict__ org, const float * __restrict__ dir, float mini, float maxi, size_t n, uint8_t * __restrict__ out) {
	extern __shared__ int s_stack[];
	const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	if(i >= n) return;
	V3 o, d;
	rb_load_ray(org, dir, i, o, d);
	const RbRay ray = rb_make_ray(o, d);
	RbHitRec best;
	bool any;
	if(ORDERED) any = rb_trace_ordered<true, false>(bvh, ray, mini, maxi, rb_make_stack(s_stack + threadIdx.x, blockDim.x), best, nullptr);
	else any = rb_trace_exact<true, false>(bvh, ray, mini, maxi, rb_make_stack(s_stack + threadIdx.x, blockDim.x), best, nullptr);
	out[i] = any ? 1 : 0;
}

// Raycaster::visible (Raycaster.cpp:226-230): !intersectsAny(p0, p1-p0, 1e-4, length(p1-p0)).
__global__ void __launch_bounds__(RB_TRACE_BLOCK) rb_k_visible(RbDeviceBvh bvh, const float * __restrict__ p0, const float * __restrict__ p1, size_t n, uint8_t * __restrict__ out) {
	extern __shared__ int s_stack[];
	const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	if(i >= n) return;
	V3 a, b;
	rb_load_ray(p0, p1, i, a, b);
	const V3 direction = b - a;
	const float maxi   = length3(direction);
	const RbRay ray	   = rb_make_ray(a, direction);
	RbHitRec best;
	// Any-hit: order independent (see rb_launch_trace), so the near-child-first walk is used.
	const bool any = rb_trace_ordered<true, false>(bvh, ray, 0.0001f, maxi, rb_make_stack(s_stack + threadIdx.x, blockDim.x), best, nullptr);
	out[i] = any ? 0 : 1;
}

// Persistent threads with a warp-level work queue: the grid is sized to fill the machine once (SMs x resident CTAs),
// every warp pulls rays from a global cursor and refills the lanes whose ray has finished as soon as fewer than
// `threshold` lanes are still traversing (one atomicAdd per refill, lanes ranked with __ballot/__popc). Rays finish
// after very different numbers of steps; without the refill a warp runs at the pace of its slowest lane.
RB_D void rb_store_any(uint8_t * out, size_t i, const RbDeviceBvh &, const RbHitRec & h) {
	out[i] = h.tri >= 0 ? 1 : 0;
}

template<bool ANY, bool ORDERED>
__global__ void __launch_bounds__(RB_TRACE_BLOCK) rb_k_trace_persistent(RbDeviceBvh bvh, const float * __restrict__ org, const float * __restrict__ dir, float mini, float maxi, unsigned int n, void * __restrict__ out, unsigned int * __restrict__ cursor, int threshold) {
	extern __shared__ int s_stack[];
	const RbStack st = rb_make_stack(s_stack + threadIdx.x, blockDim.x);
	const unsigned int lane = threadIdx.x & 31u;
	bool active	   = false;
	bool exhausted = false; // warp-uniform: the queue has no more rays
	unsigned int rayIdx = 0;
	RbRay ray;
	RbTraversal T;
	for(;;) {
		if(!exhausted) {
			const unsigned int idle = __ballot_sync(0xffffffffu, !active);
			if(idle) {
				const int nIdle	 = __popc(idle);
				const int leader = __ffs(idle) - 1;
				unsigned int base = 0;
				if((int)lane == leader) base = atomicAdd(cursor, (unsigned int)nIdle);
				base = __shfl_sync(0xffffffffu, base, leader);
				if(!active) {
					const unsigned int idx = base + __popc(idle & ((1u << lane) - 1u));
					if(idx < n) {
						V3 o, d;
						rb_load_ray(org, dir, idx, o, d);
						ray = rb_make_ray(o, d);
						rb_traversal_init(T, bvh, mini, maxi, st);
						rayIdx = idx;
						active = true;
					}
				}
				exhausted = (base + (unsigned int)nIdle) >= n;
			}
		}
		if(!__any_sync(0xffffffffu, active)) break;
		bool more;
		do {
			if(active) {
				if(rb_step<ANY, false, ORDERED>(T, bvh, ray, st, nullptr)) {
					if(ANY) rb_store_any(static_cast<uint8_t *>(out), rayIdx, bvh, T.best);
					else rb_store_hit(static_cast<rb_hit *>(out), rayIdx, bvh, T.best);
					active = false;
				}
			}
			const int nActive = __popc(__ballot_sync(0xffffffffu, active));
			more = exhausted ? (nActive > 0) : (nActive >= threshold);
		} while(more);
	}
}

// Phase-voting scheduler. Every live lane is waiting either for a node step or for a leaf step; the warp executes
// the leaf phase once at least `leafVotes` lanes wait for it (or nobody wants a node), otherwise the node phase. Lanes
// of the other kind idle for that iteration instead of the whole warp running at the pace of one lane's descent.
template<bool ANY, bool ORDERED, bool BLOCKS, int TREE = 0> // TREE: 0 none, RB_ROOTS_TREE (register mask), RB_ROOTS_TREE_BIG (mask in the stack column)
__global__ void __launch_bounds__(RB_TRACE_BLOCK) rb_k_trace_phased(RbDeviceBvh bvh, const float * __restrict__ org, const float * __restrict__ dir, float mini, float maxi, unsigned int n, void * __restrict__ out, unsigned int * __restrict__ cursor, int threshold, int leafVotes, int nodeSteps) {
	extern __shared__ int s_stack[];
	const RbStack st		= rb_make_stack(s_stack + threadIdx.x + (TREE == RB_ROOTS_TREE_BIG ? bvh.rootWords * blockDim.x : 0), blockDim.x);
	const unsigned int lane = threadIdx.x & 31u;
	bool active = false, exhausted = false;
	unsigned int rayIdx = 0;
	RbRay ray;
	RbTraversal T;
	for(;;) {
		if(!exhausted) {
			const unsigned int idle = __ballot_sync(0xffffffffu, !active);
			if(idle) {
				const int nIdle	 = __popc(idle);
				const int leader = __ffs(idle) - 1;
				unsigned int base = 0;
				if((int)lane == leader) base = atomicAdd(cursor, (unsigned int)nIdle);
				base = __shfl_sync(0xffffffffu, base, leader);
				if(!active) {
					const unsigned int idx = base + __popc(idle & ((1u << lane) - 1u));
					if(idx < n) {
						V3 o, d;
						rb_load_ray(org, dir, idx, o, d);
						ray = rb_make_ray(o, d);
						rb_traversal_init(T, bvh, mini, maxi, st);
						// TREE: scenes of 9..64 meshes — the roots this ray enters from one walk of the tree over the root
						// boxes (rb_roots_tree_pretest) instead of one test per root; same set, same visiting order.
						if(TREE == RB_ROOTS_TREE) rb_roots_tree_pretest<false>(T, bvh, ray, st, mini, maxi, nullptr);
						if(TREE == RB_ROOTS_TREE_BIG) rb_roots_tree_pretest_big<false>(T, bvh, ray, st, mini, maxi, nullptr);
						rayIdx = idx;
						active = true;
					}
				}
				exhausted = (base + (unsigned int)nIdle) >= n;
			}
		}
		unsigned int activeM = __ballot_sync(0xffffffffu, active);
		if(activeM == 0u) break;
		bool more;
		do {
			// One ballot per iteration: a live lane waits either for a leaf step or for a node step.
			const bool wantsLeaf	 = active && rb_wants_leaf(T);
			const unsigned int leafM = __ballot_sync(0xffffffffu, wantsLeaf);
			const unsigned int nodeM = activeM & ~leafM;
			bool finished = false;
			if(__popc(leafM) >= leafVotes || nodeM == 0u) {
				if(wantsLeaf) finished = rb_phase_leaf<ANY, false>(T, bvh, ray, nullptr);
			} else if(active && !wantsLeaf) {
				// Up to `nodeSteps` descents per vote: the vote, executed by all 32 lanes, costs about as much as a
				// node step; a lane that reaches a leaf (or finishes) early sits out the remaining steps.
				for(int k = 0; k < nodeSteps; ++k) {
					finished = rb_phase_node<false, ORDERED, TREE ? TREE : (BLOCKS ? RB_ROOTS_BLOCKS : RB_ROOTS_LAZY)>(T, bvh, ray, st, nullptr);
					if(finished || rb_wants_leaf(T)) break;
				}
			}
			if(finished) {
				if(ANY) rb_store_any(static_cast<uint8_t *>(out), rayIdx, bvh, T.best);
				else rb_store_hit(static_cast<rb_hit *>(out), rayIdx, bvh, T.best);
				active = false;
			}
			activeM = __ballot_sync(0xffffffffu, active);
			more	= exhausted ? (activeM != 0u) : (__popc(activeM) >= threshold);
		} while(more);
	}
}

// Phase-voting scheduler with a per-warp shared-memory ring of PREPARED rays (experiment, RB_RING=1).
// In the kernel above a refill puts a dependent global load and the ≈60 instructions of Ray::Ray (sqrt, IEEE
// divisions) on the critical path of the whole warp, executed by the few lanes that are idle; that is why refilling
// early does not pay there. Here the warp fetches 32 rays at a time with one atomicAdd, every lane prepares one of them
// (full SIMD width, one coalesced load) and parks it in the warp's ring; a lane whose ray has finished takes the next
// prepared ray with a ballot-ranked shared-memory read. Closest-hit records go the other way: (t, u, v, triangle, ray)
// is parked in a second ring and expanded into the 48-byte Hit by the full warp, 32 records at a time.
// Per ray the order of operations is untouched: identical results.
#define RB_RING_RAY_WORDS 11 // pos, dir, inv (9), ray index, regular flag
#define RB_RING_DONE_WORDS 5 // t, u, v, triangle, ray index
#define RB_RING_DONE_SLOTS 64
#define RB_RING_WARP_WORDS (RB_RING_RAY_WORDS * 32 + RB_RING_DONE_WORDS * RB_RING_DONE_SLOTS)

template<bool ANY, bool ORDERED>
__global__ void __launch_bounds__(RB_TRACE_BLOCK) rb_k_trace_ring(RbDeviceBvh bvh, const float * __restrict__ org, const float * __restrict__ dir, float mini, float maxi, unsigned int n, void * __restrict__ out, unsigned int * __restrict__ cursor, int threshold, int leafVotes, int nodeSteps) {
	extern __shared__ int s_stack[];
	const RbStack st		 = rb_make_stack(s_stack + threadIdx.x, blockDim.x);
	const unsigned int lane	 = threadIdx.x & 31u;
	const unsigned int below = (1u << lane) - 1u;
	float * ring = reinterpret_cast<float *>(s_stack + (bvh.stackDepth > 0 ? bvh.stackDepth : 1) * blockDim.x) + (threadIdx.x >> 5) * RB_RING_WARP_WORDS;
	float * done = ring + RB_RING_RAY_WORDS * 32;
	unsigned int ringHead = 0, ringCount = 0, doneHead = 0, doneCount = 0; // warp-uniform
	bool active = false, exhausted = false; // exhausted: the global queue is empty (the ring may still hold rays)
	unsigned int rayIdx = 0;
	RbRay ray;
	RbTraversal T;
	// The guard bounds the loop whatever happens: an experiment must not be able to hang the device.
	for(unsigned int guard = 0; guard < 0x20000000u; ++guard) {
		unsigned int idle = __ballot_sync(0xffffffffu, !active);
		while(idle != 0u) {
			if(ringCount == 0u) {
				if(exhausted) break;
				unsigned int base = 0;
				if(lane == 0) base = atomicAdd(cursor, 32u);
				base = __shfl_sync(0xffffffffu, base, 0);
				const unsigned int idx = base + lane;
				if(idx < n) {
					V3 o, d;
					rb_load_ray(org, dir, idx, o, d);
					const RbRay r = rb_make_ray(o, d);
					ring[0 * 32 + lane] = r.pos.x; ring[1 * 32 + lane] = r.pos.y; ring[2 * 32 + lane] = r.pos.z;
					ring[3 * 32 + lane] = r.dir.x; ring[4 * 32 + lane] = r.dir.y; ring[5 * 32 + lane] = r.dir.z;
					ring[6 * 32 + lane] = r.inv.x; ring[7 * 32 + lane] = r.inv.y; ring[8 * 32 + lane] = r.inv.z;
					ring[9 * 32 + lane]	 = __uint_as_float(idx);
					ring[10 * 32 + lane] = __uint_as_float(r.regular ? 1u : 0u);
				}
				__syncwarp();
				ringHead  = 0;
				ringCount = base < n ? min(32u, n - base) : 0u;
				exhausted = (base >= n) || (n - base <= 32u);
				if(ringCount == 0u) break;
			}
			const unsigned int take = min((unsigned int)__popc(idle), ringCount);
			const unsigned int rank = __popc(idle & below);
			if(!active && rank < take) {
				const unsigned int slot = ringHead + rank; // < 32: ringHead + ringCount never exceeds the batch
				ray.pos		= V3{ring[0 * 32 + slot], ring[1 * 32 + slot], ring[2 * 32 + slot]};
				ray.dir		= V3{ring[3 * 32 + slot], ring[4 * 32 + slot], ring[5 * 32 + slot]};
				ray.inv		= V3{ring[6 * 32 + slot], ring[7 * 32 + slot], ring[8 * 32 + slot]};
				rayIdx		= __float_as_uint(ring[9 * 32 + slot]);
				ray.regular = __float_as_uint(ring[10 * 32 + slot]) != 0u;
				ray.pxy = rb_pack2(ray.pos.x, ray.pos.y); ray.pzx = rb_pack2(ray.pos.z, ray.pos.x); ray.pyz = rb_pack2(ray.pos.y, ray.pos.z);
				ray.ixy = rb_pack2(ray.inv.x, ray.inv.y); ray.izx = rb_pack2(ray.inv.z, ray.inv.x); ray.iyz = rb_pack2(ray.inv.y, ray.inv.z);
				rb_traversal_init(T, bvh, mini, maxi, st);
				active = true;
			}
			__syncwarp(); // the slots are read before a later batch may overwrite them
			ringHead += take;
			ringCount -= take;
			idle = __ballot_sync(0xffffffffu, !active);
		}
		unsigned int activeM = __ballot_sync(0xffffffffu, active);
		if(activeM == 0u) break;
		const bool noMore = exhausted && ringCount == 0u;
		bool more;
		do {
			const bool wantsLeaf	 = active && rb_wants_leaf(T);
			const unsigned int leafM = __ballot_sync(0xffffffffu, wantsLeaf);
			const unsigned int nodeM = activeM & ~leafM;
			bool finished = false;
			if(__popc(leafM) >= leafVotes || nodeM == 0u) {
				if(wantsLeaf) finished = rb_phase_leaf<ANY, false>(T, bvh, ray, nullptr);
			} else if(active && !wantsLeaf) {
				for(int k = 0; k < nodeSteps; ++k) {
					finished = rb_phase_node<false, ORDERED, false>(T, bvh, ray, st, nullptr);
					if(finished || rb_wants_leaf(T)) break;
				}
			}
			if(finished) {
				if(ANY) rb_store_any(static_cast<uint8_t *>(out), rayIdx, bvh, T.best);
				active = false;
			}
			if(!ANY) {
				const unsigned int fin = __ballot_sync(0xffffffffu, finished);
				if(fin != 0u) {
					if(finished) {
						const unsigned int slot = (doneHead + doneCount + __popc(fin & below)) & (RB_RING_DONE_SLOTS - 1);
						done[0 * RB_RING_DONE_SLOTS + slot] = T.best.t;
						done[1 * RB_RING_DONE_SLOTS + slot] = T.best.u;
						done[2 * RB_RING_DONE_SLOTS + slot] = T.best.v;
						done[3 * RB_RING_DONE_SLOTS + slot] = __int_as_float(T.best.tri);
						done[4 * RB_RING_DONE_SLOTS + slot] = __uint_as_float(rayIdx);
					}
					doneCount += __popc(fin); // <= 31 + 32 < RB_RING_DONE_SLOTS
					__syncwarp();
					if(doneCount >= 32u) {
						const unsigned int slot = (doneHead + lane) & (RB_RING_DONE_SLOTS - 1);
						RbHitRec h;
						h.t	  = done[0 * RB_RING_DONE_SLOTS + slot];
						h.u	  = done[1 * RB_RING_DONE_SLOTS + slot];
						h.v	  = done[2 * RB_RING_DONE_SLOTS + slot];
						h.tri = __float_as_int(done[3 * RB_RING_DONE_SLOTS + slot]);
						rb_store_hit(static_cast<rb_hit *>(out), __float_as_uint(done[4 * RB_RING_DONE_SLOTS + slot]), bvh, h);
						__syncwarp();
						doneHead = (doneHead + 32u) & (RB_RING_DONE_SLOTS - 1);
						doneCount -= 32u;
					}
				}
			}
			activeM = __ballot_sync(0xffffffffu, active);
			more	= activeM != 0u && (noMore || __popc(activeM) >= threshold);
		} while(more);
	}
	if(!ANY && lane < doneCount) {
		const unsigned int slot = (doneHead + lane) & (RB_RING_DONE_SLOTS - 1);
		RbHitRec h;
		h.t	  = done[0 * RB_RING_DONE_SLOTS + slot];
		h.u	  = done[1 * RB_RING_DONE_SLOTS + slot];
		h.v	  = done[2 * RB_RING_DONE_SLOTS + slot];
		h.tri = __float_as_int(done[3 * RB_RING_DONE_SLOTS + slot]);
		rb_store_hit(static_cast<rb_hit *>(out), __float_as_uint(done[4 * RB_RING_DONE_SLOTS + slot]), bvh, h);
	}
}

// Eight-wide quantised tree (RB_TRACE_WIDE; rb_wide.h, rb_wide.cuh): persistent threads + warp work queue as in
// rb_k_trace_persistent, one rb_wide_step per iteration. Rays whose slab arithmetic can produce NaNs or overflow
// (a zero / denormal direction component, |1/d| >= 2^60, a non-finite origin) are rare: they are walked through the
// binary tree on the spot, by the exact NaN-faithful traversal.
RB_D bool rb_wide_ray_ok(const RbRay & ray) {
	const float big = 1152921504606846976.0f; // 2^60
	return ray.regular && fabsf(ray.inv.x) < big && fabsf(ray.inv.y) < big && fabsf(ray.inv.z) < big;
}

template<bool ANY, bool OCT>
__global__ void __launch_bounds__(RB_TRACE_BLOCK) rb_k_trace_wide(RbDeviceBvh bvh, const float * __restrict__ org, const float * __restrict__ dir, float mini, float maxi, unsigned int n, void * __restrict__ out, unsigned int * __restrict__ cursor, int threshold, RbDeviceWide wide) {
	extern __shared__ int s_stack[];
	const RbStack st = rb_make_stack(s_stack + threadIdx.x, blockDim.x);
	const unsigned int lane = threadIdx.x & 31u;
	bool active = false, exhausted = false;
	unsigned int rayIdx = 0;
	RbRay ray;
	RbWideTraversal T;
	for(;;) {
		if(!exhausted) {
			const unsigned int idle = __ballot_sync(0xffffffffu, !active);
			if(idle) {
				const int nIdle	 = __popc(idle);
				const int leader = __ffs(idle) - 1;
				unsigned int base = 0;
				if((int)lane == leader) base = atomicAdd(cursor, (unsigned int)nIdle);
				base = __shfl_sync(0xffffffffu, base, leader);
				if(!active) {
					const unsigned int idx = base + __popc(idle & ((1u << lane) - 1u));
					if(idx < n) {
						V3 o, d;
						rb_load_ray(org, dir, idx, o, d);
						ray = rb_make_ray(o, d);
						if(rb_wide_ray_ok(ray)) {
							rb_wide_init(T, bvh, mini, maxi, st);
							if(OCT) T.flip = rb_wide_flip(ray);
							rayIdx = idx;
							active = true;
						} else {
							RbTraversal B;
							rb_traversal_init(B, bvh, mini, maxi, st);
							if(ANY) {
								while(!rb_step_ordered<true, false, false>(B, bvh, ray, st, nullptr)) {
								}
							} else {
								while(!rb_step_exact<false, false, false>(B, bvh, ray, st, nullptr)) {
								}
							}
							if(ANY) rb_store_any(static_cast<uint8_t *>(out), idx, bvh, B.best);
							else rb_store_hit(static_cast<rb_hit *>(out), idx, bvh, B.best);
						}
					}
				}
				exhausted = (base + (unsigned int)nIdle) >= n;
			}
		}
		if(!__any_sync(0xffffffffu, active)) {
			if(exhausted) break;
			continue; // every fetched ray took the binary path: fetch again
		}
		bool more;
		do {
			if(active) {
				if(rb_wide_step<ANY, OCT>(T, bvh, wide, ray, st)) {
					if(ANY) rb_store_any(static_cast<uint8_t *>(out), rayIdx, bvh, T.best);
					else rb_store_hit(static_cast<rb_hit *>(out), rayIdx, bvh, T.best);
					active = false;
				}
			}
			const int nActive = __popc(__ballot_sync(0xffffffffu, active));
			more = exhausted ? (nActive > 0) : (nActive >= threshold);
		} while(more);
	}
}

// The wide traversal under the phase-voting scheduler of rb_k_trace_phased (RB_WIDE_PHASED=1; not yet run on a device):
// a triangle phase runs once `leafVotes` lanes wait for one (or no lane has a node to enter), otherwise one node phase.
template<bool ANY>
__global__ void __launch_bounds__(RB_TRACE_BLOCK) rb_k_trace_wide_phased(RbDeviceBvh bvh, const float * __restrict__ org, const float * __restrict__ dir, float mini, float maxi, unsigned int n, void * __restrict__ out, unsigned int * __restrict__ cursor, int threshold, int leafVotes, RbDeviceWide wide) {
	extern __shared__ int s_stack[];
	const RbStack st = rb_make_stack(s_stack + threadIdx.x, blockDim.x);
	const unsigned int lane = threadIdx.x & 31u;
	bool active = false, exhausted = false;
	unsigned int rayIdx = 0;
	RbRay ray;
	RbWideTraversal T;
	for(;;) {
		if(!exhausted) {
			const unsigned int idle = __ballot_sync(0xffffffffu, !active);
			if(idle) {
				const int nIdle	 = __popc(idle);
				const int leader = __ffs(idle) - 1;
				unsigned int base = 0;
				if((int)lane == leader) base = atomicAdd(cursor, (unsigned int)nIdle);
				base = __shfl_sync(0xffffffffu, base, leader);
				if(!active) {
					const unsigned int idx = base + __popc(idle & ((1u << lane) - 1u));
					if(idx < n) {
						V3 o, d;
						rb_load_ray(org, dir, idx, o, d);
						ray = rb_make_ray(o, d);
						if(rb_wide_ray_ok(ray)) {
							rb_wide_init(T, bvh, mini, maxi, st);
							rayIdx = idx;
							active = true;
						} else {
							RbTraversal B;
							rb_traversal_init(B, bvh, mini, maxi, st);
							if(ANY) {
								while(!rb_step_ordered<true, false, false>(B, bvh, ray, st, nullptr)) {
								}
							} else {
								while(!rb_step_exact<false, false, false>(B, bvh, ray, st, nullptr)) {
								}
							}
							if(ANY) rb_store_any(static_cast<uint8_t *>(out), idx, bvh, B.best);
							else rb_store_hit(static_cast<rb_hit *>(out), idx, bvh, B.best);
						}
					}
				}
				exhausted = (base + (unsigned int)nIdle) >= n;
			}
		}
		unsigned int activeM = __ballot_sync(0xffffffffu, active);
		if(activeM == 0u) {
			if(exhausted) break;
			continue;
		}
		bool more;
		do {
			const bool wantsTri		 = active && rb_wide_wants_triangle(T);
			const unsigned int triM	 = __ballot_sync(0xffffffffu, wantsTri);
			const unsigned int nodeM = activeM & ~triM;
			bool finished = false;
			if(__popc(triM) >= leafVotes || nodeM == 0u) {
				if(wantsTri) finished = rb_wide_phase_triangle<ANY>(T, wide, ray, st);
			} else if(active && !wantsTri) {
				finished = rb_wide_phase_node(T, bvh, wide, ray, st);
			}
			if(finished) {
				if(ANY) rb_store_any(static_cast<uint8_t *>(out), rayIdx, bvh, T.best);
				else rb_store_hit(static_cast<rb_hit *>(out), rayIdx, bvh, T.best);
				active = false;
			}
			activeM = __ballot_sync(0xffffffffu, active);
			more	= exhausted ? (activeM != 0u) : (__popc(activeM) >= threshold);
		} while(more);
	}
}

// Two rays per lane. The profile of the kernels above (profiles/) shows why a warp is slow: in the node loop only ~11
// of 32 lanes are busy — the others have reached a leaf and wait for the slowest descent. Here every lane owns two
// independent rays, A and B, each with its own registers and stack column; one iteration runs the node phase for the A
// rays that want it and then for the B rays that want it, so a lane whose A ray waits for a leaf phase keeps
// descending with B. Leaf phases are postponed until `leafVotes` of the warp's 64 slots have a leaf pending. Per ray,
// the order of operations is untouched, hence identical results. (A first version kept one "working" slot and swapped
// the register sets; the swap code, executed for two or three lanes at a time, cost more than it saved.)
struct RbSlot {
	RbRay ray;
	RbTraversal T;
	unsigned int rayIdx;
	bool live;
};

template<bool ANY>
RB_D void rb_slot_finish(RbSlot & S, const RbDeviceBvh & bvh, void * out) {
	if(ANY) rb_store_any(static_cast<uint8_t *>(out), S.rayIdx, bvh, S.T.best);
	else rb_store_hit(static_cast<rb_hit *>(out), S.rayIdx, bvh, S.T.best);
	S.live = false;
}

RB_D void rb_slot_start(RbSlot & S, const RbDeviceBvh & bvh, const float * org, const float * dir, unsigned int idx, float mini, float maxi, const RbStack & st) {
	V3 o, d;
	rb_load_ray(org, dir, idx, o, d);
	S.ray = rb_make_ray(o, d);
	rb_traversal_init(S.T, bvh, mini, maxi, st);
	S.rayIdx = idx;
	S.live	 = true;
}

template<bool ANY, bool ORDERED>
__global__ void __launch_bounds__(RB_TRACE_BLOCK) rb_k_trace_dual(RbDeviceBvh bvh, const float * __restrict__ org, const float * __restrict__ dir, float mini, float maxi, unsigned int n, void * __restrict__ out, unsigned int * __restrict__ cursor, int refillSlots, int leafVotes) {
	extern __shared__ int s_stack[];
	const RbStack stackA	= rb_make_stack(s_stack + threadIdx.x, 2 * blockDim.x);
	const RbStack stackB	= rb_make_stack(s_stack + threadIdx.x + blockDim.x, 2 * blockDim.x);
	const unsigned int lane = threadIdx.x & 31u;
	const unsigned int below = (1u << lane) - 1u;
	RbSlot A, B;
	A.live = false;
	B.live = false;
	A.rayIdx = B.rayIdx = 0;
	bool exhausted = false;
	for(;;) {
		if(!exhausted) {
			const unsigned int freeA = __ballot_sync(0xffffffffu, !A.live);
			const unsigned int freeB = __ballot_sync(0xffffffffu, !B.live);
			const int nA = __popc(freeA), nB = __popc(freeB);
			if(nA + nB >= refillSlots) {
				unsigned int base = 0;
				if(lane == 0) base = atomicAdd(cursor, (unsigned int)(nA + nB));
				base = __shfl_sync(0xffffffffu, base, 0);
				if(!A.live) {
					const unsigned int idx = base + __popc(freeA & below);
					if(idx < n) rb_slot_start(A, bvh, org, dir, idx, mini, maxi, stackA);
				}
				if(!B.live) {
					const unsigned int idx = base + (unsigned int)nA + __popc(freeB & below);
					if(idx < n) rb_slot_start(B, bvh, org, dir, idx, mini, maxi, stackB);
				}
				exhausted = (base + (unsigned int)(nA + nB)) >= n;
			}
		}
		const bool aLeaf = A.live && rb_wants_leaf(A.T), bLeaf = B.live && rb_wants_leaf(B.T);
		const bool aNode = A.live && !aLeaf, bNode = B.live && !bLeaf;
		const int nLeaf = __popc(__ballot_sync(0xffffffffu, aLeaf)) + __popc(__ballot_sync(0xffffffffu, bLeaf));
		const int nNode = __popc(__ballot_sync(0xffffffffu, aNode)) + __popc(__ballot_sync(0xffffffffu, bNode));
		if(nLeaf + nNode == 0) {
			if(exhausted) break;
			continue;
		}
		if(nLeaf >= leafVotes || nNode == 0) {
			if(aLeaf && rb_phase_leaf<ANY, false>(A.T, bvh, A.ray, nullptr)) rb_slot_finish<ANY>(A, bvh, out);
			if(bLeaf && rb_phase_leaf<ANY, false>(B.T, bvh, B.ray, nullptr)) rb_slot_finish<ANY>(B, bvh, out);
		} else {
			if(aNode && rb_phase_node<false, ORDERED>(A.T, bvh, A.ray, stackA, nullptr)) rb_slot_finish<ANY>(A, bvh, out);
			if(bNode && rb_phase_node<false, ORDERED>(B.T, bvh, B.ray, stackB, nullptr)) rb_slot_finish<ANY>(B, bvh, out);
		}
	}
}

// Instrumented copy: same loops, counts the work that defines the algorithmic bytes per ray.
template<bool ANY, bool ORDERED>
__global__ void __launch_bounds__(RB_TRACE_BLOCK) rb_k_count(RbDeviceBvh bvh, const float * __restrict__ org, const float * __restrict__ dir, float mini, float maxi, size_t n, unsigned long long * __restrict__ counters) {
	extern __shared__ int s_stack[];
	const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	RbWork w = {0u, 0u, 0u, 0u};
	if(i < n) {
		V3 o, d;
		rb_load_ray(org, dir, i, o, d);
		const RbRay ray = rb_make_ray(o, d);
		RbHitRec best;
		if(ORDERED) rb_trace_ordered<ANY, true>(bvh, ray, mini, maxi, rb_make_stack(s_stack + threadIdx.x, blockDim.x), best, &w);
		else rb_trace_exact<ANY, true>(bvh, ray, mini, maxi, rb_make_stack(s_stack + threadIdx.x, blockDim.x), best, &w);
	}
	// Warp-level reduction, one atomic per warp and counter.
	unsigned int a = w.interior, b = w.tris, c = w.roots, m = w.maxStack;
	for(int off = 16; off > 0; off >>= 1) {
		a += __shfl_down_sync(0xffffffffu, a, off);
		b += __shfl_down_sync(0xffffffffu, b, off);
		c += __shfl_down_sync(0xffffffffu, c, off);
		m = max(m, __shfl_down_sync(0xffffffffu, m, off));
	}
	if((threadIdx.x & 31) == 0) {
		atomicAdd(counters + 0, (unsigned long long)a);
		atomicAdd(counters + 1, (unsigned long long)b);
		atomicAdd(counters + 2, (unsigned long long)c);
		atomicMax(counters + 3, (unsigned long long)m);
	}
}

// ---------------------------------------------------------------- handle management

size_t rb_trace_smem_bytes(const rb_raycaster * rc, int blockThreads) {
	const int depth = rc->dev.stackDepth > 0 ? rc->dev.stackDepth : 1;
	return size_t(depth) * size_t(blockThreads) * sizeof(int);
}

// The wide tree is a derived copy: it goes whenever the binary tree it was collapsed from is replaced or re-fitted.
static void rb_wide_free_device(rb_raycaster * rc) {
	cudaFree(rc->d_wideNodes);
	cudaFree(rc->d_wideNodesOct);
	rc->d_wideNodesOct = nullptr;
	cudaFree(rc->d_wideTris);
	cudaFree(rc->d_wideRoots);
	rc->d_wideNodes = nullptr;
	rc->d_wideTris	= nullptr;
	rc->d_wideRoots = nullptr;
	rc->devWide		= RbDeviceWide{};
	rc->wideOnDevice = false;
}

void rb_raycaster_free_device(rb_raycaster * rc) {
	if(!rb_have_device() || rc->device < 0) return; // nothing was ever allocated: leave the caller's current device alone
	int previous = 0;
	cudaGetDevice(&previous);
	cudaSetDevice(rc->device);
	cudaFree(rc->d_nodes);
	cudaFree(rc->d_tris);
	cudaFree(rc->d_roots);
	cudaFree(rc->d_parent);
	cudaFree(rc->d_leafParent);
	cudaFree(rc->d_leafCode);
	cudaFree(rc->d_flags);
	cudaFree(rc->d_verts);
	cudaFree(rc->d_triVerts);
	cudaFree(rc->d_triIds);
	cudaFree(rc->d_cursors);
	rc->d_cursors = nullptr;
	rb_wide_free_device(rc);
	cudaFree(rc->d_rootBlocks);
	rc->d_rootBlocks   = nullptr;
	cudaFree(rc->d_rootTree);
	rc->d_rootTree	   = nullptr;
	rc->rootTreeNodes  = 0;
	rc->dev.rootTree   = nullptr;
	rc->dev.rootWords  = 0;
	rc->rootTreeDepth  = 0;
	rc->rootBlocksCapacity = 0;
	rc->dev.rootBlocks = nullptr;
	rc->dev.rootLevels = 0;
	rc->d_nodes = rc->d_tris = rc->d_roots = nullptr;
	rc->d_parent	 = nullptr;
	rc->d_leafParent = nullptr;
	rc->d_leafCode	 = nullptr;
	rc->d_flags		 = nullptr;
	rc->d_verts		 = nullptr;
	rc->d_triVerts	 = nullptr;
	rc->d_triIds	 = nullptr;
	rc->dev		   = RbDeviceBvh{};
	rc->device	   = -1;
	// No device tree, nothing to query: a rebuild or upload that fails half-way must leave a handle that answers
	// "hierarchy not built", not one whose kernels run on null pointers. Set again at the end of a successful upload.
	rc->built = false;
	if(rc->traceMode == RB_TRACE_WIDE) rc->traceMode = RB_TRACE_EXACT;
	cudaSetDevice(previous);
}

rb_status rb_raycaster_upload_flat(rb_raycaster * rc, const RbFlatBvh & flat) {
	RB_REQUIRE_DEVICE();
	rb_raycaster_free_device(rc);
	RB_CUDA(cudaGetDevice(&rc->device));
	rc->nInterior = flat.nodes.size() / 16;
	rc->nTris	  = flat.tris.size() / 12 - flat.dummyTris;
	rc->nVerts	  = rc->verts.size() / 3;
	// Allocations are padded so that empty arrays still give valid pointers.
	RB_CUDA(cudaMalloc(&rc->d_nodes, flat.nodes.size() * sizeof(float) + 64));
	RB_CUDA(cudaMalloc(&rc->d_tris, flat.tris.size() * sizeof(float) + 64));
	RB_CUDA(cudaMalloc(&rc->d_roots, flat.roots.size() * sizeof(float) + 64));
	RB_CUDA(cudaMalloc(&rc->d_parent, flat.parent.size() * sizeof(int) + 64));
	RB_CUDA(cudaMalloc(&rc->d_verts, rc->verts.size() * sizeof(float) + 64));
	RB_CUDA(cudaMalloc(&rc->d_triVerts, flat.triVerts.size() * sizeof(uint32_t) + 64));
	RB_CUDA(cudaMalloc(&rc->d_triIds, flat.triIds.size() * sizeof(uint32_t) + 64));
	RB_CUDA(cudaMalloc(&rc->d_leafParent, flat.leafParent.size() * sizeof(int) + 64));
	RB_CUDA(cudaMalloc(&rc->d_leafCode, flat.leafCode.size() * sizeof(int) + 64));
	RB_CUDA(cudaMalloc(&rc->d_flags, (flat.parent.size() + 1) * sizeof(unsigned int)));
	rc->nLeaves = flat.leafCode.size();
	RB_CUDA(cudaMalloc(&rc->d_cursors, RB_CURSOR_RING * sizeof(unsigned int)));
	RB_CUDA(cudaMemcpy(rc->d_nodes, flat.nodes.data(), flat.nodes.size() * sizeof(float), cudaMemcpyHostToDevice));
	RB_CUDA(cudaMemcpy(rc->d_tris, flat.tris.data(), flat.tris.size() * sizeof(float), cudaMemcpyHostToDevice));
	RB_CUDA(cudaMemcpy(rc->d_roots, flat.roots.data(), flat.roots.size() * sizeof(float), cudaMemcpyHostToDevice));
	RB_CUDA(cudaMemcpy(rc->d_parent, flat.parent.data(), flat.parent.size() * sizeof(int), cudaMemcpyHostToDevice));
	RB_CUDA(cudaMemcpy(rc->d_verts, rc->verts.data(), rc->verts.size() * sizeof(float), cudaMemcpyHostToDevice));
	RB_CUDA(cudaMemcpy(rc->d_triVerts, flat.triVerts.data(), flat.triVerts.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
	RB_CUDA(cudaMemcpy(rc->d_triIds, flat.triIds.data(), flat.triIds.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
	RB_CUDA(cudaMemcpy(rc->d_leafParent, flat.leafParent.data(), flat.leafParent.size() * sizeof(int), cudaMemcpyHostToDevice));
	RB_CUDA(cudaMemcpy(rc->d_leafCode, flat.leafCode.data(), flat.leafCode.size() * sizeof(int), cudaMemcpyHostToDevice));
	rc->dev.nodes	   = rc->d_nodes;
	rc->dev.tris	   = rc->d_tris;
	rc->dev.roots	   = rc->d_roots;
	rc->dev.nroots	   = int(rc->meshCount);
	rc->dev.stackDepth = rc->maxDepth + 1;
	rc->rootBlocksWanted = rb_root_blocks_worthwhile(flat.roots.data(), flat.roots.size() / 8);
	rb_status st		 = rb_root_tree_topology(rc, flat.roots.data(), flat.roots.size() / 8);
	if(st) return st;
	// The ray's stack column also serves the walk of the root tree (one pending node per level) and, beyond 64 meshes,
	// holds the mask of entered roots in front of the stack proper.
	rc->dev.stackDepth = rb_stack_entries(rc);
	st = rb_build_root_levels(rc, cudaStreamPerThread);
	if(st) return st;
	RB_CUDA(cudaStreamSynchronize(cudaStreamPerThread));
	rc->built		   = true;
	rc->buildSerial++;
	rc->hostMirrorStale = false;
	return RB_OK;
}

extern "C" rb_status rb_raycaster_create(rb_raycaster ** out) {
	if(!out) return rb_fail(RB_ERR_INVALID, "rb_raycaster_create: null output");
	*out = new rb_raycaster();
	// Scheduler tuning (DESIGN.md "traversal kernels"): RB_PERSISTENT=0 selects one ray per thread,
	// RB_FETCH_THRESHOLD the number of live lanes below which a warp refills from the queue.
	if(const char * e = getenv("RB_PERSISTENT")) (*out)->persistent = atoi(e) != 0;
	if(const char * e = getenv("RB_FETCH_THRESHOLD")) (*out)->fetchThreshold = atoi(e);
	if(const char * e = getenv("RB_LEAF_VOTES")) (*out)->leafVotes = atoi(e);
	if(const char * e = getenv("RB_NODE_STEPS")) (*out)->nodeSteps = atoi(e);
	if(const char * e = getenv("RB_DUAL")) (*out)->dual = atoi(e) != 0;
	if(const char * e = getenv("RB_RING")) (*out)->ring = atoi(e) != 0;
	if(const char * e = getenv("RB_WIDE_PHASED")) (*out)->widePhased = atoi(e) != 0;
	if(const char * e = getenv("RB_WIDE_OCTANT")) (*out)->wideOctant = atoi(e);
	if(const char * e = getenv("RB_DUAL_REFILL")) (*out)->dualRefill = atoi(e);
	return RB_OK;
}

extern "C" void rb_raycaster_destroy(rb_raycaster * rc) {
	if(!rc) return;
	rb_raycaster_free_device(rc);
	delete rc;
}

// vec3(model * vec4(v, 1)) as Raycaster::addMesh applies it (Raycaster.cpp:22-28); an exact identity copies.
// `model != glm::mat4(1.0f)` (Raycaster.cpp:22) is an element-wise float comparison.
static void rb_transform_positions(float * dst, const float * positions, size_t nverts, const float * model) {
	static const float identity[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
	bool isIdentity = true;
	if(model) {
		for(int k = 0; k < 16; ++k) isIdentity = isIdentity && (model[k] == identity[k]);
	}
	if(isIdentity) {
		if(nverts) std::memcpy(dst, positions, nverts * 3 * sizeof(float));
		return;
	}
	// glm evaluates (m0*x + m1*y) + (m2*z + m3*w) per component (type_mat4x4.inl:561-571).
	const float * m = model;
	rb_parallel_for(nverts, size_t(1) << 16, [&](size_t begin, size_t end) {
		for(size_t v = begin; v < end; ++v) {
			const float x = positions[3 * v], y = positions[3 * v + 1], z = positions[3 * v + 2];
			for(int r = 0; r < 3; ++r) {
				const float a0 = m[0 + r] * x + m[4 + r] * y;
				const float a1 = m[8 + r] * z + m[12 + r] * 1.0f;
				dst[3 * v + r] = a0 + a1;
			}
		}
	});
}

// BoundingBox(v0,v1,v2) (Bounds.cpp:12-15): min(min(v0,v1),v2), glm::min(x,y) = (y<x)?y:x.
static void rb_triangle_box(rb_tri & ti, const float * verts) {
	const float * a = verts + 3 * ti.v0;
	const float * b = verts + 3 * ti.v1;
	const float * c = verts + 3 * ti.v2;
	for(int k = 0; k < 3; ++k) {
		const float mn01 = (b[k] < a[k]) ? b[k] : a[k];
		const float mx01 = (a[k] < b[k]) ? b[k] : a[k];
		ti.bmin[k]		 = (c[k] < mn01) ? c[k] : mn01;
		ti.bmax[k]		 = (mx01 < c[k]) ? c[k] : mx01;
	}
}

extern "C" rb_status rb_raycaster_add_mesh(rb_raycaster * rc, const float * positions, size_t nverts, const uint32_t * indices, size_t nindices, const float * model) {
	if(!rc || (!positions && nverts) || (!indices && nindices)) return rb_fail(RB_ERR_INVALID, "rb_raycaster_add_mesh: null argument");
	if(rc->meshFirst.size() != rc->meshCount) return rb_fail(RB_ERR_INVALID, "rb_raycaster_add_mesh: the uploaded hierarchy on this handle mixes its meshes under %u root(s); meshes cannot be appended to it", rc->meshCount);
	const size_t indexOffset = rc->verts.size() / 3;
	rc->verts.resize((indexOffset + nverts) * 3);
	rb_transform_positions(rc->verts.data() + indexOffset * 3, positions, nverts, model);
	const size_t ntri		  = nindices / 3;
	const size_t startTriangle = rc->tris.size();
	rc->tris.resize(startTriangle + ntri);
	const size_t nvTotal = rc->verts.size() / 3;
	std::atomic<size_t> badTriangle{SIZE_MAX};
	const uint32_t meshId = rc->meshCount;
	rb_parallel_for(ntri, size_t(1) << 16, [&](size_t begin, size_t end) {
		for(size_t t = begin; t < end; ++t) {
			rb_tri & ti = rc->tris[startTriangle + t];
			std::memset(&ti, 0, sizeof(ti));
			ti.v0	   = indexOffset + indices[3 * t + 0];
			ti.v1	   = indexOffset + indices[3 * t + 1];
			ti.v2	   = indexOffset + indices[3 * t + 2];
			ti.localId = 3 * t;
			ti.meshId  = meshId;
			if(ti.v0 >= nvTotal || ti.v1 >= nvTotal || ti.v2 >= nvTotal) {
				size_t seen = badTriangle.load();
				while(t < seen && !badTriangle.compare_exchange_weak(seen, t)) {}
				return;
			}
			rb_triangle_box(ti, rc->verts.data());
		}
	});
	if(badTriangle.load() != SIZE_MAX) {
		rc->tris.resize(startTriangle);
		rc->verts.resize(indexOffset * 3);
		return rb_fail(RB_ERR_INVALID, "rb_raycaster_add_mesh: index out of range in triangle %zu", badTriangle.load());
	}
	// One root per mesh, appended after the existing roots: a rebuild restarts from these (Raycaster.cpp:43-46).
	rc->meshFirst.push_back(startTriangle);
	rc->meshTris.push_back(ntri);
	rc->meshVertFirst.push_back(indexOffset);
	rc->meshVerts.push_back(nverts);
	rc->meshCount++;
	rc->built = false;
	return RB_OK;
}

// Incremental geometry: mesh `mesh` (addMesh order) gets new model-space positions and / or a new model matrix, same
// topology. The world-space vertices are recomputed exactly as addMesh computes them and only this mesh's range goes
// to the device; the boxes follow at the next rb_raycaster_refit (pass NULL positions there: "fit what the handle holds").
extern "C" rb_status rb_raycaster_update_mesh(rb_raycaster * rc, uint32_t mesh, const float * positions, size_t nverts, const float * model) {
	if(!rc || !positions) return rb_fail(RB_ERR_INVALID, "rb_raycaster_update_mesh: null argument");
	if(mesh >= rc->meshVertFirst.size()) return rb_fail(RB_ERR_INVALID, "rb_raycaster_update_mesh: mesh %u unknown (%zu meshes were added through rb_raycaster_add_mesh)", mesh, rc->meshVertFirst.size());
	if(nverts != rc->meshVerts[mesh]) return rb_fail(RB_ERR_INVALID, "rb_raycaster_update_mesh: mesh %u has %zu vertices, %zu given", mesh, size_t(rc->meshVerts[mesh]), nverts);
	RB_REQUIRE_DEVICE();
	if(!rc->built || !rc->d_verts) return rb_fail(RB_ERR_INVALID, "rb_raycaster_update_mesh: hierarchy not built");
	const size_t first = size_t(rc->meshVertFirst[mesh]);
	float * dst		   = rc->verts.data() + first * 3;
	rb_transform_positions(dst, positions, nverts, model);
	rc->triBoxesStale = true;
	RB_CUDA(cudaSetDevice(rc->device));
	if(nverts) RB_CUDA(cudaMemcpyAsync(rc->d_verts + first * 3, dst, nverts * 3 * sizeof(float), cudaMemcpyHostToDevice, cudaStreamPerThread));
	RB_CUDA(cudaStreamSynchronize(cudaStreamPerThread)); // `verts` is pageable memory the caller may update again
	return RB_OK;
}

extern "C" rb_status rb_raycaster_build(rb_raycaster * rc, rb_build_mode mode) {
	if(!rc) return rb_fail(RB_ERR_INVALID, "rb_raycaster_build: null handle");
	rc->buildMode = mode;
	if(rc->triBoxesStale) {
		// Vertices moved since addMesh (refit / update_mesh): the triangle boxes the builders start from are redone.
		rb_parallel_for(rc->tris.size(), size_t(1) << 16, [&](size_t begin, size_t end) {
			for(size_t t = begin; t < end; ++t) rb_triangle_box(rc->tris[t], rc->verts.data());
		});
		rc->triBoxesStale = false;
	}
	if(rc->meshFirst.size() != rc->meshCount || rc->meshTris.size() != rc->meshCount) return rb_fail(RB_ERR_INVALID, "rb_raycaster_build: the uploaded hierarchy on this handle mixes its meshes under %u root(s); there are no per-mesh triangle ranges to rebuild from", rc->meshCount);
	if(mode == RB_BUILD_REFERENCE || mode == RB_BUILD_SAH) {
		// The host mirror is built first so that it can be read back (rb_raycaster_download_bvh) even on a machine
		// without a GPU; tracing still needs the device copy.
		rb_host_build_reference(rc);
		RB_REQUIRE_DEVICE();
		RbFlatBvh flat;
		std::string why;
		if(!rb_host_flatten(rc, flat, &why)) return rb_fail(RB_ERR_INVALID, "rb_raycaster_build: %s", why.c_str());
		return rb_raycaster_upload_flat(rc, flat);
	}
	if(mode == RB_BUILD_LBVH) {
		RB_REQUIRE_DEVICE();
		return rb_lbvh_build(rc);
	}
	return rb_fail(RB_ERR_INVALID, "rb_raycaster_build: unknown mode %d", int(mode));
}

static bool rb_reconstruct_mesh_ranges(rb_raycaster * rc, std::string * why) {
	const size_t nt = rc->tris.size(), nv = rc->verts.size() / 3, nm = rc->meshCount;
	rc->meshFirst.assign(nm, 0);
	rc->meshTris.assign(nm, 0);
	rc->meshVertFirst.assign(nm, 0);
	rc->meshVerts.assign(nm, 0);
	std::vector<uint64_t> vmin(nm, UINT64_MAX);
	size_t t = 0;
	for(size_t m = 0; m < nm; ++m) {
		rc->meshFirst[m] = t;
		while(t < nt && rc->tris[t].meshId == m) {
			const rb_tri & ti = rc->tris[t++];
			vmin[m] = std::min(vmin[m], std::min(uint64_t(ti.v0), std::min(uint64_t(ti.v1), uint64_t(ti.v2))));
		}
		rc->meshTris[m] = t - rc->meshFirst[m];
	}
	if(t != nt) {
		*why = "triangle " + std::to_string(t) + " carries mesh id " + std::to_string(rc->tris[t].meshId) + ": the triangles of a mesh must be contiguous and the meshes in order, with ids below mesh_count";
		return false;
	}
	// Vertex ranges: mesh m starts at the lowest vertex its triangles use (vertices nobody uses stay with the mesh before).
	uint64_t next = nv;
	for(size_t m = nm; m-- > 0;) {
		const uint64_t first = m == 0 ? 0 : std::min(vmin[m], next);
		if(first > next) {
			*why = "vertex indices of mesh " + std::to_string(m) + " out of range";
			return false;
		}
		rc->meshVertFirst[m] = first;
		rc->meshVerts[m]	 = next - first;
		next				 = first;
	}
	return true;
}

extern "C" rb_status rb_raycaster_upload_bvh(rb_raycaster * rc, const rb_node * nodes, size_t nnodes, const rb_tri * tris, size_t ntris, const float * vertices, size_t nverts, uint32_t mesh_count) {
	if(!rc || (!nodes && nnodes) || (!tris && ntris) || (!vertices && nverts)) return rb_fail(RB_ERR_INVALID, "rb_raycaster_upload_bvh: null argument");
	RB_REQUIRE_DEVICE();
	rb_raycaster_free_device(rc); // the host arrays are replaced below: whatever tree is on the device no longer matches them
	rc->nodes.assign(nodes, nodes + nnodes);
	rc->meshFirst.clear();
	rc->meshTris.clear();
	rc->meshVertFirst.clear();
	rc->meshVerts.clear();
	rc->triBoxesStale = true; // a later rebuild starts from triangle boxes recomputed on `vertices`, not from the caller's records
	rc->tris.assign(tris, tris + ntris);
	rc->verts.assign(vertices, vertices + nverts * 3);
	rc->meshCount = mesh_count;
	rc->buildMode = RB_BUILD_REFERENCE;
	std::string why;
	// The per-mesh triangle and vertex ranges that rb_raycaster_add_mesh keeps (a later rb_raycaster_build restarts from
	// them, a later add_mesh appends to them) are reconstructed from the records: addMesh appends, and the reference's
	// build permutes triangles inside their mesh's range only (Raycaster.cpp:30-46,100-125), so both are contiguous and in
	// mesh order in any hierarchy the reference could have produced.
	// A hierarchy with mixed meshes under one root (a downloaded LBVH tree) has no such ranges: it can be traced, re-fitted
	// and read back, but rb_raycaster_build(REFERENCE / SAH) and rb_raycaster_add_mesh on this handle fail with a clear error.
	if(!rb_reconstruct_mesh_ranges(rc, &why)) {
		rc->meshFirst.clear();
		rc->meshTris.clear();
		rc->meshVertFirst.clear();
		rc->meshVerts.clear();
	}
	const int depth = rb_host_tree_depth(rc, &why);
	if(depth < 0) return rb_fail(RB_ERR_INVALID, "rb_raycaster_upload_bvh: %s", why.c_str());
	rc->maxDepth = depth;
	RbFlatBvh flat;
	if(!rb_host_flatten(rc, flat, &why)) return rb_fail(RB_ERR_INVALID, "rb_raycaster_upload_bvh: %s", why.c_str());
	return rb_raycaster_upload_flat(rc, flat);
}

extern "C" rb_status rb_raycaster_sizes(const rb_raycaster * rc, uint64_t out[5]) {
	if(!rc || !out) return rb_fail(RB_ERR_INVALID, "rb_raycaster_sizes: null argument");
	if(rc->buildMode == RB_BUILD_LBVH && rc->built) {
		const uint64_t nl = rc->lbvhLeaves; // a proper binary tree: one record per leaf and per live interior node
		out[0] = nl ? 2 * nl - 1 : 0;
		out[3] = 1; // one tree over all meshes
	} else {
		out[0] = rc->nodes.size();
		out[3] = rc->meshCount;
	}
	out[1] = rc->tris.size();
	out[2] = rc->verts.size() / 3;
	out[4] = uint64_t(rc->maxDepth);
	return RB_OK;
}

extern "C" rb_status rb_raycaster_download_bvh(const rb_raycaster * rc, rb_node * nodes, rb_tri * tris, float * vertices) {
	if(!rc) return rb_fail(RB_ERR_INVALID, "rb_raycaster_download_bvh: null handle");
	if(rc->built && (rc->buildMode == RB_BUILD_LBVH || rc->hostMirrorStale)) {
		return rb_lbvh_download(rc, nodes, tris, vertices);
	}
	if(nodes && !rc->nodes.empty()) std::memcpy(nodes, rc->nodes.data(), rc->nodes.size() * sizeof(rb_node));
	if(tris && !rc->tris.empty()) std::memcpy(tris, rc->tris.data(), rc->tris.size() * sizeof(rb_tri));
	if(vertices && !rc->verts.empty()) std::memcpy(vertices, rc->verts.data(), rc->verts.size() * sizeof(float));
	return RB_OK;
}

extern "C" rb_status rb_raycaster_refit(rb_raycaster * rc, const float * world_positions, size_t nverts) {
	if(!rc) return rb_fail(RB_ERR_INVALID, "rb_raycaster_refit: null handle");
	if(!world_positions && nverts) return rb_fail(RB_ERR_INVALID, "rb_raycaster_refit: null positions with a vertex count");
	RB_REQUIRE_DEVICE();
	if(!rc->built) return rb_fail(RB_ERR_INVALID, "rb_raycaster_refit: hierarchy not built");
	if(rc->wideOnDevice) { // its quantised boxes describe the old positions
		RB_CUDA(cudaSetDevice(rc->device));
		rb_wide_free_device(rc);
	}
	rc->wide.reset();
	if(rc->areaSerial != rc->buildSerial) { // baseline of rb_raycaster_tree_cost: the tree as built, taken before the first refit touches it
		double built[2];
		rb_status as = rb_raycaster_tree_cost(rc, built);
		if(as) return as;
	}
	rc->areaStale = true;
	return rb_lbvh_refit(rc, world_positions, nverts);
}

extern "C" rb_status rb_raycaster_build_times(const rb_raycaster * rc, double out[2]) {
	if(!rc || !out) return rb_fail(RB_ERR_INVALID, "rb_raycaster_build_times: null argument");
	out[0] = rc->buildDeviceMs;
	out[1] = rc->refitDeviceMs;
	return RB_OK;
}

extern "C" rb_status rb_raycaster_set_trace_mode(rb_raycaster * rc, rb_trace_mode mode) {
	if(!rc) return rb_fail(RB_ERR_INVALID, "rb_raycaster_set_trace_mode: null handle");
	if(mode != RB_TRACE_EXACT && mode != RB_TRACE_ORDERED && mode != RB_TRACE_WIDE) return rb_fail(RB_ERR_INVALID, "unknown trace mode %d", int(mode));
	if(mode == RB_TRACE_WIDE && !rc->wideOnDevice) return rb_fail(RB_ERR_INVALID, "RB_TRACE_WIDE: no wide tree on the device (call rb_raycaster_wide_build after rb_raycaster_build)");
	rc->traceMode = mode;
	return RB_OK;
}

extern "C" rb_status rb_raycaster_wide_build(rb_raycaster * rc, uint64_t out[4]) {
	if(!rc) return rb_fail(RB_ERR_INVALID, "rb_raycaster_wide_build: null handle");
	if(rc->buildMode == RB_BUILD_LBVH || rc->hostMirrorStale || rc->nodes.empty()) return rb_fail(RB_ERR_INVALID, "rb_raycaster_wide_build: needs a host-built hierarchy (RB_BUILD_REFERENCE or RB_BUILD_SAH), not re-fitted since");
	const bool timing = getenv("RB_BUILD_TIMING") != nullptr;
	const auto t0	  = std::chrono::steady_clock::now();
	RbFlatBvh flat;
	std::string why;
	if(!rb_host_flatten(rc, flat, &why)) return rb_fail(RB_ERR_INVALID, "rb_raycaster_wide_build: %s", why.c_str());
	const auto t1 = std::chrono::steady_clock::now();
	auto wide	  = std::make_unique<RbWideBvh>();
	if(!rb_wide_collapse(flat, rc->meshCount, *wide, &why)) return rb_fail(RB_ERR_INVALID, "rb_raycaster_wide_build: %s", why.c_str());
	if(timing) fprintf(stderr, "[rb_raycaster_wide_build] flatten %.1f ms, collapse %.1f ms\n", std::chrono::duration<double, std::milli>(t1 - t0).count(), std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t1).count());
	rc->wide = std::move(wide);
	if(out) {
		out[0] = rc->wide->nodes.size() / RB_WIDE_NODE_WORDS;
		out[1] = rc->wide->tris.size() / 12;
		out[2] = rc->wide->rootNode.size();
		out[3] = uint64_t(rc->wide->depth);
	}
	if(!rc->built || rc->device < 0) return RB_OK; // host only (no device tree yet): sizes and download still work
	RB_CUDA(cudaSetDevice(rc->device));
	rb_wide_free_device(rc);
	const RbWideBvh & w = *rc->wide;
	RB_CUDA(cudaMalloc(&rc->d_wideNodes, w.nodes.size() * sizeof(uint32_t) + 64));
	RB_CUDA(cudaMalloc(&rc->d_wideNodesOct, w.nodesByOctant.size() * sizeof(uint32_t) + 64));
	RB_CUDA(cudaMalloc(&rc->d_wideTris, w.tris.size() * sizeof(float) + 64));
	RB_CUDA(cudaMalloc(&rc->d_wideRoots, w.rootNode.size() * sizeof(int32_t) + 64));
	RB_CUDA(cudaMemcpy(rc->d_wideNodes, w.nodes.data(), w.nodes.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
	RB_CUDA(cudaMemcpy(rc->d_wideNodesOct, w.nodesByOctant.data(), w.nodesByOctant.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
	RB_CUDA(cudaMemcpy(rc->d_wideTris, w.tris.data(), w.tris.size() * sizeof(float), cudaMemcpyHostToDevice));
	RB_CUDA(cudaMemcpy(rc->d_wideRoots, w.rootNode.data(), w.rootNode.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
	rc->devWide.nodes	   = rc->d_wideNodes;
	rc->devWide.nodesByOctant = rc->d_wideNodesOct;
	rc->devWide.tris	   = rc->d_wideTris;
	rc->devWide.rootNode   = rc->d_wideRoots;
	rc->devWide.stackDepth = w.depth + 1;
	rc->devWide.oneBits	   = 0x3f800000u;
	rc->wideOnDevice	   = true;
	return RB_OK;
}

extern "C" rb_status rb_raycaster_wide_download(const rb_raycaster * rc, uint32_t * nodes, float * tris, int32_t * root_nodes) {
	if(!rc) return rb_fail(RB_ERR_INVALID, "rb_raycaster_wide_download: null handle");
	if(!rc->wide) return rb_fail(RB_ERR_INVALID, "rb_raycaster_wide_download: no wide tree (call rb_raycaster_wide_build)");
	if(nodes) std::memcpy(nodes, rc->wide->nodes.data(), rc->wide->nodes.size() * sizeof(uint32_t));
	if(tris) std::memcpy(tris, rc->wide->tris.data(), rc->wide->tris.size() * sizeof(float));
	if(root_nodes) std::memcpy(root_nodes, rc->wide->rootNode.data(), rc->wide->rootNode.size() * sizeof(int32_t));
	return RB_OK;
}

extern "C" rb_status rb_raycaster_wide_download_by_octant(const rb_raycaster * rc, uint32_t * nodes) {
	if(!rc || !nodes) return rb_fail(RB_ERR_INVALID, "rb_raycaster_wide_download_by_octant: null argument");
	if(!rc->wide) return rb_fail(RB_ERR_INVALID, "rb_raycaster_wide_download_by_octant: no wide tree (call rb_raycaster_wide_build)");
	std::memcpy(nodes, rc->wide->nodesByOctant.data(), rc->wide->nodesByOctant.size() * sizeof(uint32_t));
	return RB_OK;
}

extern "C" uint64_t rb_raycaster_launch_count(const rb_raycaster * rc) {
	return rc ? rc->launches.load() : 0;
}

// ---------------------------------------------------------------- queries

static rb_status rb_check_query(const rb_raycaster * rc, const char * what) {
	if(!rc) return rb_fail(RB_ERR_INVALID, "%s: null handle", what);
	RB_REQUIRE_DEVICE();
	if(!rc->built) return rb_fail(RB_ERR_INVALID, "%s: hierarchy not built (call rb_raycaster_build)", what);
	return RB_OK;
}

template<typename K>
static rb_status rb_prepare_kernel(K kernel, size_t smem) {
	if(smem > 48 * 1024) {
		RB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
	}
	return RB_OK;
}

// Launch of the traversal kernels. The scheduler is selected per handle (rb_internal.h): one ray per thread, or one of the
// persistent work-queue kernels (plain, phase-voting [default], two rays per lane).
// Persistent launch shared by the three work-queue kernels: grid = SMs x resident CTAs (never more CTAs than there are
// rays for), a fresh cursor from the handle's ring, batches of < 2^32 rays per launch (the cursor is 32 bit).
template<bool ANY, typename Kernel, typename... Tuning>
static rb_status rb_launch_persistent(const rb_raycaster * rc, Kernel kernel, size_t smem, int raysPerThread, const float * d_origins, const float * d_directions, float mini, float maxi, size_t nrays, void * d_out, cudaStream_t s, Tuning... tuning) {
	if(smem > 48 * 1024) RB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
	int perSm = 0, sms = 0;
	RB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, kernel, RB_TRACE_BLOCK, smem));
	RB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, rc->device));
	if(perSm < 1) return rb_fail(RB_ERR_CUDA, "traversal kernel does not fit on an SM (stack depth %d)", rc->dev.stackDepth);
	const size_t chunkMax = 0xC0000000ull;
	const size_t perBlock = size_t(RB_TRACE_BLOCK) * raysPerThread;
	for(size_t done = 0; done < nrays; done += chunkMax) {
		const size_t n			  = (nrays - done < chunkMax) ? (nrays - done) : chunkMax;
		const size_t blocksNeeded = (n + perBlock - 1) / perBlock;
		const unsigned grid		  = unsigned(blocksNeeded < size_t(perSm) * sms ? blocksNeeded : size_t(perSm) * sms);
		unsigned int * d_cursor	  = rc->d_cursors + (rc->cursorNext.fetch_add(1) % RB_CURSOR_RING);
		RB_CUDA(cudaMemsetAsync(d_cursor, 0, sizeof(unsigned int), s));
		void * outp = ANY ? static_cast<void *>(static_cast<uint8_t *>(d_out) + done) : static_cast<void *>(static_cast<rb_hit *>(d_out) + done);
		kernel<<<grid, RB_TRACE_BLOCK, smem, s>>>(rc->dev, d_origins + 3 * done, d_directions + 3 * done, mini, maxi, (unsigned int)n, outp, d_cursor, tuning...);
		RB_CUDA(cudaGetLastError());
		rc->launches++;
	}
	return RB_OK;
}

template<bool ANY>
static rb_status rb_launch_trace(const rb_raycaster * rc, const float * d_origins, const float * d_directions, float mini, float maxi, size_t nrays, void * d_out, cudaStream_t s) {
	const size_t smem  = rb_trace_smem_bytes(rc, RB_TRACE_BLOCK);
	// An any-hit query never shrinks `maxi`, so which nodes are reachable — and therefore whether some triangle is
	// accepted — does not depend on the visiting order: the near-child-first order is used for it unconditionally
	// (it finds an occluder sooner), also in RB_TRACE_EXACT mode. Closest-hit queries follow the handle's mode.
	const bool ordered = ANY || rc->traceMode == RB_TRACE_ORDERED;
	RB_CUDA(cudaSetDevice(rc->device)); // the stream and the pointers must belong to the device that holds the tree
	if(rc->traceMode == RB_TRACE_WIDE) {
		if(!rc->wideOnDevice) return rb_fail(RB_ERR_INVALID, "RB_TRACE_WIDE: the wide tree is gone (the hierarchy was rebuilt or re-fitted): call rb_raycaster_wide_build again");
		const int entries	  = 5 * rc->devWide.stackDepth > rc->dev.stackDepth ? 5 * rc->devWide.stackDepth : (rc->dev.stackDepth > 0 ? rc->dev.stackDepth : 1);
		const size_t wideSmem = size_t(entries) * RB_TRACE_BLOCK * sizeof(int);
		if(rc->widePhased) return rb_launch_persistent<ANY>(rc, rb_k_trace_wide_phased<ANY>, wideSmem, 1, d_origins, d_directions, mini, maxi, nrays, d_out, s, rc->fetchThreshold, rc->leafVotes > 0 ? rc->leafVotes : 6, rc->devWide);
		// Slots by direction (RB_WIDE_OCTANT: 1 = closest-hit queries, which then agree with the reference except on
		// exact-distance ties, as RB_TRACE_ORDERED does — 18 % fewer steps on the incoherent C2 rays; 2 = any-hit queries
		// too, whose result cannot depend on the order but which gain nothing from it on C2).
		if(rc->wideOctant >= (ANY ? 2 : 1)) return rb_launch_persistent<ANY>(rc, rb_k_trace_wide<ANY, true>, wideSmem, 1, d_origins, d_directions, mini, maxi, nrays, d_out, s, rc->fetchThreshold, rc->devWide);
		return rb_launch_persistent<ANY>(rc, rb_k_trace_wide<ANY, false>, wideSmem, 1, d_origins, d_directions, mini, maxi, nrays, d_out, s, rc->fetchThreshold, rc->devWide);
	}
	if(rc->persistent && rc->dual) {
		auto kernel = ordered ? rb_k_trace_dual<ANY, true> : rb_k_trace_dual<ANY, false>;
		return rb_launch_persistent<ANY>(rc, kernel, 2 * smem, 2, d_origins, d_directions, mini, maxi, nrays, d_out, s, rc->dualRefill, rc->leafVotes > 0 ? rc->leafVotes : 24);
	}
	if(rc->persistent && rc->leafVotes > 0 && rc->ring && rc->dev.rootLevels == 0) {
		auto kernel = ordered ? rb_k_trace_ring<ANY, true> : rb_k_trace_ring<ANY, false>;
		const size_t ringBytes = size_t(RB_TRACE_BLOCK / 32) * RB_RING_WARP_WORDS * sizeof(float);
		return rb_launch_persistent<ANY>(rc, kernel, smem + ringBytes, 1, d_origins, d_directions, mini, maxi, nrays, d_out, s, rc->fetchThreshold, rc->leafVotes, rc->nodeSteps > 0 ? rc->nodeSteps : 1);
	}
	if(rc->persistent && rc->leafVotes > 0) {
		// Handles that carry root blocks (many meshes in a spatially coherent order) use the kernels compiled with them.
		auto kernel = rc->dev.rootLevels > 0 ? (ordered ? rb_k_trace_phased<ANY, true, true> : rb_k_trace_phased<ANY, false, true>)
											 : (ordered ? rb_k_trace_phased<ANY, true, false> : rb_k_trace_phased<ANY, false, false>);
		// A root tree (9..64 meshes) serves every query whose root tests see a fixed maxi: any-hit, and closest-hit in the
		// reference's order (the near-child-first closest-hit order tests roots against the shrinking maxi: lazy loop).
		if(rc->dev.rootTree && (ANY || !ordered)) {
			if(rc->dev.rootWords > 0) kernel = ordered ? rb_k_trace_phased<ANY, true, false, RB_ROOTS_TREE_BIG> : rb_k_trace_phased<ANY, false, false, RB_ROOTS_TREE_BIG>;
			else kernel = ordered ? rb_k_trace_phased<ANY, true, false, RB_ROOTS_TREE> : rb_k_trace_phased<ANY, false, false, RB_ROOTS_TREE>;
		}
		return rb_launch_persistent<ANY>(rc, kernel, smem, 1, d_origins, d_directions, mini, maxi, nrays, d_out, s, rc->fetchThreshold, rc->leafVotes, rc->nodeSteps > 0 ? rc->nodeSteps : 1);
	}
	if(rc->persistent) {
		auto kernel = ordered ? rb_k_trace_persistent<ANY, true> : rb_k_trace_persistent<ANY, false>;
		return rb_launch_persistent<ANY>(rc, kernel, smem, 1, d_origins, d_directions, mini, maxi, nrays, d_out, s, rc->fetchThreshold);
	}
	rb_status st;
	const unsigned grid = unsigned((nrays + RB_TRACE_BLOCK - 1) / RB_TRACE_BLOCK);
	if(ANY) {
		auto kernel = ordered ? rb_k_any<true> : rb_k_any<false>;
		st = rb_prepare_kernel(kernel, smem);
		if(st) return st;
		kernel<<<grid, RB_TRACE_BLOCK, smem, s>>>(rc->dev, d_origins, d_directions, mini, maxi, nrays, static_cast<uint8_t *>(d_out));
	} else {
		auto kernel = ordered ? rb_k_closest<true> : rb_k_closest<false>;
		st = rb_prepare_kernel(kernel, smem);
		if(st) return st;
		kernel<<<grid, RB_TRACE_BLOCK, smem, s>>>(rc->dev, d_origins, d_directions, mini, maxi, nrays, static_cast<rb_hit *>(d_out));
	}
	RB_CUDA(cudaGetLastError());
	rc->launches++;
	return RB_OK;
}

extern "C" rb_status rb_raycaster_intersect_device(const rb_raycaster * rc, const float * d_origins, const float * d_directions, float mini, float maxi, size_t nrays, rb_hit * d_out, void * stream) {
	rb_status st = rb_check_query(rc, "rb_raycaster_intersect_device");
	if(st) return st;
	if(nrays == 0) return RB_OK;
	return rb_launch_trace<false>(rc, d_origins, d_directions, mini, maxi, nrays, d_out, static_cast<cudaStream_t>(stream));
}

extern "C" rb_status rb_raycaster_intersect_any_device(const rb_raycaster * rc, const float * d_origins, const float * d_directions, float mini, float maxi, size_t nrays, uint8_t * d_out, void * stream) {
	rb_status st = rb_check_query(rc, "rb_raycaster_intersect_any_device");
	if(st) return st;
	if(nrays == 0) return RB_OK;
	return rb_launch_trace<true>(rc, d_origins, d_directions, mini, maxi, nrays, d_out, static_cast<cudaStream_t>(stream));
}

extern "C" rb_status rb_raycaster_count_work_device(const rb_raycaster * rc, const float * d_origins, const float * d_directions, float mini, float maxi, size_t nrays, int any_hit, uint64_t counters[4], void * stream) {
	rb_status st = rb_check_query(rc, "rb_raycaster_count_work_device");
	if(st) return st;
	if(!counters) return rb_fail(RB_ERR_INVALID, "rb_raycaster_count_work_device: null counters");
	cudaStream_t s = static_cast<cudaStream_t>(stream);
	unsigned long long * d_c = nullptr;
	RB_CUDA(cudaMallocAsync(&d_c, 4 * sizeof(unsigned long long), s));
	RB_CUDA(cudaMemsetAsync(d_c, 0, 4 * sizeof(unsigned long long), s));
	if(nrays) {
		const size_t smem	= rb_trace_smem_bytes(rc, RB_TRACE_BLOCK);
		const unsigned grid = unsigned((nrays + RB_TRACE_BLOCK - 1) / RB_TRACE_BLOCK);
		const bool ordered	= any_hit || rc->traceMode == RB_TRACE_ORDERED; // the order rb_launch_trace uses
#define RB_LAUNCH_COUNT(A, O)                                                                                   \
	do {                                                                                                        \
		st = rb_prepare_kernel(rb_k_count<A, O>, smem);                                                         \
		if(st) return st;                                                                                       \
		rb_k_count<A, O><<<grid, RB_TRACE_BLOCK, smem, s>>>(rc->dev, d_origins, d_directions, mini, maxi, nrays, d_c); \
	} while(0)
		if(any_hit && ordered) RB_LAUNCH_COUNT(true, true);
		else if(any_hit) RB_LAUNCH_COUNT(true, false);
		else if(ordered) RB_LAUNCH_COUNT(false, true);
		else RB_LAUNCH_COUNT(false, false);
#undef RB_LAUNCH_COUNT
		RB_CUDA(cudaGetLastError());
		rc->launches++;
	}
	unsigned long long h[4];
	RB_CUDA(cudaMemcpyAsync(h, d_c, sizeof(h), cudaMemcpyDeviceToHost, s));
	RB_CUDA(cudaStreamSynchronize(s));
	RB_CUDA(cudaFreeAsync(d_c, s));
	for(int k = 0; k < 4; ++k) counters[k] = h[k];
	return RB_OK;
}

// Host-pointer entry points: stage through stream-ordered allocations on the calling thread's own stream, so that
// concurrent callers (the reference traces from many threads) never share scratch memory.
namespace {
struct Scratch {
	cudaStream_t s;
	std::vector<void *> ptrs;
	explicit Scratch(cudaStream_t st) : s(st) {}
	~Scratch() {
		for(void * p : ptrs) cudaFreeAsync(p, s);
	}
	template<typename T>
	cudaError_t alloc(T ** p, size_t bytes) {
		void * raw		= nullptr;
		cudaError_t e = cudaMallocAsync(&raw, bytes ? bytes : 16, s);
		if(e == cudaSuccess) ptrs.push_back(raw);
		*p = static_cast<T *>(raw);
		return e;
	}
};
} // namespace

// Host-pointer batches are cut into chunks that flow through a small ring of device buffers, each slot on its own
// stream, so that the upload of chunk c+1, the traversal of chunk c and the download of chunk c-1 overlap (two copy
// engines + SMs). The ring belongs to the calling host thread: concurrent callers never share buffers or streams.
namespace {
constexpr int kPipeSlotsMax = 8;
// Rays per chunk and slots in the ring (RB_PIPE_CHUNK / RB_PIPE_SLOTS override them, for tuning).
static const size_t kPipeChunk = [] { const char * e = getenv("RB_PIPE_CHUNK"); const long v = e ? atol(e) : 0; return v >= 1024 ? size_t(v) : (size_t(1) << 19); }();
static const int kPipeSlots	   = [] { const char * e = getenv("RB_PIPE_SLOTS"); const int v = e ? atoi(e) : 0; return (v >= 1 && v <= kPipeSlotsMax) ? v : 3; }();

struct HostPipe {
	int device = -1;
	cudaStream_t stream[kPipeSlotsMax] = {};
	cudaEvent_t done[kPipeSlotsMax]	= {};
	float * d_org[kPipeSlotsMax]		= {};
	float * d_dir[kPipeSlotsMax]		= {};
	void * d_out[kPipeSlotsMax]		= {};
	bool ready = false;
	size_t nextChunk = 0; // chunks rotate through the slots across calls, so that asynchronous calls queue up behind each other

	rb_status ensure(int dev) {
		if(ready && dev == device) return RB_OK;
		release();
		device = dev;
		const rb_status st = create();
		if(st) release(); // a ring that came up half way (out of memory at slot 2) is taken down again, nothing leaks
		else ready = true;
		return st;
	}
	rb_status create() {
		for(int k = 0; k < kPipeSlots; ++k) {
			RB_CUDA(cudaStreamCreateWithFlags(&stream[k], cudaStreamNonBlocking));
			RB_CUDA(cudaEventCreateWithFlags(&done[k], cudaEventDisableTiming));
			RB_CUDA(cudaMalloc(&d_org[k], kPipeChunk * 3 * sizeof(float)));
			RB_CUDA(cudaMalloc(&d_dir[k], kPipeChunk * 3 * sizeof(float)));
			RB_CUDA(cudaMalloc(&d_out[k], kPipeChunk * sizeof(rb_hit)));
		}
		return RB_OK;
	}
	void release() {
		for(int k = 0; k < kPipeSlotsMax; ++k) {
			if(stream[k]) cudaStreamDestroy(stream[k]);
			if(done[k]) cudaEventDestroy(done[k]);
			cudaFree(d_org[k]);
			cudaFree(d_dir[k]);
			cudaFree(d_out[k]);
			stream[k] = nullptr; done[k] = nullptr; d_org[k] = d_dir[k] = nullptr; d_out[k] = nullptr;
		}
		ready = false;
	}
	~HostPipe() {
		// The CUDA context may already be gone at thread/process exit; leaking a few buffers there is harmless.
	}
};

thread_local HostPipe t_pipe;

template<bool ANY>
rb_status rb_host_batch(const rb_raycaster * rc, const float * origins, const float * directions, float mini, float maxi, size_t nrays, void * out, bool wait = true) {
	RB_CUDA(cudaSetDevice(rc->device));
	rb_status st = t_pipe.ensure(rc->device);
	if(st) return st;
	const size_t outStride = ANY ? sizeof(uint8_t) : sizeof(rb_hit);
	for(size_t begin = 0; begin < nrays; begin += kPipeChunk) {
		const int k = int(t_pipe.nextChunk++ % kPipeSlots);
		const size_t n	= (nrays - begin < kPipeChunk) ? (nrays - begin) : kPipeChunk;
		cudaStream_t s	= t_pipe.stream[k];
		// The slot's previous chunk must have left the device before its buffers are overwritten: same stream, so
		// stream order already guarantees it.
		RB_CUDA(cudaMemcpyAsync(t_pipe.d_org[k], origins + 3 * begin, n * 3 * sizeof(float), cudaMemcpyHostToDevice, s));
		RB_CUDA(cudaMemcpyAsync(t_pipe.d_dir[k], directions + 3 * begin, n * 3 * sizeof(float), cudaMemcpyHostToDevice, s));
		st = rb_launch_trace<ANY>(rc, t_pipe.d_org[k], t_pipe.d_dir[k], mini, maxi, n, t_pipe.d_out[k], s);
		if(st) return st;
		RB_CUDA(cudaMemcpyAsync(static_cast<uint8_t *>(out) + begin * outStride, t_pipe.d_out[k], n * outStride, cudaMemcpyDeviceToHost, s));
	}
	if(wait) {
		for(int k = 0; k < kPipeSlots; ++k) RB_CUDA(cudaStreamSynchronize(t_pipe.stream[k]));
	}
	return RB_OK;
}
} // namespace

// Asynchronous forms: the call returns once its chunks are queued on the calling thread's ring; results are complete after
// rb_raycaster_sync from the SAME thread. Consecutive asynchronous calls queue up behind each other in call order, so a caller
// that issues its download-heavy queries first (closest hit: 48 bytes per ray come back) and its upload-only ones last (any
// hit: 1 byte) keeps both directions of the host link busy from one thread — the four calls of the C2 step overlap into
// max(up, down) instead of paying for each call's own ramp-up and drain. Host buffers must stay valid (and pinned, for the
// copies to be asynchronous at all) until the sync.
extern "C" rb_status rb_raycaster_intersect_async(const rb_raycaster * rc, const float * origins, const float * directions, float mini, float maxi, size_t nrays, rb_hit * out) {
	rb_status st = rb_check_query(rc, "rb_raycaster_intersect_async");
	if(st) return st;
	if(nrays == 0) return RB_OK;
	if(!origins || !directions || !out) return rb_fail(RB_ERR_INVALID, "rb_raycaster_intersect_async: null argument");
	return rb_host_batch<false>(rc, origins, directions, mini, maxi, nrays, out, false);
}

extern "C" rb_status rb_raycaster_intersect_any_async(const rb_raycaster * rc, const float * origins, const float * directions, float mini, float maxi, size_t nrays, uint8_t * out) {
	rb_status st = rb_check_query(rc, "rb_raycaster_intersect_any_async");
	if(st) return st;
	if(nrays == 0) return RB_OK;
	if(!origins || !directions || !out) return rb_fail(RB_ERR_INVALID, "rb_raycaster_intersect_any_async: null argument");
	return rb_host_batch<true>(rc, origins, directions, mini, maxi, nrays, out, false);
}

extern "C" rb_status rb_raycaster_sync(const rb_raycaster * rc) {
	if(!rc) return rb_fail(RB_ERR_INVALID, "rb_raycaster_sync: null handle");
	RB_REQUIRE_DEVICE();
	if(!t_pipe.ready) return RB_OK; // nothing was ever queued from this thread
	RB_CUDA(cudaSetDevice(t_pipe.device));
	for(int k = 0; k < kPipeSlots; ++k) RB_CUDA(cudaStreamSynchronize(t_pipe.stream[k]));
	return RB_OK;
}

// Primary rays of a pinhole camera, generated on the device: pixel (x, y) -> the ray from `eye` through the centre of the
// pixel, with the reference's own arithmetic (PathTracer.cpp:193-198: ndc = screenPos / (W, H); worldPos = corner +
// ndc.x * dx + ndc.y * dy; direction = worldPos - eye, normalised by Ray::Ray). Nothing goes up the host link but the camera.
__global__ void rb_k_primary_rays(rb_camera cam, uint32_t width, uint32_t height, size_t first, size_t n, float * __restrict__ org, float * __restrict__ dir) {
	for(size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
		const size_t p	  = first + i;
		const uint32_t x  = uint32_t(p % width), y = uint32_t(p / width);
		const float ndcx  = (float(x) + 0.5f) / float(width);
		const float ndcy  = (float(y) + 0.5f) / float(height);
		const float wx = (cam.corner[0] + ndcx * cam.dx[0]) + ndcy * cam.dy[0];
		const float wy = (cam.corner[1] + ndcx * cam.dx[1]) + ndcy * cam.dy[1];
		const float wz = (cam.corner[2] + ndcx * cam.dx[2]) + ndcy * cam.dy[2];
		org[3 * i + 0] = cam.eye[0]; org[3 * i + 1] = cam.eye[1]; org[3 * i + 2] = cam.eye[2];
		dir[3 * i + 0] = wx - cam.eye[0]; dir[3 * i + 1] = wy - cam.eye[1]; dir[3 * i + 2] = wz - cam.eye[2];
	}
}

namespace {
template<bool ANY>
rb_status rb_host_primary(const rb_raycaster * rc, const rb_camera * cam, uint32_t width, uint32_t height, float mini, float maxi, void * out) {
	RB_CUDA(cudaSetDevice(rc->device));
	rb_status st = t_pipe.ensure(rc->device);
	if(st) return st;
	const size_t nrays = size_t(width) * height, outStride = ANY ? sizeof(uint8_t) : sizeof(rb_hit);
	for(size_t begin = 0; begin < nrays; begin += kPipeChunk) {
		const int k = int(t_pipe.nextChunk++ % kPipeSlots);
		const size_t n = (nrays - begin < kPipeChunk) ? (nrays - begin) : kPipeChunk;
		cudaStream_t s = t_pipe.stream[k];
		rb_k_primary_rays<<<unsigned((n + 255) / 256 < 1184 ? (n + 255) / 256 : 1184), 256, 0, s>>>(*cam, width, height, begin, n, t_pipe.d_org[k], t_pipe.d_dir[k]);
		RB_CUDA(cudaGetLastError());
		rc->launches++;
		st = rb_launch_trace<ANY>(rc, t_pipe.d_org[k], t_pipe.d_dir[k], mini, maxi, n, t_pipe.d_out[k], s);
		if(st) return st;
		RB_CUDA(cudaMemcpyAsync(static_cast<uint8_t *>(out) + begin * outStride, t_pipe.d_out[k], n * outStride, cudaMemcpyDeviceToHost, s));
	}
	for(int k = 0; k < kPipeSlots; ++k) RB_CUDA(cudaStreamSynchronize(t_pipe.stream[k]));
	return RB_OK;
}
} // namespace

extern "C" rb_status rb_raycaster_intersect_primary(const rb_raycaster * rc, const rb_camera * camera, uint32_t width, uint32_t height, float mini, float maxi, rb_hit * out) {
	rb_status st = rb_check_query(rc, "rb_raycaster_intersect_primary");
	if(st) return st;
	if(!camera || !out) return rb_fail(RB_ERR_INVALID, "rb_raycaster_intersect_primary: null argument");
	if(width == 0 || height == 0) return RB_OK;
	return rb_host_primary<false>(rc, camera, width, height, mini, maxi, out);
}

extern "C" rb_status rb_raycaster_intersect_any_primary(const rb_raycaster * rc, const rb_camera * camera, uint32_t width, uint32_t height, float mini, float maxi, uint8_t * out) {
	rb_status st = rb_check_query(rc, "rb_raycaster_intersect_any_primary");
	if(st) return st;
	if(!camera || !out) return rb_fail(RB_ERR_INVALID, "rb_raycaster_intersect_any_primary: null argument");
	if(width == 0 || height == 0) return RB_OK;
	return rb_host_primary<true>(rc, camera, width, height, mini, maxi, out);
}

extern "C" rb_status rb_raycaster_intersect(const rb_raycaster * rc, const float * origins, const float * directions, float mini, float maxi, size_t nrays, rb_hit * out) {
	rb_status st = rb_check_query(rc, "rb_raycaster_intersect");
	if(st) return st;
	if(nrays == 0) return RB_OK;
	if(!origins || !directions || !out) return rb_fail(RB_ERR_INVALID, "rb_raycaster_intersect: null argument");
	return rb_host_batch<false>(rc, origins, directions, mini, maxi, nrays, out);
}

extern "C" rb_status rb_raycaster_intersect_any(const rb_raycaster * rc, const float * origins, const float * directions, float mini, float maxi, size_t nrays, uint8_t * out) {
	rb_status st = rb_check_query(rc, "rb_raycaster_intersect_any");
	if(st) return st;
	if(nrays == 0) return RB_OK;
	if(!origins || !directions || !out) return rb_fail(RB_ERR_INVALID, "rb_raycaster_intersect_any: null argument");
	return rb_host_batch<true>(rc, origins, directions, mini, maxi, nrays, out);
}

extern "C" rb_status rb_raycaster_visible(const rb_raycaster * rc, const float * p0, const float * p1, size_t n, uint8_t * out) {
	rb_status st = rb_check_query(rc, "rb_raycaster_visible");
	if(st) return st;
	if(n == 0) return RB_OK;
	if(!p0 || !p1 || !out) return rb_fail(RB_ERR_INVALID, "rb_raycaster_visible: null argument");
	RB_CUDA(cudaSetDevice(rc->device));
	cudaStream_t s = cudaStreamPerThread;
	Scratch scratch(s);
	float *d_a = nullptr, *d_b = nullptr;
	uint8_t * d_r = nullptr;
	RB_CUDA(scratch.alloc(&d_a, n * 3 * sizeof(float)));
	RB_CUDA(scratch.alloc(&d_b, n * 3 * sizeof(float)));
	RB_CUDA(scratch.alloc(&d_r, n));
	RB_CUDA(cudaMemcpyAsync(d_a, p0, n * 3 * sizeof(float), cudaMemcpyHostToDevice, s));
	RB_CUDA(cudaMemcpyAsync(d_b, p1, n * 3 * sizeof(float), cudaMemcpyHostToDevice, s));
	const size_t smem	= rb_trace_smem_bytes(rc, RB_TRACE_BLOCK);
	const unsigned grid = unsigned((n + RB_TRACE_BLOCK - 1) / RB_TRACE_BLOCK);
	st = rb_prepare_kernel(rb_k_visible, smem);
	if(st) return st;
	rb_k_visible<<<grid, RB_TRACE_BLOCK, smem, s>>>(rc->dev, d_a, d_b, n, d_r);
	RB_CUDA(cudaGetLastError());
	rc->launches++;
	RB_CUDA(cudaMemcpyAsync(out, d_r, n, cudaMemcpyDeviceToHost, s));
	RB_CUDA(cudaStreamSynchronize(s));
	return RB_OK;
}

// ---------------------------------------------------------------- ray debugging: visited nodes of one ray
// RaycasterVisualisation::getRayLevels (src/engine/raycaster/RaycasterVisualisation.cpp:41-77) on the device tree: the
// closest-hit loop in the reference's order, recording every node it pops — its box and depth (roots are depth 0) — so
// that a viewer can draw the levels a ray went through without reading the hierarchy back. On the device a node's box
// is stored with its parent (or in the root table), so the stack carries where to find it. One thread: this is a
// click-to-inspect query, not a throughput path.
struct RbLevelEntry {
	int ref, depth, src; // src: (parent interior << 1) | side, or ~root slot
};

__global__ void rb_k_ray_levels(RbDeviceBvh bvh, const float * __restrict__ org, const float * __restrict__ dir, float mini, float maxi, RbLevelEntry * __restrict__ stack, unsigned int capacity,
								float * __restrict__ boxes, unsigned int * __restrict__ depths, unsigned int * __restrict__ count, rb_hit * __restrict__ hit) {
	if(blockIdx.x != 0 || threadIdx.x != 0) return;
	const RbRay ray = rb_make_ray(V3{org[0], org[1], org[2]}, V3{dir[0], dir[1], dir[2]});
	int top = 0;
	for(int m = 0; m < bvh.nroots; ++m) {
		const float4 lo = bvh.roots[2 * m], hi = bvh.roots[2 * m + 1];
		if(!rb_box_test(ray, lo.x, lo.y, lo.z, hi.x, hi.y, hi.z, mini, maxi)) continue;
		stack[top++] = RbLevelEntry{__float_as_int(lo.w), 0, ~m};
	}
	RbHitRec best{FLT_MAX, 0.0f, 0.0f, -1};
	unsigned int visited = 0;
	while(top > 0) {
		const RbLevelEntry e = stack[--top];
		if(visited < capacity) {
			float * b = boxes + 6 * (size_t)visited;
			if(e.src < 0) {
				const float4 lo = bvh.roots[2 * (~e.src)], hi = bvh.roots[2 * (~e.src) + 1];
				b[0] = lo.x; b[1] = lo.y; b[2] = lo.z; b[3] = hi.x; b[4] = hi.y; b[5] = hi.z;
			} else {
				const float * n = reinterpret_cast<const float *>(bvh.nodes + 4 * (size_t)(e.src >> 1)) + 6 * (e.src & 1);
				for(int k = 0; k < 6; ++k) b[k] = n[k];
			}
			depths[visited] = (unsigned int)e.depth;
		}
		++visited;
		if(e.ref == RB_REF_NONE) continue; // empty leaf
		if(e.ref < 0) {
			const int code = ~e.ref, first = code >> RB_LEAF_BITS, n = (code & (RB_LEAF_MAX_TRIS - 1)) + 1;
			for(int k = 0; k < n; ++k) {
				const float4 * tp = bvh.tris + 3 * (size_t)(first + k);
				const float4 t0 = tp[0], t1 = tp[1], t2 = tp[2];
				float t, u, v;
				if(rb_tri_test(ray, V3{t0.x, t0.y, t0.z}, V3{t1.x, t1.y, t1.z}, V3{t2.x, t2.y, t2.z}, mini, maxi, t, u, v) && t < best.t) {
					best = RbHitRec{t, u, v, first + k};
					maxi = t;
				}
			}
			continue;
		}
		const float4 * np = bvh.nodes + 4 * (size_t)e.ref;
		const float4 a = np[0], b4 = np[1], c = np[2], d4 = np[3];
		if(rb_box_test(ray, a.x, a.y, a.z, a.w, b4.x, b4.y, mini, maxi)) stack[top++] = RbLevelEntry{__float_as_int(d4.x), e.depth + 1, (e.ref << 1) | 0};
		if(rb_box_test(ray, b4.z, b4.w, c.x, c.y, c.z, c.w, mini, maxi)) stack[top++] = RbLevelEntry{__float_as_int(d4.y), e.depth + 1, (e.ref << 1) | 1};
	}
	*count = visited;
	rb_store_hit(hit, 0, bvh, best);
}

extern "C" rb_status rb_raycaster_ray_levels(const rb_raycaster * rc, const float * origin, const float * direction, float mini, float maxi, uint32_t capacity, float * boxes6, uint32_t * depths,
											 uint32_t * count, rb_hit * hit) {
	rb_status st = rb_check_query(rc, "rb_raycaster_ray_levels");
	if(st) return st;
	if(!origin || !direction || !count || !hit || (capacity && (!boxes6 || !depths))) return rb_fail(RB_ERR_INVALID, "rb_raycaster_ray_levels: null argument");
	RB_CUDA(cudaSetDevice(rc->device));
	cudaStream_t s = cudaStreamPerThread;
	Scratch scratch(s);
	float *d_ray = nullptr, *d_boxes = nullptr;
	unsigned int *d_depths = nullptr, *d_count = nullptr;
	RbLevelEntry * d_stack = nullptr;
	rb_hit * d_hit = nullptr;
	const size_t stackEntries = size_t(rc->dev.nroots) + size_t(rc->dev.stackDepth) + 2;
	RB_CUDA(scratch.alloc(&d_ray, 6 * sizeof(float)));
	RB_CUDA(scratch.alloc(&d_boxes, size_t(capacity) * 6 * sizeof(float)));
	RB_CUDA(scratch.alloc(&d_depths, size_t(capacity) * sizeof(unsigned int)));
	RB_CUDA(scratch.alloc(&d_count, sizeof(unsigned int)));
	RB_CUDA(scratch.alloc(&d_stack, stackEntries * sizeof(RbLevelEntry)));
	RB_CUDA(scratch.alloc(&d_hit, sizeof(rb_hit)));
	const float ray[6] = {origin[0], origin[1], origin[2], direction[0], direction[1], direction[2]};
	RB_CUDA(cudaMemcpyAsync(d_ray, ray, sizeof(ray), cudaMemcpyHostToDevice, s));
	rb_k_ray_levels<<<1, 32, 0, s>>>(rc->dev, d_ray, d_ray + 3, mini, maxi, d_stack, capacity, d_boxes, d_depths, d_count, d_hit);
	RB_CUDA(cudaGetLastError());
	rc->launches++;
	unsigned int visited = 0;
	RB_CUDA(cudaMemcpyAsync(&visited, d_count, sizeof(visited), cudaMemcpyDeviceToHost, s));
	RB_CUDA(cudaMemcpyAsync(hit, d_hit, sizeof(rb_hit), cudaMemcpyDeviceToHost, s));
	RB_CUDA(cudaStreamSynchronize(s));
	const size_t n = visited < capacity ? visited : capacity;
	if(n) {
		RB_CUDA(cudaMemcpyAsync(boxes6, d_boxes, n * 6 * sizeof(float), cudaMemcpyDeviceToHost, s));
		RB_CUDA(cudaMemcpyAsync(depths, d_depths, n * sizeof(unsigned int), cudaMemcpyDeviceToHost, s));
		RB_CUDA(cudaStreamSynchronize(s));
	}
	*count = visited;
	return RB_OK;
}

// ---------------------------------------------------------------- KAT: Ray::Ray + Intersection::box on arrays
__global__ void rb_k_kat_box(const float * __restrict__ org, const float * __restrict__ dir, const float * __restrict__ boxes, float mini, float maxi, size_t n, uint8_t * __restrict__ out) {
	const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	if(i >= n) return;
	V3 o, d;
	rb_load_ray(org, dir, i, o, d);
	const RbRay ray = rb_make_ray(o, d);
	const float * b = boxes + 6 * i;
	out[i] = rb_box_test(ray, b[0], b[1], b[2], b[3], b[4], b[5], mini, maxi) ? 1 : 0;
}

extern "C" rb_status rb_kat_box_test(const float * origins, const float * dirs, const float * boxes6, float mini, float maxi, size_t n, uint8_t * out) {
	RB_REQUIRE_DEVICE();
	if(n == 0) return RB_OK;
	if(!origins || !dirs || !boxes6 || !out) return rb_fail(RB_ERR_INVALID, "rb_kat_box_test: null argument");
	cudaStream_t s = cudaStreamPerThread;
	Scratch scratch(s);
	float *d_o = nullptr, *d_d = nullptr, *d_b = nullptr;
	uint8_t * d_r = nullptr;
	RB_CUDA(scratch.alloc(&d_o, n * 3 * sizeof(float)));
	RB_CUDA(scratch.alloc(&d_d, n * 3 * sizeof(float)));
	RB_CUDA(scratch.alloc(&d_b, n * 6 * sizeof(float)));
	RB_CUDA(scratch.alloc(&d_r, n));
	RB_CUDA(cudaMemcpyAsync(d_o, origins, n * 3 * sizeof(float), cudaMemcpyHostToDevice, s));
	RB_CUDA(cudaMemcpyAsync(d_d, dirs, n * 3 * sizeof(float), cudaMemcpyHostToDevice, s));
	RB_CUDA(cudaMemcpyAsync(d_b, boxes6, n * 6 * sizeof(float), cudaMemcpyHostToDevice, s));
	rb_k_kat_box<<<unsigned((n + 127) / 128), 128, 0, s>>>(d_o, d_d, d_b, mini, maxi, n, d_r);
	RB_CUDA(cudaGetLastError());
	RB_CUDA(cudaMemcpyAsync(out, d_r, n, cudaMemcpyDeviceToHost, s));
	RB_CUDA(cudaStreamSynchronize(s));
	return RB_OK;
}

// Raycaster::intersects(const Ray &, const TriangleInfos &, mini, maxi) (Raycaster.cpp:232-262) for arrays of (ray, triangle)
// pairs: the per-triangle Möller-Trumbore test on its own. `dirs` are used AS GIVEN (a Ray's already-normalised `dir`):
// normalising a second time can move a component by one ulp. out4 = (hit ? 1 : 0, t, u, v) per pair. It is what the
// drop-in's private per-triangle intersects — called by RaycasterVisualisation::getRayLevels on the leaves one ray visits
// (RaycasterVisualisation.cpp:61-66) — runs on, and a known-answer test of the triangle test.
__global__ void rb_k_kat_tri(const float * __restrict__ pos, const float * __restrict__ dir, const float * __restrict__ v9, float mini, float maxi, size_t n, float * __restrict__ out) {
	const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	if(i >= n) return;
	RbRay ray;
	ray.pos = V3{pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]};
	ray.dir = V3{dir[3 * i], dir[3 * i + 1], dir[3 * i + 2]};
	const float * v = v9 + 9 * i;
	const V3 v0 = V3{v[0], v[1], v[2]}, v1 = V3{v[3], v[4], v[5]}, v2 = V3{v[6], v[7], v[8]};
	float t = 0.0f, u = 0.0f, w = 0.0f;
	const bool hit = rb_tri_test(ray, v0, v1 - v0, v2 - v0, mini, maxi, t, u, w);
	out[4 * i + 0] = hit ? 1.0f : 0.0f;
	out[4 * i + 1] = hit ? t : 0.0f;
	out[4 * i + 2] = hit ? u : 0.0f;
	out[4 * i + 3] = hit ? w : 0.0f;
}

extern "C" rb_status rb_kat_triangle_test(const float * positions, const float * dirs, const float * vertices9, float mini, float maxi, size_t n, float * out4) {
	RB_REQUIRE_DEVICE();
	if(n == 0) return RB_OK;
	if(!positions || !dirs || !vertices9 || !out4) return rb_fail(RB_ERR_INVALID, "rb_kat_triangle_test: null argument");
	cudaStream_t s = cudaStreamPerThread;
	Scratch scratch(s);
	float *d_o = nullptr, *d_d = nullptr, *d_v = nullptr, *d_r = nullptr;
	RB_CUDA(scratch.alloc(&d_o, n * 3 * sizeof(float)));
	RB_CUDA(scratch.alloc(&d_d, n * 3 * sizeof(float)));
	RB_CUDA(scratch.alloc(&d_v, n * 9 * sizeof(float)));
	RB_CUDA(scratch.alloc(&d_r, n * 4 * sizeof(float)));
	RB_CUDA(cudaMemcpyAsync(d_o, positions, n * 3 * sizeof(float), cudaMemcpyHostToDevice, s));
	RB_CUDA(cudaMemcpyAsync(d_d, dirs, n * 3 * sizeof(float), cudaMemcpyHostToDevice, s));
	RB_CUDA(cudaMemcpyAsync(d_v, vertices9, n * 9 * sizeof(float), cudaMemcpyHostToDevice, s));
	rb_k_kat_tri<<<unsigned((n + 127) / 128), 128, 0, s>>>(d_o, d_d, d_v, mini, maxi, n, d_r);
	RB_CUDA(cudaGetLastError());
	RB_CUDA(cudaMemcpyAsync(out4, d_r, n * 4 * sizeof(float), cudaMemcpyDeviceToHost, s));
	RB_CUDA(cudaStreamSynchronize(s));
	return RB_OK;
}

// ---------------------------------------------------------------- when to stop re-fitting
// A refit keeps the topology, so after large motions boxes of unrelated triangles overlap and rays enter more of them.
// The measure is the surface-area cost of the tree the traversal actually walks: the sum, over interior nodes, of the
// areas of their two child boxes (the expected number of child boxes a random line enters is proportional to it). It is
// taken after every build (the baseline) and can be taken again after a refit; the ratio says how much more box-entering
// work the re-fitted tree asks for than when it was built.
__global__ void rb_k_tree_area(const float4 * __restrict__ nodes, size_t ninterior, double * __restrict__ out) {
	double acc = 0.0;
	for(size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < ninterior; i += (size_t)gridDim.x * blockDim.x) {
		const float4 n0 = nodes[4 * i], n1 = nodes[4 * i + 1], n2 = nodes[4 * i + 2];
		const float lx = n0.w - n0.x, ly = n1.x - n0.y, lz = n1.y - n0.z;
		const float rx = n2.y - n1.z, ry = n2.z - n1.w, rz = n2.w - n2.x;
		// empty / inverted boxes (never produced by the fit kernel for real children) count as zero
		if(lx >= 0.0f && ly >= 0.0f && lz >= 0.0f) acc += double(lx) * ly + double(ly) * lz + double(lz) * lx;
		if(rx >= 0.0f && ry >= 0.0f && rz >= 0.0f) acc += double(rx) * ry + double(ry) * rz + double(rz) * rx;
	}
	for(int off = 16; off > 0; off >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, off);
	if((threadIdx.x & 31u) == 0) atomicAdd(out, acc);
}

rb_status rb_tree_area(const rb_raycaster * rc, double * area) {
	*area = 0.0;
	if(rc->nInterior == 0) return RB_OK;
	RB_CUDA(cudaSetDevice(rc->device));
	cudaStream_t s = cudaStreamPerThread;
	Scratch scratch(s);
	double * d_a = nullptr;
	RB_CUDA(scratch.alloc(&d_a, sizeof(double)));
	RB_CUDA(cudaMemsetAsync(d_a, 0, sizeof(double), s));
	rb_k_tree_area<<<592, 256, 0, s>>>(rc->dev.nodes, rc->nInterior, d_a);
	RB_CUDA(cudaGetLastError());
	RB_CUDA(cudaMemcpyAsync(area, d_a, sizeof(double), cudaMemcpyDeviceToHost, s));
	RB_CUDA(cudaStreamSynchronize(s));
	return RB_OK;
}

extern "C" rb_status rb_raycaster_tree_cost(rb_raycaster * rc, double out[2]) {
	rb_status st = rb_check_query(rc, "rb_raycaster_tree_cost");
	if(st) return st;
	if(!out) return rb_fail(RB_ERR_INVALID, "rb_raycaster_tree_cost: null output");
	if(rc->areaSerial != rc->buildSerial) { // first question since this build: the tree as built IS the baseline, unless it was re-fitted already
		st = rb_tree_area(rc, &rc->areaAtBuild);
		if(st) return st;
		rc->areaSerial = rc->buildSerial;
		rc->areaNow	   = rc->areaAtBuild;
		rc->areaStale  = false;
	}
	if(rc->areaStale) {
		st = rb_tree_area(rc, &rc->areaNow);
		if(st) return st;
		rc->areaStale = false;
	}
	out[0] = rc->areaAtBuild;
	out[1] = rc->areaNow;
	return RB_OK;
}

