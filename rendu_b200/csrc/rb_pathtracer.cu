// rb_pathtracer.cu — PathTracer::render (reference: src/apps/pathtracer/PathTracer.cpp:159-299) as a wavefront of
// sm_100a kernels: ray generation, closest-hit traversal, shading (+ NEE set-up + BSDF sampling), shadow traversal,
// per-pixel accumulation and resolve. One path = one (pixel, stratification cell) pair; its state lives in SoA float4
// arrays in HBM between kernels, and the two traversal kernels are persistent with a warp-level work queue.
//
// The estimator is the reference's, term for term (SURVEY.md §7 "quirk fidelity"): stratified jittered pixel samples
// (:57-84), one uniformly chosen light per bounce with NEE at every depth including the last (:248-271), lobe-mixture
// BSDF sampling without MIS or Russian roulette (MaterialGGX.cpp:54-92), masked hits that consume a depth level
// (:223-226), emissive hits that add the raw texel and stop (:228-234), background only on escape (:206-209), the
// per-SAMPLE clamp min(sampleColor, 5) before accumulation (:288) and the final division by the sample count (:291).
// Two things differ by design (north_star): the RNG is counter-based (rb_shade.cuh) and a path whose attenuation has
// become exactly zero is dropped instead of being traced further (everything it could still add is multiplied by 0).
#include "rb_scene.h"
#include "rb_trace.cuh"
#include "rb_wide.cuh"

#include <cmath>
#include <cstring>

#define RB_PT_BLOCK 128
#ifndef RB_PT_SHADE_MIN_BLOCKS
#define RB_PT_SHADE_MIN_BLOCKS 7 // caps the shade kernel at 72 registers (10 bytes of spills): it is latency bound, occupancy matters
								 // more; same-box A/B of the frame: 5 blocks (96 regs) -0.4 %, 6 (80) 0, 7 (72) +1.2 %, 8 (64, 58 B spilled) -1.6 %
#endif

struct RbFrame {
	float eye[3], corner[3], dx[3], dy[3];
	uint32_t width, height;
	uint32_t samples;		 // full (rounded) sample count: grid + normalisation
	uint32_t cellsX, cellsY; // getSampleGrid
	float cellSizeX, cellSizeY;
	uint32_t depth;
	uint64_t seed;
	uint32_t tileRows, tileIndex, tileCount; // rows rendered by this call
	uint32_t sampleIndex, sampleCount;		 // stratification cells traced by this call
	uint32_t localRows;						 // number of rows selected by the tile split
	uint64_t mWidth, mTileRows, mCellsX;	 // rb_div_magic of width, tileRows, cellsX
};

struct RbBatch {
	uint32_t pixel0;  // first local pixel (row-major over the selected rows)
	uint32_t npixels; // pixels in this batch
	uint32_t sample0; // first local sample
	uint32_t schunk;  // samples per pixel in this batch
	uint64_t mSchunk; // rb_div_magic(schunk)
};

struct RbPathBuffers {
	float4 * rayO;
	float4 * rayD;
	float4 * atten;
	float4 * color;
	float4 * hit;
	uint32_t * queue[2];
	float4 * shO; // origin, maxDist
	float4 * shD; // direction, slot
	float4 * shC; // contribution
	uint32_t * counters; // see RbCounterLayout
};

// The traversal wavefronts read their rays and write their hits with the evict-first policy (ld.global.cs / st.global.cs):
// that traffic is read once, and it should not push the hierarchy out of the L2 (closest-hit wavefront -2.4 %). The shade
// kernel keeps plain accesses: with the same hints on its state it runs 5 % slower (profiles/r02_shade_diet.txt). An
// access-policy window that pins the interior nodes in a persisting part of the L2 was measured too: +0.6 % at 96 MiB,
// -20 % at 64 MiB (the set-aside starves the shading records and texels) — not kept.
template<typename T> RB_D T rb_ld(const T * p) { return __ldcs(p); }
template<typename T> RB_D void rb_st(T * p, T v) { __stcs(p, v); }

// Device counters of one batch (zeroed once per batch): queue sizes and work-queue cursors per bounce.
RB_HD uint32_t rb_cnt_queue(uint32_t did) { return 4 * did + 0; }	// paths entering bounce `did`
RB_HD uint32_t rb_cnt_shadow(uint32_t did) { return 4 * did + 1; }	// shadow rays of bounce `did`
RB_HD uint32_t rb_cnt_cursor_t(uint32_t did) { return 4 * did + 2; }
RB_HD uint32_t rb_cnt_cursor_s(uint32_t did) { return 4 * did + 3; }

// Local row -> image row for the tile split (rows come in bands of tileRows, dealt round-robin to tileCount ranks).
RB_HD uint32_t rb_tile_row(const RbFrame & f, uint32_t localRow) {
	if(f.tileCount <= 1) return localRow;
	const uint32_t band = rb_div(localRow, f.mTileRows);
	return (band * f.tileCount + f.tileIndex) * f.tileRows + (localRow - band * f.tileRows);
}

struct RbPathId {
	uint32_t x, y, pixel, sid;
};

RB_D RbPathId rb_path_id(const RbFrame & f, const RbBatch & b, uint32_t slot) {
	const uint32_t q	  = rb_div(slot, b.mSchunk);
	const uint32_t p	  = b.pixel0 + q;
	const uint32_t s	  = b.sample0 + (slot - q * b.schunk);
	const uint32_t lrow = rb_div(p, f.mWidth);
	RbPathId id;
	id.x	 = p - lrow * f.width;
	id.y	 = rb_tile_row(f, lrow);
	id.pixel = id.y * f.width + id.x;
	id.sid	 = f.sampleIndex + s * f.sampleCount;
	return id;
}

// getSamplePosition (PathTracer.cpp:71-84) + the screen -> NDC step of render (:191-193).
RB_D void rb_sample_ndc(const RbFrame & f, const RbPathId & id, uint64_t key, float & ndcx, float & ndcy) {
	const int sidy = int(rb_div(id.sid, f.mCellsX)); // sid and cellsX are non-negative: same quotient as the int division
	const int sidx = int(id.sid) - sidy * int(f.cellsX);
	const float jx = rb_rng_float(key, 0);
	const float jy = rb_rng_float(key, 1);
	const float gx = float(sidx) + jx, gy = float(sidy) + jy;
	const float lx = gx * f.cellSizeX, ly = gy * f.cellSizeY;
	const float sx = float(id.x) + lx, sy = float(id.y) + ly;
	ndcx = sx / float(f.width);
	ndcy = sy / float(f.height);
}

// ------------------------------------------------------------------------------------------------ ray generation
__global__ void __launch_bounds__(RB_PT_BLOCK) rb_k_pt_raygen(RbFrame f, RbBatch b, RbPathBuffers pb) {
	const uint32_t n = b.npixels * b.schunk;
	for(uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x; slot < n; slot += gridDim.x * blockDim.x) {
		const RbPathId id = rb_path_id(f, b, slot);
		const uint64_t key = rb_rng_key(f.seed, id.pixel, id.sid);
		float ndcx, ndcy;
		rb_sample_ndc(f, id, key, ndcx, ndcy);
		const V3 corner = V3{f.corner[0], f.corner[1], f.corner[2]}, dx = V3{f.dx[0], f.dx[1], f.dx[2]}, dy = V3{f.dy[0], f.dy[1], f.dy[2]};
		const V3 eye	= V3{f.eye[0], f.eye[1], f.eye[2]};
		const V3 worldPos = corner + ndcx * dx + ndcy * dy; // PathTracer.cpp:195
		const V3 rayDir	  = normalize3(worldPos - eye);		 // :198
		pb.rayO[slot] = make_float4(eye.x, eye.y, eye.z, 0.0f);
		pb.rayD[slot] = make_float4(rayDir.x, rayDir.y, rayDir.z, 0.0f);
		// The path's RNG key rides in the two spare words of its state (attenuation.w, colour.w): shading reads and writes
		// those vectors anyway, and deriving the key again at every bounce (slot -> pixel, sample -> two hash rounds) was
		// 7 % of the shade kernel's instructions.
		pb.atten[slot] = make_float4(1.0f, 1.0f, 1.0f, __uint_as_float(uint32_t(key)));
		pb.color[slot] = make_float4(0.0f, 0.0f, 0.0f, __uint_as_float(uint32_t(key >> 32)));
		pb.queue[0][slot] = slot;
	}
	if(blockIdx.x == 0 && threadIdx.x == 0) pb.counters[rb_cnt_queue(0)] = n;
}

// ------------------------------------------------------------------------------------------------ closest-hit wavefront
// `_raycaster.intersects(rayPos, rayDir)` with the reference's default bounds (PathTracer.cpp:204, Raycaster.hpp:65).
// COUNT = true (rb_pathtracer_profile_frame only, never the timed path): the same kernel also counts the interior nodes,
// triangles and roots it tests — the figures the roofline's algorithmic bytes are computed from — into stats[base..base+2].
RB_D void rb_pt_flush_work(const RbWork & w, unsigned int lane, unsigned long long * stats, int base) {
	unsigned int a = w.interior, b = w.tris, c = w.roots;
	for(int off = 16; off > 0; off >>= 1) {
		a += __shfl_down_sync(0xffffffffu, a, off);
		b += __shfl_down_sync(0xffffffffu, b, off);
		c += __shfl_down_sync(0xffffffffu, c, off);
	}
	if(lane == 0) {
		atomicAdd(stats + base, (unsigned long long)a);
		atomicAdd(stats + base + 1, (unsigned long long)b);
		atomicAdd(stats + base + 2, (unsigned long long)c);
	}
}

template<bool ORDERED, int ROOTS, bool COUNT = false>
__global__ void __launch_bounds__(RB_PT_BLOCK) rb_k_pt_trace(RbDeviceBvh bvh, RbPathBuffers pb, const uint32_t * __restrict__ queue, uint32_t did, int threshold, int leafVotes, int nodeSteps, unsigned long long * __restrict__ stats) {
	extern __shared__ int s_stack[];
	const RbStack st		= rb_make_stack(s_stack + threadIdx.x + (ROOTS == RB_ROOTS_TREE_BIG ? bvh.rootWords * blockDim.x : 0), blockDim.x);
	const unsigned int lane = threadIdx.x & 31u;
	const unsigned int n	= pb.counters[rb_cnt_queue(did)];
	unsigned int * cursor	= pb.counters + rb_cnt_cursor_t(did);
	if(blockIdx.x == 0 && threadIdx.x == 0 && stats) atomicAdd(stats + 0, (unsigned long long)n);
	RbWork work = {0u, 0u, 0u, 0u};
	RbWork * const wp = COUNT ? &work : nullptr;
	bool active = false, exhausted = false;
	unsigned int slot = 0;
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
						slot = rb_ld(queue + idx);
						const float4 o = rb_ld(pb.rayO + slot), d = rb_ld(pb.rayD + slot);
						ray = rb_make_ray(V3{o.x, o.y, o.z}, V3{d.x, d.y, d.z});
						rb_traversal_init(T, bvh, 0.0001f, 1e8f, st);
						if(ROOTS == RB_ROOTS_MASK) T.root = int(rb_roots_pretest<COUNT>(bvh, ray, 0.0001f, 1e8f, wp));
						if(ROOTS == RB_ROOTS_TREE) rb_roots_tree_pretest<COUNT>(T, bvh, ray, st, 0.0001f, 1e8f, wp);
						if(ROOTS == RB_ROOTS_TREE_BIG) rb_roots_tree_pretest_big<COUNT>(T, bvh, ray, st, 0.0001f, 1e8f, wp);
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
			// Phase vote (see rb_k_trace_phased): leaf phase once `leafVotes` lanes wait for it, otherwise up to
			// `nodeSteps` node steps.
			const bool wantsLeaf	 = active && rb_wants_leaf(T);
			const unsigned int leafM = __ballot_sync(0xffffffffu, wantsLeaf);
			const unsigned int nodeM = activeM & ~leafM;
			bool finished = false;
			if(__popc(leafM) >= leafVotes || nodeM == 0u) {
				if(wantsLeaf) finished = rb_phase_leaf<false, COUNT>(T, bvh, ray, wp);
			} else if(active && !wantsLeaf) {
				for(int k = 0; k < nodeSteps; ++k) {
					finished = rb_phase_node<COUNT, ORDERED, ROOTS>(T, bvh, ray, st, wp);
					if(finished || rb_wants_leaf(T)) break;
				}
			}
			if(finished) {
				rb_st(pb.hit + slot, make_float4(T.best.t, T.best.u, T.best.v, __int_as_float(T.best.tri)));
				active = false;
			}
			activeM = __ballot_sync(0xffffffffu, active);
			more	= exhausted ? (activeM != 0u) : (__popc(activeM) >= threshold);
		} while(more);
	}
	if(COUNT && stats) rb_pt_flush_work(work, lane, stats, 4);
}

// ------------------------------------------------------------------------------------------------ asynchronous work feed
// The persistent wavefront kernels above fetch work when lanes run idle: one atomic on the queue cursor, then the queue
// entry, then the ray — three dependent trips to the L2 / HBM during which the whole warp waits (the live lanes too), for
// rays that take fewer than ten node steps. The *_feed kernels take the memory out of that path. Every warp owns a small
// staging area in shared memory and keeps a pipeline of 32-ray chunks in flight with cp.async (LDGSTS): while chunk c is
// handed out to idle lanes, the rays of chunk c + 1 and the queue entries of chunk c + 2 are on their way, and the atomic
// that reserves chunk c + 3 has been issued (its result is read one chunk later, from lane 0). A refill is then two
// ld.shared.v4 per lane. Per ray nothing changes — same box tests, same order — so the hits are the same bits.
RB_D void rb_cp_async4(unsigned int dst, const void * src) { asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory"); }
RB_D void rb_cp_async16(unsigned int dst, const void * src) { asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory"); }
RB_D void rb_cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
RB_D float4 rb_lds128(unsigned int a) {
	float4 v;
	asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a) : "memory");
	return v;
}
#define RB_FEED_NONE 0xffffffffu
#define RB_FEED_TRACE_BYTES (3 * 32 * 4 + 2 * 32 * 2 * 16)	// per warp: three chunks of queue entries, two chunks of (origin, direction)
#define RB_FEED_SHADOW_BYTES (2 * 32 * 3 * 16)				// per warp: two chunks of (origin+maxDist, direction+slot, contribution)

template<bool ORDERED, int ROOTS>
__global__ void __launch_bounds__(RB_PT_BLOCK) rb_k_pt_trace_feed(RbDeviceBvh bvh, RbPathBuffers pb, const uint32_t * __restrict__ queue, uint32_t did, int threshold, int leafVotes, int nodeSteps, unsigned long long * __restrict__ stats) {
	extern __shared__ int s_stack[];
	const RbStack st		 = rb_make_stack(s_stack + threadIdx.x, blockDim.x);
	const unsigned int lane	 = threadIdx.x & 31u;
	const unsigned int below = (1u << lane) - 1u;
	const unsigned int n	 = pb.counters[rb_cnt_queue(did)];
	unsigned int * cursor	 = pb.counters + rb_cnt_cursor_t(did);
	if(blockIdx.x == 0 && threadIdx.x == 0 && stats) atomicAdd(stats + 0, (unsigned long long)n);
	// Staging area of this warp, behind the traversal stacks.
	const unsigned int area	  = (unsigned int)__cvta_generic_to_shared(s_stack) + (unsigned int)(bvh.stackDepth > 0 ? bvh.stackDepth : 1) * blockDim.x * 4u + (threadIdx.x >> 5) * RB_FEED_TRACE_BYTES;
	const unsigned int sSlots = area;				// [3][32] u32
	const unsigned int sRays  = area + 3u * 32u * 4u; // [2][32][2] float4
	auto issueSlots = [&](unsigned int buf, unsigned int base) {
		if(base < n && base + lane < n) rb_cp_async4(sSlots + (buf * 32u + lane) * 4u, queue + base + lane);
	};
	auto issueRays = [&](unsigned int rbuf, unsigned int sbuf, unsigned int base) {
		if(base < n && base + lane < n) {
			const unsigned int sl = rb_lds(sSlots + (sbuf * 32u + lane) * 4u);
			rb_cp_async16(sRays + ((rbuf * 32u + lane) * 2u + 0u) * 16u, pb.rayO + sl);
			rb_cp_async16(sRays + ((rbuf * 32u + lane) * 2u + 1u) * 16u, pb.rayD + sl);
		}
	};
	// Prologue: three chunks reserved at once, their queue entries, then the rays of the first two.
	unsigned int nb = 0;
	if(lane == 0) nb = atomicAdd(cursor, 96u);
	unsigned int b0 = __shfl_sync(0xffffffffu, nb, 0), b1 = b0 + 32u, b2 = b0 + 64u;
	issueSlots(0u, b0); issueSlots(1u, b1); issueSlots(2u, b2);
	rb_cp_async_wait_all();
	__syncwarp();
	issueRays(0u, 0u, b0); issueRays(1u, 1u, b1);
	if(lane == 0 && b2 < n) nb = atomicAdd(cursor, 32u); // chunk 3
	rb_cp_async_wait_all();
	__syncwarp();
	unsigned int c = 0, c3 = 0; // chunk being handed out, and c mod 3
	unsigned int head = 0, count = b0 < n ? min(32u, n - b0) : 0u;
	bool drained = count == 0u;

	bool active = false;
	unsigned int slot = 0;
	RbRay ray;
	RbTraversal T;
	for(;;) {
		unsigned int idle = __ballot_sync(0xffffffffu, !active);
		while(idle != 0u && !drained) {
			if(count == 0u) {
				// Chunk c is used up: move on. Everything issued one chunk ago has had a chunk's worth of traversal to land.
				__syncwarp();
				const unsigned int reserved = __shfl_sync(0xffffffffu, nb, 0);
				c += 1u; c3 = c3 == 2u ? 0u : c3 + 1u;
				b0 = b1; b1 = b2; b2 = (b2 < n) ? reserved : RB_FEED_NONE;
				rb_cp_async_wait_all();
				__syncwarp();
				head  = 0u;
				count = b0 < n ? min(32u, n - b0) : 0u;
				if(count == 0u) { drained = true; break; }
				issueRays((c + 1u) & 1u, c3 == 2u ? 0u : c3 + 1u, b1);
				issueSlots(c3 == 0u ? 2u : c3 - 1u, b2); // (c + 2) mod 3
				if(lane == 0 && b2 < n) nb = atomicAdd(cursor, 32u);
			}
			const unsigned int take = min((unsigned int)__popc(idle), count);
			const unsigned int rank = __popc(idle & below);
			if(!active && rank < take) {
				const unsigned int e = head + rank;
				slot			 = rb_lds(sSlots + (c3 * 32u + e) * 4u);
				const float4 o	 = rb_lds128(sRays + (((c & 1u) * 32u + e) * 2u + 0u) * 16u);
				const float4 d	 = rb_lds128(sRays + (((c & 1u) * 32u + e) * 2u + 1u) * 16u);
				ray = rb_make_ray(V3{o.x, o.y, o.z}, V3{d.x, d.y, d.z});
				rb_traversal_init(T, bvh, 0.0001f, 1e8f, st);
				if(ROOTS == RB_ROOTS_MASK) T.root = int(rb_roots_pretest<false>(bvh, ray, 0.0001f, 1e8f, nullptr));
				active = true;
			}
			head += take;
			count -= take;
			idle = __ballot_sync(0xffffffffu, !active);
		}
		unsigned int activeM = __ballot_sync(0xffffffffu, active);
		if(activeM == 0u) break;
		bool more;
		do {
			const bool wantsLeaf	 = active && rb_wants_leaf(T);
			const unsigned int leafM = __ballot_sync(0xffffffffu, wantsLeaf);
			const unsigned int nodeM = activeM & ~leafM;
			bool finished = false;
			if(__popc(leafM) >= leafVotes || nodeM == 0u) {
				if(wantsLeaf) finished = rb_phase_leaf<false, false>(T, bvh, ray, nullptr);
			} else if(active && !wantsLeaf) {
				for(int k = 0; k < nodeSteps; ++k) {
					finished = rb_phase_node<false, ORDERED, ROOTS>(T, bvh, ray, st, nullptr);
					if(finished || rb_wants_leaf(T)) break;
				}
			}
			if(finished) {
				rb_st(pb.hit + slot, make_float4(T.best.t, T.best.u, T.best.v, __int_as_float(T.best.tri)));
				active = false;
			}
			activeM = __ballot_sync(0xffffffffu, active);
			more	= drained ? (activeM != 0u) : (__popc(activeM) >= threshold);
		} while(more);
	}
}

// The closest-hit wavefront on the eight-wide quantised tree (rb_render_opts.trace_mode = RB_TRACE_WIDE; experiment, see
// rb_wide.h): the queue handling of rb_k_pt_trace around the phase-split wide steps. Path rays are regular except for
// exact axis-parallel directions; those walk the binary tree on the spot, as in rb_k_trace_wide.
__global__ void __launch_bounds__(RB_PT_BLOCK) rb_k_pt_trace_wide(RbDeviceBvh bvh, RbDeviceWide wide, RbPathBuffers pb, const uint32_t * __restrict__ queue, uint32_t did, int threshold, int leafVotes, unsigned long long * __restrict__ stats) {
	extern __shared__ int s_stack[];
	const RbStack st		= rb_make_stack(s_stack + threadIdx.x, blockDim.x);
	const unsigned int lane = threadIdx.x & 31u;
	const unsigned int n	= pb.counters[rb_cnt_queue(did)];
	unsigned int * cursor	= pb.counters + rb_cnt_cursor_t(did);
	if(blockIdx.x == 0 && threadIdx.x == 0 && stats) atomicAdd(stats + 0, (unsigned long long)n);
	bool active = false, exhausted = false;
	unsigned int slot = 0;
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
						slot = rb_ld(queue + idx);
						const float4 o = rb_ld(pb.rayO + slot), d = rb_ld(pb.rayD + slot);
						ray = rb_make_ray(V3{o.x, o.y, o.z}, V3{d.x, d.y, d.z});
						const float big = 1152921504606846976.0f; // 2^60, as rb_wide_ray_ok
						if(ray.regular && fabsf(ray.inv.x) < big && fabsf(ray.inv.y) < big && fabsf(ray.inv.z) < big) {
							rb_wide_init(T, bvh, 0.0001f, 1e8f, st);
							active = true;
						} else {
							RbTraversal B;
							rb_traversal_init(B, bvh, 0.0001f, 1e8f, st);
							while(!rb_step_exact<false, false, false>(B, bvh, ray, st, nullptr)) {
							}
							rb_st(pb.hit + slot, make_float4(B.best.t, B.best.u, B.best.v, __int_as_float(B.best.tri)));
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
				if(wantsTri) finished = rb_wide_phase_triangle<false>(T, wide, ray, st);
			} else if(active && !wantsTri) {
				finished = rb_wide_phase_node(T, bvh, wide, ray, st);
			}
			if(finished) {
				rb_st(pb.hit + slot, make_float4(T.best.t, T.best.u, T.best.v, __int_as_float(T.best.tri)));
				active = false;
			}
			activeM = __ballot_sync(0xffffffffu, active);
			more	= exhausted ? (activeM != 0u) : (__popc(activeM) >= threshold);
		} while(more);
	}
}

// ------------------------------------------------------------------------------------------------ shading
RB_D void rb_queue_push(uint32_t * queue, uint32_t * counter, bool want, uint32_t value) {
	// Warp-aggregated append: one atomicAdd per warp, lanes ranked with __ballot/__popc.
	const unsigned int active = __activemask();
	const unsigned int mask	  = __ballot_sync(active, want);
	if(mask == 0) return;
	const unsigned int lane	  = threadIdx.x & 31u;
	const int leader		  = __ffs(mask) - 1;
	unsigned int base		  = 0;
	if((int)lane == leader) base = atomicAdd(counter, (unsigned int)__popc(mask));
	base = __shfl_sync(active, base, leader);
	if(want) queue[base + __popc(mask & ((1u << lane) - 1u))] = value;
}

__global__ void __launch_bounds__(RB_PT_BLOCK, RB_PT_SHADE_MIN_BLOCKS) rb_k_pt_shade(RbFrame f, RbBatch b, RbPathBuffers pb, RbSceneDev sc, uint32_t did, unsigned long long * __restrict__ stats) {
	const uint32_t n		 = pb.counters[rb_cnt_queue(did)];
	// (selected, not indexed: a run-time index into an array inside a kernel parameter costs a local-memory copy of it)
	const uint32_t * queue	 = (did & 1u) ? pb.queue[1] : pb.queue[0];
	uint32_t * nextQueue	 = (did & 1u) ? pb.queue[0] : pb.queue[1];
	uint32_t * nextCounter	 = pb.counters + rb_cnt_queue(did + 1u);
	const bool lastBounce	 = (did + 1u >= f.depth);
	// Whole warps iterate together so that the warp-aggregated pushes see converged lanes.
	const uint32_t nRounded = (n + 31u) & ~31u;
	for(uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < nRounded; i += gridDim.x * blockDim.x) {
		const bool valid = i < n;
		bool continues = false;
		uint32_t slot = 0;
		if(valid) {
			slot = queue[i];
			const float4 h	= pb.hit[slot];
			const float4 ro = pb.rayO[slot], rd = pb.rayD[slot], at = pb.atten[slot], cl = pb.color[slot];
			V3 rayPos = V3{ro.x, ro.y, ro.z}, rayDir = V3{rd.x, rd.y, rd.z};
			V3 attenuation = V3{at.x, at.y, at.z}, sampleColor = V3{cl.x, cl.y, cl.z};
			const uint64_t key = (uint64_t(__float_as_uint(cl.w)) << 32) | uint64_t(__float_as_uint(at.w)); // written by raygen
			const int tri	   = __float_as_int(h.w);
			if(tri < 0) {
				// No hit: background, then the path ends (PathTracer.cpp:206-209).
				float ndcx = 0.0f, ndcy = 0.0f;
				if(did == 0 && sc.bgMode == RB_BG_IMAGE) rb_sample_ndc(f, rb_path_id(f, b, slot), key, ndcx, ndcy);
				sampleColor += attenuation * rb_eval_background(sc, rayDir, rayPos, ndcx, ndcy, did == 0);
				pb.color[slot] = make_float4(sampleColor.x, sampleColor.y, sampleColor.z, cl.w);
			} else {
				RbSurface surf;
				uint32_t meshId;
				rb_fetch_surface(sc, tri, h.y, h.z, surf, meshId);
				const RbObject obj = sc.objects[meshId];
				const V3 p		   = rayPos + h.x * rayDir; // :214
				const bool noUVs = !(obj.flags & RB_OBJ_USE_UV);
				if(noUVs) {
					surf.u = 0.5f;
					surf.v = 0.5f;
				}
				const float4 bCol = rb_image_rgbal(sc, obj.texAlbedo, surf.u, surf.v);
				if((obj.flags & RB_OBJ_MASKED) && bCol.w < 0.01f) {
					// Alpha cut-out: move to the hit point and keep casting; this spends one depth level (:223-226).
					pb.rayO[slot] = make_float4(p.x, p.y, p.z, 0.0f);
					continues	  = !lastBounce;
				} else if(obj.flags & RB_OBJ_EMISSIVE) {
					sampleColor += attenuation * V3{bCol.x, bCol.y, bCol.z}; // :228-234
					pb.color[slot] = make_float4(sampleColor.x, sampleColor.y, sampleColor.z, cl.w);
				} else {
					const M3 tbn		 = rb_build_local_frame(sc, obj, surf, rayDir);
					const V3 wo			 = normalize3(mulT(tbn, -rayDir));									// :241
					const V3 baseColor	 = V3{powf(bCol.x, 2.2f), powf(bCol.y, 2.2f), powf(bCol.z, 2.2f)}; // :242
					const float4 rmao	 = rb_image_rgbal(sc, obj.texRmao, surf.u, surf.v);
					const uint32_t dim0 = 8u * (did + 1u);
					if(sc.nLights > 0) {
						// Direct light sampling (:248-271): one light picked uniformly.
						const float ul = rb_rng_float(key, dim0 + 0u);
						int lid		   = int(ul * float(sc.nLights));
						lid			   = lid < sc.nLights ? lid : sc.nLights - 1;
						const RbLight light = sc.lights[lid];
						const V3 pShift		= p + 0.001f * tbn.c[2];
						float maxDist, falloff;
						const V3 direction = rb_light_sample(light, pShift, maxDist, falloff);
						if(falloff > 0.0f) {
							const V3 lwi		= normalize3(mulT(tbn, direction));
							// (The wi-independent terms are computed at each of the two uses: hoisted above the light
							// block they stay live across it and the kernel spills — measured, 1 % slower.)
							const RbGgxSurface gs = rb_ggx_surface(wo, baseColor, rmao.x, rmao.y);
							const V3 evalLight	= rb_ggx_eval(gs, lwi);
							const float lightPdf = 1.0f / float(sc.nLights);
							const V3 illumination = ((falloff * evalLight) * V3{light.intensity[0], light.intensity[1], light.intensity[2]}) / lightPdf;
							const V3 contrib	  = attenuation * illumination;
							if(light.castsShadow) {
								// The visibility walk runs in the shadow kernel; a contribution of exactly zero needs no ray.
								if(contrib.x != 0.0f || contrib.y != 0.0f || contrib.z != 0.0f) {
									// Appended on the spot, aggregated over the lanes that are here together: held until the end of
									// the iteration the entry would occupy twelve registers across the BSDF sampling below.
									const unsigned int am	= __activemask();
									const unsigned int ln	= threadIdx.x & 31u;
									const int leader		= __ffs(am) - 1;
									unsigned int base		= 0;
									if((int)ln == leader) base = atomicAdd(pb.counters + rb_cnt_shadow(did), (unsigned int)__popc(am));
									base					= __shfl_sync(am, base, leader);
									const unsigned int e	= base + __popc(am & ((1u << ln) - 1u));
									pb.shO[e] = make_float4(pShift.x, pShift.y, pShift.z, maxDist);
									pb.shD[e] = make_float4(direction.x, direction.y, direction.z, __uint_as_float(slot));
									pb.shC[e] = make_float4(contrib.x, contrib.y, contrib.z, 0.0f);
								}
							} else {
								sampleColor += contrib;
							}
						}
					}
					if(!lastBounce) {
						// Next direction from the BSDF (:274-285). On the last level the reference still draws it but
						// never uses the result.
						V3 wi;
						const RbGgxSurface gs = rb_ggx_surface(wo, baseColor, rmao.x, rmao.y);
						const V3 eval = rb_ggx_sample_and_eval(gs, baseColor, rmao.y, rb_rng_float(key, dim0 + 1u), rb_rng_float(key, dim0 + 2u), rb_rng_float(key, dim0 + 3u), wi);
						const V3 nextRayDir = normalize3(mul(tbn, wi));
						attenuation *= eval;
						const V3 newDir = normalize3(nextRayDir);
						pb.rayO[slot] = make_float4(p.x, p.y, p.z, 0.0f);
						pb.rayD[slot] = make_float4(newDir.x, newDir.y, newDir.z, 0.0f);
						pb.atten[slot] = make_float4(attenuation.x, attenuation.y, attenuation.z, at.w);
						continues		= (attenuation.x != 0.0f || attenuation.y != 0.0f || attenuation.z != 0.0f);
					}
					pb.color[slot] = make_float4(sampleColor.x, sampleColor.y, sampleColor.z, cl.w);
				}
			}
		}
		rb_queue_push(nextQueue, nextCounter, continues, slot);
	}
	if(blockIdx.x == 0 && threadIdx.x == 0 && stats) atomicAdd(stats + 2, (unsigned long long)n);
}

// ------------------------------------------------------------------------------------------------ shadow wavefront
// PathTracer::checkVisibility (PathTracer.cpp:124-157): closest-hit queries with (mini = 0.001, maxi = maxDist); an
// unmasked (or UV-less) occluder blocks, a masked one is looked up in its alpha channel and walked through when
// alpha < 0.01. When no object of the scene is masked the walk degenerates to "is there any hit", so the any-hit
// traversal (identical accept/reject decisions, early exit) is used instead (MASKED = false).
template<bool MASKED, bool ORDERED, int ROOTS, bool COUNT = false>
__global__ void __launch_bounds__(RB_PT_BLOCK, (!MASKED && !COUNT && ROOTS >= RB_ROOTS_TREE) ? 8 : 0) rb_k_pt_shadow(RbDeviceBvh bvh, RbPathBuffers pb, RbSceneDev sc, uint32_t did, int threshold, int leafVotes, int nodeSteps, unsigned long long * __restrict__ stats) {
	extern __shared__ int s_stack[];
	const RbStack st		= rb_make_stack(s_stack + threadIdx.x + (ROOTS == RB_ROOTS_TREE_BIG ? bvh.rootWords * blockDim.x : 0), blockDim.x);
	const unsigned int lane = threadIdx.x & 31u;
	const unsigned int n	= pb.counters[rb_cnt_shadow(did)];
	unsigned int * cursor	= pb.counters + rb_cnt_cursor_s(did);
	bool active = false, exhausted = false;
	unsigned int slot = 0;
	unsigned int queries = 0;
	RbWork work = {0u, 0u, 0u, 0u};
	RbWork * const wp = COUNT ? &work : nullptr;
	V3 lpos = v3(0.0f), dir = v3(0.0f), contrib = v3(0.0f);
	float maxDist = 0.0f;
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
						const float4 o = rb_ld(pb.shO + idx), d = rb_ld(pb.shD + idx), c = rb_ld(pb.shC + idx);
						lpos	= V3{o.x, o.y, o.z};
						maxDist = o.w;
						dir		= V3{d.x, d.y, d.z};
						slot	= __float_as_uint(d.w);
						contrib = V3{c.x, c.y, c.z};
						if(maxDist > 0.0f) {
							ray = rb_make_ray(lpos, dir);
							rb_traversal_init(T, bvh, 0.001f, maxDist, st);
							if(ROOTS == RB_ROOTS_MASK) T.root = int(rb_roots_pretest<COUNT>(bvh, ray, 0.001f, maxDist, wp));
							if(ROOTS == RB_ROOTS_TREE) rb_roots_tree_pretest<COUNT>(T, bvh, ray, st, 0.001f, maxDist, wp);
							if(ROOTS == RB_ROOTS_TREE_BIG) rb_roots_tree_pretest_big<COUNT>(T, bvh, ray, st, 0.001f, maxDist, wp);
							active = true;
							++queries;
						} else {
							// `while(maxDist > 0.0f)` never entered: visible (:127, :156).
							float4 cl = rb_ld(pb.color + slot);
							cl.x += contrib.x; cl.y += contrib.y; cl.z += contrib.z;
							rb_st(pb.color + slot, cl);
						}
					}
				}
				exhausted = (base + (unsigned int)nIdle) >= n;
			}
		}
		unsigned int activeM = __ballot_sync(0xffffffffu, active);
		if(activeM == 0u) break;
		bool more;
		do {
			// Same phase vote as the closest-hit wavefront.
			const bool wantsLeaf	 = active && rb_wants_leaf(T);
			const unsigned int leafM = __ballot_sync(0xffffffffu, wantsLeaf);
			const unsigned int nodeM = activeM & ~leafM;
			bool done = false;
			if(__popc(leafM) >= leafVotes || nodeM == 0u) {
				if(wantsLeaf) done = rb_phase_leaf<!MASKED, COUNT>(T, bvh, ray, wp);
			} else if(active && !wantsLeaf) {
				for(int k = 0; k < nodeSteps; ++k) {
					done = rb_phase_node<COUNT, ORDERED, ROOTS>(T, bvh, ray, st, wp);
					if(done || rb_wants_leaf(T)) break;
				}
			}
			if(done) {
				bool visible = false, finished = true;
				if(T.best.tri < 0) {
					visible = true;
				} else if(MASKED) {
					uint32_t meshId	   = __float_as_uint(rb_ldg4(bvh.tris + 3 * size_t(T.best.tri) + 1).w); // just read by the traversal
					const RbObject obj = sc.objects[meshId];
					if((obj.flags & RB_OBJ_MASKED) && (obj.flags & RB_OBJ_USE_UV)) {
						float lu, lv;
						rb_fetch_uv(sc, T.best.tri, T.best.u, T.best.v, lu, lv, meshId);
						const float alpha = rb_image_rgbal(sc, obj.texAlbedo, lu, lv).w;
						if(alpha < 0.01f) {
							// Transparent texel: shift to the hit, shorten the segment, cast again (:149-151).
							maxDist = maxDist - T.best.t;
							lpos	= lpos + T.best.t * dir;
							if(maxDist > 0.0f) {
								ray = rb_make_ray(lpos, dir);
								rb_traversal_init(T, bvh, 0.001f, maxDist, st);
								if(ROOTS == RB_ROOTS_MASK) T.root = int(rb_roots_pretest<COUNT>(bvh, ray, 0.001f, maxDist, wp));
								if(ROOTS == RB_ROOTS_TREE) rb_roots_tree_pretest<COUNT>(T, bvh, ray, st, 0.001f, maxDist, wp);
								if(ROOTS == RB_ROOTS_TREE_BIG) rb_roots_tree_pretest_big<COUNT>(T, bvh, ray, st, 0.001f, maxDist, wp);
							if(ROOTS == RB_ROOTS_TREE_BIG) rb_roots_tree_pretest_big<COUNT>(T, bvh, ray, st, 0.001f, maxDist, wp);
								finished = false;
								++queries;
							} else {
								visible = true;
							}
						}
					}
				}
				if(finished) {
					if(visible) {
						float4 cl = rb_ld(pb.color + slot);
						cl.x += contrib.x; cl.y += contrib.y; cl.z += contrib.z;
						rb_st(pb.color + slot, cl);
					}
					active = false;
				}
			}
			activeM = __ballot_sync(0xffffffffu, active);
			more	= exhausted ? (activeM != 0u) : (__popc(activeM) >= threshold);
		} while(more);
	}
	if(stats) {
		for(int off = 16; off > 0; off >>= 1) queries += __shfl_down_sync(0xffffffffu, queries, off);
		if(lane == 0 && queries) atomicAdd(stats + 1, (unsigned long long)queries);
		if(COUNT) rb_pt_flush_work(work, lane, stats, 7);
	}
}

// The shadow wavefront fed the same way. Its entries are dense (the shade kernel appends them), so the pipeline is one
// stage shorter: chunk c handed out, the entries of chunk c + 1 in flight, the atomic for chunk c + 2 issued.
template<bool MASKED, bool ORDERED, int ROOTS>
__global__ void __launch_bounds__(RB_PT_BLOCK, 8) rb_k_pt_shadow_feed(RbDeviceBvh bvh, RbPathBuffers pb, RbSceneDev sc, uint32_t did, int threshold, int leafVotes, int nodeSteps, unsigned long long * __restrict__ stats) {
	extern __shared__ int s_stack[];
	const RbStack st		 = rb_make_stack(s_stack + threadIdx.x, blockDim.x);
	const unsigned int lane	 = threadIdx.x & 31u;
	const unsigned int below = (1u << lane) - 1u;
	const unsigned int n	 = pb.counters[rb_cnt_shadow(did)];
	unsigned int * cursor	 = pb.counters + rb_cnt_cursor_s(did);
	const unsigned int sEnt	 = (unsigned int)__cvta_generic_to_shared(s_stack) + (unsigned int)(bvh.stackDepth > 0 ? bvh.stackDepth : 1) * blockDim.x * 4u + (threadIdx.x >> 5) * RB_FEED_SHADOW_BYTES; // [2][32][3] float4
	auto issue = [&](unsigned int buf, unsigned int base) {
		if(base < n && base + lane < n) {
			const unsigned int d = sEnt + ((buf * 32u + lane) * 3u) * 16u;
			rb_cp_async16(d, pb.shO + base + lane);
			rb_cp_async16(d + 16u, pb.shD + base + lane);
			rb_cp_async16(d + 32u, pb.shC + base + lane);
		}
	};
	unsigned int nb = 0;
	if(lane == 0) nb = atomicAdd(cursor, 64u);
	unsigned int b0 = __shfl_sync(0xffffffffu, nb, 0), b1 = b0 + 32u;
	issue(0u, b0); issue(1u, b1);
	if(lane == 0 && b1 < n) nb = atomicAdd(cursor, 32u); // chunk 2
	rb_cp_async_wait_all();
	__syncwarp();
	unsigned int c = 0, head = 0, count = b0 < n ? min(32u, n - b0) : 0u;
	bool drained = count == 0u;

	bool active = false;
	unsigned int slot = 0;
	unsigned int queries = 0;
	V3 lpos = v3(0.0f), dir = v3(0.0f), contrib = v3(0.0f);
	float maxDist = 0.0f;
	RbRay ray;
	RbTraversal T;
	for(;;) {
		unsigned int idle = __ballot_sync(0xffffffffu, !active);
		while(idle != 0u && !drained) {
			if(count == 0u) {
				__syncwarp();
				const unsigned int reserved = __shfl_sync(0xffffffffu, nb, 0);
				c += 1u;
				b0 = b1; b1 = (b1 < n) ? reserved : RB_FEED_NONE;
				rb_cp_async_wait_all();
				__syncwarp();
				head  = 0u;
				count = b0 < n ? min(32u, n - b0) : 0u;
				if(count == 0u) { drained = true; break; }
				issue((c + 1u) & 1u, b1);
				if(lane == 0 && b1 < n) nb = atomicAdd(cursor, 32u);
			}
			const unsigned int take = min((unsigned int)__popc(idle), count);
			const unsigned int rank = __popc(idle & below);
			if(!active && rank < take) {
				const unsigned int a = sEnt + (((c & 1u) * 32u + head + rank) * 3u) * 16u;
				const float4 o = rb_lds128(a), d = rb_lds128(a + 16u), cc = rb_lds128(a + 32u);
				lpos	= V3{o.x, o.y, o.z};
				maxDist = o.w;
				dir		= V3{d.x, d.y, d.z};
				slot	= __float_as_uint(d.w);
				contrib = V3{cc.x, cc.y, cc.z};
				if(maxDist > 0.0f) {
					ray = rb_make_ray(lpos, dir);
					rb_traversal_init(T, bvh, 0.001f, maxDist, st);
					if(ROOTS == RB_ROOTS_MASK) T.root = int(rb_roots_pretest<false>(bvh, ray, 0.001f, maxDist, nullptr));
					active = true;
					++queries;
				} else {
					// `while(maxDist > 0.0f)` never entered: visible (:127, :156).
					float4 cl = rb_ld(pb.color + slot);
					cl.x += contrib.x; cl.y += contrib.y; cl.z += contrib.z;
					rb_st(pb.color + slot, cl);
				}
			}
			head += take;
			count -= take;
			idle = __ballot_sync(0xffffffffu, !active);
		}
		unsigned int activeM = __ballot_sync(0xffffffffu, active);
		if(activeM == 0u) {
			if(drained) break;
			continue; // every entry taken was visible without a query: fetch more
		}
		bool more;
		do {
			const bool wantsLeaf	 = active && rb_wants_leaf(T);
			const unsigned int leafM = __ballot_sync(0xffffffffu, wantsLeaf);
			const unsigned int nodeM = activeM & ~leafM;
			bool done = false;
			if(__popc(leafM) >= leafVotes || nodeM == 0u) {
				if(wantsLeaf) done = rb_phase_leaf<!MASKED, false>(T, bvh, ray, nullptr);
			} else if(active && !wantsLeaf) {
				for(int k = 0; k < nodeSteps; ++k) {
					done = rb_phase_node<false, ORDERED, ROOTS>(T, bvh, ray, st, nullptr);
					if(done || rb_wants_leaf(T)) break;
				}
			}
			if(done) {
				bool visible = false, finished = true;
				if(T.best.tri < 0) {
					visible = true;
				} else if(MASKED) {
					uint32_t meshId	   = __float_as_uint(rb_ldg4(bvh.tris + 3 * size_t(T.best.tri) + 1).w); // just read by the traversal
					const RbObject obj = sc.objects[meshId];
					if((obj.flags & RB_OBJ_MASKED) && (obj.flags & RB_OBJ_USE_UV)) {
						float lu, lv;
						rb_fetch_uv(sc, T.best.tri, T.best.u, T.best.v, lu, lv, meshId);
						const float alpha = rb_image_rgbal(sc, obj.texAlbedo, lu, lv).w;
						if(alpha < 0.01f) {
							// Transparent texel: shift to the hit, shorten the segment, cast again (:149-151).
							maxDist = maxDist - T.best.t;
							lpos	= lpos + T.best.t * dir;
							if(maxDist > 0.0f) {
								ray = rb_make_ray(lpos, dir);
								rb_traversal_init(T, bvh, 0.001f, maxDist, st);
								if(ROOTS == RB_ROOTS_MASK) T.root = int(rb_roots_pretest<false>(bvh, ray, 0.001f, maxDist, nullptr));
								finished = false;
								++queries;
							} else {
								visible = true;
							}
						}
					}
				}
				if(finished) {
					if(visible) {
						float4 cl = rb_ld(pb.color + slot);
						cl.x += contrib.x; cl.y += contrib.y; cl.z += contrib.z;
						rb_st(pb.color + slot, cl);
					}
					active = false;
				}
			}
			activeM = __ballot_sync(0xffffffffu, active);
			more	= drained ? (activeM != 0u) : (__popc(activeM) >= threshold);
		} while(more);
	}
	if(stats) {
		for(int off = 16; off > 0; off >>= 1) queries += __shfl_down_sync(0xffffffffu, queries, off);
		if(lane == 0 && queries) atomicAdd(stats + 1, (unsigned long long)queries);
	}
}

// The shadow wavefront of an UNMASKED scene on the eight-wide tree (trace_mode RB_TRACE_WIDE; experiment): any-hit queries,
// whose answer no visiting order can change. Masked scenes keep rb_k_pt_shadow (their walk needs the closest occluder).
__global__ void __launch_bounds__(RB_PT_BLOCK) rb_k_pt_shadow_wide(RbDeviceBvh bvh, RbDeviceWide wide, RbPathBuffers pb, uint32_t did, int threshold, int leafVotes, unsigned long long * __restrict__ stats) {
	extern __shared__ int s_stack[];
	const RbStack st		= rb_make_stack(s_stack + threadIdx.x, blockDim.x);
	const unsigned int lane = threadIdx.x & 31u;
	const unsigned int n	= pb.counters[rb_cnt_shadow(did)];
	unsigned int * cursor	= pb.counters + rb_cnt_cursor_s(did);
	bool active = false, exhausted = false;
	unsigned int slot = 0, queries = 0;
	V3 contrib = v3(0.0f);
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
						const float4 o = rb_ld(pb.shO + idx), d = rb_ld(pb.shD + idx), c = rb_ld(pb.shC + idx);
						const float maxDist = o.w;
						slot	= __float_as_uint(d.w);
						contrib = V3{c.x, c.y, c.z};
						bool visible = true, decided = true;
						if(maxDist > 0.0f) {
							ray = rb_make_ray(V3{o.x, o.y, o.z}, V3{d.x, d.y, d.z});
							++queries;
							const float big = 1152921504606846976.0f; // 2^60, as rb_wide_ray_ok
							if(ray.regular && fabsf(ray.inv.x) < big && fabsf(ray.inv.y) < big && fabsf(ray.inv.z) < big) {
								rb_wide_init(T, bvh, 0.001f, maxDist, st);
								active	= true;
								decided = false;
							} else {
								RbTraversal B;
								rb_traversal_init(B, bvh, 0.001f, maxDist, st);
								while(!rb_step_ordered<true, false, false>(B, bvh, ray, st, nullptr)) {
								}
								visible = B.best.tri < 0;
							}
						}
						if(decided && visible) {
							float4 cl = rb_ld(pb.color + slot);
							cl.x += contrib.x; cl.y += contrib.y; cl.z += contrib.z;
							rb_st(pb.color + slot, cl);
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
			bool done = false;
			if(__popc(triM) >= leafVotes || nodeM == 0u) {
				if(wantsTri) done = rb_wide_phase_triangle<true>(T, wide, ray, st);
			} else if(active && !wantsTri) {
				done = rb_wide_phase_node(T, bvh, wide, ray, st);
			}
			if(done) {
				if(T.best.tri < 0) {
					float4 cl = rb_ld(pb.color + slot);
					cl.x += contrib.x; cl.y += contrib.y; cl.z += contrib.z;
					rb_st(pb.color + slot, cl);
				}
				active = false;
			}
			activeM = __ballot_sync(0xffffffffu, active);
			more	= exhausted ? (activeM != 0u) : (__popc(activeM) >= threshold);
		} while(more);
	}
	if(stats) {
		for(int off = 16; off > 0; off >>= 1) queries += __shfl_down_sync(0xffffffffu, queries, off);
		if(lane == 0 && queries) atomicAdd(stats + 1, (unsigned long long)queries);
	}
}

// ------------------------------------------------------------------------------------------------ accumulation
// `render.rgb(x,y) += glm::min(sampleColor, 5.0f)` (PathTracer.cpp:288): one warp per pixel sums the clamped samples of
// the batch in a fixed order (lane-strided, then a shuffle tree), so the image is reproducible bit for bit.
__global__ void __launch_bounds__(RB_PT_BLOCK) rb_k_pt_accumulate(RbFrame f, RbBatch b, RbPathBuffers pb, float * __restrict__ sum) {
	const unsigned int lane	 = threadIdx.x & 31u;
	const unsigned int warp	 = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	const unsigned int nwarp = (gridDim.x * blockDim.x) >> 5;
	for(uint32_t p = warp; p < b.npixels; p += nwarp) {
		float r = 0.0f, g = 0.0f, bl = 0.0f;
		for(uint32_t s = lane; s < b.schunk; s += 32u) {
			const float4 c = rb_ld(pb.color + (p * b.schunk + s));
			r += (5.0f < c.x) ? 5.0f : c.x; // glm::min(x, 5.0f)
			g += (5.0f < c.y) ? 5.0f : c.y;
			bl += (5.0f < c.z) ? 5.0f : c.z;
		}
		for(int off = 16; off > 0; off >>= 1) {
			r += __shfl_down_sync(0xffffffffu, r, off);
			g += __shfl_down_sync(0xffffffffu, g, off);
			bl += __shfl_down_sync(0xffffffffu, bl, off);
		}
		if(lane == 0) {
			const uint32_t lp	= b.pixel0 + p;
			const uint32_t lr	= rb_div(lp, f.mWidth);
			const uint32_t y	= rb_tile_row(f, lr);
			const size_t pixel = size_t(y) * f.width + (lp - lr * f.width);
			sum[3 * pixel + 0] += r;
			sum[3 * pixel + 1] += g;
			sum[3 * pixel + 2] += bl;
		}
	}
}

// `render.rgb(x,y) / float(samples)` (PathTracer.cpp:291-292) on the rows of this call's tile.
__global__ void rb_k_pt_resolve(RbFrame f, const float * __restrict__ sum, uint32_t components, float * __restrict__ pixels) {
	const uint32_t n = f.localRows * f.width;
	for(uint32_t lp = blockIdx.x * blockDim.x + threadIdx.x; lp < n; lp += gridDim.x * blockDim.x) {
		const uint32_t lr  = rb_div(lp, f.mWidth);
		const uint32_t y   = rb_tile_row(f, lr);
		const size_t pixel = size_t(y) * f.width + (lp - lr * f.width);
		for(int c = 0; c < 3; ++c) {
			const float v = pixels[pixel * components + c] + sum[3 * pixel + c];
			pixels[pixel * components + c] = v / float(f.samples);
		}
	}
}

// ------------------------------------------------------------------------------------------------ host side
// Wavefront lanes: consecutive batches alternate between buffer sets on their own streams, so that the drain of one
// batch's persistent kernels (a few long rays keep a handful of warps busy) is filled by the other batch's kernels.
#define RB_PT_LANES 4
#define RB_PT_MAX_DEPTH 64

struct rb_pathtracer {
	rb_scene * scene = nullptr;
	RbPathBuffers pb[RB_PT_LANES] = {};
	size_t capacity[RB_PT_LANES]  = {};
	cudaStream_t lane[RB_PT_LANES] = {};
	cudaStream_t side[RB_PT_LANES] = {}; // shadow wavefront of a lane: runs beside the next level's closest-hit wavefront
	cudaEvent_t fork = nullptr, join[RB_PT_LANES] = {}, shaded[RB_PT_LANES] = {}, shadowed[RB_PT_LANES] = {};
	int lanes = 2;
	unsigned long long * d_stats = nullptr; // RB_PT_NSTATS counters (rays, shaded hits; the COUNT kernels' work counters)
	rb_render_stats stats = {};
	int sms = 0;
	// rb_pathtracer_profile_frame: the frame runs on one stream, an event before every kernel (timing), or with the
	// counting copies of the traversal kernels (counting). Null during ordinary renders.
	struct Profile {
		bool timing = false, counting = false;
		std::vector<cudaEvent_t> events;
		std::vector<int> kinds; // kernel class launched after events[k]
	} * prof = nullptr;
	void * multi = nullptr; // rb_multi.cu: replicas on the other GPUs of the box (rb_render_opts.gpus > 1)
	// Frame buffers of rb_pathtracer_render (host image in / out), kept between calls: a cudaMalloc / cudaFree pair per
	// frame synchronises the device and showed up as 100-300 ms stalls in the first frames of a sequence.
	float * d_frameSum = nullptr, * d_framePixels = nullptr;
	size_t frameSumFloats = 0, framePixelFloats = 0;
	// Live session (rb_pathtracer_live_*): per-pixel sums and display bytes stay on the device between passes.
	float * d_liveSum	  = nullptr;
	uint8_t * d_liveBytes = nullptr;
	uint32_t liveW = 0, liveH = 0, livePasses = 0, livePassCount = 0, liveSamples = 0, liveDepth = 0;
	rb_camera liveCamera = {};
};

rb_scene * rb_pt_scene(rb_pathtracer * pt) { return pt->scene; }
void ** rb_pt_multi_slot(rb_pathtracer * pt) { return &pt->multi; }
void rb_pt_set_stats(rb_pathtracer * pt, const rb_render_stats & s) { pt->stats = s; }
#define RB_PT_NSTATS 12

static void rb_pt_free(rb_pathtracer * pt) {
	for(int l = 0; l < RB_PT_LANES; ++l) {
		RbPathBuffers & pb = pt->pb[l];
		cudaFree(pb.rayO);
		cudaFree(pb.rayD);
		cudaFree(pb.atten);
		cudaFree(pb.color);
		cudaFree(pb.hit);
		cudaFree(pb.queue[0]);
		cudaFree(pb.queue[1]);
		cudaFree(pb.shO);
		cudaFree(pb.shD);
		cudaFree(pb.shC);
		cudaFree(pb.counters);
		pb				= RbPathBuffers{};
		pt->capacity[l] = 0;
	}
}

static rb_status rb_pt_reserve(rb_pathtracer * pt, int l, size_t paths) {
	if(paths <= pt->capacity[l]) return RB_OK;
	RbPathBuffers & pb = pt->pb[l];
	if(pt->capacity[l]) {
		RB_CUDA(cudaDeviceSynchronize());
		cudaFree(pb.rayO); cudaFree(pb.rayD); cudaFree(pb.atten); cudaFree(pb.color); cudaFree(pb.hit);
		cudaFree(pb.queue[0]); cudaFree(pb.queue[1]); cudaFree(pb.shO); cudaFree(pb.shD); cudaFree(pb.shC); cudaFree(pb.counters);
		pb = RbPathBuffers{};
		pt->capacity[l] = 0;
	}
	// All or nothing: a partial allocation (out of memory half way) is released again, capacity stays 0.
	void ** slots[]		 = {(void **)&pb.rayO, (void **)&pb.rayD, (void **)&pb.atten, (void **)&pb.color, (void **)&pb.hit, (void **)&pb.queue[0], (void **)&pb.queue[1], (void **)&pb.shO, (void **)&pb.shD, (void **)&pb.shC, (void **)&pb.counters};
	const size_t bytes[] = {paths * sizeof(float4), paths * sizeof(float4), paths * sizeof(float4), paths * sizeof(float4), paths * sizeof(float4), paths * sizeof(uint32_t), paths * sizeof(uint32_t), paths * sizeof(float4), paths * sizeof(float4), paths * sizeof(float4), 4 * (RB_PT_MAX_DEPTH + 1) * sizeof(uint32_t)};
	for(size_t k = 0; k < sizeof(slots) / sizeof(slots[0]); ++k) {
		const cudaError_t e = cudaMalloc(slots[k], bytes[k]);
		if(e != cudaSuccess) {
			for(size_t j = 0; j < k; ++j) cudaFree(*slots[j]);
			pb = RbPathBuffers{};
			return rb_fail(RB_ERR_CUDA, "path buffers for %zu paths: cudaMalloc failed: %s", paths, cudaGetErrorString(e));
		}
	}
	pt->capacity[l] = paths;
	return RB_OK;
}

extern "C" rb_status rb_pathtracer_create(rb_scene * scene, rb_pathtracer ** out) {
	if(!scene || !out) return rb_fail(RB_ERR_INVALID, "rb_pathtracer_create: null argument");
	RB_REQUIRE_DEVICE();
	if(!scene->committed) return rb_fail(RB_ERR_INVALID, "rb_pathtracer_create: call rb_scene_commit first");
	if(cudaSetDevice(scene->raycaster->device) != cudaSuccess) return rb_fail(RB_ERR_CUDA, "rb_pathtracer_create: cannot select device %d", scene->raycaster->device);
	rb_pathtracer * pt = new rb_pathtracer();
	pt->scene		   = scene;
	cudaDeviceGetAttribute(&pt->sms, cudaDevAttrMultiProcessorCount, scene->raycaster->device);
	if(cudaMalloc(&pt->d_stats, RB_PT_NSTATS * sizeof(unsigned long long)) != cudaSuccess) {
		delete pt;
		return rb_fail(RB_ERR_CUDA, "rb_pathtracer_create: cudaMalloc failed");
	}
	bool ok = cudaEventCreateWithFlags(&pt->fork, cudaEventDisableTiming) == cudaSuccess;
	for(int l = 0; l < RB_PT_LANES; ++l) {
		ok = ok && cudaStreamCreateWithFlags(&pt->lane[l], cudaStreamNonBlocking) == cudaSuccess;
		ok = ok && cudaStreamCreateWithFlags(&pt->side[l], cudaStreamNonBlocking) == cudaSuccess;
		ok = ok && cudaEventCreateWithFlags(&pt->join[l], cudaEventDisableTiming) == cudaSuccess;
		ok = ok && cudaEventCreateWithFlags(&pt->shaded[l], cudaEventDisableTiming) == cudaSuccess;
		ok = ok && cudaEventCreateWithFlags(&pt->shadowed[l], cudaEventDisableTiming) == cudaSuccess;
	}
	if(!ok) {
		rb_pathtracer_destroy(pt);
		return rb_fail(RB_ERR_CUDA, "rb_pathtracer_create: stream setup failed");
	}
	if(const char * e = std::getenv("RB_PT_LANES")) pt->lanes = std::min(std::max(std::atoi(e), 1), RB_PT_LANES);
	*out = pt;
	return RB_OK;
}

extern "C" void rb_pathtracer_destroy(rb_pathtracer * pt) {
	if(!pt) return;
	if(pt->multi) {
		rb_pt_multi_destroy(pt->multi);
		pt->multi = nullptr;
	}
	if(rb_have_device()) {
		rb_pt_free(pt);
		cudaFree(pt->d_stats);
		cudaFree(pt->d_frameSum);
		cudaFree(pt->d_framePixels);
		cudaFree(pt->d_liveSum);
		cudaFree(pt->d_liveBytes);
		if(pt->fork) cudaEventDestroy(pt->fork);
		for(int l = 0; l < RB_PT_LANES; ++l) {
			if(pt->join[l]) cudaEventDestroy(pt->join[l]);
			if(pt->shaded[l]) cudaEventDestroy(pt->shaded[l]);
			if(pt->shadowed[l]) cudaEventDestroy(pt->shadowed[l]);
			if(pt->lane[l]) cudaStreamDestroy(pt->lane[l]);
			if(pt->side[l]) cudaStreamDestroy(pt->side[l]);
		}
	}
	delete pt;
}

uint32_t rb_pt_round_samples(uint32_t samples) { // (also used by rb_multi.cu: the peer-memory resolve divides by the same count)
	// size_t(std::pow(2, std::round(std::log2(float(samples))))) (PathTracer.cpp:169-170)
	if(samples == 0) return 1;
	return uint32_t(std::pow(2.0, std::round(std::log2(float(samples)))));
}

static rb_status rb_pt_frame(const rb_camera * cam, uint32_t width, uint32_t height, uint32_t samples, uint32_t depth, const rb_render_opts * opts, RbFrame & f) {
	if(!cam) return rb_fail(RB_ERR_INVALID, "render: null camera");
	if(width == 0 || height == 0) return rb_fail(RB_ERR_INVALID, "render: empty image");
	if(depth > RB_PT_MAX_DEPTH) return rb_fail(RB_ERR_UNSUPPORTED, "render: depth %u exceeds %d", depth, RB_PT_MAX_DEPTH);
	std::memcpy(f.eye, cam->eye, sizeof(f.eye));
	std::memcpy(f.corner, cam->corner, sizeof(f.corner));
	std::memcpy(f.dx, cam->dx, sizeof(f.dx));
	std::memcpy(f.dy, cam->dy, sizeof(f.dy));
	f.width	  = width;
	f.height  = height;
	f.samples = rb_pt_round_samples(samples);
	// getSampleGrid (PathTracer.cpp:57-69)
	const int k = int(std::floor(std::log2(double(f.samples))));
	f.cellsX = f.cellsY = uint32_t(std::pow(2, k / 2));
	if(k % 2 == 1) {
		f.cellsX = uint32_t(std::pow(2, (k + 1) / 2));
		f.cellsY = uint32_t(std::pow(2, (k - 1) / 2));
	}
	f.cellSizeX = 1.0f / float(int(f.cellsX)); // 1.0f / glm::vec2(cellCount) (:179)
	f.cellSizeY = 1.0f / float(int(f.cellsY));
	f.depth		= depth;
	f.seed		= opts ? opts->seed : 0;
	f.tileRows	= (opts && opts->tile_rows) ? opts->tile_rows : 1;
	f.tileCount = (opts && opts->tile_count) ? opts->tile_count : 1;
	f.tileIndex = opts ? opts->tile_index : 0;
	if(f.tileIndex >= f.tileCount) return rb_fail(RB_ERR_INVALID, "render: tile_index %u >= tile_count %u", f.tileIndex, f.tileCount);
	f.sampleCount = (opts && opts->sample_count) ? opts->sample_count : 1;
	f.sampleIndex = opts ? opts->sample_index : 0;
	if(f.sampleIndex >= f.sampleCount) return rb_fail(RB_ERR_INVALID, "render: sample_index %u >= sample_count %u", f.sampleIndex, f.sampleCount);
	uint32_t rows = 0;
	for(uint32_t y = 0; y < height; ++y) {
		if(f.tileCount <= 1 || (y / f.tileRows) % f.tileCount == f.tileIndex) ++rows;
	}
	f.localRows = rows;
	f.mWidth	= rb_div_magic(f.width);
	f.mTileRows = rb_div_magic(f.tileRows);
	f.mCellsX	= rb_div_magic(f.cellsX);
	return RB_OK;
}

template<typename K>
static rb_status rb_pt_grid(K kernel, size_t smem, int sms, unsigned & grid) {
	if(smem > 48 * 1024) RB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
	int perSm = 0;
	RB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, kernel, RB_PT_BLOCK, smem));
	if(perSm < 1) return rb_fail(RB_ERR_CUDA, "path-tracer kernel does not fit on an SM");
	grid = unsigned(perSm * sms);
	return RB_OK;
}

extern "C" rb_status rb_pathtracer_render_device(rb_pathtracer * pt, const rb_camera * camera, uint32_t width, uint32_t height, uint32_t samples, uint32_t depth, float * d_sum, const rb_render_opts * opts, void * stream) {
	if(!pt || !d_sum) return rb_fail(RB_ERR_INVALID, "rb_pathtracer_render_device: null argument");
	RB_REQUIRE_DEVICE();
	RbFrame f;
	rb_status st = rb_pt_frame(camera, width, height, samples, depth, opts, f);
	if(st) return st;
	cudaStream_t s		 = static_cast<cudaStream_t>(stream);
	rb_scene * sc		 = pt->scene;
	if(sc->tablesDirty) { // lights / background edited since the last render
		st = rb_scene_sync(sc);
		if(st) return st;
	}
	rb_raycaster * rc	 = sc->raycaster;
	if(!sc->committed || !sc->geometryCommitted) return rb_fail(RB_ERR_INVALID, "rb_pathtracer_render: the scene changed since rb_scene_commit (a texture or an object was added): commit it again");
	// rb_scene_raycaster() hands out the scene's raycaster; a build on it reorders the triangles under the per-triangle
	// shading records gathered at commit.
	if(!rc->built || rc->buildSerial != sc->triShadeSerial) return rb_fail(RB_ERR_INVALID, "rb_pathtracer_render: the hierarchy of the scene's raycaster was rebuilt outside rb_scene_commit: commit the scene again");
	const bool ordered	 = opts ? (opts->trace_mode == RB_TRACE_ORDERED) : false;
	const bool wideTrace = opts && opts->trace_mode == RB_TRACE_WIDE; // closest-hit wavefront — and, in unmasked scenes, the shadow wavefront — on the eight-wide tree
	if(wideTrace && !rc->wideOnDevice) return rb_fail(RB_ERR_INVALID, "rb_pathtracer_render: trace_mode RB_TRACE_WIDE needs rb_raycaster_wide_build on the scene's raycaster first");
	// Wavefront queues are shorter and more varied than the raycaster's batches: a slightly earlier refill and fewer
	// node steps per vote measured best (tools/pt_sweep.sh); RB_FETCH_THRESHOLD / RB_NODE_STEPS override both.
	const int threshold = std::getenv("RB_FETCH_THRESHOLD") ? rc->fetchThreshold : 16;
	RB_CUDA(cudaSetDevice(rc->device));
	if(!(opts && opts->accumulate)) RB_CUDA(cudaMemsetAsync(d_sum, 0, size_t(width) * height * 3 * sizeof(float), s));
	RB_CUDA(cudaMemsetAsync(pt->d_stats, 0, RB_PT_NSTATS * sizeof(unsigned long long), s));

	const uint32_t localSamples = (f.samples > f.sampleIndex) ? (f.samples - f.sampleIndex + f.sampleCount - 1) / f.sampleCount : 0;
	const uint64_t localPixels	= uint64_t(f.localRows) * width;
	uint64_t launches = 0;
	if(localSamples && localPixels && depth) {
		size_t maxPaths = (opts && opts->max_paths_in_flight) ? opts->max_paths_in_flight : (size_t(1) << 25); // 32 Mi paths x 176 B x 2 lanes = 11.8 GB of the 180 GB
		if(maxPaths < 1024) maxPaths = 1024;
		const uint32_t schunk	  = uint32_t(localSamples < maxPaths ? localSamples : maxPaths);
		uint64_t pixelsPerBatch = maxPaths / schunk;
		if(pixelsPerBatch < 1) pixelsPerBatch = 1;
		if(pixelsPerBatch > localPixels) pixelsPerBatch = localPixels;
		// One sample chunk per pixel (the usual case): batches touch disjoint pixels and may run on the two lanes. Frames
		// worth splitting are cut into at least two batches.
		const bool oneChunk = schunk == localSamples;
		const bool profTiming = pt->prof && pt->prof->timing, profCounting = pt->prof && pt->prof->counting;
		int lanes = (oneChunk && !pt->prof) ? pt->lanes : 1;
		auto mark = [&](int kind, cudaStream_t on) -> cudaError_t { // profile: an event before every kernel of the (single-stream) frame
			if(!profTiming) return cudaSuccess;
			cudaEvent_t e;
			cudaError_t ce = cudaEventCreate(&e);
			if(ce != cudaSuccess) return ce;
			pt->prof->events.push_back(e);
			pt->prof->kinds.push_back(kind);
			return cudaEventRecord(e, on);
		};
		if(lanes > 1 && pixelsPerBatch * lanes > localPixels) {
			// fewer batches than lanes: cut the frame into one batch per lane when it is worth it
			if(localPixels * schunk >= (uint64_t(1) << 20)) pixelsPerBatch = (localPixels + lanes - 1) / lanes;
			else lanes = 1;
		}
		// Frames that fill the lanes reserve the full batch size, so that a later frame at another sample count does not
		// have to reallocate.
		const size_t needed	 = size_t(pixelsPerBatch) * schunk;
		const size_t reserve = (localPixels * uint64_t(schunk) >= uint64_t(maxPaths)) ? std::max(needed, maxPaths) : needed;
		for(int l = 0; l < lanes; ++l) {
			st = rb_pt_reserve(pt, l, reserve);
			if(st) return st;
		}

		const size_t smem = rb_trace_smem_bytes(rc, RB_PT_BLOCK);
		const int wideEntries = wideTrace ? std::max(5 * rc->devWide.stackDepth, rc->dev.stackDepth > 0 ? rc->dev.stackDepth : 1) : 0;
		const size_t smemWide = size_t(wideEntries) * RB_PT_BLOCK * sizeof(int);
		unsigned gridTraceWide = 0;
		if(wideTrace) {
			st = rb_pt_grid(rb_k_pt_trace_wide, smemWide, pt->sms, gridTraceWide);
			if(st) return st;
		}
		// Grid-stride kernels are launched as exactly one wave of resident CTAs: with 8 CTAs per SM requested and 6
		// resident (shade: 80 registers) a third of the CTAs would run alone in a second round.
		unsigned gridTrace = 0, gridShadow = 0, gridShade = 0, gridFlat = unsigned(pt->sms * 8);
		st = rb_pt_grid(rb_k_pt_shade, 0, pt->sms, gridShade);
		if(st) return st;
		const bool blocks = rc->dev.rootLevels > 0; // many-mesh scenes: kernels compiled with the root-block skipping
		// A handful of roots (<= 32): all box-tested when a ray starts instead of lazily in the middle of its traversal.
		// Not in the near-child-first order of closest-hit rays, which tests roots against the shrinking maxi.
		// 9..64 roots: the set of entered roots comes from one walk of the tree over the root boxes (rb_roots_tree_pretest)
		// instead of one test per root — 52 per ray in the reference's largest bundled scene. Wins over mask and blocks.
		const bool lazyRoots = std::getenv("RB_PT_LAZY_ROOTS") != nullptr;
		const bool treeRoots = !lazyRoots && rc->dev.rootTree != nullptr;
		const bool maskRoots = !treeRoots && !blocks && rc->dev.nroots >= 2 && rc->dev.nroots <= 32 && !lazyRoots;
		auto kTrace = blocks ? (ordered ? rb_k_pt_trace<true, RB_ROOTS_BLOCKS> : rb_k_pt_trace<false, RB_ROOTS_BLOCKS>)
							 : (ordered ? rb_k_pt_trace<true, RB_ROOTS_LAZY> : (maskRoots ? rb_k_pt_trace<false, RB_ROOTS_MASK> : rb_k_pt_trace<false, RB_ROOTS_LAZY>));
		const bool bigTree = treeRoots && rc->dev.rootWords > 0; // 65..512 meshes: the mask of entered roots in the stack column
		if(treeRoots && !ordered) kTrace = bigTree ? rb_k_pt_trace<false, RB_ROOTS_TREE_BIG> : rb_k_pt_trace<false, RB_ROOTS_TREE>;
		if(profCounting) kTrace = ordered ? rb_k_pt_trace<true, RB_ROOTS_LAZY, true> : (treeRoots ? (bigTree ? rb_k_pt_trace<false, RB_ROOTS_TREE_BIG, true> : rb_k_pt_trace<false, RB_ROOTS_TREE, true>) : (maskRoots ? rb_k_pt_trace<false, RB_ROOTS_MASK, true> : rb_k_pt_trace<false, RB_ROOTS_LAZY, true>)); // (block skipping: plain root loop, same hits, counted)
		// The cp.async-fed forms of the two wavefront kernels (see "asynchronous work feed"); the counting copies of the
		// profile pass keep the plain form.
		const bool feed = !profCounting && !treeRoots && std::getenv("RB_PT_FEED") != nullptr;
		size_t smemTrace = smem, smemShadow = smem;
		if(feed) {
			kTrace = blocks ? (ordered ? rb_k_pt_trace_feed<true, RB_ROOTS_BLOCKS> : rb_k_pt_trace_feed<false, RB_ROOTS_BLOCKS>)
							: (ordered ? rb_k_pt_trace_feed<true, RB_ROOTS_LAZY> : (maskRoots ? rb_k_pt_trace_feed<false, RB_ROOTS_MASK> : rb_k_pt_trace_feed<false, RB_ROOTS_LAZY>));
			smemTrace = smem + (RB_PT_BLOCK / 32) * RB_FEED_TRACE_BYTES;
		}
		st = rb_pt_grid(kTrace, smemTrace, pt->sms, gridTrace);
		if(st) return st;
		const bool masked = sc->dev.anyMasked != 0;
		// Unmasked scenes use any-hit shadow rays, whose result is order independent: always near-child-first.
		// The shadow wavefront keeps the lazy root loop: most of its walks end at the first occluder, before the other roots
		// are looked at (measured: +5 % with the pretest, against -3 % for the closest-hit wavefront).
		auto kShadow = blocks ? (masked ? (ordered ? rb_k_pt_shadow<true, true, RB_ROOTS_BLOCKS> : rb_k_pt_shadow<true, false, RB_ROOTS_BLOCKS>) : rb_k_pt_shadow<false, true, RB_ROOTS_BLOCKS>)
							  : (masked ? (ordered ? rb_k_pt_shadow<true, true, RB_ROOTS_LAZY> : rb_k_pt_shadow<true, false, RB_ROOTS_LAZY>) : rb_k_pt_shadow<false, true, RB_ROOTS_LAZY>);
		// With a root tree the shadow wavefront uses it too wherever its root tests see a fixed maxi (any-hit walks, and the
		// masked walk in the reference's order): a visible sample otherwise tests every root of the scene.
		const bool treeShadow = treeRoots && !(masked && ordered) && !std::getenv("RB_PT_SHADOW_LAZY");
		if(treeShadow) kShadow = masked ? (bigTree ? rb_k_pt_shadow<true, false, RB_ROOTS_TREE_BIG> : rb_k_pt_shadow<true, false, RB_ROOTS_TREE>)
										: (bigTree ? rb_k_pt_shadow<false, true, RB_ROOTS_TREE_BIG> : rb_k_pt_shadow<false, true, RB_ROOTS_TREE>);
		if(profCounting) kShadow = masked ? (ordered ? rb_k_pt_shadow<true, true, RB_ROOTS_LAZY, true> : (treeShadow ? (bigTree ? rb_k_pt_shadow<true, false, RB_ROOTS_TREE_BIG, true> : rb_k_pt_shadow<true, false, RB_ROOTS_TREE, true>) : rb_k_pt_shadow<true, false, RB_ROOTS_LAZY, true>))
										  : (treeShadow ? (bigTree ? rb_k_pt_shadow<false, true, RB_ROOTS_TREE_BIG, true> : rb_k_pt_shadow<false, true, RB_ROOTS_TREE, true>) : rb_k_pt_shadow<false, true, RB_ROOTS_LAZY, true>);
		if(feed) {
			kShadow = blocks ? (masked ? (ordered ? rb_k_pt_shadow_feed<true, true, RB_ROOTS_BLOCKS> : rb_k_pt_shadow_feed<true, false, RB_ROOTS_BLOCKS>) : rb_k_pt_shadow_feed<false, true, RB_ROOTS_BLOCKS>)
							 : (masked ? (ordered ? rb_k_pt_shadow_feed<true, true, RB_ROOTS_LAZY> : rb_k_pt_shadow_feed<true, false, RB_ROOTS_LAZY>) : rb_k_pt_shadow_feed<false, true, RB_ROOTS_LAZY>);
			smemShadow = smem + (RB_PT_BLOCK / 32) * RB_FEED_SHADOW_BYTES;
		}
		st = rb_pt_grid(kShadow, smemShadow, pt->sms, gridShadow);
		if(st) return st;
		const bool wideShadow  = wideTrace && !masked; // masked scenes need the closest occluder: binary kernel
		unsigned gridShadowWide = 0;
		if(wideShadow) {
			st = rb_pt_grid(rb_k_pt_shadow_wide, smemWide, pt->sms, gridShadowWide);
			if(st) return st;
		}
		const int votes = rc->leafVotes > 0 ? rc->leafVotes : 33;
		const int steps = std::getenv("RB_NODE_STEPS") ? (rc->nodeSteps > 0 ? rc->nodeSteps : 1) : 4;

		if(lanes > 1) {
			RB_CUDA(cudaEventRecord(pt->fork, s));
			for(int l = 0; l < lanes; ++l) RB_CUDA(cudaStreamWaitEvent(pt->lane[l], pt->fork, 0));
		}
		uint64_t batch = 0;
		for(uint32_t sample0 = 0; sample0 < localSamples; sample0 += schunk) {
			const uint32_t sc_n = (localSamples - sample0 < schunk) ? (localSamples - sample0) : schunk;
			for(uint64_t pixel0 = 0; pixel0 < localPixels; pixel0 += pixelsPerBatch, ++batch) {
				RbBatch b;
				b.pixel0  = uint32_t(pixel0);
				b.npixels = uint32_t((localPixels - pixel0 < pixelsPerBatch) ? (localPixels - pixel0) : pixelsPerBatch);
				b.sample0 = sample0;
				b.schunk  = sc_n;
				b.mSchunk = rb_div_magic(sc_n);
				const int l				 = lanes > 1 ? int(batch % lanes) : 0;
				const cudaStream_t ls	 = lanes > 1 ? pt->lane[l] : s;
				const RbPathBuffers & pb = pt->pb[l];
				RB_CUDA(cudaMemsetAsync(pb.counters, 0, 4 * (RB_PT_MAX_DEPTH + 1) * sizeof(uint32_t), ls));
				RB_CUDA(mark(0, ls));
				rb_k_pt_raygen<<<gridFlat, RB_PT_BLOCK, 0, ls>>>(f, b, pb);
				++launches;
				// The shadow wavefront of level d only feeds the colour that shade(d+1) reads, so it runs on a side stream
				// next to trace(d+1): trace -> shade -> [shadow || next trace] -> shade ...
				// (small batches stay on one stream: the event traffic would cost more than the overlap gives)
				const cudaStream_t ss = (!pt->prof && uint64_t(b.npixels) * b.schunk >= (uint64_t(1) << 20)) ? pt->side[l] : ls;
				bool shadowPending	  = false;
				for(uint32_t did = 0; did < depth; ++did) {
					RB_CUDA(mark(1, ls));
					if(wideTrace) rb_k_pt_trace_wide<<<gridTraceWide, RB_PT_BLOCK, smemWide, ls>>>(rc->dev, rc->devWide, pb, pb.queue[did & 1u], did, threshold, votes > 32 ? 6 : votes, pt->d_stats);
					else kTrace<<<gridTrace, RB_PT_BLOCK, smemTrace, ls>>>(rc->dev, pb, pb.queue[did & 1u], did, threshold, votes, steps, pt->d_stats);
					if(shadowPending) RB_CUDA(cudaStreamWaitEvent(ls, pt->shadowed[l], 0)); // shade overwrites the shadow entries
					RB_CUDA(mark(2, ls));
					rb_k_pt_shade<<<gridShade, RB_PT_BLOCK, 0, ls>>>(f, b, pb, sc->dev, did, pt->d_stats);
					launches += 2;
					if(sc->dev.nLights > 0) {
						if(ss != ls) {
							RB_CUDA(cudaEventRecord(pt->shaded[l], ls));
							RB_CUDA(cudaStreamWaitEvent(ss, pt->shaded[l], 0));
						}
						RB_CUDA(mark(3, ss));
						if(wideShadow) rb_k_pt_shadow_wide<<<gridShadowWide, RB_PT_BLOCK, smemWide, ss>>>(rc->dev, rc->devWide, pb, did, threshold, votes > 32 ? 6 : votes, pt->d_stats);
						else kShadow<<<gridShadow, RB_PT_BLOCK, smemShadow, ss>>>(rc->dev, pb, sc->dev, did, threshold, votes, steps, pt->d_stats);
						if(ss != ls) {
							RB_CUDA(cudaEventRecord(pt->shadowed[l], ss));
							shadowPending = true;
						}
						++launches;
					}
				}
				if(shadowPending) RB_CUDA(cudaStreamWaitEvent(ls, pt->shadowed[l], 0));
				RB_CUDA(mark(4, ls));
				rb_k_pt_accumulate<<<gridFlat, RB_PT_BLOCK, 0, ls>>>(f, b, pb, d_sum);
				++launches;
				RB_CUDA(mark(5, ls)); // end of the batch
				RB_CUDA(cudaGetLastError());
			}
		}
		if(lanes > 1) {
			for(int l = 0; l < lanes; ++l) {
				RB_CUDA(cudaEventRecord(pt->join[l], pt->lane[l]));
				RB_CUDA(cudaStreamWaitEvent(s, pt->join[l], 0));
			}
		}
	}
	pt->stats.kernel_launches = launches;
	pt->stats.samples		  = localPixels * localSamples;
	return RB_OK;
}

extern "C" rb_status rb_pathtracer_resolve_device(rb_pathtracer * pt, const float * d_sum, uint32_t width, uint32_t height, uint32_t samples, uint32_t components, float * d_pixels, const rb_render_opts * opts, void * stream) {
	if(!pt || !d_sum || !d_pixels) return rb_fail(RB_ERR_INVALID, "rb_pathtracer_resolve_device: null argument");
	RB_REQUIRE_DEVICE();
	if(components < 3) return rb_fail(RB_ERR_INVALID, "rb_pathtracer_resolve_device: expected at least 3 components");
	RbFrame f;
	rb_camera cam;
	std::memset(&cam, 0, sizeof(cam));
	rb_status st = rb_pt_frame(&cam, width, height, samples, 1, opts, f);
	if(st) return st;
	rb_k_pt_resolve<<<unsigned(pt->sms * 4), 256, 0, static_cast<cudaStream_t>(stream)>>>(f, d_sum, components, d_pixels);
	RB_CUDA(cudaGetLastError());
	pt->stats.kernel_launches += 1;
	return RB_OK;
}

extern "C" rb_status rb_pathtracer_render(rb_pathtracer * pt, const rb_camera * camera, uint32_t width, uint32_t height, uint32_t samples, uint32_t depth, uint32_t components, float * pixels, const rb_render_opts * opts) {
	if(!pt || !pixels) return rb_fail(RB_ERR_INVALID, "rb_pathtracer_render: null argument");
	RB_REQUIRE_DEVICE();
	if(components < 3) return rb_fail(RB_ERR_INVALID, "rb_pathtracer_render: expected an RGB image (components >= 3)"); // PathTracer.cpp:166-168
	RB_CUDA(cudaSetDevice(pt->scene->raycaster->device));
	if(opts && opts->gpus > 1) {
		if(pt->scene->tablesDirty) { // the replicas are keyed to the edit counter: bring the primary up to date first
			rb_status ss = rb_scene_sync(pt->scene);
			if(ss) return ss;
		}
		return rb_pt_multi_render(pt, camera, width, height, samples, depth, components, pixels, opts);
	}
	cudaStream_t s = cudaStreamPerThread;
	const size_t npx = size_t(width) * height;
	if(pt->frameSumFloats < npx * 3) {
		cudaFree(pt->d_frameSum);
		pt->d_frameSum	   = nullptr;
		pt->frameSumFloats = 0;
		RB_CUDA(cudaMalloc(&pt->d_frameSum, npx * 3 * sizeof(float)));
		pt->frameSumFloats = npx * 3;
	}
	if(pt->framePixelFloats < npx * components) {
		cudaFree(pt->d_framePixels);
		pt->d_framePixels	 = nullptr;
		pt->framePixelFloats = 0;
		RB_CUDA(cudaMalloc(&pt->d_framePixels, npx * components * sizeof(float)));
		pt->framePixelFloats = npx * components;
	}
	float *d_sum = pt->d_frameSum, *d_pixels = pt->d_framePixels;
	cudaEvent_t e0, e1;
	cudaEventCreate(&e0);
	cudaEventCreate(&e1);
	rb_status st = RB_OK;
	do {
		if(cudaMemcpyAsync(d_pixels, pixels, npx * components * sizeof(float), cudaMemcpyHostToDevice, s) != cudaSuccess) { st = rb_fail(RB_ERR_CUDA, "upload failed"); break; }
		cudaEventRecord(e0, s);
		st = rb_pathtracer_render_device(pt, camera, width, height, samples, depth, d_sum, opts, s);
		if(st) break;
		st = rb_pathtracer_resolve_device(pt, d_sum, width, height, samples, components, d_pixels, opts, s);
		if(st) break;
		cudaEventRecord(e1, s);
		if(cudaMemcpyAsync(pixels, d_pixels, npx * components * sizeof(float), cudaMemcpyDeviceToHost, s) != cudaSuccess) { st = rb_fail(RB_ERR_CUDA, "download failed"); break; }
		unsigned long long hs[4] = {0, 0, 0, 0};
		cudaMemcpyAsync(hs, pt->d_stats, sizeof(hs), cudaMemcpyDeviceToHost, s);
		const cudaError_t e = cudaStreamSynchronize(s);
		if(e != cudaSuccess) { st = rb_fail(RB_ERR_CUDA, "render failed: %s", cudaGetErrorString(e)); break; }
		float ms = 0.0f;
		cudaEventElapsedTime(&ms, e0, e1);
		pt->stats.rays_closest	 = hs[0];
		pt->stats.rays_shadow	 = hs[1];
		pt->stats.shaded_hits	 = hs[2];
		pt->stats.device_seconds = double(ms) * 1e-3;
	} while(false);
	cudaEventDestroy(e0);
	cudaEventDestroy(e1);
	return st;
}

// One frame taken apart (bench.py's roofline; never part of a timed region): pass 1 runs the ordinary kernels on ONE
// stream with an event before each (no second lane, no side stream, so the classes do not overlap and their times add
// up to the serialised frame); pass 2 runs the counting copies of the two traversal kernels.
extern "C" rb_status rb_pathtracer_profile_frame(rb_pathtracer * pt, const rb_camera * camera, uint32_t width, uint32_t height, uint32_t samples, uint32_t depth, const rb_render_opts * opts, rb_pt_profile * out) {
	if(!pt || !camera || !out) return rb_fail(RB_ERR_INVALID, "rb_pathtracer_profile_frame: null argument");
	RB_REQUIRE_DEVICE();
	std::memset(out, 0, sizeof(*out));
	RB_CUDA(cudaSetDevice(pt->scene->raycaster->device));
	float * d_sum = nullptr;
	RB_CUDA(cudaMalloc(&d_sum, size_t(width) * height * 3 * sizeof(float)));
	rb_pathtracer::Profile prof;
	cudaStream_t s = cudaStreamPerThread;
	rb_status st   = RB_OK;
	for(int pass = 0; pass < 2 && !st; ++pass) {
		prof.timing	  = pass == 0;
		prof.counting = pass == 1;
		pt->prof	  = &prof;
		st			  = rb_pathtracer_render_device(pt, camera, width, height, samples, depth, d_sum, opts, s);
		pt->prof	  = nullptr;
		if(st) break;
		if(cudaStreamSynchronize(s) != cudaSuccess) { st = rb_fail(RB_ERR_CUDA, "rb_pathtracer_profile_frame: %s", cudaGetErrorString(cudaGetLastError())); break; }
		if(pass == 0) {
			for(size_t k = 0; k + 1 < prof.events.size(); ++k) {
				const int kind = prof.kinds[k];
				if(kind > 4) continue; // batch end marker -> next batch's first event: nothing in between
				float ms = 0.0f;
				cudaEventElapsedTime(&ms, prof.events[k], prof.events[k + 1]);
				out->ms[kind] += ms;
				out->launches[kind] += 1;
			}
		} else {
			unsigned long long hs[RB_PT_NSTATS];
			if(cudaMemcpy(hs, pt->d_stats, sizeof(hs), cudaMemcpyDeviceToHost) != cudaSuccess) { st = rb_fail(RB_ERR_CUDA, "rb_pathtracer_profile_frame: counters unreadable"); break; }
			out->rays_closest = hs[0]; out->rays_shadow = hs[1]; out->shaded_hits = hs[2];
			out->trace_interior = hs[4]; out->trace_tris = hs[5]; out->trace_roots = hs[6];
			out->shadow_interior = hs[7]; out->shadow_tris = hs[8]; out->shadow_roots = hs[9];
			out->samples = pt->stats.samples;
		}
	}
	for(cudaEvent_t e : prof.events) cudaEventDestroy(e);
	cudaFree(d_sum);
	return st;
}

extern "C" rb_status rb_pathtracer_stats(const rb_pathtracer * pt, rb_render_stats * out) {
	if(!pt || !out) return rb_fail(RB_ERR_INVALID, "rb_pathtracer_stats: null argument");
	*out = pt->stats;
	if(rb_have_device() && pt->d_stats) {
		// Counters of the last render_device call (blocks until that work has finished).
		unsigned long long hs[4] = {0, 0, 0, 0};
		RB_CUDA(cudaDeviceSynchronize());
		RB_CUDA(cudaMemcpy(hs, pt->d_stats, sizeof(hs), cudaMemcpyDeviceToHost));
		out->rays_closest = hs[0];
		out->rays_shadow  = hs[1];
		out->shaded_hits  = hs[2];
	}
	return RB_OK;
}


// ------------------------------------------------------------------------------------------------ live rendering
// PathTracerApp's "Live render" (src/apps/pathtracer/PathTracerApp.cpp:57-98,127-136) calls render() for a whole small frame
// every UI frame and uploads the float image again. Here a frame is refined pass by pass: pass p of `pass_count` traces the
// stratification cells sid % pass_count == p and ADDS them to sums that stay on the device; what comes back per displayed
// frame is 8-bit pixels (mean so far -> tonemap 1 - exp(-c) -> sRGB -> byte, one fused kernel), 3 or 4 bytes per pixel
// instead of 12 or 16 up and down. After pass_count passes the sums equal those of one render() call with `samples`.
__global__ void rb_k_live_display(const float * __restrict__ sum, size_t npixels, float cells, uint32_t components, int tonemap, int srgb, uint8_t * __restrict__ out) {
	for(size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x; p < npixels; p += (size_t)gridDim.x * blockDim.x) {
		for(uint32_t c = 0; c < 3; ++c) {
			float value = sum[3 * p + c] / cells; // render.rgb(x, y) / float(samples) (PathTracer.cpp:291-292) over the cells traced so far
			if(tonemap) value = 1.0f - expf(-value * 1.f);
			if(srgb) value = powf(value, 1.0f / 2.2f);
			out[p * components + c] = static_cast<unsigned char>(std_min(255.0f, std_max(0.0f, 255.0f * value)));
		}
		if(components == 4) out[p * 4 + 3] = 255;
	}
}

extern "C" rb_status rb_pathtracer_live_begin(rb_pathtracer * pt, uint32_t width, uint32_t height) {
	if(!pt) return rb_fail(RB_ERR_INVALID, "rb_pathtracer_live_begin: null handle");
	RB_REQUIRE_DEVICE();
	if(width == 0 || height == 0) return rb_fail(RB_ERR_INVALID, "rb_pathtracer_live_begin: empty frame");
	RB_CUDA(cudaSetDevice(pt->scene->raycaster->device));
	if(width != pt->liveW || height != pt->liveH) {
		cudaFree(pt->d_liveSum);
		cudaFree(pt->d_liveBytes);
		pt->d_liveSum = nullptr;
		pt->d_liveBytes = nullptr;
		pt->liveW = pt->liveH = 0;
		RB_CUDA(cudaMalloc(&pt->d_liveSum, size_t(width) * height * 3 * sizeof(float)));
		RB_CUDA(cudaMalloc(&pt->d_liveBytes, size_t(width) * height * 4));
		pt->liveW = width;
		pt->liveH = height;
	}
	pt->livePasses = 0; // the first pass overwrites the sums
	return RB_OK;
}

extern "C" rb_status rb_pathtracer_live_pass(rb_pathtracer * pt, const rb_camera * camera, uint32_t samples, uint32_t depth, uint32_t pass_count, const rb_render_opts * opts) {
	if(!pt || !camera) return rb_fail(RB_ERR_INVALID, "rb_pathtracer_live_pass: null argument");
	RB_REQUIRE_DEVICE();
	if(!pt->d_liveSum) return rb_fail(RB_ERR_INVALID, "rb_pathtracer_live_pass: call rb_pathtracer_live_begin first");
	if(pass_count == 0 || pass_count > samples) return rb_fail(RB_ERR_INVALID, "rb_pathtracer_live_pass: pass_count must be in [1, samples]");
	if(opts && (opts->tile_count > 1 || opts->sample_count > 1 || opts->gpus > 1)) return rb_fail(RB_ERR_INVALID, "rb_pathtracer_live_pass: the passes are the split; leave tile_* / sample_* / gpus at 0");
	// A moved camera (or other frame parameters) starts the frame over, as a new render() call does in the reference's loop.
	const bool same = pt->livePasses > 0 && std::memcmp(&pt->liveCamera, camera, sizeof(rb_camera)) == 0 && pt->liveSamples == samples && pt->liveDepth == depth && pt->livePassCount == pass_count;
	if(!same) pt->livePasses = 0;
	if(pt->livePasses >= pass_count) return RB_OK; // converged: every cell of the stratification grid has been traced
	rb_render_opts o;
	if(opts) o = *opts;
	else std::memset(&o, 0, sizeof(o));
	o.sample_index = pt->livePasses;
	o.sample_count = pass_count;
	o.accumulate   = pt->livePasses > 0 ? 1u : 0u;
	rb_status st = rb_pathtracer_render_device(pt, camera, pt->liveW, pt->liveH, samples, depth, pt->d_liveSum, &o, cudaStreamPerThread);
	if(st) return st;
	pt->liveCamera	  = *camera;
	pt->liveSamples	  = samples;
	pt->liveDepth	  = depth;
	pt->livePassCount = pass_count;
	pt->livePasses++;
	return RB_OK;
}

extern "C" rb_status rb_pathtracer_live_display(rb_pathtracer * pt, uint32_t components, int tonemap, int srgb, uint8_t * out, uint32_t * passes_done) {
	if(!pt || !out) return rb_fail(RB_ERR_INVALID, "rb_pathtracer_live_display: null argument");
	RB_REQUIRE_DEVICE();
	if(components != 3 && components != 4) return rb_fail(RB_ERR_INVALID, "rb_pathtracer_live_display: components must be 3 or 4");
	if(!pt->d_liveSum || pt->livePasses == 0) return rb_fail(RB_ERR_INVALID, "rb_pathtracer_live_display: no pass has been traced");
	RB_CUDA(cudaSetDevice(pt->scene->raycaster->device));
	cudaStream_t s = cudaStreamPerThread;
	const size_t npx = size_t(pt->liveW) * pt->liveH;
	// Cells traced so far: pass p holds the cells sid % pass_count == p of the (power-of-two) sample count render() rounds to.
	RbFrame f;
	rb_status st = rb_pt_frame(&pt->liveCamera, pt->liveW, pt->liveH, pt->liveSamples, 1, nullptr, f);
	if(st) return st;
	uint64_t cells = 0;
	for(uint32_t p = 0; p < pt->livePasses; ++p) cells += (f.samples > p) ? (f.samples - p + pt->livePassCount - 1) / pt->livePassCount : 0;
	rb_k_live_display<<<unsigned(pt->sms * 8), 256, 0, s>>>(pt->d_liveSum, npx, float(cells), components, tonemap, srgb, pt->d_liveBytes);
	RB_CUDA(cudaGetLastError());
	RB_CUDA(cudaMemcpyAsync(out, pt->d_liveBytes, npx * components, cudaMemcpyDeviceToHost, s));
	RB_CUDA(cudaStreamSynchronize(s));
	if(passes_done) *passes_done = pt->livePasses;
	return RB_OK;
}
