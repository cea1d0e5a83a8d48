/* TEST INFRASTRUCTURE — NOT PRODUCT CODE. See rendu_oracle.h.
 * The two pieces of the reference's behaviour that are *defined by libstdc++* and therefore restated with the same
 * library calls instead of plain C: the BVH build order (std::partition / std::nth_element) and the random stream
 * (std::mt19937 + std::uniform_*_distribution, Random.cpp:21-27,72-80).
 * Raycaster::updateHierarchy (Raycaster.cpp:54-142) restated. This one function is C++ because the reference's tree is
 * *defined* by libstdc++: its triangle order is whatever std::partition (:107) and std::nth_element (:122) produce on
 * the 64-byte TriangleInfos records, so the oracle calls the same algorithms on the same record type. */
#include "rendu_oracle.h"
#include <algorithm>
#include <random>
#include <vector>

namespace {
struct Set { size_t id, begin, count; };
inline float gmin(float x, float y) { return (y < x) ? y : x; }
inline float gmax(float x, float y) { return (x < y) ? y : x; }
inline float centroid(const orc_tri & t, int axis) { return 0.5f * (t.bmin[axis] + t.bmax[axis]); } /* Bounds.cpp:61-63 */
}

extern "C" size_t orc_build(orc_tri * tris, size_t ntris, const uint64_t * meshFirst, const uint64_t * meshTris, uint32_t meshCount, orc_node * nodes) {
	(void)ntris;
	size_t nnodes = 0;
	std::vector<Set> remaining;
	for(uint32_t m = 0; m < meshCount; ++m) { /* :64-69 */
		nodes[nnodes] = orc_node();
		remaining.push_back({nnodes, size_t(meshFirst[m]), size_t(meshTris[m])});
		++nnodes;
	}
	while(!remaining.empty()) {
		const Set cur = remaining.back();
		remaining.pop_back();
		orc_tri * first = tris + cur.begin;
		float lo[3], hi[3];
		for(int k = 0; k < 3; ++k) { lo[k] = first[0].bmin[k]; hi[k] = first[0].bmax[k]; }
		for(size_t t = 1; t < cur.count; ++t) { /* :79-82 */
			for(int k = 0; k < 3; ++k) { lo[k] = gmin(lo[k], first[t].bmin[k]); hi[k] = gmax(hi[k], first[t].bmax[k]); }
		}
		orc_node & node = nodes[cur.id];
		for(int k = 0; k < 3; ++k) { node.bmin[k] = lo[k]; node.bmax[k] = hi[k]; }
		if(cur.count < 3) { /* :88-91 */
			node.leaf = 1; node.left = cur.begin; node.right = cur.count;
			continue;
		}
		node.leaf = 0;
		size_t split = 0;
		const float sx = hi[0] - lo[0], sy = hi[1] - lo[1], sz = hi[2] - lo[2];
		const int axis = (sx >= sy && sx >= sz) ? 0 : (sy >= sz ? 1 : 2); /* :97-98 */
		if(cur.count >= 5) { /* :100-115 */
			float abscisse = 0.0f;
			for(size_t t = 0; t < cur.count; ++t) abscisse += centroid(first[t], axis);
			abscisse /= float(cur.count);
			orc_tri * mid = std::partition(first, first + cur.count, [abscisse, axis](const orc_tri & t0) { return centroid(t0, axis) < abscisse; });
			split = size_t(mid - first);
		}
		if(split == 0 || split == cur.count || cur.count < 5) { /* :120-125 */
			split = cur.count / 2;
			std::nth_element(first, first + split, first + cur.count, [axis](const orc_tri & a, const orc_tri & b) { return centroid(a, axis) < centroid(b, axis); });
		}
		const size_t leftPos = nnodes++;
		nodes[leftPos] = orc_node();
		nodes[cur.id].left = leftPos;
		remaining.push_back({leftPos, cur.begin, split});
		const size_t rightPos = nnodes++;
		nodes[rightPos] = orc_node();
		nodes[cur.id].right = rightPos;
		remaining.push_back({rightPos, cur.begin + split, cur.count - split});
	}
	return nnodes;
}

/* ---------------------------------------------------------------- Random (src/engine/generation/Random.cpp) */
namespace {
struct MtState { std::mt19937 mt; };
float mtFloat(void * s) { return std::uniform_real_distribution<float>(0.0f, 1.0f)(static_cast<MtState *>(s)->mt); } /* Random.cpp:25-27 */
int mtInt(void * s, int lo, int hi) { return std::uniform_int_distribution<int>(lo, hi)(static_cast<MtState *>(s)->mt); } /* :21-23 */
const orc_rng_stream kMtStream = {mtFloat, mtInt};
}

/* Random::seed(seed) (Random.cpp:9-15) seeds the shared generator and re-creates the CALLING thread's generator, which
 * draws the first local seed (LocalMT19937, :72-80). The worker thread forParallel then spawns draws the second one. */
extern "C" void * orc_mt_create(unsigned int seed) {
	std::mt19937 shared(seed);
	(void)std::uniform_int_distribution<>()(shared);
	const int workerSeed = std::uniform_int_distribution<>()(shared);
	MtState * st = new MtState();
	st->mt = std::mt19937(workerSeed);
	return st;
}
extern "C" void orc_mt_destroy(void * state) { delete static_cast<MtState *>(state); }
extern "C" const orc_rng_stream * orc_mt_stream(void) { return &kMtStream; }
extern "C" void orc_mt_floats(unsigned int seed, size_t n, float * out) {
	std::mt19937 shared(seed);
	MtState st;
	st.mt = std::mt19937(std::uniform_int_distribution<>()(shared));
	for(size_t i = 0; i < n; ++i) out[i] = mtFloat(&st);
}
