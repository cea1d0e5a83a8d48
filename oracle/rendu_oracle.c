/* TEST INFRASTRUCTURE — NOT PRODUCT CODE. See rendu_oracle.h for scope and pinning.
 * Raycaster part of the oracle. Compiled with -O2 -ffp-contract=off and no -march so that, like the reference build
 * (premake5.lua:49-51,66-69), no multiply-add is fused. glm operation order: dot = (x+y)+z, normalize = v*(1/sqrt(dot)),
 * cross as written in func_geometric.inl:70-80, glm::min(x,y)=(y<x)?y:x, glm::max(x,y)=(x<y)?y:x (func_common.inl:17-30).
 */
#include "rendu_oracle.h"
#include <float.h>
#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

static float glm_min(float x, float y) { return (y < x) ? y : x; }
static float glm_max(float x, float y) { return (x < y) ? y : x; }
static float std_max(float a, float b) { return (a < b) ? b : a; }
static float std_min(float a, float b) { return (b < a) ? b : a; }
static float dot3(const float * a, const float * b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
static void cross3(const float * x, const float * y, float * o) {
	o[0] = x[1] * y[2] - y[1] * x[2];
	o[1] = x[2] * y[0] - y[2] * x[0];
	o[2] = x[0] * y[1] - y[0] * x[1];
}

/* Intersection.cpp:3-5 */
void orc_ray(const float * origin, const float * direction, float * r) {
	const float inv = 1.0f / sqrtf(dot3(direction, direction));
	for(int k = 0; k < 3; ++k) {
		r[k]	 = origin[k];
		r[3 + k] = direction[k] * inv;
		r[6 + k] = 1.0f / r[3 + k];
	}
}

/* Intersection.cpp:22-32 */
int orc_box(const float * r, const float * bmin, const float * bmax, float mini, float maxi) {
	float minF[3], maxF[3];
	for(int k = 0; k < 3; ++k) {
		const float minRatio = (bmin[k] - r[k]) * r[6 + k];
		const float maxRatio = (bmax[k] - r[k]) * r[6 + k];
		minF[k] = glm_min(minRatio, maxRatio);
		maxF[k] = glm_max(minRatio, maxRatio);
	}
	const float closest	 = std_max(minF[0], std_max(minF[1], minF[2]));
	const float furthest = std_min(maxF[0], std_min(maxF[1], maxF[2]));
	return std_max(closest, mini) <= std_min(furthest, maxi);
}

/* Raycaster.cpp:232-262 */
int orc_tri_test(const orc_bvh * b, const float * r, const orc_tri * tri, float mini, float maxi, float * t, float * u, float * v) {
	const float * v0 = b->verts + 3 * tri->v0;
	const float * v1 = b->verts + 3 * tri->v1;
	const float * v2 = b->verts + 3 * tri->v2;
	float v01[3], v02[3], p[3], q[3], rr[3];
	for(int k = 0; k < 3; ++k) {
		v01[k] = v1[k] - v0[k];
		v02[k] = v2[k] - v0[k];
	}
	cross3(r + 3, v02, p);
	const float det = dot3(v01, p);
	if(fabsf(det) < FLT_EPSILON) return 0;
	const float invDet = 1.0f / det;
	for(int k = 0; k < 3; ++k) q[k] = r[k] - v0[k];
	*u = invDet * dot3(q, p);
	if(*u < 0.0f || *u > 1.0f) return 0;
	cross3(q, v01, rr);
	*v = invDet * dot3(r + 3, rr);
	if(*v < 0.0f || (*u + *v) > 1.0f) return 0;
	*t = invDet * dot3(v02, rr);
	return (*t > mini && *t < maxi);
}

static orc_hit miss(void) {
	orc_hit h;
	memset(&h, 0, sizeof(h));
	h.dist = FLT_MAX; /* Raycaster.cpp:6-8 */
	return h;
}

typedef struct { size_t * data; size_t size, cap; } stack_t_;
static void push(stack_t_ * s, size_t v) {
	if(s->size == s->cap) {
		s->cap	= s->cap ? 2 * s->cap : 64;
		s->data = (size_t *)realloc(s->data, s->cap * sizeof(size_t));
	}
	s->data[s->size++] = v;
}

/* Raycaster.cpp:144-185 (closest) and :187-224 (any) share this loop. */
static int traverse(const orc_bvh * b, const float * origin, const float * direction, float mini, float maxi, int any, orc_hit * best) {
	float r[9];
	orc_ray(origin, direction, r);
	stack_t_ st = {0, 0, 0};
	for(size_t nid = 0; nid < b->meshCount; ++nid) {
		if(!orc_box(r, b->nodes[nid].bmin, b->nodes[nid].bmax, mini, maxi)) continue;
		push(&st, nid);
	}
	*best = miss();
	while(st.size) {
		const orc_node * node = &b->nodes[st.data[--st.size]];
		if(node->leaf) {
			for(size_t tid = 0; tid < node->right; ++tid) {
				const orc_tri * tri = &b->tris[node->left + tid];
				float t, u, v;
				if(orc_tri_test(b, r, tri, mini, maxi, &t, &u, &v)) {
					if(any) {
						free(st.data);
						return 1;
					}
					if(t < best->dist) {
						best->hit = 1; best->dist = t; best->u = u; best->v = v;
						best->w = 1.0f - u - v; /* Raycaster.cpp:10-12 */
						best->localId = tri->localId; best->meshId = tri->meshId;
						best->internalId = (uint64_t)(node->left + tid);
						maxi = t;
					}
				}
			}
			continue;
		}
		if(orc_box(r, b->nodes[node->left].bmin, b->nodes[node->left].bmax, mini, maxi)) push(&st, node->left);
		if(orc_box(r, b->nodes[node->right].bmin, b->nodes[node->right].bmax, mini, maxi)) push(&st, node->right);
	}
	free(st.data);
	return (int)best->hit;
}

/* RaycasterVisualisation::getRayLevels (src/engine/raycaster/RaycasterVisualisation.cpp:41-77): the closest-hit loop
 * again, recording every node it pops with its depth (roots are depth 0), in pop order. At most `cap` records are
 * written; *count receives the number of nodes visited. The boxes the viewer draws are b->nodes[id].bmin/bmax. */
void orc_ray_levels(const orc_bvh * b, const float * origin, const float * direction, float mini, float maxi, size_t cap, uint64_t * nodeIds, uint32_t * depths, size_t * count, orc_hit * hit) {
	float r[9];
	orc_ray(origin, direction, r);
	stack_t_ st = {0, 0, 0}, dp = {0, 0, 0};
	for(size_t nid = 0; nid < b->meshCount; ++nid) {
		if(!orc_box(r, b->nodes[nid].bmin, b->nodes[nid].bmax, mini, maxi)) continue;
		push(&st, nid);
		push(&dp, 0);
	}
	orc_hit best = miss();
	size_t visited = 0;
	while(st.size) {
		const size_t id		  = st.data[--st.size];
		const size_t depth	  = dp.data[--dp.size];
		const orc_node * node = &b->nodes[id];
		if(visited < cap) {
			nodeIds[visited] = (uint64_t)id;
			depths[visited]	 = (uint32_t)depth;
		}
		++visited;
		if(node->leaf) {
			for(size_t tid = 0; tid < node->right; ++tid) {
				const orc_tri * tri = &b->tris[node->left + tid];
				float t, u, v;
				if(orc_tri_test(b, r, tri, mini, maxi, &t, &u, &v) && t < best.dist) {
					best.hit = 1; best.dist = t; best.u = u; best.v = v;
					best.w = 1.0f - u - v;
					best.localId = tri->localId; best.meshId = tri->meshId;
					best.internalId = (uint64_t)(node->left + tid);
					maxi = t;
				}
			}
			continue;
		}
		if(orc_box(r, b->nodes[node->left].bmin, b->nodes[node->left].bmax, mini, maxi)) { push(&st, node->left); push(&dp, depth + 1); }
		if(orc_box(r, b->nodes[node->right].bmin, b->nodes[node->right].bmax, mini, maxi)) { push(&st, node->right); push(&dp, depth + 1); }
	}
	free(st.data);
	free(dp.data);
	*count = visited;
	*hit   = best;
}

orc_hit orc_intersect(const orc_bvh * b, const float * origin, const float * direction, float mini, float maxi) {
	orc_hit h;
	traverse(b, origin, direction, mini, maxi, 0, &h);
	return h;
}

int orc_intersect_any(const orc_bvh * b, const float * origin, const float * direction, float mini, float maxi) {
	orc_hit h;
	return traverse(b, origin, direction, mini, maxi, 1, &h);
}

/* Raycaster.cpp:226-230 */
int orc_visible(const orc_bvh * b, const float * p0, const float * p1) {
	const float d[3] = {p1[0] - p0[0], p1[1] - p0[1], p1[2] - p0[2]};
	const float maxi = sqrtf(dot3(d, d));
	return !orc_intersect_any(b, p0, d, 0.0001f, maxi);
}

typedef struct {
	const orc_bvh * b; const float * org; const float * dir; float mini, maxi; size_t begin, end; orc_hit * out; uint8_t * outAny;
} job_t;

static void * run_job(void * arg) {
	job_t * j = (job_t *)arg;
	for(size_t i = j->begin; i < j->end; ++i) {
		if(j->out) j->out[i] = orc_intersect(j->b, j->org + 3 * i, j->dir + 3 * i, j->mini, j->maxi);
		else j->outAny[i] = (uint8_t)orc_intersect_any(j->b, j->org + 3 * i, j->dir + 3 * i, j->mini, j->maxi);
	}
	return NULL;
}

static void run_batch(const orc_bvh * b, const float * org, const float * dir, float mini, float maxi, size_t n, int threads, orc_hit * out, uint8_t * outAny) {
	if(threads < 1) threads = 1;
	if((size_t)threads > n) threads = n ? (int)n : 1;
	job_t * jobs	= (job_t *)calloc((size_t)threads, sizeof(job_t));
	pthread_t * tid = (pthread_t *)calloc((size_t)threads, sizeof(pthread_t));
	const size_t span = n / (size_t)threads; /* contiguous blocks, last takes the remainder (System.hpp:88-101) */
	for(int t = 0; t < threads; ++t) {
		job_t j = {b, org, dir, mini, maxi, (size_t)t * span, (t == threads - 1) ? n : (size_t)(t + 1) * span, out, outAny};
		jobs[t] = j;
		if(threads == 1) run_job(&jobs[t]);
		else pthread_create(&tid[t], NULL, run_job, &jobs[t]);
	}
	if(threads > 1) for(int t = 0; t < threads; ++t) pthread_join(tid[t], NULL);
	free(jobs);
	free(tid);
}

void orc_intersect_batch(const orc_bvh * b, const float * org, const float * dir, float mini, float maxi, size_t n, int threads, orc_hit * out) {
	run_batch(b, org, dir, mini, maxi, n, threads, out, NULL);
}

void orc_intersect_any_batch(const orc_bvh * b, const float * org, const float * dir, float mini, float maxi, size_t n, int threads, uint8_t * out) {
	run_batch(b, org, dir, mini, maxi, n, threads, NULL, out);
}

/* Raycaster.cpp:14-52; model*vec4(v,1) evaluated as glm does: (m0*x + m1*y) + (m2*z + m3*w) (type_mat4x4.inl:561-571). */
size_t orc_add_mesh(const float * positions, size_t nverts, const uint32_t * indices, size_t nindices, const float * m,
	float * vertsOut, size_t vertOffset, orc_tri * trisOut, uint32_t meshId) {
	int identity = 1;
	if(m) {
		for(int k = 0; k < 16; ++k) identity = identity && (m[k] == ((k % 5 == 0) ? 1.0f : 0.0f));
	}
	for(size_t v = 0; v < nverts; ++v) {
		float * d = vertsOut + 3 * (vertOffset + v);
		const float x = positions[3 * v], y = positions[3 * v + 1], z = positions[3 * v + 2];
		if(identity) {
			d[0] = x; d[1] = y; d[2] = z;
		} else {
			for(int r = 0; r < 3; ++r) d[r] = (m[r] * x + m[4 + r] * y) + (m[8 + r] * z + m[12 + r] * 1.0f);
		}
	}
	const size_t ntri = nindices / 3;
	for(size_t t = 0; t < ntri; ++t) {
		orc_tri * ti = &trisOut[t];
		memset(ti, 0, sizeof(*ti));
		ti->v0 = vertOffset + indices[3 * t];
		ti->v1 = vertOffset + indices[3 * t + 1];
		ti->v2 = vertOffset + indices[3 * t + 2];
		ti->localId = 3 * t;
		ti->meshId	= meshId;
		const float * a = vertsOut + 3 * ti->v0, * b = vertsOut + 3 * ti->v1, * c = vertsOut + 3 * ti->v2;
		for(int k = 0; k < 3; ++k) { /* Bounds.cpp:12-15 */
			ti->bmin[k] = glm_min(glm_min(a[k], b[k]), c[k]);
			ti->bmax[k] = glm_max(glm_max(a[k], b[k]), c[k]);
		}
	}
	return ntri;
}
