// rb_obj.cpp — Wavefront OBJ ingestion for large scenes, host side of the library.
//
// Reproduces what the reference's `Mesh(std::istream &, Mesh::Load, name)` (src/engine/resources/Mesh.cpp:100-266) makes
// of an OBJ text — same vertices in the same order, same indices, same floats — so that a mesh loaded here and handed
// to rb_raycaster_add_mesh is the mesh the reference's Resources::getMesh (ResourcesManager.cpp:306-340) would have
// given its Raycaster. What the reference does line by line with iostreams on one thread, this does in three passes
// over the text in memory: (1) chunks of lines are tokenised and their numbers converted on all host threads
// (strtof, as std::stof does underneath: identical values), (2) the per-chunk arrays are concatenated in file order,
// (3) faces are resolved — for Load::Indexed with one hash map keyed by the face token's exact text, as the reference
// keys its unordered_map, so vertices are created in the same first-use order.
// Also here: Mesh::computeNormals (Mesh.cpp:326-352) and Mesh::computeBoundingBox (:306-324) in glm's arithmetic.
// Not here: tangent frames — the reference gets them from MikkTSpace (src/libs/mikktspace, third party), which a caller
// that needs normal mapping keeps using; the path tracer's scene tables take tangents as input (rb_scene_add_object).
#include "rb_internal.h"
#include "rb_math.cuh"

#include <cerrno>
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <string>

struct rb_mesh {
	std::vector<float> positions, normals, texcoords; // xyz, xyz, uv
	std::vector<uint32_t> indices;
};

namespace {

struct Token {
	const char * p;
	uint32_t n;
};
struct TokenHash {
	size_t operator()(const Token & t) const {
		uint64_t h = 1469598103934665603ull; // FNV-1a
		for(uint32_t i = 0; i < t.n; ++i) h = (h ^ (unsigned char)t.p[i]) * 1099511628211ull;
		return size_t(h);
	}
};
struct TokenEq {
	bool operator()(const Token & a, const Token & b) const { return a.n == b.n && std::memcmp(a.p, b.p, a.n) == 0; }
};

struct Chunk {
	std::vector<float> v, vn, vt;
	std::vector<Token> faces; // three per face line
	std::string error;
};

inline bool isSpace(char c) { return c == ' ' || c == '\t' || c == '\n' || c == '\v' || c == '\f' || c == '\r'; }

// std::stof on a token: the longest valid prefix, an error where stof would throw.
inline bool parseFloat(const Token & t, float & out) {
	char * end = nullptr;
	errno	   = 0;
	out		   = std::strtof(t.p, &end); // the token is followed by white space or the terminating NUL: strtof stops there
	return end != t.p && end <= t.p + t.n && errno != ERANGE;
}

// std::stoi / stol on the text starting at p (at most n characters): leading integer, an error where they would throw.
inline bool parseIndex(const char * p, size_t n, long & out) {
	if(n == 0) return false;
	char * end = nullptr;
	errno	   = 0;
	out		   = std::strtol(p, &end, 10); // stops at '/', white space or NUL
	return end != p && errno != ERANGE;
}

void parseLines(const char * begin, const char * end, Chunk & c) {
	Token tok[5];
	const char * line = begin;
	while(line < end) {
		const char * eol = static_cast<const char *>(std::memchr(line, '\n', size_t(end - line)));
		if(!eol) eol = end;
		const char * s = line;
		const char * e = eol;
		line		   = eol + 1;
		// Mesh.cpp:118: empty lines, comments and one-character lines are skipped before anything is trimmed.
		if(e - s < 2 || *s == '#') continue;
		// First tokens of the line (white-space separated, as operator>> splits them); a face uses four, a vertex four.
		int nt = 0;
		const char * p = s;
		while(p < e && nt < 5) {
			while(p < e && isSpace(*p)) ++p;
			if(p >= e) break;
			const char * q = p;
			while(q < e && !isSpace(*q)) ++q;
			tok[nt].p = p;
			tok[nt].n = uint32_t(q - p);
			++nt;
			p = q;
		}
		if(nt == 0) continue;
		const Token & k = tok[0];
		const bool isV	= k.n == 1 && k.p[0] == 'v';
		const bool isVn = k.n == 2 && k.p[0] == 'v' && k.p[1] == 'n';
		const bool isVt = k.n == 2 && k.p[0] == 'v' && k.p[1] == 't';
		const bool isF	= k.n == 1 && k.p[0] == 'f';
		if(isV || isVn) {
			if(nt < 4) continue;
			float x, y, z;
			if(!parseFloat(tok[1], x) || !parseFloat(tok[2], y) || !parseFloat(tok[3], z)) {
				c.error = "a coordinate that is not a number";
				return;
			}
			std::vector<float> & dst = isV ? c.v : c.vn;
			dst.push_back(x);
			dst.push_back(y);
			dst.push_back(z);
		} else if(isVt) {
			if(nt < 3) continue;
			float u, v;
			if(!parseFloat(tok[1], u) || !parseFloat(tok[2], v)) {
				c.error = "a texture coordinate that is not a number";
				return;
			}
			c.vt.push_back(u);
			c.vt.push_back(1.0f - v); // Mesh.cpp:154
		} else if(isF) {
			if(nt < 4) continue;
			c.faces.push_back(tok[1]); // the first three corners only, as the reference (Mesh.cpp:158-165)
			c.faces.push_back(tok[2]);
			c.faces.push_back(tok[3]);
		}
	}
}

template<typename T>
void append(std::vector<T> & dst, const std::vector<T> & src) {
	dst.insert(dst.end(), src.begin(), src.end());
}

// The three 1-based indices of a face corner as Mesh.cpp:197-214 / :231-247 read them: position = leading integer;
// texcoord = integer after the first '/', normal = integer after the last '/' (the whole token when it has no '/').
bool cornerIndices(const Token & t, bool hasUV, bool hasNormals, size_t np, size_t nuv, size_t nn, long & ip, long & iuv, long & in, std::string & error) {
	const char * s	   = t.p;
	const char * first = static_cast<const char *>(std::memchr(s, '/', t.n));
	const char * last  = first;
	if(first) {
		for(const char * q = s + t.n - 1; q > first; --q) {
			if(*q == '/') {
				last = q;
				break;
			}
		}
	}
	if(!parseIndex(s, t.n, ip)) {
		error = "a face corner without a position index";
		return false;
	}
	ip -= 1;
	iuv = in = 0;
	if(hasUV) {
		const char * u = first ? first + 1 : s;
		if(!parseIndex(u, size_t(s + t.n - u), iuv)) {
			error = "a face corner without a texture-coordinate index in a mesh that has texture coordinates";
			return false;
		}
		iuv -= 1;
	}
	if(hasNormals) {
		const char * n = last ? last + 1 : s;
		if(!parseIndex(n, size_t(s + t.n - n), in)) {
			error = "a face corner without a normal index in a mesh that has normals";
			return false;
		}
		in -= 1;
	}
	if(ip < 0 || size_t(ip) >= np || (hasUV && (iuv < 0 || size_t(iuv) >= nuv)) || (hasNormals && (in < 0 || size_t(in) >= nn))) {
		error = "a face index out of range (relative indices are not supported, as in the reference)";
		return false;
	}
	return true;
}

} // namespace

extern "C" rb_status rb_obj_parse(const char * text, size_t nbytes, rb_obj_mode mode, rb_mesh ** out) {
	if(!out || (!text && nbytes)) return rb_fail(RB_ERR_INVALID, "rb_obj_parse: null argument");
	if(mode != RB_OBJ_EXPANDED && mode != RB_OBJ_POINTS && mode != RB_OBJ_INDEXED) return rb_fail(RB_ERR_INVALID, "rb_obj_parse: unknown mode %d", int(mode));
	const bool timing = std::getenv("RB_OBJ_TIMING") != nullptr; // developer aid: phases on stderr
	auto t0			  = std::chrono::steady_clock::now();
	auto lap		  = [&](const char * what) {
		 if(!timing) return;
		 const auto t1 = std::chrono::steady_clock::now();
		 fprintf(stderr, "[rb_obj_parse] %-28s %7.1f ms\n", what, std::chrono::duration<double, std::milli>(t1 - t0).count());
		 t0 = t1;
	};
	// One private, NUL-terminated copy: strtof / strtol need a terminator after the last token.
	// (A text that ends in white space needs none: every token is then followed by a character that stops the parse.)
	std::string buffer;
	const char * b = text;
	if(nbytes == 0 || !isSpace(text[nbytes - 1])) {
		buffer.assign(text ? text : "", nbytes);
		b = buffer.c_str();
	}
	const char * e = b + nbytes;
	lap("copy");
	// Pass 1: chunks of whole lines on all host threads.
	const size_t hw		 = std::thread::hardware_concurrency();
	size_t nchunks		 = nbytes / (size_t(1) << 20) + 1;
	if(nchunks > (hw ? hw : 1) * 4) nchunks = (hw ? hw : 1) * 4;
	std::vector<const char *> cut(nchunks + 1, e);
	cut[0] = b;
	for(size_t i = 1; i < nchunks; ++i) {
		const char * p = b + nbytes / nchunks * i;
		if(p < cut[i - 1]) p = cut[i - 1];
		const char * nl = static_cast<const char *>(std::memchr(p, '\n', size_t(e - p)));
		cut[i]			= nl ? nl + 1 : e;
	}
	std::vector<Chunk> chunks(nchunks);
	rb_parallel_for(nchunks, 1, [&](size_t begin, size_t end) {
		for(size_t i = begin; i < end; ++i) parseLines(cut[i], cut[i + 1], chunks[i]);
	});
	for(const Chunk & c : chunks) {
		if(!c.error.empty()) return rb_fail(RB_ERR_INVALID, "rb_obj_parse: %s", c.error.c_str());
	}
	lap("tokenise + convert");
	// Pass 2: file order.
	std::vector<float> v, vn, vt;
	std::vector<Token> faces;
	size_t nv = 0, nn = 0, nt = 0, nf = 0;
	for(const Chunk & c : chunks) { nv += c.v.size(); nn += c.vn.size(); nt += c.vt.size(); nf += c.faces.size(); }
	v.reserve(nv); vn.reserve(nn); vt.reserve(nt); faces.reserve(nf);
	for(const Chunk & c : chunks) { append(v, c.v); append(vn, c.vn); append(vt, c.vt); append(faces, c.faces); }
	chunks.clear();
	lap("concatenate");
	auto mesh = std::make_unique<rb_mesh>();
	if(v.empty()) { // Mesh.cpp:170: no vertices, no mesh
		*out = mesh.release();
		return RB_OK;
	}
	const bool hasUV = !vt.empty(), hasNormals = !vn.empty();
	const size_t np = v.size() / 3, nuv = vt.size() / 2, nnor = vn.size() / 3;
	std::string error;
	if(mode == RB_OBJ_POINTS) {
		mesh->positions = std::move(v);
		mesh->normals	= std::move(vn);
		mesh->texcoords = std::move(vt);
	} else {
		// Pass 3: faces. (a) every corner's three indices, on all threads; (b) Expanded: one new vertex per corner —
		// Indexed: one per distinct corner TEXT, in first-use order (sequential: the order is the result);
		// (c) the new vertices' attributes gathered on all threads.
		struct Corner {
			uint32_t p, uv, n;
		};
		std::vector<Corner> corner(faces.size());
		std::vector<uint32_t> hash(mode == RB_OBJ_INDEXED ? faces.size() : 0);
		std::vector<std::string> errors(faces.size() / (size_t(1) << 16) + 1);
		rb_parallel_for(faces.size(), size_t(1) << 16, [&](size_t begin, size_t end) {
			for(size_t i = begin; i < end; ++i) {
				long ip, iuv, in;
				std::string why;
				if(!cornerIndices(faces[i], hasUV, hasNormals, np, nuv, nnor, ip, iuv, in, why)) {
					errors[begin >> 16] = why; // ranges are at least one grain long: no two of them start in the same 64 K block
					return;
				}
				corner[i] = Corner{uint32_t(ip), uint32_t(iuv), uint32_t(in)};
				if(mode == RB_OBJ_INDEXED) hash[i] = uint32_t(TokenHash()(faces[i]) >> 17);
			}
		});
		lap("corner indices");
		std::vector<uint32_t> firstUse; // per new vertex: the face corner that created it
		mesh->indices.resize(faces.size());
		if(mode == RB_OBJ_EXPANDED) {
			firstUse.resize(faces.size());
			for(size_t i = 0; i < faces.size(); ++i) firstUse[i] = mesh->indices[i] = uint32_t(i);
		} else {
			// Open-addressing table of vertex ids keyed by the corner's text (its hash was computed above, in parallel).
			size_t cap = 1;
			while(cap < 2 * faces.size() + 16) cap <<= 1;
			std::vector<uint32_t> table(cap, 0u); // vertex id + 1, 0 = free
			const TokenEq same;
			for(size_t i = 0; i < faces.size(); ++i) {
				size_t slot = hash[i] & (cap - 1);
				for(;;) {
					const uint32_t entry = table[slot];
					if(entry == 0u) {
						// A corner that fails to parse only matters where the reference would parse it: at its first use.
						const uint32_t id = uint32_t(firstUse.size());
						table[slot]		  = id + 1u;
						firstUse.push_back(uint32_t(i));
						mesh->indices[i] = id;
						break;
					}
					if(hash[firstUse[entry - 1u]] == hash[i] && same(faces[firstUse[entry - 1u]], faces[i])) {
						mesh->indices[i] = entry - 1u;
						break;
					}
					slot = (slot + 1) & (cap - 1);
				}
			}
		}
		lap("first-use order (hash map)");
		for(const std::string & why : errors) {
			if(!why.empty()) return rb_fail(RB_ERR_INVALID, "rb_obj_parse: %s", why.c_str());
		}
		const size_t nverts = firstUse.size();
		mesh->positions.resize(3 * nverts);
		if(hasUV) mesh->texcoords.resize(2 * nverts);
		if(hasNormals) mesh->normals.resize(3 * nverts);
		rb_parallel_for(nverts, size_t(1) << 16, [&](size_t begin, size_t end) {
			for(size_t k = begin; k < end; ++k) {
				const Corner & c = corner[firstUse[k]];
				std::memcpy(&mesh->positions[3 * k], &v[3 * size_t(c.p)], 12);
				if(hasUV) std::memcpy(&mesh->texcoords[2 * k], &vt[2 * size_t(c.uv)], 8);
				if(hasNormals) std::memcpy(&mesh->normals[3 * k], &vn[3 * size_t(c.n)], 12);
			}
		});
	}
	lap("gather");
	*out = mesh.release();
	return RB_OK;
}

extern "C" void rb_mesh_destroy(rb_mesh * mesh) {
	delete mesh;
}

extern "C" rb_status rb_mesh_sizes(const rb_mesh * mesh, uint64_t out[4]) {
	if(!mesh || !out) return rb_fail(RB_ERR_INVALID, "rb_mesh_sizes: null argument");
	out[0] = mesh->positions.size() / 3;
	out[1] = mesh->normals.size() / 3;
	out[2] = mesh->texcoords.size() / 2;
	out[3] = mesh->indices.size();
	return RB_OK;
}

extern "C" rb_status rb_mesh_read(const rb_mesh * mesh, float * positions, float * normals, float * texcoords, uint32_t * indices, float bbox[6]) {
	if(!mesh) return rb_fail(RB_ERR_INVALID, "rb_mesh_read: null mesh");
	if(positions) std::memcpy(positions, mesh->positions.data(), mesh->positions.size() * sizeof(float));
	if(normals) std::memcpy(normals, mesh->normals.data(), mesh->normals.size() * sizeof(float));
	if(texcoords) std::memcpy(texcoords, mesh->texcoords.data(), mesh->texcoords.size() * sizeof(float));
	if(indices) std::memcpy(indices, mesh->indices.data(), mesh->indices.size() * sizeof(uint32_t));
	if(bbox) {
		// Mesh::computeBoundingBox (Mesh.cpp:306-324): glm::min / glm::max component by component; an empty mesh keeps the
		// default BoundingBox, which is inverted (min = FLT_MAX, max = -FLT_MAX; Bounds.hpp).
		for(int k = 0; k < 3; ++k) {
			bbox[k]		= FLT_MAX;
			bbox[3 + k] = -FLT_MAX;
		}
		const size_t n = mesh->positions.size() / 3;
		if(n) {
			const float * p = mesh->positions.data();
			for(int a = 0; a < 3; ++a) bbox[a] = bbox[3 + a] = p[a];
			for(size_t i = 1; i < n; ++i) {
				for(int a = 0; a < 3; ++a) {
					bbox[a]		= glm_min(bbox[a], p[3 * i + a]);
					bbox[3 + a] = glm_max(bbox[3 + a], p[3 * i + a]);
				}
			}
		}
	}
	return RB_OK;
}

// Mesh::computeNormals (Mesh.cpp:326-352): per face cross(normalize(v1 - v0), normalize(v2 - v0)) summed into its three
// vertices in face order, then normalised — sequential, because fp32 sums depend on their order.
extern "C" rb_status rb_mesh_compute_normals(rb_mesh * mesh) {
	if(!mesh) return rb_fail(RB_ERR_INVALID, "rb_mesh_compute_normals: null mesh");
	const size_t n = mesh->positions.size() / 3;
	mesh->normals.assign(3 * n, 0.0f);
	const float * P = mesh->positions.data();
	float * N		= mesh->normals.data();
	auto at			= [&](uint32_t i) { return V3{P[3 * size_t(i)], P[3 * size_t(i) + 1], P[3 * size_t(i) + 2]}; };
	for(size_t t = 0; t + 2 < mesh->indices.size(); t += 3) {
		const uint32_t i0 = mesh->indices[t], i1 = mesh->indices[t + 1], i2 = mesh->indices[t + 2];
		if(i0 >= n || i1 >= n || i2 >= n) return rb_fail(RB_ERR_INVALID, "rb_mesh_compute_normals: index out of range");
		const V3 v0 = at(i0);
		const V3 nrm = cross3(normalize3(at(i1) - v0), normalize3(at(i2) - v0));
		const uint32_t ids[3] = {i0, i1, i2};
		for(uint32_t i : ids) {
			N[3 * size_t(i)] += nrm.x;
			N[3 * size_t(i) + 1] += nrm.y;
			N[3 * size_t(i) + 2] += nrm.z;
		}
	}
	for(size_t i = 0; i < n; ++i) {
		const V3 r = normalize3(V3{N[3 * i], N[3 * i + 1], N[3 * i + 2]});
		N[3 * i] = r.x; N[3 * i + 1] = r.y; N[3 * i + 2] = r.z;
	}
	return RB_OK;
}
