// rb_types.h — plain structs shared by host-only and device translation units.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

struct RbDeviceBvh {
	const float4 * nodes;
	const float4 * tris;
	const float4 * roots;
	int nroots;
	int stackDepth; // entries a traversal stack needs (max interior depth + 1)
};

#define RB_REF_NONE 0x7fffffff
#define RB_LEAF_BITS 3
#define RB_LEAF_MAX_TRIS (1 << RB_LEAF_BITS)

