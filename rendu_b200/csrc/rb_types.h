// rb_types.h — plain structs shared by host-only and device translation units.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define RB_MAX_ROOT_LEVELS 24

struct RbDeviceBvh {
	const float4 * nodes;
	const float4 * tris;
	const float4 * roots;
	int nroots;
	int stackDepth; // entries a traversal stack needs (max interior depth + 1)
	// Scenes with many meshes: union boxes of aligned blocks of 2^k roots, k = 1..rootLevels (rb_next_root). Level k
	// starts at block rootLevelOffset[k]; two float4 per block, laid out like a root. rootLevels = 0: not built.
	const float4 * rootBlocks;
	int rootLevels;
	int rootLevelOffset[RB_MAX_ROOT_LEVELS + 1];
};

#define RB_REF_NONE 0x7fffffff
#define RB_LEAF_BITS 3
#define RB_LEAF_MAX_TRIS (1 << RB_LEAF_BITS)

