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
	// Scenes with 9..64 meshes: a small binary tree over the root boxes in the interior-node layout (child boxes and
	// references in the parent; reference < 0: root ~reference). rb_roots_tree_pretest walks it once per ray and returns
	// the bit mask of the roots whose box the ray enters. nullptr: not built.
	const float4 * rootTree;
	// 65..512 meshes: the mask does not fit two registers; it lives in the first `rootWords` entries of the ray's stack
	// column (the traversal stack proper starts behind them). 0: register mask.
	int rootWords;
};
#define RB_ROOT_TREE_MIN 9
#define RB_ROOT_TREE_MAX 64		// register mask up to here ...
#define RB_ROOT_TREE_BIG_MAX 512 // ... shared-memory mask up to here (16 stack entries per ray)

#define RB_REF_NONE 0x7fffffff
#define RB_LEAF_BITS 3
#define RB_LEAF_MAX_TRIS (1 << RB_LEAF_BITS)

