// plan.hpp -- host-side traversal plan for the tile-resident depth-first sweeps.
//
// The reference walks the tree node by node over whole 4xL matrices (eigen/eigen.j2:122-157,
// driven by `postorder` / `child_parent`).  Site patterns are independent, so here every CTA
// walks the WHOLE tree for its own tile of patterns, depth first, keeping the partials that are
// still waiting for their sibling in a shared-memory stack.  Visiting the child with the larger
// stack need first (post-order) / smaller need first (pre-order) bounds the stack depth by the
// tree's Strahler number <= log2(S)+1, so the stack always fits in the 227 KB of an SM.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace phylo {

// One internal node of the post-order sweep.  Children are node ids (0-based; < S means tip).
// Slots index the shared-memory stack; -1 for a tip child (read from the tip codes instead).
// The partial of the node computed at step i is also written to scratch row i.
struct alignas(16) PostStep {
    int32_t a, b;        // children, a is visited first
    int32_t sa, sb;      // stack slots holding the children's partials
    int32_t so;          // stack slot receiving this node's partial
    int32_t node;        // this node's id
    int32_t pad0, pad1;
};

// One internal node of the pre-order sweep: reads q(node), emits q for internal children and
// the 4x4 gradient statistics of both child branches.
struct alignas(16) PreStep {
    int32_t node, a, b;  // node and children ids
    int32_t sn;          // stack slot holding q(node)
    int32_t sa, sb;      // stack slots receiving q(a), q(b); -1 for tips
    int32_t rown;        // scratch row of this node (rescale exponent written by the post-order)
    int32_t rowa, rowb;  // scratch rows of the children's partials; -1 for tips
    int32_t pad0, pad1, pad2;
};

struct Plan {
    int S = 0, nnode = 0, root = 0;
    std::vector<PostStep> post;  // S-1 steps
    std::vector<PreStep> pre;    // S-1 steps
    int depth_post = 0, depth_pre = 0;
    int depth() const { return depth_post > depth_pre ? depth_post : depth_pre; }
};

// peel: [S-1][3] 1-based (child1, child2, parent) in post-order, root = 2S-1
// (phylostan/utils.py:59-81).  Returns false and sets err on malformed input.
bool build_plan(int S, const int32_t* peel, Plan& plan, std::string& err);

}  // namespace phylo
