// plan.hpp -- host-side traversal plan for the tile-resident depth-first sweeps.
//
// The reference walks the tree node by node over whole 4xL matrices (eigen/eigen.j2:122-157,
// driven by `postorder` / `child_parent`).  Site patterns are independent, so here every CTA
// walks the WHOLE tree for its own tile of patterns, depth first.  The most recent result (the
// "top of stack", TOS) stays in registers; partials that are still waiting for their sibling sit
// in a shared-memory stack.  Visiting the child with the larger stack need first (post-order) /
// smaller need first (pre-order) bounds the number of live vectors by the tree's Strahler number
// <= log2(S)+1, so the shared-memory stack needs at most that minus one slots.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace phylo {

enum : int32_t { kSrcTip = -1, kSrcTos = -2 };  // operand sources other than a shared-memory slot (>= 0)

// One internal node of the post-order sweep.  Children are node ids (0-based; < S means tip).
// The result always goes to the TOS registers and to scratch row = step index.
struct alignas(16) PostStep {
    int32_t a, b;          // children, a is visited first
    int32_t src_a, src_b;  // kSrcTip | kSrcTos | shared-memory slot
    int32_t spill;         // slot that receives the previous TOS before it is overwritten, or -1
    int32_t node;          // this node's id
    int32_t pad0, pad1;
};

// One internal node of the pre-order sweep: reads q(node), keeps q(a) in the TOS registers when a is
// internal (a is always processed next), pushes q(b) to shared memory when b is internal, and emits
// the 4x4 gradient statistics of both child branches.
struct alignas(16) PreStep {
    int32_t node, a, b;    // node and children ids; a descends first
    int32_t src_n;         // kSrcTos | shared-memory slot holding q(node)
    int32_t dst_b;         // shared-memory slot receiving q(b), or -1 (b is a tip)
    int32_t a_internal;    // 1 when q(a) must be produced (into TOS)
    int32_t rown;          // scratch row of this node (rescale exponent written by the post-order); -1: not visited by it
    int32_t rowa, rowb;    // scratch rows of the children's partials; -1 for tips (and for leafified nodes, see build_plan)
    int32_t parkn, parkb;  // a scratch row of this node / of b that may hold a parked q (every internal node has one)
    int32_t pad2;
};

struct Plan {
    int S = 0, nnode = 0, root = 0;
    std::vector<PostStep> post;  // S-1 steps
    std::vector<PreStep> pre;    // S-1 steps
    int depth_post = 0, depth_pre = 0;  // shared-memory slots needed (TOS excluded)
    int depth() const { return depth_post > depth_pre ? depth_post : depth_pre; }
};

// peel: [S-1][3] 1-based (child1, child2, parent) in post-order, root = 2S-1
// (phylostan/utils.py:59-81).  Returns false and sets err on malformed input.
// leaf (optional, [2S-1]): internal nodes the POST-order treats as leaves -- their message to the parent comes from a
// table (phylo_b200.cu: message tables), so neither they nor anything below them gets a post-order step; a leafified
// child has source kSrcTip.  The pre-order still visits every internal node; nodes without a post-order step have
// rown = -1 and a parking row behind the post-order's rows.
bool build_plan(int S, const int32_t* peel, Plan& plan, std::string& err, const std::vector<char>* leaf = nullptr);

}  // namespace phylo
