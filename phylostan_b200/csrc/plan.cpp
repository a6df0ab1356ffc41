// plan.cpp -- see plan.hpp.  Pure host C++ (no CUDA); exercised by the CPU test-suite through
// phylo_b200_plan().
#include "plan.hpp"

#include <algorithm>

namespace phylo {

bool build_plan(int S, const int32_t* peel, Plan& plan, std::string& err, const std::vector<char>* leaf) {
    if (S < 2) { err = "need at least 2 tips"; return false; }
    if (!peel) { err = "peel is NULL"; return false; }
    const int nn = 2 * S - 1;
    plan = Plan();
    plan.S = S;
    plan.nnode = nn;
    std::vector<int> left(nn, -1), right(nn, -1), parent(nn, -1);
    std::vector<char> defined(nn, 0);
    for (int k = 0; k < S; ++k) defined[k] = 1;
    for (int i = 0; i < S - 1; ++i) {
        int a = peel[3 * i] - 1, b = peel[3 * i + 1] - 1, p = peel[3 * i + 2] - 1;
        if (a < 0 || a >= nn || b < 0 || b >= nn || p < S || p >= nn || a == b) {
            err = "peel row " + std::to_string(i) + " has node ids out of range";
            return false;
        }
        if (!defined[a] || !defined[b]) {
            err = "peel is not in post-order (row " + std::to_string(i) + " uses an undefined child)";
            return false;
        }
        if (defined[p] || parent[a] >= 0 || parent[b] >= 0) {
            err = "peel row " + std::to_string(i) + " repeats a node";
            return false;
        }
        defined[p] = 1;
        left[p] = a;
        right[p] = b;
        parent[a] = parent[b] = p;
    }
    plan.root = peel[3 * (S - 2) + 2] - 1;
    if (plan.root != nn - 1) { err = "root must be node 2S-1 (last peel row)"; return false; }

    // stack needs (children precede parents in peel, so one forward pass suffices)
    // post-order view: a leafified node counts as a tip
    auto inner = [&](int n) { return n >= S && !(leaf && (*leaf)[n]); };
    std::vector<int> need_post(nn, 0), need_pre(nn, 0);
    for (int i = 0; i < S - 1; ++i) {
        int p = peel[3 * i + 2] - 1, a = left[p], b = right[p];
        bool ia = a >= S, ib = b >= S;
        if (ia && ib) {
            int hi2 = std::max(need_pre[a], need_pre[b]), lo2 = std::min(need_pre[a], need_pre[b]);
            need_pre[p] = std::max(lo2 + 1, hi2);
        } else if (ia || ib) {
            int c = ia ? a : b;
            need_pre[p] = std::max(1, need_pre[c]);
        } else {
            need_pre[p] = 1;
        }
        ia = inner(a); ib = inner(b);
        if (ia && ib) {
            int hi = std::max(need_post[a], need_post[b]), lo = std::min(need_post[a], need_post[b]);
            need_post[p] = std::max(hi, lo + 1);
        } else if (ia || ib) {
            int c = ia ? a : b;
            need_post[p] = std::max(1, need_post[c]);
        } else {
            need_post[p] = 1;
        }
    }

    // ---- post-order: iterative DFS, larger stack need first; TOS in registers ---------------
    std::vector<int> row_of(nn, -1);
    {
        struct Frame { int node, stage; int first, second; };
        std::vector<Frame> st;
        auto order = [&](int n, int& first, int& second) {
            int a = left[n], b = right[n];
            bool ia = inner(a), ib = inner(b);
            if (ia && ib) {
                if (need_post[b] > need_post[a]) std::swap(a, b);
            } else if (ib) {
                std::swap(a, b);  // the internal child is visited first
            }
            first = a;
            second = b;
        };
        int sp = 0, maxsp = 0;      // shared-memory entries
        bool tos_live = false;      // a finished partial sits in the TOS registers
        Frame f0{plan.root, 0, 0, 0};
        order(plan.root, f0.first, f0.second);
        st.push_back(f0);
        while (!st.empty()) {
            Frame& f = st.back();
            if (f.stage < 2) {
                int c = f.stage == 0 ? f.first : f.second;
                ++f.stage;
                if (inner(c)) {
                    Frame g{c, 0, 0, 0};
                    order(c, g.first, g.second);
                    st.push_back(g);
                }
                continue;
            }
            PostStep ps{};
            ps.a = f.first;
            ps.b = f.second;
            ps.node = f.node;
            ps.spill = -1;
            const bool ia = inner(ps.a), ib = inner(ps.b);
            if (ia && ib) {  // b was computed last (TOS); a waits on top of the shared-memory stack
                ps.src_a = sp - 1;
                ps.src_b = kSrcTos;
                sp -= 1;
            } else if (ia) {  // the only internal child was computed by the previous step
                ps.src_a = kSrcTos;
                ps.src_b = kSrcTip;
            } else {  // cherry: whatever is in the TOS still waits for its sibling
                ps.src_a = ps.src_b = kSrcTip;
                if (tos_live) {
                    ps.spill = sp;
                    sp += 1;
                }
            }
            tos_live = true;
            maxsp = std::max(maxsp, sp);
            row_of[f.node] = (int)plan.post.size();
            plan.post.push_back(ps);
            st.pop_back();
        }
        plan.depth_post = maxsp;
    }

    // every internal node's parking row: its post-order row, or one behind them when the post-order skips it
    std::vector<int> park_of(nn, -1);
    {
        int next = (int)plan.post.size();
        for (int n = S; n < nn; ++n) park_of[n] = row_of[n] >= 0 ? row_of[n] : next++;
    }
    // ---- pre-order: DFS from the root, smaller stack need first; q(first child) stays in TOS ----
    {
        std::vector<int> pending;  // nodes whose q sits in shared memory, slot = position
        int maxsp = 0;
        int cur = plan.root, src = kSrcTos;
        while (true) {
            int a = left[cur], b = right[cur];
            bool ia = a >= S, ib = b >= S;
            if (ia && ib) {
                if (need_pre[b] < need_pre[a]) std::swap(a, b);  // a descends first
            } else if (ib) {
                std::swap(a, b);
                std::swap(ia, ib);
            }
            PreStep ps{};
            ps.node = cur; ps.a = a; ps.b = b; ps.src_n = src;
            ps.a_internal = ia ? 1 : 0;
            ps.rown = row_of[cur];
            ps.rowa = ia ? row_of[a] : -1;
            ps.rowb = ib ? row_of[b] : -1;
            ps.parkn = park_of[cur];
            ps.parkb = ib ? park_of[b] : -1;
            ps.dst_b = -1;
            if (ib) {
                ps.dst_b = (int)pending.size();
                pending.push_back(b);
                maxsp = std::max(maxsp, (int)pending.size());
            }
            plan.pre.push_back(ps);
            if (ia) {
                cur = a;
                src = kSrcTos;
            } else if (!pending.empty()) {
                cur = pending.back();
                pending.pop_back();
                src = (int)pending.size();
            } else {
                break;
            }
        }
        plan.depth_pre = maxsp;
    }
    if ((!leaf && (int)plan.post.size() != S - 1) || (int)plan.pre.size() != S - 1 || plan.post.empty()) {
        err = "peel does not describe a single binary tree";
        return false;
    }
    return true;
}

}  // namespace phylo
