// plan.cpp -- see plan.hpp.  Pure host C++ (no CUDA); exercised by the CPU test-suite through
// phylo_b200_plan().
#include "plan.hpp"

#include <algorithm>
#include <functional>
#include <queue>

namespace phylo {

bool build_plan(int S, const int32_t* peel, Plan& plan, std::string& err) {
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
    std::vector<int> need_post(nn, 0), need_pre(nn, 0);
    for (int i = 0; i < S - 1; ++i) {
        int p = peel[3 * i + 2] - 1, a = left[p], b = right[p];
        bool ia = a >= S, ib = b >= S;
        if (ia && ib) {
            int hi = std::max(need_post[a], need_post[b]), lo = std::min(need_post[a], need_post[b]);
            need_post[p] = std::max(hi, lo + 1);
            int hi2 = std::max(need_pre[a], need_pre[b]), lo2 = std::min(need_pre[a], need_pre[b]);
            need_pre[p] = std::max(lo2 + 1, hi2);
        } else if (ia || ib) {
            int c = ia ? a : b;
            need_post[p] = std::max(1, need_post[c]);
            need_pre[p] = std::max(1, need_pre[c]);
        } else {
            need_post[p] = need_pre[p] = 1;
        }
    }

    // ---- post-order: iterative DFS, larger stack need first -------------------------------
    std::vector<int> row_of(nn, -1);
    {
        struct Frame { int node, stage; int first, second; };
        std::vector<Frame> st;
        auto order = [&](int n, int& first, int& second) {
            int a = left[n], b = right[n];
            bool ia = a >= S, ib = b >= S;
            if (ia && ib) {
                if (need_post[b] > need_post[a]) std::swap(a, b);
            } else if (ib) {
                std::swap(a, b);  // the internal child is visited first
            }
            first = a;
            second = b;
        };
        int sp = 0, maxsp = 0;
        Frame f0{plan.root, 0, 0, 0};
        order(plan.root, f0.first, f0.second);
        st.push_back(f0);
        while (!st.empty()) {
            Frame& f = st.back();
            if (f.stage < 2) {
                int c = f.stage == 0 ? f.first : f.second;
                ++f.stage;
                if (c >= S) {
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
            bool ia = ps.a >= S, ib = ps.b >= S;
            if (ia && ib) {
                ps.sa = sp - 2; ps.sb = sp - 1; ps.so = sp - 2; sp -= 1;
            } else if (ia) {
                ps.sa = sp - 1; ps.sb = -1; ps.so = sp - 1;
            } else {
                ps.sa = ps.sb = -1; ps.so = sp; sp += 1;
            }
            maxsp = std::max(maxsp, sp);
            row_of[f.node] = (int)plan.post.size();
            plan.post.push_back(ps);
            st.pop_back();
        }
        plan.depth_post = maxsp;
    }

    // ---- pre-order: DFS from the root, smaller stack need first, lowest free slot ---------
    {
        std::priority_queue<int, std::vector<int>, std::greater<int>> free_slots;
        int next_slot = 0, maxslot = 0;
        auto alloc = [&]() {
            if (!free_slots.empty()) { int s = free_slots.top(); free_slots.pop(); return s; }
            return next_slot++;
        };
        std::vector<std::pair<int, int>> st;  // (node, slot)
        st.push_back({plan.root, alloc()});
        maxslot = next_slot;
        while (!st.empty()) {
            auto [n, sn] = st.back();
            st.pop_back();
            int a = left[n], b = right[n];
            bool ia = a >= S, ib = b >= S;
            if (ia && ib && need_pre[b] < need_pre[a]) std::swap(a, b);  // a descends first
            else if (!ia && ib) { std::swap(a, b); std::swap(ia, ib); }
            ia = a >= S; ib = b >= S;
            PreStep ps{};
            ps.node = n; ps.a = a; ps.b = b; ps.sn = sn;
            ps.rown = row_of[n];
            ps.rowa = ia ? row_of[a] : -1;
            ps.rowb = ib ? row_of[b] : -1;
            free_slots.push(sn);  // q(node) is in registers before the children are written
            ps.sa = ia ? alloc() : -1;
            ps.sb = ib ? alloc() : -1;
            maxslot = std::max(maxslot, next_slot);
            if (ib) st.push_back({b, ps.sb});
            if (ia) st.push_back({a, ps.sa});
            plan.pre.push_back(ps);
        }
        plan.depth_pre = maxslot;
    }
    if ((int)plan.post.size() != S - 1 || (int)plan.pre.size() != S - 1) {
        err = "peel does not describe a single binary tree";
        return false;
    }
    return true;
}

}  // namespace phylo
