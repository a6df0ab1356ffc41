// phylo_b200.cu -- host side of libphylo_b200.so: context, parameter packing, launches, C ABI.
// See include/phylo_b200.h for the contract and the reference interfaces each entry replaces.
#include "phylo_b200.h"

#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <map>
#include <new>
#include <string>
#include <vector>

#include "kernels.cuh"
#include "plan.hpp"
#include "subst.hpp"

using namespace phylo;

namespace {

thread_local std::string g_err;

int fail(int code, const std::string& msg) {
    g_err = msg;
    return code;
}

#define CU_TRY(expr)                                                                              \
    do {                                                                                          \
        cudaError_t e_ = (expr);                                                                  \
        if (e_ != cudaSuccess)                                                                    \
            return fail(PHYLO_B200_ECUDA, std::string(#expr) + ": " + cudaGetErrorString(e_));    \
    } while (0)

constexpr int kPadPatterns = 512;  // pattern axis padded so every tile shape divides it
constexpr int kMaxCategories = 16;

template <class T>
struct DevBuf {
    T* p = nullptr;
    size_t n = 0;
    cudaError_t ensure(size_t count) {
        if (count <= n) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr;
        n = 0;
        cudaError_t e = cudaMalloc(&p, count * sizeof(T));
        if (e == cudaSuccess) n = count;
        return e;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        n = 0;
    }
};

template <class T>
struct PinnedBuf {
    T* p = nullptr;
    size_t n = 0;
    cudaError_t ensure(size_t count) {
        if (count <= n) return cudaSuccess;
        if (p) cudaFreeHost(p);
        p = nullptr;
        n = 0;
        cudaError_t e = cudaMallocHost(&p, count * sizeof(T));
        if (e == cudaSuccess) n = count;
        return e;
    }
    void release() {
        if (p) cudaFreeHost(p);
        p = nullptr;
        n = 0;
    }
};

}  // namespace

constexpr double kMsgTauMax = 12.0;  // e^12 * 2^-53 = 1.8e-11 relative: two decades inside the 1e-8 gradient tolerance
constexpr bool kDefaultPostTables = true;
constexpr int kDefaultSweepTm = 0;  // fp64 K = 4 gradient runs: 0 shared-memory stack, 2 / 3 tensor-memory stack

struct phylo_b200_ctx {
    int S = 0, L = 0, C = 0, model = 0, flags = 0, device = 0;
    int nn = 0, bcount = 0, nsubst = 0, nout = 0, Lpad = 0;
    bool rooted = false, normalize = true, jc_closed = false;
    bool tips_simple = false;  // every tip cell one-hot or all ones -> column-index fast path
    int off_subst = 0, off_freqs = 0, off_rs = 0, off_ps = 0;
    Plan plan;
    ParamLayout lay{};
    int num_sms = 0;
    size_t smem_optin = 0;

    // static device data
    DevBuf<uint8_t> d_tips;
    // TR kernels: the tip codes once more per sweep, [block of 32 K patterns][slot][32 K] in the order the sweep
    // consumes them (built on the device the first time a tiling with this K runs)
    DevBuf<uint8_t> d_tips_post, d_tips_pre;
    DevBuf<int32_t> d_tip_order;  // [2][S]: tip (0-based) of every slot, post-order then pre-order
    int tipring_K = 0;            // K the copies above were built for (0: none); the pre-order copy only for gradient runs
    bool tipring_pre = false;
    DevBuf<double> d_weights;
    DevBuf<PostStep> d_post;
    DevBuf<PreStep> d_pre;
    // per-batch device data
    DevBuf<double> d_params, d_G, d_out;
    DevBuf<unsigned char> d_spost, d_spre;  // per-(draw, category) instruction streams
    DevBuf<int32_t> d_node_pos, d_node_row;
    DevBuf<double2> d_scratch;
    DevBuf<uint8_t> d_dscr;
    PinnedBuf<double> h_params, h_out;

    // tiling (user request, 0 = auto) and the resolved launch shape of the last run
    int prec = 64;  // 64: product path; 32: optional fp32-with-scaling mode
    int req_K = 0, req_PB = 0, req_cap = 0;
    int K = 1, PB = 1, NT = 0, grid = 0, ntiles = 0;
    int slots = 0;  // shared-memory stack slots of the last run (< plan.depth(): the rest is parked in HBM)
    bool jc_run = false;  // the last resolved launch is the JC69 scalar-statistic kernel
    int req_tm = kDefaultSweepTm;  // tensor-memory-stack sweep for fp64 K = 4 gradient runs: 0 off, 2 / 3 = resident CTAs per SM
    int tm = 0;           // what the last resolved launch uses (0: shared-memory stack)
    // Message-statistic gradient sweep (kernels.cu, MSG): chosen per run when the handle allows it (fp64, simple tips,
    // 128-thread CTAs) AND the packed batch does: its contraction amplifies rounding by
    // e^{|l_i - l_j| t_b r_c} (times the conditioning of the eigenvector matrices), so tau_bound = max over the batch
    // of (max_b t_b)(max_c r_c)(l_max - l_min) + log(|m1|_F |m2|_F / 4) must stay below kMsgTauMax.  PHYLO_B200_MSG=0 turns it off.
    bool use_msg = true;
    double tau_unit = 0.0;    // max over the batch of (max_c r_c)(l_max - l_min): tau_bound per unit of branch length
    double tau_cond = 0.0;    // max over the batch of log(|m1|_F |m2|_F / 4): the eigenvector matrices' share of the bound
    double tau_bound = 0.0;   // of the batch packed last (front ends that compute branch lengths on the device set
                              // it from their own inputs)
    bool msg_run = false;     // the last resolved launch uses the message statistic
    // Cherry tables (message-statistic runs; kernels.cu K3b): a cherry's message to its parent comes from a 25-entry
    // table per (draw, category, cherry) instead of a scratch row: a third less scratch traffic, +4 % (short runs) to
    // +6 % (sustained, power-capped) on config 3.  On by default; phylo_b200_set_cherry_tables(h, 0) / PHYLO_B200_CHERRY=0
    // turn it off.
    bool use_cherry = true, cherry_run = false;
    int ncherry = 0, tab_entries = 0;            // table nodes (cherries and pitchforks) and their entries per (draw, category)
    DevBuf<int32_t> d_node_cherry, d_cherries;   // [nn] node -> table-node index or -1; [ncherry][kTabRec]
    DevBuf<int32_t> d_tab_off;                   // [ncherry] first entry of every table node
    DevBuf<uint8_t> d_ctips;                     // [ncherry][Lpad] combined codes 5 x + y / 25 x + 5 y + z (built at the first such run)
    bool ctips_built = false;
    DevBuf<double> d_ctab;                       // [B][C][tab_entries][4]
    // Post-order message tables: a second plan whose post-order treats the table nodes as leaves (their message to the
    // parent is the same table entry the pre-order gathers), so the post-order sweep runs half as many steps on a
    // coalescent tree.  Table nodes are then never rescaled.  PHYLO_B200_POST_TABLES=0 turns it off.
    Plan planB;
    bool has_planB = false, use_post_tables = kDefaultPostTables, post_tables_run = false;
    DevBuf<PostStep> d_postB;
    DevBuf<PreStep> d_preB;
    DevBuf<int32_t> d_node_rowB;
    size_t smem = 0;
    int last_launches = 0;

    cudaStream_t own_stream = nullptr, stream = nullptr;
    // eval_batch replays a captured CUDA graph (H2D copy, memsets, the three kernels, D2H copy): for
    // fluA-sized problems the six stream calls cost more CPU time than the GPU needs to run them
    struct EvalGraph { cudaGraphExec_t exec = nullptr; std::vector<unsigned long long> sig; };
    std::map<std::vector<int>, EvalGraph> graphs;  // (B, want_grad, front-end job) -> executable graph + what it baked in
    bool use_graphs = true;
    bool use_jc_scalar = true;  // JC69 gradient runs use the scalar-statistic sweep (PHYLO_B200_NO_JC_SCALAR=1: generic)
    bool timing = false;
    cudaEvent_t ev[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    bool ev_valid = false, ev_has_contract = false;

    // clock-tree front end on the device (eval_heights_batch / eval_ratios_batch): the pre-order map and what is
    // derived from it are uploaded once per map, the per-draw inputs every call
    std::vector<int32_t> c_map;
    std::vector<double> c_lowers;
    bool c_has_lowers = false, c_valid = false;
    DevBuf<int32_t> d_cmap, d_ckids, d_crow;
    DevBuf<double> d_clowers, d_cin, d_hwork, d_hout;
    PinnedBuf<double> h_cin, h_hout, h_hts;

    // multi-device handle (phylo_b200_create_multi): this context is pattern shard 0, `peers` are the shards on
    // the other devices.  Every evaluation forks from this context's stream (fork_ev), runs all shards on their
    // own streams and joins on peer_ev[i]; the peers' result rows are then added to d_out here -- read in place
    // over NVLink (peer_direct[i]) or from d_stage after a peer copy.
    std::vector<phylo_b200_ctx*> peers;
    std::vector<cudaEvent_t> peer_ev;
    std::vector<char> peer_direct;
    cudaEvent_t fork_ev = nullptr;
    DevBuf<double> d_stage;
    bool is_peer = false;  // peers need no pinned staging buffers of their own

    ~phylo_b200_ctx() {
        for (size_t i = 0; i < peers.size(); ++i) {
            if (i < peer_ev.size() && peer_ev[i]) { cudaSetDevice(peers[i]->device); cudaEventDestroy(peer_ev[i]); }
            delete peers[i];
        }
        cudaSetDevice(device);
        if (fork_ev) cudaEventDestroy(fork_ev);
        d_stage.release();
        d_cmap.release(); d_ckids.release(); d_crow.release(); d_clowers.release(); d_cin.release();
        d_hwork.release(); d_hout.release(); h_cin.release(); h_hout.release(); h_hts.release();
        d_tips.release(); d_weights.release(); d_post.release(); d_pre.release();
        d_tips_post.release(); d_tips_pre.release(); d_tip_order.release();
        d_params.release(); d_G.release(); d_out.release();
        d_spost.release(); d_spre.release(); d_node_pos.release(); d_node_row.release();
        d_scratch.release(); d_dscr.release();
        d_node_cherry.release(); d_cherries.release(); d_tab_off.release(); d_ctips.release(); d_ctab.release();
        d_postB.release(); d_preB.release(); d_node_rowB.release();
        h_params.release(); h_out.release();
        for (auto& g : graphs) if (g.second.exec) cudaGraphExecDestroy(g.second.exec);
        for (auto& e : ev) if (e) cudaEventDestroy(e);
        if (own_stream) cudaStreamDestroy(own_stream);
    }
};

namespace {

// Resolve (K, PB) -> launch shape.  Everything (occupancy, slots, grid, scratch) is derived for the kernel
// instantiation that run_enqueue launches, the JC69 scalar-statistic variant included.
int resolve_tiling(phylo_b200_ctx* h, int B, bool grad) {
    const int C = h->C, Dfull = h->plan.depth();
    // JC69 gradient runs use the scalar-statistic sweep when the whole stack fits on chip
    bool jc = grad && h->prec == 64 && h->model == PHYLO_B200_JC69 && h->use_jc_scalar && h->req_cap == 0;
    // Gradient runs may cap the shared-memory stack: the top stack positions (reached rarely, and only
    // briefly) are parked in the CTA's HBM scratch, which holds every partial anyway.  A deep tree
    // (stack depth 7+, thousands of taxa) then still gets the widest tile twice per SM.
    // Value-only runs have no scratch and keep the whole stack in shared memory.
    const int kMaxParked = 2;
    int Dmin = Dfull;
    if (grad) Dmin = h->req_cap > 0 ? std::min(Dfull, h->req_cap) : std::max(std::min(Dfull, 2), Dfull - kMaxParked);
    const int Dmax = grad && h->req_cap > 0 ? Dmin : Dfull;
    // largest slot count in [Dmin, Dmax] reaching `want` CTAs per SM; -1 when none does
    auto slots_for = [&](int k, int nt, int want, bool j) {
        for (int dd = Dmax; dd >= (j ? Dfull : Dmin); --dd) {
            int occ = 0;
            const size_t sm = sweep_smem_bytes(dd, k, nt, h->prec, j, h->tips_simple, grad);
            if (sm <= h->smem_optin &&
                sweep_occupancy(h->prec, h->tips_simple, k, grad, dd < Dfull, nt, sm, &occ, j) == cudaSuccess &&
                occ >= want)
                return dd;
            if (dd == 0) break;
        }
        return -1;
    };
    int K = h->req_K, PB = h->req_PB;
    if (K == 0) {
        // K patterns per lane share the per-step work (record decode, matrix loads, the 4x4 reduction) but
        // lengthen a warp's instruction chain and shrink the number of CTAs.  Estimate the run time of each
        // K as (waves of CTAs) x (time of one tile): measured tile times relative to K = 1 are 1 : 1.1 : 1.35
        // when everything runs in a single wave (latency-bound) and 1 : 1.40 : 1.62 with the SMs saturated
        // (142 / 123 / 86 evaluations/s at K = 4 / 2 / 1 on the 1000 x 100k bench).  K > 1 wants two CTAs
        // per SM; gradient runs may park stack positions to get there.
        const int nt = 32 * C * (PB ? PB : 1);
        K = 1;
        if (nt <= 128) {
            static const double kLat[5] = {0, 1.0, 1.1, 0, 1.35}, kSat[5] = {0, 1.0, 1.40, 0, 1.62};
            double best = 1e300;
            for (int k : {1, 2, 4}) {
                int dd = jc ? slots_for(k, nt, k == 1 ? 1 : 2, true) : -1;
                const bool j = dd >= 0;
                if (!j) dd = slots_for(k, nt, k == 1 ? 1 : 2, false);
                if (dd < 0) continue;
                int occ = 0;
                if (sweep_occupancy(h->prec, h->tips_simple, k, grad, dd < Dfull, nt,
                                    sweep_smem_bytes(dd, k, nt, h->prec, j, h->tips_simple, grad), &occ, j) != cudaSuccess || occ < 1)
                    continue;
                const double items = (double)B * ((h->L + 32 * k - 1) / (32 * k));
                const double w = items / ((double)occ * h->num_sms);
                const double est = w <= 1.0 ? kLat[k] : std::ceil(w) * kSat[k];  // items are dealt out statically
                if (est < best) { best = est; K = k; }
            }
        } else if ((long long)B * ((h->L + 31) / 32) >= 8LL * h->num_sms) {
            K = 2;
        }
    }
    if (K != 1 && K != 2 && K != 4) return fail(PHYLO_B200_EINVAL, "patterns_per_thread must be 1, 2 or 4");
    if (PB == 0) PB = 1;
    if (PB != 1 && PB != 2 && PB != 4) return fail(PHYLO_B200_EINVAL, "pattern_blocks must be 1, 2 or 4");
    while (PB > 1 && 32 * C * PB > sweep_max_threads(K)) PB >>= 1;
    while (K > 1 && 32 * C * PB > sweep_max_threads(K)) K >>= 1;
    int NT = 32 * C * PB;
    if (NT > sweep_max_threads(K)) return fail(PHYLO_B200_EINVAL, "too many rate categories for one CTA");
    int D = -1;
    bool jrun = false;
    for (;;) {
        if (jc) {  // the scalar-statistic kernel keeps the whole stack on chip
            D = slots_for(K, NT, 2, true);
            if (D < 0) D = slots_for(K, NT, 1, true);
            jrun = D >= 0;
        }
        if (D < 0) D = slots_for(K, NT, 2, false);  // two CTAs per SM if any allowed slot count gives that
        if (D < 0) D = slots_for(K, NT, 1, false);
        if (D >= 0) break;
        if (K > 1) K >>= 1;
        else if (PB > 1) { PB >>= 1; NT = 32 * C * PB; }
        else
            return fail(PHYLO_B200_EINVAL,
                        "tree too deep for the shared-memory stack (depth " + std::to_string(Dfull) + ")");
    }
    size_t smem = sweep_smem_bytes(D, K, NT, h->prec, jrun, h->tips_simple, grad);
    const int tpat = PB * 32 * K;
    h->ntiles = (h->L + tpat - 1) / tpat;
    int occ = 0;
    // the stack in tensor memory: three (or two) CTAs per SM, positions beyond its slots parked like above
    int tm = 0;
    if (h->req_tm && sweep_tm_available(h->prec, K, NT, grad, jrun) && h->req_cap == 0) {
        const int dt = std::min(Dfull, sweep_tm_slots(K, h->req_tm));
        if (Dfull - dt <= kMaxParked && sweep_tm_smem_bytes(K) <= h->smem_optin &&
            sweep_tm_prepare(h->tips_simple, K, h->req_tm, &occ) == cudaSuccess && occ >= 1) {
            tm = h->req_tm; D = dt; smem = sweep_tm_smem_bytes(K);
        }
    }
    // (not at K = 1, the latency-bound shapes: there a tip's message lookup sits on the step's dependent chain and the
    // plain statistic is 7 % faster -- fluA 144 against 154 us per gradient, profiles/r2_logs/r2_latency_final.log)
    const bool msg = h->use_msg && K > 1 && sweep_msg_available(h->prec, h->tips_simple, grad, D < Dfull, NT, jrun) &&
                     h->tau_bound < kMsgTauMax;
    if (tm && msg) CU_TRY(sweep_tm_prepare(h->tips_simple, K, tm, &occ, true));
    h->K = K; h->PB = PB; h->NT = NT; h->smem = smem; h->slots = D; h->jc_run = jrun; h->tm = tm; h->msg_run = msg;
    if (!tm) CU_TRY(sweep_occupancy(h->prec, h->tips_simple, K, grad, D < Dfull, NT, smem, &occ, jrun, msg));
    if (occ < 1) return fail(PHYLO_B200_ECUDA, "sweep kernel does not fit on an SM");
    const long long items = (long long)B * h->ntiles;
    h->grid = (int)std::min<long long>(items, (long long)occ * h->num_sms);
    return 0;
}

int ensure_batch(phylo_b200_ctx* h, int B) {
    CU_TRY(h->d_params.ensure((size_t)B * h->lay.stride));
    CU_TRY(h->d_spost.ensure((size_t)B * h->C * (h->S - 1) * kRecBytes));  // sized for the larger (fp64) record
    CU_TRY(h->d_spre.ensure((size_t)B * h->C * (h->S - 1) * kRecBytes));
    CU_TRY(h->d_G.ensure((size_t)B * h->nn * h->C * 16));
    CU_TRY(h->d_out.ensure((size_t)B * h->nout));
    if (!h->is_peer) {
        CU_TRY(h->h_params.ensure((size_t)B * h->lay.stride));
        CU_TRY(h->h_out.ensure((size_t)B * h->nout));
        if (!h->peers.empty()) CU_TRY(h->d_stage.ensure(h->peers.size() * (size_t)B * h->nout));
    }
    return 0;
}

// Table nodes of a plan: the internal nodes with two (cherry) or three (pitchfork: a cherry and a tip) tips below them,
// in post-order.  node_tab[n] = index or -1; rec = [ntab][kTabRec] (kernels.cuh); tab_off[k] = first table entry of
// node k (25 or 125 each); returns the entries per (draw, category).
int find_table_nodes(const Plan& plan, int S, int max_tips, std::vector<int32_t>& node_tab, std::vector<int32_t>& rec,
                     std::vector<int32_t>& tab_off) {
    const int nn = 2 * S - 1;
    node_tab.assign((size_t)nn, -1);
    rec.clear();
    tab_off.clear();
    std::vector<std::vector<int32_t>> below((size_t)nn);  // tips below a node, a's before b's; empty = more than max_tips
    for (int k = 0; k < S; ++k) below[k] = {k};
    int entries = 0;
    for (const PostStep& p : plan.post) {
        const auto &ta = below[p.a], &tb = below[p.b];
        if (ta.empty() || tb.empty() || (int)(ta.size() + tb.size()) > max_tips) continue;
        std::vector<int32_t> t = ta;
        t.insert(t.end(), tb.begin(), tb.end());
        below[p.node] = t;
        const int nt = (int)t.size();
        node_tab[p.node] = (int32_t)tab_off.size();
        tab_off.push_back(entries);
        rec.insert(rec.end(), {p.node, nt, t[0], t[1], nt == 3 ? t[2] : -1, nt == 3 ? (ta.size() == 2 ? p.a : p.b) : -1,
                               nt == 3 && ta.size() == 1 ? 1 : 0, entries});
        entries += nt == 3 ? 125 : 25;
    }
    return entries;
}

// on_device: tipmask / weights are DEVICE pointers on `device` (phylo_b200_create_device)
int create_common(phylo_b200_handle* out, int S, int L, int C, int model, int flags, const int32_t* peel,
                  const uint8_t* tipmask, const double* tipdata, const double* weights, int device,
                  bool on_device = false) {
    if (!out) return fail(PHYLO_B200_EINVAL, "out handle is NULL");
    *out = nullptr;
    if (S < 2 || L < 1 || C < 1) return fail(PHYLO_B200_EINVAL, "need S >= 2, L >= 1, C >= 1");
    if (C > kMaxCategories) return fail(PHYLO_B200_EINVAL, "at most 16 rate categories");
    if (model < PHYLO_B200_JC69 || model > PHYLO_B200_GTR) return fail(PHYLO_B200_EINVAL, "unknown model");
    if (!peel || (!tipmask && !tipdata)) return fail(PHYLO_B200_EINVAL, "peel / tip data is NULL");
    const bool rooted = flags & PHYLO_B200_ROOTED;
    if (!rooted && S < 3) return fail(PHYLO_B200_EINVAL, "an unrooted tree needs S >= 3");

    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0 || device < 0 || device >= ndev)
        return fail(PHYLO_B200_ENODEV, "no CUDA device " + std::to_string(device) +
                                           " (libphylo_b200 has no CPU fallback)");
    cudaDeviceProp prop;
    CU_TRY(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
        return fail(PHYLO_B200_ENODEV, std::string("device is sm_") + std::to_string(prop.major * 10 + prop.minor) +
                                           ", this library is built for sm_100a only");
    CU_TRY(cudaSetDevice(device));

    phylo_b200_ctx* h = new (std::nothrow) phylo_b200_ctx();
    if (!h) return fail(PHYLO_B200_ENOMEM, "out of host memory");
    h->S = S; h->L = L; h->C = C; h->model = model; h->flags = flags; h->device = device;
    h->rooted = rooted;
    h->normalize = !(flags & PHYLO_B200_NO_NORMQ);
    h->jc_closed = model == PHYLO_B200_JC69 && h->normalize;
    h->nn = 2 * S - 1;
    h->bcount = rooted ? 2 * S - 2 : 2 * S - 3;
    h->nsubst = n_subst(model);
    h->off_subst = 1 + h->bcount;
    h->off_freqs = h->off_subst + h->nsubst;
    h->off_rs = h->off_freqs + 4;
    h->off_ps = h->off_rs + C;
    h->nout = h->off_ps + C;
    h->num_sms = prop.multiProcessorCount;
    h->smem_optin = prop.sharedMemPerBlockOptin;

    std::string err;
    if (!build_plan(S, peel, h->plan, err)) { delete h; return fail(PHYLO_B200_EINVAL, err); }
    if (!rooted && peel[3 * (S - 2) + 1] != 2 * S - 2) {
        delete h;
        return fail(PHYLO_B200_EINVAL, "unrooted: the last peel row must list node 2S-2 second "
                                       "(phylostan/phylostan.py:264-267)");
    }

    ParamLayout& lay = h->lay;
    lay.nn = h->nn; lay.C = C; lay.nsubst = h->nsubst;
    lay.ntheta = model == PHYLO_B200_JC69 ? 0 : h->nsubst + 4;
    int o = 0;
    lay.off_t = o; o += h->nn;
    lay.off_rs = o; o += C;
    lay.off_ps = o; o += C;
    lay.off_pi = o; o += 4;
    lay.off_lam = o; o += 4;
    lay.off_m1 = o; o += 16;
    lay.off_m2 = o; o += 16;
    lay.off_Q = o; o += 16;
    lay.off_X = o; o += 16 * lay.ntheta;
    lay.stride = (o + 1) & ~1;

    // static data: tip codes [S][Lpad] (padding = all-ambiguous, weight 0), weights, step lists
    h->Lpad = ((L + kPadPatterns - 1) / kPadPatterns) * kPadPatterns;
    auto up = [&](auto& buf, const auto& vec) -> cudaError_t {
        cudaError_t e = buf.ensure(vec.size());
        if (e != cudaSuccess) return e;
        return cudaMemcpy(buf.p, vec.data(), vec.size() * sizeof(vec[0]), cudaMemcpyHostToDevice);
    };
    if (on_device) {
        // the alignment already lives on this GPU: pad / classify / re-code it there
        DevBuf<int> d_flags;
        int hf[2] = {0, 0};
        cudaError_t e = cudaSuccess;
        if ((e = h->d_tips.ensure((size_t)S * h->Lpad)) != cudaSuccess ||
            (e = h->d_weights.ensure((size_t)h->Lpad)) != cudaSuccess || (e = d_flags.ensure(2)) != cudaSuccess ||
            (e = cudaMemset(d_flags.p, 0, 2 * sizeof(int))) != cudaSuccess) {
            d_flags.release(); delete h;
            return fail(PHYLO_B200_ECUDA, std::string("device setup: ") + cudaGetErrorString(e));
        }
        launch_tips_prepare(tipmask, h->d_tips.p, S, L, h->Lpad, weights, h->d_weights.p, d_flags.p, nullptr);
        if ((e = cudaGetLastError()) != cudaSuccess ||
            (e = cudaMemcpy(hf, d_flags.p, sizeof hf, cudaMemcpyDeviceToHost)) != cudaSuccess) {
            d_flags.release(); delete h;
            return fail(PHYLO_B200_ECUDA, std::string("tip preparation: ") + cudaGetErrorString(e));
        }
        d_flags.release();
        if (hf[1]) { delete h; return fail(PHYLO_B200_EDOMAIN, "non-finite pattern weight"); }
        h->tips_simple = hf[0] == 0;
        if (h->tips_simple) {
            launch_tips_index(h->d_tips.p, S, h->Lpad, nullptr);
            if ((e = cudaGetLastError()) != cudaSuccess || (e = cudaDeviceSynchronize()) != cudaSuccess) {
                delete h;
                return fail(PHYLO_B200_ECUDA, std::string("tip preparation: ") + cudaGetErrorString(e));
            }
        }
    } else {
        std::vector<uint8_t> tips((size_t)S * h->Lpad, 0xF);
        for (int s = 0; s < S; ++s)
            for (int l = 0; l < L; ++l) {
                uint8_t m;
                if (tipmask) {
                    m = tipmask[(size_t)s * L + l] & 0xF;
                } else {
                    const double* t = tipdata + ((size_t)s * L + l) * 4;
                    m = (uint8_t)((t[0] != 0.0) | ((t[1] != 0.0) << 1) | ((t[2] != 0.0) << 2) | ((t[3] != 0.0) << 3));
                }
                tips[(size_t)s * h->Lpad + l] = m;
            }
        // reference-encoded alignments only hold one-hot or all-ones cells (phylostan/utils.py:180-188):
        // store the column index 0..3 / 4 instead of the mask and use the tip fast path
        h->tips_simple = true;
        for (uint8_t m : tips)
            if (!(m == 15 || m == 1 || m == 2 || m == 4 || m == 8)) { h->tips_simple = false; break; }
        if (h->tips_simple)
            for (uint8_t& m : tips) m = m == 15 ? 4 : (m == 1 ? 0 : (m == 2 ? 1 : (m == 4 ? 2 : 3)));
        std::vector<double> w((size_t)h->Lpad, 0.0);
        for (int l = 0; l < L; ++l) {
            w[l] = weights ? weights[l] : 1.0;
            if (!std::isfinite(w[l])) { delete h; return fail(PHYLO_B200_EDOMAIN, "non-finite pattern weight"); }
        }
        cudaError_t e = cudaSuccess;
        if ((e = up(h->d_tips, tips)) != cudaSuccess || (e = up(h->d_weights, w)) != cudaSuccess) {
            delete h;
            return fail(PHYLO_B200_ECUDA, std::string("device setup: ") + cudaGetErrorString(e));
        }
    }
    std::vector<int32_t> node_pos((size_t)h->nn, 0);  // non-root node -> 2 * post step + child slot
    for (size_t i = 0; i < h->plan.post.size(); ++i) {
        node_pos[h->plan.post[i].a] = (int32_t)(2 * i);
        node_pos[h->plan.post[i].b] = (int32_t)(2 * i + 1);
    }
    std::vector<int32_t> node_row((size_t)h->nn, -1);  // internal node -> its own post-order step = scratch row
    for (size_t i = 0; i < h->plan.post.size(); ++i) node_row[h->plan.post[i].node] = (int32_t)i;
    // table nodes: internal nodes with two (cherry) or three (pitchfork: a cherry and a tip) tips below them
    std::vector<int32_t> node_cherry, cherries, tab_off;
    {
        int max_tips = 3;
        if (const char* tt = std::getenv("PHYLO_B200_TABLE_TIPS")) max_tips = tt[0] == '2' ? 2 : 3;
        h->tab_entries = find_table_nodes(h->plan, S, max_tips, node_cherry, cherries, tab_off);
        h->ncherry = (int)tab_off.size();
    }
    cudaError_t e = cudaSuccess;
    if (h->ncherry > 0 && node_cherry[h->plan.root] < 0) {  // the second plan: table nodes are leaves of its post-order
        std::vector<char> leaf((size_t)h->nn, 0);
        for (int n = 0; n < h->nn; ++n) leaf[n] = node_cherry[n] >= 0;
        std::string err2;
        if (build_plan(S, peel, h->planB, err2, &leaf) && h->planB.depth() <= h->plan.depth()) {
            std::vector<int32_t> rowB((size_t)h->nn, -1);
            for (size_t i = 0; i < h->planB.post.size(); ++i) rowB[h->planB.post[i].node] = (int32_t)i;
            h->has_planB = (e = up(h->d_postB, h->planB.post)) == cudaSuccess && (e = up(h->d_preB, h->planB.pre)) == cudaSuccess &&
                           (e = up(h->d_node_rowB, rowB)) == cudaSuccess;
            if (e != cudaSuccess) { delete h; return fail(PHYLO_B200_ECUDA, std::string("device setup: ") + cudaGetErrorString(e)); }
        }
    }
    if ((e = up(h->d_node_cherry, node_cherry)) != cudaSuccess || (e = up(h->d_cherries, cherries)) != cudaSuccess ||
        (e = up(h->d_tab_off, tab_off)) != cudaSuccess) {
        delete h;
        return fail(PHYLO_B200_ECUDA, std::string("device setup: ") + cudaGetErrorString(e));
    }
    if ((e = up(h->d_node_pos, node_pos)) != cudaSuccess || (e = up(h->d_node_row, node_row)) != cudaSuccess ||
        (e = up(h->d_post, h->plan.post)) != cudaSuccess || (e = up(h->d_pre, h->plan.pre)) != cudaSuccess ||
        (e = cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking)) != cudaSuccess) {
        delete h;
        return fail(PHYLO_B200_ECUDA, std::string("device setup: ") + cudaGetErrorString(e));
    }
    h->stream = h->own_stream;
    if (const char* ng = std::getenv("PHYLO_B200_NO_GRAPH")) h->use_graphs = !(ng[0] && ng[0] != '0');
    if (const char* nj = std::getenv("PHYLO_B200_NO_JC_SCALAR")) h->use_jc_scalar = !(nj[0] && nj[0] != '0');
    if (const char* ms = std::getenv("PHYLO_B200_MSG")) h->use_msg = !(ms[0] == '0');
    if (const char* ch = std::getenv("PHYLO_B200_CHERRY")) h->use_cherry = !(ch[0] == '0');
    if (const char* pt = std::getenv("PHYLO_B200_POST_TABLES")) h->use_post_tables = !(pt[0] == '0');
    if (const char* tm = std::getenv("PHYLO_B200_SWEEP_TM")) h->req_tm = tm[0] == '3' ? 3 : tm[0] == '2' ? 2 : 0;
    for (auto& ev : h->ev)
        if ((e = cudaEventCreate(&ev)) != cudaSuccess) {
            delete h;
            return fail(PHYLO_B200_ECUDA, std::string("cudaEventCreate: ") + cudaGetErrorString(e));
        }
    *out = h;
    return 0;
}

// Pack one draw into the parameter block (host).  Returns false on out-of-domain input.
bool pack_draw(const phylo_b200_ctx* h, const double* blens, const double* subst, const double* freqs,
               const double* rs, const double* ps, double* dst, std::string& why) {
    const ParamLayout& lay = h->lay;
    std::memset(dst, 0, sizeof(double) * lay.stride);
    for (int b = 0; b < h->bcount; ++b) {
        if (!(blens[b] >= 0.0) || !std::isfinite(blens[b])) { why = "branch length " + std::to_string(b) + " is negative or not finite"; return false; }
        dst[lay.off_t + b] = blens[b];
    }
    for (int c = 0; c < h->C; ++c) {
        const double r = rs ? rs[c] : 1.0, p = ps ? ps[c] : 1.0 / h->C;
        if (!(r >= 0.0) || !std::isfinite(r) || !(p >= 0.0) || !std::isfinite(p)) { why = "site rate / proportion is negative or not finite"; return false; }
        dst[lay.off_rs + c] = r;
        dst[lay.off_ps + c] = p;
    }
    Derived dv;
    if (!derive(h->model, h->normalize, subst, freqs, dv)) { why = "substitution parameters out of domain"; return false; }
    std::memcpy(dst + lay.off_pi, dv.pi, sizeof dv.pi);
    std::memcpy(dst + lay.off_lam, dv.lam, sizeof dv.lam);
    std::memcpy(dst + lay.off_m1, dv.m1, sizeof dv.m1);
    std::memcpy(dst + lay.off_m2, dv.m2, sizeof dv.m2);
    std::memcpy(dst + lay.off_Q, dv.Q, sizeof dv.Q);
    for (int k = 0; k < lay.ntheta; ++k) std::memcpy(dst + lay.off_X + 16 * k, dv.X[k], sizeof dv.X[k]);
    return true;
}

}  // namespace

// ------------------------------------------------------------------------------------------ C ABI

extern "C" {

int phylo_b200_abi_version(void) { return PHYLO_B200_ABI_VERSION; }

static phylo_b200_handle g_default = nullptr;
int phylo_b200_set_default(phylo_b200_handle h) { g_default = h; return 0; }
phylo_b200_handle phylo_b200_get_default(void) { return g_default; }

const char* phylo_b200_last_error(void) { return g_err.c_str(); }

int phylo_b200_create(phylo_b200_handle* out, int S, int L, int C, int model, int flags, const int32_t* peel,
                      const uint8_t* tipmask, const double* weights, int device) {
    return create_common(out, S, L, C, model, flags, peel, tipmask, nullptr, weights, device);
}

int phylo_b200_create_tipdata(phylo_b200_handle* out, int S, int L, int C, int model, int flags,
                              const int32_t* peel, const double* tipdata, const double* weights, int device) {
    return create_common(out, S, L, C, model, flags, peel, nullptr, tipdata, weights, device);
}

int phylo_b200_create_device(phylo_b200_handle* out, int S, int L, int C, int model, int flags, const int32_t* peel,
                             const uint8_t* d_tipmask, const double* d_weights, int device) {
    if (!d_tipmask) return fail(PHYLO_B200_EINVAL, "create_device: tip masks are NULL");
    return create_common(out, S, L, C, model, flags, peel, d_tipmask, nullptr, d_weights, device, true);
}

int phylo_b200_create_multi(phylo_b200_handle* out, int S, int L, int C, int model, int flags, const int32_t* peel,
                            const uint8_t* tipmask, const double* weights, const int* devices, int ndev) {
    if (!out) return fail(PHYLO_B200_EINVAL, "out handle is NULL");
    *out = nullptr;
    if (!devices || ndev < 1 || ndev > kMaxPeers + 1)
        return fail(PHYLO_B200_EINVAL, "create_multi: need 1 <= ndev <= " + std::to_string(kMaxPeers + 1) + " devices");
    if (ndev == 1) return create_common(out, S, L, C, model, flags, peel, tipmask, nullptr, weights, devices[0]);
    if (!tipmask || S < 2 || L < ndev) return fail(PHYLO_B200_EINVAL, "create_multi: need tip masks and at least one pattern per device");
    std::vector<phylo_b200_ctx*> ctx;
    auto bail = [&](int rc) {
        const std::string keep = g_err;
        for (auto* c : ctx) delete c;
        g_err = keep;
        return rc;
    };
    for (int i = 0; i < ndev; ++i) {
        const int lo = (int)((long long)L * i / ndev), hi = (int)((long long)L * (i + 1) / ndev), Li = hi - lo;
        std::vector<uint8_t> tm((size_t)S * Li);
        for (int s = 0; s < S; ++s) std::memcpy(tm.data() + (size_t)s * Li, tipmask + (size_t)s * L + lo, (size_t)Li);
        phylo_b200_ctx* c = nullptr;
        if (int rc = create_common(&c, S, Li, C, model, flags, peel, tm.data(), nullptr, weights ? weights + lo : nullptr,
                                   devices[i]))
            return bail(rc);
        ctx.push_back(c);
    }
    phylo_b200_ctx* h = ctx[0];
    const char* nop = std::getenv("PHYLO_B200_NO_P2P");
    const bool allow_p2p = !(nop && nop[0] && nop[0] != '0');
    cudaError_t e = cudaSetDevice(h->device);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->fork_ev, cudaEventDisableTiming);
    for (int i = 1; i < ndev && e == cudaSuccess; ++i) {
        phylo_b200_ctx* p = ctx[i];
        bool direct = p->device == h->device;
        if (!direct && allow_p2p) {
            int can = 0;
            if (cudaDeviceCanAccessPeer(&can, h->device, p->device) == cudaSuccess && can) {
                cudaSetDevice(h->device);
                const cudaError_t pe = cudaDeviceEnablePeerAccess(p->device, 0);
                if (pe == cudaSuccess || pe == cudaErrorPeerAccessAlreadyEnabled) direct = true;
                (void)cudaGetLastError();
            }
        }
        cudaEvent_t ev = nullptr;
        e = cudaSetDevice(p->device);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
        p->is_peer = true;
        h->peers.push_back(p);
        h->peer_ev.push_back(ev);
        h->peer_direct.push_back(direct ? 1 : 0);
    }
    if (e != cudaSuccess) {
        for (int i = (int)h->peers.size() + 1; i < ndev; ++i) delete ctx[i];  // not yet owned by h
        delete h;
        return fail(PHYLO_B200_ECUDA, std::string("create_multi: ") + cudaGetErrorString(e));
    }
    cudaSetDevice(h->device);
    h->use_graphs = false;
    *out = h;
    return 0;
}

void phylo_b200_destroy(phylo_b200_handle h) {
    if (h && h == g_default) g_default = nullptr;
    delete h;
}

int phylo_b200_bcount(phylo_b200_handle h) { return h ? h->bcount : PHYLO_B200_EINVAL; }
int phylo_b200_nsubst(phylo_b200_handle h) { return h ? h->nsubst : PHYLO_B200_EINVAL; }
int phylo_b200_ncat(phylo_b200_handle h) { return h ? h->C : PHYLO_B200_EINVAL; }
int phylo_b200_nout(phylo_b200_handle h) { return h ? h->nout : PHYLO_B200_EINVAL; }

int phylo_b200_set_stream(phylo_b200_handle h, void* stream) {
    if (!h) return fail(PHYLO_B200_EINVAL, "NULL handle");
    h->stream = stream ? (cudaStream_t)stream : h->own_stream;
    return 0;
}

int phylo_b200_sync(phylo_b200_handle h) {
    if (!h) return fail(PHYLO_B200_EINVAL, "NULL handle");
    CU_TRY(cudaSetDevice(h->device));
    CU_TRY(cudaStreamSynchronize(h->stream));
    return 0;
}

int phylo_b200_set_tiling(phylo_b200_handle h, int patterns_per_thread, int pattern_blocks) {
    if (!h) return fail(PHYLO_B200_EINVAL, "NULL handle");
    if (patterns_per_thread != 0 && patterns_per_thread != 1 && patterns_per_thread != 2 && patterns_per_thread != 4)
        return fail(PHYLO_B200_EINVAL, "patterns_per_thread must be 0, 1, 2 or 4");
    if (pattern_blocks != 0 && pattern_blocks != 1 && pattern_blocks != 2 && pattern_blocks != 4)
        return fail(PHYLO_B200_EINVAL, "pattern_blocks must be 0, 1, 2 or 4");
    h->req_K = patterns_per_thread;
    h->req_PB = pattern_blocks;
    for (auto* p : h->peers) { p->req_K = patterns_per_thread; p->req_PB = pattern_blocks; }
    return 0;
}

int phylo_b200_set_sweep_variant(phylo_b200_handle h, int ctas_per_sm) {
    if (!h) return fail(PHYLO_B200_EINVAL, "NULL handle");
    if (ctas_per_sm != -1 && ctas_per_sm != 0 && ctas_per_sm != 2 && ctas_per_sm != 3)
        return fail(PHYLO_B200_EINVAL, "sweep variant must be -1 (default), 0, 2 or 3");
    h->req_tm = ctas_per_sm < 0 ? kDefaultSweepTm : ctas_per_sm;
    for (auto* p : h->peers) p->req_tm = h->req_tm;
    return 0;
}

int phylo_b200_set_cherry_tables(phylo_b200_handle h, int enabled) {
    if (!h) return fail(PHYLO_B200_EINVAL, "NULL handle");
    h->use_cherry = enabled != 0;
    for (auto* p : h->peers) p->use_cherry = h->use_cherry;
    return 0;
}

int phylo_b200_set_stack_slots(phylo_b200_handle h, int slots) {
    if (!h) return fail(PHYLO_B200_EINVAL, "NULL handle");
    if (slots < 0) return fail(PHYLO_B200_EINVAL, "slots must be >= 0 (0 = automatic)");
    h->req_cap = slots;
    for (auto* p : h->peers) p->req_cap = slots;
    return 0;
}

int phylo_b200_set_precision(phylo_b200_handle h, int bits) {
    if (!h) return fail(PHYLO_B200_EINVAL, "NULL handle");
    if (bits != 32 && bits != 64) return fail(PHYLO_B200_EINVAL, "precision must be 64 or 32");
    h->prec = bits;
    for (auto* p : h->peers) p->prec = bits;
    return 0;
}

int phylo_b200_set_timing(phylo_b200_handle h, int enabled) {
    if (!h) return fail(PHYLO_B200_EINVAL, "NULL handle");
    h->timing = enabled != 0;
    h->ev_valid = false;
    return 0;
}

int phylo_b200_get_timing(phylo_b200_handle h, double ms[4]) {
    if (!h || !ms) return fail(PHYLO_B200_EINVAL, "NULL argument");
    if (!h->ev_valid) return fail(PHYLO_B200_EINVAL, "no timed run recorded (call set_timing(1) then run)");
    CU_TRY(cudaSetDevice(h->device));
    CU_TRY(cudaEventSynchronize(h->ev[4]));
    float t = 0;
    CU_TRY(cudaEventElapsedTime(&t, h->ev[0], h->ev[1])); ms[0] = t;
    CU_TRY(cudaEventElapsedTime(&t, h->ev[1], h->ev[2])); ms[1] = t;
    ms[2] = 0;
    if (h->ev_has_contract) { CU_TRY(cudaEventElapsedTime(&t, h->ev[2], h->ev[3])); ms[2] = t; }
    CU_TRY(cudaEventElapsedTime(&t, h->ev[0], h->ev[4])); ms[3] = t;
    return 0;
}

long long phylo_b200_info(phylo_b200_handle h, int what) {
    if (!h) return PHYLO_B200_EINVAL;
    switch (what) {
        case 0: return h->plan.depth();
        case 1: return h->K;
        case 2: return h->NT;
        case 3: return h->grid;
        case 4: return (long long)h->smem;
        case 5: return h->Lpad;
        case 6: return h->last_launches;
        case 7: return (long long)(h->d_scratch.n * sizeof(double2) + h->d_dscr.n);
        case 8: return h->plan.depth_post;
        case 9: return h->plan.depth_pre;
        case 10: return h->ntiles;
        case 11: return h->slots;
        case 12: return 1 + (long long)h->peers.size();
        case 13: return h->tm;
        case 14: return h->msg_run ? 1 : 0;
        case 15: return h->cherry_run ? 1 : 0;
        case 16: return h->post_tables_run ? 1 : 0;
    }
    return PHYLO_B200_EINVAL;
}

namespace {
// validate, size the per-batch buffers and pack the draws into the pinned staging buffer (no copy yet)
// status != NULL: a draw that fails validation is marked 1 and gets the parameter block of the first valid draw
// (so the kernels have something harmless to chew on); *nvalid = number of valid draws
int pack_batch(phylo_b200_ctx* h, int B, const double* blens, const double* subst, const double* freqs,
               const double* rs, const double* ps, int32_t* status = nullptr, int* nvalid = nullptr) {
    if (!h || B < 1 || !blens) return fail(PHYLO_B200_EINVAL, "upload: bad arguments");
    if (h->nsubst > 0 && !subst) return fail(PHYLO_B200_EINVAL, "upload: subst is NULL");
    if (h->model != PHYLO_B200_JC69 && !freqs) return fail(PHYLO_B200_EINVAL, "upload: freqs is NULL");
    for (auto* p : h->peers) {  // the shards on the other devices read the same packed block
        CU_TRY(cudaSetDevice(p->device));
        if (int rc = ensure_batch(p, B)) return rc;
        CU_TRY(cudaStreamSynchronize(p->stream));
    }
    CU_TRY(cudaSetDevice(h->device));
    if (int rc = ensure_batch(h, B)) return rc;
    // the pinned staging buffer may still feed a copy in flight
    CU_TRY(cudaStreamSynchronize(h->stream));
    std::string why;
    int first_ok = -1, ok = 0;
    h->tau_bound = 0.0;
    h->tau_unit = 0.0;
    h->tau_cond = 0.0;
    for (int d = 0; d < B; ++d) {
        const bool good = pack_draw(h, blens + (size_t)d * h->bcount, subst ? subst + (size_t)d * h->nsubst : nullptr,
                                    freqs ? freqs + (size_t)d * 4 : nullptr, rs ? rs + (size_t)d * h->C : nullptr,
                                    ps ? ps + (size_t)d * h->C : nullptr, h->h_params.p + (size_t)d * h->lay.stride, why);
        if (!good && !status) return fail(PHYLO_B200_EDOMAIN, "draw " + std::to_string(d) + ": " + why);
        if (status) status[d] = good ? 0 : 1;
        if (good) { ++ok; if (first_ok < 0) first_ok = d; }
        if (good) {
            const double* pd = h->h_params.p + (size_t)d * h->lay.stride;
            double tmax = 0.0, rmax = 0.0, lmin = 0.0, lmax = 0.0;
            for (int b = 0; b < h->bcount; ++b) tmax = std::max(tmax, pd[h->lay.off_t + b]);
            for (int c = 0; c < h->C; ++c) rmax = std::max(rmax, pd[h->lay.off_rs + c]);
            for (int k = 0; k < 4; ++k) { lmin = std::min(lmin, pd[h->lay.off_lam + k]); lmax = std::max(lmax, pd[h->lay.off_lam + k]); }
            // the eigenvector matrices' conditioning multiplies the same rounding error (skewed frequencies): it is
            // counted in the exponent, log(|m1|_F |m2|_F / 4) being 0 for equal frequencies
            double n1 = 0.0, n2 = 0.0;
            for (int k = 0; k < 16; ++k) { n1 += pd[h->lay.off_m1 + k] * pd[h->lay.off_m1 + k]; n2 += pd[h->lay.off_m2 + k] * pd[h->lay.off_m2 + k]; }
            const double cnd = std::max(0.0, 0.5 * std::log(n1 * n2) - std::log(4.0));
            h->tau_unit = std::max(h->tau_unit, rmax * (lmax - lmin));
            h->tau_cond = std::max(h->tau_cond, cnd);
            h->tau_bound = std::max(h->tau_bound, tmax * rmax * (lmax - lmin) + cnd);
        }
    }
    for (auto* p : h->peers) { p->tau_bound = h->tau_bound; p->tau_unit = h->tau_unit; p->tau_cond = h->tau_cond; }
    if (status && first_ok >= 0)
        for (int d = 0; d < B; ++d)
            if (status[d])
                std::memcpy(h->h_params.p + (size_t)d * h->lay.stride, h->h_params.p + (size_t)first_ok * h->lay.stride,
                            sizeof(double) * h->lay.stride);
    if (nvalid) *nvalid = ok;
    return 0;
}
}  // namespace

int phylo_b200_upload(phylo_b200_handle h, int B, const double* blens, const double* subst, const double* freqs,
                      const double* rs, const double* ps) {
    if (int rc = pack_batch(h, B, blens, subst, freqs, rs, ps)) return rc;
    for (auto* p : h->peers) {
        CU_TRY(cudaSetDevice(p->device));
        CU_TRY(cudaMemcpyAsync(p->d_params.p, h->h_params.p, sizeof(double) * B * h->lay.stride, cudaMemcpyHostToDevice,
                               p->stream));
    }
    CU_TRY(cudaSetDevice(h->device));
    CU_TRY(cudaMemcpyAsync(h->d_params.p, h->h_params.p, sizeof(double) * B * h->lay.stride, cudaMemcpyHostToDevice,
                           h->stream));
    return 0;
}

namespace {

// launch shape + scratch for B draws: everything that allocates or queries, i.e. all that may not
// happen while a stream is being captured
int run_prepare(phylo_b200_ctx* h, int B, bool grad) {
    if ((size_t)B * h->lay.stride > h->d_params.n) return fail(PHYLO_B200_EINVAL, "run: upload B draws first");
    if (int rc = resolve_tiling(h, B, grad)) return rc;
    if (sweep_uses_tipring(h->tips_simple, h->NT, grad) && (h->tipring_K != h->K || (grad && !h->tipring_pre))) {
        const int S = h->S, T = 32 * h->K, nblocks = h->Lpad / T;
        if (!h->d_tip_order.p) {  // slot -> tip, in the order each sweep meets its tip children (a before b)
            std::vector<int32_t> order;
            for (const PostStep& p : h->plan.post) {
                if (p.a < S) order.push_back(p.a);
                if (p.b < S) order.push_back(p.b);
            }
            for (const PreStep& p : h->plan.pre) {
                if (p.a < S) order.push_back(p.a);
                if (p.b < S) order.push_back(p.b);
            }
            if ((int)order.size() != 2 * S) return fail(PHYLO_B200_EINVAL, "internal: every tip must be a child exactly once");
            CU_TRY(h->d_tip_order.ensure(order.size()));
            CU_TRY(cudaMemcpy(h->d_tip_order.p, order.data(), order.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
        }
        CU_TRY(cudaStreamSynchronize(h->stream));  // a sweep reading the old copies may still run
        if (h->tipring_K != h->K) {
            h->tipring_K = 0; h->tipring_pre = false;
            CU_TRY(h->d_tips_post.ensure((size_t)S * h->Lpad));
            launch_tips_reorder(h->d_tips.p, h->d_tips_post.p, h->d_tip_order.p, S, h->Lpad, T, nblocks, h->stream);
            CU_TRY(cudaGetLastError());
        }
        if (grad && !h->tipring_pre) {
            CU_TRY(h->d_tips_pre.ensure((size_t)S * h->Lpad));
            launch_tips_reorder(h->d_tips.p, h->d_tips_pre.p, h->d_tip_order.p + S, S, h->Lpad, T, nblocks, h->stream);
            CU_TRY(cudaGetLastError());
            h->tipring_pre = true;
        }
        h->tipring_K = h->K;
    }
    h->cherry_run = grad && h->msg_run && !h->tm && h->use_cherry && h->ncherry > 0 && h->tips_simple &&
                    sweep_cherry_available(h->K);
    h->post_tables_run = h->cherry_run && h->use_post_tables && h->has_planB;
    if (h->cherry_run) {
        if (!h->ctips_built) {
            CU_TRY(h->d_ctips.ensure((size_t)h->ncherry * h->Lpad));
            launch_cherry_codes(h->d_tips.p, h->d_ctips.p, h->d_cherries.p, h->ncherry, h->Lpad, h->stream);
            CU_TRY(cudaGetLastError());
            h->ctips_built = true;
        }
        CU_TRY(h->d_ctab.ensure((size_t)B * h->C * h->tab_entries * 4));
    }
    if (grad) {
        const size_t rows = (size_t)h->grid * (h->S - 1) * h->K * h->NT;
        CU_TRY(h->d_scratch.ensure(rows * 2));  // 16-byte vectors; fp32 uses half of them
        CU_TRY(h->d_dscr.ensure(rows));
    }
    return 0;
}

int run_enqueue(phylo_b200_ctx* h, int B, bool grad);

// a clock-tree front-end call in flight: what the kernels of one shard need besides the context
struct ClockJob {
    int autocorr, ratios, has_extra, nrates, in_ld, hout_ld, want_heights;
};

ClockArgs clock_args(const phylo_b200_ctx* c, const ClockJob& j, int B) {
    ClockArgs a{};
    a.map = c->d_cmap.p; a.kids = c->d_ckids.p; a.row_of = c->d_crow.p;
    a.lowers = c->c_has_lowers ? c->d_clowers.p : nullptr;
    a.in = c->d_cin.p; a.params = c->d_params.p; a.out = c->d_out.p; a.hwork = c->d_hwork.p; a.hout = c->d_hout.p;
    a.S = c->S; a.nn = c->nn; a.nrates = j.nrates; a.autocorr = j.autocorr; a.ratios = j.ratios; a.has_extra = j.has_extra;
    a.B = B; a.in_ld = j.in_ld; a.hout_ld = j.hout_ld; a.off_t = c->lay.off_t; a.stride = c->lay.stride; a.nout = c->nout;
    return a;
}

// H2D of the front end's per-draw inputs (staged in `src`'s pinned block) and the heights -> blens kernel, on
// c's stream; the packed parameter block must already be on its way (same stream)
int clock_forward_enqueue(phylo_b200_ctx* c, const phylo_b200_ctx* src, const ClockJob& j, int B) {
    CU_TRY(cudaMemcpyAsync(c->d_cin.p, src->h_cin.p, sizeof(double) * B * j.in_ld, cudaMemcpyHostToDevice, c->stream));
    launch_clock_forward(clock_args(c, j, B), c->stream);
    CU_TRY(cudaGetLastError());
    return 0;
}

// run_prepare on every shard of the handle
int run_prepare_all(phylo_b200_ctx* h, int B, bool grad) {
    for (auto* p : h->peers) {
        CU_TRY(cudaSetDevice(p->device));
        if (int rc = run_prepare(p, B, grad)) return rc;
    }
    CU_TRY(cudaSetDevice(h->device));
    return run_prepare(h, B, grad);
}

// Multi-device handle: fork from this context's stream, run every shard on its own device and stream, join,
// and add the peers' result rows to d_out on this device.  with_copies: H2D of the packed parameters (from
// this context's pinned block) before, D2H of the summed rows after.
int multi_enqueue(phylo_b200_ctx* h, int B, bool grad, bool with_copies, const ClockJob* job = nullptr) {
    const size_t in_bytes = sizeof(double) * B * h->lay.stride, count = (size_t)B * h->nout;
    CU_TRY(cudaSetDevice(h->device));
    CU_TRY(cudaEventRecord(h->fork_ev, h->stream));
    for (size_t i = 0; i < h->peers.size(); ++i) {
        phylo_b200_ctx* p = h->peers[i];
        CU_TRY(cudaSetDevice(p->device));
        CU_TRY(cudaStreamWaitEvent(p->stream, h->fork_ev, 0));  // also orders the peer's memsets after the last sum
        if (with_copies) CU_TRY(cudaMemcpyAsync(p->d_params.p, h->h_params.p, in_bytes, cudaMemcpyHostToDevice, p->stream));
        if (job) { if (int rc = clock_forward_enqueue(p, h, *job, B)) return rc; }
        if (int rc = run_enqueue(p, B, grad)) return rc;
        CU_TRY(cudaEventRecord(h->peer_ev[i], p->stream));
    }
    CU_TRY(cudaSetDevice(h->device));
    if (with_copies) CU_TRY(cudaMemcpyAsync(h->d_params.p, h->h_params.p, in_bytes, cudaMemcpyHostToDevice, h->stream));
    if (job) { if (int rc = clock_forward_enqueue(h, h, *job, B)) return rc; }
    if (int rc = run_enqueue(h, B, grad)) return rc;
    PeerRows rows{};
    rows.n = (int)h->peers.size();
    for (size_t i = 0; i < h->peers.size(); ++i) {
        phylo_b200_ctx* p = h->peers[i];
        CU_TRY(cudaStreamWaitEvent(h->stream, h->peer_ev[i], 0));
        if (h->peer_direct[i]) {
            rows.src[i] = p->d_out.p;
        } else {
            double* dst = h->d_stage.p + i * count;
            CU_TRY(cudaMemcpyPeerAsync(dst, h->device, p->d_out.p, p->device, sizeof(double) * count, h->stream));
            rows.src[i] = dst;
        }
    }
    launch_peer_sum(h->d_out.p, rows, count, h->stream);
    CU_TRY(cudaGetLastError());
    h->last_launches += 1;
    if (with_copies)
        CU_TRY(cudaMemcpyAsync(h->h_out.p, h->d_out.p, sizeof(double) * count, cudaMemcpyDeviceToHost, h->stream));
    return 0;
}

}  // namespace

int phylo_b200_run(phylo_b200_handle h, int B, int want_grad) {
    if (!h || B < 1) return fail(PHYLO_B200_EINVAL, "run: bad arguments");
    if (int rc = run_prepare_all(h, B, want_grad != 0)) return rc;
    if (!h->peers.empty()) return multi_enqueue(h, B, want_grad != 0, false);
    return run_enqueue(h, B, want_grad != 0);
}

namespace {

int run_enqueue(phylo_b200_ctx* h, int B, bool grad) {
    cudaStream_t st = h->stream;
    CU_TRY(cudaMemsetAsync(h->d_out.p, 0, sizeof(double) * B * h->nout, st));
    if (grad) CU_TRY(cudaMemsetAsync(h->d_G.p, 0, sizeof(double) * B * h->nn * h->C * 16, st));
    if (h->timing) CU_TRY(cudaEventRecord(h->ev[0], st));
    StreamArgs sa{};
    const bool ptab = grad && h->msg_run && h->cherry_run && h->post_tables_run;
    sa.params = h->d_params.p; sa.post = ptab ? h->d_postB.p : h->d_post.p; sa.pre = ptab ? h->d_preB.p : h->d_pre.p;
    sa.npost = ptab ? (int)h->planB.post.size() : h->S - 1; sa.post_tables = ptab ? 1 : 0;
    sa.spost = h->d_spost.p; sa.spre = h->d_spre.p; sa.lay = h->lay;
    sa.nsteps = h->S - 1; sa.bcount = h->bcount; sa.jc_closed = h->jc_closed; sa.B = B;
    const int VP = h->prec == 32 ? 1 : 2;  // 16-byte vectors per 4-state entry
    sa.Lpad = h->Lpad; sa.SS = h->K * VP * h->NT; sa.KNT = h->K * h->NT;
    sa.S = h->S; sa.tips_simple = h->tips_simple;
    sa.node_row = ptab ? h->d_node_rowB.p : h->d_node_row.p;
    sa.slots = grad ? h->slots : h->plan.depth();
    sa.slot_stride = grad && h->tm ? 8 * h->K : sa.SS;
    const bool msg = grad && h->msg_run;
    sa.msg = msg ? 1 : 0;
    const bool cherry = msg && h->cherry_run;
    sa.node_cherry = cherry ? h->d_node_cherry.p : nullptr;
    sa.ctips_off = cherry ? (long long)(h->d_ctips.p - h->d_tips.p) : 0;
    sa.ctab = cherry ? h->d_ctab.p : nullptr; sa.tab_off = h->d_tab_off.p; sa.tab_entries = h->tab_entries;
    launch_stream(sa, h->prec, st);
    CU_TRY(cudaGetLastError());
    if (cherry) {
        CherryArgs ch{};
        ch.params = h->d_params.p; ch.cherries = h->d_cherries.p; ch.ctab = h->d_ctab.p; ch.lay = h->lay;
        ch.B = B; ch.C = h->C; ch.ncherry = h->ncherry; ch.tab_entries = h->tab_entries; ch.bcount = h->bcount; ch.jc_closed = h->jc_closed;
        ch.norescale = ptab ? 1 : 0;
        launch_cherry_tables(ch, st);
        CU_TRY(cudaGetLastError());
    }
    if (h->timing) CU_TRY(cudaEventRecord(h->ev[1], st));

    SweepArgs a{};
    a.tips = h->d_tips.p; a.tips_post = h->d_tips_post.p; a.tips_pre = h->d_tips_pre.p;
    a.weights = h->d_weights.p; a.params = h->d_params.p;
    a.spost = h->d_spost.p; a.spre = h->d_spre.p;
    a.scratch = h->d_scratch.p; a.dscr = h->d_dscr.p; a.G = h->d_G.p; a.out = h->d_out.p;
    a.lay = h->lay;
    a.scratch_stride = (long long)(h->S - 1) * h->K * VP * h->NT;
    a.dscr_stride = (long long)(h->S - 1) * h->K * h->NT;
    a.S = h->S; a.nsteps = h->S - 1; a.npost = sa.npost; a.Lpad = h->Lpad; a.ntiles = h->ntiles; a.nitems = B * h->ntiles;
    a.C = h->C; a.nn = h->nn; a.nout = h->nout; a.D = h->slots;
    a.stack_bytes = (int)sweep_stack_bytes(h->slots, h->K, h->NT, h->prec);
    a.off_out_freqs = h->off_freqs; a.off_out_ps = h->off_ps;
    const bool deep = grad && h->slots < h->plan.depth();
    const bool jc = grad && h->jc_run;
    if (grad && h->tm) CU_TRY(launch_sweep_tm(a, h->tips_simple, h->K, h->tm, h->grid, st, msg));
    else CU_TRY(launch_sweep(a, h->prec, h->tips_simple, h->K, grad, deep, h->grid, h->NT, h->smem, st, jc, msg, cherry));
    if (h->timing) CU_TRY(cudaEventRecord(h->ev[2], st));
    h->last_launches = cherry ? 3 : 2;
    if (grad) {
        ContractArgs ca{};
        ca.spost = h->d_spost.p; ca.node_pos = h->d_node_pos.p; ca.nsteps = h->S - 1; ca.params = h->d_params.p; ca.G = h->d_G.p; ca.out = h->d_out.p; ca.lay = h->lay;
        ca.S = h->S; ca.tips_simple = h->tips_simple; ca.jc_scalar = jc ? 1 : 0; ca.msg = msg ? 1 : 0;
        ca.bcount = h->bcount; ca.C = h->C; ca.nn = h->nn; ca.nout = h->nout; ca.nsubst = h->nsubst;
        ca.off_out_subst = h->off_subst; ca.off_out_freqs = h->off_freqs; ca.off_out_rs = h->off_rs;
        launch_contract(ca, h->prec, B, st);
        CU_TRY(cudaGetLastError());
        if (h->timing) CU_TRY(cudaEventRecord(h->ev[3], st));
        h->last_launches = cherry ? 4 : 3;
    }
    if (h->timing) {
        CU_TRY(cudaEventRecord(h->ev[4], st));
        h->ev_valid = true;
        h->ev_has_contract = grad;
    }
    return 0;
}

// everything a captured eval graph has baked into its nodes; any change means capture again
std::vector<unsigned long long> graph_signature(const phylo_b200_ctx* h) {
    auto u = [](const void* p) { return (unsigned long long)(uintptr_t)p; };
    return {u(h->d_params.p), u(h->d_spost.p), u(h->d_spre.p), u(h->d_G.p), u(h->d_out.p), u(h->d_scratch.p),
            u(h->d_dscr.p), u(h->h_params.p), u(h->h_out.p), u(h->stream), (unsigned long long)h->K,
            (unsigned long long)h->NT, (unsigned long long)h->grid, (unsigned long long)h->smem,
            (unsigned long long)h->slots, (unsigned long long)h->prec, (unsigned long long)h->ntiles,
            (unsigned long long)h->jc_run, u(h->d_tips_post.p), u(h->d_tips_pre.p), (unsigned long long)h->tm,
            (unsigned long long)h->msg_run, (unsigned long long)h->cherry_run, u(h->d_ctab.p), u(h->d_ctips.p),
            (unsigned long long)h->post_tables_run};
}

// H2D of the packed parameters, the kernels, D2H of the result rows -- as one graph launch when possible
// the part of a front-end call that follows the likelihood kernels: reverse sweep, D2H of its results
int clock_tail_enqueue(phylo_b200_ctx* h, const ClockJob& job, int B, bool grad) {
    if (grad) {
        launch_clock_reverse(clock_args(h, job, B), h->stream);
        CU_TRY(cudaGetLastError());
    }
    CU_TRY(cudaMemcpyAsync(h->h_hout.p, h->d_hout.p, sizeof(double) * B * job.hout_ld, cudaMemcpyDeviceToHost, h->stream));
    if (job.want_heights)  // [heights | adjoint] rows of the work buffer: only the first half is wanted
        CU_TRY(cudaMemcpy2DAsync(h->h_hts.p, sizeof(double) * (h->S - 1), h->d_hwork.p, sizeof(double) * 2 * (h->S - 1),
                                 sizeof(double) * (h->S - 1), B, cudaMemcpyDeviceToHost, h->stream));
    return 0;
}

// everything between the packed host blocks and the host result blocks, on h's stream
int eval_sequence(phylo_b200_ctx* h, int B, bool grad, const ClockJob* job) {
    const size_t in_bytes = sizeof(double) * B * h->lay.stride, out_bytes = sizeof(double) * B * h->nout;
    CU_TRY(cudaMemcpyAsync(h->d_params.p, h->h_params.p, in_bytes, cudaMemcpyHostToDevice, h->stream));
    if (job) { if (int rc = clock_forward_enqueue(h, h, *job, B)) return rc; }
    if (int rc = run_enqueue(h, B, grad)) return rc;
    CU_TRY(cudaMemcpyAsync(h->h_out.p, h->d_out.p, out_bytes, cudaMemcpyDeviceToHost, h->stream));
    if (job) { if (int rc = clock_tail_enqueue(h, *job, B, grad)) return rc; }
    return 0;
}

int eval_enqueue(phylo_b200_ctx* h, int B, bool grad, const ClockJob* job = nullptr) {
    if (!h->peers.empty()) {  // one stream per device: plain launches
        if (int rc = multi_enqueue(h, B, grad, true, job)) return rc;
        return job ? clock_tail_enqueue(h, *job, B, grad) : 0;
    }
    const bool graphable = h->use_graphs && !h->timing && h->stream != nullptr;
    const int nlaunch = (grad ? 3 : 2) + (grad && h->cherry_run ? 1 : 0) + (job ? (grad ? 2 : 1) : 0);
    if (!graphable) {
        if (int rc = eval_sequence(h, B, grad, job)) return rc;
        h->last_launches = nlaunch;
        return 0;
    }
    std::vector<int> key{B, grad ? 1 : 0};
    if (job) key.insert(key.end(), {1, job->autocorr, job->ratios, job->has_extra, job->nrates, job->want_heights});
    auto& g = h->graphs[key];
    auto sig = graph_signature(h);
    if (job) {
        auto u = [](const void* p) { return (unsigned long long)(uintptr_t)p; };
        sig.insert(sig.end(), {u(h->d_cin.p), u(h->d_hwork.p), u(h->d_hout.p), u(h->h_cin.p), u(h->h_hout.p), u(h->h_hts.p),
                               u(h->d_cmap.p), u(h->d_ckids.p), u(h->d_crow.p), u(h->d_clowers.p),
                               (unsigned long long)h->c_has_lowers});
    }
    if (!g.exec || g.sig != sig) {
        if (g.exec) { cudaGraphExecDestroy(g.exec); g.exec = nullptr; }
        if (h->graphs.size() > 16) {  // a caller cycling through many batch sizes: keep the cache small
            for (auto it = h->graphs.begin(); it != h->graphs.end();) {
                if (&it->second != &g) { if (it->second.exec) cudaGraphExecDestroy(it->second.exec); it = h->graphs.erase(it); }
                else ++it;
            }
        }
        cudaGraph_t graph = nullptr;
        if (cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeRelaxed) != cudaSuccess) {
            (void)cudaGetLastError();  // a stream that cannot be captured (e.g. the legacy default stream): plain launches
            h->use_graphs = false;
            h->graphs.erase(key);
            return eval_enqueue(h, B, grad, job);
        }
        const int rc = eval_sequence(h, B, grad, job);
        const std::string why = g_err;
        const cudaError_t e2 = cudaStreamEndCapture(h->stream, &graph);
        if (rc) { if (graph) cudaGraphDestroy(graph); (void)cudaGetLastError(); return fail(rc, why); }
        if (e2 != cudaSuccess) {
            if (graph) cudaGraphDestroy(graph);
            return fail(PHYLO_B200_ECUDA, std::string("graph capture: ") + cudaGetErrorString(e2));
        }
        const cudaError_t e = cudaGraphInstantiate(&g.exec, graph, 0);
        cudaGraphDestroy(graph);
        if (e != cudaSuccess) { g.exec = nullptr; return fail(PHYLO_B200_ECUDA, std::string("graph instantiate: ") + cudaGetErrorString(e)); }
        g.sig = sig;
    }
    CU_TRY(cudaGraphLaunch(g.exec, h->stream));
    h->last_launches = nlaunch;
    return 0;
}

}  // namespace

int phylo_b200_device_out(phylo_b200_handle h, void** dptr, int* ld) {
    if (!h || !dptr) return fail(PHYLO_B200_EINVAL, "NULL argument");
    *dptr = h->d_out.p;
    if (ld) *ld = h->nout;
    return h->d_out.p ? 0 : fail(PHYLO_B200_EINVAL, "no output buffer yet (upload first)");
}

int phylo_b200_download(phylo_b200_handle h, int B, double* out) {
    if (!h || B < 1 || !out) return fail(PHYLO_B200_EINVAL, "download: bad arguments");
    if ((size_t)B * h->nout > h->d_out.n) return fail(PHYLO_B200_EINVAL, "download: nothing was run for B draws");
    CU_TRY(cudaSetDevice(h->device));
    CU_TRY(cudaMemcpyAsync(h->h_out.p, h->d_out.p, sizeof(double) * B * h->nout, cudaMemcpyDeviceToHost, h->stream));
    CU_TRY(cudaStreamSynchronize(h->stream));
    std::memcpy(out, h->h_out.p, sizeof(double) * B * h->nout);
    return 0;
}

static int eval_batch_impl(phylo_b200_handle h, int B, const double* blens, const double* subst, const double* freqs,
                           const double* rs, const double* ps, int want_grad, double* logp, double* g_blens,
                           double* g_subst, double* g_freqs, double* g_rs, double* g_ps, int32_t* status) {
    if (!h || !logp) return fail(PHYLO_B200_EINVAL, "eval: NULL handle or logp");
    int nvalid = B;
    if (int rc = pack_batch(h, B, blens, subst, freqs, rs, ps, status, &nvalid)) return rc;
    if (nvalid > 0) {  // (status calls only) nothing to run when every draw was rejected
        if (int rc = run_prepare_all(h, B, want_grad != 0)) return rc;
        if (int rc = eval_enqueue(h, B, want_grad != 0)) return rc;
        CU_TRY(cudaStreamSynchronize(h->stream));
    }
    bool finite = true;
    for (int d = 0; d < B; ++d) {
        const double* o = h->h_out.p + (size_t)d * h->nout;
        const bool ran = nvalid > 0 && !(status && status[d]);
        const bool good = ran && std::isfinite(o[0]);
        finite = finite && (good || !ran);
        if (status && ran && !good) status[d] = 2;
        const bool blank = status && !good;  // a rejected draw: -inf and a zero gradient
        logp[d] = blank ? -std::numeric_limits<double>::infinity() : o[0];
        if (!want_grad) continue;
        auto put = [&](double* dst, size_t n, const double* src) {
            if (blank) std::memset(dst, 0, sizeof(double) * n);
            else std::memcpy(dst, src, sizeof(double) * n);
        };
        if (g_blens) put(g_blens + (size_t)d * h->bcount, h->bcount, o + 1);
        if (g_subst && h->nsubst) put(g_subst + (size_t)d * h->nsubst, h->nsubst, o + h->off_subst);
        if (g_freqs) put(g_freqs + (size_t)d * 4, 4, o + h->off_freqs);
        if (g_rs) put(g_rs + (size_t)d * h->C, h->C, o + h->off_rs);
        if (g_ps) put(g_ps + (size_t)d * h->C, h->C, o + h->off_ps);
    }
    if (!status && !finite) return fail(PHYLO_B200_EDOMAIN, "log-likelihood is not finite (impossible pattern or underflow)");
    return 0;
}

int phylo_b200_eval_batch(phylo_b200_handle h, int B, const double* blens, const double* subst,
                          const double* freqs, const double* rs, const double* ps, int want_grad, double* logp,
                          double* g_blens, double* g_subst, double* g_freqs, double* g_rs, double* g_ps) {
    return eval_batch_impl(h, B, blens, subst, freqs, rs, ps, want_grad, logp, g_blens, g_subst, g_freqs, g_rs, g_ps,
                           nullptr);
}

int phylo_b200_eval_batch_status(phylo_b200_handle h, int B, const double* blens, const double* subst,
                                 const double* freqs, const double* rs, const double* ps, int want_grad, double* logp,
                                 double* g_blens, double* g_subst, double* g_freqs, double* g_rs, double* g_ps,
                                 int32_t* status) {
    if (!status) return fail(PHYLO_B200_EINVAL, "eval_batch_status: status is NULL");
    return eval_batch_impl(h, B, blens, subst, freqs, rs, ps, want_grad, logp, g_blens, g_subst, g_freqs, g_rs, g_ps,
                           status);
}

int phylo_b200_eval(phylo_b200_handle h, const double* blens, const double* subst, const double* freqs,
                    const double* rs, const double* ps, int want_grad, double* logp, double* g_blens,
                    double* g_subst, double* g_freqs, double* g_rs, double* g_ps) {
    return phylo_b200_eval_batch(h, 1, blens, subst, freqs, rs, ps, want_grad, logp, g_blens, g_subst, g_freqs,
                                 g_rs, g_ps);
}

// The pre-order map of phylostan/utils.py:84-90: row 0 = (root, 0); every other node exactly once, after its
// parent, under an internal parent with exactly two children.  One check for every front-end entry.
static int validate_map(const char* who, int S, const int32_t* map) {
    const int nn = 2 * S - 1;
    if (map[0] != nn) return fail(PHYLO_B200_EINVAL, std::string(who) + ": map row 0 must hold the root (node 2S-1)");
    std::vector<int> row(nn + 1, -1), nkids(nn + 1, 0);
    row[nn] = 0;
    for (int j = 1; j < nn; ++j) {
        const int node = map[2 * j], par = map[2 * j + 1];
        if (node < 1 || node >= nn || par <= S || par > nn)
            return fail(PHYLO_B200_EINVAL, std::string(who) + ": malformed pre-order map row " + std::to_string(j));
        if (row[node] >= 0)
            return fail(PHYLO_B200_EINVAL, std::string(who) + ": node " + std::to_string(node) + " appears twice in the map");
        if (row[par] < 0)
            return fail(PHYLO_B200_EINVAL, std::string(who) + ": map row " + std::to_string(j) + " precedes its parent's row");
        if (++nkids[par] > 2)
            return fail(PHYLO_B200_EINVAL, std::string(who) + ": node " + std::to_string(par) + " has more than two children");
        row[node] = j;
    }
    return 0;  // nn - 1 distinct nodes in nn - 1 rows: every non-root node appears exactly once
}

// heights -> branch lengths on the host (O(S)), likelihood on the device, chain rule back on the host.
// autocorr == 0: generate_script.py:660-679 (strict / uncorrelated clocks);
// autocorr == 1: generate_script.py:682-708 (branch rate = mean of the rates at its two ends).
static int eval_heights_impl(phylo_b200_handle h, int autocorr, const int32_t* map, const double* heights,
                             const double* lowers, const double* rates, int nrates, const double* subst,
                             const double* freqs, const double* rs, const double* ps, int want_grad, double* logp,
                             double* g_heights, double* g_rates, double* g_subst, double* g_freqs, double* g_rs,
                             double* g_ps) {
    const char* who = autocorr ? "eval_heights_autocorr" : "eval_heights";
    if (!h || !map || !heights || !rates || !logp) return fail(PHYLO_B200_EINVAL, std::string(who) + ": NULL argument");
    if (!h->rooted) return fail(PHYLO_B200_EINVAL, std::string(who) + " needs a rooted (clock) handle");
    const int S = h->S, nn = h->nn;
    if (autocorr ? nrates != h->bcount : (nrates != 1 && nrates != h->bcount))
        return fail(PHYLO_B200_EINVAL, std::string(who) + (autocorr ? ": nrates must be 2S-2" : ": nrates must be 1 or 2S-2"));
    std::vector<double> blens(h->bcount, 0.0), gb(h->bcount, 0.0);
    if (int rc = validate_map(who, S, map)) return rc;
    // rate multiplier of the branch above `node`, and the (up to two) rate slots it reads
    const int first = map[2] - 1;  // map[2,1] of the Stan program: the first child of the root
    auto slots = [&](int j, int& s0, int& s1) {
        const int node = map[2 * j], par = map[2 * j + 1];
        if (!autocorr) { s0 = nrates == 1 ? 0 : node - 1; s1 = -1; return; }
        s0 = node - 1;
        s1 = j == 1 ? -1 : (par == nn ? first : par - 1);
    };
    auto span = [&](int j) {
        const int node = map[2 * j], par = map[2 * j + 1];
        const double lo = node > S ? heights[node - S - 1] : (lowers ? lowers[node - 1] : 0.0);
        return heights[par - S - 1] - lo;
    };
    for (int j = 1; j < nn; ++j) {
        int s0, s1;
        slots(j, s0, s1);
        const double r = s1 < 0 ? rates[s0] : 0.5 * (rates[s0] + rates[s1]);
        blens[map[2 * j] - 1] = r * span(j);
    }
    int rc = phylo_b200_eval(h, blens.data(), subst, freqs, rs, ps, want_grad, logp, gb.data(), g_subst, g_freqs, g_rs,
                             g_ps);
    if (rc || !want_grad) return rc;
    if (g_heights) std::fill(g_heights, g_heights + (S - 1), 0.0);
    if (g_rates) std::fill(g_rates, g_rates + nrates, 0.0);
    for (int j = 1; j < nn; ++j) {
        const int node = map[2 * j], par = map[2 * j + 1];
        int s0, s1;
        slots(j, s0, s1);
        const double r = s1 < 0 ? rates[s0] : 0.5 * (rates[s0] + rates[s1]);
        const double g = gb[node - 1];
        if (g_heights) {
            g_heights[par - S - 1] += r * g;
            if (node > S) g_heights[node - S - 1] -= r * g;
        }
        if (g_rates) {
            const double d = span(j) * g;
            if (s1 < 0) g_rates[s0] += d;
            else { g_rates[s0] += 0.5 * d; g_rates[s1] += 0.5 * d; }
        }
    }
    return 0;
}

int phylo_b200_eval_heights(phylo_b200_handle h, const int32_t* map, const double* heights, const double* lowers,
                            const double* rates, int nrates, const double* subst, const double* freqs,
                            const double* rs, const double* ps, int want_grad, double* logp, double* g_heights,
                            double* g_rates, double* g_subst, double* g_freqs, double* g_rs, double* g_ps) {
    return eval_heights_impl(h, 0, map, heights, lowers, rates, nrates, subst, freqs, rs, ps, want_grad, logp,
                             g_heights, g_rates, g_subst, g_freqs, g_rs, g_ps);
}

int phylo_b200_eval_heights_autocorr(phylo_b200_handle h, const int32_t* map, const double* heights,
                                     const double* lowers, const double* rates, int nrates, const double* subst,
                                     const double* freqs, const double* rs, const double* ps, int want_grad,
                                     double* logp, double* g_heights, double* g_rates, double* g_subst,
                                     double* g_freqs, double* g_rs, double* g_ps) {
    return eval_heights_impl(h, 1, map, heights, lowers, rates, nrates, subst, freqs, rs, ps, want_grad, logp,
                             g_heights, g_rates, g_subst, g_freqs, g_rs, g_ps);
}

// Batched front end with the heights -> blens step and its chain rule ON THE DEVICE: B draws, one H2D of
// [heights | rates (| hbar_extra)], clock_forward -> the three likelihood kernels -> clock_reverse, one D2H.
// ratios != 0: `heights` holds [B][S-1] = S-2 proportions then the root height; the ratio transform, its
// log-Jacobian and their reverse sweep run on the device too.
static int clock_batch_impl(phylo_b200_handle h, const char* who, int autocorr, int ratios, const int32_t* map, int B,
                            const double* heights, const double* lowers, const double* rates, int nrates,
                            const double* subst, const double* freqs, const double* rs, const double* ps,
                            const double* hbar_extra, int want_grad, double* logp, double* logjac, double* heights_out,
                            double* g_heights, double* g_props, double* g_root, double* g_rates, double* g_subst,
                            double* g_freqs, double* g_rs, double* g_ps) {
    if (!h || !map || !heights || !rates || !logp || B < 1) return fail(PHYLO_B200_EINVAL, std::string(who) + ": bad arguments");
    if (!h->rooted) return fail(PHYLO_B200_EINVAL, std::string(who) + " needs a rooted (clock) handle");
    const int S = h->S, nn = h->nn;
    if (autocorr ? nrates != h->bcount : (nrates != 1 && nrates != h->bcount))
        return fail(PHYLO_B200_EINVAL, std::string(who) + (autocorr ? ": nrates must be 2S-2" : ": nrates must be 1 or 2S-2"));
    // static part: (re)upload when the map or the sampling dates change
    const bool same = h->c_valid && h->c_map.size() == (size_t)2 * nn && std::equal(map, map + 2 * nn, h->c_map.begin()) &&
                      h->c_has_lowers == (lowers != nullptr) &&
                      (!lowers || std::equal(lowers, lowers + nn, h->c_lowers.begin()));
    if (!same) {
        if (int rc = validate_map(who, S, map)) return rc;
        h->c_valid = false;
        std::vector<int32_t> row(nn, 0), kids(2 * (size_t)(S - 1), 0), nk(S - 1, 0);
        for (int j = 1; j < nn; ++j) {
            row[map[2 * j] - 1] = j;
            const int k = map[2 * j + 1] - S - 1;
            kids[2 * k + nk[k]++] = j;
        }
        for (int k = 0; k < S - 1; ++k)
            if (nk[k] != 2) return fail(PHYLO_B200_EINVAL, std::string(who) + ": internal node without two children in the map");
        std::vector<phylo_b200_ctx*> all{h};
        all.insert(all.end(), h->peers.begin(), h->peers.end());
        for (auto* c : all) {
            CU_TRY(cudaSetDevice(c->device));
            CU_TRY(cudaStreamSynchronize(c->stream));
            CU_TRY(c->d_cmap.ensure(2 * (size_t)nn)); CU_TRY(c->d_ckids.ensure(kids.size())); CU_TRY(c->d_crow.ensure(nn));
            CU_TRY(cudaMemcpy(c->d_cmap.p, map, sizeof(int32_t) * 2 * nn, cudaMemcpyHostToDevice));
            CU_TRY(cudaMemcpy(c->d_ckids.p, kids.data(), sizeof(int32_t) * kids.size(), cudaMemcpyHostToDevice));
            CU_TRY(cudaMemcpy(c->d_crow.p, row.data(), sizeof(int32_t) * nn, cudaMemcpyHostToDevice));
            c->c_has_lowers = lowers != nullptr;
            if (lowers) {
                CU_TRY(c->d_clowers.ensure(nn));
                CU_TRY(cudaMemcpy(c->d_clowers.p, lowers, sizeof(double) * nn, cudaMemcpyHostToDevice));
            }
        }
        h->c_map.assign(map, map + 2 * nn);
        if (lowers) h->c_lowers.assign(lowers, lowers + nn); else h->c_lowers.clear();
        h->c_valid = true;
    }
    ClockJob job{};
    job.autocorr = autocorr; job.ratios = ratios; job.has_extra = hbar_extra && want_grad ? 1 : 0; job.nrates = nrates;
    job.in_ld = (S - 1) + nrates + (job.has_extra ? S - 1 : 0);
    job.hout_ld = (S - 1) + nrates + (S - 2) + 3;
    job.want_heights = heights_out != nullptr;
    // parameter blocks (branch lengths are filled in on the device) and the front end's inputs
    static thread_local std::vector<double> zeros;
    zeros.assign((size_t)B * h->bcount, 0.0);
    if (int rc = pack_batch(h, B, zeros.data(), subst, freqs, rs, ps)) return rc;
    {
        std::vector<phylo_b200_ctx*> all{h};
        all.insert(all.end(), h->peers.begin(), h->peers.end());
        for (auto* c : all) {
            CU_TRY(cudaSetDevice(c->device));
            CU_TRY(c->d_cin.ensure((size_t)B * job.in_ld));
            CU_TRY(c->d_hwork.ensure((size_t)B * 2 * (S - 1)));
            CU_TRY(c->d_hout.ensure((size_t)B * job.hout_ld));
        }
        CU_TRY(cudaSetDevice(h->device));
        CU_TRY(h->h_cin.ensure((size_t)B * job.in_ld));
        CU_TRY(h->h_hout.ensure((size_t)B * job.hout_ld));
        if (job.want_heights) CU_TRY(h->h_hts.ensure((size_t)B * (S - 1)));
    }
    for (int b = 0; b < B; ++b) {
        double* dst = h->h_cin.p + (size_t)b * job.in_ld;
        std::memcpy(dst, heights + (size_t)b * (S - 1), sizeof(double) * (S - 1));
        std::memcpy(dst + (S - 1), rates + (size_t)b * nrates, sizeof(double) * nrates);
        if (job.has_extra) std::memcpy(dst + (S - 1) + nrates, hbar_extra + (size_t)b * (S - 1), sizeof(double) * (S - 1));
    }
    {   // branch lengths are computed on the device: bound them by (largest rate) x (largest height) of the batch
        double hmax = 0.0, rmax = 0.0;
        for (size_t i = 0; i < (size_t)B * (S - 1); ++i) hmax = std::max(hmax, std::fabs(heights[i]));
        for (size_t i = 0; i < (size_t)B * nrates; ++i) rmax = std::max(rmax, std::fabs(rates[i]));
        const double tb = hmax * rmax * h->tau_unit + h->tau_cond;
        h->tau_bound = std::isfinite(tb) ? tb : 1e300;
        for (auto* p : h->peers) p->tau_bound = h->tau_bound;
    }
    const bool grad = want_grad != 0;
    if (int rc = run_prepare_all(h, B, grad)) return rc;
    if (int rc = eval_enqueue(h, B, grad, &job)) return rc;
    CU_TRY(cudaStreamSynchronize(h->stream));
    if (heights_out) std::memcpy(heights_out, h->h_hts.p, sizeof(double) * B * (S - 1));
    bool finite = true;
    for (int d = 0; d < B; ++d) {
        const double* o = h->h_out.p + (size_t)d * h->nout;
        const double* ho = h->h_hout.p + (size_t)d * job.hout_ld;
        if (ho[job.hout_ld - 1] != 0.0)
            return fail(PHYLO_B200_EDOMAIN, "draw " + std::to_string(d) + ": a branch length is negative or not finite");
        logp[d] = o[0];
        finite = finite && std::isfinite(o[0]);
        if (logjac) logjac[d] = ratios ? ho[job.hout_ld - 2] : 0.0;
        if (!grad) continue;
        if (g_heights) std::memcpy(g_heights + (size_t)d * (S - 1), ho, sizeof(double) * (S - 1));
        if (g_rates) std::memcpy(g_rates + (size_t)d * nrates, ho + (S - 1), sizeof(double) * nrates);
        if (ratios && g_props && S > 2) std::memcpy(g_props + (size_t)d * (S - 2), ho + (S - 1) + nrates, sizeof(double) * (S - 2));
        if (ratios && g_root) g_root[d] = ho[(S - 1) + nrates + (S - 2)];
        if (g_subst && h->nsubst) std::memcpy(g_subst + (size_t)d * h->nsubst, o + h->off_subst, sizeof(double) * h->nsubst);
        if (g_freqs) std::memcpy(g_freqs + (size_t)d * 4, o + h->off_freqs, sizeof(double) * 4);
        if (g_rs) std::memcpy(g_rs + (size_t)d * h->C, o + h->off_rs, sizeof(double) * h->C);
        if (g_ps) std::memcpy(g_ps + (size_t)d * h->C, o + h->off_ps, sizeof(double) * h->C);
    }
    if (!finite) return fail(PHYLO_B200_EDOMAIN, "log-likelihood is not finite (impossible pattern or underflow)");
    return 0;
}

int phylo_b200_eval_heights_batch(phylo_b200_handle h, int autocorr, const int32_t* map, int B, const double* heights,
                                  const double* lowers, const double* rates, int nrates, const double* subst,
                                  const double* freqs, const double* rs, const double* ps, int want_grad, double* logp,
                                  double* g_heights, double* g_rates, double* g_subst, double* g_freqs, double* g_rs,
                                  double* g_ps) {
    return clock_batch_impl(h, "eval_heights_batch", autocorr != 0, 0, map, B, heights, lowers, rates, nrates, subst, freqs,
                            rs, ps, nullptr, want_grad, logp, nullptr, nullptr, g_heights, nullptr, nullptr, g_rates,
                            g_subst, g_freqs, g_rs, g_ps);
}

int phylo_b200_eval_ratios_batch(phylo_b200_handle h, int autocorr, const int32_t* map, int B, const double* props,
                                 const double* root_height, const double* lowers, const double* rates, int nrates,
                                 const double* subst, const double* freqs, const double* rs, const double* ps,
                                 const double* hbar_extra, int want_grad, double* logp, double* logjac, double* heights,
                                 double* g_props, double* g_root, double* g_rates, double* g_subst, double* g_freqs,
                                 double* g_rs, double* g_ps) {
    if (!h || !props || !root_height || B < 1) return fail(PHYLO_B200_EINVAL, "eval_ratios_batch: bad arguments");
    const int S = h->S;
    static thread_local std::vector<double> pr;  // [B][S-1] = proportions then the root height
    pr.resize((size_t)B * (S - 1));
    for (int b = 0; b < B; ++b) {
        if (S > 2) std::memcpy(pr.data() + (size_t)b * (S - 1), props + (size_t)b * (S - 2), sizeof(double) * (S - 2));
        pr[(size_t)b * (S - 1) + (S - 2)] = root_height[b];
    }
    return clock_batch_impl(h, "eval_ratios_batch", autocorr != 0, 1, map, B, pr.data(), lowers, rates, nrates, subst, freqs,
                            rs, ps, hbar_extra, want_grad, logp, logjac, heights, nullptr, g_props, g_root, g_rates,
                            g_subst, g_freqs, g_rs, g_ps);
}

// Host-only: the ratio transform of the node heights (generate_script.py:711-735) and its log-Jacobian
// (:738-752) for B draws, and the reverse sweep through it.  O(B S); no GPU.
// props [B][S-2] are consumed in pre-order of the internal non-root nodes, as the Stan loop does.
int phylo_b200_ratios_forward(int S, const int32_t* map, const double* lowers, int B, const double* props,
                              const double* root_height, double* heights, double* logjac) {
    if (S < 2 || !map || B < 1 || !props || !root_height || !heights) return fail(PHYLO_B200_EINVAL, "ratios_forward: bad arguments");
    const int nn = 2 * S - 1;
    if (int rc = validate_map("ratios_forward", S, map)) return rc;
    for (int b = 0; b < B; ++b) {
        double* h = heights + (size_t)b * (S - 1);
        const double* p = props + (size_t)b * (S - 2);
        h[map[0] - S - 1] = root_height[b];
        double lj = 0.0;
        int k = 0;
        for (int j = 1; j < nn; ++j) {
            const int node = map[2 * j], par = map[2 * j + 1];
            if (node <= S) continue;
            const double lo = lowers ? lowers[node - 1] : 0.0, span = h[par - S - 1] - lo;
            h[node - S - 1] = lo + span * p[k++];
            lj += std::log(span);
        }
        if (logjac) logjac[b] = lj;
    }
    return 0;
}

// hbar [B][S-1]: adjoint of the heights coming from everything downstream (likelihood, tree prior); the
// log-Jacobian's own contribution is added here.  Outputs d/dprops [B][S-2], d/droot_height [B].
int phylo_b200_ratios_reverse(int S, const int32_t* map, const double* lowers, int B, const double* props,
                              const double* heights, double* hbar, double* g_props, double* g_root) {
    if (S < 2 || !map || B < 1 || !props || !heights || !hbar || !g_props || !g_root)
        return fail(PHYLO_B200_EINVAL, "ratios_reverse: bad arguments");
    const int nn = 2 * S - 1;
    if (int rc = validate_map("ratios_reverse", S, map)) return rc;
    std::vector<int> slot(nn, -1);
    int k = 0;
    for (int j = 1; j < nn; ++j)
        if (map[2 * j] > S) slot[j] = k++;
    for (int b = 0; b < B; ++b) {
        const double* h = heights + (size_t)b * (S - 1);
        const double* p = props + (size_t)b * (S - 2);
        double* hb = hbar + (size_t)b * (S - 1);
        double* gp = g_props + (size_t)b * (S - 2);
        for (int j = nn - 1; j >= 1; --j) {  // reverse pre-order: children before parents
            if (slot[j] < 0) continue;
            const int node = map[2 * j], par = map[2 * j + 1];
            const double lo = lowers ? lowers[node - 1] : 0.0, span = h[par - S - 1] - lo;
            const double nb = hb[node - S - 1];
            gp[slot[j]] = nb * span;
            hb[par - S - 1] += nb * p[slot[j]] + 1.0 / span;
        }
        g_root[b] = hb[map[0] - S - 1];
    }
    return 0;
}

// Host-only planning hook (no GPU): exercised by the CPU test-suite.
// post/pre receive S-1 rows of 8 / 12 int32 (the PostStep / PreStep fields); depth[2] = {post, pre}.
int phylo_b200_plan(int S, const int32_t* peel, int32_t* post, int32_t* pre, int32_t* depth) {
    Plan plan;
    std::string err;
    if (!build_plan(S, peel, plan, err)) return fail(PHYLO_B200_EINVAL, err);
    if (post) std::memcpy(post, plan.post.data(), plan.post.size() * sizeof(PostStep));
    if (pre) std::memcpy(pre, plan.pre.data(), plan.pre.size() * sizeof(PreStep));
    if (depth) { depth[0] = plan.depth_post; depth[1] = plan.depth_pre; }
    return 0;
}

// Host-only hook (no GPU): the table nodes of a tree and the second plan, whose post-order treats them as leaves.
// node_tab [2S-1]; post / pre as in phylo_b200_plan (post holds info[0] rows); info[4] = {post-order steps, table
// nodes, table entries per (draw, category), stack depth}.  Returns PHYLO_B200_EINVAL when the tree has no such plan
// (the root itself is a table node).
int phylo_b200_plan_tables(int S, const int32_t* peel, int max_tips, int32_t* node_tab, int32_t* post, int32_t* pre,
                           int32_t* info) {
    Plan plan, planB;
    std::string err;
    if (!build_plan(S, peel, plan, err)) return fail(PHYLO_B200_EINVAL, err);
    std::vector<int32_t> nt, rec, off;
    const int entries = find_table_nodes(plan, S, max_tips == 2 ? 2 : 3, nt, rec, off);
    if (off.empty() || nt[plan.root] >= 0) return fail(PHYLO_B200_EINVAL, "plan_tables: the root is a table node");
    std::vector<char> leaf(nt.size(), 0);
    for (size_t n = 0; n < nt.size(); ++n) leaf[n] = nt[n] >= 0;
    if (!build_plan(S, peel, planB, err, &leaf)) return fail(PHYLO_B200_EINVAL, err);
    if (node_tab) std::memcpy(node_tab, nt.data(), nt.size() * sizeof(int32_t));
    if (post) std::memcpy(post, planB.post.data(), planB.post.size() * sizeof(PostStep));
    if (pre) std::memcpy(pre, planB.pre.data(), planB.pre.size() * sizeof(PreStep));
    if (info) { info[0] = (int32_t)planB.post.size(); info[1] = (int32_t)off.size(); info[2] = entries; info[3] = planB.depth(); }
    return 0;
}

// Host-only model algebra hook (no GPU): Q, lambda, m1, m2 and X_theta of one draw.
// out = [pi 4 | lam 4 | m1 16 | m2 16 | Q 16 | X ntheta*16]; returns ntheta or < 0.
int phylo_b200_derive(int model, int flags, const double* subst, const double* freqs, double* out) {
    Derived dv;
    if (!derive(model, !(flags & PHYLO_B200_NO_NORMQ), subst, freqs, dv))
        return fail(PHYLO_B200_EDOMAIN, "substitution parameters out of domain");
    if (out) {
        std::memcpy(out, dv.pi, sizeof dv.pi);
        std::memcpy(out + 4, dv.lam, sizeof dv.lam);
        std::memcpy(out + 8, dv.m1, sizeof dv.m1);
        std::memcpy(out + 24, dv.m2, sizeof dv.m2);
        std::memcpy(out + 40, dv.Q, sizeof dv.Q);
        for (int k = 0; k < dv.ntheta; ++k) std::memcpy(out + 56 + 16 * k, dv.X[k], sizeof dv.X[k]);
    }
    return dv.ntheta;
}

}  // extern "C"
