// kernels.cuh -- device side of libphylo_b200: hand-written CUDA for sm_100a.
//
// Three kernels per evaluation (batch of B parameter draws):
//   stream_kernel    P(t_b r_c) = m1 diag(exp(lambda t_b r_c)) m2 for both children of every step
//                    (phylostan/generate_script.py:824-829, 880-885; JC69 closed form :765-766),
//                    written as per-(draw, category) INSTRUCTION STREAMS in traversal order:
//                    one 384-byte record [step descriptor with ready-made offsets | P_a | P_b] per
//                    internal node, one stream for the post-order and one for the pre-order sweep
//   sweep_kernel     per (draw, pattern tile): depth-first post-order partials
//                    (eigen/eigen.j2:122-141, generate_script.py:998-1005), root likelihood with
//                    per-(pattern,category) rescaling (generate_script.py:1006-1010), then the
//                    depth-first pre-order sweep (eigen/eigen.j2:144-157) accumulating the 4x4
//                    branch statistics G_b,c = sum_l (w_l/L_l) A_b p_b^T that carry every gradient
//   contract_kernel  G -> d/dblens (eigen/eigen.j2:163-166 without its times[i] factor),
//                    d/drs, d/d(rates|kappa), d/dfreqs
//
// Memory plan: one thread = one (pattern, category) (x K patterns).  The most recent vector of a
// tile stays in registers, the ones waiting for their sibling in a shared-memory stack
// [slot][k][half][thread] of double2 (bank-conflict free).  Every
// warp streams its category's records into a private shared-memory ring with cp.async, two chunks
// ahead, so the step -> matrix -> operand chain never waits on global memory; operand lines of
// step i+2 are prefetched into L2.  HBM traffic is one coalesced double2 write per internal-node
// partial in the post-order (the CTA's scratch rows) and one read of it in the pre-order, plus
// 1-byte tip codes; q never leaves the SM.
//
// Round 2, on top of that (kernels.cu has the details): the MESSAGE statistic (the rows hold P_c p_c, the pre-order
// never multiplies by P again, the contraction undoes the factor analytically), the statistics' warp sums through
// TENSOR MEMORY used as a transpose unit (tcgen05.st.32x32b in, tcgen05.ld.16x256b out), message TABLES for the
// nodes with two or three tips below them (cherry_table_kernel; both sweeps look their messages up by a combined
// tip code instead of computing, storing and re-reading them), and an opt-in variant that keeps the whole stack in
// tensor memory (sweep_tm_kernel).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "plan.hpp"

namespace phylo {

constexpr int kRecBytes = 384;   // fp64 record: [desc 64 B | P_a slot 160 B | P_b slot 160 B]
constexpr int kRecBytesF32 = 224;  // fp32 mode: [desc 64 B | P_a slot 80 B | P_b slot 80 B]
constexpr int kRecChunk = 2;     // records per cp.async group
#ifndef PHYLO_RECBUFS
#define PHYLO_RECBUFS 3
#endif
constexpr int kRecBufs = PHYLO_RECBUFS;  // ring depth in chunks (kRecBufs - 1 chunks are in flight ahead of the one being read)
constexpr int kTmOpSlots = 4;    // TMEM-stack sweep: operand slots per warp (both children of steps i and i+1)

// offsets (in doubles) inside one draw's parameter block
struct ParamLayout {
    int nn, C, ntheta, nsubst;
    int off_t;    // [nn]  branch length above node k (0 where the node carries no branch)
    int off_rs;   // [C]
    int off_ps;   // [C]
    int off_pi;   // [4]
    int off_lam;  // [4]
    int off_m1;   // [16]
    int off_m2;   // [16]
    int off_Q;    // [16]
    int off_X;    // [ntheta][16]
    int stride;
};

// Record descriptors (first 64 bytes of a record).  Offsets are ready to add to a base pointer:
// stack / scratch offsets in double2 elements (slot or row times SS = K*2*NT), tip rows in bytes.
struct alignas(16) PostRec {
    long long tip_a, tip_b;   // byte offset of the child's tip-code row (node * Lpad)
    int32_t off_a, off_b;     // shared-memory stack offset of the child's partial (when it sits in a slot)
    int32_t off_spill;        // stack offset that receives the previous TOS first, or -1
    int32_t flags;            // 1: a is a tip, 2: b is a tip, 4: a is the TOS, 8: b is the TOS, 16: a is parked in its scratch
                              //   row, 32: the previous TOS is parked above the capped stack
    int32_t row_a, row_b;     // scratch offsets of the children's own rows, -1 for tips (message-statistic sweep:
    int32_t pad0, pad1;       //   the MESSAGE P_c p_c of an internal child is stored there, at its parent's step)
    long long ctab_a, ctab_b; // post-order message tables: device address of a leafified child's table for this record's
                              //   (draw, category), or 0; such a child is tip-like (flags 1 / 2) with tip_* = its code row
};
struct alignas(16) PreRec {
    long long tip_a, tip_b;   // byte offset of the child's tip-code row
    int32_t row_a, row_b;     // scratch offset of the child's partial, or -1 (tip)
    int32_t dl_n;             // byte offset of this node's rescale exponents
    int32_t off_n;            // stack offset of q(node), or -1 when it is the TOS
    int32_t off_b;            // stack offset receiving q(b), or -1 (b is a tip)
    int32_t g_a, g_b;         // offsets (doubles) of the children's 4x4 statistics blocks
    int32_t flags;            // 1: a is internal -> q(a) becomes the TOS
    long long ctab_a, ctab_b; // cherry-table runs: device address of the child's 25 x 4 message table for this record's (draw,
                              //   category), or 0; such a child has row = -1 and tip_* = its combined-code row
};
static_assert(sizeof(PostRec) == 64 && sizeof(PreRec) == 64, "record descriptors must fit the 64-byte header");

struct StreamArgs {
    const double* params;     // [B][stride]
    const PostStep* post;     // [S-1]
    const PreStep* pre;       // [S-1]
    unsigned char* spost;     // [B][C][S-1][kRecBytes]
    unsigned char* spre;      // [B][C][S-1][kRecBytes]
    ParamLayout lay;
    int nsteps, bcount, jc_closed, B;   // nsteps: pre-order steps (S - 1)
    int npost;                // post-order steps: S - 1, or fewer when the plan leafifies the table nodes (post_tables)
    int post_tables;          // the post-order takes the messages of table nodes from their tables too
    int Lpad, SS, KNT;        // tile geometry baked into the descriptors
    int S, tips_simple;       // tips_simple: tip children get the column-major 4x5 matrix layout
    // Capped shared-memory stack (gradient runs only): stack positions >= slots (the top of a deep
    // stack, reached rarely and briefly) live in the CTA's HBM scratch instead -- a parked partial is
    // there anyway (every partial is written to its scratch row), a parked q(node) takes over the row
    // of p(node), which nobody reads any more by then.
    const int* node_row;      // [2S-1] node id -> its own post-order step (= scratch row), -1 for tips
    int slots;
    int msg;                  // message-statistic sweep: tip children of the PRE-order stream get the column-major layout too
    // cherry tables (message-statistic runs): the message of a cherry (both children tips) depends on the pattern only
    // through the 5 x 5 code pairs of its tips, so it is looked up in a per-(draw, category, cherry) table instead of
    // being stored by the post-order and re-read by the pre-order
    const int32_t* node_cherry;  // [2S-1] node -> table-node index or -1 (NULL: no tables in this run).  Table nodes are
                                 // the internal nodes with two or three tips below them (cherries and pitchforks)
    long long ctips_off;         // byte offset of the combined-code rows [ntab][Lpad] from the tip-code rows
    const double* ctab;          // [B][C][tab_entries][4]: 25 (cherry) or 125 (pitchfork) messages per table node
    const int32_t* tab_off;      // [ntab] first entry of every table node inside a (draw, category) block
    int tab_entries;
    int slot_stride;          // offset unit of an ON-CHIP stack slot: SS (shared-memory stack, in 16-byte vectors) or
                              // 8 K (tensor-memory stack, in 32-bit columns)
};

struct SweepArgs {
    const uint8_t* tips;      // [S][Lpad] 4-bit state masks, or column indices 0..4 when every tip is simple
    const uint8_t* tips_post; // TR kernels: the same codes as [tile][slot][32 K] in post-order / pre-order consumption
    const uint8_t* tips_pre;  //   order (slot = s-th tip child met by the sweep, child a before child b)
    const double* weights;    // [Lpad]
    const double* params;     // [B][stride]
    const unsigned char* spost;
    const unsigned char* spre;
    void* scratch;            // [grid][S-1][K][NT] 4-state entries: double2 x 2 (fp64) or float4 (fp32)
    uint8_t* dscr;            // [grid][S-1][K][NT]   rescale exponents (units of 2^64)
    double* G;                // [B][nn][C][16]
    double* out;              // [B][nout]
    ParamLayout lay;
    long long scratch_stride, dscr_stride;   // per CTA, in 16-byte vectors / bytes
    int S, nsteps, Lpad, ntiles, nitems, C, nn, nout, D;
    int npost;                // post-order steps (<= nsteps, see StreamArgs)
    int stack_bytes;          // size of the stack region at the head of dynamic shared memory
    int off_out_freqs, off_out_ps;
};

// message tables: one thread per (draw, category, table node)
constexpr int kTabRec = 8;    // ints per table node: node, ntips (2 | 3), t0, t1, t2 (-1), inner cherry node (-1),
                              //   shape (pitchfork: 0 = ((t0, t1), t2), 1 = (t0, (t1, t2))), first table entry
struct CherryArgs {
    const double* params;
    const int32_t* cherries;  // [ntab][kTabRec]; tips 0-based, in the post-order steps' child order (a before b)
    double* ctab;
    ParamLayout lay;
    int B, C, ncherry, tab_entries, bcount, jc_closed;
    int norescale;            // post-order tables: table nodes are never rescaled (neither sweep visits their partials)
};
void launch_cherry_tables(const CherryArgs& a, cudaStream_t stream);
// combined codes of the tips below every table node, 5 x + y or 25 x + 5 y + z, [ntab][Lpad] (tip codes must be column
// indices 0..4)
void launch_cherry_codes(const uint8_t* d_tips, uint8_t* d_ctips, const int32_t* d_cherries, int ncherry, int Lpad,
                         cudaStream_t stream);

struct ContractArgs {
    const unsigned char* spost;
    const int32_t* node_pos;  // [nn] 2 * post step + child slot of every non-root node
    const double* params;
    const double* G;
    double* out;
    ParamLayout lay;
    int bcount, C, nn, nout, nsubst, nsteps, S, tips_simple;
    int jc_scalar;            // entry 0 of a G block holds <G, Q P> / mu (JC69 scalar-statistic sweep)
    int msg;                  // the G blocks hold the message statistic G~ = sum A mu^T = G P^T
    int off_out_subst, off_out_freqs, off_out_rs;
};

// prec: 64 (product path) or 32 (optional fp32-with-scaling mode); K in {1,2,4}; nthreads <= 512
void launch_stream(const StreamArgs& a, int prec, cudaStream_t stream);
// deep: the stream parks stack entries in the HBM scratch (gradient runs only)
// jc: scalar-statistic gradient kernel of JC69 handles (fp64, gradient, not deep; ContractArgs::jc_scalar must agree)
// msg: message-statistic gradient kernel (fp64, simple tips, 128-thread CTAs; StreamArgs::msg and
// ContractArgs::msg must agree) -- see sweep_msg_available
// ch (with msg): the instantiation that takes the messages of cherries from tables (sweep_cherry_available)
cudaError_t launch_sweep(const SweepArgs& a, int prec, bool tips, int K, bool grad, bool deep, int grid, int nthreads,
                         size_t smem, cudaStream_t stream, bool jc = false, bool msg = false, bool ch = false);
cudaError_t sweep_occupancy(int prec, bool tips, int K, bool grad, bool deep, int nthreads, size_t smem,
                            int* blocks_per_sm, bool jc = false, bool msg = false, bool ch = false);
bool sweep_cherry_available(int K);
bool sweep_msg_available(int prec, bool tips, bool grad, bool deep, int nthreads, bool jc);
// Gradient sweep with the stack in tensor memory (fp64, K = 4, 128-thread CTAs): `ctas` = 2 or 3 resident CTAs per
// SM, sweep_tm_slots stack slots on chip (positions beyond them are parked in the scratch like the `deep` variant)
bool sweep_tm_available(int prec, int K, int nthreads, bool grad, bool jc);
int sweep_tm_slots(int K, int ctas);
size_t sweep_tm_smem_bytes(int K);
cudaError_t launch_sweep_tm(const SweepArgs& a, bool tips, int K, int ctas, int grid, cudaStream_t stream, bool msg = false);
cudaError_t sweep_tm_prepare(bool tips, int K, int ctas, int* blocks_per_sm, bool msg = false);
void launch_contract(const ContractArgs& a, int prec, int B, cudaStream_t stream);
// device-resident tip masks [S][L] / weights [L] (NULL: ones) -> padded rows; d_flags[2]: {some cell is not
// simple, some weight is not finite}
void launch_tips_prepare(const uint8_t* d_src, uint8_t* d_dst, int S, int L, int Lpad, const double* d_w, double* d_wdst,
                         int* d_flags, cudaStream_t stream);
void launch_tips_index(uint8_t* d_tips, int S, int Lpad, cudaStream_t stream);  // masks -> column indices

// multi-device handles: out[i] += sum_p src[p][i]; the sources are the other shards' result blocks (peer
// memory read over NVLink, or staged copies of them on this device)
constexpr int kMaxPeers = 15;
struct PeerRows {
    const double* src[kMaxPeers];
    int n;
};
void launch_peer_sum(double* out, const PeerRows& peers, size_t count, cudaStream_t stream);

// Clock-tree front end on the device (phylo_b200_eval_heights_batch / _ratios_batch): node heights (or their
// ratio parametrisation) and clock rates of B draws -> branch lengths written straight into the packed
// parameter blocks (clock_forward), and the chain rule back from d/dblens (clock_reverse).
// generate_script.py:660-679 (heights_to_blens), :682-708 (autocorrelated), :711-752 (transform + log-Jacobian).
struct ClockArgs {
    const int32_t* map;       // [nn][2] (node, parent), pre-order, 1-based; row 0 = root
    const int32_t* kids;      // [S-1][2] pre-order rows of the two children of internal node S+1+k
    const int32_t* row_of;    // [nn] pre-order row of node k+1
    const double* lowers;     // [nn] or NULL
    const double* in;         // [B][in_ld]: heights S-1 (or props S-2, root 1) | rates nrates | hbar_extra S-1
    double* params;           // [B][stride] packed parameter blocks (off_t receives the branch lengths)
    const double* out;        // [B][nout] result rows of the likelihood (d/dblens at 1 + node - 1)
    double* hwork;            // [B][2(S-1)] heights and their adjoint
    double* hout;             // [B][hout_ld]: g_heights S-1 | g_rates nrates | g_props S-2 | g_root | logjac | status
    int S, nn, nrates, autocorr, ratios, has_extra, B;
    int in_ld, hout_ld, off_t, stride, nout;
};
void launch_clock_forward(const ClockArgs& a, cudaStream_t stream);
void launch_clock_reverse(const ClockArgs& a, cudaStream_t stream);

// jc: the scalar-statistic kernel needs no reduction rows
size_t sweep_smem_bytes(int D, int K, int nthreads, int prec, bool jc = false, bool tips = false, bool grad = true);
bool sweep_uses_tipring(bool tips, int nthreads, bool grad);  // value-only kernels, 128-thread CTAs, simple-tip handles
// [S][Lpad] code rows -> [ntiles][S][T] with slot s = tip d_order[s] (0-based), T = patterns per tile
void launch_tips_reorder(const uint8_t* d_tips, uint8_t* d_dst, const int32_t* d_order, int S, int Lpad, int T, int ntiles,
                         cudaStream_t stream);
size_t sweep_stack_bytes(int D, int K, int nthreads, int prec);  // first region of the above
int record_bytes(int prec);
int sweep_max_threads(int K);  // largest CTA the K-variant is compiled for

}  // namespace phylo
