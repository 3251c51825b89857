// lpx_knapsack.cu — best-first Branch & Bound for the 0/1 knapsack:
// BranchAndBoundKnapsack.Solve, R/Models/BranchAndBoundKnapsack.cs:58-407.
//
// The hot function is ComputeRelaxation (:431-491): a greedy fractional bound whose floating
// point result depends on the ORDER of its additions (fixed-to-1 items in original index order,
// then undecided items in ratio-rank order).  One warp evaluates one node: the lanes copy the
// parent's assignment vector into the child's pool slot (16 bytes per load) and stage it in shared
// memory; then either lane 0 performs the two sums in exactly the reference's order, or — when the
// data are integers whose partial sums are exact, so that the order cannot matter — the whole warp
// forms them by reduction / scan (segment-ranked when the weights are non-negative).
//
// GPU node pool: assignment vectors live in a device pool (n bytes per node).  Each round the
// host speculates on the top-K heap nodes and asks the device for their subtrees down to depth D
// in ONE cooperative launch (knap_round_kernel: a grid barrier between levels, a level reads its
// parents' results on the device; inputs and results travel through pinned host memory directly).
// Relaxations are pure functions of the node, so the host then COMMITS in the reference's order:
// its own max-heap with the reference's sift rules (:494-547), left child before right child,
// incumbent updates exactly where the C# code makes them.  Speculative results that were not
// consumed stay cached under their heap node.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>
#include <queue>
#include <vector>

#include <cooperative_groups.h>

#include "lpx_common.cuh"
#include "lpx_knap.hpp"
#include "lpx_runtime.hpp"
#include "lpx_stream.hpp"

namespace lpx {

#define KN_EPS 1e-9  // BranchAndBoundKnapsack.EPS (:56)

enum { KF_INFEASIBLE = 1, KF_ALLINT = 2, KF_SKIPPED = 4, KF_EARLY = 8 };
static const int KN_MAX_CHUNKS = 64;
static const int KN_CHUNK_SLOTS = 1 << 15;
static const int KN_MAX_LEVELS = 24;

struct KnIn {
    int inst;
    int parent_slot;     // pool slot of the parent's assignment (-1: take it from parent_pending)
    int parent_pending;  // index in this round's arrays of a parent evaluated in an earlier level
    int var;             // original index to fix (-1: none = root; -2: from the parent's fractional item)
    int side;            // value to fix it to
    int out_slot;
    int parent_out_slot; // out_slot of parent_pending (filled by run_plan; saves a dependent read)
    int foreign;         // sharded tree: 1 = another rank evaluates this relaxation
    double best;         // the instance's incumbent when the round was planned (filled by run_plan)
};

struct KnOut {
    double bound, weight, frac;
    int frac_rank, break_rank;
    int flags, var;
};

struct KnParams {
    const double* w_s;   // [inst][n] weight by ratio rank
    const double* p_s;   // [inst][n] profit by ratio rank
    const int* orig_s;   // [inst][n] original index by ratio rank
    const double* w_o;   // [inst][n] weight by original index
    const double* p_o;   // [inst][n]
    const double* cap;   // [inst]
    const int* exact;    // [inst] 1: all weights/profits are integers with exact sums (order-free adds)
    int n;
    signed char* chunks[KN_MAX_CHUNKS];
    const KnIn* in;      // this round's evaluations, sorted by level; read straight from pinned host memory
    KnOut* out;          // device copy of the results (children look their parent up here)
    KnOut* out_host;     // the same results written through to pinned host memory
    int nlev;
    int pass;            // sharded tree: 0 = evaluate own entries, 1 = materialise the foreign ones (after the merge)
    int lev_first[KN_MAX_LEVELS], lev_count[KN_MAX_LEVELS];  // evaluations [first, first+count) form a level
};

__device__ __forceinline__ signed char* kn_slot(const KnParams& P, int slot) {
    return P.chunks[slot / KN_CHUNK_SLOTS] + (size_t)(slot % KN_CHUNK_SLOTS) * P.n;
}

__device__ __forceinline__ void kn_store(const KnParams& P, int e, const KnOut& o) {
    P.out[e] = o;
    P.out_host[e] = o;
}

// One ComputeRelaxation (BranchAndBoundKnapsack.cs:425-500) by one warp; sa = its n-byte shared scratch.
__device__ void knap_eval_one(const KnParams& P, int e, signed char* sa) {
    const int lane = threadIdx.x & 31;
    const KnIn in = P.in[e];
    const int n = P.n;
    if (P.pass == 0 && in.foreign) {  // another rank's relaxation: all-zero words for the merge
        if (lane == 0) {
            KnOut zero;
            memset(&zero, 0, sizeof zero);
            P.out[e] = zero;
        }
        return;
    }
    if (P.pass == 1) {
        // after the merge: give the foreign nodes their assignment vectors, so that every rank's pool
        // is complete, and pass their results on to the host
        if (!in.foreign) return;
        const KnOut mine = P.out[e];
        if (lane == 0) P.out_host[e] = mine;
        if (mine.flags & KF_SKIPPED) return;  // its owner wrote no assignment either
        int pslot = in.parent_slot, pvar = in.var;
        if (in.parent_pending >= 0) {
            pslot = in.parent_out_slot;
            pvar = P.orig_s[(size_t)in.inst * n + P.out[in.parent_pending].frac_rank];
        }
        const signed char* src = kn_slot(P, pslot);
        signed char* dst = kn_slot(P, in.out_slot);
        for (int i = lane; i < n; i += 32) {
            signed char a = src[i];
            if (i == pvar) a = (signed char)in.side;
            dst[i] = a;
        }
        return;
    }
    KnOut o;
    o.bound = o.weight = o.frac = 0.0;
    o.frac_rank = -1;
    o.break_rank = n;
    o.flags = 0;
    o.var = in.var;

    int parent_slot = in.parent_slot, var = in.var;
    const double best = in.best;
    if (in.parent_pending >= 0) {
        const KnOut po = P.out[in.parent_pending];
        // children are only wanted below a node the reference could push: feasible, fractional,
        // bound above the incumbent (the incumbent only grows, so a skip stays valid)
        const bool expandable = !(po.flags & (KF_INFEASIBLE | KF_ALLINT | KF_SKIPPED)) && po.frac_rank >= 0 &&
                                po.bound > best + KN_EPS;
        if (!expandable) {
            o.flags = KF_SKIPPED;
            if (lane == 0) kn_store(P, e, o);
            return;
        }
        parent_slot = in.parent_out_slot;
        var = P.orig_s[(size_t)in.inst * n + po.frac_rank];
        o.var = var;
    }
    const signed char* src = kn_slot(P, parent_slot);
    signed char* dst = kn_slot(P, in.out_slot);
    if ((n & 15) == 0) {
        // 16 bytes per load: the parent's assignment is the longest dependent fetch of a relaxation
        const int4* s4 = reinterpret_cast<const int4*>(src);
        int4* d4 = reinterpret_cast<int4*>(dst);
        int4* a4 = reinterpret_cast<int4*>(sa);
        for (int q = lane; q < (n >> 4); q += 32) {
            const int4 v = s4[q];
            a4[q] = v;
            d4[q] = v;
        }
        __syncwarp();
        if (lane == 0 && var >= 0) {
            sa[var] = (signed char)in.side;
            dst[var] = (signed char)in.side;
        }
    } else {
        for (int i = lane; i < n; i += 32) {
            signed char a = src[i];
            if (i == var) a = (signed char)in.side;
            sa[i] = a;
            dst[i] = a;
        }
    }
    __syncwarp();
    const double* w_o = P.w_o + (size_t)in.inst * n;
    const double* p_o = P.p_o + (size_t)in.inst * n;
    const double* w_s = P.w_s + (size_t)in.inst * n;
    const double* p_s = P.p_s + (size_t)in.inst * n;
    const int* orig_s = P.orig_s + (size_t)in.inst * n;
    const double capacity = P.cap[in.inst];
    const double limit = __dadd_rn(capacity, KN_EPS);

    if (P.exact[in.inst]) {
        // Integer data whose every partial sum is exactly representable: floating point addition
        // is then exact and order-free, so the two sums can be formed by the whole warp (reduction
        // and prefix scan) with results identical to the reference's sequential order.  Data that
        // do not qualify take the sequential path below.
        double w1 = 0.0, p1 = 0.0;
        for (int i = lane; i < n; i += 32)
            if (sa[i] == 1) {
                w1 += w_o[i];
                p1 += p_o[i];
            }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            w1 += __shfl_xor_sync(0xffffffffu, w1, off);
            p1 += __shfl_xor_sync(0xffffffffu, p1, off);
        }
        double weight = w1, profit = p1;
        if (weight > limit) {
            o.bound = profit;
            o.weight = weight;
            o.flags = KF_INFEASIBLE | KF_EARLY;
            if (lane == 0) kn_store(P, e, o);
            return;
        }
        int s_break = n;
        int scan_lo = 0, scan_hi = n;
        bool nothing_breaks = false;
        if (P.exact[in.inst] == 2) {
            // Non-negative weights on top: the running weight is monotone, so the first item that does
            // not fit lies in the first SEGMENT whose inclusive total does not fit.  Every lane sums one
            // contiguous segment (independent loads, no cross-lane dependency), one warp scan ranks the
            // segments, and only the breaking segment (<= ceil(n/32) items) is walked item by item below
            // — 1 + 2 scans instead of n/32.
            const int seg = (n + 31) / 32;
            const int s0 = min(n, lane * seg), s1 = min(n, s0 + seg);
            double lw = 0.0, lp = 0.0;
#pragma unroll 4
            for (int s = s0; s < s1; s++)
                if (sa[orig_s[s]] < 0) {
                    lw += w_s[s];
                    lp += p_s[s];
                }
            double cw = lw, cp = lp;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                const double ow = __shfl_up_sync(0xffffffffu, cw, off);
                const double op = __shfl_up_sync(0xffffffffu, cp, off);
                if (lane >= off) {
                    cw += ow;
                    cp += op;
                }
            }
            const unsigned over = __ballot_sync(0xffffffffu, !((weight + cw) <= limit));
            if (over == 0u) {
                weight += __shfl_sync(0xffffffffu, cw, 31);
                profit += __shfl_sync(0xffffffffu, cp, 31);
                nothing_breaks = true;
            } else {
                const int L = __ffs(over) - 1;
                weight += __shfl_sync(0xffffffffu, cw - lw, L);
                profit += __shfl_sync(0xffffffffu, cp - lp, L);
                scan_lo = min(n, L * seg);
                scan_hi = min(n, scan_lo + seg);
            }
        }
        for (int base = scan_lo; base < scan_hi && !nothing_breaks; base += 32) {
            const int s = base + lane;
            const bool und = s < scan_hi && sa[orig_s[s]] < 0;
            const double wv = und ? w_s[s] : 0.0, pv = und ? p_s[s] : 0.0;
            double cw = wv, cp = pv;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                const double ow = __shfl_up_sync(0xffffffffu, cw, off);
                const double op = __shfl_up_sync(0xffffffffu, cp, off);
                if (lane >= off) {
                    cw += ow;
                    cp += op;
                }
            }
            const bool fail = und && !((weight + cw) <= limit);
            const unsigned mask = __ballot_sync(0xffffffffu, fail);
            if (mask) {
                const int first = __ffs(mask) - 1;
                s_break = base + first;
                weight = __shfl_sync(0xffffffffu, weight + (cw - wv), first);
                profit = __shfl_sync(0xffffffffu, profit + (cp - pv), first);
                break;
            }
            weight += __shfl_sync(0xffffffffu, cw, 31);
            profit += __shfl_sync(0xffffffffu, cp, 31);
        }
        if (lane != 0) return;
        if (s_break < n) {
            const double wi = w_s[s_break];
            const double remain = __dsub_rn(capacity, weight);
            if (remain > KN_EPS && wi > KN_EPS) {
                const double frac = __ddiv_rn(remain, wi);
                profit = __dadd_rn(profit, __dmul_rn(p_s[s_break], frac));
                weight = __dadd_rn(weight, __dmul_rn(wi, frac));
                o.frac_rank = s_break;
                o.frac = frac;
            }
        }
        o.break_rank = s_break;
        o.bound = profit;
        o.weight = weight;
        bool allint = true;
        if (o.frac_rank >= 0) allint = fabs(__dsub_rn(o.frac, rint(o.frac))) < KN_EPS;
        if (allint) o.flags |= KF_ALLINT;
        if (weight > limit) o.flags |= KF_INFEASIBLE;
        kn_store(P, e, o);
        return;
    }
    if (lane != 0) return;

    // 1) items fixed to 1, in original index order (:442-452)
    double weight = 0.0, profit = 0.0;
    for (int i = 0; i < n; i++) {
        if (sa[i] == 1) {
            weight = __dadd_rn(weight, w_o[i]);
            profit = __dadd_rn(profit, p_o[i]);
        }
    }
    if (weight > limit) {  // :455-456
        o.bound = profit;
        o.weight = weight;
        o.flags = KF_INFEASIBLE | KF_EARLY;
        kn_store(P, e, o);
        return;
    }
    // 2) undecided items greedily in ratio order (:459-488)
    int s = 0;
    for (; s < n; s++) {
        if (sa[orig_s[s]] >= 0) continue;  // fixed to 1 (counted) or to 0
        const double wi = w_s[s];
        const double t = __dadd_rn(weight, wi);
        if (t <= limit) {
            weight = t;
            profit = __dadd_rn(profit, p_s[s]);
        } else {
            const double remain = __dsub_rn(capacity, weight);
            if (remain > KN_EPS && wi > KN_EPS) {
                const double frac = __ddiv_rn(remain, wi);
                profit = __dadd_rn(profit, __dmul_rn(p_s[s], frac));
                weight = __dadd_rn(weight, __dmul_rn(wi, frac));
                o.frac_rank = s;
                o.frac = frac;
            }
            break;
        }
    }
    o.break_rank = s;
    o.bound = profit;
    o.weight = weight;
    bool allint = true;  // relaxed.All(v => |v - Round(v)| < EPS): only the fractional item can fail
    if (o.frac_rank >= 0) allint = fabs(__dsub_rn(o.frac, rint(o.frac))) < KN_EPS;
    if (allint) o.flags |= KF_ALLINT;
    if (weight > limit) o.flags |= KF_INFEASIBLE;
    kn_store(P, e, o);
}

// One round = all levels of the plan in ONE cooperative launch: 8 warps per CTA, a warp per relaxation,
// a grid barrier between levels (children read their parent's result and assignment).  Inputs come
// straight from pinned host memory and results are written through to it, so a round costs one launch
// and one stream synchronisation instead of two uploads, a launch per level and a download.
// Dynamic shared memory = 8 * n bytes.
__global__ void __launch_bounds__(256) knap_round_kernel(const KnParams P) {
    namespace cg = cooperative_groups;
    cg::grid_group grid = cg::this_grid();
    extern __shared__ __align__(16) signed char sm_assign[];
    const int warp = threadIdx.x >> 5;
    signed char* sa = sm_assign + (size_t)warp * P.n;
    const int gw = blockIdx.x * 8 + warp, nw = gridDim.x * 8;
    for (int lv = 0; lv < P.nlev; lv++) {
        for (int idx = gw; idx < P.lev_count[lv]; idx += nw) knap_eval_one(P, P.lev_first[lv] + idx, sa);
        __syncwarp();
        if (lv + 1 < P.nlev) {
            __threadfence();
            grid.sync();
        }
    }
}

namespace {

struct EvalRec {
    KnOut r{};
    int slot = -1;
    int child[2] = {-1, -1};
    bool pending = false;
    bool alive = false;
};

struct HeapNode {
    double bound;
    int eval;
    std::vector<int> label;
};

int cmp_double(double a, double b) {  // double.CompareTo
    if (a < b) return -1;
    if (a > b) return 1;
    if (a == b) return 0;
    if (std::isnan(a)) return std::isnan(b) ? 0 : -1;
    return 1;
}

// SimpleMaxHeap<Node> (:494-547), same swaps, same tie behaviour.
struct MaxHeap {
    std::vector<HeapNode> data;
    static int cmp(const HeapNode& a, const HeapNode& b) { return cmp_double(a.bound, b.bound); }
    void push(HeapNode item) {
        data.push_back(std::move(item));
        int ci = (int)data.size() - 1;
        while (ci > 0) {
            const int pi = (ci - 1) / 2;
            if (cmp(data[ci], data[pi]) <= 0) break;
            std::swap(data[ci], data[pi]);
            ci = pi;
        }
    }
    HeapNode pop() {
        int li = (int)data.size() - 1;
        std::swap(data[0], data[li]);
        HeapNode ret = std::move(data[li]);
        data.pop_back();
        li = (int)data.size() - 1;
        int i = 0;
        while (true) {
            const int l = 2 * i + 1, r = 2 * i + 2;
            int largest = i;
            if (l <= li && cmp(data[l], data[largest]) > 0) largest = l;
            if (r <= li && cmp(data[r], data[largest]) > 0) largest = r;
            if (largest == i) break;
            std::swap(data[i], data[largest]);
            i = largest;
        }
        return ret;
    }
};

struct KInstance {
    MaxHeap heap;
    double best = -std::numeric_limits<double>::infinity();
    std::vector<int> best_x;
    long long evals = 0, pops = 0;
    int pop_index = 0;
    bool done = false;
    int root_eval = -1;
    std::vector<int> orig_by_rank;
};

struct KnDriver {
    int count, n;
    const double *profit, *weight, *capacity;
    lpx_options opt;
    lpx_knap_pop_fn on_pop;
    void* user;
    int spec_nodes = 16, spec_depth = 2;  // measured best of a sweep on 16 x 2000-item instances (tools/gpu_probe.py knapsweep)
    bool force_sequential = false;  // tests: take the ordered-summation path even for integer data

    std::vector<KInstance> inst;
    std::vector<EvalRec> recs;
    std::vector<int> free_recs;
    std::vector<int> free_slots;
    int n_chunks = 0, next_slot = 0;
    signed char* chunks[KN_MAX_CHUNKS] = {};
    KnParams P{};
    double *d_ws, *d_ps, *d_wo, *d_po, *d_cap;
    int *d_orig, *d_exact;

    // Pool chunks (65 MB at n = 2000) are kept between calls: cudaMalloc / cudaFree of half a gigabyte per
    // call cost more than the search itself in a process that holds many other allocations.
    struct ChunkCache {
        std::vector<signed char*> idle;
        size_t bytes = 0;
    };
    static ChunkCache& chunk_cache() {
        static ChunkCache c;
        return c;
    }
    ~KnDriver() {
        ChunkCache& cc = chunk_cache();
        const size_t bytes = (size_t)KN_CHUNK_SLOTS * n;
        for (int k = 0; k < n_chunks; k++) {
            if (cc.bytes == bytes && cc.idle.size() < 16) cc.idle.push_back(chunks[k]);
            else cudaFree(chunks[k]);
        }
    }

    int alloc_slot() {
        if (!free_slots.empty()) {
            int s = free_slots.back();
            free_slots.pop_back();
            return s;
        }
        if (next_slot >= n_chunks * KN_CHUNK_SLOTS) {
            if (n_chunks >= KN_MAX_CHUNKS) return -1;
            ChunkCache& cc = chunk_cache();
            const size_t bytes = (size_t)KN_CHUNK_SLOTS * n;
            if (cc.bytes != bytes) {  // another item count: the idle chunks have the wrong size
                for (signed char* q : cc.idle) cudaFree(q);
                cc.idle.clear();
                cc.bytes = bytes;
            }
            void* p = nullptr;
            if (!cc.idle.empty()) {
                p = cc.idle.back();
                cc.idle.pop_back();
            } else if (cudaMalloc(&p, bytes) != cudaSuccess) {
                cudaGetLastError();
                return -1;
            }
            chunks[n_chunks++] = (signed char*)p;
        }
        return next_slot++;
    }
    int alloc_rec() {
        int id;
        if (!free_recs.empty()) {
            id = free_recs.back();
            free_recs.pop_back();
        } else {
            id = (int)recs.size();
            recs.emplace_back();
        }
        recs[id] = EvalRec();
        recs[id].alive = true;
        return id;
    }
    // release an evaluation, everything cached below it, and their pool slots
    void free_tree(int id) {
        if (id < 0) return;
        std::vector<int> st{id};
        while (!st.empty()) {
            const int k = st.back();
            st.pop_back();
            EvalRec& r = recs[k];
            if (!r.alive) continue;
            for (int s = 0; s < 2; s++)
                if (r.child[s] >= 0) st.push_back(r.child[s]);
            if (r.slot >= 0) free_slots.push_back(r.slot);
            r.alive = false;
            free_recs.push_back(k);
        }
    }
    void free_self(int id) {  // the node was expanded: its children live on
        EvalRec& r = recs[id];
        if (r.slot >= 0) free_slots.push_back(r.slot);
        r.alive = false;
        free_recs.push_back(id);
    }

    signed char* slot_ptr(int slot) { return chunks[slot / KN_CHUNK_SLOTS] + (size_t)(slot % KN_CHUNK_SLOTS) * n; }

    int fetch_assigned(int slot, std::vector<signed char>& out) {
        out.resize(n);
        LPX_CUDA(cudaMemcpy(out.data(), slot_ptr(slot), n, cudaMemcpyDeviceToHost));
        return LPX_OK;
    }

    // relaxed[i] of ComputeRelaxation, rebuilt from the assignment and where the greedy pass stopped
    void relaxed_from(const KInstance& I, const std::vector<signed char>& a, const KnOut& r, std::vector<double>& relaxed) {
        relaxed.assign(n, 0.0);
        for (int i = 0; i < n; i++)
            if (a[i] == 1) relaxed[i] = 1.0;
        if (r.flags & KF_EARLY) return;  // fixed items alone exceed the capacity (:455-456)
        for (int s = 0; s < r.break_rank && s < n; s++) {
            const int o = I.orig_by_rank[s];
            if (a[o] < 0) relaxed[o] = 1.0;
        }
        if (r.frac_rank >= 0) relaxed[I.orig_by_rank[r.frac_rank]] = r.frac;
    }

    struct Plan {
        std::vector<int> roots;  // heap-node evaluations planned under (for pruning skipped results)
        std::vector<KnIn> in;
        std::vector<int> rec;    // EvalRec id per planned evaluation
        std::vector<int> level;
        std::vector<int> owner;  // sharded tree: ordinal of the speculative subtree the evaluation belongs to
        int subtrees = 0;
    };

    // plan the subtree of evaluation `id` (complete) down to `depth` levels
    void plan_below(Plan& pl, int k, int id, int depth) {
        struct Item { int rec; int pending; int depth; int level; };
        std::vector<Item> q{{id, -1, depth, 0}};
        const KInstance& I = inst[k];
        while (!q.empty()) {
            Item it = q.back();
            q.pop_back();
            if (it.depth <= 0) continue;
            EvalRec& r = recs[it.rec];
            if (!r.pending) {
                const bool expandable = !(r.r.flags & (KF_INFEASIBLE | KF_ALLINT)) && r.r.frac_rank >= 0 &&
                                        r.r.bound > I.best + KN_EPS;
                // the popped node itself is expanded whenever it has a fractional item (:147,180)
                const bool is_heap_node = it.rec == id;
                if (!(expandable || (is_heap_node && r.r.frac_rank >= 0))) continue;
            }
            for (int side = 0; side < 2; side++) {
                int ch = recs[it.rec].child[side];
                int ch_pending = -1, ch_level = it.level;
                if (ch < 0) {
                    const int slot = alloc_slot();
                    if (slot < 0) {
                        // Pool exhausted: plan what we have — except in the sharded-tree mode, where every rank must
                        // build the SAME plan (the merge is an all-reduce of total * 5 words): there it is a hard
                        // error, raised identically on whichever ranks see it, never a silently shorter plan.
                        if (shard_world > 1) plan_failed = true;
                        return;
                    }
                    ch = alloc_rec();
                    EvalRec& c = recs[ch];
                    c.slot = slot;
                    c.pending = true;
                    recs[it.rec].child[side] = ch;
                    KnIn in;
                    in.inst = k;
                    in.side = side;
                    in.out_slot = slot;
                    if (recs[it.rec].pending) {
                        in.parent_slot = -1;
                        in.parent_pending = it.pending;
                        in.var = -2;
                    } else {
                        in.parent_slot = recs[it.rec].slot;
                        in.parent_pending = -1;
                        in.var = I.orig_by_rank[recs[it.rec].r.frac_rank];
                    }
                    ch_pending = (int)pl.in.size();
                    pl.in.push_back(in);
                    pl.rec.push_back(ch);
                    pl.level.push_back(it.level);
                    pl.owner.push_back(pl.subtrees);
                    ch_level = it.level + 1;
                }
                q.push_back({ch, ch_pending, it.depth - 1, ch_level});
            }
        }
        pl.subtrees++;
    }

    int coop_ctas = 1;
    bool plan_failed = false;
    int shard_world = 1, shard_rank = 0;  // > 1: one tree over all ranks of the communicator
    double tr_stage = 0, tr_launch = 0, tr_sync = 0, tr_unpack = 0;  // LPX_KNAP_TRACE: inside the device rounds
    long tr_evals = 0;

    int run_plan(Plan& pl) {
        Runtime& r = rt();
        const auto tp0 = std::chrono::steady_clock::now();
        const int total = (int)pl.in.size();
        if (total == 0) return LPX_OK;
        // order by level (stable), remapping parent_pending
        std::vector<int> order(total), pos(total);
        for (int i = 0; i < total; i++) order[i] = i;
        std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return pl.level[a] < pl.level[b]; });
        for (int i = 0; i < total; i++) pos[order[i]] = i;
        KnIn* h_in = ws_pin_as<KnIn>(WS_KN_AUX, total);
        KnOut* h_out = ws_pin_as<KnOut>(WS_KN_OUT, total);
        KnOut* d_out = ws_dev_as<KnOut>(WS_KN_OUT, total);
        if (!h_in || !h_out || !d_out) return LPX_E_CUDA;
        // the kernel reads its inputs from, and writes its results through to, pinned host memory
        KnIn* m_in = nullptr;
        KnOut* m_out = nullptr;
        LPX_CUDA(cudaHostGetDevicePointer((void**)&m_in, h_in, 0));
        LPX_CUDA(cudaHostGetDevicePointer((void**)&m_out, h_out, 0));
        std::vector<int> rec_sorted(total);
        std::vector<int> level_first, level_count;
        for (int i = 0; i < total; i++) {
            KnIn in = pl.in[order[i]];
            in.parent_out_slot = -1;
            if (in.parent_pending >= 0) {
                in.parent_out_slot = pl.in[in.parent_pending].out_slot;
                in.parent_pending = pos[in.parent_pending];
            }
            in.foreign = (shard_world > 1 && pl.owner[order[i]] % shard_world != shard_rank) ? 1 : 0;
            in.best = inst[in.inst].best;
            h_in[i] = in;
            rec_sorted[i] = pl.rec[order[i]];
            const int lv = pl.level[order[i]];
            if ((int)level_first.size() <= lv) {
                level_first.resize(lv + 1, i);
                level_count.resize(lv + 1, 0);
            }
            level_count[lv]++;
        }
        if ((int)level_first.size() > KN_MAX_LEVELS) {
            set_error("knapsack: speculation depth exceeds the round kernel's level table");
            return LPX_E_CAPACITY;
        }
        cudaStream_t s = r.stream;
        for (int k = 0; k < n_chunks; k++) P.chunks[k] = chunks[k];
        P.in = m_in;
        P.out = d_out;
        P.out_host = m_out;
        P.nlev = 0;
        int widest = 0;
        for (size_t lv = 0; lv < level_first.size(); lv++) {
            if (level_count[lv] == 0) continue;
            P.lev_first[P.nlev] = level_first[lv];
            P.lev_count[P.nlev] = level_count[lv];
            P.nlev++;
            widest = std::max(widest, level_count[lv]);
        }
        const size_t smem = (size_t)8 * n;
        const int grid = std::max(1, std::min((widest + 7) / 8, coop_ctas));
        void* args[] = {(void*)&P};
        const auto tp1 = std::chrono::steady_clock::now();
        P.pass = 0;
        LPX_CUDA(cudaLaunchCooperativeKernel((const void*)knap_round_kernel, dim3(grid), dim3(256), args, smem, s));
        count_launch();
        if (shard_world > 1) {
            // the exchange step of the round: merge every rank's relaxations, then complete the pools
            static_assert(sizeof(KnOut) % 8 == 0, "KnOut is merged as 8-byte words");
            int rc = comm_merge_u64(reinterpret_cast<unsigned long long*>(d_out), (size_t)total * (sizeof(KnOut) / 8), s);
            if (rc != LPX_OK) return rc;
            P.pass = 1;
            LPX_CUDA(cudaLaunchCooperativeKernel((const void*)knap_round_kernel, dim3(grid), dim3(256), args, smem, s));
            count_launch();
        }
        const auto tp2 = std::chrono::steady_clock::now();
        LPX_CUDA(cudaStreamSynchronize(s));
        const auto tp3 = std::chrono::steady_clock::now();
        tr_stage += std::chrono::duration<double>(tp1 - tp0).count();
        tr_launch += std::chrono::duration<double>(tp2 - tp1).count();
        tr_sync += std::chrono::duration<double>(tp3 - tp2).count();
        tr_evals += total;
        for (int i = 0; i < total; i++) {
            EvalRec& rec = recs[rec_sorted[i]];
            rec.r = h_out[i];
            rec.pending = false;
        }
        return LPX_OK;
    }

    // remove skipped children below `id` so that a later plan re-evaluates them if needed
    void prune_skipped(int id) {
        std::vector<int> st{id};
        while (!st.empty()) {
            const int k = st.back();
            st.pop_back();
            for (int s = 0; s < 2; s++) {
                const int ch = recs[k].child[s];
                if (ch < 0) continue;
                if (recs[ch].r.flags & KF_SKIPPED) {
                    free_tree(ch);
                    recs[k].child[s] = -1;
                } else {
                    st.push_back(ch);
                }
            }
        }
    }

    void fill_eval(lpx_knap_eval& ev, const EvalRec& r, int pop_index, int child, int decision,
                   const signed char* assigned) {
        ev.pop_index = pop_index;
        ev.child = child;
        ev.var = r.r.var;
        ev.bound = r.r.bound;
        ev.weight = r.r.weight;
        ev.frac_rank = r.r.frac_rank;
        ev.frac = r.r.frac;
        ev.break_rank = r.r.break_rank;
        ev.decision = decision;
        ev.assigned = assigned;
    }

    // Commit as many pops of instance k as the cached evaluations allow.
    int commit(int k) {
        KInstance& I = inst[k];
        std::vector<signed char> a_node, a_child[2];
        std::vector<double> relaxed;
        while (!I.heap.data.empty()) {
            const HeapNode& top = I.heap.data[0];
            const bool skip = top.bound <= I.best + KN_EPS;
            const EvalRec& tr = recs[top.eval];
            if (!skip && tr.r.frac_rank >= 0 && (tr.child[0] < 0 || tr.child[1] < 0 || recs[tr.child[0]].pending ||
                                                 recs[tr.child[1]].pending))
                return LPX_OK;  // children not evaluated yet: next round
            HeapNode node = I.heap.pop();
            I.pops++;
            if (skip) {  // :124
                free_tree(node.eval);
                continue;
            }
            const int this_pop = I.pop_index++;
            EvalRec cur = recs[node.eval];
            lpx_knap_pop pop;
            std::memset(&pop, 0, sizeof pop);
            pop.pop_index = this_pop;
            pop.label_len = (int)node.label.size();
            pop.label = node.label.data();
            const bool want_cb = on_pop != nullptr;
            if (want_cb) {
                int rc = fetch_assigned(cur.slot, a_node);
                if (rc != LPX_OK) return rc;
            }
            if (cur.r.frac_rank < 0) {  // :147-177
                int closed = 3;
                if (cur.r.weight <= capacity[k] + KN_EPS) {
                    if (cur.r.bound > I.best + KN_EPS) {
                        I.best = cur.r.bound;
                        if (!want_cb) {
                            int rc = fetch_assigned(cur.slot, a_node);
                            if (rc != LPX_OK) return rc;
                        }
                        relaxed_from(I, a_node, cur.r, relaxed);
                        for (int i = 0; i < n; i++) I.best_x[i] = relaxed[i] >= 0.5 ? 1 : 0;
                        closed = 1;
                    } else {
                        closed = 2;
                    }
                }
                if (want_cb) {
                    fill_eval(pop.relax, cur, this_pop, -1, LPX_KN_ROOT, a_node.data());
                    pop.closed = closed;
                    on_pop(&pop, nullptr, nullptr, user);
                }
                free_tree(node.eval);
                continue;
            }
            lpx_knap_eval evs[2];
            for (int side = 0; side < 2; side++) {
                const int ce = cur.child[side];
                const EvalRec cr = recs[ce];
                I.evals++;
                int decision;
                if (want_cb) {
                    int rc = fetch_assigned(cr.slot, a_child[side]);
                    if (rc != LPX_OK) return rc;
                }
                if (cr.r.weight > capacity[k] + KN_EPS) {
                    decision = LPX_KN_INFEASIBLE;
                    free_tree(ce);
                } else if (cr.r.bound > I.best + KN_EPS) {
                    if (cr.r.flags & KF_ALLINT) {
                        decision = LPX_KN_CANDIDATE_INT;
                        if (cr.r.bound > I.best + KN_EPS) {
                            I.best = cr.r.bound;
                            if (!want_cb) {
                                int rc = fetch_assigned(cr.slot, a_child[side]);
                                if (rc != LPX_OK) return rc;
                            }
                            relaxed_from(I, a_child[side], cr.r, relaxed);
                            for (int i = 0; i < n; i++) I.best_x[i] = (int)std::nearbyint(relaxed[i]);
                        }
                        free_tree(ce);
                    } else {
                        decision = LPX_KN_PUSHED;
                        HeapNode hn;
                        hn.bound = cr.r.bound;
                        hn.eval = ce;
                        if (want_cb) {
                            if (node.label.size() == 1 && node.label[0] == 0) hn.label = {side + 1};
                            else {
                                hn.label = node.label;
                                hn.label.push_back(side + 1);
                            }
                        }
                        I.heap.push(std::move(hn));
                    }
                } else {
                    decision = LPX_KN_DROPPED;
                    free_tree(ce);
                }
                if (want_cb) fill_eval(evs[side], cr, this_pop, side, decision, a_child[side].data());
            }
            if (want_cb) {
                fill_eval(pop.relax, cur, this_pop, -1, LPX_KN_ROOT, a_node.data());
                pop.closed = 0;
                on_pop(&pop, &evs[0], &evs[1], user);
            }
            free_self(node.eval);
        }
        I.done = true;
        return LPX_OK;
    }

    int run() {
        Runtime& r = rt();
        inst.resize(count);
        // ratio ordering on the host (:75-79): stable OrderByDescending(Ratio).ThenByDescending(Profit)
        std::vector<double> h_ws((size_t)count * n), h_ps((size_t)count * n);
        std::vector<int> h_orig((size_t)count * n);
        for (int k = 0; k < count; k++) {
            const double* p = profit + (size_t)k * n;
            const double* w = weight + (size_t)k * n;
            std::vector<int> idx(n);
            for (int i = 0; i < n; i++) idx[i] = i;
            auto ratio = [&](int i) { return w[i] > 0 ? p[i] / w[i] : std::numeric_limits<double>::infinity(); };
            std::stable_sort(idx.begin(), idx.end(), [&](int a, int b) {
                const int c = cmp_double(ratio(a), ratio(b));
                if (c != 0) return c > 0;
                return cmp_double(p[a], p[b]) > 0;
            });
            inst[k].orig_by_rank = idx;
            inst[k].best_x.assign(n, 0);
            for (int s = 0; s < n; s++) {
                h_ws[(size_t)k * n + s] = w[idx[s]];
                h_ps[(size_t)k * n + s] = p[idx[s]];
                h_orig[(size_t)k * n + s] = idx[s];
            }
        }
        const size_t cn = (size_t)count * n;
        // order-free summation is allowed only where it is provably exact: integer weights and
        // profits whose absolute sums stay below 2^53
        std::vector<int> h_exact(count);
        for (int k = 0; k < count; k++) {
            double aw = 0, ap = 0;
            bool ok = true;
            for (int i = 0; i < n && ok; i++) {
                const double wv = weight[(size_t)k * n + i], pv = profit[(size_t)k * n + i];
                ok = std::isfinite(wv) && std::isfinite(pv) && wv == std::nearbyint(wv) && pv == std::nearbyint(pv);
                aw += std::fabs(wv);
                ap += std::fabs(pv);
            }
            h_exact[k] = ok && aw < 4503599627370496.0 && ap < 4503599627370496.0 && !force_sequential;
            if (h_exact[k]) {
                bool nonneg = true;
                for (int i = 0; i < n; i++) nonneg = nonneg && weight[(size_t)k * n + i] >= 0.0;
                if (nonneg) h_exact[k] = 2;  // lets the kernel rank whole segments before walking one
            }
        }
        d_exact = ws_dev_as<int>(WS_KN_EXACT, count);
        if (!d_exact) return LPX_E_CUDA;
        LPX_CUDA(cudaMemcpyAsync(d_exact, h_exact.data(), (size_t)count * 4, cudaMemcpyHostToDevice, rt().stream));
        d_ws = ws_dev_as<double>(WS_KN_ITEMS, cn * 4);
        d_orig = ws_dev_as<int>(WS_KN_ASSIGN, cn);
        d_cap = ws_dev_as<double>(WS_MISC1, count);
        if (!d_ws || !d_orig || !d_cap) return LPX_E_CUDA;
        d_ps = d_ws + cn;
        d_wo = d_ws + 2 * cn;
        d_po = d_ws + 3 * cn;
        cudaStream_t s = r.stream;
        LPX_CUDA(cudaMemcpyAsync(d_ws, h_ws.data(), cn * 8, cudaMemcpyHostToDevice, s));
        LPX_CUDA(cudaMemcpyAsync(d_ps, h_ps.data(), cn * 8, cudaMemcpyHostToDevice, s));
        LPX_CUDA(cudaMemcpyAsync(d_wo, weight, cn * 8, cudaMemcpyHostToDevice, s));
        LPX_CUDA(cudaMemcpyAsync(d_po, profit, cn * 8, cudaMemcpyHostToDevice, s));
        LPX_CUDA(cudaMemcpyAsync(d_orig, h_orig.data(), cn * 4, cudaMemcpyHostToDevice, s));
        LPX_CUDA(cudaMemcpyAsync(d_cap, capacity, (size_t)count * 8, cudaMemcpyHostToDevice, s));
        LPX_CUDA(cudaStreamSynchronize(s));
        P.w_s = d_ws;
        P.p_s = d_ps;
        P.orig_s = d_orig;
        P.w_o = d_wo;
        P.p_o = d_po;
        P.cap = d_cap;
        P.exact = d_exact;
        P.n = n;
        LPX_CUDA(cudaFuncSetAttribute(knap_round_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(8 * n)));
        {   // how many CTAs of the round kernel can be resident at once (a cooperative launch needs them all)
            int per_sm = 0;
            LPX_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, knap_round_kernel, 256, (size_t)8 * n));
            coop_ctas = std::max(1, per_sm * r.sms);
        }

        // roots: all undecided (:101-113)
        {
            Plan pl;
            const int seed_slot = alloc_slot();
            if (seed_slot < 0) {
                set_error("knapsack: cannot allocate the node pool");
                return LPX_E_CUDA;
            }
            LPX_CUDA(cudaMemset(slot_ptr(seed_slot), 0xFF, n));
            for (int k = 0; k < count; k++) {
                const int slot = alloc_slot();
                const int id = alloc_rec();
                recs[id].slot = slot;
                recs[id].pending = true;
                inst[k].root_eval = id;
                KnIn in;
                in.inst = k;
                in.parent_slot = seed_slot;
                in.parent_pending = -1;
                in.var = -1;
                in.side = 0;
                in.out_slot = slot;
                pl.in.push_back(in);
                pl.rec.push_back(id);
                pl.level.push_back(0);
                pl.owner.push_back(pl.subtrees++);
            }
            int rc = run_plan(pl);
            if (rc != LPX_OK) return rc;
            for (int k = 0; k < count; k++) {
                HeapNode hn;
                hn.bound = recs[inst[k].root_eval].r.bound;
                hn.eval = inst[k].root_eval;
                hn.label = {0};
                inst[k].heap.push(std::move(hn));
                inst[k].evals++;
            }
        }

        int stalled = 0;
        double tr[3] = {0, 0, 0};  // LPX_KNAP_TRACE=1: seconds in plan / device rounds / commit
        long tr_rounds = 0;
        auto now = [] { return std::chrono::steady_clock::now(); };
        auto secs = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) {
            return std::chrono::duration<double>(b - a).count();
        };
        while (true) {
            const auto t0 = now();
            Plan pl;
            bool any = false;
            for (int k = 0; k < count; k++) {
                KInstance& I = inst[k];
                if (I.done) continue;
                any = true;
                // best-first peek at the top-K heap entries without disturbing the heap
                auto cmpq = [&](int a, int b) { return cmp_double(I.heap.data[a].bound, I.heap.data[b].bound) < 0; };
                std::priority_queue<int, std::vector<int>, decltype(cmpq)> pq(cmpq);
                if (!I.heap.data.empty()) pq.push(0);
                int taken = 0;
                while (!pq.empty() && taken < spec_nodes) {
                    const int i = pq.top();
                    pq.pop();
                    const int l = 2 * i + 1, rr = 2 * i + 2;
                    if (l < (int)I.heap.data.size()) pq.push(l);
                    if (rr < (int)I.heap.data.size()) pq.push(rr);
                    const HeapNode& hn = I.heap.data[i];
                    if (hn.bound <= I.best + KN_EPS) continue;
                    // the front runner gets the deep look-ahead, the others one level
                    plan_below(pl, k, hn.eval, taken == 0 ? spec_depth : 1);
                    pl.roots.push_back(hn.eval);
                    taken++;
                }
            }
            if (!any) {
                if (getenv("LPX_KNAP_TRACE"))
                    fprintf(stderr,
                            "[knap trace] %ld rounds: plan %.3f s, device %.3f s (stage %.3f, launch %.3f, wait %.3f; %ld "
                            "relaxations), commit %.3f s\n",
                            tr_rounds, tr[0], tr[1], tr_stage, tr_launch, tr_sync, tr_evals, tr[2]);
                break;
            }
            const auto t1 = now();
            if (shard_world > 1) {
                // agree on the plan before any collective is sized by it
                double bad = plan_failed ? 1.0 : 0.0;
                int rcb = lpx_comm_allreduce_max(&bad, 1);
                if (rcb != LPX_OK) return rcb;
                if (bad != 0.0) {
                    set_error("knapsack (sharded tree): node pool exhausted on a rank; the plans would differ");
                    return LPX_E_CAPACITY;
                }
            }
            int rc = run_plan(pl);
            if (rc != LPX_OK) return rc;
            const auto t2 = now();
            tr_rounds++;
            bool progressed = false;
            // unlink evaluations the device skipped, so they can be planned again later
            for (int id : pl.roots) prune_skipped(id);
            for (int k = 0; k < count; k++) {
                KInstance& I = inst[k];
                if (I.done) continue;
                const long long before = I.pops;
                rc = commit(k);
                if (rc != LPX_OK) return rc;
                if (I.pops != before || I.done) progressed = true;
            }
            if (shard_world > 1) {
                // the 8-byte incumbent max-allreduce of the batch: with replicated commits it must be a no-op
                std::vector<double> bests(count), mine(count);
                for (int k = 0; k < count; k++) bests[k] = mine[k] = inst[k].best;
                rc = lpx_comm_allreduce_max(bests.data(), count);
                if (rc != LPX_OK) return rc;
                for (int k = 0; k < count; k++)
                    if (cmp_double(bests[k], mine[k]) != 0) {
                        set_error("knapsack (sharded tree): ranks disagree on the incumbent");
                        return LPX_E_NCCL;
                    }
            }
            tr[0] += secs(t0, t1);
            tr[1] += secs(t1, t2);
            tr[2] += secs(t2, now());
            stalled = progressed ? 0 : stalled + 1;
            if (stalled > 2 || (!progressed && pl.in.empty())) {
                set_error("knapsack: no progress possible (node pool exhausted?)");
                return LPX_E_CAPACITY;
            }
        }
        return LPX_OK;
    }
};

}  // namespace

// idle pool chunks go with the device (lpx_shutdown)
void knapsack_release_cache() {
    KnDriver::ChunkCache& cc = KnDriver::chunk_cache();
    for (signed char* q : cc.idle) cudaFree(q);
    cc.idle.clear();
    cc.bytes = 0;
    knapsack_dev_release_cache();
}
}  // namespace lpx

using namespace lpx;

static int knapsack_entry(int count, int n, const double* profit, const double* weight, const double* capacity,
                          const lpx_options* opt, int* found, double* best_value, int* best_x, long long* n_evals,
                          long long* n_pops, int* rank_order, lpx_knap_pop_fn on_pop, void* user) {
    if (count < 1 || n < 1 || !profit || !weight || !capacity) {
        set_error("lpx_bnb_knapsack: bad arguments");
        return LPX_E_BAD_ARGS;
    }
    if ((size_t)8 * n > (size_t)200 * 1024) {
        set_error("lpx_bnb_knapsack: more than 25600 items per instance is not supported");
        return LPX_E_CAPACITY;
    }
    int rc = ensure_device();
    if (rc != LPX_OK) return rc;
    std::lock_guard<std::recursive_mutex> lk(rt().mu);
    {
        lpx_options o;
        lpx_default_options(&o);
        if (opt) o = *opt;
        // The search loop runs on the device (lpx_knap_dev.cu).  Only the one-tree-over-all-ranks mode still
        // plans rounds on the host, because its exchange step is a host-driven NCCL collective.
        if (!(o.knap_shard_tree && comm_world() > 1))
            return knapsack_search_device(count, n, profit, weight, capacity, o, found, best_value, best_x, n_evals,
                                          n_pops, rank_order, on_pop, user);
    }
    KnDriver d;
    d.count = count;
    d.n = n;
    d.profit = profit;
    d.weight = weight;
    d.capacity = capacity;
    lpx_default_options(&d.opt);
    if (opt) d.opt = *opt;
    if (d.opt.knap_spec_nodes > 0) d.spec_nodes = d.opt.knap_spec_nodes;
    if (d.opt.knap_spec_depth > 0) d.spec_depth = d.opt.knap_spec_depth;
    d.force_sequential = d.opt.knap_ordered_sums != 0;
    if (d.opt.knap_shard_tree && comm_world() > 1) {
        d.shard_world = comm_world();
        d.shard_rank = comm_rank();
    }
    d.on_pop = on_pop;
    d.user = user;
    rc = d.run();
    if (rc != LPX_OK) return rc;
    for (int k = 0; k < count; k++) {
        const KInstance& I = d.inst[k];
        const bool none = std::isinf(I.best) && I.best < 0;
        if (found) found[k] = none ? 0 : 1;
        if (best_value) best_value[k] = I.best;
        if (best_x)
            for (int i = 0; i < n; i++) best_x[(size_t)k * n + i] = I.best_x[i];
        if (n_evals) n_evals[k] = I.evals;
        if (n_pops) n_pops[k] = I.pops;
        if (rank_order && k == 0)
            for (int s = 0; s < n; s++) rank_order[s] = I.orig_by_rank[s];
    }
    return LPX_OK;
}

extern "C" {

int lpx_bnb_knapsack(int n, const double* profit, const double* weight, double capacity, const lpx_options* opt,
                     int* found, double* best_value, int* best_x, long long* n_evals, long long* n_pops,
                     int* rank_order, lpx_knap_pop_fn on_pop, void* user) {
    return knapsack_entry(1, n, profit, weight, &capacity, opt, found, best_value, best_x, n_evals, n_pops,
                          rank_order, on_pop, user);
}

int lpx_bnb_knapsack_batched(int count, int n, const double* profit, const double* weight, const double* capacity,
                             const lpx_options* opt, int* found, double* best_value, int* best_x,
                             long long* n_evals, long long* n_pops) {
    return knapsack_entry(count, n, profit, weight, capacity, opt, found, best_value, best_x, n_evals, n_pops, nullptr,
                          nullptr, nullptr);
}

}  // extern "C"
