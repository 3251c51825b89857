// lpx_knap_dev.cu — device-resident best-first Branch & Bound for the 0/1 knapsack:
// BranchAndBoundKnapsack.Solve, R/Models/BranchAndBoundKnapsack.cs:58-407, with the whole search loop
// on the GPU.
//
// One WARP owns one instance from the root to the last pop: it pops the node with the largest bound
// from a heap that replays SimpleMaxHeap's swaps (:494-547), evaluates both children
// (ComputeRelaxation, :431-491), commits them left then right exactly as the C# loop does (:207-327:
// infeasible | integral -> incumbent | push | drop) and goes on — no launch, grid barrier or host
// round trip per pop.  A persistent grid of warps takes instances from a queue, so a batch keeps
// every SM busy until its last instance ends.
//
// Node = two bit masks + 72 bytes: ONE (items fixed to 1, by ORIGINAL index: the order of the
// reference's first summation) and DEC (items decided either way, by RATIO RANK: the order of its
// greedy pass), the relaxation's result, and the two partial states that make children cheap:
//   (w1, p1)  the sums over the fixed-to-1 items               -> a left child (x_k = 0) inherits them
//   (wb, pb)  the greedy sums just before the break item k     -> a left child RESUMES the greedy
//             pass there: its prefix is the parent's, operation for operation.
// A right child (x_k = 1) sums the fixed items again in original index order and restarts the pass.
//
// Summation order is part of the result in floating point.  Ordered mode: the whole warp walks the
// ranks 32 at a time, every lane computing the same dependent DADD chain speculatively through the
// block (lane j keeps the j-th partial sum) and one ballot finds the first item that does not fit —
// a pure add chain, no compare or branch on it.  Exact mode (integer weights >= 0 and integer
// profits whose sums stay below 2^52: every partial sum is exact, so any order gives the reference's
// bits): prefix-sum tables by rank minus the decided items, one warp scan, one 32-item walk.
//
// The heap lives in global memory with its first levels cached in shared memory; a push loads the
// whole ancestor path at once and decides the climb with one ballot.
//
// The host (knapsack_search_device) only sorts the items, sizes the node pools, launches, and — when
// a callback is attached — replays the pop / evaluation records the kernel appended, pausing and
// resuming the kernel whenever the record buffer fills.
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <thread>
#include <unordered_map>
#include <vector>

#include <cuda_pipeline.h>

#include "lpx_common.cuh"
#include "lpx_knap.hpp"
#include "lpx_runtime.hpp"

namespace lpx {

#define KS_EPS 1e-9  // BranchAndBoundKnapsack.EPS (:56)

enum { KS_F_INFEASIBLE = 1, KS_F_ALLINT = 2, KS_F_EARLY = 8 };
enum { KS_RUNNING = 0, KS_DONE = 1, KS_PAUSED = 2, KS_OVERFLOW = 3 };

struct KsMeta {  // 72 bytes per node
    double bound, weight, frac;  // relaxation profit (== bound), weight, fraction of the break item
    double w1, p1;               // sums over the items fixed to 1 (original index order)
    double wb, pb;               // greedy sums just before the break item
    int frac_rank, break_rank, flags, var;
};

struct KsRec {  // one trace record (pop or child evaluation), 64 bytes
    int kind;   // 0 pop, 1 evaluation
    int node;   // pop: the node taken; evaluation: the node pushed, -1 if none
    int code;   // pop: closed (0 expanded, 1 best candidate, 2 candidate, 3 infeasible); evaluation: LPX_KN_*
    int var, child, frac_rank, break_rank, flags;
    double bound, weight, frac;
    double pad;
};

struct KsState {  // per instance, in global memory: everything a resumed kernel needs
    double best;
    long long evals, pops;
    int heap_size, free_top, next_fresh, status;
    int started, best_kind, trace_count, pad;
    KsMeta best_meta;
};

struct KsHeapEntry {  // global-memory part of the heap: one 16-byte load per entry
    unsigned long long key;  // ks_key(bound)
    int n, pad;
};

struct KsParams {
    // item tables, [inst][n]
    const double *w_s, *p_s, *w_o, *p_o, *cap;
    const int *orig_s, *exact;
    int n, W, MS, count;  // W mask words per set; MS words per node (ONE then DEC), a multiple of 4
    // per-slot arenas: slot q serves instance todo[q]
    const int* todo;
    const int* slots;   // arena slot of queue entry q (an instance keeps its slot when it is resumed in a shorter queue)
    int n_todo;
    int pop_budget;     // > 0: an instance pauses after this many pops in one launch (stragglers of a large batch
                        // are resumed with a pair of warps and an SM's heap cache each)
    int* next;          // work-queue head
    int node_cap;       // nodes (= heap entries = free-stack entries) per slot
    KsHeapEntry* heap;  // [slot][node_cap]
    int* free_stack;
    KsMeta* meta;
    unsigned* masks;    // [slot][node_cap][MS]   ONE (orig index) then DEC (rank)
    unsigned* best_masks;  // [inst][2 W]
    KsState* state;     // [inst]
    KsRec* trace;       // [slot][trace_cap] or null
    int trace_cap;
    int HS;             // heap entries cached in shared memory per warp
    int stage_items;    // 1: every warp keeps its instance's rank-ordered (weight, profit) tables in shared memory
    long long* prof;    // LPX_KNAP_PROF=1: clock64() sums of instance 0 per phase
};

// Heap order = double.CompareTo on the bounds (:99): NaN below everything and equal to itself, -0.0
// equal to +0.0.  The heap stores an integer key with exactly that order, so its compares are integer
// compares: NaN -> 0, -0.0 -> the key of +0.0, everything else the usual order-preserving map (> 0).
__device__ __forceinline__ unsigned long long ks_key(double x) {
    if (x != x) return 0ULL;
    return dkey(x + 0.0);  // -0.0 + 0.0 = +0.0, every other value unchanged
}

// Everything below is executed by all 32 lanes with warp-uniform control flow: values called
// "uniform" are identical in every lane (same loads, shuffles from one lane), stores are lane 0's.
struct KsInst {
    const double *w_s, *p_s, *w_o, *p_o;
    const int* orig_s;
    double capacity, limit;
    int n, W, exact;
    const unsigned *one, *dec;  // the popped node's masks, staged in shared memory
    double2* st_in;             // per-warp staging: 32 (w, p) operands of the add chain
    double2* st_out;            // and its 32 partial sums
};

struct KsEval {
    double bound, weight, frac, w1, p1, wb, pb;
    int frac_rank, break_rank, flags;
};

__device__ __forceinline__ KsEval ks_early(double w1, double p1, int n) {
    // the fixed items alone exceed the capacity (:455-456)
    KsEval o;
    o.bound = p1;
    o.weight = w1;
    o.frac = 0.0;
    o.w1 = w1;
    o.p1 = p1;
    o.wb = w1;
    o.pb = p1;
    o.frac_rank = -1;
    o.break_rank = n;
    o.flags = KS_F_INFEASIBLE | KS_F_EARLY;
    return o;
}

// The break item and the flags (:476-487).  (weight, profit) = greedy sums before rank s_break.
__device__ __forceinline__ KsEval ks_finish(const KsInst& I, double w1, double p1, double weight, double profit,
                                            int s_break) {
    KsEval o;
    o.w1 = w1;
    o.p1 = p1;
    o.wb = weight;
    o.pb = profit;
    o.frac_rank = -1;
    o.frac = 0.0;
    o.break_rank = s_break;
    o.flags = 0;
    if (s_break < I.n) {
        const double wi = I.w_s[s_break];
        const double remain = __dsub_rn(I.capacity, weight);
        if (remain > KS_EPS && wi > KS_EPS) {
            const double frac = __ddiv_rn(remain, wi);
            profit = __dadd_rn(profit, __dmul_rn(I.p_s[s_break], frac));
            weight = __dadd_rn(weight, __dmul_rn(wi, frac));
            o.frac_rank = s_break;
            o.frac = frac;
        }
    }
    o.bound = profit;
    o.weight = weight;
    bool allint = true;  // relaxed.All(v => |v - Round(v)| < EPS): only the fractional item can fail
    if (o.frac_rank >= 0) allint = fabs(__dsub_rn(o.frac, rint(o.frac))) < KS_EPS;
    if (allint) o.flags |= KS_F_ALLINT;
    if (weight > I.limit) o.flags |= KS_F_INFEASIBLE;
    return o;
}

// DEC word b of the child: the parent's word with the branching item's rank bit added.
__device__ __forceinline__ unsigned ks_dec_word(const KsInst& I, int b, int sfw, unsigned sfbit) {
    return I.dec[b] | (b == sfw ? sfbit : 0u);
}

// Greedy pass in the reference's order (:459-488) from rank `start` with running sums (weight, profit).
// 32 ranks per step: lane j stages item 32 b + j (zero if decided — adding +0.0 changes no bit, the
// sums are never -0.0) in shared memory; every lane then runs the same 32-add chain — one broadcast
// load and two DADDs per item, nothing else on the dependent chain —
// and the first partial sum that does not fit marks the break item (the compare sets a bit off the chain).
__device__ __forceinline__ KsEval ks_eval_chain(const KsInst& I, int sfw, unsigned sfbit, int start, double weight,
                                                double profit, double w1, double p1) {
    const int lane = threadIdx.x & 31;
    const int n = I.n;
    int b = start >> 5;
    auto fetch = [&](int blk, double& wv, double& pv) {
        const int r = (blk << 5) + lane;
        wv = 0.0;
        pv = 0.0;
        if (blk < I.W && r >= start && r < n) {
            const unsigned m = ks_dec_word(I, blk, sfw, sfbit);
            if (!((m >> lane) & 1u)) {
                wv = I.w_s[r];
                pv = I.p_s[r];
            }
        }
    };
    // Two staging buffers: while the chain runs over block b, block b + 1 is being written and block b + 2
    // fetched — none of that depends on the running sums, so it fills the chain's latency slots.
    double wn, pn;
    fetch(b, wn, pn);
    ((b & 1) ? I.st_out : I.st_in)[lane] = make_double2(wn, pn);
    unsigned neg_cur = __ballot_sync(0xffffffffu, wn < 0.0);  // lanes of the block with a negative weight
    fetch(b + 1, wn, pn);
    unsigned neg_next = __ballot_sync(0xffffffffu, wn < 0.0);
    __syncwarp();
    const bool int_cmp = I.limit >= 0.0;
    const long long lim_bits = __double_as_longlong(I.limit);
    for (; b < I.W; b++) {
        const double2* cur = (b & 1) ? I.st_out : I.st_in;
        ((b & 1) ? I.st_in : I.st_out)[lane] = make_double2(wn, pn);
        fetch(b + 2, wn, pn);
        double Wc = weight, Pc = profit;
        unsigned fail = 0u;  // bit j: the sum through item j does not fit (zeros never flip it first)
        // Non-negative weights in the block (the usual case): the partial sums are monotone (fl(a + b) >= a for
        // b >= 0), so the block's LAST sum decides whether anything in it fails and the chain carries no compare
        // at all; only the block that holds the break item is run again with the per-item test.
        const bool monotone = neg_cur == 0u;
        bool per_item = !monotone;
        if (monotone) {
#pragma unroll
            for (int j0 = 0; j0 < 32; j0 += 8) {
                double2 v[8];
#pragma unroll
                for (int j = 0; j < 8; j++) v[j] = cur[j0 + j];
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    Wc = __dadd_rn(Wc, v[j].x);
                    Pc = __dadd_rn(Pc, v[j].y);
                }
            }
            if (!(Wc <= I.limit)) {
                per_item = true;
                Wc = weight;
                Pc = profit;
            }
        }
        if (per_item && int_cmp) {
            // limit >= 0: "!(W <= limit)" is a signed compare of the bit patterns (a negative W fits, +NaN does
            // not), which keeps the FP64 pipe for the two add chains alone
#pragma unroll
            for (int j0 = 0; j0 < 32; j0 += 8) {
                double2 v[8];
#pragma unroll
                for (int j = 0; j < 8; j++) v[j] = cur[j0 + j];
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    Wc = __dadd_rn(Wc, v[j].x);
                    Pc = __dadd_rn(Pc, v[j].y);
                    if (__double_as_longlong(Wc) > lim_bits) fail |= 1u << (j0 + j);
                }
            }
        } else if (per_item) {
#pragma unroll
            for (int j0 = 0; j0 < 32; j0 += 8) {
                double2 v[8];
#pragma unroll
                for (int j = 0; j < 8; j++) v[j] = cur[j0 + j];
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    Wc = __dadd_rn(Wc, v[j].x);
                    Pc = __dadd_rn(Pc, v[j].y);
                    if (!(Wc <= I.limit)) fail |= 1u << (j0 + j);
                }
            }
        }
        if (fail) {
            // the break item; the sums before it are the chain again, up to it (once per evaluation)
            const int jb = __ffs(fail) - 1;
            for (int j = 0; j < jb; j++) {
                const double2 v = cur[j];
                weight = __dadd_rn(weight, v.x);
                profit = __dadd_rn(profit, v.y);
            }
            __syncwarp();
            return ks_finish(I, w1, p1, weight, profit, (b << 5) + jb);
        }
        weight = Wc;
        profit = Pc;
        neg_cur = neg_next;
        neg_next = __ballot_sync(0xffffffffu, wn < 0.0);  // block b + 2's operands have long arrived
        __syncwarp();
    }
    return ks_finish(I, w1, p1, weight, profit, n);
}

// Sums over the items fixed to 1 in ORIGINAL index order (:442-452); kw/kbit add the branching item.
// 32 fixed items per round: a warp scan of the per-word bit counts numbers them in order, lane t finds
// the t-th of the round (binary search over the scan, then the n-th set bit of that word) and fetches
// it — all 32 loads in flight together — and every lane runs the same pure add chain over the staged
// operands (unused entries are +0.0: no bit changes, the sums are never -0.0).
__device__ __forceinline__ void ks_phase1_ordered(const KsInst& I, int kw, unsigned kbit, double& weight, double& profit) {
    const int lane = threadIdx.x & 31;
    weight = 0.0;
    profit = 0.0;
    for (int base = 0; base < I.W; base += 32) {
        const int bw = base + lane;
        const unsigned mine = bw < I.W ? (I.one[bw] | (bw == kw ? kbit : 0u)) : 0u;
        const int cnt = __popc(mine);
        int incl = cnt;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const int o = __shfl_up_sync(0xffffffffu, incl, off);
            if (lane >= off) incl += o;
        }
        const int total = __shfl_sync(0xffffffffu, incl, 31);
        for (int lo = 0; lo < total; lo += 32) {
            const int ord = lo + lane;
            int L = 0;  // first lane whose inclusive count exceeds ord
#pragma unroll
            for (int step = 16; step > 0; step >>= 1) {
                const int probe = __shfl_sync(0xffffffffu, incl, L + step - 1);
                if (probe <= ord) L += step;
            }
            const unsigned word = __shfl_sync(0xffffffffu, mine, L & 31);
            const int before = __shfl_sync(0xffffffffu, incl - cnt, L & 31);
            double2 v = make_double2(0.0, 0.0);
            if (ord < total) {
                const int idx = ((base + L) << 5) + (int)__fns(word, 0, ord - before + 1);
                v = make_double2(I.w_o[idx], I.p_o[idx]);
            }
            I.st_in[lane] = v;
            __syncwarp();
#pragma unroll
            for (int j0 = 0; j0 < 32; j0 += 8) {
                double2 u[8];
#pragma unroll
                for (int j = 0; j < 8; j++) u[j] = I.st_in[j0 + j];
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    weight = __dadd_rn(weight, u[j].x);
                    profit = __dadd_rn(profit, u[j].y);
                }
            }
            __syncwarp();
        }
    }
}

__device__ __forceinline__ void ks_scan2(double& cw, double& cp) {
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const double ow = __shfl_up_sync(0xffffffffu, cw, off);
        const double op = __shfl_up_sync(0xffffffffu, cp, off);
        if (lane >= off) {
            cw += ow;
            cp += op;
        }
    }
}

// Exact mode (integer data, weights >= 0, sums below 2^52: every sum below is exact, so it has the
// reference's bits in any order, and the running weight is monotone).
// Forward: the greedy pass from rank `start` with sums (weight, profit), 32 ranks per warp scan.
__device__ __forceinline__ KsEval ks_walk_exact(const KsInst& I, int sfw, unsigned sfbit, int start, double weight,
                                                double profit, double w1, double p1) {
    const int lane = threadIdx.x & 31;
    const int n = I.n;
    for (int b = start >> 5; b < I.W; b++) {
        const int r = (b << 5) + lane;
        const unsigned m = ks_dec_word(I, b, sfw, sfbit);
        const bool und = r >= start && r < n && !((m >> lane) & 1u);
        const double wv = und ? I.w_s[r] : 0.0, pv = und ? I.p_s[r] : 0.0;
        double cw = wv, cp = pv;
        ks_scan2(cw, cp);
        const unsigned fail = __ballot_sync(0xffffffffu, und && !((weight + cw) <= I.limit));
        if (fail) {
            const int j = __ffs(fail) - 1;
            weight += __shfl_sync(0xffffffffu, cw - wv, j);
            profit += __shfl_sync(0xffffffffu, cp - pv, j);
            return ks_finish(I, w1, p1, weight, profit, (b << 5) + j);
        }
        weight += __shfl_sync(0xffffffffu, cw, 31);
        profit += __shfl_sync(0xffffffffu, cp, 31);
    }
    return ks_finish(I, w1, p1, weight, profit, n);
}

// Right child in exact mode: the fixed sums grow by the branching item (w1r = w1 + w_k), so the break
// item can only move to a LOWER rank than the parent's (sf).  (X, Y) = undecided prefix sums before
// rank sf (= parent's wb - w1, pb - p1).  Walk the 32-rank blocks backwards from sf: the failing
// items form a suffix of the undecided ranks, the first of them is the new break item.
__device__ __forceinline__ KsEval ks_right_exact(const KsInst& I, int sf, unsigned sfbit, double w1r, double p1r, double X,
                                                 double Y) {
    const int lane = threadIdx.x & 31;
    int cand = -1;
    double cand_w = 0.0, cand_p = 0.0;
    double Tw = X, Tp = Y;  // undecided prefix sums before the end of the range still to look at
    for (int b = sf >> 5; b >= 0; b--) {
        const int r = (b << 5) + lane;
        const unsigned m = I.dec[b];
        const bool und = r < sf && !((m >> lane) & 1u);
        const double wv = und ? I.w_s[r] : 0.0, pv = und ? I.p_s[r] : 0.0;
        double cw = wv, cp = pv;
        ks_scan2(cw, cp);
        const double base_w = Tw - __shfl_sync(0xffffffffu, cw, 31), base_p = Tp - __shfl_sync(0xffffffffu, cp, 31);
        const unsigned fail = __ballot_sync(0xffffffffu, und && !((w1r + (base_w + cw)) <= I.limit));
        const unsigned umask = __ballot_sync(0xffffffffu, und);
        if (fail) {
            const int j = __ffs(fail) - 1;
            cand = (b << 5) + j;
            cand_w = w1r + __shfl_sync(0xffffffffu, base_w + (cw - wv), j);
            cand_p = p1r + __shfl_sync(0xffffffffu, base_p + (cp - pv), j);
            if (umask & ((1u << j) - 1u)) break;  // an undecided item before it fits: nothing earlier fails
        } else if (umask) {
            break;
        }
        Tw = base_w;
        Tp = base_p;
    }
    if (cand >= 0) return ks_finish(I, w1r, p1r, cand_w, cand_p, cand);
    // everything before sf still fits: go on behind it
    return ks_walk_exact(I, sf >> 5, sfbit, sf + 1, w1r + X, p1r + Y, w1r, p1r);
}

// ---- the heap: SimpleMaxHeap<Node> (:494-547), same swaps, same ties -------------------------------
struct KsHeap {
    unsigned long long* sk;  // first HS entries (shared memory)
    int* sn;
    KsHeapEntry* g;  // the rest (global memory, indexed by absolute position)
    int HS;
    __device__ __forceinline__ void get(int i, unsigned long long& kv, int& nv) const {
        if (i < HS) {
            kv = sk[i];
            nv = sn[i];
        } else {
            const int4 raw = *reinterpret_cast<const int4*>(g + i);
            kv = ((unsigned long long)(unsigned)raw.y << 32) | (unsigned)raw.x;
            nv = raw.z;
        }
    }
    __device__ __forceinline__ void set(int i, unsigned long long kv, int nv) const {
        if (i < HS) {
            sk[i] = kv;
            sn[i] = nv;
        } else {
            int4 raw;
            raw.x = (int)(unsigned)kv;
            raw.y = (int)(unsigned)(kv >> 32);
            raw.z = nv;
            raw.w = 0;
            *reinterpret_cast<int4*>(g + i) = raw;
        }
    }
};

// Push (:502-513): the new entry climbs while it is strictly greater than its parent.  The ancestor
// path is known up front: lane t loads ancestor t, one ballot finds where the climb stops, the
// passed ancestors move down one step each.
__device__ __forceinline__ void ks_heap_push(const KsHeap& H, int& size, unsigned long long xk, int xn) {
    const int lane = threadIdx.x & 31;
    const int ci = size;
    const unsigned path = (unsigned)ci + 1u;  // 1-based position; ancestor t = path >> (t + 1)
    const unsigned a1 = lane < 31 ? (path >> (lane + 1)) : 0u;
    const bool valid = a1 >= 1u;
    unsigned long long ek = 0ULL;
    int en = 0;
    if (valid) H.get((int)a1 - 1, ek, en);
    const bool climbs = valid && xk > ek;
    const unsigned stop = __ballot_sync(0xffffffffu, !climbs);  // lane 31 never climbs: stop != 0
    const int k = __ffs(stop) - 1;                                // ancestors 0 .. k-1 move down
    if (lane < k) H.set((int)(path >> lane) - 1, ek, en);         // ancestor t takes the place of step t-1
    if (lane == 0) H.set((int)(path >> k) - 1, xk, xn);
    size = ci + 1;
    __syncwarp();
}

// Pop (:515-542), second half: the last entry (xk, xn) replaces the root and sinks among positions
// 0 .. last; a child must be strictly greater to pass, the right child strictly greater than the left.
// Every lane does the same work and stores the same values: no divergence on the path.
__device__ __forceinline__ void ks_heap_sink(const KsHeap& H, int last, unsigned long long xk, int xn) {
    int i = 0;
    while (true) {
        const int l = 2 * i + 1, r = l + 1;
        if (l > last) break;
        unsigned long long lk, rk = 0ULL;
        int ln, rn = 0;
        H.get(l, lk, ln);
        if (r <= last) H.get(r, rk, rn);  // rk = 0 otherwise: never strictly greater
        unsigned long long bk = xk;
        int largest = i, mv = xn;
        if (lk > bk) {
            largest = l;
            bk = lk;
            mv = ln;
        }
        if (rk > bk) {
            largest = r;
            bk = rk;
            mv = rn;
        }
        if (largest == i) break;
        H.set(i, bk, mv);
        i = largest;
    }
    H.set(i, xk, xn);
    __syncwarp();
}

__device__ __forceinline__ void ks_store_meta(KsMeta* dst, const KsEval& e, int var) {
    KsMeta m;
    m.bound = e.bound;
    m.weight = e.weight;
    m.frac = e.frac;
    m.w1 = e.w1;
    m.p1 = e.p1;
    m.wb = e.wb;
    m.pb = e.pb;
    m.frac_rank = e.frac_rank;
    m.break_rank = e.break_rank;
    m.flags = e.flags;
    m.var = var;
    *dst = m;
}

constexpr int KS_MAX_WARPS = 16;  // warps (instances in flight) per CTA, one CTA per SM

// shared memory per warp: heap cache (HS x 12 bytes), the popped node's masks (MS words), chain staging (1 KB)
__host__ __device__ inline size_t ks_warp_smem(int HS, int MS, int n_staged = 0) {
    return (((size_t)HS * 12 + 15) & ~(size_t)15) + (size_t)MS * 4 + 1024 + (size_t)n_staged * 16;
}
// A PAIR of warps per instance (few instances: latency matters) adds the helper warp's chain staging, the job
// the main warp posts and the result the helper returns.
struct KsJob {
    int kind;  // 0: no more instances, 1: evaluate the right child, 2: a new instance (k)
    int k, sf, kv;
    double w1, p1, wb, pb;
};
constexpr size_t KS_PAIR_EXTRA = 1024 + 64 + 128;
__host__ __device__ inline size_t ks_unit_smem(int HS, int MS, int n_staged, bool paired) {
    return ks_warp_smem(HS, MS, n_staged) + (paired ? KS_PAIR_EXTRA : 0);
}

// The instance as a warp sees it (pointers into the item tables, the staged masks, its own chain buffers).
__device__ __forceinline__ KsInst ks_inst(const KsParams& P, int k, const unsigned* sm_mask, double2* st_in, double2* st_out,
                                          const double* sm_items, int n_staged) {
    const int n = P.n;
    KsInst I;
    I.w_s = n_staged ? sm_items : P.w_s + (size_t)k * n;
    I.p_s = n_staged ? sm_items + n : P.p_s + (size_t)k * n;
    I.w_o = P.w_o + (size_t)k * n;
    I.p_o = P.p_o + (size_t)k * n;
    I.orig_s = P.orig_s + (size_t)k * n;
    I.capacity = P.cap[k];
    I.limit = __dadd_rn(I.capacity, KS_EPS);
    I.n = n;
    I.W = P.W;
    I.exact = P.exact[k];
    I.one = sm_mask;
    I.dec = sm_mask + P.W;
    I.st_in = st_in;
    I.st_out = st_out;
    return I;
}

// Right child (x_k = 1) of the node whose masks are staged: the fixed sums grow by item kv (:266-269).
__device__ __forceinline__ KsEval ks_right_child(const KsInst& I, int sf, int kv, double w1, double p1, double wb, double pb) {
    const int sfw = sf >> 5, kw = kv >> 5;
    const unsigned sfbit = 1u << (sf & 31), kbit = 1u << (kv & 31);
    double w1r, p1r;
    if (I.exact) {
        w1r = w1 + I.w_o[kv];
        p1r = p1 + I.p_o[kv];
    } else {
        ks_phase1_ordered(I, kw, kbit, w1r, p1r);
    }
    if (w1r > I.limit) return ks_early(w1r, p1r, I.n);
    return I.exact ? ks_right_exact(I, sf, sfbit, w1r, p1r, wb - w1, pb - p1)
                   : ks_eval_chain(I, sfw, sfbit, 0, w1r, p1r, w1r, p1r);
}

// PAIRED: two warps per instance.  The MAIN warp runs the search loop; for every node it expands it posts the
// right child (the long evaluation: all fixed items and the whole greedy pass again) to the HELPER warp and
// meanwhile sinks the heap and evaluates the left child itself; the two meet at a named barrier of their own
// (bar.sync 1 + pair, 64) before the commit.
template <int MAXW, bool PAIRED>
__global__ void __launch_bounds__(MAXW * 32, 1) knap_search_kernel(const KsParams P) {
    extern __shared__ __align__(16) unsigned char ks_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int unit = PAIRED ? warp >> 1 : warp;
    const bool helper = PAIRED && (warp & 1);
    const int HS = P.HS, MS = P.MS;
    const int n_staged = P.stage_items ? P.n : 0;
    unsigned char* mine_smem = ks_smem + (size_t)unit * ks_unit_smem(HS, MS, n_staged, PAIRED);
    unsigned char* pair_smem = mine_smem + ks_warp_smem(HS, MS, n_staged);  // PAIRED only
    double2* st_in = reinterpret_cast<double2*>(helper ? pair_smem : mine_smem);
    double2* st_out = st_in + 32;
    unsigned* sm_mask = reinterpret_cast<unsigned*>(mine_smem + 1024);
    unsigned long long* sk = reinterpret_cast<unsigned long long*>(mine_smem + 1024 + (size_t)MS * 4);
    int* sn = reinterpret_cast<int*>(sk + HS);
    double* sm_items = reinterpret_cast<double*>(mine_smem + ks_warp_smem(HS, MS));  // [w_s | p_s] when staged
    KsJob* job = reinterpret_cast<KsJob*>(pair_smem + 1024);
    KsEval* res = reinterpret_cast<KsEval*>(pair_smem + 1024 + 64);
    const int W = P.W, n = P.n, cap_nodes = P.node_cap;
    auto pair_sync = [&]() { asm volatile("bar.sync %0, 64;" ::"r"(1 + unit) : "memory"); };

    if (helper) {
        KsInst I = ks_inst(P, 0, sm_mask, st_in, st_out, sm_items, n_staged);
        for (;;) {
            pair_sync();  // a job is posted
            const KsJob j = *job;
            if (j.kind == 0) return;
            if (j.kind == 2) {
                I = ks_inst(P, j.k, sm_mask, st_in, st_out, sm_items, n_staged);
                continue;
            }
            const KsEval e = ks_right_child(I, j.sf, j.kv, j.w1, j.p1, j.wb, j.pb);
            if (lane == 0) *res = e;
            pair_sync();  // the result is in place
        }
    }

    for (;;) {
        int q = 0;
        if (lane == 0) q = atomicAdd(P.next, 1);
        q = __shfl_sync(0xffffffffu, q, 0);
        if (q >= P.n_todo) {
            if (PAIRED) {
                if (lane == 0) job->kind = 0;
                pair_sync();
            }
            break;
        }
        const int k = P.todo[q];
        const int slot = P.slots[q];

        if (n_staged) {  // the greedy pass reads these 2 n doubles over and over: keep them next to the ALUs
            const double* gw = P.w_s + (size_t)k * n;
            const double* gp = P.p_s + (size_t)k * n;
            for (int i = lane; i < n; i += 32) {
                sm_items[i] = gw[i];
                sm_items[n + i] = gp[i];
            }
            __syncwarp();
        }
        const KsInst I = ks_inst(P, k, sm_mask, st_in, st_out, sm_items, n_staged);
        if (PAIRED) {
            if (lane == 0) {
                job->kind = 2;
                job->k = k;
            }
            pair_sync();
        }

        KsState* S = P.state + k;
        KsHeap H;
        H.sk = sk;
        H.sn = sn;
        H.g = P.heap + (size_t)slot * cap_nodes;
        H.HS = HS;
        int* fstack = P.free_stack + (size_t)slot * cap_nodes;
        KsMeta* meta = P.meta + (size_t)slot * cap_nodes;
        unsigned* masks = P.masks + (size_t)slot * cap_nodes * MS;
        unsigned* bmask = P.best_masks + (size_t)k * 2 * W;
        KsRec* trace = P.trace ? P.trace + (size_t)slot * P.trace_cap : nullptr;

        double best = S->best;
        long long evals = S->evals, pops = S->pops;
        const long long pops_at_start = pops;
        int size = S->heap_size, free_top = S->free_top, next_fresh = S->next_fresh;
        int status = KS_RUNNING, tcount = 0;
        __syncwarp();

        auto alloc_node = [&]() -> int {
            int id = -1;
            if (free_top > 0) id = fstack[--free_top];
            else if (next_fresh < cap_nodes) id = next_fresh++;
            return id;
        };
        auto free_node = [&](int id) {
            if (lane == 0) fstack[free_top] = id;
            free_top++;
            __syncwarp();
        };
        // the incumbent's masks: the staged node's, plus the branching item for a child
        auto snapshot_best = [&](int kw, unsigned kbit, int sfw, unsigned sfbit, const KsEval& e, int var, int kind) {
            for (int b = lane; b < W; b += 32) {
                bmask[b] = I.one[b] | (b == kw ? kbit : 0u);
                bmask[W + b] = I.dec[b] | (b == sfw ? sfbit : 0u);
            }
            if (lane == 0) {
                ks_store_meta(&S->best_meta, e, var);
                S->best_kind = kind;
            }
        };
        auto record = [&](int kind, int node, int code, int var, int child, const KsEval& e) {
            if (lane == 0) {
                KsRec r;
                r.kind = kind;
                r.node = node;
                r.code = code;
                r.var = var;
                r.child = child;
                r.frac_rank = e.frac_rank;
                r.break_rank = e.break_rank;
                r.flags = e.flags;
                r.bound = e.bound;
                r.weight = e.weight;
                r.frac = e.frac;
                r.pad = 0.0;
                trace[tcount] = r;
            }
            tcount++;
        };

        if (!S->started) {
            // root: all undecided (:101-113)
            const int id = alloc_node();  // 0: a fresh pool
            unsigned* mk = masks + (size_t)id * MS;
            for (int b = lane; b < MS; b += 32) {
                mk[b] = 0u;
                sm_mask[b] = 0u;
            }
            __syncwarp();
            const KsEval e = I.exact ? ks_walk_exact(I, -1, 0u, 0, 0.0, 0.0, 0.0, 0.0)
                                     : ks_eval_chain(I, -1, 0u, 0, 0.0, 0.0, 0.0, 0.0);
            if (lane == 0) ks_store_meta(&meta[id], e, -1);
            evals = 1;
            ks_heap_push(H, size, ks_key(e.bound), id);
        } else {
            for (int i = lane; i < min(size, HS); i += 32) {
                const KsHeapEntry he = H.g[i];
                sk[i] = he.key;
                sn[i] = he.n;
            }
            __syncwarp();
        }

        long long pc[8] = {0, 0, 0, 0, 0, 0, 0, 0}, c0 = 0;
        const bool prof = P.prof != nullptr && k == 0;
#define KS_TICK(slot)                      \
    if (prof) {                            \
        const long long c1 = clock64();    \
        pc[slot] += c1 - c0;               \
        c0 = c1;                           \
    }
        while (size > 0) {
            if ((trace && tcount + 3 > P.trace_cap) || (P.pop_budget > 0 && pops - pops_at_start >= P.pop_budget)) {
                status = KS_PAUSED;
                break;
            }
            if (prof) c0 = clock64();
            // ---- pop (:120): take the root; its node record and masks are requested right away and
            // arrive while the last entry sinks
            unsigned long long tk, xk;
            int tn, xn;
            H.get(0, tk, tn);
            const KsMeta M = meta[tn];
            {
                const unsigned* src = masks + (size_t)tn * MS;
                for (int c = lane; c < (MS >> 2); c += 32)
                    __pipeline_memcpy_async(sm_mask + 4 * c, src + 4 * c, 16);
                __pipeline_commit();
            }
            const int li = size - 1;
            H.get(li, xk, xn);
            size = li;
            if (!PAIRED && li > 0) ks_heap_sink(H, li - 1, xk, xn);
            pops++;
            __pipeline_wait_prior(0);
            __syncwarp();
            if (PAIRED) {
                // the right child goes to the helper as soon as the node's record and masks are here; the heap
                // sinks (and the left child is evaluated) while it works
                if (!(M.bound <= best + KS_EPS) && M.frac_rank >= 0) {
                    if (lane == 0) {
                        job->kind = 1;
                        job->sf = M.frac_rank;
                        job->kv = I.orig_s[M.frac_rank];
                        job->w1 = M.w1;
                        job->p1 = M.p1;
                        job->wb = M.wb;
                        job->pb = M.pb;
                    }
                    pair_sync();
                }
                if (li > 0) ks_heap_sink(H, li - 1, xk, xn);
            }
            KS_TICK(0)
            if (M.bound <= best + KS_EPS) {  // :124 (node.Bound is the stored relaxation's)
                free_node(tn);
                continue;
            }
            KsEval pe;
            pe.bound = M.bound;
            pe.weight = M.weight;
            pe.frac = M.frac;
            pe.w1 = M.w1;
            pe.p1 = M.p1;
            pe.wb = M.wb;
            pe.pb = M.pb;
            pe.frac_rank = M.frac_rank;
            pe.break_rank = M.break_rank;
            pe.flags = M.flags;
            if (M.frac_rank < 0) {  // :147-177: an integer candidate (or infeasible)
                int closed = 3;
                if (M.weight <= I.limit) {
                    if (M.bound > best + KS_EPS) {
                        best = M.bound;
                        snapshot_best(-1, 0u, -1, 0u, pe, M.var, 1);
                        closed = 1;
                    } else {
                        closed = 2;
                    }
                }
                if (trace) record(0, tn, closed, M.var, -1, pe);
                __syncwarp();
                free_node(tn);
                continue;
            }
            if (trace) record(0, tn, 0, M.var, -1, pe);
            // branch on the fractional item (:180)
            const int sf = M.frac_rank;
            const int kv = I.orig_s[sf];
            const int sfw = sf >> 5, kw = kv >> 5;
            const unsigned sfbit = 1u << (sf & 31), kbit = 1u << (kv & 31);
            KsEval ev[2];
            KS_TICK(1)
            // left child, x_k = 0: same fixed items; the greedy pass resumes behind the parent's break item
            ev[0] = I.exact ? ks_walk_exact(I, sfw, sfbit, sf + 1, M.wb, M.pb, M.w1, M.p1)
                            : ks_eval_chain(I, sfw, sfbit, sf + 1, M.wb, M.pb, M.w1, M.p1);
            KS_TICK(2)
            // right child, x_k = 1
            if (PAIRED) {
                pair_sync();  // the helper's result
                ev[1] = *res;
            } else {
                ev[1] = ks_right_child(I, sf, kv, M.w1, M.p1, M.wb, M.pb);
            }
            KS_TICK(3)
            bool overflow = false;
#pragma unroll
            for (int side = 0; side < 2; side++) {
                const KsEval& e = ev[side];
                evals++;
                int decision, pushed = -1;
                if (e.weight > I.limit) {
                    decision = LPX_KN_INFEASIBLE;
                } else if (e.bound > best + KS_EPS) {
                    if (e.flags & KS_F_ALLINT) {  // allInt && feasible (:228-238, :288-298)
                        decision = LPX_KN_CANDIDATE_INT;
                        best = e.bound;
                        snapshot_best(side ? kw : -1, kbit, sfw, sfbit, e, kv, 2);
                    } else {
                        decision = LPX_KN_PUSHED;
                        const int id = alloc_node();
                        if (id < 0) {
                            overflow = true;
                            break;
                        }
                        unsigned* mk = masks + (size_t)id * MS;
                        for (int b = lane; b < W; b += 32) {
                            mk[b] = I.one[b] | ((side && b == kw) ? kbit : 0u);
                            mk[W + b] = I.dec[b] | (b == sfw ? sfbit : 0u);
                        }
                        if (lane == 0) ks_store_meta(&meta[id], e, kv);
                        ks_heap_push(H, size, ks_key(e.bound), id);
                        pushed = id;
                    }
                } else {
                    decision = LPX_KN_DROPPED;
                }
                if (trace) record(1, pushed, decision, kv, side, e);
            }
            if (overflow) {
                status = KS_OVERFLOW;
                break;
            }
            __syncwarp();
            KS_TICK(4)
            free_node(tn);
            KS_TICK(5)
        }
#undef KS_TICK
        if (prof && lane == 0)
            for (int t = 0; t < 8; t++) P.prof[t] += pc[t];
        if (status == KS_RUNNING) status = KS_DONE;
        if (status == KS_PAUSED) {  // the global array becomes the whole truth again
            for (int i = lane; i < min(size, HS); i += 32) {
                KsHeapEntry he;
                he.key = sk[i];
                he.n = sn[i];
                he.pad = 0;
                H.g[i] = he;
            }
        }
        if (lane == 0) {
            S->best = best;
            S->evals = evals;
            S->pops = pops;
            S->heap_size = size;
            S->free_top = free_top;
            S->next_fresh = next_fresh;
            S->status = status;
            S->started = 1;
            S->trace_count = tcount;
        }
        __syncwarp();
    }
}

// =====================================================================================================
// host side
// =====================================================================================================
namespace {

int cmp_double_h(double a, double b) {  // double.CompareTo
    if (a < b) return -1;
    if (a > b) return 1;
    if (a == b) return 0;
    if (std::isnan(a)) return std::isnan(b) ? 0 : -1;
    return 1;
}

struct DevBuf {  // plain cudaMalloc arena, freed on scope exit
    void* p = nullptr;
    ~DevBuf() {
        if (p) cudaFree(p);
    }
    int alloc(size_t bytes) {
        if (p) cudaFree(p);
        p = nullptr;
        LPX_CUDA(cudaMalloc(&p, bytes ? bytes : 16));
        return LPX_OK;
    }
};

// The node arena of the last call is kept: cudaMalloc / cudaFree of gigabytes per call cost more than a search.
struct ArenaCache {
    void* p = nullptr;
    size_t bytes = 0;
};
ArenaCache& arena_cache() {
    static ArenaCache c;
    return c;
}

struct TraceNode {
    std::vector<signed char> assigned;
    std::vector<int> label;
};

}  // namespace

void knapsack_dev_release_cache() {
    ArenaCache& c = arena_cache();
    if (c.p) cudaFree(c.p);
    c.p = nullptr;
    c.bytes = 0;
}

int knapsack_search_device(int count, int n, const double* profit, const double* weight, const double* capacity,
                           const lpx_options& opt, int* found, double* best_value, int* best_x, long long* n_evals,
                           long long* n_pops, int* rank_order, lpx_knap_pop_fn on_pop, void* user) {
    Runtime& r = rt();
    cudaStream_t s = r.stream;
    const int W = (n + 31) / 32;
    const size_t cn = (size_t)count * n;
    const bool force_ordered = opt.knap_ordered_sums != 0;

    // ---- item tables: ratio ordering (:75-79, stable OrderByDescending(Ratio).ThenByDescending(Profit)) ----
    std::vector<double> h_ws(cn), h_ps(cn);
    std::vector<int> h_orig(cn), h_exact(count);
    auto prepare = [&](int k) {
        const double* p = profit + (size_t)k * n;
        const double* w = weight + (size_t)k * n;
        std::vector<int> idx(n);
        std::vector<double> ratio(n);  // Item.Ratio (:19), once per item instead of twice per comparison
        for (int i = 0; i < n; i++) {
            idx[i] = i;
            ratio[i] = w[i] > 0 ? p[i] / w[i] : std::numeric_limits<double>::infinity();
        }
        std::stable_sort(idx.begin(), idx.end(), [&](int a, int b) {
            const int c = cmp_double_h(ratio[a], ratio[b]);
            if (c != 0) return c > 0;
            return cmp_double_h(p[a], p[b]) > 0;
        });
        // exact mode only where it is provably exact: integer data, weights >= 0, absolute sums below 2^52
        double aw = 0, ap = 0;
        bool ok = !force_ordered;
        for (int i = 0; i < n && ok; i++) {
            ok = std::isfinite(w[i]) && std::isfinite(p[i]) && w[i] == std::nearbyint(w[i]) && p[i] == std::nearbyint(p[i]) &&
                 w[i] >= 0.0;
            aw += std::fabs(w[i]);
            ap += std::fabs(p[i]);
        }
        h_exact[k] = (ok && aw < 4503599627370496.0 && ap < 4503599627370496.0) ? 1 : 0;
        for (int q = 0; q < n; q++) {
            h_ws[(size_t)k * n + q] = w[idx[q]];
            h_ps[(size_t)k * n + q] = p[idx[q]];
            h_orig[(size_t)k * n + q] = idx[q];
        }
        if (rank_order && k == 0)
            for (int q = 0; q < n; q++) rank_order[q] = idx[q];
    };
    {   // instances are independent: sort them on a few host threads (592 x 2000 items took as long as the search)
        unsigned hw = std::thread::hardware_concurrency();
        const int nthreads = (int)std::max(1u, std::min({hw ? hw / 2 : 1u, 8u, (unsigned)((count + 15) / 16)}));
        if (nthreads <= 1) {
            for (int k = 0; k < count; k++) prepare(k);
        } else {
            std::atomic<int> next(0);
            auto worker = [&]() {
                for (;;) {
                    const int k0 = next.fetch_add(8);
                    if (k0 >= count) break;
                    for (int k = k0; k < std::min(count, k0 + 8); k++) prepare(k);
                }
            };
            std::vector<std::thread> pool;
            for (int t = 1; t < nthreads; t++) pool.emplace_back(worker);
            worker();
            for (std::thread& th : pool) th.join();
        }
    }
    double* d_items = ws_dev_as<double>(WS_KN_ITEMS, cn * 4 + count);
    int* d_orig = ws_dev_as<int>(WS_KN_ASSIGN, cn + count);
    if (!d_items || !d_orig) return LPX_E_CUDA;
    double *d_ws = d_items, *d_ps = d_ws + cn, *d_wo = d_ps + cn, *d_po = d_wo + cn, *d_cap = d_po + cn;
    int* d_exact = d_orig + cn;
    LPX_CUDA(cudaMemcpyAsync(d_ws, h_ws.data(), cn * 8, cudaMemcpyHostToDevice, s));
    LPX_CUDA(cudaMemcpyAsync(d_ps, h_ps.data(), cn * 8, cudaMemcpyHostToDevice, s));
    LPX_CUDA(cudaMemcpyAsync(d_wo, weight, cn * 8, cudaMemcpyHostToDevice, s));
    LPX_CUDA(cudaMemcpyAsync(d_po, profit, cn * 8, cudaMemcpyHostToDevice, s));
    LPX_CUDA(cudaMemcpyAsync(d_cap, capacity, (size_t)count * 8, cudaMemcpyHostToDevice, s));
    LPX_CUDA(cudaMemcpyAsync(d_orig, h_orig.data(), cn * 4, cudaMemcpyHostToDevice, s));
    LPX_CUDA(cudaMemcpyAsync(d_exact, h_exact.data(), (size_t)count * 4, cudaMemcpyHostToDevice, s));

    // ---- per-instance state and incumbent snapshots --------------------------------------------------
    std::vector<KsState> h_state(count);
    // cached workspaces: a small instance should not pay three cudaMalloc / cudaFree pairs per call
    int rc;
    KsState* d_state = ws_dev_as<KsState>(WS_KN_OUT, count);
    unsigned* d_best = ws_dev_as<unsigned>(WS_KN_AUX, (size_t)count * 2 * W);
    int* d_todo = ws_dev_as<int>(WS_KN_EXACT, (size_t)2 * count + 1);  // queue (instances), their arena slots, head
    if (!d_state || !d_best || !d_todo) return LPX_E_CUDA;
    int* d_slots = d_todo + count;
    int* d_next = d_todo + 2 * count;

    const bool tracing = on_pop != nullptr;
    if (tracing && count != 1) {
        set_error("knapsack: the pop callback is delivered for single-instance calls only");
        return LPX_E_BAD_ARGS;
    }
    const int trace_cap = tracing ? 3 * 4096 : 0;
    const int MS = (2 * W + 3) & ~3;  // mask words per node, 16-byte granules for the asynchronous copy
    const size_t node_bytes = sizeof(KsHeapEntry) + 4 + sizeof(KsMeta) + (size_t)MS * 4;

    size_t free_b = 0, total_b = 0;
    LPX_CUDA(cudaMemGetInfo(&free_b, &total_b));
    ArenaCache& ac = arena_cache();
    const size_t avail = free_b + ac.bytes;  // the cached arena can be reused or regrown
    size_t budget1 = std::min(avail / 2, (size_t)16 << 30);
    if (const char* e = getenv("LPX_KNAP_POOL_MB")) budget1 = std::min(budget1, (size_t)std::max(16, atoi(e)) << 20);  // profiling aid

    std::vector<int> todo(count);
    for (int k = 0; k < count; k++) todo[k] = k;
    std::unordered_map<int, TraceNode> tnodes;  // tracing: live nodes of the (single) traced instance
    std::vector<long long> pop_index(count, 0);

    // One CTA per SM; its warps (one instance each) share the SM's shared memory: few instances get one
    // SM and a deep heap cache each, a large batch 16 warps per SM with 1023 cached entries each.
    const int max_dyn = r.smem_optin;
    // units per CTA (a unit = one instance in flight: one warp, or a main + helper pair of warps when there are
    // at most four instances per SM and a single instance's latency is what the caller waits for)
    const bool no_pairs = getenv("LPX_KNAP_NO_PAIRS") != nullptr || opt.knap_warps == 1;
    auto shape_for = [&](int nt, int& wpb, int& HS, bool& stage, bool& paired) {
        wpb = std::max(1, std::min(KS_MAX_WARPS, (nt + r.sms - 1) / r.sms));
        if (opt.knap_warps == 2) wpb = std::min(wpb, 4);  // forced pairs: at most four per CTA (the queue does the rest)
        paired = wpb <= 4 && !no_pairs;
        // one or two instances per SM: at most ~128 KB of the SM's 256 KB, so that their item tables (72 KB per
        // instance at 2000 items) stay in L1; more instances than that do not fit L1 anyway
        const size_t budget = wpb <= 2 ? (size_t)128 * 1024 : (size_t)max_dyn;
        stage = wpb <= 2 && (size_t)n * 16 <= (size_t)48 * 1024;
        HS = 8191;
        while (HS > 31 && (size_t)wpb * ks_unit_smem(HS, MS, stage ? n : 0, paired) > budget) HS >>= 1;
    };
    LPX_CUDA(cudaFuncSetAttribute(knap_search_kernel<8, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_dyn));
    LPX_CUDA(cudaFuncSetAttribute(knap_search_kernel<8, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_dyn));
    LPX_CUDA(cudaFuncSetAttribute(knap_search_kernel<KS_MAX_WARPS, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  max_dyn));

    for (int pass = 0; pass < 2 && !todo.empty(); pass++) {
        // pass 0: every instance with an equal share of the budget; pass 1: the few whose pool overflowed,
        // again from the root, sharing (almost) all free memory
        const size_t budget = pass == 0 ? budget1 : avail - avail / 8;
        const int nt = (int)todo.size();
        size_t cap_nodes = budget / nt / node_bytes;
        cap_nodes = std::min(cap_nodes, pass == 0 ? (size_t)1 << 20 : (size_t)1 << 26);
        if (cap_nodes < 64) {
            set_error("knapsack: not enough device memory for the node pools");
            return LPX_E_CAPACITY;
        }
        const size_t need = (size_t)nt * cap_nodes * node_bytes + (size_t)nt * trace_cap * sizeof(KsRec) + 256;
        if (ac.bytes < need) {
            if (ac.p) cudaFree(ac.p);
            ac.p = nullptr;
            ac.bytes = 0;
            LPX_CUDA(cudaMalloc(&ac.p, need));
            ac.bytes = need;
        }
        unsigned char* base = (unsigned char*)ac.p;
        KsParams P;
        std::memset(&P, 0, sizeof P);
        P.w_s = d_ws;
        P.p_s = d_ps;
        P.w_o = d_wo;
        P.p_o = d_po;
        P.cap = d_cap;
        P.orig_s = d_orig;
        P.exact = d_exact;
        P.n = n;
        P.W = W;
        P.MS = MS;
        P.count = count;
        P.todo = d_todo;
        P.slots = d_slots;
        P.n_todo = nt;
        P.next = d_next;
        P.node_cap = (int)cap_nodes;
        size_t off = 0;
        P.heap = (KsHeapEntry*)(base + off);
        off += (size_t)nt * cap_nodes * sizeof(KsHeapEntry);
        P.masks = (unsigned*)(base + off);
        off += (size_t)nt * cap_nodes * MS * 4;
        P.meta = (KsMeta*)(base + off);
        off += (size_t)nt * cap_nodes * sizeof(KsMeta);
        P.trace = trace_cap ? (KsRec*)(base + off) : nullptr;
        off += (size_t)nt * trace_cap * sizeof(KsRec);
        P.free_stack = (int*)(base + off);
        off += (size_t)nt * cap_nodes * 4;
        P.best_masks = d_best;
        P.state = d_state;
        P.trace_cap = trace_cap;
        // the queue of one launch: `queue[q]` = instance, `qslot[q]` = its arena slot (fixed for the pass)
        std::vector<int> queue = todo, qslot(nt);
        for (int q = 0; q < nt; q++) qslot[q] = q;
        DevBuf b_prof;
        static const bool want_prof = getenv("LPX_KNAP_PROF") != nullptr;
        if (want_prof) {
            if ((rc = b_prof.alloc(8 * 8)) != LPX_OK) return rc;
            LPX_CUDA(cudaMemsetAsync(b_prof.p, 0, 64, s));
            P.prof = (long long*)b_prof.p;
        }

        for (int q = 0; q < nt; q++) {
            KsState& st = h_state[todo[q]];
            std::memset(&st, 0, sizeof st);
            st.best = -std::numeric_limits<double>::infinity();
        }
        LPX_CUDA(cudaMemcpyAsync(d_state, h_state.data(), (size_t)count * sizeof(KsState), cudaMemcpyHostToDevice, s));
        if (tracing) {
            tnodes.clear();
            std::fill(pop_index.begin(), pop_index.end(), 0);
        }

        // Slot q <-> todo[q] stays fixed while the pass resumes paused instances.
        // Paused instances keep their slot: the kernel is relaunched over the same todo list and instances
        // that are already done return at once.
        std::vector<KsRec> h_trace;
        while (true) {
            // One launch over the current queue.  A large batch runs one warp per instance, 16 per SM, with a pop
            // budget: the few long searches that outlive it are paused and resumed in a short queue — a pair of
            // warps and an SM's heap cache each — instead of crawling on as lone warps after everybody else is done.
            const int nq = (int)queue.size();
            int wpb = 1, HS = 31;
            bool stage = false, paired = false;
            shape_for(nq, wpb, HS, stage, paired);
            P.HS = HS;
            P.stage_items = stage ? 1 : 0;
            P.n_todo = nq;
            P.pop_budget = (!paired && !tracing) ? 16384 : 0;
            const size_t smem = (size_t)wpb * ks_unit_smem(HS, MS, stage ? n : 0, paired);
            LPX_CUDA(cudaMemcpyAsync(d_todo, queue.data(), (size_t)nq * 4, cudaMemcpyHostToDevice, s));
            LPX_CUDA(cudaMemcpyAsync(d_slots, qslot.data(), (size_t)nq * 4, cudaMemcpyHostToDevice, s));
            LPX_CUDA(cudaMemsetAsync(d_next, 0, 4, s));
            const int threads = wpb * (paired ? 64 : 32), grid = std::max(1, std::min((nq + wpb - 1) / wpb, r.sms));
            if (paired) knap_search_kernel<8, true><<<grid, threads, smem, s>>>(P);            // <= 4 pairs per CTA
            else if (wpb <= 8) knap_search_kernel<8, false><<<grid, threads, smem, s>>>(P);    // up to 255 registers
            else knap_search_kernel<KS_MAX_WARPS, false><<<grid, threads, smem, s>>>(P);
            LPX_CUDA(cudaGetLastError());
            count_launch();
            LPX_CUDA(cudaMemcpyAsync(h_state.data(), d_state, (size_t)count * sizeof(KsState), cudaMemcpyDeviceToHost, s));
            LPX_CUDA(cudaStreamSynchronize(s));
            std::vector<int> nqueue, nslot;
            for (int q = 0; q < nq; q++) {
                const int k = queue[q];
                KsState& st = h_state[k];
                if (tracing && st.trace_count > 0) {
                    // replay the records: rebuild assignment vectors and labels, call back in order
                    h_trace.resize(st.trace_count);
                    LPX_CUDA(cudaMemcpy(h_trace.data(), P.trace + (size_t)qslot[q] * trace_cap,
                                        (size_t)st.trace_count * sizeof(KsRec), cudaMemcpyDeviceToHost));
                    if (tnodes.empty() && pop_index[k] == 0) {
                        TraceNode root;
                        root.assigned.assign(n, -1);
                        root.label = {0};
                        tnodes[0] = root;
                    }
                    auto fill = [&](lpx_knap_eval& ev, const KsRec& rec, long long pidx, int child, int decision,
                                    const signed char* assigned) {
                        ev.pop_index = (int)pidx;
                        ev.child = child;
                        ev.var = rec.var;
                        ev.bound = rec.bound;
                        ev.weight = rec.weight;
                        ev.frac_rank = rec.frac_rank;
                        ev.frac = rec.frac;
                        ev.break_rank = rec.break_rank;
                        ev.decision = decision;
                        ev.assigned = assigned;
                    };
                    for (int t = 0; t < st.trace_count;) {
                        const KsRec& pr = h_trace[t];
                        TraceNode node = tnodes[pr.node];
                        tnodes.erase(pr.node);
                        const long long pidx = pop_index[k]++;
                        lpx_knap_pop pop;
                        std::memset(&pop, 0, sizeof pop);
                        pop.pop_index = (int)pidx;
                        pop.label_len = (int)node.label.size();
                        pop.label = node.label.data();
                        pop.closed = pr.code;
                        fill(pop.relax, pr, pidx, -1, LPX_KN_ROOT, node.assigned.data());
                        if (pr.code != 0) {
                            on_pop(&pop, nullptr, nullptr, user);
                            t += 1;
                            continue;
                        }
                        lpx_knap_eval evs[2];
                        TraceNode ch[2];
                        for (int side = 0; side < 2; side++) {
                            const KsRec& er = h_trace[t + 1 + side];
                            ch[side].assigned = node.assigned;
                            ch[side].assigned[er.var] = (signed char)side;
                            if (node.label.size() == 1 && node.label[0] == 0) ch[side].label = {side + 1};
                            else {
                                ch[side].label = node.label;
                                ch[side].label.push_back(side + 1);
                            }
                            fill(evs[side], er, pidx, side, er.code, ch[side].assigned.data());
                        }
                        on_pop(&pop, &evs[0], &evs[1], user);
                        for (int side = 0; side < 2; side++) {
                            const KsRec& er = h_trace[t + 1 + side];
                            if (er.code == LPX_KN_PUSHED) tnodes[er.node] = std::move(ch[side]);
                        }
                        t += 3;
                    }
                }
                if (st.status == KS_PAUSED) {
                    nqueue.push_back(k);
                    nslot.push_back(qslot[q]);
                }
            }
            if (nqueue.empty()) break;
            queue.swap(nqueue);
            qslot.swap(nslot);
        }
        if (want_prof) {
            long long h[8];
            LPX_CUDA(cudaMemcpy(h, b_prof.p, 64, cudaMemcpyDeviceToHost));
            const double pp = (double)std::max(1LL, h_state[0].pops);
            fprintf(stderr, "[knap prof] instance 0: %lld pops; cycles per pop: heap pop %.0f, meta+branch %.0f, left %.0f, "
                            "right %.0f (+ ordered fixed sums %.0f), commit %.0f, free %.0f\n", h_state[0].pops, h[0] / pp, h[1] / pp,
                    h[2] / pp, h[3] / pp, h[6] / pp, h[4] / pp, h[5] / pp);
        }
        std::vector<int> again;
        for (int q = 0; q < nt; q++)
            if (h_state[todo[q]].status == KS_OVERFLOW) again.push_back(todo[q]);
        if (!again.empty() && (pass == 1 || tracing)) {
            set_error("knapsack: node pool exhausted (the search tree does not fit in device memory)");
            return LPX_E_CAPACITY;
        }
        todo = again;
    }
    if (ac.bytes > ((size_t)20 << 30)) knapsack_dev_release_cache();  // keep at most 20 GB parked between calls (a
    // first-pass arena is <= 16 GB: re-allocating it on every call cost more than the search: 0.9 s vs 0.36 s)

    // ---- results: best value and x* from the incumbent snapshot -------------------------------------
    std::vector<unsigned> h_best((size_t)count * 2 * W);
    LPX_CUDA(cudaMemcpy(h_best.data(), d_best, h_best.size() * 4, cudaMemcpyDeviceToHost));
    for (int k = 0; k < count; k++) {
        const KsState& st = h_state[k];
        const bool none = std::isinf(st.best) && st.best < 0;
        if (found) found[k] = none ? 0 : 1;
        if (best_value) best_value[k] = st.best;
        if (n_evals) n_evals[k] = st.evals;
        if (n_pops) n_pops[k] = st.pops;
        if (best_x) {
            int* bx = best_x + (size_t)k * n;
            for (int i = 0; i < n; i++) bx[i] = 0;
            if (!none) {
                // relaxed[] of the incumbent's ComputeRelaxation, then the reference's rounding rule
                const unsigned* one = h_best.data() + (size_t)k * 2 * W;
                const unsigned* dec = one + W;
                const int* ord = h_orig.data() + (size_t)k * n;
                const KsMeta& M = st.best_meta;
                std::vector<double> relaxed(n, 0.0);
                for (int i = 0; i < n; i++)
                    if ((one[i >> 5] >> (i & 31)) & 1u) relaxed[i] = 1.0;
                if (!(M.flags & KS_F_EARLY)) {
                    for (int q = 0; q < M.break_rank && q < n; q++)
                        if (!((dec[q >> 5] >> (q & 31)) & 1u)) relaxed[ord[q]] = 1.0;
                    if (M.frac_rank >= 0) relaxed[ord[M.frac_rank]] = M.frac;
                }
                for (int i = 0; i < n; i++)
                    bx[i] = st.best_kind == 1 ? (relaxed[i] >= 0.5 ? 1 : 0) : (int)std::nearbyint(relaxed[i]);
            }
        }
    }
    return LPX_OK;
}

}  // namespace lpx
