// xg_overlap.cu -- overlap weights: target faces x source faces -> CSR of areas.
//
// Replaces UnstructuredGrid2d.overlap -> CellTree2d.intersect_faces
// (/root/reference/xugrid/regrid/unstructured.py:109-135: build_bboxes ->
// locate_boxes (count pass, cumsum, fill pass) -> separating-axis filter ->
// Sutherland-Hodgman area -> keep area > 0) and MatrixCSR.from_triplet
// (/root/reference/xugrid/core/sparse.py:61-78,119-127).
//
// Pipeline (all on the caller's stream, ONE host synchronisation at the end):
//   pack_mesh(targets)   targets in Morton order, self-contained records
//   k_traverse           one warp walks the wide BVH ONCE for a tile of 32
//                        Morton-consecutive targets (lane = target).  Child
//                        boxes are loaded once per warp, coalesced, and tested
//                        by all lanes from shared memory; hits are compacted
//                        with warp votes into a per-warp queue that is flushed
//                        to a global pair arena (one atomic per ~100 pairs)
//   k_sat                one thread per candidate pair: separating-axis filter;
//                        survivors appended to a dense list
//   k_clip_survivors     one thread per survivor: Sutherland-Hodgman area,
//                        warp-uniform code; a positive area is written straight
//                        into the target's provisional row (rows in sorted target
//                        order, sized by the candidate counts of the walk)
//   cub ExclusiveSum     CSR row pointer in the caller's row order
//   k_move_rows (xg_overlap_finish)  provisional row -> caller's row, entries
//                        ordered by ascending source index on the way
// The arena is sized from a density guess (16 candidates per target); if the
// walk finds more, the job is re-run once with the exact size the first walk
// counted (the reference always walks twice: count pass + fill pass).
#include <cstdlib>

#include <cub/device/device_scan.cuh>
#include <cub/iterator/transform_input_iterator.cuh>

#include "xg_common.cuh"
#include "xg_geom.cuh"
#include "xg_pack.cuh"

namespace xg {

constexpr int TRAV_WARPS = 8;
constexpr int TRAV_STACK = 256;
#ifndef XG_TRAV_WIN
#define XG_TRAV_WIN 256
#endif
constexpr int TRAV_WIN = XG_TRAV_WIN;           // pairs per window of the arena (reserved ahead by a warp)
constexpr int TRAV_BLOCKS_PER_SM = 4;           // resident blocks of the persistent walk
constexpr int PAIRS_PER_TARGET_GUESS = 16;

// device counters of one job
// CTR_PAIRS / CTR_SURV count arena SLOTS (reserved windows, unused tails included);
// CTR_CAND / CTR_NSURV the candidate pairs / survivors themselves
// CTR_BASE: offset of the job's first CSR entry (0, or the last row pointer of the job before
// it when jobs are chained); CTR_NNZ_END: row pointer behind the job's last row;
// CTR_FLAGS: bit 0 = an entry would have landed beyond the caller's arrays
enum { CTR_VISITS = 0, CTR_TESTS = 1, CTR_PAIRS = 2, CTR_SURV = 3, CTR_CAND = 4, CTR_HUGE = 5, CTR_LONG = 6, CTR_NSURV = 7,
       CTR_BASE = 8, CTR_NNZ_END = 9, CTR_FLAGS = 10, CTR_N = 12 };

struct TraverseParams {
    const Box *boxes[MAX_LEVELS + 1];  // [0] = leaves, [d] = node level d-1
    int64_t count[MAX_LEVELS + 1];
    int top;                           // depth of the virtual root
};

__device__ __forceinline__ int64_t div_up_dev(int64_t a, int64_t b) { return (a + b - 1) / b; }

__device__ __forceinline__ bool boxes_intersect(const Box &a, const Box &b) {
    // strict: touching boxes do not intersect.  Bitwise &: all four compares are
    // evaluated (a short-circuit makes the second half of a shared-memory box load
    // wait for the compare on the first half)
    return (a.xmin < b.xmax) & (b.xmin < a.xmax) & (a.ymin < b.ymax) & (b.ymin < a.ymax);
}

// 16-byte asynchronous copy global -> shared (LDGSTS); completion is awaited with
// cp_async_wait_all() by the issuing thread
__device__ __forceinline__ void cp_async16(void *smem, const void *gmem) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;\n" ::"r"((unsigned)__cvta_generic_to_shared(smem)),
                 "l"(gmem)
                 : "memory");
}
__device__ __forceinline__ void cp_async8(void *smem, const void *gmem) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"((unsigned)__cvta_generic_to_shared(smem)),
                 "l"(gmem)
                 : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;\n" ::: "memory"); }
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
// waits until at most one committed group of this thread is still in flight
__device__ __forceinline__ void cp_async_wait_but_one() { asm volatile("cp.async.wait_group 1;\n" ::: "memory"); }
// loads whose position in the instruction stream matters (software pipelining):
// a volatile asm is not moved across the other volatile asms around it
__device__ __forceinline__ int2 ld_int2_here(const int2 *p) {
    int2 v;
    asm volatile("ld.global.nc.v2.s32 {%0, %1}, [%2];\n" : "=r"(v.x), "=r"(v.y) : "l"(p));
    return v;
}
__device__ __forceinline__ int64_t ld_i64_here(const int64_t *p) {
    long long v;
    asm volatile("ld.global.nc.s64 %0, [%1];\n" : "=l"(v) : "l"(p));
    return (int64_t)v;
}
__device__ __forceinline__ void schedule_fence() { asm volatile("" ::: "memory"); }
__device__ __forceinline__ int32_t atomic_add_here(int32_t *p, int32_t v) {
    int32_t old;
    asm volatile("atom.global.add.s32 %0, [%1], %2;\n" : "=r"(old) : "l"(p), "r"(v) : "memory");
    return old;
}

// pairs[i] = {sorted target position, sorted source position}, or {-1, -1} for an unused
// slot of a window.
//
// Persistent: a block takes groups of TRAV_WARPS consecutive tiles in grid-stride order, a
// warp one tile of each group.  Candidate pairs are written to the arena in WINDOWS of
// TRAV_WIN slots that a warp reserves ahead of time: the reservation (one atomicAdd on the
// arena counter) is issued right after the previous window was written and its result is
// first needed a window of work later.  The first version reserved exactly the pairs it
// had (one atomicAdd per flush and per tile, result needed at once): ncu showed 40 % of
// the kernel's stall samples on the shuffle that broadcasts that result -- several 100 k
// round trips to one hot address.  Cost of the windows: the unused tail of every warp's
// last window (< 1 % of the arena at 10 M targets), marked invalid.
__global__ void __launch_bounds__(TRAV_WARPS * 32)
k_traverse(const TraverseParams P, const Box *__restrict__ tbox, int64_t n_t, int tile_stride,
           int2 *__restrict__ pairs, unsigned long long cap, int32_t *__restrict__ cand_cnt,
           unsigned long long *__restrict__ ctr) {
    __shared__ Box s_node[TRAV_WARPS][32];
    __shared__ int32_t s_stack[TRAV_WARPS][TRAV_STACK];
    __shared__ int2 s_q[TRAV_WARPS][TRAV_WIN];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // tile_stride > 1: density sampling pass (every tile_stride-th tile, cap == 0: count only)
    const int64_t n_tiles = div_up_dev(div_up_dev(n_t, 32), tile_stride);
    const int64_t n_groups = div_up_dev(n_tiles, TRAV_WARPS);
    if ((int64_t)blockIdx.x * TRAV_WARPS + w >= n_tiles) return;  // this warp never gets a tile

    unsigned long long win = 0;  // start of the reserved window (lane 0)
    if (lane == 0) win = atomicAdd(ctr + CTR_PAIRS, (unsigned long long)TRAV_WIN);
    int qn = 0;
    unsigned visits = 0, tests = 0;
    unsigned long long real_pairs = 0;

    // writes the full window and reserves the next one
    auto flush_window = [&]() {
        __syncwarp();
        const unsigned long long base = __shfl_sync(0xffffffffu, win, 0);
        for (int i = lane; i < TRAV_WIN; i += 32)
            if (base + i < cap) pairs[base + i] = s_q[w][i];
        __syncwarp();
        if (lane == 0) win = atomicAdd(ctr + CTR_PAIRS, (unsigned long long)TRAV_WIN);
    };

    auto load_node = [&](int32_t entry) {
        const int d = entry >> 27;
        const int64_t c0 = (int64_t)(entry & 0x7ffffff) * FANOUT;
        const int64_t nc = P.count[d - 1] - c0;
        Box b;
        b.xmin = INFINITY; b.xmax = -INFINITY; b.ymin = INFINITY; b.ymax = -INFINITY;
        if (lane < nc) b = P.boxes[d - 1][c0 + lane];
        return b;
    };

    for (int64_t group = blockIdx.x; group < n_groups; group += gridDim.x) {
        const int64_t tile = (group * TRAV_WARPS + w) * tile_stride;
        const int64_t t = tile * 32 + lane;
        if (tile * 32 >= n_t) break;
        const bool valid = t < n_t;
        Box tb;
        if (valid) tb = tbox[t];
        else { tb.xmin = INFINITY; tb.xmax = -INFINITY; tb.ymin = INFINITY; tb.ymax = -INFINITY; }
        Box ub = tb;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            ub.xmin = fmin(ub.xmin, __shfl_xor_sync(0xffffffffu, ub.xmin, o));
            ub.xmax = fmax(ub.xmax, __shfl_xor_sync(0xffffffffu, ub.xmax, o));
            ub.ymin = fmin(ub.ymin, __shfl_xor_sync(0xffffffffu, ub.ymin, o));
            ub.ymax = fmax(ub.ymax, __shfl_xor_sync(0xffffffffu, ub.ymax, o));
        }
        int sp = 1;
        if (lane == 0) s_stack[w][0] = P.top << 27;
        int my_cand = 0;  // candidates of this lane's target = capacity of its provisional row
        __syncwarp();
        while (sp > 0) {
            const int32_t entry = s_stack[w][--sp];
            const int d = entry >> 27;
            const int64_t c0 = (int64_t)(entry & 0x7ffffff) * FANOUT;
            const Box mine = load_node(entry);
            ++visits;
            unsigned m = __ballot_sync(0xffffffffu, boxes_intersect(mine, ub));
            __syncwarp();
            if (m == 0) continue;
            s_node[w][lane] = mine;
            __syncwarp();
            if (d > 1) {
                while (m) {
                    const int c = __ffs(m) - 1;
                    m &= m - 1;
                    const Box b = s_node[w][c];
                    if (__any_sync(0xffffffffu, boxes_intersect(tb, b))) {
                        if (lane == 0) s_stack[w][sp] = ((d - 1) << 27) | (int32_t)(c0 + c);
                        ++sp;
                    }
                }
            } else {
                tests += __popc(m);
                // every lane collects the children that meet ITS target as a bit mask (two
                // children per trip: both boxes are loaded before either is tested); nothing
                // is voted or queued per child
                unsigned hits = 0u;
                while (m) {
                    const int ca = __ffs(m) - 1;
                    m &= m - 1;
                    const bool two = m != 0u;
                    const int cb2 = two ? __ffs(m) - 1 : ca;
                    m &= m - 1;
                    const Box ba = s_node[w][ca], bb = s_node[w][cb2];
                    const bool ha = boxes_intersect(tb, ba), hb = two & boxes_intersect(tb, bb);
                    hits |= (ha ? 1u : 0u) << ca;
                    hits |= (hb ? 1u : 0u) << cb2;
                }
                const int n_mine = __popc(hits);
                my_cand += n_mine;
                int incl = n_mine;  // inclusive prefix sum over the lanes
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int v = __shfl_up_sync(0xffffffffu, incl, o);
                    incl += lane >= o ? v : 0;
                }
                const int total = __shfl_sync(0xffffffffu, incl, 31);
                if (total > 0) {
                    // a lane appends its own pairs (the pairs of a node are grouped by target);
                    // whenever the window is full it goes out and the rest continues in the next
                    int pos = qn + incl - n_mine;
                    qn += total;
                    real_pairs += (unsigned long long)total;
                    while (true) {
                        while (hits != 0u && pos < TRAV_WIN) {
                            const int c = __ffs(hits) - 1;
                            hits &= hits - 1;
                            s_q[w][pos++] = make_int2((int32_t)t, (int32_t)(c0 + c));
                        }
                        if (qn < TRAV_WIN) break;
                        flush_window();
                        qn -= TRAV_WIN;
                        pos -= TRAV_WIN;
                    }
                }
            }
            __syncwarp();
        }
        if (valid && cand_cnt) cand_cnt[t] = my_cand;
    }
    // the last window: what is queued, the rest marked unused
    __syncwarp();
    {
        const unsigned long long base = __shfl_sync(0xffffffffu, win, 0);
        for (int i = lane; i < TRAV_WIN; i += 32)
            if (base + i < cap) pairs[base + i] = i < qn ? s_q[w][i] : make_int2(-1, -1);
    }
    if (lane == 0) {
        atomicAdd(ctr + CTR_VISITS, (unsigned long long)visits);
        atomicAdd(ctr + CTR_TESTS, (unsigned long long)tests * 32ull);
        atomicAdd(ctr + CTR_CAND, real_pairs);
    }
}

template <int K>
__device__ __forceinline__ int load_poly(const double2 *__restrict__ poly, const uint8_t *__restrict__ nv,
                                         int64_t i, int k, double2 (&v)[K]) {
    const double2 *src = poly + i * k;
#pragma unroll
    for (int j = 0; j < K; ++j)
        if (j < k) v[j] = src[j];
    if (K == 3) return 3;  // k == 3: every face is a triangle, no length lookup
    return nv[i] & 0x7f;
}

// ---- clipping: two dense kernels (the dominant cost of the build) -----------------
//   k_sat   one thread per candidate pair, straight-line separating-axis filter;
//           survivors are appended to a compact list (block-aggregated: one
//           global atomic per block, so a block's survivors stay contiguous)
//   k_clip_survivors  one thread per survivor: every lane of a warp clips a pair
//           that really intersects, so the fp64 pipe sees full warps
// Both read their element count from device memory and stride over it, so no
// host round trip separates them from the walk.

// rare slow path of clip_area_masks(), kept out of line so that it does not
// cost the main path registers
template <int KT, int KS>
__device__ __noinline__ double clip_area_fallback(const double2 *pa, int na, const double2 *pb, int nb,
                                                  double2 *bufA, double2 *bufB, int stride) {
    double2 a[KT], b[KS];
#pragma unroll
    for (int i = 0; i < KT; ++i) a[i] = i < na ? pa[i] : make_double2(0.0, 0.0);
#pragma unroll
    for (int i = 0; i < KS; ++i) b[i] = i < nb ? pb[i] : make_double2(0.0, 0.0);
    return clip_area<KT, KS>(a, na, b, nb, bufA, bufB, stride);
}

constexpr int SAT_WIN = 64;                 // survivors per window of the survivor list (reserved ahead, as in the walk)
constexpr int SAT_QCAP = SAT_WIN + 32;      // survivor queue entries per warp (one trip adds <= 32)
constexpr int SAT_RUN = 8;                  // consecutive 32-pair groups per warp (>= 4)

// One thread per candidate pair, warp-uniform grid-stride trips.  Like the clip
// kernel it is software-pipelined: the polygons of the thread's next pair are
// staged into shared memory with cp.async while the current pair is tested, the
// pair record after that is in flight to a register.  Survivors are compacted
// with a warp vote into a per-warp queue and written to the survivor list in windows
// of SAT_WIN slots reserved one window ahead (no block-wide barrier, no wait for the
// atomic; the tail of a warp's last window is marked unused with {-1, -1}).
#ifndef XG_SAT_FP32
#define XG_SAT_FP32 1
#endif
#ifndef XG_SAT_MINB
#define XG_SAT_MINB 3
#endif
template <int KT, int KS, int BT>
__global__ void __launch_bounds__(BT, (KT + KS <= 8) ? XG_SAT_MINB : 1)
k_sat(const int2 *__restrict__ pairs, unsigned long long cap, const double2 *__restrict__ tpoly,
      const uint8_t *__restrict__ tnv, int kt, const double2 *__restrict__ spoly,
      const uint8_t *__restrict__ snv, int ks, int2 *__restrict__ surv,
      unsigned long long *__restrict__ ctr) {
    extern __shared__ double2 s_buf[];
    __shared__ int2 s_q[BT / 32][SAT_QCAP];
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const unsigned lt_mask = (1u << lane) - 1u;
    double2 *stg = s_buf + tid;
    const unsigned long long n_pairs = ctr[CTR_PAIRS];
    // overflowed walk: the arena holds only part of the candidates while the provisional
    // rows were sized for all of them -- nothing downstream may run; the job is re-run
    // with the exact capacity the walk counted
    if (n_pairs > cap) return;

    constexpr int M = KT + KS;
    auto stage = [&](const int2 r, int sl) {
        const double2 *tp = tpoly + (int64_t)r.x * kt;
        const double2 *sp = spoly + (int64_t)r.y * ks;
        double2 *dst = stg + (size_t)sl * M * BT;
#pragma unroll
        for (int j = 0; j < KT; ++j)
            if (j < kt) cp_async16(dst + j * BT, tp + j);
#pragma unroll
        for (int j = 0; j < KS; ++j)
            if (j < ks) cp_async16(dst + (KT + j) * BT, sp + j);
    };
    unsigned long long win = 0;  // start of the reserved window (lane 0)
    // writes the first SAT_WIN queued survivors, moves the rest (< 32) to the front of the
    // queue and reserves the next window
    auto flush_window = [&](int &qn) {
        __syncwarp();
        const unsigned long long base = __shfl_sync(0xffffffffu, win, 0);
        for (int i = lane; i < SAT_WIN; i += 32) surv[base + i] = s_q[w][i];
        const int rest = qn - SAT_WIN;
        const int2 moved = lane < rest ? s_q[w][SAT_WIN + lane] : make_int2(0, 0);
        __syncwarp();
        if (lane < rest) s_q[w][lane] = moved;
        if (lane == 0) win = atomicAdd(ctr + CTR_SURV, (unsigned long long)SAT_WIN);
        __syncwarp();
        qn = rest;
    };

    // a warp takes SAT_RUN consecutive groups of 32 pairs before it jumps ahead, so that
    // the survivors it queues stay neighbours in the arena (same tile of targets): the
    // row scatter that follows depends on that locality (measured: 1.19 -> 1.44 ms
    // with plain grid-stride trips)
    const unsigned long long gw = (unsigned long long)blockIdx.x * (BT / 32) + w;
    const unsigned long long n_gw = (unsigned long long)gridDim.x * (BT / 32);
    const unsigned long long jump = (n_gw * SAT_RUN - (SAT_RUN - 1)) * 32ull;  // last group of a run -> next run
    // positions of this thread's current pair and of the three after it; tN = place of pN in its run
    unsigned long long p0 = gw * SAT_RUN * 32ull + lane, p1 = p0 + 32, p2 = p0 + 64, p3 = p0 + 96;
    int t3 = 3;  // (SAT_RUN >= 4)
    if (p0 - lane >= n_pairs) return;  // this warp has no pairs
    if (lane == 0) win = atomicAdd(ctr + CTR_SURV, (unsigned long long)SAT_WIN);
    // Software pipeline, two trips deep (ncu on the one-trip version: 15 % of all stall samples
    // on the wait for the staged polygons): while pair p0 is tested, the polygons of p1 are
    // landing in the other staging slot, those of p2 are requested into the slot p0 has just
    // been read from, and the record of p3 is in flight to a register.
    // (an unused slot of the arena, {-1, -1}, is staged as pair {0, 0} and never passes)
    int2 r0 = make_int2(-1, -1), r1 = make_int2(-1, -1), r2 = make_int2(-1, -1);
    int na0 = 3, nb0 = 3, na1 = 3, nb1 = 3;
    int slot = 0;
    auto stage_checked = [&](const int2 rec, int sl, int &na, int &nb) {
        const int2 safe = rec.x >= 0 ? rec : make_int2(0, 0);
        stage(safe, sl);
        if (KT > 3) na = tnv[safe.x] & 0x7f;
        if (KS > 3) nb = snv[safe.y] & 0x7f;
    };
    if (p0 < n_pairs) {
        r0 = pairs[p0];
        stage_checked(r0, 0, na0, nb0);
    }
    cp_async_commit();
    if (p1 < n_pairs) {
        r1 = pairs[p1];
        stage_checked(r1, 1, na1, nb1);
    }
    cp_async_commit();
    if (p2 < n_pairs) r2 = pairs[p2];
    int qn = 0;
    unsigned n_pass = 0;
    // (warp-uniform trip count: the first pair of the warp's group decides)
#pragma unroll 1
    while (p0 - lane < n_pairs) {
        const bool valid = p0 < n_pairs && r0.x >= 0;
        cp_async_wait_but_one();  // the older of the two groups in flight: the polygons of p0
        const double2 *cur = stg + (size_t)slot * M * BT;
        double2 a[KT], b[KS];
#pragma unroll
        for (int jj = 0; jj < KT; ++jj) a[jj] = cur[jj * BT];
#pragma unroll
        for (int jj = 0; jj < KS; ++jj) b[jj] = cur[(KT + jj) * BT];
        const int2 r_cur = r0;
        const int na_cur = na0, nb_cur = nb0;
        r0 = r1; na0 = na1; nb0 = nb1;
        r1 = r2;
        if (p2 < n_pairs) stage_checked(r1, slot, na1, nb1);
        cp_async_commit();  // (an empty group when nothing was staged keeps the count right)
        if (p3 < n_pairs) r2 = ld_int2_here(pairs + p3);
        schedule_fence();
        bool pass = false;
#if XG_SAT_FP32
        if (KT <= 4 && KS <= 4) {
            // fp32 pre-decision (xg_geom.cuh): certain for all but the pairs within rounding
            // distance of a tie; a warp that holds such a pair runs the fp64 test as well
            const int pre = valid ? polygons_intersect_fp32<KT, KS>(a, na_cur, b, nb_cur) : SAT_REJECT;
            pass = pre == SAT_ACCEPT;
            if (__any_sync(0xffffffffu, pre == SAT_AMBIGUOUS)) {
                const bool exact = polygons_intersect_uniform<KT, KS>(a, na_cur, b, nb_cur);
                if (pre == SAT_AMBIGUOUS) pass = exact;
            }
        } else
#endif
        if (valid) {
            if (KT <= 8 && KS <= 8) pass = polygons_intersect_uniform<KT, KS>(a, na_cur, b, nb_cur);
            else pass = polygons_intersect<KT, KS>(a, na_cur, b, nb_cur);
        }
        const unsigned m = __ballot_sync(0xffffffffu, pass);
        if (pass) s_q[w][qn + __popc(m & lt_mask)] = r_cur;
        qn += __popc(m);
        n_pass += __popc(m);
        if (qn >= SAT_WIN) flush_window(qn);
        p0 = p1;
        p1 = p2;
        p2 = p3;
        ++t3;
        p3 += (t3 == SAT_RUN) ? jump : 32ull;
        t3 = (t3 == SAT_RUN) ? 0 : t3;
        slot ^= 1;
    }
    cp_async_wait_all();
    // the last window: what is queued, the rest marked unused
    __syncwarp();
    {
        const unsigned long long base = __shfl_sync(0xffffffffu, win, 0);
        for (int i = lane; i < SAT_WIN; i += 32) surv[base + i] = i < qn ? s_q[w][i] : make_int2(-1, -1);
    }
    if (lane == 0) atomicAdd(ctr + CTR_NSURV, (unsigned long long)n_pass);
}

// One thread per survivor, grid-stride.  The kernel is bound by the latency of
// the polygon gathers (ncu: long-scoreboard stalls were half of all samples with
// 24 warps per SM), so the polygons of the NEXT survivor of a thread are staged
// into its shared-memory slots with cp.async while the current one is clipped,
// and the survivor record after that is already in flight to a register.
// Shared memory per thread: two (KT+KS)-vertex ping-pong lists + one staged pair.
//
// A positive area goes straight into the target's PROVISIONAL row (rows in sorted
// target order, capacity = candidates counted by the walk): slot = atomicAdd on the
// row's fill counter.  The atomic's return value is only needed one trip later
// (pend_*), and it is issued after the cp.async wait of that next trip so that it
// cannot delay the wait (both end up on the same hardware scoreboard).
template <int KT, int KS, int BT, int MINB>
__global__ void __launch_bounds__(BT, MINB)
k_clip_survivors(const int2 *__restrict__ surv, const unsigned long long *__restrict__ ctr,
                 const double2 *__restrict__ tpoly, const uint8_t *__restrict__ tnv, int kt,
                 const double2 *__restrict__ spoly, const uint8_t *__restrict__ snv, int ks,
                 const int64_t *__restrict__ prov_ptr, int32_t *__restrict__ fill_cnt,
                 int32_t *__restrict__ prov_col, double *__restrict__ prov_val) {
    extern __shared__ double2 s_buf[];
    constexpr int M = KT + KS;
    const unsigned long long n_surv = ctr[CTR_SURV];
    double2 *bufA = s_buf + threadIdx.x;
    double2 *bufB = s_buf + (size_t)M * BT + threadIdx.x;
    double2 *stg = s_buf + (size_t)2 * M * BT + threadIdx.x;
    const unsigned long long stride = (unsigned long long)gridDim.x * BT;
    unsigned long long q = (unsigned long long)blockIdx.x * BT + threadIdx.x;

    auto stage = [&](const int2 r) {
        const double2 *tp = tpoly + (int64_t)r.x * kt;
        const double2 *sp = spoly + (int64_t)r.y * ks;
#pragma unroll
        for (int j = 0; j < KT; ++j)
            if (j < kt) cp_async16(stg + j * BT, tp + j);
#pragma unroll
        for (int j = 0; j < KS; ++j)
            if (j < ks) cp_async16(stg + (KT + j) * BT, sp + j);
    };

    // loop-carried pipeline registers: r / na / nb / base describe the STAGED
    // survivor, r_next the one after it; pend_* the clipped one awaiting its slot
    int2 r = make_int2(0, 0), r_next = make_int2(0, 0);
    int64_t base = 0;    // prov_ptr of the staged survivor's target
    int na = 3, nb = 3;  // vertex counts (looked up for K > 3 only)
    // (an unused slot of the survivor list, {-1, -1}, is staged as pair {0, 0}; its area is dropped)
    auto stage_checked = [&](const int2 rec) {
        const int2 safe = rec.x >= 0 ? rec : make_int2(0, 0);
        stage(safe);
        base = ld_i64_here(prov_ptr + safe.x);
        if (KT > 3) na = tnv[safe.x] & 0x7f;
        if (KS > 3) nb = snv[safe.y] & 0x7f;
    };
    if (q < n_surv) {
        r = surv[q];
        stage_checked(r);
    }
    if (q + stride < n_surv) r_next = surv[q + stride];
    int2 pend_r = make_int2(-1, 0);  // x < 0: nothing pending
    int64_t pend_base = 0;
    double pend_ar = 0.0;
    for (; q < n_surv; q += stride) {
        cp_async_wait_all();
        uint32_t slot = 0;
        double2 a[KT], b[KS];
#pragma unroll
        for (int j = 0; j < KT; ++j) a[j] = stg[j * BT];
#pragma unroll
        for (int j = 0; j < KS; ++j) b[j] = stg[(KT + j) * BT];
        const int2 r_cur = r;
        const int64_t base_cur = base;
        const int na_cur = na, nb_cur = nb;
        // ---- prefetch
        if (q + stride < n_surv) {
            r = r_next;
            stage_checked(r);
        }
        if (q + 2 * stride < n_surv) r_next = ld_int2_here(surv + q + 2 * stride);
        schedule_fence();
        // the slot of the pending entry: requested here, consumed after the clip
        if (pend_r.x >= 0) slot = (uint32_t)atomic_add_here(fill_cnt + pend_r.x, 1);
        // ---- clip
        double ar;
        if (KT + KS <= 32) {
            bool fb = false;
            ar = clip_area_masks<KT, KS>(a, na_cur, b, nb_cur, bufA, bufB, BT, &fb);
            if (fb && r_cur.x >= 0)
                ar = clip_area_fallback<KT, KS>(tpoly + (int64_t)r_cur.x * kt, na_cur,
                                                spoly + (int64_t)r_cur.y * ks, nb_cur, bufA, bufB, BT);
        } else {
            ar = clip_area<KT, KS>(a, na_cur, b, nb_cur, bufA, bufB, BT);
        }
        schedule_fence();  // keeps the stores of the pending entry late
        // ptxas widened the slot to 64 bits for the address right behind the atomic -- a wait for
        // its whole round trip before the clip had even started (ncu: 20 % of the kernel's
        // stall samples on that one shift).  Opaque until here: nothing derived from it can
        // be computed earlier.
        asm volatile("" : "+r"(slot));
        if (pend_r.x >= 0) {
            prov_col[pend_base + slot] = pend_r.y;
            prov_val[pend_base + slot] = pend_ar;
        }
        pend_r = (ar > 0.0 && r_cur.x >= 0) ? r_cur : make_int2(-1, 0);
        pend_base = base_cur;
        pend_ar = ar;
    }
    if (pend_r.x >= 0) {
        const uint32_t slot = (uint32_t)atomicAdd(fill_cnt + pend_r.x, 1);
        prov_col[pend_base + slot] = pend_r.y;
        prov_val[pend_base + slot] = pend_ar;
    }
}

// area of a packed face with the reference's arithmetic (connectivity.py:615-633:
// shoelace relative to the first vertex of the ORIGINAL vertex order)
__device__ __forceinline__ double packed_face_area(const double2 *__restrict__ poly, uint8_t nvf) {
    const int n = nvf & 0x7f;
    const bool flip = nvf & 0x80;
    const double2 v0 = poly[flip ? n - 1 : 0];
    double det = 0.0;
    for (int j = 0; j < n; ++j) {
        const int ja = flip ? n - 1 - j : j;
        const int jn = (j + 1 < n) ? j + 1 : 0;
        const int jb = flip ? n - 1 - jn : jn;
        const double ax = poly[ja].x - v0.x, ay = poly[ja].y - v0.y;
        const double bx = poly[jb].x - v0.x, by = poly[jb].y - v0.y;
        det += ax * by - ay * bx;
    }
    return 0.5 * fabs(det);
}

// rows are counted by sorted target position; the row pointer is in the caller's
// row order: cnt_row[tperm[t] - t_begin] = fill_cnt[t]
__global__ void __launch_bounds__(256)
k_counts_to_rows(const int32_t *__restrict__ fill_cnt, const int32_t *__restrict__ tperm, int64_t t_begin,
                 int64_t n_t, int32_t *__restrict__ cnt_row) {
    const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t < n_t) cnt_row[tperm[t] - t_begin] = fill_cnt[t];
    if (t == 0) cnt_row[n_t] = 0;
}

__device__ __forceinline__ void cswap(int32_t &ka, double &va, int32_t &kb, double &vb) {
    const bool sw = ka > kb;
    const int32_t tk = sw ? kb : ka;
    const double tv = sw ? vb : va;
    kb = sw ? ka : kb;
    vb = sw ? va : vb;
    ka = tk;
    va = tv;
}

// Batcher odd-even merge sort networks over register-resident (key, value)
// slots (19 / 63 comparators; generated and checked with the 0-1 principle)
#define XG_CS(i, j) cswap(k[i], v[i], k[j], v[j])
__device__ __forceinline__ void sort_network(int32_t (&k)[8], double (&v)[8]) {
    XG_CS(0, 1); XG_CS(2, 3); XG_CS(4, 5); XG_CS(6, 7); XG_CS(0, 2); XG_CS(1, 3); XG_CS(4, 6); XG_CS(5, 7);
    XG_CS(1, 2); XG_CS(5, 6); XG_CS(0, 4); XG_CS(1, 5); XG_CS(2, 6); XG_CS(3, 7); XG_CS(2, 4); XG_CS(3, 5);
    XG_CS(1, 2); XG_CS(3, 4); XG_CS(5, 6);
}
__device__ __forceinline__ void sort_network(int32_t (&k)[16], double (&v)[16]) {
    XG_CS(0, 1); XG_CS(2, 3); XG_CS(4, 5); XG_CS(6, 7); XG_CS(8, 9); XG_CS(10, 11); XG_CS(12, 13); XG_CS(14, 15);
    XG_CS(0, 2); XG_CS(1, 3); XG_CS(4, 6); XG_CS(5, 7); XG_CS(8, 10); XG_CS(9, 11); XG_CS(12, 14); XG_CS(13, 15);
    XG_CS(1, 2); XG_CS(5, 6); XG_CS(9, 10); XG_CS(13, 14);
    XG_CS(0, 4); XG_CS(1, 5); XG_CS(2, 6); XG_CS(3, 7); XG_CS(8, 12); XG_CS(9, 13); XG_CS(10, 14); XG_CS(11, 15);
    XG_CS(2, 4); XG_CS(3, 5); XG_CS(10, 12); XG_CS(11, 13);
    XG_CS(1, 2); XG_CS(3, 4); XG_CS(5, 6); XG_CS(9, 10); XG_CS(11, 12); XG_CS(13, 14);
    XG_CS(0, 8); XG_CS(1, 9); XG_CS(2, 10); XG_CS(3, 11); XG_CS(4, 12); XG_CS(5, 13); XG_CS(6, 14); XG_CS(7, 15);
    XG_CS(4, 8); XG_CS(5, 9); XG_CS(6, 10); XG_CS(7, 11);
    XG_CS(2, 4); XG_CS(3, 5); XG_CS(6, 8); XG_CS(7, 9); XG_CS(10, 12); XG_CS(11, 13);
    XG_CS(1, 2); XG_CS(3, 4); XG_CS(5, 6); XG_CS(7, 8); XG_CS(9, 10); XG_CS(11, 12); XG_CS(13, 14);
}
#undef XG_CS

// what a row needs to go from the provisional CSR to the caller's CSR
struct RowMove {
    const int64_t *prov_ptr;   // (n_t + 1) provisional row starts, sorted target order
    const int32_t *fill_cnt;   // (n_t) entries per provisional row
    const int32_t *prov_col;   // SORTED source position per entry
    const double *prov_val;    // overlap area per entry
    __device__ __forceinline__ void entry(int64_t i, int32_t &col, double &val) const {
        col = prov_col[i];
        val = prov_val[i];
    }
    const int32_t *tperm;      // sorted target position -> caller's target face
    int64_t t_begin;
    const int32_t *sperm;      // sorted source position -> caller's source face
    const double2 *spoly;      // packed source polygons (relative weights)
    const uint8_t *snv;
    int ks;
    int relative;
    const int64_t *indptr;     // caller's row pointer
    int32_t *indices;
    double *data;
    int64_t entry_cap;         // entries the caller's arrays can hold (rows beyond it are dropped and flagged)
    unsigned long long *flags;
    __device__ __forceinline__ bool fits(int64_t dst, int64_t n) const { return dst + n <= entry_cap; }
};

__device__ __forceinline__ double entry_value(const RowMove &R, int32_t s, double a) {
    // relative weights: unstructured.py:133-134, area of the source face as connectivity.area computes it
    return R.relative ? a / packed_face_area(R.spoly + (int64_t)s * R.ks, R.snv[s]) : a;
}

template <int N>
__device__ __forceinline__ void move_row_regs(const RowMove &R, int64_t src, int64_t dst, int n) {
    int32_t k[N];
    double v[N];
#pragma unroll
    for (int i = 0; i < N; ++i) {
        k[i] = 0x7fffffff;
        v[i] = 0.0;
        if (i < n) {
            int32_t s;
            double a;
            R.entry(src + i, s, a);
            k[i] = R.sperm[s];
            v[i] = entry_value(R, s, a);
        }
    }
    sort_network(k, v);
#pragma unroll
    for (int i = 0; i < N; ++i)
        if (i < n) {
            R.indices[dst + i] = k[i];
            R.data[dst + i] = v[i];
        }
}

constexpr int SORT_SHORT = 16;       // rows up to this length are sorted by one thread in registers
constexpr int SORT_LONG_CAP = 1024;  // rows up to this length are rank-sorted by one warp in shared memory

// One thread per sorted target: its provisional row -> its row of the caller's
// CSR, entries in ascending source index (MatrixCSR rows as sparse.py:61-78 would
// hold them after a stable sort by column).  Consecutive threads read consecutive
// provisional rows; short rows (the common case at comparable resolutions) go
// through a fixed network in registers, warp-uniformly (8 or 16 slots by the
// longest row of the warp); longer rows are queued for k_move_long_rows.
__global__ void __launch_bounds__(256)
k_move_rows(const RowMove R, int64_t n_t, int32_t *__restrict__ long_rows,
            unsigned long long *__restrict__ n_long) {
    const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    int64_t src = 0, dst = 0;
    int n = 0;
    if (*R.flags != 0ull) return;  // (the rows do not fit the caller's arrays: nothing is moved)
    if (t < n_t) {
        n = R.fill_cnt[t];
        src = R.prov_ptr[t];
        dst = R.indptr[R.tperm[t] - R.t_begin];
    }
    if (n > 0 && !R.fits(dst, n)) n = 0;
    const bool is_long = n > SORT_SHORT;
    const unsigned lm = __ballot_sync(0xffffffffu, is_long);
    if (lm) {
        const int lane = threadIdx.x & 31;
        unsigned long long base = 0;
        if (lane == __ffs(lm) - 1) base = atomicAdd(n_long, (unsigned long long)__popc(lm));
        base = __shfl_sync(0xffffffffu, base, __ffs(lm) - 1);
        if (is_long) long_rows[base + __popc(lm & ((1u << lane) - 1u))] = (int32_t)t;
    }
    const int ns = is_long ? 0 : n;
    const bool wide = __any_sync(0xffffffffu, ns > 8);
    if (wide) {
        if (ns >= 1) move_row_regs<16>(R, src, dst, ns);
    } else {
        if (ns >= 1) move_row_regs<8>(R, src, dst, ns);
    }
}

// one warp per queued row: rank sort through shared memory (source indices of a
// row are distinct, so the rank of a key is its final position)
__global__ void __launch_bounds__(128)
k_move_long_rows(const RowMove R, const int32_t *__restrict__ long_rows,
                 const unsigned long long *__restrict__ n_long, int32_t *__restrict__ huge_rows,
                 unsigned long long *__restrict__ n_huge) {
    __shared__ int32_t s_key[4][SORT_LONG_CAP];
    __shared__ double s_val[4][SORT_LONG_CAP];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const unsigned long long total = *n_long;
    for (unsigned long long q = (unsigned long long)blockIdx.x * 4 + w; q < total;
         q += (unsigned long long)gridDim.x * 4) {
        const int64_t t = long_rows[q];
        const int64_t n = R.fill_cnt[t];
        const int64_t src = R.prov_ptr[t];
        const int64_t dst = R.indptr[R.tperm[t] - R.t_begin];
        // (rows that do not fit the caller's arrays never get here: k_move_rows dropped them)
        if (n <= SORT_LONG_CAP) {
            for (int i = lane; i < n; i += 32) {
                int32_t s;
                double a;
                R.entry(src + i, s, a);
                s_key[w][i] = R.sperm[s];
                s_val[w][i] = entry_value(R, s, a);
            }
            __syncwarp();
            for (int i = lane; i < n; i += 32) {
                const int32_t key = s_key[w][i];
                int rank = 0;
                for (int j = 0; j < (int)n; ++j) rank += s_key[w][j] < key ? 1 : 0;
                R.indices[dst + rank] = key;
                R.data[dst + rank] = s_val[w][i];
            }
            __syncwarp();
        } else if (lane == 0) {
            // e.g. a sliver on the hull of a Delaunay target that covers thousands of
            // source faces: a whole block sorts it (k_move_huge_rows)
            huge_rows[atomicAdd(n_huge, 1ull)] = (int32_t)t;
        }
    }
}

// one block per queued row of more than SORT_LONG_CAP entries: copy to the caller's row,
// then a bitonic sort in place (global memory, L1/L2 resident).  The network is the
// variant whose merges all run ascending (first step of a merge compares i with its
// mirror in the block), so the virtual +inf padding up to the next power of two never
// moves and needs no storage.
__global__ void __launch_bounds__(1024)
k_move_huge_rows(const RowMove R, const int32_t *__restrict__ huge_rows,
                 const unsigned long long *__restrict__ n_huge) {
    const unsigned long long total = *n_huge;
    for (unsigned long long q = blockIdx.x; q < total; q += gridDim.x) {
        const int64_t t = huge_rows[q];
        const int64_t n = R.fill_cnt[t];
        const int64_t src = R.prov_ptr[t];
        int32_t *key = R.indices + R.indptr[R.tperm[t] - R.t_begin];
        double *val = R.data + R.indptr[R.tperm[t] - R.t_begin];
        for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
            int32_t s;
            double a;
            R.entry(src + i, s, a);
            key[i] = R.sperm[s];
            val[i] = entry_value(R, s, a);
        }
        __syncthreads();
        int64_t P = 1;
        while (P < n) P <<= 1;
        for (int64_t k = 2; k <= P; k <<= 1) {
            for (int64_t j = k >> 1; j > 0; j >>= 1) {
                const bool flip = (j == (k >> 1));
                for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
                    const int64_t p = flip ? (i ^ (k - 1)) : (i ^ j);
                    if (p > i && p < n) {
                        const int32_t ki = key[i], kp = key[p];
                        if (ki > kp) {
                            const double vi = val[i], vp = val[p];
                            key[i] = kp; key[p] = ki;
                            val[i] = vp; val[p] = vi;
                        }
                    }
                }
                __syncthreads();
            }
        }
    }
}

struct CastI64 {
    __host__ __device__ __forceinline__ int64_t operator()(const int32_t &v) const { return (int64_t)v; }
};

struct ClipArgs {
    const int2 *pairs;
    unsigned long long cap;
    const double *tpoly;
    const uint8_t *tnv;
    int kt;
    const double *spoly;
    const uint8_t *snv;
    int ks;
    int2 *surv;
    unsigned long long *ctr;
    int64_t *prov_ptr;   // (n_t + 1) provisional row starts: exclusive scan of the candidate counts
    int32_t *fill_cnt;   // (n_t + 1) positive areas per sorted target
    int32_t *prov_col;
    double *prov_val;
    cudaEvent_t between;  // recorded before the clip kernel (phase timing)
};

template <int KT, int KS, int BT, int MINB>
static int launch_clip(const ClipArgs &A, cudaStream_t s) {
    int per_sm = 1;
    // (the filter keeps its own block size: it is tuned separately from the clip kernel)
    constexpr int SBT = (KT + KS <= 8) ? 256 : BT;
    const size_t smem_sat = sizeof(double2) * 2 * (size_t)(KT + KS) * SBT;  // two staged pairs per thread
    // (attributes are set once per device and instantiation: nothing but launches while a job is
    //  being captured into a graph)
    static bool attr_set[64] = {false};
    int dev = 0;
    XG_CUDA(cudaGetDevice(&dev));
    const bool set_attr = dev < 0 || dev >= 64 || !attr_set[dev];
    if (set_attr)
        XG_CUDA(cudaFuncSetAttribute(k_sat<KT, KS, SBT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_sat));
    XG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_sat<KT, KS, SBT>, SBT, smem_sat));
    if (per_sm < 1) per_sm = 1;
    k_sat<KT, KS, SBT><<<num_sms() * per_sm, SBT, smem_sat, s>>>(A.pairs, A.cap, (const double2 *)A.tpoly, A.tnv, A.kt,
                                                       (const double2 *)A.spoly, A.snv, A.ks, A.surv, A.ctr);
    XG_CHECK_LAUNCH();
    if (A.between) XG_CUDA(cudaEventRecord(A.between, s));
    size_t smem = sizeof(double2) * 3 * (size_t)(KT + KS) * BT;  // 2 ping-pong lists + 1 staged pair
    auto kern = k_clip_survivors<KT, KS, BT, MINB>;
    if (set_attr) {
        XG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        if (dev >= 0 && dev < 64) attr_set[dev] = true;
    }
    XG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, BT, smem));
    if (per_sm < 1) per_sm = 1;
    kern<<<num_sms() * per_sm, BT, smem, s>>>(A.surv, A.ctr, (const double2 *)A.tpoly, A.tnv, A.kt,
                                            (const double2 *)A.spoly, A.snv, A.ks, A.prov_ptr, A.fill_cnt, A.prov_col,
                                            A.prov_val);
    XG_CHECK_LAUNCH();
    return XG_OK;
}

}  // namespace xg

using namespace xg;

struct xg_overlap {
    const xg_index *src = nullptr;
    int64_t t_begin = 0, n_t = 0;
    int k_t = 0;
    int64_t n_cand = 0, n_surv = 0, nnz = 0;
    cudaStream_t s = nullptr;
    // device buffers
    DevBuf tperm, tbox, tpoly, tnv, pairs, surv, prov_col, cnt, prov_ptr, scan_tmp, ctr;
    DevBuf pk_keys_in, pk_keys_out, pk_ids, pk_tmp;  // scratch of the target packing
    PackScratch pk;
    size_t scan_tmp_bytes = 0;
    int64_t h_stats[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    bool stats_valid = false;
    // phase timing: ev[i] is recorded when phase i starts, the last one at the end
    cudaEvent_t ev[XG_N_PHASES + 1];
    int n_ev = 0;
    double h_ms[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    // the attempt in flight: arena capacity, and (after overlap_collect) what the device reported
    unsigned long long cap = 0, needed = 0, flags = 0;
    int64_t nnz_end = 0;
    double *h_ctr = nullptr;     // pinned slot the counters are copied to
    // xg_overlap_run: nothing is waited for until xg_overlap_result
    bool deferred = false, empty = false;
    cudaEvent_t done = nullptr;
    // xg_overlap_run with XG_RUN_GRAPH: the whole job is one CUDA graph launch
    bool timing = true;
    cudaStream_t ls = nullptr;  // where the job's operations are issued: s, or the capture stream
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t graph_exec = nullptr;
    xg_overlap() { for (int i = 0; i <= XG_N_PHASES; ++i) ev[i] = nullptr; }
    ~xg_overlap() {
        for (int i = 0; i <= XG_N_PHASES; ++i) if (ev[i]) cudaEventDestroy(ev[i]);
        if (done) {
            // (the pinned slot may still be the target of the counter copy)
            if (h_ctr && !stats_valid) cudaEventSynchronize(done);
            cudaEventDestroy(done);
        }
        if (graph_exec) cudaGraphExecDestroy(graph_exec);  // (after the wait above: never while in flight)
        if (graph) cudaGraphDestroy(graph);
        if (h_ctr) pinned_slot_release(h_ctr);
    }
    void release_scratch() {
        tperm.release(); tbox.release(); tpoly.release(); tnv.release(); pairs.release(); surv.release();
        prov_col.release(); cnt.release(); prov_ptr.release(); scan_tmp.release(); ctr.release();
        pk_keys_in.release(); pk_keys_out.release(); pk_ids.release(); pk_tmp.release();
    }
    cudaError_t mark() {
        if (!timing) { ++n_ev; return cudaSuccess; }  // (a captured job records no phase events)
        if (!ev[n_ev]) {
            cudaError_t e = cudaEventCreate(&ev[n_ev]);
            if (e != cudaSuccess) return e;
        }
        return cudaEventRecord(ev[n_ev++], ls);
    }
    void collect() {
        if (!timing) return;
        for (int i = 0; i + 1 < n_ev; ++i) {
            float ms = 0.f;
            if (cudaEventElapsedTime(&ms, ev[i], ev[i + 1]) == cudaSuccess) h_ms[i] = ms;
        }
    }
};

static TraverseParams traverse_params(const xg_index *source) {
    TraverseParams P;
    P.boxes[0] = source->leaf_box; P.count[0] = source->n_face;
    for (int l = 0; l < source->n_levels; ++l) { P.boxes[l + 1] = source->level_box[l]; P.count[l + 1] = source->level_n[l]; }
    P.top = source->n_levels + 1;
    return P;
}

// blocks of the persistent walk over `tiles` tiles of 32 targets
static unsigned traverse_grid(int64_t tiles) {
    const int64_t groups = div_up(tiles, TRAV_WARPS);
    const int64_t resident = (int64_t)num_sms() * TRAV_BLOCKS_PER_SM;
    return (unsigned)(groups < resident ? (groups > 0 ? groups : 1) : resident);
}

// Candidate density of a job whose source mesh is much finer than its targets: walk
// every SAMPLE_STRIDE-th tile without storing anything and scale the count, so that
// the pair arena is sized right the first time (otherwise the full walk runs twice).
constexpr int SAMPLE_STRIDE = 64;
static int estimate_candidates(xg_overlap *job, unsigned long long *estimate) {
    cudaStream_t s = job->s;
    const int64_t tiles = div_up(job->n_t, 32);
    const int64_t sampled = div_up(tiles, SAMPLE_STRIDE);
    XG_CUDA(cudaMemsetAsync(job->ctr.p, 0, sizeof(unsigned long long) * CTR_N, s));
    k_traverse<<<traverse_grid(sampled), TRAV_WARPS * 32, 0, s>>>(
        traverse_params(job->src), job->tbox.as<Box>(), job->n_t, SAMPLE_STRIDE, nullptr, 0ull, nullptr,
        job->ctr.as<unsigned long long>());
    XG_CHECK_LAUNCH();
    unsigned long long h = 0;
    XG_CUDA(cudaMemcpyAsync(&h, job->ctr.as<unsigned long long>() + CTR_CAND, sizeof(h), cudaMemcpyDeviceToHost, s));
    XG_CUDA(cudaStreamSynchronize(s));
    // (+ the unused tails of the last windows of the persistent walk's warps)
    *estimate = (unsigned long long)((double)h * (double)tiles / (double)sampled * 1.25) + 65536ull +
                (unsigned long long)traverse_grid(tiles) * TRAV_WARPS * TRAV_WIN;
    return XG_OK;
}

// dst[0 .. n) = *src (rows of an empty job all point at the job's first entry)
__global__ void __launch_bounds__(256)
k_fill_from(int64_t *__restrict__ dst, int64_t n, const int64_t *__restrict__ src) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) dst[i] = *src;
}

// Enqueues one attempt with a pair arena of `cap` entries: walk, filter, clip, row pointer.
// Nothing is read back.  chained: the job's first entry sits at the offset found in
// d_indptr[0] on the device (the last row pointer of the job before it) instead of 0.
static int overlap_enqueue(xg_overlap *job, unsigned long long cap, int64_t *d_indptr, bool chained) {
    const xg_index *source = job->src;
    cudaStream_t s = job->ls;        // operations
    cudaStream_t as = job->s;        // the stream the scratch blocks belong to (where the job executes)
    const int64_t n_t = job->n_t;
    const int k_t = job->k_t, ks = source->k;
    job->n_ev = 1;  // keep ev[0] (pack start), re-record the rest
    job->cap = cap;

    XG_CUDA(cudaMemsetAsync(job->ctr.p, 0, sizeof(unsigned long long) * CTR_N, s));
    if (chained)
        XG_CUDA(cudaMemcpyAsync(job->ctr.as<unsigned long long>() + CTR_BASE, d_indptr, sizeof(int64_t),
                                cudaMemcpyDeviceToDevice, s));
    // (arenas: kept when overlap_alloc sized them already; + the unused tails of the last survivor
    //  windows: at most 8 blocks of 8 warps per SM)
    XG_CUDA(job->pairs.ensure(sizeof(int2) * (size_t)cap, as));
    XG_CUDA(job->surv.ensure(sizeof(int2) * ((size_t)cap + (size_t)num_sms() * 64 * SAT_WIN), as));
    XG_CUDA(job->prov_col.ensure(sizeof(int32_t) * (size_t)cap, as));
    // cnt = [cand_cnt | fill_cnt | row_cnt], (n_t + 1) int32 each
    XG_CUDA(cudaMemsetAsync(job->cnt.p, 0, sizeof(int32_t) * 2 * (size_t)(n_t + 1), s));
    int32_t *cand_cnt = job->cnt.as<int32_t>(), *fill_cnt = cand_cnt + (n_t + 1), *row_cnt = fill_cnt + (n_t + 1);

    const TraverseParams P = traverse_params(source);
    const unsigned tiles = (unsigned)div_up(n_t, 32);
    XG_CUDA(job->mark());  // XG_PHASE_TRAVERSE
    k_traverse<<<traverse_grid(tiles), TRAV_WARPS * 32, 0, s>>>(
        P, job->tbox.as<Box>(), n_t, 1, job->pairs.as<int2>(), cap, cand_cnt,
        job->ctr.as<unsigned long long>());
    XG_CHECK_LAUNCH();
    {
        // provisional row starts (sorted target order): a row can hold every candidate of its target
        cub::TransformInputIterator<int64_t, CastI64, const int32_t *> it(cand_cnt, CastI64());
        size_t tb = job->scan_tmp_bytes;
        XG_CUDA(cub::DeviceScan::ExclusiveSum(job->scan_tmp.p, tb, it, job->prov_ptr.as<int64_t>(), n_t + 1, s));
    }

    XG_CUDA(job->mark());  // XG_PHASE_SAT
    ClipArgs A;
    A.pairs = job->pairs.as<int2>(); A.cap = cap;
    A.tpoly = job->tpoly.as<double>(); A.tnv = job->tnv.as<uint8_t>(); A.kt = k_t;
    A.spoly = source->poly; A.snv = source->nvert; A.ks = ks;
    A.surv = job->surv.as<int2>(); A.ctr = job->ctr.as<unsigned long long>();
    A.fill_cnt = fill_cnt;
    A.prov_ptr = job->prov_ptr.as<int64_t>();
    A.prov_col = job->prov_col.as<int32_t>();
    A.prov_val = job->pairs.as<double>();  // the candidate arena is dead once k_sat has run (same 8 B per entry)
    if (job->timing) {
        if (!job->ev[job->n_ev]) XG_CUDA(cudaEventCreate(&job->ev[job->n_ev]));
        A.between = job->ev[job->n_ev++];  // XG_PHASE_CLIP starts between the two kernels
    } else {
        A.between = nullptr;
        ++job->n_ev;
    }
    int rc;
    // (block size / min-blocks variants 256x2..4, 128x6..8 were measured: 3.30-3.87 ms, 256x3 best)
#ifndef XG_CLIP_BT
#define XG_CLIP_BT 256
#define XG_CLIP_MINB 2
#endif
    if (k_t == 3 && ks == 3) rc = launch_clip<3, 3, XG_CLIP_BT, XG_CLIP_MINB>(A, s);
    else if (k_t == 3 && ks == 4) rc = launch_clip<3, 4, 256, 2>(A, s);
    else if (k_t == 4 && ks == 3) rc = launch_clip<4, 3, 256, 2>(A, s);
    else if (k_t == 4 && ks == 4) rc = launch_clip<4, 4, 256, 2>(A, s);
    else if (k_t <= 8 && ks <= 8) rc = launch_clip<8, 8, 128, 2>(A, s);
    else if (k_t <= 16 && ks <= 16) rc = launch_clip<16, 16, 64, 1>(A, s);
    else rc = launch_clip<XG_MAX_N_VERTEX, XG_MAX_N_VERTEX, 32, 1>(A, s);
    if (rc) return rc;

    XG_CUDA(job->mark());  // XG_PHASE_ROWPTR
    k_counts_to_rows<<<(unsigned)div_up(n_t, 256), 256, 0, s>>>(fill_cnt, job->tperm.as<int32_t>(), job->t_begin,
                                                               n_t, row_cnt);
    XG_CHECK_LAUNCH();
    {
        // row pointer in the caller's row order, starting at the job's base offset (a value
        // that is only known on the device when jobs are chained)
        cub::TransformInputIterator<int64_t, CastI64, const int32_t *> it(row_cnt, CastI64());
        cub::FutureValue<int64_t> base(job->ctr.as<int64_t>() + CTR_BASE);  // (iterator type int64_t*: any other would be read on the HOST through operator T())
        size_t tb = job->scan_tmp_bytes;
        XG_CUDA(cub::DeviceScan::ExclusiveScan(job->scan_tmp.p, tb, it, d_indptr, cub::Sum(), base, n_t + 1, s));
    }
    XG_CUDA(cudaMemcpyAsync(job->ctr.as<unsigned long long>() + CTR_NNZ_END, d_indptr + n_t, sizeof(int64_t),
                            cudaMemcpyDeviceToDevice, s));
    return XG_OK;
}

// the counters of the attempt into a pinned host slot (asynchronous), then the end mark
static int overlap_request_counters(xg_overlap *job) {
    if (!job->h_ctr) {
        job->h_ctr = pinned_slot_acquire();
        if (!job->h_ctr) { set_error("overlap: no pinned host memory"); return XG_ERR_NOMEM; }
    }
    static_assert(sizeof(unsigned long long) * CTR_N <= 128, "a pinned slot holds 128 bytes");
    XG_CUDA(cudaMemcpyAsync(job->h_ctr, job->ctr.p, sizeof(unsigned long long) * CTR_N, cudaMemcpyDeviceToHost, job->ls));
    XG_CUDA(job->mark());  // end
    return XG_OK;
}

// host side of a finished attempt (the stream has been synchronised behind
// overlap_request_counters): counters -> job
static void overlap_collect(xg_overlap *job) {
    unsigned long long hc[CTR_N];
    memcpy(hc, job->h_ctr, sizeof(hc));
    job->needed = hc[CTR_PAIRS];
    job->flags = hc[CTR_FLAGS];
    job->collect();
    job->nnz = (int64_t)hc[CTR_NNZ_END] - (int64_t)hc[CTR_BASE];
    job->nnz_end = (int64_t)hc[CTR_NNZ_END];
    job->n_cand = (int64_t)hc[CTR_CAND];
    job->n_surv = (int64_t)hc[CTR_NSURV];
    job->h_stats[0] = job->n_t;
    job->h_stats[1] = (int64_t)hc[CTR_VISITS];
    job->h_stats[2] = (int64_t)hc[CTR_TESTS];
    job->h_stats[3] = (int64_t)hc[CTR_NSURV];
    job->h_stats[4] = job->nnz;
    job->h_stats[5] = (int64_t)hc[CTR_CAND];
}

// one attempt with a pair arena of `cap` entries, waited for; *needed receives the number
// of arena slots the walk wanted (> cap means: run again with that capacity)
static int overlap_attempt(xg_overlap *job, unsigned long long cap, int64_t *d_indptr,
                           unsigned long long *needed) {
    int rc = overlap_enqueue(job, cap, d_indptr, false);
    if (rc) return rc;
    // the one host round trip of a job: nnz sizes the caller's arrays.  Read-backs go through
    // a pinned slot (a pageable destination would make the copies synchronous one by one)
    rc = overlap_request_counters(job);
    if (rc) return rc;
    XG_CUDA(cudaStreamSynchronize(job->s));
    overlap_collect(job);
    *needed = job->needed;
    return XG_OK;
}

// xg_overlap_run: the caller's arrays were sized by an upper bound.  If the rows do not fit
// (ctr[CTR_NNZ_END] is the exact end), the job flags itself, moves nothing and -- so that an
// apply already queued behind it reads empty rows instead of unwritten entries --
// k_void_rows points every row at the job's first entry.
__global__ void k_check_capacity(unsigned long long *__restrict__ ctr, int64_t entry_cap) {
    if ((int64_t)ctr[CTR_NNZ_END] > entry_cap) ctr[CTR_FLAGS] |= 1ull;
}
__global__ void __launch_bounds__(256)
k_void_rows(int64_t *__restrict__ indptr, int64_t n_t, const unsigned long long *__restrict__ ctr) {
    if (ctr[CTR_FLAGS] == 0ull) return;
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n_t) indptr[i + 1] = (int64_t)ctr[CTR_BASE];
}

// provisional rows -> the caller's CSR arrays (xg_overlap_finish / xg_overlap_run)
static int enqueue_move_rows(xg_overlap *job, const int64_t *d_indptr, int32_t *d_indices, double *d_area,
                             int64_t entry_cap, int relative, cudaStream_t s) {
    const xg_index *src = job->src;
    RowMove R;
    int32_t *cand_cnt = job->cnt.as<int32_t>();
    R.prov_ptr = job->prov_ptr.as<int64_t>(); R.fill_cnt = cand_cnt + (job->n_t + 1);
    R.prov_col = job->prov_col.as<int32_t>(); R.prov_val = job->pairs.as<double>();
    R.tperm = job->tperm.as<int32_t>(); R.t_begin = job->t_begin;
    R.sperm = src->perm; R.spoly = (const double2 *)src->poly; R.snv = src->nvert; R.ks = src->k;
    R.relative = relative; R.indptr = d_indptr; R.indices = d_indices; R.data = d_area;
    R.entry_cap = entry_cap; R.flags = job->ctr.as<unsigned long long>() + CTR_FLAGS;
    if (entry_cap != INT64_MAX) {
        k_check_capacity<<<1, 1, 0, s>>>(job->ctr.as<unsigned long long>(), entry_cap);
        XG_CHECK_LAUNCH();
    }
    // rows longer than SORT_SHORT are queued in the (now dead) target box buffer
    unsigned long long *n_long = job->ctr.as<unsigned long long>() + CTR_LONG;
    int32_t *long_rows = job->tbox.as<int32_t>();
    k_move_rows<<<(unsigned)div_up(job->n_t, 256), 256, 0, s>>>(R, job->n_t, long_rows, n_long);
    unsigned long long *n_huge = job->ctr.as<unsigned long long>() + CTR_HUGE;
    int32_t *huge_rows = long_rows + job->n_t;  // second n_t ints of the same 32 * n_t byte buffer
    k_move_long_rows<<<num_sms() * 4, 128, 0, s>>>(R, long_rows, n_long, huge_rows, n_huge);
    k_move_huge_rows<<<num_sms(), 1024, 0, s>>>(R, huge_rows, n_huge);
    XG_CHECK_LAUNCH();
    if (entry_cap != INT64_MAX) {
        k_void_rows<<<(unsigned)div_up(job->n_t, 256), 256, 0, s>>>(const_cast<int64_t *>(d_indptr), job->n_t,
                                                                  job->ctr.as<unsigned long long>());
        XG_CHECK_LAUNCH();
    }
    return XG_OK;
}

// buffers every attempt needs + the packed targets; returns the default arena capacity
// Every scratch buffer of a job is allocated up front (nothing is allocated while a job is being
// recorded into a graph): job buffers, the scratch of the target packing, and -- for a job that
// will not be re-sized (xg_overlap_run) -- the arenas of the default capacity.
static unsigned long long default_arena(int64_t n_t) {
    unsigned long long cap = (unsigned long long)n_t * PAIRS_PER_TARGET_GUESS;
    return cap < 65536 ? 65536 : cap;
}

static int overlap_alloc(xg_overlap *job, bool with_arenas) {
    cudaStream_t s = job->s;
    const int64_t n_t = job->n_t;
    const int k_t = job->k_t;
    XG_CUDA(job->ctr.alloc(sizeof(unsigned long long) * CTR_N, s));
    XG_CUDA(job->tperm.alloc(sizeof(int32_t) * (size_t)n_t, s));
    XG_CUDA(job->tbox.alloc(sizeof(Box) * (size_t)n_t, s));
    XG_CUDA(job->tpoly.alloc(sizeof(double) * 2 * (size_t)k_t * (size_t)n_t, s));
    XG_CUDA(job->tnv.alloc((size_t)n_t, s));
    XG_CUDA(job->cnt.alloc(sizeof(int32_t) * 3 * (size_t)(n_t + 1), s));
    XG_CUDA(job->prov_ptr.alloc(sizeof(int64_t) * (size_t)(n_t + 1), s));
    {
        // one scratch buffer serves both scans of an attempt
        cub::TransformInputIterator<int64_t, CastI64, const int32_t *> it(job->cnt.as<int32_t>(), CastI64());
        size_t tb_sum = 0, tb_scan = 0;
        XG_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tb_sum, it, job->prov_ptr.as<int64_t>(), n_t + 1, s));
        cub::FutureValue<int64_t> base(job->ctr.as<int64_t>() + CTR_BASE);  // (iterator type int64_t*: any other would be read on the HOST through operator T())
        XG_CUDA(cub::DeviceScan::ExclusiveScan(nullptr, tb_scan, it, job->prov_ptr.as<int64_t>(), cub::Sum(), base,
                                               n_t + 1, s));
        job->scan_tmp_bytes = tb_sum > tb_scan ? tb_sum : tb_scan;
        XG_CUDA(job->scan_tmp.alloc(job->scan_tmp_bytes, s));
    }
    // scratch of pack_mesh (keys in / out, ids, sort)
    XG_CUDA(job->pk_keys_in.alloc(sizeof(uint32_t) * (size_t)n_t, s));
    XG_CUDA(job->pk_keys_out.alloc(sizeof(uint32_t) * (size_t)n_t, s));
    XG_CUDA(job->pk_ids.alloc(sizeof(int32_t) * (size_t)n_t, s));
    job->pk.sort_tmp_bytes = pack_sort_tmp_bytes(n_t);
    XG_CUDA(job->pk_tmp.alloc(job->pk.sort_tmp_bytes, s));
    job->pk.keys_in = job->pk_keys_in.as<uint32_t>(); job->pk.keys_out = job->pk_keys_out.as<uint32_t>();
    job->pk.ids_in = job->pk_ids.as<int32_t>(); job->pk.sort_tmp = job->pk_tmp.p;
    if (with_arenas) {
        const unsigned long long cap = default_arena(n_t);
        XG_CUDA(job->pairs.alloc(sizeof(int2) * (size_t)cap, s));
        XG_CUDA(job->surv.alloc(sizeof(int2) * ((size_t)cap + (size_t)num_sms() * 64 * SAT_WIN), s));
        XG_CUDA(job->prov_col.alloc(sizeof(int32_t) * (size_t)cap, s));
    }
    return XG_OK;
}

// targets of the job in curve order (launches only)
static int overlap_pack(xg_overlap *job, const double *d_t_xy, const void *d_t_faces, int idx_bytes,
                        const int32_t *d_t_ids, int64_t t_begin, int64_t t_end) {
    XG_CUDA(job->mark());  // XG_PHASE_PACK
    PackedMesh pm;
    pm.perm = job->tperm.as<int32_t>(); pm.box = job->tbox.as<Box>();
    pm.poly = job->tpoly.as<double>(); pm.nvert = job->tnv.as<uint8_t>();
    // (with a face list, tperm holds list positions t_begin + i: rows stay tperm - t_begin)
    return pack_mesh(d_t_xy, d_t_faces, idx_bytes, t_begin, t_end, job->k_t, job->src->d_frame, nullptr, &pm, job->ls,
                     nullptr, nullptr, d_t_ids, 0, job->s, &job->pk);
}

// arena capacity of the waiting path: the default, or a sampled estimate when the source is
// much finer than the targets
static int overlap_capacity(xg_overlap *job, const int32_t *d_t_ids, int64_t n_t_mesh, unsigned long long *cap_out) {
    const xg_index *source = job->src;
    const int64_t n_t = job->n_t;
    unsigned long long cap = default_arena(n_t);
    // "source much finer than the targets" compares meshes over the same area: an index
    // culled to these targets against the targets, a whole-mesh index against the whole
    // target mesh -- not against the few rows of one range of it
    const int64_t n_t_area = (source->n_face < source->n_face_mesh || d_t_ids) ? n_t : (n_t_mesh > n_t ? n_t_mesh : n_t);
    if (source->n_face > 2 * n_t_area && n_t >= 32 * SAMPLE_STRIDE * 8) {
        unsigned long long est = 0;
        const int rc = estimate_candidates(job, &est);  // (one host round trip)
        if (rc) return rc;
        if (est > cap) cap = est;
    }
    *cap_out = cap;
    return XG_OK;
}

static int overlap_check_args(const char *who, const xg_index *source, const void *d_indptr, const int32_t *d_t_ids,
                              int idx_bytes, int64_t n_t_face, int k_t, int64_t t_begin, int64_t t_end) {
    if (!source || !d_indptr) { set_error("%s: NULL argument", who); return XG_ERR_INVALID; }
    if (t_begin < 0 || t_end < t_begin || (!d_t_ids && t_end > n_t_face) || (idx_bytes != 4 && idx_bytes != 8)) {
        set_error("%s: invalid target range [%lld,%lld) of %lld", who, (long long)t_begin, (long long)t_end,
                  (long long)n_t_face);
        return XG_ERR_INVALID;
    }
    if (k_t < 3 || k_t > XG_MAX_N_VERTEX) { set_error("%s: unsupported k_t=%d", who, k_t); return XG_ERR_UNSUPPORTED; }
    return XG_OK;
}

// d_t_ids (optional): the targets are the faces d_t_ids[t_begin .. t_end) and row i of the
// result belongs to d_t_ids[t_begin + i]
static int overlap_begin_impl(const xg_index *source, const double *d_t_xy, int64_t n_t_node,
                              const void *d_t_faces, int idx_bytes, int64_t n_t_face, int k_t,
                              const int32_t *d_t_ids, int64_t t_begin, int64_t t_end, int64_t *d_indptr,
                              int64_t *h_nnz, void *stream, xg_overlap **job_out) {
    (void)n_t_node;
    if (!job_out || !h_nnz) { set_error("xg_overlap_begin: NULL argument"); return XG_ERR_INVALID; }
    *job_out = nullptr;
    int rc = overlap_check_args("xg_overlap_begin", source, d_indptr, d_t_ids, idx_bytes, n_t_face, k_t, t_begin, t_end);
    if (rc) return rc;
    cudaStream_t s = (cudaStream_t)stream;
    index_note_use(source, s);
    const int64_t n_t = t_end - t_begin;
    xg_overlap *job = new xg_overlap();
    job->src = source; job->t_begin = t_begin; job->n_t = n_t; job->k_t = k_t; job->s = s; job->ls = s;
#define OV_CUDA(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { rc = cuda_fail(e__, #call, __FILE__, __LINE__); goto fail; } } while (0)
    {
        if (n_t == 0 || source->n_face == 0) {
            OV_CUDA(cudaMemsetAsync(d_indptr, 0, sizeof(int64_t) * (size_t)(n_t + 1), s));
            OV_CUDA(cudaStreamSynchronize(s));
            *h_nnz = 0;
            job->h_stats[0] = n_t;
            job->stats_valid = true;
            *job_out = job;
            return XG_OK;
        }
        unsigned long long cap = 0;
        rc = overlap_alloc(job, false);
        if (rc) goto fail;
        rc = overlap_pack(job, d_t_xy, d_t_faces, idx_bytes, d_t_ids, t_begin, t_end);
        if (rc) goto fail;
        rc = overlap_capacity(job, d_t_ids, n_t_face, &cap);
        if (rc) goto fail;
        unsigned long long needed = 0;
        rc = overlap_attempt(job, cap, d_indptr, &needed);
        if (rc) goto fail;
        job->h_stats[6] = 1;  // walks over all tiles
        if (needed > cap) {
            // denser than guessed: the first walk counted exactly what is needed
            rc = overlap_attempt(job, needed, d_indptr, &needed);
            if (rc) goto fail;
            job->h_stats[6] = 2;
        }
        job->stats_valid = true;
        *h_nnz = job->nnz;
        *job_out = job;
        return XG_OK;
    }
fail:
    delete job;
    return rc;
#undef OV_CUDA
}

extern "C" int xg_overlap_begin(const xg_index *source, const double *d_t_xy, int64_t n_t_node,
                                const void *d_t_faces, int idx_bytes, int64_t n_t_face, int k_t,
                                int64_t t_begin, int64_t t_end, int64_t *d_indptr, int64_t *h_nnz,
                                void *stream, xg_overlap **job_out) {
    return overlap_begin_impl(source, d_t_xy, n_t_node, d_t_faces, idx_bytes, n_t_face, k_t, nullptr, t_begin,
                              t_end, d_indptr, h_nnz, stream, job_out);
}

extern "C" int xg_overlap_begin_ids(const xg_index *source, const double *d_t_xy, int64_t n_t_node,
                                    const void *d_t_faces, int idx_bytes, int64_t n_t_face, int k_t,
                                    const int32_t *d_t_ids, int64_t n_ids, int64_t *d_indptr,
                                    int64_t *h_nnz, void *stream, xg_overlap **job_out) {
    if (n_ids < 0 || (n_ids > 0 && !d_t_ids)) { set_error("xg_overlap_begin_ids: invalid face list"); return XG_ERR_INVALID; }
    if (n_ids == 0)  // nothing to list: an empty range of the mesh itself
        return overlap_begin_impl(source, d_t_xy, n_t_node, d_t_faces, idx_bytes, n_t_face, k_t, nullptr, 0, 0,
                                  d_indptr, h_nnz, stream, job_out);
    return overlap_begin_impl(source, d_t_xy, n_t_node, d_t_faces, idx_bytes, n_t_face, k_t, d_t_ids, 0, n_ids,
                              d_indptr, h_nnz, stream, job_out);
}

extern "C" int xg_overlap_finish(xg_overlap *job, const int64_t *d_indptr, int32_t *d_indices,
                                 double *d_area, int relative, void *stream) {
    if (!job) { set_error("xg_overlap_finish: NULL job"); return XG_ERR_INVALID; }
    cudaStream_t s = (cudaStream_t)stream;
    int rc = XG_OK;
    if (job->nnz > 0) {
        if (!d_indices || !d_area || !d_indptr) { set_error("xg_overlap_finish: NULL output"); delete job; return XG_ERR_INVALID; }
        rc = enqueue_move_rows(job, d_indptr, d_indices, d_area, INT64_MAX, relative, s);
    }
    // buffers are released stream-ordered on the stream that produced them
    if (s != job->s) cudaStreamSynchronize(s);
    delete job;
    return rc;
}

// a private stream per host thread on which jobs are recorded into graphs (never executes anything)
static cudaStream_t capture_stream() {
    static thread_local cudaStream_t cs[64] = {nullptr};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
    if (!cs[dev] && cudaStreamCreateWithFlags(&cs[dev], cudaStreamNonBlocking) != cudaSuccess) {
        cudaGetLastError();
        cs[dev] = nullptr;
    }
    return cs[dev];
}

// The whole job enqueued at once: no host round trip between the walk and the caller's
// CSR arrays, which are sized by an upper bound instead of the nnz read-back.
extern "C" int xg_overlap_run(const xg_index *source, const double *d_t_xy, int64_t n_t_node,
                              const void *d_t_faces, int idx_bytes, int64_t n_t_face, int k_t,
                              const int32_t *d_t_ids, int64_t t_begin, int64_t t_end, int flags,
                              int64_t *d_indptr, int32_t *d_indices, double *d_area, int64_t entry_cap,
                              int relative, void *stream, xg_overlap **job_out) {
    (void)n_t_node;
    const int chained = flags & XG_RUN_CHAINED;
    // One CUDA graph per job: its ~30 small launches (memsets, key / sort / pack of the targets,
    // scans, row movers) cost 25-40 us EACH while the copy engine saturates PCIe with the rest of
    // a mesh upload -- the front end fetches every launch over the same link -- which made a
    // 1 M-row job 0.5 ms slower under an upload than on resident data.  Captured, the job is one
    // submission.  Capture needs every scratch block to come from the cache: a miss ends the
    // capture and the job is enqueued directly (that fills the cache for the next time).
    bool as_graph = (flags & XG_RUN_GRAPH) != 0;
    {
        cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
        if (as_graph && (cudaStreamIsCapturing((cudaStream_t)stream, &st) != cudaSuccess || st != cudaStreamCaptureStatusNone))
            as_graph = false;  // (the caller is capturing already)
    }
    if (!job_out || !d_indices || !d_area || entry_cap < 0) { set_error("xg_overlap_run: invalid argument"); return XG_ERR_INVALID; }
    *job_out = nullptr;
    int rc = overlap_check_args("xg_overlap_run", source, d_indptr, d_t_ids, idx_bytes, n_t_face, k_t, t_begin, t_end);
    if (rc) return rc;
    cudaStream_t s = (cudaStream_t)stream;
    index_note_use(source, s);
    const int64_t n_t = t_end - t_begin;
    xg_overlap *job = new xg_overlap();
    job->src = source; job->t_begin = t_begin; job->n_t = n_t; job->k_t = k_t; job->s = s; job->ls = s;
    job->deferred = true;
    job->h_stats[0] = n_t;
    job->h_stats[6] = 1;
    bool capturing = false;
    auto fail = [&](int code) {
        if (capturing) {
            if (getenv("XG_GRAPH_DEBUG")) fprintf(stderr, "[xg] capture of an overlap job given up (rc %d): %s\n", code, xg_last_error_string());
            scratch_set_capturing(false);
            cudaGraph_t g = nullptr;
            cudaStreamEndCapture(job->ls, &g);  // (invalidated by the error: nothing was enqueued)
            if (g) cudaGraphDestroy(g);
            cudaGetLastError();
        }
        delete job;
        if (capturing)  // once more, enqueued directly
            return xg_overlap_run(source, d_t_xy, n_t_node, d_t_faces, idx_bytes, n_t_face, k_t, d_t_ids, t_begin, t_end,
                                  flags & ~XG_RUN_GRAPH, d_indptr, d_indices, d_area, entry_cap, relative, stream, job_out);
        return code;
    };
    if (n_t == 0 || source->n_face == 0) {
        // every row is empty: all row pointers = the job's first entry
        cudaError_t e = cudaSuccess;
        if (!chained) e = cudaMemsetAsync(d_indptr, 0, sizeof(int64_t) * (size_t)(n_t + 1), s);
        else if (n_t > 0) {
            k_fill_from<<<(unsigned)div_up(n_t, 256), 256, 0, s>>>(d_indptr + 1, n_t, d_indptr);
            e = cudaGetLastError();
        }
        if (e != cudaSuccess) return fail(cuda_fail(e, "xg_overlap_run (empty job)", __FILE__, __LINE__));
        job->empty = true;
        *job_out = job;
        return XG_OK;
    }
    rc = overlap_alloc(job, true);
    if (rc) return fail(rc);
    if (as_graph) {
        job->h_ctr = pinned_slot_acquire();  // (no host allocation inside the capture either)
        if (!job->h_ctr) { set_error("overlap: no pinned host memory"); return fail(XG_ERR_NOMEM); }
        // (the caller's stream may be the legacy default stream, which cannot be captured: the job
        //  is recorded on a private stream and the graph launched on the caller's)
        cudaStream_t cs = capture_stream();
        if (cs && cudaStreamBeginCapture(cs, cudaStreamCaptureModeThreadLocal) == cudaSuccess) {
            capturing = true;
            job->timing = false;
            job->ls = cs;
            scratch_set_capturing(true);
        } else {
            cudaGetLastError();
            as_graph = false;
        }
    }
    // (never sampled -- that would be a host round trip: a job that is denser than the default
    //  arena reports itself incomplete and its caller takes the waiting path)
    const unsigned long long cap = default_arena(n_t);
    rc = overlap_pack(job, d_t_xy, d_t_faces, idx_bytes, d_t_ids, t_begin, t_end);
    if (rc) return fail(rc);
    rc = overlap_enqueue(job, cap, d_indptr, chained != 0);
    if (rc) return fail(rc);
    rc = enqueue_move_rows(job, d_indptr, d_indices, d_area, entry_cap, relative, job->ls);
    if (rc) return fail(rc);
    rc = overlap_request_counters(job);
    if (rc) return fail(rc);
    if (capturing) {
        scratch_set_capturing(false);
        cudaError_t e = cudaStreamEndCapture(job->ls, &job->graph);
        job->ls = s;
        if (e == cudaSuccess) e = cudaGraphInstantiate(&job->graph_exec, job->graph, 0);
        if (e != cudaSuccess) {
            // (nothing has been enqueued; fail() with capturing still set would end the capture twice)
            if (getenv("XG_GRAPH_DEBUG")) fprintf(stderr, "[xg] end of capture / instantiate failed: %s\n", cudaGetErrorString(e));
            cudaGetLastError();
            capturing = false;
            delete job;
            return xg_overlap_run(source, d_t_xy, n_t_node, d_t_faces, idx_bytes, n_t_face, k_t, d_t_ids, t_begin, t_end,
                                  flags & ~XG_RUN_GRAPH, d_indptr, d_indices, d_area, entry_cap, relative, stream, job_out);
        }
        capturing = false;
        e = cudaGraphLaunch(job->graph_exec, s);
        if (e != cudaSuccess) return fail(cuda_fail(e, "cudaGraphLaunch(overlap job)", __FILE__, __LINE__));
        job->h_stats[7] = 1;  // launched as a graph
    }
    {
        cudaError_t e = cudaEventCreateWithFlags(&job->done, cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventRecord(job->done, s);
        if (e != cudaSuccess) return fail(cuda_fail(e, "xg_overlap_run (event)", __FILE__, __LINE__));
    }
    // the scratch goes back to the block cache in stream order: the job that follows on
    // this stream re-uses it while the host is still far ahead
    job->release_scratch();
    *job_out = job;
    return XG_OK;
}

extern "C" int xg_overlap_result(xg_overlap *job, int64_t *h_nnz, int64_t *h_nnz_end, int *h_complete) {
    if (!job || !job->deferred) { set_error("xg_overlap_result: not a job of xg_overlap_run"); return XG_ERR_INVALID; }
    if (!job->stats_valid) {
        if (!job->empty) {
            XG_CUDA(cudaEventSynchronize(job->done));
            overlap_collect(job);
        }
        job->stats_valid = true;
    }
    // incomplete: the pair arena was too small (nothing was written) or rows did not fit
    // the caller's arrays; the caller runs the job again through xg_overlap_begin
    const bool complete = job->empty || (job->needed <= job->cap && job->flags == 0);
    if (h_nnz) *h_nnz = job->nnz;
    if (h_nnz_end) *h_nnz_end = job->nnz_end;
    if (h_complete) *h_complete = complete ? 1 : 0;
    return XG_OK;
}

extern "C" int xg_overlap_stats(const xg_overlap *job, int64_t stats[8]) {
    if (!job || !stats || !job->stats_valid) { set_error("xg_overlap_stats: no stats"); return XG_ERR_INVALID; }
    for (int i = 0; i < 8; ++i) stats[i] = job->h_stats[i];
    return XG_OK;
}

extern "C" int xg_overlap_timings(const xg_overlap *job, double ms[8]) {
    if (!job || !ms || !job->stats_valid) { set_error("xg_overlap_timings: no timings"); return XG_ERR_INVALID; }
    for (int i = 0; i < 8; ++i) ms[i] = job->h_ms[i];
    return XG_OK;
}

extern "C" int xg_overlap_abort(xg_overlap *job) {
    delete job;
    return XG_OK;
}
