// xg_overlap.cu -- overlap weights: target faces x source faces -> CSR of areas.
//
// Replaces UnstructuredGrid2d.overlap -> CellTree2d.intersect_faces
// (/root/reference/xugrid/regrid/unstructured.py:109-135: build_bboxes ->
// locate_boxes (count pass, cumsum, fill pass) -> separating-axis filter ->
// Sutherland-Hodgman area -> keep area > 0) and MatrixCSR.from_triplet
// (/root/reference/xugrid/core/sparse.py:61-78,119-127).
//
// Pipeline (all on the caller's stream):
//   pack_mesh(targets)            targets in Morton order, self-contained records
//   k_traverse<false>             one warp walks the wide BVH for a TILE of 32
//                                 Morton-consecutive targets (lane = target):
//                                 child boxes are loaded once per warp, coalesced,
//                                 and tested by all lanes from shared memory;
//                                 hit masks are combined with warp votes
//   cub ExclusiveSum              candidate row pointer
//   k_traverse<true>              same walk, writes (target, source) candidates
//   k_clip                        one thread per candidate: SAT + clip area (fp64)
//   k_row_count + cub ExclusiveSum  CSR row pointer in the caller's row order
//   k_emit (xg_overlap_finish)    rows written in ascending source index
#include <cub/device/device_scan.cuh>
#include <cub/iterator/transform_input_iterator.cuh>

#include "xg_common.cuh"
#include "xg_geom.cuh"
#include "xg_pack.cuh"

namespace xg {

constexpr int TRAV_WARPS = 8;
constexpr int TRAV_STACK = 256;

struct TraverseParams {
    const Box *boxes[MAX_LEVELS + 1];  // [0] = leaves, [d] = node level d-1
    int64_t count[MAX_LEVELS + 1];
    int top;                           // depth of the virtual root
};

__device__ __forceinline__ bool boxes_intersect(const Box &a, const Box &b) {
    // strict: touching boxes do not intersect
    return a.xmin < b.xmax && b.xmin < a.xmax && a.ymin < b.ymax && b.ymin < a.ymax;
}

// stats: [0] node visits (warp-level, each one coalesced load of <= 32 boxes)
//        [1] exact leaf bbox tests  [2] candidates
template <bool FILL>
__global__ void __launch_bounds__(TRAV_WARPS * 32)
k_traverse(const TraverseParams P, const Box *__restrict__ tbox, int64_t n_t,
           int32_t *__restrict__ cand_cnt, const int64_t *__restrict__ cand_ptr,
           int32_t *__restrict__ cand_src, int32_t *__restrict__ cand_tgt,
           unsigned long long *__restrict__ stats) {
    __shared__ Box s_node[TRAV_WARPS][32];
    __shared__ int32_t s_stack[TRAV_WARPS][TRAV_STACK];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t tile = blockIdx.x * (int64_t)TRAV_WARPS + w;
    const int64_t t = tile * 32 + lane;
    if (tile * 32 >= n_t) return;
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
    int sp = 0;
    if (lane == 0) s_stack[w][0] = P.top << 27;
    sp = 1;
    int cnt = 0;
    const int64_t base = (FILL && valid) ? cand_ptr[t] : 0;
    unsigned visits = 0, tests = 0;
    __syncwarp();
    while (sp > 0) {
        const int32_t entry = s_stack[w][--sp];
        const int d = entry >> 27;
        const int64_t c0 = (int64_t)(entry & 0x7ffffff) * FANOUT;
        const Box *cb = P.boxes[d - 1];
        const int64_t nc = P.count[d - 1] - c0;
        Box mine;
        if (lane < nc) mine = cb[c0 + lane];
        else { mine.xmin = INFINITY; mine.xmax = -INFINITY; mine.ymin = INFINITY; mine.ymax = -INFINITY; }
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
            while (m) {
                const int c = __ffs(m) - 1;
                m &= m - 1;
                const Box b = s_node[w][c];
                if (boxes_intersect(tb, b)) {
                    if (FILL) {
                        cand_src[base + cnt] = (int32_t)(c0 + c);
                        cand_tgt[base + cnt] = (int32_t)t;
                    }
                    ++cnt;
                }
            }
        }
        __syncwarp();
    }
    if (!FILL) {
        if (valid) cand_cnt[t] = cnt;
        if (stats) {
            unsigned c = cnt;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
            if (lane == 0) {
                atomicAdd(stats + 0, (unsigned long long)visits);
                atomicAdd(stats + 1, (unsigned long long)tests * 32ull);
                atomicAdd(stats + 2, (unsigned long long)c);
            }
        }
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

// one thread per candidate pair: separating-axis filter, then clip area.
// area[p] = 0 for pairs that do not intersect.
template <int KT, int KS>
__global__ void __launch_bounds__(128)
k_clip(const int32_t *__restrict__ cand_tgt, const int32_t *__restrict__ cand_src, int64_t n_pairs,
       const double2 *__restrict__ tpoly, const uint8_t *__restrict__ tnv, int kt,
       const double2 *__restrict__ spoly, const uint8_t *__restrict__ snv, int ks,
       double *__restrict__ area, unsigned long long *__restrict__ stats) {
    extern __shared__ double2 s_buf[];
    const int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int stride = blockDim.x;
    double2 *bufA = s_buf + threadIdx.x;
    double2 *bufB = s_buf + (size_t)(KT + KS) * stride + threadIdx.x;
    bool reached = false;
    if (p < n_pairs) {
        double2 a[KT], b[KS];
        const int na = load_poly<KT>(tpoly, tnv, cand_tgt[p], kt, a);
        const int nb = load_poly<KS>(spoly, snv, cand_src[p], ks, b);
        double ar = 0.0;
        if (polygons_intersect<KT, KS>(a, na, b, nb)) {
            reached = true;
            ar = clip_area<KT, KS>(a, na, b, nb, bufA, bufB, stride);
        }
        area[p] = ar;
    }
    if (stats) {
        const unsigned m = __ballot_sync(0xffffffffu, reached);
        if ((threadIdx.x & 31) == 0 && m) atomicAdd(stats + 3, (unsigned long long)__popc(m));
    }
}

// positives per sorted target -> row count at the caller's row position
__global__ void __launch_bounds__(256)
k_row_count(const int64_t *__restrict__ cand_ptr, const double *__restrict__ area,
            const int32_t *__restrict__ tperm, int64_t t_begin, int64_t n_t,
            int32_t *__restrict__ row_cnt) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n_t) return;
    int c = 0;
    for (int64_t p = cand_ptr[i]; p < cand_ptr[i + 1]; ++p) c += area[p] > 0.0;
    row_cnt[tperm[i] - t_begin] = c;
}

// area of a packed face with the reference's arithmetic (connectivity.py:615-633:
// shoelace relative to the first vertex of the ORIGINAL vertex order)
__device__ __forceinline__ double packed_face_area(const double2 *__restrict__ poly, int k, uint8_t nvf) {
    const int n = nvf & 0x7f;
    const bool flip = nvf & 0x80;
    (void)k;
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

// one thread per sorted target: copy its positive pairs into the CSR row and
// order the row by ascending source face id
__global__ void __launch_bounds__(256)
k_emit(const int64_t *__restrict__ cand_ptr, const int32_t *__restrict__ cand_src,
       const double *__restrict__ area, const int32_t *__restrict__ tperm, int64_t t_begin,
       int64_t n_t, const int32_t *__restrict__ sperm, const double2 *__restrict__ spoly,
       const uint8_t *__restrict__ snv, int ks, int relative, const int64_t *__restrict__ indptr,
       int32_t *__restrict__ indices, double *__restrict__ data) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n_t) return;
    const int64_t dst = indptr[tperm[i] - t_begin];
    int n = 0;
    for (int64_t p = cand_ptr[i]; p < cand_ptr[i + 1]; ++p) {
        double a = area[p];
        if (!(a > 0.0)) continue;
        const int32_t s = cand_src[p];
        const int32_t sid = sperm[s];
        if (relative) a = a / packed_face_area(spoly + (int64_t)s * ks, ks, snv[s]);
        // insertion into the sorted prefix
        int j = n - 1;
        while (j >= 0 && indices[dst + j] > sid) {
            indices[dst + j + 1] = indices[dst + j];
            data[dst + j + 1] = data[dst + j];
            --j;
        }
        indices[dst + j + 1] = sid;
        data[dst + j + 1] = a;
        ++n;
    }
}

// Clipping as two dense kernels (the dominant cost of the build).
//   k_sat   one thread per candidate pair: separating-axis filter, straight-line
//           code.  Rejected pairs get area 0; survivors are appended to a
//           compact list as 16-byte records {pair index, target, source}
//           (block-aggregated: one global atomic per block of 256 pairs, so a
//           block's survivors stay contiguous and neighbours in space).
//   k_clip_survivors  one thread per survivor: every lane of a warp clips a
//           pair that really intersects (clip_area_dense), so the fp64 pipe
//           sees full warps and no lane idles through another lane's early out.
// The survivor count stays on the device: k_clip_survivors is launched over
// the candidate count and surplus threads exit after reading the counter.
constexpr int CLIP_T = 256;

struct __align__(16) Survivor {
    int64_t pair;
    int32_t tgt, src;
};

// rare slow path of clip_area_dense(), kept out of line so that it does not
// cost the main path registers
template <int KT, int KS>
__device__ __noinline__ double clip_area_fallback(const double2 *pa, int na, const double2 *pb, int nb,
                                                  double2 *bufA, double2 *bufB, int stride) {
    double2 a[KT], b[KS];
#pragma unroll
    for (int i = 0; i < KT; ++i) a[i] = pa[i];
#pragma unroll
    for (int i = 0; i < KS; ++i) b[i] = pb[i];
    return clip_area<KT, KS>(a, na, b, nb, bufA, bufB, stride);
}

template <int KT, int KS>
__global__ void __launch_bounds__(CLIP_T)
k_sat(const int32_t *__restrict__ cand_tgt, const int32_t *__restrict__ cand_src, int64_t n_pairs,
      const double2 *__restrict__ tpoly, const uint8_t *__restrict__ tnv, int kt,
      const double2 *__restrict__ spoly, const uint8_t *__restrict__ snv, int ks,
      double *__restrict__ area, Survivor *__restrict__ surv, unsigned long long *__restrict__ stats) {
    __shared__ int s_cnt;
    __shared__ unsigned long long s_base;
    const int tid = threadIdx.x, lane = tid & 31;
    if (tid == 0) s_cnt = 0;
    __syncthreads();
    const int64_t p = blockIdx.x * (int64_t)CLIP_T + tid;
    bool pass = false;
    int32_t t = 0, s = 0;
    if (p < n_pairs) {
        t = cand_tgt[p];
        s = cand_src[p];
        double2 a[KT], b[KS];
        const int na = load_poly<KT>(tpoly, tnv, t, kt, a);
        const int nb = load_poly<KS>(spoly, snv, s, ks, b);
        pass = polygons_intersect_uniform<KT, KS>(a, na, b, nb);
        if (!pass) area[p] = 0.0;
    }
    const unsigned m = __ballot_sync(0xffffffffu, pass);
    int pos = 0;
    if (m) {
        if (lane == __ffs(m) - 1) pos = atomicAdd(&s_cnt, __popc(m));
        pos = __shfl_sync(0xffffffffu, pos, __ffs(m) - 1) + __popc(m & ((1u << lane) - 1u));
    }
    __syncthreads();
    if (tid == 0 && s_cnt > 0) s_base = atomicAdd(stats + 3, (unsigned long long)s_cnt);
    __syncthreads();
    if (pass) {
        Survivor r;
        r.pair = p; r.tgt = t; r.src = s;
        surv[s_base + pos] = r;
    }
}

template <int KT, int KS>
__global__ void __launch_bounds__(CLIP_T, 3)
k_clip_survivors(const Survivor *__restrict__ surv, const unsigned long long *__restrict__ n_surv,
                 const double2 *__restrict__ tpoly, const uint8_t *__restrict__ tnv, int kt,
                 const double2 *__restrict__ spoly, const uint8_t *__restrict__ snv, int ks,
                 double *__restrict__ area) {
    extern __shared__ double2 s_buf[];
    const int64_t q = blockIdx.x * (int64_t)CLIP_T + threadIdx.x;
    if (q >= (int64_t)*n_surv) return;
    const Survivor r = surv[q];
    double2 *bufA = s_buf + threadIdx.x;
    double2 *bufB = s_buf + (size_t)(KT + KS) * CLIP_T + threadIdx.x;
    double2 a[KT], b[KS];
    const int na = load_poly<KT>(tpoly, tnv, r.tgt, kt, a);
    const int nb = load_poly<KS>(spoly, snv, r.src, ks, b);
    bool fb = false;
    double ar = clip_area_dense<KT, KS>(a, na, b, nb, bufA, bufB, CLIP_T, &fb);
    if (fb) ar = clip_area_fallback<KT, KS>(a, na, b, nb, bufA, bufB, CLIP_T);
    area[r.pair] = ar;
}

template <int KT, int KS>
static int launch_clip_dense(const int32_t *ct, const int32_t *cs, int64_t n_pairs,
                             const double *tpoly, const uint8_t *tnv, int kt, const double *spoly,
                             const uint8_t *snv, int ks, double *area, void *surv,
                             unsigned long long *stats, cudaStream_t s) {
    const unsigned grid = (unsigned)div_up(n_pairs, CLIP_T);
    k_sat<KT, KS><<<grid, CLIP_T, 0, s>>>(ct, cs, n_pairs, (const double2 *)tpoly, tnv, kt,
                                          (const double2 *)spoly, snv, ks, area, (Survivor *)surv, stats);
    XG_CHECK_LAUNCH();
    const size_t smem = sizeof(double2) * 2 * (size_t)(KT + KS) * CLIP_T;
    XG_CUDA(cudaFuncSetAttribute(k_clip_survivors<KT, KS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_clip_survivors<KT, KS><<<grid, CLIP_T, smem, s>>>((const Survivor *)surv, stats + 3, (const double2 *)tpoly, tnv, kt,
                                                        (const double2 *)spoly, snv, ks, area);
    XG_CHECK_LAUNCH();
    return XG_OK;
}

struct CastI64 {
    __host__ __device__ __forceinline__ int64_t operator()(const int32_t &v) const { return (int64_t)v; }
};

template <int KT, int KS>
static int launch_clip(int block, const int32_t *ct, const int32_t *cs, int64_t n_pairs,
                       const double *tpoly, const uint8_t *tnv, int kt, const double *spoly,
                       const uint8_t *snv, int ks, double *area, unsigned long long *stats,
                       cudaStream_t s) {
    const size_t smem = sizeof(double2) * 2 * (size_t)(KT + KS) * block;
    if (smem > 48 * 1024)
        XG_CUDA(cudaFuncSetAttribute(k_clip<KT, KS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_clip<KT, KS><<<(unsigned)div_up(n_pairs, block), block, smem, s>>>(
        ct, cs, n_pairs, (const double2 *)tpoly, tnv, kt, (const double2 *)spoly, snv, ks, area, stats);
    XG_CHECK_LAUNCH();
    return XG_OK;
}

}  // namespace xg

using namespace xg;

struct xg_overlap {
    const xg_index *src = nullptr;
    int64_t t_begin = 0, n_t = 0;
    int k_t = 0;
    int64_t n_cand = 0, nnz = 0;
    cudaStream_t s = nullptr;
    // device buffers
    DevBuf tperm, tbox, tpoly, tnv, cand_ptr, cand_src, cand_tgt, area, stats;
    int64_t h_stats[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    bool stats_valid = false;
    // phase timing: ev[i] is recorded when phase i starts, ev[n_ev-1] at the end
    cudaEvent_t ev[XG_N_PHASES + 1];
    int n_ev = 0;
    double h_ms[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    xg_overlap() { for (int i = 0; i <= XG_N_PHASES; ++i) ev[i] = nullptr; }
    ~xg_overlap() { for (int i = 0; i <= XG_N_PHASES; ++i) if (ev[i]) cudaEventDestroy(ev[i]); }
    cudaError_t mark() {
        cudaError_t e = cudaEventCreate(&ev[n_ev]);
        if (e != cudaSuccess) return e;
        return cudaEventRecord(ev[n_ev++], s);
    }
    void collect() {
        for (int i = 0; i + 1 < n_ev; ++i) {
            float ms = 0.f;
            if (cudaEventElapsedTime(&ms, ev[i], ev[i + 1]) == cudaSuccess) h_ms[i] = ms;
        }
    }
};

extern "C" int xg_overlap_begin(const xg_index *source, const double *d_t_xy, int64_t n_t_node,
                                const void *d_t_faces, int idx_bytes, int64_t n_t_face, int k_t,
                                int64_t t_begin, int64_t t_end, int64_t *d_indptr, int64_t *h_nnz,
                                void *stream, xg_overlap **job_out) {
    (void)n_t_node;
    if (!source || !job_out || !d_indptr || !h_nnz) { set_error("xg_overlap_begin: NULL argument"); return XG_ERR_INVALID; }
    *job_out = nullptr;
    if (t_begin < 0 || t_end < t_begin || t_end > n_t_face || (idx_bytes != 4 && idx_bytes != 8)) {
        set_error("xg_overlap_begin: invalid target range [%lld,%lld) of %lld", (long long)t_begin,
                  (long long)t_end, (long long)n_t_face);
        return XG_ERR_INVALID;
    }
    if (k_t < 3 || k_t > XG_MAX_N_VERTEX) { set_error("xg_overlap_begin: unsupported k_t=%d", k_t); return XG_ERR_UNSUPPORTED; }
    cudaStream_t s = (cudaStream_t)stream;
    const int64_t n_t = t_end - t_begin;
    xg_overlap *job = new xg_overlap();
    job->src = source; job->t_begin = t_begin; job->n_t = n_t; job->k_t = k_t; job->s = s;
    int rc = XG_OK;
#define OV_CUDA(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { rc = cuda_fail(e__, #call, __FILE__, __LINE__); goto fail; } } while (0)
    {
        OV_CUDA(job->stats.alloc(sizeof(unsigned long long) * 8, s));
        OV_CUDA(cudaMemsetAsync(job->stats.p, 0, sizeof(unsigned long long) * 8, s));
        if (n_t == 0 || source->n_face == 0) {
            OV_CUDA(cudaMemsetAsync(d_indptr, 0, sizeof(int64_t) * (size_t)(n_t + 1), s));
            OV_CUDA(cudaStreamSynchronize(s));
            *h_nnz = 0;
            job->stats_valid = true;
            *job_out = job;
            return XG_OK;
        }
        OV_CUDA(job->mark());  // XG_PHASE_PACK
        OV_CUDA(job->tperm.alloc(sizeof(int32_t) * (size_t)n_t, s));
        OV_CUDA(job->tbox.alloc(sizeof(Box) * (size_t)n_t, s));
        OV_CUDA(job->tpoly.alloc(sizeof(double) * 2 * (size_t)k_t * (size_t)n_t, s));
        OV_CUDA(job->tnv.alloc((size_t)n_t, s));
        PackedMesh pm;
        pm.perm = job->tperm.as<int32_t>(); pm.box = job->tbox.as<Box>();
        pm.poly = job->tpoly.as<double>(); pm.nvert = job->tnv.as<uint8_t>();
        rc = pack_mesh(d_t_xy, d_t_faces, idx_bytes, t_begin, t_end, k_t, source->d_frame, nullptr, &pm, s);
        if (rc) goto fail;

        TraverseParams P;
        P.boxes[0] = source->leaf_box; P.count[0] = source->n_face;
        for (int l = 0; l < source->n_levels; ++l) { P.boxes[l + 1] = source->level_box[l]; P.count[l + 1] = source->level_n[l]; }
        P.top = source->n_levels + 1;
        const unsigned tiles = (unsigned)div_up(n_t, 32);
        const unsigned tblocks = (unsigned)div_up(tiles, TRAV_WARPS);

        // count pass
        OV_CUDA(job->mark());  // XG_PHASE_COUNT
        DevBuf cnt;
        OV_CUDA(cnt.alloc(sizeof(int32_t) * (size_t)(n_t + 1), s));
        OV_CUDA(cudaMemsetAsync(cnt.as<int32_t>() + n_t, 0, sizeof(int32_t), s));
        k_traverse<false><<<tblocks, TRAV_WARPS * 32, 0, s>>>(P, job->tbox.as<Box>(), n_t, cnt.as<int32_t>(), nullptr,
                                                             nullptr, nullptr, job->stats.as<unsigned long long>());
        OV_CUDA(cudaGetLastError());
        OV_CUDA(job->mark());  // XG_PHASE_SCAN
        OV_CUDA(job->cand_ptr.alloc(sizeof(int64_t) * (size_t)(n_t + 1), s));
        {
            cub::TransformInputIterator<int64_t, CastI64, const int32_t *> it(cnt.as<int32_t>(), CastI64());
            size_t tb = 0;
            OV_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tb, it, job->cand_ptr.as<int64_t>(), n_t + 1, s));
            DevBuf tmp;
            OV_CUDA(tmp.alloc(tb, s));
            OV_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, tb, it, job->cand_ptr.as<int64_t>(), n_t + 1, s));
        }
        int64_t n_cand = 0;
        OV_CUDA(cudaMemcpyAsync(&n_cand, job->cand_ptr.as<int64_t>() + n_t, sizeof(int64_t), cudaMemcpyDeviceToHost, s));
        OV_CUDA(cudaStreamSynchronize(s));
        job->n_cand = n_cand;

        // fill pass
        OV_CUDA(job->mark());  // XG_PHASE_FILL
        OV_CUDA(job->cand_src.alloc(sizeof(int32_t) * (size_t)n_cand, s));
        OV_CUDA(job->cand_tgt.alloc(sizeof(int32_t) * (size_t)n_cand, s));
        OV_CUDA(job->area.alloc(sizeof(double) * (size_t)n_cand, s));
        k_traverse<true><<<tblocks, TRAV_WARPS * 32, 0, s>>>(P, job->tbox.as<Box>(), n_t, nullptr, job->cand_ptr.as<int64_t>(),
                                                            job->cand_src.as<int32_t>(), job->cand_tgt.as<int32_t>(), nullptr);
        OV_CUDA(cudaGetLastError());

        // clip
        OV_CUDA(job->mark());  // XG_PHASE_CLIP
        if (n_cand > 0) {
            const int ks = source->k;
            const int32_t *ct = job->cand_tgt.as<int32_t>(), *cs = job->cand_src.as<int32_t>();
            const double *tp = job->tpoly.as<double>();
            const uint8_t *tnv = job->tnv.as<uint8_t>();
            unsigned long long *st = job->stats.as<unsigned long long>();
            double *ar = job->area.as<double>();
            DevBuf surv;  // compact list of pairs that pass the separating-axis filter
            if (k_t <= 8 && ks <= 8) OV_CUDA(surv.alloc(sizeof(Survivor) * (size_t)n_cand, s));
            if (k_t == 3 && ks == 3) rc = launch_clip_dense<3, 3>(ct, cs, n_cand, tp, tnv, k_t, source->poly, source->nvert, ks, ar, surv.p, st, s);
            else if (k_t == 3 && ks == 4) rc = launch_clip_dense<3, 4>(ct, cs, n_cand, tp, tnv, k_t, source->poly, source->nvert, ks, ar, surv.p, st, s);
            else if (k_t == 4 && ks == 3) rc = launch_clip_dense<4, 3>(ct, cs, n_cand, tp, tnv, k_t, source->poly, source->nvert, ks, ar, surv.p, st, s);
            else if (k_t == 4 && ks == 4) rc = launch_clip_dense<4, 4>(ct, cs, n_cand, tp, tnv, k_t, source->poly, source->nvert, ks, ar, surv.p, st, s);
            else if (k_t <= 8 && ks <= 8) rc = launch_clip_dense<8, 8>(ct, cs, n_cand, tp, tnv, k_t, source->poly, source->nvert, ks, ar, surv.p, st, s);
            else rc = launch_clip<XG_MAX_N_VERTEX, XG_MAX_N_VERTEX>(32, ct, cs, n_cand, tp, tnv, k_t, source->poly, source->nvert, ks, ar, st, s);
            if (rc) goto fail;
        }

        // CSR row pointer in the caller's row order (cnt is reused as row counts)
        OV_CUDA(job->mark());  // XG_PHASE_ROWPTR
        k_row_count<<<(unsigned)div_up(n_t, 256), 256, 0, s>>>(job->cand_ptr.as<int64_t>(), job->area.as<double>(),
                                                              job->tperm.as<int32_t>(), t_begin, n_t, cnt.as<int32_t>());
        OV_CUDA(cudaGetLastError());
        {
            cub::TransformInputIterator<int64_t, CastI64, const int32_t *> it(cnt.as<int32_t>(), CastI64());
            size_t tb = 0;
            OV_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tb, it, d_indptr, n_t + 1, s));
            DevBuf tmp;
            OV_CUDA(tmp.alloc(tb, s));
            OV_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, tb, it, d_indptr, n_t + 1, s));
        }
        int64_t nnz = 0;
        unsigned long long hs[8];
        OV_CUDA(cudaMemcpyAsync(&nnz, d_indptr + n_t, sizeof(int64_t), cudaMemcpyDeviceToHost, s));
        OV_CUDA(cudaMemcpyAsync(hs, job->stats.p, sizeof(hs), cudaMemcpyDeviceToHost, s));
        OV_CUDA(job->mark());  // end
        OV_CUDA(cudaStreamSynchronize(s));
        job->collect();
        job->nnz = nnz;
        job->h_stats[0] = n_t;
        job->h_stats[1] = (int64_t)hs[0] * 2;  // the walk runs twice (count + fill)
        job->h_stats[2] = (int64_t)hs[1] * 2;
        job->h_stats[3] = (int64_t)hs[3];
        job->h_stats[4] = nnz;
        job->h_stats[5] = n_cand;
        job->stats_valid = true;
        *h_nnz = nnz;
        *job_out = job;
        return XG_OK;
    }
fail:
    delete job;
    return rc;
#undef OV_CUDA
}

extern "C" int xg_overlap_finish(xg_overlap *job, const int64_t *d_indptr, int32_t *d_indices,
                                 double *d_area, int relative, void *stream) {
    if (!job) { set_error("xg_overlap_finish: NULL job"); return XG_ERR_INVALID; }
    cudaStream_t s = (cudaStream_t)stream;
    int rc = XG_OK;
    if (job->nnz > 0) {
        if (!d_indices || !d_area || !d_indptr) { set_error("xg_overlap_finish: NULL output"); delete job; return XG_ERR_INVALID; }
        const xg_index *src = job->src;
        k_emit<<<(unsigned)div_up(job->n_t, 256), 256, 0, s>>>(
            job->cand_ptr.as<int64_t>(), job->cand_src.as<int32_t>(), job->area.as<double>(),
            job->tperm.as<int32_t>(), job->t_begin, job->n_t, src->perm, (const double2 *)src->poly,
            src->nvert, src->k, relative, d_indptr, d_indices, d_area);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) rc = cuda_fail(e, "k_emit", __FILE__, __LINE__);
    }
    // buffers are released stream-ordered on the stream that produced them
    if (s != job->s) cudaStreamSynchronize(s);
    delete job;
    return rc;
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
