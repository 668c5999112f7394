// xg_locate.cu -- point location and barycentric weights.
//
// Replaces CellTree2d.locate_points / compute_barycentric_weights as called from
// /root/reference/xugrid/regrid/unstructured.py:138-139,
// /root/reference/xugrid/ugrid/ugridbase.py:1305-1323 and
// /root/reference/xugrid/ugrid/ugrid2d.py:1054-1078,1099.
//
// Points are Morton-sorted (same frame as the index) so that a warp owns 32
// neighbouring points and walks the wide BVH once for all of them (lane =
// point), exactly like the face tiles of xg_overlap.cu.  A face qualifies when
// the point is inside by ray casting or within `tolerance` of one of its
// edges; boxes are therefore tested with a margin of 1.5 * tolerance (a point
// that passes the on-edge test lies within sqrt(2) * tolerance of the face).
// Among qualifying faces the smallest caller face id wins, which makes the
// answer independent of the tree.
#include <cub/device/device_radix_sort.cuh>

#include "xg_common.cuh"
#include "xg_geom.cuh"

namespace xg {

constexpr int LOC_WARPS = 8;
constexpr int LOC_STACK = 256;

struct LocateParams {
    const Box *boxes[MAX_LEVELS + 1];
    int64_t count[MAX_LEVELS + 1];
    int top;
};

__global__ void __launch_bounds__(256)
k_point_keys(const double2 *__restrict__ pts, int64_t n, const double *__restrict__ frame,
             uint32_t *__restrict__ keys, int32_t *__restrict__ ids) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double2 p = pts[i];
    const double fx0 = frame[0], fx1 = frame[1], fy0 = frame[2], fy1 = frame[3];
    const double sx = fx1 > fx0 ? 65536.0 / (fx1 - fx0) : 0.0;
    const double sy = fy1 > fy0 ? 65536.0 / (fy1 - fy0) : 0.0;
    double qx = (p.x - fx0) * sx, qy = (p.y - fy0) * sy;
    // NaN-safe clamp: fmax(NaN, 0) = 0
    qx = fmin(fmax(qx, 0.0), 65535.0);
    qy = fmin(fmax(qy, 0.0), 65535.0);
    keys[i] = curve_key16((uint32_t)qx, (uint32_t)qy);
    ids[i] = (int32_t)i;
}

__device__ __forceinline__ bool point_in_box(double px, double py, const Box &b, double m) {
    return px >= b.xmin - m && px <= b.xmax + m && py >= b.ymin - m && py <= b.ymax + m;
}

// lane = point (Morton order).  Writes the caller's face id (or -1) at the
// point's original position, and the face's sorted position in pos_out.
// ANY: the caller only asks WHETHER a face holds the point (the outside mask of the barycentric
// build): a lane stops at its first qualifying face, subtrees that only found lanes reach are
// not descended, and the warp leaves the walk as soon as all its points are placed.
template <bool ANY>
__global__ void __launch_bounds__(LOC_WARPS * 32)
k_locate(const LocateParams P, const double2 *__restrict__ pts, const int32_t *__restrict__ order,
         int64_t n, const double2 *__restrict__ poly, const uint8_t *__restrict__ nvert, int k,
         const int32_t *__restrict__ perm, double tol, int32_t *__restrict__ face_out,
         int32_t *__restrict__ pos_out) {
    __shared__ Box s_node[LOC_WARPS][32];
    __shared__ int32_t s_stack[LOC_WARPS][LOC_STACK];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t tile = blockIdx.x * (int64_t)LOC_WARPS + w;
    const int64_t i = tile * 32 + lane;
    if (tile * 32 >= n) return;
    const bool valid = i < n;
    const int32_t orig = valid ? order[i] : -1;
    double px = NAN, py = NAN;
    if (valid) { const double2 p = pts[orig]; px = p.x; py = p.y; }
    const double margin = 1.5 * tol;
    const double tol2 = tol * tol;
    // union box of the tile's points (NaN points never match anything)
    Box ub;
    ub.xmin = valid && px == px ? px : INFINITY; ub.xmax = valid && px == px ? px : -INFINITY;
    ub.ymin = valid && py == py ? py : INFINITY; ub.ymax = valid && py == py ? py : -INFINITY;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        ub.xmin = fmin(ub.xmin, __shfl_xor_sync(0xffffffffu, ub.xmin, o));
        ub.xmax = fmax(ub.xmax, __shfl_xor_sync(0xffffffffu, ub.xmax, o));
        ub.ymin = fmin(ub.ymin, __shfl_xor_sync(0xffffffffu, ub.ymin, o));
        ub.ymax = fmax(ub.ymax, __shfl_xor_sync(0xffffffffu, ub.ymax, o));
    }
    int32_t best = 0x7fffffff, best_pos = -1;
    int sp = 1;
    if (lane == 0) s_stack[w][0] = P.top << 27;
    __syncwarp();
    const bool searchable = valid && px == px && py == py;
    while (sp > 0) {
        if (ANY && __all_sync(0xffffffffu, best_pos >= 0 || !searchable)) break;
        const int32_t entry = s_stack[w][--sp];
        const int d = entry >> 27;
        const int64_t c0 = (int64_t)(entry & 0x7ffffff) * FANOUT;
        const Box *cb = P.boxes[d - 1];
        const int64_t nc = P.count[d - 1] - c0;
        Box mine;
        if (lane < nc) mine = cb[c0 + lane];
        else { mine.xmin = INFINITY; mine.xmax = -INFINITY; mine.ymin = INFINITY; mine.ymax = -INFINITY; }
        const bool pre = mine.xmin - margin <= ub.xmax && ub.xmin <= mine.xmax + margin &&
                         mine.ymin - margin <= ub.ymax && ub.ymin <= mine.ymax + margin;
        unsigned m = __ballot_sync(0xffffffffu, pre);
        __syncwarp();
        if (m == 0) continue;
        s_node[w][lane] = mine;
        __syncwarp();
        const bool looking = !(ANY && best_pos >= 0);  // (a placed point asks for nothing more)
        if (d > 1) {
            while (m) {
                const int c = __ffs(m) - 1;
                m &= m - 1;
                const Box b = s_node[w][c];
                if (__any_sync(0xffffffffu, looking && point_in_box(px, py, b, margin))) {
                    if (lane == 0) s_stack[w][sp] = ((d - 1) << 27) | (int32_t)(c0 + c);
                    ++sp;
                }
            }
        } else {
            // every lane first collects the faces whose box holds ITS point as a bit mask, then
            // works through its own one or two hits: the polygon test runs a couple of times per
            // leaf with most lanes busy, instead of once per child with the one to three lanes
            // whose point happens to be in that child's box (ncu launch list of the
            // BarycentricInterpolator build, 16.7 M points: 4.5 + 4.9 ms for the two walks)
            unsigned hits = 0u;
            while (m) {
                const int c = __ffs(m) - 1;
                m &= m - 1;
                const Box b = s_node[w][c];
                hits |= ((looking && point_in_box(px, py, b, margin)) ? 1u : 0u) << c;
            }
            while (__any_sync(0xffffffffu, hits != 0u)) {
                if (hits != 0u) {
                    const int c = __ffs(hits) - 1;
                    hits &= hits - 1;
                    const int64_t pos = c0 + c;
                    const int32_t fid = perm[pos];
                    if (fid < best &&
                        point_in_polygon_or_on_edge(px, py, poly + pos * k, nvert[pos] & 0x7f, tol2)) {
                        best = fid;
                        best_pos = (int32_t)pos;
                        if (ANY) hits = 0u;
                    }
                }
            }
        }
        __syncwarp();
    }
    if (valid) {
        face_out[orig] = best_pos >= 0 ? best : -1;
        if (pos_out) pos_out[orig] = best_pos;
    }
}

// one thread per point: weights for the vertices of the located face, written
// in the caller's vertex order (a face that was flipped to CCW is un-flipped)
__global__ void __launch_bounds__(128)
k_barycentric(const double2 *__restrict__ pts, int64_t n, const int32_t *__restrict__ pos_in,
              const double2 *__restrict__ poly, const uint8_t *__restrict__ nvert, int k, double tol,
              double *__restrict__ weights) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    double *wout = weights + i * k;
    for (int j = 0; j < k; ++j) wout[j] = 0.0;
    const int32_t pos = pos_in[i];
    if (pos < 0) return;
    const double2 p = pts[i];
    const double2 *pg = poly + (int64_t)pos * k;
    const uint8_t nvf = nvert[pos];
    const int len = nvf & 0x7f;
    const bool flip = nvf & 0x80;
    double w[XG_MAX_N_VERTEX];
    if (k == 3) bary_triangle(pg, p.x, p.y, w);
    else bary_wachspress(pg, len, p.x, p.y, tol * tol, w);
    for (int j = 0; j < len; ++j) wout[flip ? len - 1 - j : j] = w[j];
}

// ---- intersect_edges: line segments against the faces --------------------------------
// Replaces CellTree2d.intersect_edges as called from
// /root/reference/xugrid/regrid/unstructured.py:203-215 (NetworkGridder weights = length
// of every network edge inside every face).  Same walk as k_locate with lane = segment
// (Morton order of the midpoints, closed box tests) and Cyrus-Beck parametric clipping
// against the packed CCW polygon; arithmetic and tie rules as oracle/xg_oracle.c::cyrus_beck.
__global__ void __launch_bounds__(256)
k_edge_keys(const double4 *__restrict__ edges, int64_t n, const double *__restrict__ frame,
            uint32_t *__restrict__ keys, int32_t *__restrict__ ids) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double4 e = edges[i];
    const double fx0 = frame[0], fx1 = frame[1], fy0 = frame[2], fy1 = frame[3];
    const double sx = fx1 > fx0 ? 65536.0 / (fx1 - fx0) : 0.0;
    const double sy = fy1 > fy0 ? 65536.0 / (fy1 - fy0) : 0.0;
    double qx = (0.5 * (e.x + e.z) - fx0) * sx, qy = (0.5 * (e.y + e.w) - fy0) * sy;
    qx = fmin(fmax(qx, 0.0), 65535.0);
    qy = fmin(fmax(qy, 0.0), 65535.0);
    keys[i] = curve_key16((uint32_t)qx, (uint32_t)qy);
    ids[i] = (int32_t)i;
}

__device__ __forceinline__ bool boxes_touch(const Box &a, const Box &b) {
    return (a.xmin <= b.xmax) & (b.xmin <= a.xmax) & (a.ymin <= b.ymax) & (b.ymin <= a.ymax);
}

// segment a + t (b - a), t in [0, 1], clipped by the half planes of a CCW polygon
__device__ __forceinline__ bool cyrus_beck(double ax, double ay, double bx, double by,
                                           const double2 *__restrict__ poly, int n, double4 *out) {
    const double sx = bx - ax, sy = by - ay;
    double t0 = 0.0, t1 = 1.0;
    double2 q0 = poly[n - 1];
    for (int i = 0; i < n; ++i) {
        const double2 q1 = poly[i];
        const double ex = q1.x - q0.x, ey = q1.y - q0.y;
        if (!(ex == 0.0 && ey == 0.0)) {
            const double nx = -ey, ny = ex;  // inward normal
            const double num = nx * (ax - q0.x) + ny * (ay - q0.y);
            const double den = nx * sx + ny * sy;
            if (den == 0.0) {
                if (num < 0.0) return false;                                  // parallel, outside
                if (num == 0.0 && !(ex * sx + ey * sy > 0.0)) return false;  // on the edge, face on the right
            } else {
                const double t = -num / den;
                if (den > 0.0) t0 = t > t0 ? t : t0;  // entering
                else t1 = t < t1 ? t : t1;            // leaving
            }
        }
        q0 = q1;
    }
    if (!(t0 < t1)) return false;
    *out = make_double4(ax + t0 * sx, ay + t0 * sy, ax + t1 * sx, ay + t1 * sy);
    return true;
}

__global__ void __launch_bounds__(LOC_WARPS * 32)
k_intersect_edges(const LocateParams P, const double4 *__restrict__ edges, const int32_t *__restrict__ order,
                  int64_t n, const double2 *__restrict__ poly, const uint8_t *__restrict__ nvert, int k,
                  const int32_t *__restrict__ perm, unsigned long long cap,
                  unsigned long long *__restrict__ counter, int32_t *__restrict__ edge_out,
                  int32_t *__restrict__ face_out, double4 *__restrict__ xy_out) {
    __shared__ Box s_node[LOC_WARPS][32];
    __shared__ int32_t s_stack[LOC_WARPS][LOC_STACK];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t tile = blockIdx.x * (int64_t)LOC_WARPS + w;
    const int64_t i = tile * 32 + lane;
    if (tile * 32 >= n) return;
    const bool valid = i < n;
    const int32_t orig = valid ? order[i] : -1;
    double4 e = make_double4(NAN, NAN, NAN, NAN);
    if (valid) e = edges[orig];
    Box qb;  // NaN coordinates give an empty box: never matches
    qb.xmin = fmin(e.x, e.z); qb.xmax = fmax(e.x, e.z);
    qb.ymin = fmin(e.y, e.w); qb.ymax = fmax(e.y, e.w);
    const bool ok = valid && (e.x == e.x) && (e.y == e.y) && (e.z == e.z) && (e.w == e.w);
    if (!ok) { qb.xmin = INFINITY; qb.xmax = -INFINITY; qb.ymin = INFINITY; qb.ymax = -INFINITY; }
    Box ub = qb;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        ub.xmin = fmin(ub.xmin, __shfl_xor_sync(0xffffffffu, ub.xmin, o));
        ub.xmax = fmax(ub.xmax, __shfl_xor_sync(0xffffffffu, ub.xmax, o));
        ub.ymin = fmin(ub.ymin, __shfl_xor_sync(0xffffffffu, ub.ymin, o));
        ub.ymax = fmax(ub.ymax, __shfl_xor_sync(0xffffffffu, ub.ymax, o));
    }
    int sp = 1;
    if (lane == 0) s_stack[w][0] = P.top << 27;
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
        unsigned m = __ballot_sync(0xffffffffu, boxes_touch(mine, ub));
        __syncwarp();
        if (m == 0) continue;
        s_node[w][lane] = mine;
        __syncwarp();
        if (d > 1) {
            while (m) {
                const int c = __ffs(m) - 1;
                m &= m - 1;
                const Box b = s_node[w][c];
                if (__any_sync(0xffffffffu, boxes_touch(qb, b))) {
                    if (lane == 0) s_stack[w][sp] = ((d - 1) << 27) | (int32_t)(c0 + c);
                    ++sp;
                }
            }
        } else {
            while (m) {
                const int c = __ffs(m) - 1;
                m &= m - 1;
                const Box b = s_node[w][c];
                const int64_t pos = c0 + c;
                double4 seg;
                bool hit = false;
                if (boxes_touch(qb, b)) hit = cyrus_beck(e.x, e.y, e.z, e.w, poly + pos * k, nvert[pos] & 0x7f, &seg);
                // one atomic per warp and child, not one per hit (the loop over m is warp-uniform)
                const unsigned hm = __ballot_sync(0xffffffffu, hit);
                if (hm) {
                    const int leader = __ffs(hm) - 1;
                    unsigned long long base = 0;
                    if (lane == leader) base = atomicAdd(counter, (unsigned long long)__popc(hm));
                    base = __shfl_sync(0xffffffffu, base, leader);
                    const unsigned long long slot = base + __popc(hm & ((1u << lane) - 1u));
                    if (hit && slot < cap) {
                        edge_out[slot] = orig;
                        face_out[slot] = perm[pos];
                        xy_out[slot] = seg;
                    }
                }
            }
        }
        __syncwarp();
    }
}

static int locate_impl(const xg_index *ix, const double *d_points, int64_t n, double tol,
                       int32_t *d_face, int32_t *d_pos, cudaStream_t s, bool any_face = false) {
    if (n == 0) return XG_OK;
    if (ix->n_face == 0) {
        XG_CUDA(cudaMemsetAsync(d_face, 0xff, sizeof(int32_t) * (size_t)n, s));
        if (d_pos) XG_CUDA(cudaMemsetAsync(d_pos, 0xff, sizeof(int32_t) * (size_t)n, s));
        return XG_OK;
    }
    DevBuf keys_in, keys_out, ids_in, order, tmp;
    XG_CUDA(keys_in.alloc(sizeof(uint32_t) * (size_t)n, s));
    XG_CUDA(keys_out.alloc(sizeof(uint32_t) * (size_t)n, s));
    XG_CUDA(ids_in.alloc(sizeof(int32_t) * (size_t)n, s));
    XG_CUDA(order.alloc(sizeof(int32_t) * (size_t)n, s));
    k_point_keys<<<(unsigned)div_up(n, 256), 256, 0, s>>>((const double2 *)d_points, n, ix->d_frame,
                                                         keys_in.as<uint32_t>(), ids_in.as<int32_t>());
    XG_CHECK_LAUNCH();
    size_t tb = 0;
    XG_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tb, keys_in.as<uint32_t>(), keys_out.as<uint32_t>(),
                                            ids_in.as<int32_t>(), order.as<int32_t>(), n, 0, 32, s));
    XG_CUDA(tmp.alloc(tb, s));
    XG_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, tb, keys_in.as<uint32_t>(), keys_out.as<uint32_t>(),
                                            ids_in.as<int32_t>(), order.as<int32_t>(), n, 0, 32, s));
    LocateParams P;
    P.boxes[0] = ix->leaf_box; P.count[0] = ix->n_face;
    for (int l = 0; l < ix->n_levels; ++l) { P.boxes[l + 1] = ix->level_box[l]; P.count[l + 1] = ix->level_n[l]; }
    P.top = ix->n_levels + 1;
    const unsigned tiles = (unsigned)div_up(n, 32);
    if (any_face)
        k_locate<true><<<(unsigned)div_up(tiles, LOC_WARPS), LOC_WARPS * 32, 0, s>>>(
            P, (const double2 *)d_points, order.as<int32_t>(), n, (const double2 *)ix->poly, ix->nvert,
            ix->k, ix->perm, tol, d_face, d_pos);
    else
        k_locate<false><<<(unsigned)div_up(tiles, LOC_WARPS), LOC_WARPS * 32, 0, s>>>(
            P, (const double2 *)d_points, order.as<int32_t>(), n, (const double2 *)ix->poly, ix->nvert,
            ix->k, ix->perm, tol, d_face, d_pos);
    XG_CHECK_LAUNCH();
    return XG_OK;
}

}  // namespace xg

using namespace xg;

extern "C" int xg_locate_points(const xg_index *index, const double *d_points, int64_t n_point,
                                double tolerance, int32_t *d_face_index, void *stream) {
    if (!index || n_point < 0 || (n_point > 0 && (!d_points || !d_face_index))) {
        set_error("xg_locate_points: invalid argument");
        return XG_ERR_INVALID;
    }
    if (n_point >= (int64_t)1 << 31) { set_error("xg_locate_points: more than 2^31 points"); return XG_ERR_UNSUPPORTED; }
    if (tolerance < 0) { const int rc = index_host_values(index); if (rc) return rc; }
    const double tol = tolerance < 0 ? index->default_tolerance : tolerance;
    index_note_use(index, (cudaStream_t)stream);
    return locate_impl(index, d_points, n_point, tol, d_face_index, nullptr, (cudaStream_t)stream);
}

extern "C" int xg_locate_any(const xg_index *index, const double *d_points, int64_t n_point,
                             double tolerance, int32_t *d_face_index, void *stream) {
    if (!index || n_point < 0 || (n_point > 0 && (!d_points || !d_face_index))) {
        set_error("xg_locate_any: invalid argument");
        return XG_ERR_INVALID;
    }
    if (n_point >= (int64_t)1 << 31) { set_error("xg_locate_any: more than 2^31 points"); return XG_ERR_UNSUPPORTED; }
    if (tolerance < 0) { const int rc = index_host_values(index); if (rc) return rc; }
    const double tol = tolerance < 0 ? index->default_tolerance : tolerance;
    index_note_use(index, (cudaStream_t)stream);
    return locate_impl(index, d_points, n_point, tol, d_face_index, nullptr, (cudaStream_t)stream, true);
}

extern "C" int xg_barycentric(const xg_index *index, const double *d_points, int64_t n_point,
                              double tolerance, int32_t *d_face_index, double *d_weights,
                              void *stream) {
    if (!index || n_point < 0 || (n_point > 0 && (!d_points || !d_face_index || !d_weights))) {
        set_error("xg_barycentric: invalid argument");
        return XG_ERR_INVALID;
    }
    if (n_point >= (int64_t)1 << 31) { set_error("xg_barycentric: more than 2^31 points"); return XG_ERR_UNSUPPORTED; }
    if (n_point == 0) return XG_OK;
    cudaStream_t s = (cudaStream_t)stream;
    if (tolerance < 0) { const int rc = index_host_values(index); if (rc) return rc; }
    const double tol = tolerance < 0 ? index->default_tolerance : tolerance;
    index_note_use(index, s);
    DevBuf pos;
    XG_CUDA(pos.alloc(sizeof(int32_t) * (size_t)n_point, s));
    int rc = locate_impl(index, d_points, n_point, tol, d_face_index, pos.as<int32_t>(), s);
    if (rc) return rc;
    k_barycentric<<<(unsigned)div_up(n_point, 128), 128, 0, s>>>(
        (const double2 *)d_points, n_point, pos.as<int32_t>(), (const double2 *)index->poly,
        index->nvert, index->k, tol, d_weights);
    XG_CHECK_LAUNCH();
    return XG_OK;
}

extern "C" int xg_intersect_edges(const xg_index *index, const double *d_edges, int64_t n_edge,
                                  int64_t cap, int32_t *d_edge_index, int32_t *d_face_index,
                                  double *d_xy, int64_t *h_count, void *stream) {
    if (!index || !h_count || n_edge < 0 || cap < 0 || (n_edge > 0 && !d_edges) ||
        (cap > 0 && (!d_edge_index || !d_face_index || !d_xy))) {
        set_error("xg_intersect_edges: invalid argument");
        return XG_ERR_INVALID;
    }
    if (n_edge >= (int64_t)1 << 31) { set_error("xg_intersect_edges: more than 2^31 edges"); return XG_ERR_UNSUPPORTED; }
    *h_count = 0;
    if (n_edge == 0 || index->n_face == 0) return XG_OK;
    cudaStream_t s = (cudaStream_t)stream;
    index_note_use(index, s);
    const int64_t n = n_edge;
    DevBuf keys_in, keys_out, ids_in, order, tmp, counter;
    XG_CUDA(keys_in.alloc(sizeof(uint32_t) * (size_t)n, s));
    XG_CUDA(keys_out.alloc(sizeof(uint32_t) * (size_t)n, s));
    XG_CUDA(ids_in.alloc(sizeof(int32_t) * (size_t)n, s));
    XG_CUDA(order.alloc(sizeof(int32_t) * (size_t)n, s));
    XG_CUDA(counter.alloc(sizeof(unsigned long long), s));
    XG_CUDA(cudaMemsetAsync(counter.p, 0, sizeof(unsigned long long), s));
    k_edge_keys<<<(unsigned)div_up(n, 256), 256, 0, s>>>((const double4 *)d_edges, n, index->d_frame,
                                                        keys_in.as<uint32_t>(), ids_in.as<int32_t>());
    XG_CHECK_LAUNCH();
    size_t tb = 0;
    XG_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tb, keys_in.as<uint32_t>(), keys_out.as<uint32_t>(),
                                            ids_in.as<int32_t>(), order.as<int32_t>(), n, 0, 32, s));
    XG_CUDA(tmp.alloc(tb, s));
    XG_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, tb, keys_in.as<uint32_t>(), keys_out.as<uint32_t>(),
                                            ids_in.as<int32_t>(), order.as<int32_t>(), n, 0, 32, s));
    LocateParams P;
    P.boxes[0] = index->leaf_box; P.count[0] = index->n_face;
    for (int l = 0; l < index->n_levels; ++l) { P.boxes[l + 1] = index->level_box[l]; P.count[l + 1] = index->level_n[l]; }
    P.top = index->n_levels + 1;
    const unsigned tiles = (unsigned)div_up(n, 32);
    k_intersect_edges<<<(unsigned)div_up(tiles, LOC_WARPS), LOC_WARPS * 32, 0, s>>>(
        P, (const double4 *)d_edges, order.as<int32_t>(), n, (const double2 *)index->poly, index->nvert,
        index->k, index->perm, (unsigned long long)cap, counter.as<unsigned long long>(), d_edge_index,
        d_face_index, (double4 *)d_xy);
    XG_CHECK_LAUNCH();
    unsigned long long h = 0;
    XG_CUDA(cudaMemcpyAsync(&h, counter.p, sizeof(h), cudaMemcpyDeviceToHost, s));
    XG_CUDA(cudaStreamSynchronize(s));
    *h_count = (int64_t)h;
    return XG_OK;
}
