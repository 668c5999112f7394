// xg_voronoi.cu -- the preparation and the assembly around the barycentric kernel.
//
// BarycentricInterpolator (/root/reference/xugrid/regrid/unstructured.py:146-201)
// interpolates inside the centroidal Voronoi polygons around the source NODES
// (/root/reference/xugrid/ugrid/voronoi.py:330-458, voronoi_topology(add_exterior=True,
// add_vertices=True, skip_concave=True), built there with numpy / pandas / scipy.sparse).
// This file builds the same topology with dedicated kernels -- the formulation is the one
// of xugrid_b200/voronoi.py (torch tensor ops, kept as the CPU-testable statement of it;
// the two are compared element by element in tests/test_gpu_voronoi.py):
//
//   k_vor_degree / k_vor_fill / k_vor_sort_adj   node -> face adjacency (CSR), every list
//                      in ascending face number (atomics only decide the slot, the sort
//                      makes the result deterministic)
//   k_vor_exterior     an edge is exterior when no other (face, slot) has the same end
//                      points: searched in the adjacency list of one end point (no global
//                      sort of 3 x n_face edge keys); marks the two end nodes, remembers the
//                      edge as the OUT edge of its start node and the IN edge of its end
//                      node, and decides whether the projection of the face centroid on the
//                      edge becomes a vertex
//   k_vor_classify     interior / exterior flag per node, polygon sizes, validity
//   k_vor_polygons     one thread per node: members sorted by angle (around the node for an
//                      interior polygon, around the mean of the members for an exterior one),
//                      corner choice by the two polygon areas, rows / vertices written
//
// and, behind the barycentric kernel (xg_locate.cu):
//
//   k_bary_adjust_count / k_bary_fill   corner weights redistributed to the two projection
//                      vertices (unstructured.py:17-57), points outside the source mesh
//                      zeroed, positive weights counted and written as CSR rows
//                      (unstructured.py:176-201) -- no (n_point, k) boolean-mask passes.
#include <cub/device/device_scan.cuh>
#include <cub/iterator/transform_input_iterator.cuh>

#include "xg_common.cuh"

namespace xg {

constexpr double VOR_X_EPSILON = 1.0e-8;  // /root/reference/xugrid/constants.py:45

enum { VC_ERR = 0, VC_KMAX = 1, VC_N = 4 };

// vertices of a face row (fill values -1 behind them)
template <typename I>
__device__ __forceinline__ int face_len(const I *face, int k) {
    int n = 0;
#pragma unroll 1
    while (n < k && face[n] >= 0) ++n;
    return n;
}

template <typename I>
__global__ void __launch_bounds__(256)
k_vor_degree(const I *__restrict__ faces, int64_t n_face, int k, int32_t *__restrict__ deg) {
    const int64_t f = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (f >= n_face) return;
    for (int j = 0; j < k; ++j) {
        const I v = faces[f * k + j];
        if (v >= 0) atomicAdd(deg + v, 1);
    }
}

template <typename I>
__global__ void __launch_bounds__(256)
k_vor_fill(const I *__restrict__ faces, int64_t n_face, int k, const int32_t *__restrict__ ptr,
           int32_t *__restrict__ cursor, int32_t *__restrict__ adj) {
    const int64_t f = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (f >= n_face) return;
    for (int j = 0; j < k; ++j) {
        const I v = faces[f * k + j];
        if (v >= 0) adj[ptr[v] + atomicAdd(cursor + v, 1)] = (int32_t)f;
    }
}

__global__ void __launch_bounds__(256)
k_vor_sort_adj(const int32_t *__restrict__ ptr, int64_t n_node, int32_t *__restrict__ adj) {
    const int64_t n = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (n >= n_node) return;
    const int b = ptr[n], e = ptr[n + 1];
    for (int i = b + 1; i < e; ++i) {
        const int32_t v = adj[i];
        int j = i - 1;
        while (j >= b && adj[j] > v) {
            adj[j + 1] = adj[j];
            --j;
        }
        adj[j + 1] = v;
    }
}

// projection of the centroid of face f on its edge (a, b); voronoi.py (exterior part)
__device__ __forceinline__ double2 project_on_edge(const double2 *__restrict__ xy, const double2 c, int64_t a,
                                                   int64_t b) {
    const int64_t lo = a < b ? a : b, hi = a < b ? b : a;
    const double2 A = xy[lo], B = xy[hi];
    const double Vx = B.x - A.x, Vy = B.y - A.y;
    const double Ux = c.x - A.x, Uy = c.y - A.y;
    const double t = (Uy * Vy + Ux * Vx) / (Vy * Vy + Vx * Vx);
    return make_double2(A.x + t * Vx, A.y + t * Vy);
}

template <typename I>
__device__ __forceinline__ void edge_of_code(const I *__restrict__ faces, int k, int32_t code, int32_t &f,
                                             int64_t &a, int64_t &b) {
    f = code / k;
    const int j = code - f * k;
    const I *face = faces + (int64_t)f * k;
    const int len = face_len(face, k);
    a = face[j];
    b = face[j + 1 < len ? j + 1 : 0];
}

// io_cnt[n]: exterior edges that start at n (low 16 bits) / end at n (high 16 bits)
template <typename I>
__global__ void __launch_bounds__(256)
k_vor_exterior(const I *__restrict__ faces, int64_t n_face, int k, const double2 *__restrict__ xy,
               const double2 *__restrict__ cen, const int32_t *__restrict__ ptr,
               const int32_t *__restrict__ adj, int32_t *__restrict__ io_cnt, int32_t *__restrict__ out_code,
               int32_t *__restrict__ in_code, uint8_t *__restrict__ keep_flag) {
    const int64_t code = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (code >= n_face * k) return;
    const int64_t f = code / k;
    const int j = (int)(code - f * k);
    const I *face = faces + f * k;
    const int len = face_len(face, k);
    // (one exit, no `return` from inside the loops: the version that left the nested loops
    //  early read rows at wild addresses on some lanes of a mixed tri/quad mesh -- valid
    //  adjacency, compute-sanitizer pointing at the row load -- and stopped doing so as soon
    //  as a printf was added to the loop body; nvcc 12.9, not understood further)
    I a = 0, b = 0;
    bool exterior = false;
    if (j < len) {
        a = face[j];
        b = face[j + 1 < len ? j + 1 : 0];
        bool shared = a == b;  // (a repeated vertex is no edge at all)
        const int q_begin = ptr[a], q_end = ptr[a + 1];
        for (int q = q_begin; q < q_end; ++q) {
            const int32_t g = adj[q];
            const I *gf = faces + (int64_t)g * (int64_t)k;
            const int glen = face_len(gf, k);
            for (int i = 0; i < glen; ++i) {
                const I ga = gf[i], gb = gf[i + 1 < glen ? i + 1 : 0];
                const bool same = (ga == a && gb == b) || (ga == b && gb == a);
                shared = shared || (same && ga != gb && !(g == (int32_t)f && i == j));
            }
        }
        exterior = !shared;
    }
    if (!exterior) return;
    atomicAdd(io_cnt + a, 1);
    atomicAdd(io_cnt + b, 1 << 16);
    out_code[a] = (int32_t)code;
    in_code[b] = (int32_t)code;
    const double2 c = cen[f];
    const double2 P = project_on_edge(xy, c, (int64_t)a, (int64_t)b);
    const double dx = P.x - c.x, dy = P.y - c.y;
    keep_flag[code] = sqrt(dx * dx + dy * dy) > VOR_X_EPSILON * VOR_X_EPSILON ? 1 : 0;
}

__global__ void __launch_bounds__(256)
k_vor_classify(int64_t n_node, const int32_t *__restrict__ ptr, const int32_t *__restrict__ io_cnt,
               const int32_t *__restrict__ out_code, const int32_t *__restrict__ in_code,
               const uint8_t *__restrict__ keep_flag, uint8_t *__restrict__ ext_flag,
               uint8_t *__restrict__ int_flag, unsigned long long *__restrict__ ctr) {
    const int64_t n = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    int members = 0;
    bool bad = false;
    if (n < n_node) {
        const int deg = ptr[n + 1] - ptr[n];
        const int io = io_cnt[n];
        const bool ext = io != 0;
        ext_flag[n] = ext ? 1 : 0;
        int_flag[n] = (!ext && deg > 0) ? 1 : 0;
        bad = ext && io != 0x00010001;
        members = deg;
        if (ext && !bad) members += keep_flag[out_code[n]] + keep_flag[in_code[n]] + 1;
    }
    for (int o = 16; o > 0; o >>= 1) members = max(members, __shfl_xor_sync(0xffffffffu, members, o));
    const unsigned any_bad = __ballot_sync(0xffffffffu, bad);
    if ((threadIdx.x & 31) == 0) {
        atomicMax(ctr + VC_KMAX, (unsigned long long)members);
        if (any_bad) atomicOr(ctr + VC_ERR, 1ull);
    }
}

struct VorOut {
    double2 *vertices;       // [centroids | kept projections | corners]
    int32_t *poly;           // (n_int + n_ext, kmax), -1 fill
    int32_t *vertex_face;    // source face per vertex, -1 for corners
    int32_t *corner_map;     // (n_ext, 2): the projection vertices (in, out) of every corner
    int64_t n_face, n_int, n_keep;
    int kmax;
};

template <int MAXM>
__device__ __forceinline__ void sort_by_angle(int m, double (&ang)[MAXM], int32_t (&id)[MAXM], double (&px)[MAXM],
                                              double (&py)[MAXM]) {
    // stable insertion sort (ties keep the member order, like the stable sorts of voronoi.py)
    for (int i = 1; i < m; ++i) {
        const double a = ang[i], x = px[i], y = py[i];
        const int32_t v = id[i];
        int j = i - 1;
        while (j >= 0 && ang[j] > a) {
            ang[j + 1] = ang[j]; id[j + 1] = id[j]; px[j + 1] = px[j]; py[j + 1] = py[j];
            --j;
        }
        ang[j + 1] = a; id[j + 1] = v; px[j + 1] = x; py[j + 1] = y;
    }
}

template <int MAXM>
__device__ __forceinline__ double member_area(int m, const double (&px)[MAXM], const double (&py)[MAXM]) {
    double sum = 0.0;
    for (int i = 0; i < m; ++i) {
        const int nx = i + 1 < m ? i + 1 : 0;
        const double rx = px[i] - px[0], ry = py[i] - py[0];
        const double qx = px[nx] - px[0], qy = py[nx] - py[0];
        sum += rx * qy - ry * qx;
    }
    return 0.5 * fabs(sum);
}

template <typename I, int MAXM>
__global__ void __launch_bounds__(128)
k_vor_polygons(const I *__restrict__ faces, int k, const double2 *__restrict__ xy,
               const double2 *__restrict__ cen, int64_t n_node, const int32_t *__restrict__ ptr,
               const int32_t *__restrict__ adj, const int32_t *__restrict__ io_cnt,
               const int32_t *__restrict__ out_code, const int32_t *__restrict__ in_code,
               const uint8_t *__restrict__ keep_flag, const int32_t *__restrict__ keep_rank,
               const int32_t *__restrict__ ext_rank, const int32_t *__restrict__ int_rank, const VorOut O) {
    const int64_t n = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (n >= n_node) return;
    const int b = ptr[n], deg = ptr[n + 1] - b;
    if (deg == 0 && io_cnt[n] == 0) return;
    double ang[MAXM], px[MAXM], py[MAXM];
    int32_t id[MAXM];
    const double2 node = xy[n];
    int m = 0;
    for (int q = 0; q < deg; ++q) {
        const int32_t f = adj[b + q];
        const double2 c = cen[f];
        id[m] = f; px[m] = c.x; py[m] = c.y;
        ++m;
    }
    if (io_cnt[n] == 0) {
        // interior: face centroids by angle around the node
        for (int i = 0; i < m; ++i) ang[i] = atan2(py[i] - node.y, px[i] - node.x);
        sort_by_angle<MAXM>(m, ang, id, px, py);
        int32_t *row = O.poly + (int64_t)int_rank[n] * O.kmax;
        for (int i = 0; i < O.kmax; ++i) row[i] = i < m ? id[i] : -1;
        return;
    }
    // exterior: + the projections on the two exterior edges that meet here + one corner
    const int r = ext_rank[n];
    const int32_t oc = out_code[n], ic = in_code[n];
    int32_t of, inf;
    int64_t oa, ob, ia, ib;
    edge_of_code(faces, k, oc, of, oa, ob);
    edge_of_code(faces, k, ic, inf, ia, ib);
    const double2 Pout = project_on_edge(xy, cen[of], oa, ob);
    const double2 Pin = project_on_edge(xy, cen[inf], ia, ib);
    const bool keep_out = keep_flag[oc], keep_in = keep_flag[ic];
    const int32_t pid_out = keep_out ? (int32_t)(O.n_face + keep_rank[oc]) : of;  // dropped -> the centroid itself
    const int32_t pid_in = keep_in ? (int32_t)(O.n_face + keep_rank[ic]) : inf;
    const int32_t corner_id = (int32_t)(O.n_face + O.n_keep + r);
    const double2 S = make_double2(0.5 * (Pin.x + Pout.x), 0.5 * (Pin.y + Pout.y));
    if (keep_out) { id[m] = pid_out; px[m] = Pout.x; py[m] = Pout.y; ++m; }
    if (keep_in) { id[m] = pid_in; px[m] = Pin.x; py[m] = Pin.y; ++m; }
    id[m] = corner_id; px[m] = S.x; py[m] = S.y; ++m;
    double sx = 0.0, sy = 0.0;
    for (int i = 0; i < m; ++i) { sx += px[i]; sy += py[i]; }
    const double mx = sx / (double)m, my = sy / (double)m;
    for (int i = 0; i < m; ++i) ang[i] = atan2(py[i] - my, px[i] - mx);
    sort_by_angle<MAXM>(m, ang, id, px, py);
    // corner: the node itself unless that makes the polygon smaller (a concave corner)
    const double area_mid = member_area<MAXM>(m, px, py);
    int at = 0;
    for (int i = 0; i < m; ++i) at = id[i] == corner_id ? i : at;
    px[at] = node.x; py[at] = node.y;
    const double area_node = member_area<MAXM>(m, px, py);
    const bool use_node = area_node >= area_mid * (1.0 - 1e-12);
    O.vertices[corner_id] = use_node ? node : S;
    O.vertex_face[corner_id] = -1;
    O.corner_map[2 * (int64_t)r] = pid_in;
    O.corner_map[2 * (int64_t)r + 1] = pid_out;
    if (keep_out) {
        O.vertices[pid_out] = Pout;
        O.vertex_face[pid_out] = of;
    }
    int32_t *row = O.poly + (O.n_int + r) * O.kmax;
    for (int i = 0; i < O.kmax; ++i) row[i] = i < m ? id[i] : -1;
}

__global__ void __launch_bounds__(256)
k_vor_centroid_vertices(const double2 *__restrict__ cen, int64_t n_face, double2 *__restrict__ vertices,
                        int32_t *__restrict__ vertex_face) {
    const int64_t f = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (f >= n_face) return;
    vertices[f] = cen[f];
    vertex_face[f] = (int32_t)f;
}

struct FlagToInt {
    __host__ __device__ __forceinline__ int32_t operator()(const uint8_t &v) const { return (int32_t)v; }
};

// ---- behind the barycentric kernel ----------------------------------------------------
struct BaryCsr {
    const int32_t *poly_index;   // (n_point) Voronoi polygon of every point, -1 outside
    double *weights;             // (n_point, kv) barycentric weights, adjusted in place
    const int32_t *poly;         // (n_poly, kv)
    const int32_t *vertex_face;
    const int32_t *corner_map;
    const double2 *vertices;
    const int32_t *source_face;  // (n_point) face of the SOURCE mesh that holds the point, -1 outside
    int64_t corner_start;
    int kv;
};

// unstructured.py:17-57 + 176-190 for one point; returns the number of positive weights
__global__ void __launch_bounds__(256)
k_bary_adjust_count(const BaryCsr B, int64_t n_point, int32_t *__restrict__ count) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n_point) return;
    double *w = B.weights + i * B.kv;
    const int32_t pi = B.poly_index[i];
    const int32_t *conn = B.poly + (int64_t)(pi < 0 ? 0 : pi) * B.kv;
    // the first corner vertex with a positive weight hands it to its two projections
    int col = -1;
    for (int j = 0; j < B.kv; ++j)
        if (col < 0 && conn[j] >= B.corner_start && w[j] > 0.0) col = j;
    if (col >= 0) {
        const int32_t p = conn[col];
        const double wp = w[col];
        const int32_t q = B.corner_map[2 * (p - B.corner_start)], r = B.corner_map[2 * (p - B.corner_start) + 1];
        const double2 vp = B.vertices[p], vq = B.vertices[q], vr = B.vertices[r];
        const double dqx = vq.x - vp.x, dqy = vq.y - vp.y, drx = vr.x - vp.x, dry = vr.y - vp.y;
        const double p_q = sqrt(dqx * dqx + dqy * dqy);
        const double p_r = sqrt(drx * drx + dry * dry);
        const double total = p_q + p_r;
        const double weight_q = (p_r / total) * wp;
        const double weight_r = (p_q / total) * wp;
        w[col] = 0.0;
        for (int j = 0; j < B.kv; ++j) {
            const int32_t c = conn[j];
            double v = w[j];
            v = v + (c == q ? weight_q : 0.0);
            v = v + (c == r ? weight_r : 0.0);
            w[j] = v;
        }
    }
    const bool outside = B.source_face[i] < 0;
    int n = 0;
    for (int j = 0; j < B.kv; ++j) {
        if (outside) w[j] = 0.0;
        n += w[j] > 0.0 ? 1 : 0;
    }
    count[i] = n;
    if (i == 0) count[n_point] = 0;
}

__global__ void __launch_bounds__(256)
k_bary_fill(const BaryCsr B, int64_t n_point, const int64_t *__restrict__ indptr, int32_t *__restrict__ indices,
            double *__restrict__ data) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n_point) return;
    const double *w = B.weights + i * B.kv;
    const int32_t pi = B.poly_index[i];
    const int32_t *conn = B.poly + (int64_t)(pi < 0 ? 0 : pi) * B.kv;
    int64_t at = indptr[i];
    for (int j = 0; j < B.kv; ++j)
        if (w[j] > 0.0) {
            const int32_t c = conn[j];
            indices[at] = B.vertex_face[c < 0 ? 0 : c];
            data[at] = w[j];
            ++at;
        }
}

struct CountToI64 {
    __host__ __device__ __forceinline__ int64_t operator()(const int32_t &v) const { return (int64_t)v; }
};

}  // namespace xg

using namespace xg;

struct xg_voronoi {
    cudaStream_t s = nullptr;
    const double *d_xy = nullptr;
    const void *d_faces = nullptr;
    const double *d_cen = nullptr;
    int idx_bytes = 0, k = 0;
    int64_t n_node = 0, n_face = 0;
    int64_t n_ext = 0, n_int = 0, n_keep = 0;
    int kmax = 0;
    DevBuf ptr, adj, io_cnt, out_code, in_code, keep_flag, keep_rank, ext_rank, int_rank;
};

#ifdef XG_VOR_DEBUG
#define VOR_CHECKPOINT(name)                                                                        \
    do {                                                                                            \
        cudaError_t e_ = cudaStreamSynchronize(s);                                                  \
        fprintf(stderr, "[vor] %s: %s\n", name, cudaGetErrorString(e_));                            \
        if (e_ != cudaSuccess) return cuda_fail(e_, name, __FILE__, __LINE__);                      \
    } while (0)
#else
#define VOR_CHECKPOINT(name) do { } while (0)
#endif

template <typename I>
static int voronoi_begin_impl(xg_voronoi *job, const I *faces, int64_t sizes[4]) {
    cudaStream_t s = job->s;
    const int64_t N = job->n_node, F = job->n_face;
    const int k = job->k;
    const double2 *xy = (const double2 *)job->d_xy;
    const double2 *cen = (const double2 *)job->d_cen;
    DevBuf deg, cursor, ext_flag, int_flag, tmp, ctr;
    XG_CUDA(deg.alloc(sizeof(int32_t) * (size_t)(N + 1), s));
    XG_CUDA(cursor.alloc(sizeof(int32_t) * (size_t)(N + 1), s));
    XG_CUDA(job->ptr.alloc(sizeof(int32_t) * (size_t)(N + 1), s));
    XG_CUDA(job->adj.alloc(sizeof(int32_t) * (size_t)(F * k + 1), s));
    XG_CUDA(job->io_cnt.alloc(sizeof(int32_t) * (size_t)(N + 1), s));
    XG_CUDA(job->out_code.alloc(sizeof(int32_t) * (size_t)(N + 1), s));
    XG_CUDA(job->in_code.alloc(sizeof(int32_t) * (size_t)(N + 1), s));
    XG_CUDA(job->keep_flag.alloc((size_t)(F * k + 1), s));
    XG_CUDA(job->keep_rank.alloc(sizeof(int32_t) * (size_t)(F * k + 1), s));
    XG_CUDA(job->ext_rank.alloc(sizeof(int32_t) * (size_t)(N + 1), s));
    XG_CUDA(job->int_rank.alloc(sizeof(int32_t) * (size_t)(N + 1), s));
    XG_CUDA(ext_flag.alloc((size_t)(N + 1), s));
    XG_CUDA(int_flag.alloc((size_t)(N + 1), s));
    XG_CUDA(ctr.alloc(sizeof(unsigned long long) * VC_N, s));
    XG_CUDA(cudaMemsetAsync(deg.p, 0, sizeof(int32_t) * (size_t)(N + 1), s));
    XG_CUDA(cudaMemsetAsync(cursor.p, 0, sizeof(int32_t) * (size_t)(N + 1), s));
    XG_CUDA(cudaMemsetAsync(job->io_cnt.p, 0, sizeof(int32_t) * (size_t)(N + 1), s));
    XG_CUDA(cudaMemsetAsync(job->keep_flag.p, 0, (size_t)(F * k + 1), s));
    XG_CUDA(cudaMemsetAsync(ext_flag.p, 0, (size_t)(N + 1), s));
    XG_CUDA(cudaMemsetAsync(int_flag.p, 0, (size_t)(N + 1), s));
    XG_CUDA(cudaMemsetAsync(ctr.p, 0, sizeof(unsigned long long) * VC_N, s));
    size_t tb = 0, tb2 = 0;
    {
        cub::TransformInputIterator<int32_t, FlagToInt, const uint8_t *> it(job->keep_flag.as<uint8_t>(), FlagToInt());
        XG_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tb, it, job->keep_rank.as<int32_t>(), F * k + 1, s));
        XG_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tb2, deg.as<int32_t>(), job->ptr.as<int32_t>(), N + 1, s));
        if (tb2 > tb) tb = tb2;
        XG_CUDA(tmp.alloc(tb, s));
    }
    const unsigned GF = (unsigned)div_up(F, 256), GN = (unsigned)div_up(N, 256);
    k_vor_degree<I><<<GF, 256, 0, s>>>(faces, F, k, deg.as<int32_t>());
    XG_CHECK_LAUNCH();
    VOR_CHECKPOINT("degree");
    { size_t t = tb; XG_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, t, deg.as<int32_t>(), job->ptr.as<int32_t>(), N + 1, s)); }
    VOR_CHECKPOINT("scan");
#ifdef XG_VOR_DEBUG
    {
        std::vector<int32_t> hp(N + 1), hd(N + 1);
        cudaMemcpy(hp.data(), job->ptr.p, sizeof(int32_t) * (N + 1), cudaMemcpyDeviceToHost);
        cudaMemcpy(hd.data(), deg.p, sizeof(int32_t) * (N + 1), cudaMemcpyDeviceToHost);
        long long sum = 0; int bad = 0;
        for (int64_t i = 0; i < N; ++i) { if (hp[i] != sum) ++bad; sum += hd[i]; }
        fprintf(stderr, "[vor] N=%lld F=%lld k=%d deg sum %lld ptr[N]=%d mismatches %d deg[N]=%d\n", (long long)N, (long long)F, k, sum, hp[N], bad, hd[N]);
    }
#endif
    k_vor_fill<I><<<GF, 256, 0, s>>>(faces, F, k, job->ptr.as<int32_t>(), cursor.as<int32_t>(), job->adj.as<int32_t>());
    XG_CHECK_LAUNCH();
    VOR_CHECKPOINT("fill");
    k_vor_sort_adj<<<GN, 256, 0, s>>>(job->ptr.as<int32_t>(), N, job->adj.as<int32_t>());
    XG_CHECK_LAUNCH();
    VOR_CHECKPOINT("sort_adj");
#ifdef XG_VOR_DEBUG
    {
        std::vector<int32_t> ha(F * k + 1);
        cudaMemcpy(ha.data(), job->adj.p, sizeof(int32_t) * (F * k + 1), cudaMemcpyDeviceToHost);
        int bad = 0; long long first_bad = -1;
        std::vector<int32_t> hp(N + 1);
        cudaMemcpy(hp.data(), job->ptr.p, sizeof(int32_t) * (N + 1), cudaMemcpyDeviceToHost);
        for (int64_t i = 0; i < hp[N]; ++i) if (ha[i] < 0 || ha[i] >= F) { if (first_bad < 0) first_bad = i; ++bad; }
        fprintf(stderr, "[vor] adj entries out of range: %d (first at %lld, value %d)\n", bad, first_bad, first_bad >= 0 ? ha[first_bad] : 0);
    }
#endif
    k_vor_exterior<I><<<(unsigned)div_up(F * k, 256), 256, 0, s>>>(
        faces, F, k, xy, cen, job->ptr.as<int32_t>(), job->adj.as<int32_t>(), job->io_cnt.as<int32_t>(),
        job->out_code.as<int32_t>(), job->in_code.as<int32_t>(), job->keep_flag.as<uint8_t>());
    XG_CHECK_LAUNCH();
    VOR_CHECKPOINT("exterior");
    k_vor_classify<<<GN, 256, 0, s>>>(N, job->ptr.as<int32_t>(), job->io_cnt.as<int32_t>(), job->out_code.as<int32_t>(),
                                      job->in_code.as<int32_t>(), job->keep_flag.as<uint8_t>(), ext_flag.as<uint8_t>(),
                                      int_flag.as<uint8_t>(), ctr.as<unsigned long long>());
    XG_CHECK_LAUNCH();
    {
        cub::TransformInputIterator<int32_t, FlagToInt, const uint8_t *> itk(job->keep_flag.as<uint8_t>(), FlagToInt());
        cub::TransformInputIterator<int32_t, FlagToInt, const uint8_t *> ite(ext_flag.as<uint8_t>(), FlagToInt());
        cub::TransformInputIterator<int32_t, FlagToInt, const uint8_t *> iti(int_flag.as<uint8_t>(), FlagToInt());
        size_t t = tb;
        XG_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, t, itk, job->keep_rank.as<int32_t>(), F * k + 1, s));
        t = tb;
        XG_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, t, ite, job->ext_rank.as<int32_t>(), N + 1, s));
        t = tb;
        XG_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, t, iti, job->int_rank.as<int32_t>(), N + 1, s));
    }
    // sizes -> host (the caller allocates the outputs)
    double *slot = pinned_slot_acquire();
    if (!slot) { set_error("voronoi: no pinned host memory"); return XG_ERR_NOMEM; }
    int32_t *h = reinterpret_cast<int32_t *>(slot);
    unsigned long long *hc = reinterpret_cast<unsigned long long *>(slot + 4);
    cudaError_t e = cudaMemcpyAsync(h + 0, job->ext_rank.as<int32_t>() + N, sizeof(int32_t), cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaMemcpyAsync(h + 1, job->int_rank.as<int32_t>() + N, sizeof(int32_t), cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaMemcpyAsync(h + 2, job->keep_rank.as<int32_t>() + F * k, sizeof(int32_t), cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaMemcpyAsync(hc, ctr.p, sizeof(unsigned long long) * VC_N, cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    const int64_t n_ext = h[0], n_int = h[1], n_keep = h[2];
    const unsigned long long err = hc[VC_ERR], kmax = hc[VC_KMAX];
    pinned_slot_release(slot);
    if (e != cudaSuccess) return cuda_fail(e, "voronoi sizes", __FILE__, __LINE__);
    if (err) {
        set_error("xg_voronoi_begin: an exterior node is not met by exactly one incoming and one outgoing "
                  "exterior edge (the mesh exterior is not a set of simple closed loops)");
        return XG_ERR_INVALID;
    }
    if (kmax > 64) { set_error("xg_voronoi_begin: a Voronoi polygon has %llu vertices (limit 64)", kmax); return XG_ERR_UNSUPPORTED; }
    job->n_ext = n_ext; job->n_int = n_int; job->n_keep = n_keep; job->kmax = (int)kmax;
    sizes[0] = n_int; sizes[1] = n_ext; sizes[2] = n_keep; sizes[3] = (int64_t)kmax;
    return XG_OK;
}

extern "C" int xg_voronoi_begin(const double *d_xy, int64_t n_node, const void *d_faces, int idx_bytes,
                                int64_t n_face, int k, const double *d_centroids, int64_t sizes[4],
                                void *stream, xg_voronoi **out) {
    if (!out || !sizes || n_node < 0 || n_face < 0 || k < 3 || k > XG_MAX_N_VERTEX || (idx_bytes != 4 && idx_bytes != 8) ||
        (n_face > 0 && (!d_xy || !d_faces || !d_centroids))) {
        set_error("xg_voronoi_begin: invalid argument");
        return XG_ERR_INVALID;
    }
    *out = nullptr;
    if (n_face * k >= ((int64_t)1 << 31) - 1 || n_node >= ((int64_t)1 << 31) - 1) {
        set_error("xg_voronoi_begin: more than 2^31 face-node incidences");
        return XG_ERR_UNSUPPORTED;
    }
    xg_voronoi *job = new xg_voronoi();
    job->s = (cudaStream_t)stream; job->d_xy = d_xy; job->d_faces = d_faces; job->d_cen = d_centroids;
    job->idx_bytes = idx_bytes; job->k = k; job->n_node = n_node; job->n_face = n_face;
    int rc;
    if (n_face == 0 || n_node == 0) {
        sizes[0] = sizes[1] = sizes[2] = sizes[3] = 0;
        rc = XG_OK;
    } else if (idx_bytes == 4) {
        rc = voronoi_begin_impl<int32_t>(job, (const int32_t *)d_faces, sizes);
    } else {
        rc = voronoi_begin_impl<int64_t>(job, (const int64_t *)d_faces, sizes);
    }
    if (rc) { delete job; return rc; }
    *out = job;
    return XG_OK;
}

template <typename I>
static int voronoi_finish_impl(xg_voronoi *job, const I *faces, const VorOut &O, cudaStream_t s) {
    const int64_t N = job->n_node, F = job->n_face;
    k_vor_centroid_vertices<<<(unsigned)div_up(F, 256), 256, 0, s>>>((const double2 *)job->d_cen, F, O.vertices,
                                                                   O.vertex_face);
    XG_CHECK_LAUNCH();
    const unsigned G = (unsigned)div_up(N, 128);
#define XG_VOR_LAUNCH(M)                                                                                         \
    k_vor_polygons<I, M><<<G, 128, 0, s>>>(faces, job->k, (const double2 *)job->d_xy, (const double2 *)job->d_cen, N, \
                                           job->ptr.as<int32_t>(), job->adj.as<int32_t>(), job->io_cnt.as<int32_t>(),   \
                                           job->out_code.as<int32_t>(), job->in_code.as<int32_t>(),                     \
                                           job->keep_flag.as<uint8_t>(), job->keep_rank.as<int32_t>(),                  \
                                           job->ext_rank.as<int32_t>(), job->int_rank.as<int32_t>(), O)
    if (job->kmax <= 12) XG_VOR_LAUNCH(12);
    else if (job->kmax <= 24) XG_VOR_LAUNCH(24);
    else XG_VOR_LAUNCH(64);
#undef XG_VOR_LAUNCH
    XG_CHECK_LAUNCH();
    return XG_OK;
}

extern "C" int xg_voronoi_finish(xg_voronoi *job, double *d_vertices, int32_t *d_polygons, int32_t *d_vertex_face,
                                 int32_t *d_corner_map, void *stream) {
    if (!job) { set_error("xg_voronoi_finish: NULL job"); return XG_ERR_INVALID; }
    cudaStream_t s = (cudaStream_t)stream;
    int rc = XG_OK;
    if (job->n_face > 0 && job->n_node > 0) {
        if (!d_vertices || !d_vertex_face || (job->n_int + job->n_ext > 0 && !d_polygons) || (job->n_ext > 0 && !d_corner_map)) {
            set_error("xg_voronoi_finish: NULL output");
            delete job;
            return XG_ERR_INVALID;
        }
        VorOut O;
        O.vertices = (double2 *)d_vertices; O.poly = d_polygons; O.vertex_face = d_vertex_face; O.corner_map = d_corner_map;
        O.n_face = job->n_face; O.n_int = job->n_int; O.n_keep = job->n_keep; O.kmax = job->kmax;
        if (job->idx_bytes == 4) rc = voronoi_finish_impl<int32_t>(job, (const int32_t *)job->d_faces, O, s);
        else rc = voronoi_finish_impl<int64_t>(job, (const int64_t *)job->d_faces, O, s);
    }
    if (s != job->s) cudaStreamSynchronize(s);
    delete job;
    return rc;
}

extern "C" int xg_voronoi_abort(xg_voronoi *job) {
    delete job;
    return XG_OK;
}

extern "C" int xg_bary_csr_count(const int32_t *d_poly_index, double *d_weights, int64_t n_point, int kv,
                                 const int32_t *d_polygons, const int32_t *d_vertex_face,
                                 const int32_t *d_corner_map, const double *d_vertices, int64_t corner_start,
                                 const int32_t *d_source_face, int64_t *d_indptr, int64_t *h_nnz, void *stream) {
    if (n_point < 0 || kv < 1 || !d_indptr || !h_nnz ||
        (n_point > 0 && (!d_poly_index || !d_weights || !d_polygons || !d_vertex_face || !d_vertices || !d_source_face))) {
        set_error("xg_bary_csr_count: invalid argument");
        return XG_ERR_INVALID;
    }
    cudaStream_t s = (cudaStream_t)stream;
    *h_nnz = 0;
    if (n_point == 0) {
        XG_CUDA(cudaMemsetAsync(d_indptr, 0, sizeof(int64_t), s));
        return XG_OK;
    }
    BaryCsr B;
    B.poly_index = d_poly_index; B.weights = d_weights; B.poly = d_polygons; B.vertex_face = d_vertex_face;
    B.corner_map = d_corner_map; B.vertices = (const double2 *)d_vertices; B.source_face = d_source_face;
    B.corner_start = corner_start; B.kv = kv;
    DevBuf count, tmp;
    XG_CUDA(count.alloc(sizeof(int32_t) * (size_t)(n_point + 1), s));
    k_bary_adjust_count<<<(unsigned)div_up(n_point, 256), 256, 0, s>>>(B, n_point, count.as<int32_t>());
    XG_CHECK_LAUNCH();
    cub::TransformInputIterator<int64_t, CountToI64, const int32_t *> it(count.as<int32_t>(), CountToI64());
    size_t tb = 0;
    XG_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tb, it, d_indptr, n_point + 1, s));
    XG_CUDA(tmp.alloc(tb, s));
    XG_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, tb, it, d_indptr, n_point + 1, s));
    double *slot = pinned_slot_acquire();
    if (!slot) { set_error("bary csr: no pinned host memory"); return XG_ERR_NOMEM; }
    cudaError_t e = cudaMemcpyAsync(slot, d_indptr + n_point, sizeof(int64_t), cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    memcpy(h_nnz, slot, sizeof(int64_t));
    pinned_slot_release(slot);
    if (e != cudaSuccess) return cuda_fail(e, "bary csr nnz", __FILE__, __LINE__);
    return XG_OK;
}

extern "C" int xg_bary_csr_fill(const int32_t *d_poly_index, const double *d_weights, int64_t n_point, int kv,
                                const int32_t *d_polygons, const int32_t *d_vertex_face, const int64_t *d_indptr,
                                int32_t *d_indices, double *d_data, void *stream) {
    if (n_point < 0 || kv < 1 || (n_point > 0 && (!d_poly_index || !d_weights || !d_polygons || !d_vertex_face || !d_indptr))) {
        set_error("xg_bary_csr_fill: invalid argument");
        return XG_ERR_INVALID;
    }
    if (n_point == 0) return XG_OK;
    cudaStream_t s = (cudaStream_t)stream;
    BaryCsr B;
    B.poly_index = d_poly_index; B.weights = const_cast<double *>(d_weights); B.poly = d_polygons;
    B.vertex_face = d_vertex_face; B.corner_map = nullptr; B.vertices = nullptr; B.source_face = nullptr;
    B.corner_start = 0; B.kv = kv;
    k_bary_fill<<<(unsigned)div_up(n_point, 256), 256, 0, s>>>(B, n_point, d_indptr, d_indices, d_data);
    XG_CHECK_LAUNCH();
    return XG_OK;
}
