// xg_apply.cu -- weight application and per-face geometry.
//
// xg_csr_apply replaces make_regrid(func)._regrid
// (/root/reference/xugrid/regrid/regridder.py:34-69) with the reductions of
// /root/reference/xugrid/regrid/reduce.py; xg_coo_gather replaces
// CentroidLocatorRegridder._regrid (regridder.py:400-409); xg_face_area /
// xg_face_centroid replace connectivity.area / connectivity.centroids
// (/root/reference/xugrid/ugrid/connectivity.py:615-664).
//
// The reference walks the rows serially per time step (prange over n_extra
// only) with a per-thread workspace.  Here one thread owns one (row, step):
// a warp therefore reads one contiguous span of indices/weights and issues its
// value gathers together; every reduction is written as a streaming pass (or
// an O(n^2) re-read for mode / percentile) so no workspace exists at all.
// Compiled with -fmad=false: sums round exactly like the reference's.
#include "xg_common.cuh"

namespace xg {

struct Row {
    const int32_t *__restrict__ idx;
    const double *__restrict__ w;
    const double *__restrict__ src;
    int n;
    __device__ __forceinline__ double v(int i) const { return src[idx[i]]; }
};

__device__ __forceinline__ double r_mean(const Row &r) {
    double vsum = 0.0, wsum = 0.0;
    for (int i = 0; i < r.n; ++i) {
        const double v = r.v(i), w = r.w[i];
        if (isnan(v)) continue;
        vsum += w * v;
        wsum += w;
    }
    return wsum == 0 ? NAN : vsum / wsum;
}

__device__ __forceinline__ double r_harmonic(const Row &r) {
    double v_agg = 0.0, w_sum = 0.0;
    for (int i = 0; i < r.n; ++i) {
        const double v = r.v(i), w = r.w[i];
        if (isnan(v) || v == 0) continue;
        if (w > 0) { w_sum += w; v_agg += w / v; }
    }
    return (v_agg == 0 || w_sum == 0) ? NAN : w_sum / v_agg;
}

__device__ __forceinline__ double r_geometric(const Row &r) {
    double normsum = 0.0;
    for (int i = 0; i < r.n; ++i) normsum += r.w[i];
    if (normsum == 0) return NAN;
    double v_agg = 0.0, w_sum = 0.0;
    for (int i = 0; i < r.n; ++i) {
        const double v = r.v(i);
        const double w = r.w[i] / normsum;
        if (v > 0 && w > 0) { v_agg += w * log(fabs(v)); w_sum += w; }
        else if (v < 0) return NAN;
    }
    return w_sum == 0 ? NAN : exp((1.0 / w_sum) * v_agg);
}

__device__ __forceinline__ double r_sum(const Row &r) {
    double v_sum = 0.0, w_sum = 0.0;
    for (int i = 0; i < r.n; ++i) {
        const double v = r.v(i);
        if (isnan(v)) continue;
        v_sum += v;
        w_sum += r.w[i];
    }
    return w_sum == 0 ? NAN : v_sum;
}

__device__ __forceinline__ double r_minimum(const Row &r) {
    double v_min = INFINITY, w_max = 0.0;
    for (int i = 0; i < r.n; ++i) {
        const double v = r.v(i), w = r.w[i];
        if (isnan(v)) continue;
        v_min = v < v_min ? v : v_min;
        w_max = w > w_max ? w : w_max;
    }
    return w_max == 0.0 ? NAN : v_min;
}

__device__ __forceinline__ double r_maximum(const Row &r) {
    double v_max = -INFINITY, w_max = 0.0;
    for (int i = 0; i < r.n; ++i) {
        const double v = r.v(i), w = r.w[i];
        if (isnan(v)) continue;
        v_max = v > v_max ? v : v_max;
        w_max = w > w_max ? w : w_max;
    }
    return w_max == 0.0 ? NAN : v_max;
}

// Weighted mode.  reduce.py:126-158 accumulates the weight of every later
// duplicate onto the FIRST occurrence of a value (in row order) and leaves the
// other entries at their own weight; the winner is the largest accumulated
// weight, ties to the larger value.  Same sums, same order, no workspace.
__device__ __forceinline__ double r_mode(const Row &r) {
    int n_valid = 0;
    double w_max = 0.0;
    for (int i = 0; i < r.n; ++i) {
        if (isnan(r.v(i))) continue;
        w_max = r.w[i] > w_max ? r.w[i] : w_max;
        ++n_valid;
    }
    if (n_valid == 0 || w_max == 0.0) return NAN;
    double best_w = 0.0;
    double mode_value = r.v(0);
    for (int i = 0; i < r.n; ++i) {
        const double vi = r.v(i);
        if (isnan(vi)) continue;
        bool first = true;
        for (int j = 0; j < i; ++j)
            if (r.v(j) == vi) { first = false; break; }
        double acc = r.w[i];
        if (first)
            for (int j = i + 1; j < r.n; ++j)
                if (r.v(j) == vi) acc += r.w[j];
        if (acc > best_w || (acc == best_w && vi > mode_value)) { best_w = acc; mode_value = vi; }
    }
    return mode_value;
}

// reduce.py:161-203: linear interpolation between the two closest ranks of the
// non-NaN values; weights only matter through the all-zero early-out.  The
// k-th / k+1-th smallest are found by rank counting.
__device__ __forceinline__ double r_percentile(const Row &r, double p) {
    double w_max = 0.0;
    for (int i = 0; i < r.n; ++i) w_max = r.w[i] > w_max ? r.w[i] : w_max;
    if (w_max == 0.0) return NAN;
    if (p == 0) return r_minimum(r);
    if (p == 100) return r_maximum(r);
    int m = 0;
    double only = NAN;
    for (int i = 0; i < r.n; ++i) {
        const double v = r.v(i);
        if (!isnan(v)) { ++m; only = v; }
    }
    if (m == 0) return NAN;
    if (m == 1) return only;
    const double rank = 1 + (m - 1) * p / 100.0;
    const double f = floor(rank);
    const double frac = rank - f;
    const int k = (int)(f - 1);
    double lower = NAN, upper = NAN;
    for (int i = 0; i < r.n; ++i) {
        const double vi = r.v(i);
        if (isnan(vi)) continue;
        int rk = 0;
        for (int j = 0; j < r.n; ++j) {
            const double vj = r.v(j);
            if (isnan(vj)) continue;
            rk += (vj < vi) || (vj == vi && j < i);
        }
        if (rk == k) lower = vi;
        if (rk == k + 1) upper = vi;
    }
    return lower * (1 - frac) + upper * frac;
}

__device__ __forceinline__ double r_max_overlap(const Row &r) {
    double w_max = 0.0, v_max = -INFINITY;
    for (int i = 0; i < r.n; ++i) {
        const double v = r.v(i), w = r.w[i];
        if (!isnan(v)) {
            if (w > w_max || (w == w_max && v > v_max)) { w_max = w; v_max = v; }
        }
    }
    return w_max == 0.0 ? NAN : v_max;
}

__device__ __forceinline__ double r_conservative(const Row &r) {
    double v_agg = 0.0, w_sum = 0.0;
    for (int i = 0; i < r.n; ++i) {
        const double v = r.v(i), w = r.w[i];
        if (isnan(v)) continue;
        v_agg += v * w;
        w_sum += w;
    }
    return w_sum == 0 ? NAN : v_agg;
}

// Where a result goes: one array, or the same position of the output array of EVERY rank of
// the job (peer device memory mapped over NVLink / NVSwitch).  The second form is the apply
// fused with the all-gather that reassembles the output: a rank's rows are stored straight
// into all copies while the kernel runs, instead of into a send slab that a collective
// then moves (xg_csr_apply_peers).
constexpr int MAX_PEERS = 16;
struct OutSet {
    double *p[MAX_PEERS];
    int n;
    __device__ __forceinline__ void store(int64_t at, double v) const {
#pragma unroll 1
        for (int i = 0; i < n; ++i) p[i][at] = v;
    }
};

template <int METHOD>
__global__ void __launch_bounds__(256)
k_csr_apply(const int64_t *__restrict__ indptr, const int32_t *__restrict__ indices,
            const double *__restrict__ weights, const double *__restrict__ src, int64_t n_extra,
            int64_t m, int64_t n_rows, double percentile, const OutSet out, int64_t out_ld) {
    const int64_t row = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (row >= n_rows) return;
    const int64_t a = indptr[row], b = indptr[row + 1];
    Row r;
    r.idx = indices + a;
    r.w = weights + a;
    r.n = (int)(b - a);
    for (int64_t e = blockIdx.y; e < n_extra; e += gridDim.y) {
        r.src = src + e * m;
        double res = NAN;
        if (r.n > 0) {
            if (METHOD == XG_MEAN) res = r_mean(r);
            else if (METHOD == XG_HARMONIC_MEAN) res = r_harmonic(r);
            else if (METHOD == XG_GEOMETRIC_MEAN) res = r_geometric(r);
            else if (METHOD == XG_SUM) res = r_sum(r);
            else if (METHOD == XG_MINIMUM) res = r_minimum(r);
            else if (METHOD == XG_MAXIMUM) res = r_maximum(r);
            else if (METHOD == XG_MODE) res = r_mode(r);
            else if (METHOD == XG_PERCENTILE) res = r_percentile(r, percentile);
            else if (METHOD == XG_MAX_OVERLAP) res = r_max_overlap(r);
            else if (METHOD == XG_CONSERVATIVE) res = r_conservative(r);
        }
        out.store(e * out_ld + row, res);
    }
}

// Several time steps per thread for the single-pass reductions: the row's indices
// and weights are read ONCE for APPLY_E steps (in the per-step kernel they are
// re-read from HBM for every step -- 690 MB per step at 10 M rows, more than the
// L2 holds) and the APPLY_E value gathers of an entry are independent loads in
// flight together.  Per step the accumulation order over the row is unchanged, so
// results are bit-identical to the per-step kernel.
constexpr int APPLY_E = 8;

template <int METHOD>
struct StreamAcc {
    double a, b;
    __device__ __forceinline__ void init() {
        a = (METHOD == XG_MINIMUM) ? INFINITY : (METHOD == XG_MAXIMUM || METHOD == XG_MAX_OVERLAP) ? -INFINITY : 0.0;
        b = 0.0;
    }
    __device__ __forceinline__ void add(double v, double w) {
        if (METHOD == XG_MEAN) { if (!isnan(v)) { a += w * v; b += w; } }
        else if (METHOD == XG_HARMONIC_MEAN) { if (!(isnan(v) || v == 0) && w > 0) { b += w; a += w / v; } }
        else if (METHOD == XG_SUM) { if (!isnan(v)) { a += v; b += w; } }
        else if (METHOD == XG_MINIMUM) { if (!isnan(v)) { a = v < a ? v : a; b = w > b ? w : b; } }
        else if (METHOD == XG_MAXIMUM) { if (!isnan(v)) { a = v > a ? v : a; b = w > b ? w : b; } }
        else if (METHOD == XG_MAX_OVERLAP) { if (!isnan(v) && (w > b || (w == b && v > a))) { b = w; a = v; } }
        else if (METHOD == XG_CONSERVATIVE) { if (!isnan(v)) { a += v * w; b += w; } }
    }
    __device__ __forceinline__ double result() const {
        if (METHOD == XG_MEAN) return b == 0 ? NAN : a / b;
        if (METHOD == XG_HARMONIC_MEAN) return (a == 0 || b == 0) ? NAN : b / a;
        return b == 0.0 ? NAN : a;  // sum, minimum, maximum, max_overlap, conservative
    }
};

template <int METHOD>
__global__ void __launch_bounds__(256)
k_csr_apply_steps(const int64_t *__restrict__ indptr, const int32_t *__restrict__ indices,
                  const double *__restrict__ weights, const double *__restrict__ src, int64_t n_extra,
                  int64_t m, int64_t n_rows, const OutSet out, int64_t out_ld) {
    const int64_t row = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (row >= n_rows) return;
    const int64_t a = indptr[row], b = indptr[row + 1];
    for (int64_t e0 = (int64_t)blockIdx.y * APPLY_E; e0 < n_extra; e0 += (int64_t)gridDim.y * APPLY_E) {
        StreamAcc<METHOD> acc[APPLY_E];
#pragma unroll
        for (int u = 0; u < APPLY_E; ++u) acc[u].init();
        const double *s0 = src + e0 * m;
        const int ne = (int)min((int64_t)APPLY_E, n_extra - e0);
        for (int64_t i = a; i < b; ++i) {
            const int64_t c = indices[i];
            const double w = weights[i];
            double v[APPLY_E];
#pragma unroll
            for (int u = 0; u < APPLY_E; ++u) v[u] = u < ne ? s0[u * m + c] : NAN;
#pragma unroll
            for (int u = 0; u < APPLY_E; ++u) acc[u].add(v[u], w);
        }
#pragma unroll
        for (int u = 0; u < APPLY_E; ++u)
            if (u < ne) out.store((e0 + u) * out_ld + row, (b > a) ? acc[u].result() : NAN);
    }
}

__global__ void __launch_bounds__(256) k_fill_nan(double *__restrict__ out, int64_t n) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x)
        out[i] = NAN;
}

__global__ void __launch_bounds__(256)
k_coo_gather(const int32_t *__restrict__ row, const int32_t *__restrict__ col, int64_t nnz,
             const double *__restrict__ src, int64_t n_extra, int64_t m, double *__restrict__ out,
             int64_t size) {
    const int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (q >= nnz) return;
    const int64_t r = row[q], c = col[q];
    for (int64_t e = blockIdx.y; e < n_extra; e += gridDim.y) out[e * size + r] = src[e * m + c];
}

template <typename I>
__global__ void __launch_bounds__(256)
k_face_area(const double2 *__restrict__ xy, const I *__restrict__ faces, int64_t n_face, int k,
            double *__restrict__ area) {
    const int64_t f = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (f >= n_face) return;
    const I *face = faces + f * k;
    const I i0 = face[0];
    const double2 p0 = xy[i0];
    double det = 0.0;
    for (int i = 0; i < k; ++i) {
        const I ia = face[i] < 0 ? i0 : face[i];
        const I ib = (i + 1 < k) ? (face[i + 1] < 0 ? i0 : face[i + 1]) : i0;
        const double2 pa = xy[ia], pb = xy[ib];
        const double ax = pa.x - p0.x, ay = pa.y - p0.y;
        const double bx = pb.x - p0.x, by = pb.y - p0.y;
        det += ax * by - ay * bx;
    }
    area[f] = 0.5 * fabs(det);
}

template <typename I>
__global__ void __launch_bounds__(256)
k_face_centroid(const double2 *__restrict__ xy, const I *__restrict__ faces, int64_t n_face, int k,
                double2 *__restrict__ cxy) {
    const int64_t f = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (f >= n_face) return;
    const I *face = faces + f * k;
    if (k == 3) {
        double sx = 0.0, sy = 0.0;
        for (int i = 0; i < 3; ++i) {
            const double2 p = xy[face[i]];
            sx += p.x;
            sy += p.y;
        }
        cxy[f] = make_double2(sx / 3.0, sy / 3.0);
        return;
    }
    const I i0 = face[0];
    const double2 p0 = xy[i0];
    double det_sum = 0.0, cx = 0.0, cy = 0.0;
    for (int i = 0; i < k; ++i) {
        const I ia = face[i] < 0 ? i0 : face[i];
        const I ib = (i + 1 < k) ? (face[i + 1] < 0 ? i0 : face[i + 1]) : i0;
        const double2 pa = xy[ia], pb = xy[ib];
        const double ax = pa.x - p0.x, ay = pa.y - p0.y;
        const double bx = pb.x - p0.x, by = pb.y - p0.y;
        const double det = ax * by - ay * bx;
        det_sum += det;
        cx += (ax + bx) * det;
        cy += (ay + by) * det;
    }
    const double aw = 1.0 / (3.0 * det_sum);
    cxy[f] = make_double2(aw * cx + p0.x, aw * cy + p0.y);
}

// 24-bit little-endian indices (0xFFFFFF = fill) -> int32 (-1 = fill): the transport encoding
// of a connectivity whose mesh has fewer than 2^24 - 1 nodes (xg_unpack_u24).  One thread
// unpacks 4 indices = three aligned 32-bit words.
__global__ void __launch_bounds__(256)
k_unpack_u24(const uint8_t *__restrict__ packed, int64_t n, int32_t *__restrict__ out) {
    const int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;  // group of 4 values
    const int64_t i = q * 4;
    if (i >= n) return;
    auto widen = [](uint32_t v) { return v == 0xFFFFFFu ? -1 : (int32_t)v; };
    if (i + 4 <= n) {
        const uint32_t *w = reinterpret_cast<const uint32_t *>(packed + i * 3);
        const uint32_t w0 = w[0], w1 = w[1], w2 = w[2];
        int4 r;
        r.x = widen(w0 & 0xFFFFFFu);
        r.y = widen((w0 >> 24) | ((w1 & 0xFFFFu) << 8));
        r.z = widen((w1 >> 16) | ((w2 & 0xFFu) << 16));
        r.w = widen(w2 >> 8);
        *reinterpret_cast<int4 *>(out + i) = r;
    } else {
        for (int64_t j = i; j < n; ++j) {
            const uint8_t *b = packed + j * 3;
            out[j] = widen((uint32_t)b[0] | ((uint32_t)b[1] << 8) | ((uint32_t)b[2] << 16));
        }
    }
}

}  // namespace xg

using namespace xg;

extern "C" int xg_unpack_u24(const void *d_packed, int64_t n_values, int32_t *d_out, void *stream) {
    if (n_values < 0 || (n_values > 0 && (!d_packed || !d_out))) { set_error("xg_unpack_u24: invalid argument"); return XG_ERR_INVALID; }
    if (((uintptr_t)d_packed & 3u) || ((uintptr_t)d_out & 15u)) {
        set_error("xg_unpack_u24: the packed array must be 4-byte aligned, the output 16-byte aligned");
        return XG_ERR_INVALID;
    }
    if (n_values == 0) return XG_OK;
    k_unpack_u24<<<(unsigned)div_up(div_up(n_values, 4), 256), 256, 0, (cudaStream_t)stream>>>(
        (const uint8_t *)d_packed, n_values, d_out);
    XG_CHECK_LAUNCH();
    return XG_OK;
}

static int csr_apply_impl(int method, double percentile, const int64_t *d_indptr,
                          const int32_t *d_indices, const double *d_weights, const double *d_src,
                          int64_t n_extra, int64_t m, int64_t n_rows, const OutSet d_out,
                          int64_t out_ld, void *stream) {
    if (n_rows < 0 || n_extra < 0 || m < 0 || out_ld < n_rows) { set_error("xg_csr_apply: invalid size"); return XG_ERR_INVALID; }
    if (n_rows == 0 || n_extra == 0) return XG_OK;
    if (!d_indptr) { set_error("xg_csr_apply: NULL argument"); return XG_ERR_INVALID; }
    if (method == XG_PERCENTILE && !(percentile >= 0.0 && percentile <= 100.0)) {
        set_error("percentile must be in the range [0, 100], received: %g", percentile);
        return XG_ERR_INVALID;
    }
    cudaStream_t s = (cudaStream_t)stream;
    if (n_extra >= 2) {
        // several steps per thread for the single-pass reductions
        const int64_t chunks = div_up(n_extra, APPLY_E);
        const dim3 g((unsigned)div_up(n_rows, 256), (unsigned)(chunks < 65535 ? chunks : 65535));
#define XG_STEPS_CASE(M)                                                                              \
    case M:                                                                                           \
        k_csr_apply_steps<M><<<g, 256, 0, s>>>(d_indptr, d_indices, d_weights, d_src, n_extra, m,    \
                                               n_rows, d_out, out_ld);                               \
        XG_CHECK_LAUNCH();                                                                            \
        return XG_OK;
        switch (method) {
            XG_STEPS_CASE(XG_MEAN)
            XG_STEPS_CASE(XG_HARMONIC_MEAN)
            XG_STEPS_CASE(XG_SUM)
            XG_STEPS_CASE(XG_MINIMUM)
            XG_STEPS_CASE(XG_MAXIMUM)
            XG_STEPS_CASE(XG_MAX_OVERLAP)
            XG_STEPS_CASE(XG_CONSERVATIVE)
        default:
            break;  // two-pass / O(n^2) reductions: one step per thread below
        }
#undef XG_STEPS_CASE
    }
    const dim3 grid((unsigned)div_up(n_rows, 256), (unsigned)(n_extra < 65535 ? n_extra : 65535));
#define XG_APPLY_CASE(M)                                                                         \
    case M:                                                                                      \
        k_csr_apply<M><<<grid, 256, 0, s>>>(d_indptr, d_indices, d_weights, d_src, n_extra, m,   \
                                            n_rows, percentile, d_out, out_ld);                 \
        break;
    switch (method) {
        XG_APPLY_CASE(XG_MEAN)
        XG_APPLY_CASE(XG_HARMONIC_MEAN)
        XG_APPLY_CASE(XG_GEOMETRIC_MEAN)
        XG_APPLY_CASE(XG_SUM)
        XG_APPLY_CASE(XG_MINIMUM)
        XG_APPLY_CASE(XG_MAXIMUM)
        XG_APPLY_CASE(XG_MODE)
        XG_APPLY_CASE(XG_PERCENTILE)
        XG_APPLY_CASE(XG_MAX_OVERLAP)
        XG_APPLY_CASE(XG_CONSERVATIVE)
    default:
        set_error("Invalid regridding method code %d", method);
        return XG_ERR_INVALID;
    }
#undef XG_APPLY_CASE
    XG_CHECK_LAUNCH();
    return XG_OK;
}

extern "C" int xg_coo_gather(const int32_t *d_row, const int32_t *d_col, int64_t nnz,
                             const double *d_src, int64_t n_extra, int64_t m, double *d_out,
                             int64_t size, void *stream) {
    if (nnz < 0 || n_extra < 0 || size < 0 || m < 0) { set_error("xg_coo_gather: invalid size"); return XG_ERR_INVALID; }
    cudaStream_t s = (cudaStream_t)stream;
    if (n_extra * size > 0) {
        k_fill_nan<<<num_sms() * 8, 256, 0, s>>>(d_out, n_extra * size);
        XG_CHECK_LAUNCH();
    }
    if (nnz > 0 && n_extra > 0) {
        const dim3 grid((unsigned)div_up(nnz, 256), (unsigned)(n_extra < 65535 ? n_extra : 65535));
        k_coo_gather<<<grid, 256, 0, s>>>(d_row, d_col, nnz, d_src, n_extra, m, d_out, size);
        XG_CHECK_LAUNCH();
    }
    return XG_OK;
}

extern "C" int xg_face_area(const double *d_xy, const void *d_faces, int idx_bytes, int64_t n_face,
                            int k, double *d_area, void *stream) {
    if (n_face < 0 || k < 3 || (idx_bytes != 4 && idx_bytes != 8)) { set_error("xg_face_area: invalid argument"); return XG_ERR_INVALID; }
    if (n_face == 0) return XG_OK;
    cudaStream_t s = (cudaStream_t)stream;
    const unsigned G = (unsigned)div_up(n_face, 256);
    if (idx_bytes == 4) k_face_area<int32_t><<<G, 256, 0, s>>>((const double2 *)d_xy, (const int32_t *)d_faces, n_face, k, d_area);
    else k_face_area<int64_t><<<G, 256, 0, s>>>((const double2 *)d_xy, (const int64_t *)d_faces, n_face, k, d_area);
    XG_CHECK_LAUNCH();
    return XG_OK;
}

extern "C" int xg_face_centroid(const double *d_xy, const void *d_faces, int idx_bytes,
                                int64_t n_face, int k, double *d_centroid_xy, void *stream) {
    if (n_face < 0 || k < 3 || (idx_bytes != 4 && idx_bytes != 8)) { set_error("xg_face_centroid: invalid argument"); return XG_ERR_INVALID; }
    if (n_face == 0) return XG_OK;
    cudaStream_t s = (cudaStream_t)stream;
    const unsigned G = (unsigned)div_up(n_face, 256);
    if (idx_bytes == 4) k_face_centroid<int32_t><<<G, 256, 0, s>>>((const double2 *)d_xy, (const int32_t *)d_faces, n_face, k, (double2 *)d_centroid_xy);
    else k_face_centroid<int64_t><<<G, 256, 0, s>>>((const double2 *)d_xy, (const int64_t *)d_faces, n_face, k, (double2 *)d_centroid_xy);
    XG_CHECK_LAUNCH();
    return XG_OK;
}

extern "C" int xg_csr_apply(int method, double percentile, const int64_t *d_indptr,
                            const int32_t *d_indices, const double *d_weights, const double *d_src,
                            int64_t n_extra, int64_t m, int64_t n_rows, double *d_out,
                            int64_t out_ld, void *stream) {
    if (!d_out && n_rows > 0 && n_extra > 0) { set_error("xg_csr_apply: NULL argument"); return XG_ERR_INVALID; }
    OutSet O;
    O.n = 1;
    O.p[0] = d_out;
    return csr_apply_impl(method, percentile, d_indptr, d_indices, d_weights, d_src, n_extra, m, n_rows, O, out_ld, stream);
}

extern "C" int xg_csr_apply_peers(int method, double percentile, const int64_t *d_indptr,
                                  const int32_t *d_indices, const double *d_weights, const double *d_src,
                                  int64_t n_extra, int64_t m, int64_t n_rows, double *const *h_d_outs,
                                  int n_out, int64_t out_ld, void *stream) {
    if (!h_d_outs || n_out < 1 || n_out > MAX_PEERS) {
        set_error("xg_csr_apply_peers: between 1 and %d output arrays", MAX_PEERS);
        return XG_ERR_INVALID;
    }
    OutSet O;
    O.n = n_out;
    for (int i = 0; i < n_out; ++i) {
        if (!h_d_outs[i] && n_rows > 0 && n_extra > 0) { set_error("xg_csr_apply_peers: NULL output"); return XG_ERR_INVALID; }
        O.p[i] = h_d_outs[i];
    }
    return csr_apply_impl(method, percentile, d_indptr, d_indices, d_weights, d_src, n_extra, m, n_rows, O, out_ld, stream);
}
