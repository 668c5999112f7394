// xg_index.cu -- packed, Morton-ordered mesh and wide BVH ("spatial index").
//
// Replaces numba_celltree.CellTree2d.__init__ as called from
// /root/reference/xugrid/ugrid/ugrid2d.py:908-921 (bounding boxes + a bounding
// interval hierarchy built serially on the CPU).  B200 design instead:
//   1. one pass over the faces: vertex gather, bbox, 32-bit Morton key of the
//      bbox centre                                              (k_face_keys)
//   2. device radix sort of (key, face id)                      (cub)
//   3. one pass in sorted order: gather again, force CCW, write the face as a
//      self-contained record (bbox 32 B + vertices k*16 B)      (k_pack)
//      -> later kernels never touch the indexed mesh again; neighbours in
//      memory are neighbours in space, so a tile of queries reads contiguous
//      runs that TMA/coalesced loads can stream
//   4. FANOUT(32)-ary hierarchy over the sorted boxes, one warp per node,
//      level by level: no atomics, no pointers, child j of node p at level l
//      is entry 32*p + j of level l-1                           (k_reduce_level)
#include <cub/device/device_radix_sort.cuh>

#include <mutex>
#include <vector>

#include "xg_common.cuh"
#include "xg_pack.cuh"

namespace xg {

// ---- error plumbing ---------------------------------------------------------
static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int cuda_fail(cudaError_t e, const char *what, const char *file, int line) {
    set_error("CUDA error %s (%s) at %s:%d in %s", cudaGetErrorName(e), cudaGetErrorString(e),
              file, line, what);
    return XG_ERR_CUDA;
}

int num_sms() {
    static std::mutex mu;
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    std::lock_guard<std::mutex> lock(mu);
    if (cached[dev] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached[dev] = n;
    }
    return cached[dev];
}

// ---- block cache (see xg_common.cuh) ----------------------------------------------
struct CachedBlock {
    void *p;
    size_t cap;
    cudaStream_t s;
    int dev;
};
static std::mutex g_cache_mutex;
static std::vector<CachedBlock> g_cache;
static size_t g_cached_bytes = 0;
static const size_t g_cache_limit = (size_t)96 << 30;

static size_t round_size(size_t n) {
    const size_t g = n >= ((size_t)2 << 20) ? ((size_t)2 << 20) : 512;
    return (n + g - 1) / g * g;
}

static thread_local bool t_capturing = false;
void scratch_set_capturing(bool on) { t_capturing = on; }

cudaError_t scratch_alloc(void **p, size_t *cap, size_t bytes, cudaStream_t s) {
    int dev = 0;
    cudaGetDevice(&dev);
    const size_t want = round_size(bytes);
    {
        std::lock_guard<std::mutex> lock(g_cache_mutex);
        int best = -1;
        for (int i = 0; i < (int)g_cache.size(); ++i) {
            const CachedBlock &b = g_cache[i];
            if (b.dev != dev || b.s != s || b.cap < want || b.cap > want + want / 2 + ((size_t)1 << 20)) continue;
            if (best < 0 || b.cap < g_cache[best].cap) best = i;
        }
        if (best >= 0) {
            *p = g_cache[best].p;
            *cap = g_cache[best].cap;
            g_cached_bytes -= g_cache[best].cap;
            g_cache[best] = g_cache.back();
            g_cache.pop_back();
            return cudaSuccess;
        }
    }
    // while a job is being recorded into a CUDA graph nothing may be allocated: the caller gives
    // the capture up, runs directly (which fills the cache) and captures the next time
    if (t_capturing) return cudaErrorStreamCaptureUnsupported;
    *cap = want;
    return cudaMallocAsync(p, want, s);
}

void scratch_free(void *p, size_t cap, cudaStream_t s) {
    int dev = 0;
    cudaGetDevice(&dev);
    {
        std::lock_guard<std::mutex> lock(g_cache_mutex);
        if (g_cached_bytes + cap <= g_cache_limit) {
            g_cache.push_back(CachedBlock{p, cap, s, dev});
            g_cached_bytes += cap;
            return;
        }
    }
    cudaFreeAsync(p, s);
}

static void release_cache() {
    std::lock_guard<std::mutex> lock(g_cache_mutex);
    for (const CachedBlock &b : g_cache) cudaFreeAsync(b.p, b.s);
    g_cache.clear();
    g_cached_bytes = 0;
}

// ---- pinned host slots (128 bytes each) for asynchronous read-backs ---------------------------
static std::mutex g_pin_mutex;
static std::vector<double *> g_pin_free;

double *pinned_slot_acquire() {
    std::lock_guard<std::mutex> lock(g_pin_mutex);
    if (g_pin_free.empty()) {
        void *block = nullptr;
        const int slots = 64;
        if (cudaHostAlloc(&block, 128 * slots, cudaHostAllocDefault) != cudaSuccess) return nullptr;
        for (int i = 0; i < slots; ++i) g_pin_free.push_back(reinterpret_cast<double *>((char *)block + 128 * i));
    }
    double *p = g_pin_free.back();
    g_pin_free.pop_back();
    return p;
}

void pinned_slot_release(double *p) {
    if (!p) return;
    std::lock_guard<std::mutex> lock(g_pin_mutex);
    g_pin_free.push_back(p);
}

int index_host_values(const xg_index *cix) {
    xg_index *ix = const_cast<xg_index *>(cix);
    std::lock_guard<std::mutex> lock(ix->mu);
    if (ix->host_values_valid) return XG_OK;
    if (ix->frame_ready) XG_CUDA(cudaEventSynchronize(ix->frame_ready));
    if (ix->h_frame) {
        for (int i = 0; i < 4; ++i) ix->bounds[i] = ix->h_frame[i];
        ix->default_tolerance = ix->h_frame[4];
    }
    ix->host_values_valid = true;
    return XG_OK;
}

void index_note_use(const xg_index *cix, cudaStream_t s) {
    xg_index *ix = const_cast<xg_index *>(cix);
    if (s == ix->arena_stream) return;
    std::lock_guard<std::mutex> lock(ix->mu);
    for (cudaStream_t f : ix->foreign_streams)
        if (f == s) return;
    ix->foreign_streams.push_back(s);
}

// ---- kernels ------------------------------------------------------------------
__device__ __forceinline__ unsigned long long order_bits(double v) {
    unsigned long long u = (unsigned long long)__double_as_longlong(v);
    return (u >> 63) ? ~u : (u | 0x8000000000000000ull);
}
__host__ __device__ __forceinline__ double unorder_bits(unsigned long long u) {
    u = (u >> 63) ? (u & 0x7fffffffffffffffull) : ~u;
#ifdef __CUDA_ARCH__
    return __longlong_as_double((long long)u);
#else
    double d;
    memcpy(&d, &u, 8);
    return d;
#endif
}

// red[0..3] = ordered bits of xmin, xmax, ymin, ymax over the nodes
__global__ void k_node_bounds(const double2 *__restrict__ xy, int64_t n, unsigned long long *red) {
    double xmin = INFINITY, xmax = -INFINITY, ymin = INFINITY, ymax = -INFINITY;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x) {
        double2 p = xy[i];
        xmin = fmin(xmin, p.x); xmax = fmax(xmax, p.x);
        ymin = fmin(ymin, p.y); ymax = fmax(ymax, p.y);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        xmin = fmin(xmin, __shfl_xor_sync(0xffffffffu, xmin, o));
        xmax = fmax(xmax, __shfl_xor_sync(0xffffffffu, xmax, o));
        ymin = fmin(ymin, __shfl_xor_sync(0xffffffffu, ymin, o));
        ymax = fmax(ymax, __shfl_xor_sync(0xffffffffu, ymax, o));
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMin(red + 0, order_bits(xmin));
        atomicMax(red + 1, order_bits(xmax));
        atomicMin(red + 2, order_bits(ymin));
        atomicMax(red + 3, order_bits(ymax));
    }
}

__global__ void k_init_red(unsigned long long *red) {
    red[0] = ~0ull; red[1] = 0ull; red[2] = ~0ull; red[3] = 0ull; red[4] = 0ull;
}

// frame[0..3]: xmin, xmax, ymin, ymax used to quantise Morton keys (device memory)
__global__ void k_frame_from_red(const unsigned long long *red, double *frame) {
    frame[0] = unorder_bits(red[0]); frame[1] = unorder_bits(red[1]);
    frame[2] = unorder_bits(red[2]); frame[3] = unorder_bits(red[3]);
    frame[4] = sqrt(__longlong_as_double((long long)red[4])) * 1e-12;  // default tolerance
}

// optional cull: only faces whose bbox strictly intersects the clip box (device memory:
// [xmin, xmax, ymin, ymax], e.g. the bounds of a rank's target range reduced on the
// device a moment ago) are kept, appended compactly and counted

// position of every node relative to the clip box of a culled build: bit 0 = at or left of
// xmin, 1 = at or right of xmax, 2 = at or below ymin, 3 = at or above ymax.  The bbox of a
// face misses the box (strictly) exactly when all its nodes share one of the four bits, so
// the pass over ALL faces of the replicated source mesh that every rank makes reads the
// connectivity and k bytes per face; coordinates are gathered for the kept faces only.
__global__ void __launch_bounds__(256)
k_node_outcodes(const double2 *__restrict__ xy, int64_t n_node, const double *__restrict__ clip,
                uint8_t *__restrict__ code) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n_node) return;
    const double2 p = xy[i];
    code[i] = (uint8_t)((p.x <= clip[0] ? 1 : 0) | (p.x >= clip[1] ? 2 : 0) | (p.y <= clip[2] ? 4 : 0) |
                        (p.y >= clip[3] ? 8 : 0));
}

template <typename I, int KMAX>
__global__ void __launch_bounds__(256)
k_face_keys(const double2 *__restrict__ xy, const I *__restrict__ faces, int64_t f_begin,
            int64_t n, int k, const double *__restrict__ frame, uint32_t *__restrict__ keys,
            int32_t *__restrict__ ids, unsigned long long *diag2_max, const double *__restrict__ clip,
            unsigned long long *__restrict__ kept, const int32_t *__restrict__ face_ids,
            const uint8_t *__restrict__ outcode) {
    // face_ids (optional): the faces to key are face_ids[f_begin + i]; what is recorded in
    // `ids` is then the position in that list, not the face number
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    double d2 = 0.0;
    bool in_box = false;
    uint32_t my_key = 0;
    bool culled_by_nodes = false;
    if (i < n && outcode) {
        const I *face = faces + (face_ids ? (int64_t)face_ids[f_begin + i] : f_begin + i) * k;
        unsigned all = 0xfu;
#pragma unroll
        for (int j = 0; j < KMAX; ++j) {
            if (j >= k) break;
            const I idx = face[j];
            if (idx < 0) break;
            all &= outcode[idx];
        }
        culled_by_nodes = all != 0u;
    }
    if (i < n && !culled_by_nodes) {
        const I *face = faces + (face_ids ? (int64_t)face_ids[f_begin + i] : f_begin + i) * k;
        double2 p = xy[face[0]];
        double xmin = p.x, xmax = p.x, ymin = p.y, ymax = p.y;
#pragma unroll
        for (int j = 1; j < KMAX; ++j) {
            if (j >= k) break;
            const I idx = face[j];
            if (idx < 0) break;
            p = xy[idx];
            xmin = fmin(xmin, p.x); xmax = fmax(xmax, p.x);
            ymin = fmin(ymin, p.y); ymax = fmax(ymax, p.y);
        }
        // a culled build looks at every face of the mesh but keeps few (1/8 on 8 ranks):
        // key and diagonal only for the kept ones
        in_box = !clip || (xmin < clip[1] && clip[0] < xmax && ymin < clip[3] && clip[2] < ymax);
        if (in_box) {
            const double fx0 = frame[0], fx1 = frame[1], fy0 = frame[2], fy1 = frame[3];
            const double cx = 0.5 * (xmin + xmax), cy = 0.5 * (ymin + ymax);
            const double sx = fx1 > fx0 ? 65536.0 / (fx1 - fx0) : 0.0;
            const double sy = fy1 > fy0 ? 65536.0 / (fy1 - fy0) : 0.0;
            double qx = (cx - fx0) * sx, qy = (cy - fy0) * sy;
            qx = fmin(fmax(qx, 0.0), 65534.0);
            qy = fmin(fmax(qy, 0.0), 65534.0);
            my_key = curve_key16((uint32_t)qx, (uint32_t)qy);
            if (!clip) {
                keys[i] = my_key;
                ids[i] = (int32_t)(f_begin + i);
            }
            const double dx = xmax - xmin, dy = ymax - ymin;
            d2 = dx * dx + dy * dy;
        }
    }
    if (clip) {
        // culled build: kept faces are appended compactly (one atomic per warp)
        const unsigned m = __ballot_sync(0xffffffffu, in_box);
        if (m) {
            const int lane = threadIdx.x & 31;
            unsigned long long base = 0;
            if (lane == __ffs(m) - 1) base = atomicAdd(kept, (unsigned long long)__popc(m));
            base = __shfl_sync(0xffffffffu, base, __ffs(m) - 1);
            if (in_box) {
                const unsigned long long pos = base + __popc(m & ((1u << lane) - 1u));
                keys[pos] = my_key;
                ids[pos] = (int32_t)(f_begin + i);
            }
        }
    }
    if (diag2_max) {
        // block-wide max first: one global atomic per block (one per warp on a single
        // address serialised the kernel: 0.27 ms vs 0.13 ms without the reduction)
        __shared__ double s_d2[8];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) d2 = fmax(d2, __shfl_xor_sync(0xffffffffu, d2, o));
        if ((threadIdx.x & 31) == 0) s_d2[threadIdx.x >> 5] = d2;
        __syncthreads();
        if (threadIdx.x == 0) {
#pragma unroll
            for (int i = 1; i < 8; ++i) d2 = fmax(d2, s_d2[i]);
            if (d2 > 0.0) atomicMax(diag2_max, (unsigned long long)__double_as_longlong(d2));
        }
    }
}

// red[0..3] = ordered bits of the bounds of the vertices of faces [f_begin, f_begin + n)
template <typename I>
__global__ void __launch_bounds__(256)
k_face_range_bounds(const double2 *__restrict__ xy, const I *__restrict__ faces, int64_t f_begin,
                    int64_t n, int k, unsigned long long *red, const int32_t *__restrict__ face_ids) {
    double xmin = INFINITY, xmax = -INFINITY, ymin = INFINITY, ymax = -INFINITY;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x) {
        const I *face = faces + (face_ids ? (int64_t)face_ids[f_begin + i] : f_begin + i) * k;
        for (int j = 0; j < k; ++j) {
            const I idx = face[j];
            if (idx < 0) break;
            const double2 p = xy[idx];
            xmin = fmin(xmin, p.x); xmax = fmax(xmax, p.x);
            ymin = fmin(ymin, p.y); ymax = fmax(ymax, p.y);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        xmin = fmin(xmin, __shfl_xor_sync(0xffffffffu, xmin, o));
        xmax = fmax(xmax, __shfl_xor_sync(0xffffffffu, xmax, o));
        ymin = fmin(ymin, __shfl_xor_sync(0xffffffffu, ymin, o));
        ymax = fmax(ymax, __shfl_xor_sync(0xffffffffu, ymax, o));
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMin(red + 0, order_bits(xmin));
        atomicMax(red + 1, order_bits(xmax));
        atomicMin(red + 2, order_bits(ymin));
        atomicMax(red + 3, order_bits(ymax));
    }
}

// sorted position i <- face perm[i]: CCW-normalised vertices, bbox, vertex count
template <typename I, int KMAX>
__global__ void __launch_bounds__(256)
k_pack(const double2 *__restrict__ xy, const I *__restrict__ faces, int k,
       const int32_t *__restrict__ perm, int64_t n, Box *__restrict__ box,
       double2 *__restrict__ poly, uint8_t *__restrict__ nvert, const int32_t *__restrict__ face_ids) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const I *face = faces + (int64_t)(face_ids ? face_ids[perm[i]] : perm[i]) * k;
    double2 v[KMAX];
    int len = k;
#pragma unroll
    for (int j = 0; j < KMAX; ++j) {
        if (j < k) {
            const I idx = face[j];
            if (idx < 0 && j >= 3 && len == k) len = j;  // first fill after a triangle ends it
            v[j] = idx >= 0 ? xy[idx] : make_double2(0.0, 0.0);
        }
    }
    // bounding box over the polygon (stops at the first fill value)
    double xmin = v[0].x, xmax = v[0].x, ymin = v[0].y, ymax = v[0].y;
#pragma unroll
    for (int j = 1; j < KMAX; ++j)
        if (j < len) {
            xmin = fmin(xmin, v[j].x); xmax = fmax(xmax, v[j].x);
            ymin = fmin(ymin, v[j].y); ymax = fmax(ymax, v[j].y);
        }
    // orientation: first non-collinear consecutive triple decides
    bool flip = false;
    {
        double2 a = v[0], b = v[0];
#pragma unroll
        for (int j = 0; j < KMAX; ++j) {
            if (j == len - 2) a = v[j];
            if (j == len - 1) b = v[j];
        }
        bool decided = false;
#pragma unroll
        for (int j = 0; j < KMAX; ++j)
            if (j < len && !decided) {
                const double2 c = v[j];
                const double prod = (b.x - a.x) * (c.y - a.y) - (b.y - a.y) * (c.x - a.x);
                if (prod == 0.0) { a = b; b = c; }
                else { flip = prod < 0.0; decided = true; }
            }
    }
    double2 *dst = poly + i * k;
#pragma unroll
    for (int j = 0; j < KMAX; ++j)
        if (j < k) {
            double2 o;
            if (j < len) {
                o = v[j];
                if (flip) {
#pragma unroll
                    for (int q = 0; q < KMAX; ++q)
                        if (q == len - 1 - j) o = v[q];
                }
            } else {
                o = make_double2(0.0, 0.0);
            }
            dst[j] = o;
        }
    Box bx;
    bx.xmin = xmin; bx.xmax = xmax; bx.ymin = ymin; bx.ymax = ymax;
    box[i] = bx;
    nvert[i] = (uint8_t)(len | (flip ? 0x80 : 0));
}

// one warp per parent: union of up to FANOUT child boxes
__global__ void __launch_bounds__(256)
k_reduce_level(const Box *__restrict__ child, int64_t n_child, Box *__restrict__ parent,
               int64_t n_parent) {
    const int64_t w = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (w >= n_parent) return;
    const int64_t c = w * FANOUT + lane;
    double xmin = INFINITY, xmax = -INFINITY, ymin = INFINITY, ymax = -INFINITY;
    if (c < n_child) {
        const Box b = child[c];
        xmin = b.xmin; xmax = b.xmax; ymin = b.ymin; ymax = b.ymax;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        xmin = fmin(xmin, __shfl_xor_sync(0xffffffffu, xmin, o));
        xmax = fmax(xmax, __shfl_xor_sync(0xffffffffu, xmax, o));
        ymin = fmin(ymin, __shfl_xor_sync(0xffffffffu, ymin, o));
        ymax = fmax(ymax, __shfl_xor_sync(0xffffffffu, ymax, o));
    }
    if (lane == 0) {
        Box b;
        b.xmin = xmin; b.xmax = xmax; b.ymin = ymin; b.ymax = ymax;
        parent[w] = b;
    }
}

// ---- host side ------------------------------------------------------------------
template <typename I>
static int launch_face_keys(const double2 *xy, const I *faces, int64_t f_begin, int64_t n, int k,
                            const double *frame, uint32_t *keys, int32_t *ids,
                            unsigned long long *diag2, const double *clip, unsigned long long *kept,
                            const int32_t *face_ids, cudaStream_t s, const uint8_t *outcode = nullptr) {
    if (n == 0) return XG_OK;
    const int B = 256;
    const unsigned G = (unsigned)div_up(n, B);
    if (k <= 3) k_face_keys<I, 3><<<G, B, 0, s>>>(xy, faces, f_begin, n, k, frame, keys, ids, diag2, clip, kept, face_ids, outcode);
    else if (k <= 4) k_face_keys<I, 4><<<G, B, 0, s>>>(xy, faces, f_begin, n, k, frame, keys, ids, diag2, clip, kept, face_ids, outcode);
    else k_face_keys<I, XG_MAX_N_VERTEX><<<G, B, 0, s>>>(xy, faces, f_begin, n, k, frame, keys, ids, diag2, clip, kept, face_ids, outcode);
    XG_CHECK_LAUNCH();
    return XG_OK;
}

template <typename I>
static int launch_pack(const double2 *xy, const I *faces, int k, const int32_t *perm, int64_t n,
                       Box *box, double2 *poly, uint8_t *nvert, const int32_t *face_ids, cudaStream_t s) {
    if (n == 0) return XG_OK;
    const int B = 256;
    const unsigned G = (unsigned)div_up(n, B);
    if (k <= 3) k_pack<I, 3><<<G, B, 0, s>>>(xy, faces, k, perm, n, box, poly, nvert, face_ids);
    else if (k <= 4) k_pack<I, 4><<<G, B, 0, s>>>(xy, faces, k, perm, n, box, poly, nvert, face_ids);
    else if (k <= 8) k_pack<I, 8><<<G, B, 0, s>>>(xy, faces, k, perm, n, box, poly, nvert, face_ids);
    else if (k <= 16) k_pack<I, 16><<<G, B, 0, s>>>(xy, faces, k, perm, n, box, poly, nvert, face_ids);
    else k_pack<I, XG_MAX_N_VERTEX><<<G, B, 0, s>>>(xy, faces, k, perm, n, box, poly, nvert, face_ids);
    XG_CHECK_LAUNCH();
    return XG_OK;
}

static void ensure_pool_config() {
    // once per device (a process may drive several), under a lock
    static std::mutex mu;
    static std::vector<int> configured;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return;
    std::lock_guard<std::mutex> lock(mu);
    for (int d : configured)
        if (d == dev) return;
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
        uint64_t thr = UINT64_MAX;  // keep freed blocks cached: repeated builds reuse them
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
    }
    configured.push_back(dev);
}

size_t pack_sort_tmp_bytes(int64_t n) {
    size_t tb = 0;
    if (n <= 0) return 16;
    // (the temporary storage of the sort does not depend on the bit range)
    cub::DeviceRadixSort::SortPairs(nullptr, tb, (const uint32_t *)nullptr, (uint32_t *)nullptr, (const int32_t *)nullptr,
                                    (int32_t *)nullptr, n, 0, 32, (cudaStream_t)0);
    return tb ? tb : 16;
}

int pack_mesh(const double *d_xy, const void *d_faces, int idx_bytes, int64_t f_begin,
              int64_t f_end, int k, const double *d_frame, unsigned long long *d_diag2,
              PackedMesh *out, cudaStream_t s, const double *d_clip_box, unsigned long long *d_kept,
              const int32_t *d_face_ids, int64_t n_node, cudaStream_t alloc_stream, const PackScratch *scratch) {
    // alloc_stream: the stream the scratch blocks are ordered on (differs from s only while s is
    // a capture stream: the recorded work will run on alloc_stream)
    cudaStream_t as = alloc_stream ? alloc_stream : s;
    ensure_pool_config();
    const int64_t n = f_end - f_begin;
    out->n = n;
    out->k = k;
    const double *clip = (d_clip_box && d_kept) ? d_clip_box : nullptr;
    if (clip) XG_CUDA(cudaMemsetAsync(d_kept, 0, sizeof(unsigned long long), s));
    const double2 *xy = reinterpret_cast<const double2 *>(d_xy);
    DevBuf keys_in, keys_out, ids_in, tmp, outcode;
    if (clip && n_node > 0) {
        XG_CUDA(outcode.alloc((size_t)n_node, as));
        k_node_outcodes<<<(unsigned)div_up(n_node, 256), 256, 0, s>>>(xy, n_node, clip, outcode.as<uint8_t>());
        XG_CHECK_LAUNCH();
    }
    uint32_t *k_in, *k_out;
    int32_t *i_in;
    if (scratch) {
        k_in = scratch->keys_in; k_out = scratch->keys_out; i_in = scratch->ids_in;
    } else {
        XG_CUDA(keys_in.alloc(sizeof(uint32_t) * (size_t)n, as));
        XG_CUDA(keys_out.alloc(sizeof(uint32_t) * (size_t)n, as));
        XG_CUDA(ids_in.alloc(sizeof(int32_t) * (size_t)n, as));
        k_in = keys_in.as<uint32_t>(); k_out = keys_out.as<uint32_t>(); i_in = ids_in.as<int32_t>();
    }
    int rc;
    if (idx_bytes == 4)
        rc = launch_face_keys<int32_t>(xy, (const int32_t *)d_faces, f_begin, n, k, d_frame,
                                       k_in, i_in, d_diag2, clip, d_kept,
                                       d_face_ids, s, outcode.as<uint8_t>());
    else
        rc = launch_face_keys<int64_t>(xy, (const int64_t *)d_faces, f_begin, n, k, d_frame,
                                       k_in, i_in, d_diag2, clip, d_kept,
                                       d_face_ids, s, outcode.as<uint8_t>());
    if (rc) return rc;
    int64_t n_pack = n;
    if (clip) {
        // the one host round trip of a culled build: the sort below is sized on the host
        double *slot = pinned_slot_acquire();
        if (!slot) { set_error("pack_mesh: no pinned host memory"); return XG_ERR_NOMEM; }
        cudaError_t e = cudaMemcpyAsync(slot, d_kept, sizeof(unsigned long long), cudaMemcpyDeviceToHost, s);
        if (e == cudaSuccess) e = cudaStreamSynchronize(s);
        unsigned long long h_kept = 0;
        memcpy(&h_kept, slot, sizeof(h_kept));
        pinned_slot_release(slot);
        if (e != cudaSuccess) return cuda_fail(e, "kept count", __FILE__, __LINE__);
        n_pack = (int64_t)h_kept;
        out->n = n_pack;
    }
    if (n_pack > 0) {
        // only the leading Morton bits that can separate n faces are sorted (8-bit radix
        // digits: 10 M faces -> 24 bits -> 3 passes instead of 4); faces that share those
        // bits keep their input order (the sort is stable)
        int bits = 16;
        while (bits < 32 && ((int64_t)1 << bits) < n_pack) bits += 8;
        const int begin_bit = 32 - bits;
        size_t tb = 0;
        XG_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tb, k_in, k_out, i_in, out->perm, n_pack, begin_bit, 32, s));
        void *sort_tmp;
        if (scratch && scratch->sort_tmp_bytes >= tb) {
            sort_tmp = scratch->sort_tmp;
        } else {
            XG_CUDA(tmp.alloc(tb, as));
            sort_tmp = tmp.p;
        }
        XG_CUDA(cub::DeviceRadixSort::SortPairs(sort_tmp, tb, k_in, k_out, i_in, out->perm, n_pack, begin_bit, 32, s));
    }
    if (idx_bytes == 4)
        rc = launch_pack<int32_t>(xy, (const int32_t *)d_faces, k, out->perm, n_pack, out->box,
                                  reinterpret_cast<double2 *>(out->poly), out->nvert, d_face_ids, s);
    else
        rc = launch_pack<int64_t>(xy, (const int64_t *)d_faces, k, out->perm, n_pack, out->box,
                                  reinterpret_cast<double2 *>(out->poly), out->nvert, d_face_ids, s);
    return rc;
}

}  // namespace xg

using namespace xg;

extern "C" int xg_abi_version(void) { return XG_ABI_VERSION; }
extern "C" const char *xg_last_error_string(void) { return xg::g_err; }

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// a range of target faces whose bounds cull the source mesh (xg_index_create_for_targets)
struct TargetRange {
    const double *d_xy;
    const void *d_faces;
    int idx_bytes;
    int64_t begin, end;
    int k;
    const int32_t *d_ids;  // optional: the range indexes this list of faces
};

static int launch_range_bounds(const TargetRange &t, unsigned long long *d_red, cudaStream_t s) {
    const int64_t n = t.end - t.begin;
    if (n <= 0) return XG_OK;
    if (t.idx_bytes == 4)
        k_face_range_bounds<int32_t><<<num_sms() * 4, 256, 0, s>>>((const double2 *)t.d_xy, (const int32_t *)t.d_faces,
                                                               t.begin, n, t.k, d_red, t.d_ids);
    else
        k_face_range_bounds<int64_t><<<num_sms() * 4, 256, 0, s>>>((const double2 *)t.d_xy, (const int64_t *)t.d_faces,
                                                               t.begin, n, t.k, d_red, t.d_ids);
    XG_CHECK_LAUNCH();
    return XG_OK;
}

// h_clip_box (host) or targets (bounds reduced on the device, no host round trip) restrict
// the index to the source faces whose bbox strictly meets the box; both NULL: all faces
static int index_create_impl(const double *d_xy, int64_t n_node, const void *d_faces, int idx_bytes,
                             int64_t n_face, int k, const double *h_clip_box, const TargetRange *targets,
                             void *stream, xg_index **out) {
    if (!out) { set_error("xg_index_create: out is NULL"); return XG_ERR_INVALID; }
    *out = nullptr;
    if (n_face < 0 || n_node < 0 || (idx_bytes != 4 && idx_bytes != 8) ||
        (n_face > 0 && (!d_xy || !d_faces))) {
        set_error("xg_index_create: invalid argument");
        return XG_ERR_INVALID;
    }
    if (k < 3 || k > XG_MAX_N_VERTEX) {
        set_error("xg_index_create: faces must have 3..%d vertices, got %d", XG_MAX_N_VERTEX, k);
        return XG_ERR_UNSUPPORTED;
    }
    if (n_face >= (int64_t)1 << 31) {
        set_error("xg_index_create: more than 2^31 faces");
        return XG_ERR_UNSUPPORTED;
    }
    cudaStream_t s = (cudaStream_t)stream;
    ensure_pool_config();
    xg_index *ix = new xg_index();
    ix->n_face = n_face; ix->n_face_mesh = n_face; ix->n_node = n_node; ix->k = k;
    const bool culled = h_clip_box || targets;

    // level sizes
    int64_t level_n[MAX_LEVELS];
    int n_levels = 0;
    int64_t cur = n_face;
    while (n_face > 0 && n_levels < MAX_LEVELS) {
        cur = div_up(cur, FANOUT);
        level_n[n_levels++] = cur;
        if (cur <= FANOUT) break;
    }
    ix->n_levels = n_levels;

    // one arena: boxes | level boxes | poly | perm | nvert | frame(8 doubles) | red(8 u64) |
    //            clip box(4 doubles) + its reduction cells(8 u64)
    size_t off = 0;
    const size_t off_box = off; off = align_up(off + sizeof(Box) * (size_t)n_face, 256);
    size_t off_lvl[MAX_LEVELS];
    for (int l = 0; l < n_levels; ++l) {
        off_lvl[l] = off;
        off = align_up(off + sizeof(Box) * (size_t)level_n[l], 256);
    }
    const size_t off_poly = off; off = align_up(off + sizeof(double) * 2 * (size_t)k * (size_t)n_face, 256);
    const size_t off_perm = off; off = align_up(off + sizeof(int32_t) * (size_t)n_face, 256);
    const size_t off_nv = off; off = align_up(off + (size_t)n_face, 256);
    const size_t off_frame = off; off = align_up(off + sizeof(double) * 8, 256);
    const size_t off_red = off; off = align_up(off + sizeof(unsigned long long) * 8, 256);
    const size_t off_clip = off; off = align_up(off + sizeof(double) * 8, 256);
    const size_t off_clip_red = off; off = align_up(off + sizeof(unsigned long long) * 8, 256);
    ix->bytes = off;
    ix->arena_stream = s;
    cudaError_t e = scratch_alloc(&ix->arena, &ix->arena_cap, off, s);
    if (e != cudaSuccess) { delete ix; return cuda_fail(e, "scratch_alloc(index)", __FILE__, __LINE__); }
    char *base = (char *)ix->arena;
    ix->leaf_box = (Box *)(base + off_box);
    for (int l = 0; l < n_levels; ++l) { ix->level_box[l] = (Box *)(base + off_lvl[l]); ix->level_n[l] = level_n[l]; }
    ix->poly = (double *)(base + off_poly);
    ix->perm = (int32_t *)(base + off_perm);
    ix->nvert = (uint8_t *)(base + off_nv);
    double *d_frame = (double *)(base + off_frame);
    ix->d_frame = d_frame;
    unsigned long long *d_red = (unsigned long long *)(base + off_red);
    double *d_clip = (double *)(base + off_clip);
    unsigned long long *d_clip_red = (unsigned long long *)(base + off_clip_red);

    int rc = XG_OK;
    do {
        k_init_red<<<1, 1, 0, s>>>(d_red);
        if (n_node > 0)
            k_node_bounds<<<num_sms() * 4, 256, 0, s>>>((const double2 *)d_xy, n_node, d_red);
        k_frame_from_red<<<1, 1, 0, s>>>(d_red, d_frame);
        if (cudaGetLastError() != cudaSuccess) { rc = XG_ERR_CUDA; set_error("index: launch failed"); break; }
        if (h_clip_box) {
            // (a pageable source is staged by the runtime before the call returns)
            e = cudaMemcpyAsync(d_clip, h_clip_box, sizeof(double) * 4, cudaMemcpyHostToDevice, s);
            if (e != cudaSuccess) { rc = cuda_fail(e, "clip box upload", __FILE__, __LINE__); break; }
        } else if (targets) {
            k_init_red<<<1, 1, 0, s>>>(d_clip_red);
            rc = launch_range_bounds(*targets, d_clip_red, s);
            if (rc) break;
            k_frame_from_red<<<1, 1, 0, s>>>(d_clip_red, d_clip);
        }
        PackedMesh pm;
        pm.perm = ix->perm; pm.box = ix->leaf_box; pm.poly = ix->poly; pm.nvert = ix->nvert;
        rc = pack_mesh(d_xy, d_faces, idx_bytes, 0, n_face, k, d_frame, d_red + 4, &pm, s,
                       culled ? d_clip : nullptr, d_red + 5, nullptr, culled ? n_node : 0);
        if (rc) break;
        if (culled) {
            // only the faces that can meet the box are indexed
            ix->n_face = pm.n;
            n_levels = 0;
            int64_t c = pm.n;
            while (pm.n > 0 && n_levels < MAX_LEVELS) {
                c = div_up(c, FANOUT);
                level_n[n_levels] = c;
                ix->level_n[n_levels] = c;
                ++n_levels;
                if (c <= FANOUT) break;
            }
            ix->n_levels = n_levels;
        }
        // tolerance needs the max diagonal, which k_face_keys has now reduced
        k_frame_from_red<<<1, 1, 0, s>>>(d_red, d_frame);
        const Box *child = ix->leaf_box;
        int64_t n_child = ix->n_face;
        for (int l = 0; l < n_levels; ++l) {
            const int64_t np = level_n[l];
            k_reduce_level<<<(unsigned)div_up(np * 32, 256), 256, 0, s>>>(child, n_child, ix->level_box[l], np);
            child = ix->level_box[l];
            n_child = np;
        }
        if (cudaGetLastError() != cudaSuccess) { rc = XG_ERR_CUDA; set_error("index: launch failed"); break; }
        // bounds + default tolerance travel to a pinned host slot behind the build; nobody
        // waits for them here (index_host_values does, the first time they are asked for)
        ix->h_frame = pinned_slot_acquire();
        if (!ix->h_frame) { rc = XG_ERR_NOMEM; set_error("index: no pinned host memory"); break; }
        e = cudaMemcpyAsync(ix->h_frame, d_frame, sizeof(double) * 5, cudaMemcpyDeviceToHost, s);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ix->frame_ready, cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventRecord(ix->frame_ready, s);
        if (e != cudaSuccess) { rc = cuda_fail(e, "index finalize", __FILE__, __LINE__); break; }
    } while (0);
    if (rc) {
        if (ix->frame_ready) cudaEventDestroy(ix->frame_ready);
        pinned_slot_release(ix->h_frame);
        scratch_free(ix->arena, ix->arena_cap, s);
        delete ix;
        return rc;
    }
    *out = ix;
    return XG_OK;
}

extern "C" int xg_index_create(const double *d_xy, int64_t n_node, const void *d_faces,
                               int idx_bytes, int64_t n_face, int k, void *stream,
                               xg_index **out) {
    return index_create_impl(d_xy, n_node, d_faces, idx_bytes, n_face, k, nullptr, nullptr, stream, out);
}

extern "C" int xg_index_create_clipped(const double *d_xy, int64_t n_node, const void *d_faces,
                                       int idx_bytes, int64_t n_face, int k,
                                       const double *h_clip_box, void *stream, xg_index **out) {
    return index_create_impl(d_xy, n_node, d_faces, idx_bytes, n_face, k, h_clip_box, nullptr, stream, out);
}

extern "C" int xg_index_create_for_targets(const double *d_xy, int64_t n_node, const void *d_faces,
                                           int idx_bytes, int64_t n_face, int k, const double *d_t_xy,
                                           const void *d_t_faces, int t_idx_bytes, int64_t t_begin,
                                           int64_t t_end, int k_t, void *stream, xg_index **out) {
    if (!d_t_xy || !d_t_faces || t_begin < 0 || t_end < t_begin || k_t < 3 ||
        (t_idx_bytes != 4 && t_idx_bytes != 8)) {
        set_error("xg_index_create_for_targets: invalid target range");
        return XG_ERR_INVALID;
    }
    TargetRange tr;
    tr.d_xy = d_t_xy; tr.d_faces = d_t_faces; tr.idx_bytes = t_idx_bytes;
    tr.begin = t_begin; tr.end = t_end; tr.k = k_t; tr.d_ids = nullptr;
    return index_create_impl(d_xy, n_node, d_faces, idx_bytes, n_face, k, nullptr, &tr, stream, out);
}

extern "C" int xg_index_create_for_target_ids(const double *d_xy, int64_t n_node, const void *d_faces,
                                              int idx_bytes, int64_t n_face, int k, const double *d_t_xy,
                                              const void *d_t_faces, int t_idx_bytes, const int32_t *d_t_ids,
                                              int64_t n_ids, int k_t, void *stream, xg_index **out) {
    if (!d_t_xy || !d_t_faces || n_ids < 0 || (n_ids > 0 && !d_t_ids) || k_t < 3 ||
        (t_idx_bytes != 4 && t_idx_bytes != 8)) {
        set_error("xg_index_create_for_target_ids: invalid target list");
        return XG_ERR_INVALID;
    }
    TargetRange tr;
    tr.d_xy = d_t_xy; tr.d_faces = d_t_faces; tr.idx_bytes = t_idx_bytes;
    tr.begin = 0; tr.end = n_ids; tr.k = k_t; tr.d_ids = d_t_ids;
    return index_create_impl(d_xy, n_node, d_faces, idx_bytes, n_face, k, nullptr, &tr, stream, out);
}

// ---- partition of a mesh's faces along the space-filling curve ---------------------------
// A rank that owns a contiguous RANGE of the caller's face numbers owns a compact patch
// only if those numbers mean something in space.  Ownership by curve position does not
// depend on the numbering: every rank keys all faces (same frame = bounds of the nodes,
// so all ranks agree), counts them in 65 536 bins of the leading 16 key bits, cuts the
// cumulative count at multiples of n / world and takes the faces of its own bins.
namespace xg {

constexpr int SHARD_BINS = 1 << 16;

__global__ void __launch_bounds__(256)
k_key_histogram(const uint32_t *__restrict__ keys, int64_t n, unsigned int *__restrict__ hist) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) atomicAdd(hist + (keys[i] >> 16), 1u);
}

// one block: bin range [lo, hi) of `rank`; cut[r] = first bin whose cumulative count reaches
// round(r * n / world)
__global__ void __launch_bounds__(1024)
k_shard_bins(const unsigned int *__restrict__ hist, int64_t n, int rank, int world, int *__restrict__ lohi) {
    __shared__ unsigned long long s_sum[1024];
    __shared__ int s_lo, s_hi;
    const int per = SHARD_BINS / 1024;  // consecutive bins per thread
    unsigned long long local = 0;
    for (int j = 0; j < per; ++j) local += hist[threadIdx.x * per + j];
    s_sum[threadIdx.x] = local;
    if (threadIdx.x == 0) { s_lo = SHARD_BINS; s_hi = SHARD_BINS; }
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {  // inclusive scan (Hillis-Steele)
        const unsigned long long v = threadIdx.x >= o ? s_sum[threadIdx.x - o] : 0ull;
        __syncthreads();
        s_sum[threadIdx.x] += v;
        __syncthreads();
    }
    const unsigned long long want_lo = (unsigned long long)((__int128)n * rank / world);
    const unsigned long long want_hi = (unsigned long long)((__int128)n * (rank + 1) / world);
    unsigned long long cum = s_sum[threadIdx.x] - local;  // faces in bins before this thread's
    for (int j = 0; j < per; ++j) {
        const int bin = threadIdx.x * per + j;
        const unsigned long long next = cum + hist[bin];
        // the bin where the cumulative count crosses the cut starts the next shard
        if (cum < want_lo && next >= want_lo) atomicMin(&s_lo, bin + 1);
        if (cum < want_hi && next >= want_hi) atomicMin(&s_hi, bin + 1);
        cum = next;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        lohi[0] = rank == 0 ? 0 : s_lo;
        lohi[1] = rank == world - 1 ? SHARD_BINS : s_hi;
    }
}

__global__ void __launch_bounds__(256)
k_select_bins(const uint32_t *__restrict__ keys, int64_t n, const int *__restrict__ lohi,
              int32_t *__restrict__ ids, unsigned long long *__restrict__ count) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int lo = lohi[0], hi = lohi[1];
    bool mine = false;
    if (i < n) {
        const int bin = (int)(keys[i] >> 16);
        mine = bin >= lo && bin < hi;
    }
    const unsigned m = __ballot_sync(0xffffffffu, mine);
    if (m) {
        const int lane = threadIdx.x & 31;
        unsigned long long base = 0;
        if (lane == __ffs(m) - 1) base = atomicAdd(count, (unsigned long long)__popc(m));
        base = __shfl_sync(0xffffffffu, base, __ffs(m) - 1);
        if (mine) ids[base + __popc(m & ((1u << lane) - 1u))] = (int32_t)i;
    }
}

}  // namespace xg

extern "C" int xg_target_shard(const double *d_xy, int64_t n_node, const void *d_faces, int idx_bytes,
                               int64_t n_face, int k, int rank, int world, int32_t *d_ids,
                               int64_t *h_count, void *stream) {
    if (!h_count || n_face < 0 || n_node < 0 || k < 3 || (idx_bytes != 4 && idx_bytes != 8) || world < 1 ||
        rank < 0 || rank >= world || (n_face > 0 && (!d_xy || !d_faces || !d_ids))) {
        set_error("xg_target_shard: invalid argument");
        return XG_ERR_INVALID;
    }
    *h_count = 0;
    if (n_face == 0) return XG_OK;
    if (n_face >= (int64_t)1 << 31) { set_error("xg_target_shard: more than 2^31 faces"); return XG_ERR_UNSUPPORTED; }
    cudaStream_t s = (cudaStream_t)stream;
    DevBuf keys, ids_unused, hist, small;
    XG_CUDA(keys.alloc(sizeof(uint32_t) * (size_t)n_face, s));
    XG_CUDA(ids_unused.alloc(sizeof(int32_t) * (size_t)n_face, s));
    XG_CUDA(hist.alloc(sizeof(unsigned int) * SHARD_BINS, s));
    // small: red (8 u64) | frame (8 doubles) | lohi (2 ints) | count (u64)
    XG_CUDA(small.alloc(256, s));
    unsigned long long *d_red = small.as<unsigned long long>();
    double *d_frame = reinterpret_cast<double *>(d_red + 8);
    int *d_lohi = reinterpret_cast<int *>(d_frame + 8);
    unsigned long long *d_count = reinterpret_cast<unsigned long long *>(d_lohi + 2);
    XG_CUDA(cudaMemsetAsync(hist.p, 0, sizeof(unsigned int) * SHARD_BINS, s));
    XG_CUDA(cudaMemsetAsync(d_count, 0, sizeof(unsigned long long), s));
    k_init_red<<<1, 1, 0, s>>>(d_red);
    if (n_node > 0) k_node_bounds<<<num_sms() * 4, 256, 0, s>>>((const double2 *)d_xy, n_node, d_red);
    k_frame_from_red<<<1, 1, 0, s>>>(d_red, d_frame);
    XG_CHECK_LAUNCH();
    int rc;
    if (idx_bytes == 4)
        rc = launch_face_keys<int32_t>((const double2 *)d_xy, (const int32_t *)d_faces, 0, n_face, k, d_frame,
                                       keys.as<uint32_t>(), ids_unused.as<int32_t>(), nullptr, nullptr, nullptr,
                                       nullptr, s);
    else
        rc = launch_face_keys<int64_t>((const double2 *)d_xy, (const int64_t *)d_faces, 0, n_face, k, d_frame,
                                       keys.as<uint32_t>(), ids_unused.as<int32_t>(), nullptr, nullptr, nullptr,
                                       nullptr, s);
    if (rc) return rc;
    k_key_histogram<<<(unsigned)div_up(n_face, 256), 256, 0, s>>>(keys.as<uint32_t>(), n_face, hist.as<unsigned int>());
    k_shard_bins<<<1, 1024, 0, s>>>(hist.as<unsigned int>(), n_face, rank, world, d_lohi);
    k_select_bins<<<(unsigned)div_up(n_face, 256), 256, 0, s>>>(keys.as<uint32_t>(), n_face, d_lohi, d_ids, d_count);
    XG_CHECK_LAUNCH();
    double *slot = pinned_slot_acquire();
    if (!slot) { set_error("xg_target_shard: no pinned host memory"); return XG_ERR_NOMEM; }
    cudaError_t e = cudaMemcpyAsync(slot, d_count, sizeof(unsigned long long), cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    unsigned long long h = 0;
    memcpy(&h, slot, sizeof(h));
    pinned_slot_release(slot);
    if (e != cudaSuccess) return cuda_fail(e, "xg_target_shard", __FILE__, __LINE__);
    *h_count = (int64_t)h;
    return XG_OK;
}

extern "C" int xg_face_range_bounds(const double *d_xy, const void *d_faces, int idx_bytes,
                                    int64_t f_begin, int64_t f_end, int k, double h_box[4],
                                    void *stream) {
    if (!h_box || f_begin < 0 || f_end < f_begin || k < 3 || (idx_bytes != 4 && idx_bytes != 8)) {
        set_error("xg_face_range_bounds: invalid argument");
        return XG_ERR_INVALID;
    }
    cudaStream_t s = (cudaStream_t)stream;
    DevBuf red, frame;
    XG_CUDA(red.alloc(sizeof(unsigned long long) * 8, s));
    XG_CUDA(frame.alloc(sizeof(double) * 8, s));
    k_init_red<<<1, 1, 0, s>>>(red.as<unsigned long long>());
    const int64_t n = f_end - f_begin;
    if (n > 0) {
        if (idx_bytes == 4)
            k_face_range_bounds<int32_t><<<num_sms() * 4, 256, 0, s>>>((const double2 *)d_xy, (const int32_t *)d_faces,
                                                                   f_begin, n, k, red.as<unsigned long long>(), nullptr);
        else
            k_face_range_bounds<int64_t><<<num_sms() * 4, 256, 0, s>>>((const double2 *)d_xy, (const int64_t *)d_faces,
                                                                   f_begin, n, k, red.as<unsigned long long>(), nullptr);
    }
    k_frame_from_red<<<1, 1, 0, s>>>(red.as<unsigned long long>(), frame.as<double>());
    XG_CHECK_LAUNCH();
    XG_CUDA(cudaMemcpyAsync(h_box, frame.p, sizeof(double) * 4, cudaMemcpyDeviceToHost, s));
    XG_CUDA(cudaStreamSynchronize(s));
    return XG_OK;
}

extern "C" int xg_index_destroy(xg_index *ix) {
    if (!ix) return XG_OK;
    if (ix->arena) {
        // queries may have been enqueued on other streams than the creation stream: the
        // arena goes back to the block cache of the creation stream, ordered behind the
        // work of those streams by events (no device-wide synchronisation)
        for (cudaStream_t f : ix->foreign_streams) {
            cudaEvent_t ev;
            if (cudaEventCreateWithFlags(&ev, cudaEventDisableTiming) != cudaSuccess ||
                cudaEventRecord(ev, f) != cudaSuccess ||
                cudaStreamWaitEvent(ix->arena_stream, ev, 0) != cudaSuccess) {
                cudaDeviceSynchronize();  // (cannot order it: fall back to the big hammer)
            }
            cudaEventDestroy(ev);
        }
        scratch_free(ix->arena, ix->arena_cap, ix->arena_stream);
    }
    if (ix->frame_ready) {
        // the pinned slot must not be handed out again while the copy into it is in flight
        cudaEventSynchronize(ix->frame_ready);
        cudaEventDestroy(ix->frame_ready);
    }
    pinned_slot_release(ix->h_frame);
    delete ix;
    return XG_OK;
}

extern "C" int xg_release_cached_memory(void) {
    cudaDeviceSynchronize();
    xg::release_cache();
    return XG_OK;
}

extern "C" int xg_index_info(const xg_index *ix, int64_t info[8], double *tol) {
    if (!ix) { set_error("xg_index_info: NULL index"); return XG_ERR_INVALID; }
    if (tol) {
        const int rc = index_host_values(ix);
        if (rc) return rc;
    }
    if (info) {
        info[0] = ix->n_face; info[1] = ix->k; info[2] = ix->n_levels; info[3] = (int64_t)ix->bytes;
        info[4] = ix->n_node; info[5] = info[6] = info[7] = 0;
    }
    if (tol) *tol = ix->default_tolerance;
    return XG_OK;
}
