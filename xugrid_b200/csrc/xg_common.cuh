// xg_common.cuh -- shared declarations of libxgregrid (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <mutex>
#include <vector>

#include "../../include/xg_regrid.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libxgregrid is written for sm_100a (B200) only"
#endif

namespace xg {

constexpr int WARP = 32;
constexpr int FANOUT = 32;          // children per node of the packed wide BVH
constexpr int MAX_LEVELS = 8;       // 32^8 faces is far beyond HBM capacity
// SMs of the current device (148 on a B200), asked once per device: persistent kernels launch
// one or a few blocks per SM
int num_sms();

// leaf / node bounding box, celltree layout [xmin, xmax, ymin, ymax]
struct __align__(32) Box {
    double xmin, xmax, ymin, ymax;
};

// Position of the cell (x, y) of a 65536 x 65536 grid along the Hilbert curve (32 bits).
// Faces, points and segments are ordered by this key: unlike the Z-order (Morton)
// curve the Hilbert curve has no jumps, so ANY run of consecutive keys is a connected,
// compact patch.  That is what a tile of 32 consecutive targets is: with Morton keys
// the bounding box of a tile covered 52 lattice cells on average (runs that straddle a
// Z jump are huge) and met 140 source boxes; with Hilbert keys 26 cells and 76 boxes
// (10 M-face jittered lattices), i.e. half the box tests and node visits of the walk.
// XG_SPACE_CURVE=0 selects the Morton key (tuning builds only).
#ifndef XG_SPACE_CURVE
#define XG_SPACE_CURVE 1
#endif
__host__ __device__ __forceinline__ uint32_t curve_key16(uint32_t x, uint32_t y) {
#if XG_SPACE_CURVE == 0
    auto spread = [](uint32_t v) {
        v &= 0xffffu;
        v = (v | (v << 8)) & 0x00ff00ffu;
        v = (v | (v << 4)) & 0x0f0f0f0fu;
        v = (v | (v << 2)) & 0x33333333u;
        v = (v | (v << 1)) & 0x55555555u;
        return v;
    };
    return spread(x) | (spread(y) << 1);
#else
    // bit-parallel evaluation (a prefix scan over the 16 levels in log steps instead of a
    // loop over them; checked against the level-by-level definition on the host)
    uint32_t A, B, C, D;
    {
        const uint32_t a = x ^ y, b = 0xFFFFu ^ a, c = 0xFFFFu ^ (x | y), d = x & (y ^ 0xFFFFu);
        A = a | (b >> 1);
        B = (a >> 1) ^ a;
        C = ((c >> 1) ^ (b & (d >> 1))) ^ c;
        D = ((a & (c >> 1)) ^ (d >> 1)) ^ d;
    }
#pragma unroll
    for (int sh = 2; sh <= 4; sh <<= 1) {
        const uint32_t a = A, b = B, c = C, d = D;
        A = (a & (a >> sh)) ^ (b & (b >> sh));
        B = (a & (b >> sh)) ^ (b & ((a ^ b) >> sh));
        C ^= (a & (c >> sh)) ^ (b & (d >> sh));
        D ^= (b & (c >> sh)) ^ ((a ^ b) & (d >> sh));
    }
    {
        const uint32_t a = A, b = B, c = C, d = D;
        C ^= (a & (c >> 8)) ^ (b & (d >> 8));
        D ^= (b & (c >> 8)) ^ ((a ^ b) & (d >> 8));
    }
    const uint32_t a = C ^ (C >> 1), b = D ^ (D >> 1);
    const uint32_t i0 = x ^ y, i1 = b | (0xFFFFu ^ (i0 | a));
    auto spread = [](uint32_t v) {
        v &= 0xffffu;
        v = (v | (v << 8)) & 0x00ff00ffu;
        v = (v | (v << 4)) & 0x0f0f0f0fu;
        v = (v | (v << 2)) & 0x33333333u;
        v = (v | (v << 1)) & 0x55555555u;
        return v;
    };
    return (spread(i1) << 1) | spread(i0);
#endif
}

void set_error(const char *fmt, ...);
int cuda_fail(cudaError_t e, const char *what, const char *file, int line);

#define XG_CUDA(call)                                                         \
    do {                                                                      \
        cudaError_t e__ = (call);                                             \
        if (e__ != cudaSuccess) return xg::cuda_fail(e__, #call, __FILE__, __LINE__); \
    } while (0)

#define XG_CHECK_LAUNCH() XG_CUDA(cudaGetLastError())

inline int64_t div_up(int64_t a, int64_t b) { return (a + b - 1) / b; }

// Device memory for scratch buffers and index arenas comes from a small block
// cache in front of cudaMallocAsync: a released block is kept and handed to the
// next request of a similar size on the same stream.  Re-use is safe because
// all work is stream-ordered; it matters because a regridder is rebuilt with
// the same sizes over and over, and asking the driver's pool for several GB
// per call stalls the stream for milliseconds whenever no device-wide
// synchronisation happened since the last free (measured: 0.8 ms -> 7 ms).
cudaError_t scratch_alloc(void **p, size_t *cap, size_t bytes, cudaStream_t s);
void scratch_free(void *p, size_t cap, cudaStream_t s);
// while set (per host thread), a request that the cache cannot serve fails instead of allocating
void scratch_set_capturing(bool on);

// stream-ordered scratch allocation that frees itself
struct DevBuf {
    void *p = nullptr;
    cudaStream_t s = nullptr;
    size_t bytes = 0, cap = 0;
    DevBuf() {}
    DevBuf(const DevBuf &) = delete;
    DevBuf &operator=(const DevBuf &) = delete;
    ~DevBuf() { release(); }
    cudaError_t alloc(size_t n, cudaStream_t stream) {
        release();
        s = stream;
        bytes = n ? n : 16;
        return scratch_alloc(&p, &cap, bytes, stream);
    }
    void release() {
        if (p) scratch_free(p, cap, s);
        p = nullptr;
    }
    // keeps a block that is large enough
    cudaError_t ensure(size_t n, cudaStream_t stream) {
        if (p && cap >= (n ? n : 16) && s == stream) { bytes = n ? n : 16; return cudaSuccess; }
        return alloc(n, stream);
    }
    template <class T> T *as() const { return reinterpret_cast<T *>(p); }
};

// small pinned host slots for asynchronous read-backs of a few values
double *pinned_slot_acquire();
void pinned_slot_release(double *p);

}  // namespace xg

// The packed, curve-ordered mesh + wide BVH.  Everything is in SORTED face
// order except perm (sorted position -> caller's face id).
struct xg_index {
    int64_t n_face = 0, n_node = 0;   // n_face: faces INDEXED (a culled index holds fewer than the mesh)
    int64_t n_face_mesh = 0;          // faces of the mesh the index was built from
    int k = 0;
    int n_levels = 0;                 // number of node levels above the leaves
    int64_t level_n[xg::MAX_LEVELS];  // nodes per level (level 0 = groups of FANOUT leaves)
    xg::Box *level_box[xg::MAX_LEVELS];
    xg::Box *leaf_box = nullptr;      // (n_face)
    double *poly = nullptr;           // (n_face, k, 2) CCW, padded slots repeat nothing (see nvert)
    uint8_t *nvert = nullptr;         // (n_face) number of vertices; bit 7 = face was flipped to CCW
    int32_t *perm = nullptr;          // (n_face)
    // bounds of the mesh and the default tolerance are reduced on the device and copied to
    // a pinned host slot asynchronously; the host values below are valid after
    // xg::index_host_values(), which waits for that copy the first time they are needed
    // (overlap jobs never need them: they read d_frame on the device)
    double bounds[4];                 // xmin, xmax, ymin, ymax of the mesh
    double default_tolerance = 0.0;   // 1e-12 * max bbox diagonal
    double *h_frame = nullptr;        // pinned host slot (5 doubles)
    cudaEvent_t frame_ready = nullptr;
    bool host_values_valid = false;
    // streams other than the creation stream that queries were enqueued on: the arena is
    // released in stream order behind all of them, without a device-wide synchronisation
    std::mutex mu;
    std::vector<cudaStream_t> foreign_streams;
    size_t bytes = 0;
    size_t arena_cap = 0;
    cudaStream_t arena_stream = nullptr;
    void *arena = nullptr;            // single device allocation backing all arrays
    double *d_frame = nullptr;        // device copy of bounds[4] + default tolerance
};

namespace xg {
// host copies of bounds / default tolerance (waits for the creation-time copy once)
int index_host_values(const xg_index *ix);
// a query on `s` uses the index: remember streams other than the creation stream
void index_note_use(const xg_index *ix, cudaStream_t s);
}  // namespace xg
