// xg_common.cuh -- shared declarations of libxgregrid (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/xg_regrid.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libxgregrid is written for sm_100a (B200) only"
#endif

namespace xg {

constexpr int WARP = 32;
constexpr int FANOUT = 32;          // children per node of the packed wide BVH
constexpr int MAX_LEVELS = 8;       // 32^8 faces is far beyond HBM capacity
constexpr int NUM_SMS = 148;        // B200

// leaf / node bounding box, celltree layout [xmin, xmax, ymin, ymax]
struct __align__(32) Box {
    double xmin, xmax, ymin, ymax;
};

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
    template <class T> T *as() const { return reinterpret_cast<T *>(p); }
};

}  // namespace xg

// The packed, Morton-ordered mesh + wide BVH.  Everything is in SORTED face
// order except perm (sorted position -> caller's face id).
struct xg_index {
    int64_t n_face = 0, n_node = 0;
    int k = 0;
    int n_levels = 0;                 // number of node levels above the leaves
    int64_t level_n[xg::MAX_LEVELS];  // nodes per level (level 0 = groups of FANOUT leaves)
    xg::Box *level_box[xg::MAX_LEVELS];
    xg::Box *leaf_box = nullptr;      // (n_face)
    double *poly = nullptr;           // (n_face, k, 2) CCW, padded slots repeat nothing (see nvert)
    uint8_t *nvert = nullptr;         // (n_face) number of vertices; bit 7 = face was flipped to CCW
    int32_t *perm = nullptr;          // (n_face)
    double bounds[4];                 // xmin, xmax, ymin, ymax of the mesh
    double default_tolerance = 0.0;   // 1e-12 * max bbox diagonal
    size_t bytes = 0;
    size_t arena_cap = 0;
    cudaStream_t arena_stream = nullptr;
    void *arena = nullptr;            // single device allocation backing all arrays
    double *d_frame = nullptr;        // device copy of bounds[4] + default tolerance
};
