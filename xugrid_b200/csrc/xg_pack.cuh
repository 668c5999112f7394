// xg_pack.cuh -- a mesh range repacked in space-filling-curve order (see xg_index.cu).
#pragma once
#include "xg_common.cuh"

namespace xg {

// d_face_ids (optional): the faces packed are d_face_ids[f_begin .. f_end) and perm holds
// positions in that list instead of face numbers (a rank's share of the targets when they
// are partitioned along the space-filling curve).
// Output arrays are caller-allocated device memory of n = f_end - f_begin
// entries (poly: n * k * 2 doubles).
struct PackedMesh {
    int64_t n = 0;
    int k = 0;
    int32_t *perm = nullptr;   // sorted position -> caller's face id (absolute)
    Box *box = nullptr;        // bbox per sorted face
    double *poly = nullptr;    // (n, k, 2) CCW vertices per sorted face
    uint8_t *nvert = nullptr;  // vertex count | 0x80 if the face was flipped
};

// scratch of pack_mesh handed in by a caller that must not allocate while it enqueues (a job
// being recorded into a CUDA graph); sizes: n keys / ids, pack_sort_tmp_bytes(n) for the sort
struct PackScratch {
    uint32_t *keys_in = nullptr, *keys_out = nullptr;
    int32_t *ids_in = nullptr;
    void *sort_tmp = nullptr;
    size_t sort_tmp_bytes = 0;
};
size_t pack_sort_tmp_bytes(int64_t n);

// d_frame: device pointer to [xmin, xmax, ymin, ymax] used to quantise the
// Morton keys; d_diag2 (optional): running max of squared bbox diagonals.
// d_clip_box (device, [xmin, xmax, ymin, ymax]) + d_kept (optional): only faces whose
// bbox strictly intersects the box are packed (out->n becomes their count; this is the one
// point where the stream is synchronised: the sort is sized on the host).
int pack_mesh(const double *d_xy, const void *d_faces, int idx_bytes, int64_t f_begin,
              int64_t f_end, int k, const double *d_frame, unsigned long long *d_diag2,
              PackedMesh *out, cudaStream_t s, const double *d_clip_box = nullptr,
              unsigned long long *d_kept = nullptr, const int32_t *d_face_ids = nullptr,
              int64_t n_node = 0,   // n_node > 0 with a clip box: cull by node outcodes first
              cudaStream_t alloc_stream = nullptr, const PackScratch *scratch = nullptr);

}  // namespace xg
