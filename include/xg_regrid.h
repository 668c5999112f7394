/*
 * xg_regrid.h -- C ABI of libxgregrid.so: the sm_100a implementation of the
 * xugrid regridding hot path (weight construction + weight application).
 *
 * Conventions
 *  - every pointer named d_* is a DEVICE pointer owned by the caller (e.g. a
 *    torch tensor's data_ptr()); h_* is a host pointer.  The library never
 *    keeps a caller pointer after the call returns, except xg_index, which
 *    owns a private packed copy of the mesh.
 *  - every function returns 0 on success or a negative xg_status; the message
 *    of the last failure on the calling thread is xg_last_error_string().
 *    Nothing throws, nothing falls back to the CPU.
 *  - `stream` is a cudaStream_t passed as void*; work is enqueued on it.
 *    Functions that must return a size to the host synchronise that stream
 *    before returning: xg_overlap_begin (nnz; xg_overlap_run does not), the culled index builds (number
 *    of faces kept), xg_intersect_edges (count), xg_face_range_bounds;
 *    xg_index_info waits for the index the first time the default tolerance
 *    is asked for.  Everything else, including xg_index_create and
 *    xg_index_destroy, is asynchronous.
 *  - connectivity is (n_face, k) row-major with -1 as fill value and 0-based
 *    node indices (what xugrid's Ugrid2d normalises to:
 *    /root/reference/xugrid/ugrid/ugrid2d.py:103-109), as int32 or int64
 *    (idx_bytes = 4 or 8; the reference's IntDType is intp, constants.py:10).
 *  - coordinates are (n_node, 2) float64 row-major, i.e. Ugrid2d.node_coordinates
 *    (/root/reference/xugrid/ugrid/ugridbase.py:576-579).
 *  - thread safety: distinct handles / streams may be used concurrently; an
 *    xg_index is read-only after creation and may be shared (the reference
 *    caches its celltree the same way, ugrid2d.py:908-921).
 *
 * Reference interface each entry point replaces (file:line under
 * /root/reference) is given above the declaration.
 */
#ifndef XG_REGRID_H
#define XG_REGRID_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define XG_ABI_VERSION 1

typedef enum xg_status {
    XG_OK = 0,
    XG_ERR_INVALID = -1,   /* bad argument */
    XG_ERR_CUDA = -2,      /* CUDA runtime error, see xg_last_error_string() */
    XG_ERR_NOMEM = -3,
    XG_ERR_UNSUPPORTED = -4 /* e.g. more than XG_MAX_N_VERTEX vertices per face */
} xg_status;

/* faces with more vertices than this are rejected (numba_celltree: MAX_N_VERTEX = 32) */
#define XG_MAX_N_VERTEX 32

/* reduction codes of xg_csr_apply: xugrid/regrid/reduce.py:254-272 */
typedef enum xg_method {
    XG_MEAN = 0,             /* reduce.py:16-27  */
    XG_HARMONIC_MEAN = 1,    /* reduce.py:30-42  */
    XG_GEOMETRIC_MEAN = 2,   /* reduce.py:45-72  */
    XG_SUM = 3,              /* reduce.py:75-87  (unweighted!) */
    XG_MINIMUM = 4,          /* reduce.py:90-104 */
    XG_MAXIMUM = 5,          /* reduce.py:107-123 */
    XG_MODE = 6,             /* reduce.py:126-158 */
    XG_PERCENTILE = 7,       /* reduce.py:161-203, 241-251; `percentile` argument */
    XG_MAX_OVERLAP = 8,      /* reduce.py:225-238 */
    XG_CONSERVATIVE = 9      /* reduce.py:206-222 first_order_conservative / conductance */
} xg_method;

typedef struct xg_index xg_index;       /* spatial index over one mesh (device resident) */
typedef struct xg_overlap xg_overlap;   /* one in-flight overlap job */

int xg_abi_version(void);
const char *xg_last_error_string(void);
/* The library keeps released scratch blocks and index arenas for re-use by the
 * next call of similar size (same device, same stream).  This returns them to
 * the driver (synchronises the device). */
int xg_release_cached_memory(void);

/* ---- spatial index -------------------------------------------------------
 * Replaces numba_celltree.CellTree2d(vertices, faces, fill_value) as built by
 * Ugrid2d.celltree, xugrid/ugrid/ugrid2d.py:908-921.  Faces are forced
 * counter-clockwise internally (as CellTree2d does), the caller's arrays are
 * not modified. */
int xg_index_create(const double *d_xy, int64_t n_node, const void *d_faces, int idx_bytes,
                    int64_t n_face, int k, void *stream, xg_index **out);
/* Same, but only faces whose bounding box strictly intersects
 * h_clip_box = [xmin, xmax, ymin, ymax] (host pointer) are indexed: a rank that
 * owns a contiguous range of target faces builds its index over the part of a
 * replicated source mesh that its targets can touch.  Face ids reported by
 * queries are still the caller's.  NULL box = xg_index_create. */
int xg_index_create_clipped(const double *d_xy, int64_t n_node, const void *d_faces, int idx_bytes,
                            int64_t n_face, int k, const double *h_clip_box, void *stream,
                            xg_index **out);
/* Same cull, but the box is the bounding box of the target faces [t_begin, t_end) of
 * another mesh, reduced on the device: nothing about the targets travels to the host, so a
 * rank's sharded build costs one host round trip less than xg_face_range_bounds +
 * xg_index_create_clipped. */
int xg_index_create_for_targets(const double *d_xy, int64_t n_node, const void *d_faces, int idx_bytes,
                                int64_t n_face, int k, const double *d_t_xy, const void *d_t_faces,
                                int t_idx_bytes, int64_t t_begin, int64_t t_end, int k_t, void *stream,
                                xg_index **out);
/* The same with the targets given as a list of face numbers (a rank's share of a
 * partition along the space-filling curve, xg_target_shard). */
int xg_index_create_for_target_ids(const double *d_xy, int64_t n_node, const void *d_faces, int idx_bytes,
                                   int64_t n_face, int k, const double *d_t_xy, const void *d_t_faces,
                                   int t_idx_bytes, const int32_t *d_t_ids, int64_t n_ids, int k_t,
                                   void *stream, xg_index **out);
/* Partition of the faces of a mesh over `world` ranks by position along the
 * space-filling (Hilbert) curve, so that every rank owns a compact patch whatever the
 * caller's face numbering means (the reference has no counterpart).  All
 * ranks key all faces in the frame of the mesh's own nodes, count them in 65 536 bins of
 * the leading key bits and cut the cumulative count at multiples of n_face / world: the
 * shards are disjoint, cover every face and differ in size by at most one bin.  Writes
 * the face numbers of shard `rank` to d_ids (capacity n_face, no particular order) and
 * their number to *h_count (the stream is synchronised). */
int xg_target_shard(const double *d_xy, int64_t n_node, const void *d_faces, int idx_bytes,
                    int64_t n_face, int k, int rank, int world, int32_t *d_ids, int64_t *h_count,
                    void *stream);
/* h_box = [xmin, xmax, ymin, ymax] of the vertices of faces [f_begin, f_end):
 * the clip box of a target shard (synchronises the stream). */
int xg_face_range_bounds(const double *d_xy, const void *d_faces, int idx_bytes, int64_t f_begin,
                         int64_t f_end, int k, double h_box[4], void *stream);
int xg_index_destroy(xg_index *index);
/* info[0]=indexed faces info[1]=k info[2]=levels info[3]=device bytes owned
 * info[4]=n_node; tol = default tolerance = 1e-12 * max bbox diagonal
 * (xugrid/ugrid/ugridbase.py:1165-1170) */
int xg_index_info(const xg_index *index, int64_t info[8], double *default_tolerance);

/* ---- overlap weights -----------------------------------------------------
 * Replaces UnstructuredGrid2d.overlap -> CellTree2d.intersect_faces
 * (xugrid/regrid/unstructured.py:109-135) followed by MatrixCSR.from_triplet
 * (xugrid/core/sparse.py:61-78,119-127): all (target, source) face pairs whose
 * Sutherland-Hodgman intersection area is > 0, as CSR over the target rows
 * [t_begin, t_end).  The source mesh is the index; the target faces are the
 * queries.
 *
 * xg_overlap_begin does the traversal and the clipping, writes
 *   d_indptr[0 .. t_end - t_begin]  (int64, local to the range, d_indptr[0]=0)
 * and returns nnz in *h_nnz (the stream is synchronised).  The caller then
 * allocates indices/area of that size and calls xg_overlap_finish, which
 * writes them with every row sorted by ascending source index (a canonical
 * order: the reference's within-row order is its tree traversal order) and
 * releases the job.  `relative` != 0 divides each area by the area of its
 * source face (unstructured.py:133-134, connectivity.py:615-633). */
int xg_overlap_begin(const xg_index *source, const double *d_t_xy, int64_t n_t_node,
                     const void *d_t_faces, int idx_bytes, int64_t n_t_face, int k_t,
                     int64_t t_begin, int64_t t_end, int64_t *d_indptr, int64_t *h_nnz,
                     void *stream, xg_overlap **job);
/* The same for the target faces d_t_ids[0 .. n_ids) (e.g. from xg_target_shard): row i of
 * the CSR belongs to face d_t_ids[i]. */
int xg_overlap_begin_ids(const xg_index *source, const double *d_t_xy, int64_t n_t_node,
                         const void *d_t_faces, int idx_bytes, int64_t n_t_face, int k_t,
                         const int32_t *d_t_ids, int64_t n_ids, int64_t *d_indptr, int64_t *h_nnz,
                         void *stream, xg_overlap **job);
int xg_overlap_finish(xg_overlap *job, const int64_t *d_indptr, int32_t *d_indices,
                      double *d_area, int relative, void *stream);
/* The same job enqueued in one go, WITHOUT a host round trip: the caller hands in CSR arrays
 * sized by an upper bound (entry_cap entries; 16 per target row always suffices at comparable
 * resolutions, see below) instead of waiting for nnz, so the host can run ahead and queue the
 * apply and the copies behind it (the reference has no counterpart: numba code is
 * synchronous).  d_t_ids may be NULL (targets = faces [t_begin, t_end)) or a face list
 * (targets = d_t_ids[t_begin .. t_end)).  d_indptr receives the rows + 1 row pointers.
 * flags & XG_RUN_CHAINED: d_indptr[0] already holds, ON THE DEVICE, the offset of this job's first
 * entry -- pass the address of the last row pointer of the job before it and the jobs of
 * consecutive target ranges fill ONE CSR matrix (how the regridders consume a target mesh
 * that is still being uploaded range by range); otherwise the matrix starts at 0.
 * Rows are written sorted by ascending source index, as by xg_overlap_finish.
 * Nothing is waited for.  xg_overlap_result waits for the job, returns its nnz, the row
 * pointer behind its last row and whether it is complete: a job whose candidate arena or
 * whose caller arrays turned out too small is NOT complete (its rows are missing, nothing
 * was written out of bounds) and must be redone with xg_overlap_begin/finish, which
 * re-sizes.  xg_overlap_stats / _timings are valid after xg_overlap_result; release the
 * job with xg_overlap_abort. */
#define XG_RUN_CHAINED 1 /* d_indptr[0] holds the offset of the job's first entry (see above) */
#define XG_RUN_GRAPH 2   /* launch the job as ONE CUDA graph: its ~30 small launches otherwise cost
                          * 25-40 us each while a mesh upload saturates PCIe; no per-phase timings then
                          * (stats[7] = 1 when the job really went out as a graph: a scratch block that
                          * is not in the library's cache yet makes it fall back to direct launches) */
int xg_overlap_run(const xg_index *source, const double *d_t_xy, int64_t n_t_node,
                   const void *d_t_faces, int idx_bytes, int64_t n_t_face, int k_t,
                   const int32_t *d_t_ids, int64_t t_begin, int64_t t_end, int flags,
                   int64_t *d_indptr, int32_t *d_indices, double *d_area, int64_t entry_cap,
                   int relative, void *stream, xg_overlap **job);
int xg_overlap_result(xg_overlap *job, int64_t *h_nnz, int64_t *h_nnz_end, int *h_complete);
/* counters of the job for the byte model:
 * stats[0]=targets stats[1]=wide-node visits (V, one coalesced load of <= 32
 * boxes each) stats[2]=leaf bbox tests (C) stats[3]=pairs that passed the
 * separating-axis filter and reached the clipper (C_clip) stats[4]=nnz
 * stats[5]=bbox candidates stats[6]=full walks of the index (2 when the pair
 * arena had to be re-sized).  Valid after xg_overlap_begin. */
int xg_overlap_stats(const xg_overlap *job, int64_t stats[8]);
/* device time of each phase of xg_overlap_begin in milliseconds (CUDA events on
 * the job's stream), indexed by xg_phase.  Valid after xg_overlap_begin. */
typedef enum xg_phase {
    XG_PHASE_PACK = 0,      /* Morton order + packing of the target faces */
    XG_PHASE_TRAVERSE = 1,  /* BVH walk -> candidate pairs */
    XG_PHASE_SAT = 2,       /* separating-axis filter -> survivors */
    XG_PHASE_CLIP = 3,      /* Sutherland-Hodgman area of the survivors */
    XG_PHASE_ROWPTR = 4     /* CSR row pointer (+ size read-back) */
} xg_phase;
#define XG_N_PHASES 5
int xg_overlap_timings(const xg_overlap *job, double ms[8]);
int xg_overlap_abort(xg_overlap *job);

/* ---- point queries -------------------------------------------------------
 * Replaces CellTree2d.locate_points(points, tolerance)
 * (xugrid/regrid/unstructured.py:138-139, xugrid/ugrid/ugridbase.py:1305-1323,
 * xugrid/ugrid/ugrid2d.py:1099): index of a face containing each point (ray
 * casting, or within `tolerance` of an edge), -1 outside.  tolerance < 0
 * selects the default.  When several faces qualify (a point on a shared
 * edge) the smallest face index is returned. */
int xg_locate_points(const xg_index *index, const double *d_points, int64_t n_point,
                     double tolerance, int32_t *d_face_index, void *stream);
/* The same question asked only for "inside or not" -- the mask of
 * UnstructuredGrid2d.barycentric (xugrid/regrid/unstructured.py:188-190:
 * locate_points(...) == -1): d_face_index receives SOME face that holds the point (not
 * necessarily the smallest index when several qualify) or -1; the walk of a point stops at its
 * first qualifying face. */
int xg_locate_any(const xg_index *index, const double *d_points, int64_t n_point,
                  double tolerance, int32_t *d_face_index, void *stream);
/* Replaces CellTree2d.compute_barycentric_weights(points, tolerance)
 * (xugrid/ugrid/ugrid2d.py:1054-1078): triangle formula when k == 3,
 * Wachspress coordinates otherwise; d_weights is (n_point, k), zero for
 * points outside and for fill slots. */
int xg_barycentric(const xg_index *index, const double *d_points, int64_t n_point,
                   double tolerance, int32_t *d_face_index, double *d_weights, void *stream);

/* ---- Voronoi preparation of the barycentric interpolator -------------------
 * Replaces voronoi_topology(add_exterior=True, add_vertices=True, skip_concave=True)
 * (xugrid/ugrid/voronoi.py:330-458) as called from UnstructuredGrid2d.barycentric
 * (xugrid/regrid/unstructured.py:146-175): the polygons around the NODES of a mesh whose
 * vertices are the centroids of the adjacent faces (d_centroids, (n_face, 2), e.g. from
 * xg_face_centroid); around an exterior node also the projections of the boundary faces'
 * centroids on the two exterior edges that meet there and one corner vertex (the node
 * itself, or the midpoint of the two projections at a concave corner).
 * xg_voronoi_begin analyses the mesh and returns sizes[0] = interior polygons,
 * sizes[1] = exterior polygons (= exterior nodes), sizes[2] = projection vertices kept,
 * sizes[3] = largest polygon (the stream is synchronised).  The caller allocates
 *   d_vertices    (n_face + sizes[2] + sizes[1], 2) float64: [centroids | projections | corners]
 *   d_polygons    (sizes[0] + sizes[1], sizes[3]) int32, -1 fill, counter-clockwise;
 *                 interior polygons first, each group in ascending node number
 *   d_vertex_face (n_vertex) int32: the face whose value a vertex carries, -1 for corners
 *   d_corner_map  (sizes[1], 2) int32: the two projection vertices a corner sits between
 * and xg_voronoi_finish fills them and releases the job (xg_voronoi_abort: release only).
 * The caller's mesh arrays must stay alive between the two calls. */
typedef struct xg_voronoi xg_voronoi;
int xg_voronoi_begin(const double *d_xy, int64_t n_node, const void *d_faces, int idx_bytes,
                     int64_t n_face, int k, const double *d_centroids, int64_t sizes[4],
                     void *stream, xg_voronoi **job);
int xg_voronoi_finish(xg_voronoi *job, double *d_vertices, int32_t *d_polygons, int32_t *d_vertex_face,
                      int32_t *d_corner_map, void *stream);
int xg_voronoi_abort(xg_voronoi *job);
/* The assembly behind xg_barycentric over the Voronoi polygons
 * (xugrid/regrid/unstructured.py:176-201 with replace_interpolated_weights, :17-57):
 * d_weights (n_point, kv) is adjusted IN PLACE -- the weight of a corner vertex goes to its
 * two projection vertices, inversely proportional to distance; points whose d_source_face
 * (xg_locate_points on the SOURCE mesh) is -1 get zero weights -- the positive weights per
 * point are counted, d_indptr (n_point + 1) is written and nnz returned (the stream is
 * synchronised).  xg_bary_csr_fill then writes the CSR entries: column = the source face of
 * the polygon vertex, in the polygon's vertex order. */
int xg_bary_csr_count(const int32_t *d_poly_index, double *d_weights, int64_t n_point, int kv,
                      const int32_t *d_polygons, const int32_t *d_vertex_face,
                      const int32_t *d_corner_map, const double *d_vertices, int64_t corner_start,
                      const int32_t *d_source_face, int64_t *d_indptr, int64_t *h_nnz, void *stream);
int xg_bary_csr_fill(const int32_t *d_poly_index, const double *d_weights, int64_t n_point, int kv,
                     const int32_t *d_polygons, const int32_t *d_vertex_face, const int64_t *d_indptr,
                     int32_t *d_indices, double *d_data, void *stream);

/* ---- per-face geometry ----------------------------------------------------
 * connectivity.area (xugrid/ugrid/connectivity.py:615-633) and
 * connectivity.centroids (connectivity.py:636-664). */
int xg_face_area(const double *d_xy, const void *d_faces, int idx_bytes, int64_t n_face, int k,
                 double *d_area, void *stream);
int xg_face_centroid(const double *d_xy, const void *d_faces, int idx_bytes, int64_t n_face,
                     int k, double *d_centroid_xy, void *stream);

/* ---- connectivity transport ----------------------------------------------
 * A connectivity whose mesh has fewer than 2^24 - 1 nodes can cross PCIe as 3 bytes per index
 * (little-endian, 0xFFFFFF = fill) instead of 4 or 8: xg_unpack_u24 widens n_values packed
 * indices to int32 (-1 = fill) on the device.  d_packed must be 4-byte aligned (upload whole
 * groups of 4 indices), d_out 16-byte aligned.  The reference has no counterpart (its
 * connectivity never leaves host memory); the grids of this package stage the packed copy once,
 * like the reference's Ugrid2d.__init__ copies its connectivity once (ugrid2d.py:103-109). */
int xg_unpack_u24(const void *d_packed, int64_t n_values, int32_t *d_out, void *stream);

/* ---- segments against faces ----------------------------------------------
 * Replaces CellTree2d.intersect_edges as called from
 * xugrid/regrid/unstructured.py:203-215 (UnstructuredGrid2d.intersection_length,
 * the weights of NetworkGridder, xugrid/regrid/gridder.py:66-73): every
 * (segment, face) pair whose Cyrus-Beck clip has positive length.
 * d_edges is (n_edge, 2, 2) float64.  Up to `cap` results are written, in no
 * particular order: d_edge_index / d_face_index (caller's ids) and d_xy
 * (cap, 2, 2) = the clipped segment.  *h_count receives the number of pairs
 * found (the stream is synchronised); if it exceeds cap, call again with a
 * larger capacity. */
int xg_intersect_edges(const xg_index *index, const double *d_edges, int64_t n_edge, int64_t cap,
                       int32_t *d_edge_index, int32_t *d_face_index, double *d_xy,
                       int64_t *h_count, void *stream);

/* ---- weight application ---------------------------------------------------
 * Replaces make_regrid(func)._regrid(source, A, size)
 * (xugrid/regrid/regridder.py:34-69):
 *   out[e, r] = f(src[e, indices[row r]], weights[row r])   for r in [row_begin,row_end)
 * rows with no entries give NaN (regridder.py:63).  d_src is (n_extra, m)
 * row-major float64; d_out is written at d_out[e * out_ld + (r - row_begin)].
 * d_indptr is the int64 row pointer of the rows [row_begin,row_end] of ONE CSR
 * matrix, i.e. d_indptr[r - row_begin] .. d_indptr[r - row_begin + 1] index
 * d_indices/d_weights directly. */
int xg_csr_apply(int method, double percentile, const int64_t *d_indptr, const int32_t *d_indices,
                 const double *d_weights, const double *d_src, int64_t n_extra, int64_t m,
                 int64_t n_rows, double *d_out, int64_t out_ld, void *stream);
/* The apply FUSED with the all-gather that reassembles the output of a job whose target
 * rows are sharded over the GPUs of one box (the reference has no counterpart):
 * h_d_outs is a HOST array of n_out (<= 16) DEVICE pointers, one per rank of the job, each
 * the position of this rank's first column in that rank's copy of the (n_extra, n_target)
 * output (peer memory mapped over NVLink, e.g. torch symmetric memory); every result is
 * stored to all of them from inside the kernel, so the transfer overlaps the reduction and
 * no send slab, pad or collective launch is needed.  The caller orders the ranks around the
 * call (a device-side barrier before peers may be overwritten and after the kernel). */
int xg_csr_apply_peers(int method, double percentile, const int64_t *d_indptr, const int32_t *d_indices,
                       const double *d_weights, const double *d_src, int64_t n_extra, int64_t m,
                       int64_t n_rows, double *const *h_d_outs, int n_out, int64_t out_ld, void *stream);
/* Replaces CentroidLocatorRegridder._regrid (xugrid/regrid/regridder.py:400-409):
 *   out[e, row[q]] = src[e, col[q]], NaN elsewhere.  d_out is (n_extra, size). */
int xg_coo_gather(const int32_t *d_row, const int32_t *d_col, int64_t nnz, const double *d_src,
                  int64_t n_extra, int64_t m, double *d_out, int64_t size, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* XG_REGRID_H */
