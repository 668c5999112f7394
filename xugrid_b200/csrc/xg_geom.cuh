// xg_geom.cuh -- per-pair / per-point geometry on the device.
//
// Arithmetic follows the published algorithms the reference reaches through
// numba_celltree (separating axis test, Sutherland-Hodgman clip with the area
// taken as a fan of abs(cross), W. R. Franklin's pnpoly with an on-edge
// tolerance, triangle and Wachspress barycentric coordinates); call sites:
// /root/reference/xugrid/regrid/unstructured.py:124-139,
// /root/reference/xugrid/ugrid/ugrid2d.py:1054-1078.  The translation units
// that include this file are compiled with -fmad=false so that every
// multiply/add rounds separately, as LLVM does for numba on x86: overlap
// areas are then reproducible to the last bit against a CPU evaluation of
// the same expressions.
#pragma once
#include "xg_common.cuh"

namespace xg {

__device__ __forceinline__ double cross2(double ux, double uy, double vx, double vy) {
    return ux * vy - uy * vx;
}

// vertex n-1 of a register-resident polygon without dynamic indexing
template <int K>
__device__ __forceinline__ double2 last_vertex(const double2 (&p)[K], int n) {
    double2 v = p[0];
#pragma unroll
    for (int i = 1; i < K; ++i)
        if (i == n - 1) v = p[i];
    return v;
}

template <int KA, int KB>
__device__ __forceinline__ bool is_separating_axis(double nx, double ny, const double2 (&a)[KA],
                                                   int na, const double2 (&b)[KB], int nb) {
    double mina = INFINITY, maxa = -INFINITY, minb = INFINITY, maxb = -INFINITY;
#pragma unroll
    for (int i = 0; i < KA; ++i)
        if (i < na) {
            double p = a[i].x * nx + a[i].y * ny;
            mina = fmin(mina, p);
            maxa = fmax(maxa, p);
        }
#pragma unroll
    for (int i = 0; i < KB; ++i)
        if (i < nb) {
            double p = b[i].x * nx + b[i].y * ny;
            minb = fmin(minb, p);
            maxb = fmax(maxb, p);
        }
    return !(maxa > minb && maxb > mina);
}

// Separating axis theorem over the edge normals of both polygons.  Polygons
// that only touch count as separated (strict inequalities).
template <int KA, int KB>
__device__ __forceinline__ bool polygons_intersect(const double2 (&a)[KA], int na,
                                                   const double2 (&b)[KB], int nb) {
    double2 p = last_vertex(a, na);
#pragma unroll
    for (int i = 0; i < KA; ++i)
        if (i < na) {
            double2 q = a[i];
            double nx = -(q.y - p.y), ny = q.x - p.x;
            p = q;
            if (nx == 0.0 && ny == 0.0) continue;
            if (is_separating_axis(nx, ny, a, na, b, nb)) return false;
        }
    p = last_vertex(b, nb);
#pragma unroll
    for (int i = 0; i < KB; ++i)
        if (i < nb) {
            double2 q = b[i];
            double nx = -(q.y - p.y), ny = q.x - p.x;
            p = q;
            if (nx == 0.0 && ny == 0.0) continue;
            if (is_separating_axis(nx, ny, a, na, b, nb)) return false;
        }
    return true;
}

// Sutherland-Hodgman: area of (subject polygon) clipped by the convex CCW
// clipper.  The two ping-pong vertex lists live in shared memory, one column
// per thread: element v of this thread is buf[v * stride].
template <int KT, int KS>
__device__ __forceinline__ double clip_area(const double2 (&poly)[KT], int np,
                                            const double2 (&clip)[KS], int nc, double2 *bufA,
                                            double2 *bufB, int stride) {
    double2 *out = bufA, *sub = bufB;
    int n_out = np;
#pragma unroll
    for (int i = 0; i < KT; ++i)
        if (i < np) out[i * stride] = poly[i];

    double2 r = last_vertex(clip, nc);
#pragma unroll
    for (int i = 0; i < KS; ++i) {
        if (i < nc) {
            const double2 s = clip[i];
            const double Ux = s.x - r.x, Uy = s.y - r.y;
            if (Ux == 0.0 && Uy == 0.0) continue;  // r is not advanced
            const double Nx = -Uy, Ny = Ux;
            double2 *t = sub;
            sub = out;
            out = t;
            const int length = n_out;
            n_out = 0;
            double2 a = sub[(length - 1) * stride];
            bool a_in = Ux * (a.y - r.y) > Uy * (a.x - r.x);
            for (int j = 0; j < length; ++j) {
                const double2 b = sub[j * stride];
                const double Vx = b.x - a.x, Vy = b.y - a.y;
                if (Vx == 0.0 && Vy == 0.0) continue;
                bool b_in = Ux * (b.y - r.y) > Uy * (b.x - r.x);
                if (b_in != a_in) {
                    const double Wx = a.x - r.x, Wy = a.y - r.y;
                    const double nw = Nx * Wx + Ny * Wy;
                    const double nv = Nx * Vx + Ny * Vy;
                    if (nv != 0.0) {
                        const double tt = -nw / nv;
                        out[n_out * stride] = make_double2(a.x + tt * Vx, a.y + tt * Vy);
                        ++n_out;
                    } else if (a_in) {
                        // going out along a parallel edge: treat b as inside
                        b_in = true;
                    }
                }
                if (b_in) {
                    out[n_out * stride] = b;
                    ++n_out;
                }
                a = b;
                a_in = b_in;
            }
            if (n_out < 3) return 0.0;
            r = s;
        }
    }
    // area: fan from vertex 0, every triangle contributes abs(cross)
    double area = 0.0;
    const double2 a0 = out[0];
    const double2 b0 = out[stride];
    double Ux = b0.x - a0.x, Uy = b0.y - a0.y;
    for (int i = 2; i < n_out; ++i) {
        const double2 c = out[i * stride];
        const double Vx = a0.x - c.x, Vy = a0.y - c.y;
        area += fabs(cross2(Ux, Uy, Vx, Vy));
        Ux = Vx;
        Uy = Vy;
    }
    return 0.5 * area;
}

// ---------------------------------------------------------------------------
// Warp-uniform variants used by the dense clipping kernel.  Same arithmetic,
// same results as polygons_intersect() / clip_area() above, but written as
// straight-line code over fixed trip counts with predicates, so the 32 lanes
// of a warp never execute different instructions: ptxas otherwise unrolls the
// data-dependent loops with per-length entry points and the lanes serialise
// (measured: 17-53 % of lanes active, profiles/r1_v1*, r1_v2*).

template <int KA, int KB>
__device__ __forceinline__ bool axis_separates(double nx, double ny, const double2 (&a)[KA], int na,
                                               const double2 (&b)[KB], int nb) {
    double mina = a[0].x * nx + a[0].y * ny, maxa = mina;
#pragma unroll
    for (int i = 1; i < KA; ++i)
        if (i < na) {
            const double p = a[i].x * nx + a[i].y * ny;
            mina = p < mina ? p : mina;
            maxa = p > maxa ? p : maxa;
        }
    double minb = b[0].x * nx + b[0].y * ny, maxb = minb;
#pragma unroll
    for (int i = 1; i < KB; ++i)
        if (i < nb) {
            const double p = b[i].x * nx + b[i].y * ny;
            minb = p < minb ? p : minb;
            maxb = p > maxb ? p : maxb;
        }
    return !(maxa > minb && maxb > mina);
}

// all edge normals are evaluated (no per-lane early exit): 63 % of the
// candidate pairs intersect and need every axis anyway
template <int KA, int KB>
__device__ __forceinline__ bool polygons_intersect_uniform(const double2 (&a)[KA], int na,
                                                           const double2 (&b)[KB], int nb) {
    bool separated = false;
    double2 p = last_vertex(a, na);
#pragma unroll
    for (int i = 0; i < KA; ++i)
        if (i < na) {
            const double2 q = a[i];
            const double nx = -(q.y - p.y), ny = q.x - p.x;
            p = q;
            const bool sep = axis_separates(nx, ny, a, na, b, nb);
            separated |= sep && !(nx == 0.0 && ny == 0.0);
        }
    p = last_vertex(b, nb);
#pragma unroll
    for (int i = 0; i < KB; ++i)
        if (i < nb) {
            const double2 q = b[i];
            const double nx = -(q.y - p.y), ny = q.x - p.x;
            p = q;
            const bool sep = axis_separates(nx, ny, a, na, b, nb);
            separated |= sep && !(nx == 0.0 && ny == 0.0);
        }
    return !separated;
}



// ---------------------------------------------------------------------------
// fp32 pre-decision of the separating-axis test (what k_sat runs first).
//
// The reference decides "polygons intersect" from comparisons of fp64 projections
// maxa > minb && maxb > mina per edge normal.  For all but a vanishing fraction of
// the candidate pairs those comparisons are nowhere near a tie, and a much cheaper
// evaluation decides them with certainty:
//   * coordinates relative to the first vertex of A (one fp64 subtraction each, so
//     the large absolute coordinates cancel before anything is rounded to fp32),
//   * normals and projections in fp32 (twice the rate of fp64 on B200, FMA allowed),
//   * L = largest |local coordinate|, X = largest |absolute coordinate|.
// Error budget for a gap g = (projection of A) - (projection of B) on one axis:
//   fp32 evaluation vs exact arithmetic   |g32 - g*| <= 2 * 20 u32 L^2   (u32 = 2^-24)
//   the reference's fp64 evaluation       |g64 - g*| <= 2 * 16 u64 X L   (u64 = 2^-53)
// so with T = 2^-17 L^2 + 2^-47 X L  (each term rounded up by more than 2x):
//   g32 >  T  =>  g64 > 0  (the reference's comparison is true)
//   g32 < -T  =>  g64 < 0  (false; "<=": a tie in fp64 can only happen inside the band)
// A pair is REJECTED when some axis is surely separating, ACCEPTED when on every axis
// both comparisons are surely true, otherwise AMBIGUOUS and decided by the fp64 routine
// above -- the outcome is always exactly polygons_intersect()'s.  Zero-length edges
// (skipped by the reference) give all-zero projections here, i.e. never "sure".
enum { SAT_REJECT = 0, SAT_ACCEPT = 1, SAT_AMBIGUOUS = 2 };

template <int KA, int KB>
__device__ __forceinline__ int polygons_intersect_fp32(const double2 (&a)[KA], int na,
                                                       const double2 (&b)[KB], int nb) {
    const double ox = a[0].x, oy = a[0].y;
    float ax[KA], ay[KA], bx[KB], by[KB];
    float L = 0.f;
#pragma unroll
    for (int i = 0; i < KA; ++i) {
        ax[i] = (float)(a[i].x - ox);
        ay[i] = (float)(a[i].y - oy);
        if (i < na) L = fmaxf(L, fmaxf(fabsf(ax[i]), fabsf(ay[i])));
    }
#pragma unroll
    for (int i = 0; i < KB; ++i) {
        bx[i] = (float)(b[i].x - ox);
        by[i] = (float)(b[i].y - oy);
        if (i < nb) L = fmaxf(L, fmaxf(fabsf(bx[i]), fabsf(by[i])));
    }
    const float X = (float)fmax(fabs(ox), fabs(oy)) * 1.0000002f + L;
    const float T = 7.6293945e-6f * L * L + 7.1054274e-15f * X * L;  // 2^-17 L^2 + 2^-47 X L
    // (fp32 range: local coordinates far below 1e-15 or above 1e15 are left to fp64)
    const bool in_range = L > 1e-15f && L < 1e15f && X < 1e30f;
    bool sure_sep = false, all_sure_ovl = true;
    auto axis = [&](float nx, float ny) {
        float mina = ax[0] * nx + ay[0] * ny, maxa = mina;
#pragma unroll
        for (int i = 1; i < KA; ++i)
            if (i < na) {
                const float p = ax[i] * nx + ay[i] * ny;
                mina = fminf(mina, p);
                maxa = fmaxf(maxa, p);
            }
        float minb = bx[0] * nx + by[0] * ny, maxb = minb;
#pragma unroll
        for (int i = 1; i < KB; ++i)
            if (i < nb) {
                const float p = bx[i] * nx + by[i] * ny;
                minb = fminf(minb, p);
                maxb = fmaxf(maxb, p);
            }
        const float g1 = maxa - minb, g2 = maxb - mina;  // the reference needs both > 0
        sure_sep = sure_sep | (g1 < -T) | (g2 < -T);
        all_sure_ovl = all_sure_ovl & (g1 > T) & (g2 > T);
    };
    {
        float px = ax[0], py = ay[0];
#pragma unroll
        for (int i = 1; i < KA; ++i)
            if (i == na - 1) { px = ax[i]; py = ay[i]; }
#pragma unroll
        for (int i = 0; i < KA; ++i)
            if (i < na) {
                axis(-(ay[i] - py), ax[i] - px);
                px = ax[i];
                py = ay[i];
            }
    }
    {
        float px = bx[0], py = by[0];
#pragma unroll
        for (int i = 1; i < KB; ++i)
            if (i == nb - 1) { px = bx[i]; py = by[i]; }
#pragma unroll
        for (int i = 0; i < KB; ++i)
            if (i < nb) {
                axis(-(by[i] - py), bx[i] - px);
                px = bx[i];
                py = by[i];
            }
    }
    if (!in_range) return SAT_AMBIGUOUS;
    if (sure_sep) return SAT_REJECT;
    return all_sure_ovl ? SAT_ACCEPT : SAT_AMBIGUOUS;
}

// Bit-mask formulation of the same clip (what k_clip_survivors runs).
// For a subject without repeated vertices the sequential Sutherland-Hodgman
// loop is a pure function of the per-vertex `inside` flags:
//   * the in/out state flips where inside(j) != inside(j-1) (cyclically);
//     a convex subject has 0 or 2 flips
//   * vertex j lands at  popc(inside below j) + popc(flips at or below j),
//     the intersection of flip j right before it
// so a stage is: load the subject into registers, one `inside` evaluation per
// vertex, a handful of integer mask operations, the two intersections (the
// only divisions, evaluated together), and predicated stores to computed
// slots.  The code is straight-line: the two vertex lists have FIXED roles
// (stage i reads list i & 1 and writes the other one; an unchanged subject is
// copied through), so no per-lane buffer pointers are swapped, and a stage has
// no branch on the number of flips: lanes without a crossing run the same
// instructions with their stores predicated off and a safe denominator.
// Identical floating-point operations on the lanes that matter, identical
// results.  Executed instructions per pair (smsp__inst_executed, profiles/):
// 1 133 for the per-vertex loop this replaced, 880 for a branchy mask version,
// 750 here.  Lanes with a repeated subject vertex, more than two flips, a zero
// denominator or a zero-length clip edge set *fallback and are re-run through
// clip_area().
template <int KT, int KS>
__device__ __forceinline__ double clip_area_masks(const double2 (&poly)[KT], int np,
                                                   const double2 (&clip)[KS], int nc, double2 *bufA,
                                                   double2 *bufB, int stride, bool *fallback) {
    constexpr int M = KT + KS;
    int length = np;
#pragma unroll
    for (int i = 0; i < KT; ++i)
        if (i < np) bufA[i * stride] = poly[i];

    bool fb = false;
    double2 r = last_vertex(clip, nc);
#pragma unroll
    for (int i = 0; i < KS; ++i) {
        double2 *cur = (i & 1) ? bufB : bufA;
        double2 *oth = (i & 1) ? bufA : bufB;
        const double2 s = clip[i];
        const double Ux = s.x - r.x, Uy = s.y - r.y;
        const bool live = i < nc;  // (KS > nc only for a short face in a padded mesh)
        fb = fb || (live && Ux == 0.0 && Uy == 0.0);
        const double Nx = -Uy, Ny = Ux;
        // ---- subject into registers (stage 0: it is the caller's polygon)
        double2 v[M];
        double2 prev;
        if (i == 0) {
#pragma unroll
            for (int j = 0; j < KT; ++j) v[j] = poly[j];
            prev = last_vertex(poly, np);
        } else {
#pragma unroll
            for (int j = 0; j < M; ++j)
                if (j < KT + i) v[j] = cur[j * stride];  // (slots >= length hold stale data: masked below)
            prev = cur[(length > 0 ? length - 1 : 0) * stride];
        }
        unsigned inm = 0u;
        bool dup = false;
#pragma unroll
        for (int j = 0; j < M; ++j)
            if (j < KT + i) {
                const bool on = j < length;
                const bool in = on & (Ux * (v[j].y - r.y) > Uy * (v[j].x - r.x));
                inm |= (in ? 1u : 0u) << j;
                dup = dup | (on & (v[j].x == prev.x) & (v[j].y == prev.y));
                prev = v[j];
            }
        const unsigned full = length > 0 ? (1u << length) - 1u : 0u;
        const unsigned flipm = length > 0 ? (inm ^ ((inm << 1) | (inm >> (length - 1)))) & full : 0u;
        const int nflip = __popc(flipm);
        fb = fb || (live && (dup || nflip > 2));
        const bool cut = nflip == 2;
        const int cj0 = cut ? __ffs(flipm) - 1 : 0, cj1 = cut ? 31 - __clz(flipm) : 0;
        const int cp0 = cj0 ? cj0 - 1 : (length > 0 ? length - 1 : 0), cp1 = cj1 ? cj1 - 1 : 0;
        const double2 pa0 = cur[cp0 * stride], pb0 = cur[cj0 * stride];
        const double2 pa1 = cur[cp1 * stride], pb1 = cur[cj1 * stride];
        const double V0x = pb0.x - pa0.x, V0y = pb0.y - pa0.y;
        const double V1x = pb1.x - pa1.x, V1y = pb1.y - pa1.y;
        const double nw0 = Nx * (pa0.x - r.x) + Ny * (pa0.y - r.y);
        const double nw1 = Nx * (pa1.x - r.x) + Ny * (pa1.y - r.y);
        const double nv0 = Nx * V0x + Ny * V0y;
        const double nv1 = Nx * V1x + Ny * V1y;
        const bool zero_den = cut && (nv0 == 0.0 || nv1 == 0.0);
        fb = fb || (live && zero_den);
        const bool safe = cut && !zero_den;
        const double t0 = -nw0 / (safe ? nv0 : 1.0);
        const double t1 = -nw1 / (safe ? nv1 : 1.0);
        const double2 X0 = make_double2(pa0.x + t0 * V0x, pa0.y + t0 * V0y);
        const double2 X1 = make_double2(pa1.x + t1 * V1x, pa1.y + t1 * V1y);
        // ---- emit into the other list (a stage beyond nc copies the subject through)
        const unsigned keep = live ? inm : full;
        const unsigned fl = live ? flipm : 0u;
#pragma unroll
        for (int j = 0; j < M; ++j)
            if (j < KT + i && ((keep >> j) & 1u)) {
                const int pos = __popc(keep & ((1u << j) - 1u)) + __popc(fl & ((2u << j) - 1u));
                oth[pos * stride] = v[j];
            }
        if (live && cut) {
            oth[__popc(inm & ((1u << cj0) - 1u)) * stride] = X0;
            oth[(__popc(inm & ((1u << cj1) - 1u)) + 1) * stride] = X1;
        }
        length = __popc(keep) + (live && cut ? 2 : 0);
        if (live) r = s;
    }
    *fallback = fb;
    if (fb || length < 3) return 0.0;
    const double2 *fin = (KS & 1) ? bufB : bufA;
    double area = 0.0;
    const double2 a0 = fin[0];
    const double2 b0 = fin[stride];
    double Ux = b0.x - a0.x, Uy = b0.y - a0.y;
#pragma unroll
    for (int i = 2; i < M; ++i) {
        if (i < length) {
            const double2 c = fin[i * stride];
            const double Vx = a0.x - c.x, Vy = a0.y - c.y;
            area += fabs(cross2(Ux, Uy, Vx, Vy));
            Ux = Vx;
            Uy = Vy;
        }
    }
    return 0.5 * area;
}

// on-edge criterion: perpendicular distance to the edge line below `tol`,
// evaluated as cross^2 < tol^2 * |edge|^2 (tol2 = tol * tol)
__device__ __forceinline__ bool near_edge(double twice_area, double wx, double wy, double tol2) {
    return twice_area * twice_area < tol2 * (wx * wx + wy * wy);
}

__device__ __forceinline__ bool in_bounds(double px, double py, double ax, double ay, double bx,
                                          double by) {
    // /root/reference/xugrid/ugrid/burn.py:34-56
    const double dx = bx - ax, dy = by - ay;
    if (fabs(dx) >= fabs(dy)) {
        if (dx > 0) return ax <= px && px <= bx;
        return bx <= px && px <= ax;
    } else {
        if (dy > 0) return ay <= py && py <= by;
        return by <= py && py <= ay;
    }
}

// pnpoly ray casting; a point within tolerance of an edge counts as inside.
// The polygon is read through a pointer (global or shared), n vertices.
__device__ __forceinline__ bool point_in_polygon_or_on_edge(double px, double py,
                                                            const double2 *__restrict__ poly, int n,
                                                            double tol2) {
    double2 v0 = poly[n - 1];
    double Ux = v0.x - px, Uy = v0.y - py;
    bool c = false;
    for (int i = 0; i < n; ++i) {
        const double2 v1 = poly[i];
        const double Vx = v1.x - px, Vy = v1.y - py;
        const double twice_area = fabs(cross2(Ux, Uy, Vx, Vy));
        if (near_edge(twice_area, v1.x - v0.x, v1.y - v0.y, tol2)) {
            if (in_bounds(px, py, v0.x, v0.y, v1.x, v1.y)) return true;
        }
        if ((v0.y > py) != (v1.y > py) &&
            px < ((v1.x - v0.x) * (py - v0.y) / (v1.y - v0.y) + v0.x))
            c = !c;
        v0 = v1;
        Ux = Vx;
        Uy = Vy;
    }
    return c;
}

// triangle barycentric weights (areas of the sub-triangles)
__device__ __forceinline__ void bary_triangle(const double2 *__restrict__ t, double px, double py,
                                              double *w) {
    const double abx = t[1].x - t[0].x, aby = t[1].y - t[0].y;
    const double acx = t[2].x - t[0].x, acy = t[2].y - t[0].y;
    const double apx = px - t[0].x, apy = py - t[0].y;
    const double Aa = fabs(cross2(abx, aby, acx, acy));
    const double Ac = fabs(cross2(abx, aby, apx, apy));
    const double Ab = fabs(cross2(acx, acy, apx, apy));
    const double inv = 1.0 / Aa;
    const double wb = inv * Ab;
    const double wc = inv * Ac;
    w[0] = 1.0 - wb - wc;
    w[1] = wb;
    w[2] = wc;
}

__device__ __forceinline__ void interp_edge_case(double ax, double ay, double Ux, double Uy,
                                                 double px, double py, double *w, int n, int i,
                                                 int j) {
    const double Vx = px - ax, Vy = py - ay;
    const double f = sqrt(Vx * Vx + Vy * Vy) / sqrt(Ux * Ux + Uy * Uy);
    for (int q = 0; q < n; ++q) w[q] = 0.0;
    w[i] = 1.0 - f;
    w[j] = f;
}

// Wachspress coordinates w_i = 2 C_i / (A_{i-1} A_i), normalised; a point on
// (within tolerance of) an edge is interpolated linearly along that edge.
__device__ __forceinline__ void bary_wachspress(const double2 *__restrict__ poly, int n, double px,
                                                double py, double tol2, double *w) {
    double w_sum = 0.0;
    double2 a = poly[n - 1], b = poly[0];
    double Ux = b.x - a.x, Uy = b.y - a.y;
    double Ai = fabs(cross2(Ux, Uy, px - a.x, py - a.y));
    if (near_edge(Ai, Ux, Uy, tol2)) {
        interp_edge_case(a.x, a.y, Ux, Uy, px, py, w, n, n - 1, 0);
        return;
    }
    for (int i = 0; i < n; ++i) {
        const int i_next = (i + 1 == n) ? 0 : i + 1;
        const double2 c = poly[i_next];
        const double Tx = c.x - b.x, Ty = c.y - b.y;
        const double Aj = fabs(cross2(Tx, Ty, px - b.x, py - b.y));
        if (near_edge(Aj, Tx, Ty, tol2)) {
            interp_edge_case(b.x, b.y, Tx, Ty, px, py, w, n, i, i_next);
            return;
        }
        const double Ci = fabs(cross2(Ux, Uy, Tx, Ty));
        const double wi = 2.0 * Ci / (Ai * Aj);
        w[i] = wi;
        w_sum += wi;
        b = c;
        Ux = Tx;
        Uy = Ty;
        Ai = Aj;
    }
    for (int i = 0; i < n; ++i) w[i] /= w_sum;
}

}  // namespace xg
