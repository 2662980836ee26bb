/*
 * cwipc_oracle.c -- CPU restatement of cwipc_util's filter hot path (see cwipc_oracle.h).
 * TEST INFRASTRUCTURE ONLY; "parity unpinned" at the PCL boundary (see the header).
 *
 * Build: gcc -O2 -ffp-contract=off -fno-fast-math (no FMA contraction: the reference is built for
 * baseline x86-64, where float expressions round after every operation).
 *
 * Single-threaded on purpose: the reference never enables threading on this path
 * (src/cwipc_filters.cpp:135-140, 197-201), and this file doubles as the timed CPU baseline.
 */
#include "cwipc_oracle.h"

#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

/* The two restatements that depend on the PCL release ("lowest-confidence items") are
 * compile-time switches; the defaults are what the CUDA library implements and what the parity tests
 * assert.  `make` also builds libcwipc_oracle_alt.so with both flipped, and tests/test_oracle.py checks
 * what each switch may and may not change (see orc_config()).
 *   ORC_LEAF_ORDER_DESCENDING  0: octree leaves are visited children 0..7 (PCL >= 1.9, DFS = ascending
 *                                 Morton code); 1: children 7..0 (older releases) -- only the ORDER of the
 *                                 per-leaf blocks in the output changes, never their contents.
 *   ORC_VOXEL_SORT_UNSTABLE    0: points of one voxel are summed in input order (a stable sort);
 *                              1: in reversed order (stands in for boost spreadsort / std::sort, whose
 *                                 order among equal keys is unspecified) -- only the float rounding of the
 *                                 centroid changes (inside the 1e-5 tolerance), colours/tiles/counts do not. */
#ifndef ORC_LEAF_ORDER_DESCENDING
#define ORC_LEAF_ORDER_DESCENDING 0
#endif
#ifndef ORC_VOXEL_SORT_UNSTABLE
#define ORC_VOXEL_SORT_UNSTABLE 0
#endif
int orc_config(void) { return (ORC_LEAF_ORDER_DESCENDING ? 1 : 0) | (ORC_VOXEL_SORT_UNSTABLE ? 2 : 0); }

/* ------------------------------------------------------------------------------------------ */
/* helpers                                                                                      */
/* ------------------------------------------------------------------------------------------ */

typedef struct {
    uint64_t key;
    uint32_t idx;
} keyidx;

/* stable LSD radix sort by key (stands in for boost::sort::spreadsort, which is not stable; a stable
 * order makes the float summation order -- unspecified in the reference -- the input order) */
static void sort_keyidx(keyidx *a, size_t n, int keybits) {
    if (n < 2) return;
    keyidx *tmp = (keyidx *)malloc(n * sizeof(keyidx));
    keyidx *src = a, *dst = tmp;
    for (int shift = 0; shift < keybits; shift += 8) {
        size_t count[257];
        memset(count, 0, sizeof(count));
        for (size_t i = 0; i < n; i++) count[((src[i].key >> shift) & 0xff) + 1]++;
        for (int d = 0; d < 256; d++) count[d + 1] += count[d];
        for (size_t i = 0; i < n; i++) dst[count[(src[i].key >> shift) & 0xff]++] = src[i];
        keyidx *t = src;
        src = dst;
        dst = t;
    }
    if (src != a) memcpy(a, src, n * sizeof(keyidx));
    free(tmp);
}

static int bit_length64(uint64_t v) {
    int b = 0;
    while (v) {
        b++;
        v >>= 1;
    }
    return b;
}

/* ------------------------------------------------------------------------------------------ */
/* tilefilter                                                     ref: src/cwipc_filters.cpp:281-306 */
/* ------------------------------------------------------------------------------------------ */
long orc_tilefilter(const orc_point *in, size_t n, int tile, orc_point *out) {
    long m = 0;
    for (size_t i = 0; i < n; i++) {
        if (tile == 0 || tile == (int)in[i].tile) out[m++] = in[i]; /* `tile == pt.a`: int vs uint8 */
    }
    return m;
}

/* ------------------------------------------------------------------------------------------ */
/* cellsize heuristic                                              ref: src/cwipc_util.cpp:173-204 */
/* prevPoint is never advanced, so this is min_i |p_i - p_0|; 0 when fewer than two points.      */
/* pcl::geometry::distance = (p1 - p2).norm() on float vectors.                                  */
/* ------------------------------------------------------------------------------------------ */
float orc_min_distance_to_first(const orc_point *in, size_t n) {
    float best = INFINITY;
    for (size_t i = 1; i < n; i++) {
        float dx = in[i].x - in[0].x, dy = in[i].y - in[0].y, dz = in[i].z - in[0].z;
        float xx = dx * dx, yy = dy * dy, zz = dz * dz;
        float s = xx + yy;
        s = s + zz;
        float d = sqrtf(s);
        if (d < best) best = d;
    }
    return best == INFINITY ? 0.0f : best;
}

/* ------------------------------------------------------------------------------------------ */
/* one pcl::VoxelGrid over the points in.idx[0..m)   ref: pcl filters/impl/voxel_grid.hpp applyFilter, */
/* common/impl/accumulators.hpp (AccumulatorXYZ, AccumulatorRGBA); call sites                         */
/* src/cwipc_filters.cpp:52-74 and :135-155 (tile = OR of the contributing tiles)                     */
/* Returns the number of voxels appended to out, or -1 on index overflow (single-grid only).          */
/* ------------------------------------------------------------------------------------------ */
static long voxelgrid_subset(const orc_point *in, const uint32_t *subset, size_t m, float cs, orc_point *out, uint32_t *out_counts) {
    if (m == 0) return 0;
    const float inv = 1.0f / cs;
    float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    for (size_t j = 0; j < m; j++) {
        const orc_point *p = &in[subset[j]];
        const float c[3] = {p->x, p->y, p->z};
        for (int a = 0; a < 3; a++) {
            if (c[a] < mn[a]) mn[a] = c[a];
            if (c[a] > mx[a]) mx[a] = c[a];
        }
    }
    /* "Leaf size is too small for the input dataset": PCL copies the input, leaves leaf_layout_ empty and
     * cwipc's tile pass then throws std::out_of_range -> ERROR + NULL (src/cwipc_filters.cpp:70-80) */
    int64_t d[3];
    for (int a = 0; a < 3; a++) {
        float ext = mx[a] - mn[a];
        float scaled = ext * inv;
        d[a] = (int64_t)scaled + 1;
    }
    if (d[0] * d[1] * d[2] > (int64_t)INT32_MAX) return -1;
    int minb[3], divb[3];
    uint64_t cells = 1;
    for (int a = 0; a < 3; a++) {
        float lo = mn[a] * inv, hi = mx[a] * inv;
        minb[a] = (int)floorf(lo);
        divb[a] = (int)floorf(hi) - minb[a] + 1;
        cells *= (uint64_t)divb[a];
    }
    keyidx *kv = (keyidx *)malloc(m * sizeof(keyidx));
    for (size_t j = 0; j < m; j++) {
        const orc_point *p = &in[subset[j]];
        float fx = p->x * inv, fy = p->y * inv, fz = p->z * inv;
        /* ijk = (int)(floor(x * inv) - (float)min_b) */
        int i0 = (int)(floorf(fx) - (float)minb[0]);
        int i1 = (int)(floorf(fy) - (float)minb[1]);
        int i2 = (int)(floorf(fz) - (float)minb[2]);
        int idx = i0 + i1 * divb[0] + i2 * divb[0] * divb[1];
        kv[j].key = (uint32_t)idx;
        kv[j].idx = subset[j];
    }
    {
        int bits = bit_length64(cells - 1);
        sort_keyidx(kv, m, bits > 0 ? bits : 1);
    }
#if ORC_VOXEL_SORT_UNSTABLE
    for (size_t a = 0; a < m;) { /* reverse every run of equal keys */
        size_t e = a + 1;
        while (e < m && kv[e].key == kv[a].key) e++;
        for (size_t lo = a, hi = e - 1; lo < hi; lo++, hi--) {
            keyidx t = kv[lo];
            kv[lo] = kv[hi];
            kv[hi] = t;
        }
        a = e;
    }
#endif
    long nout = 0;
    size_t j = 0;
    while (j < m) {
        size_t e = j + 1;
        while (e < m && kv[e].key == kv[j].key) e++;
        /* pcl::CentroidPoint: float accumulators, sequential adds in sorted order */
        float sx = 0, sy = 0, sz = 0, sr = 0, sg = 0, sb = 0;
        unsigned tile = 0;
        for (size_t t = j; t < e; t++) {
            const orc_point *p = &in[kv[t].idx];
            sx += p->x;
            sy += p->y;
            sz += p->z;
            sr += (float)p->r;
            sg += (float)p->g;
            sb += (float)p->b;
            tile |= p->tile;
        }
        const float fn = (float)(e - j);
        orc_point o;
        o.x = sx / fn;
        o.y = sy / fn;
        o.z = sz / fn;
        o.r = (uint8_t)(uint32_t)(sr / fn); /* truncation, as static_cast<std::uint32_t>(r / n) */
        o.g = (uint8_t)(uint32_t)(sg / fn);
        o.b = (uint8_t)(uint32_t)(sb / fn);
        o.tile = (uint8_t)tile;
        if (out_counts) out_counts[nout] = (uint32_t)(e - j);
        out[nout++] = o;
        j = e;
    }
    free(kv);
    return nout;
}

/* interleave: bit b of x -> bit 3b+2, y -> 3b+1, z -> 3b  (octree child index = x<<2 | y<<1 | z) */
static uint64_t morton3(uint32_t x, uint32_t y, uint32_t z) {
    uint64_t code = 0;
    for (int b = 0; b < 21; b++) {
        code |= (uint64_t)((x >> b) & 1u) << (3 * b + 2);
        code |= (uint64_t)((y >> b) & 1u) << (3 * b + 1);
        code |= (uint64_t)((z >> b) & 1u) << (3 * b);
    }
    return code;
}

/* ------------------------------------------------------------------------------------------ */
/* cwipc_downsample                                                ref: src/cwipc_filters.cpp:30-172 */
/* ------------------------------------------------------------------------------------------ */
long orc_downsample(const orc_point *in, size_t n, float voxelsize, float pc_cellsize, orc_point *out, float *out_cellsize, int32_t *point_keys, uint32_t *out_counts) {
    const int single_grid = voxelsize < 0;
    float cs = single_grid ? -voxelsize : voxelsize;
    if (pc_cellsize >= cs) cs = pc_cellsize; /* :42-46, :103-107 */
    if (out_cellsize) *out_cellsize = cs;
    const float inv = 1.0f / cs;

    if (point_keys) {
        for (size_t i = 0; i < n; i++) {
            float fx = in[i].x * inv, fy = in[i].y * inv, fz = in[i].z * inv;
            point_keys[6 * i + 0] = point_keys[6 * i + 1] = point_keys[6 * i + 2] = 0;
            point_keys[6 * i + 3] = (int32_t)floorf(fx);
            point_keys[6 * i + 4] = (int32_t)floorf(fy);
            point_keys[6 * i + 5] = (int32_t)floorf(fz);
        }
    }

    if (single_grid) {
        if (n == 0) return -1; /* "VoxelGrid filter produced empty pointcloud" -> NULL (:58-62) */
        uint32_t *all = (uint32_t *)malloc(n * sizeof(uint32_t));
        for (size_t i = 0; i < n; i++) all[i] = (uint32_t)i;
        long rv = voxelgrid_subset(in, all, n, cs, out, out_counts);
        free(all);
        return rv;
    }
    if (n == 0) return 0; /* no leaves: empty, non-NULL cloud (python/test_cwipc_util.py:589-594) */

    /* ---- octree bounding box, pcl OctreePointCloud::adoptBoundingBoxToPoint / getKeyBitSize ---- */
    const float octree_cellsize = 64 * cs; /* :113-114, float */
    const double res = (double)octree_cellsize;
    const double eps = (double)FLT_EPSILON;
    double mn[3], mx[3];
    int depth = 1;
    {
        const float c[3] = {in[0].x, in[0].y, in[0].z};
        for (int a = 0; a < 3; a++) {
            mn[a] = (double)c[a] - res / 2;
            mx[a] = (double)c[a] + res / 2;
        }
        const double side = (double)(1 << 1) * res; /* max_voxels = 2 -> depth 1 */
        for (int a = 0; a < 3; a++) {
            double oversize = (side - (mx[a] - mn[a])) / 2.0;
            if (oversize > eps) {
                mn[a] -= oversize;
                mx[a] += oversize;
            }
        }
    }
    for (size_t i = 1; i < n; i++) {
        const double q[3] = {(double)in[i].x, (double)in[i].y, (double)in[i].z};
        for (;;) {
            int upper[3], any = 0;
            for (int a = 0; a < 3; a++) {
                upper[a] = q[a] >= mx[a];
                if (upper[a] || q[a] < mn[a]) any = 1;
            }
            if (!any) break;
            if (depth >= 30) return -1; /* runaway growth: non-finite input */
            double side = (double)(1 << depth) * res;
            for (int a = 0; a < 3; a++)
                if (!upper[a]) mn[a] -= side;
            depth++;
            side = (double)(1 << depth) * res - eps;
            for (int a = 0; a < 3; a++) mx[a] = mn[a] + side;
        }
    }

    /* ---- leaf key per point (genOctreeKeyforPoint, double arithmetic), leaves in DFS = Morton order ---- */
    keyidx *kv = (keyidx *)malloc(n * sizeof(keyidx));
    for (size_t i = 0; i < n; i++) {
        uint32_t lx = (uint32_t)(((double)in[i].x - mn[0]) / res);
        uint32_t ly = (uint32_t)(((double)in[i].y - mn[1]) / res);
        uint32_t lz = (uint32_t)(((double)in[i].z - mn[2]) / res);
        kv[i].key = morton3(lx, ly, lz);
#if ORC_LEAF_ORDER_DESCENDING
        kv[i].key = (((uint64_t)1 << (3 * depth)) - 1) - kv[i].key; /* children visited 7..0 at every level */
#endif
        kv[i].idx = (uint32_t)i;
        if (point_keys) {
            point_keys[6 * i + 0] = (int32_t)lx;
            point_keys[6 * i + 1] = (int32_t)ly;
            point_keys[6 * i + 2] = (int32_t)lz;
        }
    }
    sort_keyidx(kv, n, 3 * depth); /* stable: points keep input order inside a leaf */

    /* ---- a fresh VoxelGrid per leaf, results appended (:124-158) ---- */
    uint32_t *subset = (uint32_t *)malloc(n * sizeof(uint32_t));
    long nout = 0;
    size_t j = 0;
    while (j < n) {
        size_t e = j + 1;
        while (e < n && kv[e].key == kv[j].key) e++;
        for (size_t t = j; t < e; t++) subset[t - j] = kv[t].idx;
        long got = voxelgrid_subset(in, subset, e - j, cs, out + nout, out_counts ? out_counts + nout : NULL);
        if (got < 0) { /* cannot happen: a leaf spans 64 voxels per axis */
            free(subset);
            free(kv);
            return -1;
        }
        nout += got;
        j = e;
    }
    free(subset);
    free(kv);
    return nout;
}

/* ------------------------------------------------------------------------------------------ */
/* exact k nearest neighbours: kd-tree (median split on the widest axis, small leaves), float     */
/* L2_Simple distances, branch pruning in double.  Stands in for pcl::KdTreeFLANN / FLANN          */
/* KDTreeSingleIndex with eps = 0: both return the k+1 smallest distances, which is all that        */
/* StatisticalOutlierRemoval consumes.                                                              */
/* ------------------------------------------------------------------------------------------ */
#define KD_LEAF 12

typedef struct {
    int axis;       /* -1: leaf */
    float split;    /* points with coord <= split go left (ties may be on both sides: handled by pruning with >=) */
    uint32_t begin, end; /* leaf: range in perm */
    int left, right;
    float lo[3], hi[3]; /* bounding box of the node's points */
} kdnode;

typedef struct {
    const orc_point *pts;
    uint32_t *perm;
    float *px, *py, *pz; /* coordinates in perm order (cache friendly leaves) */
    kdnode *nodes;
    int nnodes, capnodes;
} kdtree;

static float coord(const orc_point *p, int a) { return a == 0 ? p->x : (a == 1 ? p->y : p->z); }

static void nth_element_perm(const orc_point *pts, uint32_t *perm, long lo, long hi, long nth, int axis) {
    /* quickselect on perm[lo..hi] */
    while (lo < hi) {
        float pivot = coord(&pts[perm[(lo + hi) / 2]], axis);
        long i = lo, j = hi;
        while (i <= j) {
            while (coord(&pts[perm[i]], axis) < pivot) i++;
            while (coord(&pts[perm[j]], axis) > pivot) j--;
            if (i <= j) {
                uint32_t t = perm[i];
                perm[i] = perm[j];
                perm[j] = t;
                i++;
                j--;
            }
        }
        if (nth <= j) hi = j;
        else if (nth >= i) lo = i;
        else return;
    }
}

static int kd_build(kdtree *t, uint32_t begin, uint32_t end) {
    if (t->nnodes == t->capnodes) {
        t->capnodes = t->capnodes ? t->capnodes * 2 : 1024;
        t->nodes = (kdnode *)realloc(t->nodes, (size_t)t->capnodes * sizeof(kdnode));
    }
    const int id = t->nnodes++;
    float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    for (uint32_t i = begin; i < end; i++) {
        for (int a = 0; a < 3; a++) {
            float c = coord(&t->pts[t->perm[i]], a);
            if (c < lo[a]) lo[a] = c;
            if (c > hi[a]) hi[a] = c;
        }
    }
    int axis = 0;
    for (int a = 1; a < 3; a++)
        if (hi[a] - lo[a] > hi[axis] - lo[axis]) axis = a;
    kdnode nd;
    memcpy(nd.lo, lo, sizeof(lo));
    memcpy(nd.hi, hi, sizeof(hi));
    nd.begin = begin;
    nd.end = end;
    nd.left = nd.right = -1;
    nd.split = 0;
    if (end - begin <= KD_LEAF || !(hi[axis] > lo[axis])) {
        nd.axis = -1;
        t->nodes[id] = nd;
        return id;
    }
    const uint32_t mid = begin + (end - begin) / 2;
    nth_element_perm(t->pts, t->perm, (long)begin, (long)end - 1, (long)mid, axis);
    nd.axis = axis;
    nd.split = coord(&t->pts[t->perm[mid]], axis);
    t->nodes[id] = nd;
    const int l = kd_build(t, begin, mid);
    const int r = kd_build(t, mid, end);
    t->nodes[id].left = l;
    t->nodes[id].right = r;
    return id;
}

/* max-heap of the kk smallest squared distances seen so far */
typedef struct {
    float *d;
    int size, cap;
} maxheap;

static void heap_push(maxheap *h, float v) {
    if (h->size < h->cap) {
        int i = h->size++;
        h->d[i] = v;
        while (i > 0 && h->d[(i - 1) / 2] < h->d[i]) {
            float t = h->d[i];
            h->d[i] = h->d[(i - 1) / 2];
            h->d[(i - 1) / 2] = t;
            i = (i - 1) / 2;
        }
    } else if (v < h->d[0]) {
        int i = 0;
        h->d[0] = v;
        for (;;) {
            int l = 2 * i + 1, r = l + 1, m = i;
            if (l < h->size && h->d[l] > h->d[m]) m = l;
            if (r < h->size && h->d[r] > h->d[m]) m = r;
            if (m == i) break;
            float t = h->d[i];
            h->d[i] = h->d[m];
            h->d[m] = t;
            i = m;
        }
    }
}

/* FLANN L2_Simple<float>: result += diff*diff over x, y, z, all in float */
static float dist2f(float ax, float ay, float az, float bx, float by, float bz) {
    float dx = ax - bx, dy = ay - by, dz = az - bz;
    float xx = dx * dx, yy = dy * dy, zz = dz * dz;
    float s = xx + yy;
    s = s + zz;
    return s;
}

/* squared distance from q to a box, in double, slightly shrunk so that pruning never discards a point
 * whose float distance could tie with the current worst */
static double box_dist2(const kdnode *nd, const float q[3]) {
    double s = 0;
    for (int a = 0; a < 3; a++) {
        double d = 0;
        if (q[a] < nd->lo[a]) d = (double)nd->lo[a] - q[a];
        else if (q[a] > nd->hi[a]) d = (double)q[a] - nd->hi[a];
        s += d * d;
    }
    return s * (1.0 - 1e-6);
}

static void kd_query(const kdtree *t, int id, const float q[3], maxheap *h) {
    const kdnode *nd = &t->nodes[id];
    if (h->size == h->cap && box_dist2(nd, q) > (double)h->d[0]) return;
    if (nd->axis < 0) {
        for (uint32_t i = nd->begin; i < nd->end; i++) heap_push(h, dist2f(q[0], q[1], q[2], t->px[i], t->py[i], t->pz[i]));
        return;
    }
    if (q[nd->axis] <= nd->split) {
        kd_query(t, nd->left, q, h);
        kd_query(t, nd->right, q, h);
    } else {
        kd_query(t, nd->right, q, h);
        kd_query(t, nd->left, q, h);
    }
}

static int cmp_float(const void *a, const void *b) {
    float x = *(const float *)a, y = *(const float *)b;
    return (x > y) - (x < y);
}

/* dist_sum += sqrt(nn_dists[k]) for k = 1..mean_k in ascending order (double), then (float)(sum / k)
 * ref: pcl filters/impl/statistical_outlier_removal.hpp (first pass) */
static float mean_from_sorted(const float *d2, int kk, int k) {
    double sum = 0.0;
    for (int j = 1; j < kk; j++) sum += sqrt((double)d2[j]);
    return (float)(sum / (double)k);
}

int orc_knn_mean_distances(const orc_point *in, size_t n, int k, float *mean_dist) {
    if (k < 1 || n <= (size_t)k) return -1;
    kdtree t;
    memset(&t, 0, sizeof(t));
    t.pts = in;
    t.perm = (uint32_t *)malloc(n * sizeof(uint32_t));
    for (size_t i = 0; i < n; i++) t.perm[i] = (uint32_t)i;
    kd_build(&t, 0, (uint32_t)n);
    t.px = (float *)malloc(n * sizeof(float));
    t.py = (float *)malloc(n * sizeof(float));
    t.pz = (float *)malloc(n * sizeof(float));
    for (size_t i = 0; i < n; i++) {
        t.px[i] = in[t.perm[i]].x;
        t.py[i] = in[t.perm[i]].y;
        t.pz[i] = in[t.perm[i]].z;
    }
    const int kk = k + 1;
    maxheap h;
    h.d = (float *)malloc((size_t)kk * sizeof(float));
    h.cap = kk;
    /* query in tree order: neighbouring queries touch the same nodes */
    for (size_t j = 0; j < n; j++) {
        const uint32_t i = t.perm[j];
        const float q[3] = {in[i].x, in[i].y, in[i].z};
        h.size = 0;
        kd_query(&t, 0, q, &h);
        qsort(h.d, (size_t)kk, sizeof(float), cmp_float);
        mean_dist[i] = mean_from_sorted(h.d, kk, k);
    }
    free(h.d);
    free(t.px);
    free(t.py);
    free(t.pz);
    free(t.perm);
    free(t.nodes);
    return 0;
}

int orc_knn_mean_distances_bruteforce(const orc_point *in, size_t n, int k, float *mean_dist) {
    if (k < 1 || n <= (size_t)k) return -1;
    const int kk = k + 1;
    maxheap h;
    h.d = (float *)malloc((size_t)kk * sizeof(float));
    h.cap = kk;
    for (size_t i = 0; i < n; i++) {
        h.size = 0;
        for (size_t j = 0; j < n; j++) heap_push(&h, dist2f(in[i].x, in[i].y, in[i].z, in[j].x, in[j].y, in[j].z));
        qsort(h.d, (size_t)kk, sizeof(float), cmp_float);
        mean_dist[i] = mean_from_sorted(h.d, kk, k);
    }
    free(h.d);
    return 0;
}

/* one pcl::StatisticalOutlierRemoval::applyFilterIndices over in[0..n): appends inliers to out */
static long sor_group(const orc_point *in, size_t n, int k, float stddev_mul, orc_point *out, double *threshold_out) {
    if (n == 0) return 0;
    if (k < 1 || n <= (size_t)k) {
        /* undefined behaviour in the reference (FLANN returns fewer than k+1 neighbours, PCL reads k+1);
         * defined by this project, for the library and the oracle alike, as keep-everything */
        memcpy(out, in, n * sizeof(orc_point));
        return (long)n;
    }
    float *d = (float *)malloc(n * sizeof(float));
    orc_knn_mean_distances(in, n, k, d);
    double sum = 0, sq_sum = 0;
    for (size_t i = 0; i < n; i++) {
        float sq = d[i] * d[i]; /* `distance * distance` is a float product */
        sum += d[i];
        sq_sum += sq;
    }
    const double nv = (double)n;
    const double mean = sum / nv;
    const double variance = (sq_sum - sum * sum / nv) / (nv - 1);
    const double stddev = sqrt(variance);
    const double threshold = mean + (double)stddev_mul * stddev;
    if (threshold_out) *threshold_out = threshold;
    long m = 0;
    for (size_t i = 0; i < n; i++) {
        if (d[i] > threshold) continue; /* float promoted to double; NaN threshold keeps everything */
        out[m++] = in[i];
    }
    free(d);
    return m;
}

/* ------------------------------------------------------------------------------------------ */
/* cwipc_remove_outliers                                           ref: src/cwipc_filters.cpp:222-278 */
/* ------------------------------------------------------------------------------------------ */
long orc_remove_outliers(const orc_point *in, size_t n, int k, float stddev_mul, int per_tile, orc_point *out, double *threshold_out) {
    if (!per_tile) return sor_group(in, n, k, stddev_mul, out, threshold_out);
    /* distinct tile values in first-appearance order (:239-249) */
    int tiles[256], ntiles = 0;
    int seen[256];
    memset(seen, 0, sizeof(seen));
    for (size_t i = 0; i < n; i++) {
        if (!seen[in[i].tile]) {
            seen[in[i].tile] = 1;
            tiles[ntiles++] = in[i].tile;
        }
    }
    orc_point *group = (orc_point *)malloc((n ? n : 1) * sizeof(orc_point));
    long total = 0;
    for (int t = 0; t < ntiles; t++) {
        /* cwipc_tilefilter(pc, tile): tile 0 selects EVERY point (:251-256, :296) */
        long cnt = orc_tilefilter(in, n, tiles[t], group);
        total += sor_group(group, (size_t)cnt, k, stddev_mul, out + total, NULL);
    }
    free(group);
    return total;
}
