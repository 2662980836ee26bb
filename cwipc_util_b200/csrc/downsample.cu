// downsample.cu -- voxel-grid downsample as hand-written sm_100a kernels.
//
// Reference semantics (src/cwipc_filters.cpp:30-172 over pcl::VoxelGrid, pcl::CentroidPoint and
// pcl::octree::OctreePointCloud, restated in SURVEY.md App. A.3/A.4 and in oracle/cwipc_oracle.c):
//   * voxel of a point:  (int)floorf(x * inv) per axis with inv = 1.0f / cellsize (float arithmetic);
//   * positive voxelsize: the cloud is first split by an octree of resolution 64*cellsize whose origin
//     is anchored on the FIRST point (leaf faces at p0 + m*res, double arithmetic); every leaf gets its
//     own voxel grid, so a voxel cut by a leaf face is emitted once per leaf.  Output order = leaves in
//     depth-first (Morton, x most significant) order, voxels inside a leaf by (z, y, x);
//   * negative voxelsize: one grid over the whole cloud, output ordered by (z, y, x);
//   * per voxel: xyz = mean, rgb = (uint)(float sum / n) (truncation), tile = OR of the tiles.
//
// Pipeline (all on one stream):
//   chunk_bbox_kernel   16 B/pt read      per-1024-point bounding boxes
//   octree_box_kernel   1 block           global bbox + exact replay of the octree's sequential
//                                         bounding-box growth (first violating point, grow, repeat)
//   voxel_keygen_kernel 16 B read + 8 B   64-bit key = [leaf Morton | voxel-in-leaf | point index]
//   radix_sort_u64      P x (8+8) B       stable LSD sort on the key bits only (index rides along)
//   voxel_reduce_kernel 8 + 16 B read     runs of equal keys -> one output point; decoupled look-back
//                       16 B/voxel write  numbers the runs, partial runs at tile edges are merged by the
//                                         last block.  Sums are 64-bit fixed point, so the result does
//                                         not depend on the order of accumulation (deterministic).
// Algorithmic bytes: 16*N in + 16*V' out.
#include <algorithm>
#include <cmath>
#include <cstring>

#include "device_utils.cuh"
#include "kernels.hpp"
#include "radix_sort.hpp"

namespace cwcu {

namespace {

constexpr int BB_CHUNK = 1024;   // points per bounding-box chunk
constexpr int BB_THREADS = 256;
constexpr int WL_RADIX = 72;     // voxel-in-leaf coordinate range per axis (64 + misalignment + guard)
constexpr int WL_BITS = 19;      // 72^3 = 373248 < 2^19
constexpr int MAX_OCTREE_DEPTH = 14;

struct OctreeBox {
    double min[3];
    double max[3];
    float gmin[3];
    float gmax[3];
    int depth;
    int error; // 1: runaway growth (non-finite input)
};

// ---- per-chunk bounding boxes --------------------------------------------------------------
__global__ void __launch_bounds__(BB_THREADS) chunk_bbox_kernel(const cwipc_point *__restrict__ pts, uint32_t n, float *__restrict__ chunk_bbox) {
    __shared__ float s_red[6][BB_THREADS / 32];
    const uint32_t base = blockIdx.x * BB_CHUNK;
    float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
#pragma unroll
    for (int j = 0; j < BB_CHUNK / BB_THREADS; j++) {
        const uint32_t i = base + j * BB_THREADS + threadIdx.x;
        if (i < n) {
            const Point16 p = ld_point_stream(pts, i);
            lo[0] = fminf(lo[0], p.x); hi[0] = fmaxf(hi[0], p.x);
            lo[1] = fminf(lo[1], p.y); hi[1] = fmaxf(hi[1], p.y);
            lo[2] = fminf(lo[2], p.z); hi[2] = fmaxf(hi[2], p.z);
        }
    }
#pragma unroll
    for (int a = 0; a < 3; a++) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lo[a] = fminf(lo[a], __shfl_xor_sync(FULL_MASK, lo[a], o));
            hi[a] = fmaxf(hi[a], __shfl_xor_sync(FULL_MASK, hi[a], o));
        }
    }
    const unsigned warp = threadIdx.x >> 5;
    if (lane_id() == 0) {
#pragma unroll
        for (int a = 0; a < 3; a++) {
            s_red[a][warp] = lo[a];
            s_red[3 + a][warp] = hi[a];
        }
    }
    __syncthreads();
    if (threadIdx.x < 6) {
        float v = s_red[threadIdx.x][0];
        for (int w = 1; w < BB_THREADS / 32; w++) v = threadIdx.x < 3 ? fminf(v, s_red[threadIdx.x][w]) : fmaxf(v, s_red[threadIdx.x][w]);
        chunk_bbox[(size_t)blockIdx.x * 6 + threadIdx.x] = v;
    }
}

// ---- global bbox + octree bounding box replay ------------------------------------------------
// PCL grows the octree box sequentially while inserting points (OctreePointCloud::adoptBoundingBoxToPoint):
// a point outside [min, max) doubles the box, moving `min` down by the old side on every axis the point
// does not exceed upward.  The final `min` (hence leaf keys and leaf visiting order) depends on the
// input order, so the replay below finds, in order, each point that violates the current box.
__device__ __forceinline__ bool violates(const double *mn, const double *mx, float lx, float ly, float lz, float hx, float hy, float hz) {
    return (double)lx < mn[0] || (double)ly < mn[1] || (double)lz < mn[2] || (double)hx >= mx[0] || (double)hy >= mx[1] || (double)hz >= mx[2];
}

__global__ void __launch_bounds__(1024) octree_box_kernel(const cwipc_point *__restrict__ pts, uint32_t n, const float *__restrict__ chunk_bbox, uint32_t nchunks, double res,
                                                           int do_octree, OctreeBox *__restrict__ out) {
    __shared__ float s_red[6][32];
    __shared__ double s_min[3], s_max[3];
    __shared__ int s_depth, s_error;
    __shared__ uint32_t s_cursor, s_first;
    const unsigned tid = threadIdx.x, warp = tid >> 5, lane = lane_id();

    // global bounding box
    float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (uint32_t c = tid; c < nchunks; c += 1024) {
#pragma unroll
        for (int a = 0; a < 3; a++) {
            lo[a] = fminf(lo[a], chunk_bbox[(size_t)c * 6 + a]);
            hi[a] = fmaxf(hi[a], chunk_bbox[(size_t)c * 6 + 3 + a]);
        }
    }
#pragma unroll
    for (int a = 0; a < 3; a++) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lo[a] = fminf(lo[a], __shfl_xor_sync(FULL_MASK, lo[a], o));
            hi[a] = fmaxf(hi[a], __shfl_xor_sync(FULL_MASK, hi[a], o));
        }
        if (lane == 0) {
            s_red[a][warp] = lo[a];
            s_red[3 + a][warp] = hi[a];
        }
    }
    __syncthreads();
    if (tid < 6) {
        float v = s_red[tid][0];
        for (int w = 1; w < 32; w++) v = tid < 3 ? fminf(v, s_red[tid][w]) : fmaxf(v, s_red[tid][w]);
        if (tid < 3) out->gmin[tid] = v;
        else out->gmax[tid - 3] = v;
    }
    if (!do_octree) {
        if (tid == 0) { out->depth = 0; out->error = 0; }
        return;
    }

    const double eps = (double)1.1920929e-07f; // std::numeric_limits<float>::epsilon(), promoted as in PCL
    if (tid == 0) {
        const Point16 p0 = ld_point(pts, 0);
        const float c[3] = {p0.x, p0.y, p0.z};
        for (int a = 0; a < 3; a++) {
            s_min[a] = (double)c[a] - res / 2;
            s_max[a] = (double)c[a] + res / 2;
        }
        // getKeyBitSize() on the empty tree: depth 1, box padded symmetrically to side 2*res
        const double side = (double)(1 << 1) * res;
        for (int a = 0; a < 3; a++) {
            const double oversize = (side - (s_max[a] - s_min[a])) / 2.0;
            if (oversize > eps) {
                s_min[a] -= oversize;
                s_max[a] += oversize;
            }
        }
        s_depth = 1;
        s_error = 0;
        s_cursor = 1;
    }
    while (true) {
        __syncthreads();
        if (tid == 0) s_first = 0xffffffffu;
        __syncthreads();
        const uint32_t cursor = s_cursor;
        // first chunk at or after the cursor whose box sticks out
        for (uint32_t c = cursor / BB_CHUNK + tid; c < nchunks; c += 1024) {
            const float *b = chunk_bbox + (size_t)c * 6;
            if (violates(s_min, s_max, b[0], b[1], b[2], b[3], b[4], b[5])) {
                atomicMin(&s_first, c);
                break;
            }
        }
        __syncthreads();
        const uint32_t cstar = s_first;
        if (cstar == 0xffffffffu) break;
        __syncthreads();
        if (tid == 0) s_first = 0xffffffffu;
        __syncthreads();
        // first violating point of that chunk, not before the cursor (BB_CHUNK == blockDim.x)
        const uint32_t i = cstar * BB_CHUNK + tid;
        if (i >= cursor && i < n) {
            const Point16 p = ld_point(pts, i);
            if (violates(s_min, s_max, p.x, p.y, p.z, p.x, p.y, p.z)) atomicMin(&s_first, i);
        }
        __syncthreads();
        const uint32_t istar = s_first;
        if (tid == 0) {
            if (istar == 0xffffffffu) {
                s_cursor = (cstar + 1) * BB_CHUNK; // only already-replayed points of this chunk stick out
            } else {
                const Point16 p = ld_point(pts, istar);
                const double q[3] = {(double)p.x, (double)p.y, (double)p.z};
                while (true) {
                    bool upper[3], any = false;
                    for (int a = 0; a < 3; a++) {
                        upper[a] = q[a] >= s_max[a];
                        any = any || upper[a] || q[a] < s_min[a];
                    }
                    if (!any) break;
                    if (s_depth >= 30) {
                        s_error = 1;
                        break;
                    }
                    double side = (double)(1 << s_depth) * res;
                    for (int a = 0; a < 3; a++)
                        if (!upper[a]) s_min[a] -= side;
                    s_depth++;
                    side = (double)(1 << s_depth) * res - eps;
                    for (int a = 0; a < 3; a++) s_max[a] = s_min[a] + side;
                }
                s_cursor = istar + 1;
            }
        }
        __syncthreads();
        if (s_error) break;
    }
    if (tid == 0) {
        for (int a = 0; a < 3; a++) {
            out->min[a] = s_min[a];
            out->max[a] = s_max[a];
        }
        out->depth = s_depth;
        out->error = s_error;
    }
}

// ---- key generation --------------------------------------------------------------------------
struct KeyParams {
    int octree;       // 1: [morton | voxel-in-leaf], 0: PCL's linear voxel index
    float inv;        // 1.0f / cellsize
    int idxbits;
    // octree mode
    double omin[3];
    double res;
    double inv_cs;    // 1.0 / (double)cellsize
    int depth;
    // single-grid mode
    int minb[3];
    int div[3];
};

__device__ __forceinline__ uint64_t spread3(uint32_t v) { // bit i -> bit 3i, v < 2^21
    uint64_t x = v & 0x1fffffu;
    x = (x | x << 32) & 0x1f00000000ffffull;
    x = (x | x << 16) & 0x1f0000ff0000ffull;
    x = (x | x << 8) & 0x100f00f00f00f00full;
    x = (x | x << 4) & 0x10c30c30c30c30c3ull;
    x = (x | x << 2) & 0x1249249249249249ull;
    return x;
}

// Returns the sort key WITHOUT index bits; *bad is set when a coordinate is out of the supported range.
__device__ __forceinline__ uint64_t voxel_key(const Point16 &p, const KeyParams &kp, bool *bad) {
    const float f[3] = {floorf(__fmul_rn(p.x, kp.inv)), floorf(__fmul_rn(p.y, kp.inv)), floorf(__fmul_rn(p.z, kp.inv))};
    if (!(fabsf(f[0]) < 4194304.f && fabsf(f[1]) < 4194304.f && fabsf(f[2]) < 4194304.f)) {
        *bad = true; // |voxel coordinate| >= 2^22 (or NaN): float voxel arithmetic is no longer exact
        return 0;
    }
    if (!kp.octree) {
        // ijk = (int)(floor(x*inv) - (float)min_b); idx = i + j*dx + k*dx*dy   (pcl voxel_grid.hpp)
        const int64_t i = (int)(f[0] - (float)kp.minb[0]), j = (int)(f[1] - (float)kp.minb[1]), k = (int)(f[2] - (float)kp.minb[2]);
        return (uint64_t)(i + j * (int64_t)kp.div[0] + k * (int64_t)kp.div[0] * (int64_t)kp.div[1]);
    }
    const float c[3] = {p.x, p.y, p.z};
    uint32_t leaf[3], w[3];
#pragma unroll
    for (int a = 0; a < 3; a++) {
        // OctreePointCloud::genOctreeKeyforPoint: (unsigned)((double)x - min) / res), double arithmetic
        const double rel = ((double)c[a] - kp.omin[a]) / kp.res;
        leaf[a] = (uint32_t)rel;
        // Any per-leaf constant keeps the (z,y,x) order inside a leaf; this one is <= every voxel
        // coordinate that can occur in the leaf, and the leaf spans < 70 voxels.
        const int origin = (int)floor((kp.omin[a] + (double)leaf[a] * kp.res) * kp.inv_cs) - 2;
        const int wi = (int)f[a] - origin;
        if (wi < 0 || wi >= WL_RADIX || rel < 0.0 || leaf[a] >= (1u << kp.depth)) *bad = true;
        w[a] = (uint32_t)min(max(wi, 0), WL_RADIX - 1);
    }
    const uint64_t morton = (spread3(leaf[0]) << 2) | (spread3(leaf[1]) << 1) | spread3(leaf[2]);
    const uint64_t wlin = ((uint64_t)w[2] * WL_RADIX + w[1]) * WL_RADIX + w[0];
    return (morton << WL_BITS) | wlin;
}

__global__ void __launch_bounds__(256) voxel_keygen_kernel(const cwipc_point *__restrict__ pts, uint32_t n, KeyParams kp, uint64_t *__restrict__ keys, int with_index,
                                                            uint32_t *__restrict__ error_flag) {
    bool bad = false;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const Point16 p = ld_point_stream(pts, i);
        const uint64_t k = voxel_key(p, kp, &bad);
        keys[i] = with_index ? ((k << kp.idxbits) | i) : k;
    }
    if (bad) atomicOr(error_flag, 1u);
}

// ---- segmented reduction of sorted runs --------------------------------------------------------
constexpr int VR_THREADS = 256;
constexpr int VR_ITEMS = 4;
constexpr int VR_TILE = VR_THREADS * VR_ITEMS; // 1024 sorted elements per tile

struct VoxelAgg { // 40 bytes
    unsigned long long sx, sy, sz; // two's-complement fixed point sums
    unsigned long long rgb;        // r | g<<21 | b<<42  (per-tile partials: 255*1024 < 2^21)
    uint32_t n;
    uint32_t tile;
};

struct TileRecord { // written by every tile, read by the fix-up
    uint32_t base;      // global number of the tile's first run
    uint32_t heads;     // runs starting in this tile
    uint32_t tail_open; // last run continues in the next tile
    uint32_t pad;
};

struct WideAgg { // carries across tiles can exceed the packed rgb field: keep r,g,b apart
    long long sx, sy, sz;
    unsigned long long r, g, b;
    unsigned long long n;
    uint32_t tile;
    uint32_t pad;
};

__device__ __forceinline__ void wide_add(WideAgg &w, const VoxelAgg &a) {
    w.sx += (long long)a.sx;
    w.sy += (long long)a.sy;
    w.sz += (long long)a.sz;
    w.r += a.rgb & 0x1fffffull;
    w.g += (a.rgb >> 21) & 0x1fffffull;
    w.b += (a.rgb >> 42) & 0x1fffffull;
    w.n += a.n;
    w.tile |= a.tile;
}

// mean xyz (exact sum, one rounding), truncated float colour average as pcl::CentroidPoint, OR of tiles
__device__ __forceinline__ Point16 finalize_voxel(const WideAgg &w, double inv_scale) {
    Point16 o;
    const double dn = (double)w.n;
    o.x = (float)(((double)w.sx * inv_scale) / dn);
    o.y = (float)(((double)w.sy * inv_scale) / dn);
    o.z = (float)(((double)w.sz * inv_scale) / dn);
    const float fn = (float)w.n;
    const uint32_t r = (uint32_t)__fdiv_rn((float)w.r, fn) & 0xffu;
    const uint32_t g = (uint32_t)__fdiv_rn((float)w.g, fn) & 0xffu;
    const uint32_t b = (uint32_t)__fdiv_rn((float)w.b, fn) & 0xffu;
    o.rgbt = r | (g << 8) | (b << 16) | ((w.tile & 0xffu) << 24);
    return o;
}

__global__ void __launch_bounds__(VR_THREADS) voxel_reduce_kernel(const uint64_t *__restrict__ sorted, uint32_t n, int idxbits, const cwipc_point *__restrict__ pts,
                                                                   double scale, double inv_scale, cwipc_point *__restrict__ out, uint32_t *__restrict__ ticket,
                                                                   uint64_t *__restrict__ status, TileRecord *__restrict__ records, VoxelAgg *__restrict__ carry_head,
                                                                   VoxelAgg *__restrict__ carry_tail, uint32_t *__restrict__ done_counter, uint32_t ntiles,
                                                                   uint32_t *__restrict__ d_total) {
    __shared__ VoxelAgg s_slot[VR_TILE + 1]; // slot 0: continuation of the previous tile's run; slot j: j-th run starting here
    __shared__ uint32_t s_warp[VR_THREADS / 32];
    __shared__ int s_tile;
    __shared__ uint32_t s_base;
    __shared__ bool s_last_block;

    if (threadIdx.x == 0) s_tile = take_ticket(ticket);
    __syncthreads();
    const int tile = s_tile;
    const uint32_t tile_base = (uint32_t)tile * VR_TILE;
    const unsigned warp = threadIdx.x >> 5, lane = lane_id();
    const uint64_t idxmask = (1ull << idxbits) - 1ull;

    // ---- keys of my 4 consecutive elements, head flags ----
    const uint32_t e0 = tile_base + threadIdx.x * VR_ITEMS;
    uint64_t k[VR_ITEMS];
    bool head[VR_ITEMS];
    uint64_t prev = (e0 > 0 && e0 <= n) ? (sorted[e0 - 1] >> idxbits) : ~0ull;
    uint32_t nheads = 0;
#pragma unroll
    for (int j = 0; j < VR_ITEMS; j++) {
        const uint32_t e = e0 + j;
        if (e < n) {
            k[j] = sorted[e];
            const uint64_t kk = k[j] >> idxbits;
            head[j] = (e == 0) || (kk != prev);
            prev = kk;
            nheads += head[j] ? 1u : 0u;
        } else {
            k[j] = 0;
            head[j] = false;
        }
    }
    // block exclusive scan of head counts
    const uint32_t incl = warp_inclusive_scan(nheads);
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    uint32_t warp_off = 0, H = 0;
#pragma unroll
    for (int w = 0; w < VR_THREADS / 32; w++) {
        const uint32_t t = s_warp[w];
        if (w < (int)warp) warp_off += t;
        H += t;
    }
    uint32_t slot = warp_off + incl - nheads; // heads before my first element == slot of a non-head first element

    // zero the slots in use
    for (uint32_t j = threadIdx.x; j <= H; j += VR_THREADS) {
        s_slot[j].sx = 0; s_slot[j].sy = 0; s_slot[j].sz = 0; s_slot[j].rgb = 0; s_slot[j].n = 0; s_slot[j].tile = 0;
    }
    // chained numbering of runs across tiles (warp 0), overlapped with the gathers below
    __syncthreads();

    // ---- gather points, accumulate sequentially, flush a partial whenever a run ends ----
    Point16 p[VR_ITEMS];
#pragma unroll
    for (int j = 0; j < VR_ITEMS; j++)
        if (e0 + j < n) p[j] = ld_point(pts, (size_t)(k[j] & idxmask));

    long long ax = 0, ay = 0, az = 0;
    unsigned long long argb = 0;
    uint32_t an = 0, at = 0;
    auto flush = [&](uint32_t s) {
        if (an) {
            atomicAdd(&s_slot[s].sx, (unsigned long long)ax);
            atomicAdd(&s_slot[s].sy, (unsigned long long)ay);
            atomicAdd(&s_slot[s].sz, (unsigned long long)az);
            atomicAdd(&s_slot[s].rgb, argb);
            atomicAdd(&s_slot[s].n, an);
            atomicOr(&s_slot[s].tile, at);
        }
        ax = ay = az = 0;
        argb = 0;
        an = at = 0;
    };
#pragma unroll
    for (int j = 0; j < VR_ITEMS; j++) {
        if (e0 + j < n) {
            if (head[j]) {
                flush(slot);
                slot++;
            }
            ax += __double2ll_rn((double)p[j].x * scale);
            ay += __double2ll_rn((double)p[j].y * scale);
            az += __double2ll_rn((double)p[j].z * scale);
            argb += (unsigned long long)pt_r(p[j]) | ((unsigned long long)pt_g(p[j]) << 21) | ((unsigned long long)pt_b(p[j]) << 42);
            an += 1;
            at |= pt_tile(p[j]);
        }
    }
    flush(slot);

    if (warp == 0) {
        const uint32_t excl = lookback_exclusive(status, tile, H);
        if (lane == 0) s_base = excl;
    }
    __syncthreads();
    const uint32_t base = s_base;

    // does the run that is open at the end of this tile continue in the next one?
    const uint32_t tile_end = min(tile_base + VR_TILE, n);
    bool tail_open = false;
    if (tile_end < n) tail_open = (sorted[tile_end] >> idxbits) == (sorted[tile_end - 1] >> idxbits);

    // ---- emit complete runs; park partial ones for the fix-up ----
    for (uint32_t j = threadIdx.x; j <= H; j += VR_THREADS) {
        const VoxelAgg a = s_slot[j];
        if (j == 0) {
            carry_head[tile] = a; // may be empty (n == 0)
        } else if (j == H && tail_open) {
            carry_tail[tile] = a;
        } else {
            WideAgg w = {0, 0, 0, 0, 0, 0, 0, 0, 0};
            wide_add(w, a);
            st_point(out, base + j - 1, finalize_voxel(w, inv_scale));
        }
    }
    if (H == 0 && threadIdx.x == 0) {
        // whole tile continues an older run: nothing starts here
    }
    if (threadIdx.x == 0) {
        TileRecord r;
        r.base = base;
        r.heads = H;
        r.tail_open = (H > 0 && tail_open) ? 1u : 0u;
        r.pad = 0;
        records[tile] = r;
        if (tile_end >= n) *d_total = base + H;
    }

    // ---- last block to finish merges the runs that cross tile boundaries ----
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last_block = (atomicAdd(done_counter, 1u) == ntiles - 1);
    __syncthreads();
    if (!s_last_block) return;
    __threadfence();
    for (uint32_t t = threadIdx.x; t < ntiles; t += VR_THREADS) {
        const TileRecord r = records[t];
        if (!r.tail_open) continue;
        WideAgg w = {0, 0, 0, 0, 0, 0, 0, 0, 0};
        wide_add(w, carry_tail[t]);
        for (uint32_t u = t + 1; u < ntiles; u++) {
            wide_add(w, carry_head[u]);
            if (records[u].heads != 0) break; // the run ends inside tile u
        }
        st_point(out, r.base + r.heads - 1, finalize_voxel(w, inv_scale));
    }
}

int bit_length(uint64_t v) {
    int b = 0;
    while (v) {
        b++;
        v >>= 1;
    }
    return b;
}

struct Plan {
    KeyParams kp;
    int keybits = 0;
    bool failed = false;
    std::string error;
    float maxabs = 0.f;
};

// Runs the two bounding-box kernels, reads the box back (one sync) and derives the key layout.
Plan make_plan(const cwipc_point *pts, size_t n, float cellsize, bool octree_split, int dev, cudaStream_t s) {
    Plan plan;
    const uint32_t nchunks = (uint32_t)div_up(n, BB_CHUNK);
    Scratch chunk_bbox((size_t)nchunks * 6 * sizeof(float), s);
    Scratch box(sizeof(OctreeBox), s);
    const float octree_cellsize = 64 * cellsize;           // ref: src/cwipc_filters.cpp:113-114 (float)
    const double res = (double)octree_cellsize;
    launch("chunk_bbox_kernel", s, 16 * (size_t)n, [&] { chunk_bbox_kernel<<<nchunks, BB_THREADS, 0, s>>>(pts, (uint32_t)n, chunk_bbox.as<float>()); });
    launch("octree_box_kernel", s, 24 * (size_t)nchunks, [&] {
        octree_box_kernel<<<1, 1024, 0, s>>>(pts, (uint32_t)n, chunk_bbox.as<float>(), nchunks, res, octree_split ? 1 : 0, box.as<OctreeBox>());
    });
    OctreeBox *h = static_cast<OctreeBox *>(thread_pinned(sizeof(OctreeBox)));
    CWCU_CHECK(cudaMemcpyAsync(h, box.p, sizeof(OctreeBox), cudaMemcpyDeviceToHost, s));
    CWCU_CHECK(cudaStreamSynchronize(s));
    const OctreeBox ob = *h;
    (void)dev;

    KeyParams &kp = plan.kp;
    memset(&kp, 0, sizeof(kp));
    kp.octree = octree_split ? 1 : 0;
    kp.inv = 1.0f / cellsize;
    kp.idxbits = std::max(1, bit_length((uint64_t)n - 1));
    for (int a = 0; a < 3; a++) plan.maxabs = std::max(plan.maxabs, std::max(std::fabs(ob.gmin[a]), std::fabs(ob.gmax[a])));
    if (!std::isfinite(plan.maxabs)) {
        plan.failed = true;
        plan.error = "pointcloud contains non-finite coordinates";
        return plan;
    }
    if (octree_split) {
        if (ob.error || ob.depth > MAX_OCTREE_DEPTH) {
            plan.failed = true;
            plan.error = "pointcloud extent too large for voxel size (octree depth " + std::to_string(ob.depth) + " > " + std::to_string(MAX_OCTREE_DEPTH) + ")";
            return plan;
        }
        for (int a = 0; a < 3; a++) kp.omin[a] = ob.min[a];
        kp.res = res;
        kp.inv_cs = 1.0 / (double)cellsize;
        kp.depth = ob.depth;
        plan.keybits = WL_BITS + 3 * ob.depth;
    } else {
        // ref: pcl VoxelGrid::applyFilter -- index-overflow guard, then min_b / div_b from the float bbox
        const float inv = kp.inv;
        const int64_t dx = (int64_t)((ob.gmax[0] - ob.gmin[0]) * inv) + 1;
        const int64_t dy = (int64_t)((ob.gmax[1] - ob.gmin[1]) * inv) + 1;
        const int64_t dz = (int64_t)((ob.gmax[2] - ob.gmin[2]) * inv) + 1;
        if (dx * dy * dz > (int64_t)INT32_MAX) {
            // PCL copies the input and leaves its leaf layout empty; cwipc's tile pass then throws out_of_range
            plan.failed = true;
            plan.error = "VoxelGrid std exception: leaf size is too small for the input dataset (integer indices would overflow)";
            return plan;
        }
        uint64_t cells = 1;
        for (int a = 0; a < 3; a++) {
            kp.minb[a] = (int)std::floor(ob.gmin[a] * inv);
            const int maxb = (int)std::floor(ob.gmax[a] * inv);
            kp.div[a] = maxb - kp.minb[a] + 1;
            cells *= (uint64_t)kp.div[a];
        }
        plan.keybits = std::max(1, bit_length(cells - 1));
    }
    if (plan.keybits + kp.idxbits > 64) {
        plan.failed = true;
        plan.error = "pointcloud too large for 64-bit voxel keys (" + std::to_string(plan.keybits) + " key bits + " + std::to_string(kp.idxbits) + " index bits)";
    }
    return plan;
}

unsigned stream_grid(size_t n, int dev) {
    return (unsigned)std::max<size_t>(1, std::min(div_up(n, 256), (size_t)sm_count(dev) * 8));
}

} // namespace

DownsampleResult downsample_points(const StoragePtr &in, float cellsize, bool octree_split, int dev, cudaStream_t s) {
    DownsampleResult result;
    const size_t n = in->count;
    if (n == 0) {
        if (octree_split) {
            // no octree leaves: an empty, non-NULL cloud (python/test_cwipc_util.py:589-594)
            result.out = std::make_shared<Storage>(dev, 0, s);
            result.out->mark_ready();
        } else {
            result.failed = true; // ref: src/cwipc_filters.cpp:58-62
            result.error = "VoxelGrid filter produced empty pointcloud";
        }
        return result;
    }
    if (!(cellsize > 0.f) || !std::isfinite(cellsize)) {
        result.failed = true;
        result.error = "invalid voxel size " + std::to_string(cellsize);
        return result;
    }
    Plan plan = make_plan(in->d_pts, n, cellsize, octree_split, dev, s);
    if (plan.failed) {
        result.failed = true;
        result.error = plan.error;
        return result;
    }
    const KeyParams &kp = plan.kp;

    Scratch keys_a(n * sizeof(uint64_t), s), keys_b(n * sizeof(uint64_t), s);
    Scratch flag(sizeof(uint32_t), s);
    CWCU_CHECK(cudaMemsetAsync(flag.p, 0, sizeof(uint32_t), s));
    launch("voxel_keygen_kernel", s, 24 * (size_t)n, [&] { voxel_keygen_kernel<<<stream_grid(n, dev), 256, 0, s>>>(in->d_pts, (uint32_t)n, kp, keys_a.as<uint64_t>(), 1, flag.as<uint32_t>()); });
    uint64_t *sorted = radix_sort_u64(keys_a.as<uint64_t>(), keys_b.as<uint64_t>(), n, kp.idxbits, kp.idxbits + plan.keybits, dev, s);

    // fixed-point scale: |x| * 2^shift * n < 2^62
    int shift = 62 - bit_length((uint64_t)n);
    if (plan.maxabs > 0.f) {
        int e;
        (void)std::frexp(plan.maxabs, &e); // maxabs < 2^e
        shift -= e;
    }
    shift = std::max(-60, std::min(shift, 100));
    const double scale = std::ldexp(1.0, shift), inv_scale = std::ldexp(1.0, -shift);

    const uint32_t ntiles = (uint32_t)div_up(n, VR_TILE);
    // [ticket | done | total | pad] u32*4, status u64*ntiles, records, carry_head, carry_tail
    const size_t off_status = 16;
    const size_t off_records = off_status + (size_t)ntiles * sizeof(uint64_t);
    const size_t off_head = off_records + (size_t)ntiles * sizeof(TileRecord);
    const size_t off_tail = off_head + (size_t)ntiles * sizeof(VoxelAgg);
    const size_t bytes = off_tail + (size_t)ntiles * sizeof(VoxelAgg);
    Scratch aux(bytes, s);
    CWCU_CHECK(cudaMemsetAsync(aux.p, 0, off_records, s));
    uint8_t *ab = aux.as<uint8_t>();
    uint32_t *ticket = reinterpret_cast<uint32_t *>(ab);
    uint32_t *done = ticket + 1, *d_total = ticket + 2;

    auto out = std::make_shared<Storage>(dev, n, s);
    launch("voxel_reduce_kernel", s, 24 * (size_t)n, [&] {
        voxel_reduce_kernel<<<ntiles, VR_THREADS, 0, s>>>(sorted, (uint32_t)n, kp.idxbits, in->d_pts, scale, inv_scale, out->d_pts, ticket,
                                                           reinterpret_cast<uint64_t *>(ab + off_status), reinterpret_cast<TileRecord *>(ab + off_records),
                                                           reinterpret_cast<VoxelAgg *>(ab + off_head), reinterpret_cast<VoxelAgg *>(ab + off_tail), done, ntiles, d_total);
    });
    uint32_t *h = static_cast<uint32_t *>(thread_pinned(2 * sizeof(uint32_t)));
    CWCU_CHECK(cudaMemcpyAsync(h, d_total, sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
    CWCU_CHECK(cudaMemcpyAsync(h + 1, flag.p, sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
    CWCU_CHECK(cudaStreamSynchronize(s));
    if (h[1] != 0) {
        result.failed = true;
        result.error = "point coordinates out of range for voxel size " + std::to_string(cellsize) + " (|x/voxelsize| must stay below 2^22)";
        return result;
    }
    out->count = h[0];
    profile_add_bytes("voxel_reduce_kernel", 16 * (size_t)h[0]); // voxels written
    out->mark_ready();
    result.out = out;
    return result;
}

void global_bbox(const cwipc_point *in, size_t n, float gmin[3], float gmax[3], int dev, cudaStream_t s) {
    (void)dev;
    const uint32_t nchunks = (uint32_t)div_up(n, BB_CHUNK);
    Scratch chunk_bbox((size_t)nchunks * 6 * sizeof(float), s);
    Scratch box(sizeof(OctreeBox), s);
    launch("chunk_bbox_kernel", s, 16 * (size_t)n, [&] { chunk_bbox_kernel<<<nchunks, BB_THREADS, 0, s>>>(in, (uint32_t)n, chunk_bbox.as<float>()); });
    launch("octree_box_kernel", s, 24 * (size_t)nchunks, [&] { octree_box_kernel<<<1, 1024, 0, s>>>(in, (uint32_t)n, chunk_bbox.as<float>(), nchunks, 1.0, 0, box.as<OctreeBox>()); });
    OctreeBox *h = static_cast<OctreeBox *>(thread_pinned(sizeof(OctreeBox)));
    CWCU_CHECK(cudaMemcpyAsync(h, box.p, sizeof(OctreeBox), cudaMemcpyDeviceToHost, s));
    CWCU_CHECK(cudaStreamSynchronize(s));
    for (int a = 0; a < 3; a++) {
        gmin[a] = h->gmin[a];
        gmax[a] = h->gmax[a];
    }
}

void downsample_keys_to_host(const StoragePtr &in, float cellsize, bool octree_split, uint64_t *host_keys, int dev, cudaStream_t s) {
    const size_t n = in->count;
    if (n == 0) return;
    Plan plan = make_plan(in->d_pts, n, cellsize, octree_split, dev, s);
    if (plan.failed) throw CudaError{cudaErrorInvalidValue, plan.error};
    Scratch keys(n * sizeof(uint64_t), s);
    Scratch flag(sizeof(uint32_t), s);
    CWCU_CHECK(cudaMemsetAsync(flag.p, 0, sizeof(uint32_t), s));
    launch("voxel_keygen_kernel", s, 24 * (size_t)n, [&] { voxel_keygen_kernel<<<stream_grid(n, dev), 256, 0, s>>>(in->d_pts, (uint32_t)n, plan.kp, keys.as<uint64_t>(), 0, flag.as<uint32_t>()); });
    CWCU_CHECK(cudaMemcpyAsync(host_keys, keys.p, n * sizeof(uint64_t), cudaMemcpyDeviceToHost, s));
    CWCU_CHECK(cudaStreamSynchronize(s));
}

} // namespace cwcu
