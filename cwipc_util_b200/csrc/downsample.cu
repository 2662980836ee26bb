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

constexpr int WL_RADIX = 72;     // voxel-in-leaf coordinate range per axis (64 + misalignment + guard)
constexpr int WL_BITS = 19;      // 72^3 = 373248 < 2^19
constexpr int MAX_OCTREE_DEPTH = 14;

struct OctreeSeed { // box state to continue from (a cloud partitioned over several GPUs is replayed part by part)
    double min[3];
    double max[3];
    int depth;
    int valid; // 0: start from this cloud's first point
};

struct OctreeBox {
    double min[3];
    double max[3];
    float gmin[3];
    float gmax[3];
    int depth;
    int error; // 1: runaway growth (non-finite input)
};

// ---- per-chunk bounding boxes (device function; the kernel follows the octree replay below) ------------
constexpr int BB_THREADS = 1024;
constexpr int BB_ITEMS = 4;
constexpr int BB_CHUNK = BB_THREADS * BB_ITEMS; // points per bounding-box chunk (one block)
__device__ __forceinline__ void chunk_bbox_block(const cwipc_point *__restrict__ pts, uint32_t n, float *chunk_bbox) {
    __shared__ float s_cb[6][BB_THREADS / 32];
    float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
#pragma unroll
    for (int j = 0; j < BB_ITEMS; j++) {
        const uint32_t i = blockIdx.x * BB_CHUNK + j * BB_THREADS + threadIdx.x;
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
            s_cb[a][warp] = lo[a];
            s_cb[3 + a][warp] = hi[a];
        }
    }
    __syncthreads();
    if (threadIdx.x < 6) {
        float v = s_cb[threadIdx.x][0];
        for (int w = 1; w < BB_THREADS / 32; w++) v = threadIdx.x < 3 ? fminf(v, s_cb[threadIdx.x][w]) : fmaxf(v, s_cb[threadIdx.x][w]);
        __stcg(chunk_bbox + (size_t)blockIdx.x * 6 + threadIdx.x, v);
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

__device__ __forceinline__ void octree_box_block(const cwipc_point *__restrict__ pts, uint32_t n, const float *chunk_bbox, uint32_t nchunks, double res, int do_octree,
                                                  const OctreeSeed &seed, OctreeBox *__restrict__ out) {
    __shared__ float s_red[6][32];
    __shared__ double s_min[3], s_max[3];
    __shared__ int s_depth, s_error;
    __shared__ uint32_t s_cursor, s_first;
    const unsigned tid = threadIdx.x, warp = tid >> 5, lane = lane_id();

    // global bounding box
    float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (uint32_t c = tid; c < nchunks; c += BB_THREADS) {
#pragma unroll
        for (int a = 0; a < 3; a++) {
            lo[a] = fminf(lo[a], __ldcg(chunk_bbox + (size_t)c * 6 + a));
            hi[a] = fmaxf(hi[a], __ldcg(chunk_bbox + (size_t)c * 6 + 3 + a));
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
    if (tid == 0 && seed.valid) {
        for (int a = 0; a < 3; a++) {
            s_min[a] = seed.min[a];
            s_max[a] = seed.max[a];
        }
        s_depth = seed.depth;
        s_error = 0;
        s_cursor = 0;
    } else if (tid == 0) {
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
        for (uint32_t c = cursor / BB_CHUNK + tid; c < nchunks; c += BB_THREADS) {
            const float *b = chunk_bbox + (size_t)c * 6;
            if (violates(s_min, s_max, __ldcg(b), __ldcg(b + 1), __ldcg(b + 2), __ldcg(b + 3), __ldcg(b + 4), __ldcg(b + 5))) {
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
        // first violating point of that chunk, not before the cursor
#pragma unroll
        for (int j = 0; j < BB_ITEMS; j++) {
            const uint32_t i = cstar * BB_CHUNK + j * BB_THREADS + tid;
            if (i >= cursor && i < n) {
                const Point16 p = ld_point(pts, i);
                if (violates(s_min, s_max, p.x, p.y, p.z, p.x, p.y, p.z)) atomicMin(&s_first, i);
            }
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

// One launch: every block boxes its 1024-point chunk; the block that finishes last reduces the chunk boxes
// and replays the octree growth (the classic "last block" pattern: fence, ticket, fence).
__global__ void __launch_bounds__(BB_THREADS) bbox_octree_kernel(const cwipc_point *__restrict__ pts, uint32_t n, float *chunk_bbox, uint32_t nchunks, double res, int do_octree,
                                                                OctreeSeed seed, uint32_t *__restrict__ done_counter, OctreeBox *__restrict__ out) {
    __shared__ bool s_last;
    chunk_bbox_block(pts, n, chunk_bbox);
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const uint32_t ticket = atomicAdd(done_counter, 1u);
        s_last = ticket == nchunks - 1;
        if (s_last) *done_counter = 0; // leave the counter clean for the next call
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    octree_box_block(pts, n, chunk_bbox, nchunks, res, do_octree, seed, out);
}

// ---- key generation --------------------------------------------------------------------------
struct KeyParams {
    int octree;       // 1: [morton | voxel-in-leaf], 0: PCL's linear voxel index
    float inv;        // 1.0f / cellsize
    int idxbits;
    // octree mode
    double omin[3];
    double res;
    double inv_res;   // 1.0 / res (filter only; ties are decided by the exact division)
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
// Leaf lookup tables (octree mode, at most 32 leaves per axis, i.e. clouds up to ~2000 voxels across): the
// leaf index of a coordinate is a step function of the float32 value, so the host finds, with the very double
// arithmetic of the generic path, the smallest float at which every step happens.  A point then costs a
// 5-compare binary search per axis instead of double-precision subtract / multiply / floor / compare.
constexpr int KT_LEAVES = 32;
struct KeyTables {
    float thr[3][KT_LEAVES];        // thr[a][m]: smallest float whose leaf index is first[a] + m + 1 (+inf beyond the last leaf)
    int origin[3][KT_LEAVES];       // voxel-coordinate origin of leaf first[a] + m
    uint64_t spread[3][KT_LEAVES];  // its Morton bits, already in the axis' position
};

__device__ __forceinline__ uint64_t voxel_key_tables(const Point16 &p, const KeyParams &kp, const KeyTables &tab, bool *bad) {
    const float f[3] = {floorf(__fmul_rn(p.x, kp.inv)), floorf(__fmul_rn(p.y, kp.inv)), floorf(__fmul_rn(p.z, kp.inv))};
    if (!(fabsf(f[0]) < 4194304.f && fabsf(f[1]) < 4194304.f && fabsf(f[2]) < 4194304.f)) {
        *bad = true;
        return 0;
    }
    const float c[3] = {p.x, p.y, p.z};
    uint64_t morton = 0;
    uint32_t w[3];
#pragma unroll
    for (int a = 0; a < 3; a++) {
        int m = 0; // number of thresholds <= c[a]
#pragma unroll
        for (int step = KT_LEAVES / 2; step > 0; step >>= 1)
            if (c[a] >= tab.thr[a][m + step - 1]) m += step;
        const int wi = (int)f[a] - tab.origin[a][m];
        if (wi < 0 || wi >= WL_RADIX) *bad = true;
        w[a] = (uint32_t)min(max(wi, 0), WL_RADIX - 1);
        morton |= tab.spread[a][m];
    }
    const uint64_t wlin = ((uint64_t)w[2] * WL_RADIX + w[1]) * WL_RADIX + w[0];
    return (morton << WL_BITS) | wlin;
}

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
        // The quotient is only needed truncated: multiply by the reciprocal, and redo it with the exact
        // division whenever the product lands within 1e-9 of a leaf face (the two differ by < 1e-11).
        const double num = (double)c[a] - kp.omin[a];
        double rel = num * kp.inv_res;
        const double fr = rel - floor(rel);
        if (!(fr > 1e-9 && fr < 1.0 - 1e-9)) rel = num / kp.res;
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

// ---- accumulation: one pass over the points into a hash table of voxels -------------------------------
// Every voxel owns one 64-byte slot (one L2 line sector pair): key, fixed-point coordinate sums, colour
// sums, point count and tile OR.  A warp first merges runs of equal keys among its 32 consecutive
// points with a segmented shuffle scan (camera and scan-order clouds are spatially coherent, so runs
// are long), then the last lane of every run finds or claims the slot (64-bit CAS, linear probing)
// and adds its partial sums with fire-and-forget L2 atomics (RED).  Sums are integers, so the result
// does not depend on the order in which points arrive (deterministic).  The thread that claims a
// slot appends [key | slot] to the list the radix sort orders afterwards.
struct __align__(64) VoxelSlot {
    unsigned long long keyp1;      // key + 1; 0 = empty
    unsigned long long sx, sy, sz; // two's-complement fixed-point sums
    unsigned long long rg;         // sum r | sum g << 32
    unsigned long long bn;         // sum b | count << 32
    uint32_t tile;                 // OR of the tile bytes
    uint32_t pad[3];
};
static_assert(sizeof(VoxelSlot) == 64, "one slot per 64 bytes");

struct TableHeader { // first 64 bytes of the zeroed workspace (layout of the whole head: runtime.hpp, ZW_HEADER_BYTES)
    uint32_t count; // claimed slots
    uint32_t error; // out-of-range coordinate seen
    uint32_t pad[14]; // pad[0]: last-block ticket of bbox_octree_kernel; pad[2..3]: barrier words of radix_fused_kernel; pad[4]: stats ticket
};

__device__ __forceinline__ uint32_t hash_key(uint64_t k) {
    k ^= k >> 33;
    k *= 0xff51afd7ed558ccdull;
    k ^= k >> 33;
    k *= 0xc4ceb9fe1a85ec53ull;
    k ^= k >> 33;
    return (uint32_t)k;
}

constexpr int VA_THREADS = 256;

template <bool TABLES>
__global__ void __launch_bounds__(VA_THREADS) voxel_accumulate_kernel(const cwipc_point *__restrict__ pts, uint32_t n, KeyParams kp, const __grid_constant__ KeyTables g_tab, float scale,
                                                                       VoxelSlot *__restrict__ table, uint32_t slot_mask, int slotbits, uint64_t *__restrict__ list,
                                                                       TableHeader *__restrict__ header) {
    __shared__ KeyTables s_tab;
    if (TABLES) {
        const uint32_t *src = reinterpret_cast<const uint32_t *>(&g_tab); // kernel parameter (constant bank) -> shared
        uint32_t *dst = reinterpret_cast<uint32_t *>(&s_tab);
        for (uint32_t i = threadIdx.x; i < sizeof(KeyTables) / 4; i += VA_THREADS) dst[i] = src[i];
        __syncthreads();
    }
    bool bad = false;
    const unsigned lane = lane_id();
    const unsigned le = lanemask_lt() | (1u << lane);
    const uint32_t warps_total = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t base = ((blockIdx.x * blockDim.x + threadIdx.x) >> 5) * 32u; base < n; base += warps_total * 32u) {
        const uint32_t i = base + lane;
        const bool valid = i < n;
        uint64_t key = ~0ull;
        long long fx = 0, fy = 0, fz = 0;
        unsigned long long rg = 0, bn = 0;
        uint32_t tl = 0;
        if (valid) {
            const Point16 p = ld_point_stream(pts, i);
            key = TABLES ? voxel_key_tables(p, kp, s_tab, &bad) : voxel_key(p, kp, &bad);
            // scale is a power of two: the float product is exact, so this is round(x * 2^s) like the double path
            fx = __float2ll_rn(__fmul_rn(p.x, scale));
            fy = __float2ll_rn(__fmul_rn(p.y, scale));
            fz = __float2ll_rn(__fmul_rn(p.z, scale));
            rg = (unsigned long long)pt_r(p) | ((unsigned long long)pt_g(p) << 32);
            bn = (unsigned long long)pt_b(p) | (1ull << 32);
            tl = pt_tile(p);
        }
        // runs of equal keys along the lanes, cut every 8 lanes -> inclusive segmented sums in 3 shuffle steps;
        // the last lane of a piece owns its total (longer runs cost one more set of atomics per 8 points)
        const uint64_t prev = __shfl_up_sync(FULL_MASK, key, 1);
        const unsigned run_heads = __ballot_sync(FULL_MASK, lane == 0 || key != prev);
        const int run_start = 31 - __clz(run_heads & le);
        const unsigned heads = __ballot_sync(FULL_MASK, (((int)lane - run_start) & 7) == 0);
        const int seg_start = 31 - __clz(heads & le);
#pragma unroll
        for (int o = 1; o < 8; o <<= 1) {
            const long long tx = __shfl_up_sync(FULL_MASK, fx, o), ty = __shfl_up_sync(FULL_MASK, fy, o), tz = __shfl_up_sync(FULL_MASK, fz, o);
            const unsigned long long trg = __shfl_up_sync(FULL_MASK, rg, o), tbn = __shfl_up_sync(FULL_MASK, bn, o);
            const uint32_t tt = __shfl_up_sync(FULL_MASK, tl, o);
            if ((int)lane - o >= seg_start) {
                fx += tx;
                fy += ty;
                fz += tz;
                rg += trg;
                bn += tbn;
                tl |= tt;
            }
        }
        const bool tail = lane == 31 || ((heads >> (lane + 1)) & 1u);
        bool claimed = false;
        uint32_t slot = 0;
        if (valid && tail) {
            const unsigned long long want = key + 1ull;
            slot = hash_key(key) & slot_mask;
            while (true) {
                unsigned long long cur = *reinterpret_cast<volatile unsigned long long *>(&table[slot].keyp1);
                if (cur == 0ull) {
                    cur = atomicCAS(&table[slot].keyp1, 0ull, want);
                    if (cur == 0ull) { // claimed: one list entry per voxel
                        claimed = true;
                        break;
                    }
                }
                if (cur == want) break;
                slot = (slot + 1u) & slot_mask;
            }
            VoxelSlot *sl = table + slot;
            atomicAdd(&sl->sx, (unsigned long long)fx);
            atomicAdd(&sl->sy, (unsigned long long)fy);
            atomicAdd(&sl->sz, (unsigned long long)fz);
            atomicAdd(&sl->rg, rg);
            atomicAdd(&sl->bn, bn);
            atomicOr(&sl->tile, tl);
        }
        // one counter update per warp for the slots it claimed
        const unsigned cm = __ballot_sync(FULL_MASK, claimed);
        if (cm) {
            uint32_t first = 0;
            if (lane == (unsigned)(__ffs(cm) - 1)) first = atomicAdd(&header->count, (uint32_t)__popc(cm));
            first = __shfl_sync(FULL_MASK, first, __ffs(cm) - 1);
            if (claimed) list[first + __popc(cm & lanemask_lt())] = (key << slotbits) | slot;
        }
    }
    if (bad) atomicOr(&header->error, 1u);
}

// mean xyz (exact sum, one rounding), truncated float colour average as pcl::CentroidPoint, OR of tiles
__device__ __forceinline__ Point16 finalize_voxel(long long sx, long long sy, long long sz, unsigned long long sr, unsigned long long sg, unsigned long long sb,
                                                   unsigned long long cnt, uint32_t tile, double inv_scale) {
    Point16 o;
    const double dn = (double)cnt;
    o.x = (float)(((double)sx * inv_scale) / dn);
    o.y = (float)(((double)sy * inv_scale) / dn);
    o.z = (float)(((double)sz * inv_scale) / dn);
    const float fn = (float)cnt;
    const uint32_t r = (uint32_t)__fdiv_rn((float)sr, fn) & 0xffu;
    const uint32_t g = (uint32_t)__fdiv_rn((float)sg, fn) & 0xffu;
    const uint32_t b = (uint32_t)__fdiv_rn((float)sb, fn) & 0xffu;
    o.rgbt = r | (g << 8) | (b << 16) | ((tile & 0xffu) << 24);
    return o;
}

// One output point per sorted list entry; the slot is read once and cleared, so the table is all zero
// again when the kernel ends (no memset between calls).
__global__ void __launch_bounds__(256) voxel_emit_kernel(const uint64_t *__restrict__ sorted, uint32_t v, uint32_t slot_mask, VoxelSlot *__restrict__ table, double inv_scale,
                                                          cwipc_point *__restrict__ out, TableHeader *__restrict__ header) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j == 0) header->count = 0;
    if (j >= v) return;
    VoxelSlot *sl = table + (uint32_t)(sorted[j] & (uint64_t)slot_mask);
    uint4 *raw = reinterpret_cast<uint4 *>(sl);
    const uint4 a = raw[0], b = raw[1], c = raw[2], d = raw[3];
    const uint4 zero = make_uint4(0u, 0u, 0u, 0u);
    raw[0] = zero;
    raw[1] = zero;
    raw[2] = zero;
    raw[3] = zero;
    const long long sx = (long long)(((unsigned long long)a.w << 32) | a.z);
    const long long sy = (long long)(((unsigned long long)b.y << 32) | b.x);
    const long long sz = (long long)(((unsigned long long)b.w << 32) | b.z);
    st_point(out, j, finalize_voxel(sx, sy, sz, c.x, c.y, c.z, c.w, d.x, inv_scale));
}

int bit_length(uint64_t v) {
    int b = 0;
    while (v) {
        b++;
        v >>= 1;
    }
    return b;
}

// the "last block" ticket of bbox_octree_kernel: a word of the thread's zeroed workspace header (pad[0])
uint32_t *bbox_counter(int dev, cudaStream_t s) {
    TableHeader *h = static_cast<TableHeader *>(thread_zeroed(dev, ZW_HEADER_BYTES, s));
    return &h->pad[0];
}

struct Plan {
    KeyParams kp;
    int keybits = 0;
    bool failed = false;
    std::string error;
    float maxabs = 0.f;
    float gmin[3] = {0, 0, 0}, gmax[3] = {0, 0, 0};
    size_t capacity = 0; // hash table slots (power of two)
    int slotbits = 0;
    bool use_tables = false;
    KeyTables tables;
};

// One launch + one readback: bounding box of the points and the octree box after inserting them in order
// (continuing from `seed` when it is valid).
OctreeBox measure_box(const cwipc_point *pts, size_t n, float cellsize, bool octree_split, const OctreeSeed &seed, int dev, cudaStream_t s) {
    const uint32_t nchunks = (uint32_t)div_up(n, BB_CHUNK);
    Scratch chunk_bbox((size_t)nchunks * 6 * sizeof(float), s);
    Scratch box(sizeof(OctreeBox), s);
    const float octree_cellsize = 64 * cellsize;           // ref: src/cwipc_filters.cpp:113-114 (float)
    const double res = (double)octree_cellsize;
    launch("bbox_octree_kernel", s, 16 * (size_t)n, [&] {
        bbox_octree_kernel<<<nchunks, BB_THREADS, 0, s>>>(pts, (uint32_t)n, chunk_bbox.as<float>(), nchunks, res, octree_split ? 1 : 0, seed, bbox_counter(dev, s), box.as<OctreeBox>());
    });
    OctreeBox *h = static_cast<OctreeBox *>(thread_pinned(sizeof(OctreeBox)));
    CWCU_CHECK(cudaMemcpyAsync(h, box.p, sizeof(OctreeBox), cudaMemcpyDeviceToHost, s));
    stream_sync(s);
    return *h;
}

// host copies of the device formulas (IEEE double arithmetic on both sides)
uint32_t host_leaf_index(float x, const KeyParams &kp, int a) { return (uint32_t)(((double)x - kp.omin[a]) / kp.res); }
uint64_t host_spread3(uint32_t v) {
    uint64_t x = v & 0x1fffffu;
    x = (x | x << 32) & 0x1f00000000ffffull;
    x = (x | x << 16) & 0x1f0000ff0000ffull;
    x = (x | x << 8) & 0x100f00f00f00f00full;
    x = (x | x << 4) & 0x10c30c30c30c30c3ull;
    x = (x | x << 2) & 0x1249249249249249ull;
    return x;
}

// Leaf lookup tables for the accumulate kernel; false when an axis spans more than KT_LEAVES leaves.
bool build_key_tables(const KeyParams &kp, const OctreeBox &ob, KeyTables &tab) {
    for (int a = 0; a < 3; a++) {
        if (!(ob.gmin[a] <= ob.gmax[a]) || (double)ob.gmin[a] < kp.omin[a]) return false;
        const uint32_t first = host_leaf_index(ob.gmin[a], kp, a), last = host_leaf_index(ob.gmax[a], kp, a);
        if (last < first || last - first + 1 > (uint32_t)KT_LEAVES || last >= (1u << kp.depth)) return false;
        for (int m = 0; m < KT_LEAVES; m++) {
            const uint32_t leaf = first + (uint32_t)m;
            tab.thr[a][m] = INFINITY;
            tab.origin[a][m] = 0;
            tab.spread[a][m] = 0;
            if (leaf > last) continue;
            tab.origin[a][m] = (int)std::floor((kp.omin[a] + (double)leaf * kp.res) * kp.inv_cs) - 2;
            tab.spread[a][m] = host_spread3(leaf) << (2 - a);
            if (leaf == last) continue;
            // smallest float whose leaf index is leaf + 1: start at the face, then walk to the exact step
            float x = (float)(kp.omin[a] + (double)(leaf + 1) * kp.res);
            for (int it = 0; it < 64 && host_leaf_index(x, kp, a) <= leaf; it++) x = std::nextafter(x, INFINITY);
            for (int it = 0; it < 64; it++) {
                const float below = std::nextafter(x, -INFINITY);
                if (!((double)below >= kp.omin[a]) || host_leaf_index(below, kp, a) <= leaf) break;
                x = below;
            }
            if (host_leaf_index(x, kp, a) != leaf + 1 || host_leaf_index(std::nextafter(x, -INFINITY), kp, a) != leaf) return false; // did not converge: generic path
            tab.thr[a][m] = x;
        }
    }
    return true;
}

// Key layout, hash-table size and limits from the measured (or supplied) boxes.
Plan derive_plan(const OctreeBox &ob, size_t n, float cellsize, bool octree_split) {
    Plan plan;
    const double res = (double)(64 * cellsize);
    KeyParams &kp = plan.kp;
    memset(&kp, 0, sizeof(kp));
    kp.octree = octree_split ? 1 : 0;
    kp.inv = 1.0f / cellsize;
    kp.idxbits = std::max(1, bit_length((uint64_t)n - 1));
    for (int a = 0; a < 3; a++) {
        plan.maxabs = std::max(plan.maxabs, std::max(std::fabs(ob.gmin[a]), std::fabs(ob.gmax[a])));
        plan.gmin[a] = ob.gmin[a];
        plan.gmax[a] = ob.gmax[a];
    }
    plan.capacity = 1024;
    while (plan.capacity < n + n / 2) plan.capacity <<= 1;
    plan.slotbits = bit_length((uint64_t)plan.capacity - 1);
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
        kp.inv_res = 1.0 / res;
        kp.inv_cs = 1.0 / (double)cellsize;
        kp.depth = ob.depth;
        plan.keybits = WL_BITS + 3 * ob.depth;
        plan.use_tables = build_key_tables(kp, ob, plan.tables);
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
    if (plan.keybits + plan.slotbits > 64) {
        plan.failed = true;
        plan.error = "pointcloud too large for 64-bit voxel keys (" + std::to_string(plan.keybits) + " key bits + " + std::to_string(plan.slotbits) + " slot bits)";
    }
    return plan;
}

unsigned stream_grid(size_t n, int dev) {
    return (unsigned)std::max<size_t>(1, std::min(div_up(n, 256), (size_t)sm_count(dev) * 8));
}

} // namespace

namespace {
DownsampleResult downsample_impl(const StoragePtr &in, float cellsize, bool octree_split, const OctreeBox *external, int dev, cudaStream_t s) {
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
    OctreeBox ob;
    if (external) {
        ob = *external;
    } else {
        OctreeSeed none;
        memset(&none, 0, sizeof(none));
        ob = measure_box(in->d_pts, n, cellsize, octree_split, none, dev, s);
    }
    Plan plan = derive_plan(ob, n, cellsize, octree_split);
    if (plan.failed) {
        result.failed = true;
        result.error = plan.error;
        return result;
    }
    const KeyParams &kp = plan.kp;

    // fixed-point scale: |x| * 2^shift * n < 2^62
    int shift = 62 - bit_length((uint64_t)n);
    if (plan.maxabs > 0.f) {
        int e;
        (void)std::frexp(plan.maxabs, &e); // maxabs < 2^e
        shift -= e;
    }
    shift = std::max(-60, std::min(shift, 100));
    const double scale = std::ldexp(1.0, shift), inv_scale = std::ldexp(1.0, -shift);

    // hash table in the thread's zeroed workspace: [header | capacity slots], load factor <= 2/3
    const size_t capacity = plan.capacity;
    const int slotbits = plan.slotbits;
    uint8_t *ws = static_cast<uint8_t *>(thread_zeroed(dev, ZW_HEADER_BYTES + capacity * sizeof(VoxelSlot), s));
    TableHeader *header = reinterpret_cast<TableHeader *>(ws);
    VoxelSlot *table = reinterpret_cast<VoxelSlot *>(ws + ZW_HEADER_BYTES);
    try {
        Scratch list(n * sizeof(uint64_t), s);
        launch("voxel_accumulate_kernel", s, 16 * (size_t)n, [&] {
            if (plan.use_tables)
                voxel_accumulate_kernel<true><<<stream_grid(n, dev), VA_THREADS, 0, s>>>(in->d_pts, (uint32_t)n, kp, plan.tables, (float)scale, table, (uint32_t)(capacity - 1), slotbits,
                                                                                          list.as<uint64_t>(), header);
            else
                voxel_accumulate_kernel<false><<<stream_grid(n, dev), VA_THREADS, 0, s>>>(in->d_pts, (uint32_t)n, kp, plan.tables, (float)scale, table, (uint32_t)(capacity - 1), slotbits,
                                                                                           list.as<uint64_t>(), header);
        });
        uint32_t *h = static_cast<uint32_t *>(thread_pinned(2 * sizeof(uint32_t)));
        CWCU_CHECK(cudaMemcpyAsync(h, header, 2 * sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
        stream_sync(s);
        const size_t v = h[0];
        if (h[1] != 0) {
            thread_zeroed_invalidate(dev);
            result.failed = true;
            result.error = "point coordinates out of range for voxel size " + std::to_string(cellsize) + " (|x/voxelsize| must stay below 2^22)";
            return result;
        }
        Scratch other(v * sizeof(uint64_t), s);
        const uint64_t *sorted = radix_sort_u64(list.as<uint64_t>(), other.as<uint64_t>(), v, slotbits, slotbits + plan.keybits, dev, s);
        auto out = std::make_shared<Storage>(dev, v, s);
        launch("voxel_emit_kernel", s, 16 * v, [&] {
            voxel_emit_kernel<<<(unsigned)std::max<size_t>(1, div_up(v, 256)), 256, 0, s>>>(sorted, (uint32_t)v, (uint32_t)(capacity - 1), table, inv_scale, out->d_pts, header);
        });
        out->count = v;
        // centroids lie inside the input's bounding box: hand it on so that a following filter need not recompute it
        out->has_bounds = true;
        for (int a = 0; a < 3; a++) {
            out->bounds_min[a] = plan.gmin[a];
            out->bounds_max[a] = plan.gmax[a];
        }
        out->mark_ready();
        result.out = out;
    } catch (...) {
        thread_zeroed_invalidate(dev);
        throw;
    }
    return result;
}

OctreeSeed no_seed() {
    OctreeSeed none;
    memset(&none, 0, sizeof(none));
    return none;
}
} // namespace

DownsampleResult downsample_points(const StoragePtr &in, float cellsize, bool octree_split, int dev, cudaStream_t s) {
    return downsample_impl(in, cellsize, octree_split, nullptr, dev, s);
}

// Partitioned clouds: the octree box and the bounding box of the WHOLE cloud are supplied by the caller.
DownsampleResult downsample_points_planned(const StoragePtr &in, float cellsize, bool octree_split, const OctreeState &state, const float bounds[6], int dev, cudaStream_t s) {
    OctreeBox ob;
    memset(&ob, 0, sizeof(ob));
    for (int a = 0; a < 3; a++) {
        ob.min[a] = state.min[a];
        ob.max[a] = state.max[a];
        ob.gmin[a] = bounds[a];
        ob.gmax[a] = bounds[3 + a];
    }
    ob.depth = state.depth;
    if (octree_split && !state.valid && in->count > 0) {
        DownsampleResult r;
        r.failed = true;
        r.error = "planned downsample needs the octree state of the whole cloud";
        return r;
    }
    return downsample_impl(in, cellsize, octree_split, &ob, dev, s);
}

// Continue the octree bounding-box replay over this cloud's points; also returns its bounding box.
void octree_replay(const cwipc_point *in, size_t n, float cellsize, OctreeState &state, float bounds[6], int dev, cudaStream_t s) {
    for (int a = 0; a < 3; a++) {
        bounds[a] = INFINITY;
        bounds[3 + a] = -INFINITY;
    }
    if (n == 0) return;
    OctreeSeed seed;
    memset(&seed, 0, sizeof(seed));
    if (state.valid) {
        for (int a = 0; a < 3; a++) {
            seed.min[a] = state.min[a];
            seed.max[a] = state.max[a];
        }
        seed.depth = state.depth;
        seed.valid = 1;
    }
    const OctreeBox ob = measure_box(in, n, cellsize, true, seed, dev, s);
    if (ob.error) throw CudaError{cudaErrorInvalidValue, "octree replay: runaway growth (non-finite coordinates?)"};
    for (int a = 0; a < 3; a++) {
        state.min[a] = ob.min[a];
        state.max[a] = ob.max[a];
        bounds[a] = ob.gmin[a];
        bounds[3 + a] = ob.gmax[a];
    }
    state.depth = ob.depth;
    state.valid = 1;
}

void global_bbox(const cwipc_point *in, size_t n, float gmin[3], float gmax[3], int dev, cudaStream_t s) {
    const uint32_t nchunks = (uint32_t)div_up(n, BB_CHUNK);
    Scratch chunk_bbox((size_t)nchunks * 6 * sizeof(float), s);
    Scratch box(sizeof(OctreeBox), s);
    launch("bbox_octree_kernel", s, 16 * (size_t)n, [&] {
        bbox_octree_kernel<<<nchunks, BB_THREADS, 0, s>>>(in, (uint32_t)n, chunk_bbox.as<float>(), nchunks, 1.0, 0, no_seed(), bbox_counter(dev, s), box.as<OctreeBox>());
    });
    OctreeBox *h = static_cast<OctreeBox *>(thread_pinned(sizeof(OctreeBox)));
    CWCU_CHECK(cudaMemcpyAsync(h, box.p, sizeof(OctreeBox), cudaMemcpyDeviceToHost, s));
    stream_sync(s);
    for (int a = 0; a < 3; a++) {
        gmin[a] = h->gmin[a];
        gmax[a] = h->gmax[a];
    }
}

void downsample_keys_to_host(const StoragePtr &in, float cellsize, bool octree_split, uint64_t *host_keys, int dev, cudaStream_t s) {
    const size_t n = in->count;
    if (n == 0) return;
    Plan plan = derive_plan(measure_box(in->d_pts, n, cellsize, octree_split, no_seed(), dev, s), n, cellsize, octree_split);
    if (plan.failed) throw CudaError{cudaErrorInvalidValue, plan.error};
    Scratch keys(n * sizeof(uint64_t), s);
    Scratch flag(sizeof(uint32_t), s);
    CWCU_CHECK(cudaMemsetAsync(flag.p, 0, sizeof(uint32_t), s));
    launch("voxel_keygen_kernel", s, 24 * (size_t)n, [&] { voxel_keygen_kernel<<<stream_grid(n, dev), 256, 0, s>>>(in->d_pts, (uint32_t)n, plan.kp, keys.as<uint64_t>(), 0, flag.as<uint32_t>()); });
    CWCU_CHECK(cudaMemcpyAsync(host_keys, keys.p, n * sizeof(uint64_t), cudaMemcpyDeviceToHost, s));
    stream_sync(s);
}

} // namespace cwcu
