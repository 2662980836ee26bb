// outliers.cu -- statistical outlier removal (exact k-nearest-neighbour statistics) on sm_100a.
//
// Reference semantics (src/cwipc_filters.cpp:181-211 over pcl::StatisticalOutlierRemoval ->
// pcl::search::KdTree -> FLANN KDTreeSingleIndex with L2_Simple<float>):
//   for every point: the k+1 smallest float distances^2  ((dx*dx + dy*dy) + dz*dz, no FMA) to ALL
//   points (itself included); d_i = (float)( sum_{j=1..k} sqrt((double)d2_j) / k );
//   mean/variance of d in double (squares taken in float); keep iff !(d_i > mean + mul*stddev).
//
// The kd-tree is replaced by a uniform grid in Morton order:
//   knn_keygen_kernel   16 B read + 8 B   key = [Morton(cell) | point index]
//   radix_sort_u64      P x (8+8) B       (shared with downsample)
//   knn_gather_kernel   8+16 B read, 16 B points re-laid out in cell order (each cell and every aligned
//                                         2^l-cube of cells is one contiguous range)
//   cell_heads_kernel   8 B read          list of occupied cells (decoupled look-back compaction)
//   knn_cell_kernel     one warp per occupied cell, one lane per query: exact top-(k+1) over the 27
//                       neighbouring cells, kept as a sorted register array (min/max insertion network).
//                       A query is final when its (k+1)-th distance is within the distance to the border
//                       of the 3x3x3 block; otherwise it is queued for a coarser level.
//   knn_far_kernel      levels 1..L: one warp per queued query, lanes stride over the candidates of the
//                       27 level-l cells that intersect the known bound, per-lane lists merged by warp
//                       min-reduction.  The top level spans the whole cloud, so every query terminates.
//   stats_kernel        sum d, sum (float)(d*d) in double, fixed two-level order (deterministic)
//   compact_kernel      keep mask + stable compaction (pointops.cu)
// Exactness never depends on the grid pitch; the pitch only moves work between the levels.
#include <algorithm>
#include <cmath>
#include <cstring>

#include "device_utils.cuh"
#include "kernels.hpp"
#include "radix_sort.hpp"

namespace cwcu {

namespace {

struct GridParams {
    float gmin[3];
    float inv_h;   // 1 / cell pitch
    float h;       // cell pitch
    int gdim[3];   // cells per axis
    int idxbits;
    int top_level; // level at which the whole grid is one cell
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
__device__ __forceinline__ uint32_t compact3(uint64_t x) { // inverse of spread3
    x &= 0x1249249249249249ull;
    x = (x | x >> 2) & 0x10c30c30c30c30c3ull;
    x = (x | x >> 4) & 0x100f00f00f00f00full;
    x = (x | x >> 8) & 0x1f0000ff0000ffull;
    x = (x | x >> 16) & 0x1f00000000ffffull;
    x = (x | x >> 32) & 0x1fffffull;
    return (uint32_t)x;
}
__device__ __forceinline__ uint64_t morton3(uint32_t x, uint32_t y, uint32_t z) { return (spread3(x) << 2) | (spread3(y) << 1) | spread3(z); }

// cell coordinate in cell units (float), monotone in the coordinate
__device__ __forceinline__ float cell_u(float x, float gmin, float inv_h) { return __fmul_rn(__fsub_rn(x, gmin), inv_h); }
__device__ __forceinline__ int cell_of(float u, int gdim) { return min(max((int)floorf(u), 0), gdim - 1); }

// FLANN L2_Simple<float>: ((dx*dx) + (dy*dy)) + (dz*dz), every operation rounded to float
__device__ __forceinline__ float dist2(const Point16 &a, const Point16 &b) {
    const float dx = __fsub_rn(a.x, b.x), dy = __fsub_rn(a.y, b.y), dz = __fsub_rn(a.z, b.z);
    return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
}

__global__ void __launch_bounds__(256) knn_keygen_kernel(const cwipc_point *__restrict__ pts, uint32_t n, GridParams gp, uint64_t *__restrict__ keys) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const Point16 p = ld_point_stream(pts, i);
        const uint32_t cx = cell_of(cell_u(p.x, gp.gmin[0], gp.inv_h), gp.gdim[0]);
        const uint32_t cy = cell_of(cell_u(p.y, gp.gmin[1], gp.inv_h), gp.gdim[1]);
        const uint32_t cz = cell_of(cell_u(p.z, gp.gmin[2], gp.inv_h), gp.gdim[2]);
        keys[i] = (morton3(cx, cy, cz) << gp.idxbits) | i;
    }
}

__global__ void __launch_bounds__(256) knn_gather_kernel(const uint64_t *__restrict__ sorted, uint32_t n, int idxbits, const cwipc_point *__restrict__ pts,
                                                          cwipc_point *__restrict__ spts) {
    const uint64_t idxmask = (1ull << idxbits) - 1ull;
    for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) st_point(spts, j, ld_point(pts, (size_t)(sorted[j] & idxmask)));
}

// ---- occupied cells: positions where the Morton code changes ---------------------------------
constexpr int CH_THREADS = 256;
constexpr int CH_ITEMS = 8;
constexpr int CH_TILE = CH_THREADS * CH_ITEMS;

__global__ void __launch_bounds__(CH_THREADS) cell_heads_kernel(const uint64_t *__restrict__ sorted, uint32_t n, int idxbits, uint32_t *__restrict__ cell_start,
                                                                 uint64_t *__restrict__ cell_code, uint32_t *__restrict__ ticket, uint64_t *__restrict__ status,
                                                                 uint32_t *__restrict__ d_ncells) {
    __shared__ int s_tile;
    __shared__ uint32_t s_warp_total[CH_THREADS / 32];
    __shared__ uint32_t s_tile_excl;
    if (threadIdx.x == 0) s_tile = take_ticket(ticket);
    __syncthreads();
    const int tile = s_tile;
    const uint32_t tile_base = (uint32_t)tile * CH_TILE;
    if (tile_base >= n) return;
    const unsigned warp = threadIdx.x >> 5, lane = lane_id();
    const uint32_t warp_base = tile_base + warp * (32 * CH_ITEMS);
    const unsigned lt = lanemask_lt();
    uint32_t rank[CH_ITEMS];
    uint64_t code[CH_ITEMS];
    unsigned headbits = 0;
    uint32_t running = 0;
#pragma unroll
    for (int i = 0; i < CH_ITEMS; i++) {
        const uint32_t e = warp_base + i * 32 + lane;
        bool head = false;
        if (e < n) {
            code[i] = sorted[e] >> idxbits;
            head = (e == 0) || ((sorted[e - 1] >> idxbits) != code[i]);
        }
        const unsigned b = __ballot_sync(FULL_MASK, head);
        rank[i] = running + __popc(b & lt);
        running += __popc(b);
        if (head) headbits |= 1u << i;
    }
    if (lane == 0) s_warp_total[warp] = running;
    __syncthreads();
    uint32_t warp_excl = 0, block_total = 0;
#pragma unroll
    for (int w = 0; w < CH_THREADS / 32; w++) {
        const uint32_t t = s_warp_total[w];
        if (w < (int)warp) warp_excl += t;
        block_total += t;
    }
    if (warp == 0) {
        const uint32_t excl = lookback_exclusive(status, tile, block_total);
        if (lane == 0) {
            s_tile_excl = excl;
            if (tile_base + CH_TILE >= n) {
                *d_ncells = excl + block_total;
                cell_start[excl + block_total] = n; // sentinel: end of the last cell
            }
        }
    }
    __syncthreads();
    const uint32_t base = s_tile_excl + warp_excl;
#pragma unroll
    for (int i = 0; i < CH_ITEMS; i++) {
        if (headbits & (1u << i)) {
            cell_start[base + rank[i]] = warp_base + i * 32 + lane;
            cell_code[base + rank[i]] = code[i];
        }
    }
}

// first index in code[0..n) with code[i] >= key
__device__ __forceinline__ uint32_t lower_bound_u64(const uint64_t *__restrict__ code, uint32_t n, uint64_t key) {
    uint32_t lo = 0, hi = n;
    while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        if (code[mid] < key) lo = mid + 1;
        else hi = mid;
    }
    return lo;
}

// sorted insertion into an ascending register array: new[j] = min(best[j], max(best[j-1], d))
template <int KCAP>
__device__ __forceinline__ void topk_insert(float (&best)[KCAP], float d) {
#pragma unroll
    for (int j = KCAP - 1; j > 0; j--) best[j] = fminf(best[j], fmaxf(best[j - 1], d));
    best[0] = fminf(best[0], d);
}

struct FarEntry {
    uint32_t q;     // index in cell (sorted) order
    uint32_t level; // level at which the query must be processed
    float bound;    // valid upper bound on the (k+1)-th squared distance, +inf if unknown
};

// smallest level l >= min_level whose pitch h*2^l covers sqrt(bound) (with slack); min_level if unbounded
__device__ __forceinline__ uint32_t level_for_bound(float bound, float h, uint32_t min_level, uint32_t top_level) {
    uint32_t l = min_level;
    if (bound < INFINITY) {
        const float r = sqrtf(bound) * 1.02f;
        float pitch = ldexpf(h, (int)l);
        while (l < top_level && pitch < r) {
            l++;
            pitch *= 2.f;
        }
    }
    return min(l, top_level);
}

// distance (in units of the level pitch) from u (cell units at that level) to the border of the
// 3x3x3 block around cell c: min(u - (c-1), (c+2) - u)
__device__ __forceinline__ float border_margin(float u, int c) { return fminf(u - (float)(c - 1), (float)(c + 2) - u); }

// ---- level 0: warp per occupied cell, lane per query ----------------------------------------------
template <int KCAP>
__global__ void __launch_bounds__(128) knn_cell_kernel(const cwipc_point *__restrict__ spts, const uint64_t *__restrict__ sorted, uint32_t n, GridParams gp, int kk, int k,
                                                        const uint32_t *__restrict__ cell_start, const uint64_t *__restrict__ cell_code,
                                                        const uint32_t *__restrict__ d_ncells, float *__restrict__ dist_out, FarEntry *__restrict__ far_list,
                                                        uint32_t *__restrict__ far_count) {
    const unsigned lane = lane_id();
    const uint32_t ncells = *d_ncells;
    const uint32_t warps_total = (gridDim.x * blockDim.x) >> 5;
    const uint64_t idxmask = (1ull << gp.idxbits) - 1ull;
    const int extra = KCAP - kk; // leading slots pinned at -inf so that best[KCAP-1] is the kk-th smallest
    for (uint32_t cell = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; cell < ncells; cell += warps_total) {
        const uint32_t q_begin = cell_start[cell], q_end = cell_start[cell + 1];
        const uint64_t code = cell_code[cell];
        const int cx = (int)compact3(code >> 2), cy = (int)compact3(code >> 1), cz = (int)compact3(code);
        // lanes 0..26 locate one neighbour cell each
        uint32_t nb_begin = 0, nb_end = 0;
        if (lane < 27) {
            const int nx = cx + (int)(lane % 3) - 1, ny = cy + (int)((lane / 3) % 3) - 1, nz = cz + (int)(lane / 9) - 1;
            if (nx >= 0 && ny >= 0 && nz >= 0 && nx < gp.gdim[0] && ny < gp.gdim[1] && nz < gp.gdim[2]) {
                const uint64_t ncode = morton3((uint32_t)nx, (uint32_t)ny, (uint32_t)nz);
                const uint32_t pos = lower_bound_u64(cell_code, ncells, ncode);
                if (pos < ncells && cell_code[pos] == ncode) {
                    nb_begin = cell_start[pos];
                    nb_end = cell_start[pos + 1];
                }
            }
        }
        for (uint32_t qb = q_begin; qb < q_end; qb += 32) {
            const uint32_t qi = qb + lane;
            const bool active = qi < q_end;
            Point16 q = ld_point(spts, active ? qi : q_begin);
            float best[KCAP];
#pragma unroll
            for (int j = 0; j < KCAP; j++) best[j] = (j < extra) ? -INFINITY : INFINITY;
            for (int nb = 0; nb < 27; nb++) {
                const uint32_t cb = __shfl_sync(FULL_MASK, nb_begin, nb), ce = __shfl_sync(FULL_MASK, nb_end, nb);
                for (uint32_t c = cb; c < ce; c++) {
                    const Point16 cand = ld_point(spts, c); // same address in every lane: one broadcast load
                    const float d2 = dist2(q, cand);
                    if (d2 < best[KCAP - 1]) topk_insert<KCAP>(best, d2);
                }
            }
            if (!active) continue;
            // exact iff the kk-th distance does not reach the border of the 3x3x3 block
            const float ux = cell_u(q.x, gp.gmin[0], gp.inv_h), uy = cell_u(q.y, gp.gmin[1], gp.inv_h), uz = cell_u(q.z, gp.gmin[2], gp.inv_h);
            const float m = fminf(fminf(border_margin(ux, cx), border_margin(uy, cy)), border_margin(uz, cz)) - 0.01f;
            const float reach = m * gp.h;
            const float worst = best[KCAP - 1];
            if (gp.top_level == 0 || (m > 0.f && worst <= reach * reach * 0.999999f)) {
                double sum = 0.0;
#pragma unroll
                for (int j = 0; j < KCAP; j++)
                    if (j > extra) sum += sqrt((double)best[j]); // j == extra is the query itself (distance 0)
                dist_out[(size_t)(sorted[qi] & idxmask)] = (float)(sum / (double)k);
            } else {
                const uint32_t slot = atomicAdd(far_count, 1u);
                FarEntry e;
                e.q = qi;
                e.bound = worst;
                e.level = level_for_bound(worst, gp.h, 1u, (uint32_t)gp.top_level);
                far_list[slot] = e;
            }
        }
    }
}

// ---- levels >= 1: warp per queued query ----------------------------------------------------------
template <int KCAP>
__global__ void __launch_bounds__(128) knn_far_kernel(const cwipc_point *__restrict__ spts, const uint64_t *__restrict__ sorted, uint32_t n, GridParams gp, int kk, int k,
                                                       uint32_t level, const uint32_t *__restrict__ cell_start, const uint64_t *__restrict__ cell_code,
                                                       const uint32_t *__restrict__ d_ncells, float *__restrict__ dist_out, FarEntry *__restrict__ far_list,
                                                       const uint32_t *__restrict__ far_count) {
    const unsigned lane = lane_id();
    const uint32_t ncells = *d_ncells;
    const uint32_t nentries = *far_count; // queued by level 0; entries are re-levelled in place
    const uint32_t warps_total = (gridDim.x * blockDim.x) >> 5;
    const uint64_t idxmask = (1ull << gp.idxbits) - 1ull;
    const float pitch = ldexpf(gp.h, (int)level);
    const float inv_pitch = ldexpf(gp.inv_h, -(int)level);
    for (uint32_t ei = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; ei < nentries; ei += warps_total) {
        const FarEntry ent = far_list[ei];
        if (ent.level != level) continue;
        const Point16 q = ld_point(spts, ent.q);
        const uint64_t code0 = sorted[ent.q] >> gp.idxbits;
        const int cx = (int)(compact3(code0 >> 2) >> level), cy = (int)(compact3(code0 >> 1) >> level), cz = (int)(compact3(code0) >> level);
        const float bound = ent.bound;
        // lanes 0..26: candidate range of one level-`level` neighbour cell (a contiguous run of fine cells)
        uint32_t nb_begin = 0, nb_end = 0;
        if (lane < 27) {
            const int nx = cx + (int)(lane % 3) - 1, ny = cy + (int)((lane / 3) % 3) - 1, nz = cz + (int)(lane / 9) - 1;
            const int gx = (gp.gdim[0] - 1) >> level, gy = (gp.gdim[1] - 1) >> level, gz = (gp.gdim[2] - 1) >> level;
            if (nx >= 0 && ny >= 0 && nz >= 0 && nx <= gx && ny <= gy && nz <= gz) {
                bool wanted = true;
                if (bound < INFINITY) {
                    // squared distance from q to the neighbour's box; skip boxes beyond the known bound
                    const float lo[3] = {gp.gmin[0] + (float)nx * pitch, gp.gmin[1] + (float)ny * pitch, gp.gmin[2] + (float)nz * pitch};
                    const float qq[3] = {q.x, q.y, q.z};
                    float bd2 = 0.f;
#pragma unroll
                    for (int a = 0; a < 3; a++) {
                        const float below = lo[a] - qq[a], above = qq[a] - (lo[a] + pitch);
                        const float d = fmaxf(fmaxf(below, above), 0.f);
                        bd2 += d * d;
                    }
                    wanted = bd2 * 0.98f <= bound;
                }
                if (wanted) {
                    const uint64_t lo_code = morton3((uint32_t)nx, (uint32_t)ny, (uint32_t)nz) << (3 * level);
                    const uint64_t hi_code = lo_code + (1ull << (3 * level));
                    const uint32_t p0 = lower_bound_u64(cell_code, ncells, lo_code);
                    const uint32_t p1 = lower_bound_u64(cell_code, ncells, hi_code);
                    nb_begin = cell_start[p0];
                    nb_end = cell_start[p1];
                }
            }
        }
        float best[KCAP];
#pragma unroll
        for (int j = 0; j < KCAP; j++) best[j] = INFINITY;
        for (int nb = 0; nb < 27; nb++) {
            const uint32_t cb = __shfl_sync(FULL_MASK, nb_begin, nb), ce = __shfl_sync(FULL_MASK, nb_end, nb);
            for (uint32_t c = cb + lane; c < ce; c += 32) {
                const float d2 = dist2(q, ld_point(spts, c));
                if (d2 <= bound && d2 < best[KCAP - 1]) topk_insert<KCAP>(best, d2);
            }
        }
        // merge the 32 ascending lists: kk rounds of warp-min + pop at the winning lane
        double sum = 0.0;
        float worst = INFINITY;
        bool enough = true;
        for (int r = 0; r < kk; r++) {
            float m = best[0];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) m = fminf(m, __shfl_xor_sync(FULL_MASK, m, o));
            if (!(m < INFINITY)) {
                enough = false;
                break;
            }
            const unsigned winners = __ballot_sync(FULL_MASK, best[0] == m);
            if ((int)lane == __ffs(winners) - 1) {
#pragma unroll
                for (int j = 0; j < KCAP - 1; j++) best[j] = best[j + 1];
                best[KCAP - 1] = INFINITY;
            }
            if (r > 0) sum += sqrt((double)m); // r == 0 is the query itself
            worst = m;
        }
        if (lane != 0) continue;
        bool done = enough;
        if (done && level < (uint32_t)gp.top_level) {
            const float ux = cell_u(q.x, gp.gmin[0], inv_pitch), uy = cell_u(q.y, gp.gmin[1], inv_pitch), uz = cell_u(q.z, gp.gmin[2], inv_pitch);
            const float m = fminf(fminf(border_margin(ux, cx), border_margin(uy, cy)), border_margin(uz, cz)) - 0.01f;
            const float reach = m * pitch;
            done = (m > 0.f) && (worst <= reach * reach * 0.999999f);
        }
        if (done || level >= (uint32_t)gp.top_level) {
            // at the top level every point has been scanned; `enough` can only be false when n < kk,
            // which the host excludes
            dist_out[(size_t)(sorted[ent.q] & idxmask)] = (float)(sum / (double)k);
        } else {
            // still open: move the entry to a coarser level (strictly later launch), in place
            FarEntry e;
            e.q = ent.q;
            e.bound = enough ? fminf(worst, bound) : bound;
            e.level = level_for_bound(e.bound, gp.h, level + 1, (uint32_t)gp.top_level);
            far_list[ei] = e;
        }
    }
}

// ---- statistics: fixed-order two-level double reduction -------------------------------------------
constexpr int ST_BLOCKS = 256;
constexpr int ST_THREADS = 256;

__global__ void __launch_bounds__(ST_THREADS) stats_kernel(const float *__restrict__ dist, uint32_t n, double *__restrict__ partial) {
    __shared__ double s_sum[ST_THREADS], s_sq[ST_THREADS];
    double sum = 0.0, sq = 0.0;
    for (uint32_t i = blockIdx.x * ST_THREADS + threadIdx.x; i < n; i += ST_BLOCKS * ST_THREADS) {
        const float d = dist[i];
        sum += (double)d;
        sq += (double)__fmul_rn(d, d); // PCL squares in float, accumulates in double
    }
    s_sum[threadIdx.x] = sum;
    s_sq[threadIdx.x] = sq;
    __syncthreads();
    for (int o = ST_THREADS / 2; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) {
            s_sum[threadIdx.x] += s_sum[threadIdx.x + o];
            s_sq[threadIdx.x] += s_sq[threadIdx.x + o];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        partial[2 * blockIdx.x] = s_sum[0];
        partial[2 * blockIdx.x + 1] = s_sq[0];
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

unsigned stream_grid(size_t n, int dev) {
    return (unsigned)std::max<size_t>(1, std::min(div_up(n, 256), (size_t)sm_count(dev) * 8));
}

// Grid pitch: ~1.6x the expected k-neighbour radius of a surface sampled at `spacing`.
GridParams choose_grid(const float gmin[3], const float gmax[3], size_t n, int k, float hint_spacing) {
    GridParams gp;
    memset(&gp, 0, sizeof(gp));
    const double ext[3] = {(double)gmax[0] - gmin[0], (double)gmax[1] - gmin[1], (double)gmax[2] - gmin[2]};
    double spacing = hint_spacing > 0.f ? (double)hint_spacing : 0.0;
    if (!(spacing > 0.0)) {
        const double area = 2.0 * (ext[0] * ext[1] + ext[1] * ext[2] + ext[2] * ext[0]);
        const double longest = std::max(ext[0], std::max(ext[1], ext[2]));
        if (area > 0.0) spacing = std::sqrt(area / (double)n);
        else if (longest > 0.0) spacing = longest / (double)n;
        else spacing = 1.0;
    }
    double h = spacing * std::sqrt((double)(k + 1) / 3.14159265358979) * 1.6;
    const double longest = std::max(ext[0], std::max(ext[1], ext[2]));
    gp.idxbits = std::max(1, bit_length((uint64_t)n - 1));
    const int max_axis_bits = std::min(13, (64 - gp.idxbits) / 3);
    if (!(h > 0.0) || !std::isfinite(h)) h = 1.0;
    while (longest / h >= (double)((1 << max_axis_bits) - 1)) h *= 2.0;
    gp.h = (float)h;
    gp.inv_h = 1.0f / gp.h;
    int maxdim = 1;
    for (int a = 0; a < 3; a++) {
        gp.gmin[a] = gmin[a];
        gp.gdim[a] = std::max(1, std::min((int)std::floor(ext[a] / (double)gp.h) + 2, 1 << max_axis_bits));
        maxdim = std::max(maxdim, gp.gdim[a]);
    }
    gp.top_level = bit_length((uint64_t)maxdim - 1); // (gdim-1) >> top_level == 0 on every axis
    return gp;
}

template <int KCAP>
void run_knn(const cwipc_point *spts, const uint64_t *sorted, size_t n, const GridParams &gp, int k, const uint32_t *cell_start, const uint64_t *cell_code,
             const uint32_t *d_ncells, float *d_dist, FarEntry *far_list, uint32_t *far_count, int dev, cudaStream_t s) {
    const int kk = k + 1;
    const unsigned grid = (unsigned)std::max<size_t>(1, std::min(div_up(n, (size_t)32), (size_t)sm_count(dev) * 16));
    launch("knn_cell_kernel", s, 28 * (size_t)n, [&] {
        knn_cell_kernel<KCAP><<<grid, 128, 0, s>>>(spts, sorted, (uint32_t)n, gp, kk, k, cell_start, cell_code, d_ncells, d_dist, far_list, far_count);
    });
    for (int level = 1; level <= gp.top_level; level++) {
        launch("knn_far_kernel", s, (size_t)0, [&] {
            knn_far_kernel<KCAP><<<(unsigned)sm_count(dev) * 8, 128, 0, s>>>(spts, sorted, (uint32_t)n, gp, kk, k, (uint32_t)level, cell_start, cell_code, d_ncells, d_dist,
                                                                            far_list, far_count);
        });
    }
}

} // namespace

void knn_mean_distances(const cwipc_point *in, size_t n, int k, float hint_spacing, float *d_dist, int dev, cudaStream_t s) {
    if (n == 0) return;
    if (k < 1) throw CudaError{cudaErrorInvalidValue, "remove_outliers: kNeighbors must be >= 1"};
    if (k + 1 > 64) throw CudaError{cudaErrorInvalidValue, "remove_outliers: kNeighbors > 63 is not supported by libcwipc_util_cuda"};
    if ((size_t)k >= n) throw CudaError{cudaErrorInvalidValue, "remove_outliers: needs more than kNeighbors points"};
    float gmin[3], gmax[3];
    global_bbox(in, n, gmin, gmax, dev, s);
    for (int a = 0; a < 3; a++)
        if (!std::isfinite(gmin[a]) || !std::isfinite(gmax[a])) throw CudaError{cudaErrorInvalidValue, "remove_outliers: pointcloud contains non-finite coordinates"};
    const GridParams gp = choose_grid(gmin, gmax, n, k, hint_spacing);
    int axis_bits = 1;
    for (int a = 0; a < 3; a++) axis_bits = std::max(axis_bits, bit_length((uint64_t)gp.gdim[a] - 1));
    const int keybits = 3 * axis_bits;

    Scratch keys_a(n * sizeof(uint64_t), s), keys_b(n * sizeof(uint64_t), s);
    launch("knn_keygen_kernel", s, 24 * (size_t)n, [&] { knn_keygen_kernel<<<stream_grid(n, dev), 256, 0, s>>>(in, (uint32_t)n, gp, keys_a.as<uint64_t>()); });
    const uint64_t *sorted = radix_sort_u64(keys_a.as<uint64_t>(), keys_b.as<uint64_t>(), n, gp.idxbits, gp.idxbits + keybits, dev, s);

    Scratch spts(n * sizeof(cwipc_point), s);
    launch("knn_gather_kernel", s, 40 * (size_t)n, [&] { knn_gather_kernel<<<stream_grid(n, dev), 256, 0, s>>>(sorted, (uint32_t)n, gp.idxbits, in, spts.as<cwipc_point>()); });

    // occupied cells
    const size_t ntiles = div_up(n, CH_TILE);
    Scratch cell_start((n + 1) * sizeof(uint32_t), s), cell_code(n * sizeof(uint64_t), s);
    // [ticket | ncells | far_count | pad] u32*4 then status u64*ntiles
    const size_t aux_bytes = 16 + ntiles * sizeof(uint64_t);
    Scratch aux(aux_bytes, s);
    CWCU_CHECK(cudaMemsetAsync(aux.p, 0, aux_bytes, s));
    uint32_t *ticket = aux.as<uint32_t>();
    uint32_t *d_ncells = ticket + 1, *far_count = ticket + 2;
    uint64_t *status = reinterpret_cast<uint64_t *>(ticket + 4);
    launch("cell_heads_kernel", s, 8 * (size_t)n, [&] {
        cell_heads_kernel<<<(unsigned)ntiles, CH_THREADS, 0, s>>>(sorted, (uint32_t)n, gp.idxbits, cell_start.as<uint32_t>(), cell_code.as<uint64_t>(), ticket, status, d_ncells);
    });

    // a query is queued at most once (by level 0) and re-levelled in place afterwards
    Scratch far_list((n + 64) * sizeof(FarEntry), s);
    const int kk = k + 1;
    auto go = [&](auto kcap) {
        run_knn<decltype(kcap)::value>(spts.as<cwipc_point>(), sorted, n, gp, k, cell_start.as<uint32_t>(), cell_code.as<uint64_t>(), d_ncells, d_dist,
                                       far_list.as<FarEntry>(), far_count, dev, s);
    };
    if (kk <= 8) go(std::integral_constant<int, 8>{});
    else if (kk <= 16) go(std::integral_constant<int, 16>{});
    else if (kk <= 32) go(std::integral_constant<int, 32>{});
    else go(std::integral_constant<int, 64>{});
}

size_t remove_outliers_points(const cwipc_point *in, size_t n, cwipc_point *out, int k, float stddev_mul, float hint_spacing, int dev, cudaStream_t s) {
    if (n == 0) return 0;
    if (k < 1 || (size_t)k >= n) {
        // n <= k is undefined behaviour in the reference (PCL reads k+1 results where FLANN returned fewer);
        // defined here as "keep everything" with a warning, and the oracle does the same.
        log(CWIPC_LOG_LEVEL_WARNING, "cwipc_remove_outliers", "fewer points than kNeighbors+1 (" + std::to_string(n) + " <= " + std::to_string(k) + "): keeping all points");
        CWCU_CHECK(cudaMemcpyAsync(out, in, n * sizeof(cwipc_point), cudaMemcpyDeviceToDevice, s));
        return n;
    }
    Scratch dist(n * sizeof(float), s);
    knn_mean_distances(in, n, k, hint_spacing, dist.as<float>(), dev, s);

    Scratch partial(2 * ST_BLOCKS * sizeof(double), s);
    launch("stats_kernel", s, 4 * (size_t)n, [&] { stats_kernel<<<ST_BLOCKS, ST_THREADS, 0, s>>>(dist.as<float>(), (uint32_t)n, partial.as<double>()); });
    double *h = static_cast<double *>(thread_pinned(2 * ST_BLOCKS * sizeof(double)));
    CWCU_CHECK(cudaMemcpyAsync(h, partial.p, 2 * ST_BLOCKS * sizeof(double), cudaMemcpyDeviceToHost, s));
    CWCU_CHECK(cudaStreamSynchronize(s));
    double sum = 0.0, sq = 0.0;
    for (int b = 0; b < ST_BLOCKS; b++) {
        sum += h[2 * b];
        sq += h[2 * b + 1];
    }
    // ref: pcl statistical_outlier_removal.hpp -- mean, unbiased variance, threshold in double
    const double dn = (double)n;
    const double mean = sum / dn;
    const double variance = (sq - sum * sum / dn) / (dn - 1.0);
    const double stddev = std::sqrt(variance);
    const double threshold = mean + (double)stddev_mul * stddev;

    Predicate p;
    p.kind = PredKind::DistanceAtMost;
    p.dist = dist.as<float>();
    p.threshold = threshold;
    return compact_points(in, n, out, p, dev, s);
}

} // namespace cwcu
