// filters.cpp -- the C-ABI filter entry points: argument rules, cellsize/timestamp propagation,
// per-tile grouping and error behaviour of the reference, with the arithmetic done by the CUDA
// kernels in pointops.cu / downsample.cu / outliers.cu.
// ref: src/cwipc_filters.cpp:30-418
#include <algorithm>
#include <cstring>

#include "kernels.hpp"
#include "pointcloud.hpp"
#include "radix_sort.hpp"

using namespace cwcu;

namespace {

// Common shape of a one-input filter: resolve storage, run `body(in, dev, stream)` -> StoragePtr,
// wrap the result with the input's timestamp and the given cellsize.
template <class Body>
cwipc_pointcloud *unary_filter(const char *who, cwipc_pointcloud *pc, Body &&body) {
    if (pc == nullptr) return nullptr; // ref: src/cwipc_filters.cpp:282-284 (silent)
    return guarded<cwipc_pointcloud *>(who, nullptr, [&]() -> cwipc_pointcloud * {
        StoragePtr in = storage_of(pc, who);
        if (!in) return nullptr;
        DeviceGuard g(in->dev);
        cudaStream_t s = thread_stream(in->dev);
        in->acquire_for_read(s);
        StoragePtr out;
        try {
            out = body(in, in->dev, s);
        } catch (...) {
            in->release_after_read(s);
            throw;
        }
        in->release_after_read(s);
        if (!out) return nullptr;
        auto *rv = new DevicePointcloud(out, pc->timestamp(), 0.f);
        rv->_set_cellsize(pc->cellsize());
        return rv;
    });
}

// a subset of a cloud lies inside any box that contains the cloud
void inherit_bounds(Storage &out, const Storage &in) {
    if (!in.has_bounds) return;
    out.has_bounds = true;
    memcpy(out.bounds_min, in.bounds_min, sizeof(out.bounds_min));
    memcpy(out.bounds_max, in.bounds_max, sizeof(out.bounds_max));
}

const float *bounds_of(const Storage &in, float box[6]) {
    if (!in.has_bounds) return nullptr;
    memcpy(box, in.bounds_min, 3 * sizeof(float));
    memcpy(box + 3, in.bounds_max, 3 * sizeof(float));
    return box;
}

StoragePtr compact_to_new(const StoragePtr &in, const Predicate &pred, int dev, cudaStream_t s) {
    auto out = std::make_shared<Storage>(dev, in->count, s);
    out->count = compact_points(in->d_pts, in->count, out->d_pts, pred, dev, s);
    inherit_bounds(*out, *in);
    out->mark_ready();
    return out;
}

} // namespace

extern "C" {

// keep iff tile == 0 || point.tile == tile; stable; empty in -> empty out.  ref: src/cwipc_filters.cpp:281-306
cwipc_pointcloud *cwipc_tilefilter(cwipc_pointcloud *pc, int tile) {
    return unary_filter("cwipc_tilefilter", pc, [&](const StoragePtr &in, int dev, cudaStream_t s) -> StoragePtr {
        if (tile == 0) return in; // keeps every point: storage is immutable, so share it
        if (tile < 0 || tile > 255) {
            // `tile == pt.a` with an 8-bit pt.a can never hold: empty result
            auto out = std::make_shared<Storage>(dev, 0, s);
            out->mark_ready();
            return out;
        }
        Predicate p;
        p.kind = PredKind::TileEquals;
        p.tile = tile;
        return compact_to_new(in, p, dev, s);
    });
}

// (tile & mask) != 0, the variant python/cwipc/registration/util.py:98-112 builds in numpy
cwipc_pointcloud *cwipc_cuda_tilefilter_masked(cwipc_pointcloud *pc, int mask) {
    return unary_filter("cwipc_tilefilter_masked", pc, [&](const StoragePtr &in, int dev, cudaStream_t s) -> StoragePtr {
        Predicate p;
        p.kind = PredKind::TileMask;
        p.tile = mask & 0xff;
        return compact_to_new(in, p, dev, s);
    });
}

// ref: src/cwipc_filters.cpp:333-360
cwipc_pointcloud *cwipc_crop(cwipc_pointcloud *pc, float bbox[6]) {
    return unary_filter("cwipc_crop", pc, [&](const StoragePtr &in, int dev, cudaStream_t s) -> StoragePtr {
        Predicate p;
        p.kind = PredKind::CropBox;
        memcpy(p.box, bbox, sizeof(p.box));
        return compact_to_new(in, p, dev, s);
    });
}

// ref: src/cwipc_filters.cpp:308-331
cwipc_pointcloud *cwipc_tilemap(cwipc_pointcloud *pc, uint8_t map[256]) {
    return unary_filter("cwipc_tilemap", pc, [&](const StoragePtr &in, int dev, cudaStream_t s) -> StoragePtr {
        auto out = std::make_shared<Storage>(dev, in->count, s);
        out->count = in->count;
        tilemap_points(in->d_pts, in->count, out->d_pts, map, s);
        out->mark_ready();
        return out;
    });
}

// ref: src/cwipc_filters.cpp:362-386
cwipc_pointcloud *cwipc_colormap(cwipc_pointcloud *pc, uint32_t clearBits, uint32_t setBits) {
    return unary_filter("cwipc_colormap", pc, [&](const StoragePtr &in, int dev, cudaStream_t s) -> StoragePtr {
        auto out = std::make_shared<Storage>(dev, in->count, s);
        out->count = in->count;
        colormap_points(in->d_pts, in->count, out->d_pts, clearBits, setBits, s);
        out->mark_ready();
        return out;
    });
}

// concatenation; timestamp and cellsize are the minima.  ref: src/cwipc_filters.cpp:388-418
cwipc_pointcloud *cwipc_join(cwipc_pointcloud *pc1, cwipc_pointcloud *pc2) {
    if (pc1 == nullptr || pc2 == nullptr) return nullptr;
    return guarded<cwipc_pointcloud *>("cwipc_join", nullptr, [&]() -> cwipc_pointcloud * {
        StoragePtr a = storage_of(pc1, "cwipc_join"), b = storage_of(pc2, "cwipc_join");
        if (!a || !b) return nullptr;
        const int dev = a->dev;
        DeviceGuard g(dev);
        cudaStream_t s = thread_stream(dev);
        a->acquire_for_read(s);
        b->acquire_for_read(s);
        auto out = std::make_shared<Storage>(dev, a->count + b->count, s);
        out->count = a->count + b->count;
        if (a->count) CWCU_CHECK(cudaMemcpyAsync(out->d_pts, a->d_pts, a->count * sizeof(cwipc_point), cudaMemcpyDeviceToDevice, s));
        if (b->count) {
            // a cloud on another device: an explicit peer copy (memory of a private pool is not reachable through UVA alone)
            if (b->dev != dev) CWCU_CHECK(cudaMemcpyPeerAsync(out->d_pts + a->count, dev, b->d_pts, b->dev, b->count * sizeof(cwipc_point), s));
            else CWCU_CHECK(cudaMemcpyAsync(out->d_pts + a->count, b->d_pts, b->count * sizeof(cwipc_point), cudaMemcpyDeviceToDevice, s));
        }
        out->mark_ready();
        a->release_after_read(s);
        b->release_after_read(s);
        auto *rv = new DevicePointcloud(out, std::min(pc1->timestamp(), pc2->timestamp()), 0.f);
        rv->_set_cellsize(std::min(pc1->cellsize(), pc2->cellsize()));
        return rv;
    });
}

// Voxel-grid downsample.  voxelsize > 0: the reference's octree-split path (a grid per 64-voxel
// octree leaf, so voxels cut by a leaf face come out once per leaf); voxelsize < 0: one global grid.
// The cloud's own cellsize wins when it is larger.  ref: src/cwipc_filters.cpp:30-172
cwipc_pointcloud *cwipc_downsample(cwipc_pointcloud *pc, float voxelsize) {
    if (pc == nullptr) return nullptr;
    const bool octree_split = !(voxelsize < 0);
    float cellsize = octree_split ? voxelsize : -voxelsize;
    return guarded<cwipc_pointcloud *>("cwipc_downsample", nullptr, [&]() -> cwipc_pointcloud * {
        const float old = pc->cellsize();
        if (old >= cellsize) cellsize = old; // ref: :42-46, :103-107
        StoragePtr in = storage_of(pc, "cwipc_downsample");
        if (!in) return nullptr;
        DeviceGuard g(in->dev);
        cudaStream_t s = thread_stream(in->dev);
        in->acquire_for_read(s);
        DownsampleResult r;
        try {
            r = downsample_points(in, cellsize, octree_split, in->dev, s);
        } catch (...) {
            in->release_after_read(s);
            throw;
        }
        in->release_after_read(s);
        if (r.failed || !r.out) {
            log(CWIPC_LOG_LEVEL_ERROR, "cwipc_downsample", r.error.empty() ? std::string("downsample failed") : r.error);
            return nullptr;
        }
        auto *rv = new DevicePointcloud(r.out, pc->timestamp(), 0.f);
        rv->_set_cellsize(cellsize);
        return rv;
    });
}

// Statistical outlier removal, whole cloud or once per distinct tile value (first-appearance order,
// tile 0 meaning "every point" exactly as cwipc_tilefilter(pc, 0) does).  ref: src/cwipc_filters.cpp:181-278
cwipc_pointcloud *cwipc_remove_outliers(cwipc_pointcloud *pc, int kNeighbors, float stddevMulThresh, bool perTile) {
    return unary_filter("cwipc_remove_outliers", pc, [&](const StoragePtr &in, int dev, cudaStream_t s) -> StoragePtr {
        const float spacing = pc->cellsize();
        const size_t n = in->count;
        float box[6];
        const float *bounds = bounds_of(*in, box);
        if (!perTile) {
            auto out = std::make_shared<Storage>(dev, n, s);
            out->count = remove_outliers_points(in->d_pts, n, out->d_pts, kNeighbors, stddevMulThresh, spacing, bounds, dev, s);
            inherit_bounds(*out, *in);
            out->mark_ready();
            return out;
        }
        std::vector<int> tiles = tiles_in_first_appearance_order(in->d_pts, n, s);
        const bool has_zero = std::find(tiles.begin(), tiles.end(), 0) != tiles.end();
        auto out = std::make_shared<Storage>(dev, has_zero ? 2 * n : n, s);
        static const bool sequential = getenv("CWIPC_CUDA_SOR_PER_TILE") && !strcmp(getenv("CWIPC_CUDA_SOR_PER_TILE"), "sequential"); // tests only
        if (!has_zero && kNeighbors >= 1 && n > (size_t)kNeighbors && !sequential) {
            // every group at once: one search index with the tile rank as a band of cells (outliers.cu)
            out->count = remove_outliers_per_tile(in->d_pts, n, out->d_pts, tiles, kNeighbors, stddevMulThresh, spacing, bounds, dev, s);
            inherit_bounds(*out, *in);
            out->mark_ready();
            return out;
        }
        // tile 0 among the values (the whole cloud is a group of its own, src/cwipc_filters.cpp:281-306) or fewer points
        // than neighbours: group after group
        Scratch group(n * sizeof(cwipc_point), s);
        size_t total = 0;
        for (int tile : tiles) {
            const cwipc_point *src = in->d_pts;
            size_t cnt = n;
            if (tile != 0) {
                Predicate p;
                p.kind = PredKind::TileEquals;
                p.tile = tile;
                cnt = compact_points(in->d_pts, n, group.as<cwipc_point>(), p, dev, s);
                src = group.as<cwipc_point>();
            }
            total += remove_outliers_points(src, cnt, out->d_pts + total, kNeighbors, stddevMulThresh, spacing, bounds, dev, s);
        }
        out->count = total;
        inherit_bounds(*out, *in);
        out->mark_ready();
        return out;
    });
}

// ---- diagnostic hooks used by the parity tests ---------------------------------------------------
int cwipc_cuda_knn_mean_distances(cwipc_pointcloud *pc, int kNeighbors, float *dist, size_t ndist) {
    if (pc == nullptr || dist == nullptr) return -1;
    return guarded<int>("cwipc_cuda_knn_mean_distances", -1, [&]() -> int {
        StoragePtr in = storage_of(pc, "cwipc_cuda_knn_mean_distances");
        if (!in || ndist < in->count) return -1;
        if (in->count == 0) return 0;
        DeviceGuard g(in->dev);
        cudaStream_t s = thread_stream(in->dev);
        in->acquire_for_read(s);
        Scratch d(in->count * sizeof(float), s);
        float box[6];
        knn_mean_distances(in->d_pts, in->count, kNeighbors, pc->cellsize(), bounds_of(*in, box), d.as<float>(), in->dev, s);
        CWCU_CHECK(cudaMemcpyAsync(dist, d.p, in->count * sizeof(float), cudaMemcpyDeviceToHost, s));
        in->release_after_read(s);
        stream_sync(s);
        return (int)in->count;
    });
}

// ---- partitioned clouds: building blocks for a cloud spread over several GPUs as x-slabs ---------------
// The collective steps (who sends what to whom) are the caller's; see cwipc_util_b200/slab.py.

int cwipc_cuda_octree_replay(cwipc_pointcloud *pc, float cellsize, struct cwipc_cuda_octree_state *state, float bounds[6]) {
    if (pc == nullptr || state == nullptr || bounds == nullptr) return -1;
    return guarded<int>("cwipc_cuda_octree_replay", -1, [&]() -> int {
        StoragePtr in = storage_of(pc, "cwipc_cuda_octree_replay");
        if (!in || !(cellsize > 0.f)) return -1;
        DeviceGuard g(in->dev);
        cudaStream_t s = thread_stream(in->dev);
        in->acquire_for_read(s);
        OctreeState st;
        memcpy(st.min, state->min, sizeof(st.min));
        memcpy(st.max, state->max, sizeof(st.max));
        st.depth = state->depth;
        st.valid = state->valid;
        st.points = state->points;
        octree_replay(in->d_pts, in->count, cellsize, st, bounds, in->dev, s);
        in->release_after_read(s);
        memcpy(state->min, st.min, sizeof(st.min));
        memcpy(state->max, st.max, sizeof(st.max));
        state->depth = st.depth;
        state->valid = st.valid;
        state->points = st.points;
        return 0;
    });
}

cwipc_pointcloud *cwipc_cuda_downsample_planned(cwipc_pointcloud *pc, float voxelsize, const struct cwipc_cuda_octree_state *state, const float bounds[6]) {
    if (pc == nullptr || state == nullptr || bounds == nullptr) return nullptr;
    const bool octree_split = !(voxelsize < 0);
    float cellsize = octree_split ? voxelsize : -voxelsize;
    return guarded<cwipc_pointcloud *>("cwipc_cuda_downsample_planned", nullptr, [&]() -> cwipc_pointcloud * {
        if (pc->cellsize() >= cellsize) cellsize = pc->cellsize();
        StoragePtr in = storage_of(pc, "cwipc_cuda_downsample_planned");
        if (!in) return nullptr;
        DeviceGuard g(in->dev);
        cudaStream_t s = thread_stream(in->dev);
        in->acquire_for_read(s);
        OctreeState st;
        memcpy(st.min, state->min, sizeof(st.min));
        memcpy(st.max, state->max, sizeof(st.max));
        st.depth = state->depth;
        st.valid = state->valid;
        st.points = state->points;
        DownsampleResult r;
        try {
            if (in->count == 0) {
                r.out = std::make_shared<Storage>(in->dev, 0, s); // an empty part contributes nothing
                r.out->mark_ready();
            } else {
                r = downsample_points_planned(in, cellsize, octree_split, st, bounds, in->dev, s);
            }
        } catch (...) {
            in->release_after_read(s);
            throw;
        }
        in->release_after_read(s);
        if (r.failed || !r.out) {
            log(CWIPC_LOG_LEVEL_ERROR, "cwipc_cuda_downsample_planned", r.error.empty() ? std::string("downsample failed") : r.error);
            return nullptr;
        }
        auto *rv = new DevicePointcloud(r.out, pc->timestamp(), 0.f);
        rv->_set_cellsize(cellsize);
        return rv;
    });
}

cwipc_pointcloud *cwipc_cuda_from_device_points(const void *dev_points, int npoint, uint64_t timestamp) {
    if (npoint < 0 || (npoint > 0 && dev_points == nullptr)) return nullptr;
    return guarded<cwipc_pointcloud *>("cwipc_cuda_from_device_points", nullptr, [&]() -> cwipc_pointcloud * {
        const int dev = current_device();
        DeviceGuard g(dev);
        cudaStream_t s = thread_stream(dev);
        auto out = std::make_shared<Storage>(dev, (size_t)npoint, s);
        out->count = (size_t)npoint;
        if (npoint) CWCU_CHECK(cudaMemcpyAsync(out->d_pts, dev_points, (size_t)npoint * sizeof(cwipc_point), cudaMemcpyDefault, s));
        out->mark_ready();
        stream_sync(s); // the caller may reuse its buffer on return
        return new DevicePointcloud(out, timestamp, 0.f);
    });
}

int cwipc_cuda_knn_query(cwipc_pointcloud *pc, int kNeighbors, int nquery, float *mean, float *kth2) {
    if (pc == nullptr || mean == nullptr || nquery < 0) return -1;
    return guarded<int>("cwipc_cuda_knn_query", -1, [&]() -> int {
        StoragePtr in = storage_of(pc, "cwipc_cuda_knn_query");
        if (!in || (size_t)nquery > in->count) return -1;
        if (nquery == 0) return 0;
        DeviceGuard g(in->dev);
        cudaStream_t s = thread_stream(in->dev);
        in->acquire_for_read(s);
        Scratch d(in->count * sizeof(float), s), kth(in->count * sizeof(float), s);
        float box[6];
        knn_mean_distances(in->d_pts, in->count, kNeighbors, pc->cellsize(), bounds_of(*in, box), d.as<float>(), in->dev, s, kth.as<float>(), (size_t)nquery);
        CWCU_CHECK(cudaMemcpyAsync(mean, d.p, (size_t)nquery * sizeof(float), cudaMemcpyDeviceToHost, s));
        if (kth2) CWCU_CHECK(cudaMemcpyAsync(kth2, kth.p, (size_t)nquery * sizeof(float), cudaMemcpyDeviceToHost, s));
        in->release_after_read(s);
        stream_sync(s);
        return nquery;
    });
}

// ---- device-resident per-point distances of a slab (so that only the few open queries ever reach the host) ----
struct cwipc_cuda_distances {
    int dev = 0;
    size_t n = 0;          // queries
    Scratch *mean = nullptr;   // float[n], owned (plain cudaMallocAsync blocks: they outlive the call that made them)
    Scratch *open_idx = nullptr; // uint32[nopen]
    Scratch *kth = nullptr;      // float[n]: (k+1)-th smallest squared distance found so far (+inf where fewer points were seen)
    size_t nopen = 0;
};

namespace {
// scratch that must survive the API call: bypass the per-call arena by allocating on a null-arena path
Scratch *persistent_scratch(size_t bytes, cudaStream_t s) {
    auto *sc = new Scratch();
    sc->p = dmalloc(bytes, s);
    sc->s = s;
    sc->arena = false;
    return sc;
}
} // namespace

cwipc_cuda_distances *cwipc_cuda_knn_query_open(cwipc_pointcloud *pc, int kNeighbors, int nquery, float x_lo, float x_hi, int *nopen) {
    if (pc == nullptr || nquery < 0 || nopen == nullptr) return nullptr;
    return guarded<cwipc_cuda_distances *>("cwipc_cuda_knn_query_open", nullptr, [&]() -> cwipc_cuda_distances * {
        StoragePtr in = storage_of(pc, "cwipc_cuda_knn_query_open");
        if (!in || (size_t)nquery > in->count) return nullptr;
        DeviceGuard g(in->dev);
        cudaStream_t s = thread_stream(in->dev);
        auto *h = new cwipc_cuda_distances();
        h->dev = in->dev;
        h->n = (size_t)nquery;
        *nopen = 0;
        if (nquery == 0) return h;
        in->acquire_for_read(s);
        h->mean = persistent_scratch((size_t)nquery * sizeof(float), s);
        h->open_idx = persistent_scratch((size_t)nquery * sizeof(uint32_t), s);
        h->kth = persistent_scratch((size_t)nquery * sizeof(float), s);
        try {
            if (in->count > (size_t)kNeighbors) {
                Scratch d(in->count * sizeof(float), s), kth(in->count * sizeof(float), s);
                float box[6];
                knn_mean_distances(in->d_pts, in->count, kNeighbors, pc->cellsize(), bounds_of(*in, box), d.as<float>(), in->dev, s, kth.as<float>(), (size_t)nquery);
                CWCU_CHECK(cudaMemcpyAsync(h->mean->p, d.p, (size_t)nquery * sizeof(float), cudaMemcpyDeviceToDevice, s));
                CWCU_CHECK(cudaMemcpyAsync(h->kth->p, kth.p, (size_t)nquery * sizeof(float), cudaMemcpyDeviceToDevice, s));
                h->nopen = mark_open_queries(in->d_pts, kth.as<float>(), (size_t)nquery, x_lo, x_hi, h->open_idx->as<uint32_t>(), in->dev, s);
            } else { // fewer points than neighbours here: every query is open, nothing is known about its neighbourhood
                CWCU_CHECK(cudaMemsetAsync(h->kth->p, 0x7f, (size_t)nquery * sizeof(float), s)); // 0x7f7f7f7f: a huge finite float
                h->nopen = mark_open_queries(in->d_pts, nullptr, (size_t)nquery, x_lo, x_hi, h->open_idx->as<uint32_t>(), in->dev, s);
            }
        } catch (...) {
            in->release_after_read(s);
            delete h->mean;
            delete h->open_idx;
            delete h->kth;
            delete h;
            throw;
        }
        in->release_after_read(s);
        *nopen = (int)h->nopen;
        return h;
    });
}

// indices (into the first nquery points), coordinates and current (k+1)-th squared distances of the open queries, to host
int cwipc_cuda_distances_open(cwipc_cuda_distances *h, cwipc_pointcloud *pc, uint32_t *idx, struct cwipc_point *points, float *kth2) {
    if (h == nullptr || pc == nullptr) return -1;
    return guarded<int>("cwipc_cuda_distances_open", -1, [&]() -> int {
        if (h->nopen == 0) return 0;
        StoragePtr in = storage_of(pc, "cwipc_cuda_distances_open");
        if (!in) return -1;
        DeviceGuard g(h->dev);
        cudaStream_t s = thread_stream(h->dev);
        in->acquire_for_read(s);
        Scratch q(h->nopen * sizeof(cwipc_point), s);
        gather_points(in->d_pts, h->open_idx->as<uint32_t>(), h->nopen, q.as<cwipc_point>(), s);
        if (idx) CWCU_CHECK(cudaMemcpyAsync(idx, h->open_idx->p, h->nopen * sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
        if (points) CWCU_CHECK(cudaMemcpyAsync(points, q.p, h->nopen * sizeof(cwipc_point), cudaMemcpyDeviceToHost, s));
        Scratch kq(h->nopen * sizeof(float), s);
        if (kth2) {
            gather_floats(h->kth->as<float>(), h->open_idx->as<uint32_t>(), h->nopen, kq.as<float>(), s);
            CWCU_CHECK(cudaMemcpyAsync(kth2, kq.p, h->nopen * sizeof(float), cudaMemcpyDeviceToHost, s));
        }
        in->release_after_read(s);
        stream_sync(s);
        return (int)h->nopen;
    });
}

// mean[idx[i]] = values[i] for the open queries (values in the order cwipc_cuda_distances_open returned them)
int cwipc_cuda_distances_patch(cwipc_cuda_distances *h, const float *values, int n) {
    if (h == nullptr || n < 0 || (size_t)n != h->nopen || (n > 0 && values == nullptr)) return -1;
    return guarded<int>("cwipc_cuda_distances_patch", -1, [&]() -> int {
        if (n == 0) return 0;
        DeviceGuard g(h->dev);
        cudaStream_t s = thread_stream(h->dev);
        Scratch v((size_t)n * sizeof(float), s);
        CWCU_CHECK(cudaMemcpyAsync(v.p, values, (size_t)n * sizeof(float), cudaMemcpyHostToDevice, s));
        scatter_floats(v.as<float>(), h->open_idx->as<uint32_t>(), (size_t)n, h->mean->as<float>(), s);
        stream_sync(s);
        return n;
    });
}

int cwipc_cuda_distances_stats(cwipc_cuda_distances *h, double sums[2]) {
    if (h == nullptr || sums == nullptr) return -1;
    return guarded<int>("cwipc_cuda_distances_stats", -1, [&]() -> int {
        DeviceGuard g(h->dev);
        distance_stats(h->n ? h->mean->as<float>() : nullptr, h->n, sums, thread_stream(h->dev));
        return 0;
    });
}

cwipc_pointcloud *cwipc_cuda_distances_filter(cwipc_pointcloud *pc, cwipc_cuda_distances *h, double threshold) {
    if (h == nullptr) return nullptr;
    return unary_filter("cwipc_cuda_distances_filter", pc, [&](const StoragePtr &in, int dev, cudaStream_t s) -> StoragePtr {
        if (h->n != in->count) throw CudaError{cudaErrorInvalidValue, "one distance per point is required"};
        Predicate p;
        p.kind = PredKind::DistanceAtMost;
        p.dist = h->n ? h->mean->as<float>() : nullptr;
        p.threshold = threshold;
        return compact_to_new(in, p, dev, s);
    });
}

void cwipc_cuda_distances_free(cwipc_cuda_distances *h) {
    if (h == nullptr) return;
    try {
        DeviceGuard g(h->dev);
        delete h->mean;
        delete h->open_idx;
        delete h->kth;
    } catch (...) {
    }
    delete h;
}

int cwipc_cuda_knn_lists(cwipc_pointcloud *pc, const struct cwipc_point *queries, const float *limits, int nq, int kNeighbors, float *lists) {
    if (pc == nullptr || nq < 0 || (nq > 0 && (queries == nullptr || lists == nullptr))) return -1;
    return guarded<int>("cwipc_cuda_knn_lists", -1, [&]() -> int {
        StoragePtr in = storage_of(pc, "cwipc_cuda_knn_lists");
        if (!in) return -1;
        if (nq == 0) return 0;
        DeviceGuard g(in->dev);
        cudaStream_t s = thread_stream(in->dev);
        in->acquire_for_read(s);
        const size_t kk = (size_t)kNeighbors + 1;
        Scratch q((size_t)nq * sizeof(cwipc_point), s), l((size_t)nq * kk * sizeof(float), s), lim(limits ? (size_t)nq * sizeof(float) : 0, s);
        CWCU_CHECK(cudaMemcpyAsync(q.p, queries, (size_t)nq * sizeof(cwipc_point), cudaMemcpyHostToDevice, s));
        if (limits) CWCU_CHECK(cudaMemcpyAsync(lim.p, limits, (size_t)nq * sizeof(float), cudaMemcpyHostToDevice, s));
        float box[6];
        knn_lists(in->d_pts, in->count, q.as<cwipc_point>(), limits ? lim.as<float>() : nullptr, (size_t)nq, kNeighbors, pc->cellsize(), bounds_of(*in, box), l.as<float>(), in->dev, s);
        CWCU_CHECK(cudaMemcpyAsync(lists, l.p, (size_t)nq * kk * sizeof(float), cudaMemcpyDeviceToHost, s));
        in->release_after_read(s);
        stream_sync(s);
        return nq;
    });
}

int cwipc_cuda_knn_merge_lists(const float *lists, int nlists, int nq, int kNeighbors, float *mean, float *kth2) {
    if (nq < 0 || nlists < 1 || (nq > 0 && (lists == nullptr || mean == nullptr))) return -1;
    return guarded<int>("cwipc_cuda_knn_merge_lists", -1, [&]() -> int {
        if (nq == 0) return 0;
        const int dev = current_device();
        DeviceGuard g(dev);
        cudaStream_t s = thread_stream(dev);
        const size_t kk = (size_t)kNeighbors + 1, total = (size_t)nlists * nq * kk;
        Scratch l(total * sizeof(float), s), m((size_t)nq * sizeof(float), s), kt((size_t)nq * sizeof(float), s);
        CWCU_CHECK(cudaMemcpyAsync(l.p, lists, total * sizeof(float), cudaMemcpyHostToDevice, s));
        knn_merge_lists(l.as<float>(), (size_t)nlists, (size_t)nq, kNeighbors, m.as<float>(), kt.as<float>(), s);
        CWCU_CHECK(cudaMemcpyAsync(mean, m.p, (size_t)nq * sizeof(float), cudaMemcpyDeviceToHost, s));
        if (kth2) CWCU_CHECK(cudaMemcpyAsync(kth2, kt.p, (size_t)nq * sizeof(float), cudaMemcpyDeviceToHost, s));
        stream_sync(s);
        return nq;
    });
}

int cwipc_cuda_distance_stats(const float *dist, size_t ndist, double sums[2]) {
    if (sums == nullptr || (ndist > 0 && dist == nullptr)) return -1;
    return guarded<int>("cwipc_cuda_distance_stats", -1, [&]() -> int {
        const int dev = current_device();
        DeviceGuard g(dev);
        cudaStream_t s = thread_stream(dev);
        Scratch d(ndist * sizeof(float), s);
        if (ndist) CWCU_CHECK(cudaMemcpyAsync(d.p, dist, ndist * sizeof(float), cudaMemcpyHostToDevice, s));
        distance_stats(d.as<float>(), ndist, sums, s);
        return 0;
    });
}

double cwipc_cuda_outlier_threshold(double sum, double sq, double n, float stddevMulThresh) { return outlier_threshold(sum, sq, n, stddevMulThresh); }

cwipc_pointcloud *cwipc_cuda_filter_by_distance(cwipc_pointcloud *pc, const float *dist, size_t ndist, double threshold) {
    return unary_filter("cwipc_cuda_filter_by_distance", pc, [&](const StoragePtr &in, int dev, cudaStream_t s) -> StoragePtr {
        if (ndist != in->count || (ndist > 0 && dist == nullptr)) throw CudaError{cudaErrorInvalidValue, "one distance per point is required"};
        Scratch d(ndist * sizeof(float), s);
        if (ndist) CWCU_CHECK(cudaMemcpyAsync(d.p, dist, ndist * sizeof(float), cudaMemcpyHostToDevice, s));
        Predicate p;
        p.kind = PredKind::DistanceAtMost;
        p.dist = d.as<float>();
        p.threshold = threshold;
        return compact_to_new(in, p, dev, s);
    });
}

// Diagnostic: the library's stable LSD radix sort on bits [begin_bit, end_bit) of host words, in place.
int cwipc_cuda_sort_u64(uint64_t *words, size_t n, int begin_bit, int end_bit) {
    if (n > 0 && words == nullptr) return -1;
    return guarded<int>("cwipc_cuda_sort_u64", -1, [&]() -> int {
        if (n == 0) return 0;
        const int dev = current_device();
        DeviceGuard g(dev);
        cudaStream_t s = thread_stream(dev);
        Scratch a(n * sizeof(uint64_t), s), b(n * sizeof(uint64_t), s);
        CWCU_CHECK(cudaMemcpyAsync(a.p, words, n * sizeof(uint64_t), cudaMemcpyHostToDevice, s));
        const uint64_t *sorted = radix_sort_u64(a.as<uint64_t>(), b.as<uint64_t>(), n, begin_bit, end_bit, dev, s);
        CWCU_CHECK(cudaMemcpyAsync(words, sorted, n * sizeof(uint64_t), cudaMemcpyDeviceToHost, s));
        stream_sync(s);
        return 0;
    });
}

int cwipc_cuda_downsample_keys(cwipc_pointcloud *pc, float voxelsize, uint64_t *keys, size_t nkeys) {
    if (pc == nullptr || keys == nullptr) return -1;
    return guarded<int>("cwipc_cuda_downsample_keys", -1, [&]() -> int {
        StoragePtr in = storage_of(pc, "cwipc_cuda_downsample_keys");
        if (!in || nkeys < in->count) return -1;
        if (in->count == 0) return 0;
        const bool octree_split = !(voxelsize < 0);
        float cellsize = octree_split ? voxelsize : -voxelsize;
        if (pc->cellsize() >= cellsize) cellsize = pc->cellsize();
        DeviceGuard g(in->dev);
        cudaStream_t s = thread_stream(in->dev);
        in->acquire_for_read(s);
        downsample_keys_to_host(in, cellsize, octree_split, keys, in->dev, s);
        in->release_after_read(s);
        return (int)in->count;
    });
}

} // extern "C"
