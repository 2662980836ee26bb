/*
 * cwipc_oracle.h -- CPU restatement of the reference's filter hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load
 * this library, and only as the checker (or the timed CPU baseline) -- never as the product path.
 *
 * PARITY UNPINNED at the PCL boundary: the reference's arithmetic for this path lives in PCL
 * (pcl::VoxelGrid, pcl::CentroidPoint, pcl::octree::OctreePointCloud, pcl::StatisticalOutlierRemoval,
 * FLANN; version not pinned by the reference and not present in this environment), the reference
 * cannot be compiled here, and its own tests hold no golden vectors for these filters
 * (python/test_cwipc_util.py:428-450,528-594 check counts and inequalities only).  This file restates
 * PCL's published algorithms from the reference's call sites (src/cwipc_filters.cpp:30-306); the
 * properties the reference's tests do pin are checked in tests/test_oracle.py.
 */
#ifndef CWIPC_ORACLE_H
#define CWIPC_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
    float x, y, z;
    uint8_t r, g, b, tile;
} orc_point; /* == struct cwipc_point, include/cwipc_util/api.h:88-96 */

/* ref: src/cwipc_filters.cpp:281-306.  out capacity n.  Returns the number of points kept. */
long orc_tilefilter(const orc_point *in, size_t n, int tile, orc_point *out);

/* ref: src/cwipc_util.cpp:173-204 (the heuristic behind _set_cellsize(<0)). */
float orc_min_distance_to_first(const orc_point *in, size_t n);

/* ref: src/cwipc_filters.cpp:30-172.  voxelsize > 0: octree split + per-leaf grids; < 0: one grid.
 * pc_cellsize: the input cloud's cellsize() (the larger of the two is used and returned in
 * *out_cellsize).  out capacity n.  Optional outputs (may be NULL):
 *   point_keys[6*i..]  = leaf key x,y,z (0 in single-grid mode) and voxel x,y,z of input point i
 *   out_counts[j]      = number of input points merged into output point j
 * Returns the number of output points, or -1 where the reference returns NULL (single-grid mode on
 * an empty cloud or on a grid whose index would overflow). */
long orc_downsample(const orc_point *in, size_t n, float voxelsize, float pc_cellsize, orc_point *out, float *out_cellsize, int32_t *point_keys, uint32_t *out_counts);

/* First pass of pcl::StatisticalOutlierRemoval: mean distance to the k nearest neighbours (exact kNN
 * with k+1 results including the query, float L2_Simple distances, double sum of square roots).
 * Requires n > k.  Returns 0, or -1 on bad arguments. */
int orc_knn_mean_distances(const orc_point *in, size_t n, int k, float *mean_dist);

/* ref: src/cwipc_filters.cpp:181-278.  out capacity 2n (tile 0 in per-tile mode re-emits the whole
 * cloud).  Returns the number of points kept.  threshold_out (may be NULL) receives the distance
 * threshold of the whole-cloud pass (perTile == 0 only). */
long orc_remove_outliers(const orc_point *in, size_t n, int k, float stddev_mul, int per_tile, orc_point *out, double *threshold_out);

/* bit 0: ORC_LEAF_ORDER_DESCENDING, bit 1: ORC_VOXEL_SORT_UNSTABLE (compile-time switches, see cwipc_oracle.c).
 * The default build returns 0. */
int orc_config(void);

/* O(n^2) kNN used to pin the kd-tree on small inputs. */
int orc_knn_mean_distances_bruteforce(const orc_point *in, size_t n, int k, float *mean_dist);

#ifdef __cplusplus
}
#endif
#endif
