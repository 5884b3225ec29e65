/*
 * dsm_cell_loop.h — the reference's per-cell DSM loop restated once (TEST INFRASTRUCTURE; see amb_oracle.h),
 * parameterised by the radius-search back end:
 *   - dsm_oracle.cc      : dependency-free exact bucket search (the portable restatement)
 *   - ref_nanoflann_dsm.cc: the reference's vendored nanoflann.hpp, compiled verbatim (oracle/_ref)
 *
 * Follows dsm.cc:113-184 (multi-thread lambda; the single-thread twin dsm.cc:54-111 has the same arithmetic).
 *
 * Searcher contract (mirrors nanoflann::RadiusResultSet + findNeighbors, nanoflann.hpp:134-183,929-946):
 *   void search(double threshold, double qx, double qy, std::vector<std::pair<int,double>>* out) const
 *   clears *out, then appends (point index, d2) for every point with d2 < threshold (strict), where
 *   d2 = 0 + (qx-px)*(qx-px); d2 += (qy-py)*(qy-py)  — the patched 2-D L2_Adaptor (nanoflann.hpp:317-328).
 *   The order of *out is the back end's traversal order (unsorted, as findNeighbors leaves it).
 */
#ifndef AMB_ORACLE_DSM_CELL_LOOP_H_
#define AMB_ORACLE_DSM_CELL_LOOP_H_

#include <atomic>
#include <utility>
#include <vector>

#include "oracle_common.h"

namespace ambo {

struct DsmPoint {
  double x, y, z;
}; /* PointCloud<double>::Point, utils-nearest-neighbor.h:24-31 */

/* dsm.cc:39-45: x = p(0) - center_northing, y = p(1) - center_easting (sic: the swap is the reference's). */
inline void fillShiftedPoints(const double* xyz, size_t n, double center_easting, double center_northing,
                              std::vector<DsmPoint>* pts) {
  pts->resize(n);
  for (size_t i = 0; i < n; ++i) {
    (*pts)[i].x = xyz[3 * i + 0] - center_northing;
    (*pts)[i].y = xyz[3 * i + 1] - center_easting;
    (*pts)[i].z = xyz[3 * i + 2];
  }
}

template <typename Searcher>
int runDsmCellLoop(const amb_geometry& g, float* elevation, const std::vector<DsmPoint>& pts,
                   const Searcher& searcher, int32_t interpolation_radius, int32_t num_threads,
                   int64_t cell_begin, int64_t cell_end, int32_t* neighbour_count, int8_t* threshold_index) {
  const int rows = g.rows;
  std::atomic<int> status(AMB_OK);

  auto cells = [&](int64_t lo, int64_t hi) {
    std::vector<std::pair<int, double> > indices_dists;
    for (int64_t k = cell_begin + lo; k < cell_begin + hi; ++k) {
      /* GridMapIterator linear index -> Index (column-major), dsm.cc:120. */
      const int i = static_cast<int>(k % rows);
      const int j = static_cast<int>(k / rows);
      double qx, qy;
      cellPosition(g, i, j, &qx, &qy); /* dsm.cc:124-125 */

      /* dsm.cc:127-131: RadiusResultSet(settings_.interpolation_radius) — int converted to double. */
      searcher.search(static_cast<double>(interpolation_radius), qx, qy, &indices_dists);
      int level = 0;
      /* dsm.cc:133-144.  `tmp` shares (and clears) the same vector, so result_set.size() sees tmp's hits. */
      {
        double lambda = 1.0;
        int it = 0;
        while (indices_dists.size() == 0u) {
          searcher.search(lambda * interpolation_radius, qx, qy, &indices_dists);
          level = it;
          ++it;
          lambda *= 1.1;
          if (lambda * interpolation_radius > 7.0) break;
        }
      }
      const size_t cnt = indices_dists.size();
      if (neighbour_count) neighbour_count[k] = static_cast<int32_t>(cnt);
      if (threshold_index) threshold_index[k] = cnt > 0 ? static_cast<int8_t>(level) : static_cast<int8_t>(-1);
      if (cnt > 0u) {
        /* dsm.cc:148-172 */
        double idw_numerator = 0.0;
        double idw_denominator = 0.0;
        for (const std::pair<int, double>& s : indices_dists) {
          const double distance = s.second;
          const double height = pts[s.first].z;
          if (!(distance > 0.0)) { /* CHECK(distances[i] > 0.0), dsm.cc:165 -> abort in the reference */
            status.store(AMB_ERR_COINCIDENT_POINT);
            continue;
          }
          idw_numerator += height / distance;
          idw_denominator += 1.0 / distance;
        }
        const double idw_height = idw_numerator / idw_denominator;
        elevation[k] = static_cast<float>(idw_height); /* layer_elevation(x, y) = idw_height, dsm.cc:172 */
      }
    }
  };

  const int64_t n_cells = cell_end - cell_begin;
  if (num_threads < 0) {
    cells(0, n_cells);
  } else {
    parFor(n_cells, cells, resolveThreads(num_threads)); /* dsm.cc:177-179 */
  }
  return status.load();
}

/* ortho::OrthoFromPcl::process cell loop (ortho-from-pcl.cc:52-108).  pts[i].z holds double(intensities[i]). */
template <typename Searcher>
int runOrthoFromPclCellLoop(const amb_geometry& g, float* ortho, const std::vector<DsmPoint>& pts,
                            const Searcher& searcher, int32_t interpolation_radius, bool adaptive,
                            int32_t num_threads, int64_t cell_begin, int64_t cell_end) {
  const int rows = g.rows;
  auto cells = [&](int64_t lo, int64_t hi) {
    std::vector<std::pair<int, double> > indices_dists;
    for (int64_t k = cell_begin + lo; k < cell_begin + hi; ++k) {
      const int i = static_cast<int>(k % rows);
      const int j = static_cast<int>(k / rows);
      double qx, qy;
      cellPosition(g, i, j, &qx, &qy);                                                      /* :53-54 */
      searcher.search(static_cast<double>(interpolation_radius), qx, qy, &indices_dists);  /* :58-61 */
      if (adaptive) { /* :63-72 — int lambda = 10; lambda *= 10 (unbounded in the reference) */
        long long lambda = 10;
        while (indices_dists.size() == 0u && lambda < (1ll << 50)) {
          searcher.search(static_cast<double>(lambda * interpolation_radius), qx, qy, &indices_dists);
          lambda *= 10;
        }
      }
      if (indices_dists.size() > 0u) { /* :73-105 */
        double idw_numerator = 0.0;
        double idw_denominator = 0.0;
        bool idw_perfect_match = false;
        for (const std::pair<int, double>& s : indices_dists) {
          const double distance = s.second;
          const double height = pts[s.first].z;
          if (distance == 0.0) { /* perfect match, no interpolation needed (:90-96) */
            idw_numerator = height;
            idw_denominator = 1.0;
            idw_perfect_match = true;
          }
          if (!idw_perfect_match) {
            idw_numerator += height / distance;
            idw_denominator += 1.0 / distance;
          }
        }
        const double idw_height = idw_numerator / idw_denominator;
        ortho[k] = static_cast<float>(idw_height); /* layer_ortho(x, y) = idw_height, :104 */
      }
    }
  };
  const int64_t n_cells = cell_end - cell_begin;
  if (num_threads <= 0) {
    cells(0, n_cells); /* the reference loop is single-threaded */
  } else {
    parFor(n_cells, cells, resolveThreads(num_threads));
  }
  return AMB_OK;
}

}  // namespace ambo
#endif
