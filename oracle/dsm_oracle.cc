/*
 * dsm_oracle.cc — dependency-free CPU restatement of dsm::Dsm::process (TEST INFRASTRUCTURE; see amb_oracle.h).
 *
 * Reference: aerial_mapper_dsm/src/dsm.cc:186-201 (process), :36-52 (point shift + kd-tree), :113-184 (cell loop).
 * The reference answers each cell's radius query with nanoflann's kd-tree (exact search, eps = 0); the set it
 * returns is exactly { p : d2(p) < threshold } with d2 evaluated by the patched 2-D L2_Adaptor
 * (nanoflann.hpp:317-328).  This file produces the same SET with a uniform bucket grid instead of a tree; only
 * the order of the IDW summation differs (bucket order instead of kd-traversal order), i.e. the last bits of the
 * double sums before the float store.  oracle/_ref (ref_nanoflann_dsm.cc) keeps the reference's own tree and
 * order; tests/test_oracle_dsm.py checks the two against each other.
 */
#include <algorithm>
#include <limits>

#include "dsm_cell_loop.h"

namespace {

using ambo::DsmPoint;

/* Exact radius search over a uniform bucket grid covering the map plus the largest threshold's reach.
 * Points outside that region can never satisfy d2 < threshold for any cell centre and are dropped. */
class BucketSearcher {
 public:
  BucketSearcher(const std::vector<DsmPoint>& pts, const amb_geometry& g, double max_threshold) : pts_(pts) {
    const double reach = std::sqrt(max_threshold) + 4.0 * g.resolution;
    size_ = std::max(g.resolution, std::sqrt(max_threshold) * 0.5);
    double x_hi, y_hi, x_lo, y_lo;
    ambo::cellPosition(g, 0, 0, &x_hi, &y_hi);
    ambo::cellPosition(g, g.rows - 1, g.cols - 1, &x_lo, &y_lo);
    x0_ = x_lo - reach;
    y0_ = y_lo - reach;
    nx_ = static_cast<int64_t>(std::floor((x_hi + reach - x0_) / size_)) + 1;
    ny_ = static_cast<int64_t>(std::floor((y_hi + reach - y0_) / size_)) + 1;
    start_.assign(static_cast<size_t>(nx_ * ny_ + 1), 0);
    std::vector<int64_t> bucket_of(pts.size(), -1);
    for (size_t p = 0; p < pts.size(); ++p) {
      const double fx = std::floor((pts[p].x - x0_) / size_);
      const double fy = std::floor((pts[p].y - y0_) / size_);
      if (!(fx >= 0.0 && fy >= 0.0 && fx < static_cast<double>(nx_) && fy < static_cast<double>(ny_))) continue;
      const int64_t b = static_cast<int64_t>(fx) + static_cast<int64_t>(fy) * nx_;
      bucket_of[p] = b;
      ++start_[static_cast<size_t>(b) + 1];
    }
    for (size_t b = 0; b + 1 < start_.size(); ++b) start_[b + 1] += start_[b];
    order_.resize(static_cast<size_t>(start_.back()));
    std::vector<int64_t> cursor(start_.begin(), start_.end() - 1);
    for (size_t p = 0; p < pts.size(); ++p) { /* stable: ascending original index inside a bucket */
      if (bucket_of[p] < 0) continue;
      order_[static_cast<size_t>(cursor[static_cast<size_t>(bucket_of[p])]++)] = static_cast<int>(p);
    }
  }

  void search(double threshold, double qx, double qy, std::vector<std::pair<int, double> >* out) const {
    out->clear(); /* RadiusResultSet ctor -> init() -> clear(), nanoflann.hpp:140-150 */
    const double reach = std::sqrt(threshold) + size_ * 1e-6 + 1e-9;
    int64_t bx0 = static_cast<int64_t>(std::floor((qx - reach - x0_) / size_));
    int64_t bx1 = static_cast<int64_t>(std::floor((qx + reach - x0_) / size_));
    int64_t by0 = static_cast<int64_t>(std::floor((qy - reach - y0_) / size_));
    int64_t by1 = static_cast<int64_t>(std::floor((qy + reach - y0_) / size_));
    bx0 = std::max<int64_t>(bx0, 0);
    by0 = std::max<int64_t>(by0, 0);
    bx1 = std::min<int64_t>(bx1, nx_ - 1);
    by1 = std::min<int64_t>(by1, ny_ - 1);
    for (int64_t by = by0; by <= by1; ++by) {
      for (int64_t bx = bx0; bx <= bx1; ++bx) {
        const size_t b = static_cast<size_t>(bx + by * nx_);
        for (int64_t s = start_[b]; s < start_[b + 1]; ++s) {
          const int idx = order_[static_cast<size_t>(s)];
          /* L2_Adaptor<double,...>::operator() with size == 2 (nanoflann.hpp:304-330): result = 0;
           * result += diff0*diff0 for dim 0, then dim 1; diff = query - point. */
          double result = 0.0;
          const double diff0 = qx - pts_[idx].x;
          result += diff0 * diff0;
          const double diff1 = qy - pts_[idx].y;
          result += diff1 * diff1;
          /* searchLevel: if (dist < worst_dist) addPoint; addPoint: if (dist < radius) push_back
           * (nanoflann.hpp:1264-1267,156-158); worstDist() == radius. */
          if (result < threshold) out->push_back(std::make_pair(idx, result));
        }
      }
    }
  }

 private:
  const std::vector<DsmPoint>& pts_;
  double size_, x0_, y0_;
  int64_t nx_, ny_;
  std::vector<int64_t> start_;
  std::vector<int> order_;
};

}  // namespace

extern "C" int ambo_dsm_process(const amb_geometry* geom, float* elevation, const double* xyz, size_t n,
                                int32_t interpolation_radius, double center_easting, double center_northing,
                                int32_t num_threads, int64_t cell_begin, int64_t cell_end,
                                int32_t* neighbour_count, int8_t* threshold_index, double* seconds) {
  if (!geom || !elevation || geom->rows <= 0 || geom->cols <= 0) return AMB_ERR_INVALID_ARGUMENT;
  if (n == 0) return AMB_ERR_EMPTY; /* dsm.cc:189-192: LOG(WARNING), return; layers untouched */
  if (!xyz || interpolation_radius < 1) return AMB_ERR_INVALID_ARGUMENT;
  const int64_t total = static_cast<int64_t>(geom->rows) * geom->cols;
  if (cell_begin < 0 || cell_end > total || cell_begin > cell_end) return AMB_ERR_SIZE_MISMATCH;

  const double t0 = ambo::now();
  std::vector<DsmPoint> pts;
  ambo::fillShiftedPoints(xyz, n, center_easting, center_northing, &pts); /* dsm.cc:39-45 */
  const std::vector<double> thr = ambo::dsmThresholds(interpolation_radius);
  BucketSearcher searcher(pts, *geom, *std::max_element(thr.begin(), thr.end()));
  const double t1 = ambo::now();
  const int st = ambo::runDsmCellLoop(*geom, elevation, pts, searcher, interpolation_radius, num_threads,
                                      cell_begin, cell_end, neighbour_count, threshold_index);
  const double t2 = ambo::now();
  if (seconds) {
    seconds[0] = t1 - t0;
    seconds[1] = t2 - t1;
  }
  return st;
}

extern "C" int ambo_ortho_from_pcl_process(const amb_geometry* geom, float* ortho, const double* xyz,
                                           const int32_t* intensities, size_t n, int32_t interpolation_radius,
                                           int32_t use_adaptive_interpolation, int32_t num_threads,
                                           int64_t cell_begin, int64_t cell_end, double* seconds) {
  if (!geom || !ortho || geom->rows <= 0 || geom->cols <= 0) return AMB_ERR_INVALID_ARGUMENT;
  if (n == 0) return AMB_ERR_EMPTY; /* CHECK(!pointcloud.empty()), ortho-from-pcl.cc:23 */
  if (!xyz || !intensities || interpolation_radius < 1) return AMB_ERR_INVALID_ARGUMENT;
  if (use_adaptive_interpolation) return AMB_ERR_UNSUPPORTED; /* unbounded radius: nanoflann back end only */
  const int64_t total = static_cast<int64_t>(geom->rows) * geom->cols;
  if (cell_begin < 0 || cell_end > total || cell_begin > cell_end) return AMB_ERR_SIZE_MISMATCH;
  const double t0 = ambo::now();
  std::vector<DsmPoint> pts(n);
  for (size_t i = 0; i < n; ++i) { /* ortho-from-pcl.cc:29-34: no centre shift, z = double(intensity) */
    pts[i].x = xyz[3 * i + 0];
    pts[i].y = xyz[3 * i + 1];
    pts[i].z = static_cast<double>(intensities[i]);
  }
  BucketSearcher searcher(pts, *geom, static_cast<double>(interpolation_radius));
  const int st = ambo::runOrthoFromPclCellLoop(*geom, ortho, pts, searcher, interpolation_radius, false,
                                               num_threads, cell_begin, cell_end);
  if (seconds) seconds[0] = ambo::now() - t0;
  return st;
}

extern "C" int ambo_hardware_concurrency(void) { return static_cast<int>(ambo::resolveThreads(0)); }
