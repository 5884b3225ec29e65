/*
 * ref_nanoflann_dsm.cc — the DSM cell loop around the REFERENCE'S OWN neighbour search
 * (TEST INFRASTRUCTURE; see amb_oracle.h).  Built only where /root/reference is present:
 *
 *   g++ ... -I/root/reference/aerial_mapper_thirdparty/include ref_nanoflann_dsm.cc -o _ref/libamb_oracle_ref.so
 *
 * <aerial-mapper-thirdparty/nanoflann.hpp> is compiled verbatim from where it lies under /root/reference
 * (nothing is copied into this repository).  It is the only source file of the reference's hot path that can be
 * built here: dsm.cc itself needs ROS, grid_map, Eigen and glog (all absent), so the loop around the tree is the
 * restatement in dsm_cell_loop.h and the dataset adaptor below restates utils-nearest-neighbor.h:33-76 (that
 * header cannot be included: it pulls in <Eigen/Core>).
 *
 * Tree parameters as dsm.h:57-65 / dsm.cc:47-51: 2-D, L2_Adaptor<double>, kMaxLeaf = 10, SearchParams() default
 * (eps 0, unsorted), RadiusResultSet<double,int>.
 */
#include <aerial-mapper-thirdparty/nanoflann.hpp>

#include <memory>

#include "dsm_cell_loop.h"

namespace {

using ambo::DsmPoint;

struct Cloud {
  typedef double coord_t;
  std::vector<DsmPoint> pts;
};

/* PointCloudAdaptor, utils-nearest-neighbor.h:33-76. */
struct Adaptor {
  typedef double coord_t;
  const Cloud& obj;
  explicit Adaptor(const Cloud& o) : obj(o) {}
  inline size_t kdtree_get_point_count() const { return obj.pts.size(); }
  inline double kdtree_get_pt(const size_t idx, int dim) const {
    if (dim == 0) return obj.pts[idx].x;
    if (dim == 1) return obj.pts[idx].y;
    return obj.pts[idx].z;
  }
  template <class BBOX>
  bool kdtree_get_bbox(BBOX&) const {
    return false;
  }
};

typedef nanoflann::KDTreeSingleIndexAdaptor<nanoflann::L2_Adaptor<double, Adaptor>, Adaptor, 2> KdTree;

class NanoflannSearcher {
 public:
  explicit NanoflannSearcher(const KdTree& tree) : tree_(tree) {}
  void search(double threshold, double qx, double qy, std::vector<std::pair<int, double> >* out) const {
    nanoflann::RadiusResultSet<double, int> result_set(threshold, *out); /* ctor clears *out */
    const double query_pt[3] = {qx, qy, 0.0};                            /* dsm.cc:129 */
    tree_.findNeighbors(result_set, query_pt, nanoflann::SearchParams());
  }

 private:
  const KdTree& tree_;
};

}  // namespace

extern "C" int ambo_ref_dsm_process(const amb_geometry* geom, float* elevation, const double* xyz, size_t n,
                                    int32_t interpolation_radius, double center_easting,
                                    double center_northing, int32_t num_threads, int64_t cell_begin,
                                    int64_t cell_end, int32_t* neighbour_count, int8_t* threshold_index,
                                    double* seconds) {
  if (!geom || !elevation || geom->rows <= 0 || geom->cols <= 0) return AMB_ERR_INVALID_ARGUMENT;
  if (n == 0) return AMB_ERR_EMPTY;
  if (!xyz || interpolation_radius < 1) return AMB_ERR_INVALID_ARGUMENT;
  const int64_t total = static_cast<int64_t>(geom->rows) * geom->cols;
  if (cell_begin < 0 || cell_end > total || cell_begin > cell_end) return AMB_ERR_SIZE_MISMATCH;

  const double t0 = ambo::now();
  Cloud cloud;
  ambo::fillShiftedPoints(xyz, n, center_easting, center_northing, &cloud.pts); /* dsm.cc:39-45 */
  Adaptor adaptor(cloud);
  KdTree tree(2, adaptor, nanoflann::KDTreeSingleIndexAdaptorParams(10)); /* dsm.cc:47-50 */
  tree.buildIndex();                                                     /* dsm.cc:51 */
  const double t1 = ambo::now();
  NanoflannSearcher searcher(tree);
  const int st = ambo::runDsmCellLoop(*geom, elevation, cloud.pts, searcher, interpolation_radius, num_threads,
                                      cell_begin, cell_end, neighbour_count, threshold_index);
  const double t2 = ambo::now();
  if (seconds) {
    seconds[0] = t1 - t0;
    seconds[1] = t2 - t1;
  }
  return st;
}

extern "C" int ambo_ref_ortho_from_pcl_process(const amb_geometry* geom, float* ortho, const double* xyz,
                                               const int32_t* intensities, size_t n,
                                               int32_t interpolation_radius, int32_t use_adaptive_interpolation,
                                               int32_t num_threads, int64_t cell_begin, int64_t cell_end,
                                               double* seconds) {
  if (!geom || !ortho || geom->rows <= 0 || geom->cols <= 0) return AMB_ERR_INVALID_ARGUMENT;
  if (n == 0) return AMB_ERR_EMPTY;
  if (!xyz || !intensities || interpolation_radius < 1) return AMB_ERR_INVALID_ARGUMENT;
  const int64_t total = static_cast<int64_t>(geom->rows) * geom->cols;
  if (cell_begin < 0 || cell_end > total || cell_begin > cell_end) return AMB_ERR_SIZE_MISMATCH;
  const double t0 = ambo::now();
  Cloud cloud;
  cloud.pts.resize(n);
  for (size_t i = 0; i < n; ++i) { /* ortho-from-pcl.cc:29-34 */
    cloud.pts[i].x = xyz[3 * i + 0];
    cloud.pts[i].y = xyz[3 * i + 1];
    cloud.pts[i].z = static_cast<double>(intensities[i]);
  }
  Adaptor adaptor(cloud);
  KdTree tree(2, adaptor, nanoflann::KDTreeSingleIndexAdaptorParams(10)); /* :37-46 */
  tree.buildIndex();
  NanoflannSearcher searcher(tree);
  const int st = ambo::runOrthoFromPclCellLoop(*geom, ortho, cloud.pts, searcher, interpolation_radius,
                                               use_adaptive_interpolation != 0, num_threads, cell_begin, cell_end);
  if (seconds) seconds[0] = ambo::now() - t0;
  return st;
}
