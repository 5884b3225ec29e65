// Drop-in replacement of <aerial-mapper-dsm/dsm.h> (reference aerial_mapper_dsm/include/aerial-mapper-dsm/dsm.h):
// same namespace, struct, class and signatures, so aerial_mapper_demos (main-dsm.cc:103-107,
// main-ortho-backward-grid.cc:129-133, main-ortho-backward-grid-incremental.cc:117-120,153) compile unchanged;
// the work happens on the GPU behind include/aerial_mapper_b200.h.  Link with -laerial_mapper_b200.
#ifndef DSM_H_
#define DSM_H_

#include "../amb_shim_common.h"

namespace dsm {

struct Settings {  // dsm.h:25-32, field for field
  EIGEN_MAKE_ALIGNED_OPERATOR_NEW
  int interpolation_radius = 1.0;
  bool adaptive_interpolation = false;
  double center_easting = 0.0;
  double center_northing = 0.0;
  bool use_multi_threads = true;
};

class Dsm {
 public:
  EIGEN_MAKE_ALIGNED_OPERATOR_NEW

  Dsm(const Settings& settings, grid_map::GridMap* map) : settings_(settings) {
    CHECK(map);  // dsm.cc:22
    context_.get(*map);
  }

  // dsm.cc:186-201
  void process(const AlignedType<std::vector, Eigen::Vector3d>::type& point_cloud, grid_map::GridMap* map) {
    if (point_cloud.empty()) {
      LOG(WARNING) << "Passed empty point cloud to DSM module";  // dsm.cc:189-192
      return;
    }
    CHECK(map);  // dsm.cc:194
    static_assert(sizeof(Eigen::Vector3d) == 3 * sizeof(double), "AoS double[3] expected");
    amb_shim::Trace trace("Dsm::process");
    context_.get(*map);
    context_.upload(map, "elevation", AMB_LAYER_ELEVATION);  // cells without neighbours keep their value
    trace.step("upload (elevation)");
    const double* xyz = &point_cloud[0](0);
    context_.dsmProcess(xyz, point_cloud.size(), settings_.interpolation_radius, settings_.center_easting,
                        settings_.center_northing);
    trace.step("amb_dsm_process");
    context_.download(map, "elevation", AMB_LAYER_ELEVATION);
    trace.step("download (elevation)");
  }

 private:
  Settings settings_;
  amb_shim::Context context_;
};

}  // namespace dsm
#endif  // DSM_H_
