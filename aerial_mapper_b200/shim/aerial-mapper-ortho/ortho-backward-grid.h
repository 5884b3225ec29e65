// Drop-in replacement of <aerial-mapper-ortho/ortho-backward-grid.h> (reference
// aerial_mapper_ortho/include/aerial-mapper-ortho/ortho-backward-grid.h): same namespace, struct, class and
// signatures (callers: main-ortho-backward-grid.cc:136-141, main-ortho-backward-grid-incremental.cc:134-136,157).
// NB the reference defines three different ortho::Settings in three headers: never include two.
#ifndef ORTHO_BACKWARD_GRID_H_
#define ORTHO_BACKWARD_GRID_H_

#include "../amb_shim_common.h"

namespace ortho {

struct Settings {  // ortho-backward-grid.h:32-41, field for field
  EIGEN_MAKE_ALIGNED_OPERATOR_NEW
  bool show_orthomosaic_opencv = true;
  bool save_orthomosaic_jpg = true;
  std::string orthomosaic_jpg_filename = "";
  double orthomosaic_elevation_m = 0.0;
  bool use_digital_elevation_map = true;
  bool colored_ortho = false;
  bool use_multi_threads = true;
};

class OrthoBackwardGrid {
 public:
  EIGEN_MAKE_ALIGNED_OPERATOR_NEW
  OrthoBackwardGrid(const std::shared_ptr<aslam::NCamera> ncameras, const Settings& settings,
                    grid_map::GridMap* map = nullptr)
      : ncameras_(ncameras), settings_(settings), context_(new amb_shim::Context()) {
    CHECK(ncameras_);  // ortho-backward-grid.cc:27
    (void)map;         // the reference only builds its sample table from it (:30-39)
  }

  // ortho-backward-grid.cc:223-239
  void process(const Poses& T_G_Bs, const Images& images, grid_map::GridMap* map) const {
    CHECK(!T_G_Bs.empty());
    CHECK(T_G_Bs.size() == images.size());
    CHECK(map);
    const aslam::Camera& camera = ncameras_->getCamera(kFrameIdx);
    amb_camera cam;
    cam.width = static_cast<int32_t>(camera.imageWidth());
    cam.height = static_cast<int32_t>(camera.imageHeight());
    cam.fu = camera.getParameters()(0);  // aslam::PinholeCamera parameters: fu, fv, cu, cv
    cam.fv = camera.getParameters()(1);
    cam.cu = camera.getParameters()(2);
    cam.cv = camera.getParameters()(3);
    cam.reserved_ = 0;
    for (int k = 0; k < 4; ++k) cam.dist[k] = 0.0;
    switch (camera.getDistortion().getType()) {
      case aslam::Distortion::Type::kNoDistortion: cam.dist_type = AMB_DIST_NONE; break;
      case aslam::Distortion::Type::kRadTan: cam.dist_type = AMB_DIST_RADTAN; break;
      case aslam::Distortion::Type::kEquidistant: cam.dist_type = AMB_DIST_EQUIDISTANT; break;
      case aslam::Distortion::Type::kFisheye: cam.dist_type = AMB_DIST_FOV; break;
      default: LOG(FATAL) << "aerial_mapper_b200: unsupported distortion model"; cam.dist_type = AMB_DIST_NONE;
    }
    if (cam.dist_type != AMB_DIST_NONE)
      for (int k = 0; k < 4; ++k) cam.dist[k] = camera.getDistortion().getParameters()(k);
    const auto& T_C_B = ncameras_->get_T_C_B(kFrameIdx);
    cam.q_C_B[0] = T_C_B.getRotation().w();
    cam.q_C_B[1] = T_C_B.getRotation().x();
    cam.q_C_B[2] = T_C_B.getRotation().y();
    cam.q_C_B[3] = T_C_B.getRotation().z();
    for (int k = 0; k < 3; ++k) cam.t_C_B[k] = T_C_B.getPosition()(k);

    const size_t n = T_G_Bs.size();
    std::vector<double> poses(7 * n);
    std::vector<const uint8_t*> rasters(n);
    const int channels = settings_.colored_ortho ? 3 : 1;
    for (size_t i = 0; i < n; ++i) {
      for (int k = 0; k < 3; ++k) poses[7 * i + k] = T_G_Bs[i].getPosition()(k);
      poses[7 * i + 3] = T_G_Bs[i].getRotation().w();
      poses[7 * i + 4] = T_G_Bs[i].getRotation().x();
      poses[7 * i + 5] = T_G_Bs[i].getRotation().y();
      poses[7 * i + 6] = T_G_Bs[i].getRotation().z();
      CHECK(images[i].rows == cam.height && images[i].cols == cam.width && images[i].channels() == channels);
      CHECK(static_cast<size_t>(images[i].step) == static_cast<size_t>(images[0].step));
      rasters[i] = images[i].data;
    }
    amb_shim::Trace trace("OrthoBackwardGrid::process");
    context_->get(*map);
    const char* out = settings_.colored_ortho ? "colored_ortho" : "ortho";
    const int out_id = settings_.colored_ortho ? AMB_LAYER_COLORED_ORTHO : AMB_LAYER_ORTHO;
    context_->upload(map, "elevation", AMB_LAYER_ELEVATION);
    context_->upload(map, "elevation_angle", AMB_LAYER_ELEVATION_ANGLE);
    context_->upload(map, "observation_index", AMB_LAYER_OBSERVATION_INDEX);
    context_->upload(map, out, out_id);
    trace.step("uploads (elevation, elevation_angle, observation_index, output layer)");
    context_->orthoProcess(&cam, poses.data(), rasters.data(), n, channels, static_cast<size_t>(images[0].step),
                           settings_.colored_ortho ? 1 : 0);
    trace.step("amb_ortho_process");
    context_->download(map, "elevation_angle", AMB_LAYER_ELEVATION_ANGLE);
    context_->download(map, "observation_index", AMB_LAYER_OBSERVATION_INDEX);
    context_->download(map, out, out_id);
    trace.step("downloads (elevation_angle, observation_index, output layer)");
  }

 private:
  std::shared_ptr<aslam::NCamera> ncameras_;
  static constexpr size_t kFrameIdx = 0u;
  Settings settings_;
  std::shared_ptr<amb_shim::Context> context_;  // process() is const in the reference
};

}  // namespace ortho
#endif  // ORTHO_BACKWARD_GRID_H_
