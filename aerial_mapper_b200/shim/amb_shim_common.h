// amb_shim_common.h — glue shared by the two drop-in class headers.
//
// Real build (inside the aerial_mapper catkin workspace): pulls the same third-party headers the reference's own
// headers pull (dsm.h:13-21, ortho-backward-grid.h:14-28).  Stand-alone build (-DAMB_SHIM_MINI, used by this
// repository's tests because Eigen / grid_map / aslam_cv2 / minkindr / OpenCV / glog are absent here): tiny
// stand-ins with the same names and the handful of members the marshalling code touches (mini/amb_mini_deps.h).
#ifndef AMB_SHIM_COMMON_H_
#define AMB_SHIM_COMMON_H_

#include <cstdio>
#include <cstdlib>
#include <memory>
#include <string>
#include <vector>

#include "../../include/aerial_mapper_b200.h"

#ifdef AMB_SHIM_MINI
#include "mini/amb_mini_deps.h"
#else
#include <aerial-mapper-io/aerial-mapper-io.h>            // Pose, Poses, Image, Images
#include <aerial-mapper-utils/utils-nearest-neighbor.h>  // AlignedType
#include <aslam/cameras/camera.h>
#include <aslam/cameras/camera-pinhole.h>
#include <aslam/cameras/ncamera.h>
#include <Eigen/Dense>
#include <glog/logging.h>
#include <grid_map_core/GridMap.hpp>
#endif

namespace amb_shim {

// The reference aborts through glog CHECK on contract violations (SURVEY.md §8b "Error convention"); a non-zero
// status from the C ABI is turned back into that.
inline void checkStatus(int status, const amb_ctx* ctx, const char* what) {
  if (status == AMB_OK) return;
  std::fprintf(stderr, "aerial_mapper_b200: %s failed: %s %s\n", what, amb_status_string(status),
               ctx ? amb_last_error(ctx) : "");
  CHECK(status == AMB_OK);
}

inline amb_geometry geometryOf(const grid_map::GridMap& map) {
  amb_geometry g;
  g.rows = map.getSize()(0);
  g.cols = map.getSize()(1);
  g.resolution = map.getResolution();
  g.length_x = map.getLength()(0);
  g.length_y = map.getLength()(1);
  g.pos_x = map.getPosition()(0);
  g.pos_y = map.getPosition()(1);
  return g;
}

// One device context per (class instance, map geometry) — or, with AMB_SHIM_GPUS=N (N > 1) in the environment, one
// amb_multi spreading the map's column stripes over N GPUs of this process (same results bit for bit, see
// include/aerial_mapper_b200.h "several GPUs of one process").  Host layers stay authoritative like in the reference:
// every process() uploads what it reads and downloads what it writes.
class Context {
 public:
  Context() : ctx_(nullptr), multi_(nullptr) {}
  ~Context() { release(); }
  // creates the backend for this map's geometry if needed; returns the single context (nullptr in multi-GPU mode)
  amb_ctx* get(const grid_map::GridMap& map) {
    const amb_geometry g = geometryOf(map);
    if ((ctx_ || multi_) && (g.rows != geom_.rows || g.cols != geom_.cols || g.resolution != geom_.resolution ||
                             g.pos_x != geom_.pos_x || g.pos_y != geom_.pos_y))
      release();
    if (!ctx_ && !multi_) {
      geom_ = g;
      int gpus = 1;
      if (const char* e = std::getenv("AMB_SHIM_GPUS")) gpus = std::atoi(e);
      if (gpus > 1) {
        checkStatus(amb_multi_create(&g, gpus, &multi_), nullptr, "amb_multi_create");
      } else {
        int device = 0;
        if (const char* e = std::getenv("AMB_DEVICE")) device = std::atoi(e);
        checkStatus(amb_create(&g, device, 0, g.cols, &ctx_), nullptr, "amb_create");
      }
    }
    return ctx_;
  }
  void upload(grid_map::GridMap* map, const char* layer, int id) {
    if (multi_) {
      checkMulti(amb_multi_upload_layer(multi_, id, (*map)[layer].data()), "amb_multi_upload_layer");
    } else {
      checkStatus(amb_upload_layer(ctx_, id, (*map)[layer].data()), ctx_, "amb_upload_layer");
    }
  }
  void download(grid_map::GridMap* map, const char* layer, int id) {
    if (multi_) {
      checkMulti(amb_multi_download_layer(multi_, id, (*map)[layer].data()), "amb_multi_download_layer");
    } else {
      checkStatus(amb_download_layer(ctx_, id, (*map)[layer].data()), ctx_, "amb_download_layer");
    }
  }
  void dsmProcess(const double* xyz, size_t n, int32_t radius, double center_easting, double center_northing) {
    if (multi_) {
      checkMulti(amb_multi_dsm_process(multi_, xyz, n, radius, center_easting, center_northing), "amb_multi_dsm_process");
    } else {
      checkStatus(amb_dsm_process(ctx_, xyz, n, radius, center_easting, center_northing), ctx_, "amb_dsm_process");
    }
  }
  void orthoProcess(const amb_camera* cam, const double* poses, const uint8_t* const* rasters, size_t n, int32_t channels,
                    size_t row_step, int32_t colored) {
    if (multi_) {
      checkMulti(amb_multi_ortho_process(multi_, cam, poses, rasters, n, channels, row_step, colored),
                 "amb_multi_ortho_process");
    } else {
      checkStatus(amb_ortho_process(ctx_, cam, poses, rasters, n, channels, row_step, colored), ctx_, "amb_ortho_process");
    }
  }

 private:
  Context(const Context&);
  Context& operator=(const Context&);
  void release() {
    if (ctx_) amb_destroy(ctx_);
    if (multi_) amb_multi_destroy(multi_);
    ctx_ = nullptr;
    multi_ = nullptr;
  }
  void checkMulti(int status, const char* what) {
    if (status == AMB_OK) return;
    std::fprintf(stderr, "aerial_mapper_b200: %s failed: %s %s\n", what, amb_status_string(status),
                 amb_multi_last_error(multi_));
    CHECK(status == AMB_OK);
  }
  amb_ctx* ctx_;
  amb_multi* multi_;
  amb_geometry geom_;
};

}  // namespace amb_shim
#endif
