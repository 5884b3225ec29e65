/*
 * refsrc_glue_main.cc — C entry points around the REFERENCE'S OWN dsm::Dsm and ortho::OrthoBackwardGrid
 * (TEST INFRASTRUCTURE; see amb_oracle.h).  Built only where /root/reference is present (Makefile target `refsrc`):
 *
 *   g++ ... -Irefsrc_stubs -I/root/reference/<pkg>/include ...
 *       /root/reference/aerial_mapper_dsm/src/dsm.cc
 *       /root/reference/aerial_mapper_ortho/src/ortho-backward-grid.cc
 *       /root/reference/aerial_mapper_utils/src/utils-common.cc   refsrc_glue_main.cc
 *       -o _ref/libamb_refsrc_main.so
 *
 * The three reference sources (and every reference header they pull in: dsm.h, ortho-backward-grid.h,
 * aerial-mapper-io.h, utils-common.h, utils-nearest-neighbor.h, nanoflann.hpp) are compiled verbatim from where
 * they lie; nothing is copied into this repository.  Their absent third-party dependencies are the stand-ins in
 * refsrc_stubs/ (see amb_refsrc_deps.h for exactly what is restated there).  This file only marshals plain
 * buffers into the reference's argument types and calls its public API:
 *   dsm::Dsm(settings, &map).process(point_cloud, &map)                      dsm.h:36-42
 *   ortho::OrthoBackwardGrid(ncameras, settings, &map).process(T_G_Bs, images, &map)   ortho-backward-grid.h:45-50
 */
#include <aerial-mapper-dsm/dsm.h>
#include <aerial-mapper-ortho/ortho-backward-grid.h>

#define AMB_EXPORT extern "C" __attribute__((visibility("default")))

namespace {

thread_local std::string g_last_error;

int fail(int code, const std::string& what) {
  g_last_error = what;
  return code;
}

bool validRange(const amb_geometry* g, int64_t b, int64_t e) {
  const int64_t total = static_cast<int64_t>(g->rows) * g->cols;
  return b >= 0 && e <= total && b <= e;
}

}  // namespace

AMB_EXPORT const char* ambo_refsrc_last_error(void) { return g_last_error.c_str(); }

/* seconds (nullable double[2]): [0] Dsm constructor (one sample per cell, dsm.cc:23-33),
 *                               [1] process() = kd-tree fill + cell loop (dsm.cc:186-201). */
AMB_EXPORT int ambo_refsrc_dsm_process(const amb_geometry* geom, float* elevation, const double* xyz, size_t n,
                                       int32_t interpolation_radius, double center_easting,
                                       double center_northing, int32_t use_multi_threads, int64_t cell_begin,
                                       int64_t cell_end, double* seconds) {
  if (!geom || !elevation || geom->rows <= 0 || geom->cols <= 0) return AMB_ERR_INVALID_ARGUMENT;
  if (n > 0 && !xyz) return AMB_ERR_INVALID_ARGUMENT;
  if (!validRange(geom, cell_begin, cell_end)) return AMB_ERR_SIZE_MISMATCH;
  try {
    grid_map::GridMap map(*geom);
    map.ambAddLayer("elevation", elevation);
    map.ambSetIterationRange(cell_begin, cell_end);

    AlignedType<std::vector, Eigen::Vector3d>::type point_cloud(n);
    for (size_t i = 0; i < n; ++i) point_cloud[i] = Eigen::Vector3d(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);

    dsm::Settings settings;
    settings.interpolation_radius = interpolation_radius;
    settings.center_easting = center_easting;
    settings.center_northing = center_northing;
    settings.use_multi_threads = use_multi_threads != 0;

    const double t0 = ambo::now();
    dsm::Dsm digital_surface_map(settings, &map);
    const double t1 = ambo::now();
    digital_surface_map.process(point_cloud, &map);
    const double t2 = ambo::now();
    if (seconds) {
      seconds[0] = t1 - t0;
      seconds[1] = t2 - t1;
    }
  } catch (const ambref::CheckFailed& e) {
    return fail(AMB_ERR_CHECK_FAILED, e.what());
  } catch (const std::exception& e) {
    return fail(AMB_ERR_INVALID_ARGUMENT, e.what());
  }
  return AMB_OK;
}

/* All layers full rows*cols column-major.  num_observations / ortho / colored_ortho may be NULL: the reference
 * fetches all six layers by name (ortho-backward-grid.cc:134-139), so a zero-filled scratch layer stands in.
 * seconds (nullable double[2]): [0] constructor, [1] process(). */
AMB_EXPORT int ambo_refsrc_ortho_process(const amb_geometry* geom, const float* elevation, float* elevation_angle,
                                         float* observation_index, float* num_observations, float* ortho,
                                         float* colored_ortho, const amb_camera* camera, const double* T_G_B,
                                         const uint8_t* const* images, size_t n, int32_t channels,
                                         size_t row_step, int32_t colored_ortho_flag, int32_t use_multi_threads,
                                         int64_t cell_begin, int64_t cell_end, double* seconds) {
  if (!geom || !camera || !elevation || !elevation_angle || !observation_index) return AMB_ERR_INVALID_ARGUMENT;
  if (geom->rows <= 0 || geom->cols <= 0) return AMB_ERR_INVALID_ARGUMENT;
  if (n > 0 && (!T_G_B || !images)) return AMB_ERR_INVALID_ARGUMENT;
  if (colored_ortho_flag ? (channels != 3 || !colored_ortho) : (channels != 1 || !ortho))
    return AMB_ERR_SIZE_MISMATCH;
  if (!validRange(geom, cell_begin, cell_end)) return AMB_ERR_SIZE_MISMATCH;
  const size_t total = static_cast<size_t>(geom->rows) * geom->cols;
  /* calloc: untouched pages of a scratch layer are never materialised */
  std::unique_ptr<float, void (*)(void*)> scratch_obs(nullptr, std::free), scratch_color(nullptr, std::free);
  if (!num_observations) {
    scratch_obs.reset(static_cast<float*>(std::calloc(total, sizeof(float))));
    num_observations = scratch_obs.get();
  }
  float** unselected = colored_ortho_flag ? &ortho : &colored_ortho;
  if (!*unselected) {
    scratch_color.reset(static_cast<float*>(std::calloc(total, sizeof(float))));
    *unselected = scratch_color.get();
  }
  if (!num_observations || !ortho || !colored_ortho) return AMB_ERR_INVALID_ARGUMENT;
  try {
    grid_map::GridMap map(*geom);
    map.ambAddLayer("elevation", const_cast<float*>(elevation)); /* read-only in the reference (:137) */
    map.ambAddLayer("elevation_angle", elevation_angle);
    map.ambAddLayer("observation_index", observation_index);
    map.ambAddLayer("num_observations", num_observations);
    map.ambAddLayer("ortho", ortho);
    map.ambAddLayer("colored_ortho", colored_ortho);
    map.ambSetIterationRange(cell_begin, cell_end);

    Poses T_G_Bs;
    Images frames;
    for (size_t i = 0; i < n; ++i) {
      T_G_Bs.push_back(Pose(ambo::tp::poseFromRow(T_G_B + 7 * i)));
      frames.push_back(cv::Mat(camera->height, camera->width, images[i], row_step));
    }
    std::shared_ptr<aslam::NCamera> ncameras(new aslam::NCamera(*camera));

    ortho::Settings settings;
    settings.show_orthomosaic_opencv = false;
    settings.save_orthomosaic_jpg = false;
    settings.colored_ortho = colored_ortho_flag != 0;
    settings.use_multi_threads = use_multi_threads != 0;

    const double t0 = ambo::now();
    ortho::OrthoBackwardGrid mosaic(ncameras, settings, &map);
    const double t1 = ambo::now();
    mosaic.process(T_G_Bs, frames, &map);
    const double t2 = ambo::now();
    if (seconds) {
      seconds[0] = t1 - t0;
      seconds[1] = t2 - t1;
    }
  } catch (const ambref::CheckFailed& e) {
    return fail(AMB_ERR_CHECK_FAILED, e.what());
  } catch (const std::exception& e) {
    return fail(AMB_ERR_INVALID_ARGUMENT, e.what());
  }
  return AMB_OK;
}

/* utils::parFor's thread count (dsm.cc:178, ortho-backward-grid.cc:215). */
AMB_EXPORT int ambo_refsrc_hardware_concurrency(void) { return static_cast<int>(std::thread::hardware_concurrency()); }
