/*
 * refsrc_glue_pcl.cc — C entry point around the REFERENCE'S OWN ortho::OrthoFromPcl (TEST INFRASTRUCTURE; see
 * amb_oracle.h and refsrc_glue_main.cc).  A separate shared object (_ref/libamb_refsrc_pcl.so) because
 * ortho-from-pcl.h and ortho-backward-grid.h both define `ortho::Settings` with different members.
 *
 *   /root/reference/aerial_mapper_ortho/src/ortho-from-pcl.cc + aerial_mapper_utils/src/utils-common.cc, verbatim,
 *   against refsrc_stubs/;  API: ortho::OrthoFromPcl(settings).process(pointcloud, intensities, &map)
 *   (ortho-from-pcl.h:36-45).
 */
#include <aerial-mapper-ortho/ortho-from-pcl.h>

#define AMB_EXPORT extern "C" __attribute__((visibility("default")))

namespace {
thread_local std::string g_last_error;
}

AMB_EXPORT const char* ambo_refsrc_pcl_last_error(void) { return g_last_error.c_str(); }

AMB_EXPORT int ambo_refsrc_ortho_from_pcl_process(const amb_geometry* geom, float* ortho, const double* xyz,
                                                  const int32_t* intensities, size_t n,
                                                  int32_t interpolation_radius, int32_t use_adaptive_interpolation,
                                                  int64_t cell_begin, int64_t cell_end, double* seconds) {
  if (!geom || !ortho || geom->rows <= 0 || geom->cols <= 0) return AMB_ERR_INVALID_ARGUMENT;
  if (n > 0 && (!xyz || !intensities)) return AMB_ERR_INVALID_ARGUMENT;
  const int64_t total = static_cast<int64_t>(geom->rows) * geom->cols;
  if (cell_begin < 0 || cell_end > total || cell_begin > cell_end) return AMB_ERR_SIZE_MISMATCH;
  try {
    grid_map::GridMap map(*geom);
    map.ambAddLayer("ortho", ortho);
    map.ambSetIterationRange(cell_begin, cell_end);
    AlignedType<std::vector, Eigen::Vector3d>::type pointcloud(n);
    for (size_t i = 0; i < n; ++i) pointcloud[i] = Eigen::Vector3d(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);
    const std::vector<int> point_intensities(intensities, intensities + n);

    ortho::Settings settings;
    settings.interpolation_radius = interpolation_radius;
    settings.use_adaptive_interpolation = use_adaptive_interpolation != 0;

    const double t0 = ambo::now();
    ortho::OrthoFromPcl mosaic(settings);
    mosaic.process(pointcloud, point_intensities, &map);
    if (seconds) seconds[0] = ambo::now() - t0;
  } catch (const ambref::CheckFailed& e) {
    g_last_error = e.what();
    return AMB_ERR_CHECK_FAILED;
  } catch (const std::exception& e) {
    g_last_error = e.what();
    return AMB_ERR_INVALID_ARGUMENT;
  }
  return AMB_OK;
}
