/*
 * stereo_oracle.cc — CPU restatement of stereo::Densifier::computePointCloud (TEST INFRASTRUCTURE; see
 * amb_oracle.h).  Reference: aerial_mapper_dense_pcl/src/densifier.cpp:25-108 ("next" row N3).  Only the outputs
 * the mapping path consumes are produced (point_cloud_eigen, point_cloud_intensities); the ROS message fill is
 * visualisation transport.
 */
#include <cmath>
#include <cstddef>
#include <cstdint>

#include "amb_oracle.h"

extern "C" int ambo_stereo_reproject(const float* disparity, size_t disparity_stride, const uint8_t* image_left,
                                     size_t image_stride, int32_t width, int32_t height, const double* K,
                                     double baseline, const double* R_G_C, const double* t_G_C1,
                                     float max_invalid_disparity, double* out_xyz, int32_t* out_intensity,
                                     size_t capacity, size_t* out_count) {
  if (!disparity || !image_left || !K || !R_G_C || !t_G_C1 || !out_xyz || !out_intensity || !out_count)
    return AMB_ERR_INVALID_ARGUMENT;
  if (baseline == 0.0) return AMB_ERR_CHECK_FAILED; /* CHECK_NE(baseline, 0.0), :39 */
  const double fx = K[0], fy = K[1], cx = K[2], cy = K[3];
  /* Q = [1 0 0 -cx; 0 fx/fy 0 -cy*(fx/fy); 0 0 0 fx; 0 0 1/baseline 0]  (:45-47) */
  const double q03 = -cx, q11 = fx / fy, q13 = -cy * (fx / fy), q23 = fx, q32 = 1.0 / baseline;
  size_t n = 0;
  for (int v = 0; v < height; ++v) {
    const float* drow = disparity + static_cast<size_t>(v) * disparity_stride;
    const uint8_t* irow = image_left + static_cast<size_t>(v) * image_stride;
    for (int u = 0; u < width; ++u) {
      if (drow[u] > max_invalid_disparity) {          /* :61 */
        const double w = q32 * drow[u];               /* :63 */
        const double x1 = (u + q03) / w;              /* :69-70 */
        const double y1 = (q11 * v + q13) / w;
        const double z1 = q23 / w;
        /* rectified_stereo_pair.R_G_C * point_r1 + stereo_pair.t_G_C1 (:73-74) */
        const double X = ((R_G_C[0] * x1 + R_G_C[1] * y1) + R_G_C[2] * z1) + t_G_C1[0];
        const double Y = ((R_G_C[3] * x1 + R_G_C[4] * y1) + R_G_C[5] * z1) + t_G_C1[1];
        const double Z = ((R_G_C[6] * x1 + R_G_C[7] * y1) + R_G_C[8] * z1) + t_G_C1[2];
        const float z = static_cast<float>(Z);
        if (!std::isinf(z)) {                         /* :78 */
          if (n < capacity) {
            out_xyz[3 * n + 0] = X;                   /* point_cloud_eigen.push_back(point_G), :93 */
            out_xyz[3 * n + 1] = Y;
            out_xyz[3 * n + 2] = Z;
            out_intensity[n] = irow[u];               /* point_cloud_intensities.push_back(gray), :94 */
          }
          ++n;
        }
      }
    }
  }
  *out_count = n;
  return AMB_OK;
}
