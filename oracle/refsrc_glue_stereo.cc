/*
 * refsrc_glue_stereo.cc — C entry point around the REFERENCE'S OWN stereo::Densifier::computePointCloud
 * (TEST INFRASTRUCTURE; see amb_oracle.h and refsrc_glue_main.cc).  Built only where /root/reference is present:
 *   /root/reference/aerial_mapper_dense_pcl/src/densifier.cpp, verbatim, with the dense-pcl headers it includes
 *   (densifier.h, common.h, block-matching-*.h), against refsrc_stubs/  ->  _ref/libamb_refsrc_stereo.so
 * Block matching (cv::StereoBM / StereoSGBM ::compute, block-matching-*.cpp) is OpenCV and not part of this build.
 */
#include <aerial-mapper-dense-pcl/densifier.h>

#define AMB_EXPORT extern "C" __attribute__((visibility("default")))

namespace stereo {
/* ODR-used (address taken, densifier.cpp:97-105) static constexpr members need a definition in C++11 */
constexpr float Densifier::kInvalidPoint;
constexpr int Densifier::kPositionX;
constexpr int Densifier::kPositionY;
constexpr int Densifier::kPositionZ;
constexpr int Densifier::kPositionIntensity;
constexpr int Densifier::kMaxInvalidDisparity;
constexpr size_t Densifier::kSizeOfFloat;
constexpr size_t Densifier::kSizeOfUint32T;
/* block-matching-*.cpp (OpenCV wrappers) are out of scope; the vtables still need these two */
void BlockMatchingBM::computeDisparityMap(const RectifiedStereoPair&, DensifiedStereoPair*) const {
  throw ambref::CheckFailed("block matching is not part of this build");
}
void BlockMatchingSGBM::computeDisparityMap(const RectifiedStereoPair&, DensifiedStereoPair*) const {
  throw ambref::CheckFailed("block matching is not part of this build");
}
}  // namespace stereo

namespace {
thread_local std::string g_last_error;
}

AMB_EXPORT const char* ambo_refsrc_stereo_last_error(void) { return g_last_error.c_str(); }

/* K = {fx, fy, cx, cy}; R_G_C row-major; strides in elements per row; same contract as ambo_stereo_reproject. */
AMB_EXPORT int ambo_refsrc_stereo_reproject(const float* disparity, size_t disparity_stride,
                                            const uint8_t* image_left, size_t image_stride, int32_t width,
                                            int32_t height, const double* K, double baseline, const double* R_G_C,
                                            const double* t_G_C1, double* out_xyz, int32_t* out_intensity,
                                            size_t capacity, size_t* out_count) {
  if (!disparity || !image_left || !K || !R_G_C || !t_G_C1 || !out_xyz || !out_intensity || !out_count)
    return AMB_ERR_INVALID_ARGUMENT;
  try {
    const cv::Size resolution(width, height);
    stereo::StereoRigParameters rig;
    rig.K << K[0], 0.0, K[2], 0.0, K[1], K[3], 0.0, 0.0, 1.0;
    rig.image_size = resolution;
    rig.t_G_C1 = Eigen::Vector3d(t_G_C1[0], t_G_C1[1], t_G_C1[2]);
    stereo::RectifiedStereoPair rectified;
    rectified.baseline = baseline;
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) rectified.R_G_C(i, j) = R_G_C[3 * i + j];
    rectified.image_left = cv::Mat(height, width, image_left, image_stride);
    stereo::DensifiedStereoPair densified;
    densified.disparity_map = cv::Mat(height, width, reinterpret_cast<const uchar*>(disparity),
                                      disparity_stride * sizeof(float));
    sensor_msgs::PointCloud2 message; /* x y z rgb, 16 bytes per point (stereo.cpp); offsets start at point_step */
    message.point_step = 16;
    message.width = static_cast<uint32_t>(width);
    message.height = static_cast<uint32_t>(height);
    message.row_step = message.point_step * message.width;
    message.data.resize(static_cast<size_t>(message.row_step) * message.height + message.point_step);

    stereo::BlockMatchingParameters block_matching;
    stereo::Densifier densifier(block_matching, resolution);
    densifier.computePointCloud(rig, rectified, &densified, message);

    const size_t n = densified.point_cloud_eigen.size();
    *out_count = n;
    for (size_t k = 0; k < n && k < capacity; ++k) {
      for (int c = 0; c < 3; ++c) out_xyz[3 * k + c] = densified.point_cloud_eigen[k](c);
      out_intensity[k] = densified.point_cloud_intensities[k];
    }
  } catch (const ambref::CheckFailed& e) {
    g_last_error = e.what();
    return AMB_ERR_CHECK_FAILED;
  } catch (const std::exception& e) {
    g_last_error = e.what();
    return AMB_ERR_INVALID_ARGUMENT;
  }
  return AMB_OK;
}
