/*
 * ortho_oracle.cc — CPU restatement of ortho::OrthoBackwardGrid::process (TEST INFRASTRUCTURE; see amb_oracle.h).
 *
 * Reference: aerial_mapper_ortho/src/ortho-backward-grid.cc:223-239 (process: pose composition + dispatch),
 * :128-221 (multi-thread cell loop; the single-thread twin :42-126 has the same arithmetic).
 *
 * External arithmetic (minkindr, Eigen quaternions, aslam_cv2 pinhole + distortion, grid_map colour packing — sources
 * NOT under /root/reference) lives in thirdparty_math.h, shared with the stand-in headers oracle/_ref's build of the
 * reference's own ortho-backward-grid.cc is compiled against (refsrc_stubs/).
 */
#include <atomic>

#include "oracle_common.h"
#include "thirdparty_math.h"

namespace {

using namespace ambo::tp; /* Quat, Vec3, Transformation, project3, ... : the absent dependencies' arithmetic */

/* ortho-backward-grid.cc:164-171 */
inline bool keypointVisible(const amb_camera& cam, ProjectionStatus st, double kx, double ky) {
  return (kx >= 0.0) && (ky >= 0.0) && (kx < static_cast<double>(cam.width)) &&
         (ky < static_cast<double>(cam.height)) && (st != POINT_BEHIND_CAMERA) && (st != PROJECTION_INVALID);
}

inline uint32_t packColor(uint8_t b, uint8_t g, uint8_t r) {
  /* ortho-backward-grid.cc:196-201: Eigen::Vector3f(float(rgb[2])/255.0, float(rgb[1])/255.0,
   * float(rgb[0])/255.0) — float / double literal evaluates in double, stored as float; then
   * grid_map::colorVectorToValue (thirdparty_math.h). */
  const float fr = static_cast<float>(static_cast<double>(static_cast<float>(r)) / 255.0);
  const float fg = static_cast<float>(static_cast<double>(static_cast<float>(g)) / 255.0);
  const float fb = static_cast<float>(static_cast<double>(static_cast<float>(b)) / 255.0);
  return colorVectorToBits(fr, fg, fb);
}

}  // namespace

extern "C" int ambo_ortho_process(const amb_geometry* geom, const float* elevation, float* elevation_angle,
                                  float* observation_index, float* ortho, float* colored_ortho,
                                  const amb_camera* camera, const double* T_G_B, const uint8_t* const* images,
                                  size_t n, int32_t channels, size_t row_step, int32_t colored_ortho_flag,
                                  int32_t num_threads, int64_t cell_begin, int64_t cell_end, double* seconds) {
  if (!geom || !camera || !elevation || !elevation_angle || !observation_index) return AMB_ERR_INVALID_ARGUMENT;
  if (n == 0) return AMB_ERR_EMPTY; /* CHECK(!T_G_Bs.empty()), ortho-backward-grid.cc:225 */
  if (!T_G_B || !images) return AMB_ERR_INVALID_ARGUMENT;
  if (colored_ortho_flag ? (channels != 3 || !colored_ortho) : (channels != 1 || !ortho))
    return AMB_ERR_SIZE_MISMATCH;
  const int64_t total = static_cast<int64_t>(geom->rows) * geom->cols;
  if (cell_begin < 0 || cell_end > total || cell_begin > cell_end) return AMB_ERR_SIZE_MISMATCH;
  const amb_camera& cam = *camera;
  const amb_geometry& g = *geom;

  /* ortho-backward-grid.cc:230-233: T_G_C = T_G_B * T_C_B(0)^-1 */
  const Transformation T_B_C = cameraExtrinsics(cam).inverse();
  std::vector<Transformation> T_G_Cs(n);
  for (size_t i = 0; i < n; ++i) T_G_Cs[i] = poseFromRow(T_G_B + 7 * i) * T_B_C;

  std::atomic<int> status(AMB_OK);
  const int rows = g.rows;
  const double t0 = ambo::now();

  auto cells = [&](int64_t lo, int64_t hi) {
    for (int64_t k = cell_begin + lo; k < cell_begin + hi; ++k) {
      const int ci = static_cast<int>(k % rows);
      const int cj = static_cast<int>(k / rows);
      double px, py;
      ambo::cellPosition(g, ci, cj, &px, &py); /* :149-150 */
      const Vec3 landmark = {px, py, static_cast<double>(elevation[k])}; /* :152-153 */

      for (size_t i = 0; i < n; ++i) { /* :156 */
        /* The reference recomputes inverse() per cell x frame (:157-158); same value every time. */
        const Vec3 C_landmark = T_G_Cs[i].inverse().transform(landmark);
        double kx, ky;
        const ProjectionStatus st = project3(cam, C_landmark, &kx, &ky); /* :159-161 */
        if (!keypointVisible(cam, st, kx, ky)) continue;                 /* :164-172 */
        const Vec3& u = C_landmark;
        const double norm_u = std::sqrt(u.x * u.x + u.y * u.y + u.z * u.z); /* :175 */
        const double alpha = std::asin(std::fabs(u.z) / norm_u);           /* :177 */
        if (!(alpha > 0.0)) { /* CHECK(alpha > 0.0), :178 */
          status.store(AMB_ERR_CHECK_FAILED);
          continue;
        }
        /* :180 — double alpha against the float32 layer value promoted to double */
        if (std::fabs(alpha) > static_cast<double>(elevation_angle[k])) {
          elevation_angle[k] = static_cast<float>(std::fabs(alpha)); /* :181 */
          observation_index[k] = static_cast<float>(i);              /* :182 */
          /* :183 num_observations += num_observations: 0 stays 0 — layer untouched. */
          /* :186-193 — second projection gives the same keypoint; round half away from zero, clamp. */
          const int kp_y = std::min(static_cast<int>(std::round(ky)), cam.height - 1);
          const int kp_x = std::min(static_cast<int>(std::round(kx)), cam.width - 1);
          const uint8_t* px_ptr =
              images[i] + static_cast<size_t>(kp_y) * row_step + static_cast<size_t>(kp_x) * channels;
          if (colored_ortho_flag) { /* :194-202 */
            const uint32_t packed = packColor(px_ptr[0], px_ptr[1], px_ptr[2]);
            std::memcpy(&colored_ortho[k], &packed, sizeof(float));
          } else { /* :203-206 */
            const double gray_value = px_ptr[0];
            ortho[k] = static_cast<float>(gray_value);
          }
        }
      }
    }
  };

  const int64_t n_cells = cell_end - cell_begin;
  if (num_threads < 0) {
    cells(0, n_cells);
  } else {
    ambo::parFor(n_cells, cells, ambo::resolveThreads(num_threads)); /* :214-216 */
  }
  if (seconds) seconds[0] = ambo::now() - t0;
  return status.load();
}

extern "C" int ambo_project3(const amb_camera* camera, const double* p_C, double* keypoint) {
  const Vec3 p = {p_C[0], p_C[1], p_C[2]};
  double kx, ky;
  const ProjectionStatus st = project3(*camera, p, &kx, &ky);
  keypoint[0] = kx;
  keypoint[1] = ky;
  return keypointVisible(*camera, st, kx, ky) ? 1 : 0;
}

extern "C" int ambo_transform_to_camera(const amb_camera* camera, const double* T_G_B, const double* p_G,
                                        double* p_C) {
  const Transformation T_G_C = poseFromRow(T_G_B) * cameraExtrinsics(*camera).inverse();
  const Vec3 p = {p_G[0], p_G[1], p_G[2]};
  const Vec3 c = T_G_C.inverse().transform(p);
  p_C[0] = c.x;
  p_C[1] = c.y;
  p_C[2] = c.z;
  return AMB_OK;
}

extern "C" uint32_t ambo_pack_color(uint8_t b, uint8_t g, uint8_t r) { return packColor(b, g, r); }
