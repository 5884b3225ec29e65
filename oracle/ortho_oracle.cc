/*
 * ortho_oracle.cc — CPU restatement of ortho::OrthoBackwardGrid::process (TEST INFRASTRUCTURE; see amb_oracle.h).
 *
 * Reference: aerial_mapper_ortho/src/ortho-backward-grid.cc:223-239 (process: pose composition + dispatch),
 * :128-221 (multi-thread cell loop; the single-thread twin :42-126 has the same arithmetic).
 *
 * External arithmetic (sources NOT under /root/reference; restated from upstream knowledge, un-versioned deps
 * install/dependencies_https.rosinstall:1,11):
 *   minkindr  QuatTransformation: operator*, inverse(), transform()  -> struct Transformation below
 *   Eigen     Quaternion product and Quaternion::_transformVector     -> quatMul / quatRotate
 *   aslam_cv2 PinholeCamera::project3 + RadTan / Equidistant distortion -> project3()
 *   grid_map  colorVectorToValue                                       -> ambo_pack_color
 * Cross-checked in tests/test_oracle_ortho.py against cv2.projectPoints, cv2.fisheye.projectPoints and
 * scipy.spatial.transform.Rotation.
 */
#include <atomic>

#include "oracle_common.h"

namespace {

struct Quat {
  double w, x, y, z;
};
struct Vec3 {
  double x, y, z;
};

/* Eigen quaternion product (Eigen/src/Geometry/Quaternion.h, quat_product<..., double>). */
inline Quat quatMul(const Quat& a, const Quat& b) {
  Quat r;
  r.w = a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z;
  r.x = a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y;
  r.y = a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z;
  r.z = a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x;
  return r;
}

inline Vec3 cross(const Vec3& a, const Vec3& b) {
  Vec3 r;
  r.x = a.y * b.z - a.z * b.y;
  r.y = a.z * b.x - a.x * b.z;
  r.z = a.x * b.y - a.y * b.x;
  return r;
}

/* Eigen QuaternionBase::_transformVector: uv = q.vec x v; uv += uv; v + q.w*uv + q.vec x uv. */
inline Vec3 quatRotate(const Quat& q, const Vec3& v) {
  const Vec3 qv = {q.x, q.y, q.z};
  Vec3 uv = cross(qv, v);
  uv.x += uv.x;
  uv.y += uv.y;
  uv.z += uv.z;
  const Vec3 c = cross(qv, uv);
  Vec3 r;
  r.x = (v.x + q.w * uv.x) + c.x;
  r.y = (v.y + q.w * uv.y) + c.y;
  r.z = (v.z + q.w * uv.z) + c.z;
  return r;
}

/* kindr::minimal::QuatTransformation (unit quaternion q_A_B + translation A_t_A_B). */
struct Transformation {
  Quat q;
  Vec3 t;
  /* transform(p) = q.rotate(p) + t */
  Vec3 transform(const Vec3& p) const {
    const Vec3 r = quatRotate(q, p);
    Vec3 o = {r.x + t.x, r.y + t.y, r.z + t.z};
    return o;
  }
  /* inverse() = (q^-1, -(q^-1).rotate(t)); unit quaternion => inverse = conjugate. */
  Transformation inverse() const {
    Transformation o;
    o.q.w = q.w;
    o.q.x = -q.x;
    o.q.y = -q.y;
    o.q.z = -q.z;
    const Vec3 r = quatRotate(o.q, t);
    o.t.x = -r.x;
    o.t.y = -r.y;
    o.t.z = -r.z;
    return o;
  }
  /* A * B = (qA*qB, tA + qA.rotate(tB)) */
  Transformation operator*(const Transformation& rhs) const {
    Transformation o;
    o.q = quatMul(q, rhs.q);
    const Vec3 r = quatRotate(q, rhs.t);
    o.t.x = t.x + r.x;
    o.t.y = t.y + r.y;
    o.t.z = t.z + r.z;
    return o;
  }
};

inline Transformation poseFromRow(const double* r) { /* x y z qw qx qy qz, aerial-mapper-io.cc:110 */
  Transformation T;
  T.t.x = r[0];
  T.t.y = r[1];
  T.t.z = r[2];
  T.q.w = r[3];
  T.q.x = r[4];
  T.q.y = r[5];
  T.q.z = r[6];
  return T;
}

inline Transformation cameraExtrinsics(const amb_camera& cam) {
  Transformation T;
  T.q.w = cam.q_C_B[0];
  T.q.x = cam.q_C_B[1];
  T.q.y = cam.q_C_B[2];
  T.q.z = cam.q_C_B[3];
  T.t.x = cam.t_C_B[0];
  T.t.y = cam.t_C_B[1];
  T.t.z = cam.t_C_B[2];
  return T;
}

enum ProjectionStatus { KEYPOINT_VISIBLE, KEYPOINT_OUTSIDE_IMAGE_BOX, POINT_BEHIND_CAMERA, PROJECTION_INVALID };

/* aslam::PinholeCamera::project3Functional + evaluateProjectionResult (kMinimumDepth = 1e-10). */
inline ProjectionStatus project3(const amb_camera& cam, const Vec3& p, double* kx, double* ky) {
  const double rz = 1.0 / p.z;
  double x = p.x * rz;
  double y = p.y * rz;
  if (cam.dist_type == AMB_DIST_RADTAN) {
    /* aslam::RadTanDistortion::distortUsingExternalCoefficients */
    const double k1 = cam.dist[0], k2 = cam.dist[1], p1 = cam.dist[2], p2 = cam.dist[3];
    const double mx2_u = x * x;
    const double my2_u = y * y;
    const double mxy_u = x * y;
    const double rho2_u = mx2_u + my2_u;
    const double rad_dist_u = k1 * rho2_u + k2 * rho2_u * rho2_u;
    x += x * rad_dist_u + 2.0 * p1 * mxy_u + p2 * (rho2_u + 2.0 * mx2_u);
    y += y * rad_dist_u + 2.0 * p2 * mxy_u + p1 * (rho2_u + 2.0 * my2_u);
  } else if (cam.dist_type == AMB_DIST_EQUIDISTANT) {
    /* aslam::EquidistantDistortion::distortUsingExternalCoefficients */
    const double k1 = cam.dist[0], k2 = cam.dist[1], k3 = cam.dist[2], k4 = cam.dist[3];
    const double x2 = x * x;
    const double y2 = y * y;
    const double r = std::sqrt(x2 + y2);
    if (r > 1e-8) {
      const double theta = std::atan(r);
      const double theta2 = theta * theta;
      const double theta4 = theta2 * theta2;
      const double theta6 = theta4 * theta2;
      const double theta8 = theta4 * theta4;
      const double thetad = theta * (1.0 + k1 * theta2 + k2 * theta4 + k3 * theta6 + k4 * theta8);
      const double scaling = thetad / r;
      x *= scaling;
      y *= scaling;
    }
  }
  *kx = cam.fu * x + cam.cu;
  *ky = cam.fv * y + cam.cv;
  const bool visibility = (*kx >= 0.0) && (*ky >= 0.0) && (*kx < static_cast<double>(cam.width)) &&
                          (*ky < static_cast<double>(cam.height));
  if (visibility && (p.z > 1e-10)) return KEYPOINT_VISIBLE;
  if (!visibility && (p.z > 1e-10)) return KEYPOINT_OUTSIDE_IMAGE_BOX;
  if (p.z < 0.0) return POINT_BEHIND_CAMERA;
  return PROJECTION_INVALID;
}

/* ortho-backward-grid.cc:164-171 */
inline bool keypointVisible(const amb_camera& cam, ProjectionStatus st, double kx, double ky) {
  return (kx >= 0.0) && (ky >= 0.0) && (kx < static_cast<double>(cam.width)) &&
         (ky < static_cast<double>(cam.height)) && (st != POINT_BEHIND_CAMERA) && (st != PROJECTION_INVALID);
}

inline uint32_t packColor(uint8_t b, uint8_t g, uint8_t r) {
  /* ortho-backward-grid.cc:196-201: Eigen::Vector3f(float(rgb[2])/255.0, float(rgb[1])/255.0,
   * float(rgb[0])/255.0) — float / double literal evaluates in double, stored as float.
   * grid_map::colorVectorToValue(Vector3f): Vector3i = (v * 255.0).cast<int>() (float product, truncation), then
   * (t0 << 16) | (t1 << 8) | t2 reinterpreted as float. */
  const float fr = static_cast<float>(static_cast<double>(static_cast<float>(r)) / 255.0);
  const float fg = static_cast<float>(static_cast<double>(static_cast<float>(g)) / 255.0);
  const float fb = static_cast<float>(static_cast<double>(static_cast<float>(b)) / 255.0);
  const int t0 = static_cast<int>(fr * 255.0f);
  const int t1 = static_cast<int>(fg * 255.0f);
  const int t2 = static_cast<int>(fb * 255.0f);
  return (static_cast<uint32_t>(t0) << 16) | (static_cast<uint32_t>(t1) << 8) | static_cast<uint32_t>(t2);
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
