/*
 * thirdparty_math.h — the arithmetic of the reference's ABSENT third-party dependencies, restated once
 * (TEST INFRASTRUCTURE; see amb_oracle.h).  Shared by
 *   - ortho_oracle.cc            : the dependency-free restatement of the ortho cell loop, and
 *   - refsrc_stubs/amb_refsrc_deps.h : the stand-in headers the reference's OWN translation units (dsm.cc,
 *                                  ortho-backward-grid.cc, ortho-from-pcl.cc) are compiled against in oracle/_ref.
 *
 * Sources NOT under /root/reference (un-versioned catkin dependencies, install/dependencies_https.rosinstall:1,9,11);
 * restated from their published upstream sources:
 *   minkindr   kindr::minimal::QuatTransformation: operator*, inverse(), transform()
 *   Eigen      Quaternion product (quat_product<double>) and QuaternionBase::_transformVector
 *   aslam_cv2  PinholeCamera::project3Functional + evaluateProjectionResult, RadTan / Equidistant distortion
 *   grid_map   colorVectorToValue(Vector3f)
 * Cross-checked in tests/test_oracle_ortho.py against cv2.projectPoints, cv2.fisheye.projectPoints and
 * scipy.spatial.transform.Rotation.  Build with -ffp-contract=off.
 * NOT cross-checked: the FOV ("fisheye") distortion branch below.  OpenCV has no such model and the aslam_cv2 source is not
 * available here; it is written from recollection of upstream distortion-fisheye.cc (thresholds 1e-5, limit 2 tan(w/2)/w).
 * A GPU-vs-oracle match on that branch shows the CUDA path equals THIS restatement, nothing more.
 */
#ifndef AMB_ORACLE_THIRDPARTY_MATH_H_
#define AMB_ORACLE_THIRDPARTY_MATH_H_

#include <cmath>
#include <cstdint>

#include "../include/aerial_mapper_b200.h" /* amb_camera, AMB_DIST_* */

namespace ambo {
namespace tp {

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

/* aslam::ProjectionResult::Status (aslam/cameras/camera.h). */
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
  } else if (cam.dist_type == AMB_DIST_FOV) {
    /* aslam::FisheyeDistortion::distortUsingExternalCoefficients — FROM RECOLLECTION, unverified (see header) */
    const double w = cam.dist[0];
    const double r_u = std::sqrt(x * x + y * y);
    const double tanwhalf = std::tan(w / 2.);
    const double atan_wrd = std::atan(2. * tanwhalf * r_u);
    double r_rd;
    if (w * w < 1e-5) {
      r_rd = 1.0;
    } else if (r_u * r_u < 1e-5) {
      r_rd = 2. * tanwhalf / w;
    } else {
      r_rd = atan_wrd / (r_u * w);
    }
    x *= r_rd;
    y *= r_rd;
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

/* grid_map::colorVectorToValue(const Eigen::Vector3f&, float&) (grid_map_core GridMapMath.cpp):
 * Vector3i = (v * 255.0).cast<int>() — a float product (Eigen converts the literal to the vector's scalar), then
 * truncation; (t0 << 16) + (t1 << 8) + t2 reinterpreted as float.  Returns the bit pattern. */
inline uint32_t colorVectorToBits(float c0, float c1, float c2) {
  const int t0 = static_cast<int>(c0 * 255.0f);
  const int t1 = static_cast<int>(c1 * 255.0f);
  const int t2 = static_cast<int>(c2 * 255.0f);
  return (static_cast<uint32_t>(t0) << 16) | (static_cast<uint32_t>(t1) << 8) | static_cast<uint32_t>(t2);
}

}  // namespace tp
}  // namespace ambo
#endif
