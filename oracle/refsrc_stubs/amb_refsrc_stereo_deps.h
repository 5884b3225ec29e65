/*
 * amb_refsrc_stereo_deps.h — extra stand-ins (TEST INFRASTRUCTURE; see amb_refsrc_deps.h) for compiling the
 * reference's aerial_mapper_dense_pcl/src/densifier.cpp VERBATIM into oracle/_ref/libamb_refsrc_stereo.so:
 * the members of Eigen (fixed 3x3 / 4x4 matrices with the comma initialiser), OpenCV (cv::Size, cv::Mat::ptr/size,
 * cv::Mat_<Vec3f>, StereoBM / StereoSGBM parameter holders) and ROS (sensor_msgs::PointCloud2) that translation unit
 * and the headers it includes touch.
 *
 * Reference code in that build: Densifier::computePointCloud — the Q matrix, the per-pixel loop, the validity rule
 * (disparity > kMaxInvalidDisparity, !isinf((float)z)), the raster order of point_cloud_eigen / intensities.
 * Restated here: Eigen's Matrix3d * Vector3d product ((m0*x + m1*y) + m2*z per row) and vector sum.
 * Block matching itself (cv::StereoBM / StereoSGBM ::compute) is OpenCV and out of scope: the holders only keep
 * parameters, and BlockMatching*::computeDisparityMap is not part of this build (the glue defines it to throw).
 */
#ifndef AMB_REFSRC_STEREO_DEPS_H_
#define AMB_REFSRC_STEREO_DEPS_H_

#include <amb_refsrc_deps.h>

/* ------------------------------------------------------------------ Eigen: small fixed matrices ------------- */
namespace Eigen {

template <typename T, int R, int C>
struct AmbMat;

template <typename T, int R, int C>
struct AmbCommaInit {
  AmbMat<T, R, C>* m;
  int next;
  AmbCommaInit& operator,(T v) {
    m->a[next / C][next % C] = v; /* row-major fill order, like Eigen's CommaInitializer */
    ++next;
    return *this;
  }
  AmbMat<T, R, C>& finished() { return *m; }
};

template <typename T, int R, int C>
struct AmbMat {
  T a[R][C];
  AmbMat() {}
  T& operator()(Index i, Index j) { return a[i][j]; }
  const T& operator()(Index i, Index j) const { return a[i][j]; }
  AmbCommaInit<T, R, C> operator<<(T v) {
    a[0][0] = v;
    AmbCommaInit<T, R, C> c;
    c.m = this;
    c.next = 1;
    return c;
  }
};
typedef AmbMat<double, 3, 3> Matrix3d;
typedef AmbMat<double, 4, 4> Matrix4d;

/* Matrix3d * Vector3d, Vector3d + Vector3d (the only products densifier.cpp forms, :73-74) */
inline Vector3d operator*(const Matrix3d& m, const Vector3d& v) {
  return Vector3d((m(0, 0) * v(0) + m(0, 1) * v(1)) + m(0, 2) * v(2), (m(1, 0) * v(0) + m(1, 1) * v(1)) + m(1, 2) * v(2),
                  (m(2, 0) * v(0) + m(2, 1) * v(1)) + m(2, 2) * v(2));
}
inline Vector3d operator+(const Vector3d& a, const Vector3d& b) { return Vector3d(a(0) + b(0), a(1) + b(1), a(2) + b(2)); }

}  // namespace Eigen

/* ------------------------------------------------------------------ OpenCV extras --------------------------- */
#define CV_8U 0
#define CV_32F 5
#define CV_8UC1 0
#define CV_32FC1 5
namespace cv {

struct Vec3f {
  float val[3];
  Vec3f() {}
  Vec3f(float a, float b, float c) {
    val[0] = a;
    val[1] = b;
    val[2] = c;
  }
};

template <typename T>
class Mat_ {
 public:
  void create(const Size& s) { storage_.resize(static_cast<size_t>(s.width) * s.height); }
  void setTo(const T& v) {
    for (size_t k = 0; k < storage_.size(); ++k) storage_[k] = v;
  }

 private:
  std::vector<T> storage_;
};

template <typename T>
using Ptr = std::shared_ptr<T>;

struct StereoBM {
  static Ptr<StereoBM> create(int, int) { return Ptr<StereoBM>(new StereoBM()); }
  void setMinDisparity(int) {}
  void setNumDisparities(int) {}
  void setPreFilterCap(int) {}
  void setPreFilterSize(int) {}
  void setUniquenessRatio(int) {}
  void setTextureThreshold(int) {}
  void setSpeckleWindowSize(int) {}
  void setSpeckleRange(int) {}
  void setDisp12MaxDiff(int) {}
  void setBlockSize(int) {}
};
struct StereoSGBM {
  static Ptr<StereoSGBM> create(int, int, int) { return Ptr<StereoSGBM>(new StereoSGBM()); }
  void setMinDisparity(int) {}
  void setNumDisparities(int) {}
  void setPreFilterCap(int) {}
  void setUniquenessRatio(int) {}
  void setSpeckleWindowSize(int) {}
  void setSpeckleRange(int) {}
  void setDisp12MaxDiff(int) {}
  void setP1(int) {}
  void setP2(int) {}
  void setBlockSize(int) {}
};

}  // namespace cv

/* ------------------------------------------------------------------ ROS messages ---------------------------- */
namespace sensor_msgs {
struct PointCloud2 {
  struct Header {
    ros::Time stamp;
  } header;
  uint32_t height, width, point_step, row_step;
  std::vector<uint8_t> data;
  PointCloud2() : height(0), width(0), point_step(0), row_step(0) {}
};
}  // namespace sensor_msgs

#endif /* AMB_REFSRC_STEREO_DEPS_H_ */
