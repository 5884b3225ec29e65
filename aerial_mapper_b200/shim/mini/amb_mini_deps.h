// amb_mini_deps.h — stand-ins for Eigen / grid_map / aslam_cv2 / minkindr / OpenCV / glog with just the members
// the shim's marshalling code touches, so the drop-in headers can be compiled and tested in this repository where
// none of those libraries exist.  NOT used in a real aerial_mapper workspace (build without -DAMB_SHIM_MINI).
#ifndef AMB_MINI_DEPS_H_
#define AMB_MINI_DEPS_H_

#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <iostream>
#include <limits>
#include <map>
#include <memory>
#include <sstream>
#include <string>
#include <vector>

// ---- glog ----
struct AmbMiniLog {
  bool fatal;
  std::ostringstream s;
  explicit AmbMiniLog(bool f) : fatal(f) {}
  ~AmbMiniLog() {
    std::cerr << s.str() << std::endl;
    if (fatal) std::abort();
  }
  template <typename T>
  AmbMiniLog& operator<<(const T& v) {
    s << v;
    return *this;
  }
};
#define WARNING false
#define FATAL true
#define LOG(severity) AmbMiniLog(severity)
#define CHECK(cond) \
  if (!(cond)) AmbMiniLog(true) << "Check failed: " #cond " "
#define EIGEN_MAKE_ALIGNED_OPERATOR_NEW

// ---- Eigen ----
namespace Eigen {
template <int N>
struct VecD {
  double v[N];
  double& operator()(int i) { return v[i]; }
  const double& operator()(int i) const { return v[i]; }
  double x() const { return v[0]; }
  double y() const { return v[1]; }
};
typedef VecD<3> Vector3d;
typedef VecD<2> Vector2d;
typedef VecD<4> Vector4d;
struct Array2i {
  int v[2];
  int operator()(int i) const { return v[i]; }
};
template <typename T>
using aligned_allocator = std::allocator<T>;
// column-major float matrix (Eigen::MatrixXf)
class MatrixXf {
 public:
  MatrixXf() : rows_(0), cols_(0) {}
  void resize(int r, int c) {
    rows_ = r;
    cols_ = c;
    d_.assign(static_cast<size_t>(r) * c, 0.f);
  }
  void setConstant(float v) { d_.assign(d_.size(), v); }
  float& operator()(int i, int j) { return d_[static_cast<size_t>(j) * rows_ + i]; }
  float operator()(int i, int j) const { return d_[static_cast<size_t>(j) * rows_ + i]; }
  float* data() { return d_.data(); }
  const float* data() const { return d_.data(); }
  int rows() const { return rows_; }
  int cols() const { return cols_; }

 private:
  int rows_, cols_;
  std::vector<float> d_;
};
}  // namespace Eigen

template <template <typename, typename> class Container, typename Type>
struct AlignedType {  // utils-nearest-neighbor.h:18-21
  typedef Container<Type, Eigen::aligned_allocator<Type> > type;
};

// ---- grid_map ----
namespace grid_map {
typedef Eigen::MatrixXf Matrix;
typedef Eigen::Vector2d Position;
typedef Eigen::Vector2d Length;
typedef Eigen::Array2i Size;
class GridMap {
 public:
  explicit GridMap(const std::vector<std::string>& layers) : resolution_(0) {
    for (const auto& l : layers) data_[l];
    size_.v[0] = size_.v[1] = 0;
  }
  void setFrameId(const std::string& f) { frame_ = f; }
  // grid_map::GridMap::setGeometry: size = round(length / resolution); length = size * resolution; clearAll().
  void setGeometry(const Length& length, double resolution, const Position& position) {
    size_.v[0] = static_cast<int>(std::round(length(0) / resolution));
    size_.v[1] = static_cast<int>(std::round(length(1) / resolution));
    resolution_ = resolution;
    length_.v[0] = size_.v[0] * resolution;
    length_.v[1] = size_.v[1] * resolution;
    position_ = position;
    for (auto& kv : data_) {
      kv.second.resize(size_.v[0], size_.v[1]);
      kv.second.setConstant(std::numeric_limits<float>::quiet_NaN());
    }
  }
  const Size& getSize() const { return size_; }
  const Length& getLength() const { return length_; }
  double getResolution() const { return resolution_; }
  const Position& getPosition() const { return position_; }
  Matrix& operator[](const std::string& layer) { return data_.at(layer); }
  const Matrix& operator[](const std::string& layer) const { return data_.at(layer); }

 private:
  std::map<std::string, Matrix> data_;
  Size size_;
  Length length_;
  Position position_;
  double resolution_;
  std::string frame_;
};
}  // namespace grid_map

// ---- minkindr ----
namespace kindr {
namespace minimal {
struct RotationQuaternion {
  double q[4];  // w x y z
  double w() const { return q[0]; }
  double x() const { return q[1]; }
  double y() const { return q[2]; }
  double z() const { return q[3]; }
};
class QuatTransformation {
 public:
  QuatTransformation() {
    r_.q[0] = 1;
    r_.q[1] = r_.q[2] = r_.q[3] = 0;
    t_.v[0] = t_.v[1] = t_.v[2] = 0;
  }
  QuatTransformation(double w, double x, double y, double z, double tx, double ty, double tz) {
    r_.q[0] = w; r_.q[1] = x; r_.q[2] = y; r_.q[3] = z;
    t_.v[0] = tx; t_.v[1] = ty; t_.v[2] = tz;
  }
  const RotationQuaternion& getRotation() const { return r_; }
  const Eigen::Vector3d& getPosition() const { return t_; }

 private:
  RotationQuaternion r_;
  Eigen::Vector3d t_;
};
}  // namespace minimal
}  // namespace kindr

// ---- aslam_cv2 ----
namespace aslam {
class Distortion {
 public:
  enum class Type { kNoDistortion = 0, kEquidistant = 1, kFisheye = 2, kRadTan = 3 };
  Distortion(Type t, const Eigen::Vector4d& p) : type_(t), p_(p) {}
  Type getType() const { return type_; }
  const Eigen::Vector4d& getParameters() const { return p_; }

 private:
  Type type_;
  Eigen::Vector4d p_;
};
class Camera {
 public:
  Camera(unsigned w, unsigned h, const Eigen::Vector4d& intrinsics, const Distortion& d)
      : w_(w), h_(h), k_(intrinsics), d_(d) {}
  unsigned imageWidth() const { return w_; }
  unsigned imageHeight() const { return h_; }
  const Eigen::Vector4d& getParameters() const { return k_; }
  const Distortion& getDistortion() const { return d_; }

 private:
  unsigned w_, h_;
  Eigen::Vector4d k_;
  Distortion d_;
};
class NCamera {
 public:
  NCamera(const Camera& c, const kindr::minimal::QuatTransformation& T_C_B) : c_(c), T_C_B_(T_C_B) {}
  const Camera& getCamera(size_t) const { return c_; }
  const kindr::minimal::QuatTransformation& get_T_C_B(size_t) const { return T_C_B_; }

 private:
  Camera c_;
  kindr::minimal::QuatTransformation T_C_B_;
};
}  // namespace aslam

// ---- OpenCV ----
namespace cv {
struct Mat {
  int rows, cols;
  int channels_;
  size_t step;
  unsigned char* data;
  std::shared_ptr<std::vector<unsigned char> > storage;
  Mat() : rows(0), cols(0), channels_(1), step(0), data(nullptr) {}
  Mat(int r, int c, int ch) : rows(r), cols(c), channels_(ch), step(static_cast<size_t>(c) * ch) {
    storage.reset(new std::vector<unsigned char>(static_cast<size_t>(r) * step));
    data = storage->data();
  }
  int channels() const { return channels_; }
};
}  // namespace cv

typedef kindr::minimal::QuatTransformation Pose;  // aerial-mapper-io.h:17-20
typedef std::vector<Pose> Poses;
typedef cv::Mat Image;
typedef std::vector<Image> Images;

#endif
