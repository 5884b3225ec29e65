/*
 * amb_refsrc_deps.h — stand-ins for the reference's ABSENT third-party libraries (TEST INFRASTRUCTURE; see
 * ../amb_oracle.h), so that the reference's OWN translation units
 *     aerial_mapper_dsm/src/dsm.cc, aerial_mapper_ortho/src/ortho-backward-grid.cc,
 *     aerial_mapper_ortho/src/ortho-from-pcl.cc, aerial_mapper_utils/src/utils-common.cc
 * (plus the headers they include from /root/reference, nanoflann.hpp among them) compile VERBATIM, from where they
 * lie, into oracle/_ref/ (Makefile target `refsrc`).  Every file next to this one is a one-line forwarder named
 * after the third-party header the reference includes (<Eigen/Dense>, <glog/logging.h>, <ros/ros.h>,
 * <grid_map_core/GridMap.hpp>, <aslam/cameras/camera.h>, <opencv2/highgui/highgui.hpp>, ...).
 *
 * What is the reference's own code in that build: the kd-tree fill, the radius queries and retry loop, the IDW
 * sums, the per-cell / per-frame loops, the visibility predicate, the float32 running maximum, the pixel rounding
 * and clamping, the colour vector construction, utils::parFor and its block partition.
 * What is restated here (no sources under /root/reference; un-versioned deps, install/dependencies_https.rosinstall):
 * only the members those translation units touch —
 *   grid_map   GridMap::getPosition / at / operator[], GridMapIterator, colorVectorToValue   (-> oracle_common.h,
 *   minkindr   QuatTransformation::inverse / transform / operator*                               thirdparty_math.h)
 *   aslam_cv2  Camera::project3, imageWidth/Height, NCamera::getCamera / get_T_C_B
 *   Eigen      fixed-size vectors, Array2i;  OpenCV cv::Mat::at;  glog CHECK/LOG/VLOG;  ros::Time
 * with the same arithmetic the dependency-free oracle uses, so the two differ exactly where the restated LOOPS differ
 * from the reference's.
 *
 * Extensions the glue needs (prefixed amb*): GridMap wraps caller-owned layer buffers and can restrict
 * GridMapIterator to a linear-index sub-range (the bench's bounded sample).
 * The stand-ins also carry the CALLER-side members the reference's demos use to set a job up (GridMap(layers) +
 * setGeometry + setConstant, aslam::Camera/Distortion/NCamera constructors, Pose(w,x,y,z,t), owning cv::Mat), with
 * the same spelling as aerial_mapper_b200/shim/mini/amb_mini_deps.h, so that ONE caller source
 * (tests/cpp/shim_demo.cc = the batch demo's call sequence) builds against the reference's headers + sources here
 * and against the drop-in headers + CUDA library there (tests/test_shim.py).
 * glog FATAL (a failed CHECK) throws ambref::CheckFailed instead of aborting; on a worker thread of utils::parFor
 * that terminates the process, exactly like the reference's abort.
 */
#ifndef AMB_REFSRC_DEPS_H_
#define AMB_REFSRC_DEPS_H_

#include <chrono>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <vector>

#include "../oracle_common.h"
#include "../thirdparty_math.h"

/* ------------------------------------------------------------------ glog ------------------------------------ */
namespace ambref {

struct CheckFailed : public std::runtime_error {
  explicit CheckFailed(const std::string& what) : std::runtime_error(what) {}
};

enum Severity { SEV_INFO = 0, SEV_WARNING = 1, SEV_ERROR = 2, SEV_FATAL = 3, SEV_VLOG = 4 };

inline bool verbose() {
  static const bool v = std::getenv("AMB_REFSRC_VERBOSE") != nullptr;
  return v;
}

class LogLine {
 public:
  LogLine(Severity severity, const char* file, int line) : severity_(severity) {
    stream_ << file << ":" << line << "] ";
  }
  ~LogLine() noexcept(false) {
    if (severity_ == SEV_FATAL) throw CheckFailed(stream_.str());
    if (verbose() || severity_ == SEV_ERROR) std::cerr << stream_.str() << std::endl;
  }
  std::ostream& stream() { return stream_; }

 private:
  Severity severity_;
  std::ostringstream stream_;
};

struct Voidify {
  void operator&(std::ostream&) {}
};

}  // namespace ambref

#define LOG(severity) ::ambref::LogLine(::ambref::SEV_##severity, __FILE__, __LINE__).stream()
#define VLOG(level) ::ambref::LogLine(::ambref::SEV_VLOG, __FILE__, __LINE__).stream()
#define CHECK(condition)        \
  (condition) ? (void)0         \
              : ::ambref::Voidify() & ::ambref::LogLine(::ambref::SEV_FATAL, __FILE__, __LINE__).stream() \
                                          << "Check failed: " #condition " "
#define CHECK_GT(a, b) CHECK((a) > (b))
#define CHECK_GE(a, b) CHECK((a) >= (b))
#define CHECK_LT(a, b) CHECK((a) < (b))
#define CHECK_LE(a, b) CHECK((a) <= (b))
#define CHECK_EQ(a, b) CHECK((a) == (b))
#define CHECK_NE(a, b) CHECK((a) != (b))
#define CHECK_NOTNULL(p) (p)

/* ------------------------------------------------------------------ ros ------------------------------------- */
namespace ros {
struct Duration {
  double seconds;
};
inline std::ostream& operator<<(std::ostream& o, const Duration& d) { return o << d.seconds; }
struct Time {
  double seconds;
  static Time now() {
    Time t;
    t.seconds = ambo::now();
    return t;
  }
  Duration operator-(const Time& o) const {
    Duration d;
    d.seconds = seconds - o.seconds;
    return d;
  }
};
}  // namespace ros

/* ------------------------------------------------------------------ Eigen ----------------------------------- */
#define EIGEN_MAKE_ALIGNED_OPERATOR_NEW
namespace Eigen {
typedef std::ptrdiff_t Index;

template <typename T, int N>
struct AmbVec {
  T v[N];
  AmbVec() {}
  AmbVec(T a, T b) {
    v[0] = a;
    v[1] = b;
  }
  AmbVec(T a, T b, T c) {
    v[0] = a;
    v[1] = b;
    v[2] = c;
  }
  T& operator()(Index i) { return v[i]; }
  const T& operator()(Index i) const { return v[i]; }
  T& operator[](Index i) { return v[i]; }
  const T& operator[](Index i) const { return v[i]; }
  const T& x() const { return v[0]; }
  const T& y() const { return v[1]; }
  const T& z() const { return v[2]; }
  T& x() { return v[0]; }
  T& y() { return v[1]; }
  T& z() { return v[2]; }
};
typedef AmbVec<double, 2> Vector2d;
typedef AmbVec<double, 3> Vector3d;
typedef AmbVec<double, 4> Vector4d;
typedef AmbVec<float, 3> Vector3f;
typedef AmbVec<int, 2> Array2i;

template <typename T>
using aligned_allocator = std::allocator<T>;
}  // namespace Eigen

/* ------------------------------------------------------------------ grid_map -------------------------------- */
namespace grid_map {
typedef Eigen::Array2i Index;
typedef Eigen::Array2i Size;
typedef Eigen::Vector2d Position;
typedef Eigen::Vector2d Length;

/* grid_map::Matrix = Eigen::MatrixXf: column-major float32; a view of the caller's buffer (glue) or owning (demo). */
class Matrix {
 public:
  Matrix() : data_(nullptr), rows_(0), cols_(0) {}
  Matrix(float* data, Eigen::Index rows, Eigen::Index cols) : data_(data), rows_(rows), cols_(cols) {}
  void resize(Eigen::Index rows, Eigen::Index cols) {
    own_.reset(new std::vector<float>(static_cast<size_t>(rows) * static_cast<size_t>(cols), 0.0f));
    data_ = own_->data();
    rows_ = rows;
    cols_ = cols;
  }
  void setConstant(float value) {
    for (Eigen::Index k = 0; k < rows_ * cols_; ++k) data_[k] = value;
  }
  float& operator()(Eigen::Index i, Eigen::Index j) { return data_[j * rows_ + i]; }
  const float& operator()(Eigen::Index i, Eigen::Index j) const { return data_[j * rows_ + i]; }
  float* data() { return data_; }
  const float* data() const { return data_; }
  Eigen::Index rows() const { return rows_; }
  Eigen::Index cols() const { return cols_; }

 private:
  float* data_;
  Eigen::Index rows_, cols_;
  std::shared_ptr<std::vector<float> > own_;
};

class GridMap {
 public:
  explicit GridMap(const amb_geometry& g) : geometry_(g), iter_begin_(0) {
    size_ = Size(g.rows, g.cols);
    iter_end_ = static_cast<int64_t>(g.rows) * g.cols;
  }
  /* caller side, as the demos use it (aerial-mapper-grid-map.cc:23-33) */
  explicit GridMap(const std::vector<std::string>& layers) : iter_begin_(0), iter_end_(0) {
    std::memset(&geometry_, 0, sizeof(geometry_));
    size_ = Size(0, 0);
    for (size_t k = 0; k < layers.size(); ++k) layers_[layers[k]];
  }
  void setFrameId(const std::string& frame_id) { frame_id_ = frame_id; }
  const std::string& getFrameId() const { return frame_id_; }
  /* grid_map::GridMap::setGeometry: size = round(length / resolution), length = size * resolution, every layer
   * resized and set to NaN (clearAll), startIndex = 0. */
  void setGeometry(const Length& length, double resolution, const Position& position) {
    geometry_.rows = static_cast<int32_t>(std::round(length(0) / resolution));
    geometry_.cols = static_cast<int32_t>(std::round(length(1) / resolution));
    geometry_.resolution = resolution;
    geometry_.length_x = geometry_.rows * resolution;
    geometry_.length_y = geometry_.cols * resolution;
    geometry_.pos_x = position(0);
    geometry_.pos_y = position(1);
    size_ = Size(geometry_.rows, geometry_.cols);
    iter_begin_ = 0;
    iter_end_ = static_cast<int64_t>(geometry_.rows) * geometry_.cols;
    for (std::unordered_map<std::string, Matrix>::iterator it = layers_.begin(); it != layers_.end(); ++it) {
      it->second.resize(geometry_.rows, geometry_.cols);
      it->second.setConstant(std::nanf(""));
    }
  }
  Length getLength() const { return Length(geometry_.length_x, geometry_.length_y); }
  Position getPosition() const { return Position(geometry_.pos_x, geometry_.pos_y); }
  /* amb* = glue-only extensions */
  void ambAddLayer(const std::string& name, float* data) {
    layers_[name] = Matrix(data, geometry_.rows, geometry_.cols);
  }
  void ambSetIterationRange(int64_t begin, int64_t end) {
    iter_begin_ = begin;
    iter_end_ = end;
  }
  int64_t ambIterationBegin() const { return iter_begin_; }
  int64_t ambIterationEnd() const { return iter_end_; }

  /* grid_map::GridMap::getPosition(index, position) -> getPositionFromIndex (oracle_common.h cellPosition). */
  bool getPosition(const Index& index, Position& position) const {
    if (index(0) < 0 || index(1) < 0 || index(0) >= size_(0) || index(1) >= size_(1)) return false;
    ambo::cellPosition(geometry_, index(0), index(1), &position.x(), &position.y());
    return true;
  }
  const Size& getSize() const { return size_; }
  double getResolution() const { return geometry_.resolution; }
  Matrix& operator[](const std::string& layer) { return find(layer); }
  const Matrix& operator[](const std::string& layer) const { return const_cast<GridMap*>(this)->find(layer); }
  float& at(const std::string& layer, const Index& index) { return find(layer)(index(0), index(1)); }

 private:
  Matrix& find(const std::string& layer) {
    std::unordered_map<std::string, Matrix>::iterator it = layers_.find(layer);
    if (it == layers_.end()) throw std::out_of_range("GridMap: layer '" + layer + "' does not exist");
    return it->second;
  }
  amb_geometry geometry_;
  Size size_;
  std::unordered_map<std::string, Matrix> layers_;
  int64_t iter_begin_, iter_end_;
  std::string frame_id_;
};

/* grid_map::GridMapIterator: linear index over the column-major buffer, index = (k % rows, k / rows)
 * (getIndexFromLinearIndex; startIndex 0). */
class GridMapIterator {
 public:
  explicit GridMapIterator(const GridMap& map)
      : rows_(map.getSize()(0)), linear_(map.ambIterationBegin()), end_(map.ambIterationEnd()) {}
  const Index operator*() const {
    return Index(static_cast<int>(linear_ % rows_), static_cast<int>(linear_ / rows_));
  }
  GridMapIterator& operator++() {
    ++linear_;
    return *this;
  }
  bool isPastEnd() const { return linear_ >= end_; }

 private:
  int64_t rows_, linear_, end_;
};

/* grid_map::colorVectorToValue(const Eigen::Vector3f&, float&) -> thirdparty_math.h colorVectorToBits. */
inline void colorVectorToValue(const Eigen::Vector3f& color_vector, float& color_value) {
  const uint32_t bits = ambo::tp::colorVectorToBits(color_vector(0), color_vector(1), color_vector(2));
  std::memcpy(&color_value, &bits, sizeof(float));
}
}  // namespace grid_map

/* ------------------------------------------------------------------ minkindr -------------------------------- */
namespace kindr {
namespace minimal {
class QuatTransformation {
 public:
  QuatTransformation() {
    T_.q.w = 1.0;
    T_.q.x = T_.q.y = T_.q.z = 0.0;
    T_.t.x = T_.t.y = T_.t.z = 0.0;
  }
  explicit QuatTransformation(const ambo::tp::Transformation& T) : T_(T) {}
  /* caller side: unit quaternion (w, x, y, z) + translation, the order of aerial-mapper-io.cc:110-117 */
  QuatTransformation(double w, double x, double y, double z, double tx, double ty, double tz) {
    T_.q.w = w;
    T_.q.x = x;
    T_.q.y = y;
    T_.q.z = z;
    T_.t.x = tx;
    T_.t.y = ty;
    T_.t.z = tz;
  }
  struct RotationQuaternion {
    ambo::tp::Quat q;
    double w() const { return q.w; }
    double x() const { return q.x; }
    double y() const { return q.y; }
    double z() const { return q.z; }
  };
  RotationQuaternion getRotation() const {
    RotationQuaternion r;
    r.q = T_.q;
    return r;
  }
  Eigen::Vector3d getPosition() const { return Eigen::Vector3d(T_.t.x, T_.t.y, T_.t.z); }
  const ambo::tp::Transformation& ambTransformation() const { return T_; }
  QuatTransformation inverse() const { return QuatTransformation(T_.inverse()); }
  Eigen::Vector3d transform(const Eigen::Vector3d& p) const {
    const ambo::tp::Vec3 in = {p(0), p(1), p(2)};
    const ambo::tp::Vec3 out = T_.transform(in);
    return Eigen::Vector3d(out.x, out.y, out.z);
  }
  QuatTransformation operator*(const QuatTransformation& rhs) const { return QuatTransformation(T_ * rhs.T_); }

 private:
  ambo::tp::Transformation T_;
};
}  // namespace minimal
}  // namespace kindr

/* ------------------------------------------------------------------ aslam_cv2 ------------------------------- */
namespace aslam {
typedef kindr::minimal::QuatTransformation Transformation;

struct ProjectionResult {
  enum Status { KEYPOINT_VISIBLE, KEYPOINT_OUTSIDE_IMAGE_BOX, POINT_BEHIND_CAMERA, PROJECTION_INVALID, UNINITIALIZED };
  ProjectionResult() : status_(UNINITIALIZED) {}
  explicit ProjectionResult(Status s) : status_(s) {}
  Status getDetailedStatus() const { return status_; }

 private:
  Status status_;
};

class Distortion {
 public:
  enum class Type { kNoDistortion = 0, kEquidistant = 1, kFisheye = 2, kRadTan = 3 };
  Distortion(Type type, const Eigen::Vector4d& parameters) : type_(type), parameters_(parameters) {}
  Type getType() const { return type_; }
  const Eigen::Vector4d& getParameters() const { return parameters_; }

 private:
  Type type_;
  Eigen::Vector4d parameters_;
};

/* aslam::PinholeCamera with its distortion, as one concrete class (the reference only calls through Camera&). */
class Camera {
 public:
  explicit Camera(const amb_camera& c) : c_(c), intrinsics_(), distortion_(Distortion::Type::kNoDistortion, Eigen::Vector4d()) {
    for (int k = 0; k < 4; ++k) distortion_params_(k) = c.dist[k];
    intrinsics_(0) = c.fu;
    intrinsics_(1) = c.fv;
    intrinsics_(2) = c.cu;
    intrinsics_(3) = c.cv;
    const Distortion::Type t = c.dist_type == AMB_DIST_RADTAN        ? Distortion::Type::kRadTan
                               : c.dist_type == AMB_DIST_EQUIDISTANT ? Distortion::Type::kEquidistant
                               : c.dist_type == AMB_DIST_FOV         ? Distortion::Type::kFisheye
                                                                     : Distortion::Type::kNoDistortion;
    distortion_ = Distortion(t, distortion_params_);
  }
  /* caller side: PinholeCamera(intrinsics fu fv cu cv, width, height, distortion) */
  Camera(unsigned width, unsigned height, const Eigen::Vector4d& intrinsics, const Distortion& distortion)
      : intrinsics_(intrinsics), distortion_(distortion) {
    std::memset(&c_, 0, sizeof(c_));
    c_.width = static_cast<int32_t>(width);
    c_.height = static_cast<int32_t>(height);
    c_.fu = intrinsics(0);
    c_.fv = intrinsics(1);
    c_.cu = intrinsics(2);
    c_.cv = intrinsics(3);
    c_.dist_type = distortion.getType() == Distortion::Type::kRadTan        ? AMB_DIST_RADTAN
                   : distortion.getType() == Distortion::Type::kEquidistant ? AMB_DIST_EQUIDISTANT
                   : distortion.getType() == Distortion::Type::kFisheye     ? AMB_DIST_FOV
                                                                            : AMB_DIST_NONE;
    for (int k = 0; k < 4; ++k) c_.dist[k] = distortion.getParameters()(k);
    c_.q_C_B[0] = 1.0;
  }
  const Eigen::Vector4d& getParameters() const { return intrinsics_; }
  const Distortion& getDistortion() const { return distortion_; }
  const ProjectionResult project3(const Eigen::Vector3d& point_3d, Eigen::Vector2d* out_keypoint) const {
    const ambo::tp::Vec3 p = {point_3d(0), point_3d(1), point_3d(2)};
    double kx, ky;
    const ambo::tp::ProjectionStatus st = ambo::tp::project3(c_, p, &kx, &ky);
    (*out_keypoint)(0) = kx;
    (*out_keypoint)(1) = ky;
    switch (st) {
      case ambo::tp::KEYPOINT_VISIBLE:
        return ProjectionResult(ProjectionResult::KEYPOINT_VISIBLE);
      case ambo::tp::KEYPOINT_OUTSIDE_IMAGE_BOX:
        return ProjectionResult(ProjectionResult::KEYPOINT_OUTSIDE_IMAGE_BOX);
      case ambo::tp::POINT_BEHIND_CAMERA:
        return ProjectionResult(ProjectionResult::POINT_BEHIND_CAMERA);
      default:
        return ProjectionResult(ProjectionResult::PROJECTION_INVALID);
    }
  }
  uint32_t imageWidth() const { return static_cast<uint32_t>(c_.width); }
  uint32_t imageHeight() const { return static_cast<uint32_t>(c_.height); }

 private:
  amb_camera c_;
  Eigen::Vector4d intrinsics_, distortion_params_;
  Distortion distortion_;
};
typedef Camera PinholeCamera;

class NCamera {
 public:
  typedef std::shared_ptr<NCamera> Ptr;
  explicit NCamera(const amb_camera& c) : camera_(c), T_C_B_(ambo::tp::cameraExtrinsics(c)) {}
  NCamera(const Camera& camera, const Transformation& T_C_B) : camera_(camera), T_C_B_(T_C_B) {}  /* caller side */
  const Camera& getCamera(size_t /*camera_index*/) const { return camera_; }
  const Transformation& get_T_C_B(size_t /*camera_index*/) const { return T_C_B_; }

 private:
  Camera camera_;
  Transformation T_C_B_;
};
}  // namespace aslam

/* ------------------------------------------------------------------ OpenCV ---------------------------------- */
typedef unsigned char uchar;
namespace cv {
struct Size {
  int width, height;
  Size() : width(0), height(0) {}
  Size(int w, int h) : width(w), height(h) {}
  bool operator==(const Size& o) const { return width == o.width && height == o.height; }
  bool operator!=(const Size& o) const { return !(*this == o); }
};
struct Vec3b {
  uchar val[3];
  const uchar& operator[](int i) const { return val[i]; }
  uchar& operator[](int i) { return val[i]; }
};
/* cv::Mat: interleaved uint8 image; a view of the caller's buffer (glue) or owning (demo). */
class Mat {
 public:
  Mat() : rows(0), cols(0), data(nullptr), step(0), channels_(1) {}
  Mat(int rows_, int cols_, const uchar* data_, size_t step_)
      : rows(rows_), cols(cols_), data(const_cast<uchar*>(data_)), step(step_),
        channels_(cols_ > 0 ? static_cast<int>(step_ / static_cast<size_t>(cols_)) : 1) {}
  Mat(int rows_, int cols_, int channels)
      : rows(rows_), cols(cols_), data(nullptr), step(static_cast<size_t>(cols_) * channels), channels_(channels) {
    own_.reset(new std::vector<uchar>(static_cast<size_t>(rows_) * step));
    data = own_->data();
  }
  template <typename T>
  const T& at(int row, int col) const {
    return *reinterpret_cast<const T*>(data + static_cast<size_t>(row) * step + static_cast<size_t>(col) * sizeof(T));
  }
  template <typename T>
  T* ptr(int row) {
    return reinterpret_cast<T*>(data + static_cast<size_t>(row) * step);
  }
  template <typename T>
  const T* ptr(int row) const {
    return reinterpret_cast<const T*>(data + static_cast<size_t>(row) * step);
  }
  Size size() const { return Size(cols, rows); }
  bool empty() const { return data == nullptr; }
  int channels() const { return channels_; }
  int rows, cols;
  uchar* data;
  size_t step;

 private:
  int channels_;
  std::shared_ptr<std::vector<uchar> > own_;
};
}  // namespace cv

#endif /* AMB_REFSRC_DEPS_H_ */
