// rectify_kernels.cu — "next" row N3, second half: the rectification-map fill of
// stereo::Rectifier::rectifyStereoPair (reference aerial_mapper_dense_pcl/src/rectifier.cpp:36-107), the step that
// precedes block matching and Densifier::computePointCloud in the incremental pipeline (stereo.cpp:149-193).
//
//   host, once per stereo pair (rectifier.cpp:43-79): Fusiello's compact rectification in double —
//     x = t_G_C2 - t_G_C1, y = R_G_C1.col(2) x x, z = x x y, R_rect = [x^ y^ z^]^T,
//     T_i = (K R_rect) (K R_G_Ci^T)^-1, T_i_inv = T_i^-1 cast to float            -> amb_stereo_rectify_setup
//   device, per rectified pixel (rectifier.cpp:80-104): [x y w]^T = T_i_inv [u v 1]^T in float32,
//     map_i = (x / w, y / w) for both cameras                                        -> rectify_maps_kernel
//
// The kernel reads 72 bytes of constants and writes 16 B per pixel (four CV_32FC1 maps): a pure HBM-write kernel,
// one thread per 4 consecutive pixels of a row, 128-bit stores, grid-stride over a grid sized to the SM count.
// Float arithmetic is un-contracted (__fmul_rn / __fadd_rn / __fdiv_rn) in the order (m0*u + m1*v) + m2, so the maps
// are bit-identical to the CPU loop for the same float homographies.  cv::remap and the contour mask (OpenCV) stay
// out of scope, like block matching.
#include <algorithm>
#include <cmath>

#include "amb_context.h"

namespace amb {
namespace {

constexpr int kRectThreads = 256;

struct RectifyParams {
  float t1[9], t2[9];  // T1_inv, T2_inv row-major
  int width, height;
  size_t stride;       // floats per map row
  float* m1x;
  float* m1y;
  float* m2x;
  float* m2y;
  int* zero_w;         // nullable: set to 1 if any w == 0 (CHECK_NE(xyw(2), 0.0), rectifier.cpp:92,99)
};

__device__ __forceinline__ bool homography(const float* T, float fu, float fv, float* mx, float* my) {
  const float x = __fadd_rn(__fadd_rn(__fmul_rn(T[0], fu), __fmul_rn(T[1], fv)), T[2]);
  const float y = __fadd_rn(__fadd_rn(__fmul_rn(T[3], fu), __fmul_rn(T[4], fv)), T[5]);
  const float w = __fadd_rn(__fadd_rn(__fmul_rn(T[6], fu), __fmul_rn(T[7], fv)), T[8]);
  *mx = __fdiv_rn(x, w);
  *my = __fdiv_rn(y, w);
  return w == 0.0f;
}

template <bool VEC4>
__global__ void __launch_bounds__(kRectThreads) rectify_maps_kernel(const __grid_constant__ RectifyParams p) {
  const int groups_per_row = (p.width + 3) >> 2;
  const size_t n_groups = static_cast<size_t>(groups_per_row) * p.height;
  bool bad = false;
  for (size_t g = blockIdx.x * static_cast<size_t>(kRectThreads) + threadIdx.x; g < n_groups;
       g += static_cast<size_t>(gridDim.x) * kRectThreads) {
    const int v = static_cast<int>(g / groups_per_row);
    const int u0 = static_cast<int>(g - static_cast<size_t>(v) * groups_per_row) << 2;
    const float fv = static_cast<float>(v);
    float a[4], b[4], c[4], d[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float fu = static_cast<float>(u0 + k);
      bad |= homography(p.t1, fu, fv, &a[k], &b[k]) && (u0 + k < p.width);
      bad |= homography(p.t2, fu, fv, &c[k], &d[k]) && (u0 + k < p.width);
    }
    const size_t off = static_cast<size_t>(v) * p.stride + u0;
    if (VEC4 && u0 + 3 < p.width) {
      *reinterpret_cast<float4*>(p.m1x + off) = make_float4(a[0], a[1], a[2], a[3]);
      *reinterpret_cast<float4*>(p.m1y + off) = make_float4(b[0], b[1], b[2], b[3]);
      *reinterpret_cast<float4*>(p.m2x + off) = make_float4(c[0], c[1], c[2], c[3]);
      *reinterpret_cast<float4*>(p.m2y + off) = make_float4(d[0], d[1], d[2], d[3]);
    } else {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (u0 + k < p.width) {
          p.m1x[off + k] = a[k];
          p.m1y[off + k] = b[k];
          p.m2x[off + k] = c[k];
          p.m2y[off + k] = d[k];
        }
      }
    }
  }
  if (bad && p.zero_w) atomicExch(p.zero_w, 1);
}

// ---- host side of rectifyStereoPair (rectifier.cpp:43-79), double precision, row-major 3x3 -------------------
inline void cross3(const double* a, const double* b, double* o) {
  o[0] = a[1] * b[2] - a[2] * b[1];
  o[1] = a[2] * b[0] - a[0] * b[2];
  o[2] = a[0] * b[1] - a[1] * b[0];
}
inline void unit3(const double* v, double* o) {
  const double n = std::sqrt((v[0] * v[0] + v[1] * v[1]) + v[2] * v[2]);
  for (int k = 0; k < 3; ++k) o[k] = v[k] / n;
}
inline void mat_mul(const double* A, const double* B, double* C) {
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) C[3 * i + j] = (A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j]) + A[3 * i + 2] * B[6 + j];
}
inline double minor_signed(const double* m, int i, int j) {  // cofactor of element (i, j), cyclic form
  const int i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (j + 1) % 3, j2 = (j + 2) % 3;
  return m[3 * i1 + j1] * m[3 * i2 + j2] - m[3 * i1 + j2] * m[3 * i2 + j1];
}
inline bool mat_inverse(const double* m, double* inv) {  // adjugate / determinant
  const double det = (minor_signed(m, 0, 0) * m[0] + minor_signed(m, 1, 0) * m[3]) + minor_signed(m, 2, 0) * m[6];
  if (det == 0.0) return false;
  const double invdet = 1.0 / det;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) inv[3 * j + i] = minor_signed(m, i, j) * invdet;
  return true;
}

bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

}  // namespace
}  // namespace amb

using namespace amb;

extern "C" int amb_stereo_rectify_setup(const double* K, const double* R_G_C1, const double* R_G_C2,
                                        const double* t_G_C1, const double* t_G_C2, double* baseline,
                                        double* R_G_C_rect, float* T1_inv, float* T2_inv) {
  if (!K || !R_G_C1 || !R_G_C2 || !t_G_C1 || !t_G_C2 || !baseline || !R_G_C_rect || !T1_inv || !T2_inv)
    return AMB_ERR_INVALID_ARGUMENT;
  const double x[3] = {t_G_C2[0] - t_G_C1[0], t_G_C2[1] - t_G_C1[1], t_G_C2[2] - t_G_C1[2]};  // rectifier.cpp:46
  *baseline = std::sqrt((x[0] * x[0] + x[1] * x[1]) + x[2] * x[2]);                            // :47
  if (*baseline == 0.0) return AMB_ERR_CHECK_FAILED;  // the densifier CHECKs it (densifier.cpp:39)
  const double z_old[3] = {R_G_C1[2], R_G_C1[5], R_G_C1[8]};  // R_G_C1.col(2)
  double y[3], z[3];
  cross3(z_old, x, y);  // :50
  cross3(x, y, z);      // :53
  unit3(x, R_G_C_rect + 0);  // rows of R_G_C_rect (:56-59)
  unit3(y, R_G_C_rect + 3);
  unit3(z, R_G_C_rect + 6);
  double KR[9], Rt[9], Q[9], Qinv[9], T[9], Tinv[9];
  mat_mul(K, R_G_C_rect, KR);  // P_rect.block<3,3>(0,0) (:64-71)
  const double* R_in[2] = {R_G_C1, R_G_C2};
  float* T_out[2] = {T1_inv, T2_inv};
  for (int c = 0; c < 2; ++c) {
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) Rt[3 * i + j] = R_in[c][3 * j + i];
    mat_mul(K, Rt, Q);  // Q_i = K R_G_Ci^T (:74-75)
    if (!mat_inverse(Q, Qinv)) return AMB_ERR_CHECK_FAILED;
    mat_mul(KR, Qinv, T);  // T_i_rect (:76-77)
    if (!mat_inverse(T, Tinv)) return AMB_ERR_CHECK_FAILED;
    for (int k = 0; k < 9; ++k) T_out[c][k] = static_cast<float>(Tinv[k]);  // .cast<float>() (:78-79)
  }
  return AMB_OK;
}

extern "C" int amb_stereo_rectify_maps_device(int device, void* stream, const float* T1_inv, const float* T2_inv,
                                              int32_t width, int32_t height, size_t map_stride, float* d_map1_x,
                                              float* d_map1_y, float* d_map2_x, float* d_map2_y,
                                              int32_t* d_zero_w_flag) {
  if (!T1_inv || !T2_inv || !d_map1_x || !d_map1_y || !d_map2_x || !d_map2_y || width <= 0 || height <= 0 ||
      map_stride < static_cast<size_t>(width))
    return AMB_ERR_INVALID_ARGUMENT;
  int n_dev = 0;
  if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev <= 0 || cudaSetDevice(device) != cudaSuccess) {
    cudaGetLastError();
    return AMB_ERR_NO_DEVICE;
  }
  RectifyParams p;
  for (int k = 0; k < 9; ++k) {
    p.t1[k] = T1_inv[k];
    p.t2[k] = T2_inv[k];
  }
  p.width = width;
  p.height = height;
  p.stride = map_stride;
  p.m1x = d_map1_x;
  p.m1y = d_map1_y;
  p.m2x = d_map2_x;
  p.m2y = d_map2_y;
  p.zero_w = d_zero_w_flag;
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
  const size_t n_groups = static_cast<size_t>((width + 3) / 4) * height;
  const size_t want = (n_groups + kRectThreads - 1) / kRectThreads;
  const int blocks = static_cast<int>(std::max<size_t>(1, std::min<size_t>(want, static_cast<size_t>(sms) * 8)));
  const bool vec4 = (map_stride % 4 == 0) && aligned16(d_map1_x) && aligned16(d_map1_y) && aligned16(d_map2_x) &&
                    aligned16(d_map2_y);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (vec4) {
    rectify_maps_kernel<true><<<blocks, kRectThreads, 0, s>>>(p);
  } else {
    rectify_maps_kernel<false><<<blocks, kRectThreads, 0, s>>>(p);
  }
  if (cudaGetLastError() != cudaSuccess) return AMB_ERR_CUDA;
  return AMB_OK;
}

extern "C" int amb_stereo_rectify_maps(int device, const float* T1_inv, const float* T2_inv, int32_t width,
                                       int32_t height, size_t map_stride, float* map1_x, float* map1_y,
                                       float* map2_x, float* map2_y) {
  if (!T1_inv || !T2_inv || !map1_x || !map1_y || !map2_x || !map2_y || width <= 0 || height <= 0 ||
      map_stride < static_cast<size_t>(width))
    return AMB_ERR_INVALID_ARGUMENT;
  int n_dev = 0;
  if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev <= 0 || cudaSetDevice(device) != cudaSuccess) {
    cudaGetLastError();
    return AMB_ERR_NO_DEVICE;
  }
  const size_t pitch = (static_cast<size_t>(width) + 3) & ~static_cast<size_t>(3);  // device rows: multiple of 4 floats
  const size_t plane = pitch * height;
  float* d_maps = nullptr;
  int32_t* d_flag = nullptr;
  cudaStream_t s = nullptr;
  int st = AMB_OK;
  auto ok = [&](cudaError_t e) {
    if (e != cudaSuccess && st == AMB_OK) {
      cudaGetLastError();
      st = AMB_ERR_CUDA;
    }
    return e == cudaSuccess;
  };
  ok(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
  ok(cudaMalloc(&d_maps, 4 * plane * sizeof(float)));
  ok(cudaMalloc(&d_flag, sizeof(int32_t)));
  int32_t flag = 0;
  if (st == AMB_OK) ok(cudaMemsetAsync(d_flag, 0, sizeof(int32_t), s));
  if (st == AMB_OK)
    st = amb_stereo_rectify_maps_device(device, s, T1_inv, T2_inv, width, height, pitch, d_maps, d_maps + plane,
                                        d_maps + 2 * plane, d_maps + 3 * plane, d_flag);
  if (st == AMB_OK) {
    float* outs[4] = {map1_x, map1_y, map2_x, map2_y};
    for (int k = 0; k < 4; ++k)
      ok(cudaMemcpy2DAsync(outs[k], map_stride * sizeof(float), d_maps + k * plane, pitch * sizeof(float),
                           static_cast<size_t>(width) * sizeof(float), height, cudaMemcpyDeviceToHost, s));
    ok(cudaMemcpyAsync(&flag, d_flag, sizeof(int32_t), cudaMemcpyDeviceToHost, s));
    ok(cudaStreamSynchronize(s));
  }
  cudaFree(d_maps);
  cudaFree(d_flag);
  if (s) cudaStreamDestroy(s);
  if (st == AMB_OK && flag) st = AMB_ERR_CHECK_FAILED;  // CHECK_NE(xyw(2), 0.0)
  return st;
}
