// shim_demo.cc — the call sequence of the reference's batch demo (main-ortho-backward-grid.cc:119-141): caller code
// that only uses the reference's public API.  ONE source, two builds (tests/test_shim.py):
//   * against the DROP-IN headers (aerial_mapper_b200/shim, -DAMB_SHIM_MINI stand-ins) + libaerial_mapper_b200.so
//     -> the CUDA path;
//   * against the REFERENCE'S OWN headers and sources (dsm.cc, ortho-backward-grid.cc compiled verbatim from
//     /root/reference with oracle/refsrc_stubs) -> oracle/_ref/libamb_reference_demo.so, the CPU path
//     (-DAMB_DEMO_AS_LIBRARY: main() gets a C name so the test can call it through ctypes).
// Reads a scenario file written by the test, writes the resulting layers back.
#include <aerial-mapper-dsm/dsm.h>
#include <aerial-mapper-ortho/ortho-backward-grid.h>

#ifdef AMB_DEMO_AS_LIBRARY
extern "C" __attribute__((visibility("default"))) int amb_demo_main(int argc, char** argv);
#define main amb_demo_main
#endif

#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <fstream>

template <typename T>
static void rd(std::ifstream& f, T* p, size_t n) {
  f.read(reinterpret_cast<char*>(p), sizeof(T) * n);
}

int main(int argc, char** argv) {
  if (argc < 3) return 2;
  std::ifstream f(argv[1], std::ios::binary);
  double geo[5];  // delta_easting, delta_northing, resolution, center_easting, center_northing
  rd(f, geo, 5);
  int64_t hdr[6];  // n_points, n_frames, width, height, channels, dist_type
  rd(f, hdr, 6);
  double cam[8];  // fu fv cu cv d0..d3
  rd(f, cam, 8);
  AlignedType<std::vector, Eigen::Vector3d>::type cloud(static_cast<size_t>(hdr[0]));
  rd(f, &cloud[0](0), 3 * cloud.size());
  std::vector<double> poses(7 * hdr[1]);
  rd(f, poses.data(), poses.size());
  Images images;
  for (int64_t k = 0; k < hdr[1]; ++k) {
    cv::Mat m(static_cast<int>(hdr[3]), static_cast<int>(hdr[2]), static_cast<int>(hdr[4]));
    rd(f, m.data, static_cast<size_t>(m.rows) * m.step);
    images.push_back(m);
  }

  // AerialGridMap::initialize (aerial-mapper-grid-map.cc:23-49)
  grid_map::GridMap map({"ortho", "elevation", "elevation_angle", "num_observations", "elevation_angle_first_view",
                         "delta", "observation_index", "observation_index_first", "colored_ortho"});
  map.setFrameId("world");
  grid_map::Length len;
  len(0) = geo[0];
  len(1) = geo[1];
  grid_map::Position pos;
  pos(0) = geo[3];
  pos(1) = geo[4];
  map.setGeometry(len, geo[2], pos);
  map["ortho"].setConstant(255);
  map["elevation_angle"].setConstant(0.0);
  map["num_observations"].setConstant(0);

  // optional third argument: repetitions of the (DSM, orthomosaic) sequence on re-initialised layers, each timed
  // (tools/shim_bench.py: what the reference's batch demo would get at the benchmark size)
  const int reps = argc > 3 ? std::atoi(argv[3]) : 1;
  typedef std::chrono::steady_clock Clock;

  dsm::Settings settings_dsm;  // main-ortho-backward-grid.cc:129-133
  settings_dsm.center_easting = 0.0;
  settings_dsm.center_northing = 0.0;
  dsm::Dsm digital_surface_map(settings_dsm, &map);

  Eigen::Vector4d intr, dist;
  for (int k = 0; k < 4; ++k) {
    intr(k) = cam[k];
    dist(k) = cam[4 + k];
  }
  const aslam::Distortion::Type types[3] = {aslam::Distortion::Type::kNoDistortion, aslam::Distortion::Type::kRadTan,
                                            aslam::Distortion::Type::kEquidistant};
  std::shared_ptr<aslam::NCamera> ncameras(new aslam::NCamera(
      aslam::Camera(static_cast<unsigned>(hdr[2]), static_cast<unsigned>(hdr[3]), intr,
                    aslam::Distortion(types[hdr[5]], dist)),
      kindr::minimal::QuatTransformation()));
  Poses T_G_Bs;
  for (int64_t k = 0; k < hdr[1]; ++k) {
    const double* p = &poses[7 * k];
    T_G_Bs.push_back(Pose(p[3], p[4], p[5], p[6], p[0], p[1], p[2]));
  }
  ortho::Settings settings_ortho;  // :136-141
  settings_ortho.colored_ortho = hdr[4] == 3;
  ortho::OrthoBackwardGrid mosaic(ncameras, settings_ortho, &map);

  for (int rep = 0; rep < reps; ++rep) {
    if (rep > 0) {  // a fresh map (AerialGridMap::initialize again), outside the timed region
      map["elevation"].setConstant(NAN);
      map["observation_index"].setConstant(NAN);
      map["colored_ortho"].setConstant(NAN);
      map["ortho"].setConstant(255);
      map["elevation_angle"].setConstant(0.0);
    }
    const Clock::time_point t0 = Clock::now();
    digital_surface_map.process(cloud, &map);
    const Clock::time_point t1 = Clock::now();
    mosaic.process(T_G_Bs, images, &map);
    const Clock::time_point t2 = Clock::now();
    if (reps > 1)
      std::printf("shim_demo rep %d: dsm.process %.2f ms, ortho.process %.2f ms, total %.2f ms\n", rep,
                  std::chrono::duration<double, std::milli>(t1 - t0).count(),
                  std::chrono::duration<double, std::milli>(t2 - t1).count(),
                  std::chrono::duration<double, std::milli>(t2 - t0).count());
  }

  std::ofstream o(argv[2], std::ios::binary);
  const char* names[5] = {"elevation", "elevation_angle", "observation_index", "ortho", "colored_ortho"};
  for (const char* n : names)
    o.write(reinterpret_cast<const char*>(map[n].data()), sizeof(float) * map[n].rows() * map[n].cols());
  std::printf("shim_demo ok %d x %d\n", map.getSize()(0), map.getSize()(1));
  return 0;
}
