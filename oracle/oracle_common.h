/*
 * oracle_common.h — shared pieces of the CPU oracle (TEST INFRASTRUCTURE; see amb_oracle.h).
 * Build with -ffp-contract=off: every expression below is meant to round exactly like the reference's
 * un-contracted x86-64 double arithmetic.
 */
#ifndef AMB_ORACLE_COMMON_H_
#define AMB_ORACLE_COMMON_H_

#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>

#include "amb_oracle.h"

namespace ambo {

/* grid_map::getPositionFromIndex (grid_map_core GridMapMath.cpp; call sites dsm.cc:124-125,
 * ortho-backward-grid.cc:149-150), restated from upstream (dependency un-versioned and absent):
 *   offset   = 0.5*length - 0.5*resolution                 (getVectorToFirstCell)
 *   position = (mapPosition + offset) + resolution * (-(double)index)
 * with startIndex = 0 (the map is never moved). Index (0,0) is the max-x / max-y corner. */
inline void cellPosition(const amb_geometry& g, int i, int j, double* x, double* y) {
  const double off_x = 0.5 * g.length_x - 0.5 * g.resolution;
  const double off_y = 0.5 * g.length_y - 0.5 * g.resolution;
  const double base_x = g.pos_x + off_x;
  const double base_y = g.pos_y + off_y;
  *x = base_x + g.resolution * (-static_cast<double>(i));
  *y = base_y + g.resolution * (-static_cast<double>(j));
}

/* The thresholds the reference's retry loop visits (dsm.cc:133-144):
 *   lambda = 1.0; while (empty) { search(lambda * radius); lambda *= 1.1; if (lambda * radius > 7.0) break; }
 * The first query (dsm.cc:127-131) uses (double)radius, which equals threshold 0. */
inline std::vector<double> dsmThresholds(int interpolation_radius) {
  std::vector<double> thr;
  double lambda = 1.0;
  while (true) {
    thr.push_back(lambda * interpolation_radius);
    lambda *= 1.1;
    if (lambda * interpolation_radius > 7.0) break;
  }
  return thr;
}

/* utils::parFor (utils-common.h:29-59): ceil(n/T) items per block, ceil(n/items) blocks, one std::thread per
 * block, join all.  The reference materialises each block as a std::vector<size_t> of indices; here a block is
 * the half-open range it would contain (identical visiting order, identical results). */
template <typename Functor>
void parFor(int64_t num_items, const Functor& functor, size_t num_threads) {
  if (num_items <= 0) return;
  if (num_threads == 0) num_threads = 1;
  const int64_t per_block = static_cast<int64_t>(
      std::ceil(static_cast<double>(num_items) / static_cast<double>(num_threads)));
  const int64_t num_blocks =
      static_cast<int64_t>(std::ceil(static_cast<double>(num_items) / static_cast<double>(per_block)));
  std::vector<std::thread> threads;
  for (int64_t b = 0; b < num_blocks; ++b) {
    const int64_t lo = b * per_block;
    const int64_t hi = std::min(num_items, lo + per_block);
    threads.push_back(std::thread([&functor, lo, hi]() { functor(lo, hi); }));
  }
  for (auto& t : threads) t.join();
}

inline size_t resolveThreads(int32_t num_threads) {
  if (num_threads > 0) return static_cast<size_t>(num_threads);
  size_t hw = std::thread::hardware_concurrency();
  return hw == 0 ? 1 : hw;
}

inline double now() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

}  // namespace ambo
#endif
