// dsm_kernels.cu — point cloud -> `elevation` layer on sm_100a.
//
// Replaces dsm::Dsm::process (reference aerial_mapper_dsm/src/dsm.cc:186-201): kd-tree build (:36-52) + one
// radius query and IDW per cell (:113-184).  The reference's result for a cell is
//     S   = { p : d2(p) < threshold }                 d2 = (qx-px)^2 + (qy-py)^2, un-contracted double
//     h   = (sum_S z/d2) / (sum_S 1/d2)                elevation(i,j) = (float)h
// with threshold = (double)interpolation_radius, or, when that set is empty, the first of the retry thresholds
// lambda_k*radius (dsm.cc:133-144) that yields a non-empty set.  Nothing in that definition needs a tree: this
// file bins the points to grid cells (one bin per cell plus an apron), and every cell gathers from the bins its
// threshold can reach.
//
// Pipeline (all on the context's stream):
//   K1 dsm_count_kernel      read AoS points, shift, bin id, atomicAdd histogram
//   K2 scan_* (3 kernels)    inclusive scan of the histogram -> bin start offsets
//   K3 dsm_scatter_kernel    write 32-byte records {x, y, z, original index} into bin order
//   K3b dsm_canon_kernel     order the records of every multi-point bin by original index, so that the IDW
//                            summation order (hence every output bit) is independent of atomic scheduling and
//                            of how the map is striped across GPUs
//   K4 dsm_gather_kernel     one 32x32 cell tile per block: the tile's bin ranges are staged through shared
//                            memory (SoA doubles), each thread walks the bin rows of its cell's window
//   K5 dsm_fill_kernel       cells with no neighbour inside the primary threshold: one warp per cell finds the
//                            minimum d2 in the fallback window, picks the reference's threshold index, re-sums
//
// Membership (d2 < threshold) is evaluated with __dmul_rn/__dadd_rn (no FMA contraction) in the reference's
// operation order, so neighbour sets and fallback levels are bit-exact decisions.  Heights differ from the CPU
// only by the order of the double-precision summation.
#include <cfloat>
#include <cmath>

#include "amb_context.h"

namespace amb {

std::vector<double> dsm_thresholds(int32_t interpolation_radius) {
  // dsm.cc:133-144: lambda = 1; while (empty) { search(lambda*radius); lambda *= 1.1; if (lambda*radius > 7) break; }
  std::vector<double> thr;
  double lambda = 1.0;
  while (true) {
    thr.push_back(lambda * interpolation_radius);
    lambda *= 1.1;
    if (lambda * interpolation_radius > 7.0) break;
  }
  return thr;
}

namespace {

constexpr int kMaxThresholds = 32;
constexpr int kMaxHalfWidth = 128;  // cells; window half-width limit (sqrt(threshold)/resolution)
constexpr int TI = 32;              // tile extent along i (rows, the contiguous axis of the layer)
constexpr int TJ = 32;              // tile extent along j
constexpr int kGatherThreads = 256;
constexpr int kStrip = 4;           // cells per thread in the gather kernel (adjacent along j)
constexpr int kScanChunk = 4096;    // elements per block in the scan kernels (256 threads x 16)
static_assert(TJ == kStrip * (kGatherThreads / 32), "one warp per strip row of the tile");

struct PointRec {  // 32 bytes, one DRAM sector
  double x, y, z;
  unsigned long long idx;
};

struct DsmPlan {
  int rows, cols_slab, col_begin;
  int P;        // apron of the bin grid in cells (reach of the largest fallback threshold)
  int W;        // window half-width of the primary threshold
  int BR, BC;   // bin grid: (rows + 2P) x (cols_slab + 2P); bin (bi, bj) <-> cell (bi - P, col_begin + bj - P)
  double base_x, base_y;  // cell centre of index 0: pos + (0.5*length - 0.5*res)   (grid_map getPosition)
  double res, inv_res;
  double shift_x, shift_y;  // dsm.cc:42-43: x -= center_northing, y -= center_easting
  double thr0;
  int n_thr;
  double thr[kMaxThresholds];
  short hw[kMaxHalfWidth + 1];   // half-width along i of the primary window at |dj|
  short hwf[kMaxHalfWidth + 1];  // same for the largest fallback threshold
};

__device__ __forceinline__ double cell_x(const DsmPlan& p, int i) {
  // position = (mapPosition + offset) + resolution * (-(double)index); un-contracted like the CPU
  return __dadd_rn(p.base_x, __dmul_rn(p.res, -static_cast<double>(i)));
}
__device__ __forceinline__ double cell_y(const DsmPlan& p, int j_global) {
  return __dadd_rn(p.base_y, __dmul_rn(p.res, -static_cast<double>(j_global)));
}

// Bin of a shifted point, or -1 if it cannot reach any cell of this slab.  A point is assigned to the cell whose
// centre is nearest; an off-by-one at a cell edge is harmless because membership is re-decided exactly by d2 and
// every window carries half a cell of slack (see half_widths()).
__device__ __forceinline__ long long bin_of(const DsmPlan& p, double px, double py) {
  const double fi = floor((p.base_x - px) * p.inv_res + 0.5);
  const double fj = floor((p.base_y - py) * p.inv_res + 0.5);
  const double bi = fi + static_cast<double>(p.P);
  const double bj = fj - static_cast<double>(p.col_begin) + static_cast<double>(p.P);
  if (!(bi >= 0.0 && bi < static_cast<double>(p.BR) && bj >= 0.0 && bj < static_cast<double>(p.BC))) return -1;
  return static_cast<long long>(bi) + static_cast<long long>(bj) * p.BR;
}

// ---------------------------------------------------------------------------------------------------------------
// K1: histogram
__global__ void __launch_bounds__(256) dsm_count_kernel(const double* __restrict__ xyz, size_t n, DsmPlan plan,
                                                        unsigned int* __restrict__ G,
                                                        unsigned int* __restrict__ counters) {
  unsigned int local = 0;
  for (size_t t = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; t < n;
       t += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const double px = xyz[3 * t + 0] - plan.shift_x;
    const double py = xyz[3 * t + 1] - plan.shift_y;
    const long long b = bin_of(plan, px, py);
    if (b >= 0) {
      atomicAdd(&G[b + 2], 1u);
      ++local;
    }
  }
  // total binned points (stats only): warp-aggregate, one atomic per warp
  for (int o = 16; o > 0; o >>= 1) local += __shfl_down_sync(0xffffffffu, local, o);
  if ((threadIdx.x & 31) == 0 && local) atomicAdd(&counters[2], local);
}

// ---------------------------------------------------------------------------------------------------------------
// K2: inclusive scan (reduce / spine / apply).  Plain HBM streaming: 3 reads + 1 write of the histogram.
__device__ __forceinline__ unsigned int block_exclusive_scan(unsigned int v, unsigned int* total) {
  // `total` must point to shared memory; valid after return.
  __shared__ unsigned int warp_sums[32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nwarps = blockDim.x >> 5;
  unsigned int inc = v;
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned int t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) warp_sums[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    const unsigned int w = lane < nwarps ? warp_sums[lane] : 0u;
    unsigned int winc = w;
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned int t = __shfl_up_sync(0xffffffffu, winc, o);
      if (lane >= o) winc += t;
    }
    warp_sums[lane] = winc - w;  // exclusive prefix of the warp totals
    if (lane == 31) *total = winc;
  }
  __syncthreads();
  const unsigned int r = warp_sums[warp] + inc - v;
  __syncthreads();  // warp_sums is reused by the next call
  return r;
}

__global__ void __launch_bounds__(256) scan_reduce_kernel(const uint4* __restrict__ in, size_t n_vec,
                                                          unsigned int* __restrict__ block_sums) {
  const size_t base = static_cast<size_t>(blockIdx.x) * (kScanChunk / 4);
  unsigned int s = 0;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const size_t v = base + k * 256 + threadIdx.x;
    if (v < n_vec) {
      const uint4 q = in[v];
      s += q.x + q.y + q.z + q.w;
    }
  }
  __shared__ unsigned int total;
  block_exclusive_scan(s, &total);
  if (threadIdx.x == 0) block_sums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(1024) scan_spine_kernel(unsigned int* __restrict__ block_sums, int n) {
  __shared__ unsigned int total;
  __shared__ unsigned int carry_s;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  for (int base = 0; base < n; base += 1024) {
    const int k = base + threadIdx.x;
    const unsigned int v = k < n ? block_sums[k] : 0u;
    const unsigned int ex = block_exclusive_scan(v, &total);
    const unsigned int carry = carry_s;
    if (k < n) block_sums[k] = ex + carry;
    __syncthreads();
    if (threadIdx.x == 0) carry_s = carry + total;
    __syncthreads();
  }
}

__global__ void __launch_bounds__(256) scan_apply_kernel(uint4* __restrict__ data, size_t n_vec,
                                                         const unsigned int* __restrict__ block_sums) {
  // blocked arrangement: thread t owns 4 consecutive uint4 (16 values) so that the scan order is the array order
  const size_t base = static_cast<size_t>(blockIdx.x) * (kScanChunk / 4) + threadIdx.x * 4;
  uint4 q[4];
  unsigned int s = 0;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    q[k] = (base + k < n_vec) ? data[base + k] : make_uint4(0, 0, 0, 0);
    s += q[k].x + q[k].y + q[k].z + q[k].w;
  }
  __shared__ unsigned int total;
  unsigned int run = block_exclusive_scan(s, &total) + block_sums[blockIdx.x];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    q[k].x += run;
    q[k].y += q[k].x;
    q[k].z += q[k].y;
    q[k].w += q[k].z;
    run = q[k].w;
    if (base + k < n_vec) data[base + k] = q[k];
  }
}

// ---------------------------------------------------------------------------------------------------------------
// K3: scatter.  H = G + 1 holds start(b) before this kernel; atomicAdd turns H[b] into start(b+1), i.e.
// afterwards G[b] = start(b) for b in [0, nb].
__global__ void __launch_bounds__(256) dsm_scatter_kernel(const double* __restrict__ xyz, size_t n, DsmPlan plan,
                                                          unsigned int* __restrict__ G,
                                                          PointRec* __restrict__ rec) {
  for (size_t t = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; t < n;
       t += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const double px = xyz[3 * t + 0] - plan.shift_x;
    const double py = xyz[3 * t + 1] - plan.shift_y;
    const double pz = xyz[3 * t + 2];
    const long long b = bin_of(plan, px, py);
    if (b >= 0) {
      const unsigned int pos = atomicAdd(&G[b + 1], 1u);
      double2* dst = reinterpret_cast<double2*>(rec + pos);
      dst[0] = make_double2(px, py);
      dst[1] = make_double2(pz, __longlong_as_double(static_cast<long long>(t)));
    }
  }
}

// K3b: canonical order inside every bin (ascending original index = what a stable sort would give).
__global__ void __launch_bounds__(256) dsm_canon_kernel(const unsigned int* __restrict__ G, size_t nb,
                                                        PointRec* __restrict__ rec) {
  const size_t b = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
  if (b >= nb) return;
  const unsigned int s = G[b], e = G[b + 1];
  if (e - s < 2) return;
  for (unsigned int a = s + 1; a < e; ++a) {  // insertion sort; bins hold O(1) points for aerial densities
    const PointRec key = rec[a];
    unsigned int k = a;
    while (k > s && rec[k - 1].idx > key.idx) {
      rec[k] = rec[k - 1];
      --k;
    }
    if (k != a) rec[k] = key;
  }
}

// Reciprocal of a positive normal double to ~1 ulp: MUFU.RCP64H seed (rcp.approx.ftz.f64, ~2^-23) + one cubic
// correction step.  Replaces the two IEEE divisions per neighbour of the reference (heights/d2 and 1/d2) by one reciprocal
// and one FMA; the IDW height changes by O(1e-16) relative, far inside the float32 layer's rounding.
__device__ __forceinline__ double fast_rcp(double x) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  const double e = fma(-x, r, 1.0);  // |e| <= 2^-23
  const double t = fma(e, e, e);     // e + e^2: r*(1 + e + e^2) leaves a relative error e^3 <= 2^-69
  return fma(r, t, r);
}

// ---------------------------------------------------------------------------------------------------------------
// K4: tile gather
struct GatherArgs {
  const unsigned int* G;
  const PointRec* rec;
  float* elevation;
  unsigned int* empty_list;
  unsigned int* counters;  // [0] empty count, [1] error flag
  int* dbg_count;
  signed char* dbg_level;
  int capacity;  // points that fit the shared-memory stage
  int tiles_i;
};

// Per-thread work of the gather kernel.  A thread owns kStrip cells with the same i and adjacent j; their windows
// overlap in all but one bin row, so every staged point is loaded once, (qx - px)^2 is computed once, and only
// the (qy - py)^2 + compare + accumulate part is per cell.  The thread walks the kStrip + 2W bin rows of its
// strip as ONE flattened loop (lanes stay busy until their own candidates run out); per row, `srow` gives the
// half-width of the widest window among the strip's cells and the mask of cells that row can reach.
// Accumulation order per cell: bin rows ascending, points ascending inside a row == canonical bin order.
template <bool STAGED, bool DEBUG>
__device__ __forceinline__ void gather_strip(const DsmPlan& plan, const GatherArgs& args, int off_srow, int off_sxy,
                                             int off_spz, int i0, int j0) {
  extern __shared__ __align__(16) unsigned char smem_raw[];  // addressed by byte offset: plain LDS, no generic ptr
  const int W = plan.W;
  const int NI = TI + 2 * W + 1;
  const int ti = threadIdx.x & 31;
  const int tq = threadIdx.x >> 5;  // warp index == strip index along j (warp-uniform)
  const int i = i0 + ti;
  const int jl0 = j0 + kStrip * tq;
  const double qx = cell_x(plan, i);
  const double thr0 = plan.thr0;
  double qy[kStrip], num[kStrip], den[kStrip];
  int cnt[kStrip];
  unsigned int cellmask = 0;
#pragma unroll
  for (int m = 0; m < kStrip; ++m) {
    qy[m] = cell_y(plan, plan.col_begin + jl0 + m);
    num[m] = 0.0;
    den[m] = 0.0;
    cnt[m] = 0;
    if (i < plan.rows && jl0 + m < plan.cols_slab) cellmask |= 1u << m;
  }
  const int n_rows = kStrip + 2 * W;
  int r = -1;                                                    // row of the strip's window
  int row_addr = ((kStrip * tq - 1) * NI + ti + W) * 4;          // byte offset of soff[row][ti + W]
  unsigned int k = 0, e = 0;
  if (cellmask) {
    for (;;) {
      if (k >= e) {
        bool found = false;
        while (r + 1 < n_rows) {
          ++r;
          row_addr += NI * 4;
          const int info = *reinterpret_cast<const int*>(smem_raw + off_srow + 4 * r);  // (h << 8) | mask, or -1
          if (info < 0) continue;
          const int h4 = (info >> 8) * 4;
          k = *reinterpret_cast<const unsigned int*>(smem_raw + row_addr - h4);
          e = *reinterpret_cast<const unsigned int*>(smem_raw + row_addr + h4 + 4);
          if (k < e) {
            found = true;
            break;
          }
        }
        if (!found) break;
      }
      double2 p;
      double pz;
      if (STAGED) {
        p = *reinterpret_cast<const double2*>(smem_raw + off_sxy + 16 * k);
        pz = *reinterpret_cast<const double*>(smem_raw + off_spz + 8 * k);
      } else {
        p = __ldg(reinterpret_cast<const double2*>(args.rec + k));
        pz = __ldg(reinterpret_cast<const double*>(args.rec + k) + 2);
      }
      const double dx = qx - p.x;
      const double dx2 = __dmul_rn(dx, dx);
      // Branch-free: every cell of the strip tests every staged point of the strip's rows.  A bin row that is out
      // of a cell's reach (|dj| > W) holds only points with |dy| > sqrt(threshold), so it can never hit; a miss
      // adds exactly 0 (w = 0: fma(z, 0, num) == num, den + 0 == den), so the sums are bit-identical to
      // visiting only the reachable rows.
#pragma unroll
      for (int m = 0; m < kStrip; ++m) {
        const double dy = qy[m] - p.y;
        const double d2 = __dadd_rn(dx2, __dmul_rn(dy, dy));  // L2_Adaptor (nanoflann.hpp:325-328), un-contracted
        const bool hit = d2 < thr0;                           // RadiusResultSet::addPoint (:156-158)
        const double w = hit ? fast_rcp(d2) : 0.0;            // 1.0 / distances[i]         (dsm.cc:167)
        num[m] = fma(pz, w, num[m]);                          // heights[i] / distances[i]  (dsm.cc:166), z * (1/d2)
        den[m] += w;
        if (DEBUG) cnt[m] += hit ? 1 : 0;
      }
      ++k;
    }
  }
  bool coincident = false;
#pragma unroll
  for (int m = 0; m < kStrip; ++m) {
    const bool valid = (cellmask >> m) & 1u;
    const size_t cell = static_cast<size_t>(jl0 + m) * plan.rows + i;
    const bool has = den[m] > 0.0 || den[m] != den[m];  // any hit adds a positive weight (NaN: coincident point)
    if (valid) {
      if (has) {
        // d2 == 0 (a point exactly on the cell centre: reference CHECK(distances[i] > 0.0) aborts, dsm.cc:165)
        // is the only way a weight becomes inf/NaN: any d2 > 0 is >= 1e-26 for metre-scale coordinates.
        if (!(den[m] < DBL_MAX)) coincident = true;
        args.elevation[cell] = __double2float_rn(__ddiv_rn(num[m], den[m]));  // dsm.cc:171-172
      }
      if (DEBUG) {
        args.dbg_count[cell] = cnt[m];
        args.dbg_level[cell] = has ? 0 : -1;
      }
    }
    // cells the primary threshold left empty go to the retry pass: one atomic per warp
    const bool is_empty = valid && !has;
    const unsigned int mask = __ballot_sync(0xffffffffu, is_empty);
    if (mask) {
      const int leader = __ffs(mask) - 1;
      unsigned int base = 0;
      if (ti == leader) base = atomicAdd(&args.counters[0], static_cast<unsigned int>(__popc(mask)));
      base = __shfl_sync(0xffffffffu, base, leader);
      if (is_empty) args.empty_list[base + __popc(mask & ((1u << ti) - 1u))] = static_cast<unsigned int>(cell);
    }
  }
  if (coincident) atomicExch(&args.counters[1], 1u);
}

__global__ void __launch_bounds__(kGatherThreads)
    dsm_gather_kernel(const __grid_constant__ DsmPlan plan, const GatherArgs args) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int W = plan.W;
  const int NC = TJ + 2 * W;      // staged bin columns
  const int NI = TI + 2 * W + 1;  // offsets per column (one past the last bin)
  unsigned int* soff = reinterpret_cast<unsigned int*>(smem_raw);           // [NC][NI]
  unsigned int* colbase = soff + NC * NI;                                    // [NC + 1] local start of a column
  unsigned int* colg0 = colbase + (NC + 1);                                  // [NC] global start of a column
  int* srow = reinterpret_cast<int*>(colg0 + NC);                            // [kStrip + 2W] strip row table
  size_t off_bytes = (static_cast<size_t>(NC) * NI + 2 * NC + 1 + kStrip + 2 * W) * sizeof(unsigned int);
  off_bytes = (off_bytes + 15) & ~static_cast<size_t>(15);
  double2* sxy = reinterpret_cast<double2*>(smem_raw + off_bytes);  // (x, y) of the staged points
  double* spz = reinterpret_cast<double*>(sxy + args.capacity);

  const int tile_i = blockIdx.x % args.tiles_i;
  const int tile_j = blockIdx.x / args.tiles_i;
  const int i0 = tile_i * TI;
  const int j0 = tile_j * TJ;  // local column in the slab

  for (int r = threadIdx.x; r < kStrip + 2 * W; r += kGatherThreads) {
    // strip row r is bin row (first cell's j) - W + r; cell m of the strip sees it at dj = r - W - m
    int hmax = -1, mask = 0;
    for (int m = 0; m < kStrip; ++m) {
      const int dj = r - W - m;
      const int ad = dj < 0 ? -dj : dj;
      const int h = ad <= W ? plan.hw[ad] : -1;
      if (h >= 0) {
        mask |= 1 << m;
        hmax = max(hmax, h);
      }
    }
    srow[r] = hmax < 0 ? -1 : ((hmax << 8) | mask);
  }
  // 1. bin start offsets of the tile + apron (global values)
  for (int e = threadIdx.x; e < NC * NI; e += kGatherThreads) {
    const int jj = e / NI, ii = e - jj * NI;
    const int bj = j0 - W + jj + plan.P;  // >= 0 because P >= W
    int bi = i0 - W + ii + plan.P;
    unsigned int v;
    if (bj >= plan.BC) {
      v = 0xffffffffu;  // marks "no such column": handled below as an empty column
    } else {
      bi = min(bi, plan.BR);  // bin BR of a row is the first bin of the next row: its start ends this row
      v = args.G[static_cast<size_t>(bj) * plan.BR + bi];
    }
    soff[e] = v;
  }
  __syncthreads();
  // 2. column lengths -> local bases
  if (threadIdx.x < 32) {
    unsigned int carry = 0;
    for (int base = 0; base < NC; base += 32) {
      const int jj = base + threadIdx.x;
      unsigned int len = 0, g0 = 0;
      if (jj < NC) {
        g0 = soff[jj * NI];
        const unsigned int g1 = soff[jj * NI + NI - 1];
        len = (g0 == 0xffffffffu) ? 0u : (g1 - g0);
        if (g0 == 0xffffffffu) g0 = 0;
      }
      unsigned int inc = len;
      for (int o = 1; o < 32; o <<= 1) {
        const unsigned int t = __shfl_up_sync(0xffffffffu, inc, o);
        if ((threadIdx.x & 31) >= o) inc += t;
      }
      if (jj < NC) {
        colbase[jj] = carry + inc - len;
        colg0[jj] = g0;
      }
      carry += __shfl_sync(0xffffffffu, inc, 31);
    }
    if (threadIdx.x == 0) colbase[NC] = carry;
  }
  __syncthreads();
  const unsigned int tile_points = colbase[NC];
  const bool staged = tile_points <= static_cast<unsigned int>(args.capacity);
  // 3. stage the records (SoA doubles) and rebase the offsets
  if (staged) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int jj = warp; jj < NC; jj += kGatherThreads / 32) {
      const unsigned int len = colbase[jj + 1] - colbase[jj];
      const PointRec* src = args.rec + colg0[jj];
      const unsigned int dst = colbase[jj];
      for (unsigned int k = lane; k < len; k += 32) {
        sxy[dst + k] = __ldg(reinterpret_cast<const double2*>(src + k));
        spz[dst + k] = __ldg(reinterpret_cast<const double*>(src + k) + 2);
      }
    }
  }
  __syncthreads();
  for (int e = threadIdx.x; e < NC * NI; e += kGatherThreads) {
    const int jj = e / NI;
    const unsigned int v = soff[e];
    if (v == 0xffffffffu) {
      soff[e] = 0;  // empty column: every range [0,0)
    } else if (staged) {
      soff[e] = v - colg0[jj] + colbase[jj];
    }
  }
  __syncthreads();

  // 4. per-cell gather: every thread owns a 1 x kStrip strip of cells (same i, adjacent j)
  const int off_srow = static_cast<int>(reinterpret_cast<unsigned char*>(srow) - smem_raw);
  const int off_sxy = static_cast<int>(off_bytes);
  const int off_spz = off_sxy + 16 * args.capacity;
  if (args.dbg_count) {  // tests: also record result_set.size() per cell
    if (staged) {
      gather_strip<true, true>(plan, args, off_srow, off_sxy, off_spz, i0, j0);
    } else {
      gather_strip<false, true>(plan, args, off_srow, off_sxy, off_spz, i0, j0);
    }
  } else if (staged) {
    gather_strip<true, false>(plan, args, off_srow, off_sxy, off_spz, i0, j0);
  } else {
    gather_strip<false, false>(plan, args, off_srow, off_sxy, off_spz, i0, j0);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// K5: expanding-radius retry for the cells the primary threshold left empty (dsm.cc:133-144).
struct FillArgs {
  const unsigned int* G;
  const PointRec* rec;
  float* elevation;
  const unsigned int* empty_list;
  unsigned int* counters;
  int* dbg_count;
  signed char* dbg_level;
};

__global__ void __launch_bounds__(256) dsm_fill_kernel(const __grid_constant__ DsmPlan plan, const FillArgs args) {
  const int lane = threadIdx.x & 31;
  const unsigned int n_empty = args.counters[0];
  const unsigned int warps_total = gridDim.x * (blockDim.x >> 5);
  const int P = plan.P;
  for (unsigned int c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); c < n_empty; c += warps_total) {
    const unsigned int cell = args.empty_list[c];
    const int i = static_cast<int>(cell % static_cast<unsigned int>(plan.rows));
    const int jl = static_cast<int>(cell / static_cast<unsigned int>(plan.rows));
    const double qx = cell_x(plan, i);
    const double qy = cell_y(plan, plan.col_begin + jl);
    // pass 1: smallest d2 in the window of the largest threshold
    double dmin = DBL_MAX;
    for (int dj = -P; dj <= P; ++dj) {
      const int h = plan.hwf[dj < 0 ? -dj : dj];
      if (h < 0) continue;
      const size_t row = static_cast<size_t>(jl + P + dj) * plan.BR;
      const unsigned int a = args.G[row + (i + P - h)];
      const unsigned int b = args.G[row + (i + P + h + 1)];
      for (unsigned int k = a + lane; k < b; k += 32) {
        const double2 v = __ldg(reinterpret_cast<const double2*>(args.rec + k));
        const double dx = qx - v.x;
        const double dy = qy - v.y;
        dmin = fmin(dmin, __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));
      }
    }
    for (int o = 16; o > 0; o >>= 1) dmin = fmin(dmin, __shfl_xor_sync(0xffffffffu, dmin, o));
    // first retry threshold whose (strict) ball is non-empty
    int level = -1;
    for (int k = 0; k < plan.n_thr; ++k) {
      if (dmin < plan.thr[k]) {
        level = k;
        break;
      }
    }
    if (level < 0) continue;  // stays untouched (NaN or the previous elevation)
    const double thr = plan.thr[level];
    double num = 0.0, den = 0.0;
    int cnt = 0;
    bool coincident = false;
    for (int dj = -P; dj <= P; ++dj) {
      const int h = plan.hwf[dj < 0 ? -dj : dj];
      if (h < 0) continue;
      const size_t row = static_cast<size_t>(jl + P + dj) * plan.BR;
      const unsigned int a = args.G[row + (i + P - h)];
      const unsigned int b = args.G[row + (i + P + h + 1)];
      for (unsigned int k = a + lane; k < b; k += 32) {
        const double2 v = __ldg(reinterpret_cast<const double2*>(args.rec + k));
        const double dx = qx - v.x;
        const double dy = qy - v.y;
        const double d2 = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
        if (d2 < thr) {
          const double pz = __ldg(reinterpret_cast<const double*>(args.rec + k) + 2);
          ++cnt;
          const double w = fast_rcp(d2);
          num = fma(pz, w, num);
          den += w;
          coincident |= !(d2 > 0.0);
        }
      }
    }
    for (int o = 16; o > 0; o >>= 1) {
      num += __shfl_xor_sync(0xffffffffu, num, o);
      den += __shfl_xor_sync(0xffffffffu, den, o);
      cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    }
    coincident = __any_sync(0xffffffffu, coincident);
    if (lane == 0) {
      if (coincident) atomicExch(&args.counters[1], 1u);
      args.elevation[cell] = __double2float_rn(__ddiv_rn(num, den));
      if (args.dbg_count) {
        args.dbg_count[cell] = cnt;
        args.dbg_level[cell] = static_cast<signed char>(level);
      }
    }
  }
}

// Half-width (in bins along i) of the window that can hold a point with d2 < thr, for a bin row |dj| away.
// A point binned to cell b lies within (0.5 + slack) cells of that cell's centre along each axis, so its
// distance along j to the query centre is at least (|dj| - 0.5 - slack) cells; what is left of the threshold
// bounds the reach along i.  -1: that bin row cannot contribute.
void half_widths(double thr, double res, int W, short* out) {
  const double slack = 1e-6;
  const double r = std::sqrt(thr) / res;  // reach in cells
  for (int dj = 0; dj <= kMaxHalfWidth; ++dj) {
    if (dj > W) {
      out[dj] = -1;
      continue;
    }
    const double dy = std::max(0.0, dj - 0.5 - slack);
    const double rem2 = r * r * (1.0 + 1e-12) - dy * dy;
    if (rem2 <= 0.0) {
      out[dj] = -1;
      continue;
    }
    int h = static_cast<int>(std::floor(std::sqrt(rem2) + 0.5 + slack));
    out[dj] = static_cast<short>(std::min(h, W));
  }
}

}  // namespace

int dsm_run(amb_ctx* ctx, const double* d_xyz, size_t n, int32_t interpolation_radius, double center_easting,
            double center_northing) {
  const amb_geometry& g = ctx->geom;
  if (n == 0) return AMB_ERR_EMPTY;
  if (interpolation_radius < 1 || n >= size_t(0xffffffffu)) return AMB_ERR_INVALID_ARGUMENT;
  int st = ensure_layer(ctx, AMB_LAYER_ELEVATION);
  if (st != AMB_OK) return st;

  DsmPlan plan;
  std::memset(&plan, 0, sizeof(plan));
  const std::vector<double> thr = dsm_thresholds(interpolation_radius);
  if (thr.size() > static_cast<size_t>(kMaxThresholds)) return AMB_ERR_UNSUPPORTED;
  double thr_max = 0.0;
  for (size_t k = 0; k < thr.size(); ++k) {
    plan.thr[k] = thr[k];
    thr_max = std::max(thr_max, thr[k]);
  }
  plan.n_thr = static_cast<int>(thr.size());
  plan.thr0 = static_cast<double>(interpolation_radius);
  plan.rows = g.rows;
  plan.cols_slab = ctx->col_end - ctx->col_begin;
  plan.col_begin = ctx->col_begin;
  plan.res = g.resolution;
  plan.inv_res = 1.0 / g.resolution;
  plan.base_x = g.pos_x + (0.5 * g.length_x - 0.5 * g.resolution);
  plan.base_y = g.pos_y + (0.5 * g.length_y - 0.5 * g.resolution);
  plan.shift_x = center_northing;  // dsm.cc:42 (sic)
  plan.shift_y = center_easting;   // dsm.cc:43
  const double slack = 1e-6;
  plan.W = static_cast<int>(std::floor(std::sqrt(plan.thr0) / g.resolution + 0.5 + slack));
  plan.P = static_cast<int>(std::floor(std::sqrt(thr_max) / g.resolution + 0.5 + slack));
  plan.P = std::max(plan.P, plan.W);
  if (plan.P > kMaxHalfWidth) return AMB_ERR_UNSUPPORTED;  // resolution far finer than the search radius
  plan.BR = plan.rows + 2 * plan.P;
  plan.BC = plan.cols_slab + 2 * plan.P;
  half_widths(plan.thr0, g.resolution, plan.W, plan.hw);
  half_widths(thr_max, g.resolution, plan.P, plan.hwf);

  const size_t nb = static_cast<size_t>(plan.BR) * plan.BC;
  const size_t g_elems = ((nb + 2 + 3) / 4) * 4;  // uint4-aligned length
  const size_t n_vec = g_elems / 4;
  const int scan_blocks = static_cast<int>((n_vec + kScanChunk / 4 - 1) / (kScanChunk / 4));
  const size_t cells = ctx->slab_cells();

  AMB_CUDA(ctx, ctx->bin_starts.reserve(g_elems * sizeof(unsigned int)));
  AMB_CUDA(ctx, ctx->block_sums.reserve(static_cast<size_t>(scan_blocks) * sizeof(unsigned int)));
  AMB_CUDA(ctx, ctx->records.reserve(n * sizeof(PointRec)));
  AMB_CUDA(ctx, ctx->empty_cells.reserve(cells * sizeof(unsigned int)));
  AMB_CUDA(ctx, ctx->counters.reserve(64));
  if (ctx->dsm_debug) {
    AMB_CUDA(ctx, ctx->dbg_count.reserve(cells * sizeof(int)));
    AMB_CUDA(ctx, ctx->dbg_level.reserve(cells));
  }
  cudaStream_t s = ctx->stream;
  unsigned int* G = ctx->bin_starts.as<unsigned int>();
  unsigned int* counters = ctx->counters.as<unsigned int>();
  PointRec* rec = ctx->records.as<PointRec>();

  AMB_CUDA(ctx, cudaMemsetAsync(G, 0, g_elems * sizeof(unsigned int), s));
  AMB_CUDA(ctx, cudaMemsetAsync(counters, 0, 64, s));
  if (ctx->dsm_debug) {
    AMB_CUDA(ctx, cudaMemsetAsync(ctx->dbg_count.ptr, 0xff, cells * sizeof(int), s));
    AMB_CUDA(ctx, cudaMemsetAsync(ctx->dbg_level.ptr, 0xff, cells, s));
  }
  ctx->dsm_launches = 0;

  const int stream_grid = kNumSMsB200 * 8;
  dsm_count_kernel<<<stream_grid, 256, 0, s>>>(d_xyz, n, plan, G, counters);
  scan_reduce_kernel<<<scan_blocks, 256, 0, s>>>(reinterpret_cast<const uint4*>(G), n_vec,
                                                 ctx->block_sums.as<unsigned int>());
  scan_spine_kernel<<<1, 1024, 0, s>>>(ctx->block_sums.as<unsigned int>(), scan_blocks);
  scan_apply_kernel<<<scan_blocks, 256, 0, s>>>(reinterpret_cast<uint4*>(G), n_vec,
                                                ctx->block_sums.as<unsigned int>());
  dsm_scatter_kernel<<<stream_grid, 256, 0, s>>>(d_xyz, n, plan, G, rec);
  dsm_canon_kernel<<<static_cast<unsigned int>((nb + 255) / 256), 256, 0, s>>>(G, nb, rec);
  ctx->dsm_launches += 6;
  AMB_CUDA(ctx, cudaEventRecord(ctx->events[EV_DSM_BIN_END], s));

  // gather
  const int W = plan.W;
  const int NC = TJ + 2 * W, NI = TI + 2 * W + 1;
  size_t off_bytes = (static_cast<size_t>(NC) * NI + 2 * NC + 1 + kStrip + 2 * W) * sizeof(unsigned int);
  off_bytes = (off_bytes + 15) & ~static_cast<size_t>(15);
  // Stage size: 1.6x the expected number of points in a tile + apron under a uniform density (the tail of a
  // Poisson count is far inside that; denser tiles fall back to global reads), clamped to [24 KB, 100 KB] so that
  // several blocks stay resident per SM (227 KB usable).
  if (off_bytes + 3 * sizeof(double) * 64 > 200 * 1024) return AMB_ERR_UNSUPPORTED;
  const double avg_per_bin = static_cast<double>(n) / static_cast<double>(nb);
  const double expect = avg_per_bin * static_cast<double>(NC) * static_cast<double>(NI - 1);
  size_t smem = off_bytes + static_cast<size_t>(1.6 * expect + 64.0) * 3 * sizeof(double);
  smem = std::min<size_t>(std::max<size_t>(smem, 24 * 1024), 100 * 1024);
  smem = std::max(smem, off_bytes + 3 * sizeof(double) * 64);
  smem = (smem + 1023) & ~static_cast<size_t>(1023);
  int capacity = static_cast<int>((smem - off_bytes) / (3 * sizeof(double)));
  GatherArgs ga;
  ga.G = G;
  ga.rec = rec;
  ga.elevation = ctx->layers[AMB_LAYER_ELEVATION];
  ga.empty_list = ctx->empty_cells.as<unsigned int>();
  ga.counters = counters;
  ga.dbg_count = ctx->dsm_debug ? ctx->dbg_count.as<int>() : nullptr;
  ga.dbg_level = ctx->dsm_debug ? ctx->dbg_level.as<signed char>() : nullptr;
  ga.capacity = capacity;
  ga.tiles_i = (plan.rows + TI - 1) / TI;
  const int tiles_j = (plan.cols_slab + TJ - 1) / TJ;
  AMB_CUDA(ctx, cudaFuncSetAttribute(dsm_gather_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     static_cast<int>(smem)));
  dsm_gather_kernel<<<ga.tiles_i * tiles_j, kGatherThreads, smem, s>>>(plan, ga);
  ctx->dsm_launches += 1;
  AMB_CUDA(ctx, cudaEventRecord(ctx->events[EV_DSM_GATHER_END], s));

  FillArgs fa;
  fa.G = G;
  fa.rec = rec;
  fa.elevation = ga.elevation;
  fa.empty_list = ga.empty_list;
  fa.counters = counters;
  fa.dbg_count = ga.dbg_count;
  fa.dbg_level = ga.dbg_level;
  dsm_fill_kernel<<<kNumSMsB200 * 8, 256, 0, s>>>(plan, fa);
  ctx->dsm_launches += 1;
  AMB_CUDA(ctx, cudaEventRecord(ctx->events[EV_DSM_FILL_END], s));
  AMB_CUDA(ctx, cudaGetLastError());
  ctx->dsm_debug_valid = ctx->dsm_debug;
  return AMB_OK;
}

}  // namespace amb
