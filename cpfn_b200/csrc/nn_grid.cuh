// Exact 3-nearest-neighbour search of a small "known" cloud (<= kNnGridMax points: the sampled centroids of a
// feature-propagation layer) through a uniform grid built in shared memory by every CTA.
//
// Semantics are the reference's (interpolate_gpu.cu:22-58): squared distances with its rounding sequence, the three
// smallest in ascending order, equal distances resolved in favour of the smaller index -- the sequential scan with
// strict '<' keeps exactly the three smallest (distance, index) pairs in lexicographic order, which is what the
// order-independent insertion below keeps.  The grid only decides WHICH candidates are looked at: cells are visited
// in growing cubes around the query's cell until the third-best distance is strictly smaller than the distance from
// the query to the nearest face of the visited cube that has cells behind it (with a relative margin that covers the
// rounding of both sides), so no unvisited point can enter or tie the result.  Worst case: the whole cloud.
#pragma once

#include <math.h>

#include "common.cuh"

namespace cpfn {

constexpr int kNnGridMax = 2048;      // known points staged + sorted in shared memory
constexpr int kNnGridMin = 384;       // below this the exhaustive scan (broadcast shared-memory loads, no divergence) wins:
                                      // measured 11 us (scan) vs 35 us (grid) at m = 128, 60 vs 55 us at m = 512
constexpr int kNnGridMaxG = 12;       // cells per axis <= 12 -> <= 1728 cells

struct NnGrid {
  float minx, miny, minz, cell, inv_cell;
  int G;
};

__host__ __device__ inline int nn_grid_cells_per_axis(int m) {
  int g = 2;
  while (g < kNnGridMaxG && g * g * g < m) ++g;
  return g;
}

// shared memory the grid needs: sorted records (float4: x, y, z, index bits), cell starts, fill cursors, cell ids
__host__ __device__ inline size_t nn_grid_smem_bytes(int m) {
  const int g = nn_grid_cells_per_axis(m);
  const size_t ncell = static_cast<size_t>(g) * g * g;
  return sizeof(float4) * static_cast<size_t>(m) + sizeof(int) * (2 * ncell + 2) + sizeof(int) * static_cast<size_t>(m);
}

__device__ __forceinline__ int nn_cell_coord(float v, float lo, float inv_cell, int G) {
  const int c = static_cast<int>(floorf((v - lo) * inv_cell));
  return c < 0 ? 0 : (c >= G ? G - 1 : c);
}

// Block-cooperative build (all threads of the CTA, blockDim.x >= 32).  `raw` = dynamic shared memory of
// nn_grid_smem_bytes(m) bytes, 16-byte aligned.
__device__ inline void nn_grid_build(const float *__restrict__ known, int m, unsigned char *raw, NnGrid &grid,
                                     float4 *&recs, int *&cell_start) {
  __shared__ NnGrid s_grid;
  const int G = nn_grid_cells_per_axis(m);
  const int ncell = G * G * G;
  recs = reinterpret_cast<float4 *>(raw);
  cell_start = reinterpret_cast<int *>(recs + m);           // [ncell + 1]
  int *fill = cell_start + ncell + 1;                       // [ncell + 1]
  int *cell_of = fill + ncell + 1;                          // [m]
  const int t = threadIdx.x, nt = blockDim.x;
  if (t < 32) {                                             // bounding box: one warp, m is small
    float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (int k = t; k < m; k += 32) {
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        const float v = __ldg(known + static_cast<size_t>(k) * 3 + a);
        lo[a] = fminf(lo[a], v); hi[a] = fmaxf(hi[a], v);
      }
    }
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        lo[a] = fminf(lo[a], __shfl_xor_sync(0xffffffffu, lo[a], o));
        hi[a] = fmaxf(hi[a], __shfl_xor_sync(0xffffffffu, hi[a], o));
      }
    if (t == 0) {
      const float ext = fmaxf(fmaxf(hi[0] - lo[0], hi[1] - lo[1]), hi[2] - lo[2]);
      const float cell = fmaxf(ext / static_cast<float>(G), 1e-20f) * 1.0001f;     // every point falls inside G cells
      s_grid.minx = lo[0]; s_grid.miny = lo[1]; s_grid.minz = lo[2];
      s_grid.cell = cell; s_grid.inv_cell = 1.0f / cell; s_grid.G = G;
    }
  }
  for (int c = t; c <= ncell; c += nt) cell_start[c] = 0;
  __syncthreads();
  grid = s_grid;
  for (int k = t; k < m; k += nt) {                         // counts, shifted by one cell
    const float x = __ldg(known + static_cast<size_t>(k) * 3), y = __ldg(known + static_cast<size_t>(k) * 3 + 1),
                z = __ldg(known + static_cast<size_t>(k) * 3 + 2);
    const int c = (nn_cell_coord(z, grid.minz, grid.inv_cell, G) * G + nn_cell_coord(y, grid.miny, grid.inv_cell, G)) * G +
                  nn_cell_coord(x, grid.minx, grid.inv_cell, G);
    cell_of[k] = c;
    atomicAdd(&cell_start[c + 1], 1);
  }
  __syncthreads();
  if (t < 32) {                                             // inclusive scan: cell_start[c] = first record of cell c
    int carry = 0;
    for (int c0 = 0; c0 <= ncell; c0 += 32) {
      const int c = c0 + t;
      int inc = c <= ncell ? cell_start[c] : 0;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int u = __shfl_up_sync(0xffffffffu, inc, o);
        if (t >= o) inc += u;
      }
      if (c <= ncell) { cell_start[c] = carry + inc; fill[c] = carry + inc; }
      carry += __shfl_sync(0xffffffffu, inc, 31);
    }
  }
  __syncthreads();
  for (int k = t; k < m; k += nt) {                         // scatter (order inside a cell is irrelevant)
    const int pos = atomicAdd(&fill[cell_of[k]], 1);
    recs[pos] = make_float4(__ldg(known + static_cast<size_t>(k) * 3), __ldg(known + static_cast<size_t>(k) * 3 + 1),
                            __ldg(known + static_cast<size_t>(k) * 3 + 2), __int_as_float(k));
  }
  __syncthreads();
}

struct Nn3 {
  float d1, d2, d3;
  int i1, i2, i3;
};

__device__ __forceinline__ bool nn_less(float d, int k, float b, int i) { return d < b || (d == b && k < i); }

__device__ __forceinline__ void nn_insert(Nn3 &r, float d, int k) {
  if (!nn_less(d, k, r.d3, r.i3)) return;
  if (nn_less(d, k, r.d1, r.i1)) {
    r.d3 = r.d2; r.i3 = r.i2; r.d2 = r.d1; r.i2 = r.i1; r.d1 = d; r.i1 = k;
  } else if (nn_less(d, k, r.d2, r.i2)) {
    r.d3 = r.d2; r.i3 = r.i2; r.d2 = d; r.i2 = k;
  } else {
    r.d3 = d; r.i3 = k;
  }
}

// One thread, one query.
__device__ inline Nn3 nn_grid_query(float ux, float uy, float uz, const NnGrid &g, const float4 *__restrict__ recs,
                                    const int *__restrict__ cell_start) {
  Nn3 r;
  r.d1 = r.d2 = r.d3 = INFINITY;
  r.i1 = r.i2 = r.i3 = 0x7FFFFFFF;                      // any real index beats the placeholder on an inf tie
  const int G = g.G;
  const int cx = nn_cell_coord(ux, g.minx, g.inv_cell, G), cy = nn_cell_coord(uy, g.miny, g.inv_cell, G),
            cz = nn_cell_coord(uz, g.minz, g.inv_cell, G);
  for (int ring = 0; ring < G; ++ring) {
    const int x0 = max(cx - ring, 0), x1 = min(cx + ring, G - 1), y0 = max(cy - ring, 0), y1 = min(cy + ring, G - 1),
              z0 = max(cz - ring, 0), z1 = min(cz + ring, G - 1);
    for (int z = z0; z <= z1; ++z)
      for (int y = y0; y <= y1; ++y) {
        // only the shell of the cube is new: full x-run on the z / y faces, the two end cells otherwise
        const bool face = (z == cz - ring || z == cz + ring || y == cy - ring || y == cy + ring);
        const int row = (z * G + y) * G;
        if (face || ring == 0) {
          for (int k = cell_start[row + x0]; k < cell_start[row + x1 + 1]; ++k) {
            const float4 c = recs[k];
            nn_insert(r, sqdist3(ux, uy, uz, c.x, c.y, c.z), __float_as_int(c.w));
          }
        } else {
          if (cx - ring >= 0)
            for (int k = cell_start[row + cx - ring]; k < cell_start[row + cx - ring + 1]; ++k) {
              const float4 c = recs[k];
              nn_insert(r, sqdist3(ux, uy, uz, c.x, c.y, c.z), __float_as_int(c.w));
            }
          if (cx + ring <= G - 1)
            for (int k = cell_start[row + cx + ring]; k < cell_start[row + cx + ring + 1]; ++k) {
              const float4 c = recs[k];
              nn_insert(r, sqdist3(ux, uy, uz, c.x, c.y, c.z), __float_as_int(c.w));
            }
        }
      }
    if (x0 == 0 && x1 == G - 1 && y0 == 0 && y1 == G - 1 && z0 == 0 && z1 == G - 1) break;   // everything visited
    // distance to the nearest face of the visited cube that has unvisited cells behind it
    float dout = INFINITY;
    if (cx - ring > 0) dout = fminf(dout, ux - (g.minx + static_cast<float>(cx - ring) * g.cell));
    if (cx + ring < G - 1) dout = fminf(dout, (g.minx + static_cast<float>(cx + ring + 1) * g.cell) - ux);
    if (cy - ring > 0) dout = fminf(dout, uy - (g.miny + static_cast<float>(cy - ring) * g.cell));
    if (cy + ring < G - 1) dout = fminf(dout, (g.miny + static_cast<float>(cy + ring + 1) * g.cell) - uy);
    if (cz - ring > 0) dout = fminf(dout, uz - (g.minz + static_cast<float>(cz - ring) * g.cell));
    if (cz + ring < G - 1) dout = fminf(dout, (g.minz + static_cast<float>(cz + ring + 1) * g.cell) - uz);
    // conservative: the bound shrinks by 1e-4 relative (cell arithmetic and the distance are both rounded)
    const float safe = dout > 0.f ? dout * 0.9999f : 0.f;
    if (r.d3 < safe * safe) break;
  }
  if (r.i1 == 0x7FFFFFFF) r.i1 = 0;                       // fewer than three known points: index 0, distance inf
  if (r.i2 == 0x7FFFFFFF) r.i2 = 0;
  if (r.i3 == 0x7FFFFFFF) r.i3 = 0;
  return r;
}

}  // namespace cpfn
