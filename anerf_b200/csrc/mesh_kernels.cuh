// marching cubes on the density volume that is already resident in HBM (reference:
// mcubes.marching_cubes(sigma, threshold) on a host copy, run_render.py:983-986).  Two passes over the cells --
// count triangles per cell, then (after an exclusive scan) emit them -- both HBM-bound: one thread per cell, cells
// ordered so that a warp reads consecutive voxels of the volume's fastest axis.
//   vertex on the cube edge (a, b): p_a + (iso - f_a) / (f_b - f_a) (p_b - p_a)   (index coordinates, like mcubes)
//   a corner is inside when f > iso
// The case table is generated (anerf_b200/mc_table.py -> mc_table.inc).  Each emitted vertex carries the id of its
// volume edge (lower corner * 3 + axis) so that the caller can weld duplicates into an indexed mesh.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace anerf {
namespace mesh {

#include "mc_table.inc"

struct McArgs {
  const float* vol;
  int n0, n1, n2;                 // voxels per axis
  long long s0, s1, s2;           // element strides of the volume (any layout; the cell order follows axis 2)
  float iso;
  int* counts;                    // [cells]     pass 1 out
  const long long* offsets;       // [cells]     pass 2 in: exclusive scan of counts
  float* verts;                   // [T,3,3]     pass 2 out
  long long* keys;                // [T,3]       pass 2 out
  long long n_cells;
};

#ifdef __CUDACC__
__constant__ signed char c_tri_count[256];
__constant__ signed char c_tri_table[256][ANERF_MC_MAX_TRIS * 3];
__constant__ signed char c_edge_corner[12][2];
__constant__ signed char c_edge_axis[12];

__device__ __forceinline__ int mc_case(const McArgs& a, int i, int j, int k, float (&f)[8]) {
  int cs = 0;
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    f[c] = __ldg(a.vol + (long long)(i + (c & 1)) * a.s0 + (long long)(j + ((c >> 1) & 1)) * a.s1 + (long long)(k + ((c >> 2) & 1)) * a.s2);
    cs |= (f[c] > a.iso) ? (1 << c) : 0;
  }
  return cs;
}

__global__ void mc_count_kernel(const McArgs a) {
  const long long cell = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (cell >= a.n_cells) return;
  const int m2 = a.n2 - 1, m1 = a.n1 - 1;
  const int k = (int)(cell % m2), j = (int)((cell / m2) % m1), i = (int)(cell / ((long long)m2 * m1));
  float f[8];
  a.counts[cell] = c_tri_count[mc_case(a, i, j, k, f)];
}

__global__ void mc_emit_kernel(const McArgs a) {
  const long long cell = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (cell >= a.n_cells) return;
  const int m2 = a.n2 - 1, m1 = a.n1 - 1;
  const int k = (int)(cell % m2), j = (int)((cell / m2) % m1), i = (int)(cell / ((long long)m2 * m1));
  float f[8];
  const int cs = mc_case(a, i, j, k, f);
  const int nt = c_tri_count[cs];
  if (nt == 0) return;
  long long t0 = a.offsets[cell];
  for (int t = 0; t < nt; ++t) {
#pragma unroll
    for (int v = 0; v < 3; ++v) {
      const int e = c_tri_table[cs][3 * t + v];
      const int ca = c_edge_corner[e][0], cb = c_edge_corner[e][1], ax = c_edge_axis[e];
      const float fa = f[ca], fb = f[cb];
      const float u = (a.iso - fa) / (fb - fa);
      const int pi = i + (ca & 1), pj = j + ((ca >> 1) & 1), pk = k + ((ca >> 2) & 1);
      float* o = a.verts + ((t0 + t) * 3 + v) * 3;
      o[0] = (float)pi + (ax == 0 ? u : 0.f);
      o[1] = (float)pj + (ax == 1 ? u : 0.f);
      o[2] = (float)pk + (ax == 2 ? u : 0.f);
      a.keys[(t0 + t) * 3 + v] = (((long long)pi * a.n1 + pj) * a.n2 + pk) * 3 + ax;
    }
  }
}
#endif

}  // namespace mesh
}  // namespace anerf
