// Probe for round 2 (half-tile forward kernel): where do the accumulator rows of a cta_group::2 MMA land in TMEM
// for M = 128 (64 rows per CTA) and M = 256 (128 rows per CTA), and can the accumulator be placed at a TMEM lane
// offset (second half-tile in lanes 64..127)?
//
// One cluster of 2 CTAs.  A(m, k): A[m][0] = m, A[m][1] = 1, rest 0;  B(n, k): B[n][0] = 256, B[n][1] = n, rest 0
// (all exact in fp16)  =>  D[m][n] = 256 m + n exactly.  Before the MMA every TMEM cell of the first 128 columns is
// set to -1, afterwards each CTA dumps lanes 0..127 x columns 0..127 to global memory; the host decodes (m, n) per cell.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -shared -Xcompiler -fPIC -I../../anerf_b200/csrc -o tmem_layout_probe.so tmem_layout_probe.cu
#include "tc_sm100.cuh"

using namespace anerf;

__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
        "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]),
        "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]),
        "r"(v[30]), "r"(v[31])
      : "memory");
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

// M: 128 or 256 (rows of the pair's MMA); N: 64; lane_off: TMEM lane offset of the accumulator (0 or 64); col_off: column offset
__global__ void __launch_bounds__(160, 1) tmem_layout_probe_kernel(int M, int N, int lane_off, int col_off, float* out /*[2][128][128]*/,
                                                                    DeviceStatus* status) {
  extern __shared__ __align__(1024) uint8_t smem[];
  // A operand: rows_per_cta x K=16 fp16, K-major no-swizzle core matrices [k/8][row/8][8 rows][8 elements]
  // B operand: (N/2) x K=16, same layout
  const int rows = M / 2, nh = N / 2;
  __half* A = reinterpret_cast<__half*>(smem);
  __half* B = reinterpret_cast<__half*>(smem + 8192);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 16384);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + 16384 + 64);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = cluster_ctarank();
  for (int i = tid; i < 4096; i += blockDim.x) { A[i] = __float2half(0.f); B[i] = __float2half(0.f); }
  __syncthreads();
  for (int r = tid; r < rows; r += blockDim.x) {
    const int m = (int)rank * rows + r;
    const int off = (r >> 3) * 64 + (r & 7) * 8;                 // k group 0 of this row (elements)
    A[off + 0] = __float2half((float)m);
    A[off + 1] = __float2half(1.f);
  }
  for (int r = tid; r < nh; r += blockDim.x) {
    const int n = (int)rank * nh + r;
    const int off = (r >> 3) * 64 + (r & 7) * 8;
    B[off + 0] = __float2half(256.f);
    B[off + 1] = __float2half((float)n);
  }
  if (tid == 0) { mbar_init(bar, 1); fence_mbar_init(); }
  if (warp == 4) tmem_alloc(tmem_slot, 512);
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after_sync();
  const uint32_t tb = *tmem_slot;
  if (warp < 4) {                      // sentinel -1 in lanes 32*warp .. +31, columns 0..127
    uint32_t v[32];
    for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(-1.f);
    for (int c = 0; c < 128; c += 32) tmem_st32(tb + ((uint32_t)(warp * 32) << 16) + c, v);
  }
  tc_fence_before_sync();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after_sync();
  if (warp == 4 && rank == 0) {
    if (elect_one()) {
      const uint32_t id = make_idesc_f16(0u, 0u, (uint32_t)M, (uint32_t)N);
      // K-major, no swizzle: LBO = distance between the two k groups of the K=16 slab, SBO = between 8-row groups
      const uint64_t da = smem_desc(smem_u32(A), (uint32_t)(rows / 8) * 128u, 128u);
      const uint64_t db = smem_desc(smem_u32(B), (uint32_t)(nh / 8) * 128u, 128u);
      umma_f16(tb + ((uint32_t)lane_off << 16) + (uint32_t)col_off, da, db, id, 0u);
      umma_commit(bar);
    }
    __syncwarp();
  }
  mbar_wait(bar, 0, status, 900);
  tc_fence_after_sync();
  if (warp < 4) {
    uint32_t v[32];
    for (int c = 0; c < 128; c += 32) {
      tmem_ld32(tb + ((uint32_t)(warp * 32) << 16) + c, v);
      tmem_ld_wait();
      for (int i = 0; i < 32; ++i) out[((size_t)rank * 128 + warp * 32 + lane) * 128 + c + i] = __uint_as_float(v[i]);
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  cluster_sync_all();
  if (warp == 4) tmem_dealloc(tb, 512);
}

extern "C" int tmem_layout_probe(int M, int N, int lane_off, int col_off, float* out_dev) {
  DeviceStatus* st_h = nullptr; DeviceStatus* st_d = nullptr;
  cudaHostAlloc((void**)&st_h, sizeof(DeviceStatus), cudaHostAllocMapped);
  memset(st_h, 0, sizeof(DeviceStatus));
  cudaHostGetDevicePointer((void**)&st_d, st_h, 0);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(2); cfg.blockDim = dim3(160); cfg.dynamicSmemBytes = 16384 + 128;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, tmem_layout_probe_kernel, M, N, lane_off, col_off, out_dev, st_d);
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  int rc = (e == cudaSuccess && st_h->code == 0) ? 0 : (int)e + 1000 * (int)st_h->code;
  cudaFreeHost(st_h);
  return rc;
}
