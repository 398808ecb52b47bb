// Microbenchmark: how long does one thread take between successive TMA tensor loads, as a function of the box size,
// and how long until the bytes have landed?  (Design input for the recurrence kernels' activation ring.)
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I autovc_b200/csrc scripts/micro/tma_issue.cu -o /tmp/tma_issue -lcuda
#include <cstdio>
#include <vector>
#include "tc_common.cuh"
namespace avc { void set_error(const char* fmt, ...) {} }
using namespace avc;

__device__ __forceinline__ unsigned long long clk() { unsigned long long t; asm volatile("mov.u64 %0, %%clock64;" : "=l"(t)); return t; }

// variant: number of TMA instructions `n`, bytes per instruction `bytes`; rows0 advance per instruction
__global__ void __launch_bounds__(64, 1) issue_kernel(const __grid_constant__ CUtensorMap map, int n, int bytes, int dim, unsigned long long* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + (base - raw) + 196608);
  const uint32_t bar0 = smem_u32(bars);
  if (threadIdx.x == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map) : "memory");
    for (int i = 0; i < 32; ++i) mbar_init(bar0 + 8u * i, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x < 32) {
    unsigned long long t_issue[33], t_land[33];
    for (int rep = 0; rep < 3; ++rep) {
      const unsigned long long t0 = clk();
      for (int i = 0; i < n; ++i) {
        if (elect_one()) {
          mbar_expect_tx(bar0 + 8u * i, bytes);
          if (dim == 3) tma_load_3d(base + i * bytes, &map, bar0 + 8u * i, (i * (bytes / 128 / 64)) % 16 * 64, 0, 0);
          else tma_load_4d(base + i * bytes, &map, bar0 + 8u * i, 0, 0, (i * (bytes / 8192)) % 16, 0);
        }
        __syncwarp();
        t_issue[i] = clk() - t0;
      }
      for (int i = 0; i < n; ++i) {
        mbar_wait(bar0 + 8u * i, rep & 1);
        t_land[i] = clk() - t0;
      }
      __syncwarp();
    }
    if (threadIdx.x == 0 && blockIdx.x == 0) {
      for (int i = 0; i < n; ++i) { out[i] = t_issue[i]; out[32 + i] = t_land[i]; }
    }
  }
}

int main() {
  const int rows = 256, K = 1024;
  __nv_bfloat16* x; cudaMalloc(&x, (size_t)rows * K * 2); cudaMemset(x, 0, (size_t)rows * K * 2);
  unsigned long long* out; cudaMalloc(&out, 64 * 8);
  PFN_encodeTiled enc = get_encode();
  struct V { const char* name; int box_rows; int groups; } vs[] = {{"16 rows x 64 (2 KB)", 16, 1}, {"64 rows x 64 (8 KB)", 64, 1}, {"128 rows x 64 (16 KB)", 128, 1},
                                                                    {"64 rows x 64 x 2 groups (16 KB)", 64, 2}, {"64 rows x 64 x 4 groups (32 KB)", 64, 4}, {"16 rows x 64 x 4 groups (8 KB)", 16, 4}};
  cudaFuncSetAttribute(issue_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 201 * 1024);
  for (int grid : {1, 64}) {
    for (auto& v : vs) {
      CUtensorMap m;
      int dim = 3;
      if (v.groups == 1) {
        make_map3(&m, x, K, rows, 1, K, (uint64_t)rows * K, 64, v.box_rows);
      } else {
        dim = 4;
        cuuint64_t dims[4] = {64, (cuuint64_t)rows, (cuuint64_t)K / 64, 1};
        cuuint64_t strides[3] = {(cuuint64_t)K * 2, 128, (cuuint64_t)rows * K * 2};
        cuuint32_t box[4] = {64, (cuuint32_t)v.box_rows, (cuuint32_t)v.groups, 1};
        cuuint32_t es[4] = {1, 1, 1, 1};
        CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, x, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); continue; }
      }
      const int bytes = v.box_rows * 128 * v.groups;
      const int n = std::min(16, 131072 / bytes);       // 128 KB in total (or 16 instructions)
      issue_kernel<<<grid, 64, 201 * 1024>>>(m, n, bytes, dim, out);
      cudaError_t e = cudaGetLastError();
      if (e == cudaSuccess) e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("%s: %s\n", v.name, cudaGetErrorString(e)); return 1; }
      unsigned long long h[64]; cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
      printf("grid %2d  %-34s n=%2d  issue:", grid, v.name, n);
      for (int i = 0; i < n; ++i) printf(" %llu", h[i]);
      printf("  | land:");
      for (int i = 0; i < n; ++i) printf(" %llu", h[32 + i]);
      printf("\n");
    }
  }
  return 0;
}
