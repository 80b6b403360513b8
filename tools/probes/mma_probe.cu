// Developer probe (not part of the library): issue cost of tcgen05.mma kind::tf32, M = 128, K = 8, as a function of N,
// of where A comes from (TMEM / shared memory) and of how many independent accumulators the instructions rotate over.
// build: nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -I xeofs_b200/csrc tools/probes/mma_probe.cu -o gpurun_out/mma_probe
#include <cstdio>
#include "tc_common.cuh"
using namespace xb;

namespace xb { __host__ __device__ void set_error(const char*, ...) {} }

__global__ void __launch_bounds__(128, 1) probe(int n, int nacc, int a_smem, int iters, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 64 * 1024 / 4; i += blockDim.x) ((float*)smem)[i] = 1.0f;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  fence_proxy_async_smem();
  if (warp == 0) tmem_alloc(&slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = slot;
  if (warp == 0) {
    const uint32_t idesc = make_idesc(n);
    const uint64_t bdesc = make_b_desc(smem_u32(smem));
    const uint64_t adesc = make_b_desc(smem_u32(smem) + 32768);
    long long t0 = 0, t1 = 0;
    if (elect_one()) {
      t0 = clock64();
      for (int i = 0; i < iters; ++i) {
        const uint32_t d = tm + (uint32_t)(i % nacc) * (uint32_t)((n + 31) / 32 * 32);
        if (a_smem) mma_tf32_ss(d, adesc, bdesc, idesc, i >= nacc);
        else mma_tf32_ts(d, tm + 448 + (i & 3) * 8, bdesc, idesc, i >= nacc);
      }
      mma_commit(&bar);
    }
    __syncwarp();
    mbar_wait(&bar, 0);
    t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tm, 512);
}

int main() {
  long long* out;
  cudaMallocManaged(&out, 8);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 70 * 1024);
  const int iters = 2048;
  printf("tcgen05.mma kind::tf32 M=128 K=8: cycles per instruction (issue loop of %d, one CTA per SM on all SMs)\n", iters);
  for (int a_smem = 0; a_smem < 2; ++a_smem)
    for (int n : {32, 64, 112, 128, 224, 256})
      for (int nacc : {1, 2, 4}) {
        if ((n + 31) / 32 * 32 * nacc > 448) continue;
        probe<<<148, 128, 70 * 1024>>>(n, nacc, a_smem, iters, out);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
        printf("A from %s  N=%3d  accumulators=%d : %6.1f cycles/MMA  (ideal at 1956 MAC/clk: %5.1f)\n", a_smem ? "smem" : "TMEM", n, nacc,
               (double)out[0] / iters, 128.0 * n * 8 / 1956.0);
      }
  return 0;
}
