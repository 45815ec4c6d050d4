// Microbenchmark: issue rate of tcgen05.mma (kind::f16, bf16, M=128) as a function of N, operand source (SS / TS)
// and accumulator rotation.  One CTA per SM, operands = zero-filled smem, `iters` MMAs issued back to back by one
// thread, one commit at the end; cycles measured with clock64 around issue+completion.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../frankenstein_b200/csrc umma_rate.cu -o umma_rate
#include <cstdio>
#include <cstdlib>
#include "common.cuh"
void fk_set_last_error(const char*, const char*, int) {}
void fk_count_launch(int) {}
using namespace fk;

__global__ void __launch_bounds__(128, 1) rate_kernel(int N, int iters, int ts_mode, int nacc, int swz64, int style, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar, bar2;
  __shared__ uint32_t tslot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (warp == 0 && lane == 0) { mbar_init(&bar, 1); mbar_init(&bar2, 1); fence_mbar_init(); }
  if (warp == 1) { tmem_alloc(&tslot, 512); tmem_relinquish(); }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = __shfl_sync(0xffffffffu, tslot, 0);
  if (style == 2) {
    // two independent issuer threads (warps 0 and 2), half of the MMAs each, different accumulators
    if ((warp == 0 || warp == 2) && lane == 0) {
      const uint32_t idesc = umma_idesc_bf16(128, N);
      const uint32_t a = smem_u32(smem), b = smem_u32(smem + 32768);
      uint64_t* mybar = warp == 0 ? &bar : &bar2;
      const long long t0 = clock64();
      for (int i = 0; i < iters / 2; ++i) {
        const int kk = i & 3;
        const uint32_t d = tb + (warp == 0 ? 0 : 224) ;
        const uint64_t bd = swz64 ? umma_desc_sw64(b + (kk & 1) * 32) : umma_desc_sw128(b + kk * 32);
        if (ts_mode) umma_bf16_ts(d, tb + 448 + kk * 8, bd, idesc, i >= 1);
        else umma_bf16(d, swz64 ? umma_desc_sw64(a + (kk & 1) * 32) : umma_desc_sw128(a + kk * 32), bd, idesc, i >= 1);
      }
      umma_commit(mybar);
      mbar_wait(mybar, 0);
      const long long t1 = clock64();
      if (blockIdx.x == 0 && warp == 0) out[0] = t1 - t0;
    }
  } else if (warp == 0 && style == 1) {
    // CUTLASS style: the whole warp runs the loop (uniform control flow, operands can live in uniform registers),
    // only the instruction itself is predicated on one elected lane
    const uint32_t idesc = umma_idesc_bf16(128, N);
    const uint32_t a = smem_u32(smem), b = smem_u32(smem + 32768);
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      const int kk = i & 3;
      const uint32_t d = tb + ((nacc == 2 && (i & 1)) ? 224u : 0u);
      const uint64_t bd = swz64 ? umma_desc_sw64(b + (kk & 1) * 32) : umma_desc_sw128(b + kk * 32);
      const uint64_t ad = swz64 ? umma_desc_sw64(a + (kk & 1) * 32) : umma_desc_sw128(a + kk * 32);
      if (elect_one()) {
        if (ts_mode) umma_bf16_ts(d, tb + 448 + kk * 8, bd, idesc, i >= 2);
        else umma_bf16(d, ad, bd, idesc, i >= 2);
      }
      __syncwarp();
    }
    if (elect_one()) umma_commit(&bar);
    __syncwarp();
    mbar_wait(&bar, 0);
    const long long t1 = clock64();
    if (blockIdx.x == 0 && lane == 0) out[0] = t1 - t0;
  } else if (warp == 0 && lane == 0) {
    const uint32_t idesc = umma_idesc_bf16(128, N);
    const uint32_t a = smem_u32(smem), b = smem_u32(smem + 32768);
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      const int kk = i & 3;
      const uint32_t d = tb + ((nacc == 2 && (i & 1)) ? 224u : 0u);
      const uint64_t bd = swz64 ? umma_desc_sw64(b + (kk & 1) * 32) : umma_desc_sw128(b + kk * 32);
      if (ts_mode) umma_bf16_ts(d, tb + 448 + kk * 8, bd, idesc, i >= 2);
      else umma_bf16(d, swz64 ? umma_desc_sw64(a + (kk & 1) * 32) : umma_desc_sw128(a + kk * 32), bd, idesc, i >= 2);
    }
    umma_commit(&bar);
    mbar_wait(&bar, 0);
    const long long t1 = clock64();
    if (blockIdx.x == 0) out[0] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tb, 512); }
}

int main() {
  long long* out;
  cudaMalloc(&out, 8);
  cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
  const int iters = 4096;
  printf("style,N,mode,swizzle,nacc,cycles_per_mma,ideal\n");
  for (int style = 0; style < 3; ++style)
  for (int swz64 = 0; swz64 < 2; ++swz64)
    for (int ts = 0; ts < 2; ++ts)
      for (int N : {32, 64, 128, 256})
        for (int nacc : {1, 2}) {
          if (nacc == 2 && N > 128) continue;
          if (style == 2 && (nacc != 1 || N > 128)) continue;
          rate_kernel<<<148, 128, 96 * 1024>>>(N, iters, ts, nacc, swz64, style, out);
          long long c = 0;
          cudaError_t e = cudaMemcpy(&c, out, 8, cudaMemcpyDeviceToHost);
          if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
          printf("%s,%d,%s,%s,%d,%.1f,%d\n", style == 2 ? "two-issuers" : (style ? "warp+elect" : "lane0"), N, ts ? "TS" : "SS", swz64 ? "sw64" : "sw128", nacc, double(c) / iters, N / 2);
        }
  return 0;
}
