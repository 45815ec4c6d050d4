// Microbenchmark 2: tensor-pipe cost of one tcgen05.mma (kind::f16, bf16, M=128, K=16) as a function of N when the
// issue path is NOT the limiter: the whole warp runs the loop, one elected lane issues 8 back-to-back MMAs per
// iteration with descriptors that ptxas keeps in uniform registers (bare UTCHMMA, no R2UR waterfall).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../frankenstein_b200/csrc umma_rate2.cu -o umma_rate2
#include <cstdio>
#include "common.cuh"
void fk_set_last_error(const char*, const char*, int) {}
void fk_count_launch(int) {}
using namespace fk;

template <int N, bool TS, int NISS>
__global__ void __launch_bounds__(128, 1) rate_kernel(int iters, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar[2];
  __shared__ uint32_t tslot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (warp == 0 && lane == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); fence_mbar_init(); }
  if (warp == 1) { tmem_alloc(&tslot, 512); tmem_relinquish(); }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = __shfl_sync(0xffffffffu, tslot, 0);
  const int w = __shfl_sync(0xffffffffu, warp, 0);
  if (w == 0 || (NISS == 2 && w == 2)) {
    constexpr uint32_t idesc = umma_idesc_bf16(128, N);
    const uint32_t a = smem_u32(smem), b = smem_u32(smem + 32768);
    const uint32_t d = tb + (w == 0 ? 0u : 256u);         // each issuer its own accumulator
    const uint64_t ad = umma_desc_sw128(a), bd = umma_desc_sw128(b);
    const long long t0 = clock64();
    for (int i = 0; i < iters / 8 / NISS; ++i) {
      if (elect_one()) {
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
          if (TS) umma_bf16_ts(d, tb + 448 + (kk & 3) * 8, bd + 2 * (kk & 3), idesc, 1u);
          else umma_bf16(d, ad + 2 * (kk & 3), bd + 2 * (kk & 3), idesc, 1u);
        }
      }
      __syncwarp();
    }
    if (elect_one()) umma_commit(&bar[w == 0 ? 0 : 1]);
    __syncwarp();
    mbar_wait(&bar[w == 0 ? 0 : 1], 0);
    const long long t1 = clock64();
    if (blockIdx.x == 0 && w == 0 && lane == 0) out[0] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tb, 512); }
}

template <int N, bool TS, int NISS>
void run(long long* out) {
  const int iters = 8192;
  cudaFuncSetAttribute(rate_kernel<N, TS, NISS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
  rate_kernel<N, TS, NISS><<<148, 128, 96 * 1024>>>(iters, out);
  long long c = 0;
  cudaError_t e = cudaMemcpy(&c, out, 8, cudaMemcpyDeviceToHost);
  if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return; }
  printf("%d,%s,%d,%.1f,%d\n", N, TS ? "TS" : "SS", NISS, double(c) / iters, N / 2);
}

int main() {
  long long* out;
  cudaMalloc(&out, 8);
  printf("N,mode,issuers,cycles_per_mma,ideal\n");
  run<16, false, 1>(out); run<32, false, 1>(out); run<64, false, 1>(out); run<128, false, 1>(out); run<256, false, 1>(out);
  run<16, true, 1>(out); run<32, true, 1>(out); run<64, true, 1>(out); run<128, true, 1>(out); run<256, true, 1>(out);
  run<32, false, 2>(out); run<64, false, 2>(out); run<128, false, 2>(out);
  run<32, true, 2>(out); run<64, true, 2>(out); run<128, true, 2>(out);
  return 0;
}
