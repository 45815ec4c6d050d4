// Microbenchmark: TMEM -> register (tcgen05.ld 32x32b.x16 / .x32) and register -> TMEM (tcgen05.st) throughput per SM
// as a function of the number of warps issuing (4, 8, 16 = 1, 2, 4 per scheduler / TMEM lane quarter).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../frankenstein_b200/csrc tmem_rate.cu -o tmem_rate
#include <cstdio>
#include "common.cuh"
void fk_set_last_error(const char*, const char*, int) {}
void fk_count_launch(int) {}
using namespace fk;

template <int X, bool STORE>
__global__ void __launch_bounds__(512, 1) k(int iters, unsigned* out, long long* cyc) {
  __shared__ uint32_t tslot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) { tmem_alloc(&tslot, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = tslot + (static_cast<uint32_t>((warp & 3) * 32) << 16);
  uint32_t r[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) r[i] = threadIdx.x + i;
  unsigned acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    const uint32_t a = tb + ((it * X) & 255) + (warp >> 2) * 0;
    if (STORE) {
      if (X == 16) tmem_st16(a, *reinterpret_cast<uint32_t(*)[16]>(&r));
      else tmem_st32(a, r);
    } else {
      if (X == 16) {
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                       "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]) : "r"(a) : "memory");
      } else {
        tmem_ld32(a, r);
      }
    }
    if ((it & 3) == 3) {
      if (STORE) tmem_wait_st(); else asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      acc += r[0] + r[X - 1];
    }
  }
  if (STORE) tmem_wait_st(); else asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  const long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc + r[3];
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tslot, 512); }
}

template <int X, bool STORE>
void run(int threads, unsigned* out, long long* cyc) {
  const int iters = 4096;
  k<X, STORE><<<148, threads>>>(iters, out, cyc);
  long long c = 0;
  cudaError_t e = cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
  if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return; }
  const double bytes = double(iters) * threads * X * 4;
  printf("%s x%d, %2d warps: %.1f B/clk/SM (%.1f cycles per warp instruction)\n", STORE ? "tcgen05.st" : "tcgen05.ld", X, threads / 32,
         bytes / c, double(c) / iters);
}

int main() {
  unsigned* out; long long* cyc;
  cudaMalloc(&out, 148 * 512 * 4); cudaMalloc(&cyc, 8);
  for (int th : {128, 256, 512}) { run<16, false>(th, out, cyc); run<32, false>(th, out, cyc); }
  for (int th : {128, 256, 512}) { run<16, true>(th, out, cyc); run<32, true>(th, out, cyc); }
  return 0;
}
