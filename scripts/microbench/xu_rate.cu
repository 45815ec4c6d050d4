// Microbenchmark: per-SM throughput of MUFU.EX2 (ex2.approx.ftz.f32), F2FP (cvt.rn.bf16x2.f32) and a mixed loop,
// 16 warps per SM (4 per scheduler), 8 independent chains per thread.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 xu_rate.cu -o xu_rate
#include <cstdio>
#include <cuda_bf16.h>

template <int OP>
__global__ void __launch_bounds__(512, 1) k(int iters, float* out, long long* cyc) {
  float x[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) x[i] = -1.0f - 0.001f * (threadIdx.x + i);
  unsigned acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (OP == 0) {
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[i]));
      } else if (OP == 1) {
        unsigned r;
        asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(x[i]), "f"(x[(i + 1) & 7]));
        acc ^= r;
      } else if (OP == 2) {
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[i]));
        x[i] = fmaf(x[i], 0.5f, -1.25f);
        x[i] = x[i] * 1.0001f;
        x[i] = x[i] - 0.001f;
      } else if (OP == 3) {
        asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(x[i]));
      } else if (OP == 4) {                     // two half-precision exponentials per instruction
        unsigned r = __float_as_uint(x[i]);
        asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(r));
        x[i] = __uint_as_float(r);
      } else if (OP == 5) {
        unsigned r = __float_as_uint(x[i]);
        asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(r));
        x[i] = __uint_as_float(r);
      } else if (OP == 6) {                     // the softmax inner step with packed half exponentials:
        unsigned r;                             // pack two fp32 arguments, one MUFU, (result stays f16x2)
        asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(x[i]), "f"(x[(i + 1) & 7]));
        asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(r));
        acc ^= r;
      } else if (OP == 7) {
        unsigned r;
        asm volatile("tanh.approx.f16x2 %0, %1;" : "=r"(r) : "r"(__float_as_uint(x[i])));
        x[i] = __uint_as_float(r);
      }
    }
  }
  const long long t1 = clock64();
  __syncthreads();
  float s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s + acc;
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

template <int OP>
void run(const char* name, float* out, long long* cyc) {
  const int iters = 2048;
  k<OP><<<148, 512>>>(iters, out, cyc);
  long long c = 0;
  cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
  const double ops = double(iters) * 8 * 512;            // thread-level ops per SM
  printf("%s: %.2f ops/clk/SM (%.1f cycles per warp instruction per scheduler)\n", name, ops / c, c / (double(iters) * 8 * 4));
}

int main() {
  float* out; long long* cyc;
  cudaMalloc(&out, 148 * 512 * 4); cudaMalloc(&cyc, 8);
  run<0>("MUFU.EX2", out, cyc);
  run<1>("F2FP.BF16.PACK_AB", out, cyc);
  run<2>("EX2 + 3 FP32 ops", out, cyc);
  run<3>("MUFU.RCP", out, cyc);
  run<4>("MUFU.EX2.F16x2 (instructions; x2 results)", out, cyc);
  run<5>("MUFU.EX2.BF16x2 (instructions; x2 results)", out, cyc);
  run<6>("cvt.f16x2 + EX2.F16x2 (pairs)", out, cyc);
  run<7>("MUFU.TANH.F16x2 (instructions; x2 results)", out, cyc);
  return 0;
}
