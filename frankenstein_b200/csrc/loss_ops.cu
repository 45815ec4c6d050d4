// Masked L1 reconstruction loss of the VQ-VAE (reference: models/vq_brain.py:220-227,
// SoundStream.custom_l1_loss): mean |pred - gt| over the (batch, time) rows whose ground truth is
// not all-zero (zero rows are padding).  The reference selects rows with `nonzero` (a host sync
// and a dynamic shape); here the row mask, the masked sum and the row count are one fused pass and
// nothing leaves the device.  HBM-bound: reads pred + gt once.
#include "common.cuh"

namespace fk {

constexpr int kL1Rows = 8;

__device__ __forceinline__ float ld_as_float(const float* p, long long i) { return p[i]; }
__device__ __forceinline__ float ld_as_float(const __nv_bfloat16* p, long long i) { return __bfloat162float(p[i]); }
__device__ __forceinline__ void st_from_float(float* p, long long i, float v) { p[i] = v; }
__device__ __forceinline__ void st_from_float(__nv_bfloat16* p, long long i, float v) { p[i] = __float2bfloat16_rn(v); }

template <typename T>
__global__ void __launch_bounds__(kL1Rows * 32)
masked_l1_fwd_kernel(const T* __restrict__ pred, const float* __restrict__ gt, long long R, int C,
                     unsigned char* __restrict__ row_valid, float* __restrict__ part_sum, float* __restrict__ part_cnt,
                     unsigned int* __restrict__ counter, float* __restrict__ loss, float* __restrict__ denom_out) {
  __shared__ float red[32];
  __shared__ bool is_last;
  const int lane = threadIdx.x & 31;
  const long long row = static_cast<long long>(blockIdx.x) * kL1Rows + (threadIdx.x >> 5);
  float s = 0.f, c = 0.f;
  if (row < R) {
    const long long base = row * C;
    bool nz = false;
    float acc = 0.f;
    for (int d = lane; d < C; d += 32) {
      const float g = gt[base + d];
      nz |= (g != 0.f);
      acc += fabsf(ld_as_float(pred, base + d) - g);
    }
    nz = __any_sync(0xffffffffu, nz);
    if (nz) s = acc;
    if (lane == 0) {
      row_valid[row] = nz ? 1 : 0;
      c = nz ? 1.f : 0.f;
    }
  }
  const float bs = block_sum(s, red);
  const float bc = block_sum(c, red);
  if (threadIdx.x == 0) {
    part_sum[blockIdx.x] = bs;
    part_cnt[blockIdx.x] = bc;
    __threadfence();
    is_last = (atomicAdd(counter, 1u) == gridDim.x - 1);
  }
  __syncthreads();
  if (is_last) {
    __threadfence();
    float a = 0.f, b = 0.f;
    for (unsigned int i = threadIdx.x; i < gridDim.x; i += blockDim.x) {
      a += __ldcg(part_sum + i);
      b += __ldcg(part_cnt + i);
    }
    const float ta = block_sum(a, red);
    const float tb = block_sum(b, red);
    if (threadIdx.x == 0) {
      const float denom = tb * static_cast<float>(C);
      loss[0] = ta / denom;            // 0/0 = NaN for an all-padding batch, like torch.mean of an empty tensor
      denom_out[0] = denom;
      *counter = 0u;
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
masked_l1_bwd_kernel(const T* __restrict__ pred, const float* __restrict__ gt, const unsigned char* __restrict__ row_valid,
                     const float* __restrict__ g_loss, const float* __restrict__ denom, long long R, int C,
                     T* __restrict__ grad_pred) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= R * C) return;
  const long long row = i / C;
  float g = 0.f;
  if (row_valid[row]) {
    const float diff = ld_as_float(pred, i) - gt[i];
    const float sgn = (diff > 0.f) ? 1.f : ((diff < 0.f) ? -1.f : 0.f);
    g = sgn * g_loss[0] / denom[0];
  }
  st_from_float(grad_pred, i, g);
}

}  // namespace fk

using namespace fk;

extern "C" __attribute__((visibility("default"))) long long fk_masked_l1_partials(long long R) { return (R + kL1Rows - 1) / kL1Rows; }

extern "C" __attribute__((visibility("default"))) int fk_masked_l1_forward(const void* pred, int dtype, const float* gt, long long R, int C,
                                    unsigned char* row_valid, float* part_sum, float* part_cnt, unsigned int* counter,
                                    float* loss, float* denom, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  FK_REQUIRE(R > 0 && C > 0, "fk_masked_l1_forward: bad shape");
  FK_REQUIRE(pred && gt && row_valid && part_sum && part_cnt && counter && loss && denom, "fk_masked_l1_forward: null pointer");
  const unsigned grid = static_cast<unsigned>(fk_masked_l1_partials(R));
  if (dtype == 0)
    masked_l1_fwd_kernel<float><<<grid, kL1Rows * 32, 0, stream>>>(static_cast<const float*>(pred), gt, R, C, row_valid,
                                                                   part_sum, part_cnt, counter, loss, denom);
  else if (dtype == 1)
    masked_l1_fwd_kernel<__nv_bfloat16><<<grid, kL1Rows * 32, 0, stream>>>(static_cast<const __nv_bfloat16*>(pred), gt, R,
                                                                           C, row_valid, part_sum, part_cnt, counter, loss, denom);
  else
    FK_REQUIRE(false, "fk_masked_l1_forward: dtype must be 0 (f32) or 1 (bf16)");
  FK_CHECK_LAUNCH();
  fk_count_launch();
  return FK_OK;
}

extern "C" __attribute__((visibility("default"))) int fk_masked_l1_backward(const void* pred, int dtype, const float* gt, const unsigned char* row_valid,
                                     const float* g_loss, const float* denom, long long R, int C, void* grad_pred,
                                     void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  FK_REQUIRE(R > 0 && C > 0, "fk_masked_l1_backward: bad shape");
  FK_REQUIRE(pred && gt && row_valid && g_loss && denom && grad_pred, "fk_masked_l1_backward: null pointer");
  const unsigned grid = static_cast<unsigned>((R * C + 255) / 256);
  if (dtype == 0)
    masked_l1_bwd_kernel<float><<<grid, 256, 0, stream>>>(static_cast<const float*>(pred), gt, row_valid, g_loss, denom, R,
                                                          C, static_cast<float*>(grad_pred));
  else if (dtype == 1)
    masked_l1_bwd_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(pred), gt, row_valid,
                                                                  g_loss, denom, R, C, static_cast<__nv_bfloat16*>(grad_pred));
  else
    FK_REQUIRE(false, "fk_masked_l1_backward: dtype must be 0 (f32) or 1 (bf16)");
  FK_CHECK_LAUNCH();
  fk_count_launch();
  return FK_OK;
}
