// Input side of the hot path on the device: the reference's NumPy preprocessing of the
// Utah-array features, utils/data_utils.py -- process_signal (:115-156: channel concat of spike power and threshold
// crossings, per-block z-score, Gaussian smoothing over time), z_score_per_block_scaling (:78-109),
// pad_truncate_brain_list (:243-267) and the float32 cast of BrainDataset.__getitem__ (:335-344) -- as three
// bandwidth-bound kernels over RAGGED trials stored back to back ([sum_T, C1] and [sum_T, C2] fp32 + row offsets):
//
//   input_trial_moments_kernel   per trial and channel: sum of x (pass 0) or of (x - block mean)^2 (pass 1), fp64
//   input_block_reduce_kernel    per block and channel: the trials' partial sums added in trial order (deterministic)
//                                -> mean (pass 0); -> std with the reference's zero rule (pass 1)
//   input_normalize_kernel       (x - mean) / std, optional 9-tap Gaussian (sigma 1, scipy 'reflect' boundary) over the
//                                trial's own bins, zero padding / truncation to T_out bins, fp32 or bf16 out -- one pass:
//                                the z-scored, smoothed and padded intermediates of the reference never exist
//
// Two passes over the input for the statistics (the numerically standard mean-then-deviations form np.std uses; a
// constant channel gives exactly std = 0 -> 1, as in the reference), one for the output: 3 reads + 1 write of the batch.
#include "common.cuh"
#include "fk_b200.h"

namespace fk {

constexpr int kInThreads = 256;        // 32 channel quads x 8 row groups
constexpr int kInRowGroups = 8;

__device__ __forceinline__ float4 load_quad(const float* __restrict__ volt, const float* __restrict__ spk, long long row, int c,
                                            int C1, int C2) {
  // channel c (multiple of 4) of the concatenated [volt | spk] row
  return (c < C1) ? *reinterpret_cast<const float4*>(volt + row * C1 + c)
                  : *reinterpret_cast<const float4*>(spk + row * C2 + (c - C1));
}

__global__ void __launch_bounds__(kInThreads)
input_trial_moments_kernel(const float* __restrict__ volt, const float* __restrict__ spk, const long long* __restrict__ offsets,
                           const int* __restrict__ block_id, int C1, int C2, const double* __restrict__ mean,
                           double* __restrict__ part) {
  __shared__ double red[kInRowGroups][128];
  const int trial = blockIdx.x, C = C1 + C2;
  const int cq = threadIdx.x & 31, rg = threadIdx.x >> 5;
  const int c = blockIdx.y * 128 + cq * 4;
  const long long r0 = offsets[trial], r1 = offsets[trial + 1];
  double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
  if (c < C) {
    double m0 = 0, m1 = 0, m2 = 0, m3 = 0;
    if (mean != nullptr) {
      const double* mp = mean + static_cast<long long>(block_id[trial]) * C + c;
      m0 = mp[0]; m1 = mp[1]; m2 = mp[2]; m3 = mp[3];
    }
    for (long long r = r0 + rg; r < r1; r += kInRowGroups) {
      const float4 v = load_quad(volt, spk, r, c, C1, C2);
      if (mean == nullptr) {
        a0 += v.x; a1 += v.y; a2 += v.z; a3 += v.w;
      } else {
        const double d0 = v.x - m0, d1 = v.y - m1, d2 = v.z - m2, d3 = v.w - m3;
        a0 += d0 * d0; a1 += d1 * d1; a2 += d2 * d2; a3 += d3 * d3;
      }
    }
  }
  red[rg][cq * 4 + 0] = a0; red[rg][cq * 4 + 1] = a1; red[rg][cq * 4 + 2] = a2; red[rg][cq * 4 + 3] = a3;
  __syncthreads();
  if (threadIdx.x < 128) {
    const int cc = blockIdx.y * 128 + threadIdx.x;
    if (cc < C) {
      double s = 0;
#pragma unroll
      for (int g = 0; g < kInRowGroups; ++g) s += red[g][threadIdx.x];     // fixed order
      part[static_cast<long long>(trial) * C + cc] = s;
    }
  }
}

__global__ void __launch_bounds__(256)
input_block_reduce_kernel(const double* __restrict__ part, const long long* __restrict__ offsets, const int* __restrict__ block_id,
                          int n_trials, int C, int pass, int zero_policy, double* __restrict__ mean_d,
                          float* __restrict__ mean_f, float* __restrict__ std_f) {
  const int blk = blockIdx.x;
  const int c = blockIdx.y * 256 + threadIdx.x;
  if (c >= C) return;
  double s = 0;
  long long n = 0;
  for (int t = 0; t < n_trials; ++t) {
    if (block_id[t] != blk) continue;
    s += part[static_cast<long long>(t) * C + c];
    n += offsets[t + 1] - offsets[t];
  }
  const long long o = static_cast<long long>(blk) * C + c;
  if (pass == 0) {
    const double m = n > 0 ? s / static_cast<double>(n) : 0.0;
    mean_d[o] = m;
    mean_f[o] = static_cast<float>(m);
  } else {
    float sd = n > 0 ? static_cast<float>(sqrt(s / static_cast<double>(n))) : 0.f;
    // process_signal: `block_std[block_std == 0] = 1`; StandardScaler (_handle_zeros_in_scale): scale < 10 eps -> 1
    if (zero_policy == 0 ? (sd == 0.f) : (sd < 10.f * 1.1920929e-07f)) sd = 1.f;
    std_f[o] = sd;
  }
}

struct GaussW { double w[9]; };

template <typename OutT>
__device__ __forceinline__ void store_quad(OutT* dst, float4 v);
template <>
__device__ __forceinline__ void store_quad<float>(float* dst, float4 v) { *reinterpret_cast<float4*>(dst) = v; }
template <>
__device__ __forceinline__ void store_quad<__nv_bfloat16>(__nv_bfloat16* dst, float4 v) {
  __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
  *reinterpret_cast<uint2*>(dst) = make_uint2(*reinterpret_cast<uint32_t*>(&a), *reinterpret_cast<uint32_t*>(&b));
}

constexpr int kInTimeChunk = 32;

// grid (time chunks of the OUTPUT, trial, channel slabs of 128 * 2); thread = one channel quad, one chunk of output bins
template <typename OutT, bool kSmooth>
__global__ void __launch_bounds__(64)
input_normalize_kernel(const float* __restrict__ volt, const float* __restrict__ spk, const long long* __restrict__ offsets,
                       const int* __restrict__ block_id, const float* __restrict__ mean, const float* __restrict__ stdv, int C1,
                       int C2, int T_out, OutT* __restrict__ out, const GaussW gw) {
  const int C = C1 + C2;
  const int trial = blockIdx.y;
  const int c = (blockIdx.z * 64 + threadIdx.x) * 4;
  if (c >= C) return;
  const long long r0 = offsets[trial];
  const int T = static_cast<int>(offsets[trial + 1] - r0);
  const int t0 = blockIdx.x * kInTimeChunk;
  const int t1 = min(t0 + kInTimeChunk, T_out);
  OutT* dst = out + (static_cast<long long>(trial) * T_out) * C + c;
  const int t_valid = min(t1, T);                       // bins below carry data, the rest of the chunk is padding
  if (t0 < t_valid) {
    const float4 m = *reinterpret_cast<const float4*>(mean + static_cast<long long>(block_id[trial]) * C + c);
    const float4 s = *reinterpret_cast<const float4*>(stdv + static_cast<long long>(block_id[trial]) * C + c);
    auto zrow = [&](int t) {                            // z-scored bin t of this trial (fp32, as the reference holds it)
      const float4 v = load_quad(volt, spk, r0 + t, c, C1, C2);
      return make_float4((v.x - m.x) / s.x, (v.y - m.y) / s.y, (v.z - m.z) / s.z, (v.w - m.w) / s.w);
    };
    if (!kSmooth) {
      for (int t = t0; t < t_valid; ++t) store_quad<OutT>(dst + static_cast<long long>(t) * C, zrow(t));
    } else {
      // scipy.ndimage.gaussian_filter1d(sigma=1): 9 taps, boundary mode 'reflect' (d c b a | a b c d | d c b a) over
      // the trial's own T bins (the filter runs before padding / truncation in the reference), fp64 accumulation
      auto reflect = [&](int i) {
        if (T == 1) return 0;
        const int period = 2 * T;
        i %= period;
        if (i < 0) i += period;
        return i < T ? i : period - 1 - i;
      };
      float4 win[9];                                    // sliding window: bins t-4 .. t+4
#pragma unroll
      for (int k = 0; k < 8; ++k) win[k + 1] = zrow(reflect(t0 - 4 + k));
      for (int t = t0; t < t_valid; ++t) {
#pragma unroll
        for (int k = 0; k < 8; ++k) win[k] = win[k + 1];
        win[8] = zrow(reflect(t + 4));
        double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
#pragma unroll
        for (int k = 0; k < 9; ++k) {
          a0 += gw.w[k] * win[k].x; a1 += gw.w[k] * win[k].y; a2 += gw.w[k] * win[k].z; a3 += gw.w[k] * win[k].w;
        }
        store_quad<OutT>(dst + static_cast<long long>(t) * C,
                         make_float4(static_cast<float>(a0), static_cast<float>(a1), static_cast<float>(a2), static_cast<float>(a3)));
      }
    }
  }
  for (int t = max(t0, t_valid); t < t1; ++t) store_quad<OutT>(dst + static_cast<long long>(t) * C, make_float4(0.f, 0.f, 0.f, 0.f));
}

}  // namespace fk

using namespace fk;

#define FK_API extern "C" __attribute__((visibility("default")))

static bool input_args_ok(const void* volt, const void* spk, const void* offsets, const void* block_id, int n_trials, int C1, int C2) {
  return volt && offsets && block_id && n_trials > 0 && C1 > 0 && C2 >= 0 && C1 % 4 == 0 && C2 % 4 == 0 && (C2 == 0 || spk) &&
         (reinterpret_cast<uintptr_t>(volt) & 15) == 0 && (reinterpret_cast<uintptr_t>(spk) & 15) == 0;
}

FK_API int fk_input_trial_moments(const float* volt, const float* spk, const long long* offsets, const int* block_id,
                                  int n_trials, int C1, int C2, const double* mean, double* part, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  FK_REQUIRE(input_args_ok(volt, spk, offsets, block_id, n_trials, C1, C2) && part,
             "fk_input_trial_moments: bad argument (channel counts must be multiples of 4, 16-byte aligned rows)");
  const int C = C1 + C2;
  const dim3 grid(static_cast<unsigned>(n_trials), static_cast<unsigned>((C + 127) / 128));
  input_trial_moments_kernel<<<grid, kInThreads, 0, stream>>>(volt, spk, offsets, block_id, C1, C2, mean, part);
  FK_CHECK_LAUNCH();
  fk_count_launch();
  return FK_OK;
}

FK_API int fk_input_block_reduce(const double* part, const long long* offsets, const int* block_id, int n_trials, int n_blocks,
                                 int C, int pass, int zero_policy, double* mean_d, float* mean_f, float* std_f, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  FK_REQUIRE(part && offsets && block_id && n_trials > 0 && n_blocks > 0 && C > 0 && (pass == 0 || pass == 1) &&
                 (zero_policy == 0 || zero_policy == 1),
             "fk_input_block_reduce: bad argument");
  FK_REQUIRE(pass == 0 ? (mean_d && mean_f) : (std_f != nullptr), "fk_input_block_reduce: missing output for this pass");
  const dim3 grid(static_cast<unsigned>(n_blocks), static_cast<unsigned>((C + 255) / 256));
  input_block_reduce_kernel<<<grid, 256, 0, stream>>>(part, offsets, block_id, n_trials, C, pass, zero_policy, mean_d, mean_f, std_f);
  FK_CHECK_LAUNCH();
  fk_count_launch();
  return FK_OK;
}

FK_API int fk_input_normalize(const float* volt, const float* spk, const long long* offsets, const int* block_id, const float* mean,
                              const float* stdv, int n_trials, int C1, int C2, int T_out, int smooth, void* out, int out_dtype,
                              void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  FK_REQUIRE(input_args_ok(volt, spk, offsets, block_id, n_trials, C1, C2) && mean && stdv && out && T_out > 0,
             "fk_input_normalize: bad argument (channel counts must be multiples of 4, 16-byte aligned rows)");
  FK_REQUIRE(out_dtype == 0 || out_dtype == 1, "fk_input_normalize: out_dtype 0 = f32, 1 = bf16");
  FK_REQUIRE(n_trials <= 65535, "fk_input_normalize: at most 65535 trials per call");
  const int C = C1 + C2;
  GaussW gw;
  {
    // scipy.ndimage._filters._gaussian_kernel1d(sigma = 1, order = 0, radius = int(4.0 * sigma + 0.5) = 4)
    double sum = 0;
    for (int k = 0; k < 9; ++k) { gw.w[k] = exp(-0.5 * (k - 4) * (k - 4)); sum += gw.w[k]; }
    for (int k = 0; k < 9; ++k) gw.w[k] /= sum;
  }
  const dim3 grid(static_cast<unsigned>((T_out + kInTimeChunk - 1) / kInTimeChunk), static_cast<unsigned>(n_trials),
                  static_cast<unsigned>((C / 4 + 63) / 64));
  if (out_dtype == 0) {
    if (smooth) input_normalize_kernel<float, true><<<grid, 64, 0, stream>>>(volt, spk, offsets, block_id, mean, stdv, C1, C2, T_out, static_cast<float*>(out), gw);
    else input_normalize_kernel<float, false><<<grid, 64, 0, stream>>>(volt, spk, offsets, block_id, mean, stdv, C1, C2, T_out, static_cast<float*>(out), gw);
  } else {
    if (smooth) input_normalize_kernel<__nv_bfloat16, true><<<grid, 64, 0, stream>>>(volt, spk, offsets, block_id, mean, stdv, C1, C2, T_out, static_cast<__nv_bfloat16*>(out), gw);
    else input_normalize_kernel<__nv_bfloat16, false><<<grid, 64, 0, stream>>>(volt, spk, offsets, block_id, mean, stdv, C1, C2, T_out, static_cast<__nv_bfloat16*>(out), gw);
  }
  FK_CHECK_LAUNCH();
  fk_count_launch();
  return FK_OK;
}
