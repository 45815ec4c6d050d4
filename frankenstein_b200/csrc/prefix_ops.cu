// Output side of the hot path: the hand-off of the perceiver's prefix tokens to the GPT-2
// decoder, models/gpt2_model.py:178-196 --
//     tok_emb = wte(idx); tok_emb = cat([prefix, tok_emb], dim=1); x = tok_emb + wpe(arange(t_ctx + t))
// as ONE bandwidth-bound pass (the reference materialises wte(idx), the concatenation, wpe(pos) and the sum: four passes
// over [B, t_ctx + t, n_embd]), and its backward: d prefix = g[:, :t_ctx] (a view, no kernel), d wpe[p] = sum_b g[b, p],
// d wte[idx[b, t]] += g[b, t_ctx + t] (fp32 red.global, as torch's embedding backward).
#include "common.cuh"
#include "fk_b200.h"

namespace fk {

template <typename PT>
__device__ __forceinline__ float4 load4(const PT* p);
template <>
__device__ __forceinline__ float4 load4<float>(const float* p) { return *reinterpret_cast<const float4*>(p); }
template <>
__device__ __forceinline__ float4 load4<__nv_bfloat16>(const __nv_bfloat16* p) {
  const uint2 w = *reinterpret_cast<const uint2*>(p);
  const __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162*>(&w.x), b = *reinterpret_cast<const __nv_bfloat162*>(&w.y);
  return make_float4(__bfloat162float(a.x), __bfloat162float(a.y), __bfloat162float(b.x), __bfloat162float(b.y));
}
template <typename OT>
__device__ __forceinline__ void store4(OT* p, float4 v);
template <>
__device__ __forceinline__ void store4<float>(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
template <>
__device__ __forceinline__ void store4<__nv_bfloat16>(__nv_bfloat16* p, float4 v) {
  __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
  *reinterpret_cast<uint2*>(p) = make_uint2(*reinterpret_cast<uint32_t*>(&a), *reinterpret_cast<uint32_t*>(&b));
}

// one thread per channel quad of one output token
template <typename PT, typename OT>
__global__ void __launch_bounds__(256)
prefix_embed_fwd_kernel(const PT* __restrict__ prefix, const long long* __restrict__ idx, const float* __restrict__ wte,
                        const float* __restrict__ wpe, OT* __restrict__ out, int B, int Tc, int T, int D, int V) {
  const int D4 = D >> 2;
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long total = static_cast<long long>(B) * (Tc + T) * D4;
  if (i >= total) return;
  const int c = static_cast<int>(i % D4) * 4;
  const long long tok = i / D4;
  const int p = static_cast<int>(tok % (Tc + T)), b = static_cast<int>(tok / (Tc + T));
  float4 v;
  if (p < Tc) {
    v = load4<PT>(prefix + (static_cast<long long>(b) * Tc + p) * D + c);
  } else {
    long long id = idx[static_cast<long long>(b) * T + (p - Tc)];
    id = id < 0 ? 0 : (id >= V ? V - 1 : id);                     // (the host wrapper rejects out-of-range ids)
    v = *reinterpret_cast<const float4*>(wte + id * D + c);
  }
  const float4 e = *reinterpret_cast<const float4*>(wpe + static_cast<long long>(p) * D + c);
  store4<OT>(out + tok * D + c, make_float4(v.x + e.x, v.y + e.y, v.z + e.z, v.w + e.w));
}

// one thread per channel quad of one POSITION: batch sum for d wpe (fixed order), scatter-add into d wte
__global__ void __launch_bounds__(256)
prefix_embed_bwd_kernel(const float* __restrict__ g, const long long* __restrict__ idx, float* __restrict__ dwte,
                        float* __restrict__ dwpe, int B, int Tc, int T, int D, int V) {
  const int D4 = D >> 2;
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= static_cast<long long>(Tc + T) * D4) return;
  const int c = static_cast<int>(i % D4) * 4;
  const int p = static_cast<int>(i / D4);
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int b = 0; b < B; ++b) {
    const float4 v = *reinterpret_cast<const float4*>(g + (static_cast<long long>(b) * (Tc + T) + p) * D + c);
    acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    if (p >= Tc) {
      long long id = idx[static_cast<long long>(b) * T + (p - Tc)];
      id = id < 0 ? 0 : (id >= V ? V - 1 : id);
      float* d = dwte + id * D + c;
      atomicAdd(d, v.x); atomicAdd(d + 1, v.y); atomicAdd(d + 2, v.z); atomicAdd(d + 3, v.w);
    }
  }
  float* d = dwpe + static_cast<long long>(p) * D + c;       // dwpe rows [0, Tc + T) are owned by this launch: plain add
  d[0] += acc.x; d[1] += acc.y; d[2] += acc.z; d[3] += acc.w;
}

}  // namespace fk

using namespace fk;

#define FK_API extern "C" __attribute__((visibility("default")))

FK_API int fk_prefix_embed_forward(const void* prefix, int prefix_dtype, const long long* idx, const float* wte, const float* wpe,
                                   void* out, int out_dtype, int B, int Tc, int T, int D, int V, int P, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  FK_REQUIRE(wte && wpe && out && B > 0 && Tc >= 0 && T >= 0 && Tc + T > 0 && D > 0 && D % 4 == 0 && V > 0,
             "fk_prefix_embed_forward: bad argument (n_embd % 4 == 0)");
  FK_REQUIRE((Tc == 0 || prefix) && (T == 0 || idx), "fk_prefix_embed_forward: prefix / idx missing");
  FK_REQUIRE(Tc + T <= P, "fk_prefix_embed_forward: sequence longer than the position table (block_size)");
  FK_REQUIRE((prefix_dtype == 0 || prefix_dtype == 1) && (out_dtype == 0 || out_dtype == 1), "fk_prefix_embed_forward: dtype 0 = f32, 1 = bf16");
  const long long total = static_cast<long long>(B) * (Tc + T) * (D / 4);
  const unsigned grid = static_cast<unsigned>((total + 255) / 256);
  const __nv_bfloat16* pb = static_cast<const __nv_bfloat16*>(prefix);
  const float* pf = static_cast<const float*>(prefix);
  if (prefix_dtype == 0 && out_dtype == 0) prefix_embed_fwd_kernel<float, float><<<grid, 256, 0, stream>>>(pf, idx, wte, wpe, static_cast<float*>(out), B, Tc, T, D, V);
  else if (prefix_dtype == 0) prefix_embed_fwd_kernel<float, __nv_bfloat16><<<grid, 256, 0, stream>>>(pf, idx, wte, wpe, static_cast<__nv_bfloat16*>(out), B, Tc, T, D, V);
  else if (out_dtype == 0) prefix_embed_fwd_kernel<__nv_bfloat16, float><<<grid, 256, 0, stream>>>(pb, idx, wte, wpe, static_cast<float*>(out), B, Tc, T, D, V);
  else prefix_embed_fwd_kernel<__nv_bfloat16, __nv_bfloat16><<<grid, 256, 0, stream>>>(pb, idx, wte, wpe, static_cast<__nv_bfloat16*>(out), B, Tc, T, D, V);
  FK_CHECK_LAUNCH();
  fk_count_launch();
  return FK_OK;
}

FK_API int fk_prefix_embed_backward(const float* g, const long long* idx, float* dwte, float* dwpe, int B, int Tc, int T, int D,
                                    int V, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  FK_REQUIRE(g && dwte && dwpe && B > 0 && Tc >= 0 && T >= 0 && Tc + T > 0 && D > 0 && D % 4 == 0 && V > 0 && (T == 0 || idx),
             "fk_prefix_embed_backward: bad argument");
  const long long total = static_cast<long long>(Tc + T) * (D / 4);
  prefix_embed_bwd_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, stream>>>(g, idx, dwte, dwpe, B, Tc, T, D, V);
  FK_CHECK_LAUNCH();
  fk_count_launch();
  return FK_OK;
}
