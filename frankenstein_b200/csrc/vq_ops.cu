// Bandwidth-bound kernels around the codeword search: operand preparation, exact fp32 re-score +
// gather + straight-through + commitment loss, EMA statistics, EMA finalize + dead-code reset,
// backward, k-means mean update, perplexity.  All are warp-per-row kernels with 16-byte
// coalesced accesses; roofline = HBM.
//
// Reference semantics: vector_quantize_pytorch.VectorQuantize as configured at
// models/vq_brain.py:184-193 (restated in oracle/vector_quantize_ref.py).
#include "common.cuh"

namespace fk {

constexpr int kRowsPerBlock = 8;   // 8 warps, one row each

__device__ __forceinline__ float to_f(float v) { return v; }
__device__ __forceinline__ float to_f(__nv_bfloat16 v) { return __bfloat162float(v); }
__device__ __forceinline__ float to_f(__half v) { return __half2float(v); }

// ------------------------------------------------------------------------------------------
// K0: input transform.  cosine: xn = e / max(|e|, 1e-12) (F.normalize); Euclidean: xn = e.
// Writes the fp32 row (if xn != null), the zero-padded bf16 row for the tensor cores and 1/|e|.
// ------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(kRowsPerBlock * 32)
vq_prepare_input_kernel(const T* __restrict__ e, long long N, int D, int Dp, int cosine, float* __restrict__ xn,
                        __nv_bfloat16* __restrict__ xb, float* __restrict__ inv_norm) {
  const int lane = threadIdx.x & 31;
  const long long row = static_cast<long long>(blockIdx.x) * kRowsPerBlock + (threadIdx.x >> 5);
  if (row >= N) return;
  const T* er = e + row * D;
  float denom = 1.f;
  if (cosine) {
    float ss = 0.f;
    for (int d = lane; d < D; d += 32) { const float v = to_f(er[d]); ss += v * v; }
    ss = warp_sum(ss);
    denom = fmaxf(sqrtf(ss), 1e-12f);   // F.normalize: x / max(|x|, eps)
    if (lane == 0 && inv_norm) inv_norm[row] = 1.f / denom;
  }
  for (int d = lane; d < Dp; d += 32) {
    float v = 0.f;
    if (d < D) {
      v = cosine ? to_f(er[d]) / denom : to_f(er[d]);
      if (xn) xn[row * D + d] = v;
    }
    xb[row * Dp + d] = __float2bfloat16_rn(v);
  }
}

// ------------------------------------------------------------------------------------------
// codebook operand: bf16 copy (zero padded to Dp) and c2 = |bf16(c)|^2 (Euclidean) / 0 (cosine),
// padded with +inf up to Kpad.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kRowsPerBlock * 32)
vq_prepare_codebook_kernel(const float* __restrict__ embed, int K, int D, int Dp, int Kpad, int cosine,
                           __nv_bfloat16* __restrict__ cb, float* __restrict__ c2pad) {
  const int lane = threadIdx.x & 31;
  const int k = blockIdx.x * kRowsPerBlock + (threadIdx.x >> 5);
  if (k >= Kpad) return;
  if (k >= K) {
    if (lane == 0) c2pad[k] = __int_as_float(0x7f800000);
    return;
  }
  float ss = 0.f;
  for (int d = lane; d < Dp; d += 32) {
    const __nv_bfloat16 b = __float2bfloat16_rn(d < D ? embed[static_cast<long long>(k) * D + d] : 0.f);
    cb[static_cast<long long>(k) * Dp + d] = b;
    const float f = __bfloat162float(b);
    ss += f * f;
  }
  ss = warp_sum(ss);
  if (lane == 0) c2pad[k] = cosine ? 0.f : ss;
}

// ------------------------------------------------------------------------------------------
// K2-K4: merge the candidate slots of a row, take the (up to) 4 best approximate keys, re-score in
// exact fp32 every candidate that lies within the bf16 error margin of the best key, pick with the
// reference's formula and lowest-index tie-break, gather the fp32 codeword, straight-through
// output, commitment-loss partial sums.
//   Euclidean (upstream cdist): d = sqrt(max(|x|^2 + |c|^2 - 2 x.c, 0)), minimise
//   cosine                    : s = x.c, maximise
// Margin: the bf16 rounding of both operands perturbs x.c by a zero-mean error of standard
// deviation ~ 0.8 * 2^-9 * sqrt(sum x_d^2 c_d^2) <= 0.8 * 2^-9 * |x||c|; the key carries alpha
// (<= 2) times that.  margin = |x||c| * 2^-9 * 32 / sqrt(D) is > 12 sigma for non-sparse vectors.
// ------------------------------------------------------------------------------------------
constexpr int kFinishCand = 4;   // must match kCand of the search kernel

struct FinishParams {
  const float* xn;       // [N, D]
  const float* embed;    // [K, D]
  const float* cand_val; // [N, S, 4]
  const int* cand_idx;   // [N, S, 4]
  long long* indices;    // [N]
  float* quantize;       // [N, D]
  float* partials;       // [gridDim.x]
  unsigned int* counter; // zero on entry, zero on exit
  float* loss;           // [1]
  long long N;
  int K, D, S, cosine, training;
  float loss_scale;      // commitment_weight / (N * D)
  float margin_scale;    // 2^-9 * 32 / sqrt(D)
};

__device__ __forceinline__ bool pair_less(float va, int ia, float vb, int ib) {
  return (va < vb) || (va == vb && ia < ib);
}

__global__ void __launch_bounds__(kRowsPerBlock * 32)
vq_finish_kernel(const FinishParams p) {
  __shared__ float red[32];
  __shared__ bool is_last;
  const int lane = threadIdx.x & 31;
  const long long row = static_cast<long long>(blockIdx.x) * kRowsPerBlock + (threadIdx.x >> 5);
  const float inf = __int_as_float(0x7f800000);
  float sq_err = 0.f;
  if (row < p.N) {
    // ---- the 4 smallest (key, index) pairs over all slots, one warp-wide selection round each ----
    const int C = p.S * kFinishCand;
    const float* cv = p.cand_val + row * C;
    const int* ci = p.cand_idx + row * C;
    float sel_v[kFinishCand];
    int sel_i[kFinishCand];
    float prev_v = -inf;
    int prev_i = -1;
#pragma unroll
    for (int r = 0; r < kFinishCand; ++r) {
      float bv = inf;
      int bi = 0x7fffffff;
      for (int c = lane; c < C; c += 32) {
        const int j = ci[c];
        const float v = cv[c];
        if (j < 0 || j >= p.K || !(v < inf)) continue;
        if (r > 0 && !pair_less(prev_v, prev_i, v, j)) continue;     // already selected
        if (pair_less(v, j, bv, bi)) { bv = v; bi = j; }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (pair_less(ov, oi, bv, bi)) { bv = ov; bi = oi; }
      }
      sel_v[r] = bv;
      sel_i[r] = (bi == 0x7fffffff) ? -1 : bi;
      prev_v = bv;
      prev_i = bi;
    }
    if (sel_i[0] < 0) sel_i[0] = 0;          // all-NaN row: nothing survived a comparison
    // ---- exact fp32 scores ----
    const float* xr = p.xn + row * p.D;
    float xx = 0.f;
    for (int d = lane * 4; d < p.D; d += 128) {
      const float4 x = *reinterpret_cast<const float4*>(xr + d);
      xx += x.x * x.x + x.y * x.y + x.z * x.z + x.w * x.w;
    }
    xx = warp_sum(xx);
    int best = sel_i[0];
    float best_score = 0.f, margin = 0.f;
#pragma unroll
    for (int r = 0; r < kFinishCand; ++r) {
      if (r > 0 && (sel_i[r] < 0 || !(sel_v[r] <= sel_v[0] + margin))) break;   // warp-uniform
      const float* c = p.embed + static_cast<long long>(sel_i[r]) * p.D;
      float dot = 0.f, nn = 0.f;
      for (int d = lane * 4; d < p.D; d += 128) {
        const float4 x = *reinterpret_cast<const float4*>(xr + d);
        const float4 a = *reinterpret_cast<const float4*>(c + d);
        dot += x.x * a.x + x.y * a.y + x.z * a.z + x.w * a.w;
        nn += a.x * a.x + a.y * a.y + a.z * a.z + a.w * a.w;
      }
      dot = warp_sum(dot);
      nn = warp_sum(nn);
      // smaller-is-better exact score in the reference's own arithmetic
      const float score = p.cosine ? -dot : sqrtf(fmaxf(xx + nn - 2.f * dot, 0.f));
      if (r == 0) {
        best_score = score;
        margin = sqrtf(xx * nn) * p.margin_scale + fabsf(sel_v[0]) * (1.f / 2048.f);
      } else if (score < best_score || (score == best_score && sel_i[r] < best)) {
        best_score = score;
        best = sel_i[r];
      }
    }
    const float* cq = p.embed + static_cast<long long>(best) * p.D;
    if (lane == 0) p.indices[row] = best;
    float* qr = p.quantize + row * p.D;
    for (int d = lane * 4; d < p.D && p.quantize != nullptr; d += 128) {
      const float4 x = *reinterpret_cast<const float4*>(xr + d);
      const float4 q = *reinterpret_cast<const float4*>(cq + d);
      float4 o;
      if (p.training) {
        // straight-through value x + (q - x), rounded exactly like the reference expression
        const float ex = q.x - x.x, ey = q.y - x.y, ez = q.z - x.z, ew = q.w - x.w;
        o = make_float4(x.x + ex, x.y + ey, x.z + ez, x.w + ew);
        sq_err += ex * ex + ey * ey + ez * ez + ew * ew;
      } else {
        o = q;
      }
      *reinterpret_cast<float4*>(qr + d) = o;
    }
  }
  if (!p.training) return;
  // ---- commitment loss: deterministic two-level reduction (block partials, last block sums) ----
  const float bsum = block_sum(sq_err, red);
  if (threadIdx.x == 0) {
    p.partials[blockIdx.x] = bsum;
    __threadfence();
    const unsigned int done = atomicAdd(p.counter, 1u);
    is_last = (done == gridDim.x - 1);
  }
  __syncthreads();
  if (is_last) {
    __threadfence();
    float acc = 0.f;
    for (unsigned int i = threadIdx.x; i < gridDim.x; i += blockDim.x) acc += __ldcg(p.partials + i);
    const float total = block_sum(acc, red);
    if (threadIdx.x == 0) {
      p.loss[0] = total * p.loss_scale;
      *p.counter = 0u;
    }
  }
}

// ------------------------------------------------------------------------------------------
// K5: EMA statistics.  stats = [K*D embed_sum | K bins] (one packed buffer = one all-reduce).
// A sort-free, atomic-free-on-data segmented sum keyed by the code index: the rows of each code are listed (counting
// sort: integer histogram -> exclusive scan -> fill), and ONE WARP PER CODE adds its rows in ascending row order with
// coalesced 16-byte loads, so embed_sum is bit-reproducible from run to run (fp32 atomics to [K, D] were not) and
// needs neither a memset nor a read-modify-write of the [K, D] buffer.  Workspace: int [3*K + N].
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
vq_ema_count_kernel(const long long* __restrict__ indices, long long N, int K, int* __restrict__ count) {
  const long long row = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x;
  if (row >= N) return;
  const long long k = indices[row];
  if (k >= 0 && k < K) atomicAdd(count + k, 1);          // integer: order independent
}

constexpr int kSegShort = 32;      // segments up to this many rows are summed by one warp (rank sort in registers)
constexpr int kSegSort = 1024;     // up to this many: one CTA sorts the row ids in shared memory; beyond: row scan

// single block: offs[k] = exclusive prefix of count; cursor[k] = 0; long_list = codes with more than kSegShort rows
__global__ void __launch_bounds__(1024)
vq_ema_scan_kernel(const int* __restrict__ count, int K, int* __restrict__ offs, int* __restrict__ cursor,
                   int* __restrict__ n_long, int* __restrict__ long_list) {
  __shared__ int scan[1024];
  __shared__ int s_long;
  if (threadIdx.x == 0) s_long = 0;
  const int per = (K + blockDim.x - 1) / blockDim.x;
  const int k0 = min(K, static_cast<int>(threadIdx.x) * per), k1 = min(K, k0 + per);
  int sum = 0;
  for (int k = k0; k < k1; ++k) sum += count[k];
  scan[threadIdx.x] = sum;
  __syncthreads();
  for (int off = 1; off < blockDim.x; off <<= 1) {
    const int v = (threadIdx.x >= off) ? scan[threadIdx.x - off] : 0;
    __syncthreads();
    scan[threadIdx.x] += v;
    __syncthreads();
  }
  int run = scan[threadIdx.x] - sum;
  for (int k = k0; k < k1; ++k) {
    offs[k] = run;
    cursor[k] = 0;
    const int c = count[k];
    run += c;
    if (c > kSegShort) long_list[atomicAdd(&s_long, 1)] = k;      // (any order: every long code is summed on its own)
  }
  __syncthreads();
  if (threadIdx.x == 0) *n_long = s_long;
}

__global__ void __launch_bounds__(256)
vq_ema_fill_kernel(const long long* __restrict__ indices, long long N, int K, const int* __restrict__ offs,
                   int* __restrict__ cursor, int* __restrict__ list) {
  const long long row = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x;
  if (row >= N) return;
  const long long k = indices[row];
  if (k >= 0 && k < K) list[offs[k] + atomicAdd(cursor + k, 1)] = static_cast<int>(row);   // any order; sorted below
}

// one warp per code: rows in ascending order (the canonical order that makes the fp32 sum reproducible)
__global__ void __launch_bounds__(kRowsPerBlock * 32)
vq_ema_segsum_kernel(const float* __restrict__ xn, const int* __restrict__ count, const int* __restrict__ offs,
                     const int* __restrict__ list, int K, int D, float* __restrict__ stats) {
  const int lane = threadIdx.x & 31;
  const int k = blockIdx.x * kRowsPerBlock + (threadIdx.x >> 5);
  if (k >= K) return;
  const int c = count[k];
  const int* seg = list + offs[k];
  float4 acc[2] = {make_float4(0.f, 0.f, 0.f, 0.f), make_float4(0.f, 0.f, 0.f, 0.f)};   // D <= 256: 2 float4 per lane
  auto add_row = [&](int row) {
    const float* xr = xn + static_cast<long long>(row) * D;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int d = lane * 4 + i * 128;
      if (d < D) {
        const float4 x = ldg_nc_f4(xr + d);
        acc[i].x += x.x; acc[i].y += x.y; acc[i].z += x.z; acc[i].w += x.w;
      }
    }
  };
  if (c > kSegShort) return;         // a hot code: vq_ema_segsum_long_kernel
  {
    // rank sort inside the warp: lane i holds one row id; the lane whose rank is t supplies the t-th row
    const int id = lane < c ? seg[lane] : 0x7fffffff;
    int rank = 0;
#pragma unroll 8
    for (int j = 0; j < 32; ++j) rank += (__shfl_sync(0xffffffffu, id, j) < id) ? 1 : 0;
    for (int t = 0; t < c; ++t) {
      const unsigned m = __ballot_sync(0xffffffffu, rank == t && lane < c);
      add_row(__shfl_sync(0xffffffffu, id, __ffs(m) - 1));
    }
  }
  float* dst = stats + static_cast<long long>(k) * D;
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int d = lane * 4 + i * 128;
    if (d < D) *reinterpret_cast<float4*>(dst + d) = acc[i];
  }
  if (lane == 0) stats[static_cast<long long>(K) * D + k] = static_cast<float>(c);
}

// Hot codes (more than kSegShort rows): one CTA of 32 warps per code.  The rows are brought into ascending order -- up
// to kSegSort ids by a bitonic sort in shared memory, beyond that by scanning the index array itself, which is in row
// order -- cut into 32 contiguous ranges, each summed by one warp in order, and the 32 partial sums are added in warp
// order: a fixed summation tree for a given assignment, whatever order the fill kernel's atomics produced.
__global__ void __launch_bounds__(1024)
vq_ema_segsum_long_kernel(const float* __restrict__ xn, const long long* __restrict__ indices, long long N,
                          const int* __restrict__ count, const int* __restrict__ offs, const int* __restrict__ list,
                          const int* __restrict__ n_long, const int* __restrict__ long_list, int K, int D,
                          float* __restrict__ stats) {
  __shared__ int ids[kSegSort];
  __shared__ float part[32][256];
  if (static_cast<int>(blockIdx.x) >= *n_long) return;
  const int k = long_list[blockIdx.x];
  const int c = count[k];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float4 acc[2] = {make_float4(0.f, 0.f, 0.f, 0.f), make_float4(0.f, 0.f, 0.f, 0.f)};
  auto add_row = [&](long long row) {
    const float* xr = xn + row * D;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int d = lane * 4 + i * 128;
      if (d < D) {
        const float4 x = ldg_nc_f4(xr + d);
        acc[i].x += x.x; acc[i].y += x.y; acc[i].z += x.z; acc[i].w += x.w;
      }
    }
  };
  if (c <= kSegSort) {
    ids[threadIdx.x] = static_cast<int>(threadIdx.x) < c ? list[offs[k] + threadIdx.x] : 0x7fffffff;
    __syncthreads();
    for (int size = 2; size <= kSegSort; size <<= 1) {
      for (int stride = size >> 1; stride > 0; stride >>= 1) {
        const int i = threadIdx.x, j = i ^ stride;
        if (j > i) {
          const int a = ids[i], b = ids[j];
          const bool up = (i & size) == 0;
          if ((a > b) == up) { ids[i] = b; ids[j] = a; }
        }
        __syncthreads();
      }
    }
    const int per = (c + 31) / 32;
    const int t0 = warp * per, t1 = min(c, t0 + per);
    for (int t = t0; t < t1; ++t) add_row(ids[t]);
  } else {
    // the index array is in row order: warp w takes the w-th 32nd of the rows and adds the ones assigned to this code
    const long long per = ((N + 31) / 32 + 31) / 32 * 32;
    const long long r0 = warp * per, r1 = min(N, r0 + per);
    for (long long r = r0; r < r1; r += 32) {
      const bool hit = (r + lane < r1) && indices[r + lane] == k;
      unsigned m = __ballot_sync(0xffffffffu, hit);
      while (m) {
        add_row(r + (__ffs(m) - 1));
        m &= m - 1;
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int d = lane * 4 + i * 128;
    if (d < D) *reinterpret_cast<float4*>(&part[warp][d]) = acc[i];
  }
  __syncthreads();
  if (static_cast<int>(threadIdx.x) < D) {
    float sum = 0.f;
#pragma unroll 8
    for (int w = 0; w < 32; ++w) sum += part[w][threadIdx.x];
    stats[static_cast<long long>(k) * D + threadIdx.x] = sum;
  }
  if (threadIdx.x == 0) stats[static_cast<long long>(K) * D + k] = static_cast<float>(c);
}

// ------------------------------------------------------------------------------------------
// K6a: cluster_size EMA, its total, expired flags and their ranks (single block, K <= 2^20).
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024)
vq_ema_cluster_kernel(const float* __restrict__ bins, float* __restrict__ cluster_size, int K, float decay,
                      float threshold, float* __restrict__ total_out, int* __restrict__ expire_rank,
                      int* __restrict__ n_expired) {
  __shared__ float red[32];
  __shared__ int scan[1024];
  const int per = (K + blockDim.x - 1) / blockDim.x;
  const int k0 = threadIdx.x * per, k1 = min(K, k0 + per);
  float sum = 0.f;
  int cnt = 0;
  for (int k = k0; k < k1; ++k) {
    const float cs = cluster_size[k];
    const float nv = cs + (1.f - decay) * (bins[k] - cs);   // torch lerp_(end, weight), weight < 0.5
    cluster_size[k] = nv;
    sum += nv;
    cnt += (nv < threshold) ? 1 : 0;
  }
  const float total = block_sum(sum, red);
  if (threadIdx.x == 0) *total_out = total;
  // exclusive scan of per-thread expired counts
  scan[threadIdx.x] = cnt;
  __syncthreads();
  for (int off = 1; off < blockDim.x; off <<= 1) {
    const int v = (threadIdx.x >= off) ? scan[threadIdx.x - off] : 0;
    __syncthreads();
    scan[threadIdx.x] += v;
    __syncthreads();
  }
  int rank = scan[threadIdx.x] - cnt;
  if (threadIdx.x == blockDim.x - 1) *n_expired = scan[threadIdx.x];
  for (int k = k0; k < k1; ++k) {
    const bool ex = cluster_size[k] < threshold;
    expire_rank[k] = ex ? rank : -1;
    rank += ex ? 1 : 0;
  }
}

// ------------------------------------------------------------------------------------------
// K6b: embed_avg EMA, Laplace-smoothed normalisation, (cosine) l2norm, dead-code replacement,
// and the next step's tensor-core operand (bf16 codebook + c2).  One warp per code.
// ------------------------------------------------------------------------------------------
struct EmaParams {
  const float* stats;        // [K*D | K] (already all-reduced)
  float* cluster_size;       // [K]  (already EMA-updated by K6a)
  float* embed_avg;          // [K, D]
  float* embed;              // [K, D]
  const float* total;        // [1]
  const int* expire_rank;    // [K]
  const long long* sample_rows;   // [n_sample] rows of xn used for dead-code replacement (may be null)
  const float* xn;           // [N, D]
  __nv_bfloat16* cb;         // [K, Dp]
  float* c2pad;              // [Kpad]
  long long N;
  int K, D, Dp, Kpad, cosine, n_sample;
  float decay, eps, threshold;
};

__global__ void __launch_bounds__(kRowsPerBlock * 32)
vq_ema_update_kernel(const EmaParams p) {
  const int lane = threadIdx.x & 31;
  const int k = blockIdx.x * kRowsPerBlock + (threadIdx.x >> 5);
  if (k >= p.Kpad) return;
  if (k >= p.K) {
    if (lane == 0) p.c2pad[k] = __int_as_float(0x7f800000);
    return;
  }
  const long long base = static_cast<long long>(k) * p.D;
  const int rank = p.expire_rank[k];
  const bool expired = rank >= 0 && p.sample_rows != nullptr && p.n_sample > 0;
  float ss = 0.f;
  if (!expired) {
    const float total = *p.total;
    const float cs = p.cluster_size[k];
    // laplace_smoothing(cs, K, eps) * total
    const float smoothed = (cs + p.eps) / (total + p.K * p.eps) * total;
    for (int d = lane; d < p.D; d += 32) {
      const float avg = p.embed_avg[base + d];
      const float nv = avg + (1.f - p.decay) * (p.stats[base + d] - avg);
      p.embed_avg[base + d] = nv;
      const float e = nv / smoothed;
      p.embed[base + d] = e;
      ss += e * e;
    }
    if (p.cosine) {
      ss = warp_sum(ss);
      const float nrm = fmaxf(sqrtf(ss), 1e-12f);
      for (int d = lane; d < p.D; d += 32) p.embed[base + d] = p.embed[base + d] / nrm;
    }
  } else {
    // upstream expire_codes_/replace: embed <- sampled row (l2norm'd for cosine), embed_avg <- row * thr,
    // cluster_size <- thr.  The EMA of embed_avg for this code is overwritten, as in the reference.
    const long long r = p.sample_rows[rank % p.n_sample];
    const float* xr = p.xn + r * p.D;
    float nrm = 1.f;
    if (p.cosine) {
      float s2 = 0.f;
      for (int d = lane; d < p.D; d += 32) s2 += xr[d] * xr[d];
      s2 = warp_sum(s2);
      nrm = fmaxf(sqrtf(s2), 1e-12f);
    }
    for (int d = lane; d < p.D; d += 32) {
      const float v = p.cosine ? xr[d] / nrm : xr[d];
      p.embed[base + d] = v;
      p.embed_avg[base + d] = v * p.threshold;
    }
    if (lane == 0) p.cluster_size[k] = p.threshold;
  }
  __syncwarp();
  // operand for the next search
  float s2 = 0.f;
  for (int d = lane; d < p.Dp; d += 32) {
    const __nv_bfloat16 b = __float2bfloat16_rn(d < p.D ? p.embed[base + d] : 0.f);
    p.cb[static_cast<long long>(k) * p.Dp + d] = b;
    const float f = __bfloat162float(b);
    s2 += f * f;
  }
  s2 = warp_sum(s2);
  if (lane == 0) p.c2pad[k] = p.cosine ? 0.f : s2;
}

// ------------------------------------------------------------------------------------------
// backward: dL/dx = g_out + g_loss * (2 w / (N D)) * (x - q); cosine chains through F.normalize:
// dL/de = (g - x (x.g)) / max(|e|, eps).
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kRowsPerBlock * 32)
vq_backward_kernel(const float* __restrict__ g_out, const float* __restrict__ g_loss, const float* __restrict__ xn,
                   const float* __restrict__ q, const float* __restrict__ inv_norm, long long N, int D, int cosine,
                   float coef, float* __restrict__ ge) {
  const int lane = threadIdx.x & 31;
  const long long row = static_cast<long long>(blockIdx.x) * kRowsPerBlock + (threadIdx.x >> 5);
  if (row >= N) return;
  const float gl = g_loss ? g_loss[0] * coef : 0.f;
  const long long base = row * D;
  float dot = 0.f;
  if (cosine) {
    for (int d = lane * 4; d < D; d += 128) {
      const float4 x = *reinterpret_cast<const float4*>(xn + base + d);
      const float4 qq = *reinterpret_cast<const float4*>(q + base + d);
      float4 g = g_out ? *reinterpret_cast<const float4*>(g_out + base + d) : make_float4(0, 0, 0, 0);
      g.x += gl * (x.x - qq.x); g.y += gl * (x.y - qq.y); g.z += gl * (x.z - qq.z); g.w += gl * (x.w - qq.w);
      dot += g.x * x.x + g.y * x.y + g.z * x.z + g.w * x.w;
    }
    dot = warp_sum(dot);
  }
  const float inv = cosine ? inv_norm[row] : 1.f;
  for (int d = lane * 4; d < D; d += 128) {
    const float4 x = *reinterpret_cast<const float4*>(xn + base + d);
    const float4 qq = *reinterpret_cast<const float4*>(q + base + d);
    float4 g = g_out ? *reinterpret_cast<const float4*>(g_out + base + d) : make_float4(0, 0, 0, 0);
    g.x += gl * (x.x - qq.x); g.y += gl * (x.y - qq.y); g.z += gl * (x.z - qq.z); g.w += gl * (x.w - qq.w);
    if (cosine) {
      g.x = (g.x - x.x * dot) * inv; g.y = (g.y - x.y * dot) * inv;
      g.z = (g.z - x.z * dot) * inv; g.w = (g.w - x.w * dot) * inv;
    }
    *reinterpret_cast<float4*>(ge + base + d) = g;
  }
}

// ------------------------------------------------------------------------------------------
// k-means mean update (upstream `kmeans`): new = sum / max(bins, 1) (cosine: l2norm), keep the
// old mean where bins == 0.  Also refreshes the tensor-core operand.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kRowsPerBlock * 32)
vq_kmeans_update_kernel(const float* __restrict__ stats, float* __restrict__ means, int K, int D, int Dp, int Kpad,
                        int cosine, __nv_bfloat16* __restrict__ cb, float* __restrict__ c2pad) {
  const int lane = threadIdx.x & 31;
  const int k = blockIdx.x * kRowsPerBlock + (threadIdx.x >> 5);
  if (k >= Kpad) return;
  if (k >= K) {
    if (lane == 0) c2pad[k] = __int_as_float(0x7f800000);
    return;
  }
  const long long base = static_cast<long long>(k) * D;
  const float bins = stats[static_cast<long long>(K) * D + k];
  if (bins != 0.f) {
    float ss = 0.f;
    for (int d = lane; d < D; d += 32) {
      const float v = stats[base + d] / bins;
      means[base + d] = v;
      ss += v * v;
    }
    if (cosine) {
      ss = warp_sum(ss);
      const float nrm = fmaxf(sqrtf(ss), 1e-12f);
      for (int d = lane; d < D; d += 32) means[base + d] = means[base + d] / nrm;
    }
  }
  __syncwarp();
  float s2 = 0.f;
  for (int d = lane; d < Dp; d += 32) {
    const __nv_bfloat16 b = __float2bfloat16_rn(d < D ? means[base + d] : 0.f);
    cb[static_cast<long long>(k) * Dp + d] = b;
    const float f = __bfloat162float(b);
    s2 += f * f;
  }
  s2 = warp_sum(s2);
  if (lane == 0) c2pad[k] = cosine ? 0.f : s2;
}

// ------------------------------------------------------------------------------------------
// K7: perplexity from the code histogram (models/vq_brain.py:238-243): exp(-sum p log(p + 1e-10)).
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024)
vq_perplexity_kernel(const float* __restrict__ bins, int K, float inv_n, float* __restrict__ out) {
  __shared__ float red[32];
  float acc = 0.f;
  for (int k = threadIdx.x; k < K; k += blockDim.x) {
    const float pk = bins[k] * inv_n;
    acc += pk * logf(pk + 1e-10f);
  }
  const float s = block_sum(acc, red);
  if (threadIdx.x == 0) out[0] = expf(-s);
}

__global__ void __launch_bounds__(256)
vq_histogram_kernel(const long long* __restrict__ indices, long long N, int K, float* __restrict__ bins) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const long long k = indices[i];
  if (k >= 0 && k < K) atomicAdd(bins + k, 1.f);
}

}  // namespace fk

using namespace fk;

static inline unsigned row_blocks(long long rows) { return static_cast<unsigned>((rows + kRowsPerBlock - 1) / kRowsPerBlock); }

extern "C" __attribute__((visibility("default"))) int fk_vq_prepare_input(const void* e, int dtype, long long N, int D, int Dp, int use_cosine, float* xn,
                                   void* x_bf16, float* inv_norm, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  FK_REQUIRE(N > 0 && D > 0 && Dp >= D, "fk_vq_prepare_input: bad shape");
  FK_REQUIRE(e && x_bf16, "fk_vq_prepare_input: null pointer");
  FK_REQUIRE(!use_cosine || inv_norm, "fk_vq_prepare_input: cosine needs inv_norm");
  auto* xb = static_cast<__nv_bfloat16*>(x_bf16);
  switch (dtype) {
    case 0: vq_prepare_input_kernel<float><<<row_blocks(N), kRowsPerBlock * 32, 0, stream>>>(
                static_cast<const float*>(e), N, D, Dp, use_cosine, xn, xb, inv_norm); break;
    case 1: vq_prepare_input_kernel<__nv_bfloat16><<<row_blocks(N), kRowsPerBlock * 32, 0, stream>>>(
                static_cast<const __nv_bfloat16*>(e), N, D, Dp, use_cosine, xn, xb, inv_norm); break;
    case 2: vq_prepare_input_kernel<__half><<<row_blocks(N), kRowsPerBlock * 32, 0, stream>>>(
                static_cast<const __half*>(e), N, D, Dp, use_cosine, xn, xb, inv_norm); break;
    default: FK_REQUIRE(false, "fk_vq_prepare_input: dtype must be 0 (f32), 1 (bf16) or 2 (f16)");
  }
  FK_CHECK_LAUNCH();
  fk_count_launch();
  return FK_OK;
}

extern "C" __attribute__((visibility("default"))) int fk_vq_prepare_codebook(const float* embed, int K, int D, int Dp, int Kpad, int use_cosine, void* cb_bf16,
                                      float* c2pad, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  FK_REQUIRE(K > 0 && D > 0 && Dp >= D && Kpad >= K, "fk_vq_prepare_codebook: bad shape");
  FK_REQUIRE(embed && cb_bf16 && c2pad, "fk_vq_prepare_codebook: null pointer");
  vq_prepare_codebook_kernel<<<row_blocks(Kpad), kRowsPerBlock * 32, 0, stream>>>(
      embed, K, D, Dp, Kpad, use_cosine, static_cast<__nv_bfloat16*>(cb_bf16), c2pad);
  FK_CHECK_LAUNCH();
  fk_count_launch();
  return FK_OK;
}

extern "C" __attribute__((visibility("default"))) int fk_vq_finish(const float* xn, const float* embed, const float* cand_val, const int* cand_idx, long long N,
                            int K, int D, int S, int use_cosine, int training, float commitment_weight,
                            long long* indices, float* quantize, float* loss, float* partials, unsigned int* counter,
                            void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  FK_REQUIRE(N > 0 && K > 0 && D > 0 && S > 0, "fk_vq_finish: bad shape");
  FK_REQUIRE(D % 4 == 0, "fk_vq_finish: D must be a multiple of 4");
  FK_REQUIRE(xn && embed && cand_val && cand_idx && indices, "fk_vq_finish: null pointer");
  FK_REQUIRE(quantize || !training, "fk_vq_finish: training needs the quantize output");
  FK_REQUIRE(!training || (loss && partials && counter), "fk_vq_finish: training needs loss/partials/counter");
  FinishParams p;
  p.xn = xn; p.embed = embed; p.cand_val = cand_val; p.cand_idx = cand_idx;
  p.indices = indices; p.quantize = quantize; p.partials = partials; p.counter = counter; p.loss = loss;
  p.N = N; p.K = K; p.D = D; p.S = S; p.cosine = use_cosine; p.training = training;
  p.loss_scale = static_cast<float>(static_cast<double>(commitment_weight) / (static_cast<double>(N) * D));
  p.margin_scale = 32.f / 512.f / sqrtf(static_cast<float>(D));
  vq_finish_kernel<<<row_blocks(N), kRowsPerBlock * 32, 0, stream>>>(p);
  FK_CHECK_LAUNCH();
  fk_count_launch();
  return FK_OK;
}

extern "C" __attribute__((visibility("default"))) long long fk_vq_finish_partials(long long N) { return (N + kRowsPerBlock - 1) / kRowsPerBlock; }

extern "C" __attribute__((visibility("default"))) long long fk_vq_ema_stats_ws(long long N, int K) {
  return 3ll * K + N + N / (kSegShort + 1) + 2;      // int32 words: count | offs | cursor | row list | hot codes | their number
}

extern "C" __attribute__((visibility("default"))) int fk_vq_ema_stats(const float* xn, const long long* indices, long long N, int K, int D, float* stats,
                               int* ws, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  FK_REQUIRE(N > 0 && N < (1ll << 31) && K > 0 && K <= (1 << 20) && D > 0 && D % 4 == 0 && D <= 256,
             "fk_vq_ema_stats: bad shape (D % 4 == 0, D <= 256, K <= 2^20)");
  FK_REQUIRE(xn && indices && stats && ws, "fk_vq_ema_stats: null pointer");
  int *count = ws, *offs = ws + K, *cursor = ws + 2 * static_cast<long long>(K), *list = ws + 3 * static_cast<long long>(K);
  if (cudaMemsetAsync(count, 0, static_cast<size_t>(K) * sizeof(int), stream) != cudaSuccess) return FK_ERR_CUDA;
  const unsigned nb = static_cast<unsigned>((N + 255) / 256);
  vq_ema_count_kernel<<<nb, 256, 0, stream>>>(indices, N, K, count);
  FK_CHECK_LAUNCH();
  int* long_list = list + N;
  const long long max_long = N / (kSegShort + 1);
  int* n_long = long_list + max_long + 1;
  vq_ema_scan_kernel<<<1, 1024, 0, stream>>>(count, K, offs, cursor, n_long, long_list);
  FK_CHECK_LAUNCH();
  vq_ema_fill_kernel<<<nb, 256, 0, stream>>>(indices, N, K, offs, cursor, list);
  FK_CHECK_LAUNCH();
  vq_ema_segsum_kernel<<<(K + kRowsPerBlock - 1) / kRowsPerBlock, kRowsPerBlock * 32, 0, stream>>>(xn, count, offs, list, K, D, stats);
  FK_CHECK_LAUNCH();
  if (max_long > 0) {
    vq_ema_segsum_long_kernel<<<static_cast<unsigned>(max_long < K ? max_long : K), 1024, 0, stream>>>(xn, indices, N, count, offs, list, n_long,
                                                                                                  long_list, K, D, stats);
    FK_CHECK_LAUNCH();
  }
  fk_count_launch(6);
  return FK_OK;
}

extern "C" __attribute__((visibility("default"))) int fk_vq_ema_update(const float* stats, float* cluster_size, float* embed_avg, float* embed, int K, int D,
                                int Dp, int Kpad, int use_cosine, float decay, float eps, float threshold,
                                const long long* sample_rows, int n_sample, const float* xn, long long N, void* cb_bf16,
                                float* c2pad, float* total_ws, int* expire_rank_ws, int* n_expired, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  FK_REQUIRE(K > 0 && K <= (1 << 20) && D > 0 && Dp >= D && Kpad >= K, "fk_vq_ema_update: bad shape");
  FK_REQUIRE(stats && cluster_size && embed_avg && embed && cb_bf16 && c2pad && total_ws && expire_rank_ws && n_expired,
             "fk_vq_ema_update: null pointer");
  FK_REQUIRE(threshold <= 0.f || sample_rows == nullptr || (xn && n_sample > 0 && N > 0),
             "fk_vq_ema_update: dead-code reset needs xn and sample rows");
  vq_ema_cluster_kernel<<<1, 1024, 0, stream>>>(stats + static_cast<long long>(K) * D, cluster_size, K, decay, threshold,
                                               total_ws, expire_rank_ws, n_expired);
  FK_CHECK_LAUNCH();
  EmaParams p;
  p.stats = stats; p.cluster_size = cluster_size; p.embed_avg = embed_avg; p.embed = embed; p.total = total_ws;
  p.expire_rank = expire_rank_ws; p.sample_rows = (threshold > 0.f) ? sample_rows : nullptr; p.xn = xn;
  p.cb = static_cast<__nv_bfloat16*>(cb_bf16); p.c2pad = c2pad; p.N = N; p.K = K; p.D = D; p.Dp = Dp; p.Kpad = Kpad;
  p.cosine = use_cosine; p.n_sample = n_sample; p.decay = decay; p.eps = eps; p.threshold = threshold;
  vq_ema_update_kernel<<<row_blocks(Kpad), kRowsPerBlock * 32, 0, stream>>>(p);
  FK_CHECK_LAUNCH();
  fk_count_launch(2);
  return FK_OK;
}

extern "C" __attribute__((visibility("default"))) int fk_vq_backward(const float* g_out, const float* g_loss, const float* xn, const float* quantize,
                              const float* inv_norm, long long N, int D, int use_cosine, float commitment_weight,
                              float* grad_in, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  FK_REQUIRE(N > 0 && D > 0 && D % 4 == 0, "fk_vq_backward: bad shape (D % 4 == 0)");
  FK_REQUIRE(xn && quantize && grad_in, "fk_vq_backward: null pointer");
  FK_REQUIRE(!use_cosine || inv_norm, "fk_vq_backward: cosine needs inv_norm");
  const float coef = static_cast<float>(2.0 * commitment_weight / (static_cast<double>(N) * D));
  vq_backward_kernel<<<row_blocks(N), kRowsPerBlock * 32, 0, stream>>>(g_out, g_loss, xn, quantize, inv_norm, N, D,
                                                                       use_cosine, coef, grad_in);
  FK_CHECK_LAUNCH();
  fk_count_launch();
  return FK_OK;
}

extern "C" __attribute__((visibility("default"))) int fk_vq_kmeans_update(const float* stats, float* means, int K, int D, int Dp, int Kpad, int use_cosine,
                                   void* cb_bf16, float* c2pad, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  FK_REQUIRE(K > 0 && D > 0 && Dp >= D && Kpad >= K, "fk_vq_kmeans_update: bad shape");
  FK_REQUIRE(stats && means && cb_bf16 && c2pad, "fk_vq_kmeans_update: null pointer");
  vq_kmeans_update_kernel<<<row_blocks(Kpad), kRowsPerBlock * 32, 0, stream>>>(
      stats, means, K, D, Dp, Kpad, use_cosine, static_cast<__nv_bfloat16*>(cb_bf16), c2pad);
  FK_CHECK_LAUNCH();
  fk_count_launch();
  return FK_OK;
}

extern "C" __attribute__((visibility("default"))) int fk_vq_perplexity(const long long* indices, long long N, int K, float* bins_ws, float* out, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  FK_REQUIRE(N > 0 && K > 0 && indices && bins_ws && out, "fk_vq_perplexity: bad argument");
  if (cudaMemsetAsync(bins_ws, 0, static_cast<size_t>(K) * sizeof(float), stream) != cudaSuccess) return FK_ERR_CUDA;
  vq_histogram_kernel<<<static_cast<unsigned>((N + 255) / 256), 256, 0, stream>>>(indices, N, K, bins_ws);
  FK_CHECK_LAUNCH();
  vq_perplexity_kernel<<<1, 1024, 0, stream>>>(bins_ws, K, 1.f / static_cast<float>(N), out);
  FK_CHECK_LAUNCH();
  fk_count_launch(3);
  return FK_OK;
}
