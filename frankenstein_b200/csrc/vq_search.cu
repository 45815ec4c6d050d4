// Nearest-codeword search: S = X * C^T on the 5th-gen tensor cores (tcgen05.mma, bf16 operands,
// fp32 accumulators in TMEM), operands staged by TMA, and the per-row arg-extremum taken in the
// epilogue straight out of TMEM -- the [N, K] similarity matrix never exists in memory.
//
// Replaces (reference call site models/vq_brain.py:209 -> vector_quantize_pytorch
// EuclideanCodebook/CosineSimCodebook.forward): `dist = -cdist(x, embed)` / `einsum('h n d, h c d -> h n c')`
// followed by `dist.argmax(-1)`; SURVEY.md section 2a rows K1+K2.
//
// Score convention: the kernel MINIMISES  v[n][k] = alpha * (x_n . c_k) + c2[k]
//   Euclidean: alpha = -2, c2[k] = |c_k|^2   (|x|^2 is row-constant, sqrt/clamp are monotone)
//   cosine   : alpha = -1, c2[k] = 0         (argmax of the similarity)
// c2 is padded to a multiple of BN with +inf so that out-of-range codes can never win.
// Because bf16 operands perturb near-ties, the kernel does not commit to one index: every epilogue
// thread keeps a running minimum per column class (32 classes = column mod 32, 3 instructions per
// element, no branches) with the tile id packed into the low mantissa bits of the key, and writes
// its 4 best classes per work segment.  fk_vq_finish merges the candidates of a row, re-scores the
// ones inside the bf16 error margin in exact fp32 and applies the reference's tie rule (lowest
// index wins).
//
// Tiling: CTA tile = 256 rows (two M=128 accumulators) x 128 codes, K-depth = D (<= 256, whole
// depth resident for X).  A work unit is one (row block, code tile) pair; the W = RB*T units are
// cut into gridDim.x contiguous ranges (persistent CTAs, perfect balance to within one unit), so a
// CTA covers at most a few row blocks ("segments") and writes one candidate slot per segment.
//
// Warp roles (640 threads): warp 0 = TMA producer, warps 1 and 3 = MMA issuers (one thread each, one per
// accumulator half), warp 2 = TMEM allocator, warps 4-19 = epilogue (TMEM lane quarter x accumulator half x column half).
#include "common.cuh"
#include "tma_host.cuh"

namespace fk {

constexpr int kBM = 256;          // rows per CTA tile (2 x UMMA_M)
constexpr int kBN = 128;          // codes per tile (UMMA_N)
constexpr int kSlabK = 64;        // bf16 elements per 128-byte swizzled row
constexpr int kXSlabBytes = kBM * 128;   // 32 KB
constexpr int kBSlabBytes = kBN * 128;   // 16 KB
constexpr int kSearchThreads = 640;       // 4 control warps + 16 epilogue warps
constexpr int kEpilogueThreads = 512;
constexpr int kMaxStages = 8;

struct SearchParams {
  const float* c2pad;   // [T * kBN]
  float* cand_val;      // [N][S][kCand] keys (approximate scores)
  int* cand_idx;        // [N][S][kCand] code indices, -1 = none
  uint32_t tag_mask;    // low mantissa bits that carry the tile tag
  float* dbg_scores;    // optional [N][T*kBN] raw accumulators (tests only)
  long long* prof;      // optional [grid][16] stall counters in cycles (kProf instantiation only)
  long long N;
  int K, nslab, T, RB, S, nstage;
  float alpha;
};

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr) : "memory");
}

__device__ __forceinline__ void tmem_wait_ld_dep(uint32_t (&r)[16]) {
  // tcgen05.wait::ld with the destination registers as in/out operands, so that no use of r[] can
  // be scheduled above the wait.
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                 "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]),
                 "+r"(r[15])
               :: "memory");
}

constexpr int kCand = 4;          // candidates written per (row, slot)

// 16 accumulator columns of this thread's row -> running class minima.  key = score with the low
// `tag` bits of the mantissa replaced by the tile tag (relative precision loss <= 2^-13, far below
// the bf16 operand noise); +inf padding columns turn into NaN keys, which fminf ignores.
__device__ __forceinline__ void scan16(const uint32_t (&r)[16], const float* c2s, float alpha, uint32_t keep_mask,
                                       uint32_t tag, float (&m)[32], int class_off, float* dbg) {
  if (dbg != nullptr) {
#pragma unroll
    for (int i = 0; i < 16; ++i) dbg[i] = __uint_as_float(r[i]);
  }
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    const float4 c = *reinterpret_cast<const float4*>(c2s + g * 4);
    const float cc[4] = {c.x, c.y, c.z, c.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float v = fmaf(__uint_as_float(r[g * 4 + e]), alpha, cc[e]);
      const float key = __uint_as_float((__float_as_uint(v) & keep_mask) | tag);
      m[class_off + g * 4 + e] = fminf(m[class_off + g * 4 + e], key);
    }
  }
}

// Stall accounting for the profiling instantiation: cycles spent inside a barrier wait are added to `acc`.
template <bool kProf>
__device__ __forceinline__ void wait_acc(uint64_t* bar, uint32_t phase, long long& acc) {
  if constexpr (kProf) {
    // only waits that were NOT already satisfied are charged (count in the high 24 bits, cycles below)
    if (mbar_test_wait(bar, phase)) return;
    const long long t = clock64();
    mbar_wait(bar, phase);
    acc += (clock64() - t) + (1ll << 40);
  } else {
    mbar_wait(bar, phase);
  }
}

// prof[cta][32]: 16+h issuer wait+fence+descriptor cycles, 18+h MMA issue cycles, 20+h commit cycles; 0 total, 1 setup, 2 producer wait empty, 3 producer wait x_empty, 4+h issuer wait x_full,
// 6+h issuer wait tmem_empty, 8+h issuer wait full, 10+h issuer loop total, 12 epilogue(warp 4) wait tmem_full,
// 13 epilogue flush, 14 epilogue named-barrier wait, 15 first accumulator ready (since start)
template <bool kProf>
__global__ void __launch_bounds__(kSearchThreads, 1)
vq_search_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_c,
                 const SearchParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // 1024-byte alignment is required by the 128B swizzle (TMA destination and UMMA descriptors).
  if ((smem_u32(smem_raw) & 1023u) != 0) __trap();
  uint8_t* smem = smem_raw;
  uint8_t* Xs = smem;                                   // nslab x 32 KB
  uint8_t* Bs = Xs + p.nslab * kXSlabBytes;             // nstage x 16 KB
  float* c2s = reinterpret_cast<float*>(Bs + p.nstage * kBSlabBytes);   // [2][kBN]
  uint64_t* bars = reinterpret_cast<uint64_t*>(c2s + 2 * kBN);
  uint64_t* full = bars;                    // [kMaxStages]
  uint64_t* empty = bars + kMaxStages;      // [kMaxStages]
  uint64_t* x_full = bars + 2 * kMaxStages;  // [4] one per k-slab of X, so the first MMAs start after 32 KB, not 128 KB
  uint64_t* x_empty = x_full + 4;
  uint64_t* tmem_full = x_full + 5;         // [2 stages][2 halves]
  uint64_t* tmem_empty = x_full + 9;        // [2 stages][2 halves]
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(x_full + 13);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long t_start = kProf ? clock64() : 0;
  long long* prof = kProf ? p.prof + static_cast<long long>(blockIdx.x) * 32 : nullptr;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_x);
    tma_prefetch_desc(&tmap_c);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < p.nstage; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 2);        // both MMA issuers (one per accumulator half) release a slab
    }
    for (int i = 0; i < 4; ++i) mbar_init(&x_full[i], 1);
    mbar_init(x_empty, 2);
    for (int i = 0; i < 4; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 8);   // one arrival per epilogue warp of that half
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_base_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_slot;
  if (kProf && threadIdx.x == 0) prof[1] = clock64() - t_start;

  // contiguous range of work units for this CTA
  const long long W = static_cast<long long>(p.RB) * p.T;
  const long long G = gridDim.x;
  const long long u_begin = (static_cast<long long>(blockIdx.x) * W) / G;
  const long long u_end = (static_cast<long long>(blockIdx.x + 1) * W) / G;

  if (warp == 0) {
    // ================================ TMA producer ================================
    // (whole warp runs the loop, one elected lane issues -- see the MMA issuers below)
    {
      int stage = 0;
      uint32_t phase = 0, seg = 0;
      long long w_empty = 0, w_xempty = 0;
      for (long long u = u_begin; u < u_end; ++seg) {
        const int rb = static_cast<int>(u / p.T), t0 = static_cast<int>(u % p.T);
        const int nt = static_cast<int>(min(static_cast<long long>(p.T - t0), u_end - u));
        wait_acc<kProf>(x_empty, (seg & 1) ^ 1, w_xempty);
        if (elect_one()) {
          for (int ks = 0; ks < p.nslab; ++ks) {
            mbar_expect_tx(&x_full[ks], kXSlabBytes);
            tma_load_2d(Xs + ks * kXSlabBytes, &tmap_x, &x_full[ks], ks * kSlabK, rb * kBM);
          }
        }
        __syncwarp();
        for (int t = t0; t < t0 + nt; ++t) {
          for (int ks = 0; ks < p.nslab; ++ks) {
            wait_acc<kProf>(&empty[stage], phase ^ 1, w_empty);
            if (elect_one()) {
              mbar_expect_tx(&full[stage], kBSlabBytes);
              tma_load_2d(Bs + stage * kBSlabBytes, &tmap_c, &full[stage], ks * kSlabK, t * kBN);
            }
            __syncwarp();
            if (++stage == p.nstage) { stage = 0; phase ^= 1; }
          }
        }
        u += nt;
      }
      if (kProf && lane == 0) { prof[2] = w_empty; prof[3] = w_xempty; }
    }
  } else if (warp == 1 || warp == 3) {
    // ================================ MMA issuers ================================
    // Two issuing threads, one per accumulator half (rows 0-127 / 128-255): measured on B200 one thread sustains
    // one tcgen05.mma per ~100 cycles while the tensor core accepts an M128 x N128 x K16 MMA every 64
    // (scripts/microbench/umma_rate.cu), so a single issuer caps the kernel at ~60 % of the MMA rate.
    // The WHOLE warp runs this loop (barrier waits by all lanes, one elected lane issues): with warp-uniform control
    // flow and operands ptxas keeps the descriptors in uniform registers and emits bare UTCHMMA instructions.  A
    // `lane == 0` branch around the loop made every MMA a divergence "waterfall" (ELECT + 5x R2UR.BROADCAST + branch,
    // ~115 cycles per issue; scripts/gpu_search_stalls.py).
    {
      const int h = (__shfl_sync(0xffffffffu, warp, 0) == 1) ? 0 : 1;
      const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
      constexpr uint32_t idesc = umma_idesc_bf16(128, kBN);
      const uint32_t xs_addr = smem_u32(Xs) + h * (128 * 128), bs_addr = smem_u32(Bs);
      int stage = 0;
      uint32_t phase = 0, seg = 0, tc = 0;
      long long w_x = 0, w_te = 0, w_full = 0, c_pre = 0, c_mma = 0, c_commit = 0;
      const long long t_loop = kProf ? clock64() : 0;
      for (long long u = u_begin; u < u_end; ++seg) {
        const int t0 = static_cast<int>(u % p.T);
        const int nt = static_cast<int>(min(static_cast<long long>(p.T - t0), u_end - u));
        for (int t = 0; t < nt; ++t, ++tc) {
          const uint32_t as = tc & 1;
          wait_acc<kProf>(&tmem_empty[as * 2 + h], ((tc >> 1) & 1) ^ 1, w_te);
          tc_fence_after();
          const uint32_t d_tmem = tmem_u + (h * 2 + as) * kBN;
          for (int ks = 0; ks < p.nslab; ++ks) {
            const long long c0 = kProf ? clock64() : 0;
            if (t == 0) wait_acc<kProf>(&x_full[ks], seg & 1, w_x);      // X slab ks of this segment has landed
            wait_acc<kProf>(&full[stage], phase, w_full);
            tc_fence_after();
            const uint64_t adesc = umma_desc_sw128(xs_addr + ks * kXSlabBytes);
            const uint64_t bdesc = umma_desc_sw128(bs_addr + stage * kBSlabBytes);
            const long long c1 = kProf ? clock64() : 0;
            if (elect_one()) {
#pragma unroll
              for (int kk = 0; kk < 4; ++kk)      // a K step of 16 bf16 = 32 bytes = +2 in the (address >> 4) field
                umma_bf16(d_tmem, adesc + 2 * kk, bdesc + 2 * kk, idesc, (ks | kk) != 0);
            }
            const long long c2 = kProf ? clock64() : 0;
            if (elect_one()) umma_commit(&empty[stage]);   // slab free once the MMAs of both halves have retired (2 arrivals)
            __syncwarp();
            if (kProf) { const long long c3 = clock64(); c_pre += c1 - c0; c_mma += c2 - c1; c_commit += c3 - c2; }
            if (++stage == p.nstage) { stage = 0; phase ^= 1; }
          }
          if (elect_one()) umma_commit(&tmem_full[as * 2 + h]);
          __syncwarp();
        }
        if (elect_one()) umma_commit(x_empty);
        __syncwarp();
        u += nt;
      }
      if (kProf && lane == 0) { prof[4 + h] = w_x; prof[6 + h] = w_te; prof[8 + h] = w_full; prof[10 + h] = clock64() - t_loop;
                   prof[16 + h] = c_pre; prof[18 + h] = c_mma; prof[20 + h] = c_commit; }
    }
  } else if (warp >= 4) {
    // ================================ epilogue ================================
    // 16 warps: q = TMEM lane quarter (fixed by warp id % 4), h = accumulator half (rows 0-127 / 128-255),
    // cpart = which 64 of the tile's 128 columns this warp scans.  Every (row, cpart) keeps its own
    // top-2 and writes its own candidate slot, so no cross-warp merge is needed here.
    const int e = warp - 4;
    const int q = e & 3, h = (e >> 2) & 1, cpart = e >> 3;
    const int etid = threadIdx.x - 128;                  // 0..511
    const float inf = __int_as_float(0x7f800000);
    float m[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) m[i] = inf;
    const uint32_t keep_mask = ~p.tag_mask;
    // c2 of the first tile; later tiles are prefetched one tile ahead into the other buffer
    if (etid < kBN && u_begin < u_end) c2s[etid] = __ldg(p.c2pad + (u_begin % p.T) * kBN + etid);
    named_bar_sync(1, kEpilogueThreads);
    uint32_t tc = 0;
    long long w_tf = 0, w_flush = 0, w_bar = 0;
    int t_seg0 = static_cast<int>(u_begin % p.T);        // first tile of the current segment
    for (long long u = u_begin; u < u_end; ++u, ++tc) {
      const int rb = static_cast<int>(u / p.T), t = static_cast<int>(u % p.T);
      const uint32_t as = tc & 1;
      float c2_next = 0.f;
      const bool have_next = (u + 1 < u_end) && etid < kBN;
      if (have_next) c2_next = __ldg(p.c2pad + ((u + 1) % p.T) * kBN + etid);
      const float* my_c2 = c2s + as * kBN + cpart * 64;
      const uint32_t tag = static_cast<uint32_t>(t - t_seg0) * 2u;
      wait_acc<kProf>(&tmem_full[as * 2 + h], (tc >> 1) & 1, w_tf);
      if (kProf && tc == 0 && warp == 4 && lane == 0) prof[15] = clock64() - t_start;
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + (h * 2 + as) * kBN + cpart * 64;
      const long long row = static_cast<long long>(rb) * kBM + h * 128 + q * 32 + lane;
      float* dbg = (p.dbg_scores != nullptr && row < p.N)
                       ? p.dbg_scores + row * (static_cast<long long>(p.T) * kBN) + static_cast<long long>(t) * kBN + cpart * 64
                       : nullptr;
      uint32_t ra[16], rb2[16];
      tmem_ld16(taddr, ra);
      tmem_wait_ld_dep(ra);
      tmem_ld16(taddr + 16, rb2);
      scan16(ra, my_c2, p.alpha, keep_mask, tag, m, 0, dbg);
      tmem_wait_ld_dep(rb2);
      tmem_ld16(taddr + 32, ra);
      scan16(rb2, my_c2 + 16, p.alpha, keep_mask, tag, m, 16, dbg ? dbg + 16 : nullptr);
      tmem_wait_ld_dep(ra);
      tmem_ld16(taddr + 48, rb2);
      scan16(ra, my_c2 + 32, p.alpha, keep_mask, tag + 1u, m, 0, dbg ? dbg + 32 : nullptr);
      tmem_wait_ld_dep(rb2);
      // all TMEM reads of this accumulator stage are done: hand it back before the last scan
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[as * 2 + h]);
      scan16(rb2, my_c2 + 48, p.alpha, keep_mask, tag + 1u, m, 16, dbg ? dbg + 48 : nullptr);
      if (have_next) c2s[(as ^ 1) * kBN + etid] = c2_next;
      const long long t_bar = kProf ? clock64() : 0;
      named_bar_sync(1, kEpilogueThreads);      // c2 of the next tile visible; this tile's c2 reads finished
      if (kProf) w_bar += clock64() - t_bar;
      const long long t_flush = kProf ? clock64() : 0;
      if (t == p.T - 1 || u + 1 == u_end) {
        // ---- end of a segment: the kCand smallest class keys of this (row, cpart) -> its candidate slot ----
        float cv[kCand];
        int ci[kCand];
#pragma unroll
        for (int r = 0; r < kCand; ++r) {
          float mn = m[0];
#pragma unroll
          for (int i = 1; i < 32; ++i) mn = fminf(mn, m[i]);
          int cls = -1;
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const bool hit = (m[i] == mn) && (cls < 0);
            cls = hit ? i : cls;
            m[i] = hit ? inf : m[i];
          }
          const uint32_t tg = __float_as_uint(mn) & p.tag_mask;
          // class -> column: classes 0-15 come from loads 0/2, 16-31 from loads 1/3; tag bit 0 = upper 32 columns
          const int col = cpart * 64 + static_cast<int>(tg & 1u) * 32 + cls;
          cv[r] = mn;
          ci[r] = (mn < inf && cls >= 0) ? (t_seg0 + static_cast<int>(tg >> 1)) * kBN + col : -1;
        }
        const long long ufirst = static_cast<long long>(rb) * p.T;
        const int first_cta = static_cast<int>(((ufirst + 1) * G - 1) / W);
        const int slot = (static_cast<int>(blockIdx.x) - first_cta) * 2 + cpart;
        if (row < p.N && slot >= 0 && slot < p.S) {
          const long long o = (row * p.S + slot) * kCand;
          *reinterpret_cast<float4*>(p.cand_val + o) = make_float4(cv[0], cv[1], cv[2], cv[3]);
          *reinterpret_cast<int4*>(p.cand_idx + o) = make_int4(ci[0], ci[1], ci[2], ci[3]);
          if (t == p.T - 1) {
            // the CTA that finishes a row block also marks the slots no CTA owns as "no candidate" (row blocks
            // touched by fewer CTAs than the widest one), so the host never has to clear cand_idx
            for (int s2 = slot + 2; s2 < p.S; s2 += 2)
              *reinterpret_cast<int4*>(p.cand_idx + (row * p.S + s2) * kCand) = make_int4(-1, -1, -1, -1);
          }
        }
#pragma unroll
        for (int i = 0; i < 32; ++i) m[i] = inf;
        t_seg0 = 0;                                   // the next segment starts a new row block at tile 0
        if (kProf) w_flush += clock64() - t_flush;
      }
    }
    if (kProf && warp == 4 && lane == 0) { prof[12] = w_tf; prof[13] = w_flush; prof[14] = w_bar; }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
  if (kProf && threadIdx.x == 0) prof[0] = clock64() - t_start;
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
static int sm_count() { return fk_sm_count(); }

}  // namespace fk

using namespace fk;

extern "C" __attribute__((visibility("default"))) int fk_vq_search_slots(long long N, int K, int max_ctas) {
  if (N <= 0 || K <= 0) return FK_ERR_BAD_ARG;
  const long long RB = (N + kBM - 1) / kBM, T = (K + kBN - 1) / kBN, W = RB * T;
  long long G = max_ctas > 0 ? max_ctas : 148;
  if (G > W) G = W;
  int S = 1;
  for (long long rb = 0; rb < RB; ++rb) {
    const long long first = ((rb * T + 1) * G - 1) / W;
    const long long last = (((rb + 1) * T - 1 + 1) * G - 1) / W;
    if (last - first + 1 > S) S = static_cast<int>(last - first + 1);
  }
  return 2 * S;   // two column halves per segment
}

static int search_launch(const void* x_bf16, const void* cb_bf16, const float* c2pad, long long N, int K, int Dp,
                         int use_cosine, float* cand_val, int* cand_idx, int S, int max_ctas, float* dbg_scores,
                         long long* prof, void* stream_);

extern "C" __attribute__((visibility("default"))) int fk_vq_search(const void* x_bf16, const void* cb_bf16, const float* c2pad, long long N, int K, int Dp,
                            int use_cosine, float* cand_val, int* cand_idx, int S, int max_ctas, void* stream_) {
  return search_launch(x_bf16, cb_bf16, c2pad, N, K, Dp, use_cosine, cand_val, cand_idx, S, max_ctas, nullptr, nullptr, stream_);
}

extern "C" __attribute__((visibility("default"))) int fk_vq_search_debug(const void* x_bf16, const void* cb_bf16, const float* c2pad, long long N, int K, int Dp,
                                  int use_cosine, float* cand_val, int* cand_idx, int S, int max_ctas, float* dbg_scores,
                                  void* stream_) {
  return search_launch(x_bf16, cb_bf16, c2pad, N, K, Dp, use_cosine, cand_val, cand_idx, S, max_ctas, dbg_scores, nullptr, stream_);
}

extern "C" __attribute__((visibility("default"))) int fk_vq_search_profile(const void* x_bf16, const void* cb_bf16, const float* c2pad, long long N, int K, int Dp,
                                    int use_cosine, float* cand_val, int* cand_idx, int S, int max_ctas, long long* prof,
                                    void* stream_) {
  FK_REQUIRE(prof != nullptr, "fk_vq_search_profile: null counter buffer");
  return search_launch(x_bf16, cb_bf16, c2pad, N, K, Dp, use_cosine, cand_val, cand_idx, S, max_ctas, nullptr, prof, stream_);
}

static int search_launch(const void* x_bf16, const void* cb_bf16, const float* c2pad, long long N, int K, int Dp,
                         int use_cosine, float* cand_val, int* cand_idx, int S, int max_ctas, float* dbg_scores,
                         long long* prof, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  FK_REQUIRE(N > 0 && K > 0, "fk_vq_search: empty problem");
  FK_REQUIRE(Dp % 64 == 0 && Dp >= 64 && Dp <= 256, "fk_vq_search: padded dim must be 64, 128, 192 or 256");
  FK_REQUIRE(x_bf16 && cb_bf16 && c2pad && cand_val && cand_idx, "fk_vq_search: null pointer");
  const long long RB = (N + kBM - 1) / kBM, T = (K + kBN - 1) / kBN, W = RB * T;
  FK_REQUIRE(RB < (1ll << 30) && T < (1 << 24), "fk_vq_search: problem too large");
  long long G = max_ctas > 0 ? max_ctas : sm_count();
  if (G > W) G = W;
  FK_REQUIRE(S == fk_vq_search_slots(N, K, static_cast<int>(G)), "fk_vq_search: slot count does not match fk_vq_search_slots");

  CUtensorMap tx, tcm;
  int rc = make_tmap_bf16_sw128(&tx, x_bf16, static_cast<uint64_t>(N), static_cast<uint64_t>(Dp), kBM);
  if (rc != FK_OK) { fk_set_last_error("cuTensorMapEncodeTiled(x) failed", __FILE__, __LINE__); return rc; }
  rc = make_tmap_bf16_sw128(&tcm, cb_bf16, static_cast<uint64_t>(K), static_cast<uint64_t>(Dp), kBN);
  if (rc != FK_OK) { fk_set_last_error("cuTensorMapEncodeTiled(codebook) failed", __FILE__, __LINE__); return rc; }

  SearchParams p;
  p.c2pad = c2pad;
  p.cand_val = cand_val;
  p.cand_idx = cand_idx;
  p.dbg_scores = dbg_scores;
  p.prof = prof;
  p.N = N;
  p.K = K;
  p.nslab = Dp / 64;
  p.T = static_cast<int>(T);
  p.RB = static_cast<int>(RB);
  p.S = S;
  p.alpha = use_cosine ? -1.f : -2.f;
  {
    uint32_t need = static_cast<uint32_t>(2 * T), bits = 1;
    while ((1u << bits) < need) ++bits;
    FK_REQUIRE(bits <= 12, "fk_vq_search: codebook too large for the packed tile tag (K <= 262144)");
    p.tag_mask = (1u << bits) - 1u;
  }
  const int fixed = p.nslab * kXSlabBytes + 2 * kBN * 4 + 256 /*barriers*/;
  int nstage = (232448 - 1024 /*static*/ - fixed) / kBSlabBytes;
  if (nstage > kMaxStages) nstage = kMaxStages;
  p.nstage = nstage;
  const int smem_bytes = fixed + nstage * kBSlabBytes;

  static bool attr_set_dev[FK_MAX_DEVICES];
  bool& attr_set = attr_set_dev[fk_device_ordinal()];
  if (!attr_set) {
    if (cudaFuncSetAttribute(vq_search_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448) != cudaSuccess ||
        cudaFuncSetAttribute(vq_search_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448) != cudaSuccess) {
      fk_set_last_error("cudaFuncSetAttribute(max dynamic smem) failed", __FILE__, __LINE__);
      return FK_ERR_CUDA;
    }
    attr_set = true;
  }
  if (prof != nullptr)
    vq_search_kernel<true><<<static_cast<unsigned>(G), kSearchThreads, smem_bytes, stream>>>(tx, tcm, p);
  else
    vq_search_kernel<false><<<static_cast<unsigned>(G), kSearchThreads, smem_bytes, stream>>>(tx, tcm, p);
  FK_CHECK_LAUNCH();
  fk_count_launch(1);
  return FK_OK;
}
