// Hand-written sm_100a GEMM for the dense DLRM arch: C[M,N] = act(A[M,K] . B[N,K]^T + bias) with
// bf16 operands staged by TMA (SWIZZLE_128B), tcgen05.mma (cta_group::1, M=128) accumulating in TMEM
// and a fused epilogue (bias, ReLU / sigmoid, ReLU-gradient mask, bf16 or fp32 store).
//
// Structure (one persistent CTA per SM, 8 warps):
//   warp 0      TMA producer      cp.async.bulk.tensor.2d -> smem ring (kStages x (A 128x64, B BNx64))
//   warp 1      MMA issuer        one elected lane issues 4 x tcgen05.mma (UMMA_K=16) per k-block,
//                                 tcgen05.commit frees the smem slot / publishes the accumulator
//   warp 2      TMEM allocator    2 accumulator stages (2 x BLOCK_N fp32 columns)
//   warps 4..7  epilogue          tcgen05.ld 32x32b (one TMEM lane == one output row per thread),
//                                 bias + activation in registers, vectorised global stores
// The accumulator is double buffered so the epilogue of tile i overlaps the MMAs of tile i+1.
//
// Both operands are K-major (row-major [rows, K]), which is the natural layout of nn.Linear
// (x [M,K], weight [N,K]); dgrad / wgrad are expressed with pre-transposed operands.
// Replaces the reference's cuBLAS nn.Linear + separate bias/ReLU kernels (modules/mlp.py:18-190).
#include <cstdlib>

#include "tcgen05.cuh"

namespace {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;   // 64 bf16 = 128 B = one SWIZZLE_128B row
constexpr int UMMA_K = 16;
constexpr int kNumThreads = 256;
constexpr int kEpiWarp0 = 4;

using namespace trb;

enum Act : int { ACT_NONE = 0, ACT_RELU = 1, ACT_SIGMOID = 2, ACT_RELU_GRAD = 3 };

struct GemmParams {
  int M, N, K;
  const float* bias;            // [N] or nullptr
  const __nv_bfloat16* mask;    // ACT_RELU_GRAD: out *= (mask[m, n] > 0)
  int64_t ld_mask;
  void* out;
  int64_t ldo;
  int out_f32;                  // 1: fp32 output, 0: bf16
  int act;
  float alpha;                  // scale applied to the accumulator before bias
  int split_k;                  // >1: K range split over CTAs, fp32 atomicAdd epilogue (out pre-zeroed)
  int tma_store;                // 1: bf16 output leaves through smem + TMA stores (tmap_o valid)
  // 1 bit per element instead of re-reading the bf16 activation as the ReLU-gradient mask (67 MB -> 4 MB for 32768 x 1024):
  uint32_t* relu_bits_out;      // ACT_RELU: bit (n % 32) of word [m, n / 32] = (out[m, n] > 0); nullptr: not wanted
  const uint32_t* mask_bits;    // ACT_RELU_GRAD: the same words of the layer whose gradient this is; nullptr: use `mask`
  int64_t ld_bits;              // words per row of either bit matrix
  // bias gradient of the layer that produced this GEMM's input gradient, folded into the TMA epilogue: every epilogue warp sums the
  // 32 rows of each [32 x 64] output slab it just wrote to smem and stores the fp32 column sums to colsum_ws[row0 / 32][n]
  // ([ceil(M / 32), N], deterministic; reduced by trb_colsum_partials). nullptr: not wanted.
  float* colsum_ws;
};

// ---- bf16 epilogue through shared memory + TMA stores ---------------------------------------------------------------
// The direct epilogue (one thread = one output row, 16 B stores 2 KB apart) costs ~6.4 us per 128 x 256 tile, more than the
// MMAs of a K <= 1024 tile: every fwd / dgrad GEMM of the DLRM MLPs was epilogue bound (tools/microbench.py gemmx). Here each
// epilogue warp owns a [32 rows x 64 cols] SWIZZLE_128B slab (double buffered): TMEM -> registers (two tcgen05.ld in flight)
// -> bias / activation / mask -> 16 B swizzled st.shared (conflict free) -> one TMA store per slab, which also clips the
// M / N tails. Bias lives in a per-warp smem copy; the ReLU-gradient mask tile is fetched with coalesced 16 B loads.
constexpr int kEpiSlabBytes = 32 * 128;                                   // [32 rows x 64 bf16]
constexpr int kEpiBytesPerWarp = 3 * kEpiSlabBytes + 256 * 4;             // 2 store slabs + mask slab + bias[256]
constexpr int kEpiBytes = 4 * kEpiBytesPerWarp;                           // 53248 B

__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

// ACT / BIAS are compile-time: with runtime checks inside the unrolled 8-piece loop the epilogue warps (one per scheduler, no
// other warp to hide latency) spent their time in branch resolution and instruction-cache misses (ncu source view).
template <int BN, int ACT, bool BIAS>
__device__ __forceinline__ void epilogue_bf16_tma_impl(const CUtensorMap* tmap_o, const GemmParams& p, uint32_t tmem_acc, int row0, int n_base, int lane,
                                                       uint8_t* epi_warp, uint32_t& store_count) {
  // plain C++ shared accesses on purpose: `asm volatile` ld/st.shared are kept in program order by the compiler, which
  // serialised the 8 global loads of the mask tile (one HBM round trip each) and cost 0.36 ms per train step
  uint8_t* store_slab = epi_warp;
  uint8_t* mask_slab = epi_warp + 2 * kEpiSlabBytes;
  float* bias_s = reinterpret_cast<float*>(epi_warp + 3 * kEpiSlabBytes);
  if constexpr (BIAS) {
#pragma unroll
    for (int i = lane * 4; i < BN; i += 128) {
      float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
      if (n_base + i < p.N) b4 = __ldg(reinterpret_cast<const float4*>(p.bias + n_base + i));
      *reinterpret_cast<float4*>(bias_s + i) = b4;
    }
    __syncwarp();
  }
  const uint32_t sw = (uint32_t) (lane & 7);
  const bool scale = p.alpha != 1.f;
  // ReLU-gradient mask: either one bit per element written by the forward epilogue of that layer (2 words per row and group,
  // no smem staging) or, as fallback, the bf16 activation itself: its [32 rows x 64 cols] tile of group g+1 is fetched (8 lanes
  // per row, full 128 B lines, 8 loads in flight) while group g is processed.
  const bool use_bits = (ACT == ACT_RELU_GRAD) && p.mask_bits != nullptr;
  const int my_row = row0 + lane;
  uint4 m[8];
  auto load_mask = [&](int n0) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int grow = row0 + i * 4 + (lane >> 3), col = n0 + (lane & 7) * 8;
      m[i] = make_uint4(0u, 0u, 0u, 0u);
      if (grow < p.M && col < p.N) m[i] = __ldg(reinterpret_cast<const uint4*>(p.mask + (int64_t) grow * p.ld_mask + col));
    }
  };
  if constexpr (ACT == ACT_RELU_GRAD) {
    if (!use_bits) load_mask(n_base);
  }
#pragma unroll 1
  for (int g = 0; g < BN / 64; ++g) {
    const int n0 = n_base + 64 * g;
    if (n0 >= p.N) break;
    uint32_t r0[32], r1[32];
    tmem_ld_32x32_nowait(tmem_acc + (uint32_t) (64 * g), r0);
    tmem_ld_32x32_nowait(tmem_acc + (uint32_t) (64 * g + 32), r1);
    uint32_t mbits0 = 0u, mbits1 = 0u;
    if constexpr (ACT == ACT_RELU_GRAD) {
      if (use_bits) {
        if (my_row < p.M) {
          const uint32_t* bw = p.mask_bits + (int64_t) my_row * p.ld_bits + (n0 >> 5);
          mbits0 = __ldg(bw);
          if (n0 + 32 < p.N) mbits1 = __ldg(bw + 1);
        }
      } else {
        // (all lanes finished reading the previous group's mask at the __syncwarp that ended the previous iteration)
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int rr = i * 4 + (lane >> 3), piece = lane & 7;
          *reinterpret_cast<uint4*>(mask_slab + rr * 128 + ((piece ^ (rr & 7)) << 4)) = m[i];
        }
        if (g + 1 < BN / 64 && n0 + 64 < p.N) load_mask(n0 + 64);
      }
    }
    uint32_t obits0 = 0u, obits1 = 0u;
    // the TMA store that read this slab two groups ago must be done with it
    if (lane == 0) bulk_wait_group_read<1>();
    __syncwarp();
    tmem_ld_wait();
    uint8_t* slab = store_slab + (store_count & 1u) * kEpiSlabBytes;
#pragma unroll
    for (int j = 0; j < 8; ++j) {  // 8 columns per 16 B piece
      float v[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) v[q] = __uint_as_float(j < 4 ? r0[j * 8 + q] : r1[(j - 4) * 8 + q]);
      if (scale) {
#pragma unroll
        for (int q = 0; q < 8; ++q) v[q] *= p.alpha;
      }
      if constexpr (BIAS) {
        const float4 b0 = *reinterpret_cast<const float4*>(bias_s + 64 * g + j * 8);
        const float4 b1 = *reinterpret_cast<const float4*>(bias_s + 64 * g + j * 8 + 4);
        v[0] += b0.x; v[1] += b0.y; v[2] += b0.z; v[3] += b0.w;
        v[4] += b1.x; v[5] += b1.y; v[6] += b1.z; v[7] += b1.w;
      }
      if constexpr (ACT == ACT_RELU) {
#pragma unroll
        for (int q = 0; q < 8; ++q) v[q] = fmaxf(v[q], 0.f);
        if (p.relu_bits_out != nullptr) {
          uint32_t b8 = 0u;
#pragma unroll
          for (int q = 0; q < 8; ++q) b8 |= (v[q] > 0.f ? 1u : 0u) << q;
          if (j < 4) obits0 |= b8 << (8 * j); else obits1 |= b8 << (8 * (j - 4));
        }
      } else if constexpr (ACT == ACT_SIGMOID) {
#pragma unroll
        for (int q = 0; q < 8; ++q) v[q] = 1.f / (1.f + __expf(-v[q]));
      } else if constexpr (ACT == ACT_RELU_GRAD) {
        if (use_bits) {
          const uint32_t b8 = (j < 4 ? (mbits0 >> (8 * j)) : (mbits1 >> (8 * (j - 4)))) & 0xFFu;
#pragma unroll
          for (int q = 0; q < 8; ++q) v[q] = ((b8 >> q) & 1u) ? v[q] : 0.f;
        } else {
          const uint4 m8 = *reinterpret_cast<const uint4*>(mask_slab + lane * 128 + (((uint32_t) j ^ sw) << 4));
          const __nv_bfloat16* mb = reinterpret_cast<const __nv_bfloat16*>(&m8);
#pragma unroll
          for (int q = 0; q < 8; ++q) v[q] = (__bfloat162float(mb[q]) > 0.f) ? v[q] : 0.f;
        }
      }
      uint4 o;
      o.x = pack_bf16x2(v[0], v[1]); o.y = pack_bf16x2(v[2], v[3]); o.z = pack_bf16x2(v[4], v[5]); o.w = pack_bf16x2(v[6], v[7]);
      *reinterpret_cast<uint4*>(slab + lane * 128 + (((uint32_t) j ^ sw) << 4)) = o;
    }
    if constexpr (ACT == ACT_RELU) {
      if (p.relu_bits_out != nullptr && my_row < p.M) {
        uint32_t* bw = p.relu_bits_out + (int64_t) my_row * p.ld_bits + (n0 >> 5);
        bw[0] = obits0;
        if (n0 + 32 < p.N) bw[1] = obits1;
      }
    }
    fence_proxy_async();
    __syncwarp();
    if (lane == 0 && row0 < p.M) {
      tma_store_2d(tmap_o, slab, n0, row0);
      bulk_commit_group();
    }
    if (p.colsum_ws != nullptr && row0 < p.M) {
      // column sums of the slab (the bf16 values that were just stored, so the result equals a column sum of the output tensor):
      // lane l owns columns 2l, 2l+1 = one 4-byte word per row; a row's 32 words are its 128 swizzled bytes -> conflict free
      float s0 = 0.f, s1 = 0.f;
      const uint32_t piece = (uint32_t) lane >> 2, sub = ((uint32_t) lane & 3u) * 4u;
      const int rmax = min(32, p.M - row0);
#pragma unroll 8
      for (int r = 0; r < rmax; ++r) {
        const uint32_t u = *reinterpret_cast<const uint32_t*>(slab + r * 128 + ((piece ^ ((uint32_t) r & 7u)) << 4) + sub);
        s0 += __uint_as_float(u << 16);
        s1 += __uint_as_float(u & 0xffff0000u);
      }
      const int col = n0 + 2 * lane;
      if (col < p.N) *reinterpret_cast<float2*>(p.colsum_ws + (int64_t) (row0 >> 5) * p.N + col) = make_float2(s0, s1);
    }
    ++store_count;
  }
}

// L2 prefetch of this warp's [32 rows x BN cols] slice of the ReLU-gradient mask, issued while the warp would otherwise idle on
// the accumulator barrier: with 16 KB of mask loads in flight per SM the masked dgrad was latency bound at ~1.6 TB/s.
template <int BN>
__device__ __forceinline__ void prefetch_mask_tile(const GemmParams& p, int row0, int n_base, int lane) {
  const int row = row0 + lane;
  if (row >= p.M) return;
  const __nv_bfloat16* mrow = p.mask + (int64_t) row * p.ld_mask;
#pragma unroll
  for (int c = 0; c < BN; c += 64)
    if (n_base + c < p.N) asm volatile("prefetch.global.L2 [%0];" ::"l"(mrow + n_base + c));
}

template <int BN>
__device__ __noinline__ void epilogue_bf16_tma(const CUtensorMap* tmap_o, const GemmParams& p, uint32_t tmem_acc, int row0, int n_base, int lane,
                                               uint8_t* epi_warp, uint32_t& store_count) {
#define TRB_EPI(A, B) epilogue_bf16_tma_impl<BN, A, B>(tmap_o, p, tmem_acc, row0, n_base, lane, epi_warp, store_count)
  const bool bias = p.bias != nullptr;
  switch (p.act) {
    case ACT_RELU: if (bias) TRB_EPI(ACT_RELU, true); else TRB_EPI(ACT_RELU, false); break;
    case ACT_SIGMOID: if (bias) TRB_EPI(ACT_SIGMOID, true); else TRB_EPI(ACT_SIGMOID, false); break;
    case ACT_RELU_GRAD: if (bias) TRB_EPI(ACT_RELU_GRAD, true); else TRB_EPI(ACT_RELU_GRAD, false); break;
    default: if (bias) TRB_EPI(ACT_NONE, true); else TRB_EPI(ACT_NONE, false); break;
  }
#undef TRB_EPI
}

template <int BLOCK_N, int kStages>
struct SmemLayout {
  static constexpr bool kEpiTma = BLOCK_N <= 128;  // the 128 x 256 variant has no smem left for the epilogue slabs
  static constexpr int kABytes = BLOCK_M * BLOCK_K * 2;
  static constexpr int kBBytes = BLOCK_N * BLOCK_K * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kEpiOffset = kStages * kStageBytes;
  static constexpr int kBarOffset = kEpiOffset + (kEpiTma ? kEpiBytes : 0);
  static constexpr int kTotal = kBarOffset + (2 * kStages + 4) * 8 + 16;
};

template <int BLOCK_N, int kStages, bool A_MN, bool B_MN>
__global__ void __launch_bounds__(kNumThreads, 1)
gemm_bf16_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                         const __grid_constant__ CUtensorMap tmap_o, const GemmParams p) {
  using L = SmemLayout<BLOCK_N, kStages>;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // dynamic smem is only guaranteed 16 B aligned: round up to 1024 B for SWIZZLE_128B
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t) 1023);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::kBarOffset);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* tmem_full_bar = empty_bar + kStages;
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);

  const int warp_idx = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  constexpr uint32_t kTmemCols = 2 * BLOCK_N;  // power of two >= 32 for BLOCK_N in {64,128,256}

  if (warp_idx == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmap_a)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmap_b)) : "memory");
  }
  if (warp_idx == 1 && lane == 0) {
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full_bar[i], 1);
      mbar_init(&tmem_empty_bar[i], 128);
    }
    fence_barrier_init();
  }
  if (warp_idx == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_smem)), "r"(kTmemCols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  const int m_tiles = (p.M + BLOCK_M - 1) / BLOCK_M;
  const int n_tiles = (p.N + BLOCK_N - 1) / BLOCK_N;
  const int k_blocks_total = (p.K + BLOCK_K - 1) / BLOCK_K;
  const int splits = p.split_k > 1 ? p.split_k : 1;
  const int kb_per_split = (k_blocks_total + splits - 1) / splits;
  const int mn_tiles = m_tiles * n_tiles;
  const int num_tiles = mn_tiles * splits;

  if (warp_idx == 0) {
    // ================= TMA producer =================
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int mn = tile % mn_tiles, ks = tile / mn_tiles;
        const int m_blk = mn / n_tiles, n_blk = mn % n_tiles;
        const int kb0 = ks * kb_per_split, kb1 = min(k_blocks_total, kb0 + kb_per_split);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * L::kStageBytes;
          uint8_t* sb = sa + L::kABytes;
          mbar_arrive_expect_tx(&full_bar[stage], L::kStageBytes);
          if constexpr (A_MN) {
            // MN-major operand: the matrix is [reduction rows, MN cols]; one 64-column atom ([64 rows x 128 B]) per load
#pragma unroll
            for (int h = 0; h < BLOCK_M / 64; ++h) tma_load_2d(&tmap_a, &full_bar[stage], sa + h * 8192, m_blk * BLOCK_M + 64 * h, kb * BLOCK_K);
          } else {
            tma_load_2d(&tmap_a, &full_bar[stage], sa, kb * BLOCK_K, m_blk * BLOCK_M);
          }
          if constexpr (B_MN) {
#pragma unroll
            for (int h = 0; h < BLOCK_N / 64; ++h) tma_load_2d(&tmap_b, &full_bar[stage], sb + h * 8192, n_blk * BLOCK_N + 64 * h, kb * BLOCK_K);
          } else {
            tma_load_2d(&tmap_b, &full_bar[stage], sb, kb * BLOCK_K, n_blk * BLOCK_N);
          }
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp_idx == 1) {
    // ================= MMA issuer =================
    constexpr uint32_t idesc = make_idesc_major(BLOCK_M, BLOCK_N, A_MN ? 1 : 0, B_MN ? 1 : 0);
    int stage = 0;
    uint32_t phase = 0;
    int accum_stage = 0;
    uint32_t accum_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      mbar_wait(&tmem_empty_bar[accum_stage], accum_phase ^ 1);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + accum_stage * BLOCK_N;
      const int ks = tile / mn_tiles;
      const int kb0 = ks * kb_per_split, kb1 = min(k_blocks_total, kb0 + kb_per_split);
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t a_addr = smem_u32(smem + stage * L::kStageBytes);
          const uint32_t b_addr = a_addr + L::kABytes;
          const uint64_t adesc = make_kmajor_desc(a_addr);
          const uint64_t bdesc = make_kmajor_desc(b_addr);
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
            // advance 16 bf16 = 32 B inside the 128 B swizzle row: +2 in the (addr >> 4) field
            // K-major: +32 B inside the 128 B swizzle row; MN-major: 16 reduction rows x 128 B = 2048 B
            const uint64_t ad = A_MN ? make_mnmajor_desc(a_addr + k * 2048, 8192, 1024) : adesc + (uint64_t) (2 * k);
            const uint64_t bd = B_MN ? make_mnmajor_desc(b_addr + k * 2048, 8192, 1024) : bdesc + (uint64_t) (2 * k);
            umma_bf16(tmem_d, ad, bd, idesc, ((kb - kb0) | k) != 0);
          }
        }
        __syncwarp();
        if (elect_one()) {
          umma_commit(&empty_bar[stage]);                       // smem slot reusable once these MMAs retire
          if (kb == kb1 - 1) umma_commit(&tmem_full_bar[accum_stage]);  // accumulator complete
        }
        __syncwarp();
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
      if (++accum_stage == 2) { accum_stage = 0; accum_phase ^= 1; }
    }
  } else if (warp_idx >= kEpiWarp0) {
    // ================= epilogue =================
    const int ew = warp_idx - kEpiWarp0;  // == warp_idx % 4 -> TMEM lanes [32*ew, 32*ew+32)
    int accum_stage = 0;
    uint32_t accum_phase = 0;
    uint32_t store_count = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int mn = tile % mn_tiles;
      const int m_blk = mn / n_tiles, n_blk = mn % n_tiles;
      if constexpr (L::kEpiTma) {
        if (p.tma_store && p.act == ACT_RELU_GRAD && p.mask_bits == nullptr) {
          // this tile's mask on the first iteration, then always the NEXT tile's: a whole epilogue ahead of its use
          if (tile == (int) blockIdx.x) prefetch_mask_tile<BLOCK_N>(p, m_blk * BLOCK_M + ew * 32, n_blk * BLOCK_N, lane);
          const int nt = tile + (int) gridDim.x;
          if (nt < num_tiles) {
            const int mn2 = nt % mn_tiles;
            prefetch_mask_tile<BLOCK_N>(p, (mn2 / n_tiles) * BLOCK_M + ew * 32, (mn2 % n_tiles) * BLOCK_N, lane);
          }
        }
      }
      mbar_wait(&tmem_full_bar[accum_stage], accum_phase);
      tc_fence_after();
      const int row = m_blk * BLOCK_M + ew * 32 + lane;
      const bool row_ok = row < p.M;
      if constexpr (L::kEpiTma) {
        if (p.tma_store) {
          epilogue_bf16_tma<BLOCK_N>(&tmap_o, p, tmem_base + ((uint32_t) (ew * 32) << 16) + (uint32_t) (accum_stage * BLOCK_N),
                                     m_blk * BLOCK_M + ew * 32, n_blk * BLOCK_N, lane, smem + L::kEpiOffset + ew * kEpiBytesPerWarp, store_count);
          tc_fence_before();
          mbar_arrive(&tmem_empty_bar[accum_stage]);
          if (++accum_stage == 2) { accum_stage = 0; accum_phase ^= 1; }
          continue;
        }
      }
#pragma unroll 1
      for (int c = 0; c < BLOCK_N / 32; ++c) {
        uint32_t r[32];
        const uint32_t taddr = tmem_base + ((uint32_t) (ew * 32) << 16) + (uint32_t) (accum_stage * BLOCK_N + c * 32);
        tmem_ld_32x32(taddr, r);
        const int n0 = n_blk * BLOCK_N + c * 32;
        if (n0 >= p.N) continue;
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]) * p.alpha;
        if (p.bias != nullptr) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            if (n0 + j < p.N) {
              const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + n0 + j));
              v[j] += b4.x; v[j + 1] += b4.y; v[j + 2] += b4.z; v[j + 3] += b4.w;
            }
          }
        }
        if (p.act == ACT_RELU) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
          if (p.relu_bits_out != nullptr && row_ok) {
            uint32_t w = 0u;
#pragma unroll
            for (int j = 0; j < 32; ++j) w |= (v[j] > 0.f ? 1u : 0u) << j;
            p.relu_bits_out[(int64_t) row * p.ld_bits + (n0 >> 5)] = w;
          }
        } else if (p.act == ACT_SIGMOID) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = 1.f / (1.f + __expf(-v[j]));
        } else if (p.act == ACT_RELU_GRAD && p.mask_bits != nullptr) {
          if (row_ok) {
            const uint32_t w = __ldg(p.mask_bits + (int64_t) row * p.ld_bits + (n0 >> 5));
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = ((w >> j) & 1u) ? v[j] : 0.f;
          }
        } else if (p.act == ACT_RELU_GRAD) {
          if (row_ok) {
            const __nv_bfloat16* mrow = p.mask + (int64_t) row * p.ld_mask + n0;
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
              if (n0 + j < p.N) {
                const uint4 m8 = *reinterpret_cast<const uint4*>(mrow + j);
                const __nv_bfloat16* mb = reinterpret_cast<const __nv_bfloat16*>(&m8);
#pragma unroll
                for (int q = 0; q < 8; ++q) v[j + q] = (__bfloat162float(mb[q]) > 0.f) ? v[j + q] : 0.f;
              }
            }
          }
        }
        if (row_ok) {
          if (splits > 1) {
            float* orow = reinterpret_cast<float*>(p.out) + (int64_t) row * p.ldo + n0;
#pragma unroll
            for (int j = 0; j < 32; j += 4)  // N % 8 == 0 and 16 B aligned rows: one vector reduction per 4 columns
              if (n0 + j < p.N)
                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(orow + j), "f"(v[j]), "f"(v[j + 1]), "f"(v[j + 2]), "f"(v[j + 3])
                             : "memory");
          } else if (p.out_f32) {
            float* orow = reinterpret_cast<float*>(p.out) + (int64_t) row * p.ldo + n0;
#pragma unroll
            for (int j = 0; j < 32; j += 4)
              if (n0 + j < p.N) *reinterpret_cast<float4*>(orow + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
          } else {
            __nv_bfloat16* orow = reinterpret_cast<__nv_bfloat16*>(p.out) + (int64_t) row * p.ldo + n0;
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
              if (n0 + j < p.N) {
                uint4 o;
                __nv_bfloat162 h0 = __floats2bfloat162_rn(v[j], v[j + 1]), h1 = __floats2bfloat162_rn(v[j + 2], v[j + 3]);
                __nv_bfloat162 h2 = __floats2bfloat162_rn(v[j + 4], v[j + 5]), h3 = __floats2bfloat162_rn(v[j + 6], v[j + 7]);
                o.x = *reinterpret_cast<uint32_t*>(&h0); o.y = *reinterpret_cast<uint32_t*>(&h1);
                o.z = *reinterpret_cast<uint32_t*>(&h2); o.w = *reinterpret_cast<uint32_t*>(&h3);
                *reinterpret_cast<uint4*>(orow + j) = o;
              }
            }
          }
        }
      }
      tc_fence_before();
      mbar_arrive(&tmem_empty_bar[accum_stage]);
      if (++accum_stage == 2) { accum_stage = 0; accum_phase ^= 1; }
    }
    if (lane == 0) bulk_wait_group<0>();  // smem slabs must outlive the last TMA stores
  }

  tc_fence_before();
  __syncthreads();
  if (warp_idx == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols));
  }
}

// ---------------------------------------------------------------------------------------------------------
// CTA-pair kernel: one (2,1,1) cluster = the two SMs of a TPC computes a 256 x 256 output tile with
// tcgen05.mma.cta_group::2 (UMMA 256 x 256 x 16). Per k-block every CTA stages only its own 128 rows of A and its
// own 128 rows of B (32 KB for 4.2 MFLOP of its half of the tile; the 1-CTA 128 x 256 tile needs 48 KB), which is
// what lifts the L2 -> smem operand-traffic bound of the kernels above. Roles per CTA are the same (warp 0 TMA,
// warp 1 MMA (leader CTA only), warp 2 TMEM alloc, warps 4..7 epilogue over the CTA's own 128 TMEM lanes).
//   full[s]      (leader)  1 arrival (leader producer, expect_tx = both CTAs' bytes); both producers' TMA complete_tx on it
//   empty[s]     (each)    1 arrival: the leader's commit, multicast to both CTAs
//   tmem_full[a] (each)    1 arrival: the leader's commit after the last k-block, multicast
//   tmem_empty[a](leader)  256 arrivals: the epilogue threads of BOTH CTAs (remote arrive from the peer)
constexpr int kPairN = 256;       // UMMA N of the pair
constexpr int kPairStages = 5;    // 5 x 32 KB (+ 52 KB of epilogue slabs)

struct PairSmem {
  static constexpr int kABytes = BLOCK_M * BLOCK_K * 2;       // this CTA's 128 rows of A
  static constexpr int kBBytes = (kPairN / 2) * BLOCK_K * 2;  // this CTA's half of B
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kEpiOffset = kPairStages * kStageBytes;
  static constexpr int kBarOffset = kEpiOffset + kEpiBytes;
  static constexpr int kTotal = kBarOffset + (2 * kPairStages + 4) * 8 + 16;
};

template <bool A_MN, bool B_MN>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kNumThreads, 1)
gemm_bf16_tcgen05_pair_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                              const __grid_constant__ CUtensorMap tmap_o, const GemmParams p) {
  using L = PairSmem;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t) 1023);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::kBarOffset);
  uint64_t* empty_bar = full_bar + kPairStages;
  uint64_t* tmem_full_bar = empty_bar + kPairStages;
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);

  const int warp_idx = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t cta_rank = cluster_ctarank();
  const bool leader = cta_rank == 0;
  constexpr uint32_t kTmemCols = 2 * kPairN;  // 2 accumulator stages x 256 fp32 columns = all 512 columns

  if (warp_idx == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmap_a)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmap_b)) : "memory");
  }
  if (warp_idx == 1 && lane == 0) {
    for (int i = 0; i < kPairStages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full_bar[i], 1);
      mbar_init(&tmem_empty_bar[i], 256);
    }
    fence_barrier_init();
  }
  if (warp_idx == 2) tmem_alloc_2sm(tmem_ptr_smem, kTmemCols);
  tc_fence_before();
  cluster_sync_all();  // barrier inits + TMEM allocation of both CTAs visible before any remote arrive / multicast commit
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  const int m_tiles = (p.M + 2 * BLOCK_M - 1) / (2 * BLOCK_M);
  const int n_tiles = (p.N + kPairN - 1) / kPairN;
  const int k_blocks_total = (p.K + BLOCK_K - 1) / BLOCK_K;
  const int splits = p.split_k > 1 ? p.split_k : 1;
  const int kb_per_split = (k_blocks_total + splits - 1) / splits;
  const int mn_tiles = m_tiles * n_tiles;
  const int num_tiles = mn_tiles * splits;
  const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;

  if (warp_idx == 0) {
    // ================= TMA producer (both CTAs; bytes are accounted on the LEADER's full barrier) =================
    if (elect_one()) {
      const uint32_t leader_full0 = mapa_shared(smem_u32(&full_bar[0]), 0);
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
        const int mn = tile % mn_tiles, ks = tile / mn_tiles;
        const int m_blk = mn / n_tiles, n_blk = mn % n_tiles;
        const int m0 = m_blk * 2 * BLOCK_M + (int) cta_rank * BLOCK_M;       // this CTA's A rows
        const int n0 = n_blk * kPairN + (int) cta_rank * (kPairN / 2);       // this CTA's B rows
        const int kb0 = ks * kb_per_split, kb1 = min(k_blocks_total, kb0 + kb_per_split);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * L::kStageBytes;
          uint8_t* sb = sa + L::kABytes;
          const uint32_t bar = leader_full0 + (uint32_t) stage * 8;
          if (leader) mbar_arrive_expect_tx(&full_bar[stage], 2 * L::kStageBytes);
          if constexpr (A_MN) {
#pragma unroll
            for (int h = 0; h < BLOCK_M / 64; ++h) tma_load_2d_2sm(&tmap_a, bar, sa + h * 8192, m0 + 64 * h, kb * BLOCK_K);
          } else {
            tma_load_2d_2sm(&tmap_a, bar, sa, kb * BLOCK_K, m0);
          }
          if constexpr (B_MN) {
#pragma unroll
            for (int h = 0; h < (kPairN / 2) / 64; ++h) tma_load_2d_2sm(&tmap_b, bar, sb + h * 8192, n0 + 64 * h, kb * BLOCK_K);
          } else {
            tma_load_2d_2sm(&tmap_b, bar, sb, kb * BLOCK_K, n0);
          }
          if (++stage == kPairStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp_idx == 1) {
    // ================= MMA issuer (leader CTA only) =================
    if (leader) {
      constexpr uint32_t idesc = make_idesc_major(2 * BLOCK_M, kPairN, A_MN ? 1 : 0, B_MN ? 1 : 0);
      int stage = 0;
      uint32_t phase = 0;
      int accum_stage = 0;
      uint32_t accum_phase = 0;
      for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
        mbar_wait(&tmem_empty_bar[accum_stage], accum_phase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + accum_stage * kPairN;
        const int ks = tile / mn_tiles;
        const int kb0 = ks * kb_per_split, kb1 = min(k_blocks_total, kb0 + kb_per_split);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          if (elect_one()) {
            const uint32_t a_addr = smem_u32(smem + stage * L::kStageBytes);
            const uint32_t b_addr = a_addr + L::kABytes;
            const uint64_t adesc = make_kmajor_desc(a_addr);
            const uint64_t bdesc = make_kmajor_desc(b_addr);
#pragma unroll
            for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
              const uint64_t ad = A_MN ? make_mnmajor_desc(a_addr + k * 2048, 8192, 1024) : adesc + (uint64_t) (2 * k);
              const uint64_t bd = B_MN ? make_mnmajor_desc(b_addr + k * 2048, 8192, 1024) : bdesc + (uint64_t) (2 * k);
              umma_bf16_2sm(tmem_d, ad, bd, idesc, ((kb - kb0) | k) != 0);
            }
          }
          __syncwarp();
          if (elect_one()) {
            umma_commit_2sm(&empty_bar[stage]);                              // both CTAs' smem slots reusable
            if (kb == kb1 - 1) umma_commit_2sm(&tmem_full_bar[accum_stage]);  // both CTAs' halves of the accumulator complete
          }
          __syncwarp();
          if (++stage == kPairStages) { stage = 0; phase ^= 1; }
        }
        if (++accum_stage == 2) { accum_stage = 0; accum_phase ^= 1; }
      }
    }
  } else if (warp_idx >= kEpiWarp0) {
    // ================= epilogue (each CTA drains its own 128 accumulator rows) =================
    const int ew = warp_idx - kEpiWarp0;
    const uint32_t leader_tmem_empty0 = mapa_shared(smem_u32(&tmem_empty_bar[0]), 0);
    int accum_stage = 0;
    uint32_t accum_phase = 0;
    uint32_t store_count = 0;
    for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
      const int mn = tile % mn_tiles;
      const int m_blk = mn / n_tiles, n_blk = mn % n_tiles;
      if (p.tma_store && p.act == ACT_RELU_GRAD && p.mask_bits == nullptr) {
        // this tile's mask on the first iteration, then always the NEXT tile's: a whole epilogue ahead of its use
        if (tile == cluster_id) prefetch_mask_tile<kPairN>(p, m_blk * 2 * BLOCK_M + (int) cta_rank * BLOCK_M + ew * 32, n_blk * kPairN, lane);
        const int nt = tile + num_clusters;
        if (nt < num_tiles) {
          const int mn2 = nt % mn_tiles;
          prefetch_mask_tile<kPairN>(p, (mn2 / n_tiles) * 2 * BLOCK_M + (int) cta_rank * BLOCK_M + ew * 32, (mn2 % n_tiles) * kPairN, lane);
        }
      }
      mbar_wait(&tmem_full_bar[accum_stage], accum_phase);
      tc_fence_after();
      const int row = m_blk * 2 * BLOCK_M + (int) cta_rank * BLOCK_M + ew * 32 + lane;
      const bool row_ok = row < p.M;
      if (p.tma_store) {
        epilogue_bf16_tma<kPairN>(&tmap_o, p, tmem_base + ((uint32_t) (ew * 32) << 16) + (uint32_t) (accum_stage * kPairN),
                                  m_blk * 2 * BLOCK_M + (int) cta_rank * BLOCK_M + ew * 32, n_blk * kPairN, lane,
                                  smem + L::kEpiOffset + ew * kEpiBytesPerWarp, store_count);
        tc_fence_before();
        mbar_arrive_cluster(leader_tmem_empty0 + (uint32_t) accum_stage * 8);
        if (++accum_stage == 2) { accum_stage = 0; accum_phase ^= 1; }
        continue;
      }
#pragma unroll 1
      for (int c = 0; c < kPairN / 32; ++c) {
        uint32_t r[32];
        const uint32_t taddr = tmem_base + ((uint32_t) (ew * 32) << 16) + (uint32_t) (accum_stage * kPairN + c * 32);
        tmem_ld_32x32(taddr, r);
        const int n0 = n_blk * kPairN + c * 32;
        if (n0 >= p.N) continue;
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]) * p.alpha;
        if (p.bias != nullptr) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            if (n0 + j < p.N) {
              const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + n0 + j));
              v[j] += b4.x; v[j + 1] += b4.y; v[j + 2] += b4.z; v[j + 3] += b4.w;
            }
          }
        }
        if (p.act == ACT_RELU) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
          if (p.relu_bits_out != nullptr && row_ok) {
            uint32_t w = 0u;
#pragma unroll
            for (int j = 0; j < 32; ++j) w |= (v[j] > 0.f ? 1u : 0u) << j;
            p.relu_bits_out[(int64_t) row * p.ld_bits + (n0 >> 5)] = w;
          }
        } else if (p.act == ACT_SIGMOID) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = 1.f / (1.f + __expf(-v[j]));
        } else if (p.act == ACT_RELU_GRAD && p.mask_bits != nullptr) {
          if (row_ok) {
            const uint32_t w = __ldg(p.mask_bits + (int64_t) row * p.ld_bits + (n0 >> 5));
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = ((w >> j) & 1u) ? v[j] : 0.f;
          }
        } else if (p.act == ACT_RELU_GRAD) {
          if (row_ok) {
            const __nv_bfloat16* mrow = p.mask + (int64_t) row * p.ld_mask + n0;
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
              if (n0 + j < p.N) {
                const uint4 m8 = *reinterpret_cast<const uint4*>(mrow + j);
                const __nv_bfloat16* mb = reinterpret_cast<const __nv_bfloat16*>(&m8);
#pragma unroll
                for (int q = 0; q < 8; ++q) v[j + q] = (__bfloat162float(mb[q]) > 0.f) ? v[j + q] : 0.f;
              }
            }
          }
        }
        if (row_ok) {
          if (splits > 1) {
            float* orow = reinterpret_cast<float*>(p.out) + (int64_t) row * p.ldo + n0;
#pragma unroll
            for (int j = 0; j < 32; j += 4)
              if (n0 + j < p.N)
                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(orow + j), "f"(v[j]), "f"(v[j + 1]), "f"(v[j + 2]), "f"(v[j + 3])
                             : "memory");
          } else if (p.out_f32) {
            float* orow = reinterpret_cast<float*>(p.out) + (int64_t) row * p.ldo + n0;
#pragma unroll
            for (int j = 0; j < 32; j += 4)
              if (n0 + j < p.N) *reinterpret_cast<float4*>(orow + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
          } else {
            __nv_bfloat16* orow = reinterpret_cast<__nv_bfloat16*>(p.out) + (int64_t) row * p.ldo + n0;
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
              if (n0 + j < p.N) {
                uint4 o;
                __nv_bfloat162 h0 = __floats2bfloat162_rn(v[j], v[j + 1]), h1 = __floats2bfloat162_rn(v[j + 2], v[j + 3]);
                __nv_bfloat162 h2 = __floats2bfloat162_rn(v[j + 4], v[j + 5]), h3 = __floats2bfloat162_rn(v[j + 6], v[j + 7]);
                o.x = *reinterpret_cast<uint32_t*>(&h0); o.y = *reinterpret_cast<uint32_t*>(&h1);
                o.z = *reinterpret_cast<uint32_t*>(&h2); o.w = *reinterpret_cast<uint32_t*>(&h3);
                *reinterpret_cast<uint4*>(orow + j) = o;
              }
            }
          }
        }
      }
      tc_fence_before();
      mbar_arrive_cluster(leader_tmem_empty0 + (uint32_t) accum_stage * 8);  // leader's MMA warp waits for both CTAs' 128 threads
      if (++accum_stage == 2) { accum_stage = 0; accum_phase ^= 1; }
    }
    if (lane == 0) bulk_wait_group<0>();
  }

  tc_fence_before();
  cluster_sync_all();  // neither CTA may free TMEM / exit while the peer can still multicast into its barriers
  if (warp_idx == 2) {
    tc_fence_after();
    tmem_dealloc_2sm(tmem_base, kTmemCols);
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}

// rows x K bf16 matrix, row pitch ld elements; box = [BLOCK_K, box_rows], SWIZZLE_128B
int make_tmap(CUtensorMap* map, const void* ptr, int64_t rows, int64_t K, int64_t ld, int box_rows, int box_cols = BLOCK_K) {
  EncodeTiledFn fn = get_encode_fn();
  if (fn == nullptr) return -10;
  cuuint64_t dims[2] = {(cuuint64_t) K, (cuuint64_t) rows};
  cuuint64_t strides[1] = {(cuuint64_t) ld * 2};
  cuuint32_t box[2] = {(cuuint32_t) box_cols, (cuuint32_t) box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : -11;
}

// bf16 output [M, N] with row pitch ldo: box = one epilogue slab [64 cols x 32 rows], SWIZZLE_128B
int make_tmap_out(CUtensorMap* map, const void* ptr, int64_t M, int64_t N, int64_t ldo) { return make_tmap(map, ptr, M, N, ldo, 32, 64); }

int g_num_sms = 0;

template <int BLOCK_N, int kStages, bool A_MN, bool B_MN>
int launch_gemm(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& to, GemmParams p, cudaStream_t stream) {
  if (!SmemLayout<BLOCK_N, kStages>::kEpiTma) p.tma_store = 0;
  using L = SmemLayout<BLOCK_N, kStages>;
  constexpr int smem_bytes = L::kTotal + 1024;
  static bool configured = false;
  if (!configured) {
    TRB_CUDA(cudaFuncSetAttribute(gemm_bf16_tcgen05_kernel<BLOCK_N, kStages, A_MN, B_MN>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    configured = true;
  }
  if (g_num_sms == 0) {
    int dev = 0;
    TRB_CUDA(cudaGetDevice(&dev));
    TRB_CUDA(cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev));
  }
  const int tiles = ((p.M + BLOCK_M - 1) / BLOCK_M) * ((p.N + BLOCK_N - 1) / BLOCK_N) * (p.split_k > 1 ? p.split_k : 1);
  const int grid = tiles < g_num_sms ? tiles : g_num_sms;
  gemm_bf16_tcgen05_kernel<BLOCK_N, kStages, A_MN, B_MN><<<grid, kNumThreads, smem_bytes, stream>>>(ta, tb, to, p);
  TRB_CHECK_LAUNCH();
  return 0;
}

template <bool A_MN, bool B_MN>
int launch_gemm_pair(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& to, const GemmParams& p, cudaStream_t stream) {
  constexpr int smem_bytes = PairSmem::kTotal + 1024;
  static bool configured = false;
  if (!configured) {
    TRB_CUDA(cudaFuncSetAttribute(gemm_bf16_tcgen05_pair_kernel<A_MN, B_MN>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    configured = true;
  }
  if (g_num_sms == 0) {
    int dev = 0;
    TRB_CUDA(cudaGetDevice(&dev));
    TRB_CUDA(cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev));
  }
  const int tiles = ((p.M + 2 * BLOCK_M - 1) / (2 * BLOCK_M)) * ((p.N + kPairN - 1) / kPairN) * (p.split_k > 1 ? p.split_k : 1);
  const int clusters = tiles < g_num_sms / 2 ? tiles : g_num_sms / 2;
  gemm_bf16_tcgen05_pair_kernel<A_MN, B_MN><<<2 * clusters, kNumThreads, smem_bytes, stream>>>(ta, tb, to, p);
  TRB_CHECK_LAUNCH();
  return 0;
}

}  // namespace

// C[M,N] = act(alpha * op(A) . op(B)^T + bias), bf16 operands, fp32 accumulation.
//   a_mn == 0: A is [M, K] row-major (K-major operand)      a_mn == 1: A is [K, M] row-major (MN-major operand)
//   b_mn == 0: B is [N, K] row-major                        b_mn == 1: B is [K, N] row-major
// Pitches lda/ldb in elements (multiples of 8). out: bf16 or fp32 with pitch ldo.
template <bool A_MN, bool B_MN>
static int gemm_dispatch(const void* A, int64_t lda, const void* B, int64_t ldb, GemmParams& p, int tile_n, cudaStream_t stream) {
  CUtensorMap ta, tb, to;
  int rc = A_MN ? make_tmap(&ta, A, p.K, p.M, lda, 64, 64) : make_tmap(&ta, A, p.M, p.K, lda, BLOCK_M);
  if (rc) return rc;
  static const int epi_tma = getenv("TRB_GEMM_EPI_TMA") ? atoi(getenv("TRB_GEMM_EPI_TMA")) : 1;
  p.tma_store = (epi_tma && !p.out_f32 && p.split_k <= 1 && (p.ldo % 8) == 0 && (reinterpret_cast<uintptr_t>(p.out) % 16) == 0) ? 1 : 0;
  if (p.tma_store && make_tmap_out(&to, p.out, p.M, p.N, p.ldo) != 0) p.tma_store = 0;  // fall back to the direct epilogue
  if (!p.tma_store) to = ta;                                                                // unused
  if (p.colsum_ws != nullptr && !p.tma_store) return -14;  // the column sums live in the TMA epilogue (bf16 output, no split-K)
  if (p.N <= 64) {
    rc = B_MN ? make_tmap(&tb, B, p.K, p.N, ldb, 64, 64) : make_tmap(&tb, B, p.N, p.K, ldb, 64);
    if (rc) return rc;
    return launch_gemm<64, 6, A_MN, B_MN>(ta, tb, to, p, stream);
  }
  // CTA-pair 256 x 256 tiles (cta_group::2): least operand traffic per FLOP. tile_n == 512: caller asked for it (split-K sized
  // for 74 clusters); tile_n == 0: taken when the tile count fills the 74 clusters for >= 2 waves. TRB_GEMM_PAIR=0 disables.
  static const int pair = getenv("TRB_GEMM_PAIR") ? atoi(getenv("TRB_GEMM_PAIR")) : 1;
  {
    const int64_t tiles_pair = (int64_t) ((p.M + 255) / 256) * ((p.N + 255) / 256) * (p.split_k > 1 ? p.split_k : 1);
    const bool use_pair = tile_n == 512 || (pair && tile_n == 0 && p.split_k <= 1 && tiles_pair >= 2 * 74 && p.N >= 256);
    if (use_pair) {
      rc = B_MN ? make_tmap(&tb, B, p.K, p.N, ldb, 64, 64) : make_tmap(&tb, B, p.N, p.K, ldb, 128);
      if (rc) return rc;
      return launch_gemm_pair<A_MN, B_MN>(ta, tb, to, p, stream);
    }
  }
  // 128 x 256 tiles halve the A-operand smem / L2 traffic per FLOP; worth it once there are enough tiles to fill the SMs
  static const int wide = getenv("TRB_GEMM_WIDE") ? atoi(getenv("TRB_GEMM_WIDE")) : 1;
  const int64_t tiles256 = (int64_t) ((p.M + BLOCK_M - 1) / BLOCK_M) * ((p.N + 255) / 256) * (p.split_k > 1 ? p.split_k : 1);
  // tile_n == 256: the caller sized its split-K for 128 x 256 tiles (ops/gemm.py: _wgrad_plan); 0: decide here
  const bool use_wide = tile_n == 256 ? true : (tile_n == 0 && p.split_k <= 1 && tiles256 >= 2 * 148);
  if (wide && use_wide && p.N % 256 == 0 && p.colsum_ws == nullptr) {  // (the 128 x 256 variant has the direct epilogue)
    rc = B_MN ? make_tmap(&tb, B, p.K, p.N, ldb, 64, 64) : make_tmap(&tb, B, p.N, p.K, ldb, 256);
    if (rc) return rc;
    return launch_gemm<256, 4, A_MN, B_MN>(ta, tb, to, p, stream);
  }
  rc = B_MN ? make_tmap(&tb, B, p.K, p.N, ldb, 64, 64) : make_tmap(&tb, B, p.N, p.K, ldb, 128);
  if (rc) return rc;
  return launch_gemm<128, 5, A_MN, B_MN>(ta, tb, to, p, stream);
}

TRB_API int trb_gemm_bf16_ex2(const void* A, int64_t lda, int a_mn, const void* B, int64_t ldb, int b_mn, void* out, int64_t ldo, int out_f32, int M,
                              int N, int K, const float* bias, int act, const void* mask, int64_t ld_mask, float alpha, int split_k, int tile_n,
                              const void* mask_bits, void* relu_bits_out, int64_t ld_bits, cudaStream_t stream);

TRB_API int trb_gemm_bf16_ex(const void* A, int64_t lda, int a_mn, const void* B, int64_t ldb, int b_mn, void* out, int64_t ldo, int out_f32, int M,
                             int N, int K, const float* bias, int act, const void* mask, int64_t ld_mask, float alpha, int split_k, int tile_n,
                             cudaStream_t stream) {
  return trb_gemm_bf16_ex2(A, lda, a_mn, B, ldb, b_mn, out, ldo, out_f32, M, N, K, bias, act, mask, ld_mask, alpha, split_k, tile_n, nullptr, nullptr, 0,
                           stream);
}

// mask_bits / relu_bits_out: int32 [M, ld_bits] bit matrices (bit n % 32 of word n / 32), see GemmParams
TRB_API int trb_gemm_bf16_ex3(const void* A, int64_t lda, int a_mn, const void* B, int64_t ldb, int b_mn, void* out, int64_t ldo, int out_f32, int M,
                              int N, int K, const float* bias, int act, const void* mask, int64_t ld_mask, float alpha, int split_k, int tile_n,
                              const void* mask_bits, void* relu_bits_out, int64_t ld_bits, float* colsum_ws, cudaStream_t stream);

TRB_API int trb_gemm_bf16_ex2(const void* A, int64_t lda, int a_mn, const void* B, int64_t ldb, int b_mn, void* out, int64_t ldo, int out_f32, int M,
                              int N, int K, const float* bias, int act, const void* mask, int64_t ld_mask, float alpha, int split_k, int tile_n,
                              const void* mask_bits, void* relu_bits_out, int64_t ld_bits, cudaStream_t stream) {
  return trb_gemm_bf16_ex3(A, lda, a_mn, B, ldb, b_mn, out, ldo, out_f32, M, N, K, bias, act, mask, ld_mask, alpha, split_k, tile_n, mask_bits, relu_bits_out,
                           ld_bits, nullptr, stream);
}

// colsum_ws: fp32 [ceil(M / 32), N] or nullptr, see GemmParams (bf16 output without split-K only: -14 otherwise)
TRB_API int trb_gemm_bf16_ex3(const void* A, int64_t lda, int a_mn, const void* B, int64_t ldb, int b_mn, void* out, int64_t ldo, int out_f32, int M,
                              int N, int K, const float* bias, int act, const void* mask, int64_t ld_mask, float alpha, int split_k, int tile_n,
                              const void* mask_bits, void* relu_bits_out, int64_t ld_bits, float* colsum_ws, cudaStream_t stream) {
  if (M <= 0 || N <= 0 || K <= 0) return 0;
  if ((lda % 8) || (ldb % 8) || (N % 8)) return -12;
  if ((!a_mn && (K % 8)) || (a_mn && (M % 8)) || (b_mn && (N % 8)) || (!b_mn && (K % 8))) return -12;
  GemmParams p;
  p.M = M; p.N = N; p.K = K;
  p.bias = bias;
  p.mask = reinterpret_cast<const __nv_bfloat16*>(mask);
  p.ld_mask = ld_mask;
  p.out = out;
  p.ldo = ldo;
  p.out_f32 = out_f32;
  p.act = act;
  p.alpha = alpha;
  p.split_k = 1;
  p.mask_bits = reinterpret_cast<const uint32_t*>(mask_bits);
  p.relu_bits_out = reinterpret_cast<uint32_t*>(relu_bits_out);
  p.ld_bits = ld_bits;
  p.colsum_ws = colsum_ws;
  if ((mask_bits != nullptr || relu_bits_out != nullptr) && ld_bits * 32 < N) return -12;
  if (split_k > 1) {
    if (!out_f32 || bias != nullptr || act != ACT_NONE) return -13;  // split-K only for plain fp32 accumulation
    const int kb = (K + BLOCK_K - 1) / BLOCK_K;
    int s = split_k < kb ? split_k : kb;
    const int per = (kb + s - 1) / s;
    s = (kb + per - 1) / per;  // no empty splits
    p.split_k = s;
    if (s > 1) TRB_CUDA(cudaMemsetAsync(out, 0, sizeof(float) * (size_t) M * (size_t) ldo, stream));
  }
  if (a_mn && b_mn) return gemm_dispatch<true, true>(A, lda, B, ldb, p, tile_n, stream);
  if (a_mn) return gemm_dispatch<true, false>(A, lda, B, ldb, p, tile_n, stream);
  if (b_mn) return gemm_dispatch<false, true>(A, lda, B, ldb, p, tile_n, stream);
  return gemm_dispatch<false, false>(A, lda, B, ldb, p, tile_n, stream);
}

TRB_API int trb_gemm_bf16(const void* A, int64_t lda, int a_mn, const void* B, int64_t ldb, int b_mn, void* out, int64_t ldo, int out_f32, int M,
                          int N, int K, const float* bias, int act, const void* mask, int64_t ld_mask, float alpha, int split_k,
                          cudaStream_t stream) {
  return trb_gemm_bf16_ex(A, lda, a_mn, B, ldb, b_mn, out, ldo, out_f32, M, N, K, bias, act, mask, ld_mask, alpha, split_k, 0, stream);
}

TRB_API int trb_gemm_bf16_tn(const void* A, int64_t lda, const void* B, int64_t ldb, void* out, int64_t ldo, int out_f32, int M, int N, int K,
                             const float* bias, int act, const void* mask, int64_t ld_mask, float alpha, int split_k, cudaStream_t stream) {
  return trb_gemm_bf16(A, lda, 0, B, ldb, 0, out, ldo, out_f32, M, N, K, bias, act, mask, ld_mask, alpha, split_k, stream);
}

// ---- helpers used around the GEMM ----------------------------------------------------------------------
// bf16 transpose: out[c, r] = in[r, c]; 32x32 smem tiles, coalesced both ways.
__global__ void __launch_bounds__(256) trb_transpose_bf16_kernel(const __nv_bfloat16* __restrict__ in, __nv_bfloat16* __restrict__ out, int rows,
                                                                   int cols, int64_t ld_in, int64_t ld_out) {
  __shared__ __nv_bfloat16 tile[32][34];
  const int bx = blockIdx.x * 32, by = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  for (int i = ty; i < 32; i += 8) {
    const int r = by + i, c = bx + tx;
    tile[i][tx] = (r < rows && c < cols) ? in[(int64_t) r * ld_in + c] : __float2bfloat16(0.f);
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    const int c = bx + i, r = by + tx;
    if (c < cols && r < rows) out[(int64_t) c * ld_out + r] = tile[tx][i];
  }
}

TRB_API int trb_transpose_bf16(const void* in, void* out, int rows, int cols, int64_t ld_in, int64_t ld_out, cudaStream_t stream) {
  if (rows == 0 || cols == 0) return 0;
  dim3 grid((cols + 31) / 32, (rows + 31) / 32);
  trb_transpose_bf16_kernel<<<grid, 256, 0, stream>>>((const __nv_bfloat16*) in, (__nv_bfloat16*) out, rows, cols, ld_in, ld_out);
  TRB_CHECK_LAUNCH();
  return 0;
}

// column sums of a bf16 [rows, cols] matrix into fp32 (bias gradient). Each thread owns 8 columns
// (one 16 B load per row) and strides over rows; 8 row lanes per block are reduced through smem and
// the block result is added atomically (blocks tile the rows so the whole chip participates).
__global__ void __launch_bounds__(256) trb_colsum_bf16_kernel(const __nv_bfloat16* __restrict__ in, float* __restrict__ out, int rows, int cols,
                                                                int64_t ld, int rows_per_block) {
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int vec = blockIdx.x * 32 + tx;  // 8-column group
  const int nvec = cols >> 3;
  const int r0 = blockIdx.y * rows_per_block;
  const int r1 = min(rows, r0 + rows_per_block);
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
  if (vec < nvec) {
    for (int r = r0 + ty; r < r1; r += 8) {
      const uint4 v = *reinterpret_cast<const uint4*>(in + (int64_t) r * ld + vec * 8);
      const __nv_bfloat16* b = reinterpret_cast<const __nv_bfloat16*>(&v);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += __bfloat162float(b[j]);
    }
  }
  __shared__ float red[8][32][9];
#pragma unroll
  for (int j = 0; j < 8; ++j) red[ty][tx][j] = acc[j];
  __syncthreads();
  if (ty == 0 && vec < nvec) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float s = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) s += red[i][tx][j];
      atomicAdd(out + vec * 8 + j, s);
    }
  }
}

// v3: contended fp32 atomics were the bottleneck of v1 (every 128 B line of `out` took rows/256 x 32 serialized L2 atomics:
// 33 us for 67 MB). Now: (a) ~4 blocks per SM, each walking a tall row slab with 8 independent 16 B loads in flight per
// thread, (b) narrow matrices fold several rows into one warp (vpr = lanes per row) so every lane loads, (c) block partials
// go to a workspace and the LAST block of a column group (ticket counter) sums them in a fixed order: no float atomics,
// deterministic result, one launch.
template <bool CLUSTER>
__device__ __forceinline__ void colsum_v3_body(const __nv_bfloat16* __restrict__ in, float* __restrict__ out, int rows, int cols, int64_t ld,
                                               int rows_per_block, int vpr_log2, float* __restrict__ partial, unsigned int* __restrict__ tickets) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int vpr = 1 << vpr_log2, rpw = 32 >> vpr_log2;  // lanes per row, rows per warp
  const int sub = lane >> vpr_log2;
  const int vec = blockIdx.x * 32 + (lane & (vpr - 1));
  const int nvec = cols >> 3;
  const int r0 = blockIdx.y * rows_per_block;
  const int r1 = min(rows, r0 + rows_per_block);
  const int step = 8 * rpw;
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
  if (vec < nvec) {
    const __nv_bfloat16* base = in + vec * 8;
    int r = r0 + warp * rpw + sub;
    constexpr int U = 8;  // independent 16 B loads in flight per thread
    for (; r + (U - 1) * step < r1; r += U * step) {
      uint4 v[U];
#pragma unroll
      for (int u = 0; u < U; ++u) v[u] = __ldg(reinterpret_cast<const uint4*>(base + (int64_t) (r + u * step) * ld));
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const __nv_bfloat162* b = reinterpret_cast<const __nv_bfloat162*>(&v[u]);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 f = __bfloat1622float2(b[j]);
          acc[2 * j] += f.x;
          acc[2 * j + 1] += f.y;
        }
      }
    }
    for (; r < r1; r += step) {
      const uint4 v = __ldg(reinterpret_cast<const uint4*>(base + (int64_t) r * ld));
      const __nv_bfloat162* b = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = __bfloat1622float2(b[j]);
        acc[2 * j] += f.x;
        acc[2 * j + 1] += f.y;
      }
    }
  }
  // fold the row lanes of a warp, then the 8 warps through smem
  for (int o = vpr; o < 32; o <<= 1) {
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], o);
  }
  __shared__ float red[8][32][9];
  __shared__ unsigned int s_ticket;
  if (sub == 0) {
#pragma unroll
    for (int j = 0; j < 8; ++j) red[warp][lane][j] = acc[j];
  }
  __syncthreads();
  // thread t owns column (blockIdx.x * 256 + t) of this block's 256-column group: vec lane t >> 3, element t & 7
  const int l = threadIdx.x >> 3, j = threadIdx.x & 7;
  const int col = blockIdx.x * 256 + threadIdx.x;
  const bool col_ok = l < vpr && col < cols;
  float s = 0.f;
  if (col_ok) {
#pragma unroll
    for (int i = 0; i < 8; ++i) s += red[i][l][j];
  }
  // CLUSTER: the 8 CTAs of a (1,8,1) cluster first fold their column sums through the leader's shared memory (DSMEM), so a
  // launch produces gridDim.y / 8 partials: short row slabs per thread (4 batches of 8 loads) AND a short last-block tail.
  int n_part = (int) gridDim.y, my_part = (int) blockIdx.y;
  if constexpr (CLUSTER) {
    __shared__ float cl[8][256];
    const uint32_t rank = cluster_ctarank();
    const uint32_t dst = mapa_shared(smem_u32(&cl[rank][threadIdx.x]), 0);
    asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(dst), "f"(s) : "memory");
    cluster_sync_all();
    if (rank != 0) return;
    s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += cl[i][threadIdx.x];
    n_part = (int) gridDim.y >> 3;
    my_part = (int) blockIdx.y >> 3;
  }
  if (n_part == 1) {
    if (col_ok) out[col] = s;
    return;
  }
  if (col_ok) partial[(int64_t) my_part * cols + col] = s;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_ticket = atomicAdd(&tickets[blockIdx.x], 1u);
  __syncthreads();
  if (s_ticket != (unsigned) n_part - 1) return;
  __threadfence();
  // last block of this column group: fixed-order sum of the partials. C = vpr * 8 columns; 256 / C thread groups split the
  // partial rows, 8 loads in flight each, then one smem pass over the groups
  {
    const int C = vpr * 8, G = 256 / C;
    const int c = threadIdx.x & (C - 1), g = threadIdx.x / C;
    const int gcol = blockIdx.x * 256 + c;
    float s = 0.f;
    if (gcol < cols) {
      const float* pc = partial + gcol;
      int y = g;
      for (; y + 7 * G < n_part; y += 8 * G) {
        float t[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) t[u] = __ldcg(pc + (int64_t) (y + u * G) * cols);
#pragma unroll
        for (int u = 0; u < 8; ++u) s += t[u];
      }
      for (; y < n_part; y += G) s += __ldcg(pc + (int64_t) y * cols);
    }
    float* fin = &red[0][0][0];  // 256 floats needed, 2304 available; all earlier reads of red are done (barriers above)
    fin[threadIdx.x] = s;
    __syncthreads();
    if (g == 0 && gcol < cols) {
      float tot = 0.f;
      for (int i = 0; i < G; ++i) tot += fin[i * C + c];
      out[gcol] = tot;
    }
  }
  if (threadIdx.x == 0) tickets[blockIdx.x] = 0;  // ready for the next launch on this stream
}

__global__ void __launch_bounds__(256) trb_colsum_bf16_v3_kernel(const __nv_bfloat16* __restrict__ in, float* __restrict__ out, int rows, int cols,
                                                                   int64_t ld, int rows_per_block, int vpr_log2, float* __restrict__ partial,
                                                                   unsigned int* __restrict__ tickets) {
  colsum_v3_body<false>(in, out, rows, cols, ld, rows_per_block, vpr_log2, partial, tickets);
}

__global__ void __cluster_dims__(1, 8, 1) __launch_bounds__(256)
trb_colsum_bf16_v4_kernel(const __nv_bfloat16* __restrict__ in, float* __restrict__ out, int rows, int cols, int64_t ld, int rows_per_block,
                          int vpr_log2, float* __restrict__ partial, unsigned int* __restrict__ tickets) {
  colsum_v3_body<true>(in, out, rows, cols, ld, rows_per_block, vpr_log2, partial, tickets);
}

// per-device workspace of the column-sum kernel (tickets + block partials); kernels that use it are stream ordered
size_t g_colsum_ws_bytes_dev[64] = {};
void* g_colsum_ws_dev[64] = {};

TRB_API int trb_colsum_bf16(const void* in, float* out, int rows, int cols, int64_t ld, cudaStream_t stream) {
  if (rows == 0 || cols == 0) return 0;
  if (cols % 8 || ld % 8) return -12;
  const int nvec = cols / 8;
  const int xblocks = (nvec + 31) / 32;
  static const int variant = [] { const char* e = getenv("TRB_COLSUM"); return e ? atoi(e) : 4; }();
  if (variant == 3 || variant == 4) {
    int vpr_log2 = 5;
    while (vpr_log2 > 0 && (1 << (vpr_log2 - 1)) >= nvec) --vpr_log2;
    const int rpw = 32 >> vpr_log2;
    int yblocks = (4 * 148 + xblocks - 1) / xblocks;
    const bool clustered = variant == 4 && rows >= 2048;
    if (clustered) {
      if (yblocks > 256) yblocks = 256;  // 8-CTA clusters fold through DSMEM: 32 partials at most reach the last-block tail
    } else if (yblocks > 64) {
      yblocks = 64;  // the last block of a column group sums yblocks partials: keep that tail short
    }
    int rows_per_block = (rows + yblocks - 1) / yblocks;
    rows_per_block = (rows_per_block + 8 * rpw - 1) / (8 * rpw) * (8 * rpw);
    yblocks = (rows + rows_per_block - 1) / rows_per_block;
    if (clustered) yblocks = (yblocks + 7) / 8 * 8;  // trailing blocks past `rows` contribute zeros
    // workspace (per process, stream-ordered use: the dense backward runs on one stream): tickets then partial sums
    const size_t need = 4096 + sizeof(float) * (size_t) yblocks * (size_t) cols;
    int dev = 0;
    TRB_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) return -12;
    size_t& g_colsum_ws_bytes = g_colsum_ws_bytes_dev[dev];
    void*& g_colsum_ws = g_colsum_ws_dev[dev];
    if (need > g_colsum_ws_bytes) {
      if (g_colsum_ws) TRB_CUDA(cudaFree(g_colsum_ws));
      g_colsum_ws_bytes = need < (8u << 20) ? (8u << 20) : need;
      TRB_CUDA(cudaMalloc(&g_colsum_ws, g_colsum_ws_bytes));
      TRB_CUDA(cudaMemset(g_colsum_ws, 0, 4096));
    }
    if (xblocks > 1024) return -12;
    dim3 grid(xblocks, yblocks);
    if (clustered)
      trb_colsum_bf16_v4_kernel<<<grid, 256, 0, stream>>>((const __nv_bfloat16*) in, out, rows, cols, ld, rows_per_block, vpr_log2,
                                                          reinterpret_cast<float*>(static_cast<char*>(g_colsum_ws) + 4096),
                                                          reinterpret_cast<unsigned int*>(g_colsum_ws));
    else
      trb_colsum_bf16_v3_kernel<<<grid, 256, 0, stream>>>((const __nv_bfloat16*) in, out, rows, cols, ld, rows_per_block, vpr_log2,
                                                          reinterpret_cast<float*>(static_cast<char*>(g_colsum_ws) + 4096),
                                                          reinterpret_cast<unsigned int*>(g_colsum_ws));
    TRB_CHECK_LAUNCH();
    return 0;
  }
  TRB_CUDA(cudaMemsetAsync(out, 0, sizeof(float) * cols, stream));
  int rows_per_block = 256;
  while ((int64_t) xblocks * ((rows + rows_per_block - 1) / rows_per_block) > 2048 && rows_per_block < 65536) rows_per_block *= 2;
  dim3 grid(xblocks, (rows + rows_per_block - 1) / rows_per_block);
  trb_colsum_bf16_kernel<<<grid, 256, 0, stream>>>((const __nv_bfloat16*) in, out, rows, cols, ld, rows_per_block);
  TRB_CHECK_LAUNCH();
  return 0;
}

// ---------------------------------------------------------------------------------------------------------------------------
// fp32 master weights -> bf16 GEMM operands (K zero-padded to a multiple of 8) for ALL layers of an MLP in one launch. The per-layer
// `.to(bf16)` + `pad` pair cost 2-3 tiny kernels per layer in front of every GEMM of the captured forward (13 launch-latency-bound
// nodes on the critical chain of the DLRM step).
// ---------------------------------------------------------------------------------------------------------------------------
constexpr int kMaxCastTensors = 16;
struct CastPadTable {
  const float* src[kMaxCastTensors];
  __nv_bfloat16* dst[kMaxCastTensors];
  int32_t rows[kMaxCastTensors], K[kMaxCastTensors], Kp[kMaxCastTensors];
};

__global__ void __launch_bounds__(256) multi_cast_pad_bf16_kernel(const CastPadTable t) {
  const int l = blockIdx.y;
  const int K = t.K[l], Kp = t.Kp[l];
  const int64_t total = (int64_t) t.rows[l] * (Kp >> 1);  // one bf16x2 per thread-iteration
  const float* __restrict__ src = t.src[l];
  __nv_bfloat162* __restrict__ dst = reinterpret_cast<__nv_bfloat162*>(t.dst[l]);
  const int half = Kp >> 1;
  for (int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t) gridDim.x * blockDim.x) {
    const int r = (int) (i / half), c = (int) (i - (int64_t) r * half) * 2;
    const float a = c < K ? src[(int64_t) r * K + c] : 0.f;
    const float b = c + 1 < K ? src[(int64_t) r * K + c + 1] : 0.f;
    dst[i] = __floats2bfloat162_rn(a, b);
  }
}

TRB_API int trb_multi_cast_pad_bf16(const void* const* src, void* const* dst, const int32_t* rows, const int32_t* K, const int32_t* Kp, int n, cudaStream_t stream) {
  if (n < 1 || n > kMaxCastTensors) return -1;
  CastPadTable t;
  int64_t biggest = 0;
  for (int i = 0; i < n; ++i) {
    if (Kp[i] % 2 || Kp[i] < K[i]) return -2;
    t.src[i] = (const float*) src[i];
    t.dst[i] = (__nv_bfloat16*) dst[i];
    t.rows[i] = rows[i];
    t.K[i] = K[i];
    t.Kp[i] = Kp[i];
    const int64_t tot = (int64_t) rows[i] * (Kp[i] / 2);
    biggest = tot > biggest ? tot : biggest;
  }
  if (biggest == 0) return 0;
  int64_t bx = (biggest + 255) / 256;
  if (bx > 148 * 4) bx = 148 * 4;
  multi_cast_pad_bf16_kernel<<<dim3((unsigned) bx, (unsigned) n), 256, 0, stream>>>(t);
  TRB_CHECK_LAUNCH();
  return 0;
}

// ---- final reduction of the epilogue column sums: out[n] = sum_p ws[p][n], fixed order (deterministic) ---------------------------------
// grid.x = 64-column groups, 8 warps stride over the P partial rows, lanes read float2 (a warp reads 256 contiguous bytes per row).
__global__ void __launch_bounds__(256) colsum_partials_kernel(const float* __restrict__ ws, int P, int N, float* __restrict__ out) {
  __shared__ float2 part[8][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int col = blockIdx.x * 64 + lane * 2;
  float2 acc = make_float2(0.f, 0.f);
  if (col < N) {
    int r = warp;
    for (; r + 24 < P; r += 32) {
      const float2 a = *reinterpret_cast<const float2*>(ws + (int64_t) r * N + col);
      const float2 b = *reinterpret_cast<const float2*>(ws + (int64_t) (r + 8) * N + col);
      const float2 c = *reinterpret_cast<const float2*>(ws + (int64_t) (r + 16) * N + col);
      const float2 d = *reinterpret_cast<const float2*>(ws + (int64_t) (r + 24) * N + col);
      acc.x += (a.x + b.x) + (c.x + d.x);
      acc.y += (a.y + b.y) + (c.y + d.y);
    }
    for (; r < P; r += 8) {
      const float2 a = *reinterpret_cast<const float2*>(ws + (int64_t) r * N + col);
      acc.x += a.x;
      acc.y += a.y;
    }
  }
  part[warp][lane] = acc;
  __syncthreads();
  if (warp == 0 && col < N) {
    float2 t = part[0][lane];
#pragma unroll
    for (int w = 1; w < 8; ++w) {
      t.x += part[w][lane].x;
      t.y += part[w][lane].y;
    }
    *reinterpret_cast<float2*>(out + col) = t;
  }
}

TRB_API int trb_colsum_partials(const float* ws, int P, int N, float* out, cudaStream_t stream) {
  if (P <= 0 || N <= 0) return 0;
  if (N % 2) return -1;
  colsum_partials_kernel<<<(unsigned) ((N + 63) / 64), 256, 0, stream>>>(ws, P, N, out);
  TRB_CHECK_LAUNCH();
  return 0;
}
