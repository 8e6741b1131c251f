// DLRM dot-interaction forward / backward on tcgen05 tensor cores (sm_100a).
//
//   X_b = [dense_b ; sparse_b]            (R = F+1 rows, D columns, R <= 32, D == 128)
//   out_b = [dense_b , strict-upper-triangle(X_b X_b^T)]         (reference models/dlrm.py:210-222)
//
// The per-sample 27x27 products are far too small for a 128-row MMA, so FOUR samples are packed into
// one 128 x 128 operand tile (32 rows per sample, unused rows zero) and ONE tcgen05.mma group
// computes T T^T; only the four diagonal 32x32 blocks are read back (tcgen05.ld: TMEM lane == packed
// row, so warp w owns sample w). The 4x redundant FLOPs are free (the kernel is HBM-bound) and the
// operands never leave shared memory.
//
// Backward: dX_b = S_b X_b with S_b the symmetric 27x27 matrix rebuilt from the triangle of the
// incoming gradient. Tile math: D[128 x 128] = blockdiag(S)[128 x 128] . T[128 x 128] where the SAME
// shared-memory image of T (row-major, SWIZZLE_128B) is consumed as an MN-major B operand — no transpose.
//
// CTAs are small (128 threads, ~38 KB smem, 128 TMEM columns) so 4 fit per SM and hide the global
// load latency of each other; each CTA loops over groups of 4 samples.
#include "tcgen05.cuh"

using namespace trb;

namespace {

constexpr int kD = 128;          // embedding dim handled by this kernel
constexpr int kRowsPerSample = 32;
constexpr int kSamples = 4;      // per MMA tile
constexpr int kTileBytes = 128 * 64 * 2;  // one [128 x 64] bf16 K-block

struct InterParams {
  const void* dense;     // [B, D] bf16
  int64_t ld_dense;
  const void* sparse;    // [B, F*D] (bf16 or fp32), row pitch ld_sparse
  int64_t ld_sparse;
  int sparse_f32;
  void* out;             // fwd: [B, ld_out] bf16
  int64_t ld_out;
  int B, F;
  int out_cols;          // D + R(R-1)/2
  // backward
  const void* gout;      // [B, ld_out] bf16
  void* g_dense;         // [B, D] bf16
  int64_t ld_gdense;
  void* g_sparse;        // [B, F*D] (bf16 or fp32)
  int64_t ld_gsparse;
  int gsparse_f32;
};

__device__ __forceinline__ int tri_offset(int i, int R) { return i * R - (i * (i + 1)) / 2; }  // first (i, i+1) slot

// Load the packed operand tile T (4 samples x 32 rows x 128 cols) into two SWIZZLE_128B K-blocks.
// Each thread moves 8 elements (16 B in smem). 128 threads: 4 samples x R rows x 16 chunks.
__device__ __forceinline__ void load_tile(const InterParams& p, uint8_t* tile, int b0, int tid) {
  const int R = p.F + 1;
  const int chunks = kSamples * R * 16;  // 16-byte chunks (8 bf16)
  for (int c = tid; c < chunks; c += 128) {
    const int s = c / (R * 16);
    const int rem = c - s * R * 16;
    const int i = rem >> 4;
    const int ch = rem & 15;  // chunk along D: columns [8*ch, 8*ch+8)
    const int b = b0 + s;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (b < p.B) {
      if (i == 0) {
        v = *reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(p.dense) + (int64_t) b * p.ld_dense + ch * 8);
      } else if (!p.sparse_f32) {
        v = *reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(p.sparse) + (int64_t) b * p.ld_sparse + (i - 1) * kD + ch * 8);
      } else {
        const float* src = reinterpret_cast<const float*>(p.sparse) + (int64_t) b * p.ld_sparse + (i - 1) * kD + ch * 8;
        const float4 f0 = *reinterpret_cast<const float4*>(src), f1 = *reinterpret_cast<const float4*>(src + 4);
        __nv_bfloat162 h0 = __floats2bfloat162_rn(f0.x, f0.y), h1 = __floats2bfloat162_rn(f0.z, f0.w);
        __nv_bfloat162 h2 = __floats2bfloat162_rn(f1.x, f1.y), h3 = __floats2bfloat162_rn(f1.z, f1.w);
        v.x = *reinterpret_cast<uint32_t*>(&h0); v.y = *reinterpret_cast<uint32_t*>(&h1);
        v.z = *reinterpret_cast<uint32_t*>(&h2); v.w = *reinterpret_cast<uint32_t*>(&h3);
      }
    }
    const int r = s * kRowsPerSample + i;
    const int kb = ch >> 3;  // which 64-column K-block
    const int k = (ch & 7) * 8;
    *reinterpret_cast<uint4*>(tile + kb * kTileBytes + sw128_offset(r, k)) = v;
  }
}

__global__ void __launch_bounds__(128, 4) interaction_fwd_kernel(const InterParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t) 1023);
  uint8_t* tile = smem;                                            // 2 x 16 KB
  __nv_bfloat16* stage = reinterpret_cast<__nv_bfloat16*>(smem + 2 * kTileBytes);  // [4][ld_out] output staging
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 2 * kTileBytes + kSamples * p.ld_out * 2);
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bar + 1);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int R = p.F + 1;

  // zero the operand tile once: padding rows (i >= R) are never written again
  for (int i = tid; i < 2 * kTileBytes / 16; i += 128) reinterpret_cast<uint4*>(tile)[i] = make_uint4(0, 0, 0, 0);
  if (tid == 0) {
    mbar_init(bar, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(tmem_ptr, 128);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = *tmem_ptr;
  constexpr uint32_t idesc = make_idesc_major(128, 128, 0, 0);
  uint32_t phase = 0;
  const int groups = (p.B + kSamples - 1) / kSamples;
  for (int g = blockIdx.x; g < groups; g += gridDim.x) {
    const int b0 = g * kSamples;
    load_tile(p, tile, b0, tid);
    fence_proxy_async();  // generic-proxy smem writes -> visible to the tensor-core (async) proxy
    __syncthreads();
    if (warp == 0) {
      if (elect_one()) {
        const uint32_t a0 = smem_u32(tile);
#pragma unroll
        for (int kb = 0; kb < 2; ++kb) {
          const uint64_t desc = make_kmajor_desc(a0 + kb * kTileBytes);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16(tmem_d, desc + (uint64_t) (2 * k), desc + (uint64_t) (2 * k), idesc, (kb | k) != 0);
        }
        umma_commit(bar);
      }
      __syncwarp();
    }
    mbar_wait(bar, phase);
    phase ^= 1;
    tc_fence_after();
    // warp w <-> sample b0+w ; lane i <-> row i ; columns [32w, 32w+32) hold Z_w[i][0..31]
    uint32_t z[32];
    tmem_ld_32x32(tmem_d + ((uint32_t) (warp * 32) << 16) + (uint32_t) (warp * 32), z);
    __nv_bfloat16* srow = stage + warp * p.ld_out;
    if (lane < R) {
      const int base = kD + tri_offset(lane, R);
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (j > lane && j < R) srow[base + (j - lane - 1)] = __float2bfloat16(__uint_as_float(z[j]));
    }
    // dense passthrough + zero padding of the row tail
    for (int c = lane; c < kD / 8; c += 32) {
      const int r = warp * kRowsPerSample;
      reinterpret_cast<uint4*>(srow)[c] = *reinterpret_cast<const uint4*>(tile + (c >> 3) * kTileBytes + sw128_offset(r, (c & 7) * 8));
    }
    for (int c = p.out_cols + lane; c < p.ld_out; c += 32) srow[c] = __float2bfloat16(0.f);
    tc_fence_before();
    __syncthreads();  // staging complete; TMEM + operand tile free for the next group
    const int vec_per_row = (int) (p.ld_out >> 3);
    for (int v = tid; v < kSamples * vec_per_row; v += 128) {
      const int s = v / vec_per_row, c = v - s * vec_per_row;
      if (b0 + s < p.B)
        reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.out) + (int64_t) (b0 + s) * p.ld_out)[c] =
            reinterpret_cast<const uint4*>(stage + s * p.ld_out)[c];
    }
    __syncthreads();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_d, 128);
}

__global__ void __launch_bounds__(128, 3) interaction_bwd_kernel(const InterParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t) 1023);
  uint8_t* tileT = smem;                     // packed X (2 K-blocks), used as MN-major B operand
  uint8_t* tileS = smem + 2 * kTileBytes;    // block-diagonal S (2 K-blocks), K-major A operand
  __nv_bfloat16* gstage = reinterpret_cast<__nv_bfloat16*>(smem + 4 * kTileBytes);  // [4][ld_out] incoming gradient rows
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 4 * kTileBytes + kSamples * p.ld_out * 2);
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bar + 1);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int R = p.F + 1;

  for (int i = tid; i < 4 * kTileBytes / 16; i += 128) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (tid == 0) {
    mbar_init(bar, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(tmem_ptr, 128);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = *tmem_ptr;
  // A: K-major (block-diagonal S), B: MN-major (packed X, [k = packed row][n = dim])
  constexpr uint32_t idesc = make_idesc_major(128, 128, 0, 1);
  uint32_t phase = 0;
  const int groups = (p.B + kSamples - 1) / kSamples;
  for (int g = blockIdx.x; g < groups; g += gridDim.x) {
    const int b0 = g * kSamples;
    load_tile(p, tileT, b0, tid);
    // stage the 4 incoming gradient rows with coalesced 16 B loads
    {
      const int vec_per_row = (int) (p.ld_out >> 3);
      for (int v = tid; v < kSamples * vec_per_row; v += 128) {
        const int sidx = v / vec_per_row, c = v - sidx * vec_per_row;
        uint4 val = make_uint4(0, 0, 0, 0);
        if (b0 + sidx < p.B) val = reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(p.gout) + (int64_t) (b0 + sidx) * p.ld_out)[c];
        reinterpret_cast<uint4*>(gstage + sidx * p.ld_out)[c] = val;
      }
    }
    __syncthreads();
    // S rows: thread (warp = sample, lane = i) writes S[i][j] for all j (symmetric, zero diagonal)
    {
      const int b = b0 + warp;
      const int r = warp * kRowsPerSample + lane;
      const __nv_bfloat16* grow = gstage + warp * p.ld_out + kD;
      if (lane < R) {
        for (int j = 0; j < R; ++j) {
          __nv_bfloat16 v = __float2bfloat16(0.f);
          if (b < p.B && j != lane) {
            const int lo = min(lane, j), hi = max(lane, j);
            v = grow[tri_offset(lo, R) + (hi - lo - 1)];
          }
          const int k = warp * kRowsPerSample + j;
          *reinterpret_cast<__nv_bfloat16*>(tileS + (k >> 6) * kTileBytes + sw128_offset(r, k & 63)) = v;
        }
      }
    }
    fence_proxy_async();
    __syncthreads();
    if (warp == 0) {
      if (elect_one()) {
        const uint32_t aS = smem_u32(tileS), aT = smem_u32(tileT);
#pragma unroll
        for (int kb = 0; kb < 2; ++kb) {
          const uint64_t adesc = make_kmajor_desc(aS + kb * kTileBytes);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            // B: K index = packed row (kb*64 + k*16 ...): 16 rows x 128 B = 2048 B per UMMA_K step inside
            // the [128 rows x 128 B] image; the two 64-column halves (MN atoms) are kTileBytes apart.
            const uint64_t bdesc = make_mnmajor_desc(aT + (kb * 64 + k * 16) * 128, kTileBytes, 1024);
            umma_bf16(tmem_d, adesc + (uint64_t) (2 * k), bdesc, idesc, (kb | k) != 0);
          }
        }
        umma_commit(bar);
      }
      __syncwarp();
    }
    mbar_wait(bar, phase);
    phase ^= 1;
    tc_fence_after();
    const int b = b0 + warp;
    const bool ok = b < p.B && lane < R;
#pragma unroll 1
    for (int c = 0; c < 4; ++c) {
      uint32_t z[32];
      tmem_ld_32x32(tmem_d + ((uint32_t) (warp * 32) << 16) + (uint32_t) (c * 32), z);
      if (!ok) continue;
      float v[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(z[j]);
      if (lane == 0) {
        // dense row: add the pass-through gradient of out[:, :D]
        const __nv_bfloat16* gd = gstage + warp * p.ld_out + c * 32;
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] += __bfloat162float(gd[j]);
        __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(p.g_dense) + (int64_t) b * p.ld_gdense + c * 32;
#pragma unroll
        for (int j = 0; j < 32; j += 2) *reinterpret_cast<__nv_bfloat162*>(dst + j) = __floats2bfloat162_rn(v[j], v[j + 1]);
      } else if (p.gsparse_f32) {
        float* dst = reinterpret_cast<float*>(p.g_sparse) + (int64_t) b * p.ld_gsparse + (lane - 1) * kD + c * 32;
#pragma unroll
        for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(dst + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
      } else {
        __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(p.g_sparse) + (int64_t) b * p.ld_gsparse + (lane - 1) * kD + c * 32;
#pragma unroll
        for (int j = 0; j < 32; j += 2) *reinterpret_cast<__nv_bfloat162*>(dst + j) = __floats2bfloat162_rn(v[j], v[j + 1]);
      }
    }
    tc_fence_before();
    __syncthreads();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_d, 128);
}

int g_sms = 0;
int num_sms() {
  if (g_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_sms, cudaDevAttrMultiProcessorCount, dev);
  }
  return g_sms;
}

}  // namespace

// out[b] = [dense_b, triu(X_b X_b^T, 1)] with ld_out >= D + R(R-1)/2 (multiple of 8, tail zero filled)
TRB_API int trb_interaction_fwd(const void* dense, int64_t ld_dense, const void* sparse, int64_t ld_sparse, int sparse_f32, void* out,
                                int64_t ld_out, int B, int F, int D, cudaStream_t stream) {
  if (D != kD || F + 1 > 32 || F < 1) return -20;
  if (B == 0) return 0;
  InterParams p = {};
  p.dense = dense; p.ld_dense = ld_dense; p.sparse = sparse; p.ld_sparse = ld_sparse; p.sparse_f32 = sparse_f32;
  p.out = out; p.ld_out = ld_out; p.B = B; p.F = F; p.out_cols = D + (F + 1) * F / 2;
  if (ld_out % 8 || ld_out < p.out_cols) return -21;
  const int smem = 2 * kTileBytes + kSamples * (int) ld_out * 2 + 64 + 1024;
  static bool cfg = false;
  if (!cfg) {
    TRB_CUDA(cudaFuncSetAttribute(interaction_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    cfg = true;
  }
  const int groups = (B + kSamples - 1) / kSamples;
  const int grid = groups < 4 * num_sms() ? groups : 4 * num_sms();
  interaction_fwd_kernel<<<grid, 128, smem, stream>>>(p);
  TRB_CHECK_LAUNCH();
  return 0;
}

TRB_API int trb_interaction_bwd(const void* dense, int64_t ld_dense, const void* sparse, int64_t ld_sparse, int sparse_f32, const void* gout,
                                int64_t ld_out, void* g_dense, int64_t ld_gdense, void* g_sparse, int64_t ld_gsparse, int gsparse_f32, int B,
                                int F, int D, cudaStream_t stream) {
  if (D != kD || F + 1 > 32 || F < 1) return -20;
  if (B == 0) return 0;
  InterParams p = {};
  p.dense = dense; p.ld_dense = ld_dense; p.sparse = sparse; p.ld_sparse = ld_sparse; p.sparse_f32 = sparse_f32;
  p.gout = gout; p.ld_out = ld_out; p.B = B; p.F = F; p.out_cols = D + (F + 1) * F / 2;
  p.g_dense = g_dense; p.ld_gdense = ld_gdense; p.g_sparse = g_sparse; p.ld_gsparse = ld_gsparse; p.gsparse_f32 = gsparse_f32;
  const int smem = 4 * kTileBytes + kSamples * (int) ld_out * 2 + 64 + 1024;
  static bool cfg = false;
  if (!cfg) {
    TRB_CUDA(cudaFuncSetAttribute(interaction_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    cfg = true;
  }
  const int groups = (B + kSamples - 1) / kSamples;
  const int grid = groups < 3 * num_sms() ? groups : 3 * num_sms();
  interaction_bwd_kernel<<<grid, 128, smem, stream>>>(p);
  TRB_CHECK_LAUNCH();
  return 0;
}
