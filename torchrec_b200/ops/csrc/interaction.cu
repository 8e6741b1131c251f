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
#include <cstdlib>

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


// ================================================================================================================
// Pipelined bf16 kernels (the production path): operand tiles arrive through cp.async (16 B LDGSTS, every thread has
// all of its loads in flight at once) into a DOUBLE-BUFFERED smem image, so the global->smem traffic of group g+1
// overlaps the MMA + epilogue of group g; results leave through a swizzled smem staging area with fully coalesced
// 16 B stores. The first-generation kernels above (synchronous loads, per-thread 4 B stores) ran at 17-27 % of the HBM
// roofline (profiles/step_breakdown_1gpu_r1.md) and remain as the fp32-I/O fallback.
// ================================================================================================================
__device__ __forceinline__ void cp_async16(uint32_t smem_addr, const void* g, bool valid) {
  const int sz = valid ? 16 : 0;  // src-size 0 -> zero fill
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_addr), "l"(g), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// issue (do not wait for) the loads of the packed tile of samples [b0, b0+4): bf16 dense + bf16 sparse.
// Thread t always moves column chunk (t & 15) of rows (t >> 4) + 8k, so the swizzle term and all strides are per-thread
// constants (the first version recomputed them per chunk and was instruction-bound, profiles/ncu_interaction_bwd_pipe_kernel_r1.md).
// EXTRA: also load gout[b][0:128] (the pass-through gradient of the dense row) as packed row R, so the backward MMA adds it
// in fp32 through S[0][R] = 1 instead of a lane-0 epilogue special case.
template <bool EXTRA>
__device__ __forceinline__ void issue_tile_loads(const InterParams& p, uint8_t* tile, int b0, int tid) {
  const int R = p.F + 1;
  const int ch = tid & 15, i0 = tid >> 4;
  const uint32_t tbase = smem_u32(tile) + (ch >> 3) * kTileBytes + ((((ch & 7) ^ (i0 & 7)) & 7) << 4);
  const __nv_bfloat16* dense = reinterpret_cast<const __nv_bfloat16*>(p.dense) + ch * 8;
  const __nv_bfloat16* sparse = reinterpret_cast<const __nv_bfloat16*>(p.sparse) + ch * 8 - kD;  // row i (>= 1) lives at + i * kD
  const __nv_bfloat16* gout = reinterpret_cast<const __nv_bfloat16*>(p.gout) + ch * 8;
#pragma unroll
  for (int s = 0; s < kSamples; ++s) {
    const int b = b0 + s;
    const bool ok = b < p.B;
    const int64_t bb = ok ? b : 0;
    const __nv_bfloat16* srow = sparse + bb * p.ld_sparse;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int i = i0 + 8 * k;
      const uint32_t dst = tbase + (uint32_t) ((s * kRowsPerSample + i) * 128);
      if (i < R) {
        cp_async16(dst, (i == 0) ? dense + bb * p.ld_dense : srow + (int64_t) i * kD, ok);
      } else if (EXTRA && i == R) {
        cp_async16(dst, gout + bb * p.ld_out, ok);
      }
    }
  }
}

// issue the loads of the 4 incoming-gradient rows (bwd)
__device__ __forceinline__ void issue_gout_loads(const InterParams& p, __nv_bfloat16* gstage, int b0, int tid) {
  const int vec_per_row = (int) (p.ld_out >> 3);
  const uint32_t gbase = smem_u32(gstage);
#pragma unroll
  for (int sidx = 0; sidx < kSamples; ++sidx) {
    const bool ok = b0 + sidx < p.B;
    const __nv_bfloat16* src = reinterpret_cast<const __nv_bfloat16*>(p.gout) + (int64_t) (ok ? b0 + sidx : 0) * p.ld_out;
    for (int c = tid; c < vec_per_row; c += 128) cp_async16(gbase + (uint32_t) (sidx * p.ld_out + c * 8) * 2, src + c * 8, ok);
  }
}

__global__ void __launch_bounds__(128, 3) interaction_fwd_pipe_kernel(const InterParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t) 1023);
  uint8_t* tiles = smem;                                                               // 2 buffers x (2 x 16 KB)
  __nv_bfloat16* stage = reinterpret_cast<__nv_bfloat16*>(smem + 4 * kTileBytes);      // [4][ld_out]
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 4 * kTileBytes + kSamples * p.ld_out * 2);
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bar + 1);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int R = p.F + 1;
  for (int i = tid; i < 4 * kTileBytes / 16; i += 128) reinterpret_cast<uint4*>(tiles)[i] = make_uint4(0, 0, 0, 0);
  if (tid == 0) { mbar_init(bar, 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc(tmem_ptr, 128);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = *tmem_ptr;
  constexpr uint32_t idesc = make_idesc_major(128, 128, 0, 0);
  uint32_t phase = 0;
  const int groups = (p.B + kSamples - 1) / kSamples;
  int cur = 0;
  if ((int) blockIdx.x < groups) issue_tile_loads<false>(p, tiles, blockIdx.x * kSamples, tid);
  cp_async_commit();
  for (int g = blockIdx.x; g < groups; g += gridDim.x, cur ^= 1) {
    const int b0 = g * kSamples;
    uint8_t* tile = tiles + cur * 2 * kTileBytes;
    cp_async_wait_all();
    fence_proxy_async();
    __syncthreads();  // tile[cur] landed; everyone is done with tile[cur^1] and with `stage`
    if (g + (int) gridDim.x < groups) issue_tile_loads<false>(p, tiles + (cur ^ 1) * 2 * kTileBytes, (g + gridDim.x) * kSamples, tid);
    cp_async_commit();
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
    // dense passthrough + zero tail while the MMA runs
    __nv_bfloat16* srow = stage + warp * p.ld_out;
    for (int c = lane; c < kD / 8; c += 32) {
      const int r = warp * kRowsPerSample;
      reinterpret_cast<uint4*>(srow)[c] = *reinterpret_cast<const uint4*>(tile + (c >> 3) * kTileBytes + sw128_offset(r, (c & 7) * 8));
    }
    for (int c = p.out_cols + lane; c < p.ld_out; c += 32) srow[c] = __float2bfloat16(0.f);
    mbar_wait(bar, phase);
    phase ^= 1;
    tc_fence_after();
    uint32_t z[32];
    tmem_ld_32x32(tmem_d + ((uint32_t) (warp * 32) << 16) + (uint32_t) (warp * 32), z);
    if (lane < R) {
      const int base = kD + tri_offset(lane, R);
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (j > lane && j < R) srow[base + (j - lane - 1)] = __float2bfloat16(__uint_as_float(z[j]));
    }
    tc_fence_before();
    __syncthreads();
    const int vec_per_row = (int) (p.ld_out >> 3);
#pragma unroll
    for (int s = 0; s < kSamples; ++s) {
      if (b0 + s >= p.B) break;
      uint4* orow = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.out) + (int64_t) (b0 + s) * p.ld_out);
      const uint4* srow4 = reinterpret_cast<const uint4*>(stage + s * p.ld_out);
      for (int c = tid; c < vec_per_row; c += 128) orow[c] = srow4[c];
    }
  }
  cp_async_wait_all();
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_d, 128);
}

template <bool EXTRA>
__global__ void __launch_bounds__(128, 2) interaction_bwd_pipe_kernel(const InterParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t) 1023);
  uint8_t* tiles = smem;                       // T: 2 buffers x 32 KB (MN-major B operand)
  uint8_t* tileS = smem + 4 * kTileBytes;      // S (K-major A operand, 32 KB); re-used as the output staging area after the MMA
  __nv_bfloat16* gst = reinterpret_cast<__nv_bfloat16*>(smem + 6 * kTileBytes);  // 2 x [4][ld_out]
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 6 * kTileBytes + 2 * kSamples * p.ld_out * 2);
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bar + 1);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int R = p.F + 1;
  for (int i = tid; i < 4 * kTileBytes / 16; i += 128) reinterpret_cast<uint4*>(tiles)[i] = make_uint4(0, 0, 0, 0);
  if (tid == 0) { mbar_init(bar, 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc(tmem_ptr, 128);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = *tmem_ptr;
  constexpr uint32_t idesc = make_idesc_major(128, 128, 0, 1);
  uint32_t phase = 0;
  const int groups = (p.B + kSamples - 1) / kSamples;
  const int gsz = kSamples * (int) p.ld_out;
  int cur = 0;
  if ((int) blockIdx.x < groups) {
    issue_tile_loads<EXTRA>(p, tiles, blockIdx.x * kSamples, tid);
    issue_gout_loads(p, gst, blockIdx.x * kSamples, tid);
  }
  cp_async_commit();
  const int r = warp * kRowsPerSample + lane;  // my row of the packed tile (sample = warp, row-in-sample = lane)
  for (int g = blockIdx.x; g < groups; g += gridDim.x, cur ^= 1) {
    const int b0 = g * kSamples;
    uint8_t* tileT = tiles + cur * 2 * kTileBytes;
    const __nv_bfloat16* gstage = gst + cur * gsz;
    cp_async_wait_all();
    fence_proxy_async();
    __syncthreads();  // T[cur] / gout[cur] landed; copy-out of the previous group finished (staging == tileS is free)
    if (g + (int) gridDim.x < groups) {
      issue_tile_loads<EXTRA>(p, tiles + (cur ^ 1) * 2 * kTileBytes, (g + gridDim.x) * kSamples, tid);
      issue_gout_loads(p, gst + (cur ^ 1) * gsz, (g + gridDim.x) * kSamples, tid);
    }
    cp_async_commit();
    // ---- S row of this thread: S[i][j] = g[triangle(min,max)] (0 on the diagonal / padding); the whole 256 B row of
    //      blockdiag(S) is written (zeros outside the sample's own 32-column block) because the staging re-use clobbers it
    {
      const __nv_bfloat16* grow = gstage + warp * p.ld_out + kD;
      const bool live = lane < R && (b0 + warp) < p.B;
      // S[i][j] = g[tri(min) + (max - min - 1)]: for j < i the offset walks down the column (stride R - j - 2), for j > i it
      // is contiguous in the row of i — no multiplications inside the loop
      uint32_t packed[16];
      int off_lo = lane - 1;
      const int off_hi = tri_offset(lane, R) - (lane + 1);
#pragma unroll
      for (int j2 = 0; j2 < 16; ++j2) {
        __nv_bfloat16 v[2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int j = 2 * j2 + h;
          v[h] = __float2bfloat16(0.f);
          if (live && j < R && j != lane) v[h] = grow[(j < lane) ? off_lo : off_hi + j];
          if (EXTRA && j == R && lane == 0) v[h] = __float2bfloat16(1.f);  // routes packed row R (pass-through gradient) into the dense row
          off_lo += R - j - 2;
        }
        __nv_bfloat162 hh = __halves2bfloat162(v[0], v[1]);
        packed[j2] = *reinterpret_cast<uint32_t*>(&hh);
      }
      // row r: K-block kb holds k in [64 kb, 64 kb + 64) as 8 chunks of 16 B; own block = k in [32 warp, 32 warp + 32)
      const int own_kb = warp >> 1, own_c0 = (warp & 1) * 4;
#pragma unroll
      for (int kb = 0; kb < 2; ++kb) {
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          uint4 val = make_uint4(0, 0, 0, 0);
          if (kb == own_kb && c >= own_c0 && c < own_c0 + 4) {
            const int q = (c - own_c0) * 4;
            val = make_uint4(packed[q], packed[q + 1], packed[q + 2], packed[q + 3]);
          }
          *reinterpret_cast<uint4*>(tileS + kb * kTileBytes + r * 128 + (((c ^ (r & 7)) & 7) << 4)) = val;
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
    // ---- epilogue: TMEM -> bf16 -> swizzled staging ([128 rows][16 chunks of 16 B], chunk ^= row & 15) -----------------
    uint8_t* ostage = tileS;
#pragma unroll 1
    for (int c = 0; c < 4; ++c) {
      uint32_t z[32];
      tmem_ld_32x32(tmem_d + ((uint32_t) (warp * 32) << 16) + (uint32_t) (c * 32), z);
      float v[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(z[j]);
      if (!EXTRA && lane == 0) {  // R == 32: no spare packed row, add the pass-through gradient of out[:, :D] here
        const __nv_bfloat16* gd = gstage + warp * p.ld_out + c * 32;
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] += __bfloat162float(gd[j]);
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        uint4 o;
        __nv_bfloat162 h0 = __floats2bfloat162_rn(v[8 * q], v[8 * q + 1]), h1 = __floats2bfloat162_rn(v[8 * q + 2], v[8 * q + 3]);
        __nv_bfloat162 h2 = __floats2bfloat162_rn(v[8 * q + 4], v[8 * q + 5]), h3 = __floats2bfloat162_rn(v[8 * q + 6], v[8 * q + 7]);
        o.x = *reinterpret_cast<uint32_t*>(&h0); o.y = *reinterpret_cast<uint32_t*>(&h1);
        o.z = *reinterpret_cast<uint32_t*>(&h2); o.w = *reinterpret_cast<uint32_t*>(&h3);
        const int chunk = c * 4 + q;
        *reinterpret_cast<uint4*>(ostage + r * 256 + (((chunk ^ lane) & 15) << 4)) = o;
      }
    }
    tc_fence_before();
    __syncthreads();
    // ---- coalesced copy-out: 16 consecutive threads write one 256 B gradient row; thread t owns chunk (t & 15) of rows (t >> 4) + 8k
    {
      const int ch = tid & 15, i0 = tid >> 4;
      __nv_bfloat16* gd_base = reinterpret_cast<__nv_bfloat16*>(p.g_dense) + ch * 8;
      __nv_bfloat16* gs_base = reinterpret_cast<__nv_bfloat16*>(p.g_sparse) + ch * 8 - kD;
#pragma unroll
      for (int s4 = 0; s4 < kSamples; ++s4) {
        const int b = b0 + s4;
        if (b >= p.B) break;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int i = i0 + 8 * k;
          if (i >= R) break;
          const uint4 val = *reinterpret_cast<const uint4*>(ostage + (s4 * kRowsPerSample + i) * 256 + (((ch ^ i) & 15) << 4));
          __nv_bfloat16* dst = (i == 0) ? gd_base + (int64_t) b * p.ld_gdense : gs_base + (int64_t) b * p.ld_gsparse + (int64_t) i * kD;
          *reinterpret_cast<uint4*>(dst) = val;
        }
      }
    }
  }
  cp_async_wait_all();
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
  const int groups = (B + kSamples - 1) / kSamples;
  if (!sparse_f32 && ld_dense % 8 == 0 && ld_sparse % 8 == 0 && !getenv("TRB_INTERACTION_LEGACY")) {
    const int smem_p = 4 * kTileBytes + kSamples * (int) ld_out * 2 + 64 + 1024;
    static bool cfg_p = false;
    if (!cfg_p) {
      TRB_CUDA(cudaFuncSetAttribute(interaction_fwd_pipe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
      cfg_p = true;
    }
    const int grid_p = groups < 3 * num_sms() ? groups : 3 * num_sms();
    interaction_fwd_pipe_kernel<<<grid_p, 128, smem_p, stream>>>(p);
    TRB_CHECK_LAUNCH();
    return 0;
  }
  const int smem = 2 * kTileBytes + kSamples * (int) ld_out * 2 + 64 + 1024;
  static bool cfg = false;
  if (!cfg) {
    TRB_CUDA(cudaFuncSetAttribute(interaction_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    cfg = true;
  }
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
  const int groups = (B + kSamples - 1) / kSamples;
  if (!sparse_f32 && !gsparse_f32 && ld_dense % 8 == 0 && ld_sparse % 8 == 0 && ld_gdense % 8 == 0 && ld_gsparse % 8 == 0 && ld_out % 8 == 0 &&
      !getenv("TRB_INTERACTION_LEGACY")) {
    const int smem_p = 6 * kTileBytes + 2 * kSamples * (int) ld_out * 2 + 64 + 1024;
    static bool cfg_p = false;
    if (!cfg_p) {
      TRB_CUDA(cudaFuncSetAttribute(interaction_bwd_pipe_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024));
      TRB_CUDA(cudaFuncSetAttribute(interaction_bwd_pipe_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024));
      cfg_p = true;
    }
    const int grid_p = groups < 2 * num_sms() ? groups : 2 * num_sms();
    if (F + 1 < 32) interaction_bwd_pipe_kernel<true><<<grid_p, 128, smem_p, stream>>>(p);
    else interaction_bwd_pipe_kernel<false><<<grid_p, 128, smem_p, stream>>>(p);
    TRB_CHECK_LAUNCH();
    return 0;
  }
  const int smem = 4 * kTileBytes + kSamples * (int) ld_out * 2 + 64 + 1024;
  static bool cfg = false;
  if (!cfg) {
    TRB_CUDA(cudaFuncSetAttribute(interaction_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    cfg = true;
  }
  const int grid = groups < 3 * num_sms() ? groups : 3 * num_sms();
  interaction_bwd_kernel<<<grid, 128, smem, stream>>>(p);
  TRB_CHECK_LAUNCH();
  return 0;
}
