// Jagged / sparse-data kernels used by KJT manipulation on the GPU (no host sync, stream ordered).
// Parity: fbgemm permute_2D_sparse_data as used by KeyedJaggedTensor.permute / dist_init
// (reference torchrec/sparse/jagged_tensor.py:2898, 3353).
#include "common.cuh"

// One 8-lane group per output segment (p, b): copy values[in_off[perm[p]*B+b] ...] -> out[out_off[p*B+b] ...].
template <typename T, typename WT>
__global__ void __launch_bounds__(256)
permute_2d_data_kernel(const int32_t* __restrict__ permute, int P, int B, const int64_t* __restrict__ in_off, const int64_t* __restrict__ out_off,
                       const T* __restrict__ values, T* __restrict__ out_values, const WT* __restrict__ weights, WT* __restrict__ out_weights) {
  const int lig = threadIdx.x & 7;
  const int64_t seg = ((int64_t) blockIdx.x * blockDim.x + threadIdx.x) >> 3;
  if (seg >= (int64_t) P * B) return;
  const int p = (int) (seg / B);
  const int b = (int) (seg - (int64_t) p * B);
  const int64_t src_seg = (int64_t) permute[p] * B + b;
  const int64_t s0 = in_off[src_seg];
  const int64_t n = in_off[src_seg + 1] - s0;
  const int64_t d0 = out_off[seg];
  for (int64_t i = lig; i < n; i += 8) {
    out_values[d0 + i] = values[s0 + i];
    if (weights != nullptr) out_weights[d0 + i] = weights[s0 + i];
  }
}

TRB_API int trb_permute_2d_data(const int32_t* permute, int P, int B, const int64_t* in_off, const int64_t* out_off, const void* values,
                                void* out_values, int value_bytes, const void* weights, void* out_weights, int weight_bytes,
                                cudaStream_t stream) {
  const int64_t segs = (int64_t) P * B;
  if (segs == 0) return 0;
  const int threads = 256;
  const unsigned blocks = (unsigned) ((segs * 8 + threads - 1) / threads);
  if (weights != nullptr && weight_bytes != 4) return -30;
#define TRB_P2D(T)                                                                                                                 \
  permute_2d_data_kernel<T, float><<<blocks, threads, 0, stream>>>(permute, P, B, in_off, out_off, (const T*) values, (T*) out_values, \
                                                                    (const float*) weights, (float*) out_weights)
  if (value_bytes == 8) TRB_P2D(int64_t);
  else if (value_bytes == 4) TRB_P2D(int32_t);
  else if (value_bytes == 2) TRB_P2D(int16_t);
  else return -31;
#undef TRB_P2D
  TRB_CHECK_LAUNCH();
  return 0;
}

// lengths [F*B] -> per-segment exclusive offsets is done with torch.cumsum; this kernel expands
// offsets into "position inside the segment" (fbgemm offsets_range) for position-weighted features.
__global__ void __launch_bounds__(256) offsets_range_kernel(const int64_t* __restrict__ offsets, int64_t n_seg, int64_t total, int64_t* __restrict__ out) {
  const int64_t seg = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
  if (seg >= n_seg) return;
  const int64_t s = offsets[seg];
  const int64_t e = seg + 1 < n_seg ? offsets[seg + 1] : total;
  for (int64_t i = s; i < e; ++i) out[i] = i - s;
}

TRB_API int trb_offsets_range(const int64_t* offsets, int64_t n_seg, int64_t total, int64_t* out, cudaStream_t stream) {
  if (n_seg == 0 || total == 0) return 0;
  offsets_range_kernel<<<(unsigned) ((n_seg + 255) / 256), 256, 0, stream>>>(offsets, n_seg, total, out);
  TRB_CHECK_LAUNCH();
  return 0;
}

// ---------------------------------------------------------------------------------------------------------------
// Row mover for the software-managed embedding cache (UVM_CACHING): dst[dst_idx[i]] = src[src_idx[i]] for i < n, rows of
// `row_bytes` (multiple of 4). Either side may be pinned HOST memory addressed zero-copy over PCIe / NVLink-C2C (the
// backing store of tables larger than HBM) or device memory (the cache). A null index array means the identity.
// Parity: the row movement inside fbgemm's lxu_cache_populate / lxu_cache_flush (reference call sites
// distributed/embedding_lookup.py:714-767, batched_embedding_kernel.py `prefetch` / `flush`).
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) trb_row_copy_kernel(uint8_t* __restrict__ dst, const int64_t* __restrict__ dst_idx, const uint8_t* __restrict__ src,
                                                           const int64_t* __restrict__ src_idx, int64_t n, int row_words) {
  // one warp per row for wide rows; consecutive lanes move consecutive 4-byte words (coalesced on both sides)
  const int lane = threadIdx.x & 31;
  const int64_t r = ((int64_t) blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (r >= n) return;
  const int64_t d = dst_idx ? dst_idx[r] : r;
  const int64_t s = src_idx ? src_idx[r] : r;
  if (d < 0 || s < 0) return;
  const uint32_t* sp = reinterpret_cast<const uint32_t*>(src) + s * row_words;
  uint32_t* dp = reinterpret_cast<uint32_t*>(dst) + d * row_words;
  if ((row_words & 3) == 0 && ((reinterpret_cast<uintptr_t>(sp) | reinterpret_cast<uintptr_t>(dp)) & 15) == 0) {
    for (int w = lane; w < (row_words >> 2); w += 32) reinterpret_cast<uint4*>(dp)[w] = reinterpret_cast<const uint4*>(sp)[w];
  } else {
    for (int w = lane; w < row_words; w += 32) dp[w] = sp[w];
  }
}

TRB_API int trb_row_copy(void* dst, const int64_t* dst_idx, const void* src, const int64_t* src_idx, int64_t n, int64_t row_bytes, cudaStream_t stream) {
  if (n <= 0) return 0;
  if (row_bytes % 4) return -2;
  const int threads = 256;
  const unsigned blocks = (unsigned) ((n * 32 + threads - 1) / threads);
  trb_row_copy_kernel<<<blocks, threads, 0, stream>>>((uint8_t*) dst, dst_idx, (const uint8_t*) src, src_idx, n, (int) (row_bytes / 4));
  TRB_CHECK_LAUNCH();
  return 0;
}
