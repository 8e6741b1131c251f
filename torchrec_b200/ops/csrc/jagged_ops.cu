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

// ---------------------------------------------------------------------------------------------------------------
// jagged [sum L, D] <-> padded dense [N, max_len, D]   (fbgemm jagged_to_padded_dense / dense_to_jagged, reference
// sparse/jagged_tensor.py:1001, 1059). One thread per 4-byte word of an output row slot; D_words = row bytes / 4.
// ---------------------------------------------------------------------------------------------------------------
// V = uint32_t (any 4-byte multiple row) or uint4 (rows of 16-byte multiples: 4x fewer threads, 16 B accesses); IDX = uint32_t when the
// output has < 2^32 vectors (one 32-bit division pair per thread instead of 64-bit ones).
template <typename V, typename IDX>
__global__ void __launch_bounds__(256) jagged_to_padded_kernel(const V* __restrict__ values, const int64_t* __restrict__ offsets, V* __restrict__ out,
                                                                int64_t N, int max_len, int D_vec, V pad) {
  const IDX i = (IDX) blockIdx.x * blockDim.x + threadIdx.x;
  const IDX total = (IDX) N * (IDX) max_len * (IDX) D_vec;
  if (i >= total) return;
  const IDX slot = i / (IDX) D_vec;
  const int w = (int) (i - slot * (IDX) D_vec);
  const IDX seg = slot / (IDX) max_len;
  const int pos = (int) (slot - seg * (IDX) max_len);
  const int64_t s = __ldg(offsets + seg), e = __ldg(offsets + seg + 1);
  out[i] = (s + pos < e) ? __ldg(values + (s + pos) * D_vec + w) : pad;
}

TRB_API int trb_jagged_to_padded_dense(const void* values, const int64_t* offsets, void* out, int64_t N, int max_len, int row_bytes, uint32_t pad_word,
                                       cudaStream_t stream) {
  if (row_bytes % 4) return -2;
  const bool v16 = row_bytes % 16 == 0 && ((uintptr_t) values % 16 == 0) && ((uintptr_t) out % 16 == 0);
  const int D_vec = row_bytes / (v16 ? 16 : 4);
  const int64_t total = N * max_len * D_vec;
  if (total == 0) return 0;
  const unsigned blocks = (unsigned) ((total + 255) / 256);
  const bool small = total < 0xffffffffLL;
  if (v16) {
    const uint4 pad = make_uint4(pad_word, pad_word, pad_word, pad_word);
    if (small) jagged_to_padded_kernel<uint4, uint32_t><<<blocks, 256, 0, stream>>>((const uint4*) values, offsets, (uint4*) out, N, max_len, D_vec, pad);
    else jagged_to_padded_kernel<uint4, int64_t><<<blocks, 256, 0, stream>>>((const uint4*) values, offsets, (uint4*) out, N, max_len, D_vec, pad);
  } else {
    if (small) jagged_to_padded_kernel<uint32_t, uint32_t><<<blocks, 256, 0, stream>>>((const uint32_t*) values, offsets, (uint32_t*) out, N, max_len, D_vec, pad_word);
    else jagged_to_padded_kernel<uint32_t, int64_t><<<blocks, 256, 0, stream>>>((const uint32_t*) values, offsets, (uint32_t*) out, N, max_len, D_vec, pad_word);
  }
  TRB_CHECK_LAUNCH();
  return 0;
}

template <typename V, typename IDX>
__global__ void __launch_bounds__(256) padded_to_jagged_kernel(const V* __restrict__ dense, const int64_t* __restrict__ offsets, V* __restrict__ out,
                                                                int64_t N, int max_len, int D_vec) {
  const IDX i = (IDX) blockIdx.x * blockDim.x + threadIdx.x;
  const IDX total = (IDX) N * (IDX) max_len * (IDX) D_vec;
  if (i >= total) return;
  const IDX slot = i / (IDX) D_vec;
  const int w = (int) (i - slot * (IDX) D_vec);
  const IDX seg = slot / (IDX) max_len;
  const int pos = (int) (slot - seg * (IDX) max_len);
  const int64_t s = __ldg(offsets + seg), e = __ldg(offsets + seg + 1);
  if (s + pos < e) out[(s + pos) * D_vec + w] = __ldg(dense + i);
}

TRB_API int trb_dense_to_jagged(const void* dense, const int64_t* offsets, void* out, int64_t N, int max_len, int row_bytes, cudaStream_t stream) {
  if (row_bytes % 4) return -2;
  const bool v16 = row_bytes % 16 == 0 && ((uintptr_t) dense % 16 == 0) && ((uintptr_t) out % 16 == 0);
  const int D_vec = row_bytes / (v16 ? 16 : 4);
  const int64_t total = N * max_len * D_vec;
  if (total == 0) return 0;
  const unsigned blocks = (unsigned) ((total + 255) / 256);
  const bool small = total < 0xffffffffLL;
  if (v16) {
    if (small) padded_to_jagged_kernel<uint4, uint32_t><<<blocks, 256, 0, stream>>>((const uint4*) dense, offsets, (uint4*) out, N, max_len, D_vec);
    else padded_to_jagged_kernel<uint4, int64_t><<<blocks, 256, 0, stream>>>((const uint4*) dense, offsets, (uint4*) out, N, max_len, D_vec);
  } else {
    if (small) padded_to_jagged_kernel<uint32_t, uint32_t><<<blocks, 256, 0, stream>>>((const uint32_t*) dense, offsets, (uint32_t*) out, N, max_len, D_vec);
    else padded_to_jagged_kernel<uint32_t, int64_t><<<blocks, 256, 0, stream>>>((const uint32_t*) dense, offsets, (uint32_t*) out, N, max_len, D_vec);
  }
  TRB_CHECK_LAUNCH();
  return 0;
}

// ---------------------------------------------------------------------------------------------------------------
// segment_sum_csr: out[s] = sum(values[csr_seg[s] * batch .. csr_seg[s+1] * batch))   (fbgemm segment_sum_csr; KJT length_per_key
// on device, mean-pooling divisors: reference sparse/jagged_tensor.py:1342, distributed/embeddingbag.py:2560). One warp per segment.
// ---------------------------------------------------------------------------------------------------------------
template <typename T, typename S>
__global__ void __launch_bounds__(256) segment_sum_csr_kernel(const T* __restrict__ values, const S* __restrict__ csr, T* __restrict__ out, int64_t n_seg, int batch) {
  const int lane = threadIdx.x & 31;
  const int64_t seg = ((int64_t) blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (seg >= n_seg) return;
  const int64_t s = (int64_t) csr[seg] * batch, e = (int64_t) csr[seg + 1] * batch;
  double acc = 0.0;
  for (int64_t i = s + lane; i < e; i += 32) acc += (double) values[i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) out[seg] = (T) acc;
}

TRB_API int trb_segment_sum_csr(const void* values, int v_dtype, const void* csr, int csr64, void* out, int64_t n_seg, int batch, cudaStream_t stream) {
  if (n_seg == 0) return 0;
  const unsigned blocks = (unsigned) ((n_seg * 32 + 255) / 256);
#define TRB_SS(VT, ST) segment_sum_csr_kernel<VT, ST><<<blocks, 256, 0, stream>>>((const VT*) values, (const ST*) csr, (VT*) out, n_seg, batch)
  if (v_dtype == TRB_F32) { if (csr64) TRB_SS(float, int64_t); else TRB_SS(float, int32_t); }
  else if (v_dtype == TRB_I64) { if (csr64) TRB_SS(int64_t, int64_t); else TRB_SS(int64_t, int32_t); }
  else if (v_dtype == TRB_I32) { if (csr64) TRB_SS(int32_t, int64_t); else TRB_SS(int32_t, int32_t); }
  else return -3;
#undef TRB_SS
  TRB_CHECK_LAUNCH();
  return 0;
}

// ---------------------------------------------------------------------------------------------------------------
// permute_pooled_embs: column-block permutation of [B, sum D] (restores the feature order after a column-wise output dist; its own
// backward with the inverse permutation). Block k of the OUTPUT starts at out_off[k], has width dims[k] and comes from in_off[k].
// Reference: fbgemm permute_pooled_embs_auto_grad, called at distributed/embeddingbag.py:1665, sharding/cw_sharding.py:305.
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) permute_pooled_kernel(const uint32_t* __restrict__ in, uint32_t* __restrict__ out, const int32_t* __restrict__ col_src, int64_t B,
                                                              int cols_words) {
  const int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * cols_words) return;
  const int c = (int) (i % cols_words);
  const int64_t b = i / cols_words;
  out[i] = in[b * cols_words + col_src[c]];
}

TRB_API int trb_permute_pooled_embs(const void* in, void* out, const int32_t* col_src_words, int64_t B, int cols_words, cudaStream_t stream) {
  const int64_t n = B * cols_words;
  if (n == 0) return 0;
  permute_pooled_kernel<<<(unsigned) ((n + 255) / 256), 256, 0, stream>>>((const uint32_t*) in, (uint32_t*) out, col_src_words, B, cols_words);
  TRB_CHECK_LAUNCH();
  return 0;
}

// ---------------------------------------------------------------------------------------------------------------
// UVM cache: id -> cache slot for a whole batch in one launch (ops/uvm.py: translate). Feature f owns positions
// [bounds[f], bounds[f+1]) of `indices`; its table's direct map slot_of_row[table][row] holds the HBM cache slot of a resident row.
// Out-of-range ids and rows that are not resident translate to -1 (the lookup kernels skip negative ids).
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) cache_translate_kernel(const void* __restrict__ indices, int idx64, const int64_t* __restrict__ bounds, int F,
                                                              const int64_t* __restrict__ map_ptrs, const int64_t* __restrict__ rows_of_feature,
                                                              void* __restrict__ out, int64_t n) {
  const int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int lo = 0, hi = F - 1;  // last feature whose start <= i
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (bounds[mid] <= i) lo = mid; else hi = mid - 1;
  }
  const int64_t id = trb_ld_idx(indices, i, idx64);
  int64_t slot = -1;
  if (id >= 0 && id < rows_of_feature[lo]) slot = (int64_t) reinterpret_cast<const int32_t*>(map_ptrs[lo])[id];
  trb_st_idx(out, i, idx64, slot);
}

TRB_API int trb_cache_translate(const void* indices, int idx64, const int64_t* bounds, int F, const int64_t* map_ptrs, const int64_t* rows_of_feature, void* out,
                                int64_t n, cudaStream_t stream) {
  if (n == 0 || F == 0) return 0;
  cache_translate_kernel<<<(unsigned) ((n + 255) / 256), 256, 0, stream>>>(indices, idx64, bounds, F, map_ptrs, rows_of_feature, out, n);
  TRB_CHECK_LAUNCH();
  return 0;
}
