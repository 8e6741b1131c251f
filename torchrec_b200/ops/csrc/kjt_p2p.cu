// Sparse-feature input dist over NVLink peer memory (hot path 1).
//
// Reference path (dist_data.py:422-996): permute -> splits all-to-all -> D2H .tolist() (host blocks on the
// GPU) -> one NCCL all-to-all per tensor -> recat permute -> cumsum. Here every rank *publishes* its routed
// KJT (offsets + values [+ weights]) in a symmetric buffer; after one device-side barrier each destination
// *pulls* the segments of its lookup units straight from the peers and writes them in the final
// [unit][source rank][sample] order while computing the output offsets — sizes never visit the host.
//
//   kernel 1 (1 CTA) : segment sizes n(u, s) = off_s[(u+1)B] - off_s[uB]  -> exclusive scan -> seg_base
//   kernel 2         : per (u, s): out_off[(u,s,b)] = seg_base + (off_s[uB+b] - off_s[uB]); copy ids (peer LDG)
#include "common.cuh"

struct KjtPullParams {
  TrbPeerPtrs off;      // per source rank: int64 offsets [U_total * B + 1]
  TrbPeerPtrs val;      // per source rank: ids
  TrbPeerPtrs wgt;      // per source rank: per-id weights (float) or nullptr
  const int32_t* units; // [U_d] global unit index of the destination's local units
  int64_t* seg_base;    // [U_d * W + 1]
  int64_t* out_off;     // [U_d * W * B + 1]
  void* out_val;
  float* out_wgt;
  int32_t* overflow;    // set to 1 when the pulled ids exceed `capacity`
  int64_t capacity;
  int32_t U_d, W, B, val_bytes;
};

__global__ void __launch_bounds__(512) kjt_pull_scan_kernel(const KjtPullParams p) {
  extern __shared__ int64_t s_n[];
  const int n_seg = p.U_d * p.W;
  for (int i = threadIdx.x; i < n_seg; i += blockDim.x) {
    const int ul = i / p.W, s = i - ul * p.W;
    const int64_t* off = reinterpret_cast<const int64_t*>(p.off.p[s]);
    const int64_t u = p.units[ul];
    s_n[i] = off[(u + 1) * p.B] - off[u * p.B];
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int64_t acc = 0;
    for (int i = 0; i < n_seg; ++i) {
      p.seg_base[i] = acc;
      acc += s_n[i];
    }
    p.seg_base[n_seg] = acc;
    p.out_off[(int64_t) n_seg * p.B] = acc;
    if (acc > p.capacity) *p.overflow = 1;
  }
}

template <typename T>
__global__ void __launch_bounds__(256) kjt_pull_copy_kernel(const KjtPullParams p) {
  const int seg = blockIdx.y;
  const int ul = seg / p.W, s = seg - ul * p.W;
  const int64_t* off = reinterpret_cast<const int64_t*>(p.off.p[s]);
  const int64_t u = p.units[ul];
  const int64_t src0 = off[u * p.B];
  const int64_t n = off[(u + 1) * p.B] - src0;
  const int64_t dst0 = p.seg_base[seg];
  const int64_t tid = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t) gridDim.x * blockDim.x;
  for (int64_t b = tid; b < p.B; b += stride) p.out_off[(int64_t) seg * p.B + b] = dst0 + (off[u * p.B + b] - src0);
  if (dst0 + n > p.capacity) return;  // overflow is reported by the scan kernel
  const T* sv = reinterpret_cast<const T*>(p.val.p[s]) + src0;
  T* dv = reinterpret_cast<T*>(p.out_val) + dst0;
  for (int64_t i = tid; i < n; i += stride) dv[i] = sv[i];
  if (p.out_wgt != nullptr) {
    const float* sw = reinterpret_cast<const float*>(p.wgt.p[s]) + src0;
    for (int64_t i = tid; i < n; i += stride) p.out_wgt[dst0 + i] = sw[i];
  }
}

TRB_API int trb_kjt_pull(void* const* off_ptrs, void* const* val_ptrs, void* const* wgt_ptrs, int W, const int32_t* units, int U_d, int B,
                         int val_bytes, int64_t* seg_base, int64_t* out_off, void* out_val, float* out_wgt, int64_t capacity, int32_t* overflow,
                         int64_t expected_per_segment, cudaStream_t stream) {
  if (W < 1 || W > TRB_MAX_PEERS) return -1;
  if (U_d == 0) return 0;
  KjtPullParams p;
  for (int i = 0; i < TRB_MAX_PEERS; ++i) {
    p.off.p[i] = i < W ? off_ptrs[i] : nullptr;
    p.val.p[i] = i < W ? val_ptrs[i] : nullptr;
    p.wgt.p[i] = (i < W && wgt_ptrs != nullptr) ? wgt_ptrs[i] : nullptr;
  }
  p.units = units; p.seg_base = seg_base; p.out_off = out_off; p.out_val = out_val; p.out_wgt = out_wgt; p.overflow = overflow;
  p.capacity = capacity; p.U_d = U_d; p.W = W; p.B = B; p.val_bytes = val_bytes;
  const int n_seg = U_d * W;
  kjt_pull_scan_kernel<<<1, 512, n_seg * sizeof(int64_t), stream>>>(p);
  TRB_CHECK_LAUNCH();
  int64_t work = expected_per_segment > B ? expected_per_segment : B;
  int bx = (int) ((work + 256 * 4 - 1) / (256 * 4));
  if (bx < 1) bx = 1;
  if (bx > 64) bx = 64;
  dim3 grid(bx, n_seg);
  if (val_bytes == 8) kjt_pull_copy_kernel<int64_t><<<grid, 256, 0, stream>>>(p);
  else if (val_bytes == 4) kjt_pull_copy_kernel<int32_t><<<grid, 256, 0, stream>>>(p);
  else return -31;
  TRB_CHECK_LAUNCH();
  return 0;
}
