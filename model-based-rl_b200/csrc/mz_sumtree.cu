// Device-resident sum-tree of the prioritized replay (SumTree, replay_buffer.py:6-66).
//
// The tree is the reference's array-embedded binary heap of 2*capacity-1 float64 sums with the
// leaves at [capacity-1, 2*capacity-1), kept in HBM.  The reference applies a batch of priority
// updates one leaf at a time (`tree[ancestor] += change`, replay_buffer.py:35-41), so the value of
// an inner node depends on the ORDER of the float64 additions.  The kernels below keep that order
// while running the batch in parallel: every tree node belongs to exactly one level, so levels are
// independent (one CTA per depth), and inside a level the first update that touches a node applies
// all later changes to that node in batch order.  Result: inner sums -- and therefore which leaf a
// given uniform lands on -- are bit-identical to the reference's.
#include "mz_common.cuh"

namespace {

constexpr int kUpdThreads = 256;
constexpr int kUpdChunk = 2048;  // updates per launch (a power of two: 16 B of shared memory each, 12 key bits)

// depth of a node in the array-embedded heap (root = 0)
MZ_DEV int node_depth(int64_t node) { return 63 - __clzll((unsigned long long)(node + 1)); }
// ancestor of `leaf` at absolute depth `depth` (-1 when the leaf is shallower): idx -> (idx-1)//2
MZ_DEV int64_t ancestor_at(int64_t leaf, int depth) {
  const int up = node_depth(leaf) - depth;
  return up < 0 ? -1 : (((leaf + 1) >> up) - 1);
}

// Both kernels below bring the updates that touch the same node together by sorting 64-bit keys (node << 12 |
// batch index) in shared memory: a node's updates end up adjacent AND in batch order, so one thread can apply them
// with the reference's sequence of float64 additions.  O(n log^2 n) compare-exchanges per CTA instead of the
// O(n^2) scans of the first version (a 500-memory add: 176 -> ~20 us).
MZ_DEV void bitonic_sort_u64(unsigned long long* keys, int N) {  // N a power of two, all threads of the CTA call
  for (int k = 2; k <= N; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int t = threadIdx.x; t < (N >> 1); t += blockDim.x) {
        const int lo = ((t & ~(j - 1)) << 1) | (t & (j - 1)), hi = lo | j;  // pair (lo, lo + j)
        const unsigned long long a = keys[lo], b = keys[hi];
        const bool up = (lo & k) == 0;
        if ((a > b) == up) {
          keys[lo] = b;
          keys[hi] = a;
        }
      }
      __syncthreads();
    }
  }
}
constexpr unsigned long long kKeySkip = ~0ull;

// change[i] = priority[i] - (value of the leaf just before update i)  (replay_buffer.py:36): the priority of the
// previous update of the same leaf in this batch, else the tree's
__global__ void sumtree_change_kernel(const double* __restrict__ tree, int n, int N,
                                      const int64_t* __restrict__ idx,
                                      const double* __restrict__ pri, double* __restrict__ change) {
  __shared__ unsigned long long s_key[kUpdChunk];
  for (int i = threadIdx.x; i < N; i += blockDim.x)
    s_key[i] = i < n ? ((unsigned long long)idx[i] << 12) | (unsigned)i : kKeySkip;
  __syncthreads();
  bitonic_sort_u64(s_key, N);
  for (int p = threadIdx.x; p < n; p += blockDim.x) {
    const unsigned long long key = s_key[p];
    const int i = (int)(key & 4095);
    const int64_t leaf = (int64_t)(key >> 12);
    const bool again = p > 0 && (int64_t)(s_key[p - 1] >> 12) == leaf;
    const double before = again ? pri[s_key[p - 1] & 4095] : tree[leaf];
    change[i] = __dsub_rn(pri[i], before);
  }
}

// blockIdx.x = depth of the nodes this CTA owns (with a capacity that is not a power of two the
// leaves sit on two depths, so ownership goes by absolute depth, not by height above the leaf)
__global__ void sumtree_apply_kernel(double* __restrict__ tree, int n, int N,
                                     const int64_t* __restrict__ idx, const double* __restrict__ pri,
                                     const double* __restrict__ change) {
  __shared__ unsigned long long s_key[kUpdChunk];
  __shared__ double s_change[kUpdChunk];
  const int depth = blockIdx.x;
  for (int i = threadIdx.x; i < N; i += blockDim.x) {
    const int64_t node = i < n ? ancestor_at(idx[i], depth) : -1;  // -1: this leaf sits above `depth`
    s_key[i] = node >= 0 ? ((unsigned long long)node << 12) | (unsigned)i : kKeySkip;
    if (i < n) s_change[i] = change[i];
  }
  __syncthreads();
  bitonic_sort_u64(s_key, N);
  for (int p = threadIdx.x; p < n; p += blockDim.x) {
    const unsigned long long key = s_key[p];
    if (key == kKeySkip) continue;
    const int64_t node = (int64_t)(key >> 12);
    if (p > 0 && (int64_t)(s_key[p - 1] >> 12) == node) continue;  // an earlier update owns this node
    int q = p;
    if (node == idx[key & 4095]) {  // the leaf itself: the last write wins
      while (q + 1 < n && (int64_t)(s_key[q + 1] >> 12) == node) ++q;
      tree[node] = pri[s_key[q] & 4095];
      continue;
    }
    double v = tree[node];
    for (; q < n && (int64_t)(s_key[q] >> 12) == node; ++q) v = __dadd_rn(v, s_change[s_key[q] & 4095]);
    tree[node] = v;
  }
}

// slot -> (window position, chunk start, chunk length): SumTree.buffer[position] = (step, history)
__global__ void sumtree_slots_kernel(int n, const int64_t* __restrict__ idx, int64_t leaf0,
                                     int64_t chunk_start, int32_t chunk_len,
                                     int64_t* __restrict__ slot_pos, int64_t* __restrict__ slot_start,
                                     int32_t* __restrict__ slot_len, int first_step) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int64_t slot = idx[i] - leaf0;
  slot_pos[slot] = chunk_start + first_step + i;
  slot_start[slot] = chunk_start;
  slot_len[slot] = chunk_len;
}

// the same for the memories of several chunks: memory i is step i - seg_begin[s] of segment s
__global__ void sumtree_slots_seg_kernel(int n, const int64_t* __restrict__ idx, int64_t leaf0, int nseg,
                                         const int32_t* __restrict__ seg_begin, const int64_t* __restrict__ seg_start,
                                         const int32_t* __restrict__ seg_len, int64_t* __restrict__ slot_pos,
                                         int64_t* __restrict__ slot_start, int32_t* __restrict__ slot_len) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int lo = 0, hi = nseg - 1;  // last segment with seg_begin <= i
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (seg_begin[mid] <= i) lo = mid; else hi = mid - 1;
  }
  const int64_t slot = idx[i] - leaf0;
  slot_pos[slot] = seg_start[lo] + (i - seg_begin[lo]);
  slot_start[slot] = seg_start[lo];
  slot_len[slot] = seg_len[lo];
}

// SumTree.get_leaf (replay_buffer.py:43-62) for the stratified values of sample_batch
// (replay_buffer.py:137-141) + the importance weights (replay_buffer.py:160-162).  One CTA.
__global__ void sumtree_sample_kernel(const double* __restrict__ tree, int64_t size, int64_t leaf0,
                                      int n, const double* __restrict__ u01,
                                      const int64_t* __restrict__ slot_pos,
                                      const int64_t* __restrict__ slot_start,
                                      const int32_t* __restrict__ slot_len, double num_memories,
                                      double beta, int u01_is_mt_words, int64_t* __restrict__ out_idx,
                                      double* __restrict__ out_pri, int64_t* __restrict__ out_pos,
                                      int64_t* __restrict__ out_start, int32_t* __restrict__ out_len,
                                      double* __restrict__ is_weights) {
  __shared__ double s_max[32];
  const double total = tree[0];
  const double segment = __ddiv_rn(total, (double)n);
  double wmax = 0.0;
  for (int b = threadIdx.x; b < n; b += blockDim.x) {
    // random.uniform(s1, s2) = s1 + (s2 - s1) * random()
    const double s1 = __dmul_rn(segment, (double)b), s2 = __dmul_rn(segment, (double)(b + 1));
    double u = u01[b];
    if (u01_is_mt_words) {  // CPython's random(): (a >> 5) * 2**26 + (b >> 6)) / 2**53 from two MT19937 outputs
      const uint2 w = reinterpret_cast<const uint2*>(u01)[b];
      u = __dmul_rn(__dadd_rn(__dmul_rn((double)(w.x >> 5), 67108864.0), (double)(w.y >> 6)), 1.0 / 9007199254740992.0);
    }
    double value = __dadd_rn(s1, __dmul_rn(__dsub_rn(s2, s1), u));
    int64_t parent = 0;
    for (;;) {
      const int64_t left = 2 * parent + 1;
      if (left >= size) break;
      const double lv = tree[left];
      if (value <= lv) {
        parent = left;
      } else {
        value = __dsub_rn(value, lv);
        parent = left + 1;
      }
    }
    const double p = tree[parent];
    out_idx[b] = parent;
    out_pri[b] = p;
    if (slot_pos) {
      const int64_t slot = parent - leaf0;
      out_pos[b] = slot_pos[slot];
      out_start[b] = slot_start[slot];
      out_len[b] = slot_len[slot];
    }
    if (is_weights) {
      const double w = pow(__dmul_rn(num_memories, __ddiv_rn(p, total)), -beta);
      is_weights[b] = w;
      wmax = fmax(wmax, w);
    }
  }
  if (!is_weights) return;
  for (int o = 16; o; o >>= 1) wmax = fmax(wmax, shfl_xor_f64<32>(wmax, o));
  if ((threadIdx.x & 31) == 0) s_max[threadIdx.x >> 5] = wmax;
  __syncthreads();
  wmax = 0.0;
  for (int w = 0; w < (int)(blockDim.x >> 5); ++w) wmax = fmax(wmax, s_max[w]);
  for (int b = threadIdx.x; b < n; b += blockDim.x) is_weights[b] = __ddiv_rn(is_weights[b], wmax);
}

// (|error| + epsilon) ** alpha like numpy on a float32 array with python-float epsilon / alpha
__global__ void priorities_kernel(int n, const float* __restrict__ errors, float epsilon, float alpha,
                                  double* __restrict__ priority) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float p = __fadd_rn(fabsf(errors[i]), epsilon);
  if (alpha != 1.0f) p = (float)pow((double)p, (double)alpha);
  priority[i] = (double)p;
}

int tree_levels(int64_t max_capacity) {
  int levels = 1;  // leaves
  for (int64_t deepest = 2 * max_capacity - 1; deepest > 1; deepest >>= 1) ++levels;
  return levels;
}

}  // namespace

extern "C" {

int mz_sumtree_update(double* tree, int64_t max_capacity, int64_t n, const int64_t* tree_idx,
                      const double* priority, double* scratch, void* stream) {
  if (!tree || max_capacity < 1 || n < 0 || (n > 0 && (!tree_idx || !priority || !scratch)))
    return MZ_ERR_BAD_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  const int levels = tree_levels(max_capacity);
  for (int64_t o = 0; o < n; o += kUpdChunk) {
    const int m = (int)((n - o) < kUpdChunk ? (n - o) : kUpdChunk);
    int N = 2;
    while (N < m) N <<= 1;  // the sort runs over the next power of two
    sumtree_change_kernel<<<1, kUpdThreads, 0, st>>>(tree, m, N, tree_idx + o, priority + o, scratch + o);
    sumtree_apply_kernel<<<levels, kUpdThreads, 0, st>>>(tree, m, N, tree_idx + o, priority + o,
                                                          scratch + o);
  }
  MZ_LAUNCH_CHECK();
  return MZ_OK;
}

int mz_sumtree_add_from(double* tree, int64_t max_capacity, int64_t n, const int64_t* tree_idx,
                        const double* priority, int64_t chunk_start, int32_t chunk_len, int32_t first_step,
                        int64_t* slot_pos, int64_t* slot_start, int32_t* slot_len, double* scratch, void* stream) {
  if (n > 0 && (!slot_pos || !slot_start || !slot_len)) return MZ_ERR_BAD_ARG;
  if (first_step < 0 || n + first_step > chunk_len) return MZ_ERR_BAD_ARG;
  const int rc = mz_sumtree_update(tree, max_capacity, n, tree_idx, priority, scratch, stream);
  if (rc != MZ_OK || n == 0) return rc;
  sumtree_slots_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      (int)n, tree_idx, max_capacity - 1, chunk_start, chunk_len, slot_pos, slot_start, slot_len, first_step);
  MZ_LAUNCH_CHECK();
  return MZ_OK;
}

int mz_sumtree_update_errors(double* tree, int64_t max_capacity, int64_t n, const int64_t* tree_idx,
                             const float* errors, double epsilon, double alpha, double* priority, double* scratch,
                             void* stream) {
  if (n < 0 || (n > 0 && (!errors || !priority))) return MZ_ERR_BAD_ARG;
  if (n == 0) return MZ_OK;
  priorities_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>((int)n, errors, (float)epsilon,
                                                                                   (float)alpha, priority);
  MZ_LAUNCH_CHECK();
  return mz_sumtree_update(tree, max_capacity, n, tree_idx, priority, scratch, stream);
}

int mz_sumtree_add_chunks(double* tree, int64_t max_capacity, int64_t n, const int64_t* tree_idx,
                          const double* priority, int32_t num_segments, const int32_t* seg_begin,
                          const int64_t* seg_start, const int32_t* seg_len, int64_t* slot_pos, int64_t* slot_start,
                          int32_t* slot_len, double* scratch, void* stream) {
  if (n > 0 && (!slot_pos || !slot_start || !slot_len || num_segments < 1 || !seg_begin || !seg_start || !seg_len))
    return MZ_ERR_BAD_ARG;
  const int rc = mz_sumtree_update(tree, max_capacity, n, tree_idx, priority, scratch, stream);
  if (rc != MZ_OK || n == 0) return rc;
  sumtree_slots_seg_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      (int)n, tree_idx, max_capacity - 1, num_segments, seg_begin, seg_start, seg_len, slot_pos, slot_start, slot_len);
  MZ_LAUNCH_CHECK();
  return MZ_OK;
}

int mz_sumtree_add(double* tree, int64_t max_capacity, int64_t n, const int64_t* tree_idx,
                   const double* priority, int64_t chunk_start, int32_t chunk_len, int64_t* slot_pos,
                   int64_t* slot_start, int32_t* slot_len, double* scratch, void* stream) {
  return mz_sumtree_add_from(tree, max_capacity, n, tree_idx, priority, chunk_start, chunk_len, 0, slot_pos,
                             slot_start, slot_len, scratch, stream);
}

int mz_sumtree_sample_mt(const double* tree, int64_t max_capacity, int32_t n, const double* u01, int32_t u01_is_mt_words,
                         const int64_t* slot_pos, const int64_t* slot_start, const int32_t* slot_len,
                         int64_t num_memories, double beta, int64_t* tree_idx, double* priority,
                         int64_t* pos, int64_t* chunk_start, int32_t* chunk_len, double* is_weights,
                         void* stream);

int mz_sumtree_sample(const double* tree, int64_t max_capacity, int32_t n, const double* u01,
                      const int64_t* slot_pos, const int64_t* slot_start, const int32_t* slot_len,
                      int64_t num_memories, double beta, int64_t* tree_idx, double* priority,
                      int64_t* pos, int64_t* chunk_start, int32_t* chunk_len, double* is_weights,
                      void* stream) {
  return mz_sumtree_sample_mt(tree, max_capacity, n, u01, 0, slot_pos, slot_start, slot_len, num_memories, beta, tree_idx,
                              priority, pos, chunk_start, chunk_len, is_weights, stream);
}

int mz_sumtree_sample_mt(const double* tree, int64_t max_capacity, int32_t n, const double* u01, int32_t u01_is_mt_words,
                         const int64_t* slot_pos, const int64_t* slot_start, const int32_t* slot_len,
                         int64_t num_memories, double beta, int64_t* tree_idx, double* priority,
                         int64_t* pos, int64_t* chunk_start, int32_t* chunk_len, double* is_weights,
                         void* stream) {
  if (!tree || max_capacity < 1 || n < 1 || !u01 || !tree_idx || !priority) return MZ_ERR_BAD_ARG;
  if (slot_pos && (!slot_start || !slot_len || !pos || !chunk_start || !chunk_len))
    return MZ_ERR_BAD_ARG;
  const int threads = n >= 1024 ? 1024 : ((n + 31) / 32) * 32;
  sumtree_sample_kernel<<<1, threads, 0, (cudaStream_t)stream>>>(
      tree, 2 * max_capacity - 1, max_capacity - 1, n, u01, slot_pos, slot_start, slot_len,
      (double)num_memories, beta, u01_is_mt_words, tree_idx, priority, pos, chunk_start, chunk_len, is_weights);
  MZ_LAUNCH_CHECK();
  return MZ_OK;
}

}  // extern "C"
