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
constexpr int kUpdChunk = 2048;  // updates per launch (16 B of shared memory each)

// depth of a node in the array-embedded heap (root = 0)
MZ_DEV int node_depth(int64_t node) { return 63 - __clzll((unsigned long long)(node + 1)); }
// ancestor of `leaf` at absolute depth `depth` (-1 when the leaf is shallower): idx -> (idx-1)//2
MZ_DEV int64_t ancestor_at(int64_t leaf, int depth) {
  const int up = node_depth(leaf) - depth;
  return up < 0 ? -1 : (((leaf + 1) >> up) - 1);
}

// change[i] = priority[i] - (value of the leaf just before update i)  (replay_buffer.py:36)
__global__ void sumtree_change_kernel(const double* __restrict__ tree, int n,
                                      const int64_t* __restrict__ idx,
                                      const double* __restrict__ pri, double* __restrict__ change) {
  __shared__ int64_t s_idx[kUpdChunk];
  for (int i = threadIdx.x; i < n; i += blockDim.x) s_idx[i] = idx[i];
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const int64_t leaf = s_idx[i];
    int j = i - 1;
    while (j >= 0 && s_idx[j] != leaf) --j;
    const double before = j >= 0 ? pri[j] : tree[leaf];
    change[i] = __dsub_rn(pri[i], before);
  }
}

// blockIdx.x = depth of the nodes this CTA owns (with a capacity that is not a power of two the
// leaves sit on two depths, so ownership goes by absolute depth, not by height above the leaf)
__global__ void sumtree_apply_kernel(double* __restrict__ tree, int n,
                                     const int64_t* __restrict__ idx, const double* __restrict__ pri,
                                     const double* __restrict__ change) {
  __shared__ int64_t s_node[kUpdChunk];
  __shared__ double s_change[kUpdChunk];
  const int depth = blockIdx.x;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    s_node[i] = ancestor_at(idx[i], depth);
    s_change[i] = change[i];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const int64_t node = s_node[i];
    if (node < 0) continue;  // this leaf sits above `depth`
    if (node == idx[i]) {    // the leaf itself
      int j = i + 1;
      while (j < n && s_node[j] != node) ++j;
      if (j == n) tree[node] = pri[i];  // the last write to a leaf wins
      continue;
    }
    int j = i - 1;
    while (j >= 0 && s_node[j] != node) --j;
    if (j >= 0) continue;  // an earlier update owns this node
    double v = tree[node];
    for (j = i; j < n; ++j)
      if (s_node[j] == node) v = __dadd_rn(v, s_change[j]);
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
                                      double beta, int64_t* __restrict__ out_idx,
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
    double value = __dadd_rn(s1, __dmul_rn(__dsub_rn(s2, s1), u01[b]));
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
    sumtree_change_kernel<<<1, kUpdThreads, 0, st>>>(tree, m, tree_idx + o, priority + o, scratch + o);
    sumtree_apply_kernel<<<levels, kUpdThreads, 0, st>>>(tree, m, tree_idx + o, priority + o,
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

int mz_sumtree_sample(const double* tree, int64_t max_capacity, int32_t n, const double* u01,
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
      (double)num_memories, beta, tree_idx, priority, pos, chunk_start, chunk_len, is_weights);
  MZ_LAUNCH_CHECK();
  return MZ_OK;
}

}  // extern "C"
