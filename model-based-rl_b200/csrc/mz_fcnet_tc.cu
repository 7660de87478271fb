// FCNetwork.recurrent_inference on the 5th-generation tensor cores (tcgen05 + TMEM), bf16 operands,
// fp32 accumulation.  One CTA evaluates the whole network for 128 rows (games):
//
//   dynamics   : [h | onehot(a)] (K1) --W1--> 2 x 512 --relu--> W2 --> reward logits (32) , h' (64)
//   LayerNorm + ReLU on h'  (fp32, one thread per row), reward = softmax-expectation + h^-1
//   prediction : h' (K3 = 64) --W3--> 2 x 512 --relu--> W4 --> value logits (32), policy logits (32)
//
// The 512-wide hidden layers are processed in 16 chunks of 128 features.  Per chunk c:
//   MMA1(c): D1[c&1] (TMEM, 128 lanes x 128 cols) = A(128 x K) * W1c^T          (tcgen05.mma)
//   EPI(c) : D1[c&1] -> registers (tcgen05.ld) -> +bias, relu, bf16 -> A2[c&1] in shared memory
//   MMA2(c): D2 (TMEM) += A2[c&1](128 x 128) * W2c^T
// Weight chunks are pre-packed in global memory as the exact shared-memory image the MMA reads
// (UMMA canonical K-major layout, no swizzle) and streamed through a 3-stage ring with one
// cp.async.bulk per chunk; the pipeline is driven by mbarriers (TMA -> MMA -> epilogue -> MMA).
//
// Warp roles (192 threads): warp 0 = TMEM allocation + bulk-copy producer, warp 1 = MMA issuer
// (one thread), warps 2..5 = epilogue (TMEM lane quarter = warp % 4).
//
// Reference semantics: networks.py:31-34, 122-174 (FCNetwork), config.py:27-33.
#include <cuda_bf16.h>
#include <math.h>

#include "mz_common.cuh"
#include "mz_transforms.cuh"

namespace {

constexpr int H = MZ_FC_HIDDEN;      // 50
constexpr int W = MZ_FC_WIDTH;       // 512
constexpr int ROWS = 128;            // rows (games) per CTA = UMMA M
constexpr int CHUNK = 128;           // hidden features per chunk = UMMA N of the first layer
constexpr int NCHUNK = 16;           // 4 heads x 4 chunks
constexpr int K3 = 64;               // padded K of the prediction first layer
constexpr int STAGES = 3;
constexpr int TC_THREADS = 192;
constexpr int N_REW = 32, N_HID = 64, N_VAL = 32, N_POL = 32;  // padded second-layer widths
constexpr int TMEM_COLS = 512;
constexpr int COL_D1 = 0;            // 2 x 128 columns
constexpr int COL_D2A = 256;         // reward / value logits (32)
constexpr int COL_D2B = 288;         // next hidden (64) / policy logits (32)

// tail parameter block (float): second-layer biases and LayerNorm affine
constexpr int T_REW_B = 0, T_DYN_B = 32, T_LN_W = 96, T_LN_B = 160, T_VAL_B = 224, T_POL_B = 256;
constexpr int TAIL_FLOATS = 288;

struct ChunkGeom {  // byte geometry of one packed chunk: [W1 | W2 | bias1]
  int k;            // K of the first layer (K1 or K3)
  int n2;           // N of the second layer
  int w1_bytes, w2_bytes, bytes;
};

__host__ __device__ inline ChunkGeom chunk_geom(int c, int k1) {
  ChunkGeom g;
  const int head = c >> 2;  // 0 reward, 1 transition, 2 value, 3 policy
  g.k = head < 2 ? k1 : K3;
  g.n2 = head == 0 ? N_REW : (head == 1 ? N_HID : (head == 2 ? N_VAL : N_POL));
  g.w1_bytes = CHUNK * g.k * 2;
  g.w2_bytes = g.n2 * CHUNK * 2;
  g.bytes = g.w1_bytes + g.w2_bytes + CHUNK * 4;
  return g;
}
__host__ __device__ inline size_t chunk_offset(int c, int k1) {
  size_t off = 0;
  for (int i = 0; i < c; ++i) off += chunk_geom(i, k1).bytes;
  return off;
}
__host__ __device__ inline int stage_bytes_for(int k1) { return chunk_geom(4, k1).bytes; }  // largest

// canonical K-major, no swizzle: 8 x 8 core matrices of 128 contiguous bytes,
// core (row_group, k_block) at ((k_block * row_groups) + row_group) * 128
__host__ __device__ inline int canon_off(int row, int k, int rows) {
  return (((k >> 3) * (rows >> 3)) + (row >> 3)) * 128 + (row & 7) * 16 + (k & 7) * 2;
}

// ---- tcgen05 wrappers --------------------------------------------------------------------------
MZ_DEV void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                   smem_u32(smem_dst)),
               "r"(ncols)
               : "memory");
}
MZ_DEV void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
MZ_DEV void tmem_dealloc(uint32_t addr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(ncols)
               : "memory");
}
MZ_DEV void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
MZ_DEV void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
MZ_DEV void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                   smem_u32(bar))
               : "memory");
}
MZ_DEV void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
MZ_DEV void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// D[tmem] (+)= A[smem desc] * B[smem desc]^T, bf16 x bf16 -> f32, issued by one thread
MZ_DEV void umma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                      uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// 32 lanes x 32 consecutive 32-bit columns -> 32 registers per thread (thread i <-> lane base + i)
MZ_DEV void tmem_ld32(uint32_t addr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
        "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),
        "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
        "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
        "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(addr)
      : "memory");
}

// shared-memory matrix descriptor: K-major, SWIZZLE_NONE, version 1 (sm_100)
MZ_DEV uint64_t make_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = (uint64_t)((smem_addr >> 4) & 0x3FFFu);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
// instruction descriptor: D = f32, A = B = bf16, both K-major, M = 128, N = n
MZ_DEV uint32_t make_idesc(int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(ROWS >> 4) << 24);
}

MZ_DEV uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 p = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&p);
}

struct TcParams {
  const uint8_t* chunks;  // packed weights
  const float* tail;
  int k1, num_actions, value_bins, reward_bins, value_min, reward_min, no_tt;
  int batch;
  const float* hidden_in;
  long long in_row_stride;
  const int32_t* in_index;
  const int32_t* actions;
  float* hidden_out;
  long long out_row_stride, out_offset;
  float *value, *reward, *logits;
};

// softmax(logits + bias) . support, then h^-1, all in registers (one thread per row)
MZ_DEV float support_to_scalar_regs(const uint32_t (&v)[32], const float* bias, int bins, int mn,
                                    int no_tt) {
  float x[32];
  float m = -INFINITY;
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    x[j] = __uint_as_float(v[j]) + bias[j];
    if (j < bins) m = fmaxf(m, x[j]);
  }
  float den = 0.0f;
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    x[j] = j < bins ? __expf(x[j] - m) : 0.0f;
    den += x[j];
  }
  float num = 0.0f;
#pragma unroll
  for (int j = 0; j < 32; ++j) num += (float)(mn + j) * __fdiv_rn(x[j], den);
  return no_tt ? num : mz_inverse_scalar_transform_f(num);
}

__global__ void __launch_bounds__(TC_THREADS, 1) fc_recurrent_tc_kernel(TcParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int k1 = p.k1;
  const int stage_bytes = stage_bytes_for(k1);
  // carve shared memory
  uint8_t* sA1 = smem;                                  // [128 x k1] bf16
  uint8_t* sA2 = sA1 + ROWS * k1 * 2;                   // 2 x [128 x 128] bf16
  uint8_t* sA3 = sA2 + 2 * ROWS * CHUNK * 2;            // [128 x 64] bf16
  uint8_t* sW = sA3 + ROWS * K3 * 2;                    // STAGES x stage_bytes
  uint64_t* bars = reinterpret_cast<uint64_t*>(sW + STAGES * stage_bytes);
  uint64_t* w_full = bars;            // [3]
  uint64_t* w_empty = bars + 3;       // [3]
  uint64_t* d1_full = bars + 6;       // [2]
  uint64_t* d1_empty = bars + 8;      // [2]
  uint64_t* a2_full = bars + 10;      // [2]
  uint64_t* a2_empty = bars + 12;     // [2]
  uint64_t* a1_ready = bars + 14;
  uint64_t* a3_ready = bars + 15;
  uint64_t* d2_full = bars + 16;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 17);

  if (threadIdx.x == 0) {
    for (int i = 0; i < 3; ++i) { mbar_init(&w_full[i], 1); mbar_init(&w_empty[i], 1); }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&d1_full[i], 1);
      mbar_init(&d1_empty[i], 128);
      mbar_init(&a2_full[i], 128);
      mbar_init(&a2_empty[i], 1);
    }
    mbar_init(a1_ready, 128);
    mbar_init(a3_ready, 128);
    mbar_init(d2_full, 1);
    mbar_fence_init();
  }
  if (warp == 0) {
    tmem_alloc(tmem_ptr, TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_ptr;

  if (warp == 0) {
    // ===== producer: stream the 16 weight chunks through the ring =====
    if (lane == 0) {
      size_t off = 0;
      for (int c = 0; c < NCHUNK; ++c) {
        const int st = c % STAGES, n = c / STAGES;
        const ChunkGeom g = chunk_geom(c, k1);
        mbar_wait(&w_empty[st], (n & 1) ^ 1);
        mbar_arrive_expect_tx(&w_full[st], (uint32_t)g.bytes);
        bulk_copy_g2s(sW + st * stage_bytes, p.chunks + off, (uint32_t)g.bytes, &w_full[st]);
        off += g.bytes;
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (single thread) =====
    if (lane == 0) {
      const uint32_t a1_addr = smem_u32(sA1), a2_addr = smem_u32(sA2), a3_addr = smem_u32(sA3);
      const uint32_t w_addr = smem_u32(sW);
      const uint32_t idesc1 = make_idesc(CHUNK);
      auto mma2 = [&](int c) {  // D2 += A2[c&1] * W2c^T
        const int st = c % STAGES, head = c >> 2;
        const ChunkGeom g = chunk_geom(c, k1);
        mbar_wait(&a2_full[c & 1], (c >> 1) & 1);
        tc_fence_after();
        const uint32_t d2 = tmem + ((head & 1) ? COL_D2B : COL_D2A);
        const uint32_t idesc2 = make_idesc(g.n2);
        const uint32_t a_base = a2_addr + (c & 1) * (ROWS * CHUNK * 2);
        const uint32_t b_base = w_addr + st * stage_bytes + g.w1_bytes;
        const uint32_t lbo_b = (g.n2 >> 3) * 128;
#pragma unroll 1
        for (int ks = 0; ks < CHUNK / 16; ++ks) {
          const uint64_t ad = make_desc(a_base + ks * 2 * (ROWS / 8) * 128, (ROWS / 8) * 128, 128);
          const uint64_t bd = make_desc(b_base + ks * 2 * lbo_b, lbo_b, 128);
          umma_bf16(d2, ad, bd, idesc2, ((c & 3) != 0 || ks != 0) ? 1u : 0u);
        }
        tc_commit(&w_empty[st]);       // chunk c's weights are no longer needed
        tc_commit(&a2_empty[c & 1]);   // A2 buffer may be overwritten
        if (c == 7 || c == 15) tc_commit(d2_full);
      };
#pragma unroll 1
      for (int c = 0; c < NCHUNK; ++c) {
        const int st = c % STAGES;
        const ChunkGeom g = chunk_geom(c, k1);
        if (c == 8) mma2(7);  // the prediction's A operand depends on the dynamics output
        mbar_wait(&w_full[st], (c / STAGES) & 1);
        if (c == 0) mbar_wait(a1_ready, 0);
        if (c == 8) mbar_wait(a3_ready, 0);
        mbar_wait(&d1_empty[c & 1], ((c >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d1 = tmem + COL_D1 + (c & 1) * CHUNK;
        const uint32_t a_base = c < 8 ? a1_addr : a3_addr;
        const uint32_t b_base = w_addr + st * stage_bytes;
#pragma unroll 1
        for (int ks = 0; ks < g.k / 16; ++ks) {
          const uint64_t ad = make_desc(a_base + ks * 2 * (ROWS / 8) * 128, (ROWS / 8) * 128, 128);
          const uint64_t bd = make_desc(b_base + ks * 2 * (CHUNK / 8) * 128, (CHUNK / 8) * 128, 128);
          umma_bf16(d1, ad, bd, idesc1, ks != 0 ? 1u : 0u);
        }
        tc_commit(&d1_full[c & 1]);
        if (c > 0 && c != 8) mma2(c - 1);
      }
      mma2(NCHUNK - 1);
    }
  } else {
    // ===== epilogue warps: one thread per row =====
    const int quarter = warp & 3;                 // TMEM lane quarter this warp may access
    const int row = quarter * 32 + lane;
    const int g = blockIdx.x * ROWS + row;
    const bool live = g < p.batch;
    const int gc = live ? g : p.batch - 1;
    const uint32_t lane_addr = tmem + ((uint32_t)(quarter * 32) << 16);
    const int a_off = (row >> 3) * 128 + (row & 7) * 16;  // + k_block * (ROWS/8) * 128

    // --- A1 = bf16([h | onehot(action) | 0]) ---
    {
      const float* src = p.hidden_in + (size_t)gc * p.in_row_stride +
                         (p.in_index ? (size_t)p.in_index[gc] * H : 0);
      const int act = p.actions[gc];
      for (int kb = 0; kb < k1 / 8; ++kb) {
        float f[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int k = kb * 8 + j;
          f[j] = k < H ? src[k] : ((k - H) == act ? 1.0f : 0.0f);
        }
        uint4 q = make_uint4(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]), pack_bf16(f[4], f[5]),
                             pack_bf16(f[6], f[7]));
        *reinterpret_cast<uint4*>(sA1 + kb * (ROWS / 8) * 128 + a_off) = q;
      }
      fence_async_smem();
      mbar_arrive(a1_ready);
    }

    uint32_t v[32];
    auto hidden_epilogue = [&](int c) {  // D1[c&1] -> relu(x + b1) -> bf16 -> A2[c&1]
      const int st = c % STAGES;
      const ChunkGeom gm = chunk_geom(c, k1);
      const float* bias = reinterpret_cast<const float*>(sW + st * stage_bytes + gm.w1_bytes + gm.w2_bytes);
      mbar_wait(&w_full[st], (c / STAGES) & 1);  // bias lives in the chunk image (async-proxy write)
      mbar_wait(&d1_full[c & 1], (c >> 1) & 1);
      tc_fence_after();
      mbar_wait(&a2_empty[c & 1], ((c >> 1) & 1) ^ 1);
      uint8_t* dst = sA2 + (c & 1) * (ROWS * CHUNK * 2) + a_off;
#pragma unroll 1
      for (int qd = 0; qd < CHUNK / 32; ++qd) {
        tmem_ld32(lane_addr + COL_D1 + (c & 1) * CHUNK + qd * 32, v);
        tmem_wait_ld();
#pragma unroll
        for (int kb = 0; kb < 4; ++kb) {
          float f[8];
#pragma unroll
          for (int j = 0; j < 8; ++j)
            f[j] = fmaxf(__uint_as_float(v[kb * 8 + j]) + bias[qd * 32 + kb * 8 + j], 0.0f);
          uint4 q = make_uint4(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]), pack_bf16(f[4], f[5]),
                               pack_bf16(f[6], f[7]));
          *reinterpret_cast<uint4*>(dst + (qd * 4 + kb) * (ROWS / 8) * 128) = q;
        }
      }
      fence_async_smem();
      mbar_arrive(&a2_full[c & 1]);
      tc_fence_before();
      mbar_arrive(&d1_empty[c & 1]);
    };

    for (int c = 0; c < 8; ++c) hidden_epilogue(c);

    // --- dynamics outputs: reward scalar, h' = relu(LN(.)) ---
    mbar_wait(d2_full, 0);
    tc_fence_after();
    tmem_ld32(lane_addr + COL_D2A, v);
    tmem_wait_ld();
    const float rew = support_to_scalar_regs(v, p.tail + T_REW_B, p.reward_bins, p.reward_min, p.no_tt);
    if (live) p.reward[g] = rew;
    {
      float hbuf[64];
      tmem_ld32(lane_addr + COL_D2B, v);
      tmem_wait_ld();
#pragma unroll
      for (int j = 0; j < 32; ++j) hbuf[j] = __uint_as_float(v[j]) + p.tail[T_DYN_B + j];
      tmem_ld32(lane_addr + COL_D2B + 32, v);
      tmem_wait_ld();
#pragma unroll
      for (int j = 0; j < 32; ++j) hbuf[32 + j] = __uint_as_float(v[j]) + p.tail[T_DYN_B + 32 + j];
      float s = 0.0f;
#pragma unroll
      for (int j = 0; j < H; ++j) s += hbuf[j];
      const float mean = s / (float)H;
      float qv = 0.0f;
#pragma unroll
      for (int j = 0; j < H; ++j) {
        const float d = hbuf[j] - mean;
        qv = fmaf(d, d, qv);
      }
      const float rstd = 1.0f / sqrtf(qv / (float)H + 1e-5f);
#pragma unroll
      for (int j = 0; j < 64; ++j)
        hbuf[j] = j < H ? fmaxf((hbuf[j] - mean) * rstd * p.tail[T_LN_W + j] + p.tail[T_LN_B + j], 0.0f)
                        : 0.0f;
      if (live) {
        float* dsth = p.hidden_out + (size_t)g * p.out_row_stride + p.out_offset;
#pragma unroll
        for (int j = 0; j < H; ++j) dsth[j] = hbuf[j];
      }
#pragma unroll
      for (int kb = 0; kb < K3 / 8; ++kb) {
        uint4 q = make_uint4(pack_bf16(hbuf[kb * 8 + 0], hbuf[kb * 8 + 1]),
                             pack_bf16(hbuf[kb * 8 + 2], hbuf[kb * 8 + 3]),
                             pack_bf16(hbuf[kb * 8 + 4], hbuf[kb * 8 + 5]),
                             pack_bf16(hbuf[kb * 8 + 6], hbuf[kb * 8 + 7]));
        *reinterpret_cast<uint4*>(sA3 + kb * (ROWS / 8) * 128 + a_off) = q;
      }
      fence_async_smem();
      tc_fence_before();
      mbar_arrive(a3_ready);
    }

    for (int c = 8; c < NCHUNK; ++c) hidden_epilogue(c);

    // --- prediction outputs: value scalar, policy logits ---
    mbar_wait(d2_full, 1);
    tc_fence_after();
    tmem_ld32(lane_addr + COL_D2A, v);
    tmem_wait_ld();
    const float val = support_to_scalar_regs(v, p.tail + T_VAL_B, p.value_bins, p.value_min, p.no_tt);
    if (live) p.value[g] = val;
    tmem_ld32(lane_addr + COL_D2B, v);
    tmem_wait_ld();
    if (live) {
      float* dl = p.logits + (size_t)g * p.num_actions;
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (j < p.num_actions) dl[j] = __uint_as_float(v[j]) + p.tail[T_POL_B + j];
    }
    tc_fence_before();
  }

  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, TMEM_COLS);
  }
}

// ---- packing: f32 reference-layout weights -> bf16 chunk images + tail parameters ---------------
__global__ void fc_tc_pack_kernel(mz_fc_weights w, int k1, uint8_t* chunks, float* tail) {
  const int A = w.num_actions;
  for (int c = 0; c < NCHUNK; ++c) {
    const ChunkGeom g = chunk_geom(c, k1);
    uint8_t* base = chunks + chunk_offset(c, k1);
    const int head = c >> 2, f0 = (c & 3) * CHUNK;
    const float* w1t = head == 0 ? w.rew_w1 : (head == 1 ? w.dyn_w1 : (head == 2 ? w.val_w1 : w.pol_w1));
    const float* b1 = head == 0 ? w.rew_b1 : (head == 1 ? w.dyn_b1 : (head == 2 ? w.val_b1 : w.pol_b1));
    const float* w2 = head == 0 ? w.rew_w2 : (head == 1 ? w.dyn_w2 : (head == 2 ? w.val_w2 : w.pol_w2));
    const int kin = head < 2 ? H + A : H;          // real input width of the first layer
    const int nout = head == 0 ? w.reward_bins : (head == 1 ? H : (head == 2 ? w.value_bins : A));
    // W1 chunk: B operand [CHUNK features x K], element (n, k) = w1t[k][f0 + n]
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < CHUNK * g.k; i += gridDim.x * blockDim.x) {
      const int n = i / g.k, k = i % g.k;
      const float x = k < kin ? w1t[(size_t)k * W + f0 + n] : 0.0f;
      *reinterpret_cast<__nv_bfloat16*>(base + canon_off(n, k, CHUNK)) = __float2bfloat16_rn(x);
    }
    // W2 chunk: B operand [n2 outputs x CHUNK], element (o, k) = w2[o][f0 + k]
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < g.n2 * CHUNK; i += gridDim.x * blockDim.x) {
      const int o = i / CHUNK, k = i % CHUNK;
      const float x = o < nout ? w2[(size_t)o * W + f0 + k] : 0.0f;
      *reinterpret_cast<__nv_bfloat16*>(base + g.w1_bytes + canon_off(o, k, g.n2)) = __float2bfloat16_rn(x);
    }
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < CHUNK; i += gridDim.x * blockDim.x)
      reinterpret_cast<float*>(base + g.w1_bytes + g.w2_bytes)[i] = b1[f0 + i];
  }
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < TAIL_FLOATS; i += gridDim.x * blockDim.x) {
    float x = 0.0f;
    if (i < T_DYN_B) x = i < w.reward_bins ? w.rew_b2[i] : 0.0f;
    else if (i < T_LN_W) x = (i - T_DYN_B) < H ? w.dyn_b2[i - T_DYN_B] : 0.0f;
    else if (i < T_LN_B) x = (i - T_LN_W) < H ? w.ln_w[i - T_LN_W] : 0.0f;
    else if (i < T_VAL_B) x = (i - T_LN_B) < H ? w.ln_b[i - T_LN_B] : 0.0f;
    else if (i < T_POL_B) x = (i - T_VAL_B) < w.value_bins ? w.val_b2[i - T_VAL_B] : 0.0f;
    else x = (i - T_POL_B) < A ? w.pol_b2[i - T_POL_B] : 0.0f;
    tail[i] = x;
  }
}

int k1_for(int A) { return (H + A + 15) / 16 * 16; }

size_t tc_smem_bytes(int k1) {
  return (size_t)ROWS * k1 * 2 + 2 * ROWS * CHUNK * 2 + ROWS * K3 * 2 + (size_t)STAGES * stage_bytes_for(k1) +
         17 * sizeof(uint64_t) + 16;
}

}  // namespace

extern "C" {

int64_t mz_fc_tc_packed_bytes(int32_t num_actions) {
  if (num_actions < 1 || num_actions > 32) return MZ_ERR_UNSUPPORTED;
  return (int64_t)chunk_offset(NCHUNK, k1_for(num_actions));
}

int32_t mz_fc_tc_tail_floats(void) { return TAIL_FLOATS; }

int mz_fc_tc_pack(const mz_fc_weights* w, void* packed, float* tail, void* stream) {
  if (!w || !packed || !tail) return MZ_ERR_BAD_ARG;
  if (w->num_actions < 1 || w->num_actions > 32 || w->value_bins < 1 || w->value_bins > 32 ||
      w->reward_bins < 1 || w->reward_bins > 32)
    return MZ_ERR_UNSUPPORTED;
  fc_tc_pack_kernel<<<64, 256, 0, (cudaStream_t)stream>>>(*w, k1_for(w->num_actions), (uint8_t*)packed, tail);
  MZ_LAUNCH_CHECK();
  return MZ_OK;
}

int mz_fc_recurrent_tc(const mz_fc_weights* w, const void* packed, const float* tail, int32_t batch,
                       const float* hidden_in, int64_t in_row_stride, const int32_t* in_index,
                       const int32_t* actions, float* hidden_out, int64_t out_row_stride,
                       int64_t out_offset, float* value, float* reward, float* logits, void* stream) {
  if (!w || !packed || !tail || batch < 1 || !hidden_in || !actions || !hidden_out || !value ||
      !reward || !logits)
    return MZ_ERR_BAD_ARG;
  if (w->num_actions < 1 || w->num_actions > 32 || w->value_bins > 32 || w->reward_bins > 32)
    return MZ_ERR_UNSUPPORTED;
  TcParams p;
  p.chunks = (const uint8_t*)packed;
  p.tail = tail;
  p.k1 = k1_for(w->num_actions);
  p.num_actions = w->num_actions;
  p.value_bins = w->value_bins;
  p.reward_bins = w->reward_bins;
  p.value_min = w->value_min;
  p.reward_min = w->reward_min;
  p.no_tt = w->no_target_transform;
  p.batch = batch;
  p.hidden_in = hidden_in;
  p.in_row_stride = in_row_stride;
  p.in_index = in_index;
  p.actions = actions;
  p.hidden_out = hidden_out;
  p.out_row_stride = out_row_stride;
  p.out_offset = out_offset;
  p.value = value;
  p.reward = reward;
  p.logits = logits;
  const size_t smem = tc_smem_bytes(p.k1);
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(fc_recurrent_tc_kernel,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
    if (e != cudaSuccess) return (int)e;
    attr = true;
  }
  if (smem > 232448) return MZ_ERR_UNSUPPORTED;
  const int grid = (batch + ROWS - 1) / ROWS;
  fc_recurrent_tc_kernel<<<grid, TC_THREADS, smem, (cudaStream_t)stream>>>(p);
  MZ_LAUNCH_CHECK();
  return MZ_OK;
}

}  // extern "C"
