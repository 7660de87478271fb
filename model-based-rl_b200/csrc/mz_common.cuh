// Shared device helpers for libmzb200 (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "mzb200.h"

#define MZ_DEV __device__ __forceinline__
#define MZ_FULL 0xffffffffu

#define MZ_LAUNCH_CHECK()                         \
  do {                                            \
    cudaError_t e__ = cudaGetLastError();         \
    if (e__ != cudaSuccess) return (int)e__;      \
  } while (0)

// ---- node record accessors (layout documented in include/mzb200.h) ---------------------------
struct MzNode {
  uint8_t* p;
  int A;
  MZ_DEV double& vsum() const { return *reinterpret_cast<double*>(p); }
  MZ_DEV int32_t& visit() const { return *reinterpret_cast<int32_t*>(p + 8); }
  MZ_DEV float& reward() const { return *reinterpret_cast<float*>(p + 12); }
  MZ_DEV double* prior() const { return reinterpret_cast<double*>(p + 16); }
  MZ_DEV int16_t* child() const { return reinterpret_cast<int16_t*>(p + 16 + 8 * A); }
};

struct MzGame {
  uint8_t* base;
  int node_bytes;
  int A;
  MZ_DEV double& mn() const { return *reinterpret_cast<double*>(base); }
  MZ_DEV double& mx() const { return *reinterpret_cast<double*>(base + 8); }
  MZ_DEV int32_t& root_to_play() const { return *reinterpret_cast<int32_t*>(base + 16); }
  MZ_DEV MzNode node(int n) const {
    return MzNode{base + MZ_GAME_HEADER_BYTES + (size_t)n * node_bytes, A};
  }
};

// ---- 64-bit shuffles ---------------------------------------------------------------------------
template <int W>
MZ_DEV double shfl_f64(double v, int src) {
  int lo = __double2loint(v), hi = __double2hiint(v);
  lo = __shfl_sync(MZ_FULL, lo, src, W);
  hi = __shfl_sync(MZ_FULL, hi, src, W);
  return __hiloint2double(hi, lo);
}
template <int W>
MZ_DEV double shfl_xor_f64(double v, int m) {
  int lo = __double2loint(v), hi = __double2hiint(v);
  lo = __shfl_xor_sync(MZ_FULL, lo, m, W);
  hi = __shfl_xor_sync(MZ_FULL, hi, m, W);
  return __hiloint2double(hi, lo);
}
