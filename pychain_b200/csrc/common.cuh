// common.cuh -- shared device/host helpers of the chain_b200 engine (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>
#include "chain_b200.h"

namespace chainb200 {

constexpr float kClampLo = -30.0f;  // loss.py:30
constexpr float kClampHi = 30.0f;
constexpr size_t kAlign = 256;

inline size_t align_up(size_t x, size_t a = kAlign) { return (x + a - 1) / a * a; }

// Per-thread launch counter (bench.py's gpu_launches).
extern thread_local int g_launch_count;
extern int g_verbose_level;

// objf = sum_b logp[b]; ok = the reference's t=0 self-check (defined in generic.cu)
void launch_finalize(const float *logp, const float *diag, const int64_t *lengths, int B, int T, int check_ab,
                     float *objf, int32_t *ok, float *diag_out, cudaStream_t stream);

struct DeviceGuard {
  int prev = -1;
  bool switched = false;
  cudaError_t err = cudaSuccess;
  explicit DeviceGuard(int device) {
    err = cudaGetDevice(&prev);
    if (err == cudaSuccess && prev != device) {
      err = cudaSetDevice(device);
      switched = (err == cudaSuccess);
    }
  }
  ~DeviceGuard() {
    if (switched) cudaSetDevice(prev);
  }
};

// ---------------------------------------------------------------- device side
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Block-wide sum, result returned to every thread.  Fixed summation order
// (lane tree, then warps 0..nw-1) => deterministic.  `scratch` needs >= 32 floats.
// Contains two __syncthreads(): safe to call back to back with the same scratch.
__device__ __forceinline__ float block_sum(float v, float *scratch) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  if (lane == 0) scratch[warp] = v;
  __syncthreads();
  float t = 0.f;
  for (int w = 0; w < nw; w++) t += scratch[w];
  __syncthreads();
  return t;
}

// log(sum_i exp(v_i)) over the block from per-thread (m, s) pairs with
// s = sum exp(v - m); m = -inf means "no terms".  Result to every thread.
__device__ __forceinline__ void lse_merge(float &m, float &s, float m2, float s2) {
  if (m2 > m) {  // one exponential per merge: the larger side keeps its scale
    s = s * expf(m - m2) + s2;  // m = -inf: exp(-inf) = 0
    m = m2;
  } else if (m2 > -INFINITY) {
    s += s2 * expf(m2 - m);
  }
}
// Running log-sum-exp: fold one term into (m, s) with a single exp.
__device__ __forceinline__ void lse_add(float &m, float &s, float t) {
  if (t > m) {
    s = s * expf(m - t) + 1.f;  // m = -inf: exp(-inf) = 0
    m = t;
  } else if (t > -INFINITY) {
    s += expf(t - m);
  }
}
__device__ __forceinline__ float block_lse(float m, float s, float *scratch /* >= 64 floats */) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float m2 = __shfl_xor_sync(0xffffffffu, m, o), s2 = __shfl_xor_sync(0xffffffffu, s, o);
    lse_merge(m, s, m2, s2);
  }
  if (lane == 0) { scratch[2 * warp] = m; scratch[2 * warp + 1] = s; }
  __syncthreads();
  float M = -INFINITY, Sx = 0.f;
  for (int w = 0; w < nw; w++) lse_merge(M, Sx, scratch[2 * w], scratch[2 * w + 1]);
  __syncthreads();
  return (M == -INFINITY) ? -INFINITY : M + logf(Sx);
}

// torch.clamp semantics (loss.py:30): NaN propagates (fminf/fmaxf would silently turn it into a bound)
__device__ __forceinline__ float clampf(float v, float lo, float hi) { return v < lo ? lo : (v > hi ? hi : v); }

}  // namespace chainb200
