// den_kernel_common.cuh -- pieces shared by the probability-domain (den_shared.cu) and the
// log-domain (den_shared_log.cu) shared-graph kernels.
#pragma once
#include <algorithm>
#include <cstdlib>

#include "common.cuh"
#include "den_blob.h"

namespace chainb200 {

struct DView {
  const int64_t *warp_info;  // [DEN_NW][3] {stream offset, pairs to run, row-code offset}
  const int32_t *rows, *split;
  const uint2 *recs;
  int n_split;
};

struct DenParams {
  DView view[kNumViews];
  const int32_t *cls_dst, *cls_pdf;  // [C_pad] destination state / pdf of every (dst,pdf) class
  int C, C_pad, R_pad;               // R_pad = S_pad + C_pad: floats per utterance of a state row incl. class sums
  const float *leaky, *init, *fin;
  const float *x;
  const int64_t *lengths;
  int B, T, D, S, S_pad, D_pad;
  float coeff;
  unsigned flags;
  float grad_scale;    // posteriors are multiplied by this as they are written
  float *logp, *grad, *diag;
  float *alpha_ws, *tot_ws;
  int num_groups, use_tma_x, max_sub;
  // records resident in shared memory (small graphs): byte offset of every warp's slot stream of every view inside
  // the CTA's record area (which follows the split-row slots), filled by the launcher from the host blob
  uint32_t rs_off[kNumViews][DEN_NW];
  int n_rep;           // copies of the blob laid out back to back on the device
  size_t rep_stride;   // bytes between copies
};

template <int G> struct Vec;
template <> struct Vec<4> { using type = float4; };
template <> struct Vec<2> { using type = float2; };
template <> struct Vec<1> { using type = float; };

template <int G> __device__ __forceinline__ void ld_vec(const float *p, float (&v)[G]);
template <> __device__ __forceinline__ void ld_vec<4>(const float *p, float (&v)[4]) {
  const float4 t = *reinterpret_cast<const float4 *>(p);
  v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
}
template <> __device__ __forceinline__ void ld_vec<2>(const float *p, float (&v)[2]) {
  const float2 t = *reinterpret_cast<const float2 *>(p);
  v[0] = t.x; v[1] = t.y;
}
template <> __device__ __forceinline__ void ld_vec<1>(const float *p, float (&v)[1]) { v[0] = *p; }

template <int G> __device__ __forceinline__ void st_vec(float *p, const float (&v)[G]);
template <> __device__ __forceinline__ void st_vec<4>(float *p, const float (&v)[4]) {
  *reinterpret_cast<float4 *>(p) = make_float4(v[0], v[1], v[2], v[3]);
}
template <> __device__ __forceinline__ void st_vec<2>(float *p, const float (&v)[2]) {
  *reinterpret_cast<float2 *>(p) = make_float2(v[0], v[1]);
}
template <> __device__ __forceinline__ void st_vec<1>(float *p, const float (&v)[1]) { *p = v[0]; }

__device__ __forceinline__ uint32_t smem_u32(const void *p);
// Shared-memory gather of one element (G floats) as a VOLATILE asm: the four operand loads of a record pair are
// issued together, ahead of the end-of-slice branch (the compiler would otherwise sink the second record's loads
// into the branch and expose their latency).
template <int G> __device__ __forceinline__ void lds_vec(uint32_t addr, float (&v)[G]);
template <> __device__ __forceinline__ void lds_vec<4>(uint32_t addr, float (&v)[4]) {
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]) : "r"(addr));
}
template <> __device__ __forceinline__ void lds_vec<2>(uint32_t addr, float (&v)[2]) {
  asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v[0]), "=f"(v[1]) : "r"(addr));
}
template <> __device__ __forceinline__ void lds_vec<1>(uint32_t addr, float (&v)[1]) {
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v[0]) : "r"(addr));
}

template <int G> __device__ __forceinline__ void sts_vec(uint32_t addr, const float (&v)[G]);
template <> __device__ __forceinline__ void sts_vec<4>(uint32_t addr, const float (&v)[4]) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]) : "memory");
}
template <> __device__ __forceinline__ void sts_vec<2>(uint32_t addr, const float (&v)[2]) {
  asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(addr), "f"(v[0]), "f"(v[1]) : "memory");
}
template <> __device__ __forceinline__ void sts_vec<1>(uint32_t addr, const float (&v)[1]) {
  asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v[0]) : "memory");
}

// one lane's PAIR of records (16 bytes); a warp reads 512 contiguous bytes
template <bool L1C>
__device__ __forceinline__ uint4 ldg_stream(const uint4 *p) {
  uint4 r;
  if (L1C) asm volatile("ld.global.nc.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  else asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}

// ---- mbarrier / TMA bulk helpers (cp.async.bulk: SASS UBLKCP) ----
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok) {
    asm volatile(
        "{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  }
}
__device__ __forceinline__ void tma_load_1d(void *smem_dst, const void *gsrc, uint32_t bytes, unsigned long long *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// sum over the block of G per-thread values; result to every thread.
template <int G>
__device__ __forceinline__ void block_sum_vec(float (&v)[G], float *red /* [DEN_NW*G] */) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int g = 0; g < G; g++) {
    v[g] = warp_sum(v[g]);
    if (lane == 0) red[warp * G + g] = v[g];
  }
  __syncthreads();
#pragma unroll
  for (int g = 0; g < G; g++) {
    float t = 0.f;
    for (int w = 0; w < DEN_NW; w++) t += red[w * G + g];
    v[g] = t;
  }
  __syncthreads();
}

struct DenWs {
  size_t alpha, tot, logp, diag, total;
};
inline DenWs den_ws_layout(const BlobHeader &h, int B, int T) {
  DenWs w;
  size_t off = 0;
  const size_t Bp = (size_t)B + 3;
  w.alpha = off; off = align_up(off + sizeof(float) * Bp * (T + 1) * h.R_pad);
  w.tot = off; off = align_up(off + sizeof(float) * Bp * (T + 1));
  w.logp = off; off = align_up(off + sizeof(float) * Bp);
  w.diag = off; off = align_up(off + sizeof(float) * 2 * Bp);
  w.total = off;
  return w;
}

inline size_t den_smem_bytes(const BlobHeader &h, int G) {
  return sizeof(float) * ((size_t)2 * h.R_pad * G + (size_t)3 * h.D_pad * G +
                          (size_t)h.max_sub * G * (h.log_domain ? 2 : 1)) + 128;
}

// Relative cost of one frame for a group of G utterances (shared-memory wavefronts of
// the vector gathers), used to pick G: time ~ waves(G) * cost(G).
inline int choose_G(const BlobHeader &h, int B, int num_sms) {
  const char *env = getenv("CHAIN_B200_G");
  if (env) {
    const int g = atoi(env);
    if ((g == 1 || g == 2 || g == 4) && den_smem_bytes(h, g) <= 227 * 1024) return g;
  }
  // measured on B200 (cfg4): a frame of a G-group costs nearly the same for every G (the gathers are
  // wavefront-bound, not byte-bound), so minimise the number of rounds first, then prefer the smaller G
  const double cost[5] = {0, 0.84, 0.88, 0, 1.0};
  int best = 0;
  double best_t = 1e30;
  for (int g : {4, 2, 1}) {
    if (den_smem_bytes(h, g) > 227 * 1024) continue;
    const int groups = (B + g - 1) / g;
    const double t = (double)((groups + num_sms - 1) / num_sms) * cost[g];
    if (t < best_t - 1e-9) { best_t = t; best = g; }
  }
  return best;
}


// defined in den_shared_log.cu
int launch_den_shared_log(const BlobHeader &h, const DenParams &p, int G, bool l1c, int grid, size_t smem, cudaStream_t stream);

// Bytes of the slot streams proper (without the register-bank run-out) of all views: what a CTA copies into shared
// memory when the graph is small enough to keep its records there.  Fills the per-warp offsets.
inline size_t resident_record_bytes(const void *blob_host, const BlobHeader &h, uint32_t (*off)[DEN_NW]) {
  size_t total = 0;
  for (int v = 0; v < kNumViews; v++) {
    const int64_t *wi = (const int64_t *)((const char *)blob_host + h.view[v].off_warp_info);
    for (int w = 0; w < DEN_NW; w++) {
      if (off) off[v][w] = (uint32_t)total;
      total += (size_t)wi[3 * w + 1] * 512;   // pairs x 32 lanes x 16 bytes
    }
  }
  return total;
}

// Do the record streams of all three views fit in L1 next to the CTA's shared memory?
inline bool records_fit_l1(const BlobHeader &h, size_t smem) {
  // n_recs includes the (never consumed) register-bank run-out of every warp stream
  const int64_t runout = (int64_t)DEN_NW * 3 * kFifoPairs * 64;
  int64_t n = 0;
  for (int i = 0; i < kNumViews; i++) n += h.view[i].n_recs - runout;
  const size_t recs = (size_t)n * 8;
  return recs + smem + 24 * 1024 <= 228 * 1024;
}

}  // namespace chainb200
