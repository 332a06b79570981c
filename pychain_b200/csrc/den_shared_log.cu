// den_shared_log.cu -- LOG-domain forward-backward for ONE graph shared by the batch, sm_100a.
//
// Same machinery as den_shared.cu (persistent CTAs, G interleaved utterances, three SELL-32
// views streamed from L2 through two register banks, TMA staging of x frames and alpha rows,
// atomic-free pdf-keyed gradient) in the log semiring of
// /root/reference/pytorch_binding/src/chain-log-domain-computation.cc:84-271: products become
// sums, sums become a running log-sum-exp (one exp per record and utterance), there is no
// leaky-HMM mixing, alpha(t) is renormalised by sc(t-1) = logsumexp_i alpha(t-1,i), and the
// gradient is gamma(t,n) = exp(logsumexp over arcs with pdf n of alpha + w + beta + x - sc)
// (or its log with CHAIN_B200_GRAD_AS_LOG).  This is the path BASELINE.json's configs[1]
// ("denominator ... log-domain") takes when the graph is shared by the batch.
#include "den_kernel_common.cuh"

namespace chainb200 {

// Branch-free running log-sum-exp with the fast exponential (ex2.approx): the arguments are
// <= 0 and the terms of one row are within a few dozen nats of their maximum, where __expf's
// absolute error is ~1e-7 -- far inside the 2e-5 gradient / 1e-4 objf tolerances.
__device__ __forceinline__ void lse_add_fast(float &m, float &s, float t) {
  const float M = fmaxf(m, t);
  if (M > -INFINITY) {
    s = s * __expf(m - M) + __expf(t - M);
    m = M;
  }
}

template <int G>
__device__ __forceinline__ void lse_rec(const uint32_t ab, const float w, const float *__restrict__ A,
                                        const float *__restrict__ Bv, float (&m)[G], float (&s)[G]) {
  float a[G], b[G];
  ld_vec<G>(A + (ab & 0xffffu) * G, a);
  ld_vec<G>(Bv + (ab >> 16) * G, b);
#pragma unroll
  for (int g = 0; g < G; g++) lse_add_fast(m[g], s[g], a[g] + w + b[g]);
}

// Two terms folded into the running (max, sum-exp) at once: three exponentials instead of four and ONE step of the
// lane's serial dependency chain per record pair instead of two (t1 = -inf: "no second term").
__device__ __forceinline__ void lse_add2_fast(float &m, float &s, float t0, float t1) {
  const float M = fmaxf(m, fmaxf(t0, t1));
  if (M > -INFINITY) {
    s = s * __expf(m - M) + __expf(t0 - M) + __expf(t1 - M);
    m = M;
  }
}

// One SELL pass in the log semiring: every lane folds A[a_k][g] + w_k + Bv[b_k][g] of its
// row's records into a running (max, sum-exp) per utterance.  Record streams and the
// two-bank register prefetch are those of den_shared.cu; the end-of-slice flag is bit 31 of
// the index word.
// EMBED: stream format (den_blob.h).  Embedded: the slice ends on a row-code slot (weight word = kCodeMarkLog).  Compact:
// bit 31 of the second slot's index word ends the slice, row codes come from the parallel stream v.rows.
template <int G, bool L1C, bool EMBED, class Epi>
__device__ __forceinline__ void sell_pass_log(const DView &v, const float *__restrict__ A,
                                              const float *__restrict__ Bv, float *__restrict__ partial, int warp,
                                              int lane, Epi &epi) {
  const uint4 *sp = reinterpret_cast<const uint4 *>(v.recs + v.warp_info[3 * warp]) + lane;  // [pair][lane]: 2 records
  const int npairs = (int)v.warp_info[3 * warp + 1];
  const int32_t *rp = EMBED ? nullptr : v.rows + v.warp_info[3 * warp + 2] + lane;
  int row_c = 0, row_n1 = 0, row_n2 = 0;
  if constexpr (!EMBED) {
    row_c = __ldg(rp); row_n1 = __ldg(rp + 32); row_n2 = __ldg(rp + 64);
    rp += 96;
  }
  uint4 fa[kFifoPairs], fb[kFifoPairs];
#pragma unroll
  for (int d = 0; d < kFifoPairs; d++) fa[d] = ldg_stream<L1C>(sp + d * 32);
  sp += kFifoPairs * 32;
  float m[G], s[G];
#pragma unroll
  for (int g = 0; g < G; g++) { m[g] = -INFINITY; s[g] = 0.f; }
  // a pair is two arcs, or -- at the end of a slice -- an arc and the lane's row-code slot (weight word = kCodeMarkLog)
  auto process = [&](const uint2 r0, const uint2 r1) {
    const bool last = EMBED ? (r1.y == kCodeMarkLog) : ((r1.x & 0x80000000u) != 0);  // warp-uniform: end of the slice
    {
      float a0[G], b0[G], t1[G];
      ld_vec<G>(A + (r0.x & 0xffffu) * G, a0);
      ld_vec<G>(Bv + (r0.x >> 16) * G, b0);
#pragma unroll
      for (int g = 0; g < G; g++) t1[g] = -INFINITY;
      if (!EMBED || !last) {   // the second slot is an arc (compact format: always; embedded: unless it is the row code)
        float a1[G], b1[G];
        const uint32_t x1 = EMBED ? r1.x : (r1.x & 0x7fffffffu);
        ld_vec<G>(A + (x1 & 0xffffu) * G, a1);
        ld_vec<G>(Bv + (x1 >> 16) * G, b1);
#pragma unroll
        for (int g = 0; g < G; g++) t1[g] = a1[g] + __uint_as_float(r1.y) + b1[g];
      }
#pragma unroll
      for (int g = 0; g < G; g++) lse_add2_fast(m[g], s[g], a0[g] + __uint_as_float(r0.y) + b0[g], t1[g]);
    }
    if (last) {
      const int row = EMBED ? (int)r1.x : row_c;
      if (row >= 0) {
        float o[G];
#pragma unroll
        for (int g = 0; g < G; g++) o[g] = (m[g] > -INFINITY) ? m[g] + logf(s[g]) : -INFINITY;
        epi(row, o);
      } else if (row <= -2) {  // piece of a split row: hand the (max, sum-exp) pairs to the owner
        st_vec<G>(partial + (size_t)(-2 - row) * 2 * G, m);
        st_vec<G>(partial + (size_t)(-2 - row) * 2 * G + G, s);
      }
#pragma unroll
      for (int g = 0; g < G; g++) { m[g] = -INFINITY; s[g] = 0.f; }
      if constexpr (!EMBED) {
        row_c = row_n1;
        row_n1 = row_n2;
        row_n2 = __ldg(rp);
        rp += 32;
      }
    }
  };
  int j = 0;
  for (; j + 2 * kFifoPairs <= npairs; j += 2 * kFifoPairs) {
#pragma unroll
    for (int d = 0; d < kFifoPairs; d++) fb[d] = ldg_stream<L1C>(sp + d * 32);
    sp += kFifoPairs * 32;
#pragma unroll
    for (int d = 0; d < kFifoPairs; d++) process(make_uint2(fa[d].x, fa[d].y), make_uint2(fa[d].z, fa[d].w));
#pragma unroll
    for (int d = 0; d < kFifoPairs; d++) fa[d] = ldg_stream<L1C>(sp + d * 32);
    sp += kFifoPairs * 32;
#pragma unroll
    for (int d = 0; d < kFifoPairs; d++) process(make_uint2(fb[d].x, fb[d].y), make_uint2(fb[d].z, fb[d].w));
  }
  const int rem = npairs - j;
  if (rem > 0) {
    if (rem > kFifoPairs) {
#pragma unroll
      for (int d = 0; d < kFifoPairs; d++) fb[d] = ldg_stream<L1C>(sp + d * 32);
    }
#pragma unroll
    for (int d = 0; d < kFifoPairs; d++)
      if (d < rem) process(make_uint2(fa[d].x, fa[d].y), make_uint2(fa[d].z, fa[d].w));
    if (rem > kFifoPairs) {
#pragma unroll
      for (int d = 0; d < kFifoPairs; d++)
        if (kFifoPairs + d < rem) process(make_uint2(fb[d].x, fb[d].y), make_uint2(fb[d].z, fb[d].w));
    }
  }
}

// Rows that were split: merge the sub-rows' (max, sum-exp) pairs in fixed order and run the epilogue.
template <int G, class Epi>
__device__ __forceinline__ void fixup_split_log(const DView &v, const float *__restrict__ partial, Epi &epi) {
  for (int i = threadIdx.x; i < v.n_split; i += DEN_NT) {
    const int row = v.split[3 * i], first = v.split[3 * i + 1], cnt = v.split[3 * i + 2];
    float m[G], s[G];
#pragma unroll
    for (int g = 0; g < G; g++) { m[g] = -INFINITY; s[g] = 0.f; }
    for (int j = 0; j < cnt; j++) {
      float m2[G], s2[G];
      ld_vec<G>(partial + (size_t)(first + j) * 2 * G, m2);
      ld_vec<G>(partial + (size_t)(first + j) * 2 * G + G, s2);
#pragma unroll
      for (int g = 0; g < G; g++) lse_merge(m[g], s[g], m2[g], s2[g]);
    }
    float o[G];
#pragma unroll
    for (int g = 0; g < G; g++) o[g] = (m[g] > -INFINITY) ? m[g] + logf(s[g]) : -INFINITY;
    epi(row, o);
  }
}

// log-sum-exp over the block of G per-thread (m, s) pairs; result to every thread.
template <int G>
__device__ __forceinline__ void block_lse_vec(float (&m)[G], float (&s)[G], float (&out)[G], float *red /* [DEN_NW*2*G] */) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int g = 0; g < G; g++) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float m2 = __shfl_xor_sync(0xffffffffu, m[g], o), s2 = __shfl_xor_sync(0xffffffffu, s[g], o);
      lse_merge(m[g], s[g], m2, s2);
    }
    if (lane == 0) { red[(warp * G + g) * 2] = m[g]; red[(warp * G + g) * 2 + 1] = s[g]; }
  }
  __syncthreads();
#pragma unroll
  for (int g = 0; g < G; g++) {
    float M = -INFINITY, Sx = 0.f;
    for (int w = 0; w < DEN_NW; w++) lse_merge(M, Sx, red[(w * G + g) * 2], red[(w * G + g) * 2 + 1]);
    out[g] = (M == -INFINITY) ? -INFINITY : M + logf(Sx);
  }
  __syncthreads();
}

constexpr int kLogWarpIssueX = DEN_NW - 1, kLogWarpFetchA = DEN_NW > 1 ? DEN_NW - 2 : 0, kLogWarpTot = DEN_NW > 2 ? DEN_NW - 3 : 0;

template <int G, bool L1C, bool EMBED>
__global__ void __launch_bounds__(DEN_NT, 1) den_shared_log_kernel(const DenParams p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ __align__(8) unsigned long long mbar[2];  // 0: alpha rows, 1: x frames
  __shared__ float red[DEN_NW * 4 * 2];

  const int S = p.S, D = p.D, T = p.T, S_pad = p.S_pad, D_pad = p.D_pad;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  float *buf0 = reinterpret_cast<float *>(smem_raw);
  float *buf1 = buf0 + (size_t)S_pad * G;
  float *x_s = buf1 + (size_t)S_pad * G;
  float *xraw = x_s + (size_t)D_pad * G;
  float *gam_s = xraw + (size_t)D_pad * G;
  float *partial = gam_s + (size_t)D_pad * G;  // [max_sub][2][G]
  const float *__restrict__ ini = p.init;
  const float *__restrict__ fin = p.fin;
  const bool clamp_x = (p.flags & CHAIN_B200_CLAMP_X) != 0;
  const bool as_log = (p.flags & CHAIN_B200_GRAD_AS_LOG) != 0;

  if (tid == 0) {
    mbar_init(&mbar[0], 1);
    mbar_init(&mbar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  uint32_t ph_a = 0, ph_x = 0;
  __syncthreads();

  for (int gi = blockIdx.x; gi < p.num_groups; gi += gridDim.x) {
    const int b0 = gi * G;
    int Lg[G];
    int Lmax = 0;
#pragma unroll
    for (int g = 0; g < G; g++) {
      long long l = (b0 + g < p.B) ? p.lengths[b0 + g] : 0;
      Lg[g] = (int)(l < 0 ? 0 : (l > T ? T : l));
      Lmax = max(Lmax, Lg[g]);
    }
    float *aws = p.alpha_ws + (size_t)gi * (T + 1) * S_pad * G;
    float *tws = p.tot_ws + (size_t)gi * (T + 1) * G;
    const float *xg[G];
#pragma unroll
    for (int g = 0; g < G; g++) xg[g] = p.x + (size_t)min(b0 + g, p.B - 1) * T * D;

    auto issue_x = [&](int f) {
      if (!p.use_tma_x) return;
      if (tid == kLogWarpIssueX * 32) {   // the single-thread duties are spread over three warps (see den_shared.cu)
        fence_proxy_async();
        uint32_t bytes = 0;
#pragma unroll
        for (int g = 0; g < G; g++) bytes += (f < Lg[g]) ? (uint32_t)D * 4u : 0u;
        mbar_expect_tx(&mbar[1], bytes);
#pragma unroll
        for (int g = 0; g < G; g++)
          if (f < Lg[g]) tma_load_1d(xraw + (size_t)g * D_pad, xg[g] + (size_t)f * D, (uint32_t)D * 4u, &mbar[1]);
      }
    };
    // clamp while transposing into the interleaved x_s[pdf][g]; dead lanes get -inf (log zero)
    auto transform_x = [&](int f) {
      if (p.use_tma_x) { mbar_wait(&mbar[1], ph_x); ph_x ^= 1u; }
      for (int n = tid; n < D; n += DEN_NT) {
        float v[G];
#pragma unroll
        for (int g = 0; g < G; g++) {
          float t = -INFINITY;
          if (f < Lg[g]) {
            t = p.use_tma_x ? xraw[(size_t)g * D_pad + n] : __ldg(xg[g] + (size_t)f * D + n);
            if (clamp_x) t = clampf(t, kClampLo, kClampHi);
          }
          v[g] = t;
        }
        st_vec<G>(x_s + (size_t)n * G, v);
      }
    };

    double logp[G];
    float lZ[G], sc[G];
#pragma unroll
    for (int g = 0; g < G; g++) { logp[g] = 0.0; lZ[g] = 0.f; sc[g] = 0.f; }

    // log Z of the utterances that end at frame t: logsumexp_i(alpha(t,i) + final_i)
    auto finish_utts = [&](const float *Arow, int t) {
      bool need = false;
#pragma unroll
      for (int g = 0; g < G; g++) need |= (Lg[g] == t);
      if (!need) return;
      float zm[G], zs[G], z[G];
#pragma unroll
      for (int g = 0; g < G; g++) { zm[g] = -INFINITY; zs[g] = 0.f; }
      for (int i = tid; i < S; i += DEN_NT) {
        float v[G];
        ld_vec<G>(Arow + (size_t)i * G, v);
        const float f = __ldg(fin + i);
#pragma unroll
        for (int g = 0; g < G; g++) lse_add(zm[g], zs[g], v[g] + f);
      }
      block_lse_vec<G>(zm, zs, z, red);
#pragma unroll
      for (int g = 0; g < G; g++)
        if (Lg[g] == t) { lZ[g] = z[g]; logp[g] += (double)z[g]; }
    };

    // ============================================================ forward
    float *A = buf0, *Bn = buf1;
    __syncthreads();
    if (Lmax > 0) issue_x(0);
    for (int i = tid; i < S_pad; i += DEN_NT) {  // alpha(0) = init for every utterance; sc(0) = 0
      float v[G];
      const float a = (i < S) ? ini[i] : -INFINITY;
#pragma unroll
      for (int g = 0; g < G; g++) v[g] = a;
      st_vec<G>(A + (size_t)i * G, v);
      st_vec<G>(aws + (size_t)i * G, v);
    }
    if (tid == 0) {
      float tv[G];
#pragma unroll
      for (int g = 0; g < G; g++) tv[g] = 0.f;
      st_vec<G>(tws, tv);
    }
    __syncthreads();
    finish_utts(A, 0);

    for (int t = 1; t <= Lmax; t++) {
      transform_x(t - 1);
      __syncthreads();
      if (t < Lmax) issue_x(t);
      float pm[G], ps[G];
#pragma unroll
      for (int g = 0; g < G; g++) { pm[g] = -INFINITY; ps[g] = 0.f; }
      auto epi_fwd = [&](int row, const float (&o)[G]) {
        float a[G];
#pragma unroll
        for (int g = 0; g < G; g++) {
          a[g] = (o[g] > -INFINITY) ? o[g] - sc[g] : -INFINITY;
          lse_add(pm[g], ps[g], a[g]);
        }
        st_vec<G>(Bn + (size_t)row * G, a);
      };
      sell_pass_log<G, L1C, EMBED>(p.view[0], A, x_s, partial, warp, lane, epi_fwd);
      if (p.view[0].n_split) { __syncthreads(); fixup_split_log<G>(p.view[0], partial, epi_fwd); }
      float nsc[G];
      block_lse_vec<G>(pm, ps, nsc, red);  // contains the barrier that publishes alpha(t)
      float *awt = aws + (size_t)t * S_pad * G;
      for (int i = tid; i < S; i += DEN_NT) {  // spill alpha(t), coalesced
        float v[G];
        ld_vec<G>(Bn + (size_t)i * G, v);
        st_vec<G>(awt + (size_t)i * G, v);
      }
#pragma unroll
      for (int g = 0; g < G; g++) {
        sc[g] = (t <= Lg[g]) ? nsc[g] : 0.f;
        if (t < Lg[g] && nsc[g] > -INFINITY) logp[g] += (double)nsc[g];
      }
      if (tid == kLogWarpTot * 32) st_vec<G>(tws + (size_t)t * G, sc);
      finish_utts(Bn, t);
      float *tmp = A; A = Bn; Bn = tmp;
      __syncthreads();
    }
    if (tid == 0) {
#pragma unroll
      for (int g = 0; g < G; g++)
        if (b0 + g < p.B) p.logp[b0 + g] = (float)logp[g];
    }

    // ============================================================ backward
    float *P = buf0, *Q = buf1;
    for (int i = tid; i < S_pad * G; i += DEN_NT) Q[i] = -INFINITY;
    if (Lmax > 0) issue_x(Lmax - 1);
    __syncthreads();
    for (int t = Lmax - 1; t >= 0; t--) {
      bool inject = false;
#pragma unroll
      for (int g = 0; g < G; g++) inject |= (Lg[g] == t + 1);
      if (tid == kLogWarpFetchA * 32) {
        fence_proxy_async();
        mbar_expect_tx(&mbar[0], (uint32_t)S_pad * G * 4u);
        tma_load_1d(P, aws + (size_t)t * S_pad * G, (uint32_t)S_pad * G * 4u, &mbar[0]);
      }
      if (inject) {  // beta(L) = final - log Z
        for (int i = tid; i < S; i += DEN_NT) {
          float v[G];
          ld_vec<G>(Q + (size_t)i * G, v);
          const float f = __ldg(fin + i);
#pragma unroll
          for (int g = 0; g < G; g++)
            if (Lg[g] == t + 1) v[g] = f - lZ[g];
          st_vec<G>(Q + (size_t)i * G, v);
        }
      }
      transform_x(t);
      float sct[G];
      ld_vec<G>(tws + (size_t)t * G, sct);
      mbar_wait(&mbar[0], ph_a); ph_a ^= 1u;
      __syncthreads();
      if (t > 0) issue_x(t - 1);

      float gsum[G];
#pragma unroll
      for (int g = 0; g < G; g++) gsum[g] = 0.f;
      auto epi_gam = [&](int row, const float (&o)[G]) {
        float xv[G], out[G];
        ld_vec<G>(x_s + (size_t)row * G, xv);
#pragma unroll
        for (int g = 0; g < G; g++) {
          const float lg = (o[g] > -INFINITY && t < Lg[g]) ? o[g] + xv[g] - sct[g] : -INFINITY;
          const float gam = (lg > -INFINITY) ? expf(lg) : 0.f;
          out[g] = as_log ? lg : gam;
          gsum[g] += gam;
        }
        st_vec<G>(gam_s + (size_t)row * G, out);
      };
      sell_pass_log<G, L1C, EMBED>(p.view[2], P, Q, partial, warp, lane, epi_gam);
      if (p.view[2].n_split) { __syncthreads(); fixup_split_log<G>(p.view[2], partial, epi_gam); }
      __syncthreads();  // gamma rows complete; everybody done reading P

      auto epi_bwd = [&](int row, const float (&o)[G]) {
        float out[G];
#pragma unroll
        for (int g = 0; g < G; g++) out[g] = (o[g] > -INFINITY && t < Lg[g]) ? o[g] - sct[g] : -INFINITY;
        st_vec<G>(P + (size_t)row * G, out);
      };
      sell_pass_log<G, L1C, EMBED>(p.view[1], Q, x_s, partial, warp, lane, epi_bwd);
      if (p.view[1].n_split) { __syncthreads(); fixup_split_log<G>(p.view[1], partial, epi_bwd); }
      {  // this frame's gradient rows: one vector read per pdf, G coalesced row writes
        const float dead = as_log ? -INFINITY : 0.f;
        for (int n = tid; n < D; n += DEN_NT) {
          float v[G];
          ld_vec<G>(gam_s + (size_t)n * G, v);
#pragma unroll
          for (int g = 0; g < G; g++)
            if (b0 + g < p.B) p.grad[((size_t)(b0 + g) * T + t) * D + n] = (t < Lg[g]) ? (as_log ? v[g] : v[g] * p.grad_scale) : dead;
        }
      }
      if (t == 0) {
        float z[G];
#pragma unroll
        for (int g = 0; g < G; g++) z[g] = gsum[g];
        block_sum_vec<G>(z, red);
        if (tid == 0) {
#pragma unroll
          for (int g = 0; g < G; g++)
            if (b0 + g < p.B) { p.diag[2 * (b0 + g)] = 1.f; p.diag[2 * (b0 + g) + 1] = z[g]; }
        }
      }
      float *tmp = P; P = Q; Q = tmp;
      __syncthreads();
    }
#pragma unroll
    for (int g = 0; g < G; g++) {
      if (b0 + g >= p.B) continue;
      if (Lg[g] == 0 && tid == 0) { p.diag[2 * (b0 + g)] = 1.f; p.diag[2 * (b0 + g) + 1] = 1.f; }
      float *gr = p.grad + ((size_t)(b0 + g) * T + Lmax) * D;
      const size_t n = (size_t)(T - Lmax) * D;
      const float dead = as_log ? -INFINITY : 0.f;
      for (size_t i = tid; i < n; i += DEN_NT) gr[i] = dead;
    }
  }
}

int launch_den_shared_log(const BlobHeader &h, const DenParams &p, int G, bool l1c, int grid, size_t smem,
                          cudaStream_t stream) {
  cudaError_t e;
#define LAUNCH_LOG(GG, LL, EE)                                                                                              \
  do {                                                                                                                      \
    e = cudaFuncSetAttribute(den_shared_log_kernel<GG, LL, EE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);   \
    if (e != cudaSuccess) return (int)e;                                                                                    \
    den_shared_log_kernel<GG, LL, EE><<<grid, DEN_NT, smem, stream>>>(p);                                                   \
  } while (0)
#define LAUNCH_LOG_G(GG)                                       \
  do {                                                         \
    if (h.embed_codes) LAUNCH_LOG(GG, false, true);            \
    else if (l1c) LAUNCH_LOG(GG, true, false);                 \
    else LAUNCH_LOG(GG, false, false);                         \
  } while (0)
  if (G == 4) LAUNCH_LOG_G(4);
  else if (G == 2) LAUNCH_LOG_G(2);
  else LAUNCH_LOG_G(1);
#undef LAUNCH_LOG_G
#undef LAUNCH_LOG
  return (int)cudaGetLastError();
}

}  // namespace chainb200
