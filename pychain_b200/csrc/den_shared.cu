// den_shared.cu -- denominator forward-backward for ONE graph shared by the batch, sm_100a.
//
// This is the headline path: the reference runs it as ChainGraphBatch(den_graph, B)
// (pychain/loss.py:99) -> pychain_C.forward_backward (pychain.cc:26-79) ->
// ChainComputation::Forward/Backward (chain-computation.cc:196-207, 332-342), i.e. one
// kernel + 3 ATen launches per frame and direction, a thread per (sequence, state) with the
// sequence index as the SLOWEST memory dimension, and per-arc thresholded atomics for the
// gradient (chain-kernels.cu:97-245).
//
// Design here:
//  * persistent CTAs, each owning a GROUP of G (1, 2 or 4) utterances for the whole
//    recursion; the G utterances are interleaved as the fastest dimension of every
//    on-chip array (alpha[state][g], x[pdf][g]) so that one arc record fetched from L2 and
//    one vector LDS per operand serve G arc-visits.
//  * the graph is packed ONCE on the host (chain_b200_den_prepare) into three sliced-ELL
//    ("SELL-32") views -- keyed by destination (alpha pass), by source (beta pass) and by
//    pdf (gamma pass) -- rows sorted by degree, 32 rows per slice, records stored
//    [slice][k][lane] as 8-byte {a:16 | b:16, weight:f32}: every warp-level record fetch is
//    one coalesced 256-byte read, every lane owns one row and accumulates in registers.
//    No atomics anywhere; summation order inside a row is the reference's CSR order.
//  * alpha'(t) rows ([S][G] floats, contiguous) are spilled with coalesced 128-bit stores
//    and brought back for the backward pass with one TMA bulk copy (cp.async.bulk +
//    mbarrier) per frame; the frames of x are staged by TMA bulk copies one frame ahead and
//    exp/clamp are applied while transposing them into the interleaved layout.
//  * leaky-HMM mixing, per-frame renormalisation, total log-likelihood, the t=0 self-check
//    and the zeroing of padded frames are all inside the one kernel.
#include <cstring>

#include "den_kernel_common.cuh"

namespace chainb200 {

// Blackwell packed fp32 math (SASS FMUL2 / FFMA2): two utterances per instruction, same rounding as
// the scalar fmaf(a * w, b, acc).
__device__ __forceinline__ void mul_fma_x2(float &acc0, float &acc1, float a0, float a1, float w, float b0, float b1) {
  asm("{\n\t.reg .b64 ra, rw, rb, rc;\n\t"
      "mov.b64 ra, {%2, %3};\n\tmov.b64 rw, {%4, %4};\n\tmov.b64 rb, {%5, %6};\n\tmov.b64 rc, {%0, %1};\n\t"
      "mul.rn.f32x2 ra, ra, rw;\n\tfma.rn.f32x2 rc, ra, rb, rc;\n\tmov.b64 {%0, %1}, rc;\n\t}"
      : "+f"(acc0), "+f"(acc1)
      : "f"(a0), "f"(a1), "f"(w), "f"(b0), "f"(b1));
}

// acc += a * w * b  /  acc += a * w
template <int G>
__device__ __forceinline__ void fma_rec(float (&acc)[G], const float (&a)[G], const float w, const float (&b)[G]) {
  if constexpr (G % 2 == 0) {
#pragma unroll
    for (int g = 0; g < G; g += 2) mul_fma_x2(acc[g], acc[g + 1], a[g], a[g + 1], w, b[g], b[g + 1]);
  } else {
#pragma unroll
    for (int g = 0; g < G; g++) acc[g] = fmaf(a[g] * w, b[g], acc[g]);
  }
}
template <int G>
__device__ __forceinline__ void fma_rec1(float (&acc)[G], const float (&a)[G], const float w) {
#pragma unroll
  for (int g = 0; g < G; g++) acc[g] = fmaf(a[g], w, acc[g]);
}

template <int G>
__device__ __forceinline__ void consume_rec(const uint32_t ab, const float w, const float *__restrict__ A,
                                            const float *__restrict__ Bv, float (&acc)[G]) {
  float a[G], b[G];
  ld_vec<G>(A + (ab & 0xffffu) * G, a);
  ld_vec<G>(Bv + (ab >> 16) * G, b);
  if constexpr (G % 2 == 0) {
#pragma unroll
    for (int g = 0; g < G; g += 2) mul_fma_x2(acc[g], acc[g + 1], a[g], a[g + 1], w, b[g], b[g + 1]);
  } else {
#pragma unroll
    for (int g = 0; g < G; g++) acc[g] = fmaf(a[g] * w, b[g], acc[g]);
  }
}

// Record pairs per register bank of the two-operand passes.  At G = 4 a pair in flight holds 16 operand registers
// next to the two banks, and with kFifoPairs = 4 ptxas spills bank entries right after their loads (which serialises
// the warp on the L2 latency); one pair less keeps the loop spill-free.
#ifndef CHAIN_B200_NF2_G4
#define CHAIN_B200_NF2_G4 (kFifoPairs - 1)
#endif
template <int G> __host__ __device__ constexpr int nf2() { return G == 4 ? CHAIN_B200_NF2_G4 : kFifoPairs; }
#ifndef CHAIN_B200_NF2A_G4
#define CHAIN_B200_NF2A_G4 CHAIN_B200_NF2_G4
#endif
template <int G> __host__ __device__ constexpr int nf2a() { return G == 4 ? CHAIN_B200_NF2A_G4 : kFifoPairs; }  // alpha pass
#ifndef CHAIN_B200_PREFETCH_AHEAD
#define CHAIN_B200_PREFETCH_AHEAD 1
#endif
constexpr int kPrefetchAhead = CHAIN_B200_PREFETCH_AHEAD;   // banks requested into L1 beyond the two register banks
#ifndef CHAIN_B200_PAIR_FENCE
#define CHAIN_B200_PAIR_FENCE 1
#endif
constexpr bool kPairFence = CHAIN_B200_PAIR_FENCE != 0;
// One-operand passes (class views) have cheap pairs: a bank must cover the L2 latency with fewer cycles of work
// per pair, and they hold 8 fewer operand registers, so their banks are deeper.  (The streams' run-out, 3 *
// kFifoPairs pairs, covers banks of up to 3 * kFifoPairs pairs.)
#ifndef CHAIN_B200_NF1_G4
#define CHAIN_B200_NF1_G4 5
#endif
template <int G> __host__ __device__ constexpr int nf1() { return G == 4 ? CHAIN_B200_NF1_G4 : kFifoPairs; }

// One-operand record (class views 3 and 4): acc += A[a] * w.
template <int G>
__device__ __forceinline__ void consume_rec1(const uint32_t ab, const float w, const float *__restrict__ A, float (&acc)[G]) {
  float a[G];
  ld_vec<G>(A + (ab & 0xffffu) * G, a);
#pragma unroll
  for (int g = 0; g < G; g++) acc[g] = fmaf(a[g], w, acc[g]);
}

// One SELL pass: every lane owns one row (or sub-row) of each of its warp's slices and accumulates
// sum_k w_k * A[a_k][g] * Bv[b_k][g] over the row's records (ONE: sum_k w_k * A[a_k][g]).  The warp's slots are one
// contiguous stream consumed in pairs by a flat counted loop; two register banks of NF pairs alternate: while one
// bank is consumed the other is being filled, so NF pairs of 128-bit record loads are always in flight per lane.
// The epilogue is a small functor passed and returned BY VALUE (its accumulators stay in registers).
//
// Bank depth is what the 128 registers of a 512-thread CTA allow: measured on B200 (cfg4, ms per call, uniform /
// biphone graph): banks of 3 (two-operand) and 5 (one-operand) pairs 8.00 / 5.84-5.91; banks of 4 and 5: 8.11-8.54 /
// 6.14-6.18 (ptxas then spills bank entries right after their loads, which serialises the warp on the L2 latency the
// banks exist to hide); 12 warps x banks of 6/8: 10.6 / 7.5; 8 warps x banks of 8/10: 9.9 / 7.6.  Compiling the pass
// as a real function (CHAIN_B200_INLINE_PASS=0: the loop gets the register file to itself, the caller's state is
// saved once per pass) measures the same as inlining it (8.00 / 5.84), with or without the pair fence below.
#ifndef CHAIN_B200_INLINE_PASS
#define CHAIN_B200_INLINE_PASS 1
#endif
#if CHAIN_B200_INLINE_PASS
#define CHAIN_B200_PASS_ATTR __forceinline__
#else
#define CHAIN_B200_PASS_ATTR __noinline__
#endif
// EMBED: stream format of the blob (den_blob.h): row codes in the last slot of every slice, or (compact format)
// a parallel stream `rp` of row codes read two slices ahead, the slice ending on the sign bit of its last arc's
// weight and the epilogue's per-row float coming from `aux_tab`.
// RES (= L1C && EMBED): the graph is small enough for its slot streams to live in shared memory (copied once per CTA,
// `sp_s`): a pair is one 128-bit LDS away, there is no L2 latency to hide and no register banks.
template <int G, bool L1C, bool EMBED, bool ONE, int NF, class Epi>
__device__ CHAIN_B200_PASS_ATTR Epi sell_pass(const uint4 *__restrict__ sp, const uint32_t sp_s, const int npairs,
                                      const uint32_t A_s, const uint32_t B_s, const uint32_t amax, const uint32_t bmax,
                                      const uint32_t part_s, const int32_t *__restrict__ rp,
                                      const float *__restrict__ aux_tab, Epi epi) {
  constexpr bool RES = L1C && EMBED;
  uint4 fa[RES ? 1 : NF], fb[RES ? 1 : NF];
  int row = 0, row_n1 = 0, row_n2 = 0;
  if constexpr (!EMBED) {
    row = __ldg(rp); row_n1 = __ldg(rp + 32); row_n2 = __ldg(rp + 64);
    rp += 96;
  }
  if constexpr (!RES) {
#pragma unroll
    for (int d = 0; d < NF; d++) fa[d] = ldg_stream<L1C>(sp + d * 32);
    sp += NF * 32;
  }
  float acc[G];
#pragma unroll
  for (int g = 0; g < G; g++) acc[g] = 0.f;
  // A pair is two arcs, or -- at the end of a slice -- an arc and the lane's ROW-CODE slot (sign bit of its second
  // word; warp-uniform): {row code, -aux}.  The code arrives through the same prefetched stream as the arcs.
  // The operands of BOTH slots are gathered before the flag is looked at: a row-code slot is not an arc, so its
  // index fields are clamped (amax / bmax) into the part of the rows that real arcs of this view read -- never into
  // a region another warp may be writing during the pass -- and what they fetch is dropped.
  auto process = [&](const uint2 r0, const uint2 r1) {
    const bool last = (r1.y & 0x80000000u) != 0;                  // warp-uniform
    const bool arc1 = !EMBED || !last;                            // compact format: the flagged slot is still an arc
    const float w1 = __uint_as_float(EMBED ? r1.y : (r1.y & 0x7fffffffu));
    if constexpr (ONE) {
      float a0[G], a1[G];
      lds_vec<G>(A_s + (r0.x & 0xffffu) * (G * 4u), a0);
      lds_vec<G>(A_s + (EMBED ? min(r1.x & 0xffffu, amax) : (r1.x & 0xffffu)) * (G * 4u), a1);
      fma_rec1<G>(acc, a0, __uint_as_float(r0.y));
      if (arc1) fma_rec1<G>(acc, a1, w1);
    } else {
      float a0[G], a1[G], b0[G], b1[G];
      lds_vec<G>(A_s + (r0.x & 0xffffu) * (G * 4u), a0);
      lds_vec<G>(A_s + (EMBED ? min(r1.x & 0xffffu, amax) : (r1.x & 0xffffu)) * (G * 4u), a1);
      lds_vec<G>(B_s + (r0.x >> 16) * (G * 4u), b0);
      lds_vec<G>(B_s + (EMBED ? min(r1.x >> 16, bmax) : (r1.x >> 16)) * (G * 4u), b1);
      fma_rec<G>(acc, a0, __uint_as_float(r0.y), b0);
      if (arc1) fma_rec<G>(acc, a1, w1, b1);
    }
    if (last) {
      if constexpr (EMBED) row = (int)r1.x;
      if (row >= 0) {
        float aux = 0.f;
        if constexpr (EMBED) aux = __uint_as_float(r1.y & 0x7fffffffu);
        else if (aux_tab) aux = __ldg(aux_tab + (row & (kRowExtraFlag - 1)));
        epi(row, acc, aux);
      } else if (row <= -2) {
        sts_vec<G>(part_s + (uint32_t)(-2 - row) * (G * 4u), acc);
      }
#pragma unroll
      for (int g = 0; g < G; g++) acc[g] = 0.f;
      if constexpr (!EMBED) {
        row = row_n1;
        row_n1 = row_n2;
        row_n2 = __ldg(rp);
        rp += 32;
      }
    }
    if constexpr (kPairFence) {
      // keep the gathers of the next pair behind this pair's arithmetic: the operand registers of ONE pair are live at
      // a time, the rest of the register file holds record banks
#pragma unroll
      for (int g = 0; g < G; g++) asm volatile("" : "+f"(acc[g]));
    }
  };
  if constexpr (RES) {
    uint32_t at = sp_s;
    int j = 0;
    for (; j + 4 <= npairs; j += 4, at += 4 * 512) {   // four pairs' slots fetched together, then consumed
      uint4 q[4];
#pragma unroll
      for (int d = 0; d < 4; d++)
        asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(q[d].x), "=r"(q[d].y), "=r"(q[d].z), "=r"(q[d].w) : "r"(at + d * 512));
#pragma unroll
      for (int d = 0; d < 4; d++) process(make_uint2(q[d].x, q[d].y), make_uint2(q[d].z, q[d].w));
    }
    for (; j < npairs; j++, at += 512) {
      uint4 q;
      asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(q.x), "=r"(q.y), "=r"(q.z), "=r"(q.w) : "r"(at));
      process(make_uint2(q.x, q.y), make_uint2(q.z, q.w));
    }
    return epi;
  }
  int j = 0;
  for (; j + 2 * NF <= npairs; j += 2 * NF) {  // full double banks: straight-line code
#pragma unroll
    for (int d = 0; d < NF; d++) fb[d] = ldg_stream<L1C>(sp + d * 32);
    sp += NF * 32;
    if constexpr (kPrefetchAhead > 0 && !L1C) {   // the banks after the next: requested into L1, no registers held
#pragma unroll
      for (int d = 0; d < kPrefetchAhead * NF; d++) asm volatile("prefetch.global.L1 [%0];" ::"l"(sp + (NF + d) * 32));
    }
#pragma unroll
    for (int d = 0; d < NF; d++) process(make_uint2(fa[d].x, fa[d].y), make_uint2(fa[d].z, fa[d].w));
#pragma unroll
    for (int d = 0; d < NF; d++) fa[d] = ldg_stream<L1C>(sp + d * 32);
    sp += NF * 32;
#pragma unroll
    for (int d = 0; d < NF; d++) process(make_uint2(fb[d].x, fb[d].y), make_uint2(fb[d].z, fb[d].w));
  }
  const int rem = npairs - j;  // < 2*NF pairs left: bank A holds the first NF of them
  if (rem > 0) {
    if (rem > NF) {
#pragma unroll
      for (int d = 0; d < NF; d++) fb[d] = ldg_stream<L1C>(sp + d * 32);
    }
#pragma unroll
    for (int d = 0; d < NF; d++)
      if (d < rem) process(make_uint2(fa[d].x, fa[d].y), make_uint2(fa[d].z, fa[d].w));
    if (rem > NF) {
#pragma unroll
      for (int d = 0; d < NF; d++)
        if (NF + d < rem) process(make_uint2(fb[d].x, fb[d].y), make_uint2(fb[d].z, fb[d].w));
    }
  }
  return epi;
}

// Launch one pass for this warp: stream start / length of view `v`.
template <int G, bool L1C, bool EMBED, bool ONE, int NF, class Epi>
__device__ __forceinline__ Epi run_pass(const DView &v, uint32_t rec_s /* this warp's resident stream of the view */,
                                        uint32_t A_s, uint32_t B_s, uint32_t amax, uint32_t bmax, uint32_t part_s,
                                        const float *aux_tab, int warp, int lane, Epi epi) {
  const uint4 *sp = reinterpret_cast<const uint4 *>(v.recs + v.warp_info[3 * warp]) + lane;  // [pair][lane]: 2 slots
  const int32_t *rp = EMBED ? nullptr : v.rows + v.warp_info[3 * warp + 2] + lane;
  return sell_pass<G, L1C, EMBED, ONE, NF, Epi>(sp, rec_s + lane * 16u, (int)v.warp_info[3 * warp + 1], A_s, B_s, amax,
                                                bmax, part_s, rp, aux_tab, epi);
}

// Hint the head of the stream this warp walks NEXT into L1 (no registers): a pass begins by waiting a full L2
// latency for its first bank, eight times per frame pair; the next pass's first pairs are requested while this warp
// still sits in the current phase's tail and barrier.  (Streaming loads hit L1 when the line is there.)
#ifndef CHAIN_B200_PREFETCH_PAIRS
#define CHAIN_B200_PREFETCH_PAIRS 3
#endif
template <bool L1C, bool EMBED>
__device__ __forceinline__ void prefetch_head(const DView &v, int warp, int lane) {
  if constexpr (!(L1C && EMBED) && CHAIN_B200_PREFETCH_PAIRS > 0) {
    const uint4 *sp = reinterpret_cast<const uint4 *>(v.recs + v.warp_info[3 * warp]) + lane;
#pragma unroll
    for (int d = 0; d < CHAIN_B200_PREFETCH_PAIRS; d++) asm volatile("prefetch.global.L1 [%0];" ::"l"(sp + d * 32));
  }
}

// ---- the epilogues of the five passes: what happens to a finished row --------------------------------------
template <int G> struct EpiClass {   // view 3: class sum -> extension of the state row (shared) + its spill (global)
  uint32_t ext_s; float *ext_g;
  __device__ __forceinline__ void operator()(int row, const float (&acc)[G], float) {
    sts_vec<G>(ext_s + (uint32_t)row * (G * 4u), acc);
    st_vec<G>(ext_g + (size_t)row * G, acc);
  }
};
template <int G> struct EpiAlpha {   // view 0: alpha(t,row), and its contribution to tot(t)
  uint32_t out_s; float part[G];
  __device__ __forceinline__ void operator()(int row, const float (&acc)[G], float) {
#pragma unroll
    for (int g = 0; g < G; g++) part[g] += acc[g];
    sts_vec<G>(out_s + (uint32_t)row * (G * 4u), acc);
  }
};
template <int G> struct EpiGamma {   // view 2: gamma(t,pdf) = x(t,pdf)/tot(t) * sum
  uint32_t xs_s, gam_s;
  __device__ __forceinline__ void operator()(int row, const float (&acc)[G], float) {
    float xv[G], o[G];
    lds_vec<G>(xs_s + (uint32_t)row * (G * 4u), xv);
#pragma unroll
    for (int g = 0; g < G; g++) o[g] = acc[g] * xv[g];
    sts_vec<G>(gam_s + (uint32_t)row * (G * 4u), o);
  }
};
template <int G> struct EpiStore {   // view 4: raw class-arc sums of beta', parked in P
  uint32_t out_s;
  __device__ __forceinline__ void operator()(int row, const float (&acc)[G], float) {
    sts_vec<G>(out_s + (uint32_t)row * (G * 4u), acc);
  }
};
template <int G> struct EpiBeta {    // view 1: beta'(t,row) (+ what view 4 parked), and sum_i beta'(t,i) leaky_i
  uint32_t out_s; float lpart[G];
  __device__ __forceinline__ void operator()(int code, const float (&acc)[G], float l /* leaky(row) */) {
    const uint32_t at = out_s + (uint32_t)(code & (kRowExtraFlag - 1)) * (G * 4u);
    float o[G];
#pragma unroll
    for (int g = 0; g < G; g++) o[g] = acc[g];
    if (code & kRowExtraFlag) {  // this state also has class arcs
      float e[G];
      lds_vec<G>(at, e);
#pragma unroll
      for (int g = 0; g < G; g++) o[g] += e[g];
    }
#pragma unroll
    for (int g = 0; g < G; g++) lpart[g] = fmaf(o[g], l, lpart[g]);
    sts_vec<G>(at, o);
  }
};

// Rows that were split: add the sub-row partial sums (fixed order) and run the epilogue.  `aux_tab`: the per-row
// float the row-code slot would have carried (leaky probabilities for the beta' pass), or null.
template <int G, class Epi>
__device__ __forceinline__ void fixup_split(const DView &v, const float *__restrict__ partial,
                                            const float *__restrict__ aux_tab, Epi &epi) {
  for (int i = threadIdx.x; i < v.n_split; i += DEN_NT) {
    const int row = v.split[3 * i], first = v.split[3 * i + 1], cnt = v.split[3 * i + 2];
    float acc[G];
#pragma unroll
    for (int g = 0; g < G; g++) acc[g] = 0.f;
    for (int j = 0; j < cnt; j++) {
      float t[G];
      ld_vec<G>(partial + (size_t)(first + j) * G, t);
#pragma unroll
      for (int g = 0; g < G; g++) acc[g] += t[g];
    }
    epi(row, acc, aux_tab ? __ldg(aux_tab + (row & (kRowExtraFlag - 1))) : 0.f);
  }
}

// L1C: the three record streams fit in the L1 next to this CTA's shared memory, so they are
// loaded with L1 allocation and hit L1 from the second frame on (small graphs are latency bound);
// otherwise they are streamed from L2 without polluting L1.
// The per-frame single-thread duties -- issuing the TMA copy of the next frame of x, issuing the reload of an alpha
// row, the log-likelihood bookkeeping -- sit at the start of a phase every other warp then waits out at the next
// barrier.  They are given to three different warps (and the bookkeeping to one lane per utterance) instead of all
// to thread 0, so no single warp carries the sum of them into every phase.
#ifndef CHAIN_B200_EARLY_LK
#define CHAIN_B200_EARLY_LK 1
#endif
#ifndef CHAIN_B200_EARLY_TOT
#define CHAIN_B200_EARLY_TOT 1
#endif
#ifndef CHAIN_B200_EARLY_CLS
#define CHAIN_B200_EARLY_CLS 1
#endif
constexpr bool kEarlyLeaky = CHAIN_B200_EARLY_LK != 0, kEarlyTot = CHAIN_B200_EARLY_TOT != 0, kEarlyCls = CHAIN_B200_EARLY_CLS != 0;
constexpr int kWarpIssueX = DEN_NW - 1, kWarpFetchA = DEN_NW > 1 ? DEN_NW - 2 : 0, kWarpTot = DEN_NW > 2 ? DEN_NW - 3 : 0;

template <int G, bool L1C, bool EMBED>
__global__ void __launch_bounds__(DEN_NT, 1) den_shared_kernel(const DenParams p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ __align__(8) unsigned long long mbar[2];  // 0: alpha rows, 1: x frames
  __shared__ float red[DEN_NW * 4];
  __shared__ float red2[DEN_NW * 4];
  __shared__ double logp_s[4];
  __shared__ float invz_s[4];
  __shared__ int Lg[4];

  const int S = p.S, D = p.D, T = p.T, S_pad = p.S_pad, D_pad = p.D_pad;
  const int R_pad = p.R_pad, C = p.C;   // a state row = [S_pad states | C_pad class sums] x G utterances
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float c = p.coeff;
  float *buf0 = reinterpret_cast<float *>(smem_raw);
  float *buf1 = buf0 + (size_t)R_pad * G;
  float *x_s = buf1 + (size_t)R_pad * G;
  float *xraw = x_s + (size_t)D_pad * G;
  float *gam_s = xraw + (size_t)D_pad * G;
  float *partial = gam_s + (size_t)D_pad * G;
  const uint32_t part_sa = smem_u32(partial);
  // records resident in shared memory (small graphs): the area after the split-row slots, 16-byte aligned
  const uint32_t rec_sa = (smem_u32(partial + (size_t)p.max_sub * G) + 15u) & ~15u;
  const float *__restrict__ lk = p.leaky;
  const float *__restrict__ ini = p.init;
  const float *__restrict__ fin = p.fin;
  const bool clamp_x = (p.flags & CHAIN_B200_CLAMP_X) && !(p.flags & CHAIN_B200_X_IS_EXP);
  const bool exp_x = !(p.flags & CHAIN_B200_X_IS_EXP);

  // the packed views are read straight from the kernel parameters (constant bank): no registers held
  const DView *views = p.view;
  if constexpr (L1C && EMBED) {   // every warp brings in the slot streams it will walk, once
    for (int v = 0; v < kNumViews; v++) {
      const uint4 *src = reinterpret_cast<const uint4 *>(views[v].recs + views[v].warp_info[3 * warp]);
      const int n = (int)views[v].warp_info[3 * warp + 1] * 32;
      const uint32_t dst = rec_sa + p.rs_off[v][warp];
      for (int i = lane; i < n; i += 32) {
        const uint4 q = __ldg(src + i);
        asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(dst + (uint32_t)i * 16u), "r"(q.x), "r"(q.y), "r"(q.z), "r"(q.w) : "memory");
      }
    }
    __syncwarp();
  }
  if (tid == 0) {
    mbar_init(&mbar[0], 1);
    mbar_init(&mbar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  uint32_t ph_a = 0, ph_x = 0;

  // graph constants: sum_i init_i and sum_i final_i * leaky_i
  float tot0, FL;
  {
    float v[2] = {0.f, 0.f};
    for (int i = tid; i < S; i += DEN_NT) { v[0] += ini[i]; v[1] += fin[i] * lk[i]; }
    block_sum_vec<2>(v, red);
    tot0 = v[0]; FL = v[1];
  }

  for (int gi = blockIdx.x; gi < p.num_groups; gi += gridDim.x) {
    const int b0 = gi * G;
    // lengths of the group's utterances: shared memory (read where needed), not registers held for the whole recursion
    int Lmax = 0;
#pragma unroll
    for (int g = 0; g < G; g++) {
      long long l = (b0 + g < p.B) ? p.lengths[b0 + g] : 0;
      const int lx = (int)(l < 0 ? 0 : (l > T ? T : l));
      if (tid == 0) Lg[g] = lx;
      Lmax = max(Lmax, lx);
    }
    float *aws = p.alpha_ws + (size_t)gi * (T + 1) * R_pad * G;
    float *tws = p.tot_ws + (size_t)gi * (T + 1) * G;
    auto xrow = [&](int g, int f) { return p.x + ((size_t)min(b0 + g, p.B - 1) * T + f) * D; };  // frame f of utterance g

    // stage one frame of x: TMA bulk copies into xraw (one per live utterance) ...
    auto issue_x = [&](int f) {
      if (!p.use_tma_x) return;
      if (tid == kWarpIssueX * 32) {
        fence_proxy_async();
        uint32_t bytes = 0;
#pragma unroll
        for (int g = 0; g < G; g++) bytes += (f < Lg[g]) ? (uint32_t)D * 4u : 0u;
        mbar_expect_tx(&mbar[1], bytes);
#pragma unroll
        for (int g = 0; g < G; g++)
          if (f < Lg[g]) tma_load_1d(xraw + (size_t)g * D_pad, xrow(g, f), (uint32_t)D * 4u, &mbar[1]);
      }
    };
    // ... then clamp / exp while transposing into the interleaved x_s[pdf][g]; dead lanes get 0.  The frame's
    // normaliser 1/tot(f) is folded in here (every use of x(f,.) is divided by tot(f): chain-computation.cc:150-173,
    // 289-309), so no epilogue multiplies by it and it is not held in registers across a pass.
    auto transform_x = [&](int f, const float (&sc)[G]) {
      if (p.use_tma_x) { mbar_wait(&mbar[1], ph_x); ph_x ^= 1u; }
      for (int n = tid; n < D; n += DEN_NT) {
        float v[G];
#pragma unroll
        for (int g = 0; g < G; g++) {
          float t = 0.f;
          if (f < Lg[g]) {
            t = p.use_tma_x ? xraw[(size_t)g * D_pad + n] : __ldg(xrow(g, f) + n);
            if (clamp_x) t = clampf(t, kClampLo, kClampHi);
            if (exp_x) t = expf(t);
            t *= sc[g];
          }
          v[g] = t;
        }
        st_vec<G>(x_s + (size_t)n * G, v);
      }
    };

    double *logp = logp_s;  // only thread 0 accumulates the log-likelihoods (shared memory: no registers held)
    float *invZ = invz_s;   // 1/Z per utterance: written once in the forward pass, read when an utterance enters the backward pass
    float inv[G];   // 1/tot of the frame about to be staged; dead during the passes
#pragma unroll
    for (int g = 0; g < G; g++) {
      if (tid == 0) { logp[g] = 0.0; invZ[g] = 0.f; }
      inv[g] = tot0 > 0.f ? 1.0f / tot0 : 0.f;
    }

    // ============================================================ forward
    float *A = buf0, *Bn = buf1;  // A: alpha'(t-1) (read), Bn: alpha(t) (written)
    __syncthreads();               // previous group is completely done with smem
    if (Lmax > 0) issue_x(0);
    // t = 0 : alpha(0) = init (same for every utterance), alpha' = alpha + c*tot*leaky
    {
      bool need_z = false;
#pragma unroll
      for (int g = 0; g < G; g++) need_z |= (Lg[g] == 0);
      float zp[G];
#pragma unroll
      for (int g = 0; g < G; g++) zp[g] = 0.f;
      const float ct = c * tot0;
      for (int i = tid; i < S_pad; i += DEN_NT) {
        float v[G];
        const float a = (i < S) ? ini[i] + ct * lk[i] : 0.f;
        const float f = (need_z && i < S) ? fin[i] : 0.f;
#pragma unroll
        for (int g = 0; g < G; g++) { v[g] = a; zp[g] += a * f; }
        st_vec<G>(A + (size_t)i * G, v);
        st_vec<G>(aws + (size_t)i * G, v);
      }
      if (tid == 0) {
        float tv[G];
#pragma unroll
        for (int g = 0; g < G; g++) { tv[g] = tot0; if (0 < Lg[g]) logp[g] += (double)logf(tot0); }
        st_vec<G>(tws, tv);
      }
      if (need_z) {
        block_sum_vec<G>(zp, red2);
#pragma unroll
        for (int g = 0; g < G; g++)
          if (Lg[g] == 0 && tid == 0) { invZ[g] = 1.0f / zp[g]; logp[g] += (double)logf(zp[g]); }
      }
    }
    __syncthreads();

    for (int t = 1; t <= Lmax; t++) {
      transform_x(t - 1, inv);   // x(t-1,.) / tot(t-1)
      if (C) {
        // class sums A_c(t-1) = sum_{arcs of class c} alpha'(t-1,src) * w: one operand per arc.  They extend the
        // state row (virtual states S_pad + c) and are spilled with it: the backward pass needs them again.
        EpiClass<G> epi_cls{smem_u32(A + (size_t)S_pad * G), aws + ((size_t)(t - 1) * R_pad + S_pad) * G};
        epi_cls = run_pass<G, L1C, EMBED, true, nf1<G>()>(views[3], rec_sa + p.rs_off[3][warp], smem_u32(A), 0u, S - 1, 0, part_sa, nullptr, warp, lane, epi_cls);
        if (p.view[3].n_split) { __syncthreads(); fixup_split<G>(p.view[3], partial, nullptr, epi_cls); }
      }
      prefetch_head<L1C, EMBED>(views[0], warp, lane);
      __syncthreads();  // x_s complete, xraw free, class sums complete
      if (t < Lmax) issue_x(t);
      EpiAlpha<G> epi_fwd;
      epi_fwd.out_s = smem_u32(Bn);
#pragma unroll
      for (int g = 0; g < G; g++) epi_fwd.part[g] = 0.f;
      epi_fwd = run_pass<G, L1C, EMBED, false, nf2a<G>()>(views[0], rec_sa + p.rs_off[0][warp], smem_u32(A), smem_u32(x_s), R_pad - 1, D - 1,
                                                          part_sa, nullptr, warp, lane, epi_fwd);
      if (p.view[0].n_split) { __syncthreads(); fixup_split<G>(p.view[0], partial, nullptr, epi_fwd); }
      prefetch_head<L1C, EMBED>(views[C ? 3 : 0], warp, lane);
#pragma unroll
      for (int g = 0; g < G; g++) {
        const float ps = warp_sum(epi_fwd.part[g]);
        if (lane == 0) red[warp * G + g] = ps;
      }
      // The leaky (and final) probabilities of the rows this thread adds them to come from L2; the first four rows'
      // worth are requested BEFORE the barrier that ends the alpha pass, so their latency runs under the wait for
      // the slowest warp instead of after it.
      bool need_z = false;
#pragma unroll
      for (int g = 0; g < G; g++) need_z |= (Lg[g] == t);
      float l[4], f[4];
      if constexpr (kEarlyLeaky) {
#pragma unroll
        for (int u = 0; u < 4; u++) {
          const int i = tid + u * DEN_NT;
          l[u] = (i < S) ? __ldg(lk + i) : 0.f;
          f[u] = (need_z && i < S) ? __ldg(fin + i) : 0.f;
        }
      }
      __syncthreads();  // alpha(t) rows + partial sums visible; everybody done reading A
      float ct[G], zp[G];
#pragma unroll
      for (int g = 0; g < G; g++) {
        float s = 0.f;
        for (int w = 0; w < DEN_NW; w++) s += red[w * G + g];
        ct[g] = c * s; zp[g] = 0.f;   // ct = c * tot(t)
        inv[g] = s > 0.f ? 1.0f / s : 0.f;
      }
      float *awt = aws + (size_t)t * R_pad * G;
      // alpha'(t) = alpha(t) + c tot(t) leaky; spill
      for (int i0 = tid; i0 < S; i0 += 4 * DEN_NT) {
        if (!kEarlyLeaky || i0 != tid) {
#pragma unroll
          for (int u = 0; u < 4; u++) {
            const int i = i0 + u * DEN_NT;
            l[u] = (i < S) ? __ldg(lk + i) : 0.f;
            f[u] = (need_z && i < S) ? __ldg(fin + i) : 0.f;
          }
        }
#pragma unroll
        for (int u = 0; u < 4; u++) {
          const int i = i0 + u * DEN_NT;
          if (i < S) {
            float v[G];
            ld_vec<G>(Bn + (size_t)i * G, v);
#pragma unroll
            for (int g = 0; g < G; g++) { v[g] = fmaf(ct[g], l[u], v[g]); zp[g] += v[g] * f[u]; }
            st_vec<G>(Bn + (size_t)i * G, v);
            st_vec<G>(awt + (size_t)i * G, v);
          }
        }
      }
      if (warp == kWarpTot && lane < G) {   // one lane per utterance (not thread 0 for all: see kWarpIssueX)
        const int g = lane;
        float s = 0.f;
        for (int w = 0; w < DEN_NW; w++) s += red[w * G + g];   // tot(t), same order as above
        if (t < Lg[g]) logp[g] += (double)logf(s);
        tws[(size_t)t * G + g] = s;
      }
      if (need_z) {
        block_sum_vec<G>(zp, red2);
#pragma unroll
        for (int g = 0; g < G; g++)
          if (Lg[g] == t && tid == 0) { invZ[g] = 1.0f / zp[g]; logp[g] += (double)logf(zp[g]); }
      }
      float *tmp = A; A = Bn; Bn = tmp;
      __syncthreads();
    }
    if (tid == 0) {
#pragma unroll
      for (int g = 0; g < G; g++)
        if (b0 + g < p.B) p.logp[b0 + g] = (float)logp[g];
    }

    // ============================================================ backward
    // P: alpha'(t) (TMA-loaded), Q: beta(t+1).  Both forward buffers are free now.
    float *P = buf0, *Q = buf1;
    auto fetch_alpha = [&](float *dst, int f) {
      if (tid == kWarpFetchA * 32) {
        fence_proxy_async();
        mbar_expect_tx(&mbar[0], (uint32_t)R_pad * G * 4u);   // [alpha'(f) | class sums A_c(f)]
        tma_load_1d(dst, aws + (size_t)f * R_pad * G, (uint32_t)R_pad * G * 4u, &mbar[0]);
      }
    };
    // the alpha rows were written with ordinary stores and come back through the async proxy (TMA)
    asm volatile("fence.proxy.async.global;" ::: "memory");
    for (int i = tid; i < S_pad * G; i += DEN_NT) Q[i] = 0.f;
    if (Lmax > 0) issue_x(Lmax - 1);
    __syncthreads();
    // tot(t) of the frame about to be processed: spilled by the forward pass, requested one phase before it is
    // needed (after the beta' pass of the previous frame) so that its L2 latency is not on the frame's critical path
    float tv[G];
#pragma unroll
    for (int g = 0; g < G; g++) tv[g] = 0.f;
    if (kEarlyTot && Lmax > 0) ld_vec<G>(tws + (size_t)(Lmax - 1) * G, tv);
    for (int t = Lmax - 1; t >= 0; t--) {
      // utterances that end at t+1 enter here: beta(L) = final/Z + c * sum_k final_k leaky_k / Z
      bool inject = false;
#pragma unroll
      for (int g = 0; g < G; g++) inject |= (Lg[g] == t + 1);
      // (fetching this row one frame ahead into a third buffer, where one fits, was measured: no gain at cfg5, and
      //  the larger shared-memory footprint cost the co-resident numerator kernel of cfg3 0.4 ms)
      fetch_alpha(P, t);
      if (inject) {
        for (int i = tid; i < S; i += DEN_NT) {
          float v[G];
          ld_vec<G>(Q + (size_t)i * G, v);
          const float f = __ldg(fin + i);
#pragma unroll
          for (int g = 0; g < G; g++)
            if (Lg[g] == t + 1) v[g] = fmaf(c * invZ[g], FL, invZ[g] * f);
          st_vec<G>(Q + (size_t)i * G, v);
        }
      }
      if constexpr (!kEarlyTot) ld_vec<G>(tws + (size_t)t * G, tv);
#pragma unroll
      for (int g = 0; g < G; g++) inv[g] = (t < Lg[g] && tv[g] > 0.f) ? 1.0f / tv[g] : 0.f;   // tv = tot(t)
      transform_x(t, inv);   // x(t,.) / tot(t)
      int kd0 = 0, kp0 = 0;   // this thread's first class: table entries requested before the barrier
      if (kEarlyCls && tid < C) { kd0 = __ldg(p.cls_dst + tid); kp0 = __ldg(p.cls_pdf + tid); }
      mbar_wait(&mbar[0], ph_a); ph_a ^= 1u;
      __syncthreads();  // P, Q, x_s ready; xraw free
      if (t > 0) issue_x(t - 1);
      // bx_c = beta(t+1, dst_c) * x(t, pdf_c) per class, next to beta in the Q row (read by view 4 below)
      for (int k = tid; k < C; k += DEN_NT) {
        float q[G], xv[G];
        const int kd = (kEarlyCls && k == tid) ? kd0 : __ldg(p.cls_dst + k), kp = (kEarlyCls && k == tid) ? kp0 : __ldg(p.cls_pdf + k);
        ld_vec<G>(Q + (size_t)kd * G, q);
        ld_vec<G>(x_s + (size_t)kp * G, xv);
#pragma unroll
        for (int g = 0; g < G; g++) q[g] *= xv[g];
        st_vec<G>(Q + (size_t)(S_pad + k) * G, q);
      }

      // ---- gamma pass: rows = pdfs; gamma(t,n) = x(t,n)/tot(t) * sum alpha'(src) p beta(dst)
      EpiGamma<G> epi_gam{smem_u32(x_s), smem_u32(gam_s)};
      epi_gam = run_pass<G, L1C, EMBED, false, nf2<G>()>(views[2], rec_sa + p.rs_off[2][warp], smem_u32(P), smem_u32(Q), R_pad - 1, S - 1,
                                                          part_sa, nullptr, warp, lane, epi_gam);
      if (p.view[2].n_split) { __syncthreads(); fixup_split<G>(p.view[2], partial, nullptr, epi_gam); }
      prefetch_head<L1C, EMBED>(views[C ? 4 : 1], warp, lane);
      __syncthreads();  // gamma rows complete; everybody done reading P; bx complete

      // ---- beta' pass, class arcs: rows = source states, one operand (bx_c) per arc; the raw sums wait in P
      if (C) {
        EpiStore<G> epi_b2{smem_u32(P)};
        epi_b2 = run_pass<G, L1C, EMBED, true, nf1<G>()>(views[4], rec_sa + p.rs_off[4][warp], smem_u32(Q), 0u, R_pad - 1, 0, part_sa, nullptr, warp, lane, epi_b2);
        if (p.view[4].n_split) { __syncthreads(); fixup_split<G>(p.view[4], partial, nullptr, epi_b2); }
        prefetch_head<L1C, EMBED>(views[1], warp, lane);
        __syncthreads();
      }

      // ---- beta' pass: rows = source states; beta'(t,i) = sum p beta(t+1,dst) x(t,pdf) / tot(t)
      EpiBeta<G> epi_bwd;
      epi_bwd.out_s = smem_u32(P);
#pragma unroll
      for (int g = 0; g < G; g++) epi_bwd.lpart[g] = 0.f;
      epi_bwd = run_pass<G, L1C, EMBED, false, nf2<G>()>(views[1], rec_sa + p.rs_off[1][warp], smem_u32(Q), smem_u32(x_s), S - 1, D - 1,
                                                          part_sa, lk, warp, lane, epi_bwd);
      if (p.view[1].n_split) { __syncthreads(); fixup_split<G>(p.view[1], partial, lk, epi_bwd); }
      prefetch_head<L1C, EMBED>(views[2], warp, lane);
#pragma unroll
      for (int g = 0; g < G; g++) {
        const float ls = warp_sum(epi_bwd.lpart[g]);
        if (lane == 0) red[warp * G + g] = ls;
      }
      // write this frame's gradient rows (coalesced over pdfs) while the other warps finish
      float gsum[G], ab[G];  // the reference's frame-0 self-check sums (chain-computation.cc:363-371)
#pragma unroll
      for (int g = 0; g < G; g++) { gsum[g] = 0.f; ab[g] = 0.f; }
      for (int n = tid; n < D; n += DEN_NT) {  // one vector read per pdf, G coalesced row writes
        float v[G];
        ld_vec<G>(gam_s + (size_t)n * G, v);
#pragma unroll
        for (int g = 0; g < G; g++) {
          if (b0 + g < p.B) p.grad[((size_t)(b0 + g) * T + t) * D + n] = (t < Lg[g]) ? v[g] * p.grad_scale : 0.f;
          gsum[g] += v[g];
        }
      }
      __syncthreads();  // beta' rows + partials visible; everybody done reading Q and gam_s
      if (kEarlyTot && t > 0) ld_vec<G>(tws + (size_t)(t - 1) * G, tv);
      float cb[G];
#pragma unroll
      for (int g = 0; g < G; g++) {
        float s = 0.f;
        for (int w = 0; w < DEN_NW; w++) s += red[w * G + g];
        cb[g] = c * s;
      }
      for (int i = tid; i < S; i += DEN_NT) {
        float v[G];
        ld_vec<G>(P + (size_t)i * G, v);
        if (t == 0) {  // alpha'(0,i) = init_i + c * tot(0) * leaky_i (the same for every utterance) times beta'(0,i)
          const float av = fmaf(c * tot0, __ldg(lk + i), __ldg(ini + i));
#pragma unroll
          for (int g = 0; g < G; g++) ab[g] = fmaf(av, v[g], ab[g]);
        }
#pragma unroll
        for (int g = 0; g < G; g++) v[g] += cb[g];
        st_vec<G>(P + (size_t)i * G, v);
      }
      if (t == 0) {
        block_sum_vec<G>(ab, red2);
        block_sum_vec<G>(gsum, red2);
        if (tid == 0) {
#pragma unroll
          for (int g = 0; g < G; g++)
            if (b0 + g < p.B) { p.diag[2 * (b0 + g)] = ab[g]; p.diag[2 * (b0 + g) + 1] = gsum[g]; }
        }
      }
      float *tmp = P; P = Q; Q = tmp;  // Q <- beta(t); P <- free
      __syncthreads();
    }
    // utterances of length 0 and padded frames t >= Lmax
#pragma unroll
    for (int g = 0; g < G; g++) {
      if (b0 + g >= p.B) continue;
      if (Lg[g] == 0 && tid == 0) { p.diag[2 * (b0 + g)] = 1.f; p.diag[2 * (b0 + g) + 1] = 1.f; }
      float *gr = p.grad + ((size_t)(b0 + g) * T + Lmax) * D;
      const size_t n = (size_t)(T - Lmax) * D;
      for (size_t i = tid; i < n; i += DEN_NT) gr[i] = 0.f;
    }
    __syncthreads();  // Lg / invZ / logp (shared memory) are rewritten at the top of the next group
  }
}


}  // namespace chainb200

using namespace chainb200;

extern "C" {

int chain_b200_den_group_width(int S, int D, int B, int device) {
  if (S <= 0 || D <= 0 || B <= 0) return 0;
  BlobHeader h;
  memset(&h, 0, sizeof(h));
  h.S_pad = (S + 3) / 4 * 4; h.D_pad = (D + 3) / 4 * 4;
  h.R_pad = h.S_pad;  // classes only use the room that is left
  h.max_sub = 256;    // allowance for split rows, unknown before packing
  int num_sms = 148;
  if (cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, device) != cudaSuccess) num_sms = 148;
  return choose_G(h, B, num_sms);
}

size_t chain_b200_den_workspace_bytes(const void *blob_host, int B, int T) {
  const BlobHeader *h = (const BlobHeader *)blob_host;
  if (!h || h->magic != kBlobMagic || h->version != kBlobVersion || B <= 0 || T <= 0) return 0;
  return den_ws_layout(*h, B, T).total;
}

int chain_b200_den_forward_backward(const void *blob_host, const void *blob_dev, int blob_copies, const float *x,
                                    const int64_t *lengths, int B, int T, int D, float coeff, unsigned flags,
                                    float grad_scale, float *objf, float *logp, float *grad, int32_t *ok, float *diag,
                                    void *ws, size_t ws_bytes, int device, void *stream_v) {
  g_launch_count = 0;
  const BlobHeader *hp = (const BlobHeader *)blob_host;
  if (!hp || !blob_dev) return CHAIN_B200_ERR_INVALID_ARG;
  if (hp->magic != kBlobMagic || hp->version != kBlobVersion || hp->nwarps != DEN_NW) return CHAIN_B200_ERR_BAD_BLOB;
  const BlobHeader h = *hp;
  if (B <= 0 || T <= 0 || D != h.D || !x || !lengths || !objf || !grad || !ok) return CHAIN_B200_ERR_INVALID_ARG;
  if (!h.log_domain && !(coeff > 0.f && coeff < 1.f)) return CHAIN_B200_ERR_INVALID_ARG;  // chain-computation.cc:68
  const DenWs w = den_ws_layout(h, B, T);
  if (!ws || ws_bytes < w.total) return CHAIN_B200_ERR_WORKSPACE;

  DeviceGuard guard(device);
  if (guard.err != cudaSuccess) return (int)guard.err;
  int num_sms = 148;
  cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, device);
  const int G = choose_G(h, B, num_sms);
  if (G == 0) return CHAIN_B200_ERR_UNSUPPORTED;
  const size_t smem = den_smem_bytes(h, G);
  cudaStream_t stream = (cudaStream_t)stream_v;
  const char *bd = (const char *)blob_dev;
  char *base = (char *)ws;

  DenParams p;
  for (int i = 0; i < kNumViews; i++) {
    const BlobView &bv = h.view[i];
    p.view[i].warp_info = (const int64_t *)(bd + bv.off_warp_info);
    p.view[i].rows = (const int32_t *)(bd + bv.off_rows);
    p.view[i].recs = (const uint2 *)(bd + bv.off_recs);
    p.view[i].split = (const int32_t *)(bd + bv.off_split);
    p.view[i].n_split = bv.n_split;
  }
  p.cls_dst = (const int32_t *)(bd + h.off_cls_dst);
  p.cls_pdf = (const int32_t *)(bd + h.off_cls_pdf);
  p.C = h.C; p.C_pad = h.C_pad; p.R_pad = h.R_pad;
  p.leaky = (const float *)(bd + h.off_leaky);
  p.init = (const float *)(bd + h.off_init);
  p.fin = (const float *)(bd + h.off_final);
  p.x = x; p.lengths = lengths;
  p.B = B; p.T = T; p.D = D; p.S = h.S; p.S_pad = h.S_pad; p.D_pad = h.D_pad;
  p.coeff = coeff; p.flags = flags; p.grad_scale = grad_scale;
  p.logp = logp ? logp : (float *)(base + w.logp);
  p.grad = grad;
  p.diag = (float *)(base + w.diag);
  p.alpha_ws = (float *)(base + w.alpha);
  p.tot_ws = (float *)(base + w.tot);
  p.num_groups = (B + G - 1) / G;
  p.use_tma_x = (D % 4 == 0) && (((uintptr_t)x & 15u) == 0);
  p.n_rep = blob_copies > 0 ? blob_copies : 1;
  p.rep_stride = (size_t)h.total_bytes;

  const int grid = std::min(p.num_groups, num_sms);
  const bool l1c = records_fit_l1(h, smem) && !getenv("CHAIN_B200_NO_L1_RECORDS");
  // small graphs in the embedded stream format keep their slot streams in shared memory (copied once per CTA)
  p.max_sub = h.max_sub;
  const size_t rec_bytes = resident_record_bytes(blob_host, h, p.rs_off);
  const bool res = h.embed_codes && !h.log_domain && smem + rec_bytes + 16 <= (size_t)227 * 1024 &&
                   !getenv("CHAIN_B200_NO_RESIDENT_RECORDS");
  const size_t smem_launch = smem + (res ? rec_bytes + 16 : 0);
  cudaError_t e;
  if (h.log_domain) {  // same blob layout, log semiring (den_shared_log.cu)
    const int st = launch_den_shared_log(h, p, G, l1c, grid, smem, stream);
    if (st != 0) return st;
    g_launch_count++;
    launch_finalize(p.logp, p.diag, lengths, B, T, 0, objf, ok, diag, stream);
    g_launch_count++;
    return (int)cudaGetLastError();
  }
#define LAUNCH_DEN(GG, LL, EE)                                                                                         \
  do {                                                                                                                \
    e = cudaFuncSetAttribute(den_shared_kernel<GG, LL, EE>, cudaFuncAttributeMaxDynamicSharedMemorySize,             \
                             (int)smem_launch);                                                                       \
    if (e != cudaSuccess) return (int)e;                                                                              \
    den_shared_kernel<GG, LL, EE><<<grid, DEN_NT, smem_launch, stream>>>(p);                                          \
  } while (0)
#define LAUNCH_DEN_G(GG)                                       \
  do {                                                         \
    if (h.embed_codes && res) LAUNCH_DEN(GG, true, true);      \
    else if (h.embed_codes) LAUNCH_DEN(GG, false, true);       \
    else if (l1c) LAUNCH_DEN(GG, true, false);                 \
    else LAUNCH_DEN(GG, false, false);                         \
  } while (0)
  if (G == 4) LAUNCH_DEN_G(4);
  else if (G == 2) LAUNCH_DEN_G(2);
  else LAUNCH_DEN_G(1);
#undef LAUNCH_DEN_G
#undef LAUNCH_DEN
  g_launch_count++;
  launch_finalize(p.logp, p.diag, lengths, B, T, 1, objf, ok, diag, stream);
  g_launch_count++;
  e = cudaGetLastError();
  return (int)e;
}

}  // extern "C"
