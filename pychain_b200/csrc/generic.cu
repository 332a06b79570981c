// generic.cu -- per-sequence-graph forward-backward (both semirings), sm_100a.
//
// One persistent CTA per utterance walks the whole alpha/beta recursion: no
// per-frame host loop, no per-frame launches (the reference issues ~8 launches
// per frame).  Replaces, for arbitrary per-sequence graphs in the
// reference's [B,K,3] layout:
//   ChainComputation            /root/reference/pytorch_binding/src/chain-computation.cc:24-391
//   ChainLogDomainComputation   /root/reference/pytorch_binding/src/chain-log-domain-computation.cc:25-304
//   the 4 kernels               chain-kernels.cu:97-245, chain-log-domain-kernels.cu:86-235
//
// Data layout per CTA (shared memory): three state rows (alpha(t-1)/alpha(t) in
// the forward pass; alpha'(t), beta(t+1), beta(t) in the backward pass) + the
// current frame of (exp) log-probs.  alpha' rows are spilled to the HBM
// workspace [B,T+1,S] for the backward pass.  The arc-posterior gradient is
// accumulated WITHOUT atomics: a third, pdf-sorted view of the arcs is built
// once per call (pdf_sort_kernel) and gamma(t,n) is a segmented sum owned by the
// thread of pdf n, written coalesced and exactly once (the reference uses
// thresholded atomicAdd / atomicCAS log-add per arc, chain-kernels.cu:53-87,
// chain-log-domain-kernels.cu:62-75).
#include <algorithm>
#include <cstdlib>

#include "common.cuh"

namespace chainb200 {

thread_local int g_launch_count = 0;
int g_verbose_level = 0;


struct GenericParams {
  const int32_t *ft, *fi, *bt, *bi;
  const float *fp, *bp, *leaky, *init, *fin;
  const float *x;
  const int64_t *lengths;
  int B, T, D, S, K;
  float coeff;
  unsigned flags;
  float grad_scale;           // posteriors are multiplied by this as they are written
  float *logp, *grad, *diag;  // diag: [B][2]
  float *alpha_ws, *tot_ws;   // [B][T+1][S], [B][T+1]
  const int32_t *pdf_ptr;     // [Bg][D+1]
  const uint2 *pdf_rec;       // [Bg][K]  {src | dst<<16, weight bits}   arcs sorted by pdf
  const uint2 *bwd_rec;       // [Bg][K]  {src | pdf<<16, weight bits}   backward-CSR order (alpha pass)
  const uint2 *fwd_rec;       // [Bg][K]  {dst | pdf<<16, weight bits}   forward-CSR order (beta pass)
  float *rows_ws;             // [B][3S+D] state rows + frame in global memory (graphs too large for shared memory)
};

// ---------------------------------------------------------------- pdf-sorted view
// Stable counting sort of one graph's arcs by pdf id.  grid = Bg, block = 128.
// smem: (D+1) ints.
__global__ void __launch_bounds__(128) pdf_sort_kernel(const int32_t *__restrict__ ft_all,
                                                       const int32_t *__restrict__ fi_all,
                                                       const float *__restrict__ fp_all,
                                                       const int32_t *__restrict__ bt_all,
                                                       const float *__restrict__ bp_all, int S, int K, int D,
                                                       int32_t *__restrict__ pdf_ptr_all,
                                                       uint2 *__restrict__ pdf_rec_all,
                                                       uint2 *__restrict__ bwd_rec_all,
                                                       uint2 *__restrict__ fwd_rec_all) {
  extern __shared__ int cnt[];  // D + 1
  __shared__ int red[4];
  const int g = blockIdx.x;
  const int32_t *ft = ft_all + (size_t)g * K * 3;
  const int32_t *fi = fi_all + (size_t)g * S * 2;
  const float *fp = fp_all + (size_t)g * K;
  int32_t *pdf_ptr = pdf_ptr_all + (size_t)g * (D + 1);
  uint2 *rec = pdf_rec_all + (size_t)g * K;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  // number of real arcs = largest row end (padded arcs of initialized_by_list lie beyond it)
  int kv = 0;
  for (int s = tid; s < S; s += blockDim.x) kv = max(kv, fi[2 * s + 1]);
  for (int o = 16; o > 0; o >>= 1) kv = max(kv, __shfl_xor_sync(0xffffffffu, kv, o));
  if (lane == 0) red[warp] = kv;
  for (int n = tid; n <= D; n += blockDim.x) cnt[n] = 0;
  __syncthreads();
  kv = min(K, max(max(red[0], red[1]), max(red[2], red[3])));
  {  // one 8-byte record per arc and direction (the reference layout needs three loads per arc)
    const int32_t *bt = bt_all + (size_t)g * K * 3;
    const float *bp = bp_all + (size_t)g * K;
    uint2 *brec = bwd_rec_all + (size_t)g * K, *frec = fwd_rec_all + (size_t)g * K;
    for (int k = tid; k < kv; k += blockDim.x) {
      brec[k] = make_uint2(((unsigned)bt[3 * k] & 0xffffu) | ((unsigned)bt[3 * k + 2] << 16), __float_as_uint(bp[k]));
      frec[k] = make_uint2(((unsigned)ft[3 * k + 1] & 0xffffu) | ((unsigned)ft[3 * k + 2] << 16), __float_as_uint(fp[k]));
    }
  }
  for (int k = tid; k < kv; k += blockDim.x) {
    const int pdf = ft[3 * k + 2];
    if (pdf >= 0 && pdf < D) atomicAdd(&cnt[pdf], 1);
  }
  __syncthreads();
  // exclusive scan of cnt[0..D) by warp 0 (chunks of 32 with a running carry)
  if (warp == 0) {
    int carry = 0;
    for (int base = 0; base < D; base += 32) {
      const int n = base + lane;
      const int c = (n < D) ? cnt[n] : 0;
      int incl = c;
      for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
      }
      if (n < D) { cnt[n] = carry + incl - c; pdf_ptr[n] = carry + incl - c; }
      carry += __shfl_sync(0xffffffffu, incl, 31);
    }
    if (lane == 0) { cnt[D] = carry; pdf_ptr[D] = carry; }
    __syncwarp();
    // stable placement: arcs in index order, 32 at a time; equal pdfs inside a
    // chunk keep their order through the match mask.
    for (int base = 0; base < kv; base += 32) {
      const int k = base + lane;
      int pdf = -1;
      if (k < kv) { pdf = ft[3 * k + 2]; if (pdf < 0 || pdf >= D) pdf = -1; }
      const unsigned peers = __match_any_sync(0xffffffffu, pdf);
      if (pdf >= 0) {
        const int rank = __popc(peers & ((1u << lane) - 1u));
        const int pos = cnt[pdf] + rank;
        const unsigned src = (unsigned)ft[3 * k], dst = (unsigned)ft[3 * k + 1];
        rec[pos] = make_uint2((src & 0xffffu) | (dst << 16), __float_as_uint(fp[k]));
      }
      __syncwarp();
      if (pdf >= 0 && (peers & ((1u << lane) - 1u)) == 0) cnt[pdf] += __popc(peers);
      __syncwarp();
    }
  }
}

// ---------------------------------------------------------------- main kernel
__device__ __forceinline__ float load_x(const float *__restrict__ p, unsigned flags, bool log_domain) {
  float v = __ldg(p);
  const bool is_exp = !log_domain && (flags & CHAIN_B200_X_IS_EXP);
  if ((flags & CHAIN_B200_CLAMP_X) && !is_exp) v = clampf(v, kClampLo, kClampHi);
  if (!log_domain && !is_exp) v = expf(v);
  return v;
}

// NT threads per CTA: 128 for small per-utterance graphs (numerators; cheap barriers),
// 1024 for graphs with hundreds of states (one state per thread).
// GROWS: the three state rows and the frame live in global memory (L2) instead of shared memory --
// the slow but complete path for graphs whose 3S+D floats exceed 227 KB (up to the 65535-state
// limit the reference itself has, chain-computation.cc:126-127).
__device__ __forceinline__ void cp_async4(float *smem_dst, const float *gsrc) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc)
               : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

template <bool LOG, int GEN_NT, bool GROWS = false>
__global__ void __launch_bounds__(GEN_NT) chain_generic_kernel(const GenericParams p) {
  extern __shared__ __align__(16) float smem[];
  __shared__ float scratch[64];
  const int S = p.S, D = p.D, T = p.T, K = p.K;
  const int b = blockIdx.x, tid = threadIdx.x;
  const size_t gb = (p.flags & CHAIN_B200_GRAPH_SHARED) ? 0 : (size_t)b;
  const int32_t *__restrict__ fi = p.fi + gb * S * 2;
  const int32_t *__restrict__ bi = p.bi + gb * S * 2;
  const float *__restrict__ lk = LOG ? nullptr : p.leaky + gb * S;
  const float *__restrict__ ini = p.init + gb * S;
  const float *__restrict__ fin = p.fin + gb * S;
  const int32_t *__restrict__ pptr = p.pdf_ptr + gb * (D + 1);
  const uint2 *__restrict__ prec = p.pdf_rec + gb * K;
  const uint2 *__restrict__ brec = p.bwd_rec + gb * K;
  const uint2 *__restrict__ frec = p.fwd_rec + gb * K;
  const float *__restrict__ x = p.x + (size_t)b * T * D;
  float *__restrict__ grad = p.grad + (size_t)b * T * D;
  float *__restrict__ alpha_ws = p.alpha_ws + (size_t)b * (T + 1) * S;
  float *__restrict__ tot_ws = p.tot_ws + (size_t)b * (T + 1);
  long long Lraw = p.lengths[b];
  const int L = (int)(Lraw < 0 ? 0 : (Lraw > T ? T : Lraw));
  const float c = p.coeff;

  float *rows = GROWS ? p.rows_ws + (size_t)blockIdx.x * ((size_t)3 * S + D) : smem;
  float *bufA = rows, *bufB = rows + S, *bufC = rows + 2 * S, *x_s = rows + 3 * S;
  // Log-domain rows in shared memory: the next frame of x (and, in the backward pass, the next
  // alpha row) is fetched by cp.async while the current frame is processed; the clamp is applied
  // when an element is used.  Layout: 4 state rows + 2 frames.
  constexpr bool PF = LOG && !GROWS;
  float *bufD = rows + 3 * S;
  float *xb[2] = {rows + 4 * S, rows + 4 * S + D};
  auto REC = [&](const uint2 *q) { return __ldg(q); };
  const bool clamp_use = PF && (p.flags & CHAIN_B200_CLAMP_X);
  auto XV = [&](int n) { return clamp_use ? clampf(x_s[n], kClampLo, kClampHi) : x_s[n]; };
  auto issue_frame = [&](int f, const float *arow, float *adst) {  // x frame f (+ optionally one alpha row)
    const float *xr = x + (size_t)f * D;
    float *dst = xb[f & 1];
    for (int n = tid; n < D; n += GEN_NT) cp_async4(dst + n, xr + n);
    if (arow)
      for (int i = tid; i < S; i += GEN_NT) cp_async4(adst + i, arow + i);
    cp_async_commit();
  };
  if (PF && L > 0) issue_frame(0, nullptr, nullptr);
  double logp = 0.0;  // only thread 0's copy is used

  // ------------------------------------------------------------ forward, t = 0
  float tot_prev;
  if (!LOG) {
    float local = 0.f;
    for (int i = tid; i < S; i += GEN_NT) { const float a = ini[i]; bufA[i] = a; local += a; }
    const float tot = block_sum(local, scratch);
    const float ct = c * tot;
    for (int i = tid; i < S; i += GEN_NT) {
      const float a = bufA[i] + ct * lk[i];
      bufA[i] = a;
      alpha_ws[i] = a;
    }
    if (tid == 0) { tot_ws[0] = tot; if (0 < L) logp += (double)logf(tot); }
    tot_prev = tot;
  } else {
    for (int i = tid; i < S; i += GEN_NT) { const float a = ini[i]; bufA[i] = a; alpha_ws[i] = a; }
    if (tid == 0) tot_ws[0] = 0.f;
    tot_prev = 0.f;  // the reference stores 0 as the frame-0 alpha-sum (chain-log-domain-computation.cc:89)
  }
  __syncthreads();

  // ------------------------------------------------------------ forward, t = 1..L
  for (int t = 1; t <= L; t++) {
    if (PF) {
      cp_async_wait_all();
      __syncthreads();  // frame t-1 has landed; the other frame buffer is free
      if (t < L) issue_frame(t, nullptr, nullptr);
      x_s = xb[(t - 1) & 1];
    } else {
      const float *xr = x + (size_t)(t - 1) * D;
      for (int n = tid; n < D; n += GEN_NT) x_s[n] = load_x(xr + n, p.flags, LOG);
      __syncthreads();
    }
    float tot;
    if (!LOG) {
      const float inv = 1.0f / tot_prev;
      float local = 0.f;
      for (int h = tid; h < S; h += GEN_NT) {
        float acc = 0.f;
        const int k1 = bi[2 * h + 1];
        for (int k = bi[2 * h]; k < k1; k++) {
          const uint2 r = REC(&brec[k]);
          acc += bufA[r.x & 0xffffu] * __uint_as_float(r.y) * x_s[r.x >> 16];
        }
        const float a = acc * inv;
        bufB[h] = a;
        local += a;
      }
      tot = block_sum(local, scratch);
      const float ct = c * tot;
      float *aw = alpha_ws + (size_t)t * S;
      for (int h = tid; h < S; h += GEN_NT) {
        const float a = bufB[h] + ct * lk[h];
        bufB[h] = a;
        aw[h] = a;
      }
      if (tid == 0) { tot_ws[t] = tot; if (t < L) logp += (double)logf(tot); }
    } else {
      float m_loc = -INFINITY, s_loc = 0.f;
      float *aw = alpha_ws + (size_t)t * S;
      for (int h = tid; h < S; h += GEN_NT) {
        const int k0 = bi[2 * h], k1 = bi[2 * h + 1];
        float m = -INFINITY, sm = 0.f;
        for (int k = k0; k < k1; k++) {
          const uint2 r = REC(&brec[k]);
          lse_add(m, sm, bufA[r.x & 0xffffu] + __uint_as_float(r.y) + XV(r.x >> 16));
        }
        const float a = (m > -INFINITY) ? m + logf(sm) - tot_prev : -INFINITY;
        bufB[h] = a;
        aw[h] = a;
        lse_merge(m_loc, s_loc, a, (a > -INFINITY) ? 1.f : 0.f);
      }
      tot = block_lse(m_loc, s_loc, scratch);
      if (tid == 0) { tot_ws[t] = tot; if (t < L && tot > -INFINITY) logp += (double)tot; }
    }
    tot_prev = tot;
    float *tmp = bufA; bufA = bufB; bufB = tmp;
    __syncthreads();
  }

  // ------------------------------------------------------------ total likelihood, beta(L)
  float inv_Z = 0.f, lZ = 0.f;
  if (!LOG) {
    float local = 0.f;
    for (int i = tid; i < S; i += GEN_NT) local += bufA[i] * fin[i];
    const float Z = block_sum(local, scratch);
    if (tid == 0) logp += (double)logf(Z);
    inv_Z = 1.0f / Z;
    local = 0.f;
    for (int i = tid; i < S; i += GEN_NT) local += (inv_Z * fin[i]) * lk[i];
    const float bs = block_sum(local, scratch);
    for (int i = tid; i < S; i += GEN_NT) bufB[i] = inv_Z * fin[i] + c * bs;
  } else {
    float m_loc = -INFINITY, s_loc = 0.f;
    for (int i = tid; i < S; i += GEN_NT) {
      const float v = bufA[i] + fin[i];
      lse_merge(m_loc, s_loc, v, (v > -INFINITY) ? 1.f : 0.f);
    }
    lZ = block_lse(m_loc, s_loc, scratch);
    if (tid == 0) logp += (double)lZ;
    for (int i = tid; i < S; i += GEN_NT) bufB[i] = fin[i] - lZ;
  }
  if (tid == 0 && p.logp) p.logp[b] = (float)logp;
  __syncthreads();
  if (PF && L > 0) {  // frame L-1 of x is still in xb[(L-1)&1]; fetch alpha(L-1) into bufA
    for (int i = tid; i < S; i += GEN_NT) cp_async4(bufA + i, alpha_ws + (size_t)(L - 1) * S + i);
    cp_async_commit();
  }

  // ------------------------------------------------------------ backward, t = L-1..0
  for (int t = L - 1; t >= 0; t--) {
    const float tot_t = tot_ws[t];
    if (PF) {
      cp_async_wait_all();
      __syncthreads();  // x(t) and alpha(t) have landed; bufD and the other frame buffer are free
      if (t > 0) issue_frame(t - 1, alpha_ws + (size_t)(t - 1) * S, bufD);
      x_s = xb[t & 1];
    } else {
      const float *xr = x + (size_t)t * D;
      const float *aw = alpha_ws + (size_t)t * S;
      for (int n = tid; n < D; n += GEN_NT) x_s[n] = load_x(xr + n, p.flags, LOG);
      for (int i = tid; i < S; i += GEN_NT) bufA[i] = aw[i];
      __syncthreads();
    }
    float *gr = grad + (size_t)t * D;
    float gsum = 0.f;
    if (!LOG) {
      const float inv = 1.0f / tot_t;
      // gamma pass: pdf-sorted arcs, thread owns pdf n
      for (int n = tid; n < D; n += GEN_NT) {
        float g = 0.f;
        const int j1 = pptr[n + 1];
        for (int j = pptr[n]; j < j1; j++) {
          const uint2 r = REC(&prec[j]);
          g += bufA[r.x & 0xffffu] * __uint_as_float(r.y) * bufB[r.x >> 16];
        }
        const float gam = g * x_s[n] * inv;
        gr[n] = gam * p.grad_scale;
        gsum += gam;
      }
      // beta' pass: thread owns source state h
      float local = 0.f;
      for (int h = tid; h < S; h += GEN_NT) {
        float acc = 0.f;
        const int k1 = fi[2 * h + 1];
        for (int k = fi[2 * h]; k < k1; k++) {
          const uint2 r = REC(&frec[k]);
          acc += __uint_as_float(r.y) * bufB[r.x & 0xffffu] * x_s[r.x >> 16];
        }
        const float bd = acc * inv;
        bufC[h] = bd;
        local += bd * lk[h];
      }
      const float bs = block_sum(local, scratch);
      if (t == 0 && p.diag) {
        float ab = 0.f;
        for (int i = tid; i < S; i += GEN_NT) ab += bufA[i] * bufC[i];
        ab = block_sum(ab, scratch);
        gsum = block_sum(gsum, scratch);
        if (tid == 0) { p.diag[2 * b] = ab; p.diag[2 * b + 1] = gsum; }
      }
      const float cb = c * bs;
      for (int h = tid; h < S; h += GEN_NT) bufC[h] += cb;
    } else {
      for (int n = tid; n < D; n += GEN_NT) {
        const int j0 = pptr[n], j1 = pptr[n + 1];
        const float xn = XV(n) - tot_t;
        float m = -INFINITY, sm = 0.f;
        for (int j = j0; j < j1; j++) {
          const uint2 r = REC(&prec[j]);
          lse_add(m, sm, bufA[r.x & 0xffffu] + __uint_as_float(r.y) + bufB[r.x >> 16] + xn);
        }
        float lg = -INFINITY, gam = 0.f;
        if (m > -INFINITY) {
          lg = m + logf(sm);
          gam = expf(lg);
        }
        gr[n] = (p.flags & CHAIN_B200_GRAD_AS_LOG) ? lg : gam * p.grad_scale;
        gsum += gam;
      }
      for (int h = tid; h < S; h += GEN_NT) {
        const int k0 = fi[2 * h], k1 = fi[2 * h + 1];
        float m = -INFINITY, sm = 0.f;
        for (int k = k0; k < k1; k++) {
          const uint2 r = REC(&frec[k]);
          lse_add(m, sm, __uint_as_float(r.y) + bufB[r.x & 0xffffu] + XV(r.x >> 16) - tot_t);
        }
        const float v = (m > -INFINITY) ? m + logf(sm) : -INFINITY;
        bufC[h] = v;
      }
      if (t == 0 && p.diag) {
        gsum = block_sum(gsum, scratch);
        if (tid == 0) { p.diag[2 * b] = 1.f; p.diag[2 * b + 1] = gsum; }
      }
    }
    float *tmp = bufB; bufB = bufC; bufC = tmp;
    if (PF) { tmp = bufA; bufA = bufD; bufD = tmp; }  // the prefetched alpha(t-1)
    __syncthreads();
  }
  if (L == 0 && p.diag && tid == 0) { p.diag[2 * b] = 1.f; p.diag[2 * b + 1] = 1.f; }

  // ------------------------------------------------------------ padded frames
  {
    const float fill = (LOG && (p.flags & CHAIN_B200_GRAD_AS_LOG)) ? -INFINITY : 0.f;
    float *g0 = grad + (size_t)L * D;
    const size_t n = (size_t)(T - L) * D;
    for (size_t i = tid; i < n; i += GEN_NT) g0[i] = fill;
  }
}

// ---------------------------------------------------------------- finalize
// objf = sum_b logp[b] (fixed order, double); ok = the reference's t=0 self-check
// (chain-computation.cc:345-391 / chain-log-domain-computation.cc:283-304).
// diag_out (8 floats): {alpha.beta sum, gamma sum, objf, valid frames, not ok, 0, 0, 0}; elements 2..4 are the
// three numbers the batch-sharded loss all-reduces (the reference's only cross-utterance operations:
// chain-computation.cc:229 and loss.py:104), ready in one contiguous buffer with no further launch.
__global__ void __launch_bounds__(256) finalize_kernel(const float *__restrict__ logp, const float *__restrict__ diag,
                                                       const int64_t *__restrict__ lengths, int B, int T,
                                                       int check_ab, float *objf, int32_t *ok, float *diag_out) {
  __shared__ double sh[4][256];
  const int tid = threadIdx.x;
  double a = 0.0, ab = 0.0, gs = 0.0, fr = 0.0;
  for (int b = tid; b < B; b += 256) {
    a += (double)logp[b];
    ab += (double)diag[2 * b];
    gs += (double)diag[2 * b + 1];
    const long long l = lengths ? lengths[b] : 0;
    fr += (double)(l < 0 ? 0 : (l > T ? T : l));
  }
  sh[0][tid] = a; sh[1][tid] = ab; sh[2][tid] = gs; sh[3][tid] = fr;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (tid < o) {
      sh[0][tid] += sh[0][tid + o]; sh[1][tid] += sh[1][tid + o]; sh[2][tid] += sh[2][tid + o];
      sh[3][tid] += sh[3][tid + o];
    }
    __syncthreads();
  }
  if (tid == 0) {
    const float obj = (float)sh[0][0], abp = (float)sh[1][0], gsm = (float)sh[2][0], bsz = (float)B;
    *objf = obj;
    int good = 1;
    // ApproxEqual(a,b,tol) is |a-b| < a*tol (base.cc:6-10); only an excess of 5% clears `ok`
    if (check_ab && !(fabsf(abp - bsz) < abp * 0.01f) && fabsf(abp - bsz) > 0.05f * bsz) good = 0;
    if (!(fabsf(gsm - bsz) < gsm * 0.01f) && (fabsf(gsm - bsz) > 0.05f * bsz || gsm - gsm != 0.f)) good = 0;
    if (!(abp - abp == 0.f) || !(obj - obj == 0.f)) good = 0;  // NaN / inf anywhere
    *ok = good;
    if (diag_out) {
      diag_out[0] = abp; diag_out[1] = gsm; diag_out[2] = obj; diag_out[3] = (float)sh[3][0];
      diag_out[4] = good ? 0.f : 1.f; diag_out[5] = diag_out[6] = diag_out[7] = 0.f;
    }
  }
}

void launch_finalize(const float *logp, const float *diag, const int64_t *lengths, int B, int T, int check_ab,
                     float *objf, int32_t *ok, float *diag_out, cudaStream_t stream) {
  finalize_kernel<<<1, 256, 0, stream>>>(logp, diag, lengths, B, T, check_ab, objf, ok, diag_out);
}

// ---------------------------------------------------------------- the tail of pychain/loss.py, fused
// loss = (objf_den - objf_num) * loss_scale (loss.py:102-104) and grad_den += grad_num (both already carry their
// +-1/sum(lengths) factor from the kernels' write-out): one launch instead of neg / sub / div / add passes.
__global__ void __launch_bounds__(256) combine_loss_kernel(const float *__restrict__ objf_num,
                                                           const float *__restrict__ objf_den, float loss_scale,
                                                           float *__restrict__ gd, const float *__restrict__ gn,
                                                           size_t n, size_t nvec, float *__restrict__ loss) {
  if (blockIdx.x == 0 && threadIdx.x == 0) *loss = (*objf_den - *objf_num) * loss_scale;
  if (!gn) return;
  const size_t stride = (size_t)gridDim.x * blockDim.x, i0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  float4 *gd4 = reinterpret_cast<float4 *>(gd);
  const float4 *gn4 = reinterpret_cast<const float4 *>(gn);
  for (size_t i = i0; i < nvec; i += stride) {
    float4 a = gd4[i];
    const float4 b = __ldg(gn4 + i);
    a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
    gd4[i] = a;
  }
  for (size_t i = 4 * nvec + i0; i < n; i += stride) gd[i] += gn[i];
}

// autograd's `input_grad * objf_grad` (loss.py:85).  A plain loss.backward() hands in 1: nothing is touched then.
__global__ void __launch_bounds__(256) scale_grad_kernel(float *__restrict__ g, size_t n, size_t nvec,
                                                         const float *__restrict__ upstream) {
  const float u = __ldg(upstream);
  if (u == 1.0f) return;
  const size_t stride = (size_t)gridDim.x * blockDim.x, i0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  float4 *g4 = reinterpret_cast<float4 *>(g);
  for (size_t i = i0; i < nvec; i += stride) {
    float4 a = g4[i];
    a.x *= u; a.y *= u; a.z *= u; a.w *= u;
    g4[i] = a;
  }
  for (size_t i = 4 * nvec + i0; i < n; i += stride) g[i] *= u;
}

// ---------------------------------------------------------------- host side
struct GenericWs {
  size_t alpha, tot, pdf_ptr, pdf_rec, bwd_rec, fwd_rec, logp, diag, rows, total;
};

static size_t generic_smem_bytes(int S, int D, bool log_domain) {
  return sizeof(float) * (log_domain ? (size_t)4 * S + 2 * (size_t)D : (size_t)3 * S + D);
}
static bool generic_rows_in_global(int S, int D, bool log_domain = true) {  // default: the larger (log) footprint
  return generic_smem_bytes(S, D, log_domain) > 227 * 1024 - 1024;
}

static GenericWs generic_ws_layout(int B, int T, int D, int S, int K, unsigned flags) {
  GenericWs w;
  const size_t Bg = (flags & CHAIN_B200_GRAPH_SHARED) ? 1 : (size_t)B;
  size_t off = 0;
  w.alpha = off; off = align_up(off + sizeof(float) * (size_t)B * (T + 1) * S);
  w.tot = off; off = align_up(off + sizeof(float) * (size_t)B * (T + 1));
  w.pdf_ptr = off; off = align_up(off + sizeof(int32_t) * Bg * (D + 1));
  w.pdf_rec = off; off = align_up(off + sizeof(uint2) * Bg * (size_t)(K > 0 ? K : 1));
  w.bwd_rec = off; off = align_up(off + sizeof(uint2) * Bg * (size_t)(K > 0 ? K : 1));
  w.fwd_rec = off; off = align_up(off + sizeof(uint2) * Bg * (size_t)(K > 0 ? K : 1));
  w.logp = off; off = align_up(off + sizeof(float) * (size_t)B);
  w.diag = off; off = align_up(off + sizeof(float) * 2 * (size_t)B);
  w.rows = off;
  if (generic_rows_in_global(S, D)) off = align_up(off + sizeof(float) * (size_t)B * ((size_t)3 * S + D));
  w.total = off;
  return w;
}

template <bool LOG>
static int run_generic(const int32_t *ft, const int32_t *fi, const float *fp, const int32_t *bt, const int32_t *bi,
                       const float *bp, const float *leaky, const float *init, const float *fin, const float *x,
                       const int64_t *lengths, int B, int T, int D, int S, int K, float coeff, unsigned flags,
                       float grad_scale, float *objf, float *logp, float *grad, int32_t *ok, float *diag, void *ws,
                       size_t ws_bytes, int device, void *stream_v) {
  g_launch_count = 0;
  if (B <= 0 || T <= 0 || D <= 0 || S <= 0 || K < 0) return CHAIN_B200_ERR_INVALID_ARG;
  if (!ft || !fi || !fp || !bt || !bi || !bp || !init || !fin || !x || !lengths || !objf || !grad || !ok)
    return CHAIN_B200_ERR_INVALID_ARG;
  if (!LOG && !leaky) return CHAIN_B200_ERR_INVALID_ARG;
  if (!LOG && !(coeff > 0.f && coeff < 1.f)) return CHAIN_B200_ERR_INVALID_ARG;  // chain-computation.cc:68
  if (S > 65535 || D > 65535) return CHAIN_B200_ERR_UNSUPPORTED;
  const GenericWs w = generic_ws_layout(B, T, D, S, K, flags);
  if (!ws || ws_bytes < w.total) return CHAIN_B200_ERR_WORKSPACE;
  const bool grows = generic_rows_in_global(S, D, LOG);
  // log domain: 4 state rows + 2 frames (cp.async prefetch); probability domain: 3 rows + 1 frame
  const size_t smem = grows ? 0 : generic_smem_bytes(S, D, LOG);
  const size_t sort_smem = sizeof(int) * (size_t)(D + 1);
  if (sort_smem > 227 * 1024 - 1024) return CHAIN_B200_ERR_UNSUPPORTED;

  DeviceGuard guard(device);
  if (guard.err != cudaSuccess) return (int)guard.err;
  cudaStream_t stream = (cudaStream_t)stream_v;
  char *base = (char *)ws;
  const int Bg = (flags & CHAIN_B200_GRAPH_SHARED) ? 1 : B;

  cudaError_t e;
  if (sort_smem > 48 * 1024) {
    e = cudaFuncSetAttribute(pdf_sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sort_smem);
    if (e != cudaSuccess) return (int)e;
  }
  pdf_sort_kernel<<<Bg, 128, sort_smem, stream>>>(ft, fi, fp, bt, bp, S, K, D, (int32_t *)(base + w.pdf_ptr),
                                                 (uint2 *)(base + w.pdf_rec), (uint2 *)(base + w.bwd_rec),
                                                 (uint2 *)(base + w.fwd_rec));
  g_launch_count++;

  GenericParams p;
  p.ft = ft; p.fi = fi; p.bt = bt; p.bi = bi; p.fp = fp; p.bp = bp;
  p.leaky = leaky; p.init = init; p.fin = fin; p.x = x; p.lengths = lengths;
  p.grad_scale = grad_scale;
  p.B = B; p.T = T; p.D = D; p.S = S; p.K = K; p.coeff = coeff; p.flags = flags;
  p.logp = logp ? logp : (float *)(base + w.logp);
  p.grad = grad;
  p.diag = (float *)(base + w.diag);
  p.alpha_ws = (float *)(base + w.alpha);
  p.tot_ws = (float *)(base + w.tot);
  p.pdf_ptr = (const int32_t *)(base + w.pdf_ptr);
  p.pdf_rec = (const uint2 *)(base + w.pdf_rec);
  p.bwd_rec = (const uint2 *)(base + w.bwd_rec);
  p.fwd_rec = (const uint2 *)(base + w.fwd_rec);
  p.rows_ws = (float *)(base + w.rows);

  const int nt = grows ? 1024 : ((S <= 160) ? 128 : (S <= 448 ? 256 : 1024));
  const size_t smem_launch = smem;
  void (*kern)(const GenericParams) = chain_generic_kernel<LOG, 256>;
  if (nt == 128) kern = chain_generic_kernel<LOG, 128>;
  if (nt == 1024) kern = chain_generic_kernel<LOG, 1024>;
  if (grows) kern = chain_generic_kernel<LOG, 1024, true>;
  if (smem_launch > 48 * 1024) {
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_launch);
    if (e != cudaSuccess) return (int)e;
  }
  kern<<<B, nt, smem_launch, stream>>>(p);
  g_launch_count++;
  launch_finalize(p.logp, p.diag, p.lengths, B, T, LOG ? 0 : 1, objf, ok, diag, stream);
  g_launch_count++;
  e = cudaGetLastError();
  return (int)e;
}

}  // namespace chainb200

using namespace chainb200;

extern "C" {

int chain_b200_version(void) { return 100; }

const char *chain_b200_status_string(int status) {
  switch (status) {
    case CHAIN_B200_OK: return "ok";
    case CHAIN_B200_ERR_INVALID_ARG: return "invalid argument";
    case CHAIN_B200_ERR_WORKSPACE: return "workspace missing or too small";
    case CHAIN_B200_ERR_UNSUPPORTED: return "graph too large (S, D <= 65535; log-domain packed graphs S, D <= 32767; D <= 57000)";
    case CHAIN_B200_ERR_BAD_BLOB: return "not a chain_b200 prepared-graph blob";
    default: return status > 0 ? cudaGetErrorString((cudaError_t)status) : "unknown error";
  }
}

void chain_b200_set_verbose_level(int level) { g_verbose_level = level; }
int chain_b200_get_verbose_level(void) { return g_verbose_level; }
int chain_b200_last_launch_count(void) { return g_launch_count; }

size_t chain_b200_workspace_bytes(int B, int T, int D, int S, int K, unsigned flags) {
  if (B <= 0 || T <= 0 || D <= 0 || S <= 0 || K < 0) return 0;
  return generic_ws_layout(B, T, D, S, K, flags).total;
}

int chain_b200_forward_backward(const int32_t *ft, const int32_t *fi, const float *fp, const int32_t *bt,
                                const int32_t *bi, const float *bp, const float *leaky, const float *init,
                                const float *fin, const float *x, const int64_t *lengths, int B, int T, int D, int S,
                                int K, float coeff, unsigned flags, float grad_scale, float *objf, float *logp,
                                float *grad, int32_t *ok, float *diag, void *ws, size_t ws_bytes, int device,
                                void *stream) {
  return run_generic<false>(ft, fi, fp, bt, bi, bp, leaky, init, fin, x, lengths, B, T, D, S, K, coeff, flags,
                            grad_scale, objf, logp, grad, ok, diag, ws, ws_bytes, device, stream);
}

int chain_b200_forward_backward_log_domain(const int32_t *ft, const int32_t *fi, const float *fp, const int32_t *bt,
                                           const int32_t *bi, const float *bp, const float *init, const float *fin,
                                           const float *x, const int64_t *lengths, int B, int T, int D, int S, int K,
                                           unsigned flags, float grad_scale, float *objf, float *logp, float *grad,
                                           int32_t *ok, float *diag, void *ws, size_t ws_bytes, int device,
                                           void *stream) {
  return run_generic<true>(ft, fi, fp, bt, bi, bp, nullptr, init, fin, x, lengths, B, T, D, S, K, 0.f,
                           flags & ~CHAIN_B200_X_IS_EXP, grad_scale, objf, logp, grad, ok, diag, ws, ws_bytes, device,
                           stream);
}

int chain_b200_combine_loss(const float *objf_num, const float *objf_den, float loss_scale, float *grad_den,
                            const float *grad_num, size_t n, float *loss, int device, void *stream_v) {
  g_launch_count = 0;
  if (!objf_num || !objf_den || !loss || (n > 0 && !grad_den)) return CHAIN_B200_ERR_INVALID_ARG;
  DeviceGuard guard(device);
  if (guard.err != cudaSuccess) return (int)guard.err;
  const size_t vec = (grad_num && ((uintptr_t)grad_den & 15u) == 0 && ((uintptr_t)grad_num & 15u) == 0) ? n / 4 : 0;
  const size_t work = grad_num ? (vec ? vec + (n - 4 * vec) : n) : 1;
  const int grid = (int)std::min<size_t>((work + 255) / 256, (size_t)148 * 8);
  combine_loss_kernel<<<std::max(grid, 1), 256, 0, (cudaStream_t)stream_v>>>(objf_num, objf_den, loss_scale, grad_den,
                                                                              grad_num, n, vec, loss);
  g_launch_count++;
  return (int)cudaGetLastError();
}

int chain_b200_scale_grad(float *grad, size_t n, const float *upstream, int device, void *stream_v) {
  g_launch_count = 0;
  if (!grad || !upstream) return CHAIN_B200_ERR_INVALID_ARG;
  DeviceGuard guard(device);
  if (guard.err != cudaSuccess) return (int)guard.err;
  const size_t vec = (((uintptr_t)grad & 15u) == 0) ? n / 4 : 0;
  const int grid = (int)std::min<size_t>((std::max<size_t>(vec, n - 4 * vec) + 255) / 256, (size_t)148 * 8);
  scale_grad_kernel<<<std::max(grid, 1), 256, 0, (cudaStream_t)stream_v>>>(grad, n, vec, upstream);
  g_launch_count++;
  return (int)cudaGetLastError();
}

}  // extern "C"
