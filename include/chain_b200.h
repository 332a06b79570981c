/* chain_b200.h -- C ABI of the B200-native LF-MMI chain-loss engine.
 *
 * This is the drop-in boundary for the one hot path of pychain: the batched
 * alpha/beta forward-backward over numerator / denominator FSTs and the
 * arc-posterior gradient.  The entry points are what a binding of the
 * reference's native module would call; each one names the reference
 * interface it replaces (paths relative to /root/reference):
 *
 *   chain_b200_forward_backward             <- pychain_C.forward_backward
 *                                              (pytorch_binding/src/pychain.cc:26-79, m.def :132;
 *                                               call site pychain/loss.py:44-60)
 *   chain_b200_forward_backward_log_domain  <- pychain_C.forward_backward_log_domain
 *                                              (pychain.cc:81-129, m.def :133; call site loss.py:62-76)
 *   chain_b200_set_verbose_level            <- pychain_C.set_verbose_level
 *                                              (pychain.cc:134, base.h:40-42)
 *   chain_b200_den_*                        <- the same forward_backward, for the case the reference
 *                                              expresses as ChainGraphBatch(den_graph, B)
 *                                              (pychain/loss.py:99, pychain/graph.py:101-120): ONE graph
 *                                              shared by all B sequences.  The graph is packed once
 *                                              (chain_b200_den_prepare) instead of being replicated B times
 *                                              on the host and uploaded every call
 *                                              (chain-computation.cc:77-89).
 *
 * Conventions
 *   - plain C: pointers and sizes only, no torch / C++ types.  Every `*_dev`
 *     pointer is device memory on CUDA device `device`; every `*_host` pointer
 *     is host memory.  The caller owns every buffer (inputs, outputs,
 *     workspace); the library retains nothing after a call returns and keeps
 *     no global state except the verbosity level.
 *   - all calls are asynchronous on `stream` (a cudaStream_t passed as void*):
 *     no host synchronisation, `objf` / `ok` stay on the device.
 *   - return value: 0 = success; < 0 = CHAIN_B200_ERR_*; > 0 = a cudaError_t.
 *     Nothing is thrown across the boundary.
 *   - there is NO CPU implementation behind these entry points.
 *
 * Tensor layouts are the reference's:
 *   transitions  int32 [Bg,K,3]  rows (src, dst, pdf); forward_* grouped by src, backward_* by dst
 *   indices      int32 [Bg,S,2]  (begin,end) row range per state
 *   probs        f32   [Bg,K]    arc probability (linear) or log-probability (log domain)
 *   leaky/init/final f32 [Bg,S]
 *   x            f32   [B,T,D]   row-major, contiguous
 *   lengths      int64 [B]       valid frames per sequence, 0 <= len <= T
 * with Bg = 1 when CHAIN_B200_GRAPH_SHARED is set (batch stride 0), else Bg = B.
 * Limits: S <= 65535 (the reference's own grid.y limit, chain-computation.cc:126-127) and
 * D <= 57000.  The general entry points keep their state rows in shared memory when
 * 3*S + D floats fit in 227 KB and in global memory (slower) otherwise; the packed
 * shared-graph path needs 2*S + 3*D floats on chip (S up to ~27000 at D = 1000).
 */
#ifndef CHAIN_B200_H_
#define CHAIN_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CHAIN_B200_OK                 0
#define CHAIN_B200_ERR_INVALID_ARG   (-1)
#define CHAIN_B200_ERR_WORKSPACE     (-2) /* workspace pointer NULL or too small */
#define CHAIN_B200_ERR_UNSUPPORTED   (-3) /* graph too large for the on-chip working set */
#define CHAIN_B200_ERR_BAD_BLOB      (-4) /* prepared-graph blob is not one of ours */

/* flags */
#define CHAIN_B200_GRAPH_SHARED   1u  /* graph tensors have batch stride 0 (one graph for all sequences) */
#define CHAIN_B200_X_IS_EXP       2u  /* prob domain: x already holds exp(nnet_output), as loss.py:43 passes it */
#define CHAIN_B200_CLAMP_X        4u  /* apply clamp(x,-30,30) (loss.py:30) inside the kernel (x is raw nnet output) */
#define CHAIN_B200_GRAD_AS_LOG    8u  /* log domain: write log-posteriors (-inf where empty) like the reference;
                                         otherwise write posteriors, i.e. the caller's `.exp()` (loss.py:77) is fused */
#define CHAIN_B200_PACK_G(g)     ((unsigned)(g) << 8) /* chain_b200_den_prepare*: optimise the record order for groups of
                                         g = 1, 2 or 4 utterances (default 4); see chain_b200_den_group_width */
#define CHAIN_B200_LOG_DOMAIN    16u  /* chain_b200_den_prepare*: the graph is a log-domain graph (weights, init, final
                                         are log-probabilities, no leaky HMM; S, D <= 32767) */

int chain_b200_version(void);
const char *chain_b200_status_string(int status);
void chain_b200_set_verbose_level(int level);
int chain_b200_get_verbose_level(void);

/* Bytes of device workspace needed by the two general entry points below. */
size_t chain_b200_workspace_bytes(int B, int T, int D, int S, int K, unsigned flags);

/* Denominator-style forward-backward: probability domain with leaky HMM.
 * objf_dev   f32 [1]      sum over sequences of log p(x_b)
 * logp_dev   f32 [B]      per-sequence log p (may be NULL)
 * grad_dev   f32 [B,T,D]  d objf / d log-prob = arc posteriors summed per pdf; rows t >= len are 0.
 *                         Every element is written (no pre-zeroing needed).
 *                         The posteriors are multiplied by `grad_scale` as they are written: pass 1 for the
 *                         reference's output; ChainLoss passes +-1/sum(lengths) (loss.py:85,102-104), so that
 *                         d loss/d x leaves the kernel finished (ignored with CHAIN_B200_GRAD_AS_LOG).
 * ok_dev     i32 [1]      the reference's self-check verdict (chain-computation.cc:345-391), 1 = ok
 * diag_dev   f32 [8]      {sum_b sum_i alpha'(0,i) beta'(0,i), sum_b sum_n gamma(0,n), objf, valid frames
 *                         (sum_b min(len_b, T)), 1 - ok, 0, 0, 0} (may be NULL).  Elements 2..4 are the numbers a
 *                         batch-sharded job sums over ranks (chain-computation.cc:229, loss.py:104): one
 *                         contiguous buffer for ONE all-reduce, no further launch. */
int chain_b200_forward_backward(
    const int32_t *fwd_trans_dev, const int32_t *fwd_idx_dev, const float *fwd_probs_dev,
    const int32_t *bwd_trans_dev, const int32_t *bwd_idx_dev, const float *bwd_probs_dev,
    const float *leaky_dev, const float *init_dev, const float *final_dev,
    const float *x_dev, const int64_t *lengths_dev,
    int B, int T, int D, int S, int K, float leaky_coeff, unsigned flags, float grad_scale,
    float *objf_dev, float *logp_dev, float *grad_dev, int32_t *ok_dev, float *diag_dev,
    void *workspace_dev, size_t workspace_bytes, int device, void *stream);

/* Numerator-style forward-backward: log domain, no leaky HMM.
 * grad_dev holds log-posteriors when CHAIN_B200_GRAD_AS_LOG is set (rows t >= len are -inf),
 * posteriors otherwise (rows t >= len are 0).  diag_dev f32 [8] as above with element 0 unused (= 0). */
int chain_b200_forward_backward_log_domain(
    const int32_t *fwd_trans_dev, const int32_t *fwd_idx_dev, const float *fwd_probs_dev,
    const int32_t *bwd_trans_dev, const int32_t *bwd_idx_dev, const float *bwd_probs_dev,
    const float *init_dev, const float *final_dev,
    const float *x_dev, const int64_t *lengths_dev,
    int B, int T, int D, int S, int K, unsigned flags, float grad_scale,
    float *objf_dev, float *logp_dev, float *grad_dev, int32_t *ok_dev, float *diag_dev,
    void *workspace_dev, size_t workspace_bytes, int device, void *stream);

/* ---- shared denominator graph: pack once, run many times -------------------
 * chain_b200_den_prepare_bytes: size of the packed blob for this graph (host CSR
 *   arrays in the layout above with Bg = 1), or < 0 on error.  flags: 0 for a
 *   probability-domain graph with leaky HMM, CHAIN_B200_LOG_DOMAIN for a log-domain
 *   graph (leaky_host may be NULL); the same flags must be given to _prepare.
 *   A log-domain blob makes chain_b200_den_forward_backward compute what
 *   chain_b200_forward_backward_log_domain computes (leaky_coeff ignored,
 *   CHAIN_B200_GRAD_AS_LOG honoured).
 * chain_b200_den_prepare: fill `blob_host` (position independent; the caller
 *   uploads it verbatim to the device, 256-byte aligned).
 * chain_b200_den_workspace_bytes: device workspace for a batch of B x T.
 * chain_b200_den_forward_backward: same outputs as chain_b200_forward_backward;
 *   needs both the host copy of the blob (read for launch geometry only) and
 *   the device copy.  `blob_dev` may hold `blob_copies` >= 1 identical copies of
 *   the blob back to back (stride = the blob's byte size, which is a multiple of
 *   256): every CTA streams the same records every frame, and spreading the CTAs
 *   over several copies (8 is a good number) avoids L2 hot-spotting. */
int64_t chain_b200_den_prepare_bytes(
    const int32_t *fwd_trans_host, const int32_t *fwd_idx_host,
    const int32_t *bwd_trans_host, const int32_t *bwd_idx_host,
    int S, int K, int D, unsigned flags);

int chain_b200_den_prepare(
    const int32_t *fwd_trans_host, const int32_t *fwd_idx_host, const float *fwd_probs_host,
    const int32_t *bwd_trans_host, const int32_t *bwd_idx_host, const float *bwd_probs_host,
    const float *leaky_host, const float *init_host, const float *final_host,
    int S, int K, int D, unsigned flags, void *blob_host, int64_t blob_bytes);

size_t chain_b200_den_workspace_bytes(const void *blob_host, int B, int T);

/* Utterances per CTA (1, 2 or 4) the shared-graph kernels will use for a batch of B on `device`
 * (0 if the graph does not fit on chip).  Packing the blob with CHAIN_B200_PACK_G(that value) gives the
 * record order whose shared-memory bank conflicts are minimised for that group width. */
int chain_b200_den_group_width(int S, int D, int B, int device);

int chain_b200_den_forward_backward(
    const void *blob_host, const void *blob_dev, int blob_copies,
    const float *x_dev, const int64_t *lengths_dev,
    int B, int T, int D, float leaky_coeff, unsigned flags, float grad_scale,
    float *objf_dev, float *logp_dev, float *grad_dev, int32_t *ok_dev, float *diag_dev,
    void *workspace_dev, size_t workspace_bytes, int device, void *stream);

/* ---- the rest of pychain/loss.py:97-105, fused ------------------------------
 * chain_b200_combine_loss: grad_den[i] += grad_num[i] for i < n (both already scaled, see grad_scale) and
 *   loss[0] = (objf_den[0] - objf_num[0]) * loss_scale, i.e. -(num - den) [/ sum(lengths)] (loss.py:102-104): one
 *   launch instead of the neg / sub / div / add elementwise passes.  grad_num may be NULL (loss only).
 * chain_b200_scale_grad: grad[i] *= upstream[0] -- autograd's `grad * objf_grad` (loss.py:85); the kernel returns
 *   without touching memory when upstream[0] == 1 (a plain loss.backward()). */
int chain_b200_combine_loss(const float *objf_num_dev, const float *objf_den_dev, float loss_scale,
                            float *grad_den_dev, const float *grad_num_dev, size_t n, float *loss_dev,
                            int device, void *stream);
int chain_b200_scale_grad(float *grad_dev, size_t n, const float *upstream_dev, int device, void *stream);

/* Number of kernel launches the last call on this thread issued (for bench.py's
 * `gpu_launches` accounting). */
int chain_b200_last_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* CHAIN_B200_H_ */
