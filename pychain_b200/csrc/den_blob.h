// den_blob.h -- layout of a packed shared denominator graph ("blob").
//
// The blob is position independent (byte offsets from its start), built on the host by
// chain_b200_den_prepare (den_pack.cu) and uploaded verbatim.  It holds three sliced-ELL
// views of the same arc set; in view v every ROW accumulates  sum_k w_k * A[a_k] * B[b_k]:
//
//   view 0 (alpha pass)  row = destination state, a = source state (alpha'), b = pdf (x)
//   view 1 (beta pass)   row = source state,      a = destination  (beta),   b = pdf (x)
//   view 2 (gamma pass)  row = pdf,               a = source (alpha'),       b = destination (beta)
//
// 32 rows form a slice (one row per lane); slices are dealt to the DEN_NW warps of a CTA and
// each warp's records are stored as ONE contiguous stream in processing order
// ([slice][k][lane], 8 bytes each: {a:16 | b:16, weight:f32}).  Records are consumed in
// pairs; the SIGN BIT of the weight of the second record of a pair marks "last pair of this
// slice" (weights are probabilities, >= 0), so the kernel's inner loop is one flat counted
// loop over pairs with a register FIFO of kFifoPairs pairs of loads in flight -- no per-slice
// control flow, no bounds checks (each stream is padded to a multiple of kFifoPairs pairs
// plus kFifoPairs run-out pairs).  The row each lane owns in each slice comes from a
// parallel per-warp stream of row codes, read one slice ahead.
// Log-domain blobs (header.log_domain = 1; weights are log-probabilities, idle records carry -inf) keep
// the end-of-slice flag in bit 31 of the index word instead (so S, D <= 32767); their sub-rows hand
// over (max, sum-exp) pairs instead of plain partial sums.
// TWO STREAM FORMATS (header.embed_codes).  Large graphs stream their records from L2 every frame, and there the row
// code of a slice must arrive through the same prefetched stream (a separate load per slice exposes a full L2
// latency on the thin slices chain graphs are full of): the LAST slot of every lane of a slice is {row code, -aux}
// (sign bit = "this is the code"; aux = leaky(row) in view 1), at the price of one slot per slice.  Small graphs
// whose record streams stay in L1 (records_fit_l1) keep the compact format: no code slot, the sign bit of the
// last arc's weight ends the slice, row codes come from a parallel per-warp stream read two slices ahead.
// Rows whose degree exceeds the view's cap are split into sub-rows; a sub-row writes its
// partial sum to a scratch slot and the listed owner adds the slots up in order.
//
// (dst, pdf) CLASSES (probability-domain blobs).  In chain denominator graphs the pdf of an arc is (almost) a
// function of its destination (pdf = ilabel - 1 sits on the arc, fstext.cc:41, but all arcs that enter an HMM
// state from other states carry the same forward pdf).  Arcs that share destination j AND pdf n form a class c;
// for the C largest classes (>= kMinClass arcs, as many as fit on chip) the packer factorises the recursions:
//     A_c(t)      = sum_{arcs in c} alpha'(t,i) w                 one gather per arc          (view 3)
//     alpha(t+1,j) += A_c(t) x(t,n_c)                              one arc per CLASS           (view 0)
//     gamma(t,n_c) += A_c(t) beta(t+1,j_c) x(t,n_c)                one arc per CLASS           (view 2)
//     bx_c(t)      = beta(t+1,j_c) x(t,n_c);  beta'(t,i) += w bx_c one gather per arc          (view 4)
// The class sums live in an EXTENSION of the state rows (virtual states S_pad .. S_pad+C-1): alpha rows hold
// [alpha'(t,.) | A_c(t)], spilled and reloaded together, the beta row holds [beta(t+1,.) | bx_c(t)].  Arcs outside the
// selected classes keep the two-gather records of views 0-2, so a graph with random pdfs packs exactly as before.
//   view 3 (class pass)   row = class,            a = source state (alpha'),   one operand
//   view 4 (beta pass 2)  row = source state,     a = S_pad + class (bx),      one operand
// Rows of view 1 whose state also has arcs in view 4 carry bit 30 in their row code ("add what view 4 left").
#pragma once
#include <stdint.h>

namespace chainb200 {

constexpr uint32_t kBlobMagic = 0x42324348u;  // "HC2B"
constexpr uint32_t kBlobVersion = 13;
constexpr uint32_t kCodeMarkLog = 0x7fc0c0deu;  // log-domain blobs: weight word of a row-code slot (a NaN no arc can carry)
#ifndef CHAIN_B200_NW
#define CHAIN_B200_NW 16
#endif
constexpr int DEN_NW = CHAIN_B200_NW;           // warps per CTA the slice schedule is built for
constexpr int DEN_NT = DEN_NW * 32;
#ifndef CHAIN_B200_FIFO_PAIRS
#define CHAIN_B200_FIFO_PAIRS 4
#endif
constexpr int kFifoPairs = CHAIN_B200_FIFO_PAIRS;  // record pairs in flight per lane

struct BlobView {
  int32_t n_rows, n_slices, n_split, n_sub;
  int64_t n_recs;
  int64_t off_warp_info;    // int64 [DEN_NW][3]    {record offset of the stream, pairs to run (multiple of kFifoPairs), row-code offset}
  int64_t off_rows;         // int32 [(n_slices + 3*DEN_NW)*32]  per warp, per slice in processing order, + one idle slice:
                            //                      >= 0 row id; -1 idle lane; <= -2: sub-row slot (-2 - slot)
  int64_t off_recs;         // uint2 [n_recs]
  int64_t off_split;        // int32 [n_split*3]    {row, first slot, slot count}
};

constexpr int kNumViews = 5;
constexpr int32_t kRowExtraFlag = 0x40000000;  // view 1: add the partial result view 4 stored for this row

struct BlobHeader {
  uint32_t magic, version;
  int32_t S, K, D, nwarps, S_pad, D_pad, max_sub, log_domain;
  int32_t opt_G;               // group width whose bank geometry the column order was optimised for
  int32_t C, C_pad, R_pad;     // selected (dst,pdf) classes; R_pad = S_pad + C_pad = floats per utterance of a state row
  int32_t embed_codes;         // 1: the last slot of every slice carries the lane's row code (streams read from L2);
                               // 0: row codes are a separate per-warp stream and the end-of-slice flag sits on the last
                               //    arc (graphs whose records stay resident in L1 next to the CTA's shared memory)
  int32_t reserved0;
  BlobView view[kNumViews];
  int64_t off_leaky, off_init, off_final;
  int64_t off_cls_dst, off_cls_pdf;  // int32 [C_pad] destination state / pdf of each class
  int64_t total_bytes;
  double predicted_wavefronts[kNumViews];  // packer's bank-conflict model (per frame, per CTA)
  double ideal_wavefronts[kNumViews];
};

}  // namespace chainb200
