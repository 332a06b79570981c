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
// Rows whose degree exceeds the view's cap are split into sub-rows; a sub-row writes its
// partial sum to a scratch slot and the listed owner adds the slots up in order.
#pragma once
#include <stdint.h>

namespace chainb200 {

constexpr uint32_t kBlobMagic = 0x42324348u;  // "HC2B"
constexpr uint32_t kBlobVersion = 9;
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

struct BlobHeader {
  uint32_t magic, version;
  int32_t S, K, D, nwarps, S_pad, D_pad, max_sub, log_domain;
  int32_t opt_G, reserved[3];  // group width whose bank geometry the column order was optimised for
  BlobView view[3];
  int64_t off_leaky, off_init, off_final;
  int64_t total_bytes;
  double predicted_wavefronts[3];  // packer's bank-conflict model for G=4 (per frame, per CTA)
  double ideal_wavefronts[3];
};

}  // namespace chainb200
