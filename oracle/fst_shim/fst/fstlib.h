// Stand-in for the part of OpenFst's <fst/fstlib.h> that the reference's openfst_binding/src/fstext.cc touches.
//
// TEST INFRASTRUCTURE ONLY (lives under oracle/; nothing in pychain_b200/ includes it).  OpenFst is not in this
// image and there is no network, so the reference's `simplefst` extension cannot be built as its setup.py
// describes (setup.py:4-13 needs libfst.so).  This header is an independent, minimal container with the same
// names -- VectorFst<StdArc>, ArcIterator, TropicalWeight::Value(), FstHeader, FstReadOptions, Read/Write -- so that
// the reference's fstext.cc compiles UNMODIFIED, where it lies, into oracle/_ref/simplefst.so (oracle/build_ref.py).
// What that pins: FstToTensor (fstext.cc:19-117) and SetLeakyProbs (fstext.cc:120-171), i.e. the CSR construction,
// the arc order, the label -> pdf shift, the exp() of the weights and the 100-step occupancy average, all executed
// by the reference's own code.  What it does NOT pin: OpenFst itself (its arc iteration order is insertion order,
// which is what this container yields, and its binary file layout, restated below from the published format).
//
// Binary layout written/read here (OpenFst "vector" FST, file version 2, little endian):
//   int32 magic 2125659606 | string "vector" | string "standard" | int32 version | int32 flags | uint64 properties
//   | int64 start | int64 num_states | int64 num_arcs, strings being int32 length + bytes; then per state:
//   float final_weight | int64 narcs | narcs x {int32 ilabel, int32 olabel, float weight, int32 nextstate}.
// Symbol tables (flags & 3) are not supported and make Read fail.
#ifndef ORACLE_FST_SHIM_FSTLIB_H_
#define ORACLE_FST_SHIM_FSTLIB_H_

#include <cassert>
#include <cmath>
#include <cstdint>
#include <fstream>
#include <iostream>
#include <limits>
#include <string>
#include <vector>

using int32 = int32_t;   // OpenFst's <fst/types.h> puts these at global scope; fstext.cc uses `int32` unqualified
using int64 = int64_t;
using uint64 = uint64_t;

namespace fst {

constexpr int32 kFstMagicNumber = 2125659606;
constexpr int kNoStateId = -1;

class TropicalWeight {
 public:
  TropicalWeight() : v_(std::numeric_limits<float>::infinity()) {}
  TropicalWeight(float v) : v_(v) {}  // NOLINT: implicit like OpenFst's
  float Value() const { return v_; }
  static TropicalWeight Zero() { return TropicalWeight(); }
  static TropicalWeight One() { return TropicalWeight(0.f); }
 private:
  float v_;
};

struct StdArc {
  typedef TropicalWeight Weight;
  typedef int Label;
  typedef int StateId;
  StdArc() : ilabel(0), olabel(0), weight(), nextstate(kNoStateId) {}
  StdArc(int il, int ol, Weight w, int ns) : ilabel(il), olabel(ol), weight(w), nextstate(ns) {}
  int ilabel, olabel;
  Weight weight;
  int nextstate;
};

namespace shim_detail {
inline bool read_string(std::istream &is, std::string *s) {
  int32 n = 0;
  is.read(reinterpret_cast<char *>(&n), sizeof(n));
  if (!is || n < 0 || n > (1 << 20)) return false;
  s->resize(n);
  if (n) is.read(&(*s)[0], n);
  return static_cast<bool>(is);
}
inline void write_string(std::ostream &os, const std::string &s) {
  const int32 n = static_cast<int32>(s.size());
  os.write(reinterpret_cast<const char *>(&n), sizeof(n));
  os.write(s.data(), n);
}
template <class T> inline void rd(std::istream &is, T *v) { is.read(reinterpret_cast<char *>(v), sizeof(T)); }
template <class T> inline void wr(std::ostream &os, const T &v) { os.write(reinterpret_cast<const char *>(&v), sizeof(T)); }
}  // namespace shim_detail

class FstHeader {
 public:
  FstHeader() : version_(0), flags_(0), properties_(0), start_(kNoStateId), numstates_(0), numarcs_(0) {}
  bool Read(std::istream &is, const std::string &source, bool rewind = false) {
    (void)source; (void)rewind;
    int32 magic = 0;
    shim_detail::rd(is, &magic);
    if (!is || magic != kFstMagicNumber) return false;
    if (!shim_detail::read_string(is, &fsttype_) || !shim_detail::read_string(is, &arctype_)) return false;
    shim_detail::rd(is, &version_); shim_detail::rd(is, &flags_); shim_detail::rd(is, &properties_);
    shim_detail::rd(is, &start_); shim_detail::rd(is, &numstates_); shim_detail::rd(is, &numarcs_);
    return static_cast<bool>(is);
  }
  void Write(std::ostream &os) const {
    shim_detail::wr(os, kFstMagicNumber);
    shim_detail::write_string(os, fsttype_); shim_detail::write_string(os, arctype_);
    shim_detail::wr(os, version_); shim_detail::wr(os, flags_); shim_detail::wr(os, properties_);
    shim_detail::wr(os, start_); shim_detail::wr(os, numstates_); shim_detail::wr(os, numarcs_);
  }
  std::string fsttype_, arctype_;
  int32 version_, flags_;
  uint64 properties_;
  int64 start_, numstates_, numarcs_;
};

struct FstReadOptions {
  explicit FstReadOptions(const std::string &src = "<unspecified>", const FstHeader *hdr = nullptr)
      : source(src), header(hdr) {}
  std::string source;
  const FstHeader *header;
};

template <class Arc>
class VectorFst {
 public:
  typedef Arc ArcType;
  typedef typename Arc::Weight Weight;
  VectorFst() : start_(kNoStateId) {}

  int NumStates() const { return static_cast<int>(states_.size()); }
  int Start() const { return start_; }
  Weight Final(int s) const { return states_[s].final; }
  size_t NumArcs(int s) const { return states_[s].arcs.size(); }

  int AddState() { states_.emplace_back(); return NumStates() - 1; }
  void SetStart(int s) { start_ = s; }
  void SetFinal(int s, Weight w) { states_[s].final = w; }
  void AddArc(int s, const Arc &a) { states_[s].arcs.push_back(a); }
  const std::vector<Arc> &Arcs(int s) const { return states_[s].arcs; }

  static VectorFst *Read(std::istream &is, const FstReadOptions &opts) {
    FstHeader own;
    const FstHeader *h = opts.header;
    if (!h) {
      if (!own.Read(is, opts.source)) return nullptr;
      h = &own;
    }
    if (h->fsttype_ != "vector" || h->arctype_ != "standard" || (h->flags_ & 3) || h->numstates_ < 0) return nullptr;
    VectorFst *f = new VectorFst;
    f->start_ = static_cast<int>(h->start_);
    f->states_.resize(static_cast<size_t>(h->numstates_));
    for (auto &st : f->states_) {
      float fw = 0.f;
      int64 narcs = 0;
      shim_detail::rd(is, &fw);
      shim_detail::rd(is, &narcs);
      if (!is || narcs < 0) { delete f; return nullptr; }
      st.final = Weight(fw);
      st.arcs.resize(static_cast<size_t>(narcs));
      for (auto &a : st.arcs) {
        float w = 0.f;
        shim_detail::rd(is, &a.ilabel); shim_detail::rd(is, &a.olabel); shim_detail::rd(is, &w);
        shim_detail::rd(is, &a.nextstate);
        a.weight = Weight(w);
      }
      if (!is) { delete f; return nullptr; }
    }
    return f;
  }
  static VectorFst *Read(const std::string &filename) {
    std::ifstream is(filename, std::ifstream::binary);
    if (!is) return nullptr;
    return Read(is, FstReadOptions(filename));
  }
  bool Write(const std::string &filename) const {
    std::ofstream os(filename, std::ofstream::binary);
    if (!os) return false;
    FstHeader h;
    h.fsttype_ = "vector"; h.arctype_ = "standard"; h.version_ = 2; h.flags_ = 0; h.properties_ = 0;
    h.start_ = start_; h.numstates_ = NumStates(); h.numarcs_ = 0;
    for (const auto &st : states_) h.numarcs_ += static_cast<int64>(st.arcs.size());
    h.Write(os);
    for (const auto &st : states_) {
      shim_detail::wr(os, st.final.Value());
      shim_detail::wr(os, static_cast<int64>(st.arcs.size()));
      for (const auto &a : st.arcs) {
        shim_detail::wr(os, a.ilabel); shim_detail::wr(os, a.olabel); shim_detail::wr(os, a.weight.Value());
        shim_detail::wr(os, a.nextstate);
      }
    }
    return static_cast<bool>(os);
  }

 private:
  struct State {
    Weight final;
    std::vector<Arc> arcs;
  };
  std::vector<State> states_;
  int start_;
};

typedef VectorFst<StdArc> StdVectorFst;

template <class F>
class ArcIterator {
 public:
  ArcIterator(const F &fst, int s) : arcs_(fst.Arcs(s)), i_(0) {}
  bool Done() const { return i_ >= arcs_.size(); }
  void Next() { ++i_; }
  const typename F::ArcType &Value() const { return arcs_[i_]; }
 private:
  const std::vector<typename F::ArcType> &arcs_;
  size_t i_;
};

}  // namespace fst

#endif  // ORACLE_FST_SHIM_FSTLIB_H_
