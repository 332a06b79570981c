"""Graph containers of the LF-MMI chain loss: `ChainGraph` / `ChainGraphBatch`.

Drop-in for the reference's `pychain/graph.py` (ChainGraph :23-70,
ChainGraphBatch :73-194): same class names, constructor arguments, attribute
names, dtypes and tensor layouts, so code written
against `pychain.graph` keeps working.  Differences, all behind that surface:

* The reference can only build a graph from an OpenFst object through its
  `simplefst` binding (`openfst_binding/src/fstext.cc`).  OpenFst is not a
  dependency here: `Fst` below is a tiny arc-list container (also readable from
  OpenFst/AT&T text format) and the CSR construction (`fstext.cc:19-117`) and
  the leaky-HMM occupancy estimate (`fstext.cc:120-171`) are restated over
  numpy arrays.  If `simplefst` *is* importable a real `StdVectorFst` is
  accepted too and routed through it, exactly as the reference does.
* `ChainGraphBatch(graph, B)` (one shared graph, the denominator case) does not
  materialise B copies (`graph.py:101-120` uses `.repeat`): the batch tensors
  are stride-0 `expand` views of the single graph, i.e. the `[B,K,3]` layout
  contract is kept but costs O(1) memory, and the engine recognises the
  stride-0 batch dimension and runs its shared-graph kernels.
"""
from __future__ import annotations

import math
from typing import Iterable, List, Optional, Sequence

import numpy as np
import torch

try:  # optional: the reference's OpenFst binding
    import simplefst  # type: ignore
except Exception:  # pragma: no cover - not available in this image
    simplefst = None

__all__ = ["Fst", "ChainGraph", "ChainGraphBatch"]


class Fst(object):
    """Minimal epsilon-free weighted acceptor in the tropical/log semiring.

    Arc weights and final weights are stored as *log-probabilities*
    (= minus the OpenFst weight, as `fstext.cc:38,45` negates them).  `pdf` is
    `ilabel - 1` (`fstext.cc:41`).  Arc order within a state is insertion order
    (what `fst::ArcIterator` yields); states are 0..num_states-1.
    """

    def __init__(self, num_states: int, start: int, src, dst, pdf, log_weight, final_log_weight=None):
        self._num_states = int(num_states)
        self._start = int(start)
        self.src = np.asarray(src, dtype=np.int64).reshape(-1)
        self.dst = np.asarray(dst, dtype=np.int64).reshape(-1)
        self.pdf = np.asarray(pdf, dtype=np.int64).reshape(-1)
        self.log_weight = np.asarray(log_weight, dtype=np.float32).reshape(-1)
        if final_log_weight is None:
            final_log_weight = np.full(self._num_states, -np.inf, dtype=np.float32)
        self.final_log_weight = np.asarray(final_log_weight, dtype=np.float32).reshape(-1)
        n = self.src.shape[0]
        if not (self.dst.shape[0] == n and self.pdf.shape[0] == n and self.log_weight.shape[0] == n):
            raise ValueError("Fst: src/dst/pdf/log_weight must have the same length")
        if self.final_log_weight.shape[0] != self._num_states:
            raise ValueError("Fst: final_log_weight must have num_states entries")
        if n and (self.src.min() < 0 or self.src.max() >= self._num_states or
                  self.dst.min() < 0 or self.dst.max() >= self._num_states):
            raise ValueError("Fst: arc endpoint out of range")
        if n and self.pdf.min() < 0:
            raise ValueError("Fst: negative pdf id (epsilon input labels are not allowed, fstext.cc:41-42)")
        # canonical FST iteration order: by source state, stable (insertion order within a state)
        order = np.argsort(self.src, kind="stable")
        self.src, self.dst, self.pdf, self.log_weight = (
            self.src[order], self.dst[order], self.pdf[order], self.log_weight[order])

    # -- the part of the simplefst.StdVectorFst API that graph.py uses ----------
    def num_states(self) -> int:
        return self._num_states

    def start_state(self) -> int:
        return self._start

    @property
    def num_arcs(self) -> int:
        return int(self.src.shape[0])

    @classmethod
    def from_text(cls, lines: Iterable[str], acceptor: bool = False) -> "Fst":
        """Parse OpenFst text format (`fstprint`): `src dst ilabel olabel [w]`
        arc lines (`src dst label [w]` with `acceptor=True`, what `fstprint --acceptor` writes) and
        `state [w]` final lines; start = source of the first line.
        Weights are tropical/log costs, so log-prob = -w.  pdf = ilabel - 1.
        A line whose field count or label fields do not fit the chosen format raises instead of being
        misread (a 4-column acceptor line would otherwise parse as a transducer arc of weight 0)."""
        src, dst, pdf, lw = [], [], [], []
        finals = {}
        start = None
        nstates = 0
        for line in lines:
            f = line.split()
            if not f:
                continue
            if start is None:
                start = int(f[0])
            n_lab = 1 if acceptor else 2
            if len(f) > 2 + n_lab + 1:
                raise ValueError("Fst.from_text: too many fields in {!r}".format(line))
            if len(f) >= 2 + n_lab:
                try:
                    s, d = int(f[0]), int(f[1])
                    labels = [int(v) for v in f[2:2 + n_lab]]
                except ValueError:
                    raise ValueError("Fst.from_text: {!r} is not a{} arc line (pass acceptor={})".format(
                        line, "n acceptor" if acceptor else " transducer", not acceptor))
                il = labels[0]
                w = float(f[2 + n_lab]) if len(f) > 2 + n_lab else 0.0
                src.append(s), dst.append(d), pdf.append(il - 1), lw.append(-w)
                nstates = max(nstates, s + 1, d + 1)
            elif len(f) == 3 and not acceptor:
                raise ValueError("Fst.from_text: 3-field line {!r}: acceptor format? pass acceptor=True".format(line))
            else:
                s = int(f[0])
                w = float(f[1]) if len(f) >= 2 else 0.0
                finals[s] = -w
                nstates = max(nstates, s + 1)
        fin = np.full(nstates, -np.inf, dtype=np.float32)
        for s, w in finals.items():
            fin[s] = w
        return cls(nstates, start or 0, src, dst, pdf, lw, fin)

    # -- binary files: what simplefst.StdVectorFst.read / read_ark / write do (fstext.cc:6-17, 177-179) ------------
    _MAGIC = 2125659606  # OpenFst's kFstMagicNumber

    @classmethod
    def read(cls, filename: str, offset: int = 0) -> "Fst":
        """Read an OpenFst binary "vector" FST over the "standard" (tropical) arc (what `fstcompile` and Kaldi's
        chain scripts write: den.fst, the numerator FSTs).  Layout: int32 magic | string fsttype | string arctype
        | int32 version | int32 flags | uint64 properties | int64 start | int64 num_states | int64 num_arcs
        (string = int32 length + bytes), then per state `float final | int64 narcs | narcs x {int32 ilabel,
        int32 olabel, float weight, int32 nextstate}`.  Files with embedded symbol tables are refused.
        `offset` seeks first, as `read_ark` (fstext.cc:7-16) does for an entry of a Kaldi archive."""
        with open(filename, "rb") as f:
            f.seek(int(offset))
            buf = f.read()
        pos = 0

        def take(dtype, n=1):
            nonlocal pos
            a = np.frombuffer(buf, dtype=dtype, count=n, offset=pos)
            pos += a.nbytes
            return a

        def take_str():
            nonlocal pos
            n = int(take("<i4")[0])
            if n < 0 or pos + n > len(buf):
                raise ValueError("Fst.read: {}: corrupt header".format(filename))
            v = buf[pos:pos + n].decode("ascii", "replace")
            pos += n
            return v

        try:
            if int(take("<i4")[0]) != cls._MAGIC:
                raise ValueError("Fst.read: {}: not an OpenFst binary file (bad magic number)".format(filename))
            fsttype, arctype = take_str(), take_str()
            version, flags = (int(v) for v in take("<i4", 2))
            take("<u8")  # properties: recomputed from the arcs where needed
            start, nstates, narcs_total = (int(v) for v in take("<i8", 3))
            if fsttype != "vector" or arctype != "standard":
                raise ValueError("Fst.read: {}: fst type {!r} / arc type {!r}; only vector/standard is supported "
                                 "(convert with fstconvert)".format(filename, fsttype, arctype))
            if flags & 3:
                raise ValueError("Fst.read: {}: embedded symbol tables are not supported (strip them with "
                                 "fstsymbols --clear_isymbols --clear_osymbols)".format(filename))
            if flags & 4:
                raise ValueError("Fst.read: {}: aligned files are not supported".format(filename))
            arc_t = np.dtype([("il", "<i4"), ("ol", "<i4"), ("w", "<f4"), ("ns", "<i4")])
            fin = np.empty(nstates, dtype=np.float32)
            counts = np.empty(nstates, dtype=np.int64)
            chunks = []
            for s in range(nstates):
                fin[s] = take("<f4")[0]
                n = int(take("<i8")[0])
                if n < 0:
                    raise ValueError("negative arc count")
                counts[s] = n
                chunks.append(take(arc_t, n))
        except (ValueError, TypeError) as e:
            if isinstance(e, ValueError) and str(e).startswith("Fst.read"):
                raise
            raise ValueError("Fst.read: {}: truncated or corrupt file ({})".format(filename, e))
        arcs = np.concatenate(chunks) if chunks else np.empty(0, dtype=arc_t)
        src = np.repeat(np.arange(nstates, dtype=np.int64), counts)
        return cls(nstates, start, src, arcs["ns"], arcs["il"].astype(np.int64) - 1, -arcs["w"], -fin)

    @classmethod
    def read_ark(cls, filename: str, offset: int) -> "Fst":
        """`simplefst.StdVectorFst.read_ark` (fstext.cc:7-16): the FST that starts at byte `offset` of an archive."""
        return cls.read(filename, offset)

    def write(self, filename: str) -> bool:
        """Write the OpenFst binary layout `read` parses (olabel = ilabel, properties 0 = unknown)."""
        S = self._num_states
        counts = np.bincount(self.src, minlength=S).astype(np.int64)
        arc_t = np.dtype([("il", "<i4"), ("ol", "<i4"), ("w", "<f4"), ("ns", "<i4")])
        arcs = np.empty(self.num_arcs, dtype=arc_t)
        arcs["il"] = arcs["ol"] = self.pdf + 1
        arcs["w"] = -self.log_weight
        arcs["ns"] = self.dst
        ends = np.cumsum(counts)
        with open(filename, "wb") as f:
            f.write(np.array([self._MAGIC], "<i4").tobytes())
            for name in (b"vector", b"standard"):
                f.write(np.array([len(name)], "<i4").tobytes() + name)
            f.write(np.array([2, 0], "<i4").tobytes())
            f.write(np.array([0], "<u8").tobytes())
            f.write(np.array([self._start, S, self.num_arcs], "<i8").tobytes())
            fin = (-self.final_log_weight).astype("<f4")
            for s in range(S):
                f.write(fin[s:s + 1].tobytes())
                f.write(counts[s:s + 1].astype("<i8").tobytes())
                f.write(arcs[ends[s] - counts[s]:ends[s]].tobytes())
        return True

    # the static spellings `pychain/graph.py` calls on simplefst.StdVectorFst (graph.py:44,48,61)
    @staticmethod
    def fst_to_tensor(fst: "Fst", log_domain: bool = False):
        return list(fst.to_tensors(log_domain))

    @staticmethod
    def set_leaky_probs(fst: "Fst") -> torch.Tensor:
        return fst.leaky_probs()

    # -- fstext.cc:19-117 restated ---------------------------------------------
    def to_tensors(self, log_domain: bool = False):
        """CSR by source and CSR by destination, as `FstToTensor` emits them.
        Returns the same 7-tuple in the same order (`fstext.cc:109-116`)."""
        S, K = self._num_states, self.num_arcs
        fwd = np.stack([self.src, self.dst, self.pdf], axis=1).astype(np.int32).reshape(K, 3)
        out_deg = np.bincount(self.src, minlength=S)
        fwd_end = np.cumsum(out_deg)
        fwd_idx = np.stack([fwd_end - out_deg, fwd_end], axis=1).astype(np.int32)
        # transitions_in_tup[dst] is appended while iterating states/arcs in
        # forward order (fstext.cc:36-47) => stable sort of the forward list by dst
        border = np.argsort(self.dst, kind="stable")
        bwd = fwd[border]
        in_deg = np.bincount(self.dst, minlength=S)
        bwd_end = np.cumsum(in_deg)
        bwd_idx = np.stack([bwd_end - in_deg, bwd_end], axis=1).astype(np.int32)
        fwd_p = torch.from_numpy(self.log_weight.copy())
        bwd_p = torch.from_numpy(self.log_weight[border].copy())
        fin = torch.from_numpy(self.final_log_weight.copy())
        if not log_domain:  # fstext.cc:89-90, 101-102, 106-107
            fwd_p.exp_(), bwd_p.exp_(), fin.exp_()
        return (torch.from_numpy(fwd.copy()), fwd_p, torch.from_numpy(fwd_idx),
                torch.from_numpy(bwd.copy()), bwd_p, torch.from_numpy(bwd_idx), fin)

    # -- fstext.cc:120-171 restated --------------------------------------------
    def leaky_probs(self, num_iters: int = 100) -> torch.Tensor:
        """Average state occupancy over `num_iters` steps of the per-state
        normalised HMM started in the start state (float64, returned float32)."""
        S = self._num_states
        w = np.exp(self.log_weight.astype(np.float64))
        tot = np.exp(self.final_log_weight.astype(np.float64)) + np.bincount(self.src, weights=w, minlength=S)
        with np.errstate(divide="ignore"):
            nf = 1.0 / tot
        cur = np.zeros(S, dtype=np.float64)
        cur[self._start] = 1.0
        avg = np.zeros(S, dtype=np.float64)
        for _ in range(num_iters):
            avg += cur * (1.0 / num_iters)
            nxt = np.bincount(self.dst, weights=(cur * nf)[self.src] * w, minlength=S)
            cur = nxt * (1.0 / nxt.sum())
        return torch.from_numpy(avg.astype(np.float32))


def _is_simplefst(fst) -> bool:
    return simplefst is not None and isinstance(fst, simplefst.StdVectorFst)


class ChainGraph(object):
    """One numerator or denominator graph (reference `graph.py:23-70`).

    Attributes (identical to the reference): `num_states`, `num_transitions`,
    `log_domain`, `is_empty`, `start_state`, `forward_transitions [K,3] i32`,
    `forward_transition_indices [S,2] i32`, `forward_transition_probs [K] f32`,
    the three `backward_*` twins, `final_probs [S]`, `initial_probs [S]`,
    `leaky_probs [S]` (None in the log domain).
    """

    def __init__(self, fst, initial_mode="fst", final_mode="fst", log_domain=False):
        self.num_states = fst.num_states()
        assert initial_mode in ["fst", "leaky"]
        assert final_mode in ["fst", "ones"]
        self.log_domain = log_domain
        if _is_simplefst(fst):  # the reference's own route (graph.py:36-44)
            tensors = simplefst.StdVectorFst.fst_to_tensor(fst, log_domain)
            start = simplefst.StdVectorFst.start_state(fst)
            leaky_fn = lambda: simplefst.StdVectorFst.set_leaky_probs(fst)  # noqa: E731
        elif isinstance(fst, Fst):
            tensors = fst.to_tensors(log_domain)
            start = fst.start_state()
            leaky_fn = fst.leaky_probs
        else:
            raise TypeError("ChainGraph expects a pychain_b200.Fst (or a simplefst.StdVectorFst), got {}".format(type(fst)))
        (
            self.forward_transitions,
            self.forward_transition_probs,
            self.forward_transition_indices,
            self.backward_transitions,
            self.backward_transition_probs,
            self.backward_transition_indices,
            self.final_probs,
        ) = tensors

        self.num_transitions = self.forward_transitions.size(0)
        self.is_empty = (self.num_transitions == 0)
        self.start_state = start

        probs_type = self.forward_transition_probs.dtype

        if not self.is_empty:
            if log_domain:
                self.leaky_probs = None  # no leaky hmm if in log domain
                assert initial_mode == "fst", "'leaky' mode is incompatible with log domain"
                self.initial_probs = torch.full([self.num_states], float("-inf"), dtype=probs_type)
                self.initial_probs[self.start_state] = 0.0
                if final_mode == "ones":
                    self.final_probs.fill_(0.0)
            else:
                self.leaky_probs = leaky_fn()
                if initial_mode == "fst":
                    self.initial_probs = torch.zeros([self.num_states], dtype=probs_type)
                    self.initial_probs[self.start_state] = 1.0
                else:
                    self.initial_probs = self.leaky_probs.clone()
                if final_mode == "ones":
                    self.final_probs.fill_(1.0)
        else:
            raise Exception("An empty graph encountered!")
        self._engine_cache = {}  # per-device prepared (packed, device-resident) views, filled by the engine

    _GRAPH_FIELDS = (
        "forward_transitions", "forward_transition_indices", "forward_transition_probs",
        "backward_transitions", "backward_transition_indices", "backward_transition_probs",
        "final_probs", "initial_probs", "leaky_probs",
    )


class ChainGraphBatch(object):
    """A batch of graphs in the reference's padded `[B, ...]` layout
    (`graph.py:73-194`).  Same constructor arguments, attributes and `reorder`.

    The per-graph tensors are described once, in `_FIELDS`: (name, trailing shape in terms of
    K = transitions / S = states, pad value in the probability domain, pad value in the log
    domain); both constructors and `reorder` / `to` / `materialize` are driven by that table.
    """

    #        name                            shape        pad(prob)  pad(log)
    _FIELDS = (
        ("forward_transitions",            ("K", 3),     0,         0),
        ("forward_transition_indices",     ("S", 2),     0,         0),
        ("forward_transition_probs",       ("K",),       0.0,       0.0),
        ("backward_transitions",           ("K", 3),     0,         0),
        ("backward_transition_indices",    ("S", 2),     0,         0),
        ("backward_transition_probs",      ("K",),       0.0,       0.0),
        ("leaky_probs",                    ("S",),       0.0,       None),       # absent in the log domain
        ("initial_probs",                  ("S",),       0.0,       float("-inf")),
        ("final_probs",                    ("S",),       0.0,       float("-inf")),
    )
    _BATCH_FIELDS = tuple(f[0] for f in _FIELDS) + ("start_state",)

    def __init__(self, graphs, batch_size=None, max_num_transitions=None, max_num_states=None):
        self.shared_graph = None  # the single ChainGraph when all B rows are the same graph
        if isinstance(graphs, ChainGraph):
            if not batch_size:
                raise ValueError("batch size should be specified to expand a single graph")
            self.batch_size = batch_size
            self.initialized_by_one(graphs)
        elif isinstance(graphs, (list, ChainGraph)):
            for value, name in ((max_num_transitions, "max_num_transitions"), (max_num_states, "max_num_states")):
                if not value:
                    raise ValueError(
                        "{} should be specified if given a a list of ChainGraph objects to initialize from".format(name))
            self.batch_size = len(graphs)
            self.initialized_by_list(graphs, max_num_transitions, max_num_states)
        else:
            raise ValueError(
                "ChainGraphBatch should be either initialized by a single ChainGraph object or a list of "
                "ChainGraph objects but given {}".format(type(graphs)))

    def initialized_by_one(self, graph):
        """One graph for the whole batch.  The reference repeats it B times
        (`graph.py:101-120`); here every batch tensor is a stride-0 view of the one graph
        (same shapes, O(1) memory), which is also how the engine recognises a shared graph."""
        B = self.batch_size
        self.shared_graph = graph
        self.log_domain = graph.log_domain
        self.num_states, self.num_transitions = graph.num_states, graph.num_transitions
        device = getattr(self, "_device", None)
        for name, _, _, pad_log in self._FIELDS:
            t = getattr(graph, name)
            if t is not None and device is not None:
                t = self._on_device(graph, name, t, device)
            setattr(self, name, None if t is None else t.unsqueeze(0).expand(B, *t.shape))
        self.start_state = torch.full((B,), int(graph.start_state), dtype=torch.long, device=device)

    @staticmethod
    def _on_device(graph, name, t, device):
        """The single graph's tensor on `device`, moved once per (graph, device) and cached on the graph."""
        cache = graph.__dict__.setdefault("_device_tensors", {})
        key = (name, str(device), t.data_ptr(), t._version)
        hit = cache.get(key)
        if hit is None:
            hit = cache[key] = t.to(device)
        return hit

    def initialized_by_list(self, graphs, max_num_transitions, max_num_states):
        """Per-utterance graphs stacked with zero / -inf padding (`graph.py:122-175`): padded
        states have empty (0,0) arc ranges and no initial / final mass."""
        B = self.batch_size
        self.log_domain = graphs[0].log_domain
        self.num_states, self.num_transitions = max_num_states, max_num_transitions
        dims = {"K": max_num_transitions, "S": max_num_states}
        for name, shape, pad_prob, pad_log in self._FIELDS:
            pad = pad_log if self.log_domain else pad_prob
            if pad is None:
                setattr(self, name, None)
                continue
            like = getattr(graphs[0], name)
            full = [B] + [dims.get(d, d) for d in shape]
            setattr(self, name, torch.full(full, pad, dtype=like.dtype))
        self.start_state = torch.zeros([B], dtype=torch.long)
        for i, graph in enumerate(graphs):
            nt, ns = graph.num_transitions, graph.num_states
            if nt > max_num_transitions or ns > max_num_states:
                raise ValueError("graph {} ({} states, {} transitions) exceeds the batch maxima".format(i, ns, nt))
            for name, shape, _, _ in self._FIELDS:
                dst = getattr(self, name)
                if dst is not None:
                    dst[i, :(nt if shape[0] == "K" else ns)].copy_(getattr(graph, name))
            self.start_state[i] = graph.start_state

    def reorder(self, new_order):
        """Select / permute utterances (`graph.py:177-194`).  For a shared graph every row is
        the same graph, so the (stride-0) graph tensors only change their batch size."""
        if self.shared_graph is not None:
            if new_order.numel() != self.batch_size:
                self.batch_size = int(new_order.numel())
                self.initialized_by_one(self.shared_graph)
            return
        for name in self._BATCH_FIELDS:
            t = getattr(self, name)
            if t is not None:
                setattr(self, name, t.index_select(0, new_order.to(t.device)))

    # -- conveniences that the reference does not have ---------------------------
    def to(self, device):
        """Move the batch tensors to `device` once, instead of the reference's per-call
        `.cuda()` inside the extension (`chain-computation.cc:77-89`)."""
        if self.shared_graph is not None:
            # `.to()` of a stride-0 expand view would materialise B copies (and the engine would then take the
            # per-sequence kernel): move the ONE graph and re-expand; `reorder` stays on this device too
            self._device = torch.device(device)
            self.initialized_by_one(self.shared_graph)
            return self
        for name in self._BATCH_FIELDS:
            t = getattr(self, name)
            if t is not None:
                setattr(self, name, t.to(device))
        return self

    def materialize(self):
        """Contiguous `[B, ...]` copies (what the reference's `.repeat` makes); needed only to
        hand a shared-graph batch to the *reference* extension, whose CHECK_CONTIGUOUS
        (`pychain.cc:42-54`) rejects stride-0 views."""
        for name in self._BATCH_FIELDS:
            t = getattr(self, name)
            if t is not None:
                setattr(self, name, t.contiguous())
        self.shared_graph = None
        return self
