"""Flat array bundle of a `GraphLib` — what the Julia host hands to `wq_index_create`.

Field meanings follow include/whippet_b200.h::wq_index_desc, which in turn cites
src/index.jl:5-15, src/graph.jl:104-115, src/edges.jl:10-13 of the reference.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field

import numpy as np

from . import abi
from . import jls


@dataclass
class FlatIndex:
    kmer: int
    gene_offset: np.ndarray     # u32[G]
    node_base: np.ndarray       # u32[G+1]
    nodeoffset: np.ndarray      # u32[N]
    nodecoord: np.ndarray       # u32[N]
    nodelen: np.ndarray         # u32[N]
    edgetype: np.ndarray        # u8[N+G]
    edgeleft: np.ndarray        # u64[N+G]
    edgeright: np.ndarray       # u64[N+G]
    seq4: np.ndarray            # u8[n]
    bwt: np.ndarray             # u8[n]
    sentinel: int
    count: np.ndarray           # i64[4]
    sa: np.ndarray              # u32[n]
    left_ptr: np.ndarray        # u32[4^K+1]
    left_nodes: np.ndarray      # u32[2*nl]
    right_ptr: np.ndarray
    right_nodes: np.ndarray
    iso_base: np.ndarray        # u32[G+1]
    iso_node_ptr: np.ndarray    # u32[I+1]
    iso_nodes: np.ndarray       # u32[]
    gene_names: list = field(default_factory=list)
    iso_names: list = field(default_factory=list)
    gene_strand: np.ndarray | None = None   # u8[G] 1 = '+'  (lib.info[g].strand)
    gene_seqname: list = field(default_factory=list)      # lib.info[g].name (reference sequence; SAM output only)

    _ARRAYS = ["gene_offset", "node_base", "nodeoffset", "nodecoord", "nodelen", "edgetype", "edgeleft",
               "edgeright", "seq4", "bwt", "count", "sa", "left_ptr", "left_nodes", "right_ptr",
               "right_nodes", "iso_base", "iso_node_ptr", "iso_nodes"]
    _DT = dict(gene_offset="<u4", node_base="<u4", nodeoffset="<u4", nodecoord="<u4", nodelen="<u4",
               edgetype="u1", edgeleft="<u8", edgeright="<u8", seq4="u1", bwt="u1", count="<i8", sa="<u4",
               left_ptr="<u4", left_nodes="<u4", right_ptr="<u4", right_nodes="<u4", iso_base="<u4",
               iso_node_ptr="<u4", iso_nodes="<u4")

    def __post_init__(self):
        for k in self._ARRAYS:
            setattr(self, k, np.ascontiguousarray(getattr(self, k), dtype=self._DT[k]))

    # ---- sizes ----
    @property
    def n_genes(self):
        return len(self.gene_offset)

    @property
    def n_nodes(self):
        return len(self.nodelen)

    @property
    def n_isoforms(self):
        return len(self.iso_node_ptr) - 1

    @property
    def text_len(self):
        return len(self.seq4)

    def desc(self) -> abi.IndexDesc:
        d = abi.IndexDesc()
        d.kmer = self.kmer
        d.n_genes = self.n_genes
        d.n_nodes = self.n_nodes
        d.n_isoforms = self.n_isoforms
        d.text_len = self.text_len
        d.sentinel = self.sentinel
        for i in range(4):
            d.count[i] = int(self.count[i])
        for k in self._ARRAYS:
            if k == "count":
                continue
            typ = {"u1": abi.u8p, "<u4": abi.u32p, "<u8": abi.u64p}[self._DT[k]]
            setattr(d, k, abi.ptr(getattr(self, k), typ))
        d._keepalive = self
        return d

    def save(self, path):
        np.savez(path, kmer=self.kmer, sentinel=self.sentinel,
                 gene_names=np.array(self.gene_names),
                 iso_names=np.array(self.iso_names),
                 gene_strand=self.gene_strand if self.gene_strand is not None else np.zeros(0, "u1"),
                 gene_seqname=np.array(self.gene_seqname),
                 **{k: getattr(self, k) for k in self._ARRAYS})

    @classmethod
    def load(cls, path):
        z = np.load(path, allow_pickle=False)
        kw = {k: z[k] for k in cls._ARRAYS}
        return cls(kmer=int(z["kmer"]), sentinel=int(z["sentinel"]), gene_names=list(map(str, z["gene_names"])),
                   iso_names=list(map(str, z["iso_names"])),
                   gene_strand=z["gene_strand"] if len(z["gene_strand"]) else None,
                   gene_seqname=list(map(str, z["gene_seqname"])) if "gene_seqname" in z.files else [], **kw)

    # ---- from the reference's own on-disk format ----
    @classmethod
    def from_jls(cls, path):
        lib = jls.load_jls(path)
        G = len(lib.graphs)
        K = int(lib.kmer)
        node_base = [0]
        cat = {k: [] for k in ("nodeoffset", "nodecoord", "nodelen", "edgetype", "edgeleft", "edgeright", "seq4")}
        iso_base, iso_ptr, iso_nodes, iso_names = [0], [0], [], []
        for g in lib.graphs:
            nn = len(g.nodelen)
            node_base.append(node_base[-1] + nn)
            for k in ("nodeoffset", "nodecoord", "nodelen", "edgetype", "edgeleft", "edgeright"):
                cat[k].append(np.asarray(getattr(g, k)))
            cat["seq4"].append(jls.decode_seq4(g.seq))
            for bs in g.annopath:   # BitSet{bits::Vector{UInt64}, offset}
                bits = np.unpackbits(np.ascontiguousarray(bs.bits).view("u1"), bitorder="little")
                ids = np.nonzero(bits)[0] + 64 * int(bs.offset)
                iso_nodes.extend(int(x) for x in ids)
                iso_ptr.append(len(iso_nodes))
            iso_names.extend(g.annoname)
            iso_base.append(iso_base[-1] + len(g.annopath))
        # edgeleft/edgeright hold #undef garbage wherever build_edges did not assign them
        # (src/graph.jl:212-213, src/edges.jl:104-111): zero those slots so the bundle is deterministic.
        et = np.concatenate(cat["edgetype"]).astype("u1")
        el = np.concatenate(cat["edgeleft"]).astype("<u8")
        er = np.concatenate(cat["edgeright"]).astype("<u8")
        isl = (et == 4) | (et == 2) | (et == 5)
        isr = (et == 4) | (et == 1) | (et == 6)
        last = np.zeros(len(et), bool)        # the last edge of each gene is never visited by build_edges
        for gi in range(G):
            last[node_base[gi + 1] + gi] = True
        el[~isl | last] = 0
        er[~isr | last] = 0

        def csr(table):
            ptr_, nodes = [0], []
            for slot in table:
                if not isinstance(slot, jls.Undef):
                    for e in slot:
                        nodes.extend((int(e["gene"]), int(e["node"])))
                ptr_.append(len(nodes) // 2)
            return np.array(ptr_, "<u4"), np.array(nodes, "<u4")

        lp, ln = csr(lib.edges.left)
        rp, rn = csr(lib.edges.right)
        assert len(lp) == 4 ** K + 1
        fm = lib.index
        return cls(kmer=K, gene_offset=np.asarray(lib.offset), node_base=np.array(node_base),
                   nodeoffset=np.concatenate(cat["nodeoffset"]), nodecoord=np.concatenate(cat["nodecoord"]),
                   nodelen=np.concatenate(cat["nodelen"]), edgetype=et, edgeleft=el, edgeright=er,
                   seq4=np.concatenate(cat["seq4"]), bwt=jls.wavelet_decode(fm.bwt), sentinel=int(fm.sentinel),
                   count=np.asarray(fm.count), sa=np.asarray(fm.samples).astype("<u4"),
                   left_ptr=lp, left_nodes=ln, right_ptr=rp, right_nodes=rn,
                   iso_base=np.array(iso_base), iso_node_ptr=np.array(iso_ptr), iso_nodes=np.array(iso_nodes, "<u4"),
                   gene_names=list(lib.names), iso_names=iso_names,
                   gene_strand=np.array([1 if i.strand else 0 for i in lib.info], "u1"),
                   gene_seqname=[str(i.name) for i in lib.info])


# ------------------------------------------------------------------------------------------
# AlignParam (src/align.jl:2-40)
# ------------------------------------------------------------------------------------------
@dataclass
class AlignParam:
    mismatches: int = 3
    kmer_size: int = 9
    seed_try: int = 3
    seed_tolerate: int = 4
    seed_min_qual: int = 4
    seed_length: int = 18
    seed_buffer: int = 5
    seed_inc: int = 18
    seed_pair_range: int = 5000
    score_range: float = 0.05
    score_min: float = 0.7
    is_stranded: bool = False
    is_paired: bool = False
    is_pair_rc: bool = True
    is_trans_ok: bool = False
    is_circ_ok: bool = False

    @classmethod
    def from_args(cls, args: dict, ispaired=False, kmer=9):
        """AlignParam(args, ispaired; kmer) as built by the CLI (src/align.jl:27-40)."""
        return cls(args["mismatches"], kmer, args["seed-try"], args["seed-tol"], 4, args["seed-len"],
                   args["seed-buf"], args["seed-inc"], args["pair-range"], 0.05, args["score-min"],
                   args["stranded"], ispaired, not args["pair-same-strand"], False, args["circ"])

    def c(self) -> abi.AlignParamC:
        p = abi.AlignParamC()
        for k in ("mismatches", "kmer_size", "seed_try", "seed_tolerate", "seed_min_qual", "seed_length",
                  "seed_buffer", "seed_inc", "seed_pair_range"):
            setattr(p, k, int(getattr(self, k)))
        p.score_range = float(self.score_range)
        p.score_min = float(self.score_min)
        for k in ("is_stranded", "is_paired", "is_pair_rc", "is_trans_ok", "is_circ_ok"):
            setattr(p, k, 1 if getattr(self, k) else 0)
        return p


# ------------------------------------------------------------------------------------------
# Read batches: FASTQRecord after fill! (src/record.jl:13-19), packed for the C ABI
# ------------------------------------------------------------------------------------------
_CODE = np.full(256, 255, dtype=np.uint8)
for _i, _c in enumerate("ACGT"):
    _CODE[ord(_c)] = _i
    _CODE[ord(_c.lower())] = _i


class ReadBatch:
    """2-bit sequences (32 bases per u64 word, read i starts at word seq_off[i]) + phred bytes."""

    def __init__(self, length, seq_off, seq2, qual_off, qual):
        self.len = np.ascontiguousarray(length, "u1")
        self.seq_off = np.ascontiguousarray(seq_off, "<u8")
        self.seq2 = np.ascontiguousarray(seq2, "<u8")
        self.qual_off = np.ascontiguousarray(qual_off, "<u8")
        self.qual = np.ascontiguousarray(qual, "u1")

    @property
    def n(self):
        return len(self.len)

    @classmethod
    def from_codes(cls, codes: np.ndarray, quals: np.ndarray, lens: np.ndarray | None = None):
        """codes/quals: [n, R] uint8 matrices (codes 0..3); lens optional per-read lengths (<= R)."""
        n, R = codes.shape
        lens = np.full(n, R, "u1") if lens is None else np.asarray(lens, "u1")
        W = (R + 31) // 32
        words = np.zeros((n, W), dtype="<u8")
        wb = words.view("u1").reshape(n, W * 8)          # little endian: base j -> byte (j//4), bits 2*(j%4)
        for lo in range(0, n, 1 << 20):
            hi = min(n, lo + (1 << 20))
            pad = np.zeros((hi - lo, W * 32), dtype=np.uint8)
            pad[:, :R] = codes[lo:hi] & 3
            pad[np.arange(W * 32)[None, :] >= lens[lo:hi, None].astype(np.int64)] = 0
            p4 = pad.reshape(hi - lo, W * 8, 4)
            wb[lo:hi] = p4[:, :, 0] | (p4[:, :, 1] << 2) | (p4[:, :, 2] << 4) | (p4[:, :, 3] << 6)
        Q = (R + 15) // 16 * 16
        q = np.zeros((n, Q), dtype=np.uint8)
        q[:, :R] = quals
        return cls(lens, np.arange(n, dtype=np.uint64) * W, words.reshape(-1),
                   np.arange(n, dtype=np.uint64) * Q, q.reshape(-1))

    @classmethod
    def from_strings(cls, seqs, quals, qualoffset=33, fill="A"):
        """seqs: list of str; quals: list of str (FASTQ quality lines).  N -> `fill`
        (FASTQ.Reader(fill_ambiguous=DNA_A), src/reads.jl:9)."""
        n = len(seqs)
        R = max((len(s) for s in seqs), default=1)
        R = max(R, 1)
        codes = np.zeros((n, R), np.uint8)
        q = np.zeros((n, R), np.uint8)
        lens = np.zeros(n, np.uint8)
        fillc = _CODE[ord(fill)]
        for i, (s, ql) in enumerate(zip(seqs, quals)):
            c = _CODE[np.frombuffer(s.encode(), "u1")]
            c[c == 255] = fillc
            codes[i, :len(s)] = c
            q[i, :len(s)] = (np.frombuffer(ql.encode(), "u1").astype(np.int16) - qualoffset).astype(np.uint8)
            lens[i] = len(s)
        return cls.from_codes(codes, q, lens)

    def fixed_stride(self):
        """L when every read sits at the strides of a fixed_len = L batch (read i at word i * ceil(L / 32) and quality byte
        i * ((L + 3) & ~3)), i.e. when the offsets need not be handed over; 0 otherwise."""
        if self.n == 0:
            return 0
        L = int(self.len.max())
        i = np.arange(self.n, dtype=np.uint64)
        ok = L > 0 and np.array_equal(self.seq_off, i * np.uint64((L + 31) // 32)) and np.array_equal(self.qual_off, i * np.uint64((L + 3) & ~3))
        return L if ok else 0

    def to_fixed(self, pack_qual=False):
        """The same reads in the fixed_len layout (rows of ceil(L / 32) words and (L + 3) & ~3 quality bytes, L = the longest read), or
        None when the batch is not stored at constant strides already (e.g. ragged FASTQ input).  pack_qual: when the batch holds at most
        16 distinct quality values (binned instrument qualities), c() hands the qualities over as 4-bit dictionary indices
        (wq_read_batch.qual_bits = 4); `qual` keeps the phred bytes for everything that reads them on the host."""
        n = self.n
        if n == 0:
            return None
        L = int(self.len.max())
        ws, qs = (L + 31) // 32, (L + 3) & ~3
        if n > 1:
            sw, sq = int(self.seq_off[1] - self.seq_off[0]), int(self.qual_off[1] - self.qual_off[0])
            i = np.arange(n, dtype=np.uint64)
            if sw < ws or sq < L or not (np.array_equal(self.seq_off, self.seq_off[0] + i * np.uint64(sw))
                                         and np.array_equal(self.qual_off, self.qual_off[0] + i * np.uint64(sq))):
                return None
        else:
            sw, sq = ws, max(qs, L)
        s0, q0 = int(self.seq_off[0]), int(self.qual_off[0])
        seq = np.ascontiguousarray(np.lib.stride_tricks.as_strided(self.seq2[s0:], (n, ws), (8 * sw, 8)))
        qual = np.zeros((n, qs), "u1")
        src = np.lib.stride_tricks.as_strided(self.qual[q0:], (n, min(qs, sq)), (sq, 1))
        qual[:, :src.shape[1]] = src
        if qs > L:                                             # padding bytes beyond the longest read are zero, as every packer leaves them
            qual[:, L:] = 0
        i = np.arange(n, dtype=np.uint64)
        out = ReadBatch(self.len, i * np.uint64(ws), seq.reshape(-1), i * np.uint64(qs), qual.reshape(-1))
        out.prefer_fixed = True                                # c() then hands the batch over without its offset arrays
        if pack_qual:
            valid = np.arange(qs)[None, :] < self.len[:, None]
            vals = np.unique(qual[valid])
            if len(vals) <= 16:
                lut = np.zeros(256, "u1")
                lut[vals] = np.arange(len(vals), dtype="u1")
                idx = np.where(valid, lut[qual], 0).astype("u1")
                prow = ((L + 1) // 2 + 3) & ~3
                even = np.zeros((n, prow * 2), "u1")
                even[:, :qs] = idx
                out.qual4 = np.ascontiguousarray((even[:, 0::2] | (even[:, 1::2] << 4)).reshape(-1))
                out.qual_dict = np.zeros(16, "u1")
                out.qual_dict[:len(vals)] = vals
        return out

    def c(self, fixed_len: int = 0) -> abi.ReadBatchC:
        """The C struct.  fixed_len = L > 0 (see fixed_stride) leaves seq_off / qual_off NULL: the library derives them on the device."""
        if not fixed_len and getattr(self, "prefer_fixed", False):
            fixed_len = self.fixed_stride()
        b = abi.ReadBatchC()
        b.n_reads = self.n
        b.fixed_len = int(fixed_len)
        b.len = abi.ptr(self.len, abi.u8p)
        b.seq2 = abi.ptr(self.seq2, abi.u64p)
        b.seq2_words = len(self.seq2)
        if fixed_len and getattr(self, "qual4", None) is not None:
            b.qual = abi.ptr(self.qual4, abi.u8p)
            b.qual_bytes = len(self.qual4)
            b.qual_bits = 4
            for k in range(16):
                b.qual_dict[k] = int(self.qual_dict[k])
        else:
            b.qual = abi.ptr(self.qual, abi.u8p)
            b.qual_bytes = len(self.qual)
        if not fixed_len:
            b.seq_off = abi.ptr(self.seq_off, abi.u64p)
            b.qual_off = abi.ptr(self.qual_off, abi.u64p)
        b._keepalive = self
        return b

    def slice(self, lo, hi):
        """Reads lo..hi-1 as a new batch sharing the packed buffers."""
        return ReadBatch(self.len[lo:hi], self.seq_off[lo:hi], self.seq2, self.qual_off[lo:hi], self.qual)
