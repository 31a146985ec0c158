"""TEST INFRASTRUCTURE ONLY: ctypes driver of tests/emu/wq_host_emu.cu (the library's HOST stages run on given count
stores; no kernel is launched)."""
import ctypes as C
import os
import subprocess

import numpy as np

import whippet_jl_b200 as wq
from whippet_jl_b200.build import nvcc_path

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libwq_host_emu.so")
_CSRC = os.path.join(_HERE, "..", "..", "whippet.jl_b200", "csrc")


def build():
    srcs = [os.path.join(_HERE, "wq_host_emu.cu")] + [os.path.join(_CSRC, f) for f in
                                                      ("wq_em.cuh", "wq_events.h", "wq_sam.h", "wq_host_quant.h", "wq_host_index.h", "wq_types.h", "wq_kernels.cuh", "wq_inflate.h")]
    if not os.path.exists(_SO) or os.path.getmtime(_SO) < max(map(os.path.getmtime, srcs)):
        subprocess.check_call([nvcc_path(), "-gencode", "arch=compute_100a,code=sm_100a", "-O2", "-std=c++17", "-fmad=false",
                               "-Xcompiler", "-fPIC", "-shared", "-o", _SO, srcs[0], "-lz"])
    return _SO


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class HostEmu:
    def __init__(self, index: wq.FlatIndex):
        build()
        self.L = C.CDLL(_SO)
        self.L.hemu_create.restype = C.c_void_p
        self.L.hemu_error.restype = C.c_char_p
        for f in ("hemu_pack", "hemu_n_classes", "hemu_class", "hemu_edge_adds", "hemu_events", "hemu_event_path"):
            getattr(self.L, f).restype = C.c_uint64
        self.L.hemu_event.restype = C.c_int64
        self.index = index
        self._d = index.desc()
        self.h = C.c_void_p(self.L.hemu_create(C.byref(self._d)))
        assert self.h

    def close(self):
        if self.h:
            self.L.hemu_destroy(self.h)
            self.h = None

    def set_counts(self, node, ekey, ecount, longs: dict, multis: dict):
        def flat(d):
            keys = list(d)
            ptr_ = np.zeros(len(keys) + 1, "<u8")
            ptr_[1:] = np.cumsum([len(k) for k in keys])
            words = np.array([w for k in keys for w in k] or [0], "<u4")
            cnt = np.array([int(d[k]) for k in keys] or [0], "<u8")
            return len(keys), ptr_, words, cnt
        node = np.ascontiguousarray(node, "<u8")
        ekey, ecount = np.ascontiguousarray(ekey, "<u8"), np.ascontiguousarray(ecount, "<u8")
        nl, lp, lw, lc = flat(longs)
        nm, mp, mw, mc = flat(multis)
        self.L.hemu_set_counts(self.h, _p(node), C.c_uint64(len(ekey)), _p(ekey), _p(ecount),
                               C.c_uint64(nl), _p(lp), _p(lw), _p(lc), C.c_uint64(nm), _p(mp), _p(mw), _p(mc))

    def set_bias(self, node_pg, edge_pg, long_pg, multi_pg):
        """--biascorrect: (primer-adjusted count, gc factor) arrays in the orders given to set_counts (multi-map records in sorted key order)."""
        arrs = [np.ascontiguousarray(a, "<f8") for pg in (node_pg, edge_pg, long_pg, multi_pg) for a in pg]
        self.L.hemu_set_bias(self.h, *[_p(a if len(a) else np.zeros(1)) for a in arrs])

    def build_classes(self, part=0, nparts=1):
        rc = self.L.hemu_build_classes(self.h, C.c_uint32(part), C.c_uint32(nparts))
        assert rc == 0, self.L.hemu_error(self.h)

    def share(self) -> bytes:
        ptr_ = C.c_void_p()
        n = self.L.hemu_pack(self.h, C.byref(ptr_))
        return C.string_at(ptr_, n)

    def assemble(self, shares):
        keep = [C.create_string_buffer(s, len(s)) for s in shares]
        bufs = (C.c_void_p * len(shares))(*[C.addressof(k) for k in keep])
        lens = (C.c_uint64 * len(shares))(*[len(s) for s in shares])
        rc = self.L.hemu_assemble(self.h, bufs, lens, C.c_uint32(len(shares)))
        assert rc == 0, self.L.hemu_error(self.h)

    def classes(self):
        """([(ids, count, isdone, postassign)], n_multi) in canonical order."""
        nm = C.c_uint32()
        n = self.L.hemu_n_classes(self.h, C.byref(nm))
        ids = np.zeros(4096, "<i4")
        out = []
        for c in range(n):
            cnt, fl = C.c_double(), C.c_int32()
            k = self.L.hemu_class(self.h, C.c_uint64(c), _p(ids), C.c_uint64(len(ids)), C.byref(cnt), C.byref(fl))
            out.append((ids[:k].tolist(), cnt.value, bool(fl.value & 1), bool(fl.value & 2)))
        return out, int(nm.value)

    def iso_count(self):
        a = np.zeros(self.index.n_isoforms)
        self.L.hemu_iso_count(self.h, _p(a))
        return a

    def assign_ambig(self, props, genetpm):
        """assign_ambig! with the EM's results given (class proportions flattened in canonical class order, gene TPM).
        Returns (node_value[N], edge_key, edge_value) with the additions merged."""
        props, genetpm = np.ascontiguousarray(props, "<f8"), np.ascontiguousarray(genetpm, "<f8")
        rc = self.L.hemu_assign_ambig(self.h, _p(props), C.c_uint64(len(props)), _p(genetpm))
        assert rc == 0, self.L.hemu_error(self.h)
        self.L.hemu_values.restype = C.c_uint64
        n = self.L.hemu_values(self.h, None, None, None)
        node, ek, ev = np.zeros(self.index.n_nodes), np.zeros(n, "<u8"), np.zeros(n)
        self.L.hemu_values(self.h, _p(node), _p(ek), _p(ev))
        return node, ek, ev

    def set_event_pieces(self, k):
        """Build every gene in pieces of k visited nodes (0 = whole genes), as the library does for genes with many nodes."""
        self.L.hemu_set_event_pieces(int(k))

    def event_cuts(self, ekey, evalue, nparts, threads=4):
        """The event stage's work list and the ranks' cuts (wq_event_make_pieces / wq_event_cuts) for sorted edge keys and their
        values.  Returns (pieces[n, 4] = g_lo, g_hi, node_lo, node_hi; cum[n + 1] running cost estimate; cuts[nparts + 1])."""
        ekey, evalue = np.ascontiguousarray(ekey, "<u8"), np.ascontiguousarray(evalue, "<f8")
        self.L.hemu_event_cuts.restype = C.c_uint64
        n = self.L.hemu_event_cuts(self.h, C.c_uint64(len(ekey)), _p(ekey), _p(evalue), int(threads), int(nparts), None, None, None)
        pieces, cum, cuts = np.zeros((n, 4), "<u4"), np.zeros(n + 1), np.zeros(nparts + 1, "<u4")
        self.L.hemu_event_cuts(self.h, C.c_uint64(len(ekey)), _p(ekey), _p(evalue), int(threads), int(nparts), _p(pieces), _p(cum), _p(cuts))
        return pieces, cum, cuts

    def events(self, node_value, ekey, evalue, readlen):
        """Event construction of every gene.  Returns a list of dicts (gene, node, motif, kind, flags, n_utr, psi, total,
        problem, inc paths, exc paths / utr paths) and the EM problems as tuples for oracle.em_psi."""
        node_value = np.ascontiguousarray(node_value, "<f8")
        ekey, evalue = np.ascontiguousarray(ekey, "<u8"), np.ascontiguousarray(evalue, "<f8")
        n = self.L.hemu_events(self.h, _p(node_value), C.c_uint64(len(ekey)), _p(ekey), _p(evalue), C.c_int64(readlen))
        hdr, vals, buf = (C.c_uint32 * 9)(), (C.c_double * 2)(), np.zeros(70000, "<u4")
        out = []
        for i in range(n):
            prob = self.L.hemu_event(self.h, C.c_uint64(i), hdr, vals)
            paths = []
            for k in range(hdr[8]):
                m = self.L.hemu_event_path(self.h, C.c_uint64(i), C.c_uint64(k), _p(buf), C.c_uint64(len(buf)))
                paths.append(tuple(int(x) for x in buf[:m]))
            out.append(dict(gene=hdr[0], node=hdr[1], motif=hdr[2], kind=hdr[3], flags=hdr[4], n_utr=hdr[5], n_inc=hdr[6], n_exc=hdr[7],
                            psi=vals[0], total=vals[1], problem=int(prob), paths=paths))
        return out

    def events_part(self, node_value, ekey, evalue, readlen, part, nparts, threads=4):
        """Event construction of rank `part` of `nparts` (its range of the work list, as the library cuts it).  Returns the records as
        (gene, node, motif, kind, flags, n_paths, psi, total) tuples."""
        node_value = np.ascontiguousarray(node_value, "<f8")
        ekey, evalue = np.ascontiguousarray(ekey, "<u8"), np.ascontiguousarray(evalue, "<f8")
        self.L.hemu_events_part.restype = C.c_uint64
        n = self.L.hemu_events_part(self.h, _p(node_value), C.c_uint64(len(ekey)), _p(ekey), _p(evalue), C.c_int64(readlen), int(threads), int(nparts), int(part))
        hdr, vals = (C.c_uint32 * 9)(), (C.c_double * 2)()
        out = []
        for i in range(n):
            self.L.hemu_event(self.h, C.c_uint64(i), hdr, vals)
            out.append((hdr[0], hdr[1], hdr[2], hdr[3], hdr[4], hdr[8], vals[0], vals[1]))
        return out

    def problem(self, p):
        sizes = (C.c_uint32 * 5)()
        self.L.hemu_problem(self.h, C.c_uint64(p), sizes, None, None, None, None, None)
        npth, ninc, namb, tandem, nap = (int(x) for x in sizes)
        cnt, ln = np.zeros(npth), np.zeros(npth)
        aptr, apaths, amult = np.zeros(namb + 1, "<u4"), np.zeros(max(1, nap), "<u4"), np.zeros(max(1, namb))
        self.L.hemu_problem(self.h, C.c_uint64(p), sizes, _p(cnt), _p(ln), _p(aptr), _p(apaths), _p(amult))
        sets = [[int(x) for x in apaths[int(aptr[a]):int(aptr[a + 1])]] for a in range(namb)]
        return ninc, cnt, ln, sets, amult[:namb].tolist(), bool(tandem)


def psi_cluster(problems, C_blocks, threads, readlen, sig=4, maxit=1500):
    """The cluster PSI kernel's phases (wq_k_em_psi_cl, csrc/wq_em.cuh) run on the CPU for `C_blocks` blocks of `threads` threads.
    problems as for GraphLibQuant.em_psi_batch.  Returns [(psi, iterations)], or None when the layout does not fit / a set names a
    path twice (the library keeps such events on the block kernel); also the shared-memory bytes one block needs."""
    L = C.CDLL(build())
    E = len(problems)
    path_ptr, amb_ptr, amb_path_ptr = [0], [0], [0]
    n_inc, cnt, ln, apaths, amult, tan = [], [], [], [], [], []
    for (ni, c, l, sets, mult, t) in problems:
        n_inc.append(ni); cnt += list(c); ln += list(l); tan.append(1 if t else 0)
        path_ptr.append(len(cnt))
        for a, m in zip(sets, mult):
            apaths += list(a); amult.append(m); amb_path_ptr.append(len(apaths))
        amb_ptr.append(len(amult))
    arr = lambda x, dt: np.ascontiguousarray(x if len(x) else [0], dt)
    a = [arr(path_ptr, "<u4"), arr(n_inc, "<u4"), arr(cnt, "<f8"), arr(ln, "<f8"), arr(amb_ptr, "<u4"), arr(amb_path_ptr, "<u4"),
         arr(apaths, "<u4"), arr(amult, "<f8"), arr(tan, "u1")]
    psi, ct, its, need = np.zeros(max(1, len(cnt))), np.zeros(max(1, len(cnt))), np.zeros(max(1, E), "<i4"), C.c_uint64()
    rc = L.hemu_psi_cluster(C.c_uint32(E), *[_p(x) for x in a], C.c_int64(readlen), C.c_int(sig), C.c_int64(maxit), C.c_uint32(C_blocks),
                            C.c_uint32(threads), _p(psi), _p(ct), _p(its), C.byref(need))
    assert rc in (0, -1), f"hemu_psi_cluster: the layout's capacities do not cover the build (rc {rc})"
    if rc != 0:
        return None, int(need.value)
    return [(psi[path_ptr[e]:path_ptr[e + 1]].copy(), int(its[e])) for e in range(E)], int(need.value)


def inflate(data: bytes, piece: int = 1 << 20, cap: int | None = None):
    """WqInflate (csrc/wq_inflate.h) on a gzip file image.  Returns (text, error message or "", members)."""
    L = C.CDLL(build())
    L.hemu_inflate.restype = C.c_int64
    cap = cap if cap is not None else max(64, 1100 * len(data) + 1024)
    out = np.zeros(cap, "u1")
    src = np.frombuffer(data, "u1") if len(data) else np.zeros(1, "u1")
    err = C.create_string_buffer(256)
    members = C.c_uint64()
    r = L.hemu_inflate(_p(src), C.c_uint64(len(data)), _p(out), C.c_uint64(cap), C.c_uint64(piece), err, C.c_uint64(256), C.byref(members))
    n = r if r >= 0 else -1 - r
    return out[:n].tobytes(), (err.value.decode() if r < 0 else ""), int(members.value)


def inflate_raw(data: bytes, out_len: int):
    """WqInflate::raw_into on a raw deflate stream.  Returns (text or None, error message, canary intact): the output buffer is followed
    by 64 canary bytes that the inflater must not touch."""
    L = C.CDLL(build())
    out = np.full(out_len + 64, 0xA5, "u1")
    src = np.frombuffer(data, "u1") if len(data) else np.zeros(1, "u1")
    err = C.create_string_buffer(256)
    rc = L.hemu_inflate_raw(_p(src), C.c_uint64(len(data)), _p(out), C.c_uint64(out_len), err, C.c_uint64(256))
    intact = bool(np.all(out[out_len:] == 0xA5))
    return (out[:out_len].tobytes() if rc == 0 else None), err.value.decode(), intact


class HostSam:
    """wq_sam.h (write_sam, src/io.jl:69-271) on the CPU: SAM text of a batch from given alignments."""

    def __init__(self, index: wq.FlatIndex):
        from whippet_jl_b200 import abi
        build()
        self.abi = abi
        self.L = C.CDLL(_SO)
        self.L.hsam_create.restype = C.c_void_p
        self.L.hsam_batch.restype = C.c_uint64
        self.L.hsam_header.restype = C.c_uint64
        self.L.hsam_text.restype = C.c_void_p
        self._d = index.desc()
        G = index.n_genes
        names = (C.c_char_p * G)(*[str(x).encode() for x in index.gene_seqname])
        st = np.ascontiguousarray(index.gene_strand, "u1")
        self.h = C.c_void_p(self.L.hsam_create(C.byref(self._d), names, _p(st)))

    def header(self):
        n = self.L.hsam_header(self.h)
        return C.string_at(self.L.hsam_text(self.h), n).decode()

    def batch(self, reads, names, res, mate=None, mate_names=None, qualoffset=33):
        abi = self.abi

        def pack(nm):
            enc = [x.encode() for x in nm]
            ln = np.array([len(x) for x in enc], "<u4")
            off = np.zeros(len(enc), "<u8")
            if len(enc) > 1:
                off[1:] = np.cumsum(ln[:-1])
            return off, ln, b"".join(enc)
        out = abi.AlignOutC()
        out.hit_start, out.nhits, out.flipped = abi.ptr(res.hit_start, abi.u32p), abi.ptr(res.nhits, abi.u8p), abi.ptr(res.flipped, abi.u8p)
        out.aln = res.aln.ctypes.data
        out.aln_mate = res.aln_mate.ctypes.data if res.aln_mate is not None else None
        out.nodes = res.nodes.ctypes.data
        out.aln_cap = out.aln_used = len(res.aln)
        out.nodes_cap = out.nodes_used = len(res.nodes)
        b = reads.c()
        o1, l1, t1 = pack(names)
        if mate is not None:
            m = mate.c()
            o2, l2, t2 = pack(mate_names)
            n = self.L.hsam_batch(self.h, C.byref(b), C.byref(m), _p(o1), _p(l1), t1, _p(o2), _p(l2), t2, C.byref(out), int(qualoffset))
        else:
            n = self.L.hsam_batch(self.h, C.byref(b), None, _p(o1), _p(l1), t1, None, None, None, C.byref(out), int(qualoffset))
        return C.string_at(self.L.hsam_text(self.h), n).decode("latin-1")

    def close(self):
        if self.h:
            self.L.hsam_destroy(self.h)
            self.h = None
