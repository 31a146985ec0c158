"""ctypes mirror of include/whippet_b200.h (struct layouts only, no logic)."""
from __future__ import annotations

import ctypes as C

import numpy as np

u8p = C.POINTER(C.c_uint8)
u32p = C.POINTER(C.c_uint32)
u64p = C.POINTER(C.c_uint64)
f64p = C.POINTER(C.c_double)
i32p = C.POINTER(C.c_int32)

WQ_MAX_PATH = 24
WQ_FAST_PATH = 12      # path capacity of the first-pass extension kernel (csrc/wq_types.h)
WQ_MAX_TOLERATE = 8


class IndexDesc(C.Structure):
    _fields_ = [
        ("kmer", C.c_int32), ("n_genes", C.c_uint32), ("n_nodes", C.c_uint32), ("n_isoforms", C.c_uint32),
        ("text_len", C.c_uint64),
        ("gene_offset", u32p), ("node_base", u32p), ("nodeoffset", u32p), ("nodecoord", u32p), ("nodelen", u32p),
        ("edgetype", u8p), ("edgeleft", u64p), ("edgeright", u64p), ("seq4", u8p), ("bwt", u8p),
        ("sentinel", C.c_uint64), ("count", C.c_int64 * 4), ("sa", u32p),
        ("left_ptr", u32p), ("left_nodes", u32p), ("right_ptr", u32p), ("right_nodes", u32p),
        ("iso_base", u32p), ("iso_node_ptr", u32p), ("iso_nodes", u32p),
    ]


class AlignParamC(C.Structure):
    _fields_ = [
        ("mismatches", C.c_int32), ("kmer_size", C.c_int32), ("seed_try", C.c_int32), ("seed_tolerate", C.c_int32),
        ("seed_min_qual", C.c_int32), ("seed_length", C.c_int32), ("seed_buffer", C.c_int32), ("seed_inc", C.c_int32),
        ("seed_pair_range", C.c_int32), ("_pad", C.c_int32),
        ("score_range", C.c_double), ("score_min", C.c_double),
        ("is_stranded", C.c_uint8), ("is_paired", C.c_uint8), ("is_pair_rc", C.c_uint8), ("is_trans_ok", C.c_uint8),
        ("is_circ_ok", C.c_uint8), ("_pad2", C.c_uint8 * 3),
    ]


class ReadBatchC(C.Structure):
    _fields_ = [
        ("n_reads", C.c_uint32), ("fixed_len", C.c_uint32),
        ("len", u8p), ("seq_off", u64p), ("seq2", u64p), ("seq2_words", C.c_uint64),
        ("qual_off", u64p), ("qual", u8p), ("qual_bytes", C.c_uint64),
        ("qual_bits", C.c_uint8), ("qual_dict", C.c_uint8 * 16), ("_pad2", C.c_uint8 * 7),
    ]


ALIGN_NODE_DT = np.dtype([("gene", "<u4"), ("node", "<u4"), ("mistolerance", "<f4"),
                          ("matches", "u1"), ("mismatches", "u1"), ("_pad", "<u2")])
ALIGNMENT_DT = np.dtype([("offset", "<u4"), ("path_start", "<u4"), ("offsetread", "u1"),
                         ("strand", "u1"), ("isvalid", "u1"), ("npath", "u1")])
assert ALIGN_NODE_DT.itemsize == 16 and ALIGNMENT_DT.itemsize == 12


class AlignOutC(C.Structure):
    _fields_ = [
        ("hit_start", u32p), ("nhits", u8p), ("flipped", u8p),
        ("aln", C.c_void_p), ("aln_mate", C.c_void_p), ("nodes", C.c_void_p),
        ("aln_cap", C.c_uint64), ("nodes_cap", C.c_uint64), ("aln_used", C.c_uint64), ("nodes_used", C.c_uint64),
    ]


class MapStatsC(C.Structure):
    _fields_ = [("total", C.c_uint64), ("mapped", C.c_uint64), ("multi", C.c_uint64),
                ("sum_readlen", C.c_uint64), ("path_overflow", C.c_uint64), ("mean_readlen", C.c_double)]


class CountsC(C.Structure):
    _fields_ = [("n_nodes", C.c_uint64), ("node_count", u64p),
                ("n_edge", C.c_uint64), ("edge_key", u64p), ("edge_count", u64p),
                ("n_circ", C.c_uint64), ("circ_key", u64p), ("circ_count", u64p)]


class BiasCountsC(C.Structure):
    _fields_ = [(k, C.c_uint64) for k in ("n_nodes", "n_edge", "n_circ", "n_long", "n_multi")] + \
               [(k, C.POINTER(C.c_double)) for k in ("node_primer", "node_gc", "edge_primer", "edge_gc", "circ_primer", "circ_gc",
                                                     "long_primer", "long_gc", "multi_primer", "multi_gc", "fore", "back")]


class KeyedC(C.Structure):
    _fields_ = [("n_rec", C.c_uint64), ("n_words", C.c_uint64),
                ("rec_ptr", u64p), ("words", u32p), ("count", u64p)]


class EmParamC(C.Structure):
    _fields_ = [("readlen", C.c_int64), ("sig", C.c_int64), ("maxit", C.c_int64)]


class ValuesC(C.Structure):
    _fields_ = [("n_nodes", C.c_uint64), ("node_value", f64p),
                ("n_edge", C.c_uint64), ("edge_key", u64p), ("edge_value", f64p),
                ("n_circ", C.c_uint64), ("circ_key", u64p), ("circ_value", f64p)]


class PsiBatchC(C.Structure):
    _fields_ = [("n_events", C.c_uint32), ("_pad", C.c_uint32),
                ("path_ptr", u32p), ("n_inc", u32p), ("path_count", f64p), ("path_length", f64p),
                ("amb_ptr", u32p), ("amb_path_ptr", u32p), ("amb_paths", u32p), ("amb_mult", f64p),
                ("is_tandem", u8p),
                ("readlen", C.c_int64), ("sig", C.c_int64), ("maxit", C.c_int64),
                ("psi", f64p), ("path_total", f64p), ("iterations", i32p)]


def ptr(a: np.ndarray, typ):
    """Borrow a numpy array's memory as a typed C pointer (array must stay alive)."""
    if a is None:
        return C.cast(None, typ)
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(typ)


EVENT_DT = np.dtype([("gene", "<u4"), ("node", "<u4"), ("motif", "u1"), ("kind", "u1"), ("flags", "u1"), ("_pad", "u1"),
                     ("n_utr", "<u4"), ("n_inc", "<u4"), ("n_exc", "<u4"), ("value_start", "<u4"), ("path_start", "<u4"),
                     ("iterations", "<i4"), ("_pad2", "<u4"), ("psi", "<f8"), ("total", "<f8"),
                     ("ci_lo", "<f8"), ("ci_hi", "<f8")])
assert EVENT_DT.itemsize == 72


class EventsOutC(C.Structure):
    _fields_ = [("n_events", C.c_uint64), ("events", C.c_void_p), ("n_values", C.c_uint64), ("values", f64p),
                ("n_paths", C.c_uint64), ("path_ptr", u64p), ("n_path_nodes", C.c_uint64), ("path_nodes", u32p)]
