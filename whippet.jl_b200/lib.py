"""ctypes binding of csrc/libwhippet_b200.so (the C ABI of include/whippet_b200.h).

Fails loudly when the CUDA library is missing: there is no CPU path behind it."""
from __future__ import annotations

import ctypes as C
import os

from . import abi
from .build import CUDA_SO

_lib = None


class WhippetB200Error(RuntimeError):
    pass


def load():
    global _lib
    if _lib is not None:
        return _lib
    so = os.environ.get("WQ_LIB", CUDA_SO)      # experiments: alternative builds of the same library
    if not os.path.exists(so):
        raise WhippetB200Error(f"{CUDA_SO} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                               "(there is no CPU fallback)")
    L = C.CDLL(so)
    vp = C.c_void_p
    L.wq_last_error.restype = C.c_char_p
    L.wq_index_create.argtypes = [C.POINTER(abi.IndexDesc), C.c_int, C.POINTER(vp)]
    L.wq_set_biascorrect.argtypes = [vp, C.c_int]
    L.wq_bias_download.argtypes = [vp, C.POINTER(abi.BiasCountsC)]
    L.wq_index_write.argtypes = [C.POINTER(abi.IndexDesc), C.POINTER(C.c_char_p), abi.u8p, C.c_char_p]
    L.wq_index_load.argtypes = [C.c_char_p, C.c_int, C.POINTER(vp)]
    L.wq_destroy.argtypes = [vp]
    L.wq_destroy.restype = None
    L.wq_quant_reset.argtypes = [vp]
    L.wq_align_se.argtypes = [vp, C.POINTER(abi.AlignParamC), C.POINTER(abi.ReadBatchC), C.POINTER(abi.AlignOutC)]
    L.wq_align_pe.argtypes = [vp, C.POINTER(abi.AlignParamC), C.POINTER(abi.ReadBatchC), C.POINTER(abi.ReadBatchC),
                              C.POINTER(abi.AlignOutC)]
    L.wq_align_se_device.argtypes = [vp, C.POINTER(abi.AlignParamC), C.POINTER(abi.ReadBatchC)]
    L.wq_align_pe_device.argtypes = [vp, C.POINTER(abi.AlignParamC), C.POINTER(abi.ReadBatchC), C.POINTER(abi.ReadBatchC)]
    L.wq_map_stats_get.argtypes = [vp, C.POINTER(abi.MapStatsC)]
    L.wq_counts_device_ptr.argtypes = [vp, C.POINTER(vp), C.POINTER(C.c_uint64)]
    L.wq_counts_download.argtypes = [vp, C.POINTER(abi.CountsC)]
    L.wq_keyed_download.argtypes = [vp, C.c_int, C.POINTER(abi.KeyedC)]
    L.wq_counts_merge.argtypes = [vp, C.POINTER(abi.CountsC), C.POINTER(abi.KeyedC), C.POINTER(abi.KeyedC)]
    L.wq_counts_refresh_dense.argtypes = [vp]
    L.wq_em_tpm.argtypes = [vp, C.POINTER(abi.EmParamC), abi.f64p, abi.f64p, abi.f64p, C.POINTER(C.c_int64)]
    L.wq_assign_ambig.argtypes = [vp, C.POINTER(abi.ValuesC)]
    L.wq_em_psi_batch.argtypes = [vp, C.POINTER(abi.PsiBatchC)]
    L.wq_process_events.argtypes = [vp, C.c_int64, C.POINTER(abi.EventsOutC)]
    L.wq_process_events_view.argtypes = [vp, C.c_int64, C.POINTER(abi.EventsOutC)]
    L.wq_last_kernel_ms.argtypes = [vp, C.POINTER(C.c_float)]
    L.wq_pinned_alloc.argtypes = [C.POINTER(vp), C.c_uint64]
    L.wq_pinned_free.argtypes = [vp]
    L.wq_set_stream.argtypes = [vp, vp]
    L.wq_sparse_pack.argtypes = [vp, vp, C.c_uint64, C.POINTER(C.c_uint64)]
    L.wq_sparse_merge.argtypes = [vp, vp, C.c_uint64]
    L.wq_sparse_device_pack.argtypes = [vp, C.POINTER(vp), C.POINTER(C.c_uint64)]
    L.wq_sparse_device_merge.argtypes = [vp, vp, C.c_uint64]
    L.wq_set_gene_partition.argtypes = [vp, C.c_uint32, C.c_uint32]
    L.wq_classes_build_local.argtypes = [vp]
    L.wq_classes_pack.argtypes = [vp, C.POINTER(vp), C.POINTER(C.c_uint64)]
    L.wq_classes_assemble.argtypes = [vp, C.POINTER(vp), C.POINTER(C.c_uint64), C.c_uint32]
    L.wq_fastq_open.argtypes = [C.c_char_p, C.c_int, C.POINTER(vp)]
    L.wq_fastq_from_memory.argtypes = [C.c_char_p, C.c_uint64, C.c_int, C.POINTER(vp)]
    L.wq_fastq_next.argtypes = [vp, C.c_uint32, C.POINTER(abi.ReadBatchC)]
    L.wq_fastq_names.argtypes = [vp, C.POINTER(C.POINTER(C.c_uint64)), C.POINTER(C.POINTER(C.c_uint32)), C.POINTER(vp)]
    L.wq_fastq_close.argtypes = [vp]
    L.wq_fastq_close.restype = None
    L.wq_process_fastq.argtypes = [vp, C.POINTER(abi.AlignParamC), C.c_char_p, C.c_char_p, C.c_int, C.c_uint32, C.POINTER(abi.MapStatsC)]
    L.wq_process_fastq_sam.argtypes = [vp, C.POINTER(abi.AlignParamC), C.c_char_p, C.c_char_p, C.c_int, C.c_uint32, C.c_int, C.POINTER(abi.MapStatsC)]
    L.wq_index_set_gene_info.argtypes = [vp, C.POINTER(C.c_char_p), abi.u8p]
    L.wq_sam_header.argtypes = [vp, C.c_char_p, C.c_uint64, C.POINTER(C.c_uint64)]
    L.wq_sam_batch.argtypes = [vp, C.POINTER(abi.ReadBatchC), C.POINTER(abi.ReadBatchC),
                               abi.u64p, abi.u32p, C.c_char_p, abi.u64p, abi.u32p, C.c_char_p,
                               C.POINTER(abi.AlignOutC), C.c_int, C.c_char_p, C.c_uint64, C.POINTER(C.c_uint64)]
    L.wq_comm_unique_id.argtypes = [vp]
    L.wq_comm_init_rank.argtypes = [vp, C.c_int, C.c_int, C.c_char_p]
    L.wq_counts_sync.argtypes = [vp]
    L.wq_classes_sync.argtypes = [vp]
    L.wq_comm_destroy.argtypes = [vp]
    L.wq_launch_count.argtypes = [vp]
    L.wq_launch_count.restype = C.c_uint64
    L.wq_last_deferred.argtypes = [vp]
    L.wq_last_deferred.restype = C.c_uint64
    L.wq_quant_kernel_stats.argtypes = [vp, C.POINTER(C.c_double)]
    _lib = L
    return L


def check(rc):
    if rc != 0:
        raise WhippetB200Error(f"libwhippet_b200 error {rc}: {load().wq_last_error().decode(errors='replace')}")
