# WhippetB200.jl — `ccall` binding of libwhippet_b200.so (include/whippet_b200.h) for the reference's Julia host.
#
# UNTESTED: Julia is not installed in the build image nor on the GPU box; the same ABI is exercised from Python
# ctypes (whippet.jl_b200/lib.py, quant.py) by every GPU test.  Struct layouts mirror the header field for field
# (all fields naturally aligned, explicit padding as in C).  See INTEGRATION.md for where each call goes in
# bin/whippet-quant.jl / src/reads.jl.
module WhippetB200

const LIB = get(ENV, "WHIPPET_B200_LIB", "libwhippet_b200.so")

struct IndexDesc            # wq_index_desc
   kmer::Int32; n_genes::UInt32; n_nodes::UInt32; n_isoforms::UInt32; text_len::UInt64
   gene_offset::Ptr{UInt32}; node_base::Ptr{UInt32}; nodeoffset::Ptr{UInt32}; nodecoord::Ptr{UInt32}; nodelen::Ptr{UInt32}
   edgetype::Ptr{UInt8}; edgeleft::Ptr{UInt64}; edgeright::Ptr{UInt64}; seq4::Ptr{UInt8}; bwt::Ptr{UInt8}
   sentinel::UInt64; count::NTuple{4,Int64}; sa::Ptr{UInt32}
   left_ptr::Ptr{UInt32}; left_nodes::Ptr{UInt32}; right_ptr::Ptr{UInt32}; right_nodes::Ptr{UInt32}
   iso_base::Ptr{UInt32}; iso_node_ptr::Ptr{UInt32}; iso_nodes::Ptr{UInt32}
end

struct AlignParamC          # wq_align_param (AlignParam, src/align.jl:2-19)
   mismatches::Int32; kmer_size::Int32; seed_try::Int32; seed_tolerate::Int32; seed_min_qual::Int32
   seed_length::Int32; seed_buffer::Int32; seed_inc::Int32; seed_pair_range::Int32; _pad::Int32
   score_range::Float64; score_min::Float64
   is_stranded::UInt8; is_paired::UInt8; is_pair_rc::UInt8; is_trans_ok::UInt8; is_circ_ok::UInt8
   _pad2::NTuple{3,UInt8}
end

struct ReadBatch            # wq_read_batch
   n_reads::UInt32; fixed_len::UInt32      # fixed_len = L > 0: read i at word i * cld(L, 32) / quality byte i * ((L + 3) & ~3); offsets may be C_NULL
   len::Ptr{UInt8}; seq_off::Ptr{UInt64}; seq2::Ptr{UInt64}; seq2_words::UInt64
   qual_off::Ptr{UInt64}; qual::Ptr{UInt8}; qual_bytes::UInt64
   qual_bits::UInt8; qual_dict::NTuple{16,UInt8}; _pad2::NTuple{7,UInt8}      # qual_bits = 4: two 4-bit indices into qual_dict per byte (fixed_len batches)
end
ReadBatch(n, fixed_len, len, seq_off, seq2, seq2_words, qual_off, qual, qual_bytes) =
   ReadBatch(n, fixed_len, len, seq_off, seq2, seq2_words, qual_off, qual, qual_bytes, 0x00, ntuple(_ -> 0x00, 16), ntuple(_ -> 0x00, 7))

struct AlignNode            # wq_align_node (SGAlignNode + SGAlignScore)
   gene::UInt32; node::UInt32; mistolerance::Float32; matches::UInt8; mismatches::UInt8; _pad::UInt16
end

struct Alignment            # wq_alignment (SGAlignment; the path is nodes[path_start+1 : path_start+npath])
   offset::UInt32; path_start::UInt32; offsetread::UInt8; strand::UInt8; isvalid::UInt8; npath::UInt8
end

mutable struct AlignOut     # wq_align_out
   hit_start::Ptr{UInt32}; nhits::Ptr{UInt8}; flipped::Ptr{UInt8}
   aln::Ptr{Alignment}; aln_mate::Ptr{Alignment}; nodes::Ptr{AlignNode}
   aln_cap::UInt64; nodes_cap::UInt64; aln_used::UInt64; nodes_used::UInt64
end

struct MapStats             # wq_map_stats
   total::UInt64; mapped::UInt64; multi::UInt64; sum_readlen::UInt64; path_overflow::UInt64; mean_readlen::Float64
end

struct EmParam; readlen::Int64; sig::Int64; maxit::Int64; end          # wq_em_param

mutable struct Counts       # wq_counts
   n_nodes::UInt64; node_count::Ptr{UInt64}
   n_edge::UInt64;  edge_key::Ptr{UInt64}; edge_count::Ptr{UInt64}
   n_circ::UInt64;  circ_key::Ptr{UInt64}; circ_count::Ptr{UInt64}
end

mutable struct Keyed        # wq_keyed
   n_rec::UInt64; n_words::UInt64; rec_ptr::Ptr{UInt64}; words::Ptr{UInt32}; count::Ptr{UInt64}
end

mutable struct Values       # wq_values (stores after assign_ambig!)
   n_nodes::UInt64; node_value::Ptr{Float64}
   n_edge::UInt64;  edge_key::Ptr{UInt64}; edge_value::Ptr{Float64}
   n_circ::UInt64;  circ_key::Ptr{UInt64}; circ_value::Ptr{Float64}
end

struct Event                # wq_event
   gene::UInt32; node::UInt32; motif::UInt8; kind::UInt8; flags::UInt8; _pad::UInt8
   n_utr::UInt32; n_inc::UInt32; n_exc::UInt32; value_start::UInt32; path_start::UInt32
   iterations::Int32; _pad2::UInt32
   psi::Float64; total::Float64; ci_lo::Float64; ci_hi::Float64
end

mutable struct EventsOut    # wq_events_out
   n_events::UInt64; events::Ptr{Event}
   n_values::UInt64; values::Ptr{Float64}
   n_paths::UInt64;  path_ptr::Ptr{UInt64}
   n_path_nodes::UInt64; path_nodes::Ptr{UInt32}
end
EventsOut() = EventsOut(0, C_NULL, 0, C_NULL, 0, C_NULL, 0, C_NULL)

last_error() = unsafe_string(ccall((:wq_last_error, LIB), Cstring, ()))
check(rc) = rc == 0 ? nothing : error("libwhippet_b200 error $rc: $(last_error())")

# ---- index ------------------------------------------------------------------------------------------------
"""`desc` fields must stay alive (GC.@preserve the arrays they point into) for the duration of the call."""
function index_create(desc::IndexDesc, device::Integer = 0)
   h = Ref{Ptr{Cvoid}}(C_NULL)
   check(ccall((:wq_index_create, LIB), Cint, (Ref{IndexDesc}, Cint, Ref{Ptr{Cvoid}}), desc, device, h))
   h[]
end
"""The index in the library's native on-disk form (csrc/wq_index_file.h).  `index_write` is host-only (no GPU needed): run it once after
`whippet-index.jl` has built the GraphLib; `index_load` then replaces `open(deserialize, indexname)` + `index_create` at the start of
every `whippet-quant.jl` run (bin/whippet-quant.jl:105).  `seqnames` / `strands` = lib.info[g].name / .strand ('+' = 1), for SAM output."""
function index_write(desc::IndexDesc, path::String, seqnames::Union{Vector{String},Nothing} = nothing, strands::Union{Vector{UInt8},Nothing} = nothing)
   if seqnames === nothing || strands === nothing
      check(ccall((:wq_index_write, LIB), Cint, (Ref{IndexDesc}, Ptr{Cstring}, Ptr{UInt8}, Cstring), desc, C_NULL, C_NULL, path))
   else
      check(ccall((:wq_index_write, LIB), Cint, (Ref{IndexDesc}, Ptr{Cstring}, Ptr{UInt8}, Cstring), desc, seqnames, strands, path))
   end
end
function index_load(path::String, device::Integer = 0)
   h = Ref{Ptr{Cvoid}}(C_NULL)
   check(ccall((:wq_index_load, LIB), Cint, (Cstring, Cint, Ref{Ptr{Cvoid}}), path, device, h))
   h[]
end
destroy(h) = ccall((:wq_destroy, LIB), Cvoid, (Ptr{Cvoid},), h)

# ---- --biascorrect (bin/whippet-quant.jl:116-122: CounterType = JointBiasCounter, BiasModelType = JointBiasMod) ----
"""Call before the first batch.  The quant stage calls (em_tpm, assign_ambig, process_events) then work with the Float64 counts of
primer_adjust! / gc_adjust! (bin/whippet-quant.jl:156-169) without further calls from the host."""
set_biascorrect(h, on::Bool = true) = check(ccall((:wq_set_biascorrect, LIB), Cint, (Ptr{Cvoid}, Cint), h, on ? 1 : 0))
mutable struct BiasCounts   # wq_bias_counts
   n_nodes::UInt64; n_edge::UInt64; n_circ::UInt64; n_long::UInt64; n_multi::UInt64
   node_primer::Ptr{Float64}; node_gc::Ptr{Float64}; edge_primer::Ptr{Float64}; edge_gc::Ptr{Float64}
   circ_primer::Ptr{Float64}; circ_gc::Ptr{Float64}; long_primer::Ptr{Float64}; long_gc::Ptr{Float64}
   multi_primer::Ptr{Float64}; multi_gc::Ptr{Float64}; fore::Ptr{Float64}; back::Ptr{Float64}
end
"""Every counter's `count` after primer_adjust! and the factor gc_adjust! multiplies it by (< 0: none), in the orders of
counts_download / keyed_download, for refilling JointBiasCounter.count of SpliceGraphQuant / MultiMapping before the writers run."""
function bias_download(h)
   o = BiasCounts(0, 0, 0, 0, 0, ntuple(_ -> Ptr{Float64}(C_NULL), 12)...)
   check(ccall((:wq_bias_download, LIB), Cint, (Ptr{Cvoid}, Ref{BiasCounts}), h, o))             # sizes
   mk(n) = (Vector{Float64}(undef, n), Vector{Float64}(undef, n))
   node, edge, circ, long, multi = mk(o.n_nodes), mk(o.n_edge), mk(o.n_circ), mk(o.n_long), mk(o.n_multi)
   fore, back = Vector{Float64}(undef, 4096), Vector{Float64}(undef, 4096)
   GC.@preserve node edge circ long multi fore back begin
      o.node_primer, o.node_gc = pointer(node[1]), pointer(node[2]); o.edge_primer, o.edge_gc = pointer(edge[1]), pointer(edge[2])
      o.circ_primer, o.circ_gc = pointer(circ[1]), pointer(circ[2]); o.long_primer, o.long_gc = pointer(long[1]), pointer(long[2])
      o.multi_primer, o.multi_gc = pointer(multi[1]), pointer(multi[2]); o.fore, o.back = pointer(fore), pointer(back)
      check(ccall((:wq_bias_download, LIB), Cint, (Ptr{Cvoid}, Ref{BiasCounts}), h, o))
   end
   (node = node, edge = edge, circ = circ, long = long, multi = multi, fore = fore, back = back)
end
quant_reset(h) = check(ccall((:wq_quant_reset, LIB), Cint, (Ptr{Cvoid},), h))
set_gene_info(h, seqnames::Vector{String}, strands::Vector{UInt8}) =
   check(ccall((:wq_index_set_gene_info, LIB), Cint, (Ptr{Cvoid}, Ptr{Cstring}, Ptr{UInt8}), h, seqnames, strands))

# ---- alignment + counting (process_reads! / process_paired_reads! bodies) --------------------------------------
align_se(h, p::AlignParamC, b::ReadBatch, out = C_NULL) =
   check(ccall((:wq_align_se, LIB), Cint, (Ptr{Cvoid}, Ref{AlignParamC}, Ref{ReadBatch}, Ptr{Cvoid}), h, p, b, out))
align_pe(h, p::AlignParamC, f::ReadBatch, r::ReadBatch, out = C_NULL) =
   check(ccall((:wq_align_pe, LIB), Cint, (Ptr{Cvoid}, Ref{AlignParamC}, Ref{ReadBatch}, Ref{ReadBatch}, Ptr{Cvoid}), h, p, f, r, out))
function map_stats(h)
   st = Ref(MapStats(0, 0, 0, 0, 0, 0.0))
   check(ccall((:wq_map_stats_get, LIB), Cint, (Ptr{Cvoid}, Ref{MapStats}), h, st))
   st[]
end
"""Whole files (gzip or plain); `rev = nothing` for single end.  Returns (mapped, total, mean_readlen)."""
function process_fastq(h, p::AlignParamC, fwd::String, rev::Union{String,Nothing} = nothing; qualoffset = 33, batch_reads = 0,
                       sam::Bool = false)
   st = Ref(MapStats(0, 0, 0, 0, 0, 0.0))
   samfd = sam ? Cint(1) : Cint(-1)          # `--sam` writes to stdout (src/reads.jl:49-52); needs set_gene_info first
   check(ccall((:wq_process_fastq_sam, LIB), Cint, (Ptr{Cvoid}, Ref{AlignParamC}, Cstring, Cstring, Cint, UInt32, Cint, Ref{MapStats}),
               h, p, fwd, rev === nothing ? C_NULL : rev, qualoffset, batch_reads, samfd, st))
   s = st[]
   s.mapped, s.total, s.mean_readlen      # the running mean of src/reads.jl:69, floored by the caller as bin/whippet-quant.jl:145 does
end

# ---- count stores (refill SpliceGraphQuant.node/edge/circ/long and MultiMapping.map for the writers) ----------------
function counts_download(h)
   c = Counts(0, C_NULL, 0, C_NULL, C_NULL, 0, C_NULL, C_NULL)
   check(ccall((:wq_counts_download, LIB), Cint, (Ptr{Cvoid}, Ref{Counts}), h, c))          # sizes
   node = Vector{UInt64}(undef, c.n_nodes); ek = Vector{UInt64}(undef, c.n_edge); ec = similar(ek)
   ck = Vector{UInt64}(undef, c.n_circ); cc = similar(ck)
   GC.@preserve node ek ec ck cc begin
      c.node_count = pointer(node); c.edge_key = pointer(ek); c.edge_count = pointer(ec); c.circ_key = pointer(ck); c.circ_count = pointer(cc)
      check(ccall((:wq_counts_download, LIB), Cint, (Ptr{Cvoid}, Ref{Counts}), h, c))
   end
   node, ek, ec, ck, cc      # keys: gene << 32 | lnode << 16 | rnode
end
function keyed_download(h, kind::Integer)                     # 0 = long paths, 1 = multi-map classes
   k = Keyed(0, 0, C_NULL, C_NULL, C_NULL)
   check(ccall((:wq_keyed_download, LIB), Cint, (Ptr{Cvoid}, Cint, Ref{Keyed}), h, kind, k))
   ptr = Vector{UInt64}(undef, k.n_rec + 1); words = Vector{UInt32}(undef, max(k.n_words, 1)); cnt = Vector{UInt64}(undef, max(k.n_rec, 1))
   GC.@preserve ptr words cnt begin
      k.rec_ptr = pointer(ptr); k.words = pointer(words); k.count = pointer(cnt)
      check(ccall((:wq_keyed_download, LIB), Cint, (Ptr{Cvoid}, Cint, Ref{Keyed}), h, kind, k))
   end
   ptr, words, cnt
end

# ---- quantification (bin/whippet-quant.jl:161-181) ----------------------------------------------------------------
"""build_equivalence_classes! + calculate_tpm! + gene_em! + set_gene_tpm!; fills tpm / count / genetpm, returns iterations."""
function em_tpm!(h, tpm::Vector{Float64}, count::Vector{Float64}, genetpm::Vector{Float64}; readlen, sig = 1, maxit = 10000)
   it = Ref{Int64}(0)
   check(ccall((:wq_em_tpm, LIB), Cint, (Ptr{Cvoid}, Ref{EmParam}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ref{Int64}),
               h, EmParam(readlen, sig, maxit), tpm, count, genetpm, it))
   it[]
end
function assign_ambig(h)
   v = Values(0, C_NULL, 0, C_NULL, C_NULL, 0, C_NULL, C_NULL)
   check(ccall((:wq_assign_ambig, LIB), Cint, (Ptr{Cvoid}, Ref{Values}), h, v))             # sizes (and the work)
   node = Vector{Float64}(undef, v.n_nodes); ek = Vector{UInt64}(undef, v.n_edge); ev = Vector{Float64}(undef, v.n_edge)
   ck = Vector{UInt64}(undef, v.n_circ); cv = Vector{Float64}(undef, v.n_circ)
   GC.@preserve node ek ev ck cv begin
      v.node_value = pointer(node); v.edge_key = pointer(ek); v.edge_value = pointer(ev); v.circ_key = pointer(ck); v.circ_value = pointer(cv)
      check(ccall((:wq_assign_ambig, LIB), Cint, (Ptr{Cvoid}, Ref{Values}), h, v))
   end
   node, ek, ev, ck, cv
end
"""process_events without the writers: arrays owned by the handle (valid until the count state changes), zero copy."""
function process_events_view(h, readlen::Integer)
   o = EventsOut()
   check(ccall((:wq_process_events_view, LIB), Cint, (Ptr{Cvoid}, Int64, Ref{EventsOut}), h, readlen, o))
   events = unsafe_wrap(Array, o.events, o.n_events)
   values = unsafe_wrap(Array, o.values, o.n_values)
   path_ptr = unsafe_wrap(Array, o.path_ptr, o.n_paths + 1)
   path_nodes = unsafe_wrap(Array, o.path_nodes, o.n_path_nodes)
   events, values, path_ptr, path_nodes
   # event e: kind 1 -> PSI line (psi, total, ci_lo, ci_hi; paths path_start .. path_start + n_inc + n_exc - 1, their psi in
   # values[value_start+1 ...]); kind 2 -> tandem-UTR lines; kind 0 -> output_empty; kind 3 -> no line
end

# ---- multi-GPU, exchange inside the library (one process per GPU; libnccl.so.2 is opened at run time) ------------------
"""Rank 0: 128 bytes (ncclUniqueId) to hand to the other ranks by any means (a file, a socket, MPI)."""
function comm_unique_id()
   id = Vector{UInt8}(undef, 128)
   check(ccall((:wq_comm_unique_id, LIB), Cint, (Ptr{UInt8},), id))
   id
end
comm_init(h, nranks::Integer, rank::Integer, id::Vector{UInt8}) =
   check(ccall((:wq_comm_init_rank, LIB), Cint, (Ptr{Cvoid}, Cint, Cint, Ptr{UInt8}), h, nranks, rank, id))
"""After the last batch: one ncclAllReduce of the dense counts, one ncclAllGather of the sparse stores, merged on the device."""
counts_sync(h) = check(ccall((:wq_counts_sync, LIB), Cint, (Ptr{Cvoid},), h))
"""Class building partitioned by gene: this rank's share, all-gathered, assembled.  Then em_tpm! / assign_ambig / process_events."""
classes_sync(h) = check(ccall((:wq_classes_sync, LIB), Cint, (Ptr{Cvoid},), h))
comm_destroy(h) = check(ccall((:wq_comm_destroy, LIB), Cint, (Ptr{Cvoid},), h))

# ---- multi-GPU, step by step (for a host that runs its own NCCL calls) ---------------------------------------------------
function counts_device_ptr(h)
   p = Ref{Ptr{Cvoid}}(C_NULL); n = Ref{UInt64}(0)
   check(ccall((:wq_counts_device_ptr, LIB), Cint, (Ptr{Cvoid}, Ref{Ptr{Cvoid}}, Ref{UInt64}), h, p, n))
   p[], n[]                  # ncclAllReduce(sum, uint64) this vector in place
end
function sparse_device_pack(h)
   p = Ref{Ptr{Cvoid}}(C_NULL); n = Ref{UInt64}(0)
   check(ccall((:wq_sparse_device_pack, LIB), Cint, (Ptr{Cvoid}, Ref{Ptr{Cvoid}}, Ref{UInt64}), h, p, n))
   p[], n[]                  # device buffer: ncclAllGather, then sparse_device_merge for every other rank's buffer
end
sparse_device_merge(h, dbuf::Ptr{Cvoid}, bytes::Integer) =
   check(ccall((:wq_sparse_device_merge, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, UInt64), h, dbuf, bytes))
set_gene_partition(h, part::Integer, nparts::Integer) =
   check(ccall((:wq_set_gene_partition, LIB), Cint, (Ptr{Cvoid}, UInt32, UInt32), h, part, nparts))
classes_build_local(h) = check(ccall((:wq_classes_build_local, LIB), Cint, (Ptr{Cvoid},), h))
function classes_pack(h)
   p = Ref{Ptr{Cvoid}}(C_NULL); n = Ref{UInt64}(0)
   check(ccall((:wq_classes_pack, LIB), Cint, (Ptr{Cvoid}, Ref{Ptr{Cvoid}}, Ref{UInt64}), h, p, n))
   unsafe_wrap(Array, Ptr{UInt8}(p[]), n[])       # host buffer owned by the handle: all-gather it between ranks
end
classes_assemble(h, shares::Vector{Vector{UInt8}}) = GC.@preserve shares check(ccall((:wq_classes_assemble, LIB), Cint,
   (Ptr{Cvoid}, Ptr{Ptr{Cvoid}}, Ptr{UInt64}, UInt32), h, [Ptr{Cvoid}(pointer(s)) for s in shares], UInt64[length(s) for s in shares], length(shares)))

# ---- SAM (process_reads! with sam = true) -----------------------------------------------------------------------------
function sam_header(h)
   n = Ref{UInt64}(0)
   check(ccall((:wq_sam_header, LIB), Cint, (Ptr{Cvoid}, Ptr{UInt8}, UInt64, Ref{UInt64}), h, C_NULL, 0, n))
   buf = Vector{UInt8}(undef, n[])
   check(ccall((:wq_sam_header, LIB), Cint, (Ptr{Cvoid}, Ptr{UInt8}, UInt64, Ref{UInt64}), h, buf, n[], n))
   String(buf)
end

end # module
