"""Input fabrication for tests and benchmarks (NOT part of the product library).

The reference's indices come from `whippet-index.jl` (GTF + genome FASTA -> GraphLib); neither Julia nor the GENCODE
files exist offline, and index construction is outside the accelerated path.  This package
builds GraphLib-shaped inputs so the hot path has something to run on:

  synth.py       random gene models, read simulators (single / paired / arbitrary node paths)
  graphbuild.py  the SpliceGraph constructor (src/graph.jl:143-230) restated, applied to the transcript structures of
                 the reference's own anno/refseq_hg19.flat.gz (committed as a compact fixture, anno/), with '+' and '-'
                 strand genes, retained introns and zero-length nodes as the real annotation has them
  sa_build.cpp   suffix array + BWT of the graph text (FMIndex layout of SURVEY.md section 8c)
"""
