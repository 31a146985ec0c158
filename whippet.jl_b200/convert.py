"""Index format converter: the reference's `.jls` (Julia Serialization of GraphLib,
bin/whippet-index.jl:97-101) <-> the flat array bundle (`.npz`, what wq_index_create takes) -> the library's native on-disk
index (`.wqx`, wq_index_write / wq_index_load).

    python -m whippet_jl_b200.convert graph.jls graph.wqx      # what a deployment runs once per index
    python -m whippet_jl_b200.convert graph.jls graph.npz
    python -m whippet_jl_b200.convert graph.npz graph.jls      # writer: not authoritative (see jls.py), for round-trip tests

Host only: no GPU is needed for any direction."""
from __future__ import annotations

import sys

from . import jls
from .flat import FlatIndex


def load_any(path: str) -> FlatIndex:
    if path.endswith(".jls"):
        return FlatIndex.from_jls(path)
    if path.endswith(".npz"):
        return FlatIndex.load(path)
    raise ValueError(f"cannot read {path}: expected .jls or .npz")


def convert(src: str, dst: str):
    idx = load_any(src)
    if dst.endswith(".wqx"):
        from .quant import write_index
        write_index(idx, dst)
    elif dst.endswith(".npz"):
        idx.save(dst)
    elif dst.endswith(".jls"):
        with open(dst, "wb") as fh:
            fh.write(jls.dump_jls(jls.graphlib_tree(idx)))
    else:
        raise ValueError(f"cannot write {dst}: expected .wqx, .npz or .jls")
    return idx


def main(argv=None):
    argv = sys.argv[1:] if argv is None else argv
    if len(argv) != 2:
        print(__doc__)
        return 2
    idx = convert(argv[0], argv[1])
    print(f"{argv[0]} -> {argv[1]}: {idx.n_genes} genes, {idx.n_nodes} nodes, {idx.n_isoforms} isoforms, text length {idx.text_len}, K = {idx.kmer}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
