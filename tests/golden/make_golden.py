"""Regenerates the committed golden fixtures from the reference checkout (this container only:
/root/reference does not exist on the GPU box).

  test_index_flat.npz   flat bundle decoded from /root/reference/test/test_index.jls
  kat_reads.json        the ten known-answer reads of /root/reference/test/runtests.jl:290-330
                        (name encodes CIGAR % leftmost,rightmost coordinate % exon path [:rc])
"""
import json
import os
import re
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
import whippet_jl_b200 as wq  # noqa: E402

REF = "/root/reference"

idx = wq.FlatIndex.from_jls(os.path.join(REF, "test", "test_index.jls"))
idx.save(os.path.join(HERE, "test_index_flat.npz"))

src = open(os.path.join(REF, "test", "runtests.jl")).read()
m = re.search(r'fastq = IOBuffer\("(.*?)"\)', src, re.S)
lines = m.group(1).strip().split("\n")
reads = [dict(name=lines[i][1:], seq=lines[i + 1], qual=lines[i + 3]) for i in range(0, len(lines), 4)]
assert len(reads) == 10
json.dump(dict(source="test/runtests.jl:290-330",
               param=dict(source="test/runtests.jl:283-284", values=[1, 2, 4, 4, 4, 5, 1, 2, 1000, 0.05, 0.7,
                                                                    False, False, True, False, True]),
               fill_ambiguous="C", reads=reads),
          open(os.path.join(HERE, "kat_reads.json"), "w"), indent=1)
print("wrote golden fixtures")
