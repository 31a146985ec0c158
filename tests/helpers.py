"""Test-side helpers (restatements of reference output code used only to check known answers)."""
import numpy as np


def cigar_string(aln, index, gene, strand, readlen):
    """cigar_string, /root/reference/src/io.jl:111-147.  `aln` = (offset, offsetread, strand, isvalid, path)
    with path entries (gene,node,matches,mismatches,mistol_bits).  Returns (cigar, curpos)."""
    offset, offsetread, _, _, path = aln
    nb = int(index.node_base[gene - 1])
    nodecoord = lambda i: int(index.nodecoord[nb + i - 1])
    nodelen = lambda i: int(index.nodelen[nb + i - 1])
    nodeoffset = lambda i: int(index.nodeoffset[nb + i - 1])
    curpos = offset
    running = 0
    total = offsetread - 1
    pad = offsetread - 1
    cigar = [f"{pad}S" if pad > 0 else ""]
    for idx in range(len(path)):
        m = path[idx][2] + path[idx][3]
        total += m
        curpos += m
        if idx < len(path) - 1:
            curi, nexti = path[idx][1], path[idx + 1][1]
            if strand:
                intron = nodecoord(nexti) - (nodecoord(curi) + nodelen(curi) - 1) - 1
            else:
                intron = nodecoord(curi) - (nodecoord(nexti) + nodelen(nexti) - 1) - 1
            if intron >= 1:
                cigar.append(f"{running + m}M")
                cigar.append(f"{intron}N")
                running = 0
                curpos = nodeoffset(nexti)
            else:
                running += m
        else:
            cigar.append(f"{running + m}M")
    if total < readlen:
        cigar.append(f"{readlen - total}S")
    if not strand:
        cigar.reverse()
    return "".join(cigar), curpos


def identity(aln, readlen):
    """identity(), src/align.jl:104-114 in Float32."""
    s = np.float32(0)
    for (_, _, m, _, tolbits) in aln[4]:
        s = np.float32(s + (np.float32(m) - np.uint32(tolbits).view(np.float32)))
    return np.float32(s / np.float32(readlen))
