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


# ---- SAM output (test-side restatement of /root/reference/src/io.jl:69-271, used to check wq_sam_batch) ----
def _score(aln):
    """score(align), src/align.jl:90-112: sum over the path of matches - mistolerance in Float32."""
    s = np.float32(0)
    for (_, _, m, _, tolbits) in aln[4]:
        s = np.float32(s + (np.float32(m) - np.uint32(tolbits).view(np.float32)))
    return s


def _revcomp(seq):
    return seq[::-1].translate(str.maketrans("ACGT", "TGCA"))


def write_sam_one(index, name, read, aln, mapq=0, paired=False, fwd_mate=True, supplemental=False, qualoffset=33):
    """write_sam for one alignment, src/io.jl:151-190.  `read` = [seq, qual bytes] (mutated in place like the
    reference's FASTQRecord).  Returns the line or ''."""
    offset, offsetread, _, isvalid, path = aln
    if not (isvalid and len(path) >= 1):
        return ""
    gene = path[0][0]
    strand = bool(index.gene_strand[gene - 1])
    nodeind = path[0][1] if strand else path[-1][1]
    if path[-1][1] < nodeind or path[-1][1] < path[0][1]:
        return ""
    cigar, endpos = cigar_string(aln, index, gene, strand, len(read[0]))
    nb = int(index.node_base[gene - 1])
    noff, nlen = int(index.nodeoffset[nb + nodeind - 1]), int(index.nodelen[nb + nodeind - 1])
    off = (offset - noff) if strand else (nlen - (endpos - noff))
    off &= 0xFFFFFFFF
    if not strand and not supplemental:
        read[0], read[1] = _revcomp(read[0]), read[1][::-1]
    if paired:
        flag = 0x01 | 0x02 | (0x40 if fwd_mate else 0x80) | (0 if strand else 0x10)
    else:
        flag = 0 if strand else 0x10
    if supplemental:
        flag |= 0x100
    pos = (int(index.nodecoord[nb + nodeind - 1]) + off) & 0xFFFFFFFF
    ident = name.split(" ")[0].split("\t")[0]
    fields = [ident, str(flag), index.gene_seqname[gene - 1], str(pos), str(mapq), cigar, "*", "0", "0"]
    if supplemental:
        fields += ["*", "*"]
    else:
        fields += [read[0], bytes((q + qualoffset) & 0xFF for q in read[1]).decode("latin-1")]
    return "\t".join(fields) + "\n"


def write_sam_reads(index, names, seqs, quals, res, mate=None, qualoffset=33):
    """process_reads! / process_paired_reads! with sam=true (src/reads.jl:59-67, 106-121; src/io.jl:221-258).
    seqs/quals: FASTQ sequences (already N-filled, upper case) and phred byte lists; mate = (names, seqs, quals)."""
    out = []
    for i in range(len(names)):
        alns = res.read(i)
        if not alns:
            continue
        r1 = [seqs[i], list(quals[i])]
        if res.flipped[i] & 1:
            r1 = [_revcomp(r1[0]), r1[1][::-1]]
        if mate is None:
            if len(alns) == 1:
                out.append(write_sam_one(index, names[i], r1, alns[0], qualoffset=qualoffset))
                continue
            sc = [_score(a) for a in alns]
            best = 0
            for k in range(1, len(alns)):
                if sc[k] > sc[best]:
                    best = k
            mq = 10 * len(alns)
            out.append(write_sam_one(index, names[i], r1, alns[best], mapq=mq, qualoffset=qualoffset))
            for k in range(len(alns)):
                if k != best:
                    out.append(write_sam_one(index, names[i], r1, alns[k], mapq=mq, supplemental=True, qualoffset=qualoffset))
            continue
        mnames, mseqs, mquals = mate
        mal = res.read(i, mate=True)
        r2 = [mseqs[i], list(mquals[i])]
        if res.flipped[i] & 2:
            r2 = [_revcomp(r2[0]), r2[1][::-1]]
        if len(alns) == 1:
            out.append(write_sam_one(index, names[i], r1, alns[0], paired=True, fwd_mate=True, qualoffset=qualoffset))
            out.append(write_sam_one(index, mnames[i], r2, mal[0], paired=True, fwd_mate=False, qualoffset=qualoffset))
            continue
        sc = [np.float32(_score(a) + _score(b)) for a, b in zip(alns, mal)]
        best = 0
        for k in range(1, len(alns)):
            if sc[k] > sc[best]:
                best = k
        mq = 10 * len(alns)
        out.append(write_sam_one(index, names[i], r1, alns[best], mapq=mq, paired=True, fwd_mate=True, qualoffset=qualoffset))
        out.append(write_sam_one(index, mnames[i], r2, mal[best], mapq=mq, paired=True, fwd_mate=False, qualoffset=qualoffset))
        for k in range(len(alns)):
            if k != best:
                out.append(write_sam_one(index, names[i], r1, alns[k], mapq=mq, paired=True, fwd_mate=True, supplemental=True, qualoffset=qualoffset))
                out.append(write_sam_one(index, mnames[i], r2, mal[k], mapq=mq, paired=True, fwd_mate=True, supplemental=True, qualoffset=qualoffset))
    return "".join(out)
