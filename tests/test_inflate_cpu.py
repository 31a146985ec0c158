"""WqInflate (csrc/wq_inflate.h: the gzip inflate of the FASTQ reader, src/reads.jl:2-23 hands this to Libz in the reference) against zlib
on the CPU, through the test harness: every compression level and strategy (stored, fixed-Huffman and dynamic blocks, long codes with
sub-tables), FASTQ-like and random and highly repetitive text, outputs larger than the inflater's window, multi-member files, header
fields, and malformed input (truncation at every kind of position, flipped bits: an error or a CRC mismatch, never a crash)."""
import gzip
import io
import zlib

import numpy as np
import pytest

from emu.host_emu import inflate, inflate_raw


def gz(data: bytes, level=6, strategy=zlib.Z_DEFAULT_STRATEGY, wbits=31, memlevel=8) -> bytes:
    c = zlib.compressobj(level, zlib.DEFLATED, wbits, memlevel, strategy)
    return c.compress(data) + c.flush()


def fastq_text(n, seed, readlen=101, nq=8):
    rng = np.random.default_rng(seed)
    s = np.frombuffer(b"ACGT", "u1")[rng.integers(0, 4, (n, readlen))]
    q = np.frombuffer(b"#,:AFGJKLMNOPQRS"[:nq], "u1")[rng.integers(0, nq, (n, readlen))]
    out = bytearray()
    for i in range(n):
        out += b"@instrument:run:flowcell:1:%d:%d:%d 1:N:0:ACGT\n" % (i // 1000, i * 7 % 9973, i) + s[i].tobytes() + b"\n+\n" + q[i].tobytes() + b"\n"
    return bytes(out)


def texts():
    rng = np.random.default_rng(3)
    return {
        "empty": b"",
        "one": b"A",
        "fastq": fastq_text(3000, 1),
        "fastq_40q": fastq_text(1500, 2, readlen=151, nq=16),
        "random": rng.integers(0, 256, 300000, dtype=np.uint8).tobytes(),
        "zeros": bytes(700000),
        "period3": b"abc" * 100000,
        "skewed": rng.choice(np.arange(256, dtype=np.uint8), 400000, p=np.r_[[0.5, 0.25, 0.125], np.full(253, 0.125 / 253)]).tobytes(),
        "words": b" ".join(rng.choice([b"alpha", b"beta", b"gamma", b"delta", b"epsilon", b"zeta"], 60000).tolist()),
    }


@pytest.mark.parametrize("name", list(texts()))
def test_inflate_equals_zlib_all_levels_and_strategies(name):
    data = texts()[name]
    for level in (0, 1, 2, 4, 6, 9):
        for strategy in (zlib.Z_DEFAULT_STRATEGY, zlib.Z_FILTERED, zlib.Z_HUFFMAN_ONLY, zlib.Z_RLE, zlib.Z_FIXED):
            comp = gz(data, level, strategy)
            got, err, members = inflate(comp, piece=70001)
            assert err == "" and members == 1, (name, level, strategy, err)
            assert got == data, (name, level, strategy)


def test_window_slides_and_small_reads():
    """Output far beyond the inflater's 4 MB window, long-distance matches across the slides, and reads of odd sizes."""
    rng = np.random.default_rng(8)
    block = rng.integers(0, 4, 30000, dtype=np.uint8).tobytes()
    data = b"".join(block[(i * 131) % 1000:] + bytes([65 + i % 26]) * (i % 300) for i in range(700))      # ~20 MB with matches 30 k back
    assert len(data) > 3 * (4 << 20)
    for level in (1, 6):
        comp = gz(data, level)
        for piece in (1 << 20, 4097, 333333):
            got, err, _ = inflate(comp, piece=piece, cap=len(data) + 64)
            assert err == "" and got == data, (level, piece)
    small = fastq_text(50, 4)
    got, err, _ = inflate(gz(small), piece=1)
    assert err == "" and got == small


def test_small_windows_and_memory_levels():
    data = fastq_text(2000, 5)
    for wbits in (25, 28, 31):                       # 512-byte to 32 KB history
        for mem in (1, 9):                           # memLevel 1: tiny blocks, many dynamic headers
            got, err, _ = inflate(gz(data, 6, wbits=wbits, memlevel=mem))
            assert err == "" and got == data, (wbits, mem)


def test_members_header_fields_and_trailing_bytes():
    a, b, c = fastq_text(400, 6), b"", fastq_text(300, 7)
    buf = io.BytesIO()
    with gzip.GzipFile(filename="reads_1.fastq", mode="wb", fileobj=buf, mtime=12345) as f:      # FNAME set
        f.write(a)
    named = buf.getvalue()
    extra = bytearray(gz(c))                         # add FEXTRA + FCOMMENT + FHCRC by hand
    hdr = bytes(extra[:10])
    flagged = bytearray(hdr)
    flagged[3] |= 4 | 16 | 2
    body = bytes(flagged) + b"\x05\x00hello" + b"a comment\x00"
    body += (zlib.crc32(body) & 0xFFFF).to_bytes(2, "little") + bytes(extra[10:])
    multi = named + gz(b) + body
    got, err, members = inflate(multi)
    assert err == "" and members == 3 and got == a + b + c
    got, err, members = inflate(multi + bytes(37))   # zero padding after the last member (tape blocks, bgzip -- ignored like gzread does)
    assert err == "" and members == 3 and got == a + b + c
    got, err, members = inflate(multi + b"junk that is not a member")
    assert err == "" and got == a + b + c


def test_malformed_input_is_an_error_not_a_crash():
    data = fastq_text(800, 9)
    comp = gz(data, 6)
    rng = np.random.default_rng(10)
    for cut in [0, 1, 5, 9, 10, 11, 12, 20, 100, len(comp) // 2, len(comp) - 9, len(comp) - 8, len(comp) - 1]:
        got, err, _ = inflate(comp[:cut])
        assert err != "", cut
        assert data.startswith(got), cut              # what was produced before the end of the input is right
    bad = bytearray(comp)
    bad[-5] ^= 0x10                                    # CRC
    assert "CRC" in inflate(bytes(bad))[1]
    bad = bytearray(comp)
    bad[-2] ^= 0x01                                    # ISIZE
    assert "length" in inflate(bytes(bad))[1]
    assert inflate(b"\x1f\x8b\x07" + comp[3:])[1] != ""
    assert inflate(b"plain text, not gzip")[1] != ""
    flips = 0
    for _ in range(300):                               # random bit flips in the deflate stream: error (almost always) or identical text, never a crash
        bad = bytearray(comp)
        pos = int(rng.integers(10, len(comp) - 8))
        bad[pos] ^= 1 << int(rng.integers(0, 8))
        got, err, _ = inflate(bytes(bad), cap=4 * len(data) + 4096)
        assert err != "" or got == data
        flips += err != ""
    assert flips >= 290
    for _ in range(200):                               # random garbage behind a valid header
        junk = comp[:10] + rng.integers(0, 256, int(rng.integers(1, 400)), dtype=np.uint8).tobytes()
        got, err, _ = inflate(junk, cap=1 << 22)
        assert err != ""


def test_raw_stream_into_exact_destination():
    """raw_into (one deflate stream straight into the caller's memory, as the blocked-gzip reader inflates a block between its
    neighbours): right text for block-sized and larger inputs of every kind, nothing written behind the announced length (canary), and
    a wrong announced length, truncation or damage is an error."""
    rng = np.random.default_rng(21)
    cases = [b"", b"A", fastq_text(280, 11)[:65280], fastq_text(30, 12), bytes(5000), b"abc" * 700, rng.integers(0, 256, 3000, dtype=np.uint8).tobytes(),
             fastq_text(2500, 13), b"x" * 299, b"y" * 300, b"z" * 301, fastq_text(2, 14)[:150]]
    for data in cases:
        for level, strategy in ((6, zlib.Z_DEFAULT_STRATEGY), (1, zlib.Z_DEFAULT_STRATEGY), (9, zlib.Z_DEFAULT_STRATEGY), (0, zlib.Z_DEFAULT_STRATEGY),
                                (6, zlib.Z_FIXED), (6, zlib.Z_HUFFMAN_ONLY), (6, zlib.Z_RLE)):
            comp = gz(data, level, strategy, wbits=-15)
            got, err, intact = inflate_raw(comp, len(data))
            assert intact and err == "" and got == data, (len(data), level, strategy, err)
            if len(data) > 1:
                for wrong in (len(data) - 1, len(data) + 1):
                    got, err, intact = inflate_raw(comp, wrong)
                    assert intact and got is None and err != "", (len(data), wrong)
                got, err, intact = inflate_raw(comp[:max(1, len(comp) * 2 // 3)], len(data))
                assert intact and got is None
    data = fastq_text(250, 15)
    comp = gz(data, 6, wbits=-15)
    for _ in range(300):
        bad = bytearray(comp)
        bad[int(rng.integers(0, len(comp)))] ^= 1 << int(rng.integers(0, 8))
        got, err, intact = inflate_raw(bytes(bad), len(data))
        assert intact                                  # no checksum at this level (the reader compares the block's announced length): wrong text is possible, overrun is not


def test_crc32_by_carryless_multiplication_equals_zlib():
    """The trailer check's CRC-32 (wq_crc32: PCLMULQDQ folding for the bulk, zlib for the tail) against zlib.crc32: every length around
    the 16- and 64-byte steps, odd alignments, running values carried from piece to piece."""
    import ctypes as C
    from emu import host_emu
    L = C.CDLL(host_emu.build())
    L.hemu_crc32.restype = C.c_uint32
    rng = np.random.default_rng(31)
    buf = rng.integers(0, 256, 300000, dtype=np.uint8)
    base = buf.ctypes.data
    for n in list(range(0, 200)) + [255, 256, 257, 1000, 4095, 4096, 65537, 299000]:
        for off in (0, 1, 7, 13):
            want = zlib.crc32(buf[off:off + n].tobytes())
            assert L.hemu_crc32(C.c_uint32(0), C.c_void_p(base + off), C.c_uint64(n)) == want, (n, off)
    crc, want, pos = 0, 0, 0
    while pos < len(buf):
        n = int(rng.integers(1, 5000))
        crc = L.hemu_crc32(C.c_uint32(crc), C.c_void_p(base + pos), C.c_uint64(min(n, len(buf) - pos)))
        want = zlib.crc32(buf[pos:pos + n].tobytes(), want)
        pos += n
        assert crc == want
