"""CPU tests of the FASTQ readers (host side of bch_count_fastq, reached through the bch_scan_fastq / bch_split_fastq test
hooks): plain text through the block reader and through the mapped-file splitter, gzip with one and with several members,
and bgzip (BGZF) input whose members are inflated in parallel must all yield the same records, in file order
(input.rs:24-89, 115-148)."""
import gzip
import random
import zlib

import pytest

import ngs_barcode_count_b200 as bc
from helpers import bgzf_compress


def make_fastq(n, seed, crlf=False, last_newline=True):
    rng = random.Random(seed)
    recs, crc, bases = [], zlib.crc32(b""), 0
    eol = "\r\n" if crlf else "\n"
    for i in range(n):
        length = rng.randint(1, 151)
        seq = "".join(rng.choice("ACGTN") for _ in range(length))
        qual = "".join(chr(33 + rng.randint(0, 41)) for _ in range(length))
        recs.append(f"@read{i} extra{eol}{seq}{eol}+{eol}{qual}{eol}")
        crc = zlib.crc32(qual.encode(), zlib.crc32(seq.encode(), crc))
        bases += length
    text = "".join(recs)
    if not last_newline:
        text = text[:-len(eol)]
    return text.encode(), (n, bases, crc)


@pytest.mark.parametrize("crlf,last_newline", [(False, True), (True, True), (False, False)])
def test_block_reader_plain_gzip_bgzip_agree(tmp_path, crlf, last_newline):
    data, want = make_fastq(30000, 5, crlf=crlf, last_newline=last_newline)
    plain = tmp_path / "r.fastq"
    plain.write_bytes(data)
    assert bc.scan_fastq(str(plain)) == want
    if not last_newline:
        # the reference's gzip path pops the last character of the last record as "the newline" (input.rs:133-137):
        # without a final newline that is the last quality character
        text = data.decode()
        last_q = text[text.rfind("\n") + 1:]
        recs = text.split("\n")
        crc = zlib.crc32(b"")
        for i in range(0, len(recs), 4):
            q = recs[i + 3] if i + 4 < len(recs) else last_q[:-1]
            crc = zlib.crc32(q.encode(), zlib.crc32(recs[i + 1].encode(), crc))
        want = (want[0], want[1], crc)
    one = tmp_path / "one.fastq.gz"
    one.write_bytes(gzip.compress(data))
    assert bc.scan_fastq(str(one)) == want
    cut = data.find(b"\n@read", len(data) // 3) + 1
    two = tmp_path / "two.fastq.gz"
    two.write_bytes(gzip.compress(data[:cut]) + gzip.compress(data[cut:]))  # MultiGzDecoder semantics (input.rs:63)
    assert bc.scan_fastq(str(two)) == want
    for chunk, threads in ((700, 1), (3000, 3), (65280, 8), (20000, 0)):
        bg = tmp_path / f"bg{chunk}.fastq.gz"
        bg.write_bytes(bgzf_compress(data, chunk=chunk))
        assert bc.scan_fastq(str(bg), threads=threads) == want, (chunk, threads)


def test_block_reader_large_bgzip_spans_blocks(tmp_path):
    """More than the reader's 8 MB block: members are consumed block by block, records straddle block boundaries."""
    data, want = make_fastq(120000, 9)
    assert len(data) > (8 << 20)
    bg = tmp_path / "big.fastq.gz"
    bg.write_bytes(bgzf_compress(data, chunk=60000, level=1))
    assert bc.scan_fastq(str(bg), threads=4) == want


def test_block_reader_rejects_corrupt_bgzip(tmp_path):
    data, _ = make_fastq(5000, 3)
    blob = bytearray(bgzf_compress(data, chunk=10000))
    bad = tmp_path / "bad.fastq.gz"
    flipped = bytearray(blob)
    flipped[len(flipped) // 2] ^= 0x55  # inside a deflate stream: inflate or the CRC has to notice
    bad.write_bytes(bytes(flipped))
    with pytest.raises(bc.BcError):
        bc.scan_fastq(str(bad), threads=2)
    trunc = tmp_path / "trunc.fastq.gz"
    trunc.write_bytes(bytes(blob[:len(blob) // 2]))
    with pytest.raises(bc.BcError):
        bc.scan_fastq(str(trunc), threads=2)


def test_block_reader_rejects_other_extensions(tmp_path):
    p = tmp_path / "reads.txt"
    p.write_bytes(b"@r\nACGT\n+\nIIII\n")
    with pytest.raises(bc.BcError, match="only works with"):
        bc.scan_fastq(str(p))


@pytest.mark.parametrize("n_records", [1, 2, 3, 5, 20, 700])
@pytest.mark.parametrize("last_newline", [True, False])
def test_mapped_splitter_small_files_many_threads(tmp_path, n_records, last_newline):
    """The plain-file path cuts every block into one slice per host thread.  A file with fewer records than threads
    leaves slices empty; the slice that really ends at the end of the file is the one that may lack the final newline
    (the reference's BufRead::lines() yields that last line, input.rs:44)."""
    data, want = make_fastq(n_records, 100 + n_records, last_newline=last_newline)
    p = tmp_path / "r.fastq"
    p.write_bytes(data)
    for threads in (1, 2, 16, 64):
        assert bc.split_fastq(str(p), threads=threads) == want, threads
        assert bc.split_fastq(str(p), threads=threads, min_slice=1) == want, threads  # one slice per thread: most are empty
        assert bc.split_fastq(str(p), threads=threads, block_bytes=2000, min_slice=16) == want, threads


@pytest.mark.parametrize("crlf,last_newline", [(False, True), (True, True), (False, False), (True, False)])
def test_mapped_splitter_equals_block_reader(tmp_path, crlf, last_newline):
    data, want = make_fastq(40000, 21, crlf=crlf, last_newline=last_newline)
    p = tmp_path / "r.fastq"
    p.write_bytes(data)
    assert bc.scan_fastq(str(p)) == want
    for threads, block in ((1, 0), (7, 0), (16, 1 << 20), (64, 300_000), (3, 70_000)):
        assert bc.split_fastq(str(p), threads=threads, block_bytes=block) == want, (threads, block)
        assert bc.split_fastq(str(p), threads=threads, block_bytes=block, min_slice=500) == want, (threads, block)


def test_mapped_splitter_quality_lines_starting_with_at(tmp_path):
    """'@' is a legal quality character (Phred 31): a quality line that starts with it must not be taken for a header."""
    rng = random.Random(4)
    recs, crc, bases = [], zlib.crc32(b""), 0
    for i in range(5000):
        n = rng.randint(20, 60)
        seq = "".join(rng.choice("ACGT") for _ in range(n))
        qual = "@" + "".join(rng.choice("@+IF") for _ in range(n - 1))
        recs.append(f"@r{i}\n{seq}\n+\n{qual}\n")
        crc = zlib.crc32(qual.encode(), zlib.crc32(seq.encode(), crc))
        bases += n
    p = tmp_path / "at.fastq"
    p.write_bytes("".join(recs).encode())
    for threads in (1, 5, 32):
        assert bc.split_fastq(str(p), threads=threads, block_bytes=100_000, min_slice=1000) == (5000, bases, crc)


def test_mapped_splitter_reports_malformed_records(tmp_path):
    good, _ = make_fastq(3000, 8)
    lines = good.decode().split("\n")
    del lines[4001]  # a record in the middle loses its sequence line
    p = tmp_path / "bad.fastq"
    p.write_bytes("\n".join(lines).encode())
    with pytest.raises(bc.BcError):
        for threads in (4, 16):
            bc.split_fastq(str(p), threads=threads, block_bytes=50_000, min_slice=1000)


# ---- the one-pass walker behind bch_count_fastq on plain files (frame + pack a cache-sized chunk at a time) ------------------
def digest_of(data):
    """what bch_walk_fastq reports for a well-formed FASTQ text"""
    lines = data.decode().replace("\r\n", "\n").split("\n")
    n = bases = dig = 0
    for i in range(0, len(lines) - 3, 4):
        seq, qual = lines[i + 1], lines[i + 3]
        dig += zlib.crc32(qual.encode(), zlib.crc32(seq.encode(), zlib.crc32(b"")))
        bases += len(seq)
        n += 1
    return n, bases, dig & ((1 << 64) - 1)


@pytest.mark.parametrize("crlf,last_newline", [(False, True), (True, True), (False, False), (True, False)])
def test_walker_fills_every_row_once(tmp_path, crlf, last_newline):
    data, want = make_fastq(40000, 33, crlf=crlf, last_newline=last_newline)
    p = tmp_path / "r.fastq"
    p.write_bytes(data)
    n, bases, dig = digest_of(data)
    assert (n, bases) == want[:2]
    for threads, chunk, rows in ((1, 0, 1 << 16), (7, 4096, 5000), (16, 700, 37), (64, 100_000, 1000), (3, 64, 3), (16, 300, 1)):
        got = bc.walk_fastq(str(p), threads=threads, chunk_bytes=chunk, batch_rows=rows)
        assert got[:3] == (n, bases, dig), (threads, chunk, rows)
        assert got[3] >= (n + rows - 1) // rows  # batches: full ones, except that an empty tail range may end one early


@pytest.mark.parametrize("n_records", [1, 2, 3, 5, 20, 700])
@pytest.mark.parametrize("last_newline", [True, False])
def test_walker_small_files_many_threads(tmp_path, n_records, last_newline):
    data, _ = make_fastq(n_records, 200 + n_records, last_newline=last_newline)
    p = tmp_path / "r.fastq"
    p.write_bytes(data)
    want = digest_of(data)
    for threads in (1, 2, 16, 64):
        for chunk, rows in ((0, 1 << 16), (64, 4), (200, 1), (1000, 3)):
            assert bc.walk_fastq(str(p), threads=threads, chunk_bytes=chunk, batch_rows=rows)[:3] == want, (threads, chunk, rows)


def test_walker_quality_lines_starting_with_at_and_malformed_records(tmp_path):
    rng = random.Random(4)
    recs = []
    for i in range(5000):
        n = rng.randint(20, 60)
        seq = "".join(rng.choice("ACGT") for _ in range(n))
        qual = "@" + "".join(rng.choice("@+IF") for _ in range(n - 1))
        recs.append(f"@r{i}\n{seq}\n+\n{qual}\n")
    data = "".join(recs).encode()
    p = tmp_path / "at.fastq"
    p.write_bytes(data)
    for threads in (1, 5, 32):
        assert bc.walk_fastq(str(p), threads=threads, chunk_bytes=1000, batch_rows=777)[:3] == digest_of(data)
    good, _ = make_fastq(3000, 8)
    lines = good.decode().split("\n")
    del lines[4001]  # a record in the middle loses its sequence line
    bad = tmp_path / "bad.fastq"
    bad.write_bytes("\n".join(lines).encode())
    with pytest.raises(bc.BcError):
        bc.walk_fastq(str(bad), threads=4, chunk_bytes=5000, batch_rows=500)


@pytest.mark.parametrize("level", [0, 1, 2])
def test_line_end_scanner_simd_levels(tmp_path, level):
    """scalar / AVX2 / AVX-512 line-end scanners frame the same records (CRLF, no final newline, tiny chunks)"""
    lib = bc.lib()
    try:
        lib.bch_set_simd_level(level)
        for crlf, last_newline in ((False, True), (True, False)):
            data, want = make_fastq(6000, 77, crlf=crlf, last_newline=last_newline)
            p = tmp_path / f"r{int(crlf)}.fastq"
            p.write_bytes(data)
            assert bc.split_fastq(str(p), threads=5, block_bytes=200_000, min_slice=500) == want
            for chunk, rows in ((0, 1 << 16), (333, 50), (64, 7)):
                assert bc.walk_fastq(str(p), threads=6, chunk_bytes=chunk, batch_rows=rows)[:3] == digest_of(data)
    finally:
        lib.bch_set_simd_level(2)


@pytest.mark.parametrize("tail", [b"\n", b"\n\n", b"\n\n\n", b"@partial\nACGT\n", b"@partial\nACGT\n+\n", b"\n\n\n\n"])
def test_readers_agree_on_ragged_file_ends(tmp_path, tail):
    """Blank lines and a partial record after the last whole one: the block reader, the mapped-file splitter and the
    one-pass walker frame the same records (the reference takes lines four at a time and never posts an incomplete group,
    input.rs:44-60; four blank lines ARE a group: an empty read)."""
    data, want = make_fastq(1000, 3)
    p = tmp_path / "t.fastq"
    p.write_bytes(data + tail)
    extra = 1 if tail == b"\n\n\n\n" else 0
    a = bc.scan_fastq(str(p))
    b = bc.split_fastq(str(p), threads=4, block_bytes=50_000, min_slice=1000)
    c = bc.walk_fastq(str(p), threads=4, chunk_bytes=5000, batch_rows=300)
    assert a[:2] == b[:2] == c[:2] == (want[0] + extra, want[1])
    assert a[2] == b[2]
