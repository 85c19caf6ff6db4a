"""bgzf (block gzip, SAM spec 4.1: what bgzip / htslib write and what .vcf.gz files are) on the host side:
walking block headers, and a writer used by the tests and the bench to make compressed workloads.
The DEFLATE payloads themselves are inflated on the GPU (csrc/bvcf_inflate.cuh, bvcf_resident_inflate_bgzf)."""
from __future__ import annotations

import struct
import zlib
from concurrent.futures import ThreadPoolExecutor
from typing import Iterator, Tuple

MAGIC = b"\x1f\x8b\x08\x04"
EOF_BLOCK = bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000")
MAX_TEXT = 65280  # text bytes per block (bgzip's choice: the compressed block always fits 64 KiB)


def is_bgzf(head: bytes) -> bool:
    return len(head) >= 18 and head[:4] == MAGIC and head[12:14] == b"BC"


def block_size(buf, p: int) -> int:
    """total size of the bgzf block that starts at buf[p] (0 when its header is not complete yet)"""
    if len(buf) - p < 18:
        return 0
    if bytes(buf[p:p + 4]) != MAGIC:
        raise ValueError("not a bgzf block")
    xlen = buf[p + 10] | (buf[p + 11] << 8)
    if len(buf) - p < 12 + xlen:
        return 0
    q, end = p + 12, p + 12 + xlen
    while q + 4 <= end:
        slen = buf[q + 2] | (buf[q + 3] << 8)
        if buf[q] == 66 and buf[q + 1] == 67 and slen == 2:
            return (buf[q + 4] | (buf[q + 5] << 8)) + 1
        q += 4 + slen
    raise ValueError("gzip member without the bgzf BC subfield")


def _one_block(args) -> bytes:
    chunk, level, strategy = args
    c = zlib.compressobj(level, zlib.DEFLATED, -15, 9, strategy)
    payload = c.compress(chunk) + c.flush()
    bsize = len(payload) + 26
    assert bsize <= 65536
    return (MAGIC + b"\x00\x00\x00\x00\x00\xff\x06\x00BC\x02\x00" + struct.pack("<H", bsize - 1) + payload +
            struct.pack("<II", zlib.crc32(chunk), len(chunk)))


def compress(data, level: int = 6, block_text: int = MAX_TEXT, threads: int = 0, strategy: int = zlib.Z_DEFAULT_STRATEGY,
             eof: bool = True) -> bytes:
    """data -> bgzf bytes (independent 64 KiB blocks; zlib drops the GIL, so blocks are compressed in parallel)"""
    import os

    mv = memoryview(data)
    jobs = [(mv[i:i + block_text], level, strategy) for i in range(0, len(mv), block_text)]
    if len(jobs) < 4:
        parts = [_one_block(j) for j in jobs]
    else:
        with ThreadPoolExecutor(threads or (os.cpu_count() or 1)) as ex:
            parts = list(ex.map(_one_block, jobs, chunksize=16))
    return b"".join(parts) + (EOF_BLOCK if eof else b"")


def inflate_host(buf: bytes, max_text: int) -> bytes:
    """the first max_text text bytes of a bgzf buffer, on the host (zlib): only used to read the VCF preamble"""
    out = []
    n = 0
    p = 0
    while p < len(buf) and n < max_text:
        bs = block_size(buf, p)
        if bs == 0 or p + bs > len(buf):
            break
        xlen = buf[p + 10] | (buf[p + 11] << 8)
        text = zlib.decompress(bytes(buf[p + 12 + xlen:p + bs - 8]), -15)
        out.append(text)
        n += len(text)
        p += bs
    return b"".join(out)
