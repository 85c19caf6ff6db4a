"""Multi-GPU sharding of the per-line transform.

Lines are independent, so the data region of a VCF is cut into N newline-aligned byte ranges, one per GPU;
each GPU runs the same kernel pipeline on its range with its own host thread and streams, and the host
writes the shard outputs back in shard order == input order.  There is no cross-shard reduction, hence no
collective (no NCCL): the only thing every shard needs is the small constant context (config + header).

`partition` is pure host logic and is what the world_size-2 gloo test exercises on CPU.
"""
from __future__ import annotations

import threading
from typing import BinaryIO, List, Optional, Sequence, Tuple

from .host import Config, Transformer, parse_preamble


def partition(data, begin: int, end: int, n: int) -> List[Tuple[int, int]]:
    """Split data[begin:end] into n byte ranges of ~equal size whose boundaries follow a newline.

    `data` only needs find(b"\\n", lo, hi) (bytes, bytearray, mmap).  Ranges are contiguous, cover [begin, end)
    exactly, and every range except possibly the last ends just after a b"\\n" (so does the last one when the
    input is newline-terminated).  Empty ranges are possible when there are fewer lines than shards."""
    cuts = [begin]
    for k in range(1, n):
        target = begin + (end - begin) * k // n
        target = max(target, cuts[-1])
        if target >= end:
            cuts.append(end)
            continue
        if target == begin or data[target - 1:target] == b"\n":
            cuts.append(target)
            continue
        nl = data.find(b"\n", target, end)
        cuts.append(end if nl < 0 else nl + 1)
    cuts.append(end)
    return [(cuts[i], cuts[i + 1]) for i in range(n)]


def chunk_ranges(data, lo: int, hi: int, chunk_bytes: int) -> List[Tuple[int, int]]:
    """Newline-aligned sub-chunks of one shard (what bvcf_submit is fed)."""
    out = []
    p = lo
    while p < hi:
        q = min(p + chunk_bytes, hi)
        if q < hi:
            nl = data.rfind(b"\n", p, q)
            if nl < 0:  # a single line longer than the chunk: extend to its end
                nl = data.find(b"\n", q, hi)
                q = hi if nl < 0 else nl + 1
            else:
                q = nl + 1
        out.append((p, q))
        p = q
    return out


def _run_shard(cfg: Config, device: int, eol_width: int, chrom_line: bytes, data, ranges, out: list, err: list):
    try:
        c = Config(**{**cfg.__dict__, "device": device})
        with Transformer(c, eol_width=eol_width, max_chunk_bytes=max(r[1] - r[0] for r in ranges) + 1 if ranges else 0) as tr:
            tr.set_header(chrom_line)
            seq_in = seq_out = 0
            held = {}
            while seq_out < len(ranges):
                while seq_in < len(ranges) and seq_in - seq_out < tr.n_slots:
                    lo, hi = ranges[seq_in]
                    held[seq_in] = bytes(data[lo:hi])
                    tr.submit(seq_in, held[seq_in])
                    seq_in += 1
                res = tr.collect(seq_out)
                held.pop(seq_out, None)
                out.append(res)
                seq_out += 1
    except Exception as e:  # surfaced by the caller
        err.append(e)


def read_vcf_multi(config: Config, data, writer: Optional[BinaryIO], devices: Sequence[int]) -> dict:
    """readVcf (main.go:241-396) over several GPUs of one box.  `data` is the whole uncompressed VCF as a
    bytes-like object or mmap.  Rows are written in input order."""
    width, chrom_line, off = parse_preamble(bytes(data[:min(len(data), 64 << 20)]))
    end = len(data)
    # an unterminated last line is dropped (main.go:354-357)
    last_nl = data.rfind(b"\n", off, end)
    end = off if last_nl < 0 else last_nl + 1
    shards = partition(data, off, end, len(devices))
    results = [[] for _ in devices]
    errors: list = []
    threads = []
    for i, dev in enumerate(devices):
        lo, hi = shards[i]
        ranges = chunk_ranges(data, lo, hi, max(int(config.chunkBytes), 1 << 16))
        t = threading.Thread(target=_run_shard, args=(config, dev, width, chrom_line, data, ranges, results[i], errors))
        t.start()
        threads.append(t)
    for t in threads:
        t.join()
    if errors:
        raise errors[0]
    totals = {"n_lines": 0, "n_records": 0, "n_rows": 0, "out_bytes": 0, "in_bytes": end - off, "shards": shards}
    for shard in results:  # shard order == input order
        for res in shard:
            if writer is not None and not config.noOut:
                writer.write(res.tsv)
            totals["n_lines"] += res.n_lines
            totals["n_records"] += res.n_records
            totals["n_rows"] += res.n_rows
            totals["out_bytes"] += len(res.tsv)
    return totals
