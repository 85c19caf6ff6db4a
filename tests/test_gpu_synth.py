"""Synthetic BASELINE.json configs (C2..C5) at oracle-sized scales: device-resident CUDA path vs CPU oracle."""
import hashlib

import pytest

pytestmark = pytest.mark.gpu


def _run_resident(cfg_kw, seed, n_samples, shape, n_lines, oracle_kw, subchunk=0, first=0):
    from bystro_vcf_b200 import Config, Transformer, synth
    from oracle import oracle as O

    c = Config()
    c.allowedFilters = {"PASS": True, ".": True}
    for k, v in cfg_kw.items():
        setattr(c, k, v)
    tr = Transformer(c, resident_subchunk_bytes=subchunk)
    chrom = synth.chrom_line(seed, n_samples)
    tr.set_header(chrom)
    _, need = synth.device_lines(seed, n_samples, shape, first, n_lines, 0, 0, 0)
    d_in, _ = tr.resident_alloc(need, max(need // 4, 1 << 20))
    got, _ = synth.device_lines(seed, n_samples, shape, first, n_lines, d_in, need, 0)
    assert got == need
    stats, times = tr.resident_run(need)
    out = tr.resident_download(0, stats["out_bytes"])
    body = tr.resident_peek(0, need)
    # the device generator and the host generator are the same function of (seed, line)
    assert body == synth.host_lines(seed, n_samples, shape, first, n_lines)
    ref = O.process_block(O.OracleConfig(**oracle_kw), chrom, body, threads=8)
    assert stats["n_lines"] == ref.n_lines == n_lines
    assert stats["n_rows"] == ref.n_rows
    assert hashlib.md5(out).hexdigest() == hashlib.md5(ref.tsv).hexdigest()
    assert out == ref.tsv
    tr.close()
    return stats, times


def test_c2_chr1_shape():
    stats, _ = _run_resident({}, 20130502, 2504, "chr1", 20000, {})
    assert stats["n_rows"] > 19000


def test_c2_many_subchunks_and_offset_shard():
    # sub-chunks of 8 MiB: lines straddle every sub-chunk boundary; shard starting at line 1,000,000
    _run_resident({}, 20130502, 2504, "chr1", 6000, {}, subchunk=8 << 20, first=1_000_000)


def test_c3_sites_only():
    stats, _ = _run_resident({}, 50, 0, "sites", 300000, {})
    assert stats["n_rows"] > 300000


def test_c4_biobank_width_slice():
    # 200,000 samples, 2 % missing: E-notation floats, 6-digit ac/an, ~4,000-name missing lists
    _run_resident({}, 200000, 200000, "biobank", 24, {})


def test_c5_filters_keepinfo():
    _run_resident({"keepInfo": True, "allowedFilters": None, "excludedFilters": {"LowQual": True}}, 20130502, 2504,
                  "chr1_filters", 8000, {"keep_info": True, "allow": None, "exclude": ["LowQual"]})
    _run_resident({"keepInfo": True, "keepID": True}, 20130502, 2504, "chr1_filters", 4000,
                  {"keep_info": True, "keep_id": True})
