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


def _resident_md5(n_lines, subchunk, first=0, cfg_kw=None, shape="chr1", n_samples=2504, seed=20130502):
    """One resident run of a device-generated workload: (stats, md5 of the TSV, transformer, input length)."""
    from bystro_vcf_b200 import Config, Transformer, synth

    c = Config()
    c.allowedFilters = {"PASS": True, ".": True}
    for k, v in (cfg_kw or {}).items():
        setattr(c, k, v)
    tr = Transformer(c, resident_subchunk_bytes=subchunk)
    tr.set_header(synth.chrom_line(seed, n_samples))
    _, need = synth.device_lines(seed, n_samples, shape, first, n_lines, 0, 0, 0)
    d_in, _ = tr.resident_alloc(need, max(need // 6, 1 << 20))
    got, _ = synth.device_lines(seed, n_samples, shape, first, n_lines, d_in, need, 0)
    assert got == need
    stats, _ = tr.resident_run(need)
    h = hashlib.md5()
    step = 256 << 20
    for off in range(0, stats["out_bytes"], step):
        h.update(tr.resident_download(off, min(step, stats["out_bytes"] - off)))
    return stats, h.hexdigest(), tr, need


def test_c2_quarter_million_lines_vs_oracle():
    """2.5 GB of the BASELINE workload (250,000 variants x 2,504 samples, every tier of the scan kernel, long and
    short rows, multi-allelic records): byte parity with the oracle, compared by md5."""
    import ctypes as C

    import numpy as np

    from bystro_vcf_b200 import _lib, synth
    from oracle import oracle as O

    n = 250_000
    stats, md5, tr, need = _resident_md5(n, 0)
    host = C.c_void_p()
    _lib.check(_lib.lib().bvcf_host_alloc(C.byref(host), need), None, "bvcf_host_alloc")
    try:
        tr.resident_peek(0, need, host.value)
        ref = O.process_block(O.OracleConfig(), synth.chrom_line(20130502, 2504), host.value, need, 1, 16)
        assert stats["n_lines"] == ref.n_lines == n
        assert stats["n_rows"] == ref.n_rows
        assert stats["out_bytes"] == len(ref.tsv)
        assert md5 == hashlib.md5(ref.tsv).hexdigest()
    finally:
        _lib.lib().bvcf_host_free(host)
        tr.close()


def test_c2_subchunking_is_invisible_at_scale():
    """Size-independent property at a scale the oracle is not run on: 1,000,000 variants (10 GB) give the same bytes
    whether the resident run is cut into 1 GiB sub-chunks or taken in one piece."""
    n = 1_000_000
    s1, m1, t1, _ = _resident_md5(n, 1 << 30)
    t1.close()
    s2, m2, t2, _ = _resident_md5(n, 16 << 30)
    t2.close()
    assert s1["n_lines"] == s2["n_lines"] == n
    assert s1["n_rows"] == s2["n_rows"] and s1["out_bytes"] == s2["out_bytes"]
    assert m1 == m2
