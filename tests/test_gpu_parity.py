"""CUDA path vs CPU oracle / reference vectors.  Every call goes through the C ABI (libbvcf.so)."""
import hashlib
import io

import pytest

import ref_vectors as V

pytestmark = pytest.mark.gpu


def _cfg(**kw):
    from bystro_vcf_b200 import Config

    c = Config()
    m = {"keep_id": "keepID", "keep_info": "keepInfo", "keep_pos": "keepPos"}
    for k, v in kw.items():
        if k == "allow":
            c.allowedFilters = None if v is None else {x: True for x in v}
        elif k == "exclude":
            c.excludedFilters = None if v is None else {x: True for x in v}
        else:
            setattr(c, m[k], v)
    if "allow" not in kw:
        c.allowedFilters = {"PASS": True, ".": True}
    return c


def gpu_rows(vcf: bytes, **kw) -> bytes:
    from bystro_vcf_b200 import read_vcf

    out = io.BytesIO()
    read_vcf(_cfg(**kw), io.BytesIO(vcf), out)
    return out.getvalue()


def oracle_rows(vcf: bytes, **kw) -> bytes:
    from oracle import oracle as O

    return O.read_vcf(O.OracleConfig(**kw), vcf).tsv


def rows_of(tsv: bytes):
    return [r.split("\t") for r in tsv.decode().split("\n")[:-1]] if tsv else []


@pytest.mark.parametrize("name,cfg,vcf,exp", V.STREAM_CASES, ids=[c[0] for c in V.STREAM_CASES])
def test_reference_stream_vectors(name, cfg, vcf, exp):
    got = gpu_rows(vcf, **cfg)
    assert rows_of(got) == exp
    assert got == oracle_rows(vcf, **cfg)


@pytest.mark.parametrize("fields,a,nhom,nhet,nmiss,ac,an", V.GT_VECTORS)
def test_genotype_vectors(fields, a, nhom, nhet, nmiss, ac, an):
    """makeHetHomozygotes vectors (main_test.go:652-951) through a synthetic one-line VCF whose ALT list
    makes `a` an output allele."""
    n_alt = int(a)
    alts = ",".join(["C", "G", "T"][:max(n_alt, 1)])
    hdr = V.HDR8 + ["FORMAT"] + ["S%d" % i for i in range(len(fields))]
    vcf = V._vcf(hdr, [["1", "100", ".", "A", alts, ".", "PASS", ".", "GT"] + fields])
    got = rows_of(gpu_rows(vcf))
    exp = rows_of(oracle_rows(vcf))
    assert got == exp
    row = [r for r in got if r[4] == ["C", "G", "T"][n_alt - 1]]
    if ac == 0:
        assert row == []
    else:
        r = row[0]
        assert (r[12], r[13]) == (str(ac), str(an))
        assert (0 if r[6] == "!" else len(r[6].split(";"))) == nhet
        assert (0 if r[8] == "!" else len(r[8].split(";"))) == nhom
        assert (0 if r[10] == "!" else len(r[10].split(";"))) == nmiss


@pytest.mark.parametrize("inp,exp", V.ALLELE_VECTORS)
def test_allele_vectors(inp, exp):
    pos, ref, alt = inp
    vcf = V._vcf(V.HDR8, [["chr1", pos, ".", ref, alt, ".", "PASS", "."]])
    got = rows_of(gpu_rows(vcf))
    typ, poss, refs, alts, _ = exp
    assert [(r[0], r[1], r[2], r[3], r[4]) for r in got] == [("chr1", p, typ, r_, a_) for p, r_, a_ in zip(poss, refs, alts)]


def test_golden_chr1(chr1_fixture):
    got = gpu_rows(chr1_fixture)
    assert len(got) == 20370676
    assert hashlib.md5(got).hexdigest() == V.GOLDEN_MD5_INPUT_ORDER


def test_golden_chr1_keepid_keepinfo_small_chunks(chr1_fixture):
    from bystro_vcf_b200 import read_vcf

    cfg = _cfg(keep_id=True, keep_info=True)
    cfg.chunkBytes = 7 << 20  # many chunks, lines straddle every read boundary
    out = io.BytesIO()
    st = read_vcf(cfg, io.BytesIO(chr1_fixture), out)
    assert st["n_lines"] == 19747 and st["n_rows"] == 19821
    assert hashlib.md5(out.getvalue()).hexdigest() == V.GOLDEN_MD5_KEEPID_KEEPINFO


def test_query_fixture_matches_oracle(query_fixture):
    for kw in ({}, {"allow": None, "keep_pos": True, "keep_id": True, "keep_info": True}):
        got = gpu_rows(query_fixture, **kw)
        exp = oracle_rows(query_fixture, **kw)
        assert got == exp
