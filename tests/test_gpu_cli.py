"""The drop-in CLIs (C++ host binary and `python -m bystro_vcf_b200`) on a real GPU: same flags as the
reference, header first, rows in input order, reference-format log lines on stderr."""
import hashlib
import os
import subprocess
import sys

import pytest

import ref_vectors as V

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "bystro_vcf_b200", "bin", "bystro-vcf-b200")


def _run(cmd, data):
    p = subprocess.run(cmd, input=data, stdout=subprocess.PIPE, stderr=subprocess.PIPE, cwd=ROOT, timeout=600)
    assert p.returncode == 0, p.stderr.decode()[-2000:]
    return p.stdout, p.stderr.decode()


@pytest.mark.parametrize("host", ["cpp", "python"])
def test_cli_golden_chr1(chr1_fixture, host):
    cmd = [BIN] if host == "cpp" else [sys.executable, "-m", "bystro_vcf_b200"]
    out, _ = _run(cmd + ["--chunkBytes", str(9 << 20)] if host == "cpp" else cmd, chr1_fixture)
    head, _, body = out.partition(b"\n")
    assert head.decode().split("\t") == V.BASE_HEADER
    assert hashlib.md5(body).hexdigest() == V.GOLDEN_MD5_INPUT_ORDER


def test_cpp_cli_flags_and_files(chr1_fixture, tmp_path):
    inp = tmp_path / "in.vcf"
    inp.write_bytes(chr1_fixture[:40 << 20].rsplit(b"\n", 1)[0] + b"\n")
    outp = tmp_path / "out.tsv"
    _run([BIN, "--in", str(inp), "--out", str(outp), "--keepId", "--keepInfo", "--keepPos=true", "--emptyField", "NA",
          "--fieldDelimiter", "|", "--allowFilter", "*", "--excludeFilter", "q10, LowQual"], b"")
    from oracle import oracle as O

    ref = O.read_vcf(O.OracleConfig(empty_field="NA", field_delim="|", keep_id=True, keep_info=True, keep_pos=True,
                                    allow=None, exclude=["q10", "LowQual"]), inp.read_bytes())
    got = outp.read_bytes()
    head, _, body = got.partition(b"\n")
    assert head.decode().split("\t") == V.BASE_HEADER + ["vcfPos", "id", "alleleIdx", "info"]
    assert body == ref.tsv


def test_cpp_cli_log_lines_match_reference_formats():
    recs = [["1", "5", ".", "A", "A", ".", "PASS", "."], ["1", "6", ".", "A", "<DEL>", ".", "PASS", "."],
            ["1", "7", ".", "AT", "G", ".", "PASS", "."], ["1", "x", ".", "AT", "A", ".", "PASS", "."],
            ["1", "9", ".", "A", "GT,C,N", ".", "PASS", "."], ["1", "10", ".", "TAGCTT", "TAC,T", ".", "PASS", "."],
            ["1", "y", ".", "AT", "A,ATT", ".", "PASS", "."], ["1", "12", ".", "AT", "C,ATT", ".", "PASS", "."]]
    vcf = V._vcf(V.HDR8, recs)
    out, err = _run([BIN], vcf)
    lines = sorted(err.strip().split("\n"))
    # formats per call site: main.go:730 ("%s:%s : %s"), :737/:782/:798 ("ALT #%d"), :835/:934 ("ALT#%d"), :827 (no ALT)
    exp = sorted(["1:5 : REF == ALT", "1:6 ALT #1 ALT not ACTG", "1:7 ALT #1 1st base REF != ALT",
                  "1:x ALT #1 Invalid POS", "1:9 ALT #1 1st base ALT != REF", "1:9 ALT #3 ALT not ACTG",
                  "1:10 ALT#1 Mixed indel/snp sites not supported", "1:y Invalid POS", "1:12 ALT#1 1st base REF != ALT"])
    assert lines == exp


def test_cpp_cli_not_a_vcf():
    p = subprocess.run([BIN], input=b"hello\nworld\n", stdout=subprocess.PIPE, stderr=subprocess.PIPE, cwd=ROOT)
    assert p.returncode == 1 and b"Not a VCF file" in p.stderr
    assert p.stdout.decode().rstrip("\n").split("\t") == V.BASE_HEADER  # header is printed first (main.go:199)
