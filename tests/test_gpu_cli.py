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


def test_python_cli_log_lines_match_reference_formats(tmp_path):
    """the Python CLI used to print "line N ALT #k msg"; now the reference's formats, and --err is honoured"""
    recs = [["1", "5", ".", "A", "A", ".", "PASS", "."], ["1", "6", ".", "A", "<DEL>", ".", "PASS", "."],
            ["1", "10", ".", "TAGCTT", "TAC,T", ".", "PASS", "."], ["1", "y", ".", "AT", "A,ATT", ".", "PASS", "."]]
    vcf = V._vcf(V.HDR8, recs)
    errp = tmp_path / "err.log"
    out, err = _run([sys.executable, "-m", "bystro_vcf_b200", "--err", str(errp)], vcf)
    assert sorted(errp.read_text().strip().split("\n")) == sorted(
        ["1:5 : REF == ALT", "1:6 ALT #1 ALT not ACTG", "1:10 ALT#1 Mixed indel/snp sites not supported", "1:y Invalid POS"])


def _tiled(chr1_fixture, tiles, first_mb=12):
    """header + the first `first_mb` MB of the fixture's data lines, `tiles` times over"""
    from bystro_vcf_b200 import parse_preamble

    _, _, off = parse_preamble(chr1_fixture)
    body = chr1_fixture[off:off + (first_mb << 20)].rsplit(b"\n", 1)[0] + b"\n"
    return chr1_fixture[:off] + body * tiles


def _devices():
    """two workers: two GPUs when the box has them, the same GPU twice otherwise"""
    import torch

    return [0, 1] if torch.cuda.device_count() >= 2 else [0, 0]


@pytest.mark.parametrize("host", ["cpp-mmap", "cpp-pipe", "python"])
def test_multi_gpu_hosts_match_single(chr1_fixture, tmp_path, host):
    """the multi-GPU product hosts (chunk k -> GPU k mod N, rows written in input order) give the
    single-GPU bytes: C++ binary with a memory-mapped --in file and with a pipe, shard.read_vcf_multi."""
    import io

    from bystro_vcf_b200 import Config, read_vcf
    from bystro_vcf_b200.shard import read_vcf_multi

    vcf = _tiled(chr1_fixture, 8)
    c = Config()
    c.allowedFilters = {"PASS": True, ".": True}
    c.keepID = c.keepInfo = True
    c.chunkBytes = 7 << 20
    single = io.BytesIO()
    st = read_vcf(c, io.BytesIO(vcf), single)
    devs = _devices()
    if host == "python":
        out = io.BytesIO()
        st2 = read_vcf_multi(c, vcf, out, devs)
        assert st2["n_rows"] == st["n_rows"] and st2["n_chunks"] > 8
        got = out.getvalue()
    else:
        flags = ["--keepId", "--keepInfo", "--chunkBytes", str(7 << 20), "--devices", ",".join(map(str, devs))]
        if host == "cpp-mmap":
            inp = tmp_path / "in.vcf"
            inp.write_bytes(vcf)
            raw, _ = _run([BIN, "--in", str(inp)] + flags, b"")
        else:
            raw, _ = _run([BIN] + flags, vcf)
        got = raw.partition(b"\n")[2]
    assert hashlib.md5(got).hexdigest() == hashlib.md5(single.getvalue()).hexdigest()
    assert len(got) == st["out_bytes"]


@pytest.mark.parametrize("host", ["cpp", "python-multi"])
def test_dosage_output_file_from_the_hosts(chr1_fixture, tmp_path, host):
    """--dosageOutput in the drop-in binary (libarrow of the pyarrow wheel) and on the multi-GPU
    Python path: an Arrow IPC file, zstd, `locus` + one non-nullable int8 column per sample, batches of at most 5,000
    rows, read back like main_test.go:2947-2976 and compared with the oracle's matrix."""
    import numpy as np
    import pyarrow as pa

    from oracle import oracle as O

    vcf = _tiled(chr1_fixture, 6, first_mb=20)  # ~11,800 rows: three batches
    ref = O.read_vcf(O.OracleConfig(want_dosage=True), vcf)
    path = tmp_path / "dosage.feather"
    devs = _devices()
    if host == "cpp":
        raw, _ = _run([BIN, "--dosageOutput", str(path), "--chunkBytes", str(16 << 20), "--devices", ",".join(map(str, devs))], vcf)
        assert raw.partition(b"\n")[2] == ref.tsv
    else:
        import io

        from bystro_vcf_b200 import Config
        from bystro_vcf_b200.shard import read_vcf_multi

        c = Config()
        c.allowedFilters = {"PASS": True, ".": True}
        c.dosageMatrixOutPath = str(path)
        c.chunkBytes = 16 << 20
        out = io.BytesIO()
        read_vcf_multi(c, vcf, out, devs)
        assert out.getvalue() == ref.tsv
    rd = pa.ipc.open_file(str(path))
    assert rd.schema.names[0] == "locus" and len(rd.schema.names) == 2505
    assert all(not f.nullable for f in rd.schema) and rd.schema.field(1).type == pa.int8()
    assert rd.num_record_batches >= 3 and all(rd.get_batch(i).num_rows <= 5000 for i in range(rd.num_record_batches))
    tab = rd.read_all()
    assert [x.encode() for x in tab.column(0).to_pylist()] == ref.loci
    mat = np.stack([tab.column(j + 1).to_numpy() for j in range(2504)], axis=1)
    assert np.array_equal(mat, ref.dosage)


def test_cpp_cli_noout_dosage_only_and_sample_list(chr1_fixture, tmp_path):
    """--noOut with --dosageOutput (main.go:160-166), no samples -> empty dosage file (main.go:308-318)"""
    import pyarrow as pa

    vcf = _tiled(chr1_fixture, 1, first_mb=6)
    path = tmp_path / "d.feather"
    raw, _ = _run([BIN, "--noOut", "--dosageOutput", str(path)], vcf)
    assert raw == b""
    assert pa.ipc.open_file(str(path)).read_all().num_rows > 500
    sites = V._vcf(V.HDR8, [["1", "5", ".", "A", "G", ".", "PASS", "."]])
    p2 = tmp_path / "e.feather"
    raw, err = _run([BIN, "--dosageOutput", str(p2)], sites)
    assert p2.read_bytes() == b"" and raw.count(b"\n") == 2 and "No samples found" in err


def test_file_to_file_wall_clock(chr1_fixture, tmp_path):
    """the whole main.go:134-217 contract timed around the binary -- file in (tmpfs when there is one),
    file out, read + staging + H2D + kernels + D2H + write; the number is printed for the log."""
    import time

    d = "/dev/shm" if os.path.isdir("/dev/shm") and os.access("/dev/shm", os.W_OK) else str(tmp_path)
    inp, outp = os.path.join(d, "bvcf_f2f_in.vcf"), os.path.join(d, "bvcf_f2f_out.tsv")
    vcf = _tiled(chr1_fixture, 10, first_mb=100)  # ~1 GB
    try:
        with open(inp, "wb") as f:
            f.write(vcf)
        if os.path.exists(outp):
            os.remove(outp)
        t0 = time.perf_counter()
        _run([BIN, "--in", inp, "--out", outp], b"")
        dt = time.perf_counter() - t0
        n_lines = vcf.count(b"\n")
        print("file->file: %.2f GB in %.2f s = %.2f GB/s, %.2f M variants/s (context creation included)"
              % (len(vcf) / 1e9, dt, len(vcf) / 1e9 / dt, n_lines / 1e6 / dt))
        from bystro_vcf_b200 import parse_preamble

        _, _, off = parse_preamble(vcf)
        one = oracle_body = None
        from oracle import oracle as O

        one = O.read_vcf(O.OracleConfig(), vcf[:off] + vcf[off:off + (100 << 20)].rsplit(b"\n", 1)[0] + b"\n").tsv
        got = open(outp, "rb").read().partition(b"\n")[2]
        assert got == one * 10
    finally:
        for p in (inp, outp):
            if os.path.exists(p):
                os.remove(p)


@pytest.mark.parametrize("how", ["cpp-file", "cpp-pipe", "python"])
def test_cli_bgzf_input(chr1_fixture, tmp_path, how):
    """`--in x.vcf.gz` / a bgzf pipe: the compressed bytes go to the GPU, rows equal the plain path's"""
    from bystro_vcf_b200 import bgzf

    vcf = chr1_fixture[:50 << 20].rsplit(b"\n", 1)[0] + b"\n"
    comp = bgzf.compress(vcf)
    flags = ["--keepId", "--keepInfo"]
    plain, _ = _run([BIN] + flags, vcf)
    if how == "cpp-file":
        p = tmp_path / "in.vcf.gz"
        p.write_bytes(comp)
        got, _ = _run([BIN, "--in", str(p), "--chunkBytes", str(2 << 20)] + flags, b"")
    elif how == "cpp-pipe":
        got, _ = _run([BIN] + flags, comp)
    else:
        got, _ = _run([sys.executable, "-m", "bystro_vcf_b200"] + flags, comp)
    assert got == plain


def test_bgzf_input_with_dosage_and_diagnostics(tmp_path):
    """the bgzf path is the whole transform: rows, the Arrow dosage file and the reference's log lines"""
    import numpy as np
    import pyarrow as pa

    from bystro_vcf_b200 import bgzf
    from oracle import oracle as O

    n = 40
    hdr = V.HDR8 + ["FORMAT"] + ["SM%05d" % i for i in range(n)]
    recs = []
    for i in range(3000):
        gts = ["0|1" if (i + j) % 17 == 0 else ("1|1" if (i * j) % 29 == 1 else "0|0") for j in range(n)]
        alt = "<DEL>" if i % 500 == 3 else ("G,T" if i % 7 == 0 else "G")
        recs.append(["1", str(100 + i), "rs%d" % i, "A", alt, ".", "PASS", "DP=%d" % i, "GT"] + gts)
    vcf = V._vcf(hdr, recs)
    ref = O.read_vcf(O.OracleConfig(want_dosage=True), vcf)
    comp = bgzf.compress(vcf, block_text=20000)
    for host in ("cpp", "python"):
        path = tmp_path / ("d_%s.feather" % host)
        cmd = [BIN] if host == "cpp" else [sys.executable, "-m", "bystro_vcf_b200"]
        raw, err = _run(cmd + ["--dosageOutput", str(path)], comp)
        assert raw.partition(b"\n")[2] == ref.tsv
        tab = pa.ipc.open_file(str(path)).read_all()
        assert [x.encode() for x in tab.column(0).to_pylist()] == ref.loci
        assert np.array_equal(np.stack([tab.column(j + 1).to_numpy() for j in range(n)], axis=1), ref.dosage)
        assert sorted(err.strip().split("\n")) == sorted("1:%d ALT #1 ALT not ACTG" % (100 + i) for i in range(3000) if i % 500 == 3)


@pytest.mark.parametrize("host", ["cpp", "python"])
@pytest.mark.parametrize("compressed_in", [False, True])
def test_bgzf_out_is_the_whole_pipeline(chr1_fixture, tmp_path, host, compressed_in):
    """--bgzfOut: header + rows leave as bgzf blocks deflated on the GPU; with .vcf.gz input the one command stands for
    `pigz -d -c in.vcf.gz | bystro-vcf ... | pigz -c` (README.md:10).  gzip reads the file back to the plain rows."""
    import gzip

    from bystro_vcf_b200 import bgzf

    vcf = chr1_fixture[:30 << 20].rsplit(b"\n", 1)[0] + b"\n"
    flags = ["--keepId"]
    plain, _ = _run([BIN] + flags, vcf)
    cmd = [BIN] if host == "cpp" else [sys.executable, "-m", "bystro_vcf_b200"]
    out = tmp_path / "rows.tsv.gz"
    data = bgzf.compress(vcf) if compressed_in else vcf
    got, _ = _run(cmd + flags + ["--bgzfOut", "--out", str(out)], data)
    assert got == b""
    raw = out.read_bytes()
    assert raw.endswith(bgzf.EOF_BLOCK) and bgzf.is_bgzf(raw)
    assert gzip.decompress(raw) == plain
    assert len(raw) < len(plain) * 0.9
    # every member is a bgzf block of at most 64 KiB
    p = n = 0
    while p < len(raw):
        bs = bgzf.block_size(raw, p)
        assert 0 < bs <= 65536
        p += bs
        n += 1
    assert p == len(raw) and n > 10


def test_bgzf_out_refuses_several_gpus(tmp_path):
    for cmd in ([BIN], [sys.executable, "-m", "bystro_vcf_b200"]):
        r = subprocess.run(cmd + ["--bgzfOut", "--gpus", "2"], input=b"##fileformat=VCFv4.1\n", capture_output=True, cwd=ROOT)
        assert r.returncode == 1 and b"--bgzfOut runs on one GPU" in r.stderr
