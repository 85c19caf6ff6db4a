"""ctypes wrapper over oracle/liboracle.so -- the CPU ORACLE (test infrastructure only).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The product package (bystro_vcf_b200) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass, field
from typing import List, Optional, Sequence

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "bvcf_oracle.c")
    hdr = os.path.join(_HERE, "bvcf_oracle.h")
    if (
        force
        or not os.path.exists(_LIB_PATH)
        or os.path.getmtime(_LIB_PATH) < max(os.path.getmtime(src), os.path.getmtime(hdr))
    ):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _LIB_PATH


class _Config(C.Structure):
    _fields_ = [
        ("empty_field", C.c_char_p),
        ("field_delim", C.c_char_p),
        ("keep_id", C.c_int),
        ("keep_info", C.c_int),
        ("keep_pos", C.c_int),
        ("allow", C.POINTER(C.c_char_p)),
        ("n_allow", C.c_int),
        ("exclude", C.POINTER(C.c_char_p)),
        ("n_exclude", C.c_int),
        ("want_tsv", C.c_int),
        ("want_dosage", C.c_int),
        ("normalize_dots", C.c_int),
    ]


class _Diag(C.Structure):
    _fields_ = [("line_no", C.c_uint64), ("alt_no", C.c_int32), ("code", C.c_int32)]


class _Result(C.Structure):
    _fields_ = [
        ("tsv", C.POINTER(C.c_char)),
        ("tsv_len", C.c_size_t),
        ("n_rows", C.c_uint64),
        ("n_lines", C.c_uint64),
        ("loci", C.POINTER(C.c_char)),
        ("loci_len", C.c_size_t),
        ("dosage", C.POINTER(C.c_int8)),
        ("n_dosage_rows", C.c_uint64),
        ("n_samples", C.c_uint32),
        ("diags", C.POINTER(_Diag)),
        ("n_diags", C.c_size_t),
        ("error", C.c_int),
    ]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        L.oracle_header.restype = C.c_void_p
        L.oracle_header.argtypes = [C.POINTER(_Config)]
        L.oracle_read_vcf.restype = C.c_int
        L.oracle_read_vcf.argtypes = [C.POINTER(_Config), C.c_char_p, C.c_size_t, C.c_int, C.POINTER(_Result)]
        L.oracle_process_block.restype = C.c_int
        L.oracle_process_block.argtypes = [
            C.POINTER(_Config), C.c_char_p, C.c_size_t, C.c_int, C.c_void_p, C.c_size_t, C.c_int,
            C.POINTER(_Result),
        ]
        L.oracle_free_result.argtypes = [C.POINTER(_Result)]
        L.oracle_get_alleles.restype = C.c_int
        L.oracle_alt_is_valid.restype = C.c_int
        L.oracle_alt_is_valid.argtypes = [C.c_char_p, C.c_size_t]
        L.oracle_format_float.argtypes = [C.c_double, C.c_char_p]
        L.oracle_tr_tv.restype = C.c_char_p
        L.oracle_tr_tv.argtypes = [C.c_char, C.c_char_p, C.c_size_t]
        _lib = L
    return _lib


@dataclass
class OracleConfig:
    """Mirror of the reference's Config fields read by processLines (main.go:494-503)."""

    empty_field: str = "!"
    field_delim: str = ";"
    keep_id: bool = False
    keep_info: bool = False
    keep_pos: bool = False
    allow: Optional[Sequence[str]] = ("PASS", ".")  # None => nil map => allow all
    exclude: Optional[Sequence[str]] = None
    want_tsv: bool = True
    want_dosage: bool = False
    normalize_dots: bool = True

    def _c(self):
        keep = []
        c = _Config()
        c.empty_field = self.empty_field.encode()
        c.field_delim = self.field_delim.encode()
        c.keep_id, c.keep_info, c.keep_pos = int(self.keep_id), int(self.keep_info), int(self.keep_pos)
        if self.allow is None:
            c.n_allow = -1
            c.allow = None
        else:
            arr = (C.c_char_p * max(1, len(self.allow)))(*[a.encode() for a in self.allow])
            keep.append(arr)
            c.allow = arr
            c.n_allow = len(self.allow)
        if not self.exclude:
            c.n_exclude = 0
            c.exclude = None
        else:
            arr = (C.c_char_p * len(self.exclude))(*[a.encode() for a in self.exclude])
            keep.append(arr)
            c.exclude = arr
            c.n_exclude = len(self.exclude)
        c.want_tsv = int(self.want_tsv)
        c.want_dosage = int(self.want_dosage)
        c.normalize_dots = int(self.normalize_dots)
        return c, keep


@dataclass
class OracleResult:
    tsv: bytes = b""
    n_rows: int = 0
    n_lines: int = 0
    loci: List[bytes] = field(default_factory=list)
    dosage: Optional["object"] = None  # numpy int8 [rows, samples]
    diags: List[tuple] = field(default_factory=list)
    error: int = 0


def _unpack(r: _Result) -> OracleResult:
    import numpy as np

    out = OracleResult()
    out.error = r.error
    out.tsv = C.string_at(r.tsv, r.tsv_len) if r.tsv_len else b""
    out.n_rows = r.n_rows
    out.n_lines = r.n_lines
    if r.loci_len:
        out.loci = C.string_at(r.loci, r.loci_len).split(b"\n")[:-1]
    if r.n_dosage_rows and r.n_samples:
        n = r.n_dosage_rows * r.n_samples
        buf = C.string_at(r.dosage, n)
        out.dosage = np.frombuffer(buf, dtype=np.int8).reshape(r.n_dosage_rows, r.n_samples).copy()
    out.diags = [(r.diags[i].line_no, r.diags[i].alt_no, r.diags[i].code) for i in range(r.n_diags)]
    return out


def header(cfg: OracleConfig) -> str:
    c, _keep = cfg._c()
    p = lib().oracle_header(C.byref(c))
    s = C.string_at(p).decode()
    return s


def read_vcf(cfg: OracleConfig, data: bytes, threads: int = 1) -> OracleResult:
    """readVcf (main.go:241): whole VCF text in, body rows (input order) out."""
    c, _keep = cfg._c()
    r = _Result()
    lib().oracle_read_vcf(C.byref(c), data, len(data), threads, C.byref(r))
    out = _unpack(r)
    lib().oracle_free_result(C.byref(r))
    return out


def process_block(cfg: OracleConfig, chrom_line: bytes, block, length: Optional[int] = None, eol_width: int = 1,
                  threads: int = 1) -> OracleResult:
    """processLines contract (main.go:476): header line + newline-terminated data lines."""
    c, _keep = cfg._c()
    r = _Result()
    if isinstance(block, (bytes, bytearray)):
        length = len(block)
        ptr = C.cast(C.c_char_p(bytes(block)), C.c_void_p)
        keep = block
    else:  # raw address (e.g. numpy array .ctypes.data)
        ptr = C.c_void_p(int(block))
        keep = None
    lib().oracle_process_block(C.byref(c), chrom_line, len(chrom_line), eol_width, ptr, length, threads, C.byref(r))
    out = _unpack(r)
    lib().oracle_free_result(C.byref(r))
    del keep
    return out


def get_alleles(chrom: str, pos: str, ref: str, alt: str):
    """getAlleles (main.go:723) -> (type, [pos], [ref byte], [alt], [altIdx])"""
    cap = 256
    type_out = C.create_string_buffer(32)
    pos_out = ((C.c_char * 24) * cap)()
    ref_out = C.create_string_buffer(cap)
    alt_out = (C.c_void_p * cap)()
    idx_out = (C.c_int * cap)()
    n = lib().oracle_get_alleles(chrom.encode(), pos.encode(), ref.encode(), alt.encode(), type_out, cap, pos_out,
                                 ref_out, alt_out, idx_out)
    libc = C.CDLL(None)
    libc.free.argtypes = [C.c_void_p]
    poss, refs, alts, idxs = [], [], [], []
    for i in range(min(n, cap)):
        poss.append(pos_out[i].value.decode())
        refs.append(ref_out.raw[i:i + 1].decode())
        alts.append(C.string_at(alt_out[i]).decode())
        libc.free(alt_out[i])
        idxs.append(idx_out[i])
    return type_out.value.decode(), poss, refs, alts, idxs


def het_hom(fields: Sequence[str], allele_num: str, names: Sequence[str]):
    """makeHetHomozygotes (main.go:1042) -> (homs, hets, missing, dosages, ac, an)"""
    n = len(fields)
    arr = (C.c_char_p * max(1, n))(*[f.encode() for f in fields])
    flags = (C.c_uint8 * max(1, n))()
    dos = (C.c_int8 * max(1, n))()
    ac, an = C.c_int(), C.c_int()
    lib().oracle_het_hom(arr, n, allele_num.encode(), flags, dos, C.byref(ac), C.byref(an))
    homs = [names[i] for i in range(n) if flags[i] == 2]
    hets = [names[i] for i in range(n) if flags[i] == 1]
    miss = [names[i] for i in range(n) if flags[i] == 3]
    return homs, hets, miss, [dos[i] for i in range(n)], ac.value, an.value


def alt_is_valid(alt: str) -> bool:
    b = alt.encode()
    return bool(lib().oracle_alt_is_valid(b, len(b)))


def format_float(q: float) -> str:
    buf = C.create_string_buffer(32)
    lib().oracle_format_float(q, buf)
    return buf.value.decode()


def tr_tv(ref: str, alt: str) -> str:
    b = alt.encode()
    return lib().oracle_tr_tv(ref.encode()[:1], b, len(b)).decode()
