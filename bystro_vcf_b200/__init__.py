"""bystro_vcf_b200 -- B200-native drop-in for bystro-vcf's per-line VCF transform.

Hand-written sm_100a CUDA kernels behind a C ABI (include/bvcf.h, lib/libbvcf.so); this package is the
host-side mirror of the reference's interface for that path (see host.py).  There is no CPU fallback.
"""
from .host import (Config, NotAVcfError, Transformer, header, parse_preamble, read_vcf, setup,  # noqa: F401
                   string_header, DIAG_TEXT)
from ._lib import BvcfError, LIB_PATH  # noqa: F401

__all__ = ["Config", "Transformer", "setup", "header", "string_header", "read_vcf", "parse_preamble",
           "NotAVcfError", "BvcfError", "DIAG_TEXT", "LIB_PATH"]
