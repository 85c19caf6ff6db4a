"""`python -m bystro_vcf_b200 [flags]` -- main() of the reference (main.go:134-217) on the GPU path.

Same flags, same stdin/stdout behaviour: header line first (main.go:199), then rows in input order."""
from __future__ import annotations

import sys

from .host import NotAVcfError, DIAG_TEXT, read_vcf, setup, string_header


def main(argv=None) -> int:
    cfg = setup(argv)
    if cfg.noOut and cfg.outPath:
        print("Cannot specify --noOut and --out", file=sys.stderr)  # main.go:160
        return 1
    if cfg.noOut and not cfg.dosageMatrixOutPath:
        print("When specifying --noOut, must specify --dosageOutput", file=sys.stderr)  # main.go:164
        return 1
    inp = open(cfg.inPath, "rb") if cfg.inPath else sys.stdin.buffer
    out = None
    if not cfg.noOut:
        out = open(cfg.outPath, "r+b" if False else "wb") if cfg.outPath else sys.stdout.buffer
        out.write(string_header(cfg).encode() + b"\n")  # main.go:199

    def log(line_no, alt_no, code):  # the reference's log.Printf sites (main.go:730-986)
        print("line %d ALT #%d %s" % (line_no, alt_no, DIAG_TEXT.get(code, "?")), file=sys.stderr)

    try:
        read_vcf(cfg, inp, out, diag_sink=log)
    except NotAVcfError as e:
        print(str(e), file=sys.stderr)  # log.Fatal main.go:263,293
        return 1
    finally:
        if out is not None:
            out.flush()
    return 0


if __name__ == "__main__":
    sys.exit(main())
