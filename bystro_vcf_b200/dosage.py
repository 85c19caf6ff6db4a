"""Arrow IPC file framing of the dosage matrix: host-side only.

The GPU produces the data the reference feeds its Arrow writer (main.go:576-584): one "chrom:pos:ref:alt"
locus string + n_samples int8 per emitted row.  Here that data is framed like the reference does
(main.go:320-336, arrow/arrow.go:24-137): Arrow IPC *file* (Feather v2), zstd, schema `locus: utf8` +
one non-nullable `int8` column per sample, record batches of at most 5,000 rows."""
from __future__ import annotations

from typing import List, Sequence


class DosageWriter:
    CHUNK_ROWS = 5000  # arrow.go: NewArrowRowBuilder(arrowWriter, 5e3) main.go:517

    def __init__(self, path: str, sample_names: Sequence[bytes]):
        import pyarrow as pa

        self._pa = pa
        names = [n.decode() if isinstance(n, bytes) else n for n in sample_names]
        fields = [pa.field("locus", pa.string(), nullable=False)] + [pa.field(n, pa.int8(), nullable=False) for n in names]
        self.schema = pa.schema(fields)
        opts = pa.ipc.IpcWriteOptions(compression="zstd")
        self._sink = pa.OSFile(path, "wb")
        self._writer = pa.ipc.new_file(self._sink, self.schema, options=opts)
        self.n_rows = 0

    def write(self, loci: List[bytes], dosage) -> None:
        """dosage: numpy int8 [rows, samples] in row order."""
        pa = self._pa
        n = len(loci)
        for lo in range(0, n, self.CHUNK_ROWS):
            hi = min(lo + self.CHUNK_ROWS, n)
            cols = [pa.array([x.decode() for x in loci[lo:hi]], type=pa.string())]
            block = dosage[lo:hi]
            cols += [pa.array(block[:, j], type=pa.int8()) for j in range(block.shape[1])]
            self._writer.write_batch(pa.record_batch(cols, schema=self.schema))
        self.n_rows += n

    def close(self) -> None:
        self._writer.close()
        self._sink.close()
