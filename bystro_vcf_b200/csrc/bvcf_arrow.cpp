// bvcf_arrow.cpp -- the dosage matrix as an Arrow IPC *file* with zstd-compressed buffers, the way the reference
// writes it (main.go:320-336, arrow/arrow.go:24-137): schema `locus: utf8` + one `int8` column per sample, all
// non-nullable, record batches of at most 5,000 rows (NewArrowRowBuilder(arrowWriter, 5e3), main.go:517).
// Links against the libarrow that ships inside the pyarrow wheel (no Arrow C++ install in this image); compiled
// as C++20 because Arrow 24's headers ask for it.  The host binary talks to it through the small C-style
// interface of bvcf_arrow.h so that bvcf_host.cpp stays plain C++17.
#include "bvcf_arrow.h"

#include <arrow/api.h>
#include <arrow/io/file.h>
#include <arrow/ipc/writer.h>
#include <arrow/util/compression.h>

#include <cstring>
#include <memory>
#include <string>
#include <vector>

struct bvcf_arrow_writer {
  std::shared_ptr<arrow::Schema> schema;
  std::shared_ptr<arrow::io::FileOutputStream> out;
  std::shared_ptr<arrow::ipc::RecordBatchWriter> writer;
  uint32_t n_samples = 0;
  // rows waiting for the next 5,000-row batch: column-major so that a column is one contiguous buffer
  uint64_t n_pending = 0;
  std::string loci;                  // concatenated
  std::vector<int32_t> loci_off{0};  // n_pending + 1
  std::vector<int8_t> cols;          // n_samples x BATCH_ROWS
  std::string error;
};

namespace {
constexpr uint64_t BATCH_ROWS = 5000;

bool flush(bvcf_arrow_writer *w) {
  if (w->n_pending == 0) return true;
  const int64_t n = (int64_t)w->n_pending;
  std::vector<std::shared_ptr<arrow::Array>> arrays;
  arrays.reserve(w->n_samples + 1);
  {
    auto offs = arrow::Buffer::Wrap(w->loci_off.data(), (size_t)n + 1);
    auto data = std::make_shared<arrow::Buffer>((const uint8_t *)w->loci.data(), (int64_t)w->loci.size());
    arrays.push_back(std::make_shared<arrow::StringArray>(n, offs, data));
  }
  for (uint32_t j = 0; j < w->n_samples; j++) {
    auto buf = std::make_shared<arrow::Buffer>((const uint8_t *)(w->cols.data() + (size_t)j * BATCH_ROWS), n);
    arrays.push_back(std::make_shared<arrow::Int8Array>(n, buf));
  }
  auto batch = arrow::RecordBatch::Make(w->schema, n, std::move(arrays));
  auto st = w->writer->WriteRecordBatch(*batch);
  if (!st.ok()) { w->error = st.ToString(); return false; }
  w->n_pending = 0;
  w->loci.clear();
  w->loci_off.assign(1, 0);
  return true;
}
}  // namespace

extern "C" {

bvcf_arrow_writer *bvcf_arrow_open(const char *path, const char *const *sample_names, uint32_t n_samples, char *err, size_t err_cap) {
  auto fail = [&](const std::string &m) -> bvcf_arrow_writer * {
    if (err && err_cap) { strncpy(err, m.c_str(), err_cap - 1); err[err_cap - 1] = 0; }
    return nullptr;
  };
  auto w = std::make_unique<bvcf_arrow_writer>();
  std::vector<std::shared_ptr<arrow::Field>> fields;
  fields.push_back(arrow::field("locus", arrow::utf8(), /*nullable=*/false));
  for (uint32_t i = 0; i < n_samples; i++) fields.push_back(arrow::field(sample_names[i], arrow::int8(), false));
  w->schema = arrow::schema(std::move(fields));
  auto out = arrow::io::FileOutputStream::Open(path);
  if (!out.ok()) return fail(out.status().ToString());
  w->out = *out;
  auto opts = arrow::ipc::IpcWriteOptions::Defaults();
  auto codec = arrow::util::Codec::Create(arrow::Compression::ZSTD);  // ipc.WithZstd(), main.go:334
  if (!codec.ok()) return fail(codec.status().ToString());
  opts.codec = std::move(*codec);
  auto wr = arrow::ipc::MakeFileWriter(w->out, w->schema, opts);
  if (!wr.ok()) return fail(wr.status().ToString());
  w->writer = *wr;
  w->n_samples = n_samples;
  w->cols.resize((size_t)n_samples * BATCH_ROWS);
  return w.release();
}

int bvcf_arrow_write(bvcf_arrow_writer *w, uint64_t n_rows, const int8_t *dosage, const uint8_t *loci, const uint64_t *loci_off) {
  const uint32_t ns = w->n_samples;
  uint64_t r = 0;
  while (r < n_rows) {
    const uint64_t take = std::min<uint64_t>(n_rows - r, BATCH_ROWS - w->n_pending);
    // rows r .. r + take: row-major int8 -> the columns of the pending batch (tiled so that both sides stay in cache)
    constexpr uint64_t TR = 64;
    constexpr uint32_t TC = 256;
    for (uint64_t r0 = 0; r0 < take; r0 += TR) {
      const uint64_t r1 = std::min(take, r0 + TR);
      for (uint32_t c0 = 0; c0 < ns; c0 += TC) {
        const uint32_t c1 = std::min(ns, c0 + TC);
        for (uint64_t i = r0; i < r1; i++) {
          const int8_t *src = dosage + (r + i) * ns;
          int8_t *dst = w->cols.data() + w->n_pending + i;
          for (uint32_t j = c0; j < c1; j++) dst[(size_t)j * BATCH_ROWS] = src[j];
        }
      }
    }
    for (uint64_t i = 0; i < take; i++) {
      const uint64_t a = loci_off[r + i], b = loci_off[r + i + 1];
      w->loci.append((const char *)loci + a, (size_t)(b - a));
      w->loci_off.push_back((int32_t)w->loci.size());
    }
    w->n_pending += take;
    r += take;
    if (w->n_pending == BATCH_ROWS && !flush(w)) return -1;
  }
  return 0;
}

int bvcf_arrow_close(bvcf_arrow_writer *w) {
  if (!w) return 0;
  int rc = 0;
  if (!flush(w)) rc = -1;
  if (w->writer) { auto st = w->writer->Close(); if (!st.ok()) { w->error = st.ToString(); rc = -1; } }
  if (w->out) { auto st = w->out->Close(); if (!st.ok()) { w->error = st.ToString(); rc = -1; } }
  if (rc) fprintf(stderr, "dosage output: %s\n", w->error.c_str());
  delete w;
  return rc;
}

const char *bvcf_arrow_error(const bvcf_arrow_writer *w) { return w ? w->error.c_str() : ""; }

}  // extern "C"
