// bvcf_host.cpp -- `bystro-vcf-b200`: the reference's main()/readVcf (main.go:134-396) as a C++ host over the
// libbvcf C ABI.  Same flags, same stdin/stdout contract, so it drops into
//     pigz -d -c in.vcf.gz | bystro-vcf-b200 --keepId --keepInfo | pigz -c > out.gz
// The per-line work happens on the GPU(s).  Threads mirror the reference's producer + worker pool
// (main.go:345-380) with GPUs as the workers:
//   reader     finds newline-aligned chunk boundaries; stdin is read straight into pinned buffers, a regular
//              --in file is memory-mapped and only the boundaries are computed here;
//   GPU worker one per GPU (chunk k goes to GPU k mod N): stages its chunks into its own pinned ring (parallel
//              memcpy out of the mapping), keeps n_slots chunks in flight on its context, and -- when it is chunk
//              k's turn -- writes the rows, appends the dosage batch to the Arrow file and logs the diagnostics,
//              so everything leaves in input order while the other GPUs keep working.
// (The reference is Go; no Go toolchain exists in this image, so the host is C++ -- see INTEGRATION.md for
// the equivalent cgo binding.)
#include <cerrno>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <condition_variable>
#include <deque>
#include <mutex>
#include <memory>
#include <string>
#include <thread>
#include <vector>

#include <fcntl.h>
#include <zlib.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include "../../include/bvcf.h"
#ifndef BVCF_NO_ARROW
#include "bvcf_arrow.h"
#endif

namespace {

struct Config {  // main.go:63-80
  std::string inPath, outPath, dosageMatrixOutPath, sampleListPath, famPath, errPath, cpuProfile;
  std::string emptyField = "!", fieldDelimiter = ";";
  bool noOut = false, keepID = false, keepInfo = false, keepQual = false, keepPos = false;
  bool allowAll = false;  // allowedFilters == nil
  std::vector<std::string> allowed, excluded;
  // not in the reference
  int gpus = 1;
  std::vector<int> devices;  // --devices a,b,...: CUDA device of each worker (a device may be listed twice); default 0..gpus-1
  size_t chunkBytes = 64u << 20;
  bool bgzfOut = false;  // --bgzfOut: the rows leave as bgzf blocks deflated on the GPU (the `| pigz -c` of README.md:10,71)
};

std::string trim(const std::string &s) {  // strings.TrimSpace
  size_t a = 0, b = s.size();
  while (a < b && isspace((unsigned char)s[a])) a++;
  while (b > a && isspace((unsigned char)s[b - 1])) b--;
  return s.substr(a, b - a);
}
std::vector<std::string> split_trim(const std::string &s) {
  std::vector<std::string> out;
  size_t p = 0;
  for (;;) {
    size_t q = s.find(',', p);
    out.push_back(trim(s.substr(p, q == std::string::npos ? std::string::npos : q - p)));
    if (q == std::string::npos) break;
    p = q + 1;
  }
  return out;
}

[[noreturn]] void fatal(const std::string &msg) {  // log.Fatal
  fprintf(stderr, "%s\n", msg.c_str());
  exit(1);
}

// Go `flag` syntax: -x / --x, --x=v / --x v; bool flags --x or --x=true|false; stops at the first non-flag.
Config setup(int argc, char **argv) {  // main.go:82-126
  Config c;
  std::string allow = "PASS,.", exclude;
  struct SF { const char *name; std::string *dst; };
  const SF sf[] = {{"in", &c.inPath}, {"fam", &c.famPath}, {"err", &c.errPath}, {"out", &c.outPath},
                   {"dosageOutput", &c.dosageMatrixOutPath}, {"sample", &c.sampleListPath},
                   {"emptyField", &c.emptyField}, {"fieldDelimiter", &c.fieldDelimiter}, {"cpuProfile", &c.cpuProfile},
                   {"allowFilter", &allow}, {"excludeFilter", &exclude}};
  struct BF { const char *name; bool *dst; };
  const BF bf[] = {{"noOut", &c.noOut}, {"keepId", &c.keepID}, {"keepQual", &c.keepQual}, {"keepPos", &c.keepPos},
                   {"keepInfo", &c.keepInfo}, {"bgzfOut", &c.bgzfOut}};
  std::string gpus, chunk, devices;
  for (int i = 1; i < argc; i++) {
    std::string s = argv[i];
    if (s.size() < 2 || s[0] != '-') break;
    if (s == "--") break;
    std::string name = s.substr(s[1] == '-' ? 2 : 1), val;
    bool has_val = false;
    size_t eq = name.find('=');
    if (eq != std::string::npos) { val = name.substr(eq + 1); name = name.substr(0, eq); has_val = true; }
    bool done = false;
    for (auto &b : bf)
      if (name == b.name) {
        if (!has_val) *b.dst = true;
        else if (val == "1" || val == "t" || val == "T" || val == "true" || val == "TRUE" || val == "True") *b.dst = true;
        else if (val == "0" || val == "f" || val == "F" || val == "false" || val == "FALSE" || val == "False") *b.dst = false;
        else fatal("invalid boolean value \"" + val + "\" for -" + name);
        done = true;
      }
    if (done) continue;
    std::string *dst = nullptr;
    for (auto &x : sf)
      if (name == x.name) dst = x.dst;
    if (name == "gpus") dst = &gpus;          // extension: number of GPUs to use
    if (name == "chunkBytes") dst = &chunk;   // extension: host chunk size
    if (name == "devices") dst = &devices;    // extension: explicit CUDA device per worker
    if (!dst) fatal("flag provided but not defined: -" + name);
    if (!has_val) {
      if (++i >= argc) fatal("flag needs an argument: -" + name);
      val = argv[i];
    }
    *dst = val;
  }
  if (!allow.empty() && allow != "*") c.allowed = split_trim(allow);  // main.go:108-114
  else c.allowAll = true;
  if (!exclude.empty()) c.excluded = split_trim(exclude);             // main.go:117-123
  if (!gpus.empty()) c.gpus = std::max(1, atoi(gpus.c_str()));
  if (!chunk.empty()) c.chunkBytes = std::max<size_t>(1 << 16, strtoull(chunk.c_str(), nullptr, 10));
  if (!devices.empty()) {
    for (auto &d : split_trim(devices)) c.devices.push_back(atoi(d.c_str()));
    c.gpus = (int)c.devices.size();
  }
  return c;
}

std::string string_header(const Config &c) {  // main.go:219-239
  bvcf_config k{};
  k.keep_pos = c.keepPos; k.keep_id = c.keepID; k.keep_info = c.keepInfo;
  char buf[512];
  bvcf_header_line(&k, buf, sizeof buf);
  return buf;
}

void write_all(int fd, const uint8_t *p, size_t n) {
  while (n) {
    ssize_t w = write(fd, p, n);
    if (w < 0) {
      if (errno == EINTR) continue;
      fatal(std::string("write: ") + strerror(errno));
    }
    p += w;
    n -= (size_t)w;
  }
}

const char *diag_text(int code) {  // main.go:41-51
  switch (code) {
    case BVCF_DIAG_SAME: return "REF == ALT";
    case BVCF_DIAG_BAD_ALT: return "ALT not ACTG";
    case BVCF_DIAG_DEL1: case BVCF_DIAG_DEL1_LIST: return "1st base REF != ALT";
    case BVCF_DIAG_POS: case BVCF_DIAG_POS_LIST: return "Invalid POS";
    case BVCF_DIAG_INS1: return "1st base ALT != REF";
    case BVCF_DIAG_MIXED: return "Mixed indel/snp sites not supported";
  }
  return "?";
}

// one of the reference's log.Printf lines (main.go:730-986); `line` points at the diagnosed line (CHROM \t POS \t ...)
void log_one_diag(const char *line, size_t avail, const bvcf_diag &d) {
  const char *end = line + avail;
  const char *t0 = (const char *)memchr(line, '\t', avail);
  if (!t0) return;
  const char *t1 = (const char *)memchr(t0 + 1, '\t', end - t0 - 1);
  if (!t1) return;
  const std::string chrom(line, t0), pos(t0 + 1, t1);
  const char *msg = diag_text(d.code);
  switch (d.code) {
    case BVCF_DIAG_SAME: fprintf(stderr, "%s:%s : %s\n", chrom.c_str(), pos.c_str(), msg); break;
    case BVCF_DIAG_MIXED: case BVCF_DIAG_DEL1_LIST:
      fprintf(stderr, "%s:%s ALT#%d %s\n", chrom.c_str(), pos.c_str(), d.alt_no, msg); break;
    case BVCF_DIAG_POS_LIST: fprintf(stderr, "%s:%s %s\n", chrom.c_str(), pos.c_str(), msg); break;
    default: fprintf(stderr, "%s:%s ALT #%d %s\n", chrom.c_str(), pos.c_str(), d.alt_no, msg);
  }
}
// the diagnostics of a chunk that is still pinned: every entry carries its line's offset
void log_diags(const uint8_t *chunk, size_t len, const bvcf_diag *d, size_t n) {
  for (size_t k = 0; k < n; k++)
    if (d[k].line_start < len) log_one_diag((const char *)chunk + d[k].line_start, len - d[k].line_start, d[k]);
}

struct Chunk {
  uint64_t seq = 0;
  uint8_t *buf = nullptr;        // pinned buffer holding the chunk (filled by the reader, or by the worker from `src`)
  const uint8_t *src = nullptr;  // memory-mapped input: where the chunk's bytes are
  size_t len = 0;
};

// bounded FIFO between the reader and one GPU worker
class ChunkQueue {
 public:
  explicit ChunkQueue(size_t cap) : cap_(cap) {}
  void push(const Chunk &c) {
    std::unique_lock<std::mutex> l(m_);
    cv_.wait(l, [&] { return q_.size() < cap_; });
    q_.push_back(c);
    cv_.notify_all();
  }
  void close() {
    std::lock_guard<std::mutex> l(m_);
    closed_ = true;
    cv_.notify_all();
  }
  // 1: got a chunk, 0: nothing right now (only when !block), -1: closed and drained
  int pop(Chunk &c, bool block) {
    std::unique_lock<std::mutex> l(m_);
    if (block) cv_.wait(l, [&] { return !q_.empty() || closed_; });
    if (q_.empty()) return closed_ ? -1 : 0;
    c = q_.front();
    q_.pop_front();
    cv_.notify_all();
    return 1;
  }

 private:
  std::mutex m_;
  std::condition_variable cv_;
  std::deque<Chunk> q_;
  size_t cap_;
  bool closed_ = false;
};

// a GPU's pinned staging buffers
class BufferPool {
 public:
  void add(uint8_t *b) { std::lock_guard<std::mutex> l(m_); free_.push_back(b); all_.push_back(b); }
  uint8_t *take() {
    std::unique_lock<std::mutex> l(m_);
    cv_.wait(l, [&] { return !free_.empty(); });
    uint8_t *b = free_.back();
    free_.pop_back();
    return b;
  }
  void give(uint8_t *b) { std::lock_guard<std::mutex> l(m_); free_.push_back(b); cv_.notify_all(); }
  ~BufferPool() { for (auto b : all_) bvcf_host_free(b); }

 private:
  std::mutex m_;
  std::condition_variable cv_;
  std::vector<uint8_t *> free_, all_;
};

// whose turn it is to write (chunk order == input order)
struct Turnstile {
  std::mutex m;
  std::condition_variable cv;
  uint64_t next = 0;
  void wait_for(uint64_t seq) { std::unique_lock<std::mutex> l(m); cv.wait(l, [&] { return next == seq; }); }
  void done() { std::lock_guard<std::mutex> l(m); next++; cv.notify_all(); }
};

// ---- bgzf input (.vcf.gz as bgzip / htslib write it): the compressed bytes go to the GPU, which inflates them ----
// Replaces the `pigz -d -c in.vcf.gz |` in front of the reference (README.md:10,46).  Groups of whole blocks are
// uploaded and inflated straight into the resident input region (bvcf_resident_inflate_bgzf), the transform runs
// there; rows, dosage batches and diagnostics come back.  One GPU, one group at a time.
bool looks_bgzf(const uint8_t *p, size_t n) {
  return n >= 18 && p[0] == 0x1f && p[1] == 0x8b && p[2] == 8 && (p[3] & 4) && p[12] == 'B' && p[13] == 'C';
}
size_t bgzf_block_size(const uint8_t *p, size_t n) {  // 0: header incomplete
  if (n < 18) return 0;
  const size_t xlen = (size_t)p[10] | ((size_t)p[11] << 8);
  if (n < 12 + xlen) return 0;
  for (size_t q = 12; q + 4 <= 12 + xlen;) {
    const size_t slen = (size_t)p[q + 2] | ((size_t)p[q + 3] << 8);
    if (p[q] == 'B' && p[q + 1] == 'C' && slen == 2) return ((size_t)p[q + 4] | ((size_t)p[q + 5] << 8)) + 1;
    q += 4 + slen;
  }
  fatal("gzip input without the bgzf BC subfield: decompress it first (gzip -dc | ...)");
}
// the first text bytes of a bgzf buffer, inflated on the host: only to read the VCF preamble
std::string bgzf_inflate_host(const uint8_t *p, size_t n, size_t want) {
  std::string out;
  size_t off = 0;
  while (off < n && out.size() < want) {
    const size_t bs = bgzf_block_size(p + off, n - off);
    if (!bs || off + bs > n) break;
    const size_t xlen = (size_t)p[off + 10] | ((size_t)p[off + 11] << 8);
    const uint32_t isize = (uint32_t)p[off + bs - 4] | ((uint32_t)p[off + bs - 3] << 8) | ((uint32_t)p[off + bs - 2] << 16) | ((uint32_t)p[off + bs - 1] << 24);
    std::string text(isize, '\0');
    z_stream z{};
    if (inflateInit2(&z, -15) != Z_OK) fatal("zlib: inflateInit2 failed");
    z.next_in = const_cast<Bytef *>(p + off + 12 + xlen); z.avail_in = (uInt)(bs - 12 - xlen - 8);
    z.next_out = (Bytef *)text.data(); z.avail_out = isize;
    const int zr = inflate(&z, Z_FINISH);
    inflateEnd(&z);
    if (zr != Z_STREAM_END && isize) fatal("corrupt bgzf block in the VCF preamble");
    out += text;
    off += bs;
  }
  return out;
}

// one bgzf block (SAM spec 4.1) around `text` (< 64 KiB), deflated on the host: the TSV header line
std::vector<uint8_t> bgzf_block_host(const uint8_t *text, size_t n) {
  std::vector<uint8_t> out(18 + compressBound(n) + 8);
  z_stream z{};
  if (deflateInit2(&z, 6, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY) != Z_OK) fatal("deflateInit2 failed");
  z.next_in = const_cast<Bytef *>(text); z.avail_in = (uInt)n;
  z.next_out = out.data() + 18; z.avail_out = (uInt)(out.size() - 18 - 8);
  if (deflate(&z, Z_FINISH) != Z_STREAM_END) fatal("deflate failed");
  const size_t payload = z.total_out;
  deflateEnd(&z);
  const size_t bsize = 18 + payload + 8;
  static const uint8_t hdr[16] = {0x1f, 0x8b, 8, 4, 0, 0, 0, 0, 0, 0xff, 6, 0, 'B', 'C', 2, 0};
  memcpy(out.data(), hdr, 16);
  out[16] = (uint8_t)((bsize - 1) & 0xFF); out[17] = (uint8_t)((bsize - 1) >> 8);
  const uint32_t crc = (uint32_t)crc32(crc32(0L, Z_NULL, 0), text, (uInt)n), isize = (uint32_t)n;
  for (int k = 0; k < 4; k++) { out[18 + payload + k] = (uint8_t)(crc >> (8 * k)); out[22 + payload + k] = (uint8_t)(isize >> (8 * k)); }
  out.resize(bsize);
  return out;
}
const uint8_t BGZF_EOF[28] = {0x1f, 0x8b, 8, 4, 0, 0, 0, 0, 0, 0xff, 6, 0, 'B', 'C', 2, 0, 0x1b, 0, 3, 0, 0, 0, 0, 0, 0, 0, 0, 0};

// readVcf over the resident regions, for compressed streams on either side: bgzf input is inflated on the GPU
// (compressed_in), --bgzfOut rows are deflated there
int run_resident(const Config &config, int in_fd, const uint8_t *map, size_t map_len, std::vector<uint8_t> &head, int out_fd,
                 bool compressed_in) {
  // compressed bytes: the mapping, or a growing buffer read from the pipe
  std::vector<uint8_t> &buf = head;
  size_t consumed = 0;  // bytes of the compressed stream already handed to the GPU
  bool eof = map != nullptr;
  auto avail = [&]() { return map ? map_len - consumed : buf.size() - consumed; };
  auto at = [&](size_t off) { return map ? map + consumed + off : buf.data() + consumed + off; };
  auto fill = [&](size_t want) {
    while (!eof && avail() < want) {
      const size_t old = buf.size();
      buf.resize(old + (8u << 20));
      const ssize_t r = read(in_fd, buf.data() + old, 8u << 20);
      if (r < 0) { if (errno == EINTR) { buf.resize(old); continue; } fatal(std::string("read: ") + strerror(errno)); }
      buf.resize(old + (size_t)r);
      if (r == 0) eof = true;
    }
  };
  // ---- preamble (main.go:250-294) from the first blocks ----
  std::string chrom_line;
  size_t data_off = 0;
  int eol_width = 1;
  for (size_t want = 1u << 20;; want *= 4) {
    fill(want);
    const std::string t = compressed_in ? bgzf_inflate_host(at(0), avail(), want * 4) : std::string((const char *)at(0), std::min(avail(), want * 4));
    const char *nl = (const char *)memchr(t.data(), '\n', t.size());
    if (!nl) { if (eof) fatal("Not a VCF file"); continue; }
    size_t first_end = nl - t.data();
    if (first_end > 0 && t[first_end - 1] == '\r') { eol_width = 2; first_end--; }
    if (!memmem(t.data(), first_end, "##fileformat=VCFv4", 18)) fatal("Not a VCF file");
    bool found = false;
    for (size_t q = (nl - t.data()) + 1; q < t.size();) {
      const char *e = (const char *)memchr(t.data() + q, '\n', t.size() - q);
      if (!e) break;
      size_t cl = (e - t.data()) - q;
      cl = cl + 1 >= (size_t)eol_width ? cl + 1 - eol_width : 0;
      if (cl >= 6 && memcmp(t.data() + q, "#CHROM", 6) == 0 && (cl == 6 || t[q + 6] == '\t')) {
        chrom_line.assign(t.data() + q, cl);
        data_off = (e - t.data()) + 1;
        found = true;
        break;
      }
      q = (e - t.data()) + 1;
    }
    if (found) break;
    if (eof) fatal("No header found");
  }
  if (!config.sampleListPath.empty() && !config.noOut) {
    FILE *f = fopen(config.sampleListPath.c_str(), "w");
    if (!f) fatal("Couldn't write sample list file");
    int field = 0;
    size_t s0 = 0;
    for (size_t i = 0; i <= chrom_line.size(); i++)
      if (i == chrom_line.size() || chrom_line[i] == '\t') {
        if (field >= 9) { std::string nm = chrom_line.substr(s0, i - s0); for (auto &ch : nm) if (ch == '.') ch = '_'; fprintf(f, "%s\n", nm.c_str()); }
        field++; s0 = i + 1;
      }
    fclose(f);
  }
  std::vector<const char *> allow_c, excl_c;
  for (auto &s : config.allowed) allow_c.push_back(s.c_str());
  for (auto &s : config.excluded) excl_c.push_back(s.c_str());
  bvcf_config bc{};
  bc.empty_field = config.emptyField.c_str(); bc.field_delim = config.fieldDelimiter.c_str();
  bc.keep_id = config.keepID; bc.keep_info = config.keepInfo; bc.keep_pos = config.keepPos;
  std::vector<std::string> sample_names;
  {
    int field = 0;
    size_t s0 = 0;
    for (size_t i = 0; i <= chrom_line.size(); i++)
      if (i == chrom_line.size() || chrom_line[i] == '\t') {
        if (field >= 9) { std::string nm = chrom_line.substr(s0, i - s0); for (auto &ch : nm) if (ch == '.') ch = '_'; sample_names.push_back(nm); }
        field++; s0 = i + 1;
      }
  }
  bool want_dosage = !config.dosageMatrixOutPath.empty();
#ifndef BVCF_NO_ARROW
  bvcf_arrow_writer *arrow = nullptr;
  if (want_dosage) {
    if (sample_names.empty()) {  // main.go:308-318
      fprintf(stderr, "No samples found in VCF file; writing empty dosage matrix file\n");
      FILE *f = fopen(config.dosageMatrixOutPath.c_str(), "w");
      if (!f) fatal(config.dosageMatrixOutPath + ": " + strerror(errno));
      fclose(f);
      want_dosage = false;
    } else {
      std::vector<const char *> nm;
      for (auto &s : sample_names) nm.push_back(s.c_str());
      char err[512];
      arrow = bvcf_arrow_open(config.dosageMatrixOutPath.c_str(), nm.data(), (uint32_t)nm.size(), err, sizeof err);
      if (!arrow) fatal(std::string("dosage output: ") + err);
    }
  }
#endif
  bc.want_tsv = !config.noOut; bc.want_dosage = want_dosage;
  bc.allow = allow_c.data(); bc.n_allow = config.allowAll ? -1 : (int)allow_c.size();
  bc.exclude = excl_c.data(); bc.n_exclude = (int)excl_c.size();
  bc.eol_width = eol_width; bc.normalize_dots = 1;
  bvcf_ctx *ctx = nullptr;
  int rc = bvcf_create(&ctx, config.devices.empty() ? 0 : config.devices[0], &bc);
  if (rc) fatal(std::string("bvcf_create: ") + bvcf_strerror(rc) + " -- a CUDA device is required, there is no CPU fallback");
  if ((rc = bvcf_set_header(ctx, chrom_line.data(), chrom_line.size()))) fatal(std::string("bvcf_set_header: ") + bvcf_strerror(rc));
  const size_t batch_text = std::max<size_t>(config.chunkBytes * 8, 64u << 20), max_line = 8u << 20;
  void *d_in, *d_out;
  if ((rc = bvcf_resident_alloc(ctx, batch_text + max_line + (1u << 20), batch_text / 4 + (64u << 20), &d_in, &d_out)))
    fatal(std::string("bvcf_resident_alloc: ") + bvcf_strerror(rc));
  uint8_t *pin = nullptr, *out_host = nullptr;   // pinned: compressed group in, rows out
  size_t pin_cap = 0, out_cap = 0;
  std::vector<uint8_t> carry, tail;
  size_t begin = data_off;
  for (;;) {
    // ---- a group of whole blocks worth about batch_text bytes of text ----
    size_t p = 0, text = 0;
    if (!compressed_in) {
      fill(batch_text);
      p = std::min(avail(), batch_text);
    }
    while (compressed_in && text < batch_text) {
      fill(p + (1u << 16) + 18);
      const size_t bs = bgzf_block_size(at(p), avail() - p);
      if (!bs || p + bs > avail()) {
        if (!eof) { fill(p + std::max<size_t>(bs, 1u << 16) + 18); continue; }
        break;
      }
      const uint8_t *e = at(p + bs - 4);
      text += (size_t)e[0] | ((size_t)e[1] << 8) | ((size_t)e[2] << 16) | ((size_t)e[3] << 24);
      p += bs;
    }
    if (p == 0) break;
    if (p > pin_cap) {
      if (pin) bvcf_host_free(pin);
      pin_cap = p + p / 4;
      if (bvcf_host_alloc((void **)&pin, pin_cap)) fatal("bvcf_host_alloc failed");
    }
    memcpy(pin, at(0), p);
    consumed += p;
    if (!map && consumed > (64u << 20)) { buf.erase(buf.begin(), buf.begin() + consumed); consumed = 0; }
    if (!carry.empty() && (rc = bvcf_resident_upload(ctx, 0, carry.data(), carry.size()))) fatal(std::string("upload: ") + bvcf_strerror(rc));
    size_t n_text = 0;
    if (compressed_in) {
      if ((rc = bvcf_resident_inflate_bgzf(ctx, pin, p, carry.size(), &n_text)))
        fatal(std::string("bgzf: ") + bvcf_strerror(rc) + " " + bvcf_last_error(ctx));
    } else {
      if ((rc = bvcf_resident_upload(ctx, carry.size(), pin, p))) fatal(std::string("upload: ") + bvcf_strerror(rc));
      n_text = p;
    }
    const size_t total = carry.size() + n_text;
    // ---- the longest newline-terminated prefix; the rest is carried into the next group ----
    const size_t tail_n = std::min(total - begin, max_line);
    tail.resize(tail_n);
    if (tail_n && (rc = bvcf_resident_peek(ctx, total - tail_n, tail.data(), tail_n))) fatal(std::string("peek: ") + bvcf_strerror(rc));
    const uint8_t *last_nl = tail_n ? (const uint8_t *)memrchr(tail.data(), '\n', tail_n) : nullptr;
    if (!last_nl) {
      if (tail_n < total - begin) fatal("a single line exceeds 8 MiB");
      carry.assign(tail.begin(), tail.end());  // no complete line yet: everything is carried
      begin = 0;
      continue;
    }
    const size_t end = total - tail_n + (size_t)(last_nl - tail.data()) + 1;
    bvcf_chunk_stats st{};
    if ((rc = bvcf_resident_run_at(ctx, begin, end, &st, nullptr))) fatal(std::string("bvcf_resident_run: ") + bvcf_strerror(rc) + " " + bvcf_last_error(ctx));
    if (!config.noOut && st.out_bytes) {
      if (st.out_bytes > out_cap) {
        if (out_host) bvcf_host_free(out_host);
        out_cap = st.out_bytes + st.out_bytes / 4;
        if (bvcf_host_alloc((void **)&out_host, out_cap)) fatal("bvcf_host_alloc failed");
      }
      if (config.bgzfOut) {
        size_t comp = 0;
        rc = bvcf_resident_download_bgzf(ctx, 0, st.out_bytes, out_host, out_cap, &comp);
        if (rc == BVCF_E_TOO_LARGE) {  // comp: the bytes needed
          bvcf_host_free(out_host);
          out_cap = comp + comp / 8;
          if (bvcf_host_alloc((void **)&out_host, out_cap)) fatal("bvcf_host_alloc failed");
          rc = bvcf_resident_download_bgzf(ctx, 0, st.out_bytes, out_host, out_cap, &comp);
        }
        if (rc) fatal(std::string("download: ") + bvcf_strerror(rc) + " " + bvcf_last_error(ctx));
        write_all(out_fd, out_host, comp);
      } else {
        if ((rc = bvcf_resident_download(ctx, 0, out_host, st.out_bytes))) fatal(std::string("download: ") + bvcf_strerror(rc));
        write_all(out_fd, out_host, st.out_bytes);
      }
    }
    {  // dosage batch and diagnostics of this group
      bvcf_dosage_batch dos;
      const bvcf_diag *dg = nullptr;
      size_t nd = 0;
      if ((rc = bvcf_resident_results(ctx, want_dosage ? &dos : nullptr, &dg, &nd))) fatal(std::string("results: ") + bvcf_strerror(rc));
#ifndef BVCF_NO_ARROW
      if (arrow && want_dosage && dos.n_rows && bvcf_arrow_write(arrow, dos.n_rows, dos.dosage, dos.loci, dos.loci_off))
        fatal(std::string("dosage output: ") + bvcf_arrow_error(arrow));
#endif
      char line[4096];
      for (size_t k = 0; k < nd; k++) {
        const size_t n = std::min<size_t>(sizeof line, end - dg[k].line_start);
        if (dg[k].line_start < end && !bvcf_resident_peek(ctx, dg[k].line_start, line, n)) log_one_diag(line, n, dg[k]);
      }
    }
    carry.assign(last_nl + 1, (const uint8_t *)tail.data() + tail_n);
    begin = 0;
  }
  // an unterminated last line (the carry) is dropped (main.go:354-357)
  if (config.bgzfOut && !config.noOut) write_all(out_fd, BGZF_EOF, sizeof BGZF_EOF);
  if (pin) bvcf_host_free(pin);
  if (out_host) bvcf_host_free(out_host);
  int rc_exit = 0;
#ifndef BVCF_NO_ARROW
  if (arrow && bvcf_arrow_close(arrow)) rc_exit = 1;
#endif
  bvcf_destroy(ctx);
  return rc_exit;
}

}  // namespace

int main(int argc, char **argv) {
  Config config = setup(argc, argv);
  int in_fd = 0;
  if (!config.inPath.empty() && (in_fd = open(config.inPath.c_str(), O_RDONLY)) < 0) fatal(config.inPath + ": " + strerror(errno));
  // main.go:150-156 re-points os.Stderr at the file, which Go's `log` never looks at again; what the flag evidently
  // means is honoured here: diagnostics are appended to it
  if (!config.errPath.empty() && !freopen(config.errPath.c_str(), "a", stderr)) fatal(config.errPath + ": " + strerror(errno));
  if (config.noOut && !config.outPath.empty()) fatal("Cannot specify --noOut and --out");                 // main.go:160
  if (config.noOut && config.dosageMatrixOutPath.empty()) fatal("When specifying --noOut, must specify --dosageOutput");  // :164
#ifdef BVCF_NO_ARROW
  if (!config.dosageMatrixOutPath.empty())
    fatal("--dosageOutput: this binary was built without libarrow (pyarrow was not found at build time); "
          "python -m bystro_vcf_b200 writes the Arrow file");
#endif
  int out_fd = 1;
  if (!config.noOut && !config.outPath.empty() &&
      (out_fd = open(config.outPath.c_str(), O_WRONLY | O_CREAT, 0644)) < 0)  // no O_TRUNC, like main.go:172
    fatal(config.outPath + ": " + strerror(errno));
  if (!config.noOut) {
    const std::string h = string_header(config) + "\n";  // main.go:199: before any input is read
    if (config.bgzfOut) {
      const std::vector<uint8_t> b = bgzf_block_host((const uint8_t *)h.data(), h.size());
      write_all(out_fd, b.data(), b.size());
    } else {
      write_all(out_fd, (const uint8_t *)h.data(), h.size());
    }
  }

  // ---- input: a regular file is memory-mapped, anything else (pipes) is read ----
  const uint8_t *map = nullptr;
  size_t map_len = 0;
  {
    struct stat sb;
    if (in_fd != 0 && fstat(in_fd, &sb) == 0 && S_ISREG(sb.st_mode) && sb.st_size > 0) {
      void *m = mmap(nullptr, (size_t)sb.st_size, PROT_READ, MAP_PRIVATE, in_fd, 0);
      if (m != MAP_FAILED) {
        map = (const uint8_t *)m;
        map_len = (size_t)sb.st_size;
        madvise(m, map_len, MADV_SEQUENTIAL);
      }
    }
  }

  std::vector<uint8_t> head;  // pipe input: what has been read so far
  {  // .vcf.gz (bgzf): another path altogether
    bool bz = false;
    if (map) bz = looks_bgzf(map, map_len);
    else {
      head.resize(1 << 16);
      size_t got = 0;
      while (got < 18) {
        const ssize_t r = read(in_fd, head.data() + got, head.size() - got);
        if (r < 0) { if (errno == EINTR) continue; fatal(std::string("read: ") + strerror(errno)); }
        if (r == 0) break;
        got += (size_t)r;
      }
      head.resize(got);
      bz = looks_bgzf(head.data(), head.size());
    }
    if (bz || config.bgzfOut) {
      if (config.bgzfOut && config.gpus > 1) fatal("--bgzfOut runs on one GPU");
      const int rc = run_resident(config, in_fd, map, map_len, head, out_fd, bz);
      if (map) munmap((void *)map, map_len);
      if (out_fd != 1) close(out_fd);
      return rc;
    }
  }
  // ---- preamble (main.go:250-294): EOL, ##fileformat check, #CHROM line ----
  size_t data_off = 0;
  int eol_width = 1;
  std::string chrom_line;
  bool eof = false;
  for (bool found = false; !found;) {
    const uint8_t *p;
    size_t n;
    if (map) {
      p = map; n = map_len; eof = true;
    } else {
      const size_t old = head.size();
      head.resize(old + (1 << 20));
      ssize_t r = read(in_fd, head.data() + old, 1 << 20);
      if (r < 0) fatal(std::string("read: ") + strerror(errno));
      head.resize(old + (size_t)r);
      if (r == 0) eof = true;
      p = head.data(); n = head.size();
    }
    if (n == 0 && eof) fatal("Not a VCF file");
    const uint8_t *nl = (const uint8_t *)memchr(p, '\n', n);
    if (!nl) { if (eof) fatal("Not a VCF file"); continue; }
    size_t first_end = nl - p;
    if (first_end > 0 && p[first_end - 1] == '\r') { eol_width = 2; first_end--; }
    if (!memmem(p, first_end, "##fileformat=VCFv4", 18)) fatal("Not a VCF file");  // main.go:256-264
    size_t q = (nl - p) + 1;
    while (q < n) {
      const uint8_t *e = (const uint8_t *)memchr(p + q, '\n', n - q);
      if (!e) break;
      size_t cl = (e - p) - q;
      cl = cl + 1 >= (size_t)eol_width ? cl + 1 - eol_width : 0;
      if (cl >= 6 && memcmp(p + q, "#CHROM", 6) == 0 && (cl == 6 || p[q + 6] == '\t')) {
        chrom_line.assign((const char *)p + q, cl);
        data_off = (e - p) + 1;
        found = true;
        break;
      }
      q = (e - p) + 1;
    }
    if (!found && eof) fatal("No header found");  // main.go:293
  }
  std::vector<std::string> sample_names;
  {
    int field = 0;
    size_t s = 0;
    for (size_t i = 0; i <= chrom_line.size(); i++)
      if (i == chrom_line.size() || chrom_line[i] == '\t') {
        if (field >= 9) {
          std::string nm = chrom_line.substr(s, i - s);
          for (auto &ch : nm) if (ch == '.') ch = '_';  // parse.NormalizeHeader
          sample_names.push_back(nm);
        }
        field++; s = i + 1;
      }
  }
  if (!config.sampleListPath.empty() && !config.noOut) {  // main.go:398-445
    FILE *f = fopen(config.sampleListPath.c_str(), "w");
    if (!f) fatal("Couldn't write sample list file");
    for (auto &nm : sample_names) fprintf(f, "%s\n", nm.c_str());
    fclose(f);
  }
  bool want_dosage = !config.dosageMatrixOutPath.empty();
#ifndef BVCF_NO_ARROW
  bvcf_arrow_writer *arrow = nullptr;
  if (want_dosage) {
    if (sample_names.empty()) {  // main.go:308-318
      fprintf(stderr, "No samples found in VCF file; writing empty dosage matrix file\n");
      FILE *f = fopen(config.dosageMatrixOutPath.c_str(), "w");
      if (!f) fatal(config.dosageMatrixOutPath + ": " + strerror(errno));
      fclose(f);
      want_dosage = false;
    } else {
      std::vector<const char *> nm;
      for (auto &s : sample_names) nm.push_back(s.c_str());
      char err[512];
      arrow = bvcf_arrow_open(config.dosageMatrixOutPath.c_str(), nm.data(), (uint32_t)nm.size(), err, sizeof err);
      if (!arrow) fatal(std::string("dosage output: ") + err);
    }
  }
#endif

  // ---- one context per GPU ----
  std::vector<const char *> allow_c, excl_c;
  for (auto &s : config.allowed) allow_c.push_back(s.c_str());
  for (auto &s : config.excluded) excl_c.push_back(s.c_str());
  bvcf_config bc{};
  bc.empty_field = config.emptyField.c_str();
  bc.field_delim = config.fieldDelimiter.c_str();
  bc.keep_id = config.keepID; bc.keep_info = config.keepInfo; bc.keep_pos = config.keepPos;
  bc.want_tsv = !config.noOut; bc.want_dosage = want_dosage;
  bc.allow = allow_c.data(); bc.n_allow = config.allowAll ? -1 : (int)allow_c.size();
  bc.exclude = excl_c.data(); bc.n_exclude = (int)excl_c.size();
  bc.eol_width = eol_width; bc.normalize_dots = 1;
  const int n_slots = 3;
  bc.n_slots = n_slots;
  const size_t cap = 2 * config.chunkBytes + (64u << 20);  // a chunk grows until it holds a newline
  bc.max_chunk_bytes = cap;
  const int n_gpu = config.gpus;
  std::vector<bvcf_ctx *> ctxs(n_gpu, nullptr);
  std::vector<BufferPool> pools(n_gpu);
  std::vector<std::unique_ptr<ChunkQueue>> queues;
  for (int g = 0; g < n_gpu; g++) {
    int rc = bvcf_create(&ctxs[g], config.devices.empty() ? g : config.devices[g], &bc);
    if (rc) fatal(std::string("bvcf_create(gpu ") + std::to_string(g) + "): " + bvcf_strerror(rc) + " -- a CUDA device is required, there is no CPU fallback");
    rc = bvcf_set_header(ctxs[g], chrom_line.data(), chrom_line.size());
    if (rc) fatal(std::string("bvcf_set_header: ") + bvcf_strerror(rc));
    for (int k = 0; k < n_slots + 2; k++) {
      uint8_t *b = nullptr;
      if (bvcf_host_alloc((void **)&b, cap)) fatal("bvcf_host_alloc failed");
      pools[g].add(b);
    }
    queues.emplace_back(new ChunkQueue(n_slots + 1));
  }

  // ---- GPU workers ----
  Turnstile turn;
  std::mutex fatal_m;
  auto worker = [&](int g) {
    bvcf_ctx *ctx = ctxs[g];
    std::deque<Chunk> inflight;
    bool closed = false;
    for (;;) {
      while (!closed && (int)inflight.size() < n_slots) {
        Chunk c;
        const int got = queues[g]->pop(c, /*block=*/inflight.empty());
        if (got < 0) { closed = true; break; }
        if (got == 0) break;
        if (c.src) {  // memory-mapped input: this thread stages its own chunk (N GPUs, N copies in parallel)
          c.buf = pools[g].take();
          memcpy(c.buf, c.src, c.len);
        }
        const int rc = bvcf_submit(ctx, c.seq, c.buf, c.len);
        if (rc) { std::lock_guard<std::mutex> l(fatal_m); fatal(std::string("bvcf_submit: ") + bvcf_strerror(rc) + " " + bvcf_last_error(ctx)); }
        inflight.push_back(c);
      }
      if (inflight.empty()) {
        if (closed) break;
        continue;
      }
      const Chunk c = inflight.front();
      inflight.pop_front();
      const uint8_t *tsv; size_t n; const bvcf_diag *dg; size_t nd;
      bvcf_dosage_batch dos;
      const int rc = bvcf_collect(ctx, c.seq, &tsv, &n, &dos, &dg, &nd, nullptr);
      if (rc) { std::lock_guard<std::mutex> l(fatal_m); fatal(std::string("bvcf_collect: ") + bvcf_strerror(rc) + " " + bvcf_last_error(ctx)); }
      turn.wait_for(c.seq);  // input order
      if (!config.noOut) write_all(out_fd, tsv, n);
#ifndef BVCF_NO_ARROW
      if (arrow && dos.n_rows && bvcf_arrow_write(arrow, dos.n_rows, dos.dosage, dos.loci, dos.loci_off))
        fatal(std::string("dosage output: ") + bvcf_arrow_error(arrow));
#endif
      log_diags(c.buf, c.len, dg, nd);
      turn.done();
      bvcf_release(ctx, c.seq);
      pools[g].give(c.buf);
    }
  };
  std::vector<std::thread> threads;
  for (int g = 0; g < n_gpu; g++) threads.emplace_back(worker, g);

  // ---- reader: newline-aligned chunks, chunk k to GPU k mod N ----
  uint64_t seq = 0;
  if (map) {
    size_t p = data_off;
    // an unterminated last line is dropped (main.go:354-357)
    size_t end = map_len;
    {
      const uint8_t *last = end > p ? (const uint8_t *)memrchr(map + p, '\n', end - p) : nullptr;
      end = last ? (size_t)(last - map) + 1 : p;
    }
    while (p < end) {
      size_t q = std::min(p + config.chunkBytes, end);
      if (q < end) {
        const uint8_t *nl = (const uint8_t *)memrchr(map + p, '\n', q - p);
        if (!nl) {  // a line longer than a chunk: extend to its end
          nl = (const uint8_t *)memchr(map + q, '\n', end - q);
          if (!nl) break;
        }
        q = (size_t)(nl - map) + 1;
      }
      if (q - p > cap) fatal("a single line exceeds the chunk capacity; raise --chunkBytes");
      Chunk c;
      c.seq = seq; c.src = map + p; c.len = q - p;
      queues[seq % n_gpu]->push(c);
      seq++;
      p = q;
    }
  } else {
    uint8_t *buf = pools[0].take();
    size_t fill = head.size() - data_off;  // bytes already in the current buffer
    if (fill > cap) fatal("header buffer larger than a chunk");
    memcpy(buf, head.data() + data_off, fill);
    head.clear(); head.shrink_to_fit();
    while (!eof || fill) {
      while (!eof && fill < config.chunkBytes) {
        ssize_t r = read(in_fd, buf + fill, std::min(cap - fill, config.chunkBytes - fill));
        if (r < 0) { if (errno == EINTR) continue; fatal(std::string("read: ") + strerror(errno)); }
        if (r == 0) { eof = true; break; }
        fill += (size_t)r;
      }
      const uint8_t *last_nl = fill ? (const uint8_t *)memrchr(buf, '\n', fill) : nullptr;
      if (!last_nl) {
        if (eof) break;  // an unterminated last line is dropped (main.go:354-357)
        if (fill >= cap) fatal("a single line exceeds the chunk capacity; raise --chunkBytes");
        ssize_t r = read(in_fd, buf + fill, cap - fill);
        if (r <= 0) { eof = true; continue; }
        fill += (size_t)r;
        continue;
      }
      const size_t cut = (last_nl - buf) + 1;
      uint8_t *next = pools[(seq + 1) % n_gpu].take();  // the next chunk's buffer comes from its GPU's ring
      memcpy(next, buf + cut, fill - cut);              // carry the partial line
      Chunk c;
      c.seq = seq; c.buf = buf; c.len = cut;
      queues[seq % n_gpu]->push(c);
      seq++;
      fill -= cut;
      buf = next;
    }
    pools[seq % n_gpu].give(buf);
  }
  for (auto &q : queues) q->close();
  for (auto &t : threads) t.join();
  int rc_exit = 0;
#ifndef BVCF_NO_ARROW
  if (arrow && bvcf_arrow_close(arrow)) rc_exit = 1;
#endif
  for (auto c : ctxs) bvcf_destroy(c);
  if (map) munmap((void *)map, map_len);
  if (out_fd != 1) close(out_fd);
  return rc_exit;
}
