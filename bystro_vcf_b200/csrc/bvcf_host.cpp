// bvcf_host.cpp -- `bystro-vcf-b200`: the reference's main()/readVcf (main.go:134-396) as a C++ host over the
// libbvcf C ABI.  Same flags, same stdin/stdout contract, so it drops into
//     pigz -d -c in.vcf.gz | bystro-vcf-b200 --keepId --keepInfo | pigz -c > out.gz
// The per-line work happens on the GPU(s); this file only finds the header, cuts newline-aligned chunks into
// pinned buffers, keeps n_slots chunks in flight per GPU and writes the rows back in input order.
// (The reference is Go; no Go toolchain exists in this image, so the host is C++ -- see INTEGRATION.md for
// the equivalent cgo binding.)
#include <cerrno>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include <fcntl.h>
#include <unistd.h>

#include "../../include/bvcf.h"

namespace {

struct Config {  // main.go:63-80
  std::string inPath, outPath, dosageMatrixOutPath, sampleListPath, famPath, errPath, cpuProfile;
  std::string emptyField = "!", fieldDelimiter = ";";
  bool noOut = false, keepID = false, keepInfo = false, keepQual = false, keepPos = false;
  bool allowAll = false;  // allowedFilters == nil
  std::vector<std::string> allowed, excluded;
  // not in the reference
  int gpus = 1;
  size_t chunkBytes = 64u << 20;
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
                   {"keepInfo", &c.keepInfo}};
  std::string gpus, chunk;
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

// the reference's log.Printf lines (main.go:730-986); chrom:pos are looked up in the chunk that is still pinned
void log_diags(const uint8_t *chunk, size_t len, const bvcf_diag *d, size_t n) {
  if (!n) return;
  std::vector<size_t> starts{0};
  for (size_t i = 0; i < len; i++)
    if (chunk[i] == '\n') starts.push_back(i + 1);
  for (size_t k = 0; k < n; k++) {
    if (d[k].line_no >= starts.size()) continue;
    const char *l = (const char *)chunk + starts[d[k].line_no];
    const char *end = (const char *)chunk + len;
    const char *t0 = (const char *)memchr(l, '\t', end - l);
    if (!t0) continue;
    const char *t1 = (const char *)memchr(t0 + 1, '\t', end - t0 - 1);
    if (!t1) continue;
    const std::string chrom(l, t0), pos(t0 + 1, t1);
    const char *msg = diag_text(d[k].code);
    switch (d[k].code) {
      case BVCF_DIAG_SAME: fprintf(stderr, "%s:%s : %s\n", chrom.c_str(), pos.c_str(), msg); break;
      case BVCF_DIAG_MIXED: case BVCF_DIAG_DEL1_LIST:
        fprintf(stderr, "%s:%s ALT#%d %s\n", chrom.c_str(), pos.c_str(), d[k].alt_no, msg); break;
      case BVCF_DIAG_POS_LIST: fprintf(stderr, "%s:%s %s\n", chrom.c_str(), pos.c_str(), msg); break;
      default: fprintf(stderr, "%s:%s ALT #%d %s\n", chrom.c_str(), pos.c_str(), d[k].alt_no, msg);
    }
  }
}

struct InFlight {
  uint64_t seq;
  int gpu;
  uint8_t *buf;
  size_t len;
};

}  // namespace

int main(int argc, char **argv) {
  Config config = setup(argc, argv);
  int in_fd = 0;
  if (!config.inPath.empty() && (in_fd = open(config.inPath.c_str(), O_RDONLY)) < 0) fatal(config.inPath + ": " + strerror(errno));
  if (!config.errPath.empty() && !freopen(config.errPath.c_str(), "a", stderr)) fatal(config.errPath + ": " + strerror(errno));
  if (config.noOut && !config.outPath.empty()) fatal("Cannot specify --noOut and --out");                 // main.go:160
  if (config.noOut && config.dosageMatrixOutPath.empty()) fatal("When specifying --noOut, must specify --dosageOutput");  // :164
  if (!config.dosageMatrixOutPath.empty())
    fatal("--dosageOutput: the Arrow IPC writer lives in the Python host (python -m bystro_vcf_b200); this binary writes the TSV");
  int out_fd = 1;
  if (!config.noOut && !config.outPath.empty() &&
      (out_fd = open(config.outPath.c_str(), O_WRONLY | O_CREAT, 0644)) < 0)  // no O_TRUNC, like main.go:172
    fatal(config.outPath + ": " + strerror(errno));
  if (!config.noOut) {
    const std::string h = string_header(config) + "\n";  // main.go:199: before any input is read
    write_all(out_fd, (const uint8_t *)h.data(), h.size());
  }

  // ---- preamble (main.go:250-294): EOL, ##fileformat check, #CHROM line ----
  std::vector<uint8_t> head;
  size_t data_off = 0;
  int eol_width = 1;
  std::string chrom_line;
  bool eof = false;
  for (bool found = false; !found;) {
    const size_t old = head.size();
    head.resize(old + (1 << 20));
    ssize_t r = read(in_fd, head.data() + old, 1 << 20);
    if (r < 0) fatal(std::string("read: ") + strerror(errno));
    head.resize(old + (size_t)r);
    if (r == 0) eof = true;
    const uint8_t *p = head.data();
    const size_t n = head.size();
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
  if (!config.sampleListPath.empty() && !config.noOut) {  // main.go:398-445
    FILE *f = fopen(config.sampleListPath.c_str(), "w");
    if (!f) fatal("Couldn't write sample list file");
    int field = 0;
    size_t s = 0;
    for (size_t i = 0; i <= chrom_line.size(); i++)
      if (i == chrom_line.size() || chrom_line[i] == '\t') {
        if (field >= 9) {
          std::string nm = chrom_line.substr(s, i - s);
          for (auto &ch : nm) if (ch == '.') ch = '_';  // parse.NormalizeHeader
          fprintf(f, "%s\n", nm.c_str());
        }
        field++; s = i + 1;
      }
    fclose(f);
  }

  // ---- one context per GPU ----
  std::vector<const char *> allow_c, excl_c;
  for (auto &s : config.allowed) allow_c.push_back(s.c_str());
  for (auto &s : config.excluded) excl_c.push_back(s.c_str());
  bvcf_config bc{};
  bc.empty_field = config.emptyField.c_str();
  bc.field_delim = config.fieldDelimiter.c_str();
  bc.keep_id = config.keepID; bc.keep_info = config.keepInfo; bc.keep_pos = config.keepPos;
  bc.want_tsv = !config.noOut; bc.want_dosage = 0;
  bc.allow = allow_c.data(); bc.n_allow = config.allowAll ? -1 : (int)allow_c.size();
  bc.exclude = excl_c.data(); bc.n_exclude = (int)excl_c.size();
  bc.eol_width = eol_width; bc.normalize_dots = 1;
  const int n_slots = 3;
  bc.n_slots = n_slots;
  const size_t cap = 2 * config.chunkBytes + (64u << 20);  // a chunk grows until it holds a newline
  bc.max_chunk_bytes = cap;
  std::vector<bvcf_ctx *> ctxs(config.gpus, nullptr);
  for (int g = 0; g < config.gpus; g++) {
    int rc = bvcf_create(&ctxs[g], g, &bc);
    if (rc) fatal(std::string("bvcf_create(gpu ") + std::to_string(g) + "): " + bvcf_strerror(rc) + " -- a CUDA device is required, there is no CPU fallback");
    rc = bvcf_set_header(ctxs[g], chrom_line.data(), chrom_line.size());
    if (rc) fatal(std::string("bvcf_set_header: ") + bvcf_strerror(rc));
  }

  // ---- chunk loop: pinned ring, n_slots chunks in flight per GPU, rows written in seq order ----
  const size_t ring_n = (size_t)config.gpus * n_slots + 1;
  std::vector<uint8_t *> ring(ring_n, nullptr);
  for (auto &b : ring)
    if (bvcf_host_alloc((void **)&b, cap)) fatal("bvcf_host_alloc failed");
  std::vector<InFlight> q;  // FIFO
  auto collect_one = [&]() {
    InFlight f = q.front();
    q.erase(q.begin());
    const uint8_t *tsv; size_t n; const bvcf_diag *dg; size_t nd;
    int rc = bvcf_collect(ctxs[f.gpu], f.seq, &tsv, &n, nullptr, &dg, &nd, nullptr);
    if (rc) fatal(std::string("bvcf_collect: ") + bvcf_strerror(rc) + " " + bvcf_last_error(ctxs[f.gpu]));
    if (!config.noOut) write_all(out_fd, tsv, n);
    log_diags(f.buf, f.len, dg, nd);
    bvcf_release(ctxs[f.gpu], f.seq);
  };
  uint64_t seq = 0;
  size_t slot = 0;
  size_t fill = head.size() - data_off;  // bytes already in the current buffer
  if (fill > cap) fatal("header buffer larger than a chunk");
  memcpy(ring[0], head.data() + data_off, fill);
  head.clear(); head.shrink_to_fit();
  while (!eof || fill) {
    uint8_t *buf = ring[slot];
    // fill up to chunkBytes (keep reading past it only if no newline has been seen yet)
    while (!eof && fill < config.chunkBytes) {
      ssize_t r = read(in_fd, buf + fill, std::min(cap - fill, config.chunkBytes - fill));
      if (r < 0) { if (errno == EINTR) continue; fatal(std::string("read: ") + strerror(errno)); }
      if (r == 0) { eof = true; break; }
      fill += (size_t)r;
    }
    const uint8_t *last_nl = (const uint8_t *)memrchr(buf, '\n', fill);
    if (!last_nl) {
      if (eof) break;  // an unterminated last line is dropped (main.go:354-357)
      if (fill >= cap) fatal("a single line exceeds the chunk capacity; raise --chunkBytes");
      ssize_t r = read(in_fd, buf + fill, cap - fill);
      if (r <= 0) { eof = true; continue; }
      fill += (size_t)r;
      continue;
    }
    const size_t cut = (last_nl - buf) + 1;
    const size_t next = (slot + 1) % ring_n;
    // the next ring buffer may still belong to an in-flight chunk: drain until it is free
    while (q.size() >= ring_n - 1) collect_one();
    memcpy(ring[next], buf + cut, fill - cut);  // carry the partial line
    const int gpu = (int)(seq % config.gpus);
    size_t on_gpu = 0;
    for (auto &f : q) on_gpu += f.gpu == gpu;
    while (on_gpu >= (size_t)n_slots) {  // FIFO order keeps the output ordered
      on_gpu -= q.front().gpu == gpu;
      collect_one();
    }
    int rc = bvcf_submit(ctxs[gpu], seq, buf, cut);
    if (rc) fatal(std::string("bvcf_submit: ") + bvcf_strerror(rc));
    q.push_back({seq, gpu, buf, cut});
    seq++;
    fill -= cut;
    slot = next;
  }
  while (!q.empty()) collect_one();
  for (auto &b : ring) bvcf_host_free(b);
  for (auto c : ctxs) bvcf_destroy(c);
  if (out_fd != 1) close(out_fd);
  return 0;
}
