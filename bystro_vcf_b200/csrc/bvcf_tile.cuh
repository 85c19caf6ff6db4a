// bvcf_tile.cuh -- north-star kernels (2)+(4) in ONE pass: the per-line fixed-field kernel and the row emitter.
//
// bvcf_tile_kernel: a CTA takes tiles of 128 consecutive records (dynamic ticket), one thread per record.
//   A  compose   FILTER allow/exclude against a shared-memory table (main.go:447-454), getAlleles (main.go:723-1038,
//                the generator of bvcf_rows.cuh), genotype summary lookup, ac == 0 skip (main.go:558), and the TEXT of
//                every row (main.go:586-695) composed with byte stores into a shared-memory arena: fixed columns,
//                float text, and -- for records with at most SMALL_EVENTS event words -- the sample-name lists too
//                (main.go:617,639,653), so that such a row is one contiguous string.  Long lists stay holes.
//   B  offsets   block scan of the records' output bytes / rows / locus bytes, then a decoupled look-back over
//                the tiles before this one (128-bit descriptors: flag | bytes, rows) gives the tile's place in the
//                output: no size pass, no separate prefix kernels, the getAlleles generator runs once.
//   C  copy-out  every thread copies its rows from the arena to the output with aligned 8-byte stores (funnel-
//                shifted shared-memory words), appends the INFO span straight from the input line, writes locus
//                strings + the small rows' dosages (main.go:576-584), and queues the rows with holes for the names
//                kernels (bvcf_names.cuh) as RowDesc entries.
// Records that do not fit the arena (very long alleles / IDs, dozens of rows per record) take the slow path: sized in
// A by the same code with stores switched off, written in C by their own thread straight to global memory.
//
// Replaces round 1's bvcf_rows_kernel<SIZE> + 3 prefix kernels + bvcf_rows_kernel<EMIT> + bvcf_names_kernel
// (+ bvcf_rows_list_kernel x2, bvcf_dosage_zero_kernel): the SIZE/EMIT pair ran the generator twice and spent 60 % of
// its instructions in an 8-byte register writer with unaligned global stores.
#pragma once
#include "bvcf_rows.cuh"

namespace bvcf {

constexpr int TILE_THREADS = 128;          // records per tile, one thread each
constexpr uint32_t TILE_ARENA = 32768;     // staging bytes per CTA
constexpr uint32_t TILE_ROWS = 256;        // staged rows per tile; thread t's first row is rows[t]
constexpr uint32_t ROW_NONE = 0xFFFFu;

struct TileParams {
  const uint8_t *in;
  DevCfg cfg;
  const LineRec *lines;        // dense, input order
  const uint32_t *events;      // sub-chunk event buffer
  const LineStats *stats;      // per record, ALT #1..3 (null when there are no samples)
  uint8_t *out;
  unsigned long long out_cap;
  RunCounters *ctr;
  ulonglong2 *tile_state;      // two descriptors per tile (zeroed before the launch): {flag | bytes, rows}, {flag | locus bytes, 0}
  RowDesc *row_desc;           // work list for the names kernels (see RowDesc)
  unsigned long long row_desc_cap;
  uint32_t long_words;         // rows of records with more event words go to the END of row_desc (CTA-per-row kernel); 0: none
  // dosage matrix (main.go:576-584)
  int8_t *dosage;              // rows x n_samples
  unsigned long long dosage_cap_rows;
  uint8_t *loci;               // locus strings "chrom:pos:ref:alt", back to back
  unsigned long long loci_cap;
  unsigned long long *loci_off;  // dosage_cap_rows entries: start of row r's locus in `loci`
  DiagSink diag;
};

// one staged row (shared memory)
struct __align__(16) TRow {
  uint32_t hole_len[3];   // bytes of the het / hom / missing list when it is NOT staged (filled by the names kernels)
  uint32_t cnt[3];        // names per list
  uint16_t hole_pos[3];   // where the hole sits in the staged string (staged bytes before it)
  uint16_t soff;          // arena offset of the staged string: TSV bytes, then the locus
  uint16_t slen;          // staged TSV bytes
  uint16_t loc_len;       // staged locus bytes
  uint16_t next;          // next row of the same record, ROW_NONE at the end
  uint16_t allele;        // ALT number
  uint16_t flags;         // bit 0: the INFO span and the EOL follow the staged bytes; bit 1: queue a RowDesc
  uint16_t pad[3];
};
static_assert(sizeof(TRow) == 48, "TRow layout");

__device__ __forceinline__ void sts8(uint32_t a, uint32_t v) { asm volatile("st.shared.u8 [%0], %1;\n" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ uint32_t tl_lds8(uint32_t a) {
  uint32_t v;
  asm volatile("ld.shared.u8 %0, [%1];\n" : "=r"(v) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ uint32_t tl_lds32(uint32_t a) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];\n" : "=r"(v) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ void st_state(ulonglong2 *p, unsigned long long x, unsigned long long y) {
  asm volatile("st.relaxed.gpu.global.v2.u64 [%0], {%1, %2};\n" ::"l"(p), "l"(x), "l"(y) : "memory");
}
__device__ __forceinline__ ulonglong2 ld_state(const ulonglong2 *p) {
  ulonglong2 v;
  asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];\n" : "=l"(v.x), "=l"(v.y) : "l"(p) : "memory");
  return v;
}

// n bytes from shared memory (address sa, any alignment) to global memory (any alignment): byte stores up to the
// first 8-byte boundary of the destination, then aligned 8-byte stores of funnel-shifted shared-memory words.
// May read up to 11 bytes past the source (the arena is padded).
__device__ __forceinline__ void copy_s2g(uint8_t *g, uint32_t sa, uint32_t n) {
  while (n && ((uintptr_t)g & 7u)) { *g++ = (uint8_t)tl_lds8(sa++); n--; }
  if (n >= 8) {
    const uint32_t sh = (sa & 3u) * 8u;
    uint32_t wa = sa & ~3u;
    uint32_t w0 = tl_lds32(wa);
    do {
      const uint32_t w1 = tl_lds32(wa + 4), w2 = tl_lds32(wa + 8);
      *reinterpret_cast<uint2 *>(g) = make_uint2(__funnelshift_r(w0, w1, sh), __funnelshift_r(w1, w2, sh));
      w0 = w2; wa += 8; g += 8; sa += 8; n -= 8;
    } while (n >= 8);
  }
  while (n) { *g++ = (uint8_t)tl_lds8(sa++); n--; }
}
// the same from global memory (a span of the input line; the input region is padded, so whole words may be read)
__device__ __forceinline__ void copy_g2g(uint8_t *g, const uint8_t *s, uint32_t n) {
  while (n && ((uintptr_t)g & 7u)) { *g++ = *s++; n--; }
  if (n >= 8) {
    const uint32_t sh = (uint32_t)((uintptr_t)s & 3u) * 8u;
    const uint32_t *wp = reinterpret_cast<const uint32_t *>((uintptr_t)s & ~(uintptr_t)3);
    uint32_t w0 = wp[0];
    do {
      const uint32_t w1 = wp[1], w2 = wp[2];
      *reinterpret_cast<uint2 *>(g) = make_uint2(__funnelshift_r(w0, w1, sh), __funnelshift_r(w1, w2, sh));
      w0 = w2; wp += 2; g += 8; s += 8; n -= 8;
    } while (n >= 8);
  }
  while (n) { *g++ = *s++; n--; }
}

// ---- byte sinks of the row composer ---------------------------------------------------------------------
// StageWriter: shared memory while `stg`, a plain byte count otherwise (slow-path records are sized with it).
struct StageWriter {
  static constexpr bool kStage = true;
  uint32_t a;   // shared-memory address of the next byte
  bool stg;
  __device__ __forceinline__ void byte(uint32_t c) { if (stg) sts8(a, c); a++; }
  __device__ __forceinline__ void packed(unsigned long long chars, int len) {  // len <= 8 characters, little-endian
    if (stg) {
#pragma unroll
      for (int k = 0; k < 8; k++) if (k < len) sts8(a + k, (uint32_t)(chars >> (8 * k)) & 0xFFu);
    }
    a += len;
  }
  __device__ __forceinline__ void span(const uint8_t *p, int len) {
    if (stg) for (int i = 0; i < len; i++) sts8(a + i, p[i]);
    a += len;
  }
};
// GlobalWriter: the slow path's second pass, byte stores straight to the output (`on` false: the output region
// is too small, the host will re-run the chunk)
struct GlobalWriter {
  static constexpr bool kStage = false;
  uint8_t *g;
  bool on;
  __device__ __forceinline__ void byte(uint32_t c) { if (on) *g = (uint8_t)c; g++; }
  __device__ __forceinline__ void packed(unsigned long long chars, int len) {
    if (on) for (int k = 0; k < len; k++) g[k] = (uint8_t)(chars >> (8 * k));
    g += len;
  }
  __device__ __forceinline__ void span(const uint8_t *p, int len) {
    if (on) for (int i = 0; i < len; i++) g[i] = p[i];
    g += len;
  }
};
template <class W>
__device__ __forceinline__ void w_dec(W &w, long long v) {  // strconv.Itoa
  unsigned long long lo, hi;
  const int len = itoa_pack(v, lo, hi);  // registers, no byte buffer in local memory
  if (len >= 0) {
    w.packed(lo, len < 8 ? len : 8);
    if (len > 8) w.packed(hi, len - 8);
  } else {
    uint8_t buf[24];
    const int l2 = itoa_dec(v, buf);
    for (int i = 0; i < l2; i++) w.byte(buf[i]);
  }
}

// what one record contributes to its tile
struct RecOut {
  unsigned long long bytes;   // TSV bytes of all its rows
  uint32_t rows, loci;        // rows, locus bytes
  uint32_t n_desc;            // RowDesc entries it will queue
  uint32_t first, last;       // its staged rows (chain through TRow::next)
  bool failed;                // slow path: nothing staged, the sizes above are still exact
};
struct TileShared {
  uint32_t arena_s;           // shared-memory address of the arena
  TRow *rows;
  uint32_t *arena_cur, *rows_cur;
};
// slow-path pass 2: where the record's rows, loci and RowDesc entries go
struct SlowOut {
  unsigned long long row;       // next row number within the sub-chunk
  unsigned long long loci_off;  // next locus byte (absolute in `loci`)
  uint32_t desc;                // next RowDesc ordinal of this record's list
  bool is_long, desc_ok;
  uint32_t big_base, long_base;
};

__device__ __forceinline__ void queue_row_desc(const TileParams &p, bool is_long, uint32_t ord, uint32_t big_base, uint32_t long_base,
                                               const RowDesc &rd) {
  const unsigned long long slot = is_long ? p.row_desc_cap - 1ull - (long_base + ord) : (unsigned long long)big_base + ord;
  p.row_desc[slot] = rd;
}

// the names of a short record's row into the staged lists (main.go:617,639,653 strings.Join): one walk over the
// record's quad events (at most SMALL_EVENTS / 2), header order
__device__ __forceinline__ void fill_small_lists(const TileParams &p, const LineRec &rec, const LineCtx &lc, uint32_t a,
                                                 uint32_t la0, uint32_t la1, uint32_t la2) {
  const DevCfg &cfg = p.cfg;
  const uint32_t *ev = p.events + rec.ev_start;
  const bool simple = !(rec.flags & 1) && a == 1;
  const uint32_t dl = (uint32_t)cfg.delim_len;
  uint32_t nh = 0, no = 0, nm = 0;
  for (uint32_t q = 0; 2 * q + 1 < rec.ev_count; q++) {
    const uint2 e = *reinterpret_cast<const uint2 *>(ev + 2 * q);
    uint32_t mh, mo, mm;
    quad_masks(e.x, e.y, a, simple, lc.L, lc.content_len, true, mh, mo, mm);
    const uint32_t s0 = (e.x & EV_SAMPLE_MASK) - EV_BASE_BIAS;
    uint32_t any = mh | mo | mm;
    while (any) {
      const uint32_t bit = any & (0u - any);
      any &= any - 1;
      const uint32_t samp = s0 + ((uint32_t)(__ffs(bit) - 1) >> 2);
      const bool is_h = (mh & bit) != 0, is_o = (mo & bit) != 0;
      uint32_t d = is_h ? la0 : (is_o ? la1 : la2);
      const uint32_t rn = is_h ? nh : (is_o ? no : nm);
      if (rn > 0) {
        for (uint32_t i = 0; i < dl; i++) sts8(d + i, cfg.delim[i]);
        d += dl;
      }
      const uint32_t nl = name_len(cfg, samp);
      if (cfg.name8) {
        const unsigned long long it = cfg.name8[samp];
#pragma unroll
        for (int i = 0; i < 7; i++) sts8(d + i, (uint32_t)(it >> (8 * i)) & 0xFFu);
      } else {
        const uint8_t *src = name_ptr(cfg, samp);
        for (uint32_t i = 0; i < nl; i++) sts8(d + i, src[i]);
      }
      d += nl;
      if (is_h) { nh++; la0 = d; } else if (is_o) { no++; la1 = d; } else { nm++; la2 = d; }
    }
  }
}

// the int8 dosages of a short record's row (main.go:1172-1178): -1 missing, else min(alleles equal to the row's, 127);
// the row was zeroed by the tile
__device__ __forceinline__ void small_dosage(const TileParams &p, const LineRec &rec, uint32_t a, int8_t *drow) {
  const DevCfg &cfg = p.cfg;
  const uint32_t *ev = p.events + rec.ev_start;
  const uint8_t *L = p.in + rec.start;
  const uint32_t content_len = rec.len >= (uint32_t)cfg.eol_width ? rec.len - (uint32_t)cfg.eol_width : 0;
  const bool simple = !(rec.flags & 1) && a == 1;
  for (uint32_t q = 0; 2 * q + 1 < rec.ev_count; q++) {
    const uint2 e = *reinterpret_cast<const uint2 *>(ev + 2 * q);
    uint32_t mh, mo, mm;
    quad_masks(e.x, e.y, a, simple, L, content_len, true, mh, mo, mm);
    const uint32_t s0 = (e.x & EV_SAMPLE_MASK) - EV_BASE_BIAS;
    uint32_t any = mh | mo | mm;
    while (any) {
      const uint32_t bit = any & (0u - any);
      any &= any - 1;
      const uint32_t samp = s0 + ((uint32_t)(__ffs(bit) - 1) >> 2);
      const bool is_h = (mh & bit) != 0, is_o = (mo & bit) != 0;
      int v = -1;
      if (is_h | is_o) {
        if (e.x & EV_COMPLEX) {
          uint32_t gt, alt;
          classify_gt_general(L + e.y, content_len > e.y ? content_len - e.y : 0, a, gt, alt);
          v = alt > 127 ? 127 : (int)alt;
        } else {
          const uint32_t sh = (uint32_t)(__ffs(bit) - 1) - 3u;  // 4 * slot
          const bool hap = ((e.y >> (16 + sh)) & 0xFu) == EV_NIB_ABSENT;
          v = is_h ? 1 : (hap ? 1 : 2);
        }
      }
      drow[samp] = (int8_t)v;
    }
  }
}

// ---- one output row (main.go:555-695) -------------------------------------------------------------------
// W = StageWriter: pass A (compose into the arena, or size only once the record has failed over to the slow path);
// W = GlobalWriter: pass C of a slow-path record.
template <class W>
__device__ __forceinline__ void tile_emit_row(const TileParams &p, const LineRec &rec, const LineCtx &lc, const OutAllele &oa,
                                              GtStats &gs, int &gs_idx, W &w, RecOut &ro, const TileShared &sh, SlowOut &so) {
  const DevCfg &cfg = p.cfg;
  const uint32_t a = (uint32_t)oa.alt_idx + 1;
  const bool has_samples = cfg.n_samples > 0;
  if (has_samples) {
    if (gs_idx != oa.alt_idx) {  // MNP bases share their ALT index: reduce once (main.go:865-868)
      if (cfg.name_fixed_w > 0 && !(rec.flags & 1)) {
        // the scan kernel's inline summary is complete: only ALT #1 occurs among the samples
        gs.n_het = a == 1 ? rec.n_het1 : 0; gs.n_hom = a == 1 ? rec.n_hom1 : 0; gs.ac = a == 1 ? rec.ac1 : 0;
        gs.n_miss = rec.n_miss; gs.an = rec.an;
        gs.het_bytes = gs.n_het * cfg.name_fixed_w; gs.hom_bytes = gs.n_hom * cfg.name_fixed_w;
        gs.miss_bytes = gs.n_miss * cfg.name_fixed_w;
      } else if (a <= (uint32_t)STAT_ALLELES) {
        const LineStats &ls = p.stats[lc.li];
        gs.n_het = ls.n_het[a - 1]; gs.n_hom = ls.n_hom[a - 1]; gs.ac = ls.ac[a - 1];
        gs.het_bytes = ls.het_bytes[a - 1]; gs.hom_bytes = ls.hom_bytes[a - 1];
        gs.n_miss = ls.n_miss; gs.an = ls.an; gs.miss_bytes = ls.miss_bytes;
      } else {
        gs = reduce_events_thread(cfg, rec, p.events + rec.ev_start, lc.L, lc.content_len, a);
      }
      gs_idx = oa.alt_idx;
    }
    if (gs.ac == 0) return;  // main.go:558
  }
  const uint32_t dl = (uint32_t)cfg.delim_len;
  const uint32_t cnts[3] = {has_samples ? gs.n_het : 0u, has_samples ? gs.n_hom : 0u, has_samples ? gs.n_miss : 0u};
  const uint32_t nb[3] = {gs.het_bytes, gs.hom_bytes, gs.miss_bytes};
  uint32_t lb[3];  // list bytes
#pragma unroll
  for (int k = 0; k < 3; k++) lb[k] = (cfg.want_tsv && cnts[k]) ? nb[k] + (cnts[k] - 1) * dl : 0u;
  const bool big = has_samples && rec.ev_count > SMALL_EVENTS;  // lists and dosages are left to the names kernels
  const bool want_locus = cfg.want_dosage && has_samples;
  const uint32_t tail = (cfg.want_tsv && cfg.keep_info) ? (uint32_t)lc.info_n + 1u : 0u;  // INFO span + EOL, never staged

  // ---- where the bytes go ----
  TRow *row = nullptr;
  uint32_t ri = ROW_NONE, a0 = 0;
  bool inline_lists = false;
  if constexpr (W::kStage) {
    if (!ro.failed) {
      // upper bound of the staged bytes: the constant pieces come to 100 bytes at most (3 "chr", 14 "\tMULTIALLELIC\t",
      // 6 ref/alt/trTv framing, 3 x 10 ratio columns, 30 ac/an/sampleMaf, 14 keepInfo framing, tabs, EOL)
      const uint32_t pos_b = oa.pos_verbatim ? (uint32_t)lc.pos_n : 20u;
      const uint32_t alt_b = oa.kind == 1 ? (uint32_t)oa.ins_n + 1u : 21u;
      uint32_t need = 0;
      if (cfg.want_tsv) {
        need = 104u + (uint32_t)lc.chrom_n + pos_b + alt_b + (has_samples ? 3u * (uint32_t)cfg.empty_len : (uint32_t)cfg.tail0_len) +
               (cfg.keep_pos ? (uint32_t)lc.pos_n : 0u) + (cfg.keep_id ? (uint32_t)lc.id_n : 0u);
        if (!big) need += lb[0] + lb[1] + lb[2];
      }
      if (want_locus) need += 8u + (uint32_t)lc.chrom_n + pos_b + alt_b;
      ri = ro.first == ROW_NONE ? threadIdx.x : atomicAdd(sh.rows_cur, 1u);
      const uint32_t off = atomicAdd(sh.arena_cur, (need + 15u) & ~15u);
      if (ri >= TILE_ROWS || off + need > TILE_ARENA) {
        ro.failed = true; ro.first = ROW_NONE;  // the whole record takes the slow path; keep sizing
      } else {
        row = &sh.rows[ri];
        a0 = sh.arena_s + off;
        inline_lists = !big && cfg.want_tsv;
      }
    }
    w.stg = row != nullptr;
    w.a = a0;
  }
  uint8_t *g_row0 = nullptr;
  if constexpr (!W::kStage) g_row0 = w.g;
  unsigned long long hole_total = 0;
  uint32_t la[3] = {0, 0, 0};                 // staged list starts (inline lists)
  uint32_t hpos[3] = {0, 0, 0}, hlen[3] = {0, 0, 0};
  unsigned long long dsts[3] = {~0ull, ~0ull, ~0ull};

  if (cfg.want_tsv) {
    // chrom (main.go:570-574)
    if (lc.chrom_n < 4 || lc.chrom[0] != 'c') w.packed(0x726863ull, 3);  // "chr"
    w.span(lc.chrom, lc.chrom_n);
    w.byte('\t');
    if (oa.pos_verbatim) w.span(lc.pos, lc.pos_n); else w_dec(w, oa.pos_val);
    if (lc.site_type != T_MULTI) {  // "\tSNP\t": the three-letter types as one piece
      const uint8_t *tt = (const uint8_t *)TYPE_TXT[lc.site_type];
      w.packed(0x09ull | ((uint64_t)tt[0] << 8) | ((uint64_t)tt[1] << 16) | ((uint64_t)tt[2] << 24) | (0x09ull << 32), 5);
    } else {
      w.byte('\t');
      w.span((const uint8_t *)TYPE_TXT[lc.site_type], TYPE_LEN[lc.site_type]);
      w.byte('\t');
    }
    const uint8_t trtv = lc.multi ? '0' : (oa.kind == 0 ? trtv_char(oa.ref, oa.alt_c) : '0');  // main.go:602-606
    if (oa.kind == 0) {  // "R\tA\tt\t" as one piece
      w.packed((uint64_t)oa.ref | (0x09ull << 8) | ((uint64_t)oa.alt_c << 16) | (0x09ull << 24) | ((uint64_t)trtv << 32) | (0x09ull << 40), 6);
    } else {
      w.byte(oa.ref);
      w.byte('\t');
      if (oa.kind == 1) { w.byte('+'); w.span(oa.ins_p, oa.ins_n); }
      else w_dec(w, oa.del_n);
      w.byte('\t');
      w.byte(trtv);
      w.byte('\t');
    }
    if (!has_samples) {  // main.go:612-616,634-637,648-651,667: "! 0 ! 0 ! 0 0 0 0"
      w.span(cfg.tail0, cfg.tail0_len);  // composed once by the host
    } else {
      const uint32_t eff = (uint32_t)cfg.n_samples - gs.n_miss;  // main.go:563
      const uint32_t den[3] = {eff, eff, (uint32_t)cfg.n_samples};
#pragma unroll
      for (int k = 0; k < 3; k++) {
        if (cnts[k] == 0) {
          w.span(cfg.empty, cfg.empty_len); w.byte('\t'); w.byte('0');
        } else {
          if constexpr (W::kStage) {
            if (inline_lists) { la[k] = w.a; w.a += lb[k]; }      // filled below
            else { hpos[k] = w.a - a0; hlen[k] = lb[k]; hole_total += lb[k]; }
          } else {
            dsts[k] = (unsigned long long)(w.g - p.out);
            w.g += lb[k];                                          // filled by the names kernels
          }
          w.byte('\t');
          int fl;
          const uint64_t ft = format_ratio_g3(cnts[k], den[k], fl);
          w.packed(ft, fl);
        }
        w.byte('\t');
      }
      w_dec(w, gs.ac); w.byte('\t');
      w_dec(w, gs.an); w.byte('\t');
      if (gs.ac == 0) w.byte('0');
      else { int fl; const uint64_t ft = format_ratio_g3(gs.ac, gs.an, fl); w.packed(ft, fl); }
    }
    if (cfg.keep_pos) { w.byte('\t'); w.span(lc.pos, lc.pos_n); }
    if (cfg.keep_id) { w.byte('\t'); w.span(lc.id, lc.id_n); }
    if (cfg.keep_info) {
      w.byte('\t'); w_dec(w, oa.alt_idx); w.byte('\t');
      if constexpr (!W::kStage) { w.span(lc.info, lc.info_n); w.byte('\n'); }  // staged rows: appended at copy-out
    } else {
      w.byte('\n');
    }
  }

  uint32_t slen = 0, loc_len = 0;
  if constexpr (W::kStage) slen = w.a - a0;
  // ---- locus "chrom:pos:ref:alt" (main.go:577) ----
  if (want_locus) {
    if constexpr (W::kStage) {
      if (lc.chrom_n < 4 || lc.chrom[0] != 'c') w.packed(0x726863ull, 3);
      w.span(lc.chrom, lc.chrom_n);
      w.byte(':');
      if (oa.pos_verbatim) w.span(lc.pos, lc.pos_n); else w_dec(w, oa.pos_val);
      w.byte(':'); w.byte(oa.ref); w.byte(':');
      if (oa.kind == 0) w.byte(oa.alt_c);
      else if (oa.kind == 1) { w.byte('+'); w.span(oa.ins_p, oa.ins_n); }
      else w_dec(w, oa.del_n);
      loc_len = w.a - a0 - slen;
    } else {
      const unsigned long long gr = p.ctr->chunk_row_base + so.row;
      GlobalWriter lw;
      lw.g = p.loci + so.loci_off;
      lw.on = w.on && gr < p.dosage_cap_rows;
      uint8_t *const l0 = lw.g;
      // bytes beyond the locus buffer are counted, not written (the host grows the buffer and re-runs the chunk)
      unsigned long long room = so.loci_off < p.loci_cap ? p.loci_cap - so.loci_off : 0ull;
      const uint32_t worst = 8u + (uint32_t)lc.chrom_n + (oa.pos_verbatim ? (uint32_t)lc.pos_n : 20u) + (oa.kind == 1 ? (uint32_t)oa.ins_n + 1u : 21u);
      if (room < worst) lw.on = false;
      if (lc.chrom_n < 4 || lc.chrom[0] != 'c') lw.packed(0x726863ull, 3);
      lw.span(lc.chrom, lc.chrom_n);
      lw.byte(':');
      if (oa.pos_verbatim) lw.span(lc.pos, lc.pos_n); else w_dec(lw, oa.pos_val);
      lw.byte(':'); lw.byte(oa.ref); lw.byte(':');
      if (oa.kind == 0) lw.byte(oa.alt_c);
      else if (oa.kind == 1) { lw.byte('+'); lw.span(oa.ins_p, oa.ins_n); }
      else w_dec(lw, oa.del_n);
      loc_len = (uint32_t)(lw.g - l0);
      if (w.on && gr < p.dosage_cap_rows) p.loci_off[gr] = so.loci_off;
      so.loci_off += loc_len;
    }
  }

  // ---- row bookkeeping ----
  if constexpr (W::kStage) {
    ro.bytes += (unsigned long long)slen + hole_total + tail;
    ro.rows++;
    ro.loci += loc_len;
    if (big) ro.n_desc++;
    if (row) {
      row->hole_len[0] = hlen[0]; row->hole_len[1] = hlen[1]; row->hole_len[2] = hlen[2];
      row->cnt[0] = cnts[0]; row->cnt[1] = cnts[1]; row->cnt[2] = cnts[2];
      row->hole_pos[0] = (uint16_t)hpos[0]; row->hole_pos[1] = (uint16_t)hpos[1]; row->hole_pos[2] = (uint16_t)hpos[2];
      row->soff = (uint16_t)(a0 - sh.arena_s); row->slen = (uint16_t)slen; row->loc_len = (uint16_t)loc_len;
      row->next = (uint16_t)ROW_NONE; row->allele = (uint16_t)a;
      row->flags = (uint16_t)((tail ? 1u : 0u) | (big ? 2u : 0u));
      if (ro.first == ROW_NONE) ro.first = ri; else sh.rows[ro.last].next = (uint16_t)ri;
      ro.last = ri;
      if (inline_lists && (cnts[0] | cnts[1] | cnts[2])) fill_small_lists(p, rec, lc, a, la[0], la[1], la[2]);
    }
  } else {
    if (has_samples) {  // every row of a slow-path record is queued for the names kernels
      if (so.desc_ok) {
        RowDesc rd;
        rd.line = lc.li; rd.allele = a;
        rd.het_dst = dsts[0]; rd.hom_dst = dsts[1]; rd.miss_dst = dsts[2];
        rd.n_het = cnts[0]; rd.n_hom = cnts[1]; rd.n_miss = cnts[2];
        rd.row = (uint32_t)so.row;
        queue_row_desc(p, so.is_long, so.desc, so.big_base, so.long_base, rd);
      }
      so.desc++;
    }
    so.row++;
    (void)g_row0;
  }
}

// ---- one record: field index, linePasses, getAlleles, one tile_emit_row per output allele ----------------
template <class W>
__device__ __forceinline__ void tile_record(const TileParams &p, uint32_t li, const LineRec &rec, W &w, RecOut &ro,
                                            const TileShared &sh, SlowOut &so, const uint8_t *s_filt, const uint32_t *s_filt_off,
                                            bool diag, const uint8_t *&info_p, uint32_t &info_n) {
  const DevCfg &cfg = p.cfg;
  const int n_filt = cfg.n_allow + cfg.n_excl;
  const uint8_t *L = p.in + rec.start;
  const uint32_t n = rec.len >= (uint32_t)cfg.eol_width ? rec.len - (uint32_t)cfg.eol_width : 0;  // main.go:535
  // ---- first eight/nine tabs (strings.Split, main.go:535) ----
  const int need = cfg.H - 1 < 9 ? cfg.H - 1 : 9;
  uint32_t t[9];
  int found = need;
  bool far = false;  // a tab beyond 64 KiB from the line start: the scan kernel could not record it
#pragma unroll
  for (int k = 0; k < 9; k++) {
    t[k] = rec.tab[k];
    far = far || (k < need && t[k] == 0xFFFFu);
  }
  if (far) {
    found = 0;
    for (uint32_t i = 0; i < n && found < need; i++)
      if (L[i] == '\t') {
#pragma unroll
        for (int k = 0; k < 9; k++) if (k == found) t[k] = i;  // static indexing keeps t[] in registers
        found++;
      }
  }
  bool pass = found >= need;  // always true for scan-kernel records; defensive
#pragma unroll
  for (int k = 0; k < 9; k++) if (k >= found) t[k] = n;
  LineCtx lc;
  lc.L = L; lc.content_len = n; lc.li = li;
  lc.chrom = L; lc.chrom_n = (int)t[0];
  lc.pos = L + t[0] + 1; lc.pos_n = (int)(t[1] - t[0] - 1);
  lc.id = L + t[1] + 1; lc.id_n = (int)(t[2] - t[1] - 1);
  const uint8_t *ref = L + t[2] + 1; const int ref_n = (int)(t[3] - t[2] - 1);
  const uint8_t *alt = L + t[3] + 1; const int alt_n = (int)(t[4] - t[3] - 1);
  const uint8_t *filt = L + t[5] + 1; const int filt_n = (int)(t[6] - t[5] - 1);
  lc.info = L + t[6] + 1; lc.info_n = (int)(t[7] - t[6] - 1);
  lc.multi = false; lc.site_type = T_SNP;
  info_p = lc.info; info_n = (uint32_t)lc.info_n;

  // ---- linePasses (main.go:447-454): exact whole-field match against the shared-memory table ----
  if (pass && (!cfg.allow_all || cfg.n_excl > 0)) {
    bool in_allow = false, in_excl = false;
    for (int k = 0; k < n_filt; k++) {
      const uint32_t o = s_filt_off[k], ln = s_filt_off[k + 1] - o;
      bool eq = (int)ln == filt_n;
      for (int i = 0; eq && i < filt_n; i++) eq = s_filt[o + i] == filt[i];
      if (eq) { if (k < cfg.n_allow) in_allow = true; else in_excl = true; }
    }
    if (!cfg.allow_all && !in_allow) pass = false;
    if (in_excl) pass = false;
  }

  // ---- getAlleles (main.go:723-1038) as a resumable generator: one converged tile_emit_row call site ----
  AlleleGen g;
  g.ref = ref; g.alt = alt; g.ref_n = ref_n; g.alt_n = alt_n;
  g.s = 0; g.alt_idx = 0; g.mnp_i = -1; g.ta = alt; g.tn = 0; g.last = false;
  g.done = !(pass && ref_n > 0 && alt_n > 0);
  g.ipos = 0;
  g.pos_ok = g.done ? false : atoi_go(lc.pos, lc.pos_n, g.ipos);
  const unsigned long long line_no = p.ctr->chunk_line_base + rec.ord;
  if (!g.done) gen_begin(g, lc, p.diag, line_no, diag);
  GtStats gs;
  gs.n_het = gs.n_hom = gs.n_miss = gs.ac = gs.an = gs.het_bytes = gs.hom_bytes = gs.miss_bytes = 0;
  int gs_idx = -1;
  OutAllele oa;
  oa.ins_p = nullptr; oa.ins_n = 0; oa.del_n = 0; oa.pos_val = 0;
  while (gen_next(g, oa, p.diag, line_no, diag)) tile_emit_row<W>(p, rec, lc, oa, gs, gs_idx, w, ro, sh, so);
}

// inclusive scan of v over the CTA's 128 threads (4 warps); total returned in `total`
__device__ __forceinline__ unsigned long long block_scan64(unsigned long long v, unsigned long long *s_w, unsigned long long &total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned long long x = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const unsigned long long t = __shfl_up_sync(FULL, x, d);
    if (lane >= d) x += t;
  }
  if (lane == 31) s_w[warp] = x;
  __syncthreads();
  unsigned long long off = 0, tot = 0;
#pragma unroll
  for (int k = 0; k < TILE_THREADS / 32; k++) {
    const unsigned long long wv = s_w[k];
    if (k < warp) off += wv;
    tot += wv;
  }
  __syncthreads();
  total = tot;
  return x + off;
}

// decoupled look-back over descriptor q (0: bytes + rows, 1: locus bytes) of the tiles before `tile`; one warp.
// Publishes this tile's aggregate, then its inclusive prefix.  Returns the exclusive prefix.
constexpr unsigned long long TS_AGG = 1ull << 62, TS_PFX = 2ull << 62, TS_VAL = (1ull << 62) - 1ull;
__device__ __forceinline__ void tile_lookback(ulonglong2 *state, uint32_t tile, int q, unsigned long long agg_x,
                                              unsigned long long agg_y, unsigned long long &ex_x, unsigned long long &ex_y, int lane) {
  ex_x = 0; ex_y = 0;
  if (tile == 0) {
    if (lane == 0) st_state(&state[q], TS_PFX | agg_x, agg_y);
    return;
  }
  if (lane == 0) st_state(&state[2ull * tile + q], TS_AGG | agg_x, agg_y);
  long long j = (long long)tile - 1;
  for (;;) {
    const long long idx = j - lane;
    ulonglong2 v;
    v.x = TS_PFX; v.y = 0;  // before the first tile: an inclusive prefix of nothing
    if (idx >= 0) {
      do { v = ld_state(&state[2ull * (unsigned long long)idx + q]); } while ((v.x >> 62) == 0);  // that tile is running: it took its ticket before ours
    }
    const uint32_t pm = __ballot_sync(FULL, (v.x >> 62) == 2ull);
    const int cut = pm ? __ffs(pm) - 1 : 31;  // the nearest tile that already knows its inclusive prefix
    ex_x += warp_sum64(lane <= cut ? (v.x & TS_VAL) : 0ull);
    ex_y += warp_sum64(lane <= cut ? v.y : 0ull);
    if (pm) break;
    j -= 32;
  }
  if (lane == 0) st_state(&state[2ull * tile + q], TS_PFX | (ex_x + agg_x), ex_y + agg_y);
}

__global__ void __launch_bounds__(TILE_THREADS, 4) bvcf_tile_kernel(const __grid_constant__ TileParams p) {
  __shared__ __align__(16) uint8_t s_arena[TILE_ARENA + 16];
  __shared__ TRow s_rows[TILE_ROWS];
  __shared__ uint8_t s_filt[FILT_SMEM];
  __shared__ uint32_t s_filt_off[65];
  __shared__ unsigned long long s_w[TILE_THREADS / 32];
  __shared__ unsigned long long s_base[4];   // tile bases: bytes, rows, locus bytes
  __shared__ uint32_t s_cur[2];              // arena bytes in use, next free row descriptor
  __shared__ uint32_t s_misc[4];             // tile ticket, big_base, long_base, desc_ok
  const DevCfg &cfg = p.cfg;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // FILTER allow/exclude table -> shared memory
  const int n_filt = cfg.n_allow + cfg.n_excl;
  for (int i = threadIdx.x; i < cfg.filt_bytes && i < FILT_SMEM; i += blockDim.x) s_filt[i] = cfg.filt_blob[i];
  for (int i = threadIdx.x; i <= n_filt && i < 65; i += blockDim.x) s_filt_off[i] = cfg.filt_off[i];
  if (p.ctr->ev_overflow | p.ctr->slot_overflow) return;  // the host grows the scratch and re-runs the chunk
  const uint32_t n_rec = p.ctr->chunk_records;
  const uint32_t n_tiles = (n_rec + TILE_THREADS - 1) / TILE_THREADS;
  const unsigned long long out_base = p.ctr->chunk_out_base, row0 = p.ctr->chunk_row_base, loci0 = p.ctr->chunk_loci_base;
  const bool has_samples = cfg.n_samples > 0;
  const bool want_locus = cfg.want_dosage && has_samples;
  TileShared sh;
  sh.arena_s = (uint32_t)__cvta_generic_to_shared(s_arena);
  sh.rows = s_rows;
  sh.arena_cur = &s_cur[0];
  sh.rows_cur = &s_cur[1];

  for (;;) {
    __syncthreads();  // the previous tile is done with the arena
    if (threadIdx.x == 0) {
      s_misc[0] = atomicAdd(&p.ctr->tile_ticket, 1u);
      s_cur[0] = 0; s_cur[1] = TILE_THREADS;
    }
    __syncthreads();
    const uint32_t tile = s_misc[0];
    if (tile >= n_tiles) break;
    const uint32_t li = tile * TILE_THREADS + threadIdx.x;
    const bool valid = li < n_rec;

    // ---- A: compose ----
    RecOut ro;
    ro.bytes = 0; ro.rows = 0; ro.loci = 0; ro.n_desc = 0; ro.first = ROW_NONE; ro.last = ROW_NONE; ro.failed = false;
    SlowOut so;
    so.row = 0; so.loci_off = 0; so.desc = 0; so.is_long = false; so.desc_ok = false; so.big_base = 0; so.long_base = 0;
    LineRec rec;
    rec.start = 0; rec.len = 0; rec.an = 0; rec.ev_start = 0; rec.ev_count = 0; rec.ord = 0; rec.flags = 0;
    rec.n_het1 = rec.n_hom1 = rec.n_miss = rec.ac1 = 0;
    const uint8_t *info_p = nullptr;
    uint32_t info_n = 0;
    if (valid) {
      rec = p.lines[li];
      StageWriter w;
      w.a = 0; w.stg = false;
      tile_record<StageWriter>(p, li, rec, w, ro, sh, so, s_filt, s_filt_off, true, info_p, info_n);
    }
    const bool is_long = p.long_words && rec.ev_count > p.long_words;
    if (ro.failed && has_samples) ro.n_desc = ro.rows;

    // ---- B: offsets within the tile, then the tile's place in the output ----
    unsigned long long tot_b, tot_rl, tot_d;
    const unsigned long long in_b = block_scan64(ro.bytes, s_w, tot_b);
    const unsigned long long in_rl = block_scan64((unsigned long long)ro.rows | ((unsigned long long)ro.loci << 32), s_w, tot_rl);
    const unsigned long long my_d = is_long ? ((unsigned long long)ro.n_desc << 32) : (unsigned long long)ro.n_desc;
    const unsigned long long in_d = block_scan64(my_d, s_w, tot_d);
    const uint32_t tile_rows = (uint32_t)tot_rl, tile_loci = (uint32_t)(tot_rl >> 32);
    const uint32_t tile_big = (uint32_t)tot_d, tile_long = (uint32_t)(tot_d >> 32);
    if (warp == 0) {
      unsigned long long ex_b, ex_r;
      tile_lookback(p.tile_state, tile, 0, tot_b, tile_rows, ex_b, ex_r, lane);
      if (lane == 0) {
        s_base[0] = ex_b; s_base[1] = ex_r;
        if (tile == n_tiles - 1) {  // the sub-chunk's totals
          p.ctr->out_cursor = out_base + ex_b + tot_b;
          p.ctr->row_cursor = row0 + ex_r + tile_rows;
          if (out_base + ex_b + tot_b > p.out_cap) p.ctr->out_overflow = 1;
        }
      }
    } else if (warp == 1) {
      if (want_locus) {
        unsigned long long ex_l, ex_0;
        tile_lookback(p.tile_state, tile, 1, tile_loci, 0ull, ex_l, ex_0, lane);
        if (lane == 0) {
          s_base[2] = ex_l;
          if (tile == n_tiles - 1) p.ctr->loci_cursor = loci0 + ex_l + tile_loci;
        }
      } else if (lane == 0) {
        s_base[2] = 0;
      }
    } else if (warp == 2 && lane == 0) {
      // RowDesc slots for this tile's queued rows: ordinary rows from the front, long ones from the end
      uint32_t ok = 1, bb = 0, lb = 0;
      const uint32_t tot = tile_big + tile_long;
      if (tot) {
        const uint32_t d0 = atomicAdd(&p.ctr->n_desc, tot);
        if ((unsigned long long)d0 + tot > p.row_desc_cap) { p.ctr->row_overflow = 1; ok = 0; }
        else {
          if (tile_big) bb = atomicAdd(&p.ctr->n_big_rows, tile_big);
          if (tile_long) lb = atomicAdd(&p.ctr->n_long_rows, tile_long);
        }
      }
      s_misc[1] = bb; s_misc[2] = lb; s_misc[3] = ok;
    }
    __syncthreads();
    const unsigned long long tile_b0 = s_base[0], tile_r0 = s_base[1], tile_l0 = s_base[2];
    const bool out_ok = out_base + tile_b0 + tot_b <= p.out_cap;  // else: flagged by the last tile, the host re-runs the chunk
    const bool desc_ok = s_misc[3] != 0;
    const uint32_t big_base = s_misc[1], long_base = s_misc[2];

    // ---- C: copy-out ----
    // this tile's dosage rows start as all-reference (0); the small rows' samples are scattered below, the queued
    // rows' by the names kernels
    if (want_locus && p.dosage && tile_rows) {
      const unsigned long long ns = (unsigned long long)cfg.n_samples;
      unsigned long long r_lo = row0 + tile_r0, r_hi = r_lo + tile_rows;
      if (r_hi > p.dosage_cap_rows) r_hi = p.dosage_cap_rows;
      if (r_lo < r_hi) {
        uint8_t *const base = reinterpret_cast<uint8_t *>(p.dosage);
        const unsigned long long b0 = r_lo * ns, b1 = r_hi * ns;
        const unsigned long long a0 = (b0 + 15ull) & ~15ull, a1 = b1 & ~15ull;  // cudaMalloc'ed: base is 256-byte aligned
        if (a0 >= a1) {
          for (unsigned long long i = b0 + threadIdx.x; i < b1; i += TILE_THREADS) base[i] = 0;
        } else {
          for (unsigned long long i = b0 + threadIdx.x; i < a0; i += TILE_THREADS) base[i] = 0;
          uint4 *v = reinterpret_cast<uint4 *>(base + a0);
          const unsigned long long nv = (a1 - a0) >> 4;
          for (unsigned long long i = threadIdx.x; i < nv; i += TILE_THREADS) v[i] = make_uint4(0u, 0u, 0u, 0u);
          for (unsigned long long i = a1 + threadIdx.x; i < b1; i += TILE_THREADS) base[i] = 0;
        }
      }
      __syncthreads();
    }
    if (valid && ro.rows) {
      unsigned long long off = out_base + tile_b0 + (in_b - ro.bytes);           // first output byte of this record
      unsigned long long r = tile_r0 + ((uint32_t)in_rl - ro.rows);              // its first row within the sub-chunk
      unsigned long long lo = loci0 + tile_l0 + ((uint32_t)(in_rl >> 32) - ro.loci);
      uint32_t d_ord = is_long ? (uint32_t)(in_d >> 32) - ro.n_desc : (uint32_t)in_d - ro.n_desc;
      if (!ro.failed) {
        for (uint32_t ri = ro.first; ri != ROW_NONE;) {
          const TRow t = s_rows[ri];
          const uint32_t sa = sh.arena_s + t.soff;
          unsigned long long dsts[3] = {~0ull, ~0ull, ~0ull};
          unsigned long long row_bytes = t.slen;
          if (out_ok) {
            uint8_t *g = p.out + off;
            uint32_t pos = 0;
#pragma unroll
            for (int k = 0; k < 3; k++) {
              if (t.hole_len[k]) {
                copy_s2g(g, sa + pos, t.hole_pos[k] - pos);
                g += t.hole_pos[k] - pos;
                pos = t.hole_pos[k];
                dsts[k] = (unsigned long long)(g - p.out);
                g += t.hole_len[k];
              }
            }
            copy_s2g(g, sa + pos, t.slen - pos);
            g += t.slen - pos;
            if (t.flags & 1u) {  // INFO straight from the input line, then the EOL (main.go:684-692)
              copy_g2g(g, info_p, info_n);
              g[info_n] = '\n';
            }
          }
          row_bytes += (unsigned long long)t.hole_len[0] + t.hole_len[1] + t.hole_len[2] + ((t.flags & 1u) ? info_n + 1u : 0u);
          const unsigned long long gr = row0 + r;
          if (want_locus && out_ok && gr < p.dosage_cap_rows) {
            if (lo + t.loc_len <= p.loci_cap) copy_s2g(p.loci + lo, sa + t.slen, t.loc_len);
            p.loci_off[gr] = lo;
            if (!(t.flags & 2u) && p.dosage) small_dosage(p, rec, t.allele, p.dosage + gr * (unsigned long long)cfg.n_samples);
          }
          if (t.flags & 2u) {
            if (desc_ok) {
              RowDesc rd;
              rd.line = li; rd.allele = t.allele;
              rd.het_dst = dsts[0]; rd.hom_dst = dsts[1]; rd.miss_dst = dsts[2];
              rd.n_het = t.cnt[0]; rd.n_hom = t.cnt[1]; rd.n_miss = t.cnt[2];
              rd.row = (uint32_t)r;
              queue_row_desc(p, is_long, d_ord, big_base, long_base, rd);
            }
            d_ord++;
          }
          off += row_bytes; r++; lo += t.loc_len;
          ri = t.next;
        }
      } else {
        // slow path: run the record again, bytes straight to global memory
        GlobalWriter gw;
        gw.g = p.out + off; gw.on = out_ok;
        so.row = r; so.loci_off = lo; so.desc = d_ord; so.is_long = is_long; so.desc_ok = desc_ok;
        so.big_base = big_base; so.long_base = long_base;
        RecOut dummy = ro;
        tile_record<GlobalWriter>(p, li, rec, gw, dummy, sh, so, s_filt, s_filt_off, false, info_p, info_n);
      }
    }
  }
}

}  // namespace bvcf
