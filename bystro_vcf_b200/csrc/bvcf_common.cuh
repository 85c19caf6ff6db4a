// bvcf_common.cuh -- shared device-side types and warp helpers for libbvcf (sm_100a).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace bvcf {

constexpr uint32_t FULL = 0xFFFFFFFFu;

// ---- line table -------------------------------------------------------------------------------
// One record per data line that has the header's field count (main.go:449).  Written by the scan
// kernel into per-range slots, then compacted into input order.
struct __align__(16) LineRec {
  uint64_t start;     // byte offset of the line in the input region
  uint32_t len;       // length including the EOL
  uint32_t an;        // non-missing allele count of the fast-classified samples (main.go:1067,1169)
  uint32_t ev_start;  // first event word of this line
  uint32_t ev_count;  // event words of this line (two per quad event)
  uint32_t ord;       // ordinal of this line among ALL lines that start in its range
  uint16_t tab[9];    // offsets of the first nine tabs from `start` (0xFFFF: beyond 64 KiB, rescan); only
                      // the first min(H-1, 9) entries are meaningful
  uint16_t flags;     // bit 0: some sample carries an ALT number > 1 or needs the general GT grammar, so the
                      // inline ALT #1 summary below is not the whole story (bvcf_line_stats_kernel serves it)
  uint32_t n_het1, n_hom1, n_miss, ac1;  // ALT #1 summary over the fast-classified samples (main.go:1042-1194)
};
static_assert(sizeof(LineRec) == 64, "LineRec layout");

// ---- genotype events ----------------------------------------------------------------------------
// The scan kernel emits one 64-bit "quad" event (two 32-bit words, 8-byte aligned) per group of four
// consecutive samples of which at least one is not plain reference, in header order:
//   word 0  bits  0..19  base sample index + EV_BASE_BIAS (the first of the four samples; may be -3..-1 for the
//                        group that straddles the start of a line's sample zone)
//           bit  30      complex: ONE sample (slot 0) whose GT needs the general grammar (polyploid, allele >= 10,
//                        mixed separators); word 1 is then the byte offset of the field from the line start
//   word 1  eight 4-bit allele codes: nibble k (k = 0..3) = first allele of sample base+k, nibble 4+k = its
//           second allele.  0 = reference / any other token, 1..9 = that allele number, 0xE = '.',
//           0xF = absent (haploid sample, second nibble only).  A sample whose two nibbles are 0 carries nothing.
// Consumers expand a slot to the 32-bit "slot word" below (sample | c1 << 20 | c2 << 25) with ev_slot_word().
constexpr uint32_t EV_SAMPLE_MASK = 0xFFFFFu;
constexpr uint32_t EV_COMPLEX = 1u << 30;
constexpr uint32_t EV_OFFSET_TAG = 1u << 31;   // slot word only: "no sample in this slot"
constexpr uint32_t EV_CODE_ABSENT = 30, EV_CODE_MISSING = 31;   // slot-word codes
constexpr uint32_t EV_NIB_MISSING = 0xEu, EV_NIB_ABSENT = 0xFu; // payload nibbles
constexpr uint32_t EV_BASE_BIAS = 3;
constexpr uint32_t MAX_SAMPLES = (1u << 20) - 16;

__device__ __forceinline__ uint32_t ev_make(uint32_t sample, uint32_t c1, uint32_t c2) {
  return sample | (c1 << 20) | (c2 << 25);
}
// one-sample quad (slot 0) from slot-word codes
__device__ __forceinline__ uint32_t ev_single_payload(uint32_t c1, uint32_t c2) {
  const uint32_t n1 = c1 == EV_CODE_MISSING ? EV_NIB_MISSING : c1;
  const uint32_t n2 = c2 == EV_CODE_MISSING ? EV_NIB_MISSING : (c2 == EV_CODE_ABSENT ? EV_NIB_ABSENT : c2);
  return n1 | (n2 << 16);
}
// slot j (0..3) of the quad (h, pl) as a slot word; EV_OFFSET_TAG when the slot carries nothing.
// Complex quads yield `sample | EV_COMPLEX` in slot 0 (the field offset is pl).
__device__ __forceinline__ uint32_t ev_slot_word(uint32_t h, uint32_t pl, int j) {
  const uint32_t samp = (h & EV_SAMPLE_MASK) + (uint32_t)j - EV_BASE_BIAS;
  if (h & EV_COMPLEX) return j == 0 ? (samp | EV_COMPLEX) : EV_OFFSET_TAG;
  const uint32_t n1 = (pl >> (4 * j)) & 0xFu, n2 = (pl >> (16 + 4 * j)) & 0xFu;
  if ((n1 | n2) == 0) return EV_OFFSET_TAG;
  if (n1 == EV_NIB_MISSING || n2 == EV_NIB_MISSING) return ev_make(samp, EV_CODE_MISSING, EV_CODE_MISSING);
  return ev_make(samp, n1, n2 == EV_NIB_ABSENT ? EV_CODE_ABSENT : n2);
}

// ---- flags written by kernels, read by the host after the run -----------------------------------
struct RunCounters {
  unsigned long long out_cursor;   // bytes of TSV written so far (device-side running offset)
  unsigned long long row_cursor;   // rows written so far
  unsigned long long loci_cursor;  // bytes of locus strings written so far (dosage output)
  unsigned long long n_lines;      // all newline-terminated lines
  unsigned long long n_records;    // lines with the right field count
  unsigned long long chunk_out_base;   // out_cursor before this sub-chunk (set by the prefix spine, mode 0)
  unsigned long long chunk_row_base;
  unsigned long long chunk_loci_base;
  unsigned long long chunk_line_base;  // n_lines before this sub-chunk (diagnostic line numbers)
  unsigned long long scratch_cursor;   // bytes of tile blocks in the scratch buffer (per sub-chunk)
  unsigned int ev_overflow;        // a range ran out of event slots
  unsigned int slot_overflow;      // a range ran out of line slots
  unsigned int out_overflow;       // output region too small
  unsigned int row_overflow;       // a sub-chunk queued more rows for the names kernels than RowDesc slots
  unsigned int n_diags;
  unsigned int chunk_records;      // records in the current sub-chunk (device-side n for grid-stride kernels)
  unsigned int n_big_recs;         // work list of the stats kernel: records with long event lists (per sub-chunk)
  unsigned int big_rec_cursor;     // next entry of the stats work list to be taken
  unsigned int tile_ticket;        // bvcf_compose_kernel: next tile of 32 records to be taken (per sub-chunk)
  unsigned int tile_ticket2;       // bvcf_copyout_kernel: the same
  unsigned int n_slow;             // slow-path records of the sub-chunk (bvcf_slow_rows_kernel's work list)
  unsigned int scratch_overflow;   // the tile blocks did not fit the scratch buffer
  unsigned int slow_overflow;      // more slow-path records than list entries
  unsigned int n_mid_rows;         // names work list: rows written by one lane each (bvcf_names_mid_kernel)
  unsigned int n_big_rows;         // names work list: rows written by a warp each (per sub-chunk)
  unsigned int big_row_cursor;     // next entry to be taken (dynamic scheduling)
  unsigned int n_long_rows;        // rows with very long event lists: written by a whole CTA (bvcf_names_long_kernel)
  unsigned int long_row_cursor;
};

// ---- configuration as the kernels see it ---------------------------------------------------------
struct DevCfg {
  int H;           // header field count (main.go:449)
  int n_samples;   // max(H-9,0)
  int eol_width;   // numChars
  int keep_id, keep_info, keep_pos, want_tsv, want_dosage;
  int allow_all;   // allowedFilters == nil
  int n_allow, n_excl;
  const uint8_t *filt_blob;   // allow strings then exclude strings, back to back
  const uint32_t *filt_off;   // n_allow + n_excl + 1 offsets
  int filt_bytes;
  uint8_t empty[64];
  int empty_len;
  uint8_t delim[64];
  int delim_len;
  uint8_t tail0[208];         // columns 7-15 of a sites-only row: "!\t0\t!\t0\t!\t0\t0\t0\t0" with the configured emptyField
  int tail0_len;
  const uint8_t *names;       // sample names back to back
  const uint32_t *name_off;   // n_samples + 1
  int name_fixed_w;           // > 0 when every sample name has this length
  const unsigned long long *name8;  // name + delimiter packed in 8 bytes per sample, when 7-char names and a
                                    // 1-char delimiter make every list item exactly 8 bytes; else null
  const uint4 *name16;              // name + delimiter zero-padded to 16 bytes per sample, when all names have one
                                    // width and an item (name + delimiter) is 5..16 bytes; else null
  int item_bytes;                   // that item size
};

// ---- warp helpers --------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t v, int lane) {
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    uint32_t t = __shfl_up_sync(FULL, v, d);
    if (lane >= d) v += t;
  }
  return v;
}
__device__ __forceinline__ uint32_t warp_sum(uint32_t v) {
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(FULL, v, d);
  return v;
}
__device__ __forceinline__ unsigned long long warp_sum64(unsigned long long v) {
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(FULL, v, d);
  return v;
}

// 4-bit mask of the bytes of w equal to the byte replicated in pat (exact, no false positives)
__device__ __forceinline__ uint32_t eq_mask4(uint32_t w, uint32_t pat) {
  uint32_t t = w ^ pat;
  uint32_t y = (t & 0x7F7F7F7Fu) + 0x7F7F7F7Fu;
  y = ~(y | t | 0x7F7F7F7Fu);               // 0x80 in every zero byte of t
  return ((y >> 7) * 0x00204081u >> 21) & 0xFu;
}
__device__ __forceinline__ uint32_t eq_mask16(const uint4 &v, uint32_t pat) {
  return eq_mask4(v.x, pat) | (eq_mask4(v.y, pat) << 4) | (eq_mask4(v.z, pat) << 8) | (eq_mask4(v.w, pat) << 12);
}

}  // namespace bvcf
