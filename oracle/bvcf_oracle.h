/*
 * bvcf_oracle.h -- CPU ORACLE for the bystro-vcf per-line transform.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under bystro_vcf_b200/ may include, link
 * or call this.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs use it, and only as the checker or as
 * the timed CPU baseline.
 *
 * It is a plain-C restatement of the reference's Go algorithm
 * (/root/reference/main.go; citations are main.go:LINE) plus the five symbols
 * of the un-vendored module github.com/bystrogenomics/bystro-utils/parse
 * @ v0.0.0-20180921004542-b5183a523f20 (go.mod:14) whose values are pinned by
 * the reference's tests and golden output.
 *
 * Parity status: PINNED against the reference's golden
 * previous_out_check/out_check_new_10_3_18.vcf.gz (19,821 rows, byte-identical
 * after the reference's own sort) and the known-answer vectors of
 * main_test.go (tests/test_oracle_*.py).  UNPINNED by the reference: E-notation
 * float text (< 1e-4), trTv on MNP rows / non-ACGT REF, NormalizeHeader, CRLF.
 */
#ifndef BVCF_ORACLE_H
#define BVCF_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
  const char *empty_field;    /* main.go:91  default "!" */
  const char *field_delim;    /* main.go:92  default ";" */
  int keep_id;                /* main.go:93 */
  int keep_info;              /* main.go:96 */
  int keep_pos;               /* main.go:95 */
  const char *const *allow;   /* main.go:108-114 */
  int n_allow;                /* < 0 => nil map (allow all) */
  const char *const *exclude; /* main.go:117-123 */
  int n_exclude;              /* 0 => nil map */
  int want_tsv;               /* !noOut */
  int want_dosage;            /* dosageMatrixOutPath != "" */
  int normalize_dots;         /* parse.NormalizeHeader: '.' -> '_' in sample names (unpinned) */
} oracle_config;

/* diagnostic codes: index into the message table main.go:41-51 */
enum {
  ORACLE_DIAG_SAME = 1,      /* "REF == ALT"                       main.go:730 */
  ORACLE_DIAG_BAD_ALT = 2,   /* "ALT not ACTG"                     main.go:737,782 */
  ORACLE_DIAG_DEL1 = 3,      /* "1st base REF != ALT"              main.go:748,835 */
  ORACLE_DIAG_POS = 4,       /* "Invalid POS"                      main.go:755,827 */
  ORACLE_DIAG_INS1 = 5,      /* "1st base ALT != REF"              main.go:798 */
  ORACLE_DIAG_MIXED = 6,     /* "Mixed indel/snp sites not supported" main.go:934,986 */
  ORACLE_DIAG_DEL1_LIST = 7, /* delError1 inside the ALT list, "ALT#%d" format   main.go:835 */
  ORACLE_DIAG_POS_LIST = 8   /* posError inside the ALT list, no ALT number      main.go:827 */
};

typedef struct {
  uint64_t line_no; /* 0-based index of the data line after the #CHROM header */
  int32_t alt_no;   /* 1-based ALT number, 0 when the message has none */
  int32_t code;
} oracle_diag;

typedef struct {
  char *tsv;          /* body rows in input order (no header line) */
  size_t tsv_len;
  uint64_t n_rows;
  uint64_t n_lines;   /* newline-terminated data lines seen */
  /* dosage matrix rows (one per emitted allele when want_dosage and samples) */
  char *loci;         /* '\n'-joined locus strings */
  size_t loci_len;
  int8_t *dosage;     /* n_dosage_rows x n_samples, row-major */
  uint64_t n_dosage_rows;
  uint32_t n_samples;
  oracle_diag *diags;
  size_t n_diags;
  int error;          /* 0 ok, 1 "Not a VCF file", 2 "No header found" */
} oracle_result;

/* main.go:219-239: the TSV header line, without the trailing newline. Returns malloc'd string. */
char *oracle_header(const oracle_config *cfg);

/* main.go:241-396 + 476-721: whole-stream transform, rows in input order.
 * threads <= 1 : single thread.  threads > 1: data lines are split into
 * `threads` contiguous blocks processed concurrently and concatenated in order. */
int oracle_read_vcf(const oracle_config *cfg, const char *in, size_t len, int threads, oracle_result *res);

/* Same per-line transform, but on a headerless block of newline-terminated data
 * lines with the "#CHROM..." line given separately (the processLines contract,
 * main.go:476).  Used as the CPU baseline and by chunk-level parity tests. */
int oracle_process_block(const oracle_config *cfg, const char *chrom_line, size_t chrom_len,
                         int eol_width, const char *block, size_t len, int threads, oracle_result *res);

void oracle_free_result(oracle_result *res);

/* unit-level entry points (main_test.go known-answer vectors) */

/* main.go:723-1038.  Returns number of output alleles; fills type (<=16 chars),
 * and for i < min(n, cap): pos[i] (<=24 chars), ref[i], alt (malloc'd strings), alt_idx[i]. */
int oracle_get_alleles(const char *chrom, const char *pos, const char *ref, const char *alt,
                       char *type_out, int cap, char (*pos_out)[24], char *ref_out, char **alt_out,
                       int *idx_out);

/* main.go:1042-1194 for `n` sample fields; flags[i] = 0 none,1 het,2 hom,3 missing; dosage optional */
void oracle_het_hom(const char *const *fields, int n, const char *allele_num, uint8_t *flags,
                    int8_t *dosage, int *ac, int *an);

int oracle_alt_is_valid(const char *alt, size_t n);                     /* main.go:456-474 */
void oracle_format_float(double q, char out[32]);                        /* FormatFloat(q,'G',3,64) */
const char *oracle_tr_tv(char ref, const char *alt, size_t alt_len);     /* parse.GetTrTv */

#ifdef __cplusplus
}
#endif
#endif
